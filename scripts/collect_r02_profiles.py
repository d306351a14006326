"""Copy the round-2 evidence from gpurun_out/ (scratch) into profiles/ (tracked) and derive
profiles/r02_traffic.json -- dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant
kernel of every workload that has an `ncu --set full` summary (gpurun_out/r02_ncu_<workload>.txt,
written by scripts/ncu_summary.py).  bench.py reads that file for `roofline.traffic`."""
import glob
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
traffic = {}
for path in sorted(glob.glob(os.path.join(SRC, "r02_ncu_*.txt"))):
    wl = os.path.basename(path)[len("r02_ncu_"):-4]
    txt = open(path).read()
    if "dram__bytes_read.sum" not in txt:
        continue
    shutil.copy(path, os.path.join(DST, os.path.basename(path)))
    tot = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        m = re.search(key + r"\s+(\w+)\s+([0-9.]+)", txt)
        tot += float(m.group(2)) * UNIT[m.group(1)]
    kern = re.search(r"Kernel Name\s+(.*)", txt).group(1).strip()
    dur = re.search(r"gpu__time_duration.sum\s+(\w+)\s+([0-9.]+)", txt)
    traffic[wl] = dict(dram_bytes_per_launch=tot, kernel=kern, ncu_duration=f"{dur.group(2)} {dur.group(1)}")
json.dump(traffic, open(os.path.join(DST, "r02_traffic.json"), "w"), indent=1, sort_keys=True)
for name in ("r02_bench_all.jsonl", "r02_launches_default_bench.csv", "r02_launches_default_bench_raw.csv",
             "r02_parity_observed.json", "r02_prefetch_sweep.jsonl", "r02_pytest_gpu.log"):
    p = os.path.join(SRC, name)
    if os.path.exists(p):
        shutil.copy(p, os.path.join(DST, name))
print(f"{len(traffic)} workloads with measured traffic")
for k, v in traffic.items():
    print(f"{k:30s} {v['dram_bytes_per_launch']/1e9:9.2f} GB  {v['ncu_duration']:>14s}  {v['kernel'][:80]}")
