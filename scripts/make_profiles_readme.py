"""profiles/README.md from a bench_all.jsonl (scripts/run_all_benches.sh).  usage: make_profiles_readme.py <jsonl> <tag>"""
import json, sys
path, tag = sys.argv[1], sys.argv[2]
rows = []


def fmt(x):
    return f"{x:.1f}" if x >= 10 else f"{x:.3g}"


for line in open(path):
    try:
        d = json.loads(line)
    except Exception:
        continue
    if "config" in d:
        rows.append(d)
out = [f"# profiles — round {tag}", "",
       "All numbers from one B200 (`gpurun`), device-resident inputs, CUDA events, >= 3 warm-up steps,",
       "inputs larger than L2 (except cfg1).  `roofline` = algorithmic bytes / step time / measured copy",
       "bandwidth (6447.8 GB/s, MEASURED_PEAKS.json).  `e2e` = public numpy API with pinned host buffers",
       "(H2D + kernels + D2H inside the timed region; row blocks are pipelined on three streams).",
       "`cpu` = oracle port (C, OpenMP over rows) on the box's 16 host cores; 1-D inputs use one core,",
       "like the reference's gufunc.  The `red_*` rows (plain NaN reductions, SURVEY 8(f) rank 1) only READ",
       "(8 bytes written per output), so they can exceed the measured COPY bandwidth used as `peak` (half",
       "reads, half writes): `roofline` above 1.0 means faster than a device-to-device copy moves the same",
       "bytes; ncu shows ~90 % of the DRAM peak for `nansum` float32.  The `quant_*` rows (nanquantile,",
       "SURVEY 8(f) rank 3) are selection: 8 reads of the data by construction, `roofline` counts one.",
       "cfg1 is launch-bound: its K steps are replayed from one CUDA graph (`config.launch`).  `mat_*` (matrix",
       "functions, SURVEY 8(f) rank 2) is the first, bit-exact but sequential-per-pair kernel: 1024 threads.", "",
       "| workload | shape | Gel/s | ms/step | roofline (of measured) | e2e Gel/s | cpu Gel/s (cores) | kernels/step |",
       "|---|---|---:|---:|---:|---:|---:|---:|"]
for d in rows:
    c = d["config"]
    out.append(f"| {c['workload']} ({c['func']}, {d['dtype']}) | {c['shape'][0]}x{c['shape'][1]}{' axis=' + str(c['axis']) if 'axis' in c else ''} | {fmt(d['value']/1e9)} | {d['ms_per_step']:.3f} | "
               f"{d['roofline']['frac']:.3g} | {fmt(d['e2e']['value']/1e9)} | {fmt(d['cpu_baseline']['value']/1e9)} ({d['cpu_baseline']['cores']}) | {d['gpu_launches']/d['steps']:.0f} |")
out += ["", "Files:", "",
        "* `*_bench_all*.jsonl` — the raw bench.py JSON lines behind the table.",
        "* `*_launches_default_bench*.csv` — ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`) of the default bench command; the row-bins kernel is ~99 % of each step.",
        "* `*_ncu_*.txt` — per-kernel summaries of `ncu --set full` captures (scripts/ncu_summary.py): duration, DRAM bytes (= algorithmic bytes: no re-reads), pipe utilisation, stall reasons, top stalled SASS instructions.",
        "* `*_launches_quant_median_long.csv` — ncu launch list of one `nanquantile` call on 2000 x 10^6 float64 (radix select: 8 x histogram + select + clear, then finish), captured before the compact target list: about 6 ms per histogram pass then, about 4 ms now.",
        ""]
open("profiles/README.md", "w").write("\n".join(out))
print("\n".join(out[:40]))
