#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"part_" -c 8 --csv --log-file gpurun_out/exp33_part_raw.csv python scripts/prof_workload.py cfg5_group_nanvar > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/exp33_part_raw.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hd = rows[h]
ik, im, iv = hd.index("Kernel Name"), hd.index("Metric Name"), hd.index("Metric Value")
for r in rows[h + 1:]:
    print(r[ik][:40], r[im], r[iv])
PY
