#!/bin/bash
mkdir -p gpurun_out
for spec in "cfg3_move_exp_nanmean scan_rowtile" "cfg3_ffill scan_rowtile"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o /tmp/r02_$1 python scripts/prof_workload.py $1 > /dev/null 2>&1
  python scripts/ncu_summary.py /tmp/r02_$1.ncu-rep 45 > gpurun_out/r02_ncu_$1.txt 2>&1
  head -30 gpurun_out/r02_ncu_$1.txt
done
