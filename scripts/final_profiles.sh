#!/bin/bash
# ncu --set full capture of the dominant kernel of one step per BASELINE workload
mkdir -p gpurun_out
for spec in "cfg2_group_nansum group_rowbins" "cfg2_group_nanstd group_rowbins" "cfg3_ffill scan_rowtile" "cfg3_move_exp_nanmean scan_rowtile" "cfg1s_move_mean move_rowtile" "cfg4_move_std move_rowtile" "cfg5_group_nansum1d group_atomic" "cfg5_group_nanvar group_atomic"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o /tmp/final_$1 python scripts/prof_workload.py $1 > /dev/null 2>&1
  python scripts/ncu_summary.py /tmp/final_$1.ncu-rep 14 > gpurun_out/ncu_$1.txt 2>&1
done
ls -la gpurun_out | grep ncu_
