#!/bin/bash
# ncu --set full capture of the reduce kernel of one step for a few red_* workloads
mkdir -p gpurun_out
for spec in "red_nansum_f32 red_rows_cta" "red_nanvar_f32 red_rows_cta" "red_nansum_f32_axis0 red_stream" "red_nanmean_f32_short red_rows_tile" "red_nanstd_f64 red_rows_cta"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o /tmp/final_$1 python scripts/prof_workload.py $1 > /dev/null 2>&1
  python scripts/ncu_summary.py /tmp/final_$1.ncu-rep 14 > gpurun_out/ncu_$1.txt 2>&1
done
ls -la gpurun_out | grep ncu_red
