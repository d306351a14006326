#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quant_ -s 2 -c 1 -o /tmp/qshort python scripts/prof_workload.py quant_quartiles_short > gpurun_out/exp26_prof.log 2>&1
python scripts/ncu_summary.py /tmp/qshort.ncu-rep 25 > gpurun_out/exp26_ncu_quant_short.txt 2>&1; head -45 gpurun_out/exp26_ncu_quant_short.txt
