#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:part_scatter -s 0 -c 1 -o /tmp/pscat python scripts/prof_workload.py cfg5_group_nanvar > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/pscat.ncu-rep 16 > gpurun_out/exp34_ncu_part_scatter.txt 2>&1; head -44 gpurun_out/exp34_ncu_part_scatter.txt
