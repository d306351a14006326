#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:move_prefix -s 1 -c 1 -o /tmp/pfx python scripts/prof_workload.py cfg4_move_std > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/pfx.ncu-rep 40 > gpurun_out/exp11_ncu_pfx_std.txt 2>&1
cat gpurun_out/exp11_ncu_pfx_std.txt
