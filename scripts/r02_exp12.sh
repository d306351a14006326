#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -x -k "move or golden" > gpurun_out/exp12_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp12_pytest.log
tail -5 gpurun_out/exp12_pytest.log
timeout 300 python scripts/r02_quick.py cfg4 > gpurun_out/exp12_cfg4.jsonl 2> gpurun_out/exp12_cfg4.err
cat gpurun_out/exp12_cfg4.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:move_prefix -s 1 -c 1 -o /tmp/pfx python scripts/prof_workload.py cfg4_move_std > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/pfx.ncu-rep 30 > gpurun_out/exp12_ncu_pfx_std.txt 2>&1
cat gpurun_out/exp12_ncu_pfx_std.txt
tail -n 5 gpurun_out/exp12_cfg4.err
