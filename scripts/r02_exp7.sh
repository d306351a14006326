#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_kwargs.py -m gpu -q -x -k "group or kwargs or out or dtype or endian or empty" > gpurun_out/exp7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp7_pytest.log
tail -4 gpurun_out/exp7_pytest.log
timeout 600 python scripts/r02_quick.py cfg2 sweep > gpurun_out/exp7_cfg2.jsonl 2> gpurun_out/exp7_cfg2.err
cat gpurun_out/exp7_cfg2.jsonl
NBG_RB2_S=2 NBG_RB2_NW=31 timeout 600 ncu --set full --clock-control none --import-source on -k regex:group_rowbins2 -s 2 -c 1 -o /tmp/rb2 python scripts/prof_workload.py cfg2_group_nansum > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/rb2.ncu-rep 16 > gpurun_out/exp7_ncu_rb2.txt 2>&1
cat gpurun_out/exp7_ncu_rb2.txt
tail -n 5 gpurun_out/exp7_cfg2.err
