#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -x -k "move or golden" > gpurun_out/exp10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp10_pytest.log
tail -15 gpurun_out/exp10_pytest.log
timeout 300 python scripts/r02_quick.py cfg4 > gpurun_out/exp10_cfg4.jsonl 2> gpurun_out/exp10_cfg4.err
NBG_PFX_OFF=1 timeout 300 python scripts/r02_quick.py cfg4 > gpurun_out/exp10_cfg4_old.jsonl 2>> gpurun_out/exp10_cfg4.err
cat gpurun_out/exp10_cfg4.jsonl; echo; cat gpurun_out/exp10_cfg4_old.jsonl
timeout 300 python scripts/r02_exp9.py > gpurun_out/exp10_exp_pieces.txt 2>&1
cat gpurun_out/exp10_exp_pieces.txt
tail -n 5 gpurun_out/exp10_cfg4.err
