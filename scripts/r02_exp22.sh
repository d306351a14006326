#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -x -k "exp" > gpurun_out/exp22_pytest.log 2>&1; tail -5 gpurun_out/exp22_pytest.log
timeout 600 python scripts/r02_quick.py cfg3x > gpurun_out/exp22_cfg3.jsonl 2> gpurun_out/exp22_cfg3.err; cat gpurun_out/exp22_cfg3.jsonl; tail -3 gpurun_out/exp22_cfg3.err
