#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -x -k "move or exp" > gpurun_out/exp31_pytest.log 2>&1; tail -3 gpurun_out/exp31_pytest.log
timeout 600 python scripts/r02_quick.py cfg3x cfg1s > gpurun_out/exp31.jsonl 2> gpurun_out/exp31.err; cat gpurun_out/exp31.jsonl; tail -3 gpurun_out/exp31.err
