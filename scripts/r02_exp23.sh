#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -x -k "group" > gpurun_out/exp23_pytest.log 2>&1; tail -4 gpurun_out/exp23_pytest.log
timeout 900 python scripts/r02_quick.py cfg5q --steps 6 > gpurun_out/exp23_cfg5.jsonl 2> gpurun_out/exp23_cfg5.err; cat gpurun_out/exp23_cfg5.jsonl; tail -3 gpurun_out/exp23_cfg5.err
