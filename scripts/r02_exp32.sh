#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "partition or high_cardinality" > gpurun_out/exp32_pytest.log 2>&1; tail -15 gpurun_out/exp32_pytest.log
timeout 900 python scripts/r02_quick.py cfg5p --steps 6 > gpurun_out/exp32_cfg5.jsonl 2> gpurun_out/exp32_cfg5.err; cat gpurun_out/exp32_cfg5.jsonl; tail -3 gpurun_out/exp32_cfg5.err
