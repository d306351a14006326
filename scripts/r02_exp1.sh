#!/bin/bash
# round-2 GPU experiment 1: regression suite + A/B timings of the new kernels / tuning knobs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/exp1_gpu.txt 2>&1
nproc >> gpurun_out/exp1_gpu.txt; free -g >> gpurun_out/exp1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/exp1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp1_pytest.log
tail -5 gpurun_out/exp1_pytest.log
timeout 600 python scripts/r02_quick.py cfg2 > gpurun_out/exp1_cfg2.jsonl 2> gpurun_out/exp1_cfg2.err
timeout 400 python scripts/r02_quick.py cfg3 > gpurun_out/exp1_cfg3.jsonl 2> gpurun_out/exp1_cfg3.err
timeout 400 python scripts/r02_quick.py cfg4 cfg1s > gpurun_out/exp1_cfg4.jsonl 2> gpurun_out/exp1_cfg4.err
timeout 600 python scripts/r02_quick.py cfg5 --steps 6 > gpurun_out/exp1_cfg5.jsonl 2> gpurun_out/exp1_cfg5.err
cat gpurun_out/exp1_cfg2.jsonl gpurun_out/exp1_cfg3.jsonl gpurun_out/exp1_cfg4.jsonl gpurun_out/exp1_cfg5.jsonl
tail -3 gpurun_out/exp1_*.err
