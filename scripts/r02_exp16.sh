#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/r02_exp16_2gpu.py > gpurun_out/exp16_breakdown.txt 2>&1
cat gpurun_out/exp16_breakdown.txt | grep -v "^\*\|OMP_NUM\|^$"
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py -m gpu -q -x -k "exp" > gpurun_out/exp16_pytest.log 2>&1; tail -3 gpurun_out/exp16_pytest.log
