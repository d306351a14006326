#!/bin/bash
# triangular matrix kernels; byte-based prefetch distance on the scan / window kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_matrix.py -q -x > gpurun_out/exp19_pytest.log 2>&1; tail -5 gpurun_out/exp19_pytest.log
timeout 600 python scripts/r02_quick.py mat > gpurun_out/exp19_mat.jsonl 2> gpurun_out/exp19_mat.err; cat gpurun_out/exp19_mat.jsonl; tail -3 gpurun_out/exp19_mat.err
timeout 600 python scripts/r02_quick.py cfg3 cfg1s > gpurun_out/exp19_cfg3.jsonl 2> gpurun_out/exp19_cfg3.err; cat gpurun_out/exp19_cfg3.jsonl; tail -3 gpurun_out/exp19_cfg3.err
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fill or exp or move" > gpurun_out/exp19_pytest2.log 2>&1; tail -3 gpurun_out/exp19_pytest2.log
