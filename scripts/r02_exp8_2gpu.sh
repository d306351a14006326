#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/exp8_gpus.txt
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -q -x > gpurun_out/exp8_pytest_dist.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp8_pytest_dist.log
tail -15 gpurun_out/exp8_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/exp8_bench_n2.json 2> gpurun_out/exp8_bench_n2.err
echo "bench rc=$?"
cat gpurun_out/exp8_bench_n2.json
tail -n 20 gpurun_out/exp8_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 --workload cfg3_ffill --no-sharded > gpurun_out/exp8_bench_n2_ffill.json 2> gpurun_out/exp8_bench_n2_ffill.err
cat gpurun_out/exp8_bench_n2_ffill.json; tail -n 5 gpurun_out/exp8_bench_n2_ffill.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/exp8_bench_n2_ref.json 2> gpurun_out/exp8_bench_n2_ref.err
cat gpurun_out/exp8_bench_n2_ref.json
