#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/exp17_ngpu.txt
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -q > gpurun_out/exp17_pytest_dist.log 2>&1
tail -3 gpurun_out/exp17_pytest_dist.log
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/exp17_bench_n$N.json 2> gpurun_out/exp17_bench_n$N.err
echo "N=$N rc=$?"; cat gpurun_out/exp17_bench_n$N.json | cut -c1-600; tail -n 3 gpurun_out/exp17_bench_n$N.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 --workload cfg3_move_exp_nanmean --no-sharded > gpurun_out/exp17_bench_n8_exp.json 2> gpurun_out/exp17_bench_n8_exp.err
cat gpurun_out/exp17_bench_n8_exp.json | cut -c1-500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 10 --warmup 3 --workload cfg5_group_nansum1d --no-sharded > gpurun_out/exp17_bench_n8_g5.json 2> gpurun_out/exp17_bench_n8_g5.err
cat gpurun_out/exp17_bench_n8_g5.json | cut -c1-500
