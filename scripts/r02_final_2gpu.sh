#!/bin/bash
# final round-2 check on 2 GPUs: NCCL tests, smoke (sharded forms), default bench with the sharded block
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_distributed_gpu.py tests/test_gpu_matrix.py -m gpu -q -x > gpurun_out/final2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final2_pytest.log; tail -3 gpurun_out/final2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.log 2>&1; tail -2 gpurun_out/final2_smoke.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/final2_bench_n2.json 2> gpurun_out/final2_bench_n2.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/final2_bench_n2.json").read().strip().splitlines()[-1])
print(d["config"]["workload"], d["n_gpus"], round(d["value"] / 1e9, 1), "Gel/s", round(d["ms_per_step"], 3), "ms", d["roofline"]["frac"], d.get("gpu_launches"))
for k, v in (d.get("sharded") or {}).items():
    print("  ", k, round(v["elements_per_s"] / 1e9, 1), "Gel/s", round(v["ms_per_step"], 3), "ms parity", v.get("parity_ok"))
PY
tail -n 5 gpurun_out/final2_bench_n2.err
timeout 300 python scripts/r02_quick.py mat > gpurun_out/final2_mat.jsonl 2>&1; grep "mat_move\|mat_exp_cov_f64" gpurun_out/final2_mat.jsonl
