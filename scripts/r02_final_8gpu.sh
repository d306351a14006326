#!/bin/bash
mkdir -p gpurun_out
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err
echo "N=$N rc=$?"; python - <<PY
import json
d = json.loads(open("gpurun_out/final_bench_n$N.json").read().strip().splitlines()[-1])
print(d["config"]["workload"], d["n_gpus"], round(d["value"] / 1e9, 1), "Gel/s", round(d["ms_per_step"], 3), "ms", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"] / 1e9, 1))
for k, v in (d.get("sharded") or {}).items():
    print("  ", k, round(v["elements_per_s"] / 1e9, 1), "Gel/s", round(v["ms_per_step"], 3), "ms parity", v.get("parity_ok"))
PY
done
