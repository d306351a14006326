#!/bin/bash
# Round-2 closing check on one B200: the -m gpu suite, smoke(), the default bench command and the
# reference arm with wall-clock times, and the bench rows whose kernels changed after the evidence run.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log; tail -3 gpurun_out/r02_pytest_gpu.log
cp gpurun_out/parity_observed.json gpurun_out/r02_parity_observed.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
s=$(date +%s); timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default bench rc=$? wall=$(( $(date +%s) - s ))s"
s=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference arm rc=$? wall=$(( $(date +%s) - s ))s"; cut -c1-400 gpurun_out/r02_bench_reference.json
: > gpurun_out/r02_bench_patch.jsonl
for w in cfg1_move_mean cfg3_move_exp_nancorr; do
  steps=10; [ $w = cfg1_move_mean ] && steps=200
  timeout 600 python bench.py --workload $w --steps $steps --warmup 3 2>/dev/null | tail -1 >> gpurun_out/r02_bench_patch.jsonl
done
for w in mat_move_cov mat_move_corr; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity 2>/dev/null | tail -1 >> gpurun_out/r02_bench_patch.jsonl
done
python - <<'PY'
import json
for line in open("gpurun_out/r02_bench_patch.jsonl"):
    d = json.loads(line); print(d["config"]["workload"], round(d["ms_per_step"], 4), "ms", round(d["roofline"]["frac"], 3))
d = json.loads(open("gpurun_out/r02_bench_default.json").read().strip().splitlines()[-1])
print("default:", d["config"]["workload"], round(d["ms_per_step"], 3), "ms", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"] / 1e9, 2), "cpu", round(d["cpu_baseline"]["value"] / 1e9, 2), d["cpu_baseline"]["kind"], d["gpu_launches"])
PY
