#!/bin/bash
# Runs bench.py for every BASELINE workload on one GPU; JSON lines -> gpurun_out/bench_all.jsonl
mkdir -p gpurun_out
: > gpurun_out/bench_all.jsonl
for w in ${WORKLOADS:-cfg2_group_nansum cfg2_group_nanmean cfg2_group_nanstd cfg1_move_mean cfg1s_move_mean cfg1s_move_sum cfg1s_move_std cfg1s_move_var cfg1s_move_cov cfg1s_move_corr cfg3_move_exp_nanmean cfg3_ffill cfg3_bfill cfg4_move_std cfg4_move_cov cfg4_move_corr cfg5_group_nansum1d cfg5_group_nanargmax cfg5_group_nanfirst cfg5_group_nanvar red_nansum_f32 red_nanmean_f32 red_nanvar_f32 red_nanmax_f32 red_nanargmax_f32 red_nansum_f64 red_nanstd_f64 red_nansum_f32_axis0 red_nanvar_f64_axis0 red_nanmean_f32_short red_nansum_f64_all quant_median_long quant_quartiles_short mat_move_cov}; do
  timeout 600 python bench.py --workload $w --steps $([ $w = cfg1_move_mean ] && echo 200 || echo ${STEPS:-5}) --warmup 3 2>gpurun_out/bench_err_$w.log | tail -1 >> gpurun_out/bench_all.jsonl || echo "{\"workload\": \"$w\", \"failed\": true}" >> gpurun_out/bench_all.jsonl
done
python - <<'PY'
import json
for line in open("gpurun_out/bench_all.jsonl"):
    try: d = json.loads(line)
    except Exception: print("bad line", line[:100]); continue
    if "config" not in d: print(d); continue
    print(f"{d['config']['workload']:24s} {d['value']/1e9:9.1f} Gel/s  {d['ms_per_step']:9.3f} ms  roofline {d['roofline']['frac']:.3f}  e2e {d['e2e']['value']/1e9:7.2f} Gel/s  cpu {d['cpu_baseline']['value']/1e9:6.2f} Gel/s ({d['cpu_baseline']['cores']} cores)  launches/step {d['gpu_launches']/d['steps']:.0f}")
PY
