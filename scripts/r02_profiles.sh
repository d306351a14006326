#!/bin/bash
# Round-2 evidence run (one B200): regression suite, every bench row, ncu launch list of the default
# bench, `ncu --set full` summaries of the dominant kernel of each BASELINE workload.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -4 gpurun_out/r02_pytest_gpu.log
cp gpurun_out/parity_observed.json gpurun_out/r02_parity_observed.json 2>/dev/null
# ---- bench rows: BASELINE configs with every leg, the 8(f) rows device-timed only
: > gpurun_out/r02_bench_all.jsonl
for w in cfg2_group_nansum cfg2_group_nanmean cfg2_group_nanstd cfg2_group_nancount cfg2_group_nanmax cfg2_group_nanargmax \
         cfg1_move_mean cfg1s_move_mean cfg1s_move_sum cfg1s_move_std cfg1s_move_var cfg1s_move_cov cfg1s_move_corr \
         cfg3_move_exp_nanmean cfg3_move_exp_nansum cfg3_move_exp_nancount cfg3_move_exp_nanvar cfg3_move_exp_nanstd cfg3_move_exp_nancov cfg3_move_exp_nancorr cfg3_move_exp_nanmean_f32 cfg3_ffill cfg3_bfill \
         cfg4_move_std cfg4_move_var cfg4_move_mean cfg4_move_cov cfg4_move_corr \
         cfg5_group_nansum1d cfg5_group_nanmean cfg5_group_nanargmax cfg5_group_nanfirst cfg5_group_nanvar; do
  if [ -n "$NBG_SKIP_BENCH" ]; then break; fi
  steps=10; [ $w = cfg1_move_mean ] && steps=200
  timeout 600 python bench.py --workload $w --steps $steps --warmup 3 2>gpurun_out/r02_bench_err_$w.log | tail -1 >> gpurun_out/r02_bench_all.jsonl
done
for w in red_nansum_f32 red_nanmean_f32 red_nanvar_f32 red_nanmax_f32 red_nanargmax_f32 red_nansum_f64 red_nanstd_f64 red_nansum_f32_axis0 red_nanvar_f64_axis0 red_nanmean_f32_short red_nansum_f64_all quant_median_long quant_quartiles_short mat_move_cov mat_move_corr; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity 2>gpurun_out/r02_bench_err_$w.log | tail -1 >> gpurun_out/r02_bench_all.jsonl
done
python - <<'PY'
import json
for line in open("gpurun_out/r02_bench_all.jsonl"):
    try: d = json.loads(line)
    except Exception: print("bad line", line[:100]); continue
    e = (d.get("e2e") or {}).get("value", 0); c = (d.get("cpu_baseline") or {}).get("value", 0)
    par = d.get("parity") or {}
    ok = all(v.get("ok", False) for v in par.values() if isinstance(v, dict)) if par else None
    print(f"{d['config']['workload']:28s} {d['value']/1e9:9.1f} Gel/s {d['ms_per_step']:9.3f} ms  roofline {d['roofline']['frac']:.3f}  e2e {e/1e9:7.2f}  cpu {c/1e9:6.2f}  parity {ok}")
PY
# ---- launch list of the default bench command (shares, not absolutes): our kernels only (the first
# 400 launches of the process are torch's input generators)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nbg|group_|scan_|move_|red_" -c 200 --csv --log-file gpurun_out/r02_launches_default_bench_raw.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02_bench_under_ncu.log 2>&1
python scripts/summarise_launches.py gpurun_out/r02_launches_default_bench_raw.csv gpurun_out/r02_launches_default_bench.csv; cat gpurun_out/r02_launches_default_bench.csv
# ---- one `ncu --set full` capture per dominant kernel (NBG_NCU_ONLY="a b c" restricts the list)
for spec in "cfg2_group_nansum group_rowbins2" "cfg2_group_nanmean group_rowbins_kernel" "cfg2_group_nanstd group_rowbins_kernel" "cfg2_group_nanargmax group_rowbins_kernel" \
            "cfg3_ffill scan_rowtile" "cfg3_move_exp_nanmean scan_rowtile" "cfg3_move_exp_nanvar scan_rowtile" "cfg3_move_exp_nancorr scan_rowtile" "cfg3_move_exp_nanmean_f32 scan_rowtile" \
            "cfg1s_move_mean move_rowtile" "cfg4_move_std move_prefix" "cfg4_move_cov move_prefix" "cfg4_move_corr move_rowtile" "cfg4_move_var move_rowtile" \
            "cfg5_group_nansum1d group_atomic" "cfg5_group_nanvar group_atomic" "cfg5_group_nanargmax group_atomic" "cfg5_group_nanfirst group_atomic" \
            "quant_median_long quant_hist" "quant_quartiles_short quant_warp_sort" "mat_move_cov mat_move_seg"; do
  set -- $spec
  if [ -n "$NBG_NCU_ONLY" ] && ! echo " $NBG_NCU_ONLY " | grep -q " $1 "; then continue; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o /tmp/r02_$1 python scripts/prof_workload.py $1 > /dev/null 2>&1
  python scripts/ncu_summary.py /tmp/r02_$1.ncu-rep 14 > gpurun_out/r02_ncu_$1.txt 2>&1
  head -3 gpurun_out/r02_ncu_$1.txt
done
