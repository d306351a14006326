#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/exp4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp4_pytest.log
tail -6 gpurun_out/exp4_pytest.log
timeout 600 python scripts/r02_quick.py cfg2 sweep > gpurun_out/exp4_cfg2.jsonl 2> gpurun_out/exp4_cfg2.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/exp4_bench_default.json 2> gpurun_out/exp4_bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/exp4_bench_ref.json 2> gpurun_out/exp4_bench_ref.err
for w in cfg3_move_exp_nanmean cfg3_bfill cfg4_move_std cfg5_group_nanvar cfg1_move_mean; do
  timeout 600 python bench.py --steps 5 --warmup 3 --workload $w > gpurun_out/exp4_bench_$w.json 2> gpurun_out/exp4_bench_$w.err
done
cat gpurun_out/exp4_cfg2.jsonl
for f in gpurun_out/exp4_bench_*.json; do echo $f; cat $f; done
for f in gpurun_out/exp4_*.err; do echo $f; tail -n 5 $f; done
