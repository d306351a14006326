#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_quantile.py -q -x > gpurun_out/exp21_pytest.log 2>&1; tail -8 gpurun_out/exp21_pytest.log
for w in quant_median_long quant_quartiles_short; do
timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity 2>gpurun_out/exp21_err_$w.log | tail -1 > gpurun_out/exp21_$w.json
python -c "
import json; d=json.load(open('gpurun_out/exp21_$w.json')); print(d['config']['workload'], d['ms_per_step'], d['roofline']['frac'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:quant -c 60 --csv --log-file gpurun_out/exp21_launches_quant_raw.csv python scripts/prof_workload.py quant_median_long > /dev/null 2>&1
python scripts/summarise_launches.py gpurun_out/exp21_launches_quant_raw.csv gpurun_out/exp21_launches_quant.csv; cat gpurun_out/exp21_launches_quant.csv
