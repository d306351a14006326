#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_quantile.py -q -x > gpurun_out/exp27_pytest.log 2>&1; tail -3 gpurun_out/exp27_pytest.log
for w in quant_quartiles_short; do
timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity 2>gpurun_out/exp27_err_$w.log | tail -1 > gpurun_out/exp27_$w.json
python -c "
import json; d=json.load(open('gpurun_out/exp27_$w.json')); print(d['config']['workload'], d['ms_per_step'], d['roofline']['frac'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quant_ -s 2 -c 1 -o /tmp/qshort python scripts/prof_workload.py quant_quartiles_short > gpurun_out/exp27_prof.log 2>&1
python scripts/ncu_summary.py /tmp/qshort.ncu-rep 12 > gpurun_out/exp27_ncu_quant_short.txt 2>&1; head -34 gpurun_out/exp27_ncu_quant_short.txt
