#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/exp5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp5_pytest.log
tail -6 gpurun_out/exp5_pytest.log
timeout 600 python scripts/r02_quick.py cfg2 sweep > gpurun_out/exp5_cfg2.jsonl 2> gpurun_out/exp5_cfg2.err
cat gpurun_out/exp5_cfg2.jsonl
NBG_RB2_S=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:group_rowbins2 -s 2 -c 1 -o /tmp/rb2 python scripts/prof_workload.py cfg2_group_nansum > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/rb2.ncu-rep 24 > gpurun_out/exp5_ncu_rb2.txt 2>&1
cp /tmp/rb2.ncu-rep gpurun_out/exp5_rb2.ncu-rep
cat gpurun_out/exp5_ncu_rb2.txt
for f in gpurun_out/exp5_*.err; do echo $f; tail -n 5 $f; done
