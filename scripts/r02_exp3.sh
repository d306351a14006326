#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/exp3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp3_pytest.log
tail -8 gpurun_out/exp3_pytest.log
timeout 600 python scripts/r02_quick.py cfg2 sweep > gpurun_out/exp3_cfg2.jsonl 2> gpurun_out/exp3_cfg2.err
timeout 600 python scripts/r02_quick.py cfg5 cfg3 --steps 6 > gpurun_out/exp3_cfg5.jsonl 2> gpurun_out/exp3_cfg5.err
cat gpurun_out/exp3_cfg2.jsonl; echo ---; cat gpurun_out/exp3_cfg5.jsonl
for f in gpurun_out/exp3_*.err; do echo $f; tail -n 3 $f; done
