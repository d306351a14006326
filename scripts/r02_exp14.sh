#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/exp14_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp14_pytest.log
tail -8 gpurun_out/exp14_pytest.log
timeout 400 python scripts/r02_quick.py cfg3 > gpurun_out/exp14_cfg3.jsonl 2> gpurun_out/exp14_cfg3.err
NBG_EXP_GATE=1 timeout 400 python scripts/r02_quick.py cfg3 > gpurun_out/exp14_cfg3_gate.jsonl 2>> gpurun_out/exp14_cfg3.err
cat gpurun_out/exp14_cfg3.jsonl; echo; cat gpurun_out/exp14_cfg3_gate.jsonl
tail -n 5 gpurun_out/exp14_cfg3.err
