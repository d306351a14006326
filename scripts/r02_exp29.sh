#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/r02_quick.py cfg2nw --steps 8 > gpurun_out/exp29_cfg2.jsonl 2> gpurun_out/exp29_cfg2.err; cat gpurun_out/exp29_cfg2.jsonl; tail -3 gpurun_out/exp29_cfg2.err
