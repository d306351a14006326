#!/bin/bash
# matrix segments (tests + timing), config-4 prefix geometry A/B, launch list of the default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_matrix.py -q -x > gpurun_out/exp18_pytest.log 2>&1; tail -5 gpurun_out/exp18_pytest.log
timeout 600 python scripts/r02_quick.py mat > gpurun_out/exp18_mat.jsonl 2> gpurun_out/exp18_mat.err; cat gpurun_out/exp18_mat.jsonl; tail -3 gpurun_out/exp18_mat.err
timeout 600 python scripts/r02_quick.py cfg4g > gpurun_out/exp18_cfg4g.jsonl 2> gpurun_out/exp18_cfg4g.err; cat gpurun_out/exp18_cfg4g.jsonl; tail -3 gpurun_out/exp18_cfg4g.err
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "move" > gpurun_out/exp18_pytest_move.log 2>&1; tail -3 gpurun_out/exp18_pytest_move.log
NBG_PFX_GEOM=1 NBG_PFX=all timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -x -k "move" > gpurun_out/exp18_pytest_move_g1.log 2>&1; tail -3 gpurun_out/exp18_pytest_move_g1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nbg|group_|scan_|move_|red_" -c 200 --csv --log-file gpurun_out/r02_launches_default_bench_raw.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02_bench_under_ncu.log 2>&1
python scripts/summarise_launches.py gpurun_out/r02_launches_default_bench_raw.csv gpurun_out/r02_launches_default_bench.csv; cat gpurun_out/r02_launches_default_bench.csv
