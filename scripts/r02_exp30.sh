#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -x -k "move and not exp" > gpurun_out/exp30_pytest.log 2>&1; tail -3 gpurun_out/exp30_pytest.log
for v in 0 1; do
  if [ $v = 1 ]; then export NBG_MOVE_BIG_TILES=1; fi
  timeout 300 python bench.py --workload cfg1_move_mean --steps 200 --warmup 5 --no-e2e --no-cpu --no-parity 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('big_tiles=$v', d['ms_per_step']*1000, 'us', d['roofline']['frac'])"
done
