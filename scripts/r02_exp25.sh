#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/exp25_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/exp25_pytest.log; tail -6 gpurun_out/exp25_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/exp25_smoke.log 2>&1; tail -2 gpurun_out/exp25_smoke.log
