#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_matrix.py -q -x > gpurun_out/exp20_pytest.log 2>&1; tail -3 gpurun_out/exp20_pytest.log
timeout 600 python scripts/r02_quick.py mat > gpurun_out/exp20_mat.jsonl 2> gpurun_out/exp20_mat.err; cat gpurun_out/exp20_mat.jsonl; tail -3 gpurun_out/exp20_mat.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mat_move_seg -s 2 -c 1 -o /tmp/mat_move python scripts/prof_mat.py move_covmatrix > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/mat_move.ncu-rep 40 > gpurun_out/exp20_ncu_mat_move_cov.txt 2>&1; head -30 gpurun_out/exp20_ncu_mat_move_cov.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mat_exp_seg -s 5 -c 1 -o /tmp/mat_exp python scripts/prof_mat.py move_exp_nancovmatrix > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/mat_exp.ncu-rep 40 > gpurun_out/exp20_ncu_mat_exp_cov.txt 2>&1; head -30 gpurun_out/exp20_ncu_mat_exp_cov.txt
