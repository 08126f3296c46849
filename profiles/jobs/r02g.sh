#!/bin/bash
# round 2, job g: K>1 loop trimmed (no saturating adds, literal thresholds); ncu of the K = 8 kernel; replay row layout experiment
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02g_gputests.log 2>&1
tail -n 8 gpurun_out/r02g_gputests.log
BENCH_EXTRA_ONLY=k8 timeout 300 python profiles/bench_extra.py > gpurun_out/r02g_extra_k8.jsonl 2> gpurun_out/r02g_extra.err
cat gpurun_out/r02g_extra_k8.jsonl
timeout 300 python profiles/replay_row_layout.py > gpurun_out/r02g_replay_row_layout.jsonl 2>> gpurun_out/r02g_extra.err
cat gpurun_out/r02g_replay_row_layout.jsonl
timeout 300 python profiles/k8_case.py > gpurun_out/r02g_k8_plain.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:boat_step_kernel -s 50 -c 1 -o gpurun_out/r02g_step_k8 python profiles/k8_case.py > gpurun_out/r02g_ncu_k8.log 2>&1
ls -la gpurun_out/r02g_step_k8.ncu-rep
