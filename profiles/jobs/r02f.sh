#!/bin/bash
# round 2, job f: rudder view off the critical path, split reset kernels; ncu evidence of the step kernel; SAC on experiment 6
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02f_gputests.log 2>&1
tail -n 12 gpurun_out/r02f_gputests.log
BENCH_EXTRA_ONLY=reset timeout 300 python profiles/bench_extra.py > gpurun_out/r02f_extra_reset.jsonl 2> gpurun_out/r02f_extra.err
BENCH_EXTRA_ONLY=k8 timeout 300 python profiles/bench_extra.py > gpurun_out/r02f_extra_k8.jsonl 2>> gpurun_out/r02f_extra.err
cat gpurun_out/r02f_extra_reset.jsonl gpurun_out/r02f_extra_k8.jsonl
timeout 300 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --no-toys --e2e-k 0 > gpurun_out/r02f_bench_1000.json 2> gpurun_out/r02f_bench_1000.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02f_bench_1000.json")); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"])
PY
# ncu: launch list of the bench command, then one --set full capture of the step kernel (after the plain run above exited 0)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 4 --warmup 3 --no-e2e > gpurun_out/r02f_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:boat_step_kernel -s 1003 -c 2 -o gpurun_out/r02f_step python bench.py --steps 4 --warmup 3 --no-e2e > gpurun_out/r02f_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
timeout 500 python examples/train_sac.py --envs 8192 --iters 40000 --warmup-iters 20 --updates-per-iter 2 --experiment 6 --buffer 33554432 --log-every 1000 --experiments-root gpurun_out/r02f_sac_exp6 > gpurun_out/r02f_sac_exp6.log 2>&1
tail -n 42 gpurun_out/r02f_sac_exp6.log | cut -c1-300
cat gpurun_out/r02f_sac_exp6/setting_6/*/console.csv | cut -c1-150 | head -45
