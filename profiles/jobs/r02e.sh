#!/bin/bash
# round 2, job e: rudder carried as an integer-valued double (cheaper K loop), hybrid reset kernel; SAC with a ring that holds > 1 episode
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02e_gputests.log 2>&1
tail -n 25 gpurun_out/r02e_gputests.log
BENCH_EXTRA_ONLY=reset timeout 300 python profiles/bench_extra.py > gpurun_out/r02e_extra_reset.jsonl 2> gpurun_out/r02e_extra.err
BENCH_EXTRA_ONLY=k8 timeout 300 python profiles/bench_extra.py > gpurun_out/r02e_extra_k8.jsonl 2>> gpurun_out/r02e_extra.err
cat gpurun_out/r02e_extra_reset.jsonl gpurun_out/r02e_extra_k8.jsonl
timeout 300 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --no-toys --e2e-k 0 > gpurun_out/r02e_bench_1000.json 2> gpurun_out/r02e_bench_1000.err
python - <<'PY'
import json
for f in ("gpurun_out/r02e_bench_1000.json",):
    d = json.load(open(f)); print(f, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"], d["episodes_finished_in_timed_region"])
PY
timeout 500 python examples/train_sac.py --envs 8192 --iters 40000 --warmup-iters 20 --updates-per-iter 2 --experiment 1 --buffer 33554432 --log-every 1000 --experiments-root gpurun_out/r02e_sac_exp1 > gpurun_out/r02e_sac_exp1.log 2>&1
tail -n 45 gpurun_out/r02e_sac_exp1.log | cut -c1-400
