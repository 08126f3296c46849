#!/bin/bash
# round 2, job d: lane-parallel wind setup (reset / K>1 queue), reference-callers test, bench idle-gap fix, first SAC learning run
timeout 900 python -m pytest tests/test_reference_callers.py tests/test_gpu_edge_cases.py tests/test_gpu_parity.py tests/test_gpu_benchmark_regime.py -m gpu -q > gpurun_out/r02d_gputests.log 2>&1
tail -n 25 gpurun_out/r02d_gputests.log
BENCH_EXTRA_ONLY=reset timeout 300 python profiles/bench_extra.py > gpurun_out/r02d_extra_reset.jsonl 2> gpurun_out/r02d_extra.err
BENCH_EXTRA_ONLY=k8 timeout 300 python profiles/bench_extra.py > gpurun_out/r02d_extra_k8.jsonl 2>> gpurun_out/r02d_extra.err
BENCH_EXTRA_ONLY=fp64 timeout 300 python profiles/bench_extra.py > gpurun_out/r02d_extra_fp64.jsonl 2>> gpurun_out/r02d_extra.err
cat gpurun_out/r02d_extra_reset.jsonl gpurun_out/r02d_extra_k8.jsonl gpurun_out/r02d_extra_fp64.jsonl
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-toys --e2e-k 0 > gpurun_out/r02d_bench_driver.json 2> gpurun_out/r02d_bench_driver.err
timeout 300 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --no-toys --e2e-k 0 > gpurun_out/r02d_bench_1000.json 2> gpurun_out/r02d_bench_1000.err
python - <<'PY'
import json
for f in ("gpurun_out/r02d_bench_driver.json", "gpurun_out/r02d_bench_1000.json"):
    d = json.load(open(f)); print(f, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"], d["episodes_finished_in_timed_region"])
PY
timeout 400 python examples/train_sac.py --envs 8192 --iters 30000 --warmup-iters 20 --updates-per-iter 2 --experiment 1 --log-every 1000 --experiments-root gpurun_out/r02d_sac_exp1 > gpurun_out/r02d_sac_exp1.log 2>&1
tail -n 35 gpurun_out/r02d_sac_exp1.log
cat gpurun_out/r02d_sac_exp1/setting_1/*/console.csv | cut -c1-200
