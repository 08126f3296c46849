#!/bin/bash
# round 2, job w: packed fp32 math (FFMA2 / FMUL2) in the fp32 sub-step: GPU suite, K = 8 rows, K = 1 line, parity report
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02w_gputests.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r02w_gputests.log
BENCH_EXTRA_ONLY=k8 timeout 300 python profiles/bench_extra.py > gpurun_out/r02w_extra_k8.jsonl 2> gpurun_out/r02w_extra.err
cat gpurun_out/r02w_extra_k8.jsonl
timeout 300 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --no-toys --e2e-k 0 > gpurun_out/r02w_bench_1000.json 2> gpurun_out/r02w_bench_1000.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02w_bench_1000.json")); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"])
PY
timeout 900 python profiles/parity_report.py big 32768 2000 1 2 3 4 5 > gpurun_out/r02w_parity_big.jsonl 2> gpurun_out/r02w_parity_big.err
cut -c1-420 gpurun_out/r02w_parity_big.jsonl
