#!/bin/bash
# round 2, job h: trimmed K loop (pointer-walked actions, unsigned bounds test, bit-pattern rudder thresholds), optional done_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02h_gputests.log 2>&1
tail -n 8 gpurun_out/r02h_gputests.log
BENCH_EXTRA_ONLY=k8 timeout 300 python profiles/bench_extra.py > gpurun_out/r02h_extra_k8.jsonl 2> gpurun_out/r02h_extra.err
cat gpurun_out/r02h_extra_k8.jsonl
timeout 300 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --no-toys --e2e-k 0 > gpurun_out/r02h_bench_1000.json 2> gpurun_out/r02h_bench_1000.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02h_bench_1000.json")); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"])
PY
