#!/bin/bash
# round 2, validation job: what the driver runs at round end (smoke, GPU tests with -x, both bench arms) + the fp64 rows
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02v_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r02v_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02v_gputests.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r02v_gputests.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02v_bench_reference.json 2> gpurun_out/r02v_bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "bench rc=$?"
BENCH_EXTRA_ONLY=fp64 timeout 300 python profiles/bench_extra.py > gpurun_out/r02v_extra_fp64.jsonl 2> gpurun_out/r02v_extra.err
cat gpurun_out/r02v_extra_fp64.jsonl
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02v_bench.json")); r = json.load(open("gpurun_out/r02v_bench_reference.json"))
print("ours", d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e_k"]["value"], d["clocks"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["value"], d["cpu_baseline"]["port"]["value"])
print("ref", r["value"], r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"], "e2e ratio", d["e2e"]["value"] / r["value"])
PY
