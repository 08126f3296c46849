#!/bin/bash
# round 2, final validation: exactly what the driver runs at round end on a fresh box
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02fin_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r02fin_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02fin_gputests.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02fin_gputests.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02fin_bench_reference.json 2> gpurun_out/r02fin_bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02fin_bench.json 2> gpurun_out/r02fin_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02fin_bench.json")); r = json.load(open("gpurun_out/r02fin_bench_reference.json"))
print("ours value %.4g ms %.4f kernel %.4f frac %.3f e2e %.4g e2e_k %.4g" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e_k"]["value"]), d["clocks"])
print("cpu_baseline", {k: v for k, v in d["cpu_baseline"].items() if k not in ("sample", "sample_1core", "port")}, "port", d["cpu_baseline"]["port"]["value"])
print("ref value %.4g" % r["value"], r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"], "e2e ratio %.0f" % (d["e2e"]["value"] / r["value"]))
PY
