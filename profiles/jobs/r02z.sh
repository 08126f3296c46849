#!/bin/bash
# round 2, job z: 2 setup warps per CTA adopted: GPU suite, the driver's bench call, a 1000-step line, setup-queue flood test timing, fp32 rows
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02z_gputests.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02z_gputests.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --no-toys > gpurun_out/r02z_bench_1000.json 2> gpurun_out/r02z_bench_1000.err
BENCH_EXTRA_ONLY=fp32 timeout 300 python profiles/bench_extra.py > gpurun_out/r02z_extra_fp32.jsonl 2> gpurun_out/r02z_extra.err
python - <<'PY'
import json
for f in ("gpurun_out/r02z_bench.json", "gpurun_out/r02z_bench_1000.json"):
    d = json.load(open(f)); print(f, "value %.4g ms %.4f kernel %.4f frac %.3f whole %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"], d["e2e"]["value"]), d["clocks"])
for l in open("gpurun_out/r02z_extra_fp32.jsonl"):
    d = json.loads(l); print("  %-30s %.4f ms  %.4g  frac %.3f" % (d["case"], d["ms"], d["rate"], d["frac_of_hbm_peak"]))
PY
timeout 120 python profiles/determinism_check.py 10 > gpurun_out/r02z_det.log 2>&1; tail -n 1 gpurun_out/r02z_det.log
timeout 120 python profiles/determinism_check.py 10 6 fp32 fused >> gpurun_out/r02z_det.log 2>&1; tail -n 1 gpurun_out/r02z_det.log
timeout 120 python profiles/determinism_check.py 10 6 fp32 k8 >> gpurun_out/r02z_det.log 2>&1; tail -n 1 gpurun_out/r02z_det.log
