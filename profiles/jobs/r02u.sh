#!/bin/bash
# round 2, job u: A/B on one box: action prefetch distance 2 (product) vs 1 (libboatenv_pf1.so) in the K > 1 kernel; step_k tests
for v in pf1 product pf1 product; do
  if [ $v = pf1 ]; then export BOATENV_LIBRARY=$PWD/sac-agent_b200/libboatenv_pf1.so; else unset BOATENV_LIBRARY; fi
  BENCH_EXTRA_ONLY=k8 timeout 300 python profiles/bench_extra.py > gpurun_out/r02u_k8_$v.jsonl 2>> gpurun_out/r02u.err
  echo "== $v"; python - <<PY
import json
for l in open("gpurun_out/r02u_k8_$v.jsonl"):
    d = json.loads(l); print("  %-45s %.4f ms  %.4g" % (d["case"], d["ms"], d["rate"]))
PY
done
unset BOATENV_LIBRARY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchmark_regime.py tests/test_gpu_edge_cases.py -x -q -m gpu -k "step_k or lane_parallel" > gpurun_out/r02u_gputests.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02u_gputests.log
