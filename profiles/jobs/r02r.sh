#!/bin/bash
# round 2, job r: A/B 3 resident CTAs per SM (libboatenv_mb3.so) for the fp32 K = 1 kernels; env_state_host event ordering test
for v in product mb3 product mb3; do
  if [ $v = product ]; then unset BOATENV_LIBRARY; else export BOATENV_LIBRARY=$PWD/sac-agent_b200/libboatenv_$v.so; fi
  BENCH_EXTRA_ONLY=fp32 timeout 300 python profiles/bench_extra.py > gpurun_out/r02r_fp32_$v.jsonl 2>> gpurun_out/r02r.err
  echo "== $v"; python - <<PY
import json
for l in open("gpurun_out/r02r_fp32_$v.jsonl"):
    d = json.loads(l); print("  %-30s %.4f ms  %.4g  frac %.3f" % (d["case"], d["ms"], d["rate"], d["frac_of_hbm_peak"]))
PY
done
unset BOATENV_LIBRARY
timeout 600 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py -x -q -m gpu -k "env_state or single_env or config0 or drop_in or latency" > gpurun_out/r02r_gputests.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02r_gputests.log
