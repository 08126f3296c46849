#!/bin/bash
# round 2, job x: A/B on ONE box: packed fp32 math (product) against the previous revision (libboatenv_nopack.so), K = 8 rows and K = 1
for v in nopack product nopack product; do
  if [ $v = nopack ]; then export BOATENV_LIBRARY=$PWD/sac-agent_b200/libboatenv_nopack.so; else unset BOATENV_LIBRARY; fi
  BENCH_EXTRA_ONLY=k8 timeout 300 python profiles/bench_extra.py > gpurun_out/r02x_k8_$v.jsonl 2>> gpurun_out/r02x.err
  echo "== $v"; python - <<PY
import json
for l in open("gpurun_out/r02x_k8_$v.jsonl"):
    d = json.loads(l); print("  %-45s %.4f ms  %.4g" % (d["case"], d["ms"], d["rate"]))
PY
done
unset BOATENV_LIBRARY
timeout 600 python -m pytest tests/test_reference_callers.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_benchmark_regime.py -x -q -m gpu > gpurun_out/r02x_gputests.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02x_gputests.log
