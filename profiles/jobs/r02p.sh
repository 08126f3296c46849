#!/bin/bash
# round 2, job p: upper bound of what a cheaper in-kernel wind setup could gain in the K = 1 kernel: ablation build that pops the
# setup requests and drops them (-DBOAT_DEBUG_SKIP_SETUP; results are wrong, only the timing is of interest)
for v in product nosetup product nosetup; do
  if [ $v = product ]; then unset BOATENV_LIBRARY; else export BOATENV_LIBRARY=$PWD/sac-agent_b200/libboatenv_$v.so; fi
  timeout 200 python bench.py --steps 600 --warmup 100 --no-e2e > gpurun_out/r02p_bench_$v.json 2>> gpurun_out/r02p.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02p_bench_$v.json") if l.startswith("{")][-1]); print("$v", "ms/step %.4f  kernel %.4f  clocks %s" % (d["ms_per_step"], d["kernel_ms"], d["clocks"]["sm_mhz"]))
PY
done
