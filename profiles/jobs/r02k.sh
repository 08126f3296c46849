#!/bin/bash
# round 2, job k: A/B on one box: small-argument trig fast path (product) vs always-reduce (libboatenv_notrig.so): K = 8 rows, K = 1; bit identity
for v in notrig product notrig product; do
  if [ $v = product ]; then unset BOATENV_LIBRARY; else export BOATENV_LIBRARY=$PWD/sac-agent_b200/libboatenv_$v.so; fi
  BENCH_EXTRA_ONLY=k8 timeout 300 python profiles/bench_extra.py > gpurun_out/r02k_k8_$v.jsonl 2>> gpurun_out/r02k.err
  echo "== $v"; python - <<PY
import json
for l in open("gpurun_out/r02k_k8_$v.jsonl"):
    d = json.loads(l); print("  %-45s %.4f ms  %.4g" % (d["case"], d["ms"], d["rate"]))
PY
  timeout 200 python bench.py --steps 600 --warmup 100 --no-e2e > gpurun_out/r02k_bench_$v.json 2>> gpurun_out/r02k.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02k_bench_$v.json") if l.startswith("{")][-1]); print("  K=1 bench", "ms/step %.4f  kernel %.4f  clocks %s" % (d["ms_per_step"], d["kernel_ms"], d["clocks"]["sm_mhz"]))
PY
done
unset BOATENV_LIBRARY
# bit identity of the two builds: 300 steps of 100k envs, big actions (rudder beyond 1.5 rad never happens while alive; s_r stays small)
python - <<'PY'
import os, subprocess, sys
code = '''
import sys; sys.path.insert(0, ".")
import torch, sac_agent_b200 as S
env = S.BatchedBoatEnv(S.load_config(base_settings__experiment=6), 100_000, seed=3, precision="fp32", device=0, auto_reset=True)
env.reset(); acc = torch.zeros(100_000, 11, device="cuda", dtype=torch.float64)
for t in range(300):
    o, r, d, i = env.step(env.uniform_actions(t, 2.0)); acc += o.double() * (t + 1)
print(float(acc.sum()), float(acc.abs().max()))
'''
outs = []
for lib in (None, os.path.join(os.getcwd(), "sac-agent_b200", "libboatenv_notrig.so")):
    e = dict(os.environ)
    if lib: e["BOATENV_LIBRARY"] = lib
    outs.append(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e).stdout.strip())
print("bit identity product vs notrig:", outs[0] == outs[1], outs)
PY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchmark_regime.py -x -q -m gpu > gpurun_out/r02k_gputests.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/r02k_gputests.log
