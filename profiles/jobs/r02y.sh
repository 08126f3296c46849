#!/bin/bash
# round 2, job y: validation of HEAD (driver's commands) + tuning-knob A/B of the K = 1 kernel on one box + 10 more parity seeds
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02y_gputests.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02y_gputests.log
for v in product sw2 sw1 st2 sw2st2 product; do
  if [ $v = product ]; then unset BOATENV_LIBRARY; else export BOATENV_LIBRARY=$PWD/sac-agent_b200/libboatenv_$v.so; fi
  timeout 200 python bench.py --steps 600 --warmup 100 --no-e2e > gpurun_out/r02y_bench_$v.json 2>> gpurun_out/r02y.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02y_bench_$v.json") if l.startswith("{")][-1]); print("$v", "ms/step %.4f  kernel %.4f  clocks %s" % (d["ms_per_step"], d["kernel_ms"], d["clocks"]["sm_mhz"]))
PY
done
unset BOATENV_LIBRARY
timeout 900 python profiles/parity_report.py big 32768 2000 6 7 8 9 10 11 12 13 14 15 > gpurun_out/r02y_parity_big.jsonl 2> gpurun_out/r02y_parity_big.err
python - <<'PY'
import json
eps = mm = 0
for l in open("gpurun_out/r02y_parity_big.jsonl"):
    d = json.loads(l)
    if d["experiment"] == 6 and d["precision"] == "fp32":
        eps += d["episodes_finished"]; mm += d["envs_with_a_termination_mismatch"]
    else:
        print(d["experiment"], d["precision"], d["episodes_finished"], d["envs_with_a_termination_mismatch"], d["mismatch_kinds"])
print("fp32 exp 6: episodes", eps, "mismatching envs", mm)
PY
