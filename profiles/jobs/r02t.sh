#!/bin/bash
# round 2, job t: regime parity for experiments 1, 3, 4, 5; the complete bench_extra table of the final build; ncu of the reset kernel
timeout 900 python -m pytest tests/test_gpu_benchmark_regime.py -x -q -m gpu > gpurun_out/r02t_gputests.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02t_gputests.log
timeout 600 python profiles/bench_extra.py > gpurun_out/r02t_extra.jsonl 2> gpurun_out/r02t_extra.err
python - <<'PY'
import json
for l in open("gpurun_out/r02t_extra.jsonl"):
    d = json.loads(l); print("  %-50s %.4f ms  %.4g %s  %s" % (d["case"], d["ms"], d["rate"], d["unit"], ("frac %.3f" % d["frac_of_hbm_peak"]) if "frac_of_hbm_peak" in d else ""))
PY
cat > /tmp/reset_case.py <<'PY'
import sys; sys.path.insert(0, ".")
import torch, sac_agent_b200 as S
env = S.BatchedBoatEnv(S.load_config(base_settings__experiment=6), 16 << 20, seed=1, precision="fp32", device=0)
for _ in range(3): env.reset()
torch.cuda.synchronize(); print("ok")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:boat_reset_kernel -s 2 -c 1 -o gpurun_out/r02t_reset python /tmp/reset_case.py > gpurun_out/r02t_ncu_reset.log 2>&1
ls -la gpurun_out/r02t_reset.ncu-rep
