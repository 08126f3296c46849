#!/bin/bash
# round 2, job s: validation after the shift cap + SAC stability probes (one GPU, 8192 envs, experiment 6)
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02s_gputests.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02s_gputests.log
timeout 600 python profiles/parity_report.py big 32768 2000 1 2 3 > gpurun_out/r02s_parity_big.jsonl 2> gpurun_out/r02s_parity_big.err
python - <<'PY'
import json
for l in open("gpurun_out/r02s_parity_big.jsonl"):
    d = json.loads(l); print(d["experiment"], d["precision"], d["seed"], d["episodes_finished"], d["envs_with_a_termination_mismatch"], d["mismatch_kinds"], "%.2e %.2e" % (d["max_scaled_obs_err_before_first_mismatch"], d["max_scaled_reward_err_before_first_mismatch"]))
PY
run() { name=$1; shift; timeout 400 python examples/train_sac.py --envs 8192 --iters 40000 --warmup-iters 20 --updates-per-iter 2 --experiment 6 --buffer 33554432 --log-every 2000 --experiments-root gpurun_out/r02s_sac_$name "$@" > gpurun_out/r02s_sac_$name.log 2>&1; echo "== $name $@"; cut -d';' -f1-3 gpurun_out/r02s_sac_$name/setting_6/*/console.csv | tr '\n' ' ' | cut -c1-1500; echo; tail -n 1 gpurun_out/r02s_sac_$name/setting_6/*/terminations.csv; }
run alpha001 --set agent__learning_rate_alpha=0.001
run alpha001_rs1 --set agent__learning_rate_alpha=0.001 --set agent__reward_scale=1
run alpha0003 --set agent__learning_rate_alpha=0.0003
