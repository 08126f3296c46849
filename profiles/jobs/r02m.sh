#!/bin/bash
# round 2, job m (8 GPUs of one box): strong scaling 16,777,216 envs in total at N = 1, 2, 4, 8; host-fabric ceiling at the
# same N; the weak-scaling line at N = 8; end-to-end SAC (BASELINE.json configs[4]: 65536 envs on 8 GPUs) with population rounds
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python bench.py --gpus 1 --steps 1000 --warmup 50 --scaling strong --no-cpu-baseline --no-toys > gpurun_out/r02m_strong_n1.json 2> gpurun_out/r02m_strong_n1.err
for N in 2 4 8; do
  timeout 300 $TR --nproc-per-node $N --master-port $((29500 + N)) bench.py --gpus $N --steps 1000 --warmup 50 --scaling strong --no-cpu-baseline --no-toys > gpurun_out/r02m_strong_n$N.json 2> gpurun_out/r02m_strong_n$N.err
done
timeout 120 python profiles/host_ceiling.py > gpurun_out/r02m_host_ceiling_n1.json 2> gpurun_out/r02m_host_ceiling.err
for N in 2 4 8; do
  timeout 120 $TR --nproc-per-node $N --master-port $((29600 + N)) profiles/host_ceiling.py > gpurun_out/r02m_host_ceiling_n$N.json 2>> gpurun_out/r02m_host_ceiling.err
done
timeout 300 $TR --nproc-per-node 8 --master-port 29700 bench.py --gpus 8 --steps 200 --warmup 20 --no-cpu-baseline --no-toys > gpurun_out/r02m_weak_n8.json 2> gpurun_out/r02m_weak_n8.err
for E in 6 1; do
  timeout 400 $TR --nproc-per-node 8 --master-port $((29800 + E)) examples/train_sac.py --envs 65536 --iters 16000 --warmup-iters 20 --updates-per-iter 2 --experiment $E --buffer 33554432 --tune --pbt-every 2000 --log-every 1000 --experiments-root gpurun_out/r02m_sac_exp$E > gpurun_out/r02m_sac_exp$E.log 2>&1
  tail -n 3 gpurun_out/r02m_sac_exp$E.log | cut -c1-1500
  cat gpurun_out/r02m_sac_exp$E/setting_$E/overview.csv
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02m_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    if "roofline" in d:
        print(f, "N", d["n_gpus"], d["scaling"], "value %.4g" % d["value"], "ms %.4f" % d["ms_per_step"], "kernel %.4f" % d["roofline"]["kernel_ms"],
              "frac %.3f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"], "e2e_k %.4g" % (d["e2e_k"] or {}).get("value", 0))
    else:
        print(f, {k: d[k] for k in ("n_gpus", "d2h_gbs_aggregate", "h2d_gbs_aggregate", "both_gbs_aggregate", "env_steps_per_s_ceiling") if k in d})
PY
