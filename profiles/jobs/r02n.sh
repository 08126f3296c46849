#!/bin/bash
# round 2, job n (8 GPUs): configs[4] with the reference's own hyper-parameters: 65536 envs, 8 replicas (a) independent, (b) with population rounds
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29811 examples/train_sac.py --envs 65536 --iters 10000 --warmup-iters 20 --updates-per-iter 2 --experiment 6 --buffer 33554432 --log-every 1000 --experiments-root gpurun_out/r02n_sac_exp6_replicas > gpurun_out/r02n_sac_exp6_replicas.log 2>&1
tail -n 2 gpurun_out/r02n_sac_exp6_replicas.log | cut -c1-1200
cat gpurun_out/r02n_sac_exp6_replicas/setting_6/overview.csv
timeout 300 $TR --nproc-per-node 8 --master-port 29812 examples/train_sac.py --envs 65536 --iters 10000 --warmup-iters 20 --updates-per-iter 2 --experiment 6 --buffer 33554432 --pbt-every 2500 --log-every 1000 --experiments-root gpurun_out/r02n_sac_exp6_pbt > gpurun_out/r02n_sac_exp6_pbt.log 2>&1
tail -n 2 gpurun_out/r02n_sac_exp6_pbt.log | cut -c1-1500
cat gpurun_out/r02n_sac_exp6_pbt/setting_6/overview.csv
for d in gpurun_out/r02n_sac_exp6_replicas gpurun_out/r02n_sac_exp6_pbt; do for f in $d/setting_6/*/terminations.csv; do tail -n 1 $f; done; done
