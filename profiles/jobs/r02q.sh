#!/bin/bash
# round 2, job q (2 GPUs): the device-guard test with a handle on device 1 while device 0 is current; bench under torchrun at N = 2
timeout 600 python -m pytest tests/test_gpu_edge_cases.py -x -q -m gpu > gpurun_out/r02q_gputests.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02q_gputests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02q_bench_n2.json 2> gpurun_out/r02q_bench_n2.err; echo "bench rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02q_bench_ref_n2.json 2> gpurun_out/r02q_bench_ref_n2.err; echo "ref rc=$?"
grep '^{' gpurun_out/r02q_bench_n2.json | cut -c1-400; grep '^{' gpurun_out/r02q_bench_ref_n2.json | cut -c1-300
