#!/usr/bin/env python
"""Host-fabric ceiling for the e2e (host-buffer) path: concurrent pinned-memory copies on N GPUs, NO kernels.

    python profiles/host_ceiling.py                                   # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        profiles/host_ceiling.py

Every rank moves exactly what one `boatenv_step_host` call moves for 16,777,216 envs -- 67 MB host->device (actions)
and 822 MB device->host (obs 44 B + reward 4 B + done 1 B per env) -- first each direction alone, then both at
once on two streams (what the chunked pipeline of step_host does), ITERS times, all ranks concurrently.
Rank 0 prints one JSON line: aggregate GB/s per direction and the env-steps/s ceiling those bytes allow
(49 B D2H + 4 B H2D per env-step).  The e2e numbers of bench.py are quoted against this line in DESIGN.md.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

N_ENVS = 16_777_216
ITERS = 10


def main():
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local_rank)
    bind = "--no-bind" not in sys.argv
    cpus = []
    if bind and world > 1:
        import sac_agent_b200 as S
        cpus = S.sharding.bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    d2h_bytes, h2d_bytes = N_ENVS * 49, N_ENVS * 4
    dev_out = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
    host_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    dev_in = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    host_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ITERS):
            fn()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt[0]) / ITERS

    def d2h():
        with torch.cuda.stream(s_out):
            host_out.copy_(dev_out, non_blocking=True)
        s_out.synchronize()

    def h2d():
        with torch.cuda.stream(s_in):
            dev_in.copy_(host_in, non_blocking=True)
        s_in.synchronize()

    def both():
        with torch.cuda.stream(s_in):
            dev_in.copy_(host_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            host_out.copy_(dev_out, non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()

    t_d2h, t_h2d, t_both = timed(d2h), timed(h2d), timed(both)
    if rank == 0:
        print(json.dumps({
            "what": "host ceiling: concurrent pinned copies, no kernels", "n_gpus": world, "envs_per_gpu": N_ENVS,
            "d2h_bytes_per_gpu": d2h_bytes, "h2d_bytes_per_gpu": h2d_bytes, "iters": ITERS,
            "d2h_gbs_aggregate": world * d2h_bytes / t_d2h / 1e9, "h2d_gbs_aggregate": world * h2d_bytes / t_h2d / 1e9,
            "both_ms": 1e3 * t_both, "both_gbs_aggregate": world * (d2h_bytes + h2d_bytes) / t_both / 1e9,
            "env_steps_per_s_ceiling": world * N_ENVS / t_both, "numa_bound_cpus_rank0": len(cpus) or None,
            "host_cpus": os.cpu_count()}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
