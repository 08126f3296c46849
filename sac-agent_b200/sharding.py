"""Multi-GPU plumbing: env instances shard trivially -- rank r owns the
contiguous block of global env ids [r*N/G, (r+1)*N/G), one process per GPU, no data-path
collective.  The only exchange is an occasional all-reduce of the 8 episode statistics
(64 bytes) over ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os

COUNTER_NAMES = ("reached_goal", "out_of_bounds", "out_of_fuel", "timeout", "rudder_broken",
                 "episodes", "return_sum", "return_sumsq")


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """(env_id_offset, n_local) of rank ``rank``: contiguous, covers [0, n_total) exactly."""
    if not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad rank/world/n_total")
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi - lo


def dist_info() -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (defaults 0, 0, 1)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def all_reduce_counters(counters, group=None) -> dict:
    """Sum the 8-element statistics vector over all ranks (in place) and name it.
    ``counters``: float64 tensor[8] -- ``BatchedBoatEnv.counters_tensor()`` on a GPU rank."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    vals = counters.detach().cpu().tolist()
    out = dict(zip(COUNTER_NAMES, vals))
    n = out["episodes"]
    out["return_mean"] = out["return_sum"] / n if n else 0.0
    out["return_var"] = max(out["return_sumsq"] / n - out["return_mean"] ** 2, 0.0) if n else 0.0
    return out


def make_sharded_env(config, n_total: int, seed: int = 0, precision: str = "fp32", auto_reset: bool = True):
    """This rank's shard of an ``n_total``-env population as a ``BatchedBoatEnv`` on
    cuda:LOCAL_RANK.  Global env ids key the Philox streams, so per-env trajectories do
    not depend on the number of GPUs."""
    from .boat_env import BatchedBoatEnv
    rank, local_rank, world = dist_info()
    offset, n_local = shard_range(n_total, rank, world)
    return BatchedBoatEnv(config, n_local, seed=seed, precision=precision, device=local_rank,
                          env_id_offset=offset, auto_reset=auto_reset)


def bind_to_gpu_numa_node(cuda_index: int) -> list[int]:
    """Pin this process to the CPUs NVML reports as local to its GPU, so that pinned host buffers
    allocated afterwards (first touch) and the threads that drive the copies sit on the GPU's NUMA
    node.  Matters for the host-buffer path (``step_host``) when 8 ranks share one host.  Returns the
    CPU list (empty if NVML or the affinity call is unavailable -- then nothing changes)."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(cuda_index)
        bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return []
