"""The reference's own per-episode CPU stepping, timed on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm).  Every
worker process imports the UNMODIFIED reference ``BoatEnv`` (environment/boat_env.py:9) from ``/root/reference`` or
from the staged copy ``oracle/_ref`` (oracle/make_ref.py) under the stub modules of ``oracle/ref_shim.py`` and runs
the loop of main.py:70-99 without the agent: ``reset()``; ``step(action)`` with float32-representable uniform(-1,1) actions
(policy A1 of SURVEY.md 8d); ``reset()`` again when done.  One env per process -- the reference's own fan-out
style (main.py:215-235 starts one Process per model).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

_ENV = None
_RNG = None


def _worker_init(experiment: int, seed_base: int) -> None:
    global _ENV, _RNG
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    import numpy as np
    from oracle import ref_shim as R
    cfg = R.load_config(base_settings__experiment=experiment)
    seed = seed_base + (os.getpid() % 100003)
    np.random.seed(seed % (2 ** 32))
    _ENV = R.make_env(cfg)
    _ENV.reset()
    _RNG = np.random.default_rng(seed)


def _worker_run(n_steps: int):
    """n_steps env-steps of the reference loop (resets included in the time).  Returns (steps, seconds, episodes)."""
    import numpy as np
    env, rng = _ENV, _RNG
    # float32-representable actions passed as float64 (a float32 action would make
    # the reference accumulate its rudder in float32)
    acts = rng.uniform(-1.0, 1.0, size=n_steps).astype(np.float32).astype(np.float64)
    episodes = 0
    t0 = time.perf_counter()
    for k in range(n_steps):
        _, _, done, _ = env.step(acts[k:k + 1])
        if done:
            env.reset()
            episodes += 1
    return n_steps, time.perf_counter() - t0, episodes


def available() -> bool:
    from oracle import ref_shim as R
    return R.reference_available()


class ReferencePool:
    """P processes, one reference BoatEnv each."""

    def __init__(self, processes: int, experiment: int = 6, seed: int = 1):
        self.processes = int(processes)
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.processes, initializer=_worker_init, initargs=(experiment, seed))

    def run(self, steps_per_process: int):
        """Every process does steps_per_process env-steps concurrently.  Returns (total steps, wall seconds,
        episodes) -- the aggregate rate is total / wall."""
        t0 = time.perf_counter()
        res = self.pool.map(_worker_run, [int(steps_per_process)] * self.processes, chunksize=1)
        wall = time.perf_counter() - t0
        return sum(r[0] for r in res), wall, sum(r[2] for r in res)

    def close(self):
        self.pool.close()
        self.pool.join()


def measure(processes: int, seconds: float, experiment: int = 6, seed: int = 1) -> dict:
    """Aggregate env-steps/s of `processes` reference envs over about `seconds` of wall time."""
    pool = ReferencePool(processes, experiment, seed)
    try:
        n, wall, _ = pool.run(2000)                      # warm-up (one episode or so) + calibration
        per_proc = max(2000, int(n / processes / wall * seconds))
        n, wall, eps = pool.run(per_proc)
    finally:
        pool.close()
    return {"value": n / wall, "processes": processes, "steps": n, "seconds": wall, "episodes": eps}
