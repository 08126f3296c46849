"""Batched versions of the reference's two integrator demos
(environment/toy_car.py:7-33, environment/toy_parachute.py:8-41).  The scripts' constants
are the default per-env parameters; ``jitter`` spreads them by +-jitter (Philox, env 0 is
never jittered), so env 0 with defaults reproduces the scripts' numbers.
"""
from __future__ import annotations

import ctypes as C

from . import _lib

CAR_DEFAULTS = dict(accel=10.0, v_limit=10.0, dtheta=0.01, dt=0.1)           # toy_car.py:7-8,11,23; Integrator.dt
PARACHUTE_DEFAULTS = dict(h0=3000.0, h1=1500.0, area_free=0.5, area_chute=25.0, mass=85.0, c_w=1.3,
                          rho=1.2, g=9.81, dt_integrator=0.1)                  # toy_parachute.py:8-15


def loop_count(t_max: float, dt: float) -> int:
    """Iterations of the scripts' ``t = 0; while t <= t_max: ...; t += dt`` loops."""
    t, n = 0.0, 0
    while t <= t_max:
        n += 1
        t += dt
    return n


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("sac_agent_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


class _Toy:
    KIND = -1
    DEFAULTS: dict = {}

    def __init__(self, n_envs=1, jitter=0.0, seed=0, precision="fp32", device=None, **params):
        torch = _torch()
        unknown = set(params) - set(self.DEFAULTS)
        if unknown:
            raise TypeError(f"unknown parameters {sorted(unknown)}")
        self.params = {**self.DEFAULTS, **params}
        self._jitter, self._seed = float(jitter), int(seed)
        self.n_envs = int(n_envs)
        self.precision = {"fp32": 32, "fp64": 64, 32: 32, 64: 64}[precision]
        self.dtype = torch.float32 if self.precision == 32 else torch.float64
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        vals = [float(self.params[k]) for k in self.DEFAULTS]
        arr = (C.c_double * len(vals))(*vals)
        self._L = _lib.lib()
        h = C.c_void_p()
        _lib.check(self._L.boattoy_create(self.KIND, self.n_envs, arr, len(vals), float(jitter), int(seed),
                                          self.precision, self.device.index, C.byref(h)), "boattoy_create")
        self._h = h
        self.out = torch.empty((self.n_envs, 4), dtype=self.dtype, device=self.device)
        self.done = torch.empty(self.n_envs, dtype=torch.uint8, device=self.device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.boattoy_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def reset(self):
        _lib.check(self._L.boattoy_reset(self._h, self._stream()), "boattoy_reset")

    def env_params(self, env_begin=0, n=None):
        """The (jittered) parameters envs [env_begin, env_begin + n) run with: float64 [n, len(DEFAULTS)] in the
        order of ``DEFAULTS`` (host computation, the same Philox expression as the kernels)."""
        import numpy as np
        n = self.n_envs - int(env_begin) if n is None else int(n)
        vals = [float(self.params[k]) for k in self.DEFAULTS]
        arr = (C.c_double * len(vals))(*vals)
        out = np.empty((n, len(vals)), dtype=np.float64)
        _lib.check(self._L.boattoy_params_host(self.KIND, arr, len(vals), self._jitter, self._seed, int(env_begin), n,
                                               out.ctypes.data), "boattoy_params_host")
        return out

    def step(self, k=1):
        """k loop iterations per env; returns (out[n_envs, 4], done[n_envs])."""
        _lib.check(self._L.boattoy_step(self._h, int(k), self.out.data_ptr(), self.done.data_ptr(),
                                        self._stream()), "boattoy_step")
        return self.out, self.done


class ToyCar(_Toy):
    """out columns: s_x, s_y, v, car_angle."""
    KIND = 0
    DEFAULTS = CAR_DEFAULTS


class ToyParachute(_Toy):
    """out columns: s, v, total_a, integrator calls; done = ground reached (s < 0)."""
    KIND = 1
    DEFAULTS = PARACHUTE_DEFAULTS
