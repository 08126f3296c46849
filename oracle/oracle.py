"""ctypes front-end of the CPU oracle (``oracle/boat_oracle.c``).

TEST INFRASTRUCTURE ONLY: importable from ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product
package (``sac-agent_b200/``) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libboat_oracle.so")

TERM_NAMES = ("", "reached_goal", "out_of_bounds", "out_of_fuel", "timeout", "rudder_broken")


class OracleParams(C.Structure):
    _fields_ = [
        ("experiment", C.c_int), ("test_mode", C.c_int),
        ("dt", C.c_double), ("t_max", C.c_double),
        ("track_width", C.c_double), ("oob_offset", C.c_double), ("goal_line", C.c_double),
        ("fuel", C.c_double),
        ("boat_m", C.c_double), ("boat_m_x", C.c_double), ("boat_m_y", C.c_double),
        ("boat_I", C.c_double), ("boat_Iz", C.c_double),
        ("propeller_diameter", C.c_double), ("wake_friction", C.c_double),
        ("c_r_front", C.c_double), ("c_r_side", C.c_double), ("thrust_deduction", C.c_double),
        ("rho", C.c_double), ("boat_area_front", C.c_double), ("boat_area_side", C.c_double),
        ("boat_l", C.c_double), ("boat_b", C.c_double), ("rudder_area", C.c_double),
        ("fixed_points", C.c_int), ("max_velocity", C.c_double), ("direction", C.c_double),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (a few hundred ms)."""
    src = os.path.join(_HERE, "boat_oracle.c")
    hdr = os.path.join(_HERE, "boat_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-Wall", "-shared", "-o", _LIB_PATH,
             src, "-lm", "-lpthread"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        dp, ip, ucp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_ubyte)
        pp = C.POINTER(OracleParams)
        L.oracle_wind_length.argtypes = [pp]
        L.oracle_wind_length.restype = C.c_int
        L.oracle_random_curve.argtypes = [dp, C.c_int, C.c_int, dp]
        L.oracle_generate_wind.argtypes = [pp, dp, dp, dp, dp]
        L.oracle_env_create.argtypes = [pp]
        L.oracle_env_create.restype = C.c_void_p
        L.oracle_env_destroy.argtypes = [C.c_void_p]
        L.oracle_env_reset.argtypes = [C.c_void_p, C.c_int, dp, dp, dp]
        L.oracle_env_reset_tables.argtypes = [C.c_void_p, C.c_int, dp, dp, dp]
        L.oracle_env_step.argtypes = [C.c_void_p, C.c_double, dp, dp, ip, ip]
        L.oracle_env_all_data.argtypes = [C.c_void_p, dp]
        L.oracle_env_wind_velocity.argtypes = [C.c_void_p]
        L.oracle_env_wind_velocity.restype = dp
        L.oracle_env_wind_angle.argtypes = [C.c_void_p]
        L.oracle_env_wind_angle.restype = dp
        L.oracle_env_episode_reward.argtypes = [C.c_void_p]
        L.oracle_env_episode_reward.restype = C.c_double
        L.oracle_rollout.argtypes = [pp, C.c_int, C.c_int, C.c_int, C.c_int, dp, ip, dp, dp, dp,
                                     ucp, ucp, ip, dp, C.c_int]
        L.oracle_rollout.restype = C.c_longlong
        L.oracle_toy_car.argtypes = [C.c_double] * 4 + [C.c_int, dp, dp]
        L.oracle_toy_car.restype = None
        L.oracle_toy_parachute.argtypes = [C.c_double] * 9 + [C.c_int, dp, dp]
        L.oracle_toy_parachute.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def params_from_config(cfg) -> OracleParams:
    """cfg: nested mapping with the layout of configs/original_config.yaml."""
    b, e, bo, w = cfg["base_settings"], cfg["boat_env"], cfg["boat"], cfg["wind"]
    p = OracleParams()
    p.experiment, p.test_mode = int(b["experiment"]), int(b["test_mode"])
    p.dt, p.t_max = float(b["dt"]), float(b["t_max"])
    p.track_width = float(e["track_width"])
    p.oob_offset = float(e["boat_out_of_bounds_offset"])
    p.goal_line = float(e["goal_line"])
    p.fuel = float(bo["fuel"])
    for k in ("boat_m", "boat_m_x", "boat_m_y", "boat_I", "boat_Iz", "propeller_diameter",
              "wake_friction", "c_r_front", "c_r_side", "thrust_deduction", "rho",
              "boat_area_front", "boat_area_side", "boat_l", "boat_b", "rudder_area"):
        setattr(p, k, float(bo[k]))
    p.fixed_points = int(w["fixed_points"])
    p.max_velocity, p.direction = float(w["max_velocity"]), float(w["direction"])
    return p


def random_curve(knots, L: int) -> np.ndarray:
    knots = np.ascontiguousarray(knots, dtype=np.float64)
    out = np.empty(L, dtype=np.float64)
    rc = lib().oracle_random_curve(_dp(knots), len(knots), L, _dp(out))
    if rc:
        raise ValueError("Please select at least 4 fixed_points in your config.")
    return out


class OracleEnv:
    """Single env with the reference's gym-style reset/step (boat_env.py:67,120)."""

    def __init__(self, params: OracleParams):
        self.params = params
        self.L = lib().oracle_wind_length(C.byref(params))
        self._h = lib().oracle_env_create(C.byref(params))
        if not self._h:
            raise MemoryError
        self.info = {"termination": "", "reached_goal": 0, "out_of_bounds": 0, "out_of_fuel": 0,
                     "rudder_broken": 0, "timeout": 0, "episode_reward": 0}

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_env_destroy(self._h)
            self._h = None

    def reset(self, s_y_start=0, knots_a=None, knots_b=None, wind_tables=None):
        obs = np.empty(11)
        fp = self.params.fixed_points
        if wind_tables is not None:
            wv = np.ascontiguousarray(wind_tables[0], dtype=np.float64)
            wa = np.ascontiguousarray(wind_tables[1], dtype=np.float64)
            assert len(wv) == self.L and len(wa) == self.L
            rc = lib().oracle_env_reset_tables(self._h, int(s_y_start), _dp(wv), _dp(wa), _dp(obs))
        else:
            ka = np.ascontiguousarray(knots_a if knots_a is not None else np.zeros(fp), np.float64)
            kb = np.ascontiguousarray(knots_b if knots_b is not None else np.zeros(fp), np.float64)
            rc = lib().oracle_env_reset(self._h, int(s_y_start), _dp(ka), _dp(kb), _dp(obs))
        if rc:
            raise ValueError(f"oracle reset failed rc={rc}")
        self.info["episode_reward"] = 0
        return obs

    def step(self, action: float):
        obs = np.empty(11)
        r, d, c = C.c_double(), C.c_int(), C.c_int()
        rc = lib().oracle_env_step(self._h, float(action), _dp(obs), C.byref(r), C.byref(d),
                                   C.byref(c))
        if rc:
            raise IndexError("stepped past the wind table")
        if d.value:
            self.info["termination"] = TERM_NAMES[c.value]
            self.info[TERM_NAMES[c.value]] += 1
        self.info["episode_reward"] = lib().oracle_env_episode_reward(self._h)
        return obs, r.value, bool(d.value), c.value

    def all_data(self) -> np.ndarray:
        out = np.empty(8)
        lib().oracle_env_all_data(self._h, _dp(out))
        return out

    def wind(self):
        wv = np.ctypeslib.as_array(lib().oracle_env_wind_velocity(self._h), (self.L,)).copy()
        wa = np.ctypeslib.as_array(lib().oracle_env_wind_angle(self._h), (self.L,)).copy()
        return wv, wa


def rollout(params: OracleParams, actions, s_y_start, knots, auto_reset=False, want_obs=True,
            n_threads=0):
    """actions [T, N] float64; s_y_start [E, N] int32; knots [E, N, 2, fp] float64.
    Returns dict(obs [T,N,11], reward [T,N], done [T,N], term [T,N], ep_len [N],
    final [N,8], steps)."""
    actions = np.ascontiguousarray(actions, dtype=np.float64)
    T, N = actions.shape
    s_y_start = np.ascontiguousarray(np.atleast_2d(s_y_start), dtype=np.int32)
    fp = params.fixed_points
    knots = np.ascontiguousarray(knots, dtype=np.float64).reshape(-1, N, 2, fp)
    E = knots.shape[0]
    assert s_y_start.shape == (E, N)
    obs = np.empty((T, N, 11)) if want_obs else None
    reward = np.empty((T, N)) if want_obs else None
    done = np.empty((T, N), dtype=np.uint8)
    term = np.empty((T, N), dtype=np.uint8)
    ep_len = np.empty(N, dtype=np.int32)
    final = np.empty((N, 8))
    n = lib().oracle_rollout(
        C.byref(params), N, T, E, int(bool(auto_reset)), _dp(actions),
        s_y_start.ctypes.data_as(C.POINTER(C.c_int)), _dp(knots), _dp(obs), _dp(reward),
        done.ctypes.data_as(C.POINTER(C.c_ubyte)), term.ctypes.data_as(C.POINTER(C.c_ubyte)),
        ep_len.ctypes.data_as(C.POINTER(C.c_int)), _dp(final), int(n_threads))
    if n < 0:
        raise RuntimeError(f"oracle_rollout failed rc={n}")
    return dict(obs=obs, reward=reward, done=done, term=term, ep_len=ep_len, final=final, steps=n)


def loop_count(t_max: float, dt: float) -> int:
    """Iterations of the toys' ``t = 0; while t <= t_max: ...; t += dt`` loop
    (toy_car.py:19-32, toy_parachute.py:16-40) with float accumulation."""
    t, n = 0.0, 0
    while t <= t_max:
        n += 1
        t += dt
    return n


def toy_car(accel=10.0, v_limit=10.0, dtheta=0.01, dt=0.1, n_iter=None):
    n_iter = loop_count(500, 0.1) if n_iter is None else n_iter
    traj = np.empty((n_iter, 2))
    out = np.empty(2)
    lib().oracle_toy_car(accel, v_limit, dtheta, dt, n_iter, _dp(traj), _dp(out))
    return traj, out


def toy_parachute(h0=3000.0, h1=1500.0, area_free=0.5, area_chute=25.0, mass=85.0, c_w=1.3,
                  rho=1.2, g=9.81, dt_integrator=0.1, max_iter=None):
    max_iter = loop_count(500, 0.01) if max_iter is None else max_iter
    traj = np.empty((max_iter, 2))
    out = np.empty(2)
    n = lib().oracle_toy_parachute(h0, h1, area_free, area_chute, mass, c_w, rho, g,
                                   dt_integrator, max_iter, _dp(traj), _dp(out))
    return traj[:n], out, n
