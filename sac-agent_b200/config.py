"""original_config.yaml semantics (utils/config_reader.py:6-14 of the reference).

The reference reads its YAML into a ``DotMap``; ``AttrDict`` offers the same attribute
access, and every entry point here also accepts a real DotMap or a plain dict.
"""
from __future__ import annotations

import copy
import ctypes as C

import yaml

# Every key of the reference's configs/original_config.yaml with the value it ships (:2-65): the hot path reads
# base_settings.test_mode/dt/t_max/experiment, boat_env.*, most of boat.* and wind.*; the SAC example reads the
# `agent:` block; the rest (n_games, render_skip_size, avg_lookback, boat.n_max/w/aspect_ratio/a/b,
# agent.layer*_size) is carried because the reference's post-processing reads a run's configs/tuned_configs.yaml
# (rendering/boat_env_render.py:31, postprocessing/replayer.py:52 use base_settings.avg_lookback).
DEFAULTS = {
    "base_settings": {"test_mode": 0, "n_games": 250, "render_skip_size": 50, "avg_lookback": 50, "dt": 0.25,
                      "t_max": 2500, "experiment": 5},
    "agent": {"learning_rate_alpha": 0.005, "learning_rate_beta": 0.0003, "gamma": 0.99,
              "tvn_parameter_modulation_tau": 0.005, "max_size": 1_000_000, "layer1_size": 256, "layer2_size": 256,
              "batch_size": 1024, "reward_scale": 10},
    "boat_env": {"track_width": 800, "boat_out_of_bounds_offset": 0, "goal_line": 3900},
    "boat": {"n_max": 30, "fuel": 15000, "w": 0.3, "boat_m": 600, "boat_m_x": 50, "boat_m_y": 100,
             "boat_I": 6_000_000, "boat_Iz": 10, "propeller_diameter": 1, "wake_friction": 0.3, "c_r_front": 0.31,
             "c_r_side": 2, "thrust_deduction": 0.3, "rho": 1, "boat_area_front": 20, "boat_area_side": 90,
             "boat_l": 15, "boat_b": 6, "rudder_area": 10, "aspect_ratio": 2, "a": 4.252, "b": 0.262},
    "wind": {"fixed_points": 8, "max_velocity": 0.5, "direction": 90},
}


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
            self[k] = v
        return v

    def __setattr__(self, k, v):
        self[k] = v


def load_config(path: str | None = None, **overrides) -> AttrDict:
    """get_config (utils/config_reader.py:6-8): a reference YAML (original_config.yaml,
    tuned_configs.yaml) is read unchanged; without a path the defaults above are used.
    ``overrides``: section__key=value."""
    if path is None:
        cfg = copy.deepcopy(DEFAULTS)
    else:
        with open(path) as f:
            cfg = yaml.safe_load(f)
    for k, v in overrides.items():
        sec, key = k.split("__", 1)
        cfg[sec][key] = v
    return AttrDict(cfg)


class BoatEnvParams(C.Structure):
    """Mirror of ``boatenv_params`` (include/boatenv.h)."""
    _fields_ = [
        ("experiment", C.c_int32), ("test_mode", C.c_int32),
        ("dt", C.c_double), ("t_max", C.c_double),
        ("track_width", C.c_double), ("oob_offset", C.c_double), ("goal_line", C.c_double),
        ("fuel", C.c_double),
        ("boat_m", C.c_double), ("boat_m_x", C.c_double), ("boat_m_y", C.c_double),
        ("boat_I", C.c_double), ("boat_Iz", C.c_double),
        ("propeller_diameter", C.c_double), ("wake_friction", C.c_double),
        ("c_r_front", C.c_double), ("c_r_side", C.c_double), ("thrust_deduction", C.c_double),
        ("rho", C.c_double), ("boat_area_front", C.c_double), ("boat_area_side", C.c_double),
        ("boat_l", C.c_double), ("boat_b", C.c_double), ("rudder_area", C.c_double),
        ("fixed_points", C.c_int32), ("_pad", C.c_int32),
        ("max_velocity", C.c_double), ("direction", C.c_double),
    ]


def _get(cfg, sec, key):
    s = cfg[sec] if isinstance(cfg, dict) else getattr(cfg, sec)
    return s[key] if isinstance(s, dict) else getattr(s, key)


def params_from_config(cfg) -> BoatEnvParams:
    """The keys the hot path reads; everything else is ignored,
    as in the reference."""
    p = BoatEnvParams()
    p.experiment = int(_get(cfg, "base_settings", "experiment"))  # wind.py:30 int(...)
    p.test_mode = int(_get(cfg, "base_settings", "test_mode"))
    p.dt = float(_get(cfg, "base_settings", "dt"))
    p.t_max = float(_get(cfg, "base_settings", "t_max"))
    p.track_width = float(_get(cfg, "boat_env", "track_width"))
    p.oob_offset = float(_get(cfg, "boat_env", "boat_out_of_bounds_offset"))
    p.goal_line = float(_get(cfg, "boat_env", "goal_line"))
    p.fuel = float(_get(cfg, "boat", "fuel"))
    for k in ("boat_m", "boat_m_x", "boat_m_y", "boat_I", "boat_Iz", "propeller_diameter",
              "wake_friction", "c_r_front", "c_r_side", "thrust_deduction", "rho",
              "boat_area_front", "boat_area_side", "boat_l", "boat_b", "rudder_area"):
        setattr(p, k, float(_get(cfg, "boat", k)))
    p.fixed_points = int(_get(cfg, "wind", "fixed_points"))
    p.max_velocity = float(_get(cfg, "wind", "max_velocity"))
    p.direction = float(_get(cfg, "wind", "direction"))
    return p
