"""Import the UNMODIFIED reference (Nilau1998/SAC-Agent) under stub modules.

TEST INFRASTRUCTURE ONLY.  This file is used (a) by ``oracle/make_golden.py`` to
generate the committed fixtures under ``tests/golden/`` and (b) by the ``not gpu``
tests that pin the C restatement (``oracle/boat_oracle.c``) against the real
reference while ``/root/reference`` is mounted (i.e. in the build container; the
GPU box has no ``/root/reference`` and nothing run there imports this module).

The reference needs ``gym``, ``dotmap``, ``matplotlib`` and ``seaborn`` which are
not installed; they are only touched for the ``Env``/``Box`` base classes, the
config container and plotting.  Recipe: SURVEY.md section 8(c).

Reference call sites exercised through this shim:
  environment/boat_env.py:10-140   BoatEnv
  environment/wind.py:12-99        Wind
  agent/buffer.py:3-35             ReplayBuffer
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(_HERE, "_ref")   # oracle/make_ref.py: byte-for-byte copies, git-ignored, travel to the GPU box


def _pick_root() -> str:
    env = os.environ.get("SAC_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile(os.path.join("/root/reference", "environment", "boat_env.py")):
        return "/root/reference"
    return STAGED_ROOT


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "environment", "boat_env.py"))


class AttrDict(dict):
    """4-line stand-in for DotMap (utils/config_reader.py:6-14)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e
        return AttrDict(v) if isinstance(v, dict) else v

    def __setattr__(self, k, v):
        self[k] = v


def _install_stubs() -> None:
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env:  # gym.Env is only a base class for BoatEnv
            pass

        gym.Env = Env
        spaces = types.ModuleType("gym.spaces")

        class Box:
            def __init__(self, low, high, shape=None, dtype=np.float32):
                low = np.asarray(low, dtype=dtype)
                high = np.asarray(high, dtype=dtype)
                if low.ndim == 0:  # gym 0.26.2: scalar bounds, no shape -> (1,)
                    low = low.reshape(1)
                    high = high.reshape(1)
                self.low, self.high, self.shape, self.dtype = low, high, low.shape, dtype

        spaces.Box = Box
        gym.spaces = spaces
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    if "dotmap" not in sys.modules:  # utils/config_reader.py:3
        dm = types.ModuleType("dotmap")
        dm.DotMap = AttrDict
        sys.modules["dotmap"] = dm
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]


def load_config(path: str | None = None, **overrides) -> AttrDict:
    """YAML -> attribute dict; ``overrides`` use ``section__key=value``."""
    path = path or os.path.join(REFERENCE_ROOT, "configs", "original_config.yaml")
    with open(path) as f:
        cfg = yaml.safe_load(f)
    for k, v in overrides.items():
        sec, key = k.split("__", 1)
        cfg[sec][key] = v
    return AttrDict(cfg)


_IMPORTED = {}


def import_reference():
    """Returns a namespace with BoatEnv, Boat, Wind, Integrator, ReplayBuffer."""
    if _IMPORTED:
        return types.SimpleNamespace(**_IMPORTED)
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from environment.boat_env import BoatEnv, Boat  # noqa
    from environment.wind import Wind  # noqa
    from environment.control_theory.control_blocks import Integrator  # noqa
    from agent.buffer import ReplayBuffer  # noqa

    _IMPORTED.update(BoatEnv=BoatEnv, Boat=Boat, Wind=Wind, Integrator=Integrator,
                     ReplayBuffer=ReplayBuffer)
    return types.SimpleNamespace(**_IMPORTED)


def fixture_dir(n: int) -> str:
    return os.path.join(REFERENCE_ROOT, "ressources", "settings_visualized",
                        f"experiment_setting_{n}")


def plot_free_experiment_dir() -> str:
    """A directory that already holds a file named reward_field.png (reward_functions.py:23-26)."""
    d = fixture_dir(1)
    if os.path.isfile(os.path.join(d, "reward_field.png")):
        return d
    return os.path.join(STAGED_ROOT, "_experiment_dir")


def make_env(cfg: AttrDict):
    """BoatEnv with an experiment_dir that already holds reward_field.png so the
    per-step os.path.exists (reward_functions.py:26) short-circuits the plot."""
    ref = import_reference()
    exp = types.SimpleNamespace(experiment_dir=plot_free_experiment_dir())
    return ref.BoatEnv(cfg, exp)


class KnotInjector:
    """Context manager: make the reference's ``np.random.randint`` /
    ``np.random.sample`` (boat_env.py:147, wind.py:78) return prescribed values so
    that the reference and the CUDA path see identical episode randomness."""

    def __init__(self, s_y_start: int, curves):
        self.s_y_start = int(s_y_start)
        self.curves = [np.asarray(c, dtype=np.float64) for c in curves]
        self._i = 0

    def __enter__(self):
        self._randint, self._sample = np.random.randint, np.random.sample

        def randint(lo, hi=None, *a, **k):
            return self.s_y_start

        def sample(n):
            c = self.curves[self._i]
            self._i += 1
            assert len(c) == n
            return c.copy()

        np.random.randint, np.random.sample = randint, sample
        return self

    def __exit__(self, *exc):
        np.random.randint, np.random.sample = self._randint, self._sample
        return False


def import_reference_agent(replay_buffer_cls=None):
    """The reference's ``ContinuousAgent`` class (agent/continuous_agent.py:9).  With ``replay_buffer_cls`` the
    module is imported with ``agent.buffer.ReplayBuffer`` replaced -- the import swap of INTEGRATION.md section 1
    (agent/continuous_agent.py:4 ``from agent.buffer import ReplayBuffer``) without touching the file."""
    import importlib
    import_reference()
    saved_buf = sys.modules.get("agent.buffer")
    saved_agent = sys.modules.pop("agent.continuous_agent", None)
    try:
        if replay_buffer_cls is not None:
            m = types.ModuleType("agent.buffer")
            m.ReplayBuffer = replay_buffer_cls
            sys.modules["agent.buffer"] = m
        mod = importlib.import_module("agent.continuous_agent")
        return mod.ContinuousAgent
    finally:
        sys.modules.pop("agent.continuous_agent", None)
        if saved_agent is not None:
            sys.modules["agent.continuous_agent"] = saved_agent
        if saved_buf is not None:
            sys.modules["agent.buffer"] = saved_buf


def import_reference_recorder():
    """postprocessing/recorder.py:7 ``Recorder``."""
    import_reference()
    from postprocessing.recorder import Recorder  # noqa
    return Recorder
