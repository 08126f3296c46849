#!/usr/bin/env python
"""Recipe: stage the UNMODIFIED reference (Nilau1998/SAC-Agent) files of the hot path and of its callers under
``oracle/_ref/`` so that the real reference travels to the GPU box with the tree.

TEST INFRASTRUCTURE ONLY.  ``oracle/_ref/`` is git-ignored (no reference source ever enters the history) but not
gpurun-ignored.  It is used by
  * ``bench.py``'s ``cpu_baseline`` leg and ``--impl reference`` arm (``kind: "reference"``): the reference's own
    ``BoatEnv`` stepped on the host cores (oracle/ref_bench.py);
  * ``tests/test_reference_callers.py``: the reference's own ``ContinuousAgent`` / ``Recorder`` / ``main.py``-style
    loop driving ``sac_agent_b200.BoatEnv`` / ``ReplayBuffer`` through the two-import swap of INTEGRATION.md section 1.
Nothing under ``sac-agent_b200/`` reads it.

    python oracle/make_ref.py [--src /root/reference] [--check]

The files are byte-for-byte copies (MANIFEST.json records their sha256); they are imported through the stub
modules of ``oracle/ref_shim.py`` (gym / dotmap / matplotlib / seaborn are not installed in this image).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
DEFAULT_SRC = "/root/reference"

# the seven hot-path files ...
HOT_PATH = [
    "environment/__init__.py", "environment/boat_env.py", "environment/wind.py", "environment/reward_functions.py",
    "environment/toy_car.py", "environment/toy_parachute.py", "environment/control_theory/__init__.py",
    "environment/control_theory/control_blocks.py", "agent/buffer.py",
]
# ... and the callers either side of the path (section 8b): agent, networks, recorder, config reader, the loop
CALLERS = [
    "agent/__init__.py", "agent/base_agent.py", "agent/continuous_agent.py", "agent/discrete_agent.py",
    "networks/__init__.py", "networks/base_network.py", "networks/networks.py",
    "postprocessing/__init__.py", "postprocessing/recorder.py",
    "utils/__init__.py", "utils/config_reader.py",
    "configs/__init__.py", "configs/original_config.yaml", "configs/hp_configs.yaml",
    "main.py",
]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def make_ref(src: str = DEFAULT_SRC, dest: str = DEST) -> str:
    if not os.path.isfile(os.path.join(src, "environment", "boat_env.py")):
        raise FileNotFoundError(f"reference not found at {src}")
    manifest = {"source": src, "files": {}}
    for rel in HOT_PATH + CALLERS:
        s = os.path.join(src, rel)
        if not os.path.isfile(s):
            continue  # optional (an __init__.py the reference does not have)
        d = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest["files"][rel] = _sha(d)
    # reward_functions.py:23-26 plots the reward field unless <experiment_dir>/reward_field.png exists: an empty
    # file of that name (not a reference file) short-circuits the per-step os.path.exists
    plot_dir = os.path.join(dest, "_experiment_dir")
    for sub in ("", "episodes", "checkpoints", "configs", "plots", "rendering"):
        os.makedirs(os.path.join(plot_dir, sub), exist_ok=True)
    open(os.path.join(plot_dir, "reward_field.png"), "ab").close()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return dest


def check(dest: str = DEST) -> bool:
    """True when every staged file still has the recorded hash."""
    try:
        with open(os.path.join(dest, "MANIFEST.json")) as f:
            man = json.load(f)
        return bool(man["files"]) and all(_sha(os.path.join(dest, rel)) == h for rel, h in man["files"].items())
    except Exception:
        return False


if __name__ == "__main__":
    src = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else DEFAULT_SRC
    if "--check" in sys.argv:
        sys.exit(0 if check() else 1)
    print(make_ref(src))
