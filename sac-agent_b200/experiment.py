"""Experiment-directory compatibility: the directory tree, the tuned-config draw
and the three summary CSVs the reference's training driver leaves behind, so that its post-processing
(`Replayer`, `BoatEnvironmentRenderer`, the `-a` average plots) and `get_experiment_config` read a
run of the batched env unchanged.

  utils/build_experiment.py:9-41      Experiment: experiments/[subdir/]experiment_<timestamp>/{plots,checkpoints,
                                      configs,rendering,episodes}, save_configs()
  utils/hyperparameter_tuner.py:9-52  HPTuner: alpha / beta / gamma / tau uniform in hp_configs.yaml's ranges,
                                      rounded to 4 digits, written as configs/tuned_configs.yaml
  utils/config_reader.py:11-14        get_experiment_config
  main.py:116-133                     console.csv, <experiments_dir>/overview.csv (appended), terminations.csv

Host-side file I/O only; nothing here touches the GPU.
"""
from __future__ import annotations

import copy
import csv
import os
import random
import time

import yaml

from .config import DEFAULTS, AttrDict

# configs/hp_configs.yaml of the reference
HP_RANGES = {"alpha": (0.001, 0.01), "beta": (0.0008, 0.008), "gamma": (0.95, 0.99), "tau": (0.001, 0.01)}
CONSOLE_COLUMNS = ["CCID Episode", "Termination", "Score", "Best Score", "Average Score", "RA", "Action RA"]  # main.py:52-53
INFO_KEYS = ["termination", "reached_goal", "out_of_bounds", "out_of_fuel", "rudder_broken", "timeout",
             "episode_reward"]  # boat_env.py:24-32, the header of terminations.csv


def _plain(cfg):
    return {k: (_plain(v) if isinstance(v, dict) else v) for k, v in dict(cfg).items()}


def get_experiment_config(experiment_dir, config_file="tuned_configs.yaml") -> AttrDict:
    """utils/config_reader.py:11-14."""
    with open(os.path.join(experiment_dir, "configs", config_file)) as f:
        return AttrDict(yaml.safe_load(f))


class HPTuner:
    """utils/hyperparameter_tuner.py:9-52.  `rng`: a `random.Random` (the reference uses the module-level one)."""

    def __init__(self, ranges=None, rng=None):
        r, g = dict(HP_RANGES if ranges is None else ranges), (rng or random)
        self.hpset = {k: round(g.uniform(*r[k]), 4) for k in ("alpha", "beta", "gamma", "tau")}

    def tuned(self, config) -> dict:
        out = copy.deepcopy(_plain(config))
        a = out.setdefault("agent", {})
        a["learning_rate_alpha"], a["learning_rate_beta"] = self.hpset["alpha"], self.hpset["beta"]
        a["gamma"], a["tvn_parameter_modulation_tau"] = self.hpset["gamma"], self.hpset["tau"]
        return out


class Experiment:
    """utils/build_experiment.py:9-41.  `root` replaces the reference's hard-wired 'experiments'."""

    SUBDIRS = ("plots", "checkpoints", "configs", "rendering", "episodes")

    def __init__(self, experiment_name="experiment", subdir=None, root="experiments", rng=None):
        self.timestamp = time.strftime("_%m-%d_%H-%M-%S%f")[:-2]
        self.experiments_dir = os.path.join(root, subdir) if subdir is not None else root
        self.tuner = HPTuner(rng=rng)
        os.makedirs(self.experiments_dir, exist_ok=True)
        self.experiment_name = experiment_name + self.timestamp
        self.experiment_dir = os.path.join(self.experiments_dir, self.experiment_name)
        for d in self.SUBDIRS:
            os.makedirs(os.path.join(self.experiment_dir, d))

    def save_configs(self, config=None):
        """configs/original_config.yaml + the tuned draw as configs/tuned_configs.yaml (:37-41)."""
        cfg = _plain(DEFAULTS if config is None else config)
        for sec, keys in DEFAULTS.items():   # a partial config still leaves a complete reference-style YAML behind
            have = cfg.setdefault(sec, {})
            for k, v in keys.items():
                have.setdefault(k, v)
        cdir = os.path.join(self.experiment_dir, "configs")
        with open(os.path.join(cdir, "original_config.yaml"), "w") as f:
            yaml.dump(cfg, f, default_flow_style=False)
        with open(os.path.join(cdir, "hp_configs.yaml"), "w") as f:
            yaml.dump({"agent": {f"{k}_{e}": v for k, (lo, hi) in HP_RANGES.items() for e, v in (("min", lo), ("max", hi))}},
                      f, default_flow_style=False)
        with open(os.path.join(cdir, "tuned_configs.yaml"), "w") as f:
            yaml.dump(self.tuner.tuned(cfg), f, default_flow_style=False)
        return get_experiment_config(self.experiment_dir, "tuned_configs.yaml")

    # -- main.py:116-133 ----------------------------------------------------------------
    def write_console(self, rows):
        """console.csv: header + one row per episode (or per logging interval of a batched run)."""
        with open(os.path.join(self.experiment_dir, "console.csv"), "x", newline="") as f:
            w = csv.writer(f, delimiter=";")
            w.writerow(CONSOLE_COLUMNS)
            w.writerows(rows)

    def append_overview(self, best_score):
        """<experiments_dir>/overview.csv: one `name;best_score` line per experiment, appended."""
        with open(os.path.join(self.experiments_dir, "overview.csv"), "a", newline="") as f:
            csv.writer(f, delimiter=";").writerow([self.experiment_name, best_score])

    def write_terminations(self, info):
        """terminations.csv: the keys and values of the env's cumulative info dict."""
        with open(os.path.join(self.experiment_dir, "terminations.csv"), "x", newline="") as f:
            w = csv.writer(f, delimiter=";")
            w.writerow(list(info.keys()))
            w.writerow(list(info.values()))


def info_from_counters(counters: dict, last_termination="") -> dict:
    """The reference's cumulative `info` dict (boat_env.py:24-32) from a BatchedBoatEnv's counters: the
    per-kind termination counts of ALL envs, `episode_reward` = their mean episode return."""
    n = counters.get("episodes", 0.0)
    out = {"termination": last_termination}
    for k in INFO_KEYS[1:-1]:
        out[k] = int(counters.get(k, 0))
    out["episode_reward"] = counters.get("return_sum", 0.0) / n if n else 0.0
    return out
