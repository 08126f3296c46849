"""Recorder-compatible export (postprocessing/recorder.py:18-56 of the reference) for selected envs
of a ``BatchedBoatEnv``: the same ``;``-separated CSVs -- per-step rows with the nine columns of
``BoatEnv.return_all_data`` (boat_env.py:128-140), one ``info.csv`` row per finished episode, the
wind table once -- so the reference's ``Replayer`` / renderer can consume GPU runs unchanged.

The single-env drop-in ``sac_agent_b200.BoatEnv`` needs none of this: the reference's own
``Recorder(env)`` works on it as is (it only touches ``experiment_dir``, ``return_all_data()``,
``info`` and ``boat.wind``).
"""
from __future__ import annotations

import csv
import os

import numpy as np

DATA_COLUMNS = ("boat_position_x", "boat_position_y", "boat_velocity_x", "boat_velocity_y", "boat_angle",
                "action_rudder", "reward", "rudder_angle", "n")
INFO_COLUMNS = ("termination", "reached_goal", "out_of_bounds", "out_of_fuel", "rudder_broken", "timeout",
                "episode_reward")
_FIELDS = ("s_x", "s_y", "v_x", "v_y", "s_r", "rudder_angle")
TERM_NAMES = ("", "reached_goal", "out_of_bounds", "out_of_fuel", "timeout", "rudder_broken")


class BatchedRecorder:
    """Call ``write_data_to_csv()`` BEFORE every ``env.step`` (like main.py:79-81: row k is the state
    after k steps, the terminal step is never written) and ``after_step(actions)`` after it."""

    def __init__(self, env, env_ids, experiment_dir):
        self.env = env
        self.ids = [int(i) for i in env_ids]
        self.dir = os.path.join(experiment_dir, "episodes")
        os.makedirs(self.dir, exist_ok=True)
        self.episode = {i: 0 for i in self.ids}
        self.info = {i: dict.fromkeys(INFO_COLUMNS[1:], 0) | {"termination": ""} for i in self.ids}
        self.last_action = {i: 0 for i in self.ids}
        self.last_reward = {i: 0 for i in self.ids}
        self._wind_written = set()

    def _name(self, i, what):
        prefix = "" if self.ids == [0] else f"env{i}_"
        return os.path.join(self.dir, prefix + what)

    def _append(self, path, header, row):
        new = not os.path.exists(path)
        with open(path, "a", newline="") as f:
            w = csv.writer(f, delimiter=";")
            if new:
                w.writerow(header)
            w.writerow(row)

    def write_data_to_csv(self):
        cols = {f: self.env.get_field(f).double().cpu().numpy() for f in _FIELDS}
        for i in self.ids:
            row = [cols["s_x"][i], cols["s_y"][i], cols["v_x"][i], cols["v_y"][i], cols["s_r"][i],
                   self.last_action[i], self.last_reward[i], cols["rudder_angle"][i], 20]
            self._append(self._name(i, f"episode_{self.episode[i]}_data.csv"), DATA_COLUMNS, row)

    def after_step(self, actions):
        """Book-keeping of one ``env.step(actions)``: last action / reward columns, and on termination the
        ``info.csv`` row of the finished episode (cumulative counters, boat_env.py:24-32)."""
        a = np.asarray(actions.detach().cpu() if hasattr(actions, "detach") else actions, dtype=np.float64).reshape(-1)
        r = self.env.reward.double().cpu().numpy()
        term = self.env.term.cpu().numpy()
        ret = None
        for i in self.ids:
            self.last_action[i], self.last_reward[i] = a[i], r[i]
            self.info[i]["episode_reward"] += r[i]
            if term[i]:
                name = TERM_NAMES[int(term[i])]
                self.info[i]["termination"] = name
                self.info[i][name] += 1
                self._append(self._name(i, "info.csv"), INFO_COLUMNS, [self.info[i][k] for k in INFO_COLUMNS])
                self.info[i]["episode_reward"] = 0
                self.episode[i] += 1
                self.last_action[i] = self.last_reward[i] = 0
        return ret

    def write_winds_to_csv(self):
        """wind.csv of each recorded env's CURRENT episode (recorder.py:45-56), written once per file."""
        for i in self.ids:
            path = self._name(i, "wind.csv")
            if path in self._wind_written or os.path.exists(path):
                continue
            wv, wa = self.env.wind_table(i)
            with open(path, "w", newline="") as f:
                w = csv.writer(f, delimiter=";")
                w.writerow(["wind_velocity", "wind_angle"])
                w.writerows(np.column_stack((wv, wa)))
            self._wind_written.add(path)
