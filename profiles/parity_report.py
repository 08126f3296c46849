#!/usr/bin/env python
"""Measured parity margins (not pass/fail -- the tests do that): for every experiment, 4096 envs x 1000
steps of small steering noise against the CPU oracle, fp64 and fp32; plus the auto-reset regime.
Metric: |a-b| / max(|b|, 1) on normalised observations and rewards.  One JSON line each."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import sac_agent_b200 as S  # noqa: E402
from oracle import oracle as O  # noqa: E402


def draws(env, episodes):
    n, fp = env.n_envs, int(env.params.fixed_points)
    s_y = np.empty((episodes, n), dtype=np.int32)
    knots = np.empty((episodes, n, 2, fp))
    for e in range(episodes):
        for i in range(n):
            s_y[e, i], knots[e, i] = env.episode_draws(i, e)
    return s_y, knots


def case(experiment, precision, n, T, scale, auto_reset, episodes):
    cfg = S.load_config(base_settings__experiment=experiment)
    env = S.BatchedBoatEnv(cfg, n, seed=1, precision=precision, device=0, auto_reset=auto_reset)
    s_y, knots = draws(env, episodes)
    env.reset()
    acts = torch.stack([env.uniform_actions(t, scale).clone() for t in range(T)])
    actions = acts.double().cpu().numpy()
    ref = O.rollout(O.params_from_config(cfg), actions, s_y, knots, auto_reset=auto_reset)
    worst_obs = worst_rew = 0.0
    mism = 0
    for t in range(T):
        obs, rew, done, info = env.step(acts[t])
        o = obs.double().cpu().numpy()
        d = ref["done"][t].astype(bool)
        if auto_reset and d.any():
            o[d] = info["final_obs"].double().cpu().numpy()[d]
        worst_obs = max(worst_obs, float((np.abs(o - ref["obs"][t]) / np.maximum(np.abs(ref["obs"][t]), 1.0)).max()))
        r = rew.double().cpu().numpy()
        worst_rew = max(worst_rew, float((np.abs(r - ref["reward"][t]) / np.maximum(np.abs(ref["reward"][t]), 1.0)).max()))
        mism += int((done.cpu().numpy() != ref["done"][t]).sum() + (info["term"].cpu().numpy() != ref["term"][t]).sum())
    print(json.dumps({"experiment": experiment, "precision": precision, "n_envs": n, "steps": T, "action_scale": scale,
                      "auto_reset": auto_reset, "episodes_finished": int(ref["done"].sum()),
                      "max_scaled_obs_err": worst_obs, "max_scaled_reward_err": worst_rew,
                      "done_or_term_mismatches": mism,
                      "tolerance": 1e-9 if precision == "fp64" else 1e-4}), flush=True)
    env.close()


if __name__ == "__main__":
    for exp in range(1, 7):
        for precision in ("fp64", "fp32"):
            case(exp, precision, 4096, 1000, 0.05, False, 1)
    for precision in ("fp64", "fp32"):
        case(6, precision, 2048, 1500, 1.0, True, 64)
