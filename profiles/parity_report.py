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


def big_case(seed, precision, n, T, chunk=4096, scale=1.0, experiment=6, episodes=48):
    """fp32 / fp64 auto-reset regime at scale: n envs in chunks (shards of one population: env_id_offset), T steps,
    everything recorded on the device and compared once per chunk.  Counts EPISODES whose end differs."""
    cfg = S.load_config(base_settings__experiment=experiment)
    p = O.params_from_config(cfg)
    tot_eps = mism_eps = 0
    worst_obs = worst_rew = 0.0
    kinds = {}
    for off in range(0, n, chunk):
        m = min(chunk, n - off)
        env = S.BatchedBoatEnv(cfg, m, seed=seed, precision=precision, device=0, auto_reset=True, env_id_offset=off)
        s_y, knots = env.episode_draws_batch(np.arange(m), episodes)
        env.reset()
        acts = torch.stack([env.uniform_actions(t, scale).clone() for t in range(T)])
        obs = torch.empty((T, m, 11), dtype=env.dtype, device=env.device)
        rew = torch.empty((T, m), dtype=env.dtype, device=env.device)
        term = torch.empty((T, m), dtype=torch.uint8, device=env.device)
        for t in range(T):
            o, r, d, info = env.step(acts[t])
            obs[t], rew[t], term[t] = o, r, info["term"]
            obs[t] = torch.where((d > 0)[:, None], info["final_obs"], o)
        ref = O.rollout(p, acts.double().cpu().numpy(), s_y, knots, auto_reset=True)
        tm = term.cpu().numpy()
        bad = tm != ref["term"]
        # an env is comparable up to its first mismatch (afterwards the two are in different episodes)
        first_bad = np.where(bad.any(axis=0), bad.argmax(axis=0), T)
        live = np.arange(T)[:, None] < first_bad[None, :]
        tot_eps += int((ref["term"] > 0).sum())
        mism_eps += int(bad.any(axis=0).sum())
        for i in np.nonzero(bad.any(axis=0))[0]:
            t0 = int(first_bad[i])
            k = f"ref={int(ref['term'][t0, i])},ours={int(tm[t0, i])}"
            kinds[k] = kinds.get(k, 0) + 1
        o = obs.double().cpu().numpy()
        e_obs = np.abs(o - ref["obs"]) / np.maximum(np.abs(ref["obs"]), 1.0)
        e_rew = np.abs(rew.double().cpu().numpy() - ref["reward"]) / np.maximum(np.abs(ref["reward"]), 1.0)
        worst_obs = max(worst_obs, float(e_obs[live].max()))
        worst_rew = max(worst_rew, float(e_rew[live].max()))
        env.close()
    print(json.dumps({"case": "big_auto_reset", "experiment": experiment, "precision": precision, "seed": seed,
                      "n_envs": n, "steps": T, "action_scale": scale, "episodes_finished": tot_eps,
                      "envs_with_a_termination_mismatch": mism_eps, "mismatch_kinds": kinds,
                      "max_scaled_obs_err_before_first_mismatch": worst_obs,
                      "max_scaled_reward_err_before_first_mismatch": worst_rew,
                      "tolerance": 1e-9 if precision == "fp64" else 1e-4}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "big":   # python profiles/parity_report.py big [n_envs] [steps] [seeds...]
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
        T = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
        seeds = [int(x) for x in sys.argv[4:]] or [1, 2, 3, 4, 5]
        for seed in seeds:
            big_case(seed, "fp32", n, T)
        big_case(seeds[0], "fp64", min(n, 8192), T)
        # goal / out-of-bounds regime: small steering noise, experiment 2 (random start), long episodes
        for seed in seeds[:2]:
            big_case(seed, "fp32", min(n, 8192), 6000, scale=0.03, experiment=2, episodes=8)
        sys.exit(0)
    for exp in range(1, 7):
        for precision in ("fp64", "fp32"):
            case(exp, precision, 4096, 1000, 0.05, False, 1)
    for precision in ("fp64", "fp32"):
        case(6, precision, 2048, 1500, 1.0, True, 64)
