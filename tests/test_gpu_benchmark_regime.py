"""CUDA-vs-oracle parity IN THE BENCHMARK REGIME (BASELINE.json configs[2]): millions of envs per launch, so that
every persistent warp walks hundreds of state blocks (stage refill, action prefetch, alternating sweep direction),
with the uniform(-1,1) policy and auto-reset in steady state (>= 400 steps).  The other parity tests stay below
75,776 envs = one block per persistent warp; the stage-refill race fixed in round 1 passed all of them.

A subset of >= 4096 env ids -- the first blocks, blocks just past one grid sweep, the middle, the last full blocks,
the ragged tail, and random ids -- is compared with the CPU oracle (boat_env.py:67-115 restated in
oracle/boat_oracle.c) step by step: done / termination code equal, observations and rewards within the north-star
tolerance, statistics of the subset equal.  Reference semantics: /root/reference/environment/boat_env.py:67-126.

These tests FAIL when the refill vote of boat_step.cuh is compiled out (-DBOAT_DEBUG_NO_REFILL_VOTE, see
profiles/r02_refill_vote_ablation.txt).
"""
import numpy as np
import pytest

from boat_testlib import scaled_err

pytestmark = pytest.mark.gpu

TOL = {"fp64": 1e-9, "fp32": 1e-4}
N_FULL = 16 * 1024 * 1024 + 19      # the benchmark's population plus a ragged last block of 19 envs


@pytest.fixture(scope="module")
def S():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sac_agent_b200 as pkg
    pkg.lib()
    return pkg


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def sample_ids(n, m_random, seed):
    """Env ids spread over the launch: first 8 blocks, the blocks a persistent warp reaches on its 2nd / 3rd / 100th
    trip (grid = 296 CTAs x 8 warps -> 2368 blocks per sweep), the middle, the last full blocks, the ragged tail."""
    blocks = [0, 1, 2, 3, 4, 5, 6, 7]
    nblk = (n + 31) // 32
    for trip in (1, 2, 3, 100, 150, 220):
        for w in (0, 1, 7, 8, 2367):
            b = trip * 2368 + w
            if b < nblk:
                blocks.append(b)
    mid = nblk // 2
    blocks += list(range(mid - 4, mid + 4)) + list(range(nblk - 9, nblk))
    ids = []
    for b in sorted(set(blocks)):
        ids += [i for i in range(b * 32, b * 32 + 32) if i < n]
    rng = np.random.default_rng(seed)
    ids += list(rng.integers(0, n, size=m_random))
    return np.unique(np.asarray(ids, dtype=np.int64))


def subset_draws(env, ids, episodes):
    s_y, knots = env.episode_draws_batch(ids, episodes)
    s1, k1 = env.episode_draws(int(ids[-1]), episodes - 1)   # the batch call == the per-pair call
    assert s_y[-1, -1] == s1 and np.array_equal(knots[-1, -1], k1)
    return s_y, knots


def np_(t):
    return t.detach().double().cpu().numpy() if t.is_floating_point() else t.detach().cpu().numpy()


def compare_with_oracle(S, O, cfg, env, ids, T, E, precision, advance, scale=1.0):
    """Steps `env` T times through `advance(env, actions)`, records the subset, compares with O.rollout."""
    import torch
    idx = torch.from_numpy(ids).to(env.device)
    M = len(ids)
    s_y, knots = subset_draws(env, ids, E)
    obs0 = np_(env.reset()[idx])
    acts = torch.empty((T, M), dtype=env.dtype, device=env.device)
    obs = torch.empty((T, M, 11), dtype=env.dtype, device=env.device)
    fin = torch.empty((T, M, 11), dtype=env.dtype, device=env.device)
    rew = torch.empty((T, M), dtype=env.dtype, device=env.device)
    done = torch.empty((T, M), dtype=torch.uint8, device=env.device)
    term = torch.empty((T, M), dtype=torch.uint8, device=env.device)
    total_done = torch.zeros((), dtype=torch.int64, device=env.device)
    for t in range(T):
        a = env.uniform_actions(t, scale)
        acts[t] = a[idx]
        o, r, d, info = advance(env, a)
        obs[t], rew[t], done[t], term[t] = o[idx], r[idx], d[idx], info["term"][idx]
        if "final_obs" in info:
            fin[t] = info["final_obs"][idx]
        total_done += d.sum()
    actions = np_(acts)
    ref = O.rollout(O.params_from_config(cfg), actions, s_y, knots, auto_reset=True)
    out = dict(obs=np_(obs), reward=np_(rew), done=np_(done), term=np_(term), final_obs=np_(fin))
    tol = TOL[precision]
    ref0 = np.zeros((M, 11)); ref0[:, 3] = 0.5; ref0[:, 9] = 0.5; ref0[:, 10] = 1.0   # boat_env.py:124 (exp 6: s_y0 = 0)
    assert np.abs(obs0 - ref0).max() <= (1e-15 if precision == "fp64" else 1e-6)
    assert np.array_equal(out["done"], ref["done"]), \
        f"{int((out['done'] != ref['done']).sum())} done mismatches in {int(ref['done'].sum())} episodes"
    assert np.array_equal(out["term"], ref["term"])
    d = ref["done"].astype(bool)
    assert d.sum() > M, "auto-reset must be in steady state (more than one episode end per sampled env)"
    assert scaled_err(out["obs"][~d], ref["obs"][~d]).max() <= tol
    assert scaled_err(out["reward"], ref["reward"]).max() <= tol
    # statistics of the subset (info dict, boat_env.py:24-32): per-kind counts equal the oracle's
    for code in range(1, 6):
        assert int((out["term"] == code).sum()) == int((ref["term"] == code).sum())
    c = env.counters()
    assert c["episodes"] == float(total_done.item())
    return out, ref, d


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_full_population_sampled_parity(S, O, precision):
    """boatenv_step, 16,777,235 envs, experiment 6, 450 steps, both sweep directions (launches alternate)."""
    cfg = S.load_config(base_settings__experiment=6)
    ids = sample_ids(N_FULL, 3000, seed=5)
    assert len(ids) >= 4096
    env = S.BatchedBoatEnv(cfg, N_FULL, seed=1, precision=precision, device=0, auto_reset=True)
    out, ref, d = compare_with_oracle(S, O, cfg, env, ids, 450, 24, precision,
                                      lambda e, a: e.step(a))
    # terminal observations go to final_obs; the obs row of a finished env is the next episode's reset observation
    assert scaled_err(out["final_obs"][d], ref["obs"][d]).max() <= TOL[precision]
    rows = out["obs"][d]
    assert np.all(rows[:, [0, 1, 2, 4, 5, 6, 7, 8]] == 0) and np.all(rows[:, [3, 9]] == 0.5) and np.all(rows[:, 10] == 1)
    env.close()


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("n", [300_003, 1_000_003])
def test_l2_resident_population_sampled_parity(S, O, precision, n):
    """The same at sizes whose state stays in the 126 MB L2 (the refill of a stage then lands within a few hundred
    cycles: the regime in which the round-1 stage-refill race showed), ~4 / ~13 blocks per persistent warp, big
    actions (an episode every ~25 steps: resets in every launch)."""
    cfg = S.load_config(base_settings__experiment=6)
    ids = sample_ids(n, 3500, seed=8)
    env = S.BatchedBoatEnv(cfg, n, seed=2, precision=precision, device=0, auto_reset=True)
    out, ref, d = compare_with_oracle(S, O, cfg, env, ids, 200, 96, precision, lambda e, a: e.step(a), scale=4.0)
    assert scaled_err(out["final_obs"][d], ref["obs"][d]).max() <= TOL[precision]
    env.close()


@pytest.mark.parametrize("experiment", [1, 3, 4, 5])
def test_other_experiments_sampled_parity_at_scale(S, O, experiment):
    """The other kernel instantiations (no wind, constant wind, one random curve as velocity / as rectified direction)
    at 4 M envs, fp32, the same sampled comparison: experiments 4 and 5 use the warp-specialised setup path with
    one curve, 1 and 3 the inline path without curves."""
    cfg = S.load_config(base_settings__experiment=experiment)
    n = 4 * 1024 * 1024 + 11
    ids = sample_ids(n, 2500, seed=10 + experiment)
    env = S.BatchedBoatEnv(cfg, n, seed=4, precision="fp32", device=0, auto_reset=True)
    out, ref, d = compare_with_oracle(S, O, cfg, env, ids, 420, 24, "fp32", lambda e, a: e.step(a))
    assert scaled_err(out["final_obs"][d], ref["obs"][d]).max() <= TOL["fp32"]
    env.close()


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_full_population_fused_store_sampled_parity(S, O, precision):
    """boatenv_step_store (step + agent.remember in one kernel) at 4 M envs over a 6 M-slot ring (wraps twice):
    the step outputs against the oracle, and the ring rows of the last launch against the recorded transition."""
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    n = 4 * 1024 * 1024 + 7
    ids = sample_ids(n, 3000, seed=6)
    env = S.BatchedBoatEnv(cfg, n, seed=2, precision=precision, device=0, auto_reset=True)
    ring = S.ReplayBuffer(6_000_001, (11,), 1, precision=precision, device=0, as_torch=True)
    prev = {}

    def advance(e, a):
        prev["obs"] = e.obs.clone()
        prev["cntr"] = ring.mem_cntr
        return ring.step_store(e, a, done_flag_mode=0)

    T = 420
    out, ref, d = compare_with_oracle(S, O, cfg, env, ids, T, 24, precision, advance)
    # ring rows written by the LAST launch: s = the previous observations, s' = the step's observation -- the
    # TERMINAL one for finished envs (main.py:81-88 stores observation_ before the reset), a, r, done
    idx = torch.from_numpy(ids).to(env.device)
    slots = (prev["cntr"] + idx) % ring.mem_size
    s, a, r, s2, dn = ring.gather(slots)
    tol = TOL[precision]
    assert torch.equal(s, prev["obs"][idx])
    last_ref = ref["obs"][T - 1]
    assert scaled_err(np_(s2), last_ref).max() <= tol
    assert scaled_err(np_(r), ref["reward"][T - 1]).max() <= tol
    assert np.array_equal(np_(dn).astype(np.uint8), ref["done"][T - 1])
    assert ring.mem_cntr == T * n
    env.close(); ring.close()


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_full_population_step_k_sampled_parity(S, O, precision):
    """boatenv_step_k (K = 8 fused sub-steps, per-sub-step actions, deferred episode-end queue) at 4 M envs: an env
    stops at its first done inside the window; the next window starts its next episode."""
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    n, K, W, E = 4 * 1024 * 1024 + 7, 8, 56, 24
    ids = sample_ids(n, 0, seed=7)[::3]          # ~400 envs: the per-env window emulation runs in Python
    ids = np.unique(np.concatenate([ids, np.arange(n - 7, n)]))
    idx = torch.from_numpy(ids).to(0)
    env = S.BatchedBoatEnv(cfg, n, seed=3, precision=precision, device=0, auto_reset=True)
    s_y, knots = subset_draws(env, ids, E)
    env.reset()
    p = O.params_from_config(cfg)
    oracles = [O.OracleEnv(p) for _ in ids]
    episode = [0] * len(ids)
    for j, o in enumerate(oracles):
        o.reset(int(s_y[0, j]), knots[0, j, 0], knots[0, j, 1])
    tol = TOL[precision]
    n_done = 0
    for w in range(W):
        acts = torch.stack([env.uniform_actions(w * K + k, 1.0).clone() for k in range(K)])
        obs, rew, done, info = env.step_k(acts, K)
        a_np, obs_np, rew_np = np_(acts[:, idx]), np_(obs[idx]), np_(rew[idx])
        done_np, term_np, steps_np = np_(done[idx]), np_(info["term"][idx]), np_(info["steps"][idx])
        for j, o in enumerate(oracles):
            r_sum, dd, code, steps, last = 0.0, False, 0, 0, None
            for k in range(K):
                last, r, dd, code = o.step(float(a_np[k, j]))
                r_sum += r
                steps += 1
                if dd:
                    break
            assert bool(done_np[j]) == dd and int(term_np[j]) == code and int(steps_np[j]) == steps
            assert abs(rew_np[j] - r_sum) <= tol * max(1.0, abs(r_sum))
            if dd:
                n_done += 1
                episode[j] += 1
                last = o.reset(int(s_y[episode[j], j]), knots[episode[j], j, 0], knots[episode[j], j, 1])
            assert scaled_err(obs_np[j], last).max() <= tol
    assert n_done > len(ids) // 2
    env.close()
