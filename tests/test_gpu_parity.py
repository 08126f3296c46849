"""Parity of the CUDA path (through the C ABI of libboatenv.so) against the CPU oracle
and the committed golden vectors of the reference.  Needs a B200: every test is
``@pytest.mark.gpu``.

Tolerances (BASELINE.json north_star): termination / step counts bit-exact; states and
rewards within 1e-9 (fp64 validation mode) and 1e-4 (fp32 production mode) over 1000
steps, measured as |a-b| / max(|b|, S_i) on the reference's own normalised observations
(S_i = 1).
"""
import json

import numpy as np
import pytest

from boat_testlib import load_golden, scaled_err

pytestmark = pytest.mark.gpu

TOL64 = 1e-9
TOL32 = 1e-4

ROLLOUTS = [f"ref_rollout_exp{e}" for e in range(1, 7)] + [
    "ref_term_rudder_exp6", "ref_term_oob_exp2", "ref_term_fuel_exp3", "ref_term_timeout_exp4",
    "ref_term_goal_exp5"]


@pytest.fixture(scope="module")
def S():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sac_agent_b200 as pkg
    pkg.lib()  # raises if libboatenv.so is missing: the GPU tests never run on a fallback
    return pkg


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def np_(t):
    return t.detach().double().cpu().numpy() if t.is_floating_point() else t.detach().cpu().numpy()


def make_env(S, cfg, n, precision, s_y=None, knots=None, **kw):
    env = S.BatchedBoatEnv(cfg, n, precision=precision, device=0, **kw)
    if s_y is not None or knots is not None:
        env.set_episode_draws(s_y, knots)
    return env


def run_steps(env, actions):
    """actions [T, N] numpy -> dict of [T, N, ...] numpy arrays."""
    import torch
    T, N = actions.shape
    a_dev = torch.from_numpy(np.ascontiguousarray(actions)).to(env.device).to(env.dtype)
    obs = torch.empty((T, N, 11), dtype=env.dtype, device=env.device)
    rew = torch.empty((T, N), dtype=env.dtype, device=env.device)
    done = torch.empty((T, N), dtype=torch.uint8, device=env.device)
    term = torch.empty((T, N), dtype=torch.uint8, device=env.device)
    fin = torch.zeros((T, N, 11), dtype=env.dtype, device=env.device)
    for t in range(T):
        o, r, d, info = env.step(a_dev[t])
        obs[t], rew[t], done[t], term[t] = o, r, d, info["term"]
        fin[t] = info["final_obs"]
    return dict(obs=np_(obs), reward=np_(rew), done=np_(done), term=np_(term), final_obs=np_(fin))


# ---------------------------------------------------------------------------------------
# 1. golden roll-outs of the unmodified reference (tests/golden/ref_*.npz)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ROLLOUTS)
def test_golden_rollouts_fp64(S, name):
    g = load_golden(name)
    T, N = g["actions"].shape
    env = make_env(S, g["config"], N, "fp64", g["s_y_start"], g["knots"], auto_reset=False)
    obs0 = np_(env.reset())
    assert np.abs(obs0 - g["obs0"]).max() < 1e-15
    out = run_steps(env, g["actions"].astype(np.float64))
    assert np.array_equal(out["done"], g["done"])
    assert np.array_equal(out["term"], g["term"])
    assert scaled_err(out["obs"], g["obs"]).max() <= TOL64
    assert scaled_err(out["reward"], g["reward"]).max() <= TOL64
    assert scaled_err(np_(env.get_field("episode_reward")), g["episode_reward"]).max() <= TOL64
    # return_all_data (boat_env.py:128-140) columns from the state fields
    last = g["all_data"][-1]
    for col, f in enumerate(("s_x", "s_y", "v_x", "v_y", "s_r")):
        scale = (3900.0, 800.0, 5.0, 2.0, 2 * np.pi)[col]
        assert scaled_err(np_(env.get_field(f)), last[:, col], scale).max() <= TOL64
    assert scaled_err(np_(env.get_field("rudder_angle")), last[:, 7], np.pi / 3).max() <= TOL64
    env.close()


@pytest.mark.parametrize("name", ROLLOUTS)
def test_golden_rollouts_fp32(S, name):
    g = load_golden(name)
    T, N = g["actions"].shape
    env = make_env(S, g["config"], N, "fp32", g["s_y_start"], g["knots"], auto_reset=False)
    obs0 = np_(env.reset())
    assert np.abs(obs0 - g["obs0"]).max() < 1e-6
    out = run_steps(env, g["actions"])
    # compare every env up to and including its first done (afterwards the reference
    # object keeps integrating a broken boat: SURVEY.md H5, the unstable regime)
    first = np.where(g["done"].any(axis=0), g["done"].argmax(axis=0), T - 1)
    live = np.arange(T)[:, None] <= first[None, :]
    assert np.array_equal(out["done"][live], g["done"][live])
    assert np.array_equal(out["term"][live], g["term"][live])
    assert scaled_err(out["obs"], g["obs"])[live].max() <= TOL32
    assert scaled_err(out["reward"], g["reward"])[live].max() <= TOL32
    env.close()


# ---------------------------------------------------------------------------------------
# 2. the reference's recorded fixtures (ressources/settings_visualized): known answers
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6])
def test_recorded_fixture_replay(S, n, precision):
    """The reference's six recorded episodes (test_mode 1: the boat runs straight to the goal line in 4959 / 4963
    steps).  fp64: 1e-11; fp32 production mode: the same STEP COUNT and termination (s_x is carried in fixed point,
    so 5000 accumulations do not move the goal crossing) and 1e-4 on the recorded quantities."""
    g = load_golden(f"fixture_exp{n}")
    tol, rtol = (1e-11, 1e-11) if precision == "fp64" else (TOL32, 2e-4)
    fp = int(g["config"]["wind"]["fixed_points"])
    knots = np.zeros((1, 2, fp))
    if g["meta"]["kind_v"] == "curve":
        knots[0, 0] = g["knots_v"]
    if g["meta"]["kind_a"] == "curve":
        knots[0, 1] = g["knots_a"]
    if g["meta"]["kind_a"] == "rect":  # exp 5: the first drawn curve is the rect source (wind.py:57)
        knots[0, 0] = g["knots_r"]
    env = make_env(S, g["config"], 1, precision, np.array([int(g["s_y_start"])]), knots, auto_reset=False)
    env.reset()
    wv, wa = env.wind_table(0)
    assert np.abs(wv[g["wind_idx"]] - g["wind_v"]).max() < 5e-14
    assert np.abs(wa[g["wind_idx"]] - g["wind_a"]).max() < 5e-13
    import torch
    zero = torch.zeros(1, dtype=env.dtype, device=env.device)
    want = {int(k): i for i, k in enumerate(g["row_idx"])}
    ref = g["rows"]
    steps, done, ep_reward = 0, False, 0.0
    offset = 0.0 if n == 6 else 0.1  # fixtures 1-5 were recorded when f_x == 0.1
    while not done and steps < 6000:
        if steps in want:
            row = ref[want[steps]]
            for col, f, scale in ((0, "s_x", 3900.0), (1, "s_y", 800.0), (2, "v_x", 5.0), (3, "v_y", 2.0),
                                  (4, "s_r", 2 * np.pi)):
                assert scaled_err(np_(env.get_field(f))[0], row[col], scale) <= tol
            if steps > 0:
                assert abs(float(rew[0]) + offset - row[6]) <= rtol
        obs, rew, d, info = env.step(zero)
        done = bool(d[0].item())
        ep_reward += float(rew[0])
        steps += 1
    assert steps == int(g["n_rows"])  # the terminal step is never written (main.py:79-81)
    assert int(info["term"][0]) == 1 and str(g["termination"]) == "reached_goal"
    if n == 6:
        assert ep_reward == pytest.approx(871.2727580297085, abs=1e-8 if precision == "fp64" else 0.05)
    env.close()


# ---------------------------------------------------------------------------------------
# 3. CUDA vs oracle on seeded inputs (BASELINE.json configs[1]: exp 3, 4096 envs, fp64)
# ---------------------------------------------------------------------------------------
def host_draws(env, episodes):
    n = env.n_envs
    fp = int(env.params.fixed_points)
    s_y = np.empty((episodes, n), dtype=np.int32)
    knots = np.empty((episodes, n, 2, fp))
    for e in range(episodes):
        for i in range(n):
            s_y[e, i], knots[e, i] = env.episode_draws(i, e)
    return s_y, knots


@pytest.mark.parametrize("experiment,precision,n,T", [(3, "fp64", 4096, 1000), (6, "fp64", 1024, 1000),
                                                      (2, "fp64", 512, 600), (6, "fp32", 4096, 1000),
                                                      (5, "fp32", 1024, 1000), (4, "fp32", 1024, 1000)])
def test_cuda_vs_oracle_seeded(S, O, experiment, precision, n, T):
    cfg = S.load_config(base_settings__experiment=experiment)
    env = make_env(S, cfg, n, precision, seed=1, auto_reset=False)
    s_y, knots = host_draws(env, 1)
    env.reset()
    import torch
    # small steering noise A2 of SURVEY.md 8(d): float32(0.05 * U(-1,1)) from Philox(seed=1)
    acts = torch.stack([env.uniform_actions(t, 0.05).clone() for t in range(T)])
    actions = np_(acts)
    ref = O.rollout(O.params_from_config(cfg), actions, s_y, knots)
    out = run_steps(env, actions)
    tol = TOL64 if precision == "fp64" else TOL32
    assert np.array_equal(out["done"], ref["done"])
    assert np.array_equal(out["term"], ref["term"])
    assert scaled_err(out["obs"], ref["obs"]).max() <= tol
    assert scaled_err(out["reward"], ref["reward"]).max() <= tol
    env.close()


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_synchronised_piece_crossings_flood_the_setup_queue(S, O, precision):
    """Worst case for the warp-specialised K = 1 kernel: t_max = 60 s gives L = 240 wind samples, i.e. a
    new spline piece every ~34 steps, and with tiny actions nobody resets, so EVERY env of every warp asks
    for new wind coefficients in the same launch -- 8 x 32 requests per CTA against a 64-slot ring (the
    producers block until the setup warps have drained it).  Timeouts then reset everybody at once."""
    cfg = S.load_config(base_settings__experiment=6, base_settings__t_max=60)
    n, T, E = 40_000, 300, 3
    env = make_env(S, cfg, n, precision, seed=21, auto_reset=True)
    s_y, knots = host_draws(env, E)
    env.reset()
    import torch
    actions = np_(torch.stack([env.uniform_actions(t, 0.01).clone() for t in range(T)]))
    ref = O.rollout(O.params_from_config(cfg), actions, s_y, knots, auto_reset=True)
    out = run_steps(env, actions)
    assert np.array_equal(out["done"], ref["done"]) and np.array_equal(out["term"], ref["term"])
    assert (ref["term"] == 4).sum() == n  # everybody times out at step 240, together
    tol = TOL64 if precision == "fp64" else TOL32
    d = ref["done"].astype(bool)
    assert scaled_err(out["obs"][~d], ref["obs"][~d]).max() <= tol
    assert scaled_err(out["final_obs"][d], ref["obs"][d]).max() <= tol
    assert scaled_err(out["reward"], ref["reward"]).max() <= tol
    env.close()


@pytest.mark.parametrize("precision,seed", [("fp64", 3), ("fp32", 1), ("fp32", 2), ("fp32", 3), ("fp32", 4), ("fp32", 5)])
def test_auto_reset_matches_oracle(S, O, precision, seed):
    """Uniform(-1,1) policy A1: episodes end by rudder_broken every ~360 steps; the kernel
    resets in place with the next episode's Philox draws.  Step counts, termination kinds
    and statistics must equal the oracle's; terminal observations go to final_obs.  No seed is special: the fp32
    mode carries the rudder in 2^-42 rad fixed point, so the pi/3 and pi/4 thresholds (boat_env.py:102,107) fall
    on the reference's step (round 1 passed only on seeds that happened not to cross within 1e-6)."""
    cfg = S.load_config(base_settings__experiment=6)
    n, T, E = 2048, 1500, 40
    env = make_env(S, cfg, n, precision, seed=seed, auto_reset=True)
    s_y, knots = env.episode_draws_batch(np.arange(n), E)
    env.reset()
    import torch
    actions = np_(torch.stack([env.uniform_actions(t, 1.0).clone() for t in range(T)]))
    ref = O.rollout(O.params_from_config(cfg), actions, s_y, knots, auto_reset=True)
    out = run_steps(env, actions)
    assert np.array_equal(out["done"], ref["done"])
    assert np.array_equal(out["term"], ref["term"])
    tol = TOL64 if precision == "fp64" else TOL32
    d = ref["done"].astype(bool)
    assert scaled_err(out["obs"][~d], ref["obs"][~d]).max() <= tol
    assert scaled_err(out["final_obs"][d], ref["obs"][d]).max() <= tol
    assert scaled_err(out["reward"], ref["reward"]).max() <= tol
    # under auto-reset obs at a done step is the NEW episode's reset observation
    reset_rows = out["obs"][d]
    assert np.all(reset_rows[:, [0, 1, 2, 4, 5, 6, 7, 8]] == 0) and np.all(reset_rows[:, 9] == 0.5)
    assert np.all(reset_rows[:, 10] == 1.0) and np.all(reset_rows[:, 3] == 0.5)
    c = env.counters()
    for code, name in enumerate(S.TERM_NAMES):
        if code:
            assert c[name] == float((ref["term"] == code).sum())
    assert c["episodes"] == float(d.sum()) and d.sum() > n
    assert np_(env.get_field("episode")).max() < E
    env.close()


def test_step_k_equals_k_single_steps(S):
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    n, K, rounds = 2048, 8, 30
    for precision in ("fp32", "fp64"):
        a = make_env(S, cfg, n, precision, seed=9, auto_reset=False)
        b = make_env(S, cfg, n, precision, seed=9, auto_reset=False)
        a.reset(); b.reset()
        for r in range(rounds):
            acts = torch.stack([a.uniform_actions(r * K + k, 0.3).clone() for k in range(K)])
            rsum = torch.zeros(n, dtype=a.dtype, device=a.device)
            alive = torch.ones(n, dtype=torch.bool, device=a.device)
            # reference semantics of the fused window: an env stops at its first done
            snap = {f: a.get_field(f).clone() for f in ("s_x", "s_y", "v_x", "rudder_angle", "index")}
            for k in range(K):
                # single steps cannot freeze individual envs, so compare only envs alive through the window
                o, rw, d, _ = a.step(acts[k])
                rsum += torch.where(alive, rw, torch.zeros_like(rw))
                alive &= d == 0
            ob, rb, db, info = b.step_k(acts, K)
            ok = alive
            assert torch.equal(o[ok], ob[ok])
            assert torch.allclose(rsum[ok], rb[ok], rtol=1e-5 if precision == "fp32" else 1e-12, atol=0)
            assert torch.all(db[ok] == 0) and torch.all(info["steps"][ok] == K)
            assert torch.all(db[~ok] == 1)
            # resynchronise b's finished envs with a (a kept stepping them): the state blob is an exact copy
            # (set_field(get_field()) rounds the fp32 mode's fixed-point rudder / positions to fp32)
            b.load_state_dict(a.state_dict())
        a.close(); b.close()


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("per_substep_actions", [False, True])
def test_step_k_auto_reset_matches_oracle(S, O, precision, per_substep_actions):
    """K fused sub-steps with auto-reset against an env-by-env emulation on the oracle: an env stops at its
    first done inside the window (reward = sum over executed sub-steps, steps = their number), the next
    window starts its next episode -- whose wind coefficients come from the deferred episode-end queue."""
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    n, K, W, E = 192, 5, 90, 80
    env = make_env(S, cfg, n, precision, seed=31, auto_reset=True)
    s_y, knots = host_draws(env, E)
    env.reset()
    p = O.params_from_config(cfg)
    oracles = [O.OracleEnv(p) for _ in range(n)]
    episode = [0] * n
    for i, o in enumerate(oracles):
        o.reset(int(s_y[0, i]), knots[0, i, 0], knots[0, i, 1])
    tol = TOL64 if precision == "fp64" else TOL32
    n_done = 0
    for w in range(W):
        if per_substep_actions:
            acts = torch.stack([env.uniform_actions(w * K + k, 1.5).clone() for k in range(K)])
        else:
            acts = env.uniform_actions(w, 1.5).clone()
        obs, rew, done, info = env.step_k(acts, K)
        a_np, obs_np, rew_np = np_(acts), np_(obs), np_(rew)
        done_np, term_np, steps_np = np_(done), np_(info["term"]), np_(info["steps"])
        for i, o in enumerate(oracles):
            r_sum, d, code, steps, last = 0.0, False, 0, 0, None
            for k in range(K):
                a = a_np[k, i] if per_substep_actions else a_np[i]
                last, r, d, code = o.step(float(a))
                r_sum += r
                steps += 1
                if d:
                    break
            assert bool(done_np[i]) == d and int(term_np[i]) == code and int(steps_np[i]) == steps
            assert abs(rew_np[i] - r_sum) <= tol * max(1.0, abs(r_sum))
            if d:
                n_done += 1
                episode[i] += 1
                last = o.reset(int(s_y[episode[i], i]), knots[episode[i], i, 0], knots[episode[i], i, 1])
            assert scaled_err(obs_np[i], last).max() <= tol
    assert n_done > n  # plenty of episode ends went through the queue
    assert env.counters()["episodes"] == n_done
    env.close()


def test_shard_invariance(S):
    """Philox is keyed by the GLOBAL env id: a 1024-env population equals two 512-env
    shards (what ranks 0 and 1 of a 2-GPU run hold), bit for bit."""
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    whole = make_env(S, cfg, 1024, "fp32", seed=11, auto_reset=True)
    lo = make_env(S, cfg, 512, "fp32", seed=11, auto_reset=True, env_id_offset=0)
    hi = make_env(S, cfg, 512, "fp32", seed=11, auto_reset=True, env_id_offset=512)
    for e in (whole, lo, hi):
        e.reset()
    for t in range(800):
        aw = whole.uniform_actions(t)
        assert torch.equal(aw[:512], lo.uniform_actions(t)) and torch.equal(aw[512:], hi.uniform_actions(t))
        ow, rw, dw, _ = whole.step(aw)
        ol, rl, dl, _ = lo.step(aw[:512])
        oh, rh, dh, _ = hi.step(aw[512:])
        assert torch.equal(ow[:512], ol) and torch.equal(ow[512:], oh)
        assert torch.equal(rw[:512], rl) and torch.equal(dw[512:], dh)
    # the action stream is keyed by global env id too, also across an odd (non quad-aligned) split
    odd_lo = make_env(S, cfg, 513, "fp32", seed=11, env_id_offset=0)
    odd_hi = make_env(S, cfg, 511, "fp32", seed=11, env_id_offset=513)
    aw = whole.uniform_actions(5, 0.7)
    assert torch.equal(aw[:513], odd_lo.uniform_actions(5, 0.7)) and torch.equal(aw[513:], odd_hi.uniform_actions(5, 0.7))
    assert float(aw.abs().max()) <= 0.7 and float(aw.std()) == pytest.approx(0.7 / 3 ** 0.5, rel=0.1)
    odd_lo.close(); odd_hi.close()
    cw, cl, ch = whole.counters(), lo.counters(), hi.counters()
    assert cw["episodes"] > 0
    for k in ("reached_goal", "out_of_bounds", "out_of_fuel", "timeout", "rudder_broken", "episodes"):
        assert cw[k] == cl[k] + ch[k]
    assert cw["return_sum"] == pytest.approx(cl["return_sum"] + ch["return_sum"], rel=1e-9)
    for e in (whole, lo, hi):
        e.close()


@pytest.mark.parametrize("experiment", range(1, 7))
def test_wind_tables_and_draws(S, O, experiment):
    """Device Philox == host Philox (boatenv_episode_draws_host), and the wind the kernels
    see == the oracle's wind.py tables for those knots."""
    cfg = S.load_config(base_settings__experiment=experiment)
    env = make_env(S, cfg, 64, "fp64", seed=77, auto_reset=False)
    obs0 = np_(env.reset())
    p = O.params_from_config(cfg)
    for i in (0, 1, 17, 63):
        s_y, knots = env.episode_draws(i, 0)
        assert -640 <= s_y < 640 and np.all((knots > 0) & (knots < 1))
        assert np.all(knots == knots.astype(np.float32))  # fp32-representable: both modes see the same knots
        o = O.OracleEnv(p)
        ref_obs0 = o.reset(s_y, knots[0], knots[1])
        assert np.abs(obs0[i] - ref_obs0).max() < 1e-15
        wv, wa = env.wind_table(i)
        rv, ra = o.wind()
        assert np.abs(wv - rv).max() < 1e-13 and np.abs(wa - ra).max() < 1e-12
    env.close()


@pytest.mark.parametrize("n,pinned", [
    (300_000, True),   # > 4 * 65536: the chunked copy/compute overlap
    (2048, True), (1000, False), (33, True), (1, False),  # <= 2048: the zero-copy path (mapped pinned staging)
    (2049, False),     # smallest size of the copy-engine path, pageable buffers
])
def test_step_host_equals_step(S, n, pinned):
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    a = make_env(S, cfg, n, "fp32", seed=2, auto_reset=True)
    b = make_env(S, cfg, n, "fp32", seed=2, auto_reset=True)
    a.reset(); b.reset()
    pin = (lambda x: x.pin_memory()) if pinned else (lambda x: x)
    act_h = pin(torch.empty(n, dtype=torch.float32))
    obs_h = pin(torch.empty((n, 11), dtype=torch.float32))
    rew_h = pin(torch.empty(n, dtype=torch.float32))
    done_h = pin(torch.empty(n, dtype=torch.uint8))
    for t in range(12):
        acts = a.uniform_actions(t, 4.0)  # big steps: resets inside the run
        act_h.copy_(acts)
        o, r, d, _ = a.step(acts)
        torch.cuda.synchronize()
        b.step_host(act_h, obs_h, rew_h, done_h)
        assert torch.equal(o.cpu(), obs_h) and torch.equal(r.cpu(), rew_h) and torch.equal(d.cpu(), done_h)
    a.close(); b.close()


@pytest.mark.parametrize("n,precision", [(300_000, "fp32"), (1000, "fp32"), (70_000, "fp64")])
def test_step_k_host_equals_step_k(S, n, precision):
    """boatenv_step_k_host (actions [K][N] from host memory, one observation per env back) == boatenv_step_k on a
    twin, bit for bit, incl. the deferred episode-end queue and the chunked copy / compute pipeline."""
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    K = 8
    a = make_env(S, cfg, n, precision, seed=2, auto_reset=True)
    b = make_env(S, cfg, n, precision, seed=2, auto_reset=True)
    a.reset(); b.reset()
    dt = a.dtype
    act_h = torch.empty((K, n), dtype=dt).pin_memory()
    obs_h = torch.empty((n, 11), dtype=dt).pin_memory()
    rew_h = torch.empty(n, dtype=dt).pin_memory()
    done_h = torch.empty(n, dtype=torch.uint8).pin_memory()
    term_h = torch.empty(n, dtype=torch.uint8).pin_memory()
    steps_h = torch.empty(n, dtype=torch.int32).pin_memory()
    for w in range(10):
        acts = torch.stack([a.uniform_actions(w * K + k, 3.0).clone() for k in range(K)])
        act_h.copy_(acts)
        o, r, d, info = a.step_k(acts, K)
        torch.cuda.synchronize()
        b.step_k_host(act_h, K, obs_h, rew_h, done_h, term_h, steps_h)
        assert torch.equal(o.cpu(), obs_h) and torch.equal(r.cpu(), rew_h) and torch.equal(d.cpu(), done_h)
        assert torch.equal(info["term"].cpu(), term_h) and torch.equal(info["steps"].cpu(), steps_h)
    assert int(done_h.sum()) > 0
    ca, cb = a.counters(), b.counters()
    assert ca["episodes"] == cb["episodes"] > 0
    a.close(); b.close()


def test_step_host_does_not_serialise_other_streams(S):
    """boatenv_step_host_stream waits for an EVENT on the caller's stream, not for the device: a long kernel queued
    on an unrelated stream is still running when step_host returns."""
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    n = 4096
    env = make_env(S, cfg, n, "fp32", seed=2, auto_reset=True)
    env.reset()
    act_h = torch.zeros(n, dtype=torch.float32).pin_memory()
    obs_h = torch.empty((n, 11), dtype=torch.float32).pin_memory()
    rew_h = torch.empty(n, dtype=torch.float32).pin_memory()
    done_h = torch.empty(n, dtype=torch.uint8).pin_memory()
    env.step_host(act_h, obs_h, rew_h, done_h)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    x = torch.randn(8192, 8192, device="cuda")
    marker = torch.cuda.Event()
    with torch.cuda.stream(side):
        for _ in range(40):          # ~100 ms of matmuls on the side stream
            x = (x @ x).clamp_(-1, 1)
        marker.record()
    env.step_host(act_h, obs_h, rew_h, done_h)
    still_running = not marker.query()
    torch.cuda.synchronize()
    assert still_running, "step_host waited for an unrelated stream (device-wide synchronise?)"
    env.close()


def test_error_behaviour(S):
    """ValueError for an unknown experiment / fixed_points < 4 (wind.py:65-67, 73-75)."""
    with pytest.raises(ValueError):
        S.BatchedBoatEnv(S.load_config(base_settings__experiment=7), 4, device=0)
    with pytest.raises(ValueError):
        S.BatchedBoatEnv(S.load_config(base_settings__experiment=4, wind__fixed_points=3), 4, device=0)
    S.BatchedBoatEnv(S.load_config(base_settings__experiment=3, wind__fixed_points=3), 4, device=0).close()
    env = S.BatchedBoatEnv(S.load_config(), 4, device=0)
    with pytest.raises(S.BoatEnvError):
        env.step(np.zeros(4, dtype=np.float32))  # step before reset
    env.close()


def test_single_env_drop_in(S, O):
    """The reference's BoatEnv object protocol: 4-tuple step, cumulative
    info with a sticky 'termination', episode_reward reset by reset(), return_all_data."""
    cfg = S.load_config(base_settings__experiment=3, boat__fuel=60)
    env = S.BoatEnv(cfg, experiment=None, seed=4, precision="fp64", device=0)
    assert env.observation_space.shape == (11,) and env.action_space.shape[0] == 1
    assert float(env.action_space.high[0]) == 1.0
    p = O.params_from_config(cfg)
    for episode in range(2):
        o = O.OracleEnv(p)
        ref = o.reset(0)
        obs = env.reset()
        assert isinstance(obs, np.ndarray) and obs.dtype == np.float64 and obs.shape == (11,)
        assert np.array_equal(obs, ref) and env.info["episode_reward"] == 0
        done, total = False, 0.0
        rng = np.random.default_rng(episode)
        while not done:
            a = np.array([np.float32(0.05 * rng.uniform(-1, 1))], dtype=np.float32)
            obs, reward, done, info = env.step(a)
            ro, rr, rd, rc = o.step(float(a[0]))
            assert scaled_err(obs, ro).max() <= TOL64 and abs(reward - rr) <= 1e-9 and done == rd
            total += reward
            assert info is env.info
        assert info["termination"] == "out_of_fuel" and info["out_of_fuel"] == episode + 1
        assert info[info["termination"]] == episode + 1  # main.py:110
        assert info["episode_reward"] == pytest.approx(total)
        d = env.return_all_data()
        assert set(d) == {"boat_position_x", "boat_position_y", "boat_velocity_x", "boat_velocity_y", "boat_angle",
                          "action_rudder", "reward", "rudder_angle", "n"}
        assert np.allclose([d["boat_position_x"], d["boat_position_y"], d["boat_velocity_x"], d["boat_velocity_y"],
                            d["boat_angle"], d["reward"], d["rudder_angle"]], o.all_data()[[0, 1, 2, 3, 4, 6, 7]],
                           rtol=1e-9, atol=1e-12)
        assert len(env.boat.wind.wind_velocity) == 10000
    env.close()


@pytest.mark.parametrize("stream", ["A1", "A2", "A3"])
def test_config0_single_env_action_streams(S, O, stream):
    """BASELINE.json configs[0] / SURVEY.md 8(d) config 1: experiment 1, ONE env (the reference's own use), one
    episode of fixed random actions from np.random.default_rng(0), through the single-env drop-in `BoatEnv`:
    A1 float32(U(-1,1)) (ends by rudder_broken after a few hundred steps), A2 float32(0.05 U(-1,1)) (thousands of
    steps), A3 zeros with test_mode = 1 (fixture 1: reached_goal after 4959 steps).  Step count, termination and
    every observation / reward against the oracle (fp64: 1e-9)."""
    over = dict(base_settings__experiment=1)
    if stream == "A3":
        over["base_settings__test_mode"] = 1
    cfg = S.load_config(**over)
    env = S.BoatEnv(cfg, experiment=None, seed=0, precision="fp64", device=0)
    o = O.OracleEnv(O.params_from_config(cfg))
    obs, ref = env.reset(), o.reset(0)
    assert np.array_equal(obs, ref)
    rng = np.random.default_rng(0)
    scale = {"A1": 1.0, "A2": 0.05, "A3": 0.0}[stream]
    steps, done, total = 0, False, 0.0
    while not done and steps < 10001:
        a = np.array([np.float32(scale * rng.uniform(-1, 1))], dtype=np.float64)   # float32-representable, passed as float64 (H4)
        obs, reward, done, info = env.step(a)
        ro, rr, rd, rc = o.step(float(a[0]))
        steps += 1
        total += reward
        assert done == rd and scaled_err(obs, ro).max() <= TOL64 and abs(reward - rr) <= TOL64 * max(1.0, abs(rr)), steps
    assert done and info["termination"] == O.TERM_NAMES[rc] and info["episode_reward"] == pytest.approx(total)
    if stream == "A1":
        assert info["termination"] == "rudder_broken" and 20 < steps < 3000
    if stream == "A3":
        assert info["termination"] == "reached_goal" and steps == 4959
    env.close()


def test_full_size_properties(S):
    """BASELINE.json configs[2] shape (exp 6, fp32, millions of envs): properties that do
    not need the oracle -- determinism, counter consistency, fuel bookkeeping."""
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    n, T = 16 * 1024 * 1024, 40   # the full per-GPU population of the benchmark
    runs = []
    for _ in range(2):
        env = make_env(S, cfg, n, "fp32", seed=1, auto_reset=True)
        env.reset()
        dones = torch.zeros(n, dtype=torch.int64, device=env.device)
        for t in range(T):
            obs, rew, done, info = env.step(env.uniform_actions(t, 4.0))  # big steps: rudder breaks within ~10 steps
            dones += done
            assert torch.all((info["term"] > 0) == (done > 0))
        c = env.counters()
        assert c["episodes"] == float(dones.sum().item()) and c["episodes"] > n
        assert c["episodes"] == sum(c[k] for k in ("reached_goal", "out_of_bounds", "out_of_fuel", "timeout",
                                                   "rudder_broken"))
        idx = env.get_field("index")
        assert torch.equal(obs[:, 10], (15000.0 - idx.float()) / 15000.0) or \
            torch.allclose(obs[:, 10], (15000.0 - idx.float()) / 15000.0, rtol=1e-6)
        assert torch.equal(env.get_field("episode").long(), dones)
        runs.append((obs.clone(), rew.clone(), c))
        env.close()
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    assert runs[0][2]["episodes"] == runs[1][2]["episodes"]


def test_checkpoint_resume_is_exact(S, tmp_path):
    """the reference saves only network weights; here the env state is one blob and
    Philox is counter-based, so save -> keep running == restore -> run again, bit for bit (incl. statistics)."""
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    n = 50_000
    env = make_env(S, cfg, n, "fp32", seed=13, auto_reset=True)
    env.reset()
    for t in range(150):
        env.step(env.uniform_actions(t, 2.0))
    sd = env.state_dict()
    torch.save(sd, tmp_path / "env.pt")
    tail_a = []
    for t in range(150, 300):
        o, r, d, info = env.step(env.uniform_actions(t, 2.0))
        tail_a.append((o.clone(), r.clone(), d.clone(), info["term"].clone()))
    ca = env.counters()
    twin = make_env(S, cfg, n, "fp32", seed=13, auto_reset=True)   # a fresh process would do exactly this
    twin.load_state_dict(torch.load(tmp_path / "env.pt"))
    assert torch.equal(twin.obs, sd["obs"])
    for t, (o, r, d, tm) in zip(range(150, 300), tail_a):
        o2, r2, d2, info2 = twin.step(twin.uniform_actions(t, 2.0))
        assert torch.equal(o2, o) and torch.equal(r2, r) and torch.equal(d2, d) and torch.equal(info2["term"], tm)
    cb = twin.counters()
    assert ca["episodes"] == cb["episodes"] > 0 and ca["rudder_broken"] == cb["rudder_broken"]
    assert cb["return_sum"] == pytest.approx(ca["return_sum"], rel=1e-9)
    other = make_env(S, cfg, n, "fp32", seed=14)
    with pytest.raises(ValueError):
        other.load_state_dict(sd)
    for e in (env, twin, other):
        e.close()


def test_recorder_compatible_export(S, tmp_path):
    """the CSVs of postprocessing/recorder.py from a GPU run.  Replaying fixture 6
    (test_mode 1) and reading the files back the way the reference's Replayer does (pandas, sep=';')
    reproduces the recorded episode_0_data.csv / wind.csv / info.csv."""
    import pandas as pd
    import torch
    g = load_golden("fixture_exp6")
    fp = int(g["config"]["wind"]["fixed_points"])
    knots = np.zeros((1, 2, fp))
    knots[0, 0], knots[0, 1] = g["knots_v"], g["knots_a"]
    env = make_env(S, g["config"], 1, "fp64", np.array([0]), knots, auto_reset=True)
    env.reset()
    rec = S.BatchedRecorder(env, [0], str(tmp_path))
    rec.write_winds_to_csv()
    zero = torch.zeros(1, dtype=env.dtype, device=env.device)
    done = False
    while not done:
        rec.write_data_to_csv()
        _, _, d, _ = env.step(zero)
        rec.after_step(zero)
        done = bool(d[0])
    data = pd.read_csv(tmp_path / "episodes" / "episode_0_data.csv", sep=";")
    assert list(data.columns) == json.loads('["boat_position_x", "boat_position_y", "boat_velocity_x", "boat_velocity_y", '
                                            '"boat_angle", "action_rudder", "reward", "rudder_angle", "n"]')
    assert len(data) == int(g["n_rows"])
    rows = data.values[g["row_idx"]]
    ref = g["rows"]
    for col, scale in ((0, 3900.0), (1, 800.0), (2, 5.0), (3, 2.0), (4, 2 * np.pi)):
        assert scaled_err(rows[:, col], ref[:, col], scale).max() <= 1e-11
    m = g["row_idx"] > 0
    assert np.abs(rows[m, 6] - ref[m, 6]).max() <= 1e-11 and np.all(rows[:, 8] == 20)
    wind = pd.read_csv(tmp_path / "episodes" / "wind.csv", sep=";")
    assert list(wind.columns) == ["wind_velocity", "wind_angle"] and len(wind) == 10000
    assert np.abs(wind.wind_velocity.values[g["wind_idx"]] - g["wind_v"]).max() < 5e-14
    info = pd.read_csv(tmp_path / "episodes" / "info.csv", sep=";")
    assert info.termination[0] == "reached_goal" and info.reached_goal[0] == 1
    assert info.episode_reward[0] == pytest.approx(871.2727580297085, abs=1e-8)
    env.close()
