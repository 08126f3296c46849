"""Edge cases of the batched step path against the oracle: ragged env counts (not a multiple of the
32-env state block / the 256-thread CTA), the smallest and largest supported spline sizes, very short
wind tables, stepping past the end of the wind table, unclipped actions, argument errors."""
import ctypes as C

import numpy as np
import pytest

from boat_testlib import scaled_err

pytestmark = pytest.mark.gpu


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


def np_(t):
    return t.detach().double().cpu().numpy() if t.is_floating_point() else t.detach().cpu().numpy()


def rollout_both(S, O, cfg, n, T, precision, scale, auto_reset, episodes, seed=5, k=1):
    import torch
    env = S.BatchedBoatEnv(cfg, n, seed=seed, precision=precision, device=0, auto_reset=auto_reset)
    fp = int(env.params.fixed_points)
    s_y = np.empty((episodes, n), dtype=np.int32)
    knots = np.empty((episodes, n, 2, fp))
    for e in range(episodes):
        for i in range(n):
            s_y[e, i], knots[e, i] = env.episode_draws(i, e)
    env.reset()
    acts = torch.stack([env.uniform_actions(t, scale).clone() for t in range(T)])
    ref = O.rollout(O.params_from_config(cfg), np_(acts), s_y, knots, auto_reset=auto_reset)
    obs = np.empty((T, n, 11)); rew = np.empty((T, n)); done = np.empty((T, n), np.uint8); term = np.empty((T, n), np.uint8)
    fin = np.zeros((T, n, 11))
    for t in range(T):
        o, r, d, info = env.step(acts[t])
        obs[t], rew[t], done[t], term[t], fin[t] = np_(o), np_(r), np_(d), np_(info["term"]), np_(info["final_obs"])
    env.close()
    return ref, dict(obs=obs, reward=rew, done=done, term=term, final_obs=fin)


def assert_parity(ref, out, tol, auto_reset):
    assert np.array_equal(out["done"], ref["done"]) and np.array_equal(out["term"], ref["term"])
    d = ref["done"].astype(bool)
    if auto_reset:
        assert scaled_err(out["obs"][~d], ref["obs"][~d]).max() <= tol
        if d.any():
            assert scaled_err(out["final_obs"][d], ref["obs"][d]).max() <= tol
    else:
        assert scaled_err(out["obs"], ref["obs"]).max() <= tol
    assert scaled_err(out["reward"], ref["reward"]).max() <= tol


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 255, 257, 1000])
@pytest.mark.parametrize("precision,tol", [("fp64", 1e-9), ("fp32", 1e-4)])
def test_ragged_env_counts(S, O, n, precision, tol):
    cfg = S.load_config(base_settings__experiment=6, base_settings__t_max=100)  # L = 400: pieces cross every 57 steps
    ref, out = rollout_both(S, O, cfg, n, 260, precision, 1.0, True, 24)
    assert_parity(ref, out, tol, True)
    assert n < 30 or ref["done"].sum() > 0


@pytest.mark.parametrize("experiment", [4, 5, 6])
@pytest.mark.parametrize("fixed_points", [4, 5, 11, 16])
def test_spline_sizes(S, O, experiment, fixed_points):
    """wind.fixed_points from the minimum the reference accepts (4, wind.py:73) to this build's maximum (16)."""
    cfg = S.load_config(base_settings__experiment=experiment, wind__fixed_points=fixed_points,
                        base_settings__t_max=250)
    ref, out = rollout_both(S, O, cfg, 96, 400, "fp64", 0.05, True, 4)
    assert_parity(ref, out, 1e-9, True)
    env = S.BatchedBoatEnv(cfg, 8, seed=5, precision="fp64", device=0)
    env.reset()
    p = O.params_from_config(cfg)
    for i in (0, 7):  # wind tables == wind.py's for these knots
        s_y, k = env.episode_draws(i, 0)
        o = O.OracleEnv(p)
        o.reset(s_y, k[0], k[1])
        wv, wa = env.wind_table(i)
        rv, ra = o.wind()
        assert np.abs(wv - rv).max() < 1e-13 and np.abs(wa - ra).max() < 1e-12
    env.close()
    if fixed_points == 16:
        with pytest.raises(S.BoatEnvError):   # valid for the reference, unsupported by this build
            S.BatchedBoatEnv(S.load_config(base_settings__experiment=6, wind__fixed_points=17), 4, device=0)


@pytest.mark.parametrize("t_max,dt", [(2.5, 0.25), (9.25, 0.25), (30, 0.5), (3, 0.125)])
def test_short_wind_tables_and_timeouts(S, O, t_max, dt):
    """L = int(t_max / dt) as small as 10: every step is a piece crossing; timeout after L steps.
    (dt must be a binary fraction: with dt = 0.1 the reference's accumulated t falls short of t_max after
    L steps and its next step indexes past the wind table -- IndexError, wind.py:24.)"""
    cfg = S.load_config(base_settings__experiment=6, base_settings__t_max=t_max, base_settings__dt=dt)
    L = int(t_max / dt)
    ref, out = rollout_both(S, O, cfg, 64, 3 * L + 5, "fp64", 0.02, True, 8)
    assert_parity(ref, out, 1e-9, True)
    assert (ref["term"] == 4).sum() >= 64 * 2  # everybody times out, repeatedly


def test_stepping_past_done_without_reset(S, O):
    """Without auto-reset the reference object keeps integrating after done; the CUDA path does too
    (the oracle's own IndexError at the end of the wind table is the reference's, wind.py:24)."""
    cfg = S.load_config(base_settings__experiment=4, boat__fuel=40)
    ref, out = rollout_both(S, O, cfg, 40, 120, "fp64", 0.05, False, 1)
    assert_parity(ref, out, 1e-9, False)
    assert (ref["term"] == 3).sum() == 40 * (120 - 40)  # out_of_fuel on every step after the 40th


def test_unclipped_actions(S, O):
    """boat_env.py:73 does not clip: |action| = 25 drives the rudder past pi/3 in one step."""
    import torch
    cfg = S.load_config(base_settings__experiment=3)
    env = S.BatchedBoatEnv(cfg, 64, seed=1, precision="fp64", device=0, auto_reset=False)
    env.reset()
    a = torch.full((64,), 25.0, dtype=torch.float64, device="cuda")
    a[::2] = -25.0
    o = O.OracleEnv(O.params_from_config(cfg))
    o.reset(0)
    ro, rr, rd, rc = o.step(25.0)
    obs, rew, done, info = env.step(a)
    assert bool(done.all()) and bool((info["term"] == 5).all()) and rd and rc == 5
    assert scaled_err(np_(obs)[1], ro).max() <= 1e-9 and abs(float(rew[1]) - rr) <= 1e-9 * abs(rr)
    assert float(rew[1]) < -200.0  # the rudder penalty of boat_env.py:107-108 applies on the terminal step too
    env.close()


def test_argument_errors(S):
    L = S.lib()
    p = S.params_from_config(S.load_config())
    h = C.c_void_p()
    assert L.boatenv_create(C.byref(p), 0, 0, 0, 32, 0, C.byref(h)) == -1        # empty population
    assert L.boatenv_create(C.byref(p), 4, 0, 0, 16, 0, C.byref(h)) == -1        # unknown precision
    assert L.boatenv_create(C.byref(p), 4, 0, -1, 32, 0, C.byref(h)) == -1       # negative env id offset
    assert L.boatenv_create(C.byref(p), 4, 0, 0, 32, 99, C.byref(h)) == -5       # no such device
    assert L.boatenv_create(None, 4, 0, 0, 32, 0, C.byref(h)) == -1
    assert L.boatenv_create(C.byref(p), 4, 0, 0, 32, 0, C.byref(h)) == 0
    assert L.boatenv_step(h, None, None, None, None, None, None, 0, None) == -6  # step before reset
    assert L.boatenv_reset(h, None, None, None) == 0
    assert L.boatenv_step(h, None, None, None, None, None, None, 0, None) == -1  # null tensors
    assert L.boatenv_get_field(h, 99, None, None) == -1
    assert L.boatenv_wind_table(h, 4, None, None, None) == -1
    assert L.boatenv_destroy(h) == 0 and L.boatenv_destroy(None) == -1
    r = C.c_void_p()
    assert L.boatreplay_create(0, 11, 1, 32, 0, C.byref(r)) == -1
    assert L.boatreplay_create(8, 11, 1, 32, 0, C.byref(r)) == 0
    assert L.boatreplay_sample(r, 4, 0, 0, 1, 1, 1, 1, 1, None, None) == -6      # empty buffer (np.random.choice(0, n))
    assert L.boatreplay_destroy(r) == 0
    t = C.c_void_p()
    arr = (C.c_double * 4)(10, 10, 0.01, 0.1)
    assert L.boattoy_create(7, 4, arr, 4, 0.0, 0, 32, 0, C.byref(t)) == -1       # unknown toy
    assert L.boattoy_create(0, 4, arr, 3, 0.0, 0, 32, 0, C.byref(t)) == -1       # wrong parameter count
    bad = S.BatchedBoatEnv(S.load_config(), 5, device=0)
    bad.reset()
    with pytest.raises(ValueError):
        bad.step_k(np.zeros((3, 5), dtype=np.float32), 4)                        # [k, N] actions with the wrong k
    bad.close()


def test_masked_reset(S, O):
    """reset(mask): BoatEnv.reset (boat_env.py:120-126) for the selected envs only -- new Boat, new Wind
    (next episode's Philox draws), reset observation row; everybody else keeps stepping undisturbed."""
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    n = 300
    a = S.BatchedBoatEnv(cfg, n, seed=9, precision="fp64", device=0, auto_reset=False)
    b = S.BatchedBoatEnv(cfg, n, seed=9, precision="fp64", device=0, auto_reset=False)
    a.reset(); b.reset()
    for t in range(40):
        acts = a.uniform_actions(t, 0.2)
        a.step(acts); b.step(acts)
    mask = torch.zeros(n, dtype=torch.bool, device="cuda")
    mask[::7] = True
    before = b.obs.clone()
    obs = b.reset(mask)
    assert torch.equal(obs[~mask], before[~mask])
    assert torch.all(obs[mask][:, [0, 1, 2, 4, 5, 6, 7, 8]] == 0) and torch.all(obs[mask][:, 3] == 0.5)
    assert torch.all(b.get_field("episode")[mask] == 1) and torch.all(b.get_field("episode")[~mask] == 0)
    assert torch.all(b.get_field("index")[mask] == 0) and torch.all(b.get_field("index")[~mask] == 40)
    with pytest.raises(ValueError):
        b.reset(mask[:10])
    p = O.params_from_config(cfg)
    for t in range(40, 70):
        acts = a.uniform_actions(t, 0.2)
        oa, ra, _, _ = a.step(acts)
        ob, rb, _, _ = b.step(acts)
        assert torch.equal(oa[~mask], ob[~mask]) and torch.equal(ra[~mask], rb[~mask])
    # a re-started env follows the oracle fed with episode 1's draws
    i = 7
    s_y, knots = b.episode_draws(i, 1)
    o = O.OracleEnv(p)
    o.reset(s_y, knots[0], knots[1])
    for t in range(40, 70):
        ro, rr, rd, rc = o.step(float(a.uniform_actions(t, 0.2)[i]))
    assert scaled_err(np_(b.obs)[i], ro).max() <= 1e-9
    a.close(); b.close()


def test_single_env_step_latency(S):
    """The single-env drop-in makes ONE blocking C-ABI call per step (pinned host buffers): it must not be
    slower than the reference's own CPU step (0.10-0.35 ms) it replaces."""
    import time
    env = S.BoatEnv(S.load_config(base_settings__experiment=6), seed=1, precision="fp64", device=0)
    env.reset()
    a = np.array([0.01], dtype=np.float32)
    for _ in range(50):
        env.step(a)
    t0 = time.perf_counter()
    for _ in range(300):
        env.step(a)
    per_step = (time.perf_counter() - t0) / 300
    t0 = time.perf_counter()
    for _ in range(300):   # main.py:79-81 with the Recorder attached: return_all_data() after every step
        env.step(a)
        d = env.return_all_data()
    per_step_rec = (time.perf_counter() - t0) / 300
    print(f"single-env BoatEnv.step: {per_step * 1e6:.0f} us; step + return_all_data: {per_step_rec * 1e6:.0f} us")
    assert per_step < 0.2e-3 and per_step_rec < 0.4e-3
    assert d["boat_position_x"] > 0 and d["n"] == 20
    env.close()


def test_env_state_host_and_zero_copy_errors(S):
    """boatenv_env_state_host: all ten fields of one env in one call, equal to the per-field reads; argument and
    state errors come back as codes (re-raised by the binding), never as a crash."""
    import ctypes as C
    cfg = S.load_config(base_settings__experiment=4)
    env = S.BatchedBoatEnv(cfg, 100, seed=2, precision="fp64", device=0, auto_reset=True)
    out = (C.c_double * 10)()
    assert env._L.boatenv_env_state_host(env._h, 0, out) != 0   # before reset: BOATENV_ESTATE
    env.reset()
    for t in range(7):
        env.step(env.uniform_actions(t, 0.5))
    for i in (0, 31, 32, 99):
        assert env._L.boatenv_env_state_host(env._h, i, out) == 0
        for name, f in S.boat_env.FIELDS.items():
            assert float(env.get_field(name)[i].item()) == out[f], (i, name)
    assert env._L.boatenv_env_state_host(env._h, 100, out) != 0 and env._L.boatenv_env_state_host(env._h, -1, out) != 0
    assert env._L.boatenv_env_state_host(env._h, 0, None) != 0
    # zero-copy host step in fp64 at the size limit, against the device step of a twin
    twin = S.BatchedBoatEnv(cfg, 100, seed=2, precision="fp64", device=0, auto_reset=True)
    twin.reset()
    for t in range(7):
        twin.step(twin.uniform_actions(t, 0.5))
    acts = env.uniform_actions(7, 0.5)
    o, r, d, _ = twin.step(acts)
    ah, oh = acts.cpu().numpy().copy(), np.empty((100, 11))
    rh, dh = np.empty(100), np.empty(100, dtype=np.uint8)
    rc = env._L.boatenv_step_host(env._h, ah.ctypes.data, oh.ctypes.data, rh.ctypes.data, dh.ctypes.data, 1)
    assert rc == 0 and np.array_equal(oh, o.cpu().numpy()) and np.array_equal(rh, r.cpu().numpy())
    assert np.array_equal(dh, d.cpu().numpy())
    env.close(); twin.close()


def test_whole_launch_step_is_deterministic(S):
    """Two identical handles stepped alternately with the same actions agree bit for bit: several state blocks per
    persistent warp (the TMA stage is refilled while the warp computes), frequent resets, a ragged last block.
    Regression test for the stage-refill race: a bulk copy served from L2 could land before the last
    shared-memory loads of the stage had returned (profiles/determinism_check.py: 16 of 40 trials failed)."""
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    n = 300_003
    for trial in range(6):
        a = S.BatchedBoatEnv(cfg, n, seed=2, precision="fp32", device=0, auto_reset=True)
        b = S.BatchedBoatEnv(cfg, n, seed=2, precision="fp32", device=0, auto_reset=True)
        a.reset(); b.reset()
        for t in range(14):
            acts = a.uniform_actions(t, 4.0)
            oa, ra, da, _ = a.step(acts)
            torch.cuda.synchronize()
            ob, rb, db, _ = b.step(acts)
            torch.cuda.synchronize()
            assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db), (trial, t)
        a.close(); b.close()


def test_fp32_exact_carriers_field_round_trip(S):
    """fp32 mode: rudder (44-bit fixed point, 2^-42 rad) and s_x / s_y (int32 fixed point) are exchanged as plain
    numbers by get_field / set_field / env_state_host; the step index shares its word with the rudder's low bits."""
    import ctypes as C
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    env = S.BatchedBoatEnv(cfg, 70, seed=5, precision="fp32", device=0, auto_reset=False)
    env.reset()
    rud = torch.linspace(-1.2, 1.2, 70, device=env.device)
    sx = torch.linspace(0.0, 3899.0, 70, device=env.device)
    sy = torch.linspace(-799.0, 799.0, 70, device=env.device)
    env.set_field("rudder_angle", rud); env.set_field("s_x", sx); env.set_field("s_y", sy)
    env.set_field("index", torch.arange(70, dtype=torch.int32, device=env.device) * 100)
    assert torch.allclose(env.get_field("rudder_angle"), rud, atol=1e-9, rtol=0)      # 2^-42 rad resolution, then fp32
    assert torch.allclose(env.get_field("s_x"), sx, atol=2e-6, rtol=0) and torch.allclose(env.get_field("s_y"), sy, atol=5e-7, rtol=0)
    assert torch.equal(env.get_field("index"), torch.arange(70, dtype=torch.int32, device=env.device) * 100)
    env.set_field("rudder_angle", rud * 0.5)                                         # must not disturb the index bits
    assert torch.equal(env.get_field("index"), torch.arange(70, dtype=torch.int32, device=env.device) * 100)
    out = (C.c_double * 10)()
    assert env._L.boatenv_env_state_host(env._h, 69, out) == 0
    assert abs(out[3] - 0.6) < 1e-7 and abs(out[4] - 3899.0) < 2e-6 and abs(out[5] - 799.0) < 5e-7 and out[8] == 6900
    # thresholds: rudder just below / above pi/3 (boat_env.py:102), s_x just below / at the goal line (:85)
    import math
    env.set_field("rudder_angle", torch.full((70,), math.pi / 3 - 0.05, device=env.device))
    env.set_field("s_x", torch.full((70,), 100.0, device=env.device)); env.set_field("s_y", torch.zeros(70, device=env.device))
    a = torch.zeros(70, device=env.device); a[1] = 0.4999; a[2] = 0.5001
    _, _, done, info = env.step(a)
    assert info["term"][:3].tolist() == [0, 0, 5] and done[:3].tolist() == [0, 0, 1]
    env.close()


def _wind_coefficients(env):
    """[n_envs, 2, 4] wind piece coefficients of every env, cut out of the state blob (layout: common.cuh)."""
    import torch
    fp32 = env.precision == 32
    nd_bytes = 32 * 16 * (2 if fp32 else 4)
    w_bytes = 32 * 16 * (1 if fp32 else 2)
    off_wa, off_wb = nd_bytes + 128, nd_bytes + 128 + w_bytes
    block = off_wb + w_bytes + 128
    nblk = (env.n_envs + 31) // 32
    blob = env.state_dict()["blob"][: nblk * block].reshape(nblk, block)
    out = []
    for off in (off_wa, off_wb):
        sec = blob[:, off:off + w_bytes].contiguous()
        if fp32:
            c = sec.view(torch.float32).reshape(nblk, 32, 4)
        else:   # two 16-byte vector rows per lane: scalars (0, 1) in row 0, (2, 3) in row 1
            c = sec.view(torch.float64).reshape(nblk, 2, 32, 2).permute(0, 2, 1, 3).reshape(nblk, 32, 4)
        out.append(c.reshape(nblk * 32, 4)[: env.n_envs])
    return torch.stack(out, dim=1)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_lane_parallel_and_cooperative_wind_setup_agree_bit_for_bit(S, precision):
    """The reset kernel and the K > 1 episode-end queue kernel set the wind up one env per LANE, the K = 1 step kernel
    one env per WARP (setup warps): same knots -> bit-identical piece coefficients, whatever path started the
    episode.  (Every episode of an env is given the same knots through the validation override.)"""
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    n = 5000
    rng = np.random.default_rng(3)
    knots = (rng.integers(1, 2 ** 23, size=(n, 2, 8)) * 2 + 1) / 2.0 ** 24        # like knot_from_word: (2k+1) 2^-24
    envs = {}
    for name in ("reset", "k1", "k8"):
        e = S.BatchedBoatEnv(cfg, n, seed=4, precision=precision, device=0, auto_reset=True)
        e.set_episode_draws(None, knots)
        e.reset()
        envs[name] = e
    base = _wind_coefficients(envs["reset"])                                      # lane-parallel (reset kernel)
    assert torch.isfinite(base).all() and float(base.abs().max()) > 0
    big = envs["k1"].uniform_actions(0, 40.0).clone()                              # |a| / 10 > pi/3 for most envs: rudder breaks at once
    _, _, d1, _ = envs["k1"].step(big)                                             # in-kernel reset: cooperative (setup warps)
    _, _, d8, _ = envs["k8"].step_k(torch.stack([big, big, big]), 3)               # deferred queue: lane-parallel
    assert int(d1.sum()) > n // 2 and int(d8.sum()) >= int(d1.sum())
    c1, c8 = _wind_coefficients(envs["k1"]), _wind_coefficients(envs["k8"])
    assert torch.equal(c1, base) and torch.equal(c8, base)
    # and they are the coefficients of the oracle's curve: wind[0] = c0 of piece 0
    wv, wa = envs["reset"].wind_table(17)
    assert abs(float(base[17, 0, 0]) - wv[0]) <= (1e-6 if precision == "fp32" else 1e-13)
    assert abs(float(base[17, 1, 0]) - wa[0]) <= (1e-5 if precision == "fp32" else 1e-12)
    for e in envs.values():
        e.close()


def test_done_can_be_derived_from_the_termination_codes(S):
    """boatenv_step with done_out = NULL: only term_out is written, done == (term != 0); everything else is identical."""
    import torch
    cfg = S.load_config(base_settings__experiment=6)
    a = S.BatchedBoatEnv(cfg, 20_000, seed=6, precision="fp32", device=0, auto_reset=True)
    b = S.BatchedBoatEnv(cfg, 20_000, seed=6, precision="fp32", device=0, auto_reset=True)
    a.reset(); b.reset()
    b.done.fill_(77)
    n_done = 0
    for t in range(30):
        acts = a.uniform_actions(t, 4.0)
        o1, r1, d1, i1 = a.step(acts)
        o2, r2, d2, i2 = b.step(acts, done_from_term=True)
        assert d2 is None and torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(i1["term"], i2["term"])
        assert torch.equal(d1, (i2["term"] != 0).to(torch.uint8))
        n_done += int(d1.sum())
    assert n_done > 0 and bool((b.done == 77).all())      # the done buffer was never written
    L = a._L
    assert L.boatenv_step(a._h, acts.data_ptr(), a.obs.data_ptr(), a.reward.data_ptr(), None, None, None, 1, None) == -1
    a.close(); b.close()


def test_abi_calls_leave_the_callers_device_alone(S):
    """Every handle-based entry point switches to the handle's device and restores the caller's.  With one
    visible GPU the handle lives on device 0 and the check is that nothing changes; with more, the handle lives on
    device 1 while torch's current device stays 0 through create / reset / step / counters / replay / toy calls."""
    import torch
    dev = 1 if torch.cuda.device_count() > 1 else 0
    torch.cuda.set_device(0)
    cfg = S.load_config(base_settings__experiment=6)
    env = S.BatchedBoatEnv(cfg, 5000, seed=1, precision="fp32", device=dev, auto_reset=True)
    assert torch.cuda.current_device() == 0
    with torch.cuda.device(dev):            # tensors and streams of the handle's device for the launches themselves
        env.reset()
        acts = env.uniform_actions(0, 2.0)
        env.step(acts)
        torch.cuda.synchronize()
    assert torch.cuda.current_device() == 0
    c = env.counters()                      # called with device 0 current: reduce + copy run on the handle's device
    assert torch.cuda.current_device() == 0 and c["episodes"] >= 0
    ring = S.ReplayBuffer(10_000, (11,), 1, precision="fp32", device=dev, as_torch=True)
    toy = S.ToyCar(n_envs=64, precision="fp32", device=dev)
    assert torch.cuda.current_device() == 0
    with torch.cuda.device(dev):
        ring.step_store(env, acts)
        toy.step(3)
        torch.cuda.synchronize()
    assert torch.cuda.current_device() == 0 and ring.mem_cntr == 5000
    env.close(); ring.close(); toy.close()
    assert torch.cuda.current_device() == 0
