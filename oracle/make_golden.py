"""Generate the committed golden fixtures under tests/golden/.

Runs ONLY in the build container, where the reference is mounted read-only at
/root/reference: it (1) condenses the reference's recorded artefacts
(ressources/settings_visualized/experiment_setting_{1..6}) and (2) runs the
UNMODIFIED reference classes (imported through oracle/ref_shim.py) on seeded
inputs, and stores inputs + outputs as small .npz files.  The GPU box has no
/root/reference; tests there read only the .npz files.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
HOT_KEYS = {
    "base_settings": ["test_mode", "dt", "t_max", "experiment"],
    "boat_env": ["track_width", "boat_out_of_bounds_offset", "goal_line"],
    "boat": ["fuel", "boat_m", "boat_m_x", "boat_m_y", "boat_I", "boat_Iz", "propeller_diameter",
             "wake_friction", "c_r_front", "c_r_side", "thrust_deduction", "rho",
             "boat_area_front", "boat_area_side", "boat_l", "boat_b", "rudder_area"],
    "wind": ["fixed_points", "max_velocity", "direction"],
}
TERM_CODE = {"": 0, "reached_goal": 1, "out_of_bounds": 2, "out_of_fuel": 3, "timeout": 4,
             "rudder_broken": 5}


def hot_config(cfg) -> dict:
    return {sec: {k: cfg[sec][k] for k in keys} for sec, keys in HOT_KEYS.items()}


def knot_values(rng, n):
    """Knot values of the form the CUDA path draws: (2*k+1) * 2**-24, k 23-bit."""
    return (2.0 * rng.integers(0, 2 ** 23, size=n) + 1.0) * 2.0 ** -24


def cardinal_basis(L: int, n_knots: int) -> np.ndarray:
    """[L, n_knots] matrix W with curve = W @ knots, built from scipy's own
    interp1d(kind='cubic') exactly as wind.py:76-85 calls it."""
    from scipy.interpolate import interp1d
    fixed = np.linspace(0, L, num=n_knots)
    rng_ = np.linspace(0, L, num=L, endpoint=True)
    W = np.empty((L, n_knots))
    for k in range(n_knots):
        e = np.zeros(n_knots)
        e[k] = 1.0
        W[:, k] = interp1d(fixed, e, kind="cubic", fill_value="extrapolate")(rng_)
    return W


def rect_knots_from_switches(wa, W):
    """Fixture 5 records only the rectified direction (pi/2 <-> 3pi/2, wind.py:92-99), which pins
    the sign of (curve - 0.25) at every sample but not the knots.  Any knot vector with that sign
    pattern (and a curve inside [0,1], so that wind.py:87 does not renormalise) reproduces the
    recorded wind table exactly; pick the one with the largest margin (a small LP)."""
    from scipy.optimize import linprog
    L, k = W.shape
    s = np.where(wa > np.pi, 1.0, -1.0)
    A = np.vstack([np.hstack([-(s[:, None] * W), np.ones((L, 1))]),
                   np.hstack([W, np.zeros((L, 1))]), np.hstack([-W, np.zeros((L, 1))])])
    b = np.concatenate([-0.25 * s, np.full(L, 0.98), np.full(L, -0.02)])
    res = linprog(c=[0] * k + [-1], A_ub=A, b_ub=b, bounds=[(0.01, 0.99)] * k + [(0, 0.2)], method="highs")
    assert res.status == 0 and res.x[k] > 1e-6, res.message
    u = res.x[:k]
    assert np.array_equal(W @ u > 0.25, wa > np.pi)
    return u, float(res.x[k])


# --------------------------------------------------------------------------
# 1. the reference's recorded fixtures
# --------------------------------------------------------------------------
def condense_fixture(n: int):
    import pandas as pd
    d = R.fixture_dir(n)
    cfg = R.load_config(os.path.join(d, "configs", "tuned_configs.yaml"))
    data = pd.read_csv(os.path.join(d, "episodes", "episode_0_data.csv"), sep=";")
    wind = pd.read_csv(os.path.join(d, "episodes", "wind.csv"), sep=";")
    info = pd.read_csv(os.path.join(d, "episodes", "info.csv"), sep=";")
    L = len(wind)
    wv, wa = wind.wind_velocity.values.astype(float), wind.wind_angle.values.astype(float)
    W = cardinal_basis(L, int(cfg.wind.fixed_points))
    rec = dict(kind_v="const", kind_a="const", const_v=float(wv[0]), const_a=float(wa[0]))
    knots_v = knots_a = knots_r = np.zeros(0)
    switches = np.zeros(0, dtype=np.int64)
    if n in (4, 6):  # velocity = curve * max_velocity   (wind.py:49,62)
        knots_v, *_ = np.linalg.lstsq(W, wv / float(cfg.wind.max_velocity), rcond=None)
        rec["kind_v"] = "curve"
        rec["resid_v"] = float(np.abs(W @ knots_v * float(cfg.wind.max_velocity) - wv).max())
    if n == 6:       # angle = curve * pi * 2             (wind.py:63)
        knots_a, *_ = np.linalg.lstsq(W, wa / (2 * np.pi), rcond=None)
        rec["kind_a"] = "curve"
        rec["resid_a"] = float(np.abs(W @ knots_a * np.pi * 2 - wa).max())
    if n == 5:       # angle in {pi/2, 3pi/2}, a few switches (wind.py:57-58,92-99)
        switches = np.flatnonzero(np.diff(wa) != 0) + 1
        rec["kind_a"] = "rect"
        rec["rect_levels"] = sorted(set(np.round(wa, 12).tolist()))
        knots_r, rec["rect_margin"] = rect_knots_from_switches(wa, W)
    rows = np.unique(np.concatenate([np.arange(0, 8), np.arange(0, len(data), 37),
                                     np.arange(len(data) - 8, len(data))]))
    cols = ["boat_position_x", "boat_position_y", "boat_velocity_x", "boat_velocity_y",
            "boat_angle", "action_rudder", "reward", "rudder_angle", "n"]
    wsub = np.unique(np.concatenate([np.arange(0, L, 53), [L - 2, L - 1]]))
    np.savez_compressed(
        os.path.join(OUT, f"fixture_exp{n}.npz"),
        config=json.dumps(hot_config(cfg)), meta=json.dumps(rec),
        n_rows=len(data), termination=str(info.termination[0]),
        episode_reward=float(info.episode_reward[0]),
        s_y_start=int(round(float(data.boat_position_y[0]))),
        knots_v=knots_v, knots_a=knots_a, knots_r=knots_r, angle_switches=switches, angle0=float(wa[0]),
        wind_idx=wsub, wind_v=wv[wsub], wind_a=wa[wsub],
        row_idx=rows, rows=data[cols].values[rows].astype(float), columns=json.dumps(cols))
    print(f"fixture {n}: rows={len(data)} L={L} {rec}")


# --------------------------------------------------------------------------
# 2. live reference roll-outs with injected episode randomness
# --------------------------------------------------------------------------
def run_reference(cfg, s_y_start, knots, actions):
    """One env, len(actions) steps (keeps stepping after done, like the reference
    object would).  Returns obs [T,11], reward [T], done [T], term [T], all_data [T,8],
    wind tables."""
    curves = [knots[0], knots[1]]
    with R.KnotInjector(s_y_start, curves):
        env = R.make_env(cfg)  # BoatEnv.__init__ builds a Boat (draws once)...
    with R.KnotInjector(s_y_start, curves):
        obs0 = env.reset()     # ...and reset() builds the one we step
    T = len(actions)
    obs = np.empty((T, 11))
    rew = np.empty(T)
    done = np.zeros(T, dtype=np.uint8)
    term = np.zeros(T, dtype=np.uint8)
    alld = np.empty((T, 8))
    before = {k: env.info[k] for k in TERM_CODE if k}
    for t in range(T):
        o, r, d, info = env.step(np.array([np.float64(actions[t])]))  # float64 actions, H4
        obs[t], rew[t], done[t] = o, r, d
        if d:
            for k in before:
                if info[k] != before[k]:
                    term[t] = TERM_CODE[k]
                    before[k] = info[k]
        a = env.return_all_data()
        alld[t] = [a["boat_position_x"], a["boat_position_y"], a["boat_velocity_x"],
                   a["boat_velocity_y"], a["boat_angle"], a["action_rudder"], a["reward"],
                   a["rudder_angle"]]
    return dict(obs0=obs0, obs=obs, reward=rew, done=done, term=term, all_data=alld,
                wind_v=np.asarray(env.boat.wind.wind_velocity, dtype=float),
                wind_a=np.asarray(env.boat.wind.wind_angle, dtype=float),
                episode_reward=float(env.info["episode_reward"]))


def rollout_case(name, overrides, n_envs, T, policy, seed):
    cfg = R.load_config(**overrides)
    fp = int(cfg.wind.fixed_points)
    rng = np.random.default_rng(seed)
    half = int(cfg.boat_env.track_width * 0.8)
    s_y = rng.integers(-half, half, size=n_envs)
    knots = knot_values(rng, n_envs * 2 * fp).reshape(n_envs, 2, fp)
    if policy == "uniform":      # A1 of SURVEY 8(d): ends by rudder_broken in ~10^2-10^3 steps
        actions = rng.uniform(-1, 1, size=(T, n_envs)).astype(np.float32)
    elif policy == "small":      # A2: long episodes
        actions = (0.05 * rng.uniform(-1, 1, size=(T, n_envs))).astype(np.float32)
    elif policy == "hard_port":  # steer one way: out_of_bounds / heading penalty
        sgn = np.where(np.arange(n_envs) % 2 == 0, 1.0, -1.0)
        actions = np.zeros((T, n_envs), dtype=np.float32)
        actions[:45] = (0.2 * sgn).astype(np.float32)   # rudder -> +-0.9 rad (> pi/4, < pi/3)
    elif policy == "zero":
        actions = np.zeros((T, n_envs), dtype=np.float32)
    else:
        raise ValueError(policy)
    outs = [run_reference(cfg, int(s_y[i]), knots[i], actions[:, i]) for i in range(n_envs)]
    L = len(outs[0]["wind_v"])
    wsub = np.unique(np.concatenate([np.arange(0, L, 41), [L - 2, L - 1]]))
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"),
        config=json.dumps(hot_config(cfg)), policy=policy,
        s_y_start=s_y.astype(np.int32), knots=knots, actions=actions,
        obs0=np.stack([o["obs0"] for o in outs]),
        obs=np.stack([o["obs"] for o in outs], axis=1),
        reward=np.stack([o["reward"] for o in outs], axis=1),
        done=np.stack([o["done"] for o in outs], axis=1),
        term=np.stack([o["term"] for o in outs], axis=1),
        all_data=np.stack([o["all_data"] for o in outs], axis=1),
        wind_idx=wsub,
        wind_v=np.stack([o["wind_v"][wsub] for o in outs]),
        wind_a=np.stack([o["wind_a"][wsub] for o in outs]),
        episode_reward=np.array([o["episode_reward"] for o in outs]))
    first = [int(np.argmax(o["done"])) + 1 if o["done"].any() else -1 for o in outs]
    kinds = sorted({int(t) for o in outs for t in o["term"] if t})
    print(f"{name}: T={T} n={n_envs} first_done={first} term_kinds={kinds}")


# --------------------------------------------------------------------------
# 3. replay buffer (agent/buffer.py)
# --------------------------------------------------------------------------
def replay_case():
    ref = R.import_reference()
    rng = np.random.default_rng(7)
    size, n_store, batch = 64, 150, 32          # wraps the ring twice
    buf = ref.ReplayBuffer(size, (11,), 1)
    S = rng.standard_normal((n_store, 11))
    S2 = rng.standard_normal((n_store, 11))
    A = rng.uniform(-1, 1, (n_store, 1))
    Rw = rng.standard_normal(n_store)
    D = rng.integers(0, 2, n_store).astype(bool)
    samples = []
    for i in range(n_store):
        buf.store_transition(S[i], A[i], Rw[i], S2[i], D[i])
        if i in (9, 63, 64, 100, 149):
            np.random.seed(1000 + i)
            max_mem = min(buf.mem_cntr, buf.mem_size)
            idx = np.random.choice(max_mem, batch)   # same draw as buffer.py:27
            np.random.seed(1000 + i)
            s, a, r, s2, d = buf.sample_buffer(batch)
            samples.append(dict(after=i + 1, idx=idx, s=s, a=a, r=r, s2=s2, d=d))
    np.savez_compressed(
        os.path.join(OUT, "replay_buffer.npz"), size=size, batch=batch,
        S=S, S2=S2, A=A, R=Rw, D=D,
        after=np.array([s["after"] for s in samples]),
        idx=np.stack([s["idx"] for s in samples]),
        s=np.stack([s["s"] for s in samples]), a=np.stack([s["a"] for s in samples]),
        r=np.stack([s["r"] for s in samples]), s2=np.stack([s["s2"] for s in samples]),
        d=np.stack([s["d"] for s in samples]),
        final_state=buf.state_memory, final_new_state=buf.new_state_memory,
        final_action=buf.action_memory, final_reward=buf.reward_memory,
        final_terminal=buf.terminal_memory, mem_cntr=buf.mem_cntr)
    print("replay_buffer: ok")


def main():
    os.makedirs(OUT, exist_ok=True)
    for n in range(1, 7):
        condense_fixture(n)
    for ex in range(1, 7):
        # long episodes, small steering noise: 1000 steps, no termination expected
        rollout_case(f"ref_rollout_exp{ex}", dict(base_settings__experiment=ex), 4, 1000,
                     "small", seed=100 + ex)
    # uniform(-1,1) policy, exp 6: rudder_broken, then keeps stepping after done
    rollout_case("ref_term_rudder_exp6", dict(base_settings__experiment=6), 6, 700, "uniform",
                 seed=21)
    # out_of_bounds: exp 2 (random start) + steering hard to one side
    rollout_case("ref_term_oob_exp2", dict(base_settings__experiment=2), 6, 1200, "hard_port",
                 seed=22)
    # out_of_fuel: tiny tank
    rollout_case("ref_term_fuel_exp3", dict(base_settings__experiment=3, boat__fuel=120), 3, 200,
                 "small", seed=23)
    # timeout: t_max = 60 s -> L = 240 wind samples
    rollout_case("ref_term_timeout_exp4", dict(base_settings__experiment=4,
                                               base_settings__t_max=60), 3, 240, "small", seed=24)
    # goal: move the goal line close
    rollout_case("ref_term_goal_exp5", dict(base_settings__experiment=5, boat_env__goal_line=300),
                 3, 500, "small", seed=25)
    replay_case()


if __name__ == "__main__":
    main()
