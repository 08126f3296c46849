"""The CPU oracle (oracle/boat_oracle.c) against the committed golden vectors:
(1) the reference's recorded fixtures, (2) roll-outs of the unmodified reference,
(3) the toy scripts' known answers.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from boat_testlib import load_golden, scaled_err

ROLLOUTS = [f"ref_rollout_exp{e}" for e in range(1, 7)] + [
    "ref_term_rudder_exp6", "ref_term_oob_exp2", "ref_term_fuel_exp3", "ref_term_timeout_exp4",
    "ref_term_goal_exp5"]


def fixture_wind_tables(g, params):
    """Rebuild the fixture's wind tables from the condensed description."""
    L = O.lib().oracle_wind_length(params)
    meta = g["meta"]
    wv = np.full(L, meta["const_v"])
    wa = np.full(L, meta["const_a"])
    if meta["kind_v"] == "curve":
        wv = O.random_curve(g["knots_v"], L) * params.max_velocity
    if meta["kind_a"] == "curve":
        wa = O.random_curve(g["knots_a"], L) * np.pi * 2
    if meta["kind_a"] == "rect":
        # knots recovered from the recorded switch pattern (make_golden.rect_knots_from_switches):
        # the oracle's own wind.py:53-58 path must reproduce the recorded table from them
        exp5 = O.OracleEnv(params)
        exp5.reset(0, g["knots_r"], np.zeros(len(g["knots_r"])))
        wa_from_knots = exp5.wind()[1]
        lo, hi = np.pi / 2, np.pi + np.pi / 2
        cur = float(g["angle0"])
        edges = [0] + list(g["angle_switches"]) + [L]
        for a, b in zip(edges[:-1], edges[1:]):
            wa[a:b] = cur
            cur = hi if abs(cur - lo) < 1e-9 else lo
        assert np.array_equal(wa_from_knots, wa)
    return wv, wa


@pytest.mark.parametrize("n", range(1, 7))
def test_recorded_fixture_replay(n):
    """6 known-answer tests: ressources/settings_visualized/experiment_setting_N
    (test_mode 1 => trajectory is a deterministic function of wind.csv)."""
    g = load_golden(f"fixture_exp{n}")
    p = O.params_from_config(g["config"])
    wv, wa = fixture_wind_tables(g, p)
    # the spline restatement reproduces the recorded wind.csv samples
    assert np.abs(wv[g["wind_idx"]] - g["wind_v"]).max() < 5e-15
    assert np.abs(wa[g["wind_idx"]] - g["wind_a"]).max() < 5e-14
    env = O.OracleEnv(p)
    env.reset(int(g["s_y_start"]), wind_tables=(wv, wa))
    rows = [env.all_data()]
    done, steps, code = False, 0, 0
    while not done:
        _, _, done, code = env.step(0.0)
        steps += 1
        if not done:
            rows.append(env.all_data())
    # row k = state after k steps; the terminal step is never written (main.py:79-81)
    assert steps == int(g["n_rows"])
    assert O.TERM_NAMES[code] == str(g["termination"]) == "reached_goal"
    rows = np.array(rows)[g["row_idx"]]
    ref = g["rows"]
    for c, scale in zip(range(5), (3900.0, 800.0, 5.0, 2.0, 2 * np.pi)):
        assert scaled_err(rows[:, c], ref[:, c], scale).max() < 1e-12
    assert np.array_equal(rows[:, 7], ref[:, 7])  # rudder frozen
    # rewards: fixtures 1-5 were recorded when f_x == 0.1
    offset = 0.0 if n == 6 else 0.1
    mask = g["row_idx"] > 0
    assert np.abs(rows[mask, 6] + offset - ref[mask, 6]).max() < 1e-12
    if n == 6:
        assert env.info["episode_reward"] == pytest.approx(float(g["episode_reward"]), abs=1e-9)
        assert env.info["episode_reward"] == pytest.approx(871.2727580297085, abs=1e-9)


@pytest.mark.parametrize("name", ROLLOUTS)
def test_reference_rollouts(name):
    g = load_golden(name)
    p = O.params_from_config(g["config"])
    T, N = g["actions"].shape
    out = O.rollout(p, g["actions"].astype(np.float64), g["s_y_start"][None], g["knots"][None])
    assert np.array_equal(out["done"], g["done"])
    assert np.array_equal(out["term"], g["term"])
    # normalised observations: scale 1 (they ARE the reference's normalised units)
    # compare while the trajectory is in the bounded regime
    err = scaled_err(out["obs"], g["obs"], 1.0)
    assert err.max() < 1e-10, err.max()
    assert scaled_err(out["reward"], g["reward"], 1.0).max() < 1e-10
    assert scaled_err(out["final"][:, 7], g["episode_reward"], 1.0).max() < 1e-9


@pytest.mark.parametrize("name", ROLLOUTS)
def test_wind_tables_match_reference(name):
    g = load_golden(name)
    p = O.params_from_config(g["config"])
    for i in range(len(g["s_y_start"])):
        env = O.OracleEnv(p)
        obs0 = env.reset(int(g["s_y_start"][i]), g["knots"][i, 0], g["knots"][i, 1])
        wv, wa = env.wind()
        assert np.abs(wv[g["wind_idx"]] - g["wind_v"][i]).max() < 2e-14
        assert np.abs(wa[g["wind_idx"]] - g["wind_a"][i]).max() < 1e-13
        assert np.abs(obs0 - g["obs0"][i]).max() < 1e-15


def test_wind_errors():
    """wind.py:65-67 and :73-75: ValueError for unknown experiment / < 4 knots."""
    g = load_golden("ref_rollout_exp6")
    cfg = g["config"]
    cfg["base_settings"]["experiment"] = 7
    env = O.OracleEnv(O.params_from_config(cfg))
    with pytest.raises(ValueError):
        env.reset(0, np.zeros(8), np.zeros(8))
    with pytest.raises(ValueError):
        O.random_curve(np.zeros(3), 100)


def test_curve_matches_scipy():
    from scipy.interpolate import interp1d
    rng = np.random.default_rng(3)
    for L in (10000, 240, 1000):
        for _ in range(20):
            u = rng.random(8)
            fixed = np.linspace(0, L, num=8)
            c = interp1d(fixed, u, kind="cubic", fill_value="extrapolate")(
                np.linspace(0, L, num=L, endpoint=True))
            if np.any((c < 0) | (c > 1)):
                c = (c - c.min()) / (c.max() - c.min())
            assert np.abs(O.random_curve(u, L) - c).max() < 1e-13


def test_toy_car_known_answers():
    """SURVEY.md 8(a) golden numbers, from the unmodified toy_car.py."""
    n = O.loop_count(500, 0.1)
    traj, out = O.toy_car()
    assert n == 5000
    assert out[0] == pytest.approx(-35.4716861557275, abs=1e-11)
    assert out[1] == pytest.approx(3.4236034791721615, abs=1e-11)
    # Scope drops the first sample (control_blocks.py:55-57): first recorded = iteration 2
    assert traj[1, 0] == pytest.approx(0.09998000066665778, abs=1e-15)
    assert traj[1, 1] == pytest.approx(0.0019998666693333083, abs=1e-15)


def test_toy_parachute_known_answers():
    traj, sv, calls = O.toy_parachute()
    assert calls == 2654
    assert traj[1, 0] == pytest.approx(2999.9019, abs=1e-9)
    assert traj[-2, 0] == pytest.approx(0.5671644323787001, abs=1e-10)
    assert sv[0] == pytest.approx(-0.0867586400200222, abs=1e-10)
    assert sv[1] == pytest.approx(-6.539230723987222, abs=1e-10)
