"""The reference's OWN callers drive the drop-in.

The unmodified reference files (staged by oracle/make_ref.py into oracle/_ref, or /root/reference in the build
container) are imported under the stub modules of oracle/ref_shim.py, and the loop body of main.py:70-99 --
``recorder.write_data_to_csv(); action = agent.choose_action(obs); obs_, r, done, info = env.step(action);
agent.remember(...); agent.learn()`` -- is run twice with the reference's ``ContinuousAgent``
(agent/continuous_agent.py:9) and ``Recorder`` (postprocessing/recorder.py:7):

  A. on the reference ``BoatEnv`` + the reference ``ReplayBuffer`` (np.random draws injected), and
  B. on ``sac_agent_b200.BoatEnv`` + ``sac_agent_b200.ReplayBuffer`` swapped in by the two imports INTEGRATION.md
     names (main.py:3 ``from environment.boat_env import BoatEnv``; agent/continuous_agent.py:4
     ``from agent.buffer import ReplayBuffer``), with the same episode draws.

Both runs start from the same torch seed, so they pick the same actions as long as their observations agree;
learning starts after ``batch_size`` stored transitions, from batches whose indices run B recorded and run A replays
(the two buffers draw from different generators).  Compared: the trajectory step by step, the buffers' contents,
the info dicts and the three CSV files the Recorder writes.
"""
import os
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STEPS = 360          # env steps of the loop ("a few hundred")
BATCH = 64           # config.agent.batch_size for this test: learn() is active from step 64 on


@pytest.fixture(scope="module")
def S():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sac_agent_b200 as pkg
    pkg.lib()
    return pkg


@pytest.fixture(scope="module")
def R():
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("the reference is neither mounted at /root/reference nor staged under oracle/_ref "
                    "(python oracle/make_ref.py, run by __graft_entry__.build() in the build container)")
    ref_shim.import_reference()
    return ref_shim


def experiment_dir(tmp_path, name):
    d = tmp_path / name
    for sub in ("episodes", "checkpoints", "configs", "plots", "rendering"):
        os.makedirs(d / sub)
    (d / "reward_field.png").write_bytes(b"")     # reward_functions.py:23-26: skip the plot
    return types.SimpleNamespace(experiment_dir=str(d))


def run_loop(env, agent, recorder, steps, action_dtype=None):
    """main.py:70-99 without the progress table: one list entry per env step.  action_dtype = np.float64 widens the
    agent's float32 action before env.step (with a float32 action the reference's own
    `rudder_angle += action[0] / 10` runs in float32, `0 + np.float32` stays float32)."""
    log = []
    episode = 0
    observation = env.reset()
    recorder.create_csvs(episode)
    done = False
    for _ in range(steps):
        if done:                                           # main.py:70-75: next game
            recorder.write_info_to_csv()
            recorder.write_winds_to_csv()
            episode += 1
            observation = env.reset()
            recorder.create_csvs(episode)
        recorder.write_data_to_csv()
        action = agent.choose_action(observation)
        if action_dtype is not None:
            action = np.asarray(action, dtype=action_dtype)
        observation_, reward, done, info = env.step(action)
        agent.remember(observation, action, reward, observation_, info["termination"] == "reached_goal")
        agent.learn()
        log.append((np.asarray(observation_, dtype=np.float64).copy(), float(reward), bool(done),
                    np.asarray(action, dtype=np.float64).copy(), float(env.boat.rudder_angle)))
        observation = observation_
    recorder.write_info_to_csv()
    recorder.write_winds_to_csv()
    return log, episode


@pytest.mark.parametrize("action_dtype,tol", [(np.float64, 1e-6), (None, 1e-3)])
def test_reference_agent_and_recorder_drive_the_drop_in(S, R, tmp_path, action_dtype, tol):
    """action_dtype None: the loop exactly as main.py runs it -- the agent's float32 actions make the REFERENCE env
    accumulate its rudder in float32 (H4), so it drifts ~1e-7 rad from the drop-in's fp64 rudder and the yaw
    acceleration follows at ~1e-5 after a few hundred steps; inside the pi/4 penalty zone the reward is
    -100 |rudder| (boat_env.py:107-108), i.e. 100 x the rudder difference: tolerance 1e-3.  np.float64: the same loop with the action widened first (the
    parity convention of SURVEY.md 8c): both envs compute in fp64 and agree to the float32 policy's rounding."""
    import pandas as pd
    import torch
    cfg = R.load_config(base_settings__experiment=6, agent__batch_size=BATCH, agent__max_size=4096)
    Recorder = R.import_reference_recorder()
    fp = int(cfg.wind.fixed_points)

    # ---- run B: the drop-in classes behind the reference's agent and recorder ------------------------------
    sampled = []                       # indices of every learn() batch, replayed by run A

    class RecordingReplayBuffer(S.ReplayBuffer):     # ReplayBuffer(max_size, input_shape, n_actions): the reference's signature
        def sample_buffer(self, batch_size):
            *out, idx = super().sample_buffer(batch_size, return_indices=True)
            sampled.append(idx.cpu().numpy().copy())
            return tuple(out)

    AgentB = R.import_reference_agent(RecordingReplayBuffer)
    exp_b = experiment_dir(tmp_path, "drop_in")
    env_b = S.BoatEnv(cfg, exp_b, seed=11, precision="fp64", device=0)
    draws = [env_b._b.episode_draws(0, e) for e in range(1, 40)]   # reset() number e starts episode e of env 0
    torch.manual_seed(1234)
    agent_b = AgentB(config=cfg, experiment_dir=exp_b.experiment_dir, input_dims=env_b.observation_space.shape, env=env_b)
    assert type(agent_b.memory).__mro__[1] is S.ReplayBuffer
    log_b, episodes_b = run_loop(env_b, agent_b, Recorder(env_b), STEPS, action_dtype)

    # ---- run A: the reference env and buffer, same draws, same torch seed, same batch indices ----------------
    AgentA = R.import_reference_agent()
    exp_a = experiment_dir(tmp_path, "reference")
    ref = R.import_reference()
    # BoatEnv.__init__ builds a Boat (boat_env.py:15) and reset() builds another: each consumes one randint and
    # two sample(fixed_points) draws.  The drop-in's constructor episode is number 0, reset() number e is episode e.
    s0, k0 = env_b._b.episode_draws(0, 0)
    order = [(s0, k0)] + draws
    cursor = {"i": 0, "phase": 0}
    real_randint, real_sample, real_choice = np.random.randint, np.random.sample, np.random.choice

    def randint(lo, hi=None, *a, **k):
        cursor["phase"] = 0
        return order[cursor["i"]][0]

    def sample(n):
        assert n == fp
        i, ph = cursor["i"], cursor["phase"]
        cursor["phase"] += 1
        if ph == 1:
            cursor["i"] += 1
        return order[i][1][ph].copy()

    replay = iter(sampled)

    def choice(a, size=None, *args, **kw):
        idx = next(replay)
        assert len(idx) == size and idx.max() < a
        return idx

    np.random.randint, np.random.sample, np.random.choice = randint, sample, choice
    try:
        env_a = ref.BoatEnv(cfg, exp_a)
        torch.manual_seed(1234)
        agent_a = AgentA(config=cfg, experiment_dir=exp_a.experiment_dir, input_dims=env_a.observation_space.shape, env=env_a)
        log_a, episodes_a = run_loop(env_a, agent_a, Recorder(env_a), STEPS, action_dtype)
    finally:
        np.random.randint, np.random.sample, np.random.choice = real_randint, real_sample, real_choice

    # ---- the two runs agree ------------------------------------------------------------------------------------
    assert episodes_a == episodes_b
    assert len(sampled) == STEPS - BATCH + 1 and next(replay, None) is None     # learn() ran from step 64 on, in both
    for t, (a, b) in enumerate(zip(log_a, log_b)):
        assert a[2] == b[2], f"done differs at step {t}"
        assert np.abs(a[3] - b[3]).max() <= tol, f"actions differ at step {t}"        # float32 policy outputs
        assert np.abs(a[0] - b[0]).max() <= tol and abs(a[1] - b[1]) <= tol * max(1.0, abs(a[1])), t
        assert abs(a[4] - b[4]) <= tol                                                    # env.boat.rudder_angle
    # before learning starts the actions are bit-identical, and so are the fp64 observations to 1e-9
    if action_dtype is not None:
        for a, b in zip(log_a[:BATCH - 1], log_b[:BATCH - 1]):
            assert np.array_equal(a[3], b[3]) and np.abs(a[0] - b[0]).max() <= 1e-9
    assert env_a.info["termination"] == env_b.info["termination"]
    for k in ("reached_goal", "out_of_bounds", "out_of_fuel", "rudder_broken", "timeout"):
        assert env_a.info[k] == env_b.info[k]
    # replay buffers: same counters, same rows (agent/buffer.py:13-22)
    ma, mb = agent_a.memory, agent_b.memory
    assert ma.mem_cntr == mb.mem_cntr == STEPS and ma.mem_size == mb.mem_size == 4096
    s, a_, r, s2, d = S.ReplayBuffer.gather(mb, np.arange(STEPS), as_torch=False)
    assert np.abs(s - ma.state_memory[:STEPS]).max() <= tol and np.abs(s2 - ma.new_state_memory[:STEPS]).max() <= tol
    assert np.abs(a_ - ma.action_memory[:STEPS]).max() <= tol and np.array_equal(d, ma.terminal_memory[:STEPS])
    assert np.abs(r - ma.reward_memory[:STEPS]).max() <= tol * np.maximum(1.0, np.abs(ma.reward_memory[:STEPS])).max()
    # the networks were trained identically (same batches, same code: the reference's learn()); with float32 actions
    # the 1e-6 differences of the inputs (H4, above) are amplified by ~300 Adam steps at lr 0.005: not compared
    for pa, pb in zip(agent_a.actor.parameters(), agent_b.actor.parameters()):
        if action_dtype is not None:
            assert torch.allclose(pa, pb, atol=1e-4, rtol=1e-3)
        else:
            assert torch.isfinite(pb).all()
    # ---- the Recorder's files (postprocessing/recorder.py:18-56), read back like replayer.py does ----------------
    for e in range(episodes_a + 1):
        fa = pd.read_csv(os.path.join(exp_a.experiment_dir, "episodes", f"episode_{e}_data.csv"), sep=";")
        fb = pd.read_csv(os.path.join(exp_b.experiment_dir, "episodes", f"episode_{e}_data.csv"), sep=";")
        assert list(fa.columns) == list(fb.columns) and len(fa) == len(fb)
        va, vb = fa.values.astype(np.float64), fb.values.astype(np.float64)                # raw metres / rad, not normalised
        assert np.all(np.abs(va - vb) <= 10 * tol * np.maximum(1.0, np.abs(va)))
    ia = pd.read_csv(os.path.join(exp_a.experiment_dir, "episodes", "info.csv"), sep=";")
    ib = pd.read_csv(os.path.join(exp_b.experiment_dir, "episodes", "info.csv"), sep=";")
    assert list(ia.columns) == list(ib.columns) and list(ia.termination) == list(ib.termination)
    assert np.abs(ia.episode_reward.values - ib.episode_reward.values).max() <= max(1e-4, tol) * np.abs(ia.episode_reward.values).max()
    wa = pd.read_csv(os.path.join(exp_a.experiment_dir, "episodes", "wind.csv"), sep=";")
    wb = pd.read_csv(os.path.join(exp_b.experiment_dir, "episodes", "wind.csv"), sep=";")
    assert list(wa.columns) == list(wb.columns) == ["wind_velocity", "wind_angle"] and len(wa) == len(wb) == 10000
    assert np.abs(wa.values - wb.values).max() <= 1e-12
    env_b.close()
