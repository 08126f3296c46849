"""SAC update parity: `SACLearner` / `ContinuousAgent` against the recorded run of
the reference's own `ContinuousAgent.learn()` (tests/golden/agent_update.npz, written by
oracle/make_agent_golden.py from agent/continuous_agent.py:96-154 + networks/networks.py).

Tolerances (fp32 learner against an fp32 reference on another BLAS): gradients 2e-4 of the tensor's
largest gradient, weights 1e-6 + 2 % of the learning rate -- the first Adam steps move every weight by
~lr * sign(g), so a weight whose gradient is within rounding of zero may differ by a fraction of lr;
at most 0.2 % of the sampled positions may exceed that (none do on the CPU).
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_agent_golden as G  # noqa: E402  (test infrastructure)
from oracle import ref_shim as R  # noqa: E402

import sac_agent_b200 as S  # noqa: E402
from sac_agent_b200.continuous_agent import ContinuousAgent, SACLearner  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "agent_update.npz")


def _learner(meta, device):
    L = SACLearner((11,), 1, np.array([1.0], dtype=np.float32), float(meta["meta/learning_rate_alpha"]),
                   float(meta["meta/learning_rate_beta"]), float(meta["meta/gamma"]),
                   float(meta["meta/tvn_parameter_modulation_tau"]), float(meta["meta/reward_scale"]),
                   device=device, torch_reference_math=(device == "cpu"))
    return L


def _load_init(L, seed):
    w0 = G.init_weights(seed)
    for n in G.NETS:
        getattr(L, n).load_state_dict({k: torch.from_numpy(v.copy()) for k, v in w0[n].items()})


def _compare(rec, L, u, lr_of):
    bad = []
    for n in G.NETS:
        net = getattr(L, n)
        for k, p in net.named_parameters():
            for kind in ("w", "g"):
                key = f"u{u}/{n}/{k}/{kind}"
                if key + "@" not in rec:
                    continue
                t = p if kind == "w" else p.grad
                ours = t.detach().float().cpu().numpy().reshape(-1)
                pos = G.sample_positions(tuple(p.shape), key)
                ref = rec[key + "@"]
                got = ours[pos]
                if kind == "g":
                    tol = 2e-4 * max(np.abs(ref).max(), 1e-12) + 1e-9
                else:
                    tol = 1e-6 + 0.02 * lr_of[n]
                frac = np.mean(np.abs(got - ref) > tol)
                if frac > 0.002:
                    bad.append((key, float(np.abs(got - ref).max()), tol, float(frac)))
                # the whole tensor through its recorded norm
                l2 = float(np.sqrt((ours.astype(np.float64) ** 2).sum()))
                assert abs(l2 - float(rec[key + "#l2"])) <= 1e-3 * max(float(rec[key + "#l2"]), 1e-9), key
    assert not bad, bad


def _run(rec, device, via_agent=False):
    seed, n_updates, batch = int(rec["meta/seed"]), int(rec["meta/n_updates"]), int(rec["meta/batch"])
    b, noise = G.make_batch(seed, batch), G.make_noise(seed, n_updates, batch)
    lr_of = {"actor": float(rec["meta/learning_rate_alpha"])}
    for n in ("critic_1", "critic_2", "value", "target_value"):
        lr_of[n] = float(rec["meta/learning_rate_beta"])
    if via_agent:
        cfg = S.load_config()
        env = type("E", (), {"action_space": S.Box(low=-1, high=1, dtype=np.float32)})()
        mem = S.ReplayBuffer(4096, (11,), 1, precision="fp32", device=0, as_torch=True)
        agent = ContinuousAgent(cfg, None, (11,), env, device=0, use_cuda_graph=True, memory=mem)
        L = agent.learner
    else:
        agent = None
        L = _learner(rec, device)
    _load_init(L, seed)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in b.items()}
    for u in range(1, n_updates + 1):
        if via_agent:
            agent.learn_from(t["state"], t["action"], t["reward"], t["new_state"], t["done"].to(torch.uint8),
                             eps=torch.from_numpy(noise[u - 1]))
        else:
            e = torch.from_numpy(noise[u - 1]).to(device)
            L.update(t["state"], t["action"], t["reward"], t["new_state"], t["done"], eps_sample=e[0], eps_rsample=e[1])
        if u in rec["meta/record_after"]:
            _compare(rec, L, u, lr_of)
    if via_agent:
        assert agent.updates == n_updates
        mem.close()


def test_state_dict_layout_matches_reference():
    rec = np.load(GOLDEN)
    L = _learner(rec, "cpu")
    ours = [f"{n}/{k}/{tuple(p.shape)}" for n in G.NETS for k, p in getattr(L, n).named_parameters()]
    assert ours == [str(x) for x in rec["meta/param_names"]]
    for p, q in zip(L.value.parameters(), L.target_value.parameters()):  # update_network_parameters(tau=1)
        assert torch.equal(p, q)


def test_learner_cpu_matches_recorded_reference_update():
    _run(np.load(GOLDEN), "cpu")


@pytest.mark.skipif(not R.reference_available(), reason="reference not mounted (GPU box)")
def test_learner_cpu_matches_live_reference_on_a_fresh_seed():
    rec = G.reference_learn(seed=11, n_updates=2, batch=256, record_after=(2,))
    _run(rec, "cpu")


def test_agent_needs_cuda():
    with pytest.raises(RuntimeError):   # the learner does not fall back to the CPU on its own
        SACLearner((11,), 1, np.array([1.0], dtype=np.float32), 5e-3, 3e-4, 0.99, 0.005, 10, device="cpu")
    if torch.cuda.is_available():
        return
    env = type("E", (), {"action_space": S.Box(low=-1, high=1, dtype=np.float32)})()
    with pytest.raises(RuntimeError):
        ContinuousAgent(S.load_config(), None, (11,), env)


@pytest.mark.gpu
def test_learner_gpu_matches_recorded_reference_update():
    _run(np.load(GOLDEN), "cuda")


@pytest.mark.gpu
def test_agent_cuda_graph_matches_recorded_reference_update():
    _run(np.load(GOLDEN), "cuda", via_agent=True)


@pytest.mark.gpu
def test_agent_api_on_batched_env():
    """choose_action / step_and_remember / learn on a BatchedBoatEnv, single-env numpy API, checkpoints."""
    import tempfile
    cfg = S.load_config(base_settings__experiment=6, agent__batch_size=256)
    env = S.BatchedBoatEnv(cfg, 2048, seed=3, precision="fp32", device=0, auto_reset=True)
    with tempfile.TemporaryDirectory() as tmp:
        agent = ContinuousAgent(cfg, tmp, env.observation_space.shape, env, device=0, seed=3)
        assert agent.get_n_actions() == 1 and float(agent.get_max_actions()[0]) == 1.0
        obs = env.reset()
        assert agent.learn() is None  # memory below one batch: continuous_agent.py:97-98
        for it in range(6):
            a = agent.choose_action_graphed(obs)
            assert a.shape == (2048, 1) and float(a.abs().max()) <= 1.0
            agent.step_and_remember(env, a.squeeze(-1))
            losses = agent.learn()
        assert agent.memory.mem_cntr == 6 * 2048 and agent.updates == 6
        assert all(torch.isfinite(x).item() for x in losses)
        a1 = agent.choose_action(np.zeros(11))  # the reference's single-observation call
        assert isinstance(a1, np.ndarray) and a1.shape == (1,)
        before = [p.detach().clone() for p in agent.actor.parameters()]
        agent.save_models()
        assert sorted(os.listdir(os.path.join(tmp, "checkpoints"))) == sorted(
            ["actor_network", "critic_network_1", "critic_network_2", "value_network", "target_value_network"])
        with torch.no_grad():
            for p in agent.actor.parameters():
                p.add_(1.0)
        agent.load_models()
        assert all(torch.equal(p, q) for p, q in zip(agent.actor.parameters(), before))
        agent.learn()  # the captured graph still works on the reloaded (same-storage) parameters
    env.close()
    agent.memory.close()


@pytest.mark.gpu
@pytest.mark.parametrize("use_graph", [False, True])
def test_whole_sac_loop_checkpoint_resume_is_exact(tmp_path, use_graph):
    """the reference checkpoints network weights only (base_network.py:13-17).  Here the env
    blob, the replay ring (rows, store counter, Philox sample counter), the Adam moments + step counter and the
    CUDA generator state are saved too: save -> keep running == fresh objects -> load -> run, bit for bit
    (observations, sampled batches, losses, weights)."""
    cfg = S.load_config(base_settings__experiment=6, agent__batch_size=256)

    def make():
        env = S.BatchedBoatEnv(cfg, 2048, seed=3, precision="fp32", device=0, auto_reset=True)
        mem = S.ReplayBuffer(30_000, (11,), 1, precision="fp32", device=0, seed=7, as_torch=True)   # wraps inside the run
        agent = ContinuousAgent(cfg, str(tmp_path), (11,), env, device=0, seed=3, memory=mem, use_cuda_graph=use_graph)
        return env, mem, agent

    def run(env, agent, iters):
        out = []
        obs = env.obs
        for _ in range(iters):
            a = agent.choose_action(obs)
            agent.step_and_remember(env, a.squeeze(-1))
            losses = agent.learn()
            out.append((env.obs.clone(), env.reward.clone(), torch.stack([x.detach().clone() for x in losses])))
        return out

    os.makedirs(tmp_path / "checkpoints", exist_ok=True)
    env, mem, agent = make()
    env.reset()
    run(env, agent, 25)                       # the ring wraps at iteration 15
    torch.save({"env": env.state_dict(), "agent": agent.state_dict()}, tmp_path / "full.pt")
    assert agent.save_training_state() == str(tmp_path / "checkpoints" / "training_state.pt")
    tail_a = run(env, agent, 12)
    weights_a = [p.detach().clone() for net in agent.learner.networks() for p in net.parameters()]
    cntr_a = mem.mem_cntr

    env2, mem2, agent2 = make()               # what a fresh process does
    ck = torch.load(tmp_path / "full.pt", weights_only=False)
    env2.load_state_dict(ck["env"])
    agent2.load_state_dict(ck["agent"])
    assert mem2.mem_cntr == 25 * 2048 and agent2.updates == 25
    tail_b = run(env2, agent2, 12)
    for (o1, r1, l1), (o2, r2, l2) in zip(tail_a, tail_b):
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(l1, l2)
    weights_b = [p.detach().clone() for net in agent2.learner.networks() for p in net.parameters()]
    assert all(torch.equal(p, q) for p, q in zip(weights_a, weights_b)) and mem2.mem_cntr == cntr_a
    idx = torch.arange(0, 30_000, 7, device="cuda")
    assert all(torch.equal(x, y) for x, y in zip(mem.gather(idx), mem2.gather(idx)))
    other = S.ReplayBuffer(4096, (11,), 1, precision="fp32", device=0, as_torch=True)
    with pytest.raises(ValueError):
        other.load_state_dict(ck["agent"]["memory"])
    for x in (env, env2, mem, mem2, other):
        x.close()


@pytest.mark.gpu
def test_load_models_refreshes_the_acting_copies(tmp_path):
    """load_models() in evaluation-only use (learn() never called): the packed bf16 blob of the tcgen05 policy and
    the acting copies of an OverlappedActorLearner follow the loaded weights."""
    cfg = S.load_config(base_settings__experiment=6, agent__batch_size=256)
    os.makedirs(tmp_path / "checkpoints", exist_ok=True)
    env = S.BatchedBoatEnv(cfg, 4096, seed=3, precision="fp32", device=0, auto_reset=True)
    mem = S.ReplayBuffer(30_000, (11,), 1, precision="fp32", device=0, as_torch=True)
    agent = ContinuousAgent(cfg, str(tmp_path), (11,), env, device=0, seed=3, memory=mem, policy_precision="tcgen05")
    obs = env.reset()
    agent.save_models()
    pipe = S.OverlappedActorLearner(agent, env)
    pol = agent.tensor_core_policy()
    eps = torch.randn(4096, 1, device="cuda")
    ref_act = pol.act(obs, eps).clone()
    with torch.no_grad():
        for p in agent.actor.parameters():
            p.add_(0.25)
    pol.refresh()
    assert not torch.equal(pol.act(obs, eps), ref_act)
    agent.load_models()                                   # weights back; no learn() afterwards
    assert torch.equal(pol.act(obs, eps), ref_act)
    torch.cuda.synchronize()
    for c in pipe.acting:
        assert all(torch.equal(p, q) for p, q in zip(c.parameters(), agent.actor.parameters()))
    env.close(); mem.close()


@pytest.mark.gpu
def test_overlapped_actor_learner():
    """Acting and learning on two streams: same env trajectory bookkeeping as the sequential loop (every
    step stores n transitions, one update per step once a batch exists), finite losses, and the acting
    copies hold exactly the actor weights the learner published (one or two updates old)."""
    cfg = S.load_config(base_settings__experiment=6, agent__batch_size=512)
    env = S.BatchedBoatEnv(cfg, 8192, seed=5, precision="fp32", device=0, auto_reset=True)
    mem = S.ReplayBuffer(100_000, (11,), 1, precision="fp32", device=0, as_torch=True)   # wraps inside the run
    agent = ContinuousAgent(cfg, None, (11,), env, device=0, seed=5, memory=mem)
    env.reset()
    pipe = S.OverlappedActorLearner(agent, env)
    snaps = []
    for it in range(40):
        losses = pipe.step()
        if it >= 37:
            pipe.finish()
            torch.cuda.synchronize()
            snaps.append([p.detach().clone() for p in agent.actor.parameters()])
    pipe.finish()
    torch.cuda.synchronize()
    assert agent.memory.mem_cntr == 40 * 8192 and agent.updates == 40
    assert all(torch.isfinite(x).item() for x in losses)
    # update(t) is published into acting[t & 1]: after step 39 acting[1] holds update 39, acting[0] update 38
    assert all(torch.equal(p, q) for p, q in zip(pipe.acting[1].parameters(), snaps[2]))
    assert all(torch.equal(p, q) for p, q in zip(pipe.acting[0].parameters(), snaps[1]))
    assert float(env.counters()["episodes"]) >= 0
    env.close(); mem.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n_actions", [1, 3])
def test_fused_gaussian_head_matches_the_torch_expression(n_actions):
    """csrc/agent_ops.cu (one kernel forward, one backward) against the element-wise PyTorch expression of
    networks.py:47-70 evaluated in float64 on the same inputs: values and both gradients."""
    from sac_agent_b200.networks import ActorNetwork
    torch.manual_seed(0)
    B = 1500
    mean = (torch.randn(B, n_actions, device="cuda") * 1.5)
    raw = torch.randn(B, n_actions, device="cuda") * 2.0
    eps = torch.randn(B, n_actions, device="cuda")
    ga, gl = torch.randn(B, n_actions, device="cuda"), torch.randn(B, 1, device="cuda")
    max_a = np.linspace(1.0, 0.5, n_actions).astype(np.float32)  # <= 1: the formula takes log(1 - action^2 + 1e-6)

    class Head(ActorNetwork):  # the two linear heads replaced by leaf tensors
        def forward(self, state):
            return self._m, self._r

    def run(dtype):
        net = Head(None, (11,), max_a, n_actions=n_actions).cuda().to(dtype)
        net._m = mean.to(dtype).clone().requires_grad_(True)
        net._r = raw.to(dtype).clone().requires_grad_(True)
        act, lp = net.sample_normal(None, reparameterize=True, eps=eps.to(dtype))
        (act * ga.to(dtype)).sum().add((lp * gl.to(dtype)).sum()).backward()
        with torch.no_grad():
            act0, lp0 = net.sample_normal(None, reparameterize=False, eps=eps.to(dtype))
        assert torch.equal(act0, act.detach()) and torch.equal(lp0, lp.detach())
        return [x.detach().double() for x in (act, lp, net._m.grad, net._r.grad)]

    fused, ref = run(torch.float32), run(torch.float64)
    # log(1 - action^2 + 1e-6) cancels catastrophically in float32 once tanh saturates (the reference has the
    # same property): rows with |u| > 3 in any action are compared on the action only
    sd = torch.exp(-5.0 + 3.5 * (torch.tanh(raw.double()) + 1.0))
    ok = ((mean.double() + eps.double() * sd).abs() <= 3.0).all(dim=1)
    assert ok.float().mean().item() > 0.4
    for name, a, b in zip(("action", "log_prob", "grad_mean", "grad_raw_std"), fused, ref):
        if name != "action":
            a, b = a[ok], b[ok]
        err = ((a - b).abs() / b.abs().clamp_min(1.0)).max().item()
        assert err < 2e-4, (name, err)


@pytest.mark.gpu
def test_device_adam_matches_torch_adam_and_polyak():
    """One libboatenv launch (Adam for every tensor + Polyak average of the marked ones) against
    torch.optim.Adam + the reference's tau * value + (1 - tau) * target, 25 steps, two learning rates."""
    from sac_agent_b200.continuous_agent import DeviceAdam
    torch.manual_seed(1)
    shapes = [(256, 11), (256,), (256, 256), (1, 256), (1,), (5000,)]
    ref = [torch.randn(s, device="cuda").requires_grad_(True) for s in shapes]
    ours = [p.detach().clone() for p in ref]
    tgt_ref = [torch.randn_like(p) for p in ref[:2]]
    tgt_ours = [t.clone() for t in tgt_ref]
    lrs = [5e-3, 5e-3, 3e-4, 3e-4, 3e-4, 3e-4]
    opt_ref = torch.optim.Adam([{"params": ref[:2], "lr": 5e-3}, {"params": ref[2:], "lr": 3e-4}])
    opt = DeviceAdam([(ours[:2], 5e-3), (ours[2:], 3e-4)], polyak=list(zip(ours[:2], tgt_ours)), tau=0.005)
    for step in range(25):
        grads = [torch.randn_like(p) * (10.0 ** ((step % 5) - 3)) for p in ref]
        for p, g in zip(ref, grads):
            p.grad = g.clone()
        opt_ref.step()
        with torch.no_grad():
            for p, t in zip(ref[:2], tgt_ref):
                t.copy_(0.005 * p + (1 - 0.005) * t)
        opt.step(grads)
    torch.cuda.synchronize()
    assert int(opt.state[0]) == 25 and int(opt.state[1]) == 0
    for p, q, lr in zip(ref, ours, lrs):
        assert (p.detach() - q).abs().max().item() <= 1e-6 + 1e-3 * lr * 25
    for t, u in zip(tgt_ref, tgt_ours):
        assert (t - u).abs().max().item() <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("n,n_actions,obs_dim", [(1, 1, 11), (127, 1, 11), (128, 1, 11), (1000, 2, 16), (70_001, 1, 11),
                                                  (300, 8, 5), (5000, 4, 11)])
def test_tcgen05_policy_matches_bf16_emulation(n, n_actions, obs_dim):
    """csrc/policy_mlp.cu against the same arithmetic spelled out in PyTorch (operands rounded to bf16, fp32
    accumulation, fp32 head) and against the fp32 policy; ragged tiles, several tiles per CTA, padded shapes."""
    from sac_agent_b200.networks import ActorNetwork, TensorCorePolicy
    torch.manual_seed(n)
    max_a = np.linspace(1.0, 0.5, n_actions).astype(np.float32)
    actor = ActorNetwork(None, (obs_dim,), max_a, n_actions=n_actions).cuda()
    pol = TensorCorePolicy(actor)
    obs, eps = torch.rand(n, obs_dim, device="cuda"), torch.randn(n, n_actions, device="cuda")
    got = pol.act(obs, eps)
    bf = lambda t: t.to(torch.bfloat16).float()  # noqa: E731
    with torch.no_grad():
        h = bf(torch.relu(bf(obs) @ bf(actor.fc1.weight).T + actor.fc1.bias))
        h = torch.relu(h @ bf(actor.fc2.weight).T + actor.fc2.bias)     # stays fp32: the heads are fp32 dot products
        mean, raw = h @ actor.mean.weight.T + actor.mean.bias, h @ actor.std.weight.T + actor.std.bias
        emu = torch.tanh(mean + eps * torch.exp(-5.0 + 3.5 * (torch.tanh(raw) + 1.0))) * actor.max_action
        ref = actor.sample_normal(obs, reparameterize=False, eps=eps)[0]
    assert got.shape == (n, n_actions)
    assert (got - emu).abs().max().item() < 2e-3     # accumulation order and bf16 re-rounding of near-ties
    assert (got - ref).abs().max().item() < 3e-2     # bf16 inputs against the fp32 policy
    # without eps: Philox draws, a fresh set per call, reproducible per (seed, call number)
    k = pol.steps
    a1, a2 = pol.act(obs).clone(), pol.act(obs).clone()
    pol.steps = k + 1
    assert not torch.equal(a1, a2) and torch.equal(pol.act(obs), a2)
    # the packed weights follow the actor after refresh()
    with torch.no_grad():
        actor.mean.bias.add_(0.5)
    stale = pol.act(obs, eps)
    pol.refresh()
    assert torch.equal(stale, got) and not torch.equal(pol.act(obs, eps), got)


@pytest.mark.gpu
def test_tcgen05_policy_draws_are_standard_normal():
    from sac_agent_b200.networks import ActorNetwork, TensorCorePolicy
    actor = ActorNetwork(None, (11,), np.array([1.0], dtype=np.float32), n_actions=1).cuda()
    with torch.no_grad():   # mean 0, log_std = -5 + 3.5 * (tanh(0) + 1) = -1.5: action = tanh(e * exp(-1.5))
        for layer in (actor.mean, actor.std):
            layer.weight.zero_(); layer.bias.zero_()
    pol = TensorCorePolicy(actor, seed=3)
    a = pol.act(torch.rand(400_000, 11, device="cuda"))
    e = torch.atanh(a.double()) / np.exp(-1.5)
    assert abs(e.mean().item()) < 0.01 and abs(e.std().item() - 1.0) < 0.01
    assert abs((e ** 4).mean().item() - 3.0) < 0.1


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("tf32", 5e-3), ("bf16", 5e-2)])
def test_reduced_precision_acting_stays_close_to_fp32(precision, tol):
    """choose_action with tensor-core inputs (acting only): same noise, actions within the rounding of the
    input format of the fp32 policy."""
    cfg = S.load_config()
    env = type("E", (), {"action_space": S.Box(low=-1, high=1, dtype=np.float32)})()
    mem = S.ReplayBuffer(2048, (11,), 1, precision="fp32", device=0, as_torch=True)
    agent = ContinuousAgent(cfg, None, (11,), env, device=0, memory=mem, policy_precision=precision)
    obs = torch.rand(4096, 11, device="cuda")
    torch.manual_seed(7)
    a_lo = agent.choose_action(obs)
    agent.policy_precision = "fp32"
    torch.manual_seed(7)
    a_hi = agent.choose_action(obs)
    assert a_lo.dtype == torch.float32 and a_lo.shape == (4096, 1)
    assert (a_lo - a_hi).abs().max().item() < tol and not torch.equal(a_lo, a_hi)
    with pytest.raises(ValueError):
        ContinuousAgent(cfg, None, (11,), env, device=0, memory=mem, policy_precision="fp8")
    mem.close()


@pytest.mark.gpu
def test_agent_kernel_argument_errors():
    """The agent entry points of the ABI reject bad arguments with codes instead of launching."""
    from sac_agent_b200 import _lib
    from sac_agent_b200.continuous_agent import DeviceAdam
    from sac_agent_b200.networks import ActorNetwork, TensorCorePolicy
    L = _lib.lib()
    x = torch.zeros(8, device="cuda")
    assert L.boatagent_policy_act(None, x.data_ptr(), None, x.data_ptr(), 0, 0, 8, 11, 1, x.data_ptr(), None) != 0
    assert L.boatagent_policy_act(x.data_ptr(), x.data_ptr(), None, x.data_ptr(), 0, 0, 0, 11, 1, x.data_ptr(), None) != 0
    assert L.boatagent_policy_act(x.data_ptr(), x.data_ptr(), None, x.data_ptr(), 0, 0, 8, 17, 1, x.data_ptr(), None) != 0
    assert L.boatagent_policy_act(x.data_ptr(), x.data_ptr(), None, x.data_ptr(), 0, 0, 8, 11, 3, x.data_ptr(), None) != 0
    assert L.boatagent_gaussian_head_forward(None, None, None, None, 4, 1, None, None, None) != 0
    assert L.boatagent_adam_polyak_step(None, 1, 0.9, 0.999, 1e-8, 0.0, x.data_ptr(), None) != 0
    with pytest.raises(ValueError):
        DeviceAdam([([torch.zeros(4, device="cuda") for _ in range(65)], 1e-3)])
    with pytest.raises(ValueError):
        DeviceAdam([([torch.zeros(4, device="cuda", dtype=torch.float64)], 1e-3)])
    with pytest.raises(ValueError):   # three actions: not one of the instantiated head counts
        TensorCorePolicy(ActorNetwork(None, (11,), np.ones(3, dtype=np.float32), n_actions=3).cuda())
    pol = TensorCorePolicy(ActorNetwork(None, (11,), np.ones(1, dtype=np.float32), n_actions=1).cuda())
    with pytest.raises(ValueError):
        pol.act(torch.zeros(4, 12, device="cuda"))
