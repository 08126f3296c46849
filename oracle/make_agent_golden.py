"""Golden vectors for the SAC update: the UNMODIFIED reference
`ContinuousAgent.learn()` (agent/continuous_agent.py:96-154, networks/networks.py) run on the CPU in
the build container, with prescribed weights, batch and Gaussian draws.

    python oracle/make_agent_golden.py        # rewrites tests/golden/agent_update.npz

What is prescribed (so that the device learner can be fed exactly the same):
  * initial weights: numpy RandomState(seed) uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)), written into the
    reference's networks through `load_state_dict` -- `init_weights()` below regenerates them anywhere;
  * the batch: the reference's ReplayBuffer arrays filled directly, `np.random.choice` pinned to
    arange(batch) for the duration of `learn()`;
  * the noise: `torch.distributions.Normal.sample / rsample` replaced by loc + eps * scale with recorded
    eps (that is their own formula; only the source of eps changes).
What is recorded after updates 1 and 3: every parameter of the five networks and the gradients the
update left on them, at up to 512 fixed positions per tensor (plus each tensor's sum and L2 norm).

TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "agent_update.npz")
NETS = ("actor", "critic_1", "critic_2", "value", "target_value")
# networks.py:24-27, 83-85, 116-118 -- (name, out_features, in_features) in creation order
LAYERS = {
    "actor": (("fc1", 256, 11), ("fc2", 256, 256), ("mean", 1, 256), ("std", 1, 256)),
    "critic_1": (("fc1", 256, 12), ("fc2", 256, 256), ("q", 1, 256)),
    "critic_2": (("fc1", 256, 12), ("fc2", 256, 256), ("q", 1, 256)),
    "value": (("fc1", 256, 11), ("fc2", 256, 256), ("v", 1, 256)),
}
SAMPLES_PER_TENSOR = 512


def init_weights(seed: int) -> dict:
    """{net: {param_name: float32 array}} for actor, critic_1, critic_2, value (target_value := value)."""
    rng = np.random.RandomState(seed)
    out = {}
    for net, layers in LAYERS.items():
        sd = {}
        for name, n_out, n_in in layers:
            k = 1.0 / np.sqrt(n_in)
            sd[f"{name}.weight"] = rng.uniform(-k, k, size=(n_out, n_in)).astype(np.float32)
            sd[f"{name}.bias"] = rng.uniform(-k, k, size=(n_out,)).astype(np.float32)
        out[net] = sd
    out["target_value"] = {k: v.copy() for k, v in out["value"].items()}
    return out


def make_batch(seed: int, batch: int) -> dict:
    rng = np.random.RandomState(seed + 1)
    return {
        "state": rng.uniform(0.0, 1.0, size=(batch, 11)).astype(np.float32),
        "action": rng.uniform(-1.0, 1.0, size=(batch, 1)).astype(np.float32),
        "reward": (-np.abs(rng.normal(0.0, 0.5, size=batch))).astype(np.float32),
        "new_state": rng.uniform(0.0, 1.0, size=(batch, 11)).astype(np.float32),
        "done": rng.uniform(size=batch) < 0.05,
    }


def make_noise(seed: int, n_updates: int, batch: int) -> np.ndarray:
    """[n_updates][2 (sample, rsample)][batch][1] standard-normal draws."""
    return np.random.RandomState(seed + 2).normal(size=(n_updates, 2, batch, 1)).astype(np.float32)


def sample_positions(shape, key: str) -> np.ndarray:
    n = int(np.prod(shape))
    if n <= SAMPLES_PER_TENSOR:
        return np.arange(n)
    h = sum(ord(c) * (i + 1) for i, c in enumerate(key)) % (2 ** 31)
    return np.sort(np.random.RandomState(h).choice(n, SAMPLES_PER_TENSOR, replace=False))


def condense(named: dict) -> dict:
    """{key: tensor-as-array} -> {key_at: sampled values, key_sum, key_l2}."""
    out = {}
    for key, arr in named.items():
        a = np.asarray(arr, dtype=np.float32)
        flat = a.reshape(-1)
        out[key + "@"] = flat[sample_positions(a.shape, key)]
        out[key + "#sum"] = np.float64(flat.astype(np.float64).sum())
        out[key + "#l2"] = np.float64(np.sqrt((flat.astype(np.float64) ** 2).sum()))
    return out


def reference_learn(seed: int = 0, n_updates: int = 3, batch: int = 1024, record_after=(1, 3)) -> dict:
    """Runs the reference's ContinuousAgent.learn() `n_updates` times; returns the recorded arrays."""
    import torch as T

    R.import_reference()  # gym / matplotlib stubs + sys.path
    from agent.continuous_agent import ContinuousAgent  # noqa: the reference, unmodified
    from gym.spaces import Box

    cfg = R.load_config()
    cfg["agent"]["batch_size"] = batch
    env = types.SimpleNamespace(action_space=Box(low=-1, high=1, dtype=np.float32))
    T.manual_seed(seed)
    with tempfile.TemporaryDirectory() as tmp:
        agent = ContinuousAgent(cfg, tmp, (11,), env)
    nets = {n: getattr(agent, n) for n in NETS}
    w0 = init_weights(seed)
    for n, net in nets.items():
        net.load_state_dict({k: T.from_numpy(v.copy()) for k, v in w0[n].items()})
    b = make_batch(seed, batch)
    mem = agent.memory  # buffer.py:7-11
    mem.state_memory[:batch] = b["state"]
    mem.new_state_memory[:batch] = b["new_state"]
    mem.action_memory[:batch] = b["action"]
    mem.reward_memory[:batch] = b["reward"]
    mem.terminal_memory[:batch] = b["done"]
    mem.mem_cntr = batch
    noise = make_noise(seed, n_updates, batch)

    draws = []
    Normal = T.distributions.Normal
    orig_sample, orig_rsample, orig_choice = Normal.sample, Normal.rsample, np.random.choice

    def rsample(self, sample_shape=T.Size()):
        return self.loc + T.from_numpy(draws.pop(0)) * self.scale

    def sample(self, sample_shape=T.Size()):
        with T.no_grad():
            return self.loc + T.from_numpy(draws.pop(0)) * self.scale

    rec = {}
    Normal.sample, Normal.rsample = sample, rsample
    np.random.choice = lambda max_mem, n, *a, **k: np.arange(n)
    try:
        for u in range(1, n_updates + 1):
            draws[:] = [noise[u - 1, 0], noise[u - 1, 1]]  # learn() calls sample (:113) before rsample (:127)
            agent.learn()
            assert not draws
            if u in record_after:
                named = {}
                for n, net in nets.items():
                    for k, p in net.named_parameters():
                        named[f"u{u}/{n}/{k}/w"] = p.detach().numpy()
                        if n != "target_value":
                            named[f"u{u}/{n}/{k}/g"] = p.grad.detach().numpy()
                rec.update(condense(named))
    finally:
        Normal.sample, Normal.rsample, np.random.choice = orig_sample, orig_rsample, orig_choice
    rec["meta/seed"], rec["meta/n_updates"], rec["meta/batch"] = np.int64(seed), np.int64(n_updates), np.int64(batch)
    rec["meta/record_after"] = np.asarray(record_after, dtype=np.int64)
    rec["meta/param_names"] = np.asarray(
        [f"{n}/{k}/{tuple(p.shape)}" for n, net in nets.items() for k, p in net.named_parameters()])
    for key in ("learning_rate_alpha", "learning_rate_beta", "gamma", "tvn_parameter_modulation_tau", "reward_scale"):
        rec["meta/" + key] = np.float64(cfg["agent"][key])
    return rec


def main():
    rec = reference_learn(seed=0, n_updates=3, batch=1024)
    np.savez_compressed(OUT, **rec)
    print(f"wrote {OUT}: {len(rec)} arrays, {os.path.getsize(OUT) / 1024:.0f} kB")


if __name__ == "__main__":
    main()
