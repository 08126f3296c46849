#!/usr/bin/env python
"""End-to-end SAC (v1: actor, twin critics, value + target value) on the GPU-resident batched BoatEnv
and the fused device replay buffer -- BASELINE.json configs[4] / SURVEY.md 8(f) rank 1.

The algorithm and hyper-parameters are the reference's (agent/continuous_agent.py:96-154,
networks/networks.py:47-70, the `agent:` block of original_config.yaml); what changes is the data
path: N envs step per launch, `ReplayBuffer.step_store` writes the transitions straight from env
state into the device ring (main.py:81-88 in one kernel), `sample_buffer(as_torch=True)` gathers
the batch on the device -- no host round trip anywhere in the loop.  The MLPs are plain PyTorch
(dense 256-wide layers: cuBLAS is the right tool; they are not part of the hot path of this repo).

    python examples/train_sac.py --envs 65536 --iters 200
    torchrun --nproc-per-node 8 examples/train_sac.py --envs 65536        # replicas: one agent per GPU shard

Multi-GPU is "replicas only" (SURVEY.md 8e): every rank trains its own agent on its own env shard,
like the reference's `-p` mode runs independent models; only the episode statistics are all-reduced.
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sac_agent_b200 as S  # noqa: E402

LOG_STD_MIN, LOG_STD_MAX = -5.0, 2.0  # networks.py:48-49


def mlp(n_in, n_out, hidden=256):
    return nn.Sequential(nn.Linear(n_in, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(), nn.Linear(hidden, n_out))


class Actor(nn.Module):
    """tanh-squashed Gaussian; the std head is squashed into [exp(-5), exp(2)] (networks.py:47-70)."""

    def __init__(self, obs_dim, act_dim, max_action=1.0):
        super().__init__()
        self.body = mlp(obs_dim, 2 * act_dim)
        self.max_action = max_action

    def sample(self, obs, reparameterize):
        mean, raw = self.body(obs).chunk(2, dim=-1)
        log_std = LOG_STD_MIN + 0.5 * (LOG_STD_MAX - LOG_STD_MIN) * (torch.tanh(raw) + 1.0)
        dist = torch.distributions.Normal(mean, log_std.exp())
        u = dist.rsample() if reparameterize else dist.sample()
        action = torch.tanh(u) * self.max_action
        log_prob = (dist.log_prob(u) - torch.log(1.0 - action.pow(2) + 1e-6)).sum(-1)
        return action, log_prob


class SAC:
    def __init__(self, cfg, obs_dim=11, act_dim=1, device="cuda"):
        a = cfg.agent
        self.gamma, self.tau, self.scale, self.batch = a.gamma, a.tvn_parameter_modulation_tau, a.reward_scale, a.batch_size
        self.actor = Actor(obs_dim, act_dim).to(device)
        self.q1, self.q2 = mlp(obs_dim + act_dim, 1).to(device), mlp(obs_dim + act_dim, 1).to(device)
        self.v, self.v_targ = mlp(obs_dim, 1).to(device), mlp(obs_dim, 1).to(device)
        self.v_targ.load_state_dict(self.v.state_dict())
        self.opt_actor = torch.optim.Adam(self.actor.parameters(), lr=a.learning_rate_alpha)
        self.opt_q = torch.optim.Adam(list(self.q1.parameters()) + list(self.q2.parameters()), lr=a.learning_rate_beta)
        self.opt_v = torch.optim.Adam(self.v.parameters(), lr=a.learning_rate_beta)

    @torch.no_grad()
    def act(self, obs):
        return self.actor.sample(obs, reparameterize=False)[0]

    def update(self, s, a, r, s2, done):
        """One SAC-v1 update (continuous_agent.py:96-154)."""
        with torch.no_grad():
            v_next = self.v_targ(s2).squeeze(-1)
            v_next[done] = 0.0
            q_hat = self.scale * r + self.gamma * v_next
            a_new, logp = self.actor.sample(s, reparameterize=False)
            q_min = torch.min(self.q1(torch.cat([s, a_new], 1)), self.q2(torch.cat([s, a_new], 1))).squeeze(-1)
            v_target = q_min - logp
        v_loss = 0.5 * F.mse_loss(self.v(s).squeeze(-1), v_target)
        self.opt_v.zero_grad(set_to_none=True); v_loss.backward(); self.opt_v.step()

        a_rep, logp = self.actor.sample(s, reparameterize=True)
        q_min = torch.min(self.q1(torch.cat([s, a_rep], 1)), self.q2(torch.cat([s, a_rep], 1))).squeeze(-1)
        actor_loss = (logp - q_min).mean()
        self.opt_actor.zero_grad(set_to_none=True); actor_loss.backward(); self.opt_actor.step()

        sa = torch.cat([s, a], 1)
        q_loss = 0.5 * F.mse_loss(self.q1(sa).squeeze(-1), q_hat) + 0.5 * F.mse_loss(self.q2(sa).squeeze(-1), q_hat)
        self.opt_q.zero_grad(set_to_none=True); q_loss.backward(); self.opt_q.step()

        with torch.no_grad():  # Polyak average of the target value net (continuous_agent.py:66-80)
            for p, pt in zip(self.v.parameters(), self.v_targ.parameters()):
                pt.mul_(1.0 - self.tau).add_(p, alpha=self.tau)
        return v_loss.item(), actor_loss.item(), q_loss.item()


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536, help="total envs over all ranks")
    ap.add_argument("--iters", type=int, default=200, help="env steps (each over every env)")
    ap.add_argument("--updates-per-iter", type=int, default=1)
    ap.add_argument("--experiment", type=int, default=5)
    ap.add_argument("--buffer", type=int, default=1_000_000, help="ring capacity per rank (original_config max_size)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--log-every", type=int, default=50)
    args = ap.parse_args(argv)

    rank, local_rank, world = S.sharding.dist_info()
    torch.cuda.set_device(local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.manual_seed(args.seed + rank)
    cfg = S.load_config(base_settings__experiment=args.experiment)
    env = S.make_sharded_env(cfg, args.envs, seed=args.seed, precision="fp32", auto_reset=True)
    buf = S.ReplayBuffer(max(args.buffer, env.n_envs), (11,), 1, precision="fp32", device=local_rank,
                         seed=args.seed + rank, as_torch=True)
    agent = SAC(cfg, device=env.device)
    env.reset()
    losses, t0 = (math.nan,) * 3, time.perf_counter()
    for it in range(1, args.iters + 1):
        actions = agent.act(env.obs).squeeze(-1)
        buf.step_store(env, actions, done_flag_mode=1)          # env.step + agent.remember, one kernel
        if buf.mem_cntr >= agent.batch:
            for _ in range(args.updates_per_iter):
                s, a, r, s2, d = buf.sample_buffer(agent.batch)
                losses = agent.update(s, a, r, s2, d)
        if it % args.log_every == 0 or it == args.iters:
            torch.cuda.synchronize()
            stats = S.all_reduce_counters(env.counters_tensor())
            dt = time.perf_counter() - t0
            if rank == 0:
                print(f"iter {it:6d}  env-steps/s {it * args.envs / dt:.3e}  updates/s {it * args.updates_per_iter / dt:7.1f}  "
                      f"episodes {stats['episodes']:.0f}  mean return {stats['return_mean']:9.2f}  goals {stats['reached_goal']:.0f}  "
                      f"losses v/pi/q {losses[0]:.3g} {losses[1]:.3g} {losses[2]:.3g}", flush=True)
    env.close(); buf.close()
    if world > 1:
        torch.distributed.destroy_process_group()
    return losses


if __name__ == "__main__":
    main()
