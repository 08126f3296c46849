#!/usr/bin/env python
"""Where one iteration of examples/train_sac.py goes (65536 envs, batch 1024, one B200): each phase timed
alone with CUDA events over 200 repetitions.  One JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import sac_agent_b200 as S  # noqa: E402


def timed(fn, iters=200, warmup=10):
    for _ in range(warmup):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    n = int(os.environ.get("SAC_ENVS", 65536))
    cfg = S.load_config(base_settings__experiment=5)
    env = S.BatchedBoatEnv(cfg, n, seed=0, precision="fp32", device=0, auto_reset=True)
    mem = S.ReplayBuffer(max(1_000_000, n), (11,), 1, precision="fp32", device=0, as_torch=True)
    out = {"envs": n, "batch": int(cfg.agent.batch_size)}
    for graph in (True, False):
        agent = S.ContinuousAgent(cfg, None, (11,), env, device=0, use_cuda_graph=graph, memory=mem)
        obs = env.reset()
        act = agent.choose_action_graphed if graph else agent.choose_action
        a = act(obs).squeeze(-1)
        agent.step_and_remember(env, a)
        agent.learn()
        tag = "graph" if graph else "eager"
        out[f"policy_ms_{tag}"] = timed(lambda: act(obs))
        out[f"update_ms_{tag}"] = timed(lambda: agent._run_update())
        out[f"learn_ms_{tag}"] = timed(lambda: agent.learn())
    out["step_store_ms"] = timed(lambda: agent.step_and_remember(env, a))
    out["sample_ms"] = timed(lambda: mem.sample_buffer(agent.batch_size, out=agent._batch))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
