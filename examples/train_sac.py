#!/usr/bin/env python
"""End-to-end SAC (v1: actor, twin critics, value + target value) on the GPU-resident batched BoatEnv
and the fused device replay buffer -- BASELINE.json configs[4] / SURVEY.md 8(f) rank 1.

The loop is the reference's main.py:72-90 with every piece on the device:

    action = agent.choose_action(obs)            -> ContinuousAgent.choose_action_graphed (N envs, one graph)
    obs_, r, done, info = env.step(action)       \\  ContinuousAgent.step_and_remember: ONE kernel
    agent.remember(obs, action, r, obs_, done)   /   (boatenv_step_store)
    agent.learn()                                -> sample-gather kernel + one captured CUDA graph

Algorithm and hyper-parameters are the reference's (agent/continuous_agent.py:96-154,
networks/networks.py, the `agent:` block of original_config.yaml); the update is checked against a
recorded run of the reference's own learn() in tests/test_agent_golden.py.

    python examples/train_sac.py --envs 65536 --iters 200
    torchrun --nproc-per-node 8 examples/train_sac.py --envs 65536        # replicas: one agent per GPU shard
    torchrun --nproc-per-node 8 examples/train_sac.py --experiments-root experiments --tune
                                          # the reference's `main.py -p 8`: a tuned_configs.yaml draw per member

Multi-GPU is "replicas only" (SURVEY.md 8e): every rank trains its own agent on its own env shard,
like the reference's `-p` mode runs independent models; only the episode statistics are all-reduced.
The last line printed by rank 0 is a JSON summary (device-timed after `--warmup-iters`).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sac_agent_b200 as S  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536, help="total envs over all ranks")
    ap.add_argument("--iters", type=int, default=200, help="env steps (each over every env)")
    ap.add_argument("--warmup-iters", type=int, default=20, help="untimed iterations (graph capture happens here)")
    ap.add_argument("--updates-per-iter", type=int, default=1)
    ap.add_argument("--experiment", type=int, default=5)
    ap.add_argument("--buffer", type=int, default=1_000_000, help="ring capacity per rank (original_config max_size)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--log-every", type=int, default=50)
    ap.add_argument("--no-graph", action="store_true", help="eager update / policy (for comparison)")
    ap.add_argument("--experiments-root", default=None,
                    help="write the reference's experiment tree (configs/, checkpoints/, console.csv, overview.csv, "
                         "terminations.csv; main.py:116-133) under this directory, one experiment per rank")
    ap.add_argument("--tune", action="store_true",
                    help="population mode (main.py -p): every rank trains with its own tuned_configs.yaml draw")
    ap.add_argument("--policy-precision", default="fp32", choices=["fp32", "tf32", "bf16", "tcgen05"],
                    help="dense layers of choose_action on tensor-core inputs (acting only; the learner stays fp32)")
    ap.add_argument("--set", action="append", default=[], metavar="SECTION__KEY=VALUE",
                    help="override a config key, e.g. --set agent__learning_rate_alpha=0.001 (original_config.yaml names)")
    ap.add_argument("--pbt-every", type=int, default=0,
                    help="population-based training (with torchrun, one member per GPU): every this many iterations the "
                         "bottom quarter adopts the best member's weights + hyper-parameters and perturbs them")
    ap.add_argument("--overlap", action="store_true",
                    help="acting and learning on two streams (OverlappedActorLearner: the policy lags one update)")
    args = ap.parse_args(argv)

    rank, local_rank, world = S.sharding.dist_info()
    torch.cuda.set_device(local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.manual_seed(args.seed + rank)
    overrides = {}
    for item in args.set:
        k, v = item.split("=", 1)
        overrides[k] = float(v) if any(ch in v for ch in ".e") else int(v)
    cfg = S.load_config(base_settings__experiment=args.experiment, **overrides)
    exp = None
    if args.experiments_root:   # utils/build_experiment.py: one experiment directory per population member
        import random
        exp = S.Experiment(experiment_name=f"experiment_r{rank}", subdir=f"setting_{args.experiment}",
                           root=args.experiments_root, rng=random.Random(args.seed + rank))
        tuned = exp.save_configs(cfg)
        if args.tune:
            cfg = tuned
    env = S.make_sharded_env(cfg, args.envs, seed=args.seed, precision="fp32", auto_reset=True)
    mem = S.ReplayBuffer(max(args.buffer, env.n_envs), (11,), 1, precision="fp32", device=local_rank,
                         seed=args.seed + rank, as_torch=True)
    agent = S.ContinuousAgent(cfg, exp.experiment_dir if exp else None, env.observation_space.shape, env,
                              device=local_rank, seed=args.seed + rank,
                              use_cuda_graph=not args.no_graph, memory=mem, policy_precision=args.policy_precision)
    act = agent.choose_action if args.no_graph else agent.choose_action_graphed
    obs = env.reset()
    pipe = S.OverlappedActorLearner(agent, env, done_flag_mode=1) if args.overlap else None
    losses = None
    rows, best, prev = [], float("-inf"), {"episodes": 0.0, "return_sum": 0.0}
    import random as _random
    pbt_prev, pbt_log, pbt_rng = {"episodes": 0.0, "return_sum": 0.0}, [], _random.Random(1000 + args.seed + rank)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(1, args.warmup_iters + args.iters + 1):
        if it == args.warmup_iters + 1:
            if pipe:
                pipe.finish()
            t0.record()
        if pipe:
            losses = pipe.step()
        else:
            actions = act(obs).squeeze(-1)
            agent.step_and_remember(env, actions, done_flag_mode=1)   # env.step + agent.remember, one kernel
            for _ in range(args.updates_per_iter):
                out = agent.learn()
                losses = out if out is not None else losses
        if args.pbt_every and world > 1 and it % args.pbt_every == 0:
            if pipe:
                pipe.sync()
            c = env.counters()
            n = c["episodes"] - pbt_prev["episodes"]
            score = (c["return_sum"] - pbt_prev["return_sum"]) / n if n else float("nan")
            pbt_prev = c
            hp, info = S.population.exploit_explore(agent.training_tensors(), agent.hyperparameters(), score, pbt_rng)
            if info["adopted_from"] is not None:
                agent.set_hyperparameters(**hp)
                agent._weights_changed()
            pbt_log.append({"iter": it, "score": score, "adopted_from": info["adopted_from"], "hp": hp,
                            "ranking": info["ranking"]})
        if pipe and args.log_every and it % args.log_every == 0:
            pipe.sync()   # counters, losses and checkpoints below run on the default stream: order it after the pipeline
        if exp and args.log_every and it % args.log_every == 0:
            # one console.csv row per logging interval: the mean return of the episodes that ended in it
            c = env.counters()
            n = c["episodes"] - prev["episodes"]
            score = (c["return_sum"] - prev["return_sum"]) / n if n else float("nan")
            prev = c
            if n and score > best:
                best = score
                agent.save_models()          # main.py:104-106 keeps the checkpoints of improving episodes
            kind = max(S.TERM_NAMES[1:], key=lambda k: c[k])
            avg = c["return_sum"] / c["episodes"] if c["episodes"] else float("nan")
            rows.append([(rank, it), f"{kind}-{int(c[kind])}", score, best, avg, "", ""])
        if rank == 0 and args.log_every and it % args.log_every == 0:
            c = env.counters()
            lv = [float(x) for x in losses] if losses is not None else [float("nan")] * 3
            print(f"iter {it:6d}  episodes {c['episodes']:.0f}  goals {c['reached_goal']:.0f}  "
                  f"losses v/pi/q {lv[0]:.3g} {lv[1]:.3g} {lv[2]:.3g}", flush=True)
    if pipe:
        pipe.finish()
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    stats = S.all_reduce_counters(env.counters_tensor())
    sec = float(ms.item()) * 1e-3
    summary = {"example": "train_sac", "overrides": overrides, "n_gpus": world, "envs_total": args.envs, "iters": args.iters,
               "updates_per_iter": args.updates_per_iter, "cuda_graph": not args.no_graph, "overlap": bool(args.overlap),
               "policy_precision": args.policy_precision,
               "env_steps_per_s": args.iters * args.envs / sec,
               "updates_per_s_per_replica": args.iters * args.updates_per_iter / sec,
               "ms_per_iter": 1e3 * sec / args.iters, "episodes": stats["episodes"],
               "mean_return": stats["return_mean"], "reached_goal": stats["reached_goal"],
               "losses_v_pi_q": [float(x) for x in losses] if losses is not None else None}
    if args.pbt_every:
        summary["pbt_rounds"] = len(pbt_log)
        summary["pbt_adoptions_rank0"] = [e for e in pbt_log if e["adopted_from"] is not None]
        summary["hyperparameters_rank0"] = agent.hyperparameters()
    if exp:
        c = env.counters()
        exp.write_console(rows)
        exp.append_overview(best)
        exp.write_terminations(S.info_from_counters(c, max(S.TERM_NAMES[1:], key=lambda k: c[k])))
        summary["experiment_dir"], summary["best_interval_return"] = exp.experiment_dir, best
        if args.tune:
            summary["hpset"] = exp.tuner.hpset
        if world > 1:   # the population's best member, like reading overview.csv after a `-p` run
            scores = [None] * world
            torch.distributed.all_gather_object(scores, (best, rank, exp.experiment_name))
            summary["population_best"] = max(scores)
    if rank == 0:
        print(json.dumps(summary), flush=True)
    env.close(); mem.close()
    if world > 1:
        torch.distributed.destroy_process_group()
    return summary


if __name__ == "__main__":
    main()
