#!/usr/bin/env python
"""Secondary measurements (not the driver's bench line): the other BASELINE.json configs and the
other rows of SURVEY.md section 8 on one B200, each with its algorithmic bytes and the fraction of
the measured HBM copy peak.  Prints one JSON object per line; run on the GPU box:

    python profiles/bench_extra.py > gpurun_out/extra.jsonl
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import sac_agent_b200 as S  # noqa: E402

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def timed(fn, iters, warmup=5):
    for _ in range(warmup):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters  # ms


def emit(name, ms, units, unit_name, bytes_per_unit=None, **extra):
    line = {"case": name, "ms": ms, "rate": units / (ms * 1e-3), "unit": unit_name + "/s", **extra}
    if bytes_per_unit:
        gbs = bytes_per_unit * units / (ms * 1e-3) / 1e9
        line.update(algorithmic_bytes_per_unit=bytes_per_unit, achieved_gbs=gbs, frac_of_hbm_peak=gbs / PEAK)
    print(json.dumps(line), flush=True)


def boat_case(name, experiment, precision, n, scale, bytes_per_step, warm, iters, k=1, per_substep=False):
    cfg = S.load_config(base_settings__experiment=experiment)
    env = S.BatchedBoatEnv(cfg, n, seed=1, precision=precision, device=0, auto_reset=True)
    env.reset()
    acts = env.uniform_actions(0, scale)
    acts_k = torch.empty((k, n), dtype=acts.dtype, device=acts.device) if per_substep else None
    t = [0]

    def fill():
        if per_substep:   # the policy emits a fresh action for every sub-step: same episode statistics as K = 1
            for q in range(k):
                env.uniform_actions(t[0] * k + q, scale, out=acts_k[q])
        else:             # action repeat (frame skip): one action held for the k sub-steps
            env.uniform_actions(t[0], scale, out=acts)

    def launch():
        if k == 1:
            env.step(acts)
        else:
            env.step_k(acts_k if per_substep else acts, k)

    def step():
        fill(); launch()
        t[0] += 1
    for _ in range(warm):
        step()

    # fresh actions every step (a repeated action tensor would drive every rudder to its limit within
    # ~20 steps); only the step launch sits between the event pairs, like in bench.py
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in pairs:
        fill()
        a.record()
        launch()
        b.record()
        t[0] += 1
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in pairs) / iters
    emit(name, ms, n * k, "env-steps", bytes_per_step, n_envs=n, precision=precision, experiment=experiment, k=k,
         episodes=env.counters()["episodes"])
    env.close()


def main():
    torch.cuda.set_device(0)
    M = 1 << 20
    if os.environ.get("BENCH_EXTRA_ONLY") == "k8":
        boat_case("exp6_fp32_16M_k8", 6, "fp32", 16 * M, 1.0, 4 + 5 + (112 + 44) / 8, 100, 30, k=8)
        boat_case("exp6_fp32_16M_k8_per_substep_actions", 6, "fp32", 16 * M, 1.0, 4 + 4 + 5 + (112 + 44) / 8, 80, 30, k=8,
                  per_substep=True)
        boat_case("exp6_fp32_16M_k8_long_episodes", 6, "fp32", 16 * M, 0.02, 4 + 5 + (112 + 44) / 8, 30, 30, k=8)
        boat_case("exp6_fp32_16M_k1_long_episodes", 6, "fp32", 16 * M, 0.02, 165, 100, 100)
        return
    if os.environ.get("BENCH_EXTRA_ONLY") == "reset":
        reset_cases(M)
        return
    if os.environ.get("BENCH_EXTRA_ONLY") == "fp32":
        boat_case("exp1_fp32_16M", 1, "fp32", 16 * M, 1.0, 133, 300, 100)
        boat_case("exp4_fp32_16M", 4, "fp32", 16 * M, 1.0, 149, 300, 100)
        boat_case("exp6_fp32_16M", 6, "fp32", 16 * M, 1.0, 165, 600, 100)
        return
    if os.environ.get("BENCH_EXTRA_ONLY") == "fp64":
        boat_case("exp3_fp64_4M", 3, "fp64", 4 * M, 0.05, 249, 50, 50)
        boat_case("exp6_fp64_4M", 6, "fp64", 4 * M, 1.0, 8 + 104 + 104 + 88 + 8 + 1, 300, 50)
        return
    if os.environ.get("BENCH_EXTRA_ONLY") == "replay":
        replay_cases(M)
        return
    # BASELINE.json configs[1]: exp 3, 4096 envs, fp64 (launch-latency bound at this size) and the same at 4M envs
    boat_case("exp3_fp64_4096", 3, "fp64", 4096, 0.05, 249, 50, 200)
    boat_case("exp3_fp64_4M", 3, "fp64", 4 * M, 0.05, 249, 50, 50)
    boat_case("exp6_fp64_4M", 6, "fp64", 4 * M, 1.0, 8 + 104 + 104 + 88 + 8 + 1, 300, 50)
    # fp32 production mode, the three state sizes of SURVEY.md 8(d)
    boat_case("exp1_fp32_16M", 1, "fp32", 16 * M, 1.0, 133, 300, 100)
    boat_case("exp4_fp32_16M", 4, "fp32", 16 * M, 1.0, 149, 300, 100)
    boat_case("exp6_fp32_16M", 6, "fp32", 16 * M, 1.0, 165, 600, 100)
    # K fused sub-steps (obs only at the end): bytes per env-step shrink, the kernel turns compute bound
    boat_case("exp6_fp32_16M_k8", 6, "fp32", 16 * M, 1.0, 4 + 5 + (112 + 44) / 8, 100, 30, k=8)
    boat_case("exp6_fp32_16M_k8_per_substep_actions", 6, "fp32", 16 * M, 1.0, 4 + 4 + 5 + (112 + 44) / 8, 80, 30, k=8,
              per_substep=True)
    # the same with long episodes (small steering noise: nobody resets within the run)
    boat_case("exp6_fp32_16M_k8_long_episodes", 6, "fp32", 16 * M, 0.02, 4 + 5 + (112 + 44) / 8, 30, 30, k=8)
    boat_case("exp6_fp32_16M_k1_long_episodes", 6, "fp32", 16 * M, 0.02, 165, 100, 100)

    reset_cases(M)
    replay_cases(M)

    # toy envs, 1M envs each, fp32 (BASELINE.json configs[3]): k iterations per launch
    car = S.ToyCar(n_envs=M, jitter=0.1, seed=0, precision="fp32", device=0)
    ms = timed(lambda: car.step(100), 20)
    emit("toy_car_1M_k100", ms, M * 100, "env-iterations")
    chute = S.ToyParachute(n_envs=M, jitter=0.1, seed=0, precision="fp32", device=0)
    ms = timed(lambda: (chute.reset(), chute.step(100)), 20)
    emit("toy_parachute_1M_k100", ms, M * 100, "env-iterations")
    car.close(); chute.close()


def reset_cases(M):
    """BoatEnv.reset (boat_env.py:120-126, a new Boat + a new Wind per env; 2.2 ms per env in the reference):
    the explicit reset of a whole population -- one warp-cooperative wind setup per env -- and a masked one."""
    for exp, n in ((1, 16 * M), (6, 16 * M)):
        cfg = S.load_config(base_settings__experiment=exp)
        env = S.BatchedBoatEnv(cfg, n, seed=1, precision="fp32", device=0, auto_reset=True)
        ms = timed(lambda: env.reset(), 5, warmup=1)
        emit(f"reset_exp{exp}_fp32_16M", ms, n, "env-resets")
        if exp == 6:
            mask = (torch.arange(n, device="cuda") % 300 == 0).to(torch.uint8)   # the steady-state reset share of one step
            ms = timed(lambda: env.reset(mask), 10, warmup=2)
            emit("reset_exp6_fp32_16M_masked_1_in_300", ms, int(mask.sum().item()), "env-resets")
        env.close()


def replay_cases(M):
    """replay buffer: batched store, sample-gather (agent/buffer.py), fused step+store"""
    n, cap = 4 * M, 16 * M
    buf = S.ReplayBuffer(cap, (11,), 1, precision="fp32", device=0, as_torch=True)
    s = torch.randn(n, 11, device="cuda")
    a = torch.randn(n, 1, device="cuda")
    r = torch.randn(n, device="cuda")
    d = torch.zeros(n, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: buf.store_batch(s, a, r, s, d), 20)
    emit("replay_store_4M_rows", ms, n, "rows", 2 * 97, note="97 B read + 97 B written per row")
    for batch in (1024, 1 << 20):
        ms = timed(lambda: buf.sample_buffer(batch), 50)
        emit(f"replay_sample_gather_{batch}", ms, batch, "rows", 2 * 97,
             note="includes the torch.empty of the five output tensors; 97 B gathered + 97 B written per row")
    for batch in (1024, 1 << 20):  # the learner's way: five preallocated outputs, nothing but the gather kernel
        outs = buf.sample_buffer(batch)
        outs = (outs[0], outs[1], outs[2], outs[3], outs[4].to(torch.uint8))
        ms = timed(lambda: buf.sample_buffer(batch, out=outs), 50)
        emit(f"replay_sample_gather_{batch}_into_static_batch", ms, batch, "rows", 2 * 97,
             note="sample_buffer(out=...): 97 B gathered + 97 B written per row")
    cfg = S.load_config(base_settings__experiment=6)
    env = S.BatchedBoatEnv(cfg, 16 * M, seed=1, precision="fp32", device=0, auto_reset=True)
    env.reset()
    big = S.ReplayBuffer(32 * M, (11,), 1, precision="fp32", device=0, as_torch=True)
    acts = env.uniform_actions(0, 1.0)
    for t in range(300):
        env.uniform_actions(t, 1.0, out=acts)
        env.step(acts)
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
    for t, (e0, e1) in enumerate(pairs, start=300):
        env.uniform_actions(t, 1.0, out=acts)
        e0.record()
        big.step_store(env, acts)
        e1.record()
    torch.cuda.synchronize()
    ms = sum(e0.elapsed_time(e1) for e0, e1 in pairs) / len(pairs)
    emit("exp6_fp32_16M_fused_step_store", ms, 16 * M, "env-steps", 165 + 44 + 97,
         note="step (165 B) + previous obs read (44 B) + transition written (97 B)")
    env.close(); big.close(); buf.close()


if __name__ == "__main__":
    main()
