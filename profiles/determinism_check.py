"""Run-to-run determinism of the whole-launch K = 1 step path: two identical handles stepped alternately with the
same actions must agree bit for bit (several state blocks per persistent warp, frequent resets).  This is the check
that exposed the stage-refill race (a bulk copy served from L2 landing before the last shared-memory loads of the
stage had returned): 16 of 40 trials failed before the vote-based dependency in boat_step.cuh, 0 of 40 after.
    python profiles/determinism_check.py <trials> [experiment] [precision] [step|fused|k8]"""
import sys; sys.path.insert(0,'.')
import torch
import sac_agent_b200 as S
trials = int(sys.argv[1])
experiment = int(sys.argv[2]) if len(sys.argv) > 2 else 6
precision = sys.argv[3] if len(sys.argv) > 3 else "fp32"
mode = sys.argv[4] if len(sys.argv) > 4 else "step"       # step | fused (step + replay store) | k8 (8 fused sub-steps)
cfg = S.load_config(base_settings__experiment=experiment)
n = 300_003; fails = 0
def mk():
    e = S.BatchedBoatEnv(cfg, n, seed=2, precision=precision, device=0, auto_reset=True); e.reset(); return e
def advance(e, ring, acts, t):
    if mode == "fused":
        ring.step_store(e, acts); return e.obs
    if mode == "k8":
        return e.step_k(torch.stack([e.uniform_actions(8 * t + q, 4.0) for q in range(8)]), 8)[0]
    return e.step(acts)[0]
for trial in range(trials):
    a, ref = mk(), mk()
    ra = rb = None
    if mode == "fused":
        ra = S.ReplayBuffer(1_000_001, (11,), 1, precision=precision, device=0, as_torch=True)
        rb = S.ReplayBuffer(1_000_001, (11,), 1, precision=precision, device=0, as_torch=True)
    bad = False
    for t in range(14):
        acts = a.uniform_actions(t, 4.0)
        o = advance(a, ra, acts, t).clone(); torch.cuda.synchronize()
        oref = advance(ref, rb, acts, t); torch.cuda.synchronize()
        if (o != oref).any():
            bad = True; break
    if mode == "fused" and not bad:
        idx = torch.arange(0, 1_000_001, 7, device="cuda")
        bad = any(not torch.equal(x, y) for x, y in zip(ra.gather(idx), rb.gather(idx)))
    fails += bad
    for x in (a, ref, ra, rb):
        if x is not None: x.close()
print(mode, "experiment", experiment, precision, ": failing trials", fails, "of", trials)
