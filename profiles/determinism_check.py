"""Run-to-run determinism of the whole-launch K = 1 step path: two identical handles stepped alternately with the
same actions must agree bit for bit (several state blocks per persistent warp, frequent resets).  This is the check
that exposed the stage-refill race (a bulk copy served from L2 landing before the last shared-memory loads of the
stage had returned): 16 of 40 trials failed before the vote-based dependency in boat_step.cuh, 0 of 40 after.
    python profiles/determinism_check.py <trials> [experiment] [precision]"""
import sys; sys.path.insert(0,'.')
import torch
import sac_agent_b200 as S
cfg = S.load_config(base_settings__experiment=int(sys.argv[2]) if len(sys.argv)>2 else 6)
n=300_000; trials=int(sys.argv[1]); fails=0
def mk():
    e=S.BatchedBoatEnv(cfg, n, seed=2, precision=sys.argv[3] if len(sys.argv)>3 else "fp32", device=0, auto_reset=True); e.reset(); return e
for trial in range(trials):
    a=mk(); ref=mk(); bad_any=False
    for t in range(14):
        acts=a.uniform_actions(t,4.0)
        o=a.step(acts)[0]; torch.cuda.synchronize()
        oref=ref.step(acts)[0]; torch.cuda.synchronize()
        if (o!=oref).any(): bad_any=True; break
    fails+=bad_any; a.close(); ref.close()
print('failing trials',fails,'of',trials)
