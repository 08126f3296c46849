"""K = 8 fused sub-steps on 16 M envs of experiment 6 (fp32), a fresh policy action per sub-step (same
episode statistics as K = 1): the ncu target for the KMULTI kernel and its follow-up setup kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sac_agent_b200 as S
cfg = S.load_config(base_settings__experiment=6)
n, K = 16 << 20, 8
env = S.BatchedBoatEnv(cfg, n, seed=1, precision="fp32", device=0, auto_reset=True)
env.reset()
acts = torch.empty((K, n), dtype=torch.float32, device="cuda")
for t in range(60):
    for q in range(K):
        env.uniform_actions(t * K + q, 1.0, out=acts[q])
    env.step_k(acts, K)
torch.cuda.synchronize()
print("ok")
