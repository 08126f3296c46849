"""K = 8 fused sub-steps on 4 M envs of experiment 6 (fp32): the ncu target for the KMULTI kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sac_agent_b200 as S
cfg = S.load_config(base_settings__experiment=6)
env = S.BatchedBoatEnv(cfg, 4 << 20, seed=1, precision="fp32", device=0, auto_reset=True)
env.reset()
acts = env.uniform_actions(0, 1.0)
for t in range(60):
    env.uniform_actions(t, 1.0, out=acts)
    env.step_k(acts, 8)
torch.cuda.synchronize()
print("ok")
