"""Small all-paths workload for compute-sanitizer (memcheck): both precisions, inline and queued slow
paths, K-fused steps, fused replay store, sample-gather, toys, host-buffer step, checkpoint."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sac_agent_b200 as S

for exp in (2, 3, 5, 6):
    for precision in ("fp32", "fp64"):
        cfg = S.load_config(base_settings__experiment=exp, base_settings__t_max=120)
        env = S.BatchedBoatEnv(cfg, 1000, seed=exp, precision=precision, device=0, auto_reset=True)
        buf = S.ReplayBuffer(4096, (11,), 1, precision=precision, device=0, as_torch=True)
        env.reset()
        for t in range(60):
            a = env.uniform_actions(t, 3.0)
            if t % 3 == 0:
                env.step(a)
            elif t % 3 == 1:
                env.step_k(a, 4)
            else:
                buf.step_store(env, a)
        buf.sample_buffer(256)
        sd = env.state_dict()
        env.load_state_dict(sd)
        env.wind_table(7)
        env.get_field("s_x"); env.counters()
        h = [torch.empty(1000, dtype=env.dtype).pin_memory(), torch.empty((1000, 11), dtype=env.dtype).pin_memory(),
             torch.empty(1000, dtype=env.dtype).pin_memory(), torch.empty(1000, dtype=torch.uint8).pin_memory()]
        h[0].zero_()
        env.step_host(*h)
        env.close(); buf.close()
car = S.ToyCar(n_envs=500, jitter=0.1, precision="fp32", device=0); car.step(50); car.close()
ch = S.ToyParachute(n_envs=500, jitter=0.1, precision="fp64", device=0); ch.step(50); ch.close()
torch.cuda.synchronize()
print("sanitize case ok", S.lib().boatenv_kernel_launches(), "launches")
