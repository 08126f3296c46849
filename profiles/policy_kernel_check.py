"""Check and time boatagent_policy_act (csrc/policy_mlp.cu): against the fp32 policy and against the same arithmetic
spelled out in PyTorch (bf16 operands, fp32 accumulation, fp32 heads), then 1,048,576 envs timed."""
import sys, time
sys.path.insert(0,'/root/repo')
import torch, numpy as np
from sac_agent_b200.networks import ActorNetwork, TensorCorePolicy
torch.manual_seed(0)
actor = ActorNetwork(None, (11,), np.array([1.0],dtype=np.float32), n_actions=1).cuda()
pol = TensorCorePolicy(actor)
for n in (1, 100, 128, 1000, 65536):
    obs = torch.rand(n, 11, device='cuda')
    eps = torch.randn(n, 1, device='cuda')
    a = pol.act(obs, eps)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = actor.sample_normal(obs, reparameterize=False, eps=eps)[0]
        # bf16-input emulation
        bf = lambda t: t.to(torch.bfloat16).float()
        h = bf(torch.relu(bf(obs) @ bf(actor.fc1.weight).T + actor.fc1.bias))
        h = torch.relu(h @ bf(actor.fc2.weight).T + actor.fc2.bias)
        mean = h @ actor.mean.weight.T + actor.mean.bias
        raw = h @ actor.std.weight.T + actor.std.bias
        emu = torch.tanh(mean + eps*torch.exp(-5+3.5*(torch.tanh(raw)+1)))
    print(n, 'vs fp32', (a-ref).abs().max().item(), 'vs bf16 emu', (a-emu).abs().max().item(), flush=True)
n=1<<20
obs=torch.rand(n,11,device='cuda'); out=torch.empty(n,1,device='cuda')
for _ in range(3): pol.act(obs, None, out)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(20): pol.act(obs, None, out)
torch.cuda.synchronize(); print('1M envs ms', (time.perf_counter()-t)/20*1e3, 'mean', out.mean().item(), 'std', out.std().item())
