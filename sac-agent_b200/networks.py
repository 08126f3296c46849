"""The three SAC-v1 networks of the reference (networks/networks.py:14-133), device-resident and
batched over envs.  Layer names, shapes and the order they are created in are the reference's
(`fc1, fc2, mean, std` / `fc1, fc2, q` / `fc1, fc2, v`, all 256 wide), so a `state_dict` saved by the
reference's `BaseNetwork.save_checkpoint` (networks/base_network.py:13-17) loads here unchanged and
vice versa.  Plain dense layers: cuBLAS is the right tool for them, there is no custom kernel here;
the B200-specific part of the learner is how it is driven (continuous_agent.py: one CUDA graph per
update, batch gathered on the device by libboatenv's replay kernels).
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

LOG_STD_MAX, LOG_STD_MIN = 2.0, -5.0   # networks.py:48-49
REPARAM_NOISE = 1e-6                   # networks.py:22
_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


class BaseNetwork(nn.Module):
    """networks/base_network.py:5-17 -- checkpoint path = <experiment_dir>/checkpoints/<name>."""

    def __init__(self, name, experiment_dir):
        super().__init__()
        self.name = name
        self.experiment_dir = experiment_dir
        self.checkpoints_dir = os.path.join(experiment_dir, "checkpoints") if experiment_dir is not None else None
        self.checkpoint_file = os.path.join(self.checkpoints_dir, name) if experiment_dir is not None else None

    def save_checkpoint(self):
        os.makedirs(self.checkpoints_dir, exist_ok=True)
        torch.save(self.state_dict(), self.checkpoint_file)

    def load_checkpoint(self):
        dev = next(self.parameters()).device
        self.load_state_dict(torch.load(self.checkpoint_file, map_location=dev))


class ActorNetwork(BaseNetwork):
    """networks.py:14-70: tanh-squashed Gaussian policy; the std head goes through tanh into
    log_std in [-5, 2]."""

    def __init__(self, experiment_dir, input_dims, max_action, fc1_dims=256, fc2_dims=256, n_actions=1,
                 name="actor_network"):
        super().__init__(name, experiment_dir)
        self.fc1 = nn.Linear(*input_dims, fc1_dims)
        self.fc2 = nn.Linear(fc1_dims, fc2_dims)
        self.mean = nn.Linear(fc2_dims, n_actions)
        self.std = nn.Linear(fc2_dims, n_actions)
        self.register_buffer("max_action", torch.as_tensor(max_action, dtype=torch.float32).reshape(-1), persistent=False)

    def forward(self, state):
        prob = F.relu(self.fc1(state))
        prob = F.relu(self.fc2(prob))
        return self.mean(prob), self.std(prob)

    def sample_normal(self, state, reparameterize=True, eps=None):
        """networks.py:47-70.  `eps` (standard-normal, shape [B, n_actions]) replaces the draw -- the
        parity tests inject the reference's noise through it."""
        mean, std = self.forward(state)
        log_std = LOG_STD_MIN + 0.5 * (LOG_STD_MAX - LOG_STD_MIN) * (torch.tanh(std) + 1.0)
        std = log_std.exp()
        if eps is None:
            eps = torch.randn_like(mean)
        if reparameterize:
            u = mean + eps * std              # Normal.rsample
        else:
            with torch.no_grad():             # Normal.sample: the draw carries no gradient
                u = mean + eps * std
        action = torch.tanh(u) * self.max_action
        # Normal.log_prob(u) - log(1 - action^2 + 1e-6), summed over the action dimensions
        log_probs = -((u - mean) ** 2) / (2.0 * std * std) - log_std - _HALF_LOG_2PI
        log_probs = log_probs - torch.log(1.0 - action.pow(2) + REPARAM_NOISE)
        return action, log_probs.sum(1, keepdim=True)


class CriticNetwork(BaseNetwork):
    """networks.py:73-104: Q(s, a) on cat([state, action])."""

    def __init__(self, experiment_dir, input_dims, n_actions, fc1_dims=256, fc2_dims=256, name="critic_network"):
        super().__init__(name, experiment_dir)
        self.fc1 = nn.Linear(input_dims[0] + n_actions, fc1_dims)
        self.fc2 = nn.Linear(fc1_dims, fc2_dims)
        self.q = nn.Linear(fc2_dims, 1)

    def forward(self, state, action):
        x = F.relu(self.fc1(torch.cat([state, action], dim=1)))
        x = F.relu(self.fc2(x))
        return self.q(x)


class ValueNetwork(BaseNetwork):
    """networks.py:107-133: V(s)."""

    def __init__(self, experiment_dir, input_dims, fc1_dims=256, fc2_dims=256, name="value_network"):
        super().__init__(name, experiment_dir)
        self.fc1 = nn.Linear(*input_dims, fc1_dims)
        self.fc2 = nn.Linear(fc1_dims, fc2_dims)
        self.v = nn.Linear(fc2_dims, 1)

    def forward(self, state):
        x = F.relu(self.fc1(state))
        x = F.relu(self.fc2(x))
        return self.v(x)
