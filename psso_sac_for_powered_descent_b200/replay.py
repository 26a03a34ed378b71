"""Device-resident replay buffer fed directly by the batched collection kernels.

Same interface as the reference's uniform `ReplayBuffer` (src/agents/sac_pytorch.py:12-47:
`add`, `sample`, `__len__`, attributes `capacity / state_dim / action_dim / position / size`),
so the stock `SACPyTorch.update` (sac_pytorch.py:411-436) consumes it unchanged - its
`.to(self.device)` calls are no-ops on tensors that already live on the GPU.  What is new is
`reserve(n)`: it hands out contiguous slots of the ring so that
`BatchedRocketEnv.collect(actor, n_steps, into=buffer)` lets `pd_collect_shared_actor` write
observations, actions, rewards, done flags and next observations of n_steps x n_envs
transitions straight into the buffer's storage (no staging copy, no host round trip).

As in the reference's collection loop (sac_pytorch_powered_descent.py:167-173) the stored `done`
is 1.0 for a successful landing only; truncations end the episode but are not terminal for the
critic's bootstrap.
"""
from __future__ import annotations

import torch


class DeviceReplayBuffer:
    def __init__(self, capacity: int, state_dim: int, action_dim: int, device="cuda"):
        self.capacity, self.state_dim, self.action_dim = int(capacity), int(state_dim), int(action_dim)
        self.device = torch.device(device)
        self.position = 0
        self.size = 0
        z = lambda *shape, dt=torch.float32: torch.zeros(*shape, dtype=dt, device=self.device)
        self.states = z(self.capacity, state_dim)
        self.actions = z(self.capacity, action_dim)
        self.rewards = z(self.capacity, 1)
        self.next_states = z(self.capacity, state_dim)
        self.dones = z(self.capacity, 1)
        self._done_u8 = z(self.capacity, dt=torch.uint8)      # the kernels write uint8 flags
        self._gen = torch.Generator(device=self.device)

    # ------------------------------------------------------------------ reference interface
    def add(self, state, action, reward, next_state, done):
        i = self.position
        f = lambda x: torch.as_tensor(x, dtype=torch.float32, device=self.device).reshape(-1)
        self.states[i] = f(state)
        self.actions[i] = f(action)
        self.rewards[i, 0] = float(reward)
        self.next_states[i] = f(next_state)
        self.dones[i, 0] = float(done)
        self.position = (self.position + 1) % self.capacity
        self.size = min(self.size + 1, self.capacity)

    def sample(self, batch_size: int):
        idx = torch.randint(0, self.size, (batch_size,), device=self.device, generator=self._gen)
        return (self.states[idx], self.actions[idx], self.rewards[idx], self.next_states[idx],
                self.dones[idx])

    def __len__(self):
        return self.size

    # ------------------------------------------------------------------ batched producer side
    def reserve(self, n: int):
        """Contiguous slots for n transitions: (start, views dict).  A block that would run over
        the end of the ring starts again at slot 0 (the tail is overwritten on the next lap)."""
        if n > self.capacity:
            raise ValueError(f"cannot reserve {n} transitions in a buffer of {self.capacity}")
        start = self.position if self.position + n <= self.capacity else 0
        sl = slice(start, start + n)
        return start, dict(obs=self.states[sl], actions=self.actions[sl], rewards=self.rewards[sl],
                           next_obs=self.next_states[sl], done=self._done_u8[sl])

    def commit(self, start: int, n: int):
        """Publish n transitions written into the slots `reserve` handed out."""
        sl = slice(start, start + n)
        self.dones[sl, 0] = self._done_u8[sl].to(torch.float32)
        self.position = (start + n) % self.capacity
        self.size = min(max(self.size, start + n), self.capacity)

    def add_batch(self, obs, actions, rewards, next_obs, done):
        """Append a batch that already lives in tensors (any leading shape)."""
        n = obs.reshape(-1, self.state_dim).shape[0]
        start, v = self.reserve(n)
        v["obs"].copy_(obs.reshape(n, self.state_dim))
        v["actions"].copy_(actions.reshape(n, self.action_dim))
        v["rewards"].copy_(rewards.reshape(n, 1))
        v["next_obs"].copy_(next_obs.reshape(n, self.state_dim))
        v["done"].copy_(done.reshape(n).to(torch.uint8))
        self.commit(start, n)
        return start
