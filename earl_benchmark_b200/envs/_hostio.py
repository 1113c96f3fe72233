"""Pinned host buffers of the numpy / CPU-tensor step path (`earl_*_step_host`), shared by the env classes.

The reference returns FRESH arrays from every `step()`, so the canonical loop

    next_obs, r, done, _ = env.step(a); buffer.add(obs, a, r, next_obs); obs = next_obs

may hold `obs` across the next call.  Returning views of ONE pinned output set would make `obs` and `next_obs` the same
memory from the second step on.  Copying 56 MB per step at 1M envs would cost more than the PCIe transfer
it follows, so the outputs alternate between TWO pinned sets instead: the arrays a step returns stay untouched during
the NEXT step and are overwritten by the one after it ("valid until the step after next").  Anything kept longer must be
copied by the caller, as a replay buffer's `add` does.
"""
import numpy as np
import torch


class HostBuffers:
    def __init__(self, n, act_dim, obs_dim):
        pin = dict(pin_memory=True)
        self.n, self.act_dim, self.obs_dim = n, act_dim, obs_dim
        self.act = torch.empty((n, act_dim), dtype=torch.float32, **pin)
        self.sets = [(torch.empty((n, obs_dim), dtype=torch.float32, **pin), torch.empty((n,), dtype=torch.float32, **pin),
                      torch.empty((n,), dtype=torch.uint8, **pin), torch.empty((n,), dtype=torch.uint8, **pin)) for _ in range(2)]
        self.turn = 0

    def stage(self, action):
        """Pinned float32 [n, act_dim] tensor holding `action` (the caller's own tensor when it already is one)."""
        if isinstance(action, torch.Tensor):
            if action.is_pinned() and action.dtype == torch.float32 and action.is_contiguous() and action.numel() == self.n * self.act_dim:
                return action
            self.act.copy_(action.reshape(self.n, self.act_dim))
        else:
            self.act.numpy()[...] = np.asarray(action, np.float32).reshape(self.n, self.act_dim)
        return self.act

    def next_outputs(self):
        self.turn ^= 1
        return self.sets[self.turn]

    @staticmethod
    def as_numpy(ho, hr, hd, hs):
        return ho.numpy(), hr.numpy(), hd.numpy().view(np.bool_), {"success": hs.numpy().view(np.bool_)}
