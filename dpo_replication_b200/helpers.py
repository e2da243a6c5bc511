"""GAC/helpers.py surface: ActionSampler, ReplayBuffer, update, normalize, denormalize."""
from __future__ import annotations

import random
from collections import namedtuple

import numpy as np
import torch

from .engine import f32
from .networks import _engine_for


class ActionSampler:
    """helpers.py:9-30: tau ~ U[0,1) then actor(states, tau, actions)."""

    def __init__(self, action_dim):
        self.dim = action_dim

    def get_actions(self, actor, states, actions=None, *, taus=None):
        states = f32(states, actor.device)
        batch_size = states.shape[0] if states.dim() > 1 else 1          # helpers.py:22-25
        if taus is None:
            taus = torch.rand(batch_size, self.dim, device=actor.device)  # helpers.py:28
        return actor(states, taus, actions)


class ReplayBuffer:
    """helpers.py:33-72: FIFO NumPy ring buffer, sampling without replacement.  Host-side
    feeder (a device-resident replay is the next row of SURVEY.md §8f)."""

    def __init__(self, obs_dim, act_dim, size):
        self.transitions = namedtuple("transition", ["s", "a", "r", "sp", "it"])
        self.obs1_buf = np.zeros([size, obs_dim], dtype=np.float32)
        self.obs2_buf = np.zeros([size, obs_dim], dtype=np.float32)
        self.acts_buf = np.zeros([size, act_dim], dtype=np.float32)
        self.rews_buf = np.zeros([size, 1], dtype=np.float32)
        self.done_buf = np.zeros([size, 1], dtype=np.float32)
        self.ptr, self.size, self.max_size = 0, 0, size
        self.size_list = range(self.size)

    def store(self, obs, act, rew, next_obs, done):
        i = self.ptr
        self.obs1_buf[i], self.obs2_buf[i], self.acts_buf[i] = obs, next_obs, act
        self.rews_buf[i], self.done_buf[i] = rew, done
        self.ptr = (i + 1) % self.max_size
        if self.size < self.max_size:
            self.size += 1
            self.size_list = range(self.size)

    def sample_batch(self, batch_size=32, device=None):
        idxs = random.sample(self.size_list, batch_size)
        conv = (lambda x: torch.from_numpy(x).to(device, non_blocking=True)) if device is not None else torch.from_numpy
        t = self.transitions
        t.s, t.a, t.r = conv(self.obs1_buf[idxs]), conv(self.acts_buf[idxs]), conv(self.rews_buf[idxs])
        t.sp, t.it = conv(self.obs2_buf[idxs]), conv(self.done_buf[idxs])
        return t


def update(target, source, rate):
    """helpers.py:75-86: t <- t*(1-rate) + s*rate for every variable pair (rate = 1 => copy).
    One 128-bit streaming kernel per network slice (gac_polyak) instead of 3 ops per variable."""
    for net in target.net_names:
        t = target.arena.net_slice(target._buf, net)
        s = source.arena.net_slice(source._buf, net)
        eng = _engine_for(target.arena.S, target.arena.A, 128, target.device)
        eng.polyak(t, s, float(rate))


def normalize(x, stats):
    """helpers.py:89-95."""
    if stats is None:
        return x
    return (x - torch.as_tensor(stats.mean, dtype=torch.float32, device=x.device)) / torch.sqrt(
        torch.as_tensor(stats.var, dtype=torch.float32, device=x.device))


def denormalize(x, stats):
    """helpers.py:98-101."""
    if stats is None:
        return x
    return x * stats.std + stats.mean
