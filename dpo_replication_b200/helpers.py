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


class DeviceReplayBuffer:
    """helpers.py:33-72 with the ring buffer resident in HBM: `store` writes one
    transition row on the device, `sample_batch` draws indices exactly as the reference does
    (random.sample over range(size), without replacement) and gathers the five arrays with ONE kernel
    (gac_replay_gather) -- no per-step host->device copy of the batch (43 KB at HalfCheetah dims)."""

    def __init__(self, obs_dim, act_dim, size, device):
        import ctypes as C

        from . import _lib
        self._C, self._lib = C, _lib
        self.device = torch.device(device)
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.transitions = namedtuple("transition", ["s", "a", "r", "sp", "it"])
        z = lambda w: torch.zeros(size, w, dtype=torch.float32, device=self.device)
        self.obs1_buf, self.obs2_buf, self.acts_buf, self.rews_buf, self.done_buf = z(obs_dim), z(obs_dim), z(act_dim), z(1), z(1)
        self.ptr, self.size, self.max_size = 0, 0, size
        self.size_list = range(self.size)

    def store(self, obs, act, rew, next_obs, done):
        i = self.ptr
        row = lambda x, w: torch.as_tensor(np.asarray(x, dtype=np.float32).reshape(w)).to(self.device, non_blocking=True)
        self.obs1_buf[i], self.obs2_buf[i], self.acts_buf[i] = row(obs, self.obs_dim), row(next_obs, self.obs_dim), row(act, self.act_dim)
        self.rews_buf[i], self.done_buf[i] = float(rew), float(done)
        self.ptr = (i + 1) % self.max_size
        if self.size < self.max_size:
            self.size += 1
            self.size_list = range(self.size)

    def sample_batch(self, batch_size=32, device=None):
        C = self._C
        idxs = torch.tensor(random.sample(self.size_list, batch_size), dtype=torch.int32).to(self.device, non_blocking=True)
        new = lambda w: torch.empty(batch_size, w, dtype=torch.float32, device=self.device)
        t = self.transitions
        t.s, t.sp, t.a, t.r, t.it = new(self.obs_dim), new(self.obs_dim), new(self.act_dim), new(1), new(1)
        p = lambda x: C.c_void_p(x.data_ptr())
        self._lib.check(self._lib.lib.gac_replay_gather(p(self.obs1_buf), p(self.obs2_buf), p(self.acts_buf), p(self.rews_buf), p(self.done_buf),
                                                        p(idxs), batch_size, self.obs_dim, self.act_dim, p(t.s), p(t.sp), p(t.a), p(t.r), p(t.it),
                                                        C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        self._last_idx = idxs
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
    as_t = lambda v: torch.as_tensor(v, dtype=torch.float32, device=x.device) if isinstance(x, torch.Tensor) else v
    return x * as_t(stats.std) + as_t(stats.mean)
