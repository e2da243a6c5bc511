"""Reference-equivalent CPU path (TensorFlow is not installable here): a torch-CPU *eager*,
op-for-op mirror of the reference's GAC train step, used

  * as an independent second implementation to validate ``oracle/gac_oracle.py``
    (torch.nn.functional GRU-cell algebra + autograd instead of hand-derived backward), and
  * as the timed CPU baseline (``bench.py`` ``cpu_baseline`` / ``--impl reference``): same
    Python A-step loop (networks.py:208), same two tilings (networks.py:464-467, agent.py:186),
    same redundant work (state embedding on all tiled rows, target_value on tiled rows,
    per-variable Adam and Polyak), no fusion, no hoisting.

TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product package.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from . import gac_oracle as O


def to_t(p, dtype=torch.float32, grad=False, device="cpu"):
    """Parameter dict -> tensors.  dtype / device generic: fp32 on the CPU is the timed baseline, fp64 on a GPU is the
    checker the full-size parity tests use (tests/test_fullsize_oracle_gpu.py)."""
    return {k: torch.tensor(np.asarray(v), dtype=dtype, device=device).requires_grad_(grad) for k, v in p.items()}


class CosineBasisLinearT:
    """networks.py:17-62 including the per-call tiled (rows, n) constant (:45-46)."""

    def __init__(self, w, b, act=None):
        self.w, self.b, self.act = w, b, act
        self.n = w.shape[0]

    def __call__(self, x):
        rows = x.shape[0]
        col = x.reshape(-1, 1).to(torch.float32)
        if col.device.type == "cpu":       # the timed CPU path keeps the reference's per-call host-side tile (networks.py:45-46)
            ipi = torch.from_numpy(np.tile(np.arange(1, self.n + 1, dtype=np.float32), (col.shape[0], 1)) * np.float32(np.pi))
        else:                              # checker on a device: same fp32 constants fl32(i) * fl32(pi), broadcast
            ipi = torch.from_numpy(np.arange(1, self.n + 1, dtype=np.float32) * np.float32(np.pi)).to(col.device)
        emb = torch.cos((col * ipi).to(self.w.dtype))
        out = emb @ self.w + self.b
        if self.act is not None:
            out = F.leaky_relu(out, self.act)
        return out.reshape(rows, -1, self.w.shape[1])


class AIQNT:
    """AutoRegressiveStochasticActor (networks.py:144-268) on a parameter dict of tensors."""

    def __init__(self, p, action_dim):
        self.p, self.action_dim = p, action_dim
        self.noise_embedding = CosineBasisLinearT(p["wn"], p["bn"])
        self.action_embedding = CosineBasisLinearT(p["wa"], p["ba"])

    def _state(self, s):
        return F.leaky_relu(F.leaky_relu(s @ self.p["ws"] + self.p["bs"], 0.01), 0.2)

    def _gru(self, x, h):
        p = self.p
        mx = x @ p["kx"] + p["gb"][0]
        mh = h @ p["u"] + p["gb"][1]
        E = O.E
        z = torch.sigmoid(mx[:, :E] + mh[:, :E])
        r = torch.sigmoid(mx[:, E:2 * E] + mh[:, E:2 * E])
        c = torch.tanh(mx[:, 2 * E:] + r * mh[:, 2 * E:])
        return z * h + (1 - z) * c

    def _head(self, x):
        p = self.p
        return torch.tanh(F.leaky_relu(x @ p["w1"] + p["b1"], 0.01) @ p["w2"] + p["b2"])

    def __call__(self, state, taus, supervise_actions=None):
        if supervise_actions is not None:
            return self._supervised(state, taus, supervise_actions)
        rows = state.shape[0]
        se = self._state(state)
        ne = self.noise_embedding(taus)
        a = torch.zeros(rows, dtype=state.dtype, device=state.device)
        h = torch.zeros(rows, O.E, dtype=state.dtype, device=state.device)
        outs = []
        for d in range(self.action_dim):
            ae = F.leaky_relu(self.action_embedding(a.reshape(rows, 1, 1)), 0.2)[:, 0, :]
            h = self._gru(torch.cat([se, ae], dim=1), h)
            a = self._head(h * ne[:, d, :])
            outs.append(a)
            a = a[:, 0]
        return torch.stack(outs, dim=1).squeeze(-1)

    def _supervised(self, state, taus, acts):
        rows = state.shape[0]
        A = self.action_dim
        acts = acts.reshape(rows, A)
        se = self._state(state)
        shifted = torch.zeros_like(acts)
        shifted[:, 1:] = acts[:, :-1]
        ae = F.leaky_relu(self.action_embedding(shifted), 0.2)
        ne = self.noise_embedding(taus)
        h = torch.zeros(rows, O.E, dtype=state.dtype, device=state.device)
        hs = []
        for d in range(A):
            h = self._gru(torch.cat([se, ae[:, d]], dim=1), h)
            hs.append(h)
        gru_out = torch.stack(hs, dim=1)
        return self._head(gru_out * ne).squeeze(-1)


class IQNT:
    """StochasticActor (networks.py:271-323) on a parameter dict of tensors."""

    def __init__(self, p, action_dim):
        self.p, self.action_dim = p, action_dim
        self.noise_embedding_layer = CosineBasisLinearT(p["wn"], p["bn"], act=0.01)

    def __call__(self, states, taus, supervise_actions=None):
        p = self.p
        se = F.leaky_relu(states @ p["ws"] + p["bs"], 0.01)
        ne = self.noise_embedding_layer(taus).reshape(states.shape[0], -1)
        me = F.leaky_relu((se * ne) @ p["wm"] + p["bm"], 0.01)
        return torch.tanh(me @ p["wo"] + p["bo"])


def fnn_t(p, x):
    h = F.leaky_relu(x @ p["w0"] + p["b0"], 0.01)
    h = F.leaky_relu(h @ p["w1"] + p["b1"], 0.01)
    return h @ p["w2"] + p["b2"]


def critic_t(p1, p2, s, a):
    x = torch.cat([s, a], dim=-1)
    return torch.minimum(fnn_t(p1, x), fnn_t(p2, x))


def huber_quantile_loss_t(pred, target, taus, w):
    ind = ((pred - target) > 0).to(pred.dtype)
    hub = F.huber_loss(pred, target, reduction="none", delta=1.0)
    return (torch.abs(taus - ind) * hub * w).mean()


def adam_t_(p, g, m, v, t, lr=1e-4, b1=0.9, b2=0.999, eps=1e-7):
    alpha = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    m.add_((g - m) * (1 - b1))
    v.add_((g * g - v) * (1 - b2))
    p.sub_(alpha * m / (torch.sqrt(v) + eps))


class TorchGAC:
    """Holds six nets + Adam slots as torch tensors; ``train_one_step`` mirrors agent.py:121-168."""

    NETS = ("actor", "fnn1", "fnn2", "value")

    def __init__(self, nets, A, K, gamma=0.99, rho=5e-3, q_normalization=0.01, mode="linear", beta=1.0,
                 dtype=torch.float32, device="cpu"):
        self.A, self.K, self.gamma, self.rho, self.qn, self.mode, self.beta = A, K, gamma, rho, q_normalization, mode, beta
        self.dtype, self.device = dtype, device
        self.n = {}
        for k in self.NETS:
            self.n[k] = to_t(nets[k], dtype, grad=True, device=device)
            self.n["target_" + k] = to_t(nets["target_" + k], dtype, device=device)
            self.n["m_" + k] = to_t(nets["m_" + k], dtype, device=device)
            self.n["v_" + k] = to_t(nets["v_" + k], dtype, device=device)
        self.t = {k: int(nets["t_" + k]) for k in self.NETS}

    def _apply(self, name, loss_vec_sum):
        params = list(self.n[name].values())
        grads = torch.autograd.grad(loss_vec_sum, params)
        self.t[name] += 1
        with torch.no_grad():
            for (k, p), g in zip(self.n[name].items(), grads):
                adam_t_(p, g, self.n["m_" + name][k], self.n["v_" + name][k], self.t[name])
        return {k: g for k, g in zip(self.n[name].keys(), grads)}

    def sample_positive(self, s, noise):
        n, K, dt = self.n, self.K, self.dtype
        T = lambda a: torch.as_tensor(a, dtype=dt, device=self.device)
        tiled = s.repeat(K, 1)
        tactor = AIQNT(n["target_actor"], self.A)
        with torch.no_grad():
            ta = tactor(tiled, T(noise["filter_tau"]))
            ta = torch.clamp(ta + T(noise["filter_normal"]) * 0.01, -1, 1)
            tq = critic_t(n["target_fnn1"], n["target_fnn2"], tiled, ta)
            ra = T(noise["filter_rand"])
            rq = critic_t(n["target_fnn1"], n["target_fnn2"], tiled, ra)
            acts = torch.cat([ta, ra], 0)
            q = torch.cat([tq, rq], 0)
            v = fnn_t(n["target_value"], tiled)
            v = torch.cat([v, v], 0)
            tiled2 = torch.cat([tiled, tiled], 0)
            idx = torch.nonzero(q.squeeze() > v.squeeze())
            i = idx[:, 0]
        return tiled2[i], acts[i], (q - v)[i], dict(q=q, v=v, idx=idx)

    def train_one_step(self, batch, noise):
        n, K, dt = self.n, self.K, self.dtype
        T = lambda a: torch.as_tensor(a, dtype=dt, device=self.device)
        s, a, r, sp, it = (T(batch[k]) for k in ("s", "a", "r", "sp", "it"))
        B = s.shape[0]
        info = {}
        # critics.train
        nz = (T(noise["critic_u"]) * 2 - 1) * self.qn
        ab = torch.clamp(a + nz, -1, 1)
        with torch.no_grad():
            yq = r + self.gamma * fnn_t(n["target_value"], sp) * (1 - it)
        x = torch.cat([s, ab], dim=-1)
        for name in ("fnn1", "fnn2"):
            loss = ((yq - fnn_t(n[name], x)) ** 2).mean(dim=-1)
            info["grad_" + name] = self._apply(name, loss.sum())
        # value.train
        tiled = s.unsqueeze(1).repeat(1, K, 1).reshape(-1, s.shape[1])
        with torch.no_grad():
            va = AIQNT(n["target_actor"], self.A)(tiled, T(noise["value_tau"]))
            q = critic_t(n["target_fnn1"], n["target_fnn2"], tiled, va).reshape(-1, K, 1)
            v_critic = q.mean(dim=1)
        loss = ((v_critic - fnn_t(n["value"], s)) ** 2).mean(dim=-1)
        info["grad_value"] = self._apply("value", loss.sum())
        # filter + actor.train
        gs, ga, adv, finfo = self.sample_positive(s, noise)
        info["filter"] = finfo
        M = adv.shape[0]
        info["M"] = M
        if M:
            taus = T(noise["actor_tau"][:M])
            if self.mode == "linear":
                w = adv / adv.sum()
            elif self.mode == "boltzmann":
                w = torch.softmax(adv / self.beta, dim=-1)
            else:
                raise NotImplementedError
            pred = AIQNT(n["actor"], self.A)(gs, taus, ga)
            loss = huber_quantile_loss_t(pred, ga, taus, w)
            info["loss_actor"] = float(loss.detach())
            info["grad_actor"] = self._apply("actor", loss)
        # Polyak, one variable at a time (helpers.py:83-86)
        with torch.no_grad():
            for name in self.NETS:
                for k, tv in n["target_" + name].items():
                    tv.copy_(tv * (1.0 - self.rho) + n[name][k] * self.rho)
        return info
