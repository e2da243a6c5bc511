"""fp64 checker used by the full-size GPU parity tests: the torch mirror of the oracle (oracle/torch_ref.py, itself
checked against oracle/gac_oracle.py at 1e-9 in tests/test_oracle.py) evaluated in float64 ON THE GPU, where the
NumPy oracle would take minutes at BASELINE.json's sizes.  Test infrastructure only.

Besides the plain forwards it exposes, per row, the *kink margin* of a teacher-forced actor pass: the smallest
|pre-activation| in front of a LeakyReLU (relative to that tensor's largest entry) and the smallest |pred - target|
(the quantile-Huber indicator).  A gradient is discontinuous across those kinks, so two correct implementations whose
forwards differ by 1e-6 can legitimately disagree on a whole row's contribution; gradient parity at 1e-4 is therefore
asserted on rows whose margin exceeds the forward tolerance -- the rows are chosen by the checker BEFORE the CUDA path
sees them, and both sides then get exactly the same rows.
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import gac_oracle as O
from oracle import torch_ref as T

E = O.E


def tile_state_major(s, K):      # networks.py:464-467: row = b*K + k
    return s.unsqueeze(1).repeat(1, K, 1).reshape(-1, s.shape[1])


def tile_sample_major(s, K):     # agent.py:186: row = k*B + b
    return s.repeat(K, 1)


def params64(p, device, grad=False):
    return T.to_t(p, torch.float64, grad=grad, device=device)


def _cosb(x, dtype):
    ipi = torch.from_numpy(np.arange(1, O.NB + 1, dtype=np.float32) * np.float32(np.pi)).to(x.device)
    return torch.cos((x.reshape(-1, 1).to(torch.float32) * ipi).to(dtype))


def supervised_margins(p, state, taus, acts):
    """AutoRegressiveStochasticActor._supervised_forward (networks.py:227-268) -> (pred (rows, A), margin (rows,))."""
    rows, A = acts.shape
    dt = p["ws"].dtype
    pre_s = state @ p["ws"] + p["bs"]
    se = F.leaky_relu(F.leaky_relu(pre_s, 0.01), 0.2)
    shifted = torch.zeros_like(acts)
    shifted[:, 1:] = acts[:, :-1]
    pre_a = (_cosb(shifted, dt) @ p["wa"] + p["ba"]).reshape(rows, A, E)
    ae = F.leaky_relu(pre_a, 0.2)
    ne = (_cosb(taus, dt) @ p["wn"] + p["bn"]).reshape(rows, A, E)
    h = torch.zeros(rows, E, dtype=dt, device=state.device)
    hs = []
    for d in range(A):
        x = torch.cat([se, ae[:, d]], dim=1)
        mx = x @ p["kx"] + p["gb"][0]
        mh = h @ p["u"] + p["gb"][1]
        z = torch.sigmoid(mx[:, :E] + mh[:, :E])
        r = torch.sigmoid(mx[:, E:2 * E] + mh[:, E:2 * E])
        c = torch.tanh(mx[:, 2 * E:] + r * mh[:, 2 * E:])
        h = z * h + (1 - z) * c
        hs.append(h)
    y1p = (torch.stack(hs, dim=1) * ne) @ p["w1"] + p["b1"]
    pred = torch.tanh(F.leaky_relu(y1p, 0.01) @ p["w2"] + p["b2"]).squeeze(-1)
    with torch.no_grad():
        rel = lambda t, dims: t.abs().amin(dim=dims) / t.abs().max()
        margin = torch.minimum(rel(pre_s, 1), rel(y1p, (1, 2)))
        if A > 1:       # step 0 embeds the constant a_{-1} = 0: identical for every row, no row-dependent kink
            margin = torch.minimum(margin, rel(pre_a[:, 1:], (1, 2)))
        margin = torch.minimum(margin, (pred - acts).abs().amin(dim=1))
    return pred, margin


def actor_loss(pred, target, taus, w):
    """IQNActor.huber_quantile_loss (networks.py:97-120): elementwise Huber, mean over all M*A elements."""
    return T.huber_quantile_loss_t(pred, target, taus, w)


def actor_grads(p, state, taus, acts, w, chunk=8192):
    """d(loss)/d(13 actor variables) with the weights outside the tape (networks.py:135-140); accumulated over row
    chunks (the loss is a sum over rows once the weights are fixed) to bound the autograd memory."""
    M, A = acts.shape
    names = list(p.keys())
    tot = {k: torch.zeros_like(v) for k, v in p.items()}
    loss_sum = 0.0
    for r0 in range(0, M, chunk):
        sl = slice(r0, min(M, r0 + chunk))
        pred, _ = supervised_margins(p, state[sl], taus[sl], acts[sl])
        ind = ((pred - acts[sl]) > 0).to(pred.dtype)
        hub = F.huber_loss(pred, acts[sl], reduction="none", delta=1.0)
        loss = (torch.abs(taus[sl] - ind) * hub * w[sl].reshape(-1, 1)).sum() / (M * A)
        g = torch.autograd.grad(loss, [p[k] for k in names])
        for k, gk in zip(names, g):
            tot[k] += gk
        loss_sum += float(loss.detach())
    return tot, loss_sum


def fnn_margin(p, x):
    """smallest relative |pre-activation| in front of the two LeakyReLUs of an FNN, per row."""
    with torch.no_grad():
        a0 = x @ p["w0"] + p["b0"]
        a1 = F.leaky_relu(a0, 0.01) @ p["w1"] + p["b1"]
        return torch.minimum(a0.abs().amin(1) / a0.abs().max(), a1.abs().amin(1) / a1.abs().max())


def td_grads(p, x, y):
    """loss_b = (y_b - f(x_b))^2, gradient of the SUM (networks.py:418-426 / 493-496, F5)."""
    loss = ((y.reshape(-1, 1) - T.fnn_t(p, x)) ** 2).mean(dim=-1).sum()
    g = torch.autograd.grad(loss, list(p.values()))
    return dict(zip(p.keys(), g)), float(loss.detach())
