"""The fp64 checker of the full-size GPU tests (tests/fp64_checker.py) against the NumPy oracle at small sizes, on the CPU:
same predictions, gradients and losses to 1e-9, and the kink margin flags exactly the rows it should."""
import numpy as np
import torch

import fp64_checker as CK
from oracle import gac_oracle as O


def _setup(S=5, A=3, M=40, seed=0):
    rng = np.random.default_rng(seed)
    p = O.init_actor(rng, S, A, bias_scale=0.05)
    s = rng.standard_normal((M, S)).astype(np.float32)
    taus = rng.uniform(size=(M, A)).astype(np.float32)
    acts = rng.uniform(-1, 1, (M, A)).astype(np.float32)
    adv = rng.uniform(0.1, 1, (M, 1)).astype(np.float32)
    return p, s, taus, acts, adv


def test_checker_matches_oracle_actor():
    p, s, taus, acts, adv = _setup()
    g_ref, loss_ref, pred_ref, _ = O.actor_train_grads(p, s, acts, adv, taus, dtype=np.float64)
    p64 = CK.params64(p, "cpu", grad=True)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    pred, margin = CK.supervised_margins(p64, t(s), t(taus), t(acts))
    assert np.abs(pred.detach().numpy() - pred_ref).max() < 1e-12
    assert margin.shape == (s.shape[0],) and bool((margin > 0).all())
    w = t(adv) / t(adv).sum()
    for chunk in (7, 1000):            # chunked accumulation is exact
        g, loss = CK.actor_grads(p64, t(s), t(taus), t(acts), w.reshape(-1), chunk=chunk)
        assert abs(loss - loss_ref) < 1e-12
        for k, v in g_ref.items():
            assert np.abs(g[k].numpy().reshape(v.shape) - v).max() <= 1e-9 * max(np.abs(v).max(), 1e-30), k


def test_checker_matches_oracle_td():
    rng = np.random.default_rng(1)
    S, A, B = 6, 2, 33
    nets = O.make_agent(rng, S, A, bias_scale=0.05)
    b = O.make_batch(rng, S, A, B)
    u = rng.uniform(size=(B, A)).astype(np.float32)
    ref = O.critic_train_grads(nets["fnn1"], nets["fnn2"], nets["target_value"], b["s"], b["a"], b["r"], b["sp"], b["it"], 0.99, 0.01, u,
                               dtype=np.float64)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    x = torch.cat([t(b["s"]), torch.clamp(t(b["a"]) + (t(u) * 2 - 1) * 0.01, -1, 1)], dim=-1)
    for name in ("fnn1", "fnn2"):
        p64 = CK.params64(nets[name], "cpu", grad=True)
        g, loss = CK.td_grads(p64, x, t(ref["yq"]))
        assert abs(loss - ref[name + "_loss"]) <= 1e-9 * abs(loss)
        for k, v in ref[name].items():
            assert np.abs(g[k].numpy().reshape(v.shape) - v).max() <= 1e-9 * max(np.abs(v).max(), 1e-30), (name, k)
        assert CK.fnn_margin(p64, x).shape == (B,)


def test_margin_flags_a_row_on_the_kink():
    p, s, taus, acts, adv = _setup(M=16)
    p64 = CK.params64(p, "cpu")
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    pred, margin = CK.supervised_margins(p64, t(s), t(taus), t(acts))
    # put row 3's target exactly on its prediction for one dimension: the Huber indicator sits on its kink
    acts2 = t(acts).clone()
    acts2[3, 2] = pred[3, 2].detach()            # the last dimension feeds no later step: the forward is unchanged
    pred2, margin2 = CK.supervised_margins(p64, t(s), t(taus), acts2)
    assert float(margin2[3]) == 0.0 and float((pred2 - pred).abs().max()) == 0.0
    assert bool((margin2[torch.arange(16) != 3] > 0).all())
