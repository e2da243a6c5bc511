"""CPU checks of the oracle itself (the reference ships no golden vectors: parity unpinned).
Two independent implementations, fp64 finite differences, algebraic identities, the reference's
own shape tests (tests/unit/GAC/networks.py) re-expressed, and the committed golden fixture."""
import copy
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import gac_oracle as O
from oracle import torch_ref as T

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden_mod():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_cosine_constants_follow_reference_rounding():
    # networks.py:45: np.arange(1,65,float32) * np.pi is an fp32 product of fl32(i) and fl32(pi) (F4)
    ipi = O.i_pi()
    assert ipi.dtype == np.float32
    ref = np.arange(1, 65, dtype=np.float32) * np.pi
    assert (ipi == ref).all()
    once = (np.arange(1, 65, dtype=np.float64) * np.pi).astype(np.float32)
    assert (ipi != once).sum() == 24


def test_reference_shape_tests():
    rng = np.random.default_rng(0)
    # test_cosine_basis_linear: (10,10) -> (10,10,400)
    w, b = rng.standard_normal((64, 400)).astype(np.float32), np.zeros(400, np.float32)
    assert O.cosine_basis_linear(rng.standard_normal((10, 10)).astype(np.float32), w, b, np.float32).shape == (10, 10, 400)
    # test_autoregressive_stochastic_actor_no_action: batch 1, S=10, A=5, taus (B,A,1)
    p = O.init_actor(rng, 10, 5)
    out = O.aiqn_sample(p, rng.standard_normal((1, 10)), rng.uniform(size=(1, 5, 1)))
    assert out.shape == (1, 5)
    # ..._with_action: batch 10, S=20, A=5, taus/actions (B,A,1)
    p = O.init_actor(rng, 20, 5)
    out, _ = O.aiqn_supervised(p, rng.standard_normal((10, 20)), rng.uniform(size=(10, 5, 1)), rng.uniform(size=(10, 5, 1)))
    assert out.shape == (10, 5)
    # test_critic: -> (10,1)
    assert O.critic_forward(O.init_fnn(rng, 25), O.init_fnn(rng, 25), rng.standard_normal((10, 20)),
                            rng.uniform(size=(10, 5))).shape == (10, 1)


def test_teacher_forcing_identity():
    rng = np.random.default_rng(1)
    p = O.init_actor(rng, 7, 4, bias_scale=0.1)
    s, tau = rng.standard_normal((32, 7)), rng.uniform(size=(32, 4))
    a = O.aiqn_sample(p, s, tau, np.float64)
    a2, _ = O.aiqn_supervised(p, s, tau, a.astype(np.float32), np.float64)
    # teacher actions are rounded to fp32 at the cosine argument, exactly as the free-running path does
    assert np.abs(a - a2).max() == 0.0


def test_gru_matches_torch_gru_cell():
    rng = np.random.default_rng(2)
    p = O.cast(O.init_actor(rng, 5, 2, bias_scale=0.1), np.float64)
    x, h = rng.standard_normal((9, 800)), rng.standard_normal((9, 400))
    hn, _ = O.gru_step(p, x, h)
    cell = torch.nn.GRUCell(800, 400).double()
    # Keras gate order [z|r|h] -> torch [r|z|n]
    perm = lambda m: np.concatenate([m[..., 400:800], m[..., :400], m[..., 800:]], axis=-1)
    with torch.no_grad():
        cell.weight_ih.copy_(torch.tensor(perm(p["kx"]).T))
        cell.weight_hh.copy_(torch.tensor(perm(p["u"]).T))
        cell.bias_ih.copy_(torch.tensor(perm(p["gb"][0])))
        cell.bias_hh.copy_(torch.tensor(perm(p["gb"][1])))
        ref = cell(torch.tensor(x), torch.tensor(h)).numpy()
    assert np.abs(hn - ref).max() < 1e-12


def test_whole_step_matches_torch_autograd_fp64():
    rng = np.random.default_rng(3)
    S, A, B, K = 5, 3, 8, 4
    nets = O.make_agent(rng, S, A, bias_scale=0.05)
    batch, noise = O.make_batch(rng, S, A, B), O.make_noise(rng, A, B, K)
    n1, info = O.train_one_step(copy.deepcopy(nets), batch, noise, K, dtype=np.float64)
    tg = T.TorchGAC(copy.deepcopy(nets), A, K, dtype=torch.float64)
    ti = tg.train_one_step(batch, noise)
    assert info["M"] == ti["M"] and info["M"] > 0
    assert (info["filter"]["idx"] == ti["filter"]["idx"].numpy()).all()
    for name in ("fnn1", "fnn2", "value", "actor"):
        for k, g in info["grad_" + name].items():
            ref = ti["grad_" + name][k].numpy()
            assert np.abs(g - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-30) + 1e-15, (name, k)
    assert abs(info["loss_actor"] - ti["loss_actor"]) < 1e-12
    # post-step parameters (Adam + Polyak), fp32 state updated from fp64 grads on both sides
    for name in ("actor", "fnn1", "fnn2", "value"):
        for k in n1[name]:
            assert np.abs(n1[name][k] - tg.n[name][k].detach().numpy()).max() < 2e-6, (name, k)
            assert np.abs(n1["target_" + name][k] - tg.n["target_" + name][k].numpy()).max() < 2e-6


def test_actor_gradient_finite_differences():
    rng = np.random.default_rng(4)
    S, A, M = 4, 3, 6
    p = O.init_actor(rng, S, A, bias_scale=0.1)
    s, acts = rng.standard_normal((M, S)), rng.uniform(-1, 1, (M, A)).astype(np.float32)
    adv, taus = rng.uniform(0.1, 1, (M, 1)), rng.uniform(size=(M, A)).astype(np.float32)
    g, loss, _, _ = O.actor_train_grads(p, s, acts, adv, taus, dtype=np.float64)

    def f(pp):
        pred, _ = O.aiqn_supervised(pp, s, taus, acts, np.float64)
        return O.quantile_huber_loss(pred, acts.astype(np.float64), taus.astype(np.float64), adv / adv.sum())[0]

    for k in ("ws", "kx", "u", "gb", "wa", "wn", "w1", "w2", "b2", "bs"):
        flat = p[k].reshape(-1)
        for j in rng.choice(flat.size, min(3, flat.size), replace=False):
            pp = {n: v.astype(np.float64).copy() for n, v in p.items()}
            eps = 1e-6
            pp[k].reshape(-1)[j] += eps
            up = f(pp)
            pp[k].reshape(-1)[j] -= 2 * eps
            dn = f(pp)
            fd = (up - dn) / (2 * eps)
            an = g[k].reshape(-1)[j]
            assert abs(fd - an) <= 1e-5 * max(abs(an), 1e-6) + 1e-9, (k, j, fd, an)


def test_fnn_gradient_is_gradient_of_sum():
    # F5: mse over a size-1 axis => (B,) vector; tape.gradient sums => d/dpred = 2 (pred - y), no 1/B
    rng = np.random.default_rng(5)
    p = O.cast(O.init_fnn(rng, 6, bias_scale=0.1), np.float64)
    x, y = rng.standard_normal((5, 6)), rng.standard_normal((5, 1))
    g = O.value_train_grads(p, x, y, np.float64)
    tp = T.to_t(p, torch.float64, grad=True)
    loss = ((torch.tensor(y) - T.fnn_t(tp, torch.tensor(x))) ** 2).mean(dim=-1).sum()
    gr = torch.autograd.grad(loss, list(tp.values()))
    for k, r in zip(tp.keys(), gr):
        assert np.abs(g[k] - r.numpy()).max() < 1e-12


def test_filter_edge_cases():
    q = np.array([1.0, 2.0, np.nan, 3.0, 3.0, -1.0], np.float32)
    v = np.array([0.5, 2.0, 0.0, np.nan, 2.5, -1.0], np.float32)
    mask, idx = O.positive_advantage_filter(q, v)
    assert idx.dtype == np.int64 and idx.shape == (2, 1)
    assert idx[:, 0].tolist() == [0, 4]             # strict >, NaN excluded
    _, idx0 = O.positive_advantage_filter(np.zeros(4), np.ones(4))
    assert idx0.shape == (0, 1)


def test_density_and_polyak_quirks():
    adv = np.array([[1.0], [3.0]])
    assert np.allclose(O.target_density("linear", adv, 1.0), [[0.25], [0.75]])
    assert (O.target_density("boltzmann", adv, 0.5) == 1).all()          # F6
    assert (T.AIQNT is not None)
    with pytest.raises(NotImplementedError):
        O.target_density("max", adv, 1.0)
    rng = np.random.default_rng(6)
    t, s = rng.standard_normal(100).astype(np.float32), rng.standard_normal(100).astype(np.float32)
    assert (O.polyak(t, s, 1.0) == s).all()                              # rate 1 => hard copy
    assert (O.polyak(t, s, 0.0) == t).all()


def test_tilings():
    s = np.arange(6).reshape(3, 2)
    assert O.tile_state_major(s, 2)[:, 0].tolist() == [0, 0, 2, 2, 4, 4]     # networks.py:464-467
    assert O.tile_sample_major(s, 2)[:, 0].tolist() == [0, 2, 4, 0, 2, 4]    # agent.py:186


def test_adam_is_keras_formula():
    p, g = np.float32(1.0), np.float32(0.5)
    p1, m1, v1 = O.adam_tf(np.array([p]), np.array([g]), np.zeros(1, np.float32), np.zeros(1, np.float32), 1)
    m, v = 0.1 * 0.5, 0.001 * 0.25
    alpha = 1e-4 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert abs(p1[0] - (1.0 - alpha * m / (np.sqrt(v) + 1e-7))) < 1e-7     # eps OUTSIDE the bias correction


def test_golden_fixture_matches_oracle():
    gm = _golden_mod()
    z = np.load(os.path.join(HERE, "golden", "gac_step_small.npz"))
    nets, batch, noise, st, tau, tch = gm.inputs()
    for name in ("actor", "target_value"):
        for k, v in nets[name].items():
            assert np.array_equal(z[f"in.{name}.{k}"], gm.probe(v)), "NumPy bit generator changed: regenerate the fixture"
    after, info = O.train_one_step(copy.deepcopy(nets), batch, noise, gm.K, dtype=np.float64)
    assert int(z["M"]) == info["M"]
    assert np.array_equal(z["idx"], info["filter"]["idx"])
    for name in ("actor", "fnn1", "fnn2", "value"):
        for k, v in info["grad_" + name].items():
            assert np.allclose(z[f"grad.{name}.{k}"], gm.probe(v), rtol=1e-9, atol=1e-14)
        for k, v in after["target_" + name].items():
            assert np.allclose(z[f"out.target_{name}.{k}"], gm.probe(v), rtol=1e-6, atol=1e-9)
    assert np.allclose(z["stage.sample"], O.aiqn_sample(nets["actor"], st, tau, np.float64), atol=1e-12)
    # fp32 oracle stays within tolerance of the fp64 fixture on the teacher-forced path
    sup32, _ = O.aiqn_supervised(nets["actor"], st, tau, tch, np.float32)
    assert np.abs(sup32 - z["stage.supervised"]).max() < 1e-5


def test_iqn_actor_matches_torch_autograd():
    """StochasticActor (IQN, networks.py:271-323): forward + hand-derived backward vs torch autograd, and the
    reference's shape test (tests/unit/GAC/networks.py:96-130: batch 10, S=20, A=5 -> (10, 5))."""
    rng = np.random.default_rng(8)
    S, A, M = 20, 5, 10
    p = O.init_iqn(rng, S, A, bias_scale=0.1)
    assert p["wn"].shape == (64, 80) and p["wm"].shape == (400, 200) and O.iqn_shapes(17, 6)["ws"] == (17, 396)
    s, taus = rng.standard_normal((M, S)), rng.uniform(size=(M, A, 1)).astype(np.float32)
    acts, adv = rng.uniform(-1, 1, (M, A)).astype(np.float32), rng.uniform(0.1, 1, (M, 1))
    out = O.iqn_forward(p, s, taus, np.float64)
    assert out.shape == (M, A)
    g, loss, pred, _ = O.iqn_train_grads(p, s, acts, adv, taus, dtype=np.float64)
    tp = T.to_t(p, torch.float64, grad=True)
    tpred = T.IQNT(tp, A)(torch.tensor(s), torch.tensor(taus.reshape(M, A), dtype=torch.float64))
    assert np.abs(tpred.detach().numpy() - pred).max() < 1e-12
    tl = T.huber_quantile_loss_t(tpred, torch.tensor(acts, dtype=torch.float64), torch.tensor(taus.reshape(M, A), dtype=torch.float64),
                                 torch.tensor(adv / adv.sum()))
    gr = torch.autograd.grad(tl, list(tp.values()))
    assert abs(float(tl.detach()) - loss) < 1e-14
    for k, r in zip(tp.keys(), gr):
        assert np.abs(g[k] - r.numpy()).max() <= 1e-10 * max(np.abs(r.numpy()).max(), 1e-30), k
