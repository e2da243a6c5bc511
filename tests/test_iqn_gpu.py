"""StochasticActor (IQN, reference GAC/networks.py:271-323 -- the CLI's default actor, main.py:78) on the GPU
vs the oracle: forward, quantile-Huber gradients, and the whole GACAgent step with actor='IQN'."""
import copy

import numpy as np
import pytest
import torch

from conftest import FWD_RTOL, GRAD_RTOL, assert_close
from oracle import gac_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gb():
    import __graft_entry__ as ge
    ge.build()
    import dpo_replication_b200 as g
    return g


def dev(x, dtype=torch.float32):
    return torch.as_tensor(np.asarray(x)).to("cuda", dtype).contiguous()


@pytest.mark.parametrize("engine", [1, 0])
@pytest.mark.parametrize("S,A,rows,nst,live", [(20, 5, 10, 10, None), (17, 6, 300, 24, None), (376, 17, 256, 40, 131), (3, 1, 64, 8, None)])
def test_iqn_forward_and_gradients(gb, engine, S, A, rows, nst, live):
    rng = np.random.default_rng(21)
    p = O.init_iqn(rng, S, A, bias_scale=0.05)
    eng = gb.Engine(S, A, max(rows, 128), 1, "cuda", engine=engine, actor_kind=1)
    ar = gb.Arena(S, A, torch.device("cuda"), 1)
    ar.load(ar.online, "actor", p)
    su = rng.standard_normal((nst, S)).astype(np.float32)
    sidx = rng.integers(0, nst, rows).astype(np.int32)
    taus = rng.uniform(size=(rows, A)).astype(np.float32)
    acts = rng.uniform(-1, 1, (rows, A)).astype(np.float32)
    adv = rng.uniform(0.05, 1.0, rows).astype(np.float32)
    base = ar.net_slice(ar.online, "actor")
    out = eng.iqn_forward(base, dev(su), dev(sidx, torch.int32), dev(taus))
    assert_close(out.cpu().numpy(), O.iqn_forward(p, su[sidx], taus, np.float64), FWD_RTOL, "iqn forward")
    n = rows if live is None else live
    d_rows = None if live is None else dev(np.array([live]), torch.int32)
    pred = eng.iqn_train_fwd(base, dev(su), dev(sidx, torch.int32), dev(taus), d_rows)
    g_ref, loss_ref, pred_ref, dpred_ref = O.iqn_train_grads(p, su[sidx][:n], acts[:n], adv[:n, None], taus[:n], "linear", 1.0, np.float64)
    assert_close(pred.cpu().numpy()[:n], pred_ref, FWD_RTOL, "iqn train forward")
    w = np.zeros(rows, np.float32); w[:n] = adv[:n] / adv[:n].astype(np.float64).sum()
    loss, dpred = eng.qhuber_loss_bwd(pred, dev(acts), dev(taus), dev(w), 1.0 / (n * A), d_rows)
    assert abs(loss.item() - loss_ref) <= GRAD_RTOL * abs(loss_ref)
    dfull = np.zeros((rows, A), np.float32); dfull[:n] = dpred_ref
    eng.iqn_bwd(base, dev(dfull), ar.net_slice(ar.grad, "actor"))
    got = ar.named(ar.grad, "actor")
    assert set(got) == set(g_ref)
    for k, g in g_ref.items():
        assert_close(got[k].cpu().numpy().reshape(g.shape), g, GRAD_RTOL, f"iqn grad {k}")
    eng.iqn_bwd(base, dev(dfull), ar.net_slice(ar.grad, "actor"), accumulate=True)
    assert_close(ar.named(ar.grad, "actor")["wm"].cpu().numpy(), 2 * g_ref["wm"], GRAD_RTOL, "accumulate")


def _oracle_step_iqn(nets, batch, noise, K, dtype=np.float64):
    """agent.py:121-168 with the IQN actor, from the oracle's building blocks."""
    B = batch["s"].shape[0]
    f = lambda a: np.asarray(a, dtype)
    info = {}
    cg = O.critic_train_grads(nets["fnn1"], nets["fnn2"], nets["target_value"], batch["s"], batch["a"], batch["r"], batch["sp"], batch["it"],
                              0.99, 0.01, noise["critic_u"], dtype)
    tiled = O.tile_state_major(f(batch["s"]), K)
    va = O.iqn_forward(nets["target_actor"], tiled, noise["value_tau"], dtype)
    vq = O.critic_forward(nets["target_fnn1"], nets["target_fnn2"], tiled, va, dtype)
    vg = O.value_train_grads(nets["value"], batch["s"], O.v_critic_from_q(vq, B, K), dtype)
    vg.pop("_loss")
    tiled = O.tile_sample_major(f(batch["s"]), K)
    ta = np.clip(O.iqn_forward(nets["target_actor"], tiled, noise["filter_tau"], dtype) + f(noise["filter_normal"]) * dtype(0.01), -1, 1)
    acts = np.concatenate([ta, f(noise["filter_rand"])], 0)
    tiled2 = np.concatenate([tiled, tiled], 0)
    q = O.critic_forward(nets["target_fnn1"], nets["target_fnn2"], tiled2, acts, dtype)
    v = O.value_forward(nets["target_value"], tiled2, dtype)
    mask, idx = O.positive_advantage_filter(q, v)
    i = idx[:, 0]
    M = len(i)
    ag, loss, _, _ = O.iqn_train_grads(nets["actor"], tiled2[i], acts[i], (q - v)[i], noise["actor_tau"][:M], "linear", 1.0, dtype)
    info.update(grad_fnn1=cg["fnn1"], grad_fnn2=cg["fnn2"], grad_value=vg, grad_actor=ag, M=M, idx=idx, q=q, v=v, loss_actor=loss)
    for name, g in (("fnn1", cg["fnn1"]), ("fnn2", cg["fnn2"]), ("value", vg), ("actor", ag)):
        nets["t_" + name] += 1
        for k in nets[name]:
            nets[name][k], nets["m_" + name][k], nets["v_" + name][k] = O.adam_tf(nets[name][k], np.asarray(g[k], np.float32),
                                                                                nets["m_" + name][k], nets["v_" + name][k], nets["t_" + name])
    for name in ("actor", "fnn1", "fnn2", "value"):
        for k in nets[name]:
            nets["target_" + name][k] = O.polyak(nets["target_" + name][k], nets[name][k], 5e-3)
    return nets, info


@pytest.mark.parametrize("engine", [1, 0])
def test_iqn_agent_step_end_to_end(gb, engine):
    # the IQN actor has no autoregressive feedback, so the whole step is comparable at any A
    S, A, B, K = 17, 6, 32, 16
    rng = np.random.default_rng(33)
    nets = O.make_agent(rng, S, A, bias_scale=0.05)
    nets["actor"], nets["target_actor"] = O.init_iqn(rng, S, A, 0.05), O.init_iqn(rng, S, A, 0.05)
    # non-trivial Adam state: with m = v = 0 the first update is sign-like (m / sqrt(v)) and amplifies
    # rounding differences of near-zero gradients
    for n in ("actor", "fnn1", "fnn2", "value"):
        nets["m_" + n] = {k: (1e-3 * rng.standard_normal(v.shape)).astype(np.float32) for k, v in nets[n].items()}
        nets["v_" + n] = {k: (1e-6 * rng.uniform(size=v.shape)).astype(np.float32) for k, v in nets[n].items()}
        nets["t_" + n] = 5
    batch, noise = O.make_batch(rng, S, A, B), O.make_noise(rng, A, B, K)
    _, probe = _oracle_step_iqn(copy.deepcopy(nets), batch, noise, K)
    nets["target_value"]["b2"] = nets["target_value"]["b2"] + np.float32(np.median(probe["q"] - probe["v"]))
    agent = gb.GACAgent(A, S, action_samples=K, batch_size=B, actor="IQN", device="cuda", engine=engine)
    assert agent.actor.module_type == "StochasticActor" and len(agent.actor.trainable_variables) == 8
    agent.load_state(nets)
    P = B * K
    taps = dict(sel_idx=torch.zeros(2 * P, dtype=torch.int64, device="cuda"), q_out=torch.zeros(2 * P, device="cuda"),
                v_out=torch.zeros(2 * P, device="cuda"))
    agent.train_on_batch(batch, noise, taps)
    torch.cuda.synchronize()
    ref, info = _oracle_step_iqn(copy.deepcopy(nets), batch, noise, K)
    st = agent.arena.stats.cpu().numpy()
    M = int(st[5])
    q, v = taps["q_out"].cpu().numpy(), taps["v_out"].cpu().numpy()
    _, idx_given = O.positive_advantage_filter(q, v)
    assert M == idx_given.shape[0] and np.array_equal(taps["sel_idx"].cpu().numpy()[:M], idx_given[:, 0])   # bit-exact given q, v
    assert_close(q, info["q"].reshape(-1), FWD_RTOL, "candidate q")
    if (np.abs(info["q"] - info["v"]) > 1e-4).all():
        assert np.array_equal(idx_given, info["idx"]) and M > 0
        scale = 1.0 / (M * A * st[4])
        got = agent.arena.named(agent.arena.grad, "actor")
        for k, g in info["grad_actor"].items():
            assert_close(got[k].cpu().numpy().reshape(g.shape) * scale, g, GRAD_RTOL, f"actor grad {k}")
        assert abs(st[3] * scale - info["loss_actor"]) <= GRAD_RTOL * abs(info["loss_actor"])
        out = agent.dump_state()
        for name in ("actor", "fnn1", "fnn2", "value"):
            for k, r in ref[name].items():
                assert_close(out[name][k].reshape(r.shape), r, 1e-5, f"param {name}.{k}")
                assert_close(out["target_" + name][k].reshape(r.shape), ref["target_" + name][k], 1e-5, f"target {name}.{k}")


def test_iqn_reference_surface(gb):
    # tests/unit/GAC/networks.py:96-130: batch 10, S=20, A=5 -> (10, 5); supervise_actions is ignored
    actor = gb.StochasticActor(20, 5, 64)
    assert len(actor.trainable_variables) == 8 and actor.noise_embed_dim == 80
    st, tau = torch.randn(10, 20) * 0.1, torch.rand(10, 5, 1)
    a = actor(st, tau)
    assert tuple(a.shape) == (10, 5) and float(a.abs().max()) <= 1.0
    assert torch.equal(a, actor(st, tau, torch.rand(10, 5, 1)))
    before = [x.clone() for x in actor.trainable_variables]
    assert actor.train(torch.randn(8, 20), torch.rand(8, 5) * 2 - 1, torch.rand(8, 1), "boltzmann", 1.0) is None
    assert any(not torch.equal(x, y) for x, y in zip(before, actor.trainable_variables))
    ag = gb.GACAgent(5, 20, buffer_size=64, action_samples=4, batch_size=8, actor="IQN")
    rng = np.random.default_rng(0)
    for _ in range(16):
        ag.store_transition(rng.standard_normal(20), rng.uniform(-1, 1, 5), 0.5, rng.standard_normal(20), False)
    assert ag.train_one_step() is None
    assert tuple(ag.get_action(torch.randn(3, 20)).shape) == (3, 5)
