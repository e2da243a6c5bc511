"""Whole train step (GACAgent.train_one_step, agent.py:121-168) on the GPU vs the oracle."""
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


def setup(S, A, B, K, seed, centre=True):
    rng = np.random.default_rng(seed)
    nets = O.make_agent(rng, S, A, bias_scale=0.05)
    batch, noise = O.make_batch(rng, S, A, B), O.make_noise(rng, A, B, K)
    if centre:   # M/2P ~ 0.5
        flt = O.sample_positive_advantage_actions(nets, batch["s"], noise, K, np.float64)
        nets["target_value"]["b2"] = nets["target_value"]["b2"] + np.float32(np.median(flt["q"] - flt["v"]))
    # non-trivial Adam state
    for n in ("actor", "fnn1", "fnn2", "value"):
        for k in nets[n]:
            nets["m_" + n][k] = (1e-3 * rng.standard_normal(nets[n][k].shape)).astype(np.float32)
            nets["v_" + n][k] = (1e-6 * rng.uniform(size=nets[n][k].shape)).astype(np.float32)
        nets["t_" + n] = 5
    return nets, batch, noise


def run_gpu(gb, nets, batch, noise, S, A, B, K, engine, chunk_rows=0, mode="linear"):
    agent = gb.GACAgent(A, S, action_samples=K, batch_size=B, mode=mode, device="cuda", engine=engine, chunk_rows=chunk_rows)
    agent.load_state(nets)
    P = B * K
    taps = dict(sel_idx=torch.zeros(2 * P, dtype=torch.int64, device="cuda"), q_out=torch.zeros(2 * P, device="cuda"),
                v_out=torch.zeros(2 * P, device="cuda"), cand_actions=torch.zeros(2 * P, A, device="cuda"),
                value_actions=torch.zeros(P, A, device="cuda"))
    agent.train_on_batch(batch, noise, taps)
    torch.cuda.synchronize()
    return agent, taps


@pytest.mark.parametrize("engine", [1, 0])
@pytest.mark.parametrize("mode,B,K", [("linear", 64, 32), ("boltzmann", 64, 32), ("linear", 10, 10), ("linear", 37, 29)])
def test_pendulum_step_end_to_end(gb, engine, mode, B, K):
    # C1: S=3, A=1, B=64, K=32 -- A = 1 has no chaotic feedback, so the whole step is comparable.
    # (10, 10) and (37, 29): row counts that are not multiples of the 128-row tiles (ragged last tiles, 3xTF32 fallbacks)
    S, A = 3, 1
    nets, batch, noise = setup(S, A, B, K, 100)
    agent, taps = run_gpu(gb, nets, batch, noise, S, A, B, K, engine, mode=mode)
    ref, info = O.train_one_step(copy.deepcopy(nets), batch, noise, K, mode=mode, dtype=np.float64)
    st = agent.arena.stats.cpu().numpy()
    M = int(st[5])
    # selection: bit-exact GIVEN the q, v the GPU computed (north_star), and equal to the oracle's
    q, v = taps["q_out"].cpu().numpy(), taps["v_out"].cpu().numpy()
    _, idx_given = O.positive_advantage_filter(q, v)
    assert M == idx_given.shape[0]
    assert np.array_equal(taps["sel_idx"].cpu().numpy()[:M], idx_given[:, 0])
    assert_close(q, info["filter"]["q"].reshape(-1), FWD_RTOL, "candidate q")
    assert_close(v, info["filter"]["v"].reshape(-1), FWD_RTOL, "candidate v")
    assert_close(taps["value_actions"].cpu().numpy(), info["value_actions"], FWD_RTOL, "value-pass actions")
    margin = np.abs(info["filter"]["q"] - info["filter"]["v"]).reshape(-1)
    if (margin > 1e-4).all():          # no candidate sits on the decision boundary => same set as the oracle
        assert np.array_equal(idx_given, info["filter"]["idx"])
        assert abs(st[0] - info["loss_fnn1"]) <= GRAD_RTOL * abs(info["loss_fnn1"])
        assert abs(st[2] - info["loss_value"]) <= GRAD_RTOL * abs(info["loss_value"])
        g = agent.arena
        scale = 1.0 / (M * A * st[4]) if mode == "linear" else 1.0 / (M * A)
        assert abs(st[3] * scale - info["loss_actor"]) <= GRAD_RTOL * abs(info["loss_actor"])
        for net in ("fnn1", "fnn2", "value", "actor"):
            got = g.named(g.grad, net)
            for k, gr in info["grad_" + net].items():
                mult = scale if net == "actor" else 1.0
                assert_close(got[k].cpu().numpy().reshape(gr.shape) * mult, gr, GRAD_RTOL, f"grad {net}.{k}")
        got = agent.dump_state()
        for name in ("actor", "fnn1", "fnn2", "value"):
            for k, r in ref[name].items():
                assert_close(got[name][k].reshape(r.shape), r, 1e-5, f"param {name}.{k}")
                assert_close(got["target_" + name][k].reshape(r.shape), ref["target_" + name][k], 1e-5, f"target {name}.{k}")
            assert got["t_" + name] == ref["t_" + name]


@pytest.mark.parametrize("engine", [1, 0])
def test_halfcheetah_dims_step_and_chunking(gb, engine):
    # HalfCheetah dims (S=17, A=6) at a small batch; the same step with the actor fwd/bwd split
    # into 128-row chunks must give the same gradients (sum order differs => tolerance)
    S, A, B, K = 17, 6, 16, 24
    nets, batch, noise = setup(S, A, B, K, 200)
    a1, t1 = run_gpu(gb, nets, batch, noise, S, A, B, K, engine)
    a2, t2 = run_gpu(gb, nets, batch, noise, S, A, B, K, engine, chunk_rows=128)
    assert a2.engine.cfg.chunk_rows == 128
    M = int(a1.arena.stats[5].item())
    assert M == int(a2.arena.stats[5].item()) and M > 128
    assert np.array_equal(t1["sel_idx"].cpu().numpy()[:M], t2["sel_idx"].cpu().numpy()[:M])
    for net in ("actor", "fnn1", "fnn2", "value"):
        g1, g2 = a1.arena.named(a1.arena.grad, net), a2.arena.named(a2.arena.grad, net)
        for k in g1:
            assert_close(g2[k].cpu().numpy(), g1[k].cpu().numpy(), 2e-5, f"chunked grad {net}.{k}")
    # actor gradient against the oracle, fed the GPU's own selection (stage-wise: F3)
    idx = t1["sel_idx"].cpu().numpy()[:M]
    acts = t1["cand_actions"].cpu().numpy()[idx]
    adv = (t1["q_out"].cpu().numpy() - t1["v_out"].cpu().numpy())[idx]
    states = np.tile(batch["s"], (2 * K, 1))[idx]
    g_ref, loss_ref, _, _ = O.actor_train_grads(nets["actor"], states, acts, adv[:, None], noise["actor_tau"][:M], "linear", 1.0, np.float64)
    st = a1.arena.stats.cpu().numpy()
    scale = 1.0 / (M * A * st[4])
    got = a1.arena.named(a1.arena.grad, "actor")
    # wa / ba sit downstream of the action embedding's LeakyReLU(0.2) gate: one pre-activation within rounding distance of 0
    # flips the gate between the fp32 kernel and the fp64 oracle and moves those two tensors by a discrete amount (a few
    # 1e-4 at ~400 rows, data dependent: the teacher actions are the GPU sampler's own draws) -- held to 1e-3 like the
    # kink-downstream tensors of tests/test_fullsize_gpu.py; everything else at 1e-4.
    for k, gr in g_ref.items():
        assert_close(got[k].cpu().numpy().reshape(gr.shape) * scale, gr, 1e-3 if k in ("wa", "ba") else GRAD_RTOL, f"actor grad {k}")
    assert abs(st[3] * scale - loss_ref) <= GRAD_RTOL * abs(loss_ref)


def test_step_with_no_positive_advantage_skips_actor(gb):
    S, A, B, K = 3, 1, 8, 4
    nets, batch, noise = setup(S, A, B, K, 300, centre=False)
    nets["target_value"]["b2"] = nets["target_value"]["b2"] + np.float32(100.0)     # V >> Q => M = 0
    agent, _ = run_gpu(gb, nets, batch, noise, S, A, B, K, 1)
    got = agent.dump_state()
    assert int(agent.arena.stats[5].item()) == 0
    assert got["t_actor"] == nets["t_actor"] and got["t_fnn1"] == nets["t_fnn1"] + 1      # agent.py:158
    for k, v in nets["actor"].items():
        assert np.array_equal(got["actor"][k].reshape(v.shape), v)                        # untouched
        assert np.array_equal(got["target_actor"][k].reshape(v.shape), O.polyak(nets["target_actor"][k], v, 5e-3))


def test_reference_surface(gb):
    """The reference's own shape tests (tests/unit/GAC/networks.py) against the B200 classes."""
    actor = gb.AutoRegressiveStochasticActor(10, 5, 64)
    assert actor.trainable_variables and len(actor.trainable_variables) == 13
    out = actor(torch.randn(1, 10) * 0.1, torch.rand(1, 5, 1))
    assert tuple(out.shape) == (1, 5)
    actor = gb.AutoRegressiveStochasticActor(20, 5, 64)
    out = actor(torch.randn(10, 20) * 0.1, torch.rand(10, 5, 1), torch.rand(10, 5, 1))
    assert tuple(out.shape) == (10, 5)
    critic = gb.Critic(20, 5)
    assert tuple(critic(torch.randn(10, 20), torch.rand(10, 5)).shape) == (10, 1)
    emb = gb.CosineBasisLinear(64, 400)
    assert tuple(emb(torch.randn(10, 10) * 0.1).shape) == (10, 10, 400)
    # update(target, source, 1.0) is a hard copy (helpers.py:75-86)
    t, s = gb.Critic(4, 2), gb.Critic(4, 2)
    gb.update(t, s, 1.0)
    assert all(torch.equal(a, b) for a, b in zip(t.trainable_variables, s.trainable_variables))
    with pytest.raises(NotImplementedError):
        gb.IQNActor(3, 1).target_density("max", torch.ones(2, 1), 1.0)
    # stand-alone .train calls mutate in place and return None
    v = gb.Value(4)
    before = [x.clone() for x in v.trainable_variables]
    a = gb.AutoRegressiveStochasticActor(4, 2)
    assert v.train(torch.randn(8, 4), a, s, action_samples=4) is None
    assert any(not torch.equal(x, y) for x, y in zip(before, v.trainable_variables))
    before = [x.clone() for x in a.trainable_variables]
    assert a.train(torch.randn(8, 4), torch.rand(8, 2) * 2 - 1, torch.rand(8, 1), "linear", 1.0) is None
    assert any(not torch.equal(x, y) for x, y in zip(before, a.trainable_variables))
    tv = gb.Value(4)
    assert s.train(torch.randn(8, 4), torch.rand(8, 2), torch.randn(8, 1), torch.randn(8, 4), torch.zeros(8, 1), tv, 0.99) is None


def test_agent_surface(gb):
    ag = gb.GACAgent(2, 4, buffer_size=64, action_samples=4, batch_size=8)
    rng = np.random.default_rng(0)
    for _ in range(20):
        ag.store_transition(rng.standard_normal(4), rng.uniform(-1, 1, 2), 0.5, rng.standard_normal(4), False)
    assert ag.replay.size == 20
    assert ag.train_one_step() is None
    a = ag.get_action(torch.randn(5, 4))
    assert tuple(a.shape) == (5, 2) and float(a.abs().max()) <= 1.0
    a = ag.select_perturbed_action(np.zeros((1, 4), np.float32), lambda: np.full((1, 2), 5.0, np.float32))
    assert float(a.max()) == 1.0
    gs, ga, adv = ag._sample_positive_advantage_actions(torch.randn(8, 4))
    assert gs.shape[0] == ga.shape[0] == adv.shape[0] and gs.shape[1] == 4 and ga.shape[1] == 2
    assert bool((adv > 0).all())
    with pytest.raises(NotImplementedError):
        gb.GACAgent(2, 4, mode="max")


def test_device_replay_buffer_matches_reference_ring(gb):
    """DeviceReplayBuffer vs the host ReplayBuffer (reference tests/unit/GAC/helpers.py:10-47 semantics: FIFO
    ring, store writes row `ptr`, sampling without replacement); the on-device gather is bit-exact."""
    import random
    S, A, size = 5, 2, 4
    dev_buf = gb.DeviceReplayBuffer(S, A, size, "cuda")
    host = gb.ReplayBuffer(S, A, size)
    rng = np.random.default_rng(3)
    for i in range(6):                                   # 6 stores into a size-4 ring: rows 0,1 overwritten
        tr = (rng.standard_normal(S).astype(np.float32), rng.uniform(-1, 1, A).astype(np.float32), float(i), rng.standard_normal(S).astype(np.float32), i == 5)
        dev_buf.store(*tr); host.store(*tr)
        if i == 0:
            assert np.array_equal(dev_buf.obs1_buf[0].cpu().numpy(), tr[0])     # test_store: row 0 written
    assert dev_buf.size == host.size == 4 and dev_buf.ptr == host.ptr == 2
    for name in ("obs1_buf", "obs2_buf", "acts_buf", "rews_buf", "done_buf"):
        assert np.array_equal(getattr(dev_buf, name).cpu().numpy(), getattr(host, name))
    random.seed(7); a = dev_buf.sample_batch(3)
    random.seed(7); b = host.sample_batch(3)
    for k in ("s", "a", "r", "sp", "it"):
        assert np.array_equal(getattr(a, k).cpu().numpy(), getattr(b, k).numpy()), k
    assert len(set(dev_buf._last_idx.cpu().tolist())) == 3                       # without replacement


@pytest.mark.parametrize("actor", ["AIQN", "IQN"])
def test_graphed_acting_path_matches_eager(gb, actor):
    """get_action for a fixed small batch replayed from one CUDA graph gives the same
    actions as the eager multi-launch path (same tau), and follows weight updates (the graph reads the arena)."""
    ag = gb.GACAgent(3, 11, action_samples=4, batch_size=8, actor=actor)
    st = torch.randn(1, 11, device="cuda")
    tau = torch.rand(1, 3, device="cuda")
    eager = ag.action_sampler.get_actions(ag.actor, st, taus=tau)
    g1 = ag.get_action_graphed(st, taus=tau)
    assert torch.equal(eager, g1)
    ag.arena.online.mul_(1.01)                      # weights change in place -> the replay must see them
    eager2 = ag.action_sampler.get_actions(ag.actor, st, taus=tau)
    g2 = ag.get_action_graphed(st, taus=tau)
    assert torch.equal(eager2, g2) and not torch.equal(g1, g2)
    assert tuple(ag.get_action_graphed(torch.randn(11)).shape) == (1, 3)


def test_graphed_step_is_bit_identical_to_eager(gb):
    """GACAgent.train_on_batch_graphed (one cudaGraphLaunch per step) against the eager path: same inputs, same draws =>
    bit-identical parameters, targets, Adam moments and counters over several steps."""
    S, A, B, K = 17, 6, 32, 16
    nets, batch, noise = setup(S, A, B, K, 400)
    a1 = gb.GACAgent(A, S, action_samples=K, batch_size=B, device="cuda")
    a2 = gb.GACAgent(A, S, action_samples=K, batch_size=B, device="cuda")
    rng = np.random.default_rng(5)
    for ag in (a1, a2):
        ag.load_state(nets)
    for step in range(3):
        b, n = O.make_batch(rng, S, A, B), O.make_noise(rng, A, B, K)
        a1.train_on_batch(b, n)
        a2.train_on_batch_graphed(b, n)
    torch.cuda.synchronize()
    for k in ("online", "target", "adam_m", "adam_v", "adam_t"):
        assert torch.equal(getattr(a1.arena, k), getattr(a2.arena, k)), k
    assert a1.arena.adam_t.cpu().tolist() == [8, 8, 8, 8]
    # device-side draws: the graphed step runs and keeps advancing
    a2.train_on_batch_graphed(b)
    torch.cuda.synchronize()
    assert a2.arena.adam_t.cpu().tolist() == [9, 9, 9, 9] and bool(torch.isfinite(a2.arena.online).all())
