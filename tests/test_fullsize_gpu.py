"""Full BASELINE.json sizes on the GPU through size-independent properties (the oracle is too slow
at these sizes): teacher-forcing identity, selection bit-exactness given q/v, engine agreement
(tcgen05 3xTF32 vs SIMT fp32), chunking invariance, Polyak identities, Adam step counters."""
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


def _agent(gb, S, A, B, K, engine=0, chunk_rows=0, seed=0):
    torch.manual_seed(seed)
    return gb.GACAgent(A, S, action_samples=K, batch_size=B, device="cuda", engine=engine, chunk_rows=chunk_rows)


def _inputs(S, A, B, K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s: torch.rand(*s, device="cuda", generator=g)
    n = lambda *s: torch.randn(*s, device="cuda", generator=g)
    P = B * K
    batch = dict(s=n(B, S), a=r(B, A) * 2 - 1, r=n(B, 1), sp=n(B, S), it=(r(B, 1) < 0.05).float())
    noise = dict(critic_u=r(B, A), value_tau=r(P, A), filter_tau=r(P, A), filter_normal=n(P, A), filter_rand=r(P, A) * 2 - 1,
                 actor_tau=r(2 * P, A))
    return batch, noise


@pytest.mark.parametrize("name,S,A,B,K", [("halfcheetah", 17, 6, 256, 64), ("humanoid", 376, 17, 256, 128)])
def test_full_config_properties(gb, name, S, A, B, K):
    P = B * K
    ag = _agent(gb, S, A, B, K, engine=0, seed=1)
    batch, noise = _inputs(S, A, B, K, 7)
    taps = dict(sel_idx=torch.zeros(2 * P, dtype=torch.int64, device="cuda"), q_out=torch.zeros(2 * P, device="cuda"),
                v_out=torch.zeros(2 * P, device="cuda"), cand_actions=torch.zeros(2 * P, A, device="cuda"),
                value_actions=torch.zeros(P, A, device="cuda"))
    ag.train_on_batch(batch, noise, taps)
    shift = (taps["q_out"] - taps["v_out"]).median()       # centre V_target: M ~ P
    ag.arena.view(ag.arena.target, "value.b2").add_(shift)
    before = {k: v.clone() for k, v in (("online", ag.arena.online), ("target", ag.arena.target))}
    t_before = ag.arena.adam_t.clone()
    ag.train_on_batch(batch, noise, taps)
    torch.cuda.synchronize()
    st = ag.arena.stats.cpu().numpy()
    M = int(st[5])
    assert 0 < M < 2 * P
    # (1) selection: bit-exact given the q, v the step used; ascending int64
    q, v = taps["q_out"].cpu().numpy(), taps["v_out"].cpu().numpy()
    _, idx = O.positive_advantage_filter(q, v)
    assert M == idx.shape[0] and np.array_equal(taps["sel_idx"].cpu().numpy()[:M], idx[:, 0])
    assert abs(st[4] - (q - v)[idx[:, 0]].astype(np.float64).sum()) <= 1e-4 * abs(st[4])
    # (2) candidates stay in [-1, 1]; sampled actions are tanh outputs
    ca = taps["cand_actions"]
    assert float(ca.abs().max()) <= 1.0 and bool(torch.isfinite(ca).all())
    # (3) Adam counters advanced once for all four nets, every parameter finite, Polyak relation holds
    assert (ag.arena.adam_t - t_before).cpu().tolist() == [1, 1, 1, 1]
    assert bool(torch.isfinite(ag.arena.online).all()) and bool(torch.isfinite(ag.arena.grad).all())
    n = ag.arena.n
    expect_t = O.polyak(before["target"][:n].cpu().numpy(), ag.arena.online[:n].cpu().numpy(), 5e-3)
    assert np.array_equal(ag.arena.target[:n].cpu().numpy(), expect_t)              # bit-exact
    # (4) teacher-forcing identity on the full row count: supervised(actor(s,tau)) == actor(s,tau)
    eng, ar = ag.engine, ag.arena
    sidx = torch.arange(B, dtype=torch.int32, device="cuda").repeat_interleave(K)
    base = ar.net_slice(ar.target, "actor")
    acts = eng.aiqn_sample(base, batch["s"], sidx, noise["value_tau"])
    rows = min(P, eng.cfg.chunk_rows if eng.cfg.chunk_rows else P)
    rows = min(rows, 32768)
    sup = eng.aiqn_supervised_fwd(base, batch["s"], sidx[:rows].contiguous(), noise["value_tau"][:rows].contiguous(), acts[:rows].contiguous())
    assert_close(sup.cpu().numpy(), acts[:rows].cpu().numpy(), 2e-5, f"{name}: teacher-forcing identity")


def test_engines_and_chunking_agree_at_halfcheetah(gb):
    """tcgen05 3xTF32 vs SIMT fp32 at the full HalfCheetah row counts, stage-wise on IDENTICAL inputs
    (the free-running sampler is chaotic, SURVEY.md F3, so whole-step comparisons across engines would
    compare different candidate sets), and row-chunking invariance of the whole step on one engine."""
    S, A, B, K = 17, 6, 256, 64
    P = B * K
    batch, noise = _inputs(S, A, B, K, 11)
    # (a) critic gradients do not involve the sampler: whole-step comparison across engines
    res = {}
    for engine in (0, 1):
        ag = _agent(gb, S, A, B, K, engine=engine, seed=3)
        io = ag.engine.step_io(ag.arena, batch, noise)
        ag.engine.step_grads(io, ag._hp)
        torch.cuda.synchronize()
        res[engine] = ag
    ar = res[1].arena
    for net in ("fnn1", "fnn2"):
        for name, t in ar.named(ar.grad, net).items():
            assert_close(res[0].arena.named(res[0].arena.grad, net)[name].cpu().numpy(), t.cpu().numpy(), GRAD_RTOL, f"tcgen05 vs SIMT {net}.{name}")
    # (b) actor: teacher-forced forward, loss gradient and BPTT on the same 32768 rows for both engines
    g = torch.Generator(device="cuda").manual_seed(5)
    M = 2 * P
    sidx = torch.randint(0, B, (M,), device="cuda", generator=g, dtype=torch.int32)
    teacher = torch.rand(M, A, device="cuda", generator=g) * 2 - 1
    taus = torch.rand(M, A, device="cuda", generator=g)
    w = torch.rand(M, device="cuda", generator=g) + 0.05
    out = {}
    for engine in (0, 1):
        ag = res[engine]
        eng, a2 = ag.engine, ag.arena
        base = a2.net_slice(a2.online, "actor")
        pred = eng.aiqn_supervised_fwd(base, batch["s"], sidx, taus, teacher)
        loss, dpred = eng.qhuber_loss_bwd(pred, teacher, taus, w, 1.0 / (M * A))
        eng.aiqn_bwd(base, dpred, a2.net_slice(a2.grad, "actor"))
        torch.cuda.synchronize()
        out[engine] = (pred.cpu().numpy(), loss.item(), {k: v.cpu().numpy().copy() for k, v in a2.named(a2.grad, "actor").items()})
    assert_close(out[0][0], out[1][0], FWD_RTOL, "tcgen05 vs SIMT supervised forward")
    assert abs(out[0][1] - out[1][1]) <= GRAD_RTOL * abs(out[1][1])
    # w2 / b2 are upstream of every LeakyReLU derivative: strict.  All other gradients pass through
    # lrelu'(y) = {1, alpha}: with ~8e7 pre-activations a few hundred lie within fp32 rounding of zero
    # and flip slope between ANY two fp32 implementations (each flip = one row's whole contribution,
    # ~1/sqrt(rows) of the tensor), so those are held to 1e-3 at this size (1e-4 in the small-size
    # oracle tests where no element sits on the kink).
    for k, v in out[1][2].items():
        assert_close(out[0][2][k], v, GRAD_RTOL if k in ("w2", "b2") else 1e-3, f"tcgen05 vs SIMT actor grad {k}")
    # (c) chunking: same engine => same sampler => same candidates; only the summation order differs
    a1 = _agent(gb, S, A, B, K, engine=0, chunk_rows=0, seed=3)
    a2 = _agent(gb, S, A, B, K, engine=0, chunk_rows=8192, seed=3)
    taps = dict(q_out=torch.zeros(2 * P, device="cuda"), v_out=torch.zeros(2 * P, device="cuda"))
    a1.engine.step_grads(a1.engine.step_io(a1.arena, batch, noise, taps), a1._hp)
    shift = (taps["q_out"] - taps["v_out"]).median()            # centre V_target: M ~ P
    for ag in (a1, a2):
        ag.arena.view(ag.arena.target, "value.b2").add_(shift)
        ag.engine.step_grads(ag.engine.step_io(ag.arena, batch, noise), ag._hp)
    torch.cuda.synchronize()
    assert a1.arena.stats[5].item() == a2.arena.stats[5].item() > 8192
    for net in ("actor", "fnn1", "fnn2", "value"):
        for name, t in a1.arena.named(a1.arena.grad, net).items():
            assert_close(a2.arena.named(a2.arena.grad, net)[name].cpu().numpy(), t.cpu().numpy(), 5e-5, f"chunked grad {net}.{name}")


def test_ant_shard_runs(gb):
    # Ant dims (S=111, A=8), one rank's shard of the global 16384 x 256 batch on 8 GPUs would be
    # B=2048, K=256; here a smaller shard exercises multi-chunk sampling (rows > 131072) and filtering
    S, A, B, K = 111, 8, 768, 256
    ag = _agent(gb, S, A, B, K, engine=0, seed=5)
    batch, noise = _inputs(S, A, B, K, 13)
    ag.train_on_batch(batch, noise)
    torch.cuda.synchronize()
    st = ag.arena.stats.cpu().numpy()
    assert np.isfinite(st).all() and bool(torch.isfinite(ag.arena.online).all())
    assert ag.engine.cfg.B * ag.engine.cfg.K * 2 > 131072
