"""GPU parity tests, stage by stage, through the C ABI: every kernel group is
fed the ORACLE's inputs for that stage and compared norm-wise at 1e-5 (forwards) / 1e-4 (losses,
gradients); the selection mask / indices and the Polyak update are bit-exact."""
import copy

import numpy as np
import pytest
import torch

from conftest import FWD_RTOL, GRAD_RTOL, assert_close, normwise_err
from oracle import gac_oracle as O

pytestmark = pytest.mark.gpu

ENGINES = [1, 0]   # 1 = SIMT fp32, 0 = tcgen05 3xTF32 (falls back per-shape to SIMT where unsupported)


def dev(x, dtype=torch.float32):
    return torch.as_tensor(np.asarray(x)).to("cuda", dtype).contiguous()


@pytest.fixture(scope="module")
def gb():
    import __graft_entry__ as ge
    ge.build()
    import dpo_replication_b200 as g
    return g


def make_engine(gb, S, A, B, K, engine, chunk_rows=0):
    return gb.Engine(S, A, B, K, "cuda", chunk_rows=chunk_rows, engine=engine)


def load_actor(gb, S, A, p):
    ar = gb.Arena(S, A, torch.device("cuda"))
    ar.load(ar.online, "actor", p)
    return ar


# ---------------------------------------------------------------------------------------------
def test_polyak_bit_exact(gb):
    eng = make_engine(gb, 3, 1, 4, 2, 1)
    rng = np.random.default_rng(0)
    for n in (1, 3, 4, 5, 1023, 4096, 100003):
        for rho in (5e-3, 1.0, 0.0, 0.37):
            t, s = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
            td, sd = dev(t), dev(s)
            eng.polyak(td, sd, rho)
            assert np.array_equal(td.cpu().numpy(), O.polyak(t, s, rho)), (n, rho)
    # unaligned views take the scalar path
    t, s = rng.standard_normal(1001).astype(np.float32), rng.standard_normal(1001).astype(np.float32)
    td, sd = dev(t), dev(s)
    eng.polyak(td[1:], sd[1:], 0.25)
    assert np.array_equal(td.cpu().numpy()[1:], O.polyak(t[1:], s[1:], 0.25)) and td[0].item() == t[0]


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("shape", [(128, 128, 64), (200, 400, 17), (384, 1200, 400), (130, 300, 23), (1, 400, 400), (257, 96, 1200)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_gemm(gb, engine, shape, ta, tb):
    M, N, K = shape
    eng = make_engine(gb, 3, 1, 4, 2, 1)
    rng = np.random.default_rng(M * 7 + N + K)
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    ref = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    try:
        c = eng.gemm(dev(a), dev(b), bool(ta), bool(tb), engine=engine)
    except gb._lib.GacError as e:
        if engine == 0 and "not supported" in str(e):
            pytest.skip("shape not on the tcgen05 engine")
        raise
    # SIMT fp32: ~1e-7 * sqrt(K).  tcgen05 3xTF32: the tensor core truncates on accumulate, one
    # truncation per 8-wide K step on the main accumulator => bound grows with K.
    tol = 2e-6 if engine == 1 else max(2e-6, 6e-9 * K)
    assert_close(c.cpu().numpy(), ref, tol, f"gemm {shape} ta={ta} tb={tb} engine={engine}")


@pytest.mark.parametrize("engine", ENGINES)
def test_adam_polyak(gb, engine):
    S, A = 5, 2
    eng = make_engine(gb, S, A, 4, 2, engine)
    ar = gb.Arena(S, A, torch.device("cuda"))
    rng = np.random.default_rng(1)
    n = ar.n
    host = {k: rng.standard_normal(n).astype(np.float32) for k in ("online", "target", "adam_m", "grad")}
    host["adam_v"] = rng.uniform(0, 1, n).astype(np.float32)
    for k, v in host.items():
        getattr(ar, k)[:n].copy_(dev(v))
    ar.adam_t.copy_(torch.tensor([3, 0, 7, 1], dtype=torch.int32))
    scale = np.array([0.5, 1.0, 1.0, 2.0], np.float32)
    skip = np.array([0, 0, 1, 0], np.int32)
    hp = eng.hparams(rho=0.01)
    eng.adam_polyak(ar.online, ar.target, ar.adam_m, ar.adam_v, ar.grad, hp, ar.adam_t, dev(scale), dev(skip, torch.int32))
    torch.cuda.synchronize()
    assert ar.adam_t.cpu().tolist() == [4, 1, 7, 2]                # skipped net keeps its counter (App. A.8)
    for i, net in enumerate(("actor", "fnn1", "fnn2", "value")):
        a, b = ar.net_range(net)
        p, g, m, v, t = (host[k][a:b] for k in ("online", "grad", "adam_m", "adam_v", "target"))
        if skip[i]:
            pn, mn, vn = p, m, v
        else:
            pn, mn, vn = O.adam_tf(p, g * scale[i], m, v, [4, 1, 7, 2][i])
        assert_close(ar.online[a:b].cpu().numpy(), pn, 1e-6, net + " p")
        assert_close(ar.adam_m[a:b].cpu().numpy(), mn, 1e-6, net + " m")
        assert_close(ar.adam_v[a:b].cpu().numpy(), vn, 1e-6, net + " v")
        got_t = ar.target[a:b].cpu().numpy()
        assert np.array_equal(got_t, O.polyak(t, ar.online[a:b].cpu().numpy(), 0.01)), net + " target (bit-exact given p)"


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("S,A,rows,nstates", [(3, 1, 64, 64), (17, 6, 300, 10), (11, 3, 129, 129), (40, 17, 256, 16)])
def test_aiqn_sampler(gb, engine, S, A, rows, nstates):
    rng = np.random.default_rng(2)
    p = O.init_actor(rng, S, A, bias_scale=0.05)
    eng = make_engine(gb, S, A, max(rows, 128), 1, engine)
    ar = load_actor(gb, S, A, p)
    su = rng.standard_normal((nstates, S)).astype(np.float32)
    sidx = rng.integers(0, nstates, rows).astype(np.int32) if nstates != rows else np.arange(rows, dtype=np.int32)
    taus = rng.uniform(size=(rows, A)).astype(np.float32)
    got = eng.aiqn_sample(ar.net_slice(ar.online, "actor"), dev(su), dev(sidx, torch.int32), dev(taus)).cpu().numpy()
    ref64, trace = O.aiqn_sample(p, su[sidx], taus, np.float64, trace=True)
    ref32 = O.aiqn_sample(p, su[sidx], taus, np.float32)
    # dimension 0 is strict; later dimensions are bounded by the oracle's own fp32-vs-fp64
    # divergence (the free-running sampler is chaotic: cos(i*pi*a) feedback)
    assert np.abs(got[:, 0] - ref64[:, 0]).max() <= FWD_RTOL
    for d in range(1, A):
        # envelope = the oracle's own fp32-vs-fp64 divergence on the same inputs; the floor covers rows where that
        # divergence happens to be tiny and is capped so that it can never become vacuous on values in (-1, 1)
        env = max(5.0 * np.abs(ref32[:, d] - ref64[:, d]).max(), min(1e-5 * 10 ** d, 1e-2))
        assert np.abs(got[:, d] - ref64[:, d]).max() <= env, (d, env)
    # stage-wise: teacher-force the ORACLE's own actions => every step sees oracle inputs
    sup = eng.aiqn_supervised_fwd(ar.net_slice(ar.online, "actor"), dev(su), dev(sidx, torch.int32), dev(taus),
                                  dev(ref64.astype(np.float32))).cpu().numpy()
    sup_ref, _ = O.aiqn_supervised(p, su[sidx], taus, ref64.astype(np.float32), np.float64)
    assert_close(sup, sup_ref, FWD_RTOL, "teacher-forced forward")


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("bounded", [False, True])
@pytest.mark.parametrize("rows,nstates,first", [(300, 10, False), (300, 10, True), (1000, 7, False), (40, 40, False)])
def test_gru_step_entry(gb, engine, rows, nstates, first, bounded):
    """gac_gru_step against the oracle's Keras-GRU cell (reset_after, [z|r|h]) on identical inputs."""
    S, A = 17, 6
    rng = np.random.default_rng(5)
    p = O.init_actor(rng, S, A, bias_scale=0.05)
    eng = make_engine(gb, S, A, 1024, 1, engine)
    ar = load_actor(gb, S, A, p)
    su = rng.standard_normal((nstates, S)).astype(np.float32)
    sidx = rng.integers(0, nstates, rows).astype(np.int32)
    if bounded:    # the actor's own embedding (|ae| <= sum|wa| + |ba|): the 3xFP16 variant may run
        ae = O.cosine_basis_linear(rng.uniform(-1, 1, (rows, 1)).astype(np.float32), p["wa"], p["ba"], np.float32, 0.2)[:, 0, :]
    else:          # arbitrary caller data, far beyond that bound: 3xTF32 variant
        ae = (5.0 * rng.standard_normal((rows, 400))).astype(np.float32)
    h = np.tanh(rng.standard_normal((rows, 400))).astype(np.float32)
    p64 = O.cast(p, np.float64)
    se, _ = O.state_embed(p64, su.astype(np.float64))
    x = np.concatenate([se[sidx], ae.astype(np.float64)], axis=1)
    ref, _ = O.gru_step(p64, x, np.zeros((rows, 400)) if first else h.astype(np.float64))
    actor = ar.net_slice(ar.online, "actor")
    got = eng.gru_step(actor, dev(su), dev(sidx, torch.int32), dev(ae), None if first else dev(h), actor_inputs=bounded)
    assert_close(got.cpu().numpy(), ref, FWD_RTOL, "gru step")
    got2 = eng.gru_step(actor, dev(su), dev(sidx, torch.int32), dev(ae), None if first else dev(h), reuse=True, actor_inputs=bounded)
    assert torch.equal(got, got2)


# The fused GRU-step kernel runs 3xFP16 with power-of-two scales derived from the weights (gru_step_tc.cuh): the
# result must not depend on the magnitudes of the four operands.  wa = ba = 0 gives a zero bound on the action
# embedding => no valid scaling => the in-kernel 3xTF32 fallback is the one that runs.
@pytest.mark.parametrize("s_kx,s_u,s_wa", [(1.0, 1.0, 1.0), (64.0, 1 / 64.0, 1.0), (1 / 256.0, 4.0, 8.0), (1.0, 1.0, 40.0),
                                           (1e-3, 1e-3, 1e-2), (1.0, 1.0, 0.0)])
def test_gru_step_operand_scales(gb, s_kx, s_u, s_wa):
    S, A, rows, nstates = 17, 6, 300, 10
    rng = np.random.default_rng(11)
    p = O.init_actor(rng, S, A, bias_scale=0.05)
    p = {k: np.array(v, copy=True) for k, v in p.items()}
    p["kx"] *= np.float32(s_kx); p["u"] *= np.float32(s_u); p["wa"] *= np.float32(s_wa); p["ba"] *= np.float32(s_wa)
    eng = make_engine(gb, S, A, 512, 1, 0)
    ar = load_actor(gb, S, A, p)
    su = rng.standard_normal((nstates, S)).astype(np.float32)
    sidx = rng.integers(0, nstates, rows).astype(np.int32)
    taus = rng.uniform(size=(rows, A)).astype(np.float32)
    teacher = rng.uniform(-1, 1, size=(rows, A)).astype(np.float32)
    sup = eng.aiqn_supervised_fwd(ar.net_slice(ar.online, "actor"), dev(su), dev(sidx, torch.int32), dev(taus), dev(teacher)).cpu().numpy()
    sup_ref, _ = O.aiqn_supervised(p, su[sidx], taus, teacher, np.float64)
    assert np.isfinite(sup).all()
    assert_close(sup, sup_ref, FWD_RTOL, f"teacher-forced forward, scales {s_kx} {s_u} {s_wa}")


# the arithmetic variant is chosen once per process from the environment: re-run the actor tests in a child process
@pytest.mark.parametrize("env", ["GAC_GRU_V1", "GAC_GRU_TF32", "GAC_NO_FUSED_GRU", "GAC_NO_FUSED_HEAD", "GAC_HEAD_TF32", "GAC_NO_F16_DGRAD", "GAC_NO_F16_WGRAD", "GAC_NO_FUSED_EMB", "GAC_NO_FUSED_CRITIC", "GAC_DGRAD_TF32",
                                 "GAC_NO_SMALL_SPLITK", "GAC_GRU_PAIR", "GAC_NO_SIDE_STREAM"])
def test_actor_paths_other_variants(env):
    import os
    import subprocess
    import sys
    e = dict(os.environ)
    e[env] = "1"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", os.path.join(root, "tests", "test_kernels_gpu.py"),
                        os.path.join(root, "tests", "test_step_gpu.py"), "-k",
                        "aiqn_sampler or actor_loss_and_bptt or golden_stage or gru_step_operand or gradient_scales or td_fwd_bwd or "
                        "pendulum_step_end_to_end or halfcheetah_dims"], env=e, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("engine", ENGINES)
def test_critic_and_value_eval(gb, engine):
    S, A, rows, nst = 17, 6, 777, 40
    rng = np.random.default_rng(3)
    p1, p2, pv = O.init_fnn(rng, S + A, 0.05), O.init_fnn(rng, S + A, 0.05), O.init_fnn(rng, S, 0.05)
    eng = make_engine(gb, S, A, 1024, 1, engine)
    ar = gb.Arena(S, A, torch.device("cuda"))
    ar.load(ar.target, "fnn1", p1); ar.load(ar.target, "fnn2", p2); ar.load(ar.target, "value", pv)
    su = rng.standard_normal((nst, S)).astype(np.float32)
    sidx = rng.integers(0, nst, rows).astype(np.int32)
    acts = rng.uniform(-1, 1, (rows, A)).astype(np.float32)
    q = eng.critic_eval(ar.net_slice(ar.target, "fnn1"), ar.net_slice(ar.target, "fnn2"), dev(su), dev(sidx, torch.int32), dev(acts))
    assert_close(q.cpu().numpy(), O.critic_forward(p1, p2, su[sidx], acts, np.float64), FWD_RTOL, "critic")
    v = eng.value_eval(ar.net_slice(ar.target, "value"), dev(su))
    assert_close(v.cpu().numpy(), O.value_forward(pv, su, np.float64), FWD_RTOL, "value")


@pytest.mark.parametrize("n", [1, 31, 1024, 1025, 5000, 65536])
def test_advantage_compaction_bit_exact(gb, n):
    A = 3
    eng = make_engine(gb, 4, A, max(n // 2 + 1, 4), 1, 1)
    rng = np.random.default_rng(n)
    q = rng.standard_normal(n).astype(np.float32)
    v = rng.standard_normal(n).astype(np.float32)
    if n >= 31:
        v[3] = q[3]                       # tie => excluded (strict >)
        q[5] = np.nan; v[7] = np.nan      # NaN => excluded
        q[11] = np.inf; v[13] = -np.inf
    acts = rng.uniform(-1, 1, (n, A)).astype(np.float32)
    sidx = rng.integers(0, 100, n).astype(np.int32)
    r = eng.advantage_compact(dev(q), dev(v), dev(acts), dev(sidx, torch.int32))
    mask, idx = O.positive_advantage_filter(q, v)
    M = int(r["m"].item())
    assert M == idx.shape[0]
    got = r["idx"][:M].cpu().numpy()
    assert got.dtype == np.int64 and np.array_equal(got, idx[:, 0])
    assert np.array_equal(r["actions"][:M].cpu().numpy(), acts[idx[:, 0]])
    assert np.array_equal(r["state_idx"][:M].cpu().numpy(), sidx[idx[:, 0]])
    with np.errstate(invalid="ignore"):
        adv = (q - v)[idx[:, 0]]
    assert np.array_equal(r["adv"][:M].cpu().numpy(), adv)
    pos = r["pos"].cpu().numpy()
    assert np.array_equal(pos, np.concatenate([[0], np.cumsum(mask)[:-1]]))
    fin = np.isfinite(adv)
    if fin.all():
        assert abs(r["sum_adv"].item() - adv.astype(np.float64).sum()) <= 1e-5 * max(np.abs(adv).sum(), 1e-30)


def test_compaction_none_and_all(gb):
    eng = make_engine(gb, 4, 2, 2048, 1, 1)
    z, o = dev(np.zeros(3000, np.float32)), dev(np.ones(3000, np.float32))
    r = eng.advantage_compact(z, o)
    assert int(r["m"].item()) == 0 and r["sum_adv"].item() == 0.0          # M == 0 is not an error (agent.py:158)
    r = eng.advantage_compact(o, z)
    assert int(r["m"].item()) == 3000 and np.array_equal(r["idx"].cpu().numpy(), np.arange(3000))


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("n_in,rows", [(23, 256), (17, 100), (3, 7)])
def test_td_fwd_bwd(gb, engine, n_in, rows):
    rng = np.random.default_rng(4)
    p = O.init_fnn(rng, n_in, 0.05)
    S = n_in
    eng = make_engine(gb, S, 1, max(rows, 128), 1, engine)
    ar = gb.Arena(S, 1, torch.device("cuda"))
    ar.load(ar.online, "value", p)
    x, y = rng.standard_normal((rows, n_in)).astype(np.float32), rng.standard_normal((rows, 1)).astype(np.float32)
    gslice = ar.net_slice(ar.grad, "value")
    loss = eng.td_fwd_bwd(ar.net_slice(ar.online, "value"), n_in, dev(x), dev(y.reshape(-1)), gslice)
    ref = O.value_train_grads(p, x, y, np.float64)
    assert abs(loss.item() - ref.pop("_loss")) <= GRAD_RTOL * abs(loss.item())
    got = ar.named(ar.grad, "value")
    for k, g in ref.items():
        assert_close(got[k].cpu().numpy().reshape(g.shape), g, GRAD_RTOL, f"fnn grad {k}")


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("S,A,M,nst,live", [(3, 1, 64, 16, None), (17, 6, 200, 24, None), (9, 3, 256, 256, 131), (5, 2, 128, 8, 0),
                                           (40, 17, 256, 16, 200)])
def test_actor_loss_and_bptt(gb, engine, S, A, M, nst, live):
    rng = np.random.default_rng(5)
    p = O.init_actor(rng, S, A, bias_scale=0.05)
    eng = make_engine(gb, S, A, max(M, 128), 1, engine)
    ar = load_actor(gb, S, A, p)
    su = rng.standard_normal((nst, S)).astype(np.float32)
    sidx = rng.integers(0, nst, M).astype(np.int32)
    taus = rng.uniform(size=(M, A)).astype(np.float32)
    acts = rng.uniform(-1, 1, (M, A)).astype(np.float32)
    adv = rng.uniform(0.05, 1.0, M).astype(np.float32)
    n = M if live is None else live
    d_rows = None if live is None else dev(np.array([live]), torch.int32)
    base = ar.net_slice(ar.online, "actor")
    pred = eng.aiqn_supervised_fwd(base, dev(su), dev(sidx, torch.int32), dev(taus), dev(acts), d_rows)
    if n == 0:
        loss, dpred = eng.qhuber_loss_bwd(pred, dev(acts), dev(taus), dev(adv), 1.0, d_rows)
        eng.aiqn_bwd(base, dpred, ar.net_slice(ar.grad, "actor"))
        torch.cuda.synchronize()
        assert loss.item() == 0.0
        assert float(ar.net_slice(ar.grad, "actor").abs().max()) == 0.0     # no live row => zero gradient
        return
    g_ref, loss_ref, pred_ref, dpred_ref = O.actor_train_grads(p, su[sidx][:n], acts[:n], adv[:n, None], taus[:n], "linear", 1.0, np.float64)
    assert_close(pred.cpu().numpy()[:n], pred_ref, FWD_RTOL, "supervised forward")
    w = adv[:n] / adv[:n].astype(np.float64).sum()
    wfull = np.zeros(M, np.float32); wfull[:n] = w
    loss, dpred = eng.qhuber_loss_bwd(pred, dev(acts), dev(taus), dev(wfull), 1.0 / (n * A), d_rows)
    assert abs(loss.item() - loss_ref) <= GRAD_RTOL * abs(loss_ref)
    assert_close(dpred.cpu().numpy()[:n], dpred_ref, GRAD_RTOL, "dL/dpred")
    # feed the ORACLE's dL/dpred to the backward (stage-wise parity)
    dfull = np.zeros((M, A), np.float32); dfull[:n] = dpred_ref
    eng.aiqn_bwd(base, dev(dfull), ar.net_slice(ar.grad, "actor"))
    got = ar.named(ar.grad, "actor")
    for k, g in g_ref.items():
        assert_close(got[k].cpu().numpy().reshape(g.shape), g, GRAD_RTOL, f"actor grad {k}")
    # accumulate semantics: a second backward with accumulate doubles the gradient
    eng.aiqn_bwd(base, dev(dfull), ar.net_slice(ar.grad, "actor"), accumulate=True)
    assert_close(ar.named(ar.grad, "actor")["u"].cpu().numpy(), 2 * g_ref["u"], GRAD_RTOL, "accumulate")


# The backward's input- and weight-gradient GEMMs run 3xFP16 with scales taken from running maxima of the gradient tensors
# (chunk sizes that are multiples of 128): the result must be scale-covariant over many orders of magnitude, and all-zero
# gradients (no usable scale) must fall back cleanly.
@pytest.mark.parametrize("gscale", [1.0, 1e-7, 3e4, 0.0])
def test_actor_bptt_gradient_scales(gb, gscale):
    S, A, M, nst = 17, 6, 256, 24
    rng = np.random.default_rng(15)
    p = O.init_actor(rng, S, A, bias_scale=0.05)
    eng = make_engine(gb, S, A, 128, 1, 0)          # chunk = 2 * 128 * 1 = 256 rows
    ar = load_actor(gb, S, A, p)
    su = rng.standard_normal((nst, S)).astype(np.float32)
    sidx = rng.integers(0, nst, M).astype(np.int32)
    taus = rng.uniform(size=(M, A)).astype(np.float32)
    acts = rng.uniform(-1, 1, (M, A)).astype(np.float32)
    adv = rng.uniform(0.05, 1.0, M).astype(np.float32)
    base = ar.net_slice(ar.online, "actor")
    pred = eng.aiqn_supervised_fwd(base, dev(su), dev(sidx, torch.int32), dev(taus), dev(acts))
    g_ref, _, pred_ref, dpred_ref = O.actor_train_grads(p, su[sidx], acts, adv[:, None], taus, "linear", 1.0, np.float64)
    assert_close(pred.cpu().numpy(), pred_ref, FWD_RTOL, "supervised forward")
    eng.aiqn_bwd(base, dev((dpred_ref * gscale).astype(np.float32)), ar.net_slice(ar.grad, "actor"))
    got = ar.named(ar.grad, "actor")
    for k, g in g_ref.items():
        x = got[k].cpu().numpy().reshape(g.shape)
        assert np.isfinite(x).all(), k
        if gscale == 0.0:
            assert float(np.abs(x).max()) == 0.0, k
        else:
            assert_close(x / gscale, g, GRAD_RTOL, f"actor grad {k} at gradient scale {gscale}")


def test_golden_stage_vectors(gb):
    import importlib.util, os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    gm = importlib.util.module_from_spec(spec); spec.loader.exec_module(gm)
    z = np.load(os.path.join(here, "golden", "gac_step_small.npz"))
    nets, batch, noise, st, tau, tch = gm.inputs()
    S, A = gm.S, gm.A
    eng = make_engine(gb, S, A, 128, 1, 0)
    ar = gb.Arena(S, A, torch.device("cuda"))
    for net in ("actor", "fnn1", "fnn2", "value"):
        ar.load(ar.online, net, nets[net])
    sidx = dev(np.arange(8), torch.int32)
    base = ar.net_slice(ar.online, "actor")
    sup = eng.aiqn_supervised_fwd(base, dev(st), sidx, dev(tau), dev(tch))
    assert_close(sup.cpu().numpy(), z["stage.supervised"], FWD_RTOL, "golden supervised")
    smp = eng.aiqn_sample(base, dev(st), sidx, dev(tau)).cpu().numpy()
    assert np.abs(smp[:, 0] - z["stage.sample"][:, 0]).max() <= FWD_RTOL
    assert np.abs(smp - z["stage.sample"]).max() <= 1e-4
    q = eng.critic_eval(ar.net_slice(ar.online, "fnn1"), ar.net_slice(ar.online, "fnn2"), dev(st), None, dev(tch))
    assert_close(q.cpu().numpy(), z["stage.q"], FWD_RTOL, "golden critic")
    v = eng.value_eval(ar.net_slice(ar.online, "value"), dev(st))
    assert_close(v.cpu().numpy(), z["stage.v"], FWD_RTOL, "golden value")
