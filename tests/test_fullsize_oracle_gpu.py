"""Stage-wise parity at BASELINE.json's FULL sizes (HalfCheetah B=256 K=64, Humanoid B=256 K=128) against the oracle's
torch mirror evaluated in fp64 on the GPU (tests/fp64_checker.py): the NumPy oracle needs minutes at these sizes, B200
fp64 needs well under a second.  Every stage of one real train step is fed the SAME inputs on both sides:

  sampler      dimension 0 free-running and every dimension teacher-forced on the CUDA path's own previous actions
               (the free-running sampler is chaotic), 1e-5;
  critic/value candidate q and v on the CUDA path's own candidate actions, 1e-5;
  filter       mask / int64 indices bit-exact given those q, v;
  TD nets      all 18 critic / value gradients of the step, 1e-4;
  actor        all 13 gradients, 1e-4, through the stage entry points on the step's own selected rows that are clear of
               every LeakyReLU / Huber-indicator kink (see fp64_checker.py), and the step's own un-normalised gradient
               over ALL selected rows at 1e-3.
"""
import numpy as np
import pytest
import torch

import fp64_checker as CK
from conftest import FWD_RTOL, GRAD_RTOL, assert_close
from oracle import gac_oracle as O
from oracle import torch_ref as T

pytestmark = pytest.mark.gpu
# Relative distance from a kink below which a row is left out of the strict gradient comparison: ~10x the CUDA path's
# forward rounding error at these sizes (5e-7 of a tensor's largest entry), i.e. rows on which a correct fp32 forward may
# still take the other branch of a LeakyReLU / of the Huber indicator than the fp64 checker does.
KINK = 5e-6
TD_KINK = 3e-6     # the 256-row TD nets: (256 x 700) pre-activations, re-seeded until none lies inside the band


@pytest.fixture(scope="module")
def gb():
    import __graft_entry__ as ge
    ge.build()
    import dpo_replication_b200 as g
    return g


def _nets(S, A, B, K, seed):
    """Oracle-initialised agent + batch whose 256 TD rows are clear of the FNN kinks: rows with a pre-activation inside the
    band (checked in fp64) are re-drawn until none is left."""
    rng = np.random.default_rng(seed)
    nets = O.make_agent(rng, S, A, bias_scale=0.05)
    batch, noise = O.make_batch(rng, S, A, B), O.make_noise(rng, A, B, K)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device="cuda")
    p64 = {n: CK.params64(nets[n], "cuda") for n in ("fnn1", "fnn2", "value")}
    for attempt in range(100):
        x = torch.cat([t(batch["s"]), torch.clamp(t(batch["a"]) + (t(noise["critic_u"]) * 2 - 1) * 0.01, -1, 1)], dim=-1)
        margin = torch.minimum(torch.minimum(CK.fnn_margin(p64["fnn1"], x), CK.fnn_margin(p64["fnn2"], x)), CK.fnn_margin(p64["value"], t(batch["s"])))
        bad = torch.nonzero(margin <= TD_KINK).reshape(-1).cpu().numpy()
        if bad.size == 0:
            return nets, batch, noise
        batch["s"][bad] = rng.standard_normal((bad.size, S)).astype(np.float32)
        batch["a"][bad] = rng.uniform(-1, 1, (bad.size, A)).astype(np.float32)
    raise AssertionError("could not clear the TD batch of kink rows")


@pytest.mark.parametrize("name,S,A,B,K", [("halfcheetah", 17, 6, 256, 64), ("humanoid", 376, 17, 256, 128)])
def test_full_size_stagewise_vs_fp64(gb, name, S, A, B, K):
    P = B * K
    dev = torch.device("cuda")
    nets, batch, noise = _nets(S, A, B, K, 4242)
    t64 = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    t32 = lambda a: torch.as_tensor(a, dtype=torch.float32, device=dev)
    agent = gb.GACAgent(A, S, action_samples=K, batch_size=B, device="cuda", engine=0)
    agent.load_state(nets)
    bd = {k: t32(v) for k, v in batch.items()}
    nd = {k: t32(v) for k, v in noise.items()}
    taps = dict(sel_idx=torch.zeros(2 * P, dtype=torch.int64, device=dev), q_out=torch.zeros(2 * P, device=dev),
                v_out=torch.zeros(2 * P, device=dev), cand_actions=torch.zeros(2 * P, A, device=dev),
                value_actions=torch.zeros(P, A, device=dev))
    eng, ar = agent.engine, agent.arena
    # centre V_target (about half of the candidates selected), identically on both sides
    eng.step_grads(eng.step_io(ar, bd, nd, taps), agent._hp)
    shift = np.float32((taps["q_out"] - taps["v_out"]).median().item())
    nets["target_value"]["b2"] = nets["target_value"]["b2"] + shift
    agent.load_state(nets)
    eng.step_grads(eng.step_io(ar, bd, nd, taps), agent._hp)
    torch.cuda.synchronize()
    st = ar.stats.cpu().numpy().astype(np.float64)
    M = int(st[5])
    assert 0.2 * 2 * P < M < 0.8 * 2 * P, M

    s64 = t64(batch["s"])
    ta, tf1, tf2, tv = (CK.params64(nets["target_" + n], dev) for n in ("actor", "fnn1", "fnn2", "value"))

    # ---- sampler (value pass, state-major tiling): teacher-forced on the CUDA path's own actions ----
    va = taps["value_actions"]
    tiled_sm = CK.tile_state_major(s64, K)
    with torch.no_grad():
        sup, _ = CK.supervised_margins(ta, tiled_sm, t64(noise["value_tau"]), va.double())
    assert_close(va[:, 0].cpu().numpy(), sup[:, 0].cpu().numpy(), FWD_RTOL, f"{name}: sampler dim 0 (free-running)")
    assert_close(va.cpu().numpy(), sup.cpu().numpy(), FWD_RTOL, f"{name}: sampler, every dim teacher-forced on its own previous actions")

    # ---- candidate q, v on the CUDA path's candidates (sample-major tiling, [target-policy ; uniform]) ----
    cand = taps["cand_actions"]
    assert float(cand.abs().max()) <= 1.0
    assert torch.equal(cand[P:], nd["filter_rand"])
    tiled2 = CK.tile_sample_major(s64, K).repeat(2, 1)
    with torch.no_grad():
        q_ref = T.critic_t(tf1, tf2, tiled2, cand.double()).reshape(-1)
        v_ref = T.fnn_t(tv, tiled2).reshape(-1)
    assert_close(taps["q_out"].cpu().numpy(), q_ref.cpu().numpy(), FWD_RTOL, f"{name}: candidate q")
    assert_close(taps["v_out"].cpu().numpy(), v_ref.cpu().numpy(), FWD_RTOL, f"{name}: candidate v")

    # ---- filter: bit-exact given q, v ----
    q, v = taps["q_out"].cpu().numpy(), taps["v_out"].cpu().numpy()
    _, idx = O.positive_advantage_filter(q, v)
    assert M == idx.shape[0] and np.array_equal(taps["sel_idx"].cpu().numpy()[:M], idx[:, 0])
    assert abs(st[4] - (q.astype(np.float64) - v)[idx[:, 0]].sum()) <= 1e-5 * abs(st[4])

    # ---- TD nets: all critic / value gradients of the step ----
    x = torch.cat([s64, torch.clamp(t64(batch["a"]) + (t64(noise["critic_u"]) * 2 - 1) * 0.01, -1, 1)], dim=-1)
    with torch.no_grad():
        yq = t64(batch["r"]) + 0.99 * T.fnn_t(tv, t64(batch["sp"])) * (1 - t64(batch["it"]))
        # Value.train target: mean over K of the per-sample min-critic on the value-pass actions (networks.py:483-485)
        v_crit = T.critic_t(tf1, tf2, tiled_sm, va.double()).reshape(B, K, 1).mean(dim=1)
    for i, (net, xin, y) in enumerate((("fnn1", x, yq), ("fnn2", x, yq), ("value", s64, v_crit))):
        g_ref, loss_ref = CK.td_grads(CK.params64(nets[net], dev, grad=True), xin, y)
        assert abs(st[i] - loss_ref) <= GRAD_RTOL * abs(loss_ref), (net, st[i], loss_ref)
        got = ar.named(ar.grad, net)
        for k, g in g_ref.items():
            assert_close(got[k].cpu().numpy().reshape(g.shape), g.cpu().numpy(), GRAD_RTOL, f"{name}: grad {net}.{k}")

    # ---- actor: the step's own selected rows ----
    sel = torch.as_tensor(idx[:, 0], device=dev)
    gs, ga = tiled2[sel], cand[sel].double()
    adv = (taps["q_out"].double() - taps["v_out"].double())[sel]
    taus = t64(noise["actor_tau"][:M])
    pa = CK.params64(nets["actor"], dev, grad=True)
    w_all = adv / adv.sum()
    g_all, loss_all = CK.actor_grads(pa, gs, taus, ga, w_all)
    scale = 1.0 / (M * A * st[4])
    assert abs(st[3] * scale - loss_all) <= GRAD_RTOL * abs(loss_all)
    got = ar.named(ar.grad, "actor")
    for k, g in g_all.items():
        # the whole selection includes rows on a kink: discrete differences of ~1/sqrt(rows) of a tensor (see fp64_checker.py)
        assert_close(got[k].double().cpu().numpy().reshape(g.shape) * scale, g.cpu().numpy(), 1e-3, f"{name}: step actor grad {k} (all {M} rows)")
    # strict: rows clear of every kink, through the stage entry points, same rows on both sides
    with torch.no_grad():
        margin = torch.cat([CK.supervised_margins(pa, gs[r0:r0 + 8192], taus[r0:r0 + 8192], ga[r0:r0 + 8192])[1] for r0 in range(0, M, 8192)])
    keep = torch.nonzero(margin > KINK).reshape(-1)
    Mk = int(keep.numel())
    assert Mk > 0.5 * M, (Mk, M)
    sidx_all = torch.arange(B, dtype=torch.int32, device=dev).repeat(2 * K)          # candidate row -> state (sample-major, twice)
    sidx_k = sidx_all[sel][keep].contiguous()
    teacher_k, taus_k = cand[sel][keep].contiguous(), nd["actor_tau"][:M][keep].contiguous()
    w_k = (adv[keep] / adv[keep].sum())
    base = ar.net_slice(ar.online, "actor")
    pred = eng.aiqn_supervised_fwd(base, bd["s"], sidx_k, taus_k, teacher_k)
    with torch.no_grad():
        pred_ref, _ = CK.supervised_margins(pa, gs[keep], taus[keep], ga[keep])
    assert_close(pred.cpu().numpy(), pred_ref.cpu().numpy(), FWD_RTOL, f"{name}: supervised forward ({Mk} rows)")
    loss, dpred = eng.qhuber_loss_bwd(pred, teacher_k, taus_k, w_k.float().contiguous(), 1.0 / (Mk * A))
    gbuf = torch.zeros_like(ar.grad)
    eng.aiqn_bwd(base, dpred, ar.net_slice(gbuf, "actor"))
    torch.cuda.synchronize()
    g_k, loss_k = CK.actor_grads(pa, gs[keep], taus[keep], ga[keep], w_k)
    assert abs(loss.item() - loss_k) <= GRAD_RTOL * abs(loss_k)
    got = ar.named(gbuf, "actor")
    for k, g in g_k.items():
        assert_close(got[k].cpu().numpy().reshape(g.shape), g.cpu().numpy(), GRAD_RTOL, f"{name}: actor grad {k} ({Mk} kink-free rows of {M})")
