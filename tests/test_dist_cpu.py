"""world_size-2 gloo test (CPU) of the data-parallel host logic: sharding the
global replay batch by state, the SUM all-reduce of [grad arena | stats], and the deferred
actor normalisation reproduce the single-process gradients.  The oracle plays the compute."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gac_oracle as O

S, A, B, K = 4, 2, 8, 3


def _setup():
    rng = np.random.default_rng(11)
    nets = O.make_agent(rng, S, A, bias_scale=0.05)
    # centre V_target so that roughly half of the candidates have positive advantage
    batch, noise = O.make_batch(rng, S, A, B), O.make_noise(rng, A, B, K)
    flt = O.sample_positive_advantage_actions(nets, batch["s"], noise, K, np.float64)
    nets["target_value"]["b2"] = nets["target_value"]["b2"] + np.float32(np.median(flt["q"] - flt["v"]))
    return nets, batch, noise


def _shard_grads(nets, batch, noise, rank, world):
    """Per-rank un-normalised gradients exactly as gac_step_grads produces them."""
    from dpo_replication_b200 import dist as gd
    b, n = gd.shard_step_inputs(batch, noise, K, rank, world)
    f64 = np.float64
    cg = O.critic_train_grads(nets["fnn1"], nets["fnn2"], nets["target_value"], b["s"], b["a"], b["r"], b["sp"], b["it"], 0.99,
                              0.01, n["critic_u"], f64)
    tiled = O.tile_state_major(b["s"].astype(f64), K)
    va = O.aiqn_sample(nets["target_actor"], tiled, n["value_tau"], f64)
    vq = O.critic_forward(nets["target_fnn1"], nets["target_fnn2"], tiled, va, f64)
    vg = O.value_train_grads(nets["value"], b["s"], O.v_critic_from_q(vq, b["s"].shape[0], K), f64)
    vg.pop("_loss")
    flt = O.sample_positive_advantage_actions(nets, b["s"], n, K, f64)
    M = flt["idx"].shape[0]
    adv = flt["advantages"].reshape(-1, 1)
    if M:
        pred, cache = O.aiqn_supervised(nets["actor"], flt["good_states"], n["actor_tau"][:M], flt["good_actions"], f64)
        # un-normalised: weights = adv, no 1/(M*A): quantile_huber_loss divides by pred.size, undo it
        _, dpred = O.quantile_huber_loss(pred, flt["good_actions"], n["actor_tau"][:M].astype(f64), adv)
        ag = O.aiqn_supervised_backward(cache, dpred * pred.size)
    else:                                                     # a rank without positive-advantage rows adds zeros
        ag = {k: np.zeros(v, f64) for k, v in O.actor_shapes(S, A).items()}
    flat = np.concatenate([np.concatenate([g[k].reshape(-1) for k in sorted(g)]) for g in (ag, cg["fnn1"], cg["fnn2"], vg)])
    stats = np.array([adv.sum(), M], f64)
    return flat, stats, flt, gd


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nets, batch, noise = _setup()
        flat, stats, flt, gd = _shard_grads(nets, batch, noise, rank, world)
        t = torch.from_numpy(np.concatenate([flat, stats]))
        gd.allreduce_grads(t)
        glob = gd.local_to_global_candidate(flt["idx"][:, 0], B // world, K, rank, world)
        q.put((rank, t.numpy(), glob))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradients_match_single_process():
    import dpo_replication_b200  # noqa: F401 (needs the built library)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda x: x[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    # both ranks hold the same reduced buffer
    assert np.array_equal(res[0][1], res[1][1])
    # single-process reference on the global batch
    nets, batch, noise = _setup()
    # NOTE actor_tau is indexed by compacted row, which differs between the sharded and the
    # single-process order; compare the critic/value parts exactly and the filter index sets.
    cg = O.critic_train_grads(nets["fnn1"], nets["fnn2"], nets["target_value"], batch["s"], batch["a"], batch["r"], batch["sp"],
                              batch["it"], 0.99, 0.01, noise["critic_u"], np.float64)
    red = res[0][1]
    n_actor = sum(int(np.prod(s)) for s in O.actor_shapes(S, A).values())
    n_fnn = sum(int(np.prod(s)) for s in O.fnn_shapes(S + A).values())
    got_f1 = red[n_actor:n_actor + n_fnn]
    ref_f1 = np.concatenate([cg["fnn1"][k].reshape(-1) for k in sorted(cg["fnn1"])])
    assert np.abs(got_f1 - ref_f1).max() <= 1e-9 * np.abs(ref_f1).max()      # SUM, not mean (F5)
    flt = O.sample_positive_advantage_actions(nets, batch["s"], noise, K, np.float64)
    glob = np.sort(np.concatenate([res[0][2], res[1][2]]))
    assert np.array_equal(glob, flt["idx"][:, 0])                           # index parity after local->global mapping
    assert red[-1] == flt["idx"].shape[0] > 0                               # M_g
    assert abs(red[-2] - flt["advantages"].sum()) < 1e-9 * abs(flt["advantages"].sum())


def test_shard_layouts():
    from dpo_replication_b200 import dist as gd
    rng = np.random.default_rng(0)
    batch, noise = O.make_batch(rng, S, A, B), O.make_noise(rng, A, B, K)
    for rank in range(2):
        b, n = gd.shard_step_inputs(batch, noise, K, rank, 2)
        sl = gd.shard_rows(B, rank, 2)
        # value_tau rows follow their states (state-major)
        assert np.array_equal(n["value_tau"], noise["value_tau"].reshape(B, K, A)[sl].reshape(-1, A))
        # filter rows: local row k*B_l + b  <->  global row k*B + (rank*B_l + b)
        Bl = B // 2
        for k in range(K):
            for bb in range(Bl):
                assert np.array_equal(n["filter_tau"][k * Bl + bb], noise["filter_tau"][k * B + sl.start + bb])
    with pytest.raises(ValueError):
        gd.shard_rows(7, 0, 2)
