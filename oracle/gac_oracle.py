"""CPU oracle for the GAC train step of gwbcho/dpo-replication.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dpo_replication_b200/`` may import this
module; it is used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker.

PARITY UNPINNED: the reference's arithmetic lives in TensorFlow 2.0, which is not
installable in this image, and the reference's own tests assert shapes only
(``tests/unit/GAC/networks.py``).  This file restates the algorithm formula by formula
 in NumPy.  It is cross-checked against an independent torch-CPU
composition (``oracle/torch_ref.py``: ``torch.nn.GRU`` + autograd) and against fp64
finite differences in ``tests/test_oracle.py``.

Every function takes a ``dtype`` (np.float32 mirrors the reference's precision,
np.float64 is the high-precision truth used for error envelopes).  All randomness of
the reference (7 ``tf.random`` sites) is an explicit input.

Variable naming (one dict per network):
  actor : ws (S,400) bs (400)            state_embedding          networks.py:159-162
          wn (64,400) bn (400)           noise_embedding          networks.py:165
          wa (64,400) ba (400)           action_embedding         networks.py:166
          kx (800,1200) u (400,1200) gb (2,1200)   rnn (GRU)      networks.py:169
          w1 (400,400) b1 (400)          dense_layer_1            networks.py:171
          w2 (400,1)  b2 (1)             dense_layer_2            networks.py:174
  fnn   : w0 (in,400) b0 (400) w1 (400,300) b1 (300) w2 (300,1) b2 (1)   networks.py:326-358
"""
from __future__ import annotations

import numpy as np

E = 400          # embedding / hidden width (networks.py:160,165,166,169,171)
H1, H2 = 400, 300  # FNN widths (networks.py:379-380,438)
NB = 64          # cosine basis functions (networks.py:145)

ACTOR_KEYS = ("wa", "ba", "w1", "b1", "w2", "b2", "wn", "bn", "kx", "u", "gb", "ws", "bs")
FNN_KEYS = ("w0", "b0", "w1", "b1", "w2", "b2")


# --------------------------------------------------------------------------------------
# shapes / init
# --------------------------------------------------------------------------------------
def actor_shapes(S, A):
    return {"wa": (NB, E), "ba": (E,), "w1": (E, E), "b1": (E,), "w2": (E, 1), "b2": (1,),
            "wn": (NB, E), "bn": (E,), "kx": (2 * E, 3 * E), "u": (E, 3 * E), "gb": (2, 3 * E),
            "ws": (S, E), "bs": (E,)}


def fnn_shapes(n_in):
    return {"w0": (n_in, H1), "b0": (H1,), "w1": (H1, H2), "b1": (H2,), "w2": (H2, 1), "b2": (1,)}


def _glorot(rng, shape):
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def _orthogonal(rng, n):
    q, r = np.linalg.qr(rng.standard_normal((n, n)))
    return (q * np.sign(np.diag(r))).astype(np.float32)


def init_actor(rng, S, A, bias_scale=0.0):
    """Keras default inits: glorot-uniform kernels, orthogonal recurrent
    kernel per gate, zero biases (bias_scale>0 draws small non-zero biases so tests see them)."""
    p = {}
    for k, shp in actor_shapes(S, A).items():
        if k == "u":
            p[k] = np.concatenate([_orthogonal(rng, E) for _ in range(3)], axis=1)
        elif len(shp) == 2 and k != "gb":
            p[k] = _glorot(rng, shp)
        else:
            p[k] = (bias_scale * rng.standard_normal(shp)).astype(np.float32)
    return p


def init_fnn(rng, n_in, bias_scale=0.0):
    p = {}
    for k, shp in fnn_shapes(n_in).items():
        p[k] = _glorot(rng, shp) if len(shp) == 2 else (bias_scale * rng.standard_normal(shp)).astype(np.float32)
    return p


def cast(p, dtype):
    return {k: np.asarray(v, dtype=dtype) for k, v in p.items()}


# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------
def i_pi(n=NB):
    """networks.py:45 -- np.arange(1,n+1,float32) * np.pi: the product is formed in fp32
    from fl32(i) and fl32(pi)."""
    return np.arange(1, n + 1, dtype=np.float32) * np.float32(np.pi)


def lrelu(x, alpha):
    """tf.nn.leaky_relu / LeakyReLU layer: x>0 ? x : alpha*x (App. A.2)."""
    return np.where(x > 0, x, np.asarray(alpha, dtype=x.dtype) * x)


def lrelu_grad(x, alpha):
    return np.where(x > 0, np.asarray(1, dtype=x.dtype), np.asarray(alpha, dtype=x.dtype))


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def cosine_basis(x, dtype, n=NB):
    """networks.py:44-48.  The argument replays the reference's fp32 rounding sequence
    fl32(fl32(x) * i_pi) for every dtype; only the cosine itself is evaluated in ``dtype``."""
    x32 = np.asarray(x, dtype=np.float32).reshape(-1, 1)
    arg = x32 * i_pi(n)[None, :]                      # fp32 product, as in the reference
    return np.cos(arg.astype(dtype))


def cosine_basis_linear(x, w, b, dtype, activation_alpha=None):
    """CosineBasisLinear.__call__, networks.py:50-62.  x (rows, a) or (rows, a, 1) ->
    (rows, a, embed)."""
    rows = x.shape[0]
    out = cosine_basis(x, dtype, w.shape[0]) @ w + b
    if activation_alpha is not None:
        out = lrelu(out, activation_alpha)
    return out.reshape(rows, -1, w.shape[1])


def state_embed(p, states):
    """networks.py:196 / :243 -- Dense(400, LeakyReLU(0.01)) then tf.nn.leaky_relu (0.2)."""
    pre = states @ p["ws"] + p["bs"]
    return lrelu(lrelu(pre, 0.01), 0.2), pre


def gru_step(p, x, h):
    """tf.keras.layers.GRU v2 defaults, reset_after=True, gate order [z|r|h] (App. A.5)."""
    mx = x @ p["kx"] + p["gb"][0]
    mh = h @ p["u"] + p["gb"][1]
    z = sigmoid(mx[:, :E] + mh[:, :E])
    r = sigmoid(mx[:, E:2 * E] + mh[:, E:2 * E])
    c = np.tanh(mx[:, 2 * E:] + r * mh[:, 2 * E:])
    hn = z * h + (1 - z) * c
    return hn, (z, r, c, mh[:, 2 * E:])


def head(p, hid):
    """dense_layer_2(dense_layer_1(.)), networks.py:221 / :266."""
    y1p = hid @ p["w1"] + p["b1"]
    y1 = lrelu(y1p, 0.01)
    op = y1 @ p["w2"] + p["b2"]
    return np.tanh(op), (y1p, y1)


# --------------------------------------------------------------------------------------
# AIQN actor
# --------------------------------------------------------------------------------------
def _flat2(x):
    x = np.asarray(x)
    return x.reshape(x.shape[0], -1)


def aiqn_sample(p, states, taus, dtype=np.float32, trace=False):
    """AutoRegressiveStochasticActor.__call__ sampling mode, networks.py:194-225.
    Returns actions (rows, A); with trace=True also the per-step (prev_action, h_prev, h)
    used for stage-wise kernel parity."""
    p = cast(p, dtype)
    states = np.asarray(states, dtype=dtype)
    taus = _flat2(taus)
    rows, A = taus.shape
    se, _ = state_embed(p, states)
    ne = cosine_basis_linear(taus, p["wn"], p["bn"], dtype)            # (rows, A, 400)
    a = np.zeros((rows,), dtype=dtype)                                   # networks.py:203
    h = np.zeros((rows, E), dtype=dtype)                                 # hidden_state=None
    acts, tr = [], []
    for d in range(A):                                                   # networks.py:208
        ae = cosine_basis_linear(a.reshape(rows, 1), p["wa"], p["ba"], dtype, 0.2)[:, 0, :]
        x = np.concatenate([se, ae], axis=1)
        hn, _ = gru_step(p, x, h)
        out, _ = head(p, hn * ne[:, d, :])
        if trace:
            tr.append((a.copy(), h.copy(), hn.copy()))
        a, h = out[:, 0], hn
        acts.append(a)
    actions = np.stack(acts, axis=1)
    return (actions, tr) if trace else actions


def aiqn_supervised(p, states, taus, actions, dtype=np.float32):
    """_supervised_forward, networks.py:227-268 (teacher forced).  Returns (pred, cache)."""
    p = cast(p, dtype)
    states = np.asarray(states, dtype=dtype)
    taus, actions = _flat2(taus), _flat2(actions)
    rows, A = taus.shape
    se, se_pre = state_embed(p, states)
    shifted = np.zeros((rows, A), dtype=np.float32)                      # networks.py:253-255
    shifted[:, 1:] = actions[:, :-1]
    ca = cosine_basis(shifted, dtype).reshape(rows, A, NB)
    ae_pre = ca @ p["wa"] + p["ba"]
    ae = lrelu(ae_pre, 0.2)
    cn = cosine_basis(taus, dtype).reshape(rows, A, NB)
    ne = cn @ p["wn"] + p["bn"]
    h = np.zeros((rows, E), dtype=dtype)
    steps, preds = [], []
    for d in range(A):
        x = np.concatenate([se, ae[:, d]], axis=1)
        hn, (z, r, c, mhh) = gru_step(p, x, h)
        u = hn * ne[:, d]
        out, (y1p, y1) = head(p, u)
        steps.append(dict(hprev=h, h=hn, z=z, r=r, c=c, mhh=mhh, u=u, y1p=y1p, y1=y1, out=out[:, 0]))
        preds.append(out[:, 0])
        h = hn
    cache = dict(p=p, states=states, se=se, se_pre=se_pre, ca=ca, ae_pre=ae_pre, ae=ae, cn=cn, ne=ne,
                 steps=steps, A=A, rows=rows)
    return np.stack(preds, axis=1), cache


def aiqn_supervised_backward(cache, dact):
    """Hand-derived BPTT for aiqn_supervised; dact = dL/d(pred) (rows, A).  The teacher
    actions and taus carry no gradient (networks.py:134-140: only trainable_variables)."""
    p, A, rows = cache["p"], cache["A"], cache["rows"]
    dt = dact.dtype
    g = {k: np.zeros_like(v) for k, v in p.items()}
    dh_next = np.zeros((rows, E), dtype=dt)
    dsx = np.zeros((rows, 3 * E), dtype=dt)
    kxs, kxa = p["kx"][:E], p["kx"][E:]
    for d in reversed(range(A)):
        st = cache["steps"][d]
        dop = (dact[:, d] * (1 - st["out"] ** 2))[:, None]
        g["w2"] += st["y1"].T @ dop
        g["b2"] += dop.sum(0)
        dy1p = (dop @ p["w2"].T) * lrelu_grad(st["y1p"], 0.01)
        g["w1"] += st["u"].T @ dy1p
        g["b1"] += dy1p.sum(0)
        du = dy1p @ p["w1"].T
        dne = du * st["h"]
        g["wn"] += cache["cn"][:, d].T @ dne
        g["bn"] += dne.sum(0)
        dh = dh_next + du * cache["ne"][:, d]
        z, r, c = st["z"], st["r"], st["c"]
        dz = dh * (st["hprev"] - c)
        dc = dh * (1 - z)
        dcp = dc * (1 - c * c)
        dzp = dz * z * (1 - z)
        drp = (dcp * st["mhh"]) * r * (1 - r)
        dmx = np.concatenate([dzp, drp, dcp], axis=1)
        dmh = np.concatenate([dzp, drp, dcp * r], axis=1)
        g["u"] += st["hprev"].T @ dmh
        g["gb"][1] += dmh.sum(0)
        g["gb"][0] += dmx.sum(0)
        g["kx"][E:] += cache["ae"][:, d].T @ dmx
        dsx += dmx
        dae_pre = (dmx @ kxa.T) * lrelu_grad(cache["ae_pre"][:, d], 0.2)
        g["wa"] += cache["ca"][:, d].T @ dae_pre
        g["ba"] += dae_pre.sum(0)
        dh_next = dh * z + dmh @ p["u"].T
    g["kx"][:E] += cache["se"].T @ dsx
    dse = dsx @ kxs.T
    pre = cache["se_pre"]
    dse_pre = dse * np.where(pre > 0, np.asarray(1, dt), np.asarray(0.01 * 0.2, dt))
    # slope of lrelu_0.2(lrelu_0.01(x)) on x<=0 is 0.2*0.01 (App. A.2)
    g["ws"] += cache["states"].T @ dse_pre
    g["bs"] += dse_pre.sum(0)
    return g


# --------------------------------------------------------------------------------------
# IQN actor (StochasticActor, networks.py:271-323) -- the reference CLI's default actor (main.py:78)
#   wm (D,200) bm (200)   merge_embedding_layer      D = (400 // A) * A
#   wn (64,d)  bn (d)     noise_embedding_layer.act_linear (LeakyReLU 0.01), d = 400 // A
#   wo (200,A) bo (A)     output_action_layer (tanh)
#   ws (S,D)   bs (D)     state_embedding_layer (LeakyReLU 0.01)
# --------------------------------------------------------------------------------------
IQN_KEYS = ("wm", "bm", "wn", "bn", "wo", "bo", "ws", "bs")


def iqn_shapes(S, A):
    d = E // A
    D = d * A
    return {"wm": (D, 200), "bm": (200,), "wn": (NB, d), "bn": (d,), "wo": (200, A), "bo": (A,), "ws": (S, D), "bs": (D,)}


def init_iqn(rng, S, A, bias_scale=0.0):
    p = {}
    for k, shp in iqn_shapes(S, A).items():
        p[k] = _glorot(rng, shp) if len(shp) == 2 else (bias_scale * rng.standard_normal(shp)).astype(np.float32)
    return p


def iqn_forward(p, states, taus, dtype=np.float32, cache=False):
    """StochasticActor.__call__, networks.py:304-323 (supervise_actions is ignored there)."""
    p = cast(p, dtype)
    states = np.asarray(states, dtype=dtype)
    taus = _flat2(taus)
    rows, A = taus.shape
    d = p["wn"].shape[1]
    se_pre = states @ p["ws"] + p["bs"]
    se = lrelu(se_pre, 0.01)                                              # :313
    cn = cosine_basis(taus, dtype)                                        # (rows*A, 64)
    ne_pre = cn @ p["wn"] + p["bn"]
    ne = lrelu(ne_pre, 0.01).reshape(rows, A * d)                         # :315-317
    merge = se * ne                                                       # :319
    me_pre = merge @ p["wm"] + p["bm"]
    me = lrelu(me_pre, 0.01)                                              # :321
    out = np.tanh(me @ p["wo"] + p["bo"])                                 # :322
    if cache:
        return out, dict(p=p, states=states, se=se, se_pre=se_pre, cn=cn, ne=ne, ne_pre=ne_pre.reshape(rows, A * d), merge=merge,
                         me=me, me_pre=me_pre, out=out, rows=rows, A=A, d=d)
    return out


def iqn_backward(c, dact):
    p, rows, A, d = c["p"], c["rows"], c["A"], c["d"]
    g = {}
    dop = dact * (1 - c["out"] ** 2)
    g["wo"] = c["me"].T @ dop
    g["bo"] = dop.sum(0)
    dme = (dop @ p["wo"].T) * lrelu_grad(c["me_pre"], 0.01)
    g["wm"] = c["merge"].T @ dme
    g["bm"] = dme.sum(0)
    dmerge = dme @ p["wm"].T
    dne = (dmerge * c["se"] * lrelu_grad(c["ne_pre"], 0.01)).reshape(rows * A, d)
    g["wn"] = c["cn"].T @ dne
    g["bn"] = dne.sum(0)
    dse = dmerge * c["ne"] * lrelu_grad(c["se_pre"], 0.01)
    g["ws"] = c["states"].T @ dse
    g["bs"] = dse.sum(0)
    return g


def iqn_train_grads(p, states, actions, advantage, taus, mode="linear", beta=1.0, dtype=np.float32):
    """IQNActor.train with the IQN forward (networks.py:122-141)."""
    adv = np.asarray(advantage, dtype).reshape(-1, 1)
    w = target_density(mode, adv, beta)
    pred, cache = iqn_forward(p, states, taus, dtype, cache=True)
    loss, dpred = quantile_huber_loss(pred, np.asarray(_flat2(actions), dtype), np.asarray(_flat2(taus), dtype), w)
    return iqn_backward(cache, dpred), float(loss), pred, dpred


# --------------------------------------------------------------------------------------
# FNN / Critic / Value
# --------------------------------------------------------------------------------------
def fnn_forward(p, x, cache=False):
    """FNN.__call__, networks.py:350-358: in->400->300->1, LeakyReLU(0.01) hidden, linear out."""
    a0 = x @ p["w0"] + p["b0"]
    h0 = lrelu(a0, 0.01)
    a1 = h0 @ p["w1"] + p["b1"]
    h1 = lrelu(a1, 0.01)
    out = h1 @ p["w2"] + p["b2"]
    return (out, (x, a0, h0, a1, h1)) if cache else out


def fnn_backward(p, cache, dout):
    x, a0, h0, a1, h1 = cache
    g = {}
    g["w2"] = h1.T @ dout
    g["b2"] = dout.sum(0)
    da1 = (dout @ p["w2"].T) * lrelu_grad(a1, 0.01)
    g["w1"] = h0.T @ da1
    g["b1"] = da1.sum(0)
    da0 = (da1 @ p["w1"].T) * lrelu_grad(a0, 0.01)
    g["w0"] = x.T @ da0
    g["b0"] = da0.sum(0)
    return g


def critic_forward(p1, p2, states, actions, dtype=np.float32):
    """Critic.__call__, networks.py:384-388: min of two FNNs on concat(s, a)."""
    x = np.concatenate([np.asarray(states, dtype), np.asarray(actions, dtype)], axis=-1)
    return np.minimum(fnn_forward(cast(p1, dtype), x), fnn_forward(cast(p2, dtype), x))


def value_forward(p, states, dtype=np.float32):
    return fnn_forward(cast(p, dtype), np.asarray(states, dtype))


def critic_train_grads(p1, p2, pv_target, states, actions, rewards, next_states, terminal_mask,
                       gamma, q_normalization, unif_noise, dtype=np.float32):
    """Critic.train, networks.py:410-426.  unif_noise ~ U[0,1) (B,A) replaces :411.
    mse over the size-1 last axis gives a (B,) vector whose tape.gradient is the gradient of
    the SUM => d/dpred = 2 (pred - y)."""
    p1, p2, pv_target = cast(p1, dtype), cast(p2, dtype), cast(pv_target, dtype)
    f = lambda a: np.asarray(a, dtype=dtype)
    noise = (f(unif_noise) * 2 - 1) * dtype(q_normalization)
    ab = np.clip(f(actions) + noise, -1, 1)
    yq = f(rewards) + dtype(gamma) * fnn_forward(pv_target, f(next_states)) * (1 - f(terminal_mask))
    x = np.concatenate([f(states), ab], axis=-1)
    out = {}
    for name, p in (("fnn1", p1), ("fnn2", p2)):
        pred, cache = fnn_forward(p, x, cache=True)
        out[name] = fnn_backward(p, cache, 2 * (pred - yq))
        out[name + "_loss"] = float(((pred - yq) ** 2).sum())
    out["yq"] = yq
    return out


def tile_state_major(states, K):
    """Value.train tiling, networks.py:464-467: row = b*K + k."""
    return np.repeat(states, K, axis=0)


def tile_sample_major(states, K):
    """_sample_positive_advantage_actions tiling, agent.py:186: row = k*B + b."""
    return np.tile(states, (K, 1))


def value_train_grads(pv, states, v_critic, dtype=np.float32):
    """Value.train, networks.py:493-496 given v_critic (B,1) = mean_K min-critic (:483-485)."""
    pv = cast(pv, dtype)
    pred, cache = fnn_forward(pv, np.asarray(states, dtype), cache=True)
    yv = np.asarray(v_critic, dtype)
    g = fnn_backward(pv, cache, 2 * (pred - yv))
    g["_loss"] = float(((pred - yv) ** 2).sum())
    return g


def v_critic_from_q(q, B, K):
    """networks.py:484-485: reshape (B,K,1), mean over K of the per-sample min."""
    return q.reshape(B, K, 1).mean(axis=1)


# --------------------------------------------------------------------------------------
# advantage filter, actor loss
# --------------------------------------------------------------------------------------
def positive_advantage_filter(q, v):
    """agent.py:207-215: tf.where(q > v) -> ascending int64 indices (M,1); strict >, NaN false."""
    q, v = np.asarray(q).reshape(-1), np.asarray(v).reshape(-1)
    with np.errstate(invalid="ignore"):
        mask = q > v
    idx = np.nonzero(mask)[0].astype(np.int64)
    return mask, idx.reshape(-1, 1)


def target_density(mode, advantage, beta):
    """IQNActor.target_density, networks.py:79-95.  boltzmann: softmax over the size-1 last
    axis of an (M,1) tensor => all ones."""
    if mode == "linear":
        return advantage / advantage.sum()
    if mode == "boltzmann":
        return np.ones_like(advantage)
    raise NotImplementedError


def quantile_huber_loss(pred, target, taus, weights):
    """IQNActor.huber_quantile_loss, networks.py:113-120 (Huber elementwise, TF 2.0/2.1; App. A.6).
    Returns (loss, dL/dpred)."""
    u = pred - target
    ind = (u > 0).astype(pred.dtype)
    au = np.abs(u)
    q = np.minimum(au, 1)
    hub = 0.5 * q * q + (au - q)
    coef = np.abs(taus - ind) * weights
    n = pred.size
    loss = (coef * hub).sum() / n
    return loss, coef * np.clip(u, -1, 1) / n


def actor_train_grads(p, states, actions, advantage, taus, mode="linear", beta=1.0, dtype=np.float32):
    """IQNActor.train, networks.py:122-141 (taus injected for :134)."""
    adv = np.asarray(advantage, dtype).reshape(-1, 1)
    w = target_density(mode, adv, beta)
    pred, cache = aiqn_supervised(p, states, taus, actions, dtype)
    loss, dpred = quantile_huber_loss(pred, np.asarray(_flat2(actions), dtype), np.asarray(_flat2(taus), dtype), w)
    g = aiqn_supervised_backward(cache, dpred)
    return g, float(loss), pred, dpred


# --------------------------------------------------------------------------------------
# optimiser / target update
# --------------------------------------------------------------------------------------
def adam_tf(p, g, m, v, t, lr=1e-4, b1=0.9, b2=0.999, eps=1e-7):
    """Keras OptimizerV2 Adam, App. A.8 (epsilon outside the bias correction).  fp32 in place
    semantics; t is the 1-based step.  Returns new (p, m, v)."""
    f = np.float32
    alpha = f(lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t))
    m = m + (g - m) * f(1 - b1)
    v = v + (g * g - v) * f(1 - b2)
    p = p - alpha * m / (np.sqrt(v) + f(eps))
    return p.astype(f), m.astype(f), v.astype(f)


def polyak(t, s, rho):
    """helpers.update, helpers.py:86: t*(1-rho) + s*rho with (1-rho) formed in Python float
    then cast to fp32; each op rounded to fp32 (App. A.9)."""
    f = np.float32
    return (t.astype(f) * f(1.0 - rho) + s.astype(f) * f(rho)).astype(f)


# --------------------------------------------------------------------------------------
# the whole train step (agent.py:121-168) with injected randomness
# --------------------------------------------------------------------------------------
def make_agent(rng, S, A, bias_scale=0.0):
    """Six independently initialised nets + zero Adam slots."""
    nets = dict(actor=init_actor(rng, S, A, bias_scale), fnn1=init_fnn(rng, S + A, bias_scale),
                fnn2=init_fnn(rng, S + A, bias_scale), value=init_fnn(rng, S, bias_scale),
                target_actor=init_actor(rng, S, A, bias_scale), target_fnn1=init_fnn(rng, S + A, bias_scale),
                target_fnn2=init_fnn(rng, S + A, bias_scale), target_value=init_fnn(rng, S, bias_scale))
    for k in ("actor", "fnn1", "fnn2", "value"):
        nets["m_" + k] = {n: np.zeros_like(x) for n, x in nets[k].items()}
        nets["v_" + k] = {n: np.zeros_like(x) for n, x in nets[k].items()}
        nets["t_" + k] = 0
    return nets


def make_batch(rng, S, A, B):
    """Synthetic replay batch, SURVEY.md §8d."""
    f = np.float32
    return dict(s=rng.standard_normal((B, S)).astype(f), a=rng.uniform(-1, 1, (B, A)).astype(f),
                r=rng.standard_normal((B, 1)).astype(f), sp=rng.standard_normal((B, S)).astype(f),
                it=(rng.uniform(size=(B, 1)) < 0.05).astype(f))


def make_noise(rng, A, B, K):
    """The injected RNG buffers for one step (SURVEY.md §8a 'RNG sites')."""
    f = np.float32
    P = B * K
    return dict(critic_u=rng.uniform(size=(B, A)).astype(f),       # networks.py:411
                value_tau=rng.uniform(size=(P, A)).astype(f),      # networks.py:473
                filter_tau=rng.uniform(size=(P, A)).astype(f),     # helpers.py:28
                filter_normal=rng.standard_normal((P, A)).astype(f),   # agent.py:189
                filter_rand=rng.uniform(-1, 1, (P, A)).astype(f),  # agent.py:194
                actor_tau=rng.uniform(size=(2 * P, A)).astype(f))  # networks.py:134 (row j <- row j)


def sample_positive_advantage_actions(nets, states, noise, K, dtype=np.float32):
    """agent.py:170-216.  Returns dict with q, v, mask, idx and gathered rows."""
    f = lambda a: np.asarray(a, dtype)
    tiled = tile_sample_major(f(states), K)
    ta = aiqn_sample(nets["target_actor"], tiled, noise["filter_tau"], dtype)
    ta = np.clip(ta + f(noise["filter_normal"]) * dtype(0.01), -1, 1)
    tq = critic_forward(nets["target_fnn1"], nets["target_fnn2"], tiled, ta, dtype)
    ra = f(noise["filter_rand"])
    rq = critic_forward(nets["target_fnn1"], nets["target_fnn2"], tiled, ra, dtype)
    acts = np.concatenate([ta, ra], 0)
    q = np.concatenate([tq, rq], 0)
    v = value_forward(nets["target_value"], tiled, dtype)
    v = np.concatenate([v, v], 0)
    tiled2 = np.concatenate([tiled, tiled], 0)
    mask, idx = positive_advantage_filter(q, v)
    i = idx[:, 0]
    return dict(q=q, v=v, actions=acts, mask=mask, idx=idx, good_states=tiled2[i], good_actions=acts[i],
                advantages=(q - v)[i])


def train_one_step(nets, batch, noise, K, gamma=0.99, rho=5e-3, q_normalization=0.01, mode="linear",
                   beta=1.0, lr=1e-4, dtype=np.float32):
    """GACAgent.train_one_step, agent.py:121-168.  Mutates and returns ``nets``; also
    returns a dict of intermediates (losses, M, q/v, indices, gradients)."""
    B = batch["s"].shape[0]
    A = batch["a"].shape[1]
    info = {}
    # critics.train (agent.py:140-149)
    cg = critic_train_grads(nets["fnn1"], nets["fnn2"], nets["target_value"], batch["s"], batch["a"], batch["r"],
                            batch["sp"], batch["it"], gamma, q_normalization, noise["critic_u"], dtype)
    info["grad_fnn1"], info["grad_fnn2"] = cg["fnn1"], cg["fnn2"]
    info["loss_fnn1"], info["loss_fnn2"] = cg["fnn1_loss"], cg["fnn2_loss"]
    # value.train (agent.py:150-155)
    tiled = tile_state_major(np.asarray(batch["s"], dtype), K)
    va = aiqn_sample(nets["target_actor"], tiled, noise["value_tau"], dtype)
    vq = critic_forward(nets["target_fnn1"], nets["target_fnn2"], tiled, va, dtype)
    v_critic = v_critic_from_q(vq, B, K)
    vg = value_train_grads(nets["value"], batch["s"], v_critic, dtype)
    info["loss_value"] = vg.pop("_loss")
    info["grad_value"], info["v_critic"], info["value_actions"] = vg, v_critic, va
    # advantage filter (agent.py:157)
    flt = sample_positive_advantage_actions(nets, batch["s"], noise, K, dtype)
    info["filter"] = flt
    M = flt["idx"].shape[0]
    info["M"] = M
    # actor.train (agent.py:158-165)
    if M:
        ag, loss, pred, dpred = actor_train_grads(nets["actor"], flt["good_states"], flt["good_actions"],
                                                  flt["advantages"], noise["actor_tau"][:M], mode, beta, dtype)
        info["grad_actor"], info["loss_actor"] = ag, loss
    # optimisers (networks.py:141,421,426,496)
    for name, g in (("fnn1", cg["fnn1"]), ("fnn2", cg["fnn2"]), ("value", vg)) + ((("actor", info["grad_actor"]),) if M else ()):
        nets["t_" + name] += 1
        for k in nets[name]:
            nets[name][k], nets["m_" + name][k], nets["v_" + name][k] = adam_tf(
                nets[name][k], np.asarray(g[k], np.float32), nets["m_" + name][k], nets["v_" + name][k],
                nets["t_" + name], lr)
    # Polyak (agent.py:166-168)
    for name in ("actor", "fnn1", "fnn2", "value"):
        for k in nets[name]:
            nets["target_" + name][k] = polyak(nets["target_" + name][k], nets[name][k], rho)
    return nets, info


# --------------------------------------------------------------------------------------
# running observation / return statistics (agent.py:72-80, 263-268)
# --------------------------------------------------------------------------------------
def running_mean_std_merge(mean, var, count, batch_mean, batch_var, batch_count):
    """utils/utils.py:25-38, as written: the cross term uses sqrt(delta) where the parallel-variance formula has
    delta**2 (NaN for a negative delta).  fp32 state, Python-float count."""
    f = np.float32
    delta = np.asarray(batch_mean, f) - np.asarray(mean, f)
    tot = count + batch_count
    new_mean = np.asarray(mean, f) + delta * f(batch_count) / f(tot)
    with np.errstate(invalid="ignore"):
        m2 = np.asarray(var, f) * f(count) + np.asarray(batch_var, f) * f(batch_count) + np.sqrt(delta) * f(count) * f(batch_count) / f(tot)
    return new_mean.astype(f), (m2 / f(tot)).astype(f), tot


def running_mean_std_update(mean, var, count, x):
    """RunningMeanStd.update, utils/utils.py:48-52: population moments over axis 0, batch_count = x.shape[0]."""
    x = np.asarray(x)
    return running_mean_std_merge(mean, var, count, np.mean(x, axis=0).astype(np.float32), np.square(np.std(x, axis=0)).astype(np.float32), x.shape[0])
