"""Synthetic inputs of the benchmark (SURVEY.md §8d "Synthetic inputs"), NumPy only.

There is no network for datasets or checkpoints, so the bench and the examples run on replay batches of the
named environment's dimensions and on Keras-default random weights: Dense glorot-uniform / zero bias, GRU
kernel glorot-uniform over (800, 1200), recurrent kernel orthogonal per gate, zero biases (what
tf.keras.layers.Dense / GRU give the reference at GAC/networks.py:159-174, 345-346).  Both arms of
``bench.py`` (this package on the GPU, the reference-equivalent path on the CPU) load the SAME state and
batches from here, so they time the same workload.
"""
from __future__ import annotations

import numpy as np

E, NB, H1, H2 = 400, 64, 400, 300


def _glorot(rng, r, c):
    lim = np.sqrt(6.0 / (r + c))
    return rng.uniform(-lim, lim, (r, c)).astype(np.float32)


def _orthogonal(rng, n):
    q, r = np.linalg.qr(rng.standard_normal((n, n)))
    return (q * np.sign(np.diag(r))).astype(np.float32)


def keras_actor(rng, S, A):
    z = lambda *s: np.zeros(s, np.float32)
    return {"wa": _glorot(rng, NB, E), "ba": z(E), "w1": _glorot(rng, E, E), "b1": z(E), "w2": _glorot(rng, E, 1), "b2": z(1),
            "wn": _glorot(rng, NB, E), "bn": z(E), "kx": _glorot(rng, 2 * E, 3 * E),
            "u": np.concatenate([_orthogonal(rng, E) for _ in range(3)], axis=1), "gb": z(2, 3 * E), "ws": _glorot(rng, S, E), "bs": z(E)}


def keras_fnn(rng, n_in):
    z = lambda *s: np.zeros(s, np.float32)
    return {"w0": _glorot(rng, n_in, H1), "b0": z(H1), "w1": _glorot(rng, H1, H2), "b1": z(H2), "w2": _glorot(rng, H2, 1), "b2": z(1)}


def agent_state(rng, S, A):
    """A freshly constructed GACAgent (agent.py:65-116): four online nets, target actor / critics hard-copied,
    target value independently initialised (F7), zero Adam slots, step counters 0.  Layout = GACAgent.load_state."""
    nets = {"actor": keras_actor(rng, S, A), "fnn1": keras_fnn(rng, S + A), "fnn2": keras_fnn(rng, S + A), "value": keras_fnn(rng, S)}
    for k in ("actor", "fnn1", "fnn2"):
        nets["target_" + k] = {n: x.copy() for n, x in nets[k].items()}
    nets["target_value"] = keras_fnn(rng, S)
    for k in ("actor", "fnn1", "fnn2", "value"):
        nets["m_" + k] = {n: np.zeros_like(x) for n, x in nets[k].items()}
        nets["v_" + k] = {n: np.zeros_like(x) for n, x in nets[k].items()}
        nets["t_" + k] = 0
    return nets


def replay_batch(rng, S, A, B):
    f = np.float32
    return dict(s=rng.standard_normal((B, S)).astype(f), a=rng.uniform(-1, 1, (B, A)).astype(f),
                r=rng.standard_normal((B, 1)).astype(f), sp=rng.standard_normal((B, S)).astype(f),
                it=(rng.uniform(size=(B, 1)) < 0.05).astype(f))


def step_noise(rng, A, B, K):
    """One step's draws at the reference's seven tf.random sites, in the layouts gac_step_io documents."""
    f = np.float32
    P = B * K
    return dict(critic_u=rng.uniform(size=(B, A)).astype(f), value_tau=rng.uniform(size=(P, A)).astype(f),
                filter_tau=rng.uniform(size=(P, A)).astype(f), filter_normal=rng.standard_normal((P, A)).astype(f),
                filter_rand=rng.uniform(-1, 1, (P, A)).astype(f), actor_tau=rng.uniform(size=(2 * P, A)).astype(f))
