"""On-disk formats either side of training.

``save_model`` / ``load_model`` mirror ``utils/utils.py:7-22`` of the reference: same names, same argument, same
directory layout (``{basedir}/actor``).  The reference writes a TensorFlow SavedModel (``tf.saved_model.save``); TensorFlow
is not part of this stack, so the directory holds a plain ``variables.npz`` (one array per variable, in the
``tf.Module`` variable order of ``include/gac_b200.h``) and a ``meta.json``.  The reference LOADS from
``{basedir}/ddpg_actor`` (``utils.py:18``), a path its own ``save_model`` never writes; ``load_model`` looks there first and
then at ``{basedir}/actor``.

``save_checkpoint`` / ``load_checkpoint`` store the whole training state of a ``GACAgent`` -- the four online nets, the four
targets, the Adam moments and step counters -- so that a resumed run continues bit-identically given the same replay
batches and random draws (the reference only ever saves the actor).
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

CHECKPOINT_VERSION = 1


class RunningMeanStd(object):
    """utils/utils.py:41-56 (OpenAI-baselines running mean / variance, parallel-variance merge), host side like the
    reference's: it is fed one transition per environment step (agent.py:263-268), far from the train step.

    The reference's merge (utils/utils.py:25-38) adds ``sqrt(delta) * count * batch_count / tot_count`` where the
    parallel-variance formula has ``delta ** 2``: NaN as soon as a batch mean falls below the running mean.
    ``bug_compatible=True`` (default) reproduces that line as written; ``False`` uses the square."""

    def __init__(self, epsilon=1e-4, shape=(), bug_compatible=True):
        self.mean = np.zeros(shape, np.float32)
        self.var = np.ones(shape, np.float32)
        self.count = epsilon
        self.bug_compatible = bug_compatible

    @property
    def std(self):                     # read by helpers.denormalize (helpers.py:98-101); the reference never defines it
        return np.sqrt(self.var)

    def update(self, x):
        x = np.asarray(x.detach().cpu() if isinstance(x, torch.Tensor) else x)
        batch_mean = np.mean(x, axis=0).astype(np.float32)             # tf.reduce_mean(x, axis=0) cast to float32
        batch_var = np.square(np.std(x, axis=0)).astype(np.float32)    # tf.square(tf.math.reduce_std(x, axis=0))
        self.update_from_moments(batch_mean, batch_var, x.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        f = np.float32
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * f(batch_count) / f(tot_count)
        m_a = self.var * f(self.count)
        m_b = batch_var * f(batch_count)
        with np.errstate(invalid="ignore"):
            cross = np.sqrt(delta) if self.bug_compatible else np.square(delta)
        M2 = m_a + m_b + cross * f(self.count) * f(batch_count) / f(tot_count)
        self.mean, self.var, self.count = new_mean.astype(f), (M2 / f(tot_count)).astype(f), tot_count


def _actor_kind_name(actor) -> str:
    return "IQN" if getattr(actor, "actor_kind", 0) == 1 else "AIQN"


def save_model(actor, basedir=None):
    """utils.py:7-15 -- export the actor's variables to ``{basedir}/actor``."""
    if not os.path.exists('models/'):          # utils.py:8-9 (kept: callers rely on the side effect)
        os.makedirs('models/')
    actor_path = "{}/actor".format(basedir)
    os.makedirs(actor_path, exist_ok=True)
    params = actor.arena.dump(actor._buf, "actor")
    np.savez(os.path.join(actor_path, "variables.npz"), **params)
    meta = {"format": "dpo_replication_b200.actor", "version": CHECKPOINT_VERSION, "actor": _actor_kind_name(actor),
            "state_dim": int(actor.arena.S), "action_dim": int(actor.arena.A),
            "variables": {k: list(v.shape) for k, v in params.items()}}
    with open(os.path.join(actor_path, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)


def load_model(basedir=None, device=None):
    """utils.py:17-22 -- returns an actor object (callable like the trained one)."""
    from .networks import AutoRegressiveStochasticActor, StochasticActor
    for name in ("ddpg_actor", "actor"):       # the reference's load path first (utils.py:18), then its save path
        actor_path = "{}/{}".format(basedir, name)
        if os.path.exists(os.path.join(actor_path, "meta.json")):
            break
    else:
        raise FileNotFoundError("no actor export under {}/ddpg_actor or {}/actor".format(basedir, basedir))
    print('Loading model from {}'.format(actor_path))
    with open(os.path.join(actor_path, "meta.json")) as f:
        meta = json.load(f)
    if meta.get("format") != "dpo_replication_b200.actor":
        raise ValueError("{} is not an actor export of this package".format(actor_path))
    cls = StochasticActor if meta["actor"] == "IQN" else AutoRegressiveStochasticActor
    actor = cls(meta["state_dim"], meta["action_dim"], device=device, init=False)
    with np.load(os.path.join(actor_path, "variables.npz")) as z:
        params = {k: z[k] for k in z.files}
    for k, shp in meta["variables"].items():
        if k not in params or list(params[k].shape) != list(shp):
            raise ValueError("variable {} missing or mis-shaped in {}".format(k, actor_path))
    actor.load_weights({"actor": params})
    return actor


def save_checkpoint(agent, path):
    """Full training state of a GACAgent in one ``.npz`` (atomic: written to a temp file, then renamed)."""
    st = agent.dump_state()
    flat = {}
    for key, val in st.items():
        if isinstance(val, dict):
            for k, v in val.items():
                flat[key + "/" + k] = v
        else:
            flat["scalar/" + key] = np.asarray(val, dtype=np.int64)
    meta = {"format": "dpo_replication_b200.checkpoint", "version": CHECKPOINT_VERSION, "actor": "IQN" if agent.actor_kind == 1 else "AIQN",
            "state_dim": int(agent.state_dim), "action_dim": int(agent.action_dim), "batch_size": int(agent.batch_size),
            "action_samples": int(agent.action_samples), "mode": agent.mode, "beta": float(agent.beta), "tau": float(agent.tau),
            "gamma": float(agent.gamma), "q_normalization": float(agent.q_normalization)}
    flat["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    tmp = path + ".tmp.npz"
    np.savez(tmp, **flat)
    os.replace(tmp, path)


def load_checkpoint(agent, path):
    """Restore a ``save_checkpoint`` file into an agent of the same architecture (checked)."""
    with np.load(path) as z:
        meta = json.loads(bytes(z["meta"]).decode())
        if meta.get("format") != "dpo_replication_b200.checkpoint":
            raise ValueError("{} is not a checkpoint of this package".format(path))
        want = {"actor": "IQN" if agent.actor_kind == 1 else "AIQN", "state_dim": int(agent.state_dim), "action_dim": int(agent.action_dim)}
        for k, v in want.items():
            if meta[k] != v:
                raise ValueError("checkpoint {} = {!r}, agent has {!r}".format(k, meta[k], v))
        st = {}
        for name in z.files:
            if name == "meta":
                continue
            key, k = name.split("/", 1)
            if key == "scalar":
                st[k] = int(z[name])
            else:
                st.setdefault(key, {})[k] = z[name]
    agent.load_state(st)
    torch.cuda.synchronize(agent.device)
    return meta
