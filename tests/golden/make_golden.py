"""Generates tests/golden/gac_step_small.npz from the fp64 oracle (TensorFlow, hence the
reference itself, cannot run in this image -- see DESIGN.md "parity unpinned").  The fixture
pins the oracle against accidental edits and is what the GPU box compares with when
/root/reference and the oracle's author are not around.

    python tests/golden/make_golden.py
"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gac_oracle as O  # noqa: E402

S, A, B, K, SEED = 3, 2, 4, 4, 20260101


def probe(x, n=96):
    """Small, position-stable summary of a tensor: [sum, sum|x|] + n evenly spaced samples."""
    x = np.asarray(x, np.float64).reshape(-1)
    pick = x[np.linspace(0, x.size - 1, min(n, x.size)).astype(np.int64)]
    return np.concatenate([[x.sum(), np.abs(x).sum()], pick])


def flat(prefix, d, out, full=False):
    for k, v in d.items():
        out[f"{prefix}.{k}"] = np.asarray(v) if full else probe(v)


def inputs():
    """All inputs are regenerated from the seed (the weights are ~16 MB); their probes are
    stored so that a change of NumPy's bit generator would be noticed."""
    rng = np.random.default_rng(SEED)
    nets = O.make_agent(rng, S, A, bias_scale=0.05)
    batch, noise = O.make_batch(rng, S, A, B), O.make_noise(rng, A, B, K)
    st = rng.standard_normal((8, S)).astype(np.float32)
    tau = rng.uniform(size=(8, A)).astype(np.float32)
    tch = rng.uniform(-1, 1, (8, A)).astype(np.float32)
    return nets, batch, noise, st, tau, tch


def main():
    nets, batch, noise, st, tau, tch = inputs()
    out = {"dims": np.array([S, A, B, K, SEED])}
    for name in ("actor", "fnn1", "fnn2", "value", "target_actor", "target_fnn1", "target_fnn2", "target_value"):
        flat("in." + name, nets[name], out)
    flat("batch", batch, out, full=True)
    flat("noise", noise, out, full=True)
    after, info = O.train_one_step(copy.deepcopy(nets), batch, noise, K, dtype=np.float64)
    for name in ("actor", "fnn1", "fnn2", "value", "target_actor", "target_fnn1", "target_fnn2", "target_value"):
        flat("out." + name, after[name], out)
    for name in ("actor", "fnn1", "fnn2", "value"):
        flat("grad." + name, info["grad_" + name], out)
    out["M"] = np.array(info["M"])
    out["idx"] = info["filter"]["idx"]
    out["q"], out["v"] = info["filter"]["q"], info["filter"]["v"]
    out["losses"] = np.array([info["loss_fnn1"], info["loss_fnn2"], info["loss_value"], info["loss_actor"]])
    # stage vectors: sampler (free-running), supervised forward, critic, value
    out["stage.states"], out["stage.taus"], out["stage.teacher"] = st, tau, tch
    out["stage.sample"] = O.aiqn_sample(nets["actor"], st, tau, np.float64)
    out["stage.supervised"] = O.aiqn_supervised(nets["actor"], st, tau, tch, np.float64)[0]
    out["stage.q"] = O.critic_forward(nets["fnn1"], nets["fnn2"], st, tch, np.float64)
    out["stage.v"] = O.value_forward(nets["value"], st, np.float64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gac_step_small.npz"), **out)
    print("wrote gac_step_small.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
