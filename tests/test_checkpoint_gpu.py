"""actor export/import (utils/utils.py:7-22 of the reference) and the full-state checkpoint."""
import os

import numpy as np
import pytest
import torch

from oracle import gac_oracle as O
from test_step_gpu import setup

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gb():
    import __graft_entry__ as ge
    ge.build()
    import dpo_replication_b200 as g
    return g


def _arenas(agent):
    ar = agent.arena
    return [ar.online.clone(), ar.target.clone(), ar.adam_m.clone(), ar.adam_v.clone(), ar.adam_t.clone()]


def test_checkpoint_resume_is_bit_identical(gb, tmp_path):
    S, A, B, K = 17, 6, 32, 8
    nets, batch, noise = setup(S, A, B, K, 7)
    rng = np.random.default_rng(8)
    batch2, noise2 = O.make_batch(rng, S, A, B), O.make_noise(rng, A, B, K)
    a = gb.GACAgent(A, S, action_samples=K, batch_size=B, device="cuda")
    a.load_state(nets)
    a.train_on_batch(batch, noise)
    path = str(tmp_path / "ckpt.npz")
    gb.save_checkpoint(a, path)
    a.train_on_batch(batch2, noise2)
    torch.cuda.synchronize()
    want = _arenas(a)
    b = gb.GACAgent(A, S, action_samples=K, batch_size=B, device="cuda")     # fresh agent, different random init
    meta = gb.load_checkpoint(b, path)
    assert meta["state_dim"] == S and meta["action_dim"] == A and meta["actor"] == "AIQN"
    b.train_on_batch(batch2, noise2)
    torch.cuda.synchronize()
    for x, y in zip(want, _arenas(b)):
        assert torch.equal(x, y)
    with pytest.raises(ValueError):
        gb.load_checkpoint(gb.GACAgent(A, S + 1, action_samples=K, batch_size=B, device="cuda"), path)


@pytest.mark.parametrize("kind", ["AIQN", "IQN"])
def test_actor_export_import(gb, tmp_path, monkeypatch, kind):
    monkeypatch.chdir(tmp_path)                 # save_model creates ./models like the reference (utils.py:8-9)
    S, A = 11, 3
    agent = gb.GACAgent(A, S, action_samples=4, batch_size=16, device="cuda", actor=kind)
    base = str(tmp_path / "run0")
    gb.save_model(agent.actor, base)
    assert os.path.isdir("models") and os.path.isfile(os.path.join(base, "actor", "variables.npz"))
    actor = gb.load_model(base)
    assert type(actor) is type(agent.actor)
    for x, y in zip(agent.actor.trainable_variables, actor.trainable_variables):
        assert torch.equal(x, y)
    states = torch.randn(40, S, device="cuda")
    taus = torch.rand(40, A, device="cuda")
    assert torch.equal(agent.actor(states, taus), actor(states, taus))
    # the reference's load path (utils.py:18) wins when it exists
    os.rename(os.path.join(base, "actor"), os.path.join(base, "ddpg_actor"))
    actor2 = gb.load_model(base)
    assert torch.equal(actor(states, taus), actor2(states, taus))
    with pytest.raises(FileNotFoundError):
        gb.load_model(str(tmp_path / "nothing"))
