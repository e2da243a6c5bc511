"""RunningMeanStd / normalize / denormalize (utils/utils.py:25-56, helpers.py:89-101) against the oracle's restatement."""
import numpy as np
import torch

from dpo_replication_b200 import RunningMeanStd, denormalize, normalize
from oracle import gac_oracle as O


def test_running_mean_std_follows_the_reference_merge():
    rng = np.random.default_rng(0)
    rms = RunningMeanStd(shape=4)
    mean, var, count = np.zeros(4, np.float32), np.ones(4, np.float32), 1e-4
    assert rms.count == 1e-4 and np.array_equal(rms.var, var)
    for i in range(5):
        x = (rng.standard_normal((7, 4)) + 3.0 * (i + 1)).astype(np.float32)     # increasing means: delta > 0, no NaN
        rms.update(x)
        mean, var, count = O.running_mean_std_update(mean, var, count, x)
        assert np.array_equal(rms.mean, mean) and np.array_equal(rms.var, var) and rms.count == count
    assert np.isfinite(rms.var).all()
    # a batch mean below the running mean: sqrt(delta) is NaN in the reference, reproduced by default ...
    x = (rng.standard_normal((7, 4)) - 50.0).astype(np.float32)
    rms.update(x)
    mean, var, count = O.running_mean_std_update(mean, var, count, x)
    assert np.isnan(rms.var).all() and np.isnan(var).all() and np.array_equal(rms.mean, mean)
    # ... and the textbook parallel-variance merge behind the switch agrees with the pooled moments
    fix = RunningMeanStd(shape=4, bug_compatible=False, epsilon=0.0)
    a, b = rng.standard_normal((9, 4)).astype(np.float32), (rng.standard_normal((5, 4)) - 2).astype(np.float32)
    fix.mean, fix.var, fix.count = a.mean(0), a.var(0), 9
    fix.update(b)
    both = np.concatenate([a, b])
    assert np.allclose(fix.mean, both.mean(0), atol=1e-6) and np.allclose(fix.var, both.var(0), atol=1e-5)


def test_single_state_update_as_store_transition_calls_it():
    # agent.py:264: obs_rms.update(state) with a rank-1 state: moments over axis 0 are SCALARS, batch_count = state_dim
    rms = RunningMeanStd(shape=3)
    s = np.array([1.0, 2.0, 6.0], np.float32)
    rms.update(s)
    m, v, c = O.running_mean_std_update(np.zeros(3, np.float32), np.ones(3, np.float32), 1e-4, s)
    assert np.array_equal(rms.mean, m) and np.array_equal(rms.var, v) and rms.count == c == 3 + 1e-4
    assert rms.mean.shape == (3,) and np.allclose(rms.mean, 3.0, atol=1e-3)


def test_normalize_denormalize():
    rms = RunningMeanStd(shape=3)
    rms.mean, rms.var = np.array([1, 2, 3], np.float32), np.array([4, 9, 16], np.float32)
    x = torch.tensor([[3.0, 5.0, 7.0]])
    y = normalize(x, rms)
    assert torch.allclose(y, torch.tensor([[1.0, 1.0, 1.0]]))
    assert torch.allclose(denormalize(y, rms), x)
    assert normalize(x, None) is x and denormalize(x, None) is x
