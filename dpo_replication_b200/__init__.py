"""dpo_replication_b200 -- the GAC train step of gwbcho/dpo-replication on B200 (sm_100a).

Host-side mirror of the reference's `GAC/networks.py` / `GAC/agent.py` / `GAC/helpers.py`
class surface over hand-written CUDA kernels behind a C ABI (include/gac_b200.h).
Importing the package loads libgac_b200.so and fails loudly if it has not been built.
"""
from . import _lib  # noqa: F401  (loads the shared library or raises)
from .agent import GACAgent
from .engine import Arena, Engine
from .helpers import ActionSampler, DeviceReplayBuffer, ReplayBuffer, denormalize, normalize, update
from .utils import RunningMeanStd, load_checkpoint, load_model, save_checkpoint, save_model
from .networks import (FNN, AutoRegressiveStochasticActor, CosineBasisLinear, Critic, IQNActor, StochasticActor, Value)

__all__ = ["GACAgent", "Arena", "Engine", "ActionSampler", "ReplayBuffer", "DeviceReplayBuffer", "update", "normalize", "denormalize", "FNN",
           "AutoRegressiveStochasticActor", "CosineBasisLinear", "Critic", "IQNActor", "StochasticActor", "Value", "save_model", "load_model",
           "save_checkpoint", "load_checkpoint", "RunningMeanStd"]
