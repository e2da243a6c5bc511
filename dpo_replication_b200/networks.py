"""The reference's network classes (GAC/networks.py) on the B200 engine.

Same names, constructor/`__call__`/`.train` signatures, defaults, mutation-in-place and
``None``-returning semantics as the reference.  Tensors are torch CUDA fp32
tensors instead of tf tensors.  Every method body is a call into libgac_b200.so; the only
torch ops here are RNG draws (the reference's ``tf.random`` sites) and shape plumbing.

Differences that are deliberate and documented:
  * variables exist from construction (Keras builds them at first call, agent.py:84-111);
  * every ``.train`` accepts optional keyword-only injected randomness (``taus=``, ``noise=``)
    so that parity tests can feed the reference's RNG draws (SURVEY.md §8a "RNG sites").
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib
from .engine import Arena, Engine, f32, keras_default_init

_ENGINES: Dict[tuple, Engine] = {}


def _engine_for(S: int, A: int, rows: int, device, kind: int = 0) -> Engine:
    """Engines for stand-alone network objects, cached by capacity (power-of-two rows)."""
    cap = 128
    while cap < rows:
        cap *= 2
    key = (S, A, cap, str(device), kind)
    if key not in _ENGINES:
        _ENGINES[key] = Engine(S, A, B=cap, K=1, device=device, chunk_rows=cap, actor_kind=kind)
    return _ENGINES[key]


def _default_device():
    if not torch.cuda.is_available():
        raise RuntimeError("dpo_replication_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device(f"cuda:{torch.cuda.current_device()}")


class _Net:
    """A network = a named slice of an Arena (its own, or the owning agent's)."""
    net_names: tuple = ()

    actor_kind = 0

    def _attach(self, arena: Optional[Arena], role: str, S: int, A: int, device, init: bool):
        self._owns_arena = arena is None
        self.arena = arena if arena is not None else Arena(S, A, device or _default_device(), self.actor_kind)
        self.role = role            # "online" or "target": which arena buffer holds the weights
        self.device = self.arena.device
        if init:
            for net in self.net_names:
                self.arena.load(self._buf, net, keras_default_init(S, A, "iqn" if (net == "actor" and self.actor_kind == 1) else net))

    @property
    def _buf(self) -> torch.Tensor:
        return getattr(self.arena, self.role)

    def _base(self, net: str, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.arena.net_slice(self._buf if buf is None else buf, net)

    @property
    def trainable_variables(self) -> List[torch.Tensor]:
        """Views in arena order (matching order on source and target, helpers.py:83-85)."""
        out = []
        for net in self.net_names:
            out += list(self.arena.named(self._buf, net).values())
        return out

    def load_weights(self, params_by_net: Dict[str, Dict[str, object]]):
        for net, p in params_by_net.items():
            self.arena.load(self._buf, net, p)

    def _adam_only(self, eng: Engine, nets: tuple, grad_scale=None, skip_actor=None):
        """Adam on `nets` of this object's arena (rho = 0 keeps the target buffer untouched)."""
        skip = torch.tensor([0 if n in nets else 1 for n in _lib.NET_NAMES], dtype=torch.int32, device=self.device)
        if skip_actor is not None:
            skip[0] = torch.maximum(skip[0], skip_actor.to(torch.int32))
        hp = eng.hparams(rho=0.0)
        eng.adam_polyak(self.arena.online, self.arena.target, self.arena.adam_m, self.arena.adam_v, self.arena.grad, hp,
                        self.arena.adam_t, grad_scale, skip)


class CosineBasisLinear:
    """GAC/networks.py:17-62: phi(x)_j = sum_i cos(x * i*pi) W[i,j] + b[j]; x (batch, a) or
    (batch, a, 1) -> (batch, a, embed_dim).  embed_dim is fixed to 400 by the engine's GEMM
    shapes when used inside the actor; the stand-alone class supports any embed_dim via the
    generic GEMM entry."""

    def __init__(self, n_basis_functions, embed_dim, activation=None, device=None):
        self.n_basis_functions, self.embed_dim, self.activation = n_basis_functions, embed_dim, activation
        self.device = device or _default_device()
        lim = (6.0 / (n_basis_functions + embed_dim)) ** 0.5
        self.kernel = ((torch.rand(n_basis_functions, embed_dim) * 2 - 1) * lim).to(self.device)
        self.bias = torch.zeros(embed_dim, device=self.device)

    @property
    def trainable_variables(self):
        return [self.kernel, self.bias]

    def __call__(self, x):
        x = f32(x, self.device)
        batch = x.shape[0]
        flat = x.reshape(-1, 1)
        ipi = (torch.arange(1, self.n_basis_functions + 1, dtype=torch.float32) * torch.tensor(3.14159274101257324, dtype=torch.float32)).to(self.device)
        emb = torch.cos(flat * ipi)            # stand-alone path: shape utility only (the actor fuses this on device)
        eng = _engine_for(3, 1, flat.shape[0], self.device)
        out = eng.gemm(emb.contiguous(), self.kernel.contiguous(), engine=1) + self.bias
        if self.activation is not None:
            out = self.activation(out)
        return out.reshape(batch, -1, self.embed_dim)


class IQNActor(_Net):
    """GAC/networks.py:65-141."""
    net_names = ("actor",)

    def __init__(self, state_dim, action_dim):
        self.state_dim, self.action_dim = state_dim, action_dim
        self.module_type = "IQNActor"

    def target_density(self, mode, advantage, beta):
        """networks.py:79-95 (boltzmann is a softmax over a size-1 axis => ones, F6)."""
        if mode == "linear":
            return advantage / advantage.sum()
        elif mode == "boltzmann":
            return torch.softmax(advantage / beta, dim=-1)
        raise NotImplementedError

    def huber_quantile_loss(self, actions, target_actions, taus, weights):
        """networks.py:97-120 via gac_qhuber_loss_bwd; returns the scalar loss tensor."""
        a, t, u = (f32(x, self.device).reshape(actions.shape[0], -1) for x in (actions, target_actions, taus))
        eng = _engine_for(self.state_dim, self.action_dim, a.shape[0], self.device)
        w = f32(weights, self.device).reshape(-1)
        loss, _ = eng.qhuber_loss_bwd(a, t, u, w, 1.0 / a.numel())
        return loss[0]

    def train(self, states, supervise_actions, advantage, mode, beta, *, taus=None):
        """networks.py:122-141: tau ~ U, weights = target_density, grads of the quantile Huber
        loss over the 13 variables, Adam(1e-4).  Mutates in place, returns None."""
        if mode not in ("linear", "boltzmann"):
            raise NotImplementedError
        s = f32(states, self.device)
        acts = f32(supervise_actions, self.device).reshape(s.shape[0], -1)
        adv = f32(advantage, self.device).reshape(-1)
        M, A = acts.shape
        if taus is None:
            taus = torch.rand(M, A, device=self.device)                      # networks.py:134
        taus = f32(taus, self.device).reshape(M, A)
        eng = _engine_for(self.state_dim, self.action_dim, M, self.device, self.actor_kind)
        sidx = torch.arange(M, dtype=torch.int32, device=self.device)
        base = self._base("actor", self.arena.online)
        iqn = self.actor_kind == 1
        pred = eng.iqn_train_fwd(base, s, sidx, taus) if iqn else eng.aiqn_supervised_fwd(base, s, sidx, taus, acts)
        w = adv / adv.sum() if mode == "linear" else torch.ones_like(adv)
        self.last_loss, dpred = eng.qhuber_loss_bwd(pred, acts, taus, w, 1.0 / (M * A))
        (eng.iqn_bwd if iqn else eng.aiqn_bwd)(base, dpred, self.arena.net_slice(self.arena.grad, "actor"), accumulate=False)
        self._adam_only(eng, ("actor",))


class AutoRegressiveStochasticActor(IQNActor):
    """GAC/networks.py:144-268 (AIQN: cosine embeddings + GRU(400) over the action dimensions)."""

    def __init__(self, state_dim, action_dim, n_basis_functions=64, *, arena=None, role="online", device=None, init=True):
        super().__init__(state_dim, action_dim)
        if n_basis_functions != 64:
            raise NotImplementedError("the B200 kernels bake n_basis_functions = 64 (reference default)")
        self.module_type = "AutoRegressiveStochasticActor"
        self._attach(arena, role, state_dim, action_dim, device, init)

    def __call__(self, state, taus, supervise_actions=None):
        s = f32(state, self.device)
        if s.dim() == 1:
            s = s.reshape(1, -1)
        rows = s.shape[0]
        taus = f32(taus, self.device).reshape(rows, -1)     # (B,A) or (B,A,1): tests/unit/GAC/networks.py:51-55
        eng = _engine_for(self.state_dim, self.action_dim, rows, self.device)
        sidx = torch.arange(rows, dtype=torch.int32, device=self.device)
        if supervise_actions is not None:
            acts = f32(supervise_actions, self.device).reshape(rows, -1)
            return eng.aiqn_supervised_fwd(self._base("actor"), s, sidx, taus, acts)
        return eng.aiqn_sample(self._base("actor"), s, sidx, taus)


class StochasticActor(IQNActor):
    """GAC/networks.py:271-323 (IQN): state embedding (S -> d*A) * cosine noise embedding (per action
    dimension, d = 400 // A) -> Dense(200) -> Dense(A, tanh).  `supervise_actions` is accepted and ignored,
    exactly as in the reference (:309)."""
    actor_kind = 1

    def __init__(self, state_dim, action_dim, n_basis_functions=64, *, arena=None, role="online", device=None, init=True):
        super().__init__(state_dim, action_dim)
        if n_basis_functions != 64:
            raise NotImplementedError("the B200 kernels bake n_basis_functions = 64 (reference default)")
        self.module_type = "StochasticActor"
        self.noise_embed_dim = 400 // action_dim
        self._attach(arena, role, state_dim, action_dim, device, init)

    def __call__(self, states, taus, supervise_actions=None):
        s = f32(states, self.device)
        if s.dim() == 1:
            s = s.reshape(1, -1)
        rows = s.shape[0]
        taus = f32(taus, self.device).reshape(rows, -1)
        eng = _engine_for(self.state_dim, self.action_dim, rows, self.device, 1)
        sidx = torch.arange(rows, dtype=torch.int32, device=self.device)
        return eng.iqn_forward(self._base("actor"), s, sidx, taus)


class FNN(_Net):
    """GAC/networks.py:326-358: in -> 400 -> 300 -> 1, LeakyReLU(0.01) on the hidden layers.
    Inside Critic/Value this is a named view of the owner's arena (`critics.fnn1`, `value.fnn`);
    constructed stand-alone it is a state-value style net on arch[0] inputs."""

    def __init__(self, arch, *, arena=None, role="online", net="value", device=None, init=True):
        if list(arch[1:]) != [400, 300, 1]:
            raise NotImplementedError("the B200 kernels implement the reference's 400/300/1 FNN")
        self.arch, self.net_names = list(arch), (net,)
        if arena is None:
            self._attach(None, role, arch[0], 1, device, init)
        else:
            self._attach(arena, role, arena.S, arena.A, device, False)

    def __call__(self, x):
        x = f32(x, self.device)
        if x.shape[1] != self.arch[0]:
            raise ValueError(f"expected {self.arch[0]} input features, got {x.shape[1]}")
        eng = _engine_for(self.arena.S, self.arena.A, x.shape[0], self.device)
        net = self.net_names[0]
        if net == "value":
            return eng.value_eval(self._base("value"), x)
        # a single critic head: evaluate the twin with itself (min(f, f) = f)
        S = self.arena.S
        return eng.critic_eval(self._base(net), self._base(net), x[:, :S].contiguous(), None, x[:, S:].contiguous())


class Critic(_Net):
    """GAC/networks.py:361-426: twin FNN critic, min of the two on concat(s, a)."""
    net_names = ("fnn1", "fnn2")

    def __init__(self, state_dim, action_dim, *, arena=None, role="online", device=None, init=True):
        self.state_dim, self.action_dim = state_dim, action_dim
        self._attach(arena, role, state_dim, action_dim, device, init)
        self.fnn1 = FNN([state_dim + action_dim, 400, 300, 1], arena=self.arena, role=role, net="fnn1")
        self.fnn2 = FNN([state_dim + action_dim, 400, 300, 1], arena=self.arena, role=role, net="fnn2")

    def __call__(self, states, actions):
        s, a = f32(states, self.device), f32(actions, self.device)
        eng = _engine_for(self.state_dim, self.action_dim, s.shape[0], self.device)
        return eng.critic_eval(self._base("fnn1"), self._base("fnn2"), s, None, a)

    def train(self, states, actions, rewards, next_states, terminal_mask, value, gamma, q_normalization=0.01, *, noise=None):
        """networks.py:390-426.  `value` is the target Value object.  Returns None."""
        s, a = f32(states, self.device), f32(actions, self.device)
        B, A = a.shape
        if noise is None:
            noise = torch.rand(B, A, device=self.device)                          # networks.py:411
        nz = (f32(noise, self.device) * 2 - 1) * q_normalization
        ab = torch.clamp(a + nz, -1, 1)
        yq = f32(rewards, self.device).reshape(B, 1) + gamma * value(next_states) * (1 - f32(terminal_mask, self.device).reshape(B, 1))
        x = torch.cat([s, ab], dim=-1).contiguous()
        eng = _engine_for(self.state_dim, self.action_dim, B, self.device)
        self.last_loss = []
        for net in ("fnn1", "fnn2"):
            self.last_loss.append(eng.td_fwd_bwd(self._base(net, self.arena.online), self.state_dim + self.action_dim, x,
                                                 yq.reshape(-1).contiguous(), self.arena.net_slice(self.arena.grad, net)))
        self._adam_only(eng, ("fnn1", "fnn2"))


class Value(_Net):
    """GAC/networks.py:429-496."""
    net_names = ("value",)

    def __init__(self, state_dim, *, arena=None, role="online", action_dim=1, device=None, init=True):
        self.state_dim = state_dim
        self._attach(arena, role, state_dim, action_dim, device, init)
        self.fnn = FNN([state_dim, 400, 300, 1], arena=self.arena, role=role, net="value")

    def __call__(self, states):
        s = f32(states, self.device)
        eng = _engine_for(self.arena.S, self.arena.A, s.shape[0], self.device)
        return eng.value_eval(self._base("value"), s)

    def train(self, states, actor, critic, action_samples=8, *, taus=None):
        """networks.py:444-496: state-major tiling, K samples from `actor`, mean over K of the
        min-critic, TD regression of V(s).  Returns None."""
        s = f32(states, self.device)
        B, K = s.shape[0], action_samples
        A = actor.action_dim
        if taus is None:
            taus = torch.rand(B * K, A, device=self.device)                       # networks.py:473
        sidx = torch.arange(B, dtype=torch.int32, device=self.device).repeat_interleave(K)   # row = b*K + k
        eng = _engine_for(actor.state_dim, A, B * K, self.device, actor.actor_kind)
        sample = eng.iqn_forward if actor.actor_kind == 1 else eng.aiqn_sample
        acts = sample(actor._base("actor"), s, sidx, f32(taus, self.device).reshape(B * K, A))
        q = eng.critic_eval(critic._base("fnn1"), critic._base("fnn2"), s, sidx, acts)
        v_critic = q.reshape(B, K, 1).mean(dim=1)                                  # networks.py:484-485
        eng_v = _engine_for(self.arena.S, self.arena.A, B, self.device)
        self.last_loss = eng_v.td_fwd_bwd(self._base("value", self.arena.online), self.state_dim, s, v_critic.reshape(-1).contiguous(),
                                          self.arena.net_slice(self.arena.grad, "value"))
        self._adam_only(eng_v, ("value",))
