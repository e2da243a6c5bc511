"""GACAgent (reference GAC/agent.py) on the B200 engine: same constructor signature and
defaults, same public methods, `train_one_step()` mutates in place and returns None."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import dist as gdist
from .engine import Arena, Engine, f32, keras_default_init
from .helpers import ActionSampler, DeviceReplayBuffer, ReplayBuffer, normalize, update
from .networks import AutoRegressiveStochasticActor, Critic, StochasticActor, Value
from .utils import RunningMeanStd


class GACAgent:
    """GAC agent; actions live in [-1, 1]^A (agent.py:18-119)."""

    def __init__(self, action_dim, state_dim, buffer_size=1000000, action_samples=10, mode='linear', beta=1, tau=5e-3,
                 q_normalization=0.01, gamma=0.99, normalize_obs=False, normalize_rewards=False, batch_size=64,
                 actor='AIQN', *args, device=None, chunk_rows=0, engine=0, replay_on_device=True, **kwargs):
        self.action_dim, self.state_dim, self.buffer_size = action_dim, state_dim, buffer_size
        self.gamma, self.action_samples, self.mode, self.beta, self.tau = gamma, action_samples, mode, beta, tau
        self.batch_size = batch_size
        self.normalize_observations, self.q_normalization, self.normalize_rewards = normalize_obs, q_normalization, normalize_rewards
        if actor not in ('AIQN', 'IQN'):
            raise NotImplementedError("actor must be 'AIQN' or 'IQN' (agent.py:65-70)")
        self.actor_kind = 1 if actor == 'IQN' else 0
        # agent.py:72-80: running statistics of observations / discounted returns (utils/utils.py:41-56)
        self.obs_rms = RunningMeanStd(shape=state_dim) if normalize_obs else None
        self.ret_rms = RunningMeanStd(shape=1) if normalize_rewards else None
        self.ret = 0
        if not torch.cuda.is_available():
            raise RuntimeError("dpo_replication_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        S, A = state_dim, action_dim
        self.arena = Arena(S, A, self.device, self.actor_kind)
        self.engine = Engine(S, A, batch_size, action_samples, self.device, chunk_rows=chunk_rows, engine=engine, actor_kind=self.actor_kind)
        mk = dict(arena=self.arena, init=False)
        actor_cls = StochasticActor if self.actor_kind == 1 else AutoRegressiveStochasticActor      # agent.py:65-70
        self.actor = actor_cls(S, A, role="online", **mk)
        self.target_actor = actor_cls(S, A, role="target", **mk)
        self.critics = Critic(S, A, role="online", **mk)
        self.target_critics = Critic(S, A, role="target", **mk)
        self.value = Value(S, role="online", **mk)
        self.target_value = Value(S, role="target", **mk)
        for net in ("actor", "fnn1", "fnn2", "value"):
            self.arena.load(self.arena.online, net, keras_default_init(S, A, "iqn" if (net == "actor" and self.actor_kind == 1) else net))
        # agent.py:114-116: hard copies.  target_value has no variables yet when update() runs in
        # the reference, so it keeps its own independent initialisation.
        update(self.target_actor, self.actor, 1.0)
        update(self.target_critics, self.critics, 1.0)
        self.arena.load(self.arena.target, "value", keras_default_init(S, A, "value"))
        # the reference keeps a host NumPy ring (helpers.py:33-72); the default here keeps it in HBM
        self.replay = DeviceReplayBuffer(S, A, buffer_size, self.device) if replay_on_device else ReplayBuffer(S, A, buffer_size)
        self.action_sampler = ActionSampler(self.actor.action_dim)
        self._hp = self.engine.hparams(gamma, tau, q_normalization, beta, mode)
        gdist.broadcast_state(self.arena)       # data-parallel replicas must start from identical parameters

    # ---- weights in / out (parity harness) ------------------------------------------------
    def load_state(self, nets: Dict[str, Dict[str, np.ndarray]]):
        """Load {actor,fnn1,fnn2,value,target_*,m_*,v_*,t_*} as produced by the oracle."""
        ar = self.arena
        for name in ("actor", "fnn1", "fnn2", "value"):
            ar.load(ar.online, name, nets[name])
            ar.load(ar.target, name, nets["target_" + name])
            if "m_" + name in nets:
                ar.load(ar.adam_m, name, nets["m_" + name])
                ar.load(ar.adam_v, name, nets["v_" + name])
        ts = [int(nets.get("t_" + n, 0)) for n in ("actor", "fnn1", "fnn2", "value")]
        ar.adam_t.copy_(torch.tensor(ts, dtype=torch.int32))
        gdist.broadcast_state(ar)

    def dump_state(self):
        ar = self.arena
        out = {}
        for name in ("actor", "fnn1", "fnn2", "value"):
            out[name] = ar.dump(ar.online, name)
            out["target_" + name] = ar.dump(ar.target, name)
            out["m_" + name] = ar.dump(ar.adam_m, name)
            out["v_" + name] = ar.dump(ar.adam_v, name)
        t = ar.adam_t.cpu().tolist()
        for i, name in enumerate(("actor", "fnn1", "fnn2", "value")):
            out["t_" + name] = t[i]
        return out

    # ---- the hot path ---------------------------------------------------------------------
    def draw_noise(self) -> Dict[str, torch.Tensor]:
        """The reference's seven tf.random sites for one step, drawn on device."""
        B, K, A, dev = self.batch_size, self.action_samples, self.action_dim, self.device
        P = B * K
        return dict(critic_u=torch.rand(B, A, device=dev), value_tau=torch.rand(P, A, device=dev),
                    filter_tau=torch.rand(P, A, device=dev), filter_normal=torch.randn(P, A, device=dev),
                    filter_rand=torch.rand(P, A, device=dev) * 2 - 1, actor_tau=torch.rand(2 * P, A, device=dev))

    def train_on_batch(self, batch: Dict, noise: Optional[Dict] = None, taps: Optional[Dict[str, torch.Tensor]] = None):
        """One train step on an explicit replay batch {s,a,r,sp,it} (+ injected randomness).
        Everything runs on the engine's stream with no host synchronisation; with an initialised
        torch.distributed group the gradient arena is all-reduced (SUM) between the two phases."""
        b = {k: f32(batch[k], self.device) for k in ("s", "a", "r", "sp", "it")}
        n = self.draw_noise() if noise is None else {k: f32(v, self.device) for k, v in noise.items()}
        io = self.engine.step_io(self.arena, b, n, taps)
        self._keep = (b, n, io)                 # keep the inputs alive until the stream is done with them
        self.engine.step_grads(io, self._hp)
        gdist.allreduce_grads(self.arena.grad)
        self.engine.step_apply(io, self._hp)

    def train_on_batch_graphed(self, batch: Dict, noise: Optional[Dict] = None):
        """`train_on_batch` replayed from CUDA graphs.  The step is ~290 kernel launches on two streams with no host
        synchronisation and nothing data-dependent on the host (the number of selected rows stays on the device), so it
        captures as is: inputs are copied into static buffers, the reference's seven tf.random draws are made in place
        outside the graph, and one cudaGraphLaunch replaces the launches -- what matters for the small configs (Pendulum: the
        eager step is launch bound) and for the host thread's time.  With a process group the step is TWO graphs,
        `gac_step_grads` and `gac_step_apply`, with the NCCL all-reduce of the gradient arena between them."""
        multi = gdist.world()[1] > 1
        dev = self.device
        if not hasattr(self, "_step_graph"):
            b0 = {k: torch.zeros_like(f32(batch[k], dev)) for k in ("s", "a", "r", "sp", "it")}
            n0 = self.draw_noise()
            io = self.engine.step_io(self.arena, b0, n0)
            cap = torch.cuda.Stream(device=dev)
            cap.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(cap):      # warm up on the capture stream: lazy module loading and one-time attribute calls
                state = [t.clone() for t in (self.arena.online, self.arena.target, self.arena.adam_m, self.arena.adam_v, self.arena.adam_t)]
                self.engine.step_grads(io, self._hp)
                self.engine.step_apply(io, self._hp)
                for t, sv in zip((self.arena.online, self.arena.target, self.arena.adam_m, self.arena.adam_v, self.arena.adam_t), state):
                    t.copy_(sv)               # the warm-up step must not count as a train step
            cap.synchronize()
            graph, graph2 = torch.cuda.CUDAGraph(), None
            with torch.cuda.graph(graph, stream=cap):
                self.engine.step_grads(io, self._hp)
                if not multi:
                    self.engine.step_apply(io, self._hp)
            if multi:
                graph2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph2, stream=cap):
                    self.engine.step_apply(io, self._hp)
            self._step_graph = (graph, graph2, b0, n0, io)
        graph, graph2, b0, n0, _ = self._step_graph
        for k, t in b0.items():
            t.copy_(f32(batch[k], dev).reshape(t.shape), non_blocking=True)
        if noise is None:                     # the reference's tf.random sites, drawn in place
            for k in ("critic_u", "value_tau", "filter_tau", "actor_tau"):
                n0[k].uniform_()
            n0["filter_normal"].normal_()
            n0["filter_rand"].uniform_(-1, 1)
        else:
            for k, t in n0.items():
                t.copy_(f32(noise[k], dev).reshape(t.shape), non_blocking=True)
        graph.replay()
        if graph2 is not None:
            gdist.allreduce_grads(self.arena.grad)
            graph2.replay()

    def train_one_step(self):
        """agent.py:121-168."""
        t = self.replay.sample_batch(self.batch_size, device=self.device)
        batch = dict(s=normalize(t.s, self.obs_rms), a=t.a, r=normalize(t.r, self.ret_rms), sp=normalize(t.sp, self.obs_rms), it=t.it)
        self.train_on_batch(batch)

    def _sample_positive_advantage_actions(self, states, *, noise: Optional[Dict] = None):
        """agent.py:170-216 -> (good_states, good_actions, advantages).  This API-compatible
        form returns dynamically shaped tensors and therefore synchronises on M; the fused
        train step keeps M on the device."""
        s = f32(states, self.device)
        B, K, A = s.shape[0], self.action_samples, self.action_dim
        P = B * K
        eng = self.engine if B == self.batch_size else Engine(self.state_dim, A, B, K, self.device, actor_kind=self.actor_kind)
        n = noise or {}
        g = lambda k, f: f32(n[k], self.device) if k in n else f()
        sidx = torch.arange(B, dtype=torch.int32, device=self.device).repeat(K)            # row = k*B + b
        ar = self.arena
        sample = eng.iqn_forward if self.actor_kind == 1 else eng.aiqn_sample
        ta = sample(ar.net_slice(ar.target, "actor"), s, sidx, g("filter_tau", lambda: torch.rand(P, A, device=self.device)))
        ta = torch.clamp(ta + g("filter_normal", lambda: torch.randn(P, A, device=self.device)) * 0.01, -1, 1)
        ra = g("filter_rand", lambda: torch.rand(P, A, device=self.device) * 2 - 1)
        acts = torch.cat([ta, ra], 0).contiguous()
        sidx2 = torch.cat([sidx, sidx]).contiguous()
        q = eng.critic_eval(ar.net_slice(ar.target, "fnn1"), ar.net_slice(ar.target, "fnn2"), s, sidx2, acts)
        v = eng.value_eval(ar.net_slice(ar.target, "value"), s)[sidx2.long()].contiguous()
        r = eng.advantage_compact(q.reshape(-1), v.reshape(-1), acts, sidx2)
        M = int(r["m"].item())
        self._last_filter = dict(q=q, v=v, idx=r["idx"][:M].reshape(-1, 1), actions=acts)
        return s[r["state_idx"][:M].long()], r["actions"][:M], r["adv"][:M].reshape(-1, 1)

    def get_action(self, states):
        """agent.py:218-228."""
        return self.action_sampler.get_actions(self.actor, states)

    # ---- acting path: the batch-1 / small-batch sampler as one CUDA graph ----------
    def get_action_graphed(self, states, *, taus=None):
        """Same result as `get_action` for a fixed batch size, replayed from a captured CUDA graph: the
        A-step autoregressive sampler is ~7 small kernels per action dimension, i.e. launch bound at
        batch 1 (main.py:119-121,195-203 call it once per environment step).  The graph is captured on
        first use per batch size; inputs are copied into static buffers, tau is drawn outside the graph."""
        s = f32(states, self.device)
        if s.dim() == 1:
            s = s.reshape(1, -1)
        rows = s.shape[0]
        if not hasattr(self, "_act_graphs"):
            self._act_graphs = {}
        if rows not in self._act_graphs:
            st_in = torch.zeros(rows, self.state_dim, device=self.device)
            tau_in = torch.zeros(rows, self.action_dim, device=self.device)
            sidx = torch.arange(rows, dtype=torch.int32, device=self.device)
            base = self.arena.net_slice(self.arena.online, "actor")
            # the engine launches on the stream it was created on: create it on the capture stream, warm it up
            # there (lazy module loading, no sync or allocation may happen inside the capture), then capture
            cap_stream = torch.cuda.Stream(device=self.device)
            cap_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(cap_stream):
                cap_eng = Engine(self.state_dim, self.action_dim, max(rows, 128), 1, self.device, chunk_rows=128, actor_kind=self.actor_kind)
                cap_fn = cap_eng.iqn_forward if self.actor_kind == 1 else cap_eng.aiqn_sample
                cap_fn(base, st_in, sidx, tau_in)
            cap_stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=cap_stream):
                out = cap_fn(base, st_in, sidx, tau_in)
            self._act_graphs[rows] = (graph, st_in, tau_in, out, cap_eng, sidx)
        graph, st_in, tau_in, out, _, _ = self._act_graphs[rows]
        st_in.copy_(s)
        tau_in.copy_(torch.rand(rows, self.action_dim, device=self.device) if taus is None else f32(taus, self.device).reshape(rows, -1))
        graph.replay()
        return out.clone()

    def select_perturbed_action(self, state, action_noise=None):
        """agent.py:230-248."""
        state = normalize(f32(state, self.device), self.obs_rms)
        action = self.action_sampler.get_actions(self.actor, state)
        if action_noise is not None:
            action = action + f32(action_noise(), self.device)
        return torch.clamp(action, -1, 1)

    def store_transition(self, state, action, reward, next_state, is_done):
        """agent.py:250-268."""
        self.replay.store(state, action, reward, next_state, is_done)
        if self.normalize_observations:
            self.obs_rms.update(state)
        if self.normalize_rewards:
            self.ret = self.ret * self.gamma + reward
            self.ret_rms.update(np.array([self.ret]))
            if is_done:
                self.ret = 0
