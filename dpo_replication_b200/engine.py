"""Host-side glue between torch tensors (device memory, streams) and the C ABI.

``Arena``  : the flat fp32 parameter arenas (online / target / Adam m / Adam v / grad) laid out
             by ``gac_layout`` with named views per variable.
``Engine`` : one ``gac_ctx`` (workspace partition) for a (S, A, B, K) configuration and thin
             typed wrappers over each entry point of include/gac_b200.h.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import Config, HParams, StepIO, check, lib

MODES = {"linear": 0, "boltzmann": 1}


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("dpo_replication_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def f32(x, device) -> torch.Tensor:
    """fp32, contiguous, on `device` (copies only when needed)."""
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    return x.to(device=device, dtype=torch.float32).contiguous()


class Arena:
    """Flat parameter storage for the four trainable nets + their targets and Adam slots."""

    def __init__(self, S: int, A: int, device, actor_kind: int = 0):
        self.S, self.A, self.device, self.actor_kind = S, A, device, actor_kind
        self.names = _lib.var_names(actor_kind)
        self.lay = _lib.layout(S, A, actor_kind)
        self.n = int(self.lay.offsets[_lib.NVARS])
        z = lambda extra=0: torch.zeros(self.n + extra, dtype=torch.float32, device=device)
        self.online, self.target, self.adam_m, self.adam_v = z(), z(), z(), z()
        self.grad = z(_lib.STEP_NSTATS)          # gradient arena + stats tail (all-reduced together)
        self.adam_t = torch.zeros(_lib.NNETS, dtype=torch.int32, device=device)

    # ---- addressing -------------------------------------------------------------------
    def net_range(self, net: str):
        i = _lib.NET_NAMES.index(net)
        return int(self.lay.net_begin[i]), int(self.lay.net_begin[i + 1])

    def net_slice(self, buf: torch.Tensor, net: str) -> torch.Tensor:
        a, b = self.net_range(net)
        return buf[a:b]

    def view(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        v = self.names.index(name)
        o, r, c = int(self.lay.offsets[v]), int(self.lay.rows[v]), int(self.lay.cols[v])
        t = buf[o:o + r * c]
        return t.view(r, c) if c > 1 else t

    def named(self, buf: torch.Tensor, net: str) -> Dict[str, torch.Tensor]:
        return {n.split(".", 1)[1]: self.view(buf, n) for n in self.names if n.startswith(net + ".") and not n.split(".", 1)[1].startswith("_")}

    @property
    def stats(self) -> torch.Tensor:
        return self.grad[self.n:]

    # ---- init / IO --------------------------------------------------------------------
    def load(self, buf: torch.Tensor, net: str, params: Dict[str, "torch.Tensor | object"]):
        for k, v in params.items():
            dst = self.view(buf, f"{net}.{k}")
            dst.copy_(f32(v, self.device).reshape(dst.shape))

    def dump(self, buf: torch.Tensor, net: str):
        return {k: v.detach().cpu().numpy().copy() for k, v in self.named(buf, net).items()}


def keras_default_init(S: int, A: int, net: str, generator=None) -> Dict[str, torch.Tensor]:
    """Keras defaults: Dense glorot-uniform / zero bias; GRU kernel glorot-uniform, recurrent
    kernel orthogonal, zero bias (reference GAC/networks.py:159-174, 345-346)."""
    def glorot(r, c):
        lim = math.sqrt(6.0 / (r + c))
        return (torch.rand(r, c, generator=generator) * 2 - 1) * lim

    def orth(n):
        q, r = torch.linalg.qr(torch.randn(n, n, generator=generator))
        return q * torch.sign(torch.diagonal(r))

    E = 400
    if net == "iqn":          # StochasticActor, reference GAC/networks.py:283-302
        d = E // A
        D = d * A
        return {"wm": glorot(D, 200), "bm": torch.zeros(200), "wn": glorot(64, d), "bn": torch.zeros(d), "wo": glorot(200, A),
                "bo": torch.zeros(A), "ws": glorot(S, D), "bs": torch.zeros(D)}
    if net == "actor":
        return {"wa": glorot(64, E), "ba": torch.zeros(E), "w1": glorot(E, E), "b1": torch.zeros(E), "w2": glorot(E, 1),
                "b2": torch.zeros(1), "wn": glorot(64, E), "bn": torch.zeros(E), "kx": glorot(2 * E, 3 * E),
                "u": torch.cat([orth(E) for _ in range(3)], dim=1), "gb": torch.zeros(2, 3 * E), "ws": glorot(S, E),
                "bs": torch.zeros(E)}
    n_in = S if net == "value" else S + A
    return {"w0": glorot(n_in, 400), "b0": torch.zeros(400), "w1": glorot(400, 300), "b1": torch.zeros(300),
            "w2": glorot(300, 1), "b2": torch.zeros(1)}


def _on_stream(fn):
    """Entry-point wrapper: make the engine's device current and bind the context to torch's CURRENT stream
    (so a caller inside `torch.cuda.stream(...)` / a graph capture gets stream-ordered work)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        with torch.cuda.device(self.device):
            self.bind_current_stream()
            return fn(self, *a, **k)
    return wrapped


class Engine:
    """One gac_ctx; every call runs on torch's current stream of the engine's device."""

    def __init__(self, S: int, A: int, B: int, K: int, device=None, chunk_rows: int = 0, engine: int = 0, actor_kind: int = 0):
        _require_cuda()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.S, self.A, self.B, self.K = S, A, B, K
        self.actor_kind = actor_kind
        self.cfg = Config(S, A, B, K, chunk_rows, engine, actor_kind)
        nbytes = lib.gac_workspace_bytes(C.byref(self.cfg))
        if nbytes == 0:
            raise _lib.GacError(f"bad configuration: {lib.gac_last_error().decode()}")
        self.workspace_bytes = int(nbytes)
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = (self.workspace.data_ptr() + 255) // 256 * 256
            self.stream = torch.cuda.current_stream(self.device)
            h = C.c_void_p()
            check(lib.gac_create(C.byref(self.cfg), C.c_void_p(base), nbytes, C.c_void_p(self.stream.cuda_stream), C.byref(h)))
        self.h = h
        self.lay = _lib.layout(S, A, actor_kind)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib.gac_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- helpers ----------------------------------------------------------------------
    def _new(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def launch_count(self) -> int:
        return int(lib.gac_launch_count(self.h))

    def bind_current_stream(self):
        """Re-bind the context when torch's current stream changed since the last call; the new stream first waits
        for everything queued on the old one (the workspace is shared)."""
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != self.stream.cuda_stream:
            if not torch.cuda.is_current_stream_capturing():
                cur.wait_stream(self.stream)
            check(lib.gac_set_stream(self.h, C.c_void_p(cur.cuda_stream)))
            self.stream = cur

    def hparams(self, gamma=0.99, rho=5e-3, q_normalization=0.01, beta=1.0, mode="linear", lr=1e-4, beta1=0.9,
                beta2=0.999, eps=1e-7) -> HParams:
        if mode not in MODES:
            raise NotImplementedError  # reference GAC/networks.py:94-95
        return HParams(gamma, rho, q_normalization, beta, MODES[mode], lr, beta1, beta2, eps)

    # ---- entry points -----------------------------------------------------------------
    @_on_stream
    def polyak(self, target: torch.Tensor, source: torch.Tensor, rho: float):
        assert target.numel() == source.numel()
        check(lib.gac_polyak(ptr(target), ptr(source), target.numel(), rho, C.c_void_p(self.stream.cuda_stream)))

    @_on_stream
    def adam_polyak(self, online, target, m, v, grad, hp: HParams, adam_t, grad_scale=None, skip=None):
        check(lib.gac_adam_polyak(self.h, ptr(online), ptr(target), ptr(m), ptr(v), ptr(grad), C.byref(hp), ptr(adam_t),
                                  ptr(grad_scale), ptr(skip)))

    @_on_stream
    def aiqn_sample(self, actor, states_u, state_idx, taus, out=None):
        rows = taus.shape[0]
        if out is None:
            out = self._new(rows, self.A)
        elif out.shape != (rows, self.A) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != taus.device:
            raise ValueError("aiqn_sample: out must be a contiguous float32 (rows, A) tensor on the engine's device")
        check(lib.gac_aiqn_sample(self.h, ptr(actor), ptr(states_u), states_u.shape[0], ptr(state_idx), ptr(taus), rows, ptr(out)))
        return out

    @_on_stream
    def gru_step(self, actor, states_u, state_idx, action_emb, h_prev=None, reuse=False, actor_inputs=False, same_inputs=False):
        """One GRU(400) step of the AIQN actor (reference GAC/networks.py:166): h' = GRU([se[state_idx] | action_emb], h_prev).
        actor_inputs: action_emb / h_prev come from the actor itself (bounded by its weights) -> 3xFP16 variant allowed."""
        rows = action_emb.shape[0]
        out = self._new(rows, 400)
        check(lib.gac_gru_step(self.h, ptr(actor), ptr(states_u), states_u.shape[0], ptr(state_idx), ptr(action_emb), ptr(h_prev), rows,
                               ptr(out), int(reuse) | (2 if actor_inputs else 0) | (4 if same_inputs else 0)))
        return out

    @_on_stream
    def iqn_forward(self, actor, states_u, state_idx, taus):
        rows = taus.shape[0]
        out = self._new(rows, self.A)
        check(lib.gac_iqn_forward(self.h, ptr(actor), ptr(states_u), states_u.shape[0], ptr(state_idx), ptr(taus), rows, ptr(out)))
        return out

    @_on_stream
    def iqn_train_fwd(self, actor, states_u, state_idx, taus, d_rows=None):
        rows = taus.shape[0]
        pred = self._new(rows, self.A)
        check(lib.gac_iqn_train_fwd(self.h, ptr(actor), ptr(states_u), states_u.shape[0], ptr(state_idx), ptr(taus), rows, ptr(d_rows),
                                    ptr(pred)))
        return pred

    @_on_stream
    def iqn_bwd(self, actor, dpred, grad_actor, accumulate=False):
        check(lib.gac_iqn_bwd(self.h, ptr(actor), ptr(dpred), ptr(grad_actor), int(accumulate)))

    @_on_stream
    def critic_eval(self, fnn1, fnn2, states_u, state_idx, actions):
        rows = actions.shape[0]
        q = self._new(rows, 1)
        check(lib.gac_critic_eval(self.h, ptr(fnn1), ptr(fnn2), ptr(states_u), ptr(state_idx), ptr(actions), rows, ptr(q)))
        return q

    @_on_stream
    def value_eval(self, value, states):
        rows = states.shape[0]
        v = self._new(rows, 1)
        check(lib.gac_value_eval(self.h, ptr(value), ptr(states), rows, ptr(v)))
        return v

    @_on_stream
    def advantage_compact(self, q, v, actions=None, state_idx=None):
        n = q.numel()
        idx = self._new(n, dtype=torch.int64)
        pos = self._new(n, dtype=torch.int32)
        m = self._new(1, dtype=torch.int32)
        sadv = self._new(1)
        adv = self._new(n)
        A = actions.shape[1] if actions is not None else 0
        act_out = self._new(n, A) if actions is not None else None
        sidx_out = self._new(n, dtype=torch.int32) if state_idx is not None else None
        check(lib.gac_advantage_compact(self.h, ptr(q), ptr(v), n, ptr(idx), ptr(pos), ptr(m), ptr(sadv), ptr(adv), ptr(actions), A,
                                        ptr(act_out), ptr(state_idx), ptr(sidx_out)))
        return dict(idx=idx, pos=pos, m=m, sum_adv=sadv, adv=adv, actions=act_out, state_idx=sidx_out)

    @_on_stream
    def aiqn_supervised_fwd(self, actor, states_u, state_idx, taus, teacher, d_rows=None):
        rows = taus.shape[0]
        pred = self._new(rows, self.A)
        check(lib.gac_aiqn_supervised_fwd(self.h, ptr(actor), ptr(states_u), states_u.shape[0], ptr(state_idx), ptr(taus), ptr(teacher),
                                          rows, ptr(d_rows), ptr(pred)))
        return pred

    @_on_stream
    def qhuber_loss_bwd(self, pred, target, taus, weights, inv_norm: float, d_rows=None):
        rows, A = pred.shape
        dpred = torch.zeros_like(pred)
        loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        check(lib.gac_qhuber_loss_bwd(self.h, ptr(pred), ptr(target), ptr(taus), ptr(weights), rows, ptr(d_rows), A, inv_norm, ptr(dpred),
                                      ptr(loss)))
        return loss, dpred

    @_on_stream
    def aiqn_bwd(self, actor, dpred, grad_actor, accumulate=False):
        check(lib.gac_aiqn_bwd(self.h, ptr(actor), ptr(dpred), ptr(grad_actor), int(accumulate)))

    @_on_stream
    def td_fwd_bwd(self, fnn, n_in, x, y, grad_fnn):
        loss = self._new(1)
        check(lib.gac_td_fwd_bwd(self.h, ptr(fnn), n_in, ptr(x), ptr(y), x.shape[0], ptr(grad_fnn), ptr(loss)))
        return loss

    @_on_stream
    def gemm(self, a, b, trans_a=False, trans_b=False, engine=0):
        M = a.shape[1] if trans_a else a.shape[0]
        K = a.shape[0] if trans_a else a.shape[1]
        N = b.shape[0] if trans_b else b.shape[1]
        c = self._new(M, N)
        check(lib.gac_gemm(self.h, engine, ptr(a), ptr(b), ptr(c), M, N, K, int(trans_a), int(trans_b)))
        return c

    def step_io(self, arena: Arena, batch: Dict[str, torch.Tensor], noise: Dict[str, torch.Tensor], taps=None) -> StepIO:
        io = StepIO()
        for k in ("online", "target", "adam_m", "adam_v", "grad", "adam_t"):
            setattr(io, k, getattr(arena, k).data_ptr())
        io.s, io.a, io.r, io.sp, io.done = (batch[k].data_ptr() for k in ("s", "a", "r", "sp", "it"))
        for k in ("critic_u", "value_tau", "filter_tau", "filter_normal", "filter_rand", "actor_tau"):
            setattr(io, k, noise[k].data_ptr())
        for k, t in (taps or {}).items():
            setattr(io, k, t.data_ptr())
        return io

    @_on_stream
    def step_grads(self, io: StepIO, hp: HParams):
        check(lib.gac_step_grads(self.h, C.byref(io), C.byref(hp)))

    @_on_stream
    def step_apply(self, io: StepIO, hp: HParams):
        check(lib.gac_step_apply(self.h, C.byref(io), C.byref(hp)))
