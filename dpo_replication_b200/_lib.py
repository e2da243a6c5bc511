"""ctypes binding of libgac_b200.so (include/gac_b200.h).

There is NO fallback: if the shared library is missing the import fails loudly, and every
compute call needs a CUDA device.  Build with ``python __graft_entry__.py`` (or
``python -m dpo_replication_b200.build``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgac_b200.so")

NVARS = 31
NNETS = 4
STEP_NSTATS = 8

VAR_NAMES = (
    "actor.wa", "actor.ba", "actor.w1", "actor.b1", "actor.w2", "actor.b2", "actor.wn", "actor.bn",
    "actor.kx", "actor.u", "actor.gb", "actor.ws", "actor.bs",
    "fnn1.w0", "fnn1.b0", "fnn1.w1", "fnn1.b1", "fnn1.w2", "fnn1.b2",
    "fnn2.w0", "fnn2.b0", "fnn2.w1", "fnn2.b1", "fnn2.w2", "fnn2.b2",
    "value.w0", "value.b0", "value.w1", "value.b1", "value.w2", "value.b2",
)
# IQN actor (StochasticActor): the first eight actor slots, the other five are empty
VAR_NAMES_IQN = ("actor.wm", "actor.bm", "actor.wn", "actor.bn", "actor.wo", "actor.bo", "actor.ws", "actor.bs",
                 "actor._8", "actor._9", "actor._10", "actor._11", "actor._12") + VAR_NAMES[13:]
NET_NAMES = ("actor", "fnn1", "fnn2", "value")
ACTOR_KINDS = {"AIQN": 0, "IQN": 1}


def var_names(kind: int):
    return VAR_NAMES_IQN if kind == 1 else VAR_NAMES


class Layout(C.Structure):
    _fields_ = [("offsets", C.c_int64 * (NVARS + 1)), ("rows", C.c_int32 * NVARS), ("cols", C.c_int32 * NVARS),
                ("net_begin", C.c_int64 * (NNETS + 1))]


class Config(C.Structure):
    _fields_ = [("S", C.c_int32), ("A", C.c_int32), ("B", C.c_int32), ("K", C.c_int32), ("chunk_rows", C.c_int32),
                ("engine", C.c_int32), ("actor_kind", C.c_int32)]


class HParams(C.Structure):
    _fields_ = [("gamma", C.c_float), ("rho", C.c_float), ("q_normalization", C.c_float), ("beta", C.c_float),
                ("mode", C.c_int32), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float)]


class StepIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "online", "target", "adam_m", "adam_v", "grad", "adam_t", "s", "a", "r", "sp", "done", "critic_u", "value_tau",
        "filter_tau", "filter_normal", "filter_rand", "actor_tau", "sel_idx", "q_out", "v_out", "cand_actions",
        "value_actions")]


# every symbol include/gac_b200.h declares: (name, restype, argtypes)
_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "gac_layout": (C.c_int, [_I, _I, C.POINTER(Layout)]),
    "gac_layout_ex": (C.c_int, [_I, _I, _I, C.POINTER(Layout)]),
    "gac_gru_step": (C.c_int, [_P, _P, _P, _I, _P, _P, _P, _I, _P, _I]),
    "gac_iqn_forward": (C.c_int, [_P, _P, _P, _I, _P, _P, _I, _P]),
    "gac_iqn_train_fwd": (C.c_int, [_P, _P, _P, _I, _P, _P, _I, _P, _P]),
    "gac_iqn_bwd": (C.c_int, [_P, _P, _P, _P, _I]),
    "gac_workspace_bytes": (C.c_size_t, [C.POINTER(Config)]),
    "gac_create": (C.c_int, [C.POINTER(Config), _P, C.c_size_t, _P, C.POINTER(_P)]),
    "gac_destroy": (None, [_P]),
    "gac_set_stream": (C.c_int, [_P, _P]),
    "gac_last_error": (C.c_char_p, []),
    "gac_abi_version": (C.c_int, []),
    "gac_polyak": (C.c_int, [_P, _P, _L, _F, _P]),
    "gac_adam_polyak": (C.c_int, [_P, _P, _P, _P, _P, _P, C.POINTER(HParams), _P, _P, _P]),
    "gac_aiqn_sample": (C.c_int, [_P, _P, _P, _I, _P, _P, _I, _P]),
    "gac_critic_eval": (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _P]),
    "gac_value_eval": (C.c_int, [_P, _P, _P, _I, _P]),
    "gac_advantage_compact": (C.c_int, [_P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P]),
    "gac_aiqn_supervised_fwd": (C.c_int, [_P, _P, _P, _I, _P, _P, _P, _I, _P, _P]),
    "gac_qhuber_loss_bwd": (C.c_int, [_P, _P, _P, _P, _P, _I, _P, _I, _F, _P, _P]),
    "gac_aiqn_bwd": (C.c_int, [_P, _P, _P, _P, _I]),
    "gac_td_fwd_bwd": (C.c_int, [_P, _P, _I, _P, _P, _I, _P, _P]),
    "gac_step_grads": (C.c_int, [_P, C.POINTER(StepIO), C.POINTER(HParams)]),
    "gac_step_apply": (C.c_int, [_P, C.POINTER(StepIO), C.POINTER(HParams)]),
    "gac_replay_gather": (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "gac_gemm": (C.c_int, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I]),
    "gac_launch_count": (_L, [_P]),
}


class GacError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension is the only compute path of dpo_replication_b200 "
            "(no CPU fallback). Build it with `python __graft_entry__.py`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()


def check(rc: int):
    if rc != 0:
        raise GacError(f"libgac_b200 error {rc}: {lib.gac_last_error().decode()}")


def layout(S: int, A: int, kind: int = 0) -> Layout:
    lay = Layout()
    check(lib.gac_layout_ex(S, A, kind, C.byref(lay)))
    return lay
