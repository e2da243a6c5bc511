"""The C-ABI library loads on a CPU-only box and exports every symbol include/gac_b200.h
declares; layout and argument validation need no GPU."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

from oracle import gac_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from dpo_replication_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "gac_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gac_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    raw = C.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert lib.lib.gac_abi_version() == 4


@pytest.mark.parametrize("S,A", [(3, 1), (17, 6), (376, 17), (111, 8)])
def test_layout_matches_reference_variable_shapes(lib, S, A):
    lay = lib.layout(S, A)
    shapes = {}
    for k, v in O.actor_shapes(S, A).items():
        shapes["actor." + k] = v
    for net, n_in in (("fnn1", S + A), ("fnn2", S + A), ("value", S)):
        for k, v in O.fnn_shapes(n_in).items():
            shapes[f"{net}.{k}"] = v
    assert len(shapes) == lib.NVARS == 31                 # 13 + 6 + 6 + 6 variables
    total = 0
    for i, name in enumerate(lib.VAR_NAMES):
        r, c = lay.rows[i], lay.cols[i]
        assert r * c == int(np.prod(shapes[name])), name
        assert lay.offsets[i] % 32 == 0                   # 128-byte aligned for 128-bit streaming
        assert lay.offsets[i + 1] - lay.offsets[i] >= r * c
        total += r * c
    # parameter counts of SURVEY.md App. D
    expect = {(3, 1): 2024204, (17, 6): 2050604, (376, 17): 2633804, (111, 8): 2202604}[(S, A)]
    assert total == expect
    assert list(lay.net_begin)[0] == 0 and lay.net_begin[4] == lay.offsets[31]


def test_argument_validation_without_gpu(lib):
    assert lib.lib.gac_layout(0, 1, C.byref(lib.Layout())) == -1
    assert b"gac_layout" in lib.lib.gac_last_error()
    cfg = lib.Config(17, 6, 256, 64, 0, 0)
    n = lib.lib.gac_workspace_bytes(C.byref(cfg))
    assert 1 << 30 < n < 40 << 30                          # a few GB at the HalfCheetah config
    assert lib.lib.gac_workspace_bytes(C.byref(lib.Config(0, 6, 1, 1, 0, 0))) == 0
    assert lib.lib.gac_polyak(None, None, 4, 0.5, None) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dpo_replication_b200 import GACAgent
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GACAgent(1, 3)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dpo_replication_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("the oracle", "").replace("by the oracle", ""), f


@pytest.mark.parametrize("S,A", [(17, 6), (376, 17), (3, 1)])
def test_iqn_layout_matches_reference_variable_shapes(lib, S, A):
    # StochasticActor (networks.py:283-302): 8 actor variables in tf.Module order, then the same three FNNs
    lay = lib.layout(S, A, 1)
    shapes = O.iqn_shapes(S, A)
    for i, k in enumerate(("wm", "bm", "wn", "bn", "wo", "bo", "ws", "bs")):
        assert lay.rows[i] * lay.cols[i] == int(np.prod(shapes[k])), k
        assert lay.offsets[i] % 32 == 0
    assert all(lay.rows[i] == 0 for i in range(8, 13))
    assert lay.net_begin[1] == lay.offsets[13] and lay.net_begin[4] == lay.offsets[31]
    assert lib.lib.gac_layout_ex(S, A, 7, C.byref(lib.Layout())) == -1


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference-equivalent CPU path) prints ONE JSON line with the contract's keys;
    it needs no GPU and never touches the CUDA library."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "pendulum", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "impl", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "pairs/s"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
