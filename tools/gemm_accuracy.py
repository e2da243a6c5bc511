"""Accuracy of the GEMM engines vs fp64 as a function of the reduction length (the tensor core
truncates on accumulate, so the 3xTF32 error grows with the number of K steps per accumulator)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import dpo_replication_b200 as g  # noqa: E402

eng = g.Engine(17, 6, 256, 64, "cuda")
torch.manual_seed(0)
for K in (512, 2048, 8192, 32768, 131072):
    M, N = 384, 400
    A = torch.randn(K, M, device="cuda")
    B = torch.randn(K, N, device="cuda")
    ref = (A.double().T @ B.double())
    for e in (0, 1):
        c = eng.gemm(A, B, True, False, engine=e).double()
        err = float((c - ref).abs().max() / ref.abs().max())
        bias = float(((c - ref) * torch.sign(ref)).mean() / ref.abs().mean())
        print(f"K={K:7d} engine={e}: normwise err {err:.2e}   signed bias (toward zero < 0) {bias:+.2e}")
