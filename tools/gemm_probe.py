"""Micro-benchmark of the GEMM engines through the C ABI (gac_gemm): times a list of shapes with
CUDA events and prints TFLOP/s.  Also the target of the `ncu --set full` captures in profiles/.
    python tools/gemm_probe.py [--engine 0|1] [--reps 20] [--shapes M,N,K,ta,tb ...]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import dpo_replication_b200 as g  # noqa: E402

DEFAULT = ["32768,1200,400,0,0", "32768,400,400,0,0", "32768,400,1200,0,1", "98304,1200,400,0,0", "1200,400,98304,1,0",
           "400,1200,98304,1,0", "32768,400,64,0,0", "49152,300,400,0,0"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", type=int, default=0)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--shapes", nargs="*", default=DEFAULT)
    a = ap.parse_args()
    eng = g.Engine(17, 6, 256, 64, "cuda")
    for sh in a.shapes:
        M, N, K, ta, tb = (int(x) for x in sh.split(","))
        A = torch.randn((K, M) if ta else (M, K), device="cuda")
        B = torch.randn((N, K) if tb else (K, N), device="cuda")
        for _ in range(3):
            eng.gemm(A, B, bool(ta), bool(tb), engine=a.engine)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            eng.gemm(A, B, bool(ta), bool(tb), engine=a.engine)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        print(f"engine={a.engine} M={M} N={N} K={K} ta={ta} tb={tb}: {ms * 1e3:9.1f} us  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s")


if __name__ == "__main__":
    main()
