"""Per-CTA phase stamps of the kx_a weight-gradient GEMM inside a real train step (set GAC_WGRAD_TIMELINE=1):
setup, loader + MMA loop, wait for the last MMAs, epilogue; per-K-block cost and CTAs per SM.
    GAC_WGRAD_TIMELINE=1 python tools/wgrad_timeline.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import dpo_replication_b200 as g  # noqa: E402


def main():
    S, A, B, K = 17, 6, 256, 64
    agent = g.GACAgent(A, S, action_samples=K, batch_size=B, device="cuda")
    rng = np.random.default_rng(0)
    batch = dict(s=rng.standard_normal((B, S)).astype(np.float32), a=rng.uniform(-1, 1, (B, A)).astype(np.float32),
                 r=rng.standard_normal((B, 1)).astype(np.float32), sp=rng.standard_normal((B, S)).astype(np.float32),
                 it=(rng.uniform(size=(B, 1)) > 0.05).astype(np.float32))
    for _ in range(3):
        agent.train_on_batch(batch)
    torch.cuda.synchronize()
    arr = agent.engine.workspace.view(torch.int64).cpu().numpy()
    hits = np.nonzero(arr == 0x600DC0DE600DC0DF)[0]
    if hits.size == 0:
        print("no stamps (set GAC_WGRAD_TIMELINE=1)")
        return
    n = int(arr[hits[0] + 1])
    rec = arr[hits[0] + 8: hits[0] + 8 + 8 * n].reshape(n, 8)
    rec = rec[rec[:, 6] > 0]
    d = np.diff(rec[:, 1:6], axis=1)
    names = ["setup", "loader+MMA loop", "wait last MMAs", "epilogue"]   # (GAC_DGRAD_TIMELINE=1: the kx_a input-gradient dot-GEMM instead)
    print(f"{rec.shape[0]} CTAs with work of {n}; K blocks per CTA: mean {rec[:, 6].mean():.1f}")
    print("  ".join(f"{nm} {d[:, i].mean():9.0f}" for i, nm in enumerate(names)), f"  total {(rec[:, 5] - rec[:, 1]).mean():9.0f} cycles")
    print(f"loop cycles per K block: {(d[:, 1] / rec[:, 6]).mean():.0f}")
    print(f"CTAs per SM: {rec.shape[0] / np.unique(rec[:, 7]).size:.2f}; kernel span {(rec[:, 0].max() - rec[:, 0].min()) / 1e3:.1f} us to the last CTA start")


if __name__ == "__main__":
    main()
