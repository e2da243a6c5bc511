"""Times the fused GRU-step kernel through the sampler entry point (gac_aiqn_sample): 6 autoregressive steps on
`rows` rows; prints ms per call.  GAC_GRU_PROBE bits (timing experiments, results garbage): 1 = one MMA pass of three,
2 = no A global loads, 4 = no weight bulk copies, 8 = no A smem stores, 16 = no MMAs except the first-touch ones.
    python tools/gru_probe.py [--rows 32768] [--reps 10]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import dpo_replication_b200 as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=32768)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--timeline", action="store_true", help="needs GAC_GRU_TIMELINE=1 in the environment")
    a = ap.parse_args()
    S, A, K = 17, 6, 64
    B = a.rows // K
    eng = g.Engine(S, A, B, K, "cuda")
    from dpo_replication_b200.engine import Arena, keras_default_init
    arena = Arena(S, A, "cuda")
    arena.load(arena.online, "actor", keras_default_init(S, A, "actor", torch.Generator().manual_seed(0)))
    actor = arena.net_slice(arena.online, "actor")
    states = torch.randn(B, S, device="cuda")
    sidx = torch.arange(B, device="cuda", dtype=torch.int32).repeat_interleave(K)
    taus = torch.rand(a.rows, A, device="cuda")
    for _ in range(2):
        eng.aiqn_sample(actor, states, sidx, taus)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        eng.aiqn_sample(actor, states, sidx, taus)
    e1.record()
    torch.cuda.synchronize()
    print(f"probe={os.environ.get('GAC_GRU_PROBE', '0')} rows={a.rows}: {e0.elapsed_time(e1) / a.reps:8.3f} ms per 6-step sample")
    if a.timeline:
        import numpy as np
        arr = eng.workspace.view(torch.int64).cpu().numpy()
        hits = np.nonzero(arr == 0x600DC0DE600DC0DE)[0]
        if hits.size == 0:
            print("no stamps (set GAC_GRU_TIMELINE=1)")
            return
        n = 7 * (a.rows // 128)
        rec = arr[hits[0] + 8: hits[0] + 8 + 8 * n].reshape(n, 8)      # stamps of the LAST launch (step d = A-1)
        t0 = rec[:, 0].min()
        names = ["setup", "loader loop", "wait accum", "phase1 (TMEM->smem)", "phase2 (gates+IO)"]
        for label, sel in (("nu=64 tiles", np.arange(n) % 7 != 6), ("nu=16 tiles", np.arange(n) % 7 == 6)):
            r = rec[sel]
            d = np.diff(r[:, 1:7], axis=1)
            print(f"{label}: n={r.shape[0]}  " + "  ".join(f"{nm} {d[:, i].mean():8.0f}" for i, nm in enumerate(names)) +
                  f"  total {(r[:, 6] - r[:, 1]).mean():8.0f} cycles")
        # per-SM: gap between one CTA's end and the next CTA's start (launch + TMEM/smem hand-over)
        gaps = []
        for sm in np.unique(rec[:, 7]):
            r = rec[rec[:, 7] == sm]
            r = r[np.argsort(r[:, 1])]
            gaps += list(r[1:, 1] - r[:-1, 6])
        gaps = np.array(gaps)
        print(f"CTAs per SM {n / np.unique(rec[:, 7]).size:.2f}; gap end->next start: mean {gaps.mean():.0f} median {np.median(gaps):.0f} cycles")
        print(f"kernel span (globaltimer): {(rec[:, 0].max() - t0) / 1e3:.1f} us to the last CTA start")


if __name__ == "__main__":
    main()
