"""Second-generation GRU step kernel (csrc/gru_step_v2.cuh) through the C-ABI entry gac_gru_step: error against an fp64
evaluation of the same step on the GPU, time per launch, and (GAC_GRU_TIMELINE=1) the MMA lane's per-CTA counters.
    python tools/gru2_probe.py [--rows 32768] [--reps 20]          (GAC_GRU_V1=1: round 1's kernel, for comparison)"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import dpo_replication_b200 as g  # noqa: E402
from dpo_replication_b200.engine import Arena, keras_default_init  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=32768)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--first", action="store_true", help="first action dimension (h_prev = 0)")
    a = ap.parse_args()
    S, A, K = 17, 6, 64
    B = max(-(-a.rows // (2 * K)), 1)
    rows = a.rows
    eng = g.Engine(S, A, B, K, "cuda")
    ar = Arena(S, A, "cuda")
    gen = torch.Generator().manual_seed(0)
    p = keras_default_init(S, A, "actor", gen)
    for k in ("ba", "b1", "bn", "gb", "bs"):
        p[k] = 0.05 * torch.randn(p[k].shape, generator=gen)
    ar.load(ar.online, "actor", p)
    actor = ar.net_slice(ar.online, "actor")
    dev = "cuda"
    states = torch.randn(B, S, device=dev)
    sidx = torch.randint(0, B, (rows,), device=dev, dtype=torch.int32)
    x = torch.rand(rows, 1, device=dev) * 2 - 1
    ipi = torch.arange(1, 65, device=dev, dtype=torch.float32) * torch.tensor(np.float32(np.pi), device=dev)
    wa, ba = p["wa"].to(dev), p["ba"].to(dev)
    ae = F.leaky_relu(torch.cos(x * ipi) @ wa + ba, 0.2).contiguous()
    h = None if a.first else torch.tanh(torch.randn(rows, 400, device=dev)).contiguous()
    out = eng.gru_step(actor, states, sidx, ae, h, actor_inputs=True)
    torch.cuda.synchronize()
    # fp64 reference of the same step
    P = {k: v.to(dev).double() for k, v in p.items()}
    se = F.leaky_relu(F.leaky_relu(states.double() @ P["ws"] + P["bs"], 0.01), 0.2)
    xx = torch.cat([se[sidx.long()], ae.double()], 1)
    mx = xx @ P["kx"] + P["gb"][0]
    hh = torch.zeros(rows, 400, device=dev, dtype=torch.float64) if h is None else h.double()
    mh = hh @ P["u"] + P["gb"][1]
    z = torch.sigmoid(mx[:, :400] + mh[:, :400]); r = torch.sigmoid(mx[:, 400:800] + mh[:, 400:800])
    c = torch.tanh(mx[:, 800:] + r * mh[:, 800:])
    ref = z * hh + (1 - z) * c
    err = (out.double() - ref).abs()
    print(f"rows={rows} first={a.first} kernel={'v1' if os.environ.get('GAC_GRU_V1') else 'v3 (cta_group::2)' if os.environ.get('GAC_GRU_PAIR') else 'v2'}: max |h - fp64| = {err.max().item():.3e}  mean = {err.mean().item():.3e}  "
          f"(max |h| = {ref.abs().max().item():.3f}); worst column {int(err.max(0).values.argmax())}, worst row {int(err.max(1).values.argmax())}")
    for _ in range(3):
        eng.gru_step(actor, states, sidx, ae, h, reuse=True, actor_inputs=True, same_inputs=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        out2 = eng.gru_step(actor, states, sidx, ae, h, reuse=True, actor_inputs=True, same_inputs=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    fl = 2.0 * rows * (400 if a.first else 800) * 1200
    print(f"  {ms * 1e3:8.1f} us per launch = {fl / (ms * 1e-3) / 1e12:6.1f} TFLOP/s algorithmic; repeat is bit-identical: {bool(torch.equal(out, out2))}")
    if os.environ.get("GAC_GRU_TIMELINE"):
        arr = eng.workspace.view(torch.int64).cpu().numpy()
        hits = np.nonzero(arr == 0x600DC0DE600DC0E4)[0]
        if hits.size:
            n = int(arr[hits[0] + 1])
            rec = arr[hits[0] + 8: hits[0] + 8 + 16 * n].reshape(n, 16)
            lead = rec[:, 0] > 0                      # the pair kernel's peer CTAs issue no MMAs (their lane counters are zero)
            L = rec[lead]
            pct = lambda k: 100 * L[:, k].sum() / L[:, 0].sum()
            tl = L[:, 4].sum()
            tl_all = tl * n / max(int(lead.sum()), 1)  # (128 rows x 80 units) tiles over all CTAs
            print(f"  SM clock {1e3 * L[:, 0].sum() / L[:, 5].sum():.0f} MHz;  per tile (cycles): MMA lane {L[:, 0].sum() / tl:.0f}"
                  f" | epilogue: wait drain {rec[:, 7].sum() / tl_all:.0f} wait accum {rec[:, 6].sum() / tl_all:.0f} load issue {rec[:, 14].sum() / tl_all:.0f} phase1 {rec[:, 8].sum() / tl_all:.0f} bar {rec[:, 9].sum() / tl_all:.0f}"
                  f" phase2 {rec[:, 10].sum() / tl_all:.0f} bar {rec[:, 11].sum() / tl_all:.0f} | drain warps: wait {rec[:, 12].sum() / tl_all:.0f} work {rec[:, 13].sum() / tl_all:.0f}")
            print(f"  MMA lane per issuing CTA ({int(lead.sum())} of {n} CTAs): cycles mean {L[:, 0].mean():.0f} max {L[:, 0].max()}  waiting for operands {pct(1):.1f}%  "
                  f"for a free accumulator {pct(2):.1f}%  for the drain warps {pct(3):.1f}%  tiles {L[:, 4].mean():.2f} (max {L[:, 4].max()})")
        else:
            print("  no counters found")


if __name__ == "__main__":
    main()
