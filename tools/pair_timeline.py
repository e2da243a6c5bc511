"""Timeline of CTA 0 of the cta_group::2 GEMM (clock64 stamps written by the kernel when
GAC_TC_TIMELINE is set): per tile, MMA-thread wait for TMEM, first full stage, last MMA issued;
epilogue start / end."""
import ctypes as C
import os
import sys

import torch

os.environ["GAC_TC_PAIR"] = "1"   # the timeline probe is about the opt-in cta_group::2 kernel
os.environ["GAC_TC_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import dpo_replication_b200 as g  # noqa: E402

eng = g.Engine(17, 6, 256, 64, "cuda")
A = torch.randn(32768, 400, device="cuda")
B = torch.randn(400, 1200, device="cuda")
for _ in range(3):
    c = eng.gemm(A, B, engine=2)
torch.cuda.synchronize()
# the stamps live at the start of the split-K scratch buffer: find it through a second engine call
# (gac_gemm with engine 2 wrote them); read them back via a raw pointer exposed by the workspace layout
ws = eng.workspace
raw = ws.view(torch.int64)
# locate: stamps are monotonically increasing clock values; search for the first block of 8 where [0]<[1]<=[2]<[3]
import numpy as np
arr = raw.cpu().numpy()
found = None
for off in range(0, arr.size - 128, 32):
    blk = arr[off:off + 128].reshape(16, 8)
    if blk[0, 0] > 1e6 and (blk[:8, 0] < blk[:8, 3]).all() and (blk[1:8, 0] >= blk[:7, 0]).all() and blk[0, 3] - blk[0, 0] < 1e7 and (blk[:8, 6] == 0).all():
        found = blk
        break
if found is None:
    print("stamps not found")
else:
    t00 = found[0, 0]
    print("tile  wait_tmem  first_full  mma_issue_done | epi_start  epi_end   (cycles from tile 0 start)")
    for i in range(6):
        r = found[i] - t00
        print(f"{i:3d} start={r[0]:8d} tmem_ok={r[1]:8d} first_full={r[2]:8d} issued={r[3]:8d} | epi_start={r[4]:8d} epi_end={r[5]:8d}")
    e = arr[off + 300:off + 304]
    print(f"epilogue group (warp 8, tile 2): tmem ld+wait+add {e[1]-e[0]}  st.shared+syncwarp {e[2]-e[1]}  ld.shared+epilogue+st.global+syncwarp {e[3]-e[2]} cycles")
