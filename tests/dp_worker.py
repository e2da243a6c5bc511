"""One data-parallel rank of tests/test_dp_gpu.py (launched with torch.distributed.run, NCCL, one GPU per rank): takes its
shard of the global batch, runs ONE train step through the public GACAgent API and writes what the parent compares."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    inp, out = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import dpo_replication_b200 as gb
    from dpo_replication_b200 import dist as gdist
    z = np.load(inp, allow_pickle=True)
    nets, batch, noise = z["nets"].item(), z["batch"].item(), z["noise"].item()
    S, A, K = int(z["S"]), int(z["A"]), int(z["K"])
    B = batch["s"].shape[0]
    Bl = B // world
    torch.manual_seed(100 + rank)                       # replicas start DIFFERENT; the constructor must broadcast rank 0's state
    agent = gb.GACAgent(A, S, action_samples=K, batch_size=Bl, device=dev)
    start = agent.arena.online.clone()
    ref = [torch.zeros_like(start) for _ in range(world)]
    dist.all_gather(ref, start)
    same_start = all(torch.equal(ref[0], r) for r in ref)
    if rank != 0:                                       # a wrong state on the non-zero ranks: load_state must broadcast rank 0's
        nets = {k: ({n: v * 0 for n, v in d.items()} if isinstance(d, dict) else d) for k, d in nets.items()}
    agent.load_state(nets)
    b, n = gdist.shard_step_inputs(batch, noise, K, rank, world)
    if f"actor_tau_rank{rank}" in z.files:              # the draws this rank's compacted rows have in the one-process run
        n["actor_tau"] = z[f"actor_tau_rank{rank}"]
    P = Bl * K
    taps = dict(sel_idx=torch.zeros(2 * P, dtype=torch.int64, device=dev), q_out=torch.zeros(2 * P, device=dev), v_out=torch.zeros(2 * P, device=dev))
    agent.train_on_batch(b, n, taps)
    torch.cuda.synchronize()
    m_local = int((taps["q_out"] > taps["v_out"]).sum().item())
    idx_local = taps["sel_idx"][:m_local].cpu().numpy()
    idx_global = gdist.local_to_global_candidate(idx_local, Bl, K, rank, world)
    st = agent.dump_state()
    flat = {f"{k}/{n}": v for k, d in st.items() if isinstance(d, dict) for n, v in d.items()}
    np.savez(f"{out}.rank{rank}.npz", idx_global=idx_global, m_local=m_local, grad=agent.arena.grad.cpu().numpy(), same_start=same_start,
             t=np.array([st["t_" + k] for k in ("actor", "fnn1", "fnn2", "value")]), **flat)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
