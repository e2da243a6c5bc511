#!/usr/bin/env python
"""Benchmark of the GAC train step (BASELINE.json): sampled (s,a) pairs/s and train steps/s
at HalfCheetah dims (S=17, A=6, B=256 states per GPU, K=64 samples) on 1/2/4/8 B200.

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
  python bench.py --impl reference ...                   # reference-equivalent CPU path

A "step" is one full GACAgent.train_one_step (agent.py:121-168): Critic.train, Value.train
(P sampler rows), the advantage filter (P sampler rows + 2P critic rows), IQNActor.train on the
M selected rows (supervised forward + BPTT), Adam x4, Polyak x3.  Data is synthetic
(SURVEY.md §8d); data-parallel ranks each own B states (weak scaling) and exchange one SUM
all-reduce of the gradient arena per step.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (S, A, B per rank, K)
    "halfcheetah": (17, 6, 256, 64),      # BASELINE.json configs[1] -- the metric's config
    "pendulum": (3, 1, 64, 32),           # configs[0]
    "humanoid": (376, 17, 256, 128),      # configs[2]
    "ant": (111, 8, 16384, 256),          # configs[3]: B is the GLOBAL batch, split over ranks (strong scaling)
    "walker2d": (17, 6, 1000000, 32),     # configs[4]: forward-only sampler sweep over 1M states x 32 + Polyak bandwidth
}
CONFIG_NUM = {"pendulum": 1, "halfcheetah": 2, "humanoid": 3, "ant": 4, "walker2d": 5}
METRIC = "GAC sampled (s,a) pairs/s through full train steps (HalfCheetah dims; steps/s in steps_per_s)"


# ---------------------------------------------------------------------------------------------
# algorithmic work: F_min counts the exact hoists, F_ref the reference as written
# ---------------------------------------------------------------------------------------------
def step_flops(S, A, B, K, M):
    P = B * K
    samp_ref = 2 * (400 * S + 1651600 * A)
    samp_min = 2 * ((400 * S + 480000) / K + 1171600 * A - 480000)
    critic = 2 * 2 * ((S + A) * 400 + 120300)
    sup_fwd = 2 * (400 * S + A * (2 * 25600 + 1440000 + 160400))
    f_ref = 2 * P * samp_ref + 3 * P * critic + 3 * M * sup_fwd
    f_min = 2 * P * samp_min + 3 * P * critic + 3 * M * sup_fwd
    return f_ref, f_min


def csrc_hash():
    """sha256 over the kernel sources: ties a measured-traffic record to the code it was captured from."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "dpo_replication_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(kernel, rows):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of `kernel`, from the `ncu --set full` capture that
    tools/ncu_traffic.py summarised into profiles/traffic.json.  Only returned when that capture was taken from the
    kernel sources as they are now (same csrc hash); otherwise null -- never a number from older code."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, "no capture"
    rec = json.load(open(p)).get(kernel)
    if not rec or rec.get("rows") != rows:
        return None, "no capture for this kernel / row count"
    if rec.get("csrc_hash") != csrc_hash():
        return None, f"capture {rec.get('source')} is from older kernel sources"
    return rec["dram_bytes_read"] + rec["dram_bytes_write"], rec.get("source")


def peaks():
    """Measured roofline denominators (driver-written MEASURED_PEAKS.json) or the stated fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (profiling recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [ln.split(",") for ln in open(self.f.name).read().strip().splitlines() if ln.count(",") >= 6]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[0]) for r in rows]
        out["sm_mhz"] = float(np.median(sm))
        out["sm_max_mhz"] = float(rows[0][1])
        out["power_w_max"] = max(float(r[2]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out["reasons"] = [n for i, n in enumerate(names) if any("Active" in r[3 + i] and "Not" not in r[3 + i] for r in rows)]
        out["samples"] = len(rows)
        return out


def _synth_module():
    """dpo_replication_b200/synth.py loaded BY PATH (NumPy only): the reference arm must not import the package, which
    would load libgac_b200.so into the CPU process."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_gac_synth", os.path.join(ROOT, "dpo_replication_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


WEIGHT_SEED = 1234


def synth(S, A, B, K, cfg_index, rank):
    """Initial agent state (identical on every rank and in both arms) and rank `rank`'s replay batch + RNG draws."""
    sy = _synth_module()
    nets = sy.agent_state(np.random.default_rng(WEIGHT_SEED), S, A)
    rng = np.random.default_rng(1000 * cfg_index + rank)         # seed = 1000 * config# + rank
    return nets, sy.replay_batch(rng, S, A, B), sy.step_noise(rng, A, B, K)


def bench_config(name, S, A, B, K, world, scaling):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": f"{name}: S={S} A={A} B={B}/GPU K={K} full GAC train step (AIQN actor, twin critic, value)",
            "global_batch": B * world, "action_samples": K, "parallelism": f"dp{world}", "scaling": scaling,
            "candidate_rows_per_gpu": 2 * B * K,
            "l2": "per-step working set (activations saved for BPTT) is GBs >> 126 MB L2; no flush needed",
            "v_target": "after one step, target/online value.b2 shifted by median(q - v) of that step's candidates (about half are selected)"}


# ---------------------------------------------------------------------------------------------
# reference arm: the reference-equivalent CPU path (TensorFlow is not installable -- DESIGN.md)
# ---------------------------------------------------------------------------------------------
def time_cpu_reference(name, S, A, B, K, steps, warmup, budget_s=None):
    """The reference-equivalent CPU path (oracle/torch_ref.py: torch-CPU eager, op for op as GAC/networks.py + agent.py)
    on the SAME initial state, batch, draws and V_target centring as the GPU arm, all host threads."""
    import torch
    from oracle import torch_ref as T
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nets, batch, noise = synth(S, A, B, K, CONFIG_NUM[name], 0)
    agent = T.TorchGAC(nets, A, K)
    info = agent.train_one_step(batch, noise)                    # the GPU arm's centring step
    shift = float((info["filter"]["q"] - info["filter"]["v"]).median())
    with torch.no_grad():
        agent.n["target_value"]["b2"] += shift
        agent.n["value"]["b2"] += shift
    for _ in range(warmup):
        agent.train_one_step(batch, noise)
    t0 = time.perf_counter()
    done, msum = 0, 0
    for _ in range(steps):
        msum += agent.train_one_step(batch, noise)["M"]
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dt / done, done, cores, msum / done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    S, A, B, K = CONFIGS[args.config]
    scaling = "weak"
    if args.config == "ant":
        B, scaling = B // world, "strong"
    if args.config in ("ant", "walker2d"):
        print(json.dumps({"impl": "reference", "unavailable": f"the CPU path is timed on the metric's config only, not on --config {args.config}"}))
        return 0
    # one step = the full B-state batch of ONE rank (at N = 1: the whole job's batch; at N > 1 the bounded sample is one
    # rank's shard of the N*B global batch -- the rate, pairs/s, does not depend on the sample size)
    sec, done, cores, m_mean = time_cpu_reference(args.config, S, A, B, K, args.steps, args.warmup)
    pairs = 2 * B * K / sec
    sample = f"{done} full train steps on {B} states x K={K}" + ("" if world == 1 else f" (one rank's shard of the {B * world}-state global batch)")
    line = {"impl": "reference", "metric": METRIC, "value": pairs, "unit": "pairs/s", "n_gpus": args.gpus, "steps": done,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.config, S, A, B, K, world, scaling),
            "steps_per_s": 1.0 / sec, "selected_rows_mean": m_mean,
            "note": "reference-equivalent CPU path (torch-CPU eager op-for-op mirror of GAC/networks.py + agent.py; TensorFlow is not installable here)",
            "cpu_baseline": {"value": pairs, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": pairs, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    import __graft_entry__ as ge
    ge.build()
    from dpo_replication_b200 import GACAgent

    S, A, B, K = CONFIGS[args.config]
    scaling = "weak"
    if args.config == "ant":              # strong scaling on the fixed global batch
        B, scaling = B // world, "strong"
    P = B * K
    torch.manual_seed(1234 + rank)        # device-side draws of the e2e loop differ per rank; the weights come from synth()
    agent = GACAgent(A, S, action_samples=K, batch_size=B, device=dev, engine=args.engine, chunk_rows=args.chunk_rows, actor=args.actor)
    nets_np, batch_np, noise_np = synth(S, A, B, K, CONFIG_NUM[args.config], rank)
    if args.actor == "AIQN":
        agent.load_state(nets_np)         # the same initial state as the reference arm (and on every rank)
    to_dev = lambda d: {k: torch.from_numpy(v).to(dev) for k, v in d.items()}
    batch, noise = to_dev(batch_np), to_dev(noise_np)

    # centre V_target so that about half of the 2P candidates have positive advantage
    taps = dict(q_out=torch.zeros(2 * P, device=dev), v_out=torch.zeros(2 * P, device=dev))
    agent.train_on_batch(batch, noise, taps)
    shift = (taps["q_out"] - taps["v_out"]).median()
    if world > 1:
        dist.all_reduce(shift)
        shift /= world
    for buf in (agent.arena.target, agent.arena.online):
        agent.arena.view(buf, "value.b2").add_(shift)

    m_log = torch.zeros(args.steps + args.warmup + 1, device=dev)

    use_graph = bool(args.graph)             # world > 1: two graphs (grads, apply) around the NCCL all-reduce
    step_fn = agent.train_on_batch_graphed if use_graph else agent.train_on_batch
    launches_per_step = None
    if use_graph:                          # the context's launch counter only sees eager launches: count one eager step
        l0 = agent.engine.launch_count()
        agent.train_on_batch(batch, noise)
        launches_per_step = agent.engine.launch_count() - l0

    def step(i):
        step_fn(batch, noise)
        m_log[i].copy_(agent.arena.stats[5])

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = agent.engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    launches = agent.engine.launch_count() - launches0
    if launches_per_step is not None:
        launches = launches_per_step * args.steps      # kernels executed by the graph replays
    clk = clocks.stop() if clocks else None
    m_mean = float(m_log[args.warmup:args.warmup + args.steps].mean().item()) / world   # stats hold the all-reduced (global) M

    # ---- end to end: host (pinned) replay batch -> H2D -> public API step -> D2H stats, every step ----
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in batch_np.items()}
    stats_host = torch.zeros(8).pin_memory()
    h2d = sum(v.numel() * 4 for v in pinned.values())

    m_e2e = torch.zeros(1, device=dev)

    def e2e_step():
        b = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        (agent.train_on_batch_graphed if use_graph else agent.train_on_batch)(b)   # RNG drawn on device, as the reference's tf.random does
        stats_host.copy_(agent.arena.stats, non_blocking=True)
        m_e2e.add_(agent.arena.stats[5])                     # fresh random draws every step => its own selected-row count

    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    m_e2e.zero_()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2.item()) / args.steps

    # ---- roofline of the dominant kernel, timed live with CUDA events on the launching stream.
    # AIQN actor on the tcgen05 engine: gru_step_v2_kernel (one GRU step of the sampler on its 2P candidate rows: both
    # input projections + gates, 3xFP16 arithmetic; profiles/r2*_launches_summary.txt: the largest single kernel of
    # the step; the loop below goes through the C-ABI entry gac_gru_step, so each iteration also pays the ~3 us stand-by
    # launch of round 1's kernel that takes over when the weights admit no fp16 scaling).  Algorithmic flops per launch = 2 * rows * (400 + 400) * 1200 (the GRU's two
    # projections).  Otherwise: the packed-weight GEMM (rows x 400) * (400 x 1200). ----
    roof = None
    if rank == 0:
        eng = agent.engine
        rows = 2 * P
        n_rep = 20
        pk = peaks()
        if args.actor == "AIQN" and args.engine == 0:
            ar = agent.arena
            actor = ar.net_slice(ar.target, "actor")
            states = torch.randn(B, S, device=dev)
            # the sampler's own row -> state map (sidx3_kernel): rows [0, P) state-major b * K + k (networks.py:464-467), rows
            # [P, 2P) sample-major k * B + b (agent.py:186) -- the gather pattern of the per-state projection matters to the kernel
            r_ = torch.arange(P, device=dev, dtype=torch.int32)
            sidx = torch.cat([r_ // K, r_ % B]).contiguous()
            wa, ba = ar.view(ar.target, "actor.wa"), ar.view(ar.target, "actor.ba")
            ae = torch.nn.functional.leaky_relu(torch.cos(torch.rand(rows, 1, device=dev) * 2 - 1) @ wa[:1] + ba, 0.2)   # inside the actor's bound
            h = torch.tanh(torch.randn(rows, 400, device=dev))
            eng.gru_step(actor, states, sidx, ae, h, actor_inputs=True)
            for _ in range(3):
                eng.gru_step(actor, states, sidx, ae, h, reuse=True, actor_inputs=True, same_inputs=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n_rep):
                eng.gru_step(actor, states, sidx, ae, h, reuse=True, actor_inputs=True, same_inputs=True)
            e1.record()
            torch.cuda.synchronize()
            kms = e0.elapsed_time(e1) / n_rep
            ach = 2.0 * rows * 800 * 1200 / (kms * 1e-3) / 1e12
            peak = pk["bf16_burst"] / 3.0
            roof = {"bound": "tensor", "kernel": "gru_step_v2_kernel (persistent, bulk-copied pre-split operands, 3xFP16: kind::f16 MMAs, 3 passes per product)",
                    "shape": f"rows={rows}: [ae | h_prev] ({rows}x800) * [kx_a ; U] (800x1200) + gates", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak, "traffic": measured_traffic("gru_step_v2_kernel", rows)[0],
                    "traffic_source": measured_traffic("gru_step_v2_kernel", rows)[1], "ms_per_launch": kms,
                    "algorithmic_bytes": rows * 400 * 4 * 3,
                    "mma_issue_bound": {"cycles_per_mma": 168, "source": "profiles/r2b_mma_rate.txt (tools/mma_rate.cu): an M = 128 kind::f16 tcgen05.mma takes ~168 SM cycles whatever N <= 256 and cta_group",
                                        "mmas_per_launch_per_sm_max": 9 * 150, "note": "N = 240 of 256 columns per instruction and 9 tiles on the busiest SM bound this formulation at ~0.86 of the peak below"},
                    "peak_source": f"{pk['source']} bf16 burst {pk['bf16_burst']} TF/s / 3 (three fp16 passes per fp32 product)"}
        else:
            a = torch.randn(rows, 400, device=dev)
            w = torch.randn(400, 1200, device=dev)
            keng = 2 if args.engine == 0 else 1          # the step's forward GEMMs use the packed-weight path
            for _ in range(3):
                eng.gemm(a, w, engine=keng)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n_rep):
                eng.gemm(a, w, engine=keng)
            e1.record()
            torch.cuda.synchronize()
            kms = e0.elapsed_time(e1) / n_rep
            tf32_peak = pk["bf16_burst"] / 2.0
            peak = tf32_peak / 3.0 if args.engine == 0 else 74.0
            ach = 2.0 * rows * 400 * 1200 / (kms * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "gemm_tc_kernel<3xTF32, packed weights> (incl. the 5 us weight pack)" if args.engine == 0 else "gemm_simt_kernel",
                    "shape": f"({rows}x400)*(400x1200)", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": None, "ms_per_launch": kms,
                    "peak_source": (f"{pk['source']} bf16 burst {pk['bf16_burst']} TF/s / 2 (TF32) / 3 (3xTF32 passes)" if args.engine == 0
                                    else "fp32 SIMT: 148 SM x 128 FMA x 2 x 1.965 GHz")}

    hbm = hbm_kernels(agent, dev) if (rank == 0 and world == 1) else None
    if world > 1:
        dist.barrier()
    ant = None
    if args.config == "halfcheetah" and args.actor == "AIQN" and not args.no_ant:
        try:
            ant = ant_strong_probe(dev, rank, world, engine=args.engine)
        except Exception as exc:                                         # the probe must never take the headline line down
            ant = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    if rank == 0:
        pairs_per_step = 2 * P * world
        value = pairs_per_step / (ms_per_step * 1e-3)
        f_ref, f_min = step_flops(S, A, B, K, m_mean) if args.actor == "AIQN" else (float("nan"), float("nan"))
        pk = peaks()
        cpu = None
        if world == 1 and not args.no_cpu_baseline and args.actor == "AIQN" and args.config in ("halfcheetah", "pendulum"):
            sec, done, cores, m_cpu = time_cpu_reference(args.config, S, A, B, K, steps=20, warmup=0, budget_s=15.0)
            cpu = {"value": 2 * B * K / sec, "unit": "pairs/s", "cores": cores, "kind": "port", "selected_rows_mean": m_cpu,
                   "sample": f"{done} full train steps of the reference-equivalent torch-CPU path on the same {B} states x K={K}, same initial state, draws and V_target centring"}
        cfg = bench_config(args.config, S, A, B, K, world, scaling)
        tf = f_min * world / (ms_per_step * 1e-3) / 1e12
        line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": cfg,
                "engine": "tcgen05: 3xFP16 (kind::f16) GRU step / heads / dgrads / large wgrads, 3xTF32 elsewhere" if args.engine == 0 else "SIMT fp32",
                "actor": args.actor, "selected_rows_mean": m_mean, "candidate_rows": 2 * P, "graph": use_graph,
                "steps_per_s": 1e3 / ms_per_step,
                "step_flops": {"algorithmic_gflop_min": f_min / 1e9, "as_written_gflop": f_ref / 1e9,
                               "achieved_tflops_min": tf,
                               # the hot kernels issue kind::f16 MMAs (3 passes per fp32 product): that is the honest whole-step ceiling
                               "frac_of_3xf16_sustained_peak": tf / world / (pk["bf16_sustained"] / 3.0),
                               "frac_of_3xtf32_sustained_peak": tf / world / (pk["bf16_sustained"] / 6.0),
                               "note": "3xf16: bf16_sustained / 3 (the issued MMA kind of the dominant kernels); 3xtf32: bf16_sustained / 6 (what a pure kind::tf32 path could reach)"},
                "clocks": clk, "gpu_launches": int(launches),
                "e2e": {"value": pairs_per_step / (e2e_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": 32, "ms_per_step": e2e_ms,
                        "selected_rows_mean": float(m_e2e.item()) / args.steps / world,
                        "note": "fresh device-side random draws every step (the kernel-only loop replays one fixed draw), so the data-dependent number of selected rows differs"},
                "roofline": roof, "cpu_baseline": cpu, "hbm_kernels": hbm}
        if ant is not None:
            line["ant_strong"] = ant
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def ant_strong_probe(dev, rank, world, steps=2, engine=0):
    """configs[3]: Ant-v2 dims, GLOBAL batch 16384 states x K=256 samples split over the ranks (strong scaling: the job
    is fixed, each rank takes 16384 / N states), full train steps with the NCCL SUM all-reduce.  Carried in every bench
    line so that the driver's N = 1, 2, 4, 8 runs give the strong-scaling curve; per-rank selected rows show the load
    imbalance the data-dependent filter causes."""
    import torch
    import torch.distributed as dist
    from dpo_replication_b200 import GACAgent
    S, A, BG, K = CONFIGS["ant"]
    B = BG // world
    P = B * K
    nets = _synth_module().agent_state(np.random.default_rng(WEIGHT_SEED), S, A)
    agent = GACAgent(A, S, action_samples=K, batch_size=B, device=dev, engine=engine)
    agent.load_state(nets)
    g = torch.Generator(device=dev).manual_seed(4000 + rank)
    r = lambda *sh: torch.rand(*sh, device=dev, generator=g)
    n = lambda *sh: torch.randn(*sh, device=dev, generator=g)
    batch = dict(s=n(B, S), a=r(B, A) * 2 - 1, r=n(B, 1), sp=n(B, S), it=(r(B, 1) < 0.05).float())
    noise = dict(critic_u=r(B, A), value_tau=r(P, A), filter_tau=r(P, A), filter_normal=n(P, A), filter_rand=r(P, A) * 2 - 1,
                 actor_tau=r(2 * P, A))
    taps = dict(q_out=torch.zeros(2 * P, device=dev), v_out=torch.zeros(2 * P, device=dev))
    agent.train_on_batch(batch, noise, taps)
    shift = (taps["q_out"] - taps["v_out"]).median()
    if world > 1:
        dist.all_reduce(shift)
        shift /= world
    for buf in (agent.arena.target, agent.arena.online):
        agent.arena.view(buf, "value.b2").add_(shift)
    del taps
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    m_local = torch.zeros(1, device=dev)
    agent.train_on_batch(batch, noise)                                   # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    io = agent.engine.step_io(agent.arena, batch, noise)
    t_grads = 0.0
    e[0].record()
    for _ in range(steps):
        # the two phases timed apart: gradients (rank-local, data dependent) and exchange + update
        e[2].record()
        agent.engine.step_grads(io, agent._hp)
        e[3].record()
        m_local.add_(agent.arena.stats[5])                               # local M, before the all-reduce sums it
        if world > 1:
            dist.all_reduce(agent.arena.grad)
        agent.engine.step_apply(io, agent._hp)
        e[3].synchronize()
        t_grads += e[2].elapsed_time(e[3])
    e[1].record()
    torch.cuda.synchronize()
    per = torch.tensor([e[0].elapsed_time(e[1]) / steps, t_grads / steps, float(m_local.item()) / steps], device=dev)
    allr = [torch.zeros_like(per) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, per)
    else:
        allr = [per]
    ms = max(float(x[0]) for x in allr)
    _, f_min = step_flops(S, A, B, K, float(per[2]))
    out = {"workload": f"ant: S={S} A={A} global batch {BG} states x K={K}, {B} states per GPU, full GAC train step",
           "scaling": "strong", "n_gpus": world, "steps": steps, "ms_per_step": ms, "pairs_per_s": 2 * BG * K / (ms * 1e-3),
           "grads_phase_ms_per_rank": [float(x[1]) for x in allr], "selected_rows_per_rank": [float(x[2]) for x in allr],
           "candidate_rows_per_rank": 2 * P, "achieved_tflops_min_rank0": f_min / (float(per[1]) * 1e-3) / 1e12,
           "note": "ms_per_step = slowest rank (max over ranks, CUDA events); grads phase excludes the all-reduce and the optimiser"}
    del agent, batch, noise
    torch.cuda.empty_cache()
    return out


def hbm_kernels(agent, dev, reps=10):
    """Achieved HBM GB/s of the streaming optimiser kernels on a stacked arena >= 1 GiB (a population
    of parameter sets): at the real 8 MB model size they are L2 resident and launch bound."""
    import torch
    eng = agent.engine
    n = 1 << 28                                  # 2^28 floats = 1 GiB per array
    t = torch.zeros(n, device=dev); s_ = torch.ones(n, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        eng.polyak(t, s_, 5e-3)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        eng.polyak(t, s_, 5e-3)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pk = peaks()
    gbs = 12.0 * n / (ms * 1e-3) / 1e9            # read t, read s, write t
    out = {"polyak": {"bytes_per_param": 12, "params": n, "ms": ms, "achieved_gbs": gbs, "peak_gbs": pk["hbm"], "frac": gbs / pk["hbm"]}}
    del t, s_
    # the model-sized fused Adam+Polyak sweep (L2 resident): report time and effective bandwidth
    ar = agent.arena
    hp = eng.hparams()
    for _ in range(2):
        eng.adam_polyak(ar.online, ar.target, ar.adam_m, ar.adam_v, ar.grad, hp, ar.adam_t)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        eng.adam_polyak(ar.online, ar.target, ar.adam_m, ar.adam_v, ar.grad, hp, ar.adam_t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out["adam_polyak_model_size"] = {"bytes_per_param": 36, "params": ar.n, "ms": ms, "effective_gbs": 36.0 * ar.n / (ms * 1e-3) / 1e9,
                                     "note": "8 MB working set: L2 resident, launch-latency bound (2 launches)"}
    return out


def run_sweep(args):
    """configs[4]: forward-only AIQN sampling over 1,000,000 DISTINCT states x 32 samples (Walker2d dims), streamed in
    chunks of 4096 states (+ one 576-state tail), + the Polyak / delayed-actor update bandwidth.  Two legs:
    `value`: all states resident in HBM, every chunk reads its own slice, draws fresh taus on the device (the reference draws
    them inside the actor, networks.py:202) and writes its own slice of the (32M, A) action array;
    `e2e`: the states start in pinned HOST memory and the actions end there -- chunk c + 1's H2D copy and chunk c - 1's D2H
    copy run on a second stream under chunk c's kernels."""
    import torch
    import __graft_entry__ as ge
    ge.build()
    from dpo_replication_b200 import GACAgent
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    S, A, NS, K = CONFIGS["walker2d"]
    chunk_states = 4096
    if args.sweep_chunks:
        NS = min(NS, args.sweep_chunks * chunk_states)
    agent = GACAgent(A, S, action_samples=K, batch_size=chunk_states, device=dev, engine=args.engine, chunk_rows=128)
    eng, ar = agent.engine, agent.arena
    base = ar.net_slice(ar.target, "actor")
    gen = torch.Generator().manual_seed(11)
    states_host = torch.randn(NS, S, generator=gen).pin_memory()
    out_host = torch.empty(NS * K, A).pin_memory()
    states_dev = states_host.to(dev)
    out_dev = torch.empty(NS * K, A, device=dev)
    dgen = torch.Generator(device=dev).manual_seed(12)
    bounds = [(c0, min(c0 + chunk_states, NS)) for c0 in range(0, NS, chunk_states)]
    sidx_full = torch.arange(chunk_states, dtype=torch.int32, device=dev).repeat_interleave(K)

    def sample(st, c0, c1, dst):
        n = c1 - c0
        taus = torch.rand(n * K, A, device=dev, generator=dgen)
        eng.aiqn_sample(base, st, sidx_full[:n * K], taus, out=dst)

    for c0, c1 in bounds[:2]:
        sample(states_dev[c0:c1], c0, c1, out_dev[c0 * K:c1 * K])
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for c0, c1 in bounds:
        sample(states_dev[c0:c1], c0, c1, out_dev[c0 * K:c1 * K])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    # every chunk produced its own actions (distinct states, distinct draws): no two chunk slices are equal, all are in (-1, 1)
    n0 = chunk_states * K
    distinct = bool(len(bounds) < 2 or not torch.equal(out_dev[:n0], out_dev[n0:2 * n0]))
    finite = bool(torch.isfinite(out_dev).all().item() and out_dev.abs().max().item() <= 1.0)

    # e2e: host -> device -> host, double buffered on a copy stream
    main_s, copy_s = torch.cuda.current_stream(), torch.cuda.Stream()
    st_buf = [torch.empty(chunk_states, S, device=dev) for _ in range(2)]
    act_buf = [torch.empty(chunk_states * K, A, device=dev) for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    torch.cuda.synchronize()
    e0.record()
    with torch.cuda.stream(copy_s):
        c0, c1 = bounds[0]
        st_buf[0][:c1 - c0].copy_(states_host[c0:c1], non_blocking=True)
        ev_in[0].record(copy_s)
    for i, (c0, c1) in enumerate(bounds):
        b = i & 1
        if i + 1 < len(bounds):                                          # prefetch the next chunk's states
            n0, n1 = bounds[i + 1]
            with torch.cuda.stream(copy_s):
                if i >= 1:
                    copy_s.wait_event(ev_done[1 - b])                    # chunk i - 1 has finished reading st_buf[1 - b]
                st_buf[1 - b][:n1 - n0].copy_(states_host[n0:n1], non_blocking=True)
                ev_in[1 - b].record(copy_s)
        main_s.wait_event(ev_in[b])
        if i >= 2:
            main_s.wait_event(ev_out[b])                                 # act_buf[b] has been copied out (chunk i - 2)
        sample(st_buf[b][:c1 - c0], c0, c1, act_buf[b][:(c1 - c0) * K])
        ev_done[b].record(main_s)
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(ev_done[b])
            out_host[c0 * K:c1 * K].copy_(act_buf[b][:(c1 - c0) * K], non_blocking=True)
            ev_out[b].record(copy_s)
    main_s.wait_stream(copy_s)
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    pairs = NS * K
    f_min = 2 * ((400 * S + 480000) / K + 1171600 * A - 480000)
    line = {"metric": "AIQN forward-only sampled (s,a) pairs/s (Walker2d dims, 1M states x 32, streamed)", "value": pairs / (ms * 1e-3),
            "unit": "pairs/s", "n_gpus": 1, "ms_total": ms, "pairs": pairs, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"walker2d sweep: S={S} A={A} states={NS} (all distinct) K={K} chunk={chunk_states} states, {len(bounds)} chunks, fresh taus per chunk"},
            "achieved_tflops_min": pairs * f_min / (ms * 1e-3) / 1e12, "gpu_launches": int(launches),
            "checks": {"chunks_distinct": distinct, "actions_finite_in_unit_box": finite},
            "e2e": {"value": pairs / (e2e_ms * 1e-3), "unit": "pairs/s", "ms_total": e2e_ms, "h2d_bytes_total": NS * S * 4, "d2h_bytes_total": NS * K * A * 4,
                    "note": "states from pinned host memory, actions back to pinned host memory, copies double-buffered on a second stream"},
            "hbm_kernels": hbm_kernels(agent, dev)}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="halfcheetah", choices=sorted(CONFIGS))
    ap.add_argument("--engine", type=int, default=0, help="0 = tcgen05 3xTF32 GEMMs, 1 = SIMT fp32")
    ap.add_argument("--chunk-rows", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ant", action="store_true", help="skip the Ant strong-scaling probe (ant_strong key)")
    ap.add_argument("--graph", type=int, default=1, help="1 = replay the train step from CUDA graphs (one graph; with ranks > 1 two graphs around the NCCL all-reduce), 0 = eager launches")
    ap.add_argument("--actor", default="AIQN", choices=["AIQN", "IQN"], help="AIQN = BASELINE.json's actor; IQN = the reference CLI default (main.py:78)")
    ap.add_argument("--sweep-chunks", type=int, default=0, help="walker2d sweep: number of 4096-state chunks (0 = all 1M states)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "walker2d":
        return run_sweep(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
