// GRU step of the AIQN actor, third generation (sm_100a): gru_step_v2.cuh as a cta_group::2 kernel.
//
//   same arithmetic as gru_step_v2.cuh (Keras GRU cell, reset_after, gate order [z|r|h]; GAC/networks.py:166,217,259)
//
// OPT-IN (GAC_GRU_PAIR=1), numerically validated, level with v2 (163.5 vs 166.3 us per 32768-row step in tools/gru2_probe.py,
// equal in bench.py's loop: profiles/r2_pair_relay_probe.txt).  History: written on the hypothesis that the SM's shared-memory
// data pipe bounds v2; the microbenchmarks that followed (tools/mma_rate.cu, tools/mma_issue.cu) showed v2's tile time was the
// issuing lane's per-instruction overhead instead, identical for cta_group 1 and 2.  Its first version lost 20 % to the
// stage relay: an explicit .release.cluster on the remote mbarrier arrive costs a MEMBAR.ALL.GPU per stage (gemm_tc.cuh,
// mbar_arrive_leader).  What the pair buys is operand traffic:
//   * cta_group::2: the two CTAs of a cluster form one M = 256 x N = 240 instruction.  Each SM keeps its own 128 rows of A
//     and only HALF of the weight tile (120 of the 240 rows), so an instruction fetches 32 + 30 instead of 32 + 60
//     wavefronts per SM, a stage is 31 KB instead of 46 KB (5 stages fit), and the weight tile crosses L2 -> SM once per
//     row-block PAIR.
//   * every MMA of a tile is the same N = 240 instruction.  v2 split the first K step of the recurrent half into an
//     accumulating N = 160 part (z, r) and an overwriting N = 80 part (g); with cta_group::2 the column halves come from
//     different CTAs and no such split maps onto them.  Instead the drain warps, which copy g = ae * kx_h out of TMEM when the
//     action-embedding half has retired, also store ZEROS over it (tcgen05.st), and the recurrent half simply accumulates.
//   * one lane of the leader CTA issues for the pair; the peer's MMA lane relays "my stage is full"; commits are multicast to
//     both CTAs' barriers; the epilogue / drain warps of both CTAs release accumulators by arriving on the leader's barriers.
// Everything else (a16 operand images, persistent static schedule, 16 epilogue + 8 drain warps, double-buffered TMEM) is v2's.
#pragma once
#include "gru_step_v2.cuh"

namespace gac {
namespace tc {

constexpr uint32_t G3_B_HALF = (uint32_t)(3 * G2_NU / 2) * 64u;     // 120 weight rows x 64 bytes: this CTA's half of hi (or lo)
constexpr uint32_t G3_B_SLOT = 2u * G3_B_HALF;                      // [hi | lo] of this CTA's half: 15 KB
constexpr uint32_t G3_STAGE = G2_A_SLOT + G3_B_SLOT;                // 31 KB
constexpr int G3_STAGES = 6;
constexpr int G3_SMEM = G3_STAGES * (int)G3_STAGE + BM * G2_SS * 4 + G2_DRAIN_SMEM + 4 * 3 * G2_NU * 4 + 1024;
static_assert((3 * G2_NU / 2) % 8 == 0 && (3 * G2_NU) % 16 == 0, "cta_group::2: N multiple of 16, each CTA's half in whole 8-row swizzle groups");
static_assert(G3_STAGE % 512 == 0 && G3_B_HALF % 512 == 0, "SWIZZLE_64B tiles start on 512-byte boundaries");

// weights for the pair kernel: slot (tile j, block kb2) = [rank 0: hi | lo][rank 1: hi | lo]; rank r holds tile rows
// [120 r, 120 r + 120) of the [z | r | h] x 80 tile (same row order as gru2_pack_kernel, hence the same TMEM columns)
__global__ void gru3_pack_kernel(const float* __restrict__ kxa, const float* __restrict__ U, unsigned char* __restrict__ out,
                                 const Gru2Scales* __restrict__ sc) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_slot = 3 * G2_NU * 4;
  if (idx >= (long long)G2_NT * 2 * G2_KB * per_slot || !sc->ok) return;
  const int c = idx % 4, r = (idx / 4) % (3 * G2_NU);
  const int slot = idx / per_slot, kb2 = slot % (2 * G2_KB), j = slot / (2 * G2_KB);
  const int half = kb2 >= G2_KB, kb = kb2 - half * G2_KB;
  const int gate = r / G2_NU, u = j * G2_NU + r % G2_NU;
  const float* W = half ? U : kxa;
  const float sw = half ? sc->s_u : sc->s_kx;
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k = kb * G2_BK + 8 * c + 2 * e;
    const float x0 = k < GAC_E ? W[(size_t)k * GAC_G + gate * GAC_E + u] * sw : 0.f;
    const float x1 = k + 1 < GAC_E ? W[(size_t)(k + 1) * GAC_G + gate * GAC_E + u] * sw : 0.f;
    split_f16x2_u(x0, x1, h[e], l[e]);
  }
  const int rank = r / (3 * G2_NU / 2), rr = r - rank * (3 * G2_NU / 2);
  unsigned char* base = out + (size_t)slot * G2_B_SLOT + (size_t)rank * G3_B_SLOT;
  const uint32_t off = sw64_off(rr, c);
  *reinterpret_cast<uint4*>(base + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(base + G3_B_HALF + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void mma_f16_2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// one K block (two K steps x three passes) / one K step as ONE instruction group, cta_group::2 (see mma_f16_kblock in v2)
template <uint32_t A_LO, uint32_t B_HI, uint32_t B_LO>
__device__ __forceinline__ void mma_f16_kblock_2(uint32_t d, uint64_t a_hi, uint32_t idesc, uint32_t acc0) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b64 al, bh, bl, ah2, al2, bh2, bl2;\n"
      "setp.ne.b32 q, %3, 0;\n"
      "setp.eq.b32 p, 1, 1;\n"
      "add.u64 al, %1, %4;\n"
      "add.u64 bh, %1, %5;\n"
      "add.u64 bl, %1, %6;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, bh, %2, q;\n"
      "add.u64 ah2, %1, 2;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], al, bh, %2, p;\n"
      "add.u64 bh2, bh, 2;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, bl, %2, p;\n"
      "add.u64 al2, al, 2;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], ah2, bh2, %2, p;\n"
      "add.u64 bl2, bl, 2;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], al2, bh2, %2, p;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], ah2, bl2, %2, p;\n"
      "}\n" ::"r"(d), "l"(a_hi), "r"(idesc), "r"(acc0), "n"(A_LO), "n"(B_HI), "n"(B_LO)
      : "memory");
}
template <uint32_t A_LO, uint32_t B_HI, uint32_t B_LO>
__device__ __forceinline__ void mma_f16_kstep_2(uint32_t d, uint64_t a_hi, uint32_t idesc, uint32_t acc0) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b64 al, bh, bl;\n"
      "setp.ne.b32 q, %3, 0;\n"
      "setp.eq.b32 p, 1, 1;\n"
      "add.u64 al, %1, %4;\n"
      "add.u64 bh, %1, %5;\n"
      "add.u64 bl, %1, %6;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, bh, %2, q;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], al, bh, %2, p;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, bl, %2, p;\n"
      "}\n" ::"r"(d), "l"(a_hi), "r"(idesc), "r"(acc0), "n"(A_LO), "n"(B_HI), "n"(B_LO)
      : "memory");
}

template <bool DBG>
__global__ void __launch_bounds__(G2_THREADS, 1) gru_step_v3_kernel(const Gru2Args g) {
  unsigned long long* const dbgp = DBG ? g.dbg : nullptr;
  const int probe = DBG ? g.probe : 0;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[G3_STAGES], bar_pfull[G3_STAGES], bar_empty[G3_STAGES], bar_tfull[2], bar_tempty[2], bar_xfull[2], bar_xfree[2],
      bar_xsaved[2];
  __shared__ uint32_t tmem_slot;

  if (!g.sc->ok) {                                                 // no valid scaling (both CTAs decide alike): plain fp32 (gru_step_v2.cuh)
    gru2_fallback(g, g.dyn_rows ? min(g.rows, *g.dyn_rows) : g.rows, g.hprev != nullptr);
    return;
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int cluster_id = (int)blockIdx.x >> 1, nclusters = (int)gridDim.x >> 1;
  const int m_end = g.dyn_rows ? min(g.rows, *g.dyn_rows) : g.rows;
  const int n_rb = (m_end + BM - 1) / BM;
  const int n_items = ((n_rb + 1) / 2) * G2_NT;                    // (row-block pair, unit tile), pair major
  const bool has_h = g.h_pk != nullptr;
  const int nkb = has_h ? 2 * G2_KB : G2_KB;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  float* stg = reinterpret_cast<float*>(smem_raw + (smem0 - smem_u32(smem_raw)) + G3_STAGES * G3_STAGE);
  float* dstg = stg + BM * G2_SS;
  float* b1_base = dstg + G2_DRAIN_WARPS * 32 * G2_DS;
  float* xh_cta = g.xh_scratch + (size_t)blockIdx.x * (G2_XH_BYTES_PER_CTA / 4);

  if (tid == 0) {
    for (int s = 0; s < G3_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);                         // this CTA's producer (+ its bytes)
      mbar_init(smem_u32(&bar_pfull[s]), 1);                        // leader only: the peer's stage is full too
      mbar_init(smem_u32(&bar_empty[s]), 1);                        // the pair's MMAs on this stage retired (multicast commit)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tfull[b]), 1);                        // multicast commit
      mbar_init(smem_u32(&bar_tempty[b]), 2 * G2_EPI_WARPS);        // leader only: both CTAs' epilogue warps
      mbar_init(smem_u32(&bar_xfull[b]), 1);                        // multicast commit
      mbar_init(smem_u32(&bar_xfree[b]), 2 * G2_DRAIN_WARPS);       // leader only: both CTAs' drain warps
      mbar_init(smem_u32(&bar_xsaved[b]), G2_DRAIN_WARPS);          // local
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == G2_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // the peer's barriers exist before anyone signals them
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (warp == G2_PROD_WARP) {
    // ------------------------------ producer: one lane, two bulk copies per K block (own A rows, own half of B) ------------------------------
    if (lane == 0) {
      int it = 0;
      for (int i = cluster_id; i < n_items; i += nclusters) {
        const int rbp = i / G2_NT, jt = i - rbp * G2_NT;
        const int rb = min(2 * rbp + (int)crank, n_rb - 1);         // a dead second row block re-reads the last live one (results unused)
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % G3_STAGES;
          const uint32_t ph = (uint32_t)(it / G3_STAGES) & 1u;
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
          const uint32_t sa = smem0 + (uint32_t)s * G3_STAGE, bar = smem_u32(&bar_full[s]);
          const int half = kb >= G2_KB, kbl = kb - half * G2_KB;
          const unsigned char* asrc = (half ? g.h_pk : g.ae_pk) + ((size_t)rb * G2_KB + kbl) * G2_A_SLOT;
          const unsigned char* bsrc = g.wpk + ((size_t)jt * 2 * G2_KB + kb) * G2_B_SLOT + (size_t)crank * G3_B_SLOT;
          mbar_expect_tx_arrive(bar, G3_STAGE);
          bulk_g2s(sa, asrc, G2_A_SLOT, bar);
          bulk_g2s(sa + G2_A_SLOT, bsrc, G3_B_SLOT, bar);
        }
      }
    }
    __syncwarp();
  } else if (warp == G2_MMA_WARP) {
    if (lane == 0 && crank == 0) {
      // ------------------------------ leader: one lane issues the pair's MMAs ------------------------------
      long long t_wait_full = 0, t_wait_acc = 0, t_wait_x = 0;
      const long long t_begin = dbgp ? clock64() : 0;
      unsigned long long gt0 = 0;
      if (dbgp) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt0));
      constexpr uint32_t id3 = (1u << 4) | ((uint32_t)((3 * G2_NU) >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // kind::f16, M = 256, N = 240
      const uint64_t desc0 = make_desc64(smem0);                   // A hi of stage 0
      int it = 0, t = 0;
      for (int i = cluster_id; i < n_items; i += nclusters, ++t) {
        const int b = t & 1;
        const uint32_t tph = (uint32_t)(t >> 1) & 1u;
        long long c0 = dbgp ? clock64() : 0;
        mbar_wait(smem_u32(&bar_tempty[b]), tph ^ 1u);             // both CTAs' epilogues are done with this accumulator (two tiles ago)
        tc_fence_after();
        if (dbgp) t_wait_acc += clock64() - c0;
        const uint32_t d0 = tmem_d + (uint32_t)b * 256u;           // columns [z | r | g], 80 each, in both CTAs' TMEM
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % G3_STAGES;
          const uint32_t ph = (uint32_t)(it / G3_STAGES) & 1u;
          c0 = dbgp ? clock64() : 0;
          mbar_wait(smem_u32(&bar_full[s]), ph);
          mbar_wait(smem_u32(&bar_pfull[s]), ph);
          tc_fence_after();
          if (dbgp) t_wait_full += clock64() - c0;
          if (kb == G2_KB) {
            // the recurrent half starts: the drain warps of both CTAs have copied g = ae * kx_h out and zeroed it
            c0 = dbgp ? clock64() : 0;
            mbar_wait(smem_u32(&bar_xfree[b]), tph);
            tc_fence_after();
            if (dbgp) t_wait_x += clock64() - c0;
          }
          const uint64_t a_hi = desc0 + (uint64_t)((uint32_t)s * (G3_STAGE >> 4));
          constexpr uint32_t C_ALO = G2_A_TILE >> 4, C_BHI = G2_A_SLOT >> 4, C_BLO = (G2_A_SLOT + G3_B_HALF) >> 4;
          const bool last_of_half = kb == G2_KB - 1 || kb == 2 * G2_KB - 1;       // 400 = 12 x 32 + 16: one K step
          if (DBG && (probe & 32)) {
          } else if (last_of_half) {
            mma_f16_kstep_2<C_ALO, C_BHI, C_BLO>(d0, a_hi, id3, 1u);
          } else {
            mma_f16_kblock_2<C_ALO, C_BHI, C_BLO>(d0, a_hi, id3, kb ? 1u : 0u);
          }
          mma_commit2_mc(smem_u32(&bar_empty[s]));                 // frees the stage in both CTAs
          if (has_h && kb == G2_KB - 1) mma_commit2_mc(smem_u32(&bar_xfull[b]));   // the action-embedding half has retired
        }
        mma_commit2_mc(smem_u32(&bar_tfull[b]));
      }
      if (dbgp) {
        unsigned long long* d = dbgp + 8 + (size_t)blockIdx.x * 16;
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt1));
        d[0] = (unsigned long long)(clock64() - t_begin); d[1] = (unsigned long long)t_wait_full; d[2] = (unsigned long long)t_wait_acc;
        d[3] = (unsigned long long)t_wait_x; d[4] = (unsigned long long)t; d[5] = gt1 - gt0;
        if (blockIdx.x == 0) { dbgp[0] = 0x600DC0DE600DC0E4ull; dbgp[1] = gridDim.x; }
      }
    } else if (lane == 0) {
      // ------------------------------ peer: relay "my stage is full" to the leader's MMA lane ------------------------------
      int it = 0;
      for (int i = cluster_id; i < n_items; i += nclusters)
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % G3_STAGES;
          mbar_wait(smem_u32(&bar_full[s]), (uint32_t)(it / G3_STAGES) & 1u);
          mbar_arrive_leader(smem_u32(&bar_pfull[s]));
        }
      if (dbgp) { unsigned long long* d = dbgp + 8 + (size_t)blockIdx.x * 16; d[0] = d[1] = d[2] = d[3] = d[4] = d[5] = 0; }
    }
    __syncwarp();
  } else if (warp == G2_PROD_WARP + 1) {
    // h_prev prefetcher (see v2): the NEXT tile's 128 x 80 block of this CTA towards L2 while the current tile's MMAs run
    if (has_h) {
      int t = 0;
      for (int i = cluster_id; i < n_items; i += nclusters, ++t) {
        if (t > 0) mbar_wait(smem_u32(&bar_tfull[(t - 1) & 1]), (uint32_t)((t - 1) >> 1) & 1u);
        const int rbp = i / G2_NT, jt = i - rbp * G2_NT, rb = 2 * rbp + (int)crank;
#pragma unroll
        for (int k = 0; k < 12; ++k) {
          const int idx = k * 32 + lane, row = idx / 3, part = idx - row * 3;
          const int m = rb * BM + row;
          if (m < m_end) {
            const float* p = g.hprev + (size_t)m * GAC_E + jt * G2_NU + part * 32;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
          }
        }
      }
    }
  } else if (warp > G2_PROD_WARP && warp < G2_DRAIN_WARP0) {
    // idle warp of the control warpgroup
  } else if (warp >= G2_DRAIN_WARP0) {
    // ------------------------------ drain warps: g = ae * kx_h out of TMEM between the two halves, zeros back in ------------------------------
    if (has_h) {
      const int quarter = warp & 3;
      const float inv_unit = g.sc->inv_unit;
      const bool stamp = dbgp && tid == G2_DRAIN_WARP0 * 32;
      long long dw = 0, dk = 0;
      int t = 0;
      for (int i = cluster_id; i < n_items; i += nclusters, ++t) {
        const int b = t & 1;
        const uint32_t tph = (uint32_t)(t >> 1) & 1u;
        long long c0 = stamp ? clock64() : 0;
        mbar_wait(smem_u32(&bar_xfull[b]), tph);
        tc_fence_after();
        long long c1 = stamp ? clock64() : 0;
        dw += c1 - c0;
        const int hsel = (warp - G2_DRAIN_WARP0) >> 2;             // which 40 of the 80 g columns
        const uint32_t d0 = tmem_d + (uint32_t)b * 256u + 2u * G2_NU + (uint32_t)(hsel * 40) + ((uint32_t)(quarter * 32) << 16);
        uint32_t v[40];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(d0));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                       "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(d0 + 16u));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                     : "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39])
                     : "r"(d0 + 32u));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        {
          // the recurrent half accumulates into g from zero
          const uint32_t zr = 0u;
          asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(d0), "r"(zr) : "memory");
          asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(d0 + 16u), "r"(zr) : "memory");
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(d0 + 32u), "r"(zr) : "memory");
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(smem_u32(&bar_xfree[b]));   // the leader's MMA lane may continue with this tile
        if (!(probe & 1)) {
          float* blk = dstg + (size_t)(warp - G2_DRAIN_WARP0) * 32 * G2_DS;
          float* dst = xh_cta + ((size_t)b * BM + quarter * 32) * G2_NU + hsel * 40;
#pragma unroll
          for (int ci = 0; ci < 5; ++ci) {
#pragma unroll
            for (int e = 0; e < 8; e += 4)
              *reinterpret_cast<float4*>(blk + lane * G2_DS + 4 * ((e >> 2) ^ ((lane >> 2) & 1))) =
                  make_float4(__uint_as_float(v[ci * 8 + e]) * inv_unit, __uint_as_float(v[ci * 8 + e + 1]) * inv_unit,
                              __uint_as_float(v[ci * 8 + e + 2]) * inv_unit, __uint_as_float(v[ci * 8 + e + 3]) * inv_unit);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int rr = (lane >> 1) + 16 * k, qd = lane & 1;
              __stcg(reinterpret_cast<float4*>(dst + (size_t)rr * G2_NU + ci * 8 + 4 * qd),
                     *reinterpret_cast<const float4*>(blk + rr * G2_DS + 4 * (qd ^ ((rr >> 2) & 1))));
            }
            __syncwarp();
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_xsaved[b]));
        if (stamp) dk += clock64() - c1;
      }
      if (stamp) { unsigned long long* d = dbgp + 8 + (size_t)blockIdx.x * 16; d[12] = (unsigned long long)dw; d[13] = (unsigned long long)dk; }
    }
  } else {
    // ------------------------------ epilogue: sixteen warps, overlapped with the next tile's main loop (gru_step_v2.cuh) ------------------------------
    const int quarter = warp & 3, sub = warp >> 2;
    const float inv_unit = g.sc->inv_unit, s_h = g.sc->s_h;
    auto tnh = [](float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); };
    // four independent quadrant groups with their own named barrier, staging rows and bias copy (see v2)
    const int grp = quarter, gw = sub;
    const int gtid = gw * 32 + lane;
    const uint32_t gbar = 1u + (uint32_t)grp;
    float* b1_s = b1_base + grp * (3 * G2_NU);
#define G3_GROUP_SYNC() asm volatile("bar.sync %0, 128;" ::"r"(gbar) : "memory")
    const int prow = 32 * grp + gw * 8 + (lane >> 2), pj = lane & 3;
    const bool stamp = dbgp && tid == 0;
    long long e_wt = 0, e_wx = 0, e_p1 = 0, e_b1 = 0, e_p2 = 0, e_b2 = 0, e_ld = 0, ec = 0;
#define G3_STAMP(acc) if (stamp) { const long long now = clock64(); acc += now - ec; ec = now; }
    int t = 0;
    for (int i = cluster_id; i < n_items; i += nclusters, ++t) {
      const int rbp = i / G2_NT, jt = i - rbp * G2_NT;
      const int m0 = (2 * rbp + (int)crank) * BM, b = t & 1;       // a dead second row block has m0 >= m_end: no loads, no stores
      if (stamp) ec = clock64();
      const uint32_t tph = (uint32_t)(t >> 1) & 1u;
      const int m = m0 + prow;
      const int sid = m < m_end ? g.sidx[m] : -1;
      if (gtid < 3 * G2_NU / 4)
        *reinterpret_cast<float4*>(b1_s + 4 * gtid) = *reinterpret_cast<const float4*>(g.b1 + (gtid / (G2_NU / 4)) * GAC_E + jt * G2_NU + 4 * (gtid % (G2_NU / 4)));
      if (has_h) mbar_wait(smem_u32(&bar_xsaved[b]), tph);
      G3_STAMP(e_wx)
      float4 sz, sr, sh, hp, xh;
      auto load_ops = [&](int q) {
        hp = make_float4(0.f, 0.f, 0.f, 0.f); xh = hp;
        if (probe & 4) { sz = sr = sh = hp; return; }
        if (sid < 0) return;
        const int ul = q * 16 + 4 * pj, u = jt * G2_NU + ul;
        const float* sx = g.sx + (size_t)sid * GAC_G + u;
        sz = __ldg(reinterpret_cast<const float4*>(sx)); sr = __ldg(reinterpret_cast<const float4*>(sx + GAC_E));
        sh = __ldg(reinterpret_cast<const float4*>(sx + 2 * GAC_E));
        if (has_h) {
          hp = __ldcg(reinterpret_cast<const float4*>(g.hprev + (size_t)m * GAC_E + u));
          xh = __ldcg(reinterpret_cast<const float4*>(xh_cta + ((size_t)b * BM + prow) * G2_NU + ul));
        }
      };
      load_ops(0);
      G3_STAMP(e_ld)
      G3_GROUP_SYNC();                                             // this tile's bias rows are in smem
      mbar_wait(smem_u32(&bar_tfull[b]), tph);
      tc_fence_after();
      G3_STAMP(e_wt)
      const uint32_t d0 = tmem_d + (uint32_t)b * 256u + ((uint32_t)(quarter * 32) << 16);
      for (int q = 0; q < G2_NU / 16; ++q) {
        const int ul = q * 16 + 4 * pj, u = jt * G2_NU + ul;
        if (sub < 3) {
          float* srow = stg + (size_t)(quarter * 32 + lane) * G2_SS + sub * 16;
          const int rot = (lane >> 1) & 3;
          const float* bsub = b1_s + sub * G2_NU + q * 16;
          const bool add_b = sub < 2 || has_h;
          uint32_t v[16];
          if (probe & 2) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = 0x3f800000u;
          } else {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(d0 + (uint32_t)(sub * G2_NU + q * 16)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          }
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (add_b) bb = *reinterpret_cast<const float4*>(bsub + e);
            *reinterpret_cast<float4*>(srow + 4 * (((e >> 2) + rot) & 3)) =
                make_float4(fmaf(__uint_as_float(v[e]), inv_unit, bb.x), fmaf(__uint_as_float(v[e + 1]), inv_unit, bb.y),
                            fmaf(__uint_as_float(v[e + 2]), inv_unit, bb.z), fmaf(__uint_as_float(v[e + 3]), inv_unit, bb.w));
          }
        }
        G3_STAMP(e_p1)
        G3_GROUP_SYNC();
        G3_STAMP(e_b1)
        {
          const float* srow = stg + (size_t)prow * G2_SS + 4 * ((pj + (prow >> 1)) & 3);
          const float4 az = *reinterpret_cast<const float4*>(srow), ar = *reinterpret_cast<const float4*>(srow + 16),
                       ag = *reinterpret_cast<const float4*>(srow + 32);
          float4 bh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!has_h) bh = *reinterpret_cast<const float4*>(b1_s + 2 * G2_NU + ul);
          const float4 axh = has_h ? xh : ag;
          const float4 ahh = has_h ? ag : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 ho, zo, ro, co, mo;
          if (probe & 16) { ho = az; zo = ar; ro = ag; co = axh; mo = ahh; } else {
#define GRU3_LANE(f)                                                                              \
  {                                                                                               \
    const float hh = ahh.f + bh.f;                                                                \
    const float pa = 1.f + __expf(-fminf(fmaxf(sz.f + az.f, -30.f), 30.f));                       \
    const float pb = 1.f + __expf(-fminf(fmaxf(sr.f + ar.f, -30.f), 30.f));                       \
    const float R = __fdividef(1.f, pa * pb);                                                     \
    const float zz = pb * R, rr = pa * R;                                                         \
    const float cc = tnh((sh.f + axh.f) + rr * hh);                                               \
    ho.f = zz * hp.f + (1.f - zz) * cc; zo.f = zz; ro.f = rr; co.f = cc; mo.f = hh;               \
  }
          GRU3_LANE(x) GRU3_LANE(y) GRU3_LANE(z) GRU3_LANE(w)
#undef GRU3_LANE
          }
          const bool live = sid >= 0 && !(probe & 8);
          const size_t o = (size_t)m * GAC_E + u;
          if (live) {
            *reinterpret_cast<float4*>(g.hout + o) = ho;
            if (g.z) {
              *reinterpret_cast<float4*>(g.z + o) = zo; *reinterpret_cast<float4*>(g.r + o) = ro;
              *reinterpret_cast<float4*>(g.c + o) = co; *reinterpret_cast<float4*>(g.mhh + o) = mo;
            }
          }
          if (g.hout_pk) {
            uint32_t h0, l0, h1, l1;
            split_f16x2_u(ho.x * s_h, ho.y * s_h, h0, l0);
            split_f16x2_u(ho.z * s_h, ho.w * s_h, h1, l1);
            if (live) {
              unsigned char* dst = g.hout_pk + a16_off(m, u);
              *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
              *reinterpret_cast<uint2*>(dst + G2_A_TILE) = make_uint2(l0, l1);
            }
          }
          if (q + 1 < G2_NU / 16) load_ops(q + 1);
        }
        G3_STAMP(e_p2)
        G3_GROUP_SYNC();
        G3_STAMP(e_b2)
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(smem_u32(&bar_tempty[b]));
    }
#undef G3_STAMP
#undef G3_GROUP_SYNC
    if (stamp) {
      unsigned long long* d = dbgp + 8 + (size_t)blockIdx.x * 16;
      d[6] = (unsigned long long)e_wt; d[7] = (unsigned long long)e_wx; d[8] = (unsigned long long)e_p1; d[9] = (unsigned long long)e_b1;
      d[10] = (unsigned long long)e_p2; d[11] = (unsigned long long)e_b2; d[14] = (unsigned long long)e_ld;
    }
  }

  tc_fence_before();
  __syncthreads();
  // neither CTA may exit (or free TMEM) while the other can still signal its barriers
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == G2_MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512) : "memory");
  }
}

}  // namespace tc

// weights -> pair-kernel tiles (scales come from gru2_pack, which must have run on the same stream before)
inline cudaError_t gru3_pack(const float* kxa, const float* U, unsigned char* out, const tc::Gru2Scales* sc, cudaStream_t st) {
  const long long total = (long long)tc::G2_NT * 2 * tc::G2_KB * 3 * tc::G2_NU * 4;
  tc::gru3_pack_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(kxa, U, out, sc);
  return cudaGetLastError();
}
inline cudaError_t launch_gru_step_v3(const tc::Gru2Args& g, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc::gru_step_v3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::G3_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gru_step_v3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::G3_SMEM);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int pairs = tc::G2_NT * cdiv(cdiv(g.rows, tc::BM), 2);
  static const int grid_env = getenv("GAC_G2_GRID") ? atoi(getenv("GAC_G2_GRID")) : 0;
  const int cap = (grid_env > 0 && grid_env < gru2_grid() ? grid_env : gru2_grid()) & ~1;
  const int grid = 2 * pairs < cap ? 2 * pairs : cap;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(tc::G2_THREADS); cfg.dynamicSmemBytes = tc::G3_SMEM; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (g.dbg || g.probe) return cudaLaunchKernelEx(&cfg, tc::gru_step_v3_kernel<true>, g);
  return cudaLaunchKernelEx(&cfg, tc::gru_step_v3_kernel<false>, g);
}

}  // namespace gac
