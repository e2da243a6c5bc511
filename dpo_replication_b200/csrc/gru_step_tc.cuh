// One GRU step of the AIQN actor as ONE tcgen05 kernel (sm_100a): both input projections and the
// gate math, with nothing but h leaving the SM.
//
//   mx = [se | ae] * kx + gb0      (state half hoisted per unique state: sx = se * kx_s + gb0, given)
//   mh = h_prev * U + gb1
//   z = sigmoid(mx_z + mh_z), r = sigmoid(mx_r + mh_r), c = tanh(mx_h + r * mh_h), h = z h_prev + (1 - z) c
//   (Keras GRU, reset_after=True, gate order [z|r|h]: GAC/networks.py:166)
//
// The unfused path writes mxa = ae*kx_a and mh = h_prev*U (rows x 1200 each) to HBM and reads them
// back in a gates kernel: ~730 MB per 32768-row step, and the GEMM epilogues (157 MB of stores
// each, not overlapped with any MMA work) take ~45% of the GEMM time.  Here a CTA owns 128 rows x
// 64 hidden units and accumulates, over K = 400 (ae) + 400 (h_prev),
//     TMEM columns  [ xh | z | r | hh ]   (64 each; main accumulator, correction accumulator at +256)
// ae-blocks multiply [kx_h | kx_z | kx_r] into columns [0, 192), h-blocks multiply [U_z | U_r | U_h]
// into columns [64, 256): z and r receive both halves, the candidate gate keeps its two parts apart
// as reset_after needs.  Same 3xTF32 split and dual accumulators as gemm_tc.cuh.  The epilogue
// stages the 128 x 256 sums through the (now idle) pipeline smem so that global accesses are
// row-contiguous, applies the gates and writes h (+ z, r, c, mh_h when the backward needs them):
// 157 MB per 32768-row step.
#pragma once
#include "gemm_tc.cuh"

namespace gac {
namespace tc {

struct GruStepArgs {
  int rows;                    // M
  const int* dyn_rows;         // optional device row count: rows >= *dyn_rows are dead
  const float* ae;             // (rows, 400) action embedding
  const float* hprev;          // (rows, 400); nullptr = first step (h_prev = 0, U never touched)
  const unsigned char* wpk;    // gru_pack_kernel output
  const float* sx;             // (n_states, 1200) hoisted state half + input bias
  const int* sidx;             // row -> state
  const float* b1;             // recurrent bias (1200)
  float* hout;                 // (rows, 400)
  float *z, *r, *c, *mhh;      // optional saves (rows, 400)
  int probe;                   // timing experiments (env GAC_GRU_PROBE): results are garbage when set
  unsigned long long* dbg;     // optional per-CTA phase stamps (tools/gru_probe.py --timeline)
};

constexpr int GRU_NU = 64;                                      // hidden units per tile
constexpr int GRU_NT = (GAC_E + GRU_NU - 1) / GRU_NU;           // 7 (the last tile has 16 units)
constexpr int GRU_KB = (GAC_E + BK - 1) / BK;                   // 13 K blocks per operand half
constexpr uint32_t GRU_BSLOT = 2u * 3u * GRU_NU * 128u;         // [hi | lo] of a (3*64 x 32) tile: 48 KB
constexpr uint32_t GRU_STAGE = 2 * A_TILE_BYTES + GRU_BSLOT;    // 80 KB
constexpr int GRU_STAGES = 2;
constexpr int GRU_SMEM = GRU_STAGES * GRU_STAGE + 1024;
constexpr size_t GRU_PACK_BYTES = (size_t)GRU_NT * 2 * GRU_KB * GRU_BSLOT;   // 8.7 MB

// Pre-split, pre-swizzled weights.  Slot (tile j, block kb2): kb2 < 13 -> rows k of kx_a, gate order [h z r];
// kb2 >= 13 -> rows k of U, gate order [z r h].  Tile row r = gate * nu + unit.
__global__ void gru_pack_kernel(const float* __restrict__ kxa, const float* __restrict__ U, unsigned char* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_slot = 3 * GRU_NU * 8;
  if (idx >= (long long)GRU_NT * 2 * GRU_KB * per_slot) return;
  const int c = idx % 8, r = (idx / 8) % (3 * GRU_NU);
  const int slot = idx / per_slot, kb2 = slot % (2 * GRU_KB), j = slot / (2 * GRU_KB);
  const int nu = min(GRU_NU, GAC_E - j * GRU_NU);
  if (r >= 3 * nu) return;
  const int half = kb2 >= GRU_KB, kb = kb2 - half * GRU_KB;
  const int gi = r / nu, u = j * GRU_NU + r % nu;
  const int gate = half ? gi : (gi + 2) % 3;                     // [z r h] for U, [h z r] for kx_a (z = 0, r = 1, h = 2)
  const float* W = half ? U : kxa;
  float h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k = kb * BK + 4 * c + e;
    const float x = k < GAC_E ? W[(size_t)k * GAC_G + gate * GAC_E + u] : 0.f;
    split_tf32(x, h[e], l[e]);
  }
  unsigned char* base = out + (size_t)slot * GRU_BSLOT;
  const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
  *reinterpret_cast<float4*>(base + off) = make_float4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<float4*>(base + (size_t)3 * nu * 128 + off) = make_float4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(THREADS, 1) gru_step_tc_kernel(const GruStepArgs g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[GRU_STAGES], bar_empty[GRU_STAGES], bar_accum;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // 1-D grid: all 64-unit tiles first (row-block major, 6 per row block: neighbours share the A rows in L2), the
  // narrow 16-unit tiles last, where they fill the tail of the last wave
  const int nmb = (g.rows + BM - 1) / BM, nfull = nmb * (GRU_NT - 1);
  const int jt = (int)blockIdx.x < nfull ? (int)blockIdx.x % (GRU_NT - 1) : GRU_NT - 1;
  const int m0 = ((int)blockIdx.x < nfull ? (int)blockIdx.x / (GRU_NT - 1) : (int)blockIdx.x - nfull) * BM;
  const int m_end = g.dyn_rows ? min(g.rows, *g.dyn_rows) : g.rows;
  if (m0 >= m_end) return;                                       // dead tile (uniform per CTA)
  unsigned long long* dbg = (g.dbg && tid == 0) ? g.dbg + 8 + (size_t)blockIdx.x * 8 : nullptr;
  if (dbg) {
    unsigned long long gt; unsigned sm;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
    dbg[0] = gt; dbg[1] = clock64(); dbg[7] = sm;
    if (blockIdx.x == 0) g.dbg[0] = 0x600DC0DE600DC0DEull;
  }
  const int nu = min(GRU_NU, GAC_E - jt * GRU_NU);               // 64, or 16 for the last tile
  const int nkb = g.hprev ? 2 * GRU_KB : GRU_KB;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_lo_off = 3u * (uint32_t)nu * 128u;

  if (tid == 0) {
    for (int s = 0; s < GRU_STAGES; ++s) { mbar_init(smem_u32(&bar_full[s]), LOADER_WARPS / 2); mbar_init(smem_u32(&bar_empty[s]), 1); }
    mbar_init(smem_u32(&bar_accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == LOADER_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  if (dbg) dbg[2] = clock64();

  if (warp < LOADER_WARPS) {
    // ------------------------------ loaders: two groups alternate K blocks ------------------------------
    const int grp = warp >> 2, t = tid & (GROUP - 1);
    for (int kb = grp; kb < nkb; kb += 2) {
      const int half = kb >= GRU_KB, k0 = (kb - half * GRU_KB) * BK, live = min(BK, GAC_E - k0);
      float xa[8][4];
      if (g.probe & 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) xa[i][0] = xa[i][1] = xa[i][2] = xa[i][3] = 1.f;
      } else
      load_operand<false, true, 8>(xa, half ? g.hprev : g.ae, GAC_E, m0, m_end, BM, k0, live, t);
      float la[8][4];
      split_operand<8>(xa, la);     // waits for the loads, not for the stage
      const int s = kb % GRU_STAGES;
      const uint32_t ph = (uint32_t)(kb / GRU_STAGES) & 1u;
      mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
      const uint32_t sa = smem0 + (uint32_t)s * GRU_STAGE;
      if (t == 0 && !(g.probe & 4)) {
        const uint32_t bytes = 2u * b_lo_off;
        const unsigned char* srcp = g.wpk + ((size_t)jt * 2 * GRU_KB + kb) * GRU_BSLOT;
        const uint32_t bar = smem_u32(&bar_full[s]);
        asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa + 2 * A_TILE_BYTES),
                     "l"(srcp), "r"(bytes), "r"(bar)
                     : "memory");
      }
      if (!(g.probe & 8)) store_split<false, 8>(xa, la, sa, sa + A_TILE_BYTES, BM, t);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
    }
    // ------------------------------ epilogue ------------------------------
    if (dbg) dbg[3] = clock64();
    // Work split of phase 2: a thread owns one quad of hidden units and walks rows (rpi rows apart); consecutive
    // lanes = consecutive units of a row, so global traffic is row-contiguous.  Nothing it reads from global
    // memory depends on the accumulators, so the loads of the first batch of rows are issued BEFORE waiting for the
    // last MMAs, and those of batch b + 1 before the gate math of batch b (8 warps per SM cannot hide a dependent
    // sidx -> sx chain per item otherwise: measured 14.7k of a tile's 59k cycles before this).
    const int quarter = warp & 3, hf = warp >> 2;
    const int n4 = nu >> 2, rpi = (LOADER_WARPS * 32) / n4;       // rows per pass: 16 (nu = 64) or 64 (nu = 16)
    const int u4 = tid % n4, rsub = tid / n4;
    const int u = jt * GRU_NU + 4 * u4;
    constexpr int IB = 4;                                         // rows per batch
    const int nbatch = (BM + IB * rpi - 1) / (IB * rpi);          // 2 or 1
    struct Batch { int sid[IB]; float4 sz[IB], sr[IB], sh[IB], hp[IB]; };
    auto load_sid = [&](Batch& q, int r0) {
#pragma unroll
      for (int b = 0; b < IB; ++b) {
        const int row = r0 + b * rpi + rsub, m = m0 + row;
        q.sid[b] = (row < BM && m < m_end) ? g.sidx[m] : -1;
      }
    };
    auto load_rows = [&](Batch& q, int r0) {
#pragma unroll
      for (int b = 0; b < IB; ++b) {
        q.hp[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q.sid[b] < 0) continue;
        const float* sx = g.sx + (size_t)q.sid[b] * GAC_G + u;
        q.sz[b] = *reinterpret_cast<const float4*>(sx); q.sr[b] = *reinterpret_cast<const float4*>(sx + GAC_E);
        q.sh[b] = *reinterpret_cast<const float4*>(sx + 2 * GAC_E);
        if (g.hprev) q.hp[b] = *reinterpret_cast<const float4*>(g.hprev + (size_t)(m0 + r0 + b * rpi + rsub) * GAC_E + u);
      }
    };
    Batch q0, q1;
    load_sid(q0, 0);
    if (nbatch > 1) load_sid(q1, IB * rpi);
    const float4 bz = *reinterpret_cast<const float4*>(g.b1 + u), br = *reinterpret_cast<const float4*>(g.b1 + GAC_E + u),
                 bh = *reinterpret_cast<const float4*>(g.b1 + 2 * GAC_E + u);
    load_rows(q0, 0);

    mbar_wait(smem_u32(&bar_accum), 0);
    tc_fence_after();
    if (dbg) dbg[4] = clock64();
    // phase 1: main + correction -> smem [128][4 nu (+4 pad)] fp32.  All MMAs have retired, so the stages are free.
    const int ncols = g.hprev ? 4 * nu : 3 * nu;                 // hh columns were never written on the first step
    const int SS = 4 * nu + 4;                                   // row stride in floats (== 4 mod 32: conflict-free v4 stores)
    float* stg = reinterpret_cast<float*>(smem_raw + (smem0 - smem_u32(smem_raw)));
    {
      float* srow = stg + (size_t)(quarter * 32 + lane) * SS;
      for (int ci = hf; ci < ncols / 16; ci += 2) {
        float v[16], w[16];
        const uint32_t ta = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ci * 16);
        tmem_ld16(ta, v);
        tmem_ld16(ta + 256u, w);
#pragma unroll
        for (int q = 0; q < 16; q += 4)
          *reinterpret_cast<float4*>(srow + ci * 16 + q) = make_float4(v[q] + w[q], v[q + 1] + w[q + 1], v[q + 2] + w[q + 2], v[q + 3] + w[q + 3]);
      }
    }
    if (nbatch > 1) load_rows(q1, IB * rpi);
    asm volatile("bar.sync 1, 256;" ::: "memory");               // the 8 epilogue warps only
    if (dbg) dbg[5] = clock64();
    // phase 2: gate math.  sigmoid / tanh through ex2.approx + fast division: |error| ~ 2e-7, far inside the 1e-5 contract
    auto sig = [](float x) { return __fdividef(1.f, 1.f + __expf(-x)); };
    auto tnh = [](float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); };
    auto gates = [&](const Batch& q, int r0) {
#pragma unroll
      for (int b = 0; b < IB; ++b) {
        if (q.sid[b] < 0) continue;
        const int row = r0 + b * rpi + rsub;
        const float* srow = stg + (size_t)row * SS + 4 * u4;
        const float4 axh = *reinterpret_cast<const float4*>(srow);
        const float4 az = *reinterpret_cast<const float4*>(srow + nu);
        const float4 ar = *reinterpret_cast<const float4*>(srow + 2 * nu);
        float4 ahh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.hprev) ahh = *reinterpret_cast<const float4*>(srow + 3 * nu);
        float4 ho, zo, ro, co, mo;
#define GRU_LANE(f)                                                                               \
  {                                                                                               \
    const float hh = ahh.f + bh.f;                                                                \
    const float zz = sig((q.sz[b].f + az.f) + bz.f), rr = sig((q.sr[b].f + ar.f) + br.f);         \
    const float cc = tnh((q.sh[b].f + axh.f) + rr * hh);                                          \
    ho.f = zz * q.hp[b].f + (1.f - zz) * cc; zo.f = zz; ro.f = rr; co.f = cc; mo.f = hh;          \
  }
        GRU_LANE(x) GRU_LANE(y) GRU_LANE(z) GRU_LANE(w)
#undef GRU_LANE
        const size_t o = (size_t)(m0 + row) * GAC_E + u;
        *reinterpret_cast<float4*>(g.hout + o) = ho;
        if (g.z) {
          *reinterpret_cast<float4*>(g.z + o) = zo; *reinterpret_cast<float4*>(g.r + o) = ro;
          *reinterpret_cast<float4*>(g.c + o) = co; *reinterpret_cast<float4*>(g.mhh + o) = mo;
        }
      }
    };
    gates(q0, 0);
    if (nbatch > 1) gates(q1, IB * rpi);
    if (dbg) dbg[6] = clock64();
  } else {
    // ------------------------------ MMA issuer (one lane) ------------------------------
    if (lane == 0) {
      const uint32_t id3 = make_idesc(3 * nu), id2 = make_idesc(2 * nu), id1 = make_idesc(nu);
      const uint32_t corr = tmem_d + 256u;
      for (int kb = 0; kb < nkb; ++kb) {
        const int half = kb >= GRU_KB, k0 = (kb - half * GRU_KB) * BK, live = min(BK, GAC_E - k0);
        const int s = kb % GRU_STAGES;
        const uint32_t ph = (uint32_t)(kb / GRU_STAGES) & 1u;
        mbar_wait(smem_u32(&bar_full[s]), ph);
        tc_fence_after();
        const uint32_t sa = smem0 + (uint32_t)s * GRU_STAGE;
        const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_TILE_BYTES);
        const uint64_t b_hi = make_desc(sa + 2 * A_TILE_BYTES), b_lo = make_desc(sa + 2 * A_TILE_BYTES + b_lo_off);
        const uint32_t dcol = half ? (uint32_t)nu : 0u;          // ae -> [xh z r], h_prev -> [z r hh]
        const int ksteps = (live + 7) / 8;
        for (int j = 0; j < ksteps; ++j) {
          const uint64_t adv = (uint64_t)(j * 2);
          if (half && kb == GRU_KB && j == 0) {
            // first touch of the hh columns: z, r keep accumulating, hh starts from zero
            const uint64_t bo = (uint64_t)((2u * (uint32_t)nu * 128u) >> 4);
            mma_tf32(corr + dcol, a_lo + adv, b_hi + adv, id2, 1u);
            mma_tf32(corr + dcol, a_hi + adv, b_lo + adv, id2, 1u);
            mma_tf32(tmem_d + dcol, a_hi + adv, b_hi + adv, id2, 1u);
            mma_tf32(corr + 3u * nu, a_lo + adv, b_hi + bo + adv, id1, 0u);
            mma_tf32(corr + 3u * nu, a_hi + adv, b_lo + bo + adv, id1, 1u);
            mma_tf32(tmem_d + 3u * nu, a_hi + adv, b_hi + bo + adv, id1, 0u);
            continue;
          }
          const uint32_t acc = (kb | j) ? 1u : 0u;
          if (g.probe & 1) { mma_tf32(tmem_d + dcol, a_hi + adv, b_hi + adv, id3, acc); continue; }
          if (g.probe & 16) continue;
          mma_tf32(corr + dcol, a_lo + adv, b_hi + adv, id3, acc);
          mma_tf32(corr + dcol, a_hi + adv, b_lo + adv, id3, 1u);
          mma_tf32(tmem_d + dcol, a_hi + adv, b_hi + adv, id3, acc);
        }
        mma_commit(smem_u32(&bar_empty[s]));
      }
      mma_commit(smem_u32(&bar_accum));
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == LOADER_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512) : "memory");
  }
}

}  // namespace tc

inline cudaError_t gru_pack(const float* kxa, const float* U, unsigned char* out, cudaStream_t st) {
  const long long total = (long long)tc::GRU_NT * 2 * tc::GRU_KB * 3 * tc::GRU_NU * 8;
  tc::gru_pack_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(kxa, U, out);
  return cudaGetLastError();
}

inline cudaError_t launch_gru_step(const tc::GruStepArgs& g, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc::gru_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GRU_SMEM);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  static const int probe = getenv("GAC_GRU_PROBE") ? atoi(getenv("GAC_GRU_PROBE")) : 0;
  tc::GruStepArgs a = g;
  a.probe = probe;
  dim3 grid(tc::GRU_NT * cdiv(g.rows, tc::BM));
  tc::gru_step_tc_kernel<<<grid, tc::THREADS, tc::GRU_SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gac
