// GRU step of the AIQN actor, second generation (sm_100a): a PERSISTENT, warp-specialised tcgen05 kernel whose
// epilogue overlaps the next tile's main loop.
//
//   same arithmetic as gru_step_tc.cuh (Keras GRU cell, reset_after, gate order [z|r|h]; GAC/networks.py:166,217,259)
//
// What round 1's kernel left on the table (profiles/r1b_gru_step_summary.txt: 21.1k of a tile's 31.8k cycles in the MMA
// loop, the rest serial): the A operand (action embedding, h_prev) was staged through registers by the same eight warps that
// later ran the epilogue, so nothing could overlap, and TMEM was full (main + correction accumulator = 512 columns).  Here:
//   * BOTH operands arrive by cp.async.bulk.  The producers of the activations -- the embedding kernel for `ae`, this
//     kernel's own epilogue for `h` -- write, next to (or instead of) their fp32 output, the pre-split, pre-scaled,
//     pre-swizzled fp16 hi / lo tiles the MMA consumes ("a16" images: [row block][K block][hi | lo], 128 rows x 64 bytes,
//     K-major SWIZZLE_64B).  One producer lane issues two bulk copies per K block; no thread touches operand data.
//   * ONE accumulator per gate: hi*hi + lo*hi + hi*lo all add into the same TMEM columns with lo = x' - hi kept UNSCALED
//     (operands are scaled to ~2^12 so lo stays a normal fp16 for every element within 2^-10 of the largest).  A tile is
//     128 rows x 64 units x [xh | z | r | hh] = 256 columns, so TMEM holds TWO tiles: the MMA lane fills one while the eight
//     epilogue warps drain the other.  The cost is three accumulate-truncations per K step instead of one (measured
//     relative shrink ~1e-8 per add, 150 adds per output: ~1e-6 against the 1e-5 contract; the parity tests are the judge).
//   * 32-deep K blocks (64-byte rows): a stage is 40 KB, four stages fit beside the epilogue's staging buffer, so three
//     stages of operand data are in flight while one is consumed -- bulk-copy latency is off the critical path.
//   * persistent grid (one CTA per SM), static round-robin over (row block, unit tile) items: consecutive CTAs share a row
//     block (its A tiles are L2-hot), every CTA sees all seven unit tiles including the narrow one.
#pragma once
#include "gru_step_tc.cuh"

namespace gac {
namespace tc {

constexpr int G2_BK = 32;                                        // fp16 K elements per block = one 64-byte swizzle row
constexpr int G2_KB = (GAC_E + G2_BK - 1) / G2_BK;               // 13 K blocks per operand half (the last has 16 live columns)
constexpr uint32_t G2_A_TILE = (uint32_t)BM * 64u;               // 8 KB per hi / lo
constexpr uint32_t G2_A_SLOT = 2u * G2_A_TILE;                   // [hi | lo] of one (row block, K block)
constexpr uint32_t G2_B_TILE = 3u * GRU_NU * 64u;                // 12 KB per hi / lo of a full 64-unit tile
constexpr uint32_t G2_B_SLOT = 2u * G2_B_TILE;
constexpr uint32_t G2_STAGE = G2_A_SLOT + G2_B_SLOT;             // 40 KB
constexpr int G2_STAGES = 4;
constexpr int G2_SS = 68;                                        // staging row stride in floats (68 = 4 mod 32: conflict-free v4 row stores)
constexpr int G2_SMEM = G2_STAGES * (int)G2_STAGE + BM * G2_SS * 4 + 1024;
constexpr int G2_EPI_WARPS = 8;
constexpr int G2_THREADS = (G2_EPI_WARPS + 2) * 32;              // 8 epilogue warps, the MMA warp, the producer warp
constexpr size_t G2_WPK_BYTES = (size_t)GRU_NT * 2 * G2_KB * G2_B_SLOT;   // 4.4 MB
inline size_t a16_bytes(long long rows) { return (size_t)((rows + BM - 1) / BM) * G2_KB * G2_A_SLOT; }

struct Gru2Scales {
  float s_ae, s_h, s_kx, s_u, inv_unit;    // x' = x * s; accumulators hold unit = s_ae * s_kx = s_h * s_u
  int ok;                                  // 0: no valid scaling, the caller's launch runs round 1's kernel instead
  unsigned int slots[3];                   // scan scratch: max |kx_a|, max |U|, bound of |ae| (float bit patterns)
  int pad[3];
};

struct Gru2Args {
  int rows;                      // M
  const int* dyn_rows;           // optional device row count
  const unsigned char* ae_pk;    // a16 image of the action embedding, scaled by s_ae
  const unsigned char* h_pk;     // a16 image of h_prev, scaled by s_h; nullptr = first step (h_prev = 0)
  const float* hprev;            // (rows, 400) fp32 h_prev for the state update; nullptr on the first step
  const unsigned char* wpk;      // gru2_pack_kernel tiles
  const Gru2Scales* sc;
  const float* sx;               // (n_states, 1200) hoisted state half + input bias
  const int* sidx;               // row -> state
  const float* b1;               // recurrent bias (1200)
  float* hout;                   // (rows, 400) fp32
  unsigned char* hout_pk;        // optional a16 image of h_out (the next step's A operand)
  float *z, *r, *c, *mhh;        // optional saves (rows, 400)
  unsigned long long* dbg;       // optional per-CTA counters: [cycles, MMA-lane wait for operands, wait for a free accumulator, tiles]
};

// K-major SWIZZLE_64B shared-memory matrix descriptor (cute: Swizzle<2,4,3> o ((8,n),2):((4,SBO),1) in 16-byte units):
// rows of 64 bytes, 8-row groups SBO = 512 bytes apart, 16-byte chunk index XOR (row >> 1) & 3; layout type 4.
__device__ __forceinline__ uint64_t make_desc64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ uint32_t sw64_off(int r, int c) { return (uint32_t)r * 64u + (uint32_t)((c ^ ((r >> 1) & 3)) << 4); }
// two values -> fp16 hi and UNSCALED fp16 lo = x - hi
__device__ __forceinline__ void split_f16x2_u(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// byte offset of the 8-byte group holding columns [col, col + 4) of row m inside an a16 image (hi part; lo is G2_A_TILE further)
__device__ __forceinline__ size_t a16_off(int m, int col) {
  const int rb = m >> 7, r = m & 127, kb = col >> 5, e = col & 31;
  return ((size_t)rb * G2_KB + kb) * G2_A_SLOT + sw64_off(r, e >> 3) + (uint32_t)((e >> 2) & 1) * 8u;
}

// scales: every operand is brought to a maximum of ~2^12 (fp16 tops out at 2^16; lo of an element 2^-10 below the maximum is
// still a normal fp16).  slots: [max |kx_a|, max |U|, bound of |ae|] from gru_amax_kernel.
__global__ void gru2_scales_kernel(Gru2Scales* __restrict__ sc) {
  const float mk = __uint_as_float(sc->slots[0]), mu = __uint_as_float(sc->slots[1]), mb = __uint_as_float(sc->slots[2]);
  sc->ok = 0; sc->s_ae = sc->s_h = sc->s_kx = sc->s_u = sc->inv_unit = 1.f;
  if (!(mk > 0.f && mu > 0.f && mb > 0.f && mk < 1e30f && mu < 1e30f && mb < 1e30f)) return;
  int ek, eu, eb;
  frexpf(mk, &ek); frexpf(mu, &eu); frexpf(mb, &eb);               // x = f * 2^e, f in [0.5, 1)
  const int lh = 12, lu = 12 - eu;                                 // |h| < 1 -> 2^12;  |U'| < 2^12
  const int unit = lh + lu;
  int lkx = 12 - ek, lae = unit - lkx;                             // |kx'| < 2^12; s_ae * s_kx == unit
  if (eb + lae > 15) { const int d = eb + lae - 15; lae -= d; lkx += d; }     // keep the scaled embedding bound below 2^15
  if (eb + lae < 8) { const int d = 8 - (eb + lae); lae += d; lkx -= d; }      // ... and its lo terms out of the subnormals
  const int kx_top = ek + lkx;
  if (kx_top > 15 || kx_top < 4 || abs(lae) > 100 || abs(lkx) > 100 || abs(lu) > 100) return;
  sc->s_ae = ldexpf(1.f, lae); sc->s_h = ldexpf(1.f, lh); sc->s_kx = ldexpf(1.f, lkx); sc->s_u = ldexpf(1.f, lu);
  sc->inv_unit = ldexpf(1.f, -unit);
  sc->ok = 1;
}

// Pre-split, pre-swizzled weights.  Slot (tile j, block kb2): kb2 < KB -> rows k of kx_a, gate order [h z r]; kb2 >= KB ->
// rows k of U, gate order [z r h].  Tile row r = gate * nu + unit; 64-byte rows (32 k).  One thread per 16-byte chunk.
__global__ void gru2_pack_kernel(const float* __restrict__ kxa, const float* __restrict__ U, unsigned char* __restrict__ out,
                                 const Gru2Scales* __restrict__ sc) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_slot = 3 * GRU_NU * 4;
  if (idx >= (long long)GRU_NT * 2 * G2_KB * per_slot || !sc->ok) return;
  const int c = idx % 4, r = (idx / 4) % (3 * GRU_NU);
  const int slot = idx / per_slot, kb2 = slot % (2 * G2_KB), j = slot / (2 * G2_KB);
  const int nu = min(GRU_NU, GAC_E - j * GRU_NU);
  if (r >= 3 * nu) return;
  const int half = kb2 >= G2_KB, kb = kb2 - half * G2_KB;
  const int gi = r / nu, u = j * GRU_NU + r % nu;
  const int gate = half ? gi : (gi + 2) % 3;                       // [z r h] for U, [h z r] for kx_a (z = 0, r = 1, h = 2)
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
  unsigned char* base = out + (size_t)slot * G2_B_SLOT;
  const uint32_t off = sw64_off(r, c);
  *reinterpret_cast<uint4*>(base + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(base + (size_t)3 * nu * 64 + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// fp32 (rows, 400) -> a16 image, for activations whose producer does not write the image itself.  One thread per 16-byte
// chunk, ordered (K block, row, chunk) so that a warp writes 512 contiguous bytes and reads eight 128-byte row segments.
__global__ void a16_pack_kernel(const float* __restrict__ X, int ldx, int rows, const int* __restrict__ dyn_rows, const float* __restrict__ scale,
                                unsigned char* __restrict__ out) {
  const int m_end = dyn_rows ? min(rows, *dyn_rows) : rows;
  const int n_rb = (m_end + BM - 1) / BM;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_rb * G2_KB * BM * 4) return;
  const int c = idx & 3, r = (idx >> 2) & 127;
  const int t = idx >> 9, kb = t % G2_KB, rb = t / G2_KB;
  const int m = rb * BM + r, k0 = kb * G2_BK + 8 * c;
  const float s = *scale;
  uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
  if (m < m_end && k0 < GAC_E) {                                    // 400 = 50 x 8: a chunk is entirely inside or outside
    const float4 a = *reinterpret_cast<const float4*>(X + (size_t)m * ldx + k0), b = *reinterpret_cast<const float4*>(X + (size_t)m * ldx + k0 + 4);
    split_f16x2_u(a.x * s, a.y * s, h[0], l[0]); split_f16x2_u(a.z * s, a.w * s, h[1], l[1]);
    split_f16x2_u(b.x * s, b.y * s, h[2], l[2]); split_f16x2_u(b.z * s, b.w * s, h[3], l[3]);
  }
  unsigned char* base = out + ((size_t)rb * G2_KB + kb) * G2_A_SLOT + sw64_off(r, c);
  *reinterpret_cast<uint4*>(base) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(base + G2_A_TILE) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void mbar_expect_tx_arrive(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(G2_THREADS, 1) gru_step_v2_kernel(const Gru2Args g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[G2_STAGES], bar_empty[G2_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_slot;

  if (!g.sc->ok) return;                                           // no valid scaling: gru_step_tc_kernel (launched next) does the step
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_end = g.dyn_rows ? min(g.rows, *g.dyn_rows) : g.rows;
  const int n_items = ((m_end + BM - 1) / BM) * GRU_NT;           // (row block, unit tile), row-block major
  const int nkb = g.h_pk ? 2 * G2_KB : G2_KB;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  float* stg = reinterpret_cast<float*>(smem_raw + (smem0 - smem_u32(smem_raw)) + G2_STAGES * G2_STAGE);

  if (tid == 0) {
    for (int s = 0; s < G2_STAGES; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&bar_tfull[b]), 1); mbar_init(smem_u32(&bar_tempty[b]), G2_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == G2_EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (warp == G2_EPI_WARPS + 1) {
    // ------------------------------ producer: one lane, two bulk copies per K block ------------------------------
    if (lane == 0) {
      int it = 0;
      for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
        const int rb = i / GRU_NT, jt = i - rb * GRU_NT;
        const int nu = min(GRU_NU, GAC_E - jt * GRU_NU);
        const uint32_t b_bytes = 2u * 3u * (uint32_t)nu * 64u;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % G2_STAGES;
          const uint32_t ph = (uint32_t)(it / G2_STAGES) & 1u;
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
          const uint32_t sa = smem0 + (uint32_t)s * G2_STAGE, bar = smem_u32(&bar_full[s]);
          const int half = kb >= G2_KB, kbl = kb - half * G2_KB;
          const unsigned char* asrc = (half ? g.h_pk : g.ae_pk) + ((size_t)rb * G2_KB + kbl) * G2_A_SLOT;
          const unsigned char* bsrc = g.wpk + ((size_t)jt * 2 * G2_KB + kb) * G2_B_SLOT;
          mbar_expect_tx_arrive(bar, G2_A_SLOT + b_bytes);
          bulk_g2s(sa, asrc, G2_A_SLOT, bar);
          bulk_g2s(sa + G2_A_SLOT, bsrc, b_bytes, bar);
        }
      }
    }
    __syncwarp();
  } else if (warp == G2_EPI_WARPS) {
    // ------------------------------ MMA issuer: one lane ------------------------------
    if (lane == 0) {
      long long t_wait_full = 0, t_wait_acc = 0;
      const long long t_begin = g.dbg ? clock64() : 0;
      int it = 0, t = 0;
      for (int i = blockIdx.x; i < n_items; i += gridDim.x, ++t) {
        const int jt = i % GRU_NT;
        const int nu = min(GRU_NU, GAC_E - jt * GRU_NU);
        const int b = t & 1;
        const uint32_t tph = (uint32_t)(t >> 1) & 1u;
        long long c0 = g.dbg ? clock64() : 0;
        mbar_wait(smem_u32(&bar_tempty[b]), tph ^ 1u);             // the epilogue has drained this accumulator (two tiles ago)
        tc_fence_after();
        if (g.dbg) t_wait_acc += clock64() - c0;
        const uint32_t d0 = tmem_d + (uint32_t)b * 256u;
        const uint32_t id3 = make_idesc_f16(3 * nu), id2 = make_idesc_f16(2 * nu), id1 = make_idesc_f16(nu);
        const uint32_t b_lo_off = 3u * (uint32_t)nu * 64u;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % G2_STAGES;
          const uint32_t ph = (uint32_t)(it / G2_STAGES) & 1u;
          c0 = g.dbg ? clock64() : 0;
          mbar_wait(smem_u32(&bar_full[s]), ph);
          tc_fence_after();
          if (g.dbg) t_wait_full += clock64() - c0;
          const uint32_t sa = smem0 + (uint32_t)s * G2_STAGE;
          const uint64_t a_hi = make_desc64(sa), a_lo = make_desc64(sa + G2_A_TILE);
          const uint64_t b_hi = make_desc64(sa + G2_A_SLOT), b_lo = make_desc64(sa + G2_A_SLOT + b_lo_off);
          const int half = kb >= G2_KB, kbl = kb - half * G2_KB, live = min(G2_BK, GAC_E - kbl * G2_BK);
          const uint32_t dcol = half ? (uint32_t)nu : 0u;          // ae -> [xh z r], h_prev -> [z r hh]
          const int ksteps = (live + 15) / 16;
          for (int j = 0; j < ksteps; ++j) {
            const uint64_t adv = (uint64_t)(j * 2);
            if (half && kbl == 0 && j == 0) {
              // first touch of the hh columns: z, r keep accumulating, hh starts from zero (stale from two tiles ago)
              const uint64_t bo = (uint64_t)((2u * (uint32_t)nu * 64u) >> 4);
              mma_f16(d0 + dcol, a_hi + adv, b_hi + adv, id2, 1u);
              mma_f16(d0 + dcol, a_lo + adv, b_hi + adv, id2, 1u);
              mma_f16(d0 + dcol, a_hi + adv, b_lo + adv, id2, 1u);
              mma_f16(d0 + 3u * nu, a_hi + adv, b_hi + bo + adv, id1, 0u);
              mma_f16(d0 + 3u * nu, a_lo + adv, b_hi + bo + adv, id1, 1u);
              mma_f16(d0 + 3u * nu, a_hi + adv, b_lo + bo + adv, id1, 1u);
              continue;
            }
            mma_f16(d0 + dcol, a_hi + adv, b_hi + adv, id3, (kb | j) ? 1u : 0u);
            mma_f16(d0 + dcol, a_lo + adv, b_hi + adv, id3, 1u);
            mma_f16(d0 + dcol, a_hi + adv, b_lo + adv, id3, 1u);
          }
          mma_commit(smem_u32(&bar_empty[s]));
        }
        mma_commit(smem_u32(&bar_tfull[b]));
      }
      if (g.dbg) {
        unsigned long long* d = g.dbg + 8 + (size_t)blockIdx.x * 4;
        d[0] = (unsigned long long)(clock64() - t_begin); d[1] = (unsigned long long)t_wait_full; d[2] = (unsigned long long)t_wait_acc;
        d[3] = (unsigned long long)t;
        if (blockIdx.x == 0) { g.dbg[0] = 0x600DC0DE600DC0E2ull; g.dbg[1] = gridDim.x; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue: eight warps, overlapped with the next tile's main loop ------------------------------
    const int quarter = warp & 3, hf = warp >> 2;
    const float inv_unit = g.sc->inv_unit, s_h = g.sc->s_h;
    const bool has_h = g.hprev != nullptr;
    auto sig = [](float x) { return __fdividef(1.f, 1.f + __expf(-x)); };
    auto tnh = [](float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); };
    // phase-2 ownership: lane -> row (lane & 7) of the warp's 8-row strip (+ 64 in the second pass), unit quad lane >> 3:
    // a quarter warp reads eight consecutive staging rows (every bank once) and a warp touches 8 rows x 64 contiguous bytes
    const int prow = warp * 8 + (lane & 7), pj = lane >> 3;
    int t = 0;
    for (int i = blockIdx.x; i < n_items; i += gridDim.x, ++t) {
      const int rb = i / GRU_NT, jt = i - rb * GRU_NT;
      const int nu = min(GRU_NU, GAC_E - jt * GRU_NU), nq = nu >> 4;
      const int m0 = rb * BM, b = t & 1;
      const uint32_t tph = (uint32_t)(t >> 1) & 1u;
      int sid[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) { const int m = m0 + prow + 64 * p; sid[p] = m < m_end ? g.sidx[m] : -1; }
      mbar_wait(smem_u32(&bar_tfull[b]), tph);
      tc_fence_after();
      const uint32_t d0 = tmem_d + (uint32_t)b * 256u + ((uint32_t)(quarter * 32) << 16);
      for (int q = 0; q < nq; ++q) {
        const int u = jt * GRU_NU + q * 16 + 4 * pj;               // this lane's four hidden units
        // operands that do not depend on the accumulators: in flight across phase 1 and the barrier
        float4 sz[2], sr[2], sh[2], hp[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          hp[p] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (sid[p] < 0) continue;
          const float* sx = g.sx + (size_t)sid[p] * GAC_G + u;
          sz[p] = *reinterpret_cast<const float4*>(sx); sr[p] = *reinterpret_cast<const float4*>(sx + GAC_E);
          sh[p] = *reinterpret_cast<const float4*>(sx + 2 * GAC_E);
          if (has_h) hp[p] = *reinterpret_cast<const float4*>(g.hprev + (size_t)(m0 + prow + 64 * p) * GAC_E + u);
        }
        const float4 bz = *reinterpret_cast<const float4*>(g.b1 + u), br = *reinterpret_cast<const float4*>(g.b1 + GAC_E + u),
                     bh = *reinterpret_cast<const float4*>(g.b1 + 2 * GAC_E + u);
        // phase 1: 16 units x [xh z r hh] of 128 rows, TMEM -> staging; warps 0-3 take xh and z, warps 4-7 r and hh
        {
          float* srow = stg + (size_t)(quarter * 32 + lane) * G2_SS;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int gi = 2 * hf + k;
            float v[16];
            if (gi == 3 && !has_h) {                               // the hh columns were never written on the first step
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = 0.f;
            } else {
              tmem_ld16(d0 + (uint32_t)(gi * nu + q * 16), v);
            }
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              *reinterpret_cast<float4*>(srow + gi * 16 + e) = make_float4(v[e] * inv_unit, v[e + 1] * inv_unit, v[e + 2] * inv_unit, v[e + 3] * inv_unit);
          }
        }
        if (q == nq - 1) {                                         // last TMEM read of this tile: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[b]));
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // phase 2: gate math and the state update, row-contiguous global traffic
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          if (sid[p] < 0) continue;
          const int row = prow + 64 * p;
          const float* srow = stg + (size_t)row * G2_SS + 4 * pj;
          const float4 axh = *reinterpret_cast<const float4*>(srow), az = *reinterpret_cast<const float4*>(srow + 16),
                       ar = *reinterpret_cast<const float4*>(srow + 32), ahh = *reinterpret_cast<const float4*>(srow + 48);
          float4 ho, zo, ro, co, mo;
#define GRU2_LANE(f)                                                                              \
  {                                                                                               \
    const float hh = ahh.f + bh.f;                                                                \
    const float zz = sig((sz[p].f + az.f) + bz.f), rr = sig((sr[p].f + ar.f) + br.f);             \
    const float cc = tnh((sh[p].f + axh.f) + rr * hh);                                            \
    ho.f = zz * hp[p].f + (1.f - zz) * cc; zo.f = zz; ro.f = rr; co.f = cc; mo.f = hh;            \
  }
          GRU2_LANE(x) GRU2_LANE(y) GRU2_LANE(z) GRU2_LANE(w)
#undef GRU2_LANE
          const int m = m0 + row;
          const size_t o = (size_t)m * GAC_E + u;
          *reinterpret_cast<float4*>(g.hout + o) = ho;
          if (g.z) {
            *reinterpret_cast<float4*>(g.z + o) = zo; *reinterpret_cast<float4*>(g.r + o) = ro;
            *reinterpret_cast<float4*>(g.c + o) = co; *reinterpret_cast<float4*>(g.mhh + o) = mo;
          }
          if (g.hout_pk) {                                         // the next step's A operand, already split / scaled / swizzled
            uint32_t h0, l0, h1, l1;
            split_f16x2_u(ho.x * s_h, ho.y * s_h, h0, l0);
            split_f16x2_u(ho.z * s_h, ho.w * s_h, h1, l1);
            unsigned char* dst = g.hout_pk + a16_off(m, u);
            *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(dst + G2_A_TILE) = make_uint2(l0, l1);
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");             // the staging rows are rewritten by the next chunk / tile
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == G2_EPI_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512) : "memory");
  }
}

}  // namespace tc

// weights -> scales + packed tiles (the scan reuses round 1's gru_amax_kernel); sc->ok == 0 afterwards means "use gru_step_tc_kernel"
inline cudaError_t gru2_pack(const float* kxa, const float* U, const float* wa, const float* ba, unsigned char* out, tc::Gru2Scales* sc,
                             cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(sc, 0, sizeof(tc::Gru2Scales), st);
  if (e != cudaSuccess) return e;
  tc::gru_amax_kernel<<<60, 256, 0, st>>>(kxa, U, wa, ba, sc->slots);
  tc::gru2_scales_kernel<<<1, 1, 0, st>>>(sc);
  const long long total = (long long)tc::GRU_NT * 2 * tc::G2_KB * 3 * tc::GRU_NU * 4;
  tc::gru2_pack_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(kxa, U, out, sc);
  return cudaGetLastError();
}
inline cudaError_t a16_pack(const float* X, int ldx, int rows, const int* dyn_rows, const float* scale, unsigned char* out, cudaStream_t st) {
  const long long total = (long long)cdiv(rows, tc::BM) * tc::G2_KB * tc::BM * 4;
  tc::a16_pack_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(X, ldx, rows, dyn_rows, scale, out);
  return cudaGetLastError();
}
inline int gru2_grid() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}
inline cudaError_t launch_gru_step_v2(const tc::Gru2Args& g, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc::gru_step_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::G2_SMEM);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int items = tc::GRU_NT * cdiv(g.rows, tc::BM);
  const int grid = items < gru2_grid() ? items : gru2_grid();
  tc::gru_step_v2_kernel<<<grid, tc::G2_THREADS, tc::G2_SMEM, st>>>(g);
  return cudaGetLastError();
}

}  // namespace gac
