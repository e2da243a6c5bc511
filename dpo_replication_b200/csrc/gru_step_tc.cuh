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
//
// Two arithmetic variants of the same kernel (template F16):
//   3xTF32 (F16 = false): hi = rna_tf32(x), lo = x - hi, kind::tf32, 8 K-elements per MMA.
//   3xFP16 (F16 = true) : x' = x * 2^e, hi = rn_fp16(x'), lo' = rn_fp16((x' - hi) * 2^11), kind::f16, 16 K-elements
//       per MMA -- the same 11 + 11 significand bits and the same three passes, but HALF the tcgen05.mma
//       instructions, and this kernel's loop is bound by MMA issue (round 1 measured ~128 cycles per M = 128 instruction
//       "whatever N and kind": 26 K-blocks x 12 MMAs = 40k of a tile's 52k cycles; round 2 found that to be the issuing
//       lane's per-instruction overhead, not the tensor core's N / 2 cycles -- gru_step_v2.cuh).  fp16 has 5 exponent bits, so every
//       operand needs a known magnitude bound; this kernel is the one place where all four have one WITHOUT any
//       activation statistics: |h_prev| < 1 (GRU state), |ae| <= max_j (sum_i |wa_ij| + |ba_j|) (LeakyReLU of a
//       cosine embedding, |cos| <= 1), and the two weight matrices are scanned when they are packed.
//       gru_scales_kernel picks power-of-two scales (exact) with s_ae * s_kx == s_h * s_U so that both halves
//       accumulate in one unit; values below 2^-14 of the scaled range land in fp16 subnormals whose absolute
//       error (2^-36 of the range after the lo term) is far below the 1e-5 contract.  If no valid scaling exists
//       (non-finite or absurd weights) the device flag `ok` stays 0, this variant exits at once and the 3xTF32
//       variant, in the same launch does the work (the kernel branches on the flag).
#pragma once
#include "gemm_tc.cuh"

namespace gac {
namespace tc {

struct GruStepArgs {
  int rows;                    // M
  const int* dyn_rows;         // optional device row count: rows >= *dyn_rows are dead
  const float* ae;             // (rows, 400) action embedding
  const float* hprev;          // (rows, 400); nullptr = first step (h_prev = 0, U never touched)
  const unsigned char* wpk;    // gru_pack_kernel<false> output (3xTF32 tiles)
  const unsigned char* wpk16;  // gru_pack_kernel<true> output (3xFP16 tiles), used when sc->ok
  const float* sx;             // (n_states, 1200) hoisted state half + input bias
  const int* sidx;             // row -> state
  const float* b1;             // recurrent bias (1200)
  float* hout;                 // (rows, 400)
  float *z, *r, *c, *mhh;      // optional saves (rows, 400)
  unsigned long long* dbg;     // optional per-CTA phase stamps (tools/gru_probe.py --timeline)
  const struct GruScales* sc;  // 3xFP16 scales (nullptr: 3xTF32 only)
  const int* skip_if;          // optional device flag: non-zero = gru_step_v2_kernel did this step, exit at once
};

struct GruScales {
  float s_ae, s_kx, s_u, inv_unit;   // x' = x * s; accumulators hold unit = s_ae*s_kx = s_u (s_h = 1); inv_unit = 1 / unit
  int ok;                            // 1: the 3xFP16 variant runs, 0: the 3xTF32 variant runs
  int pad[3];
};

constexpr int GRU_NU = 64;                                      // hidden units per tile
constexpr int GRU_NT = (GAC_E + GRU_NU - 1) / GRU_NU;           // 7 (the last tile has 16 units)
constexpr int GRU_KB = (GAC_E + BK - 1) / BK;                   // 13 K blocks per operand half
constexpr uint32_t GRU_BSLOT = 2u * 3u * GRU_NU * 128u;         // [hi | lo] of a (3*64 x 32) tile: 48 KB
constexpr uint32_t GRU_STAGE = 2 * A_TILE_BYTES + GRU_BSLOT;    // 80 KB
constexpr int GRU_STAGES = 2;
constexpr int GRU_SMEM = GRU_STAGES * GRU_STAGE + 1024;
constexpr size_t GRU_PACK_BYTES = (size_t)GRU_NT * 2 * GRU_KB * GRU_BSLOT;   // 8.7 MB
constexpr int BK16 = 64;                                        // fp16 elements per 128-byte swizzle row
constexpr int GRU_KB16 = (GAC_E + BK16 - 1) / BK16;             // 7 K blocks per operand half
constexpr size_t GRU_PACK16_BYTES = (size_t)GRU_NT * 2 * GRU_KB16 * GRU_BSLOT;   // 4.7 MB

// Power-of-two scales of the 3xFP16 variant.  Pass 1 (many blocks): maxima of |kx_a|, |U| and of the action-embedding
// bound into three uint slots (non-negative floats order like their bit patterns; the slots are zeroed before);
// any non-finite value poisons the bound with +inf.  Pass 2 (one thread): the scales.
__global__ void gru_amax_kernel(const float* __restrict__ kxa, const float* __restrict__ U, const float* __restrict__ wa,
                                const float* __restrict__ ba, unsigned int* __restrict__ slots) {
  float mk = 0.f, mu = 0.f, mb = 0.f;
  bool bad = false;
  const int n4 = GAC_E * GAC_G / 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(kxa)[i], b = reinterpret_cast<const float4*>(U)[i];
    const float ma = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)));
    const float mbb = fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)));
    bad |= !(fabsf(a.x) + fabsf(a.y) + fabsf(a.z) + fabsf(a.w) <= 3e38f) || !(fabsf(b.x) + fabsf(b.y) + fabsf(b.z) + fabsf(b.w) <= 3e38f);
    mk = fmaxf(mk, ma); mu = fmaxf(mu, mbb);
  }
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < GAC_E; j += gridDim.x * blockDim.x) {
    float sacc = fabsf(ba[j]);                                  // |ae_j| <= sum_i |wa_ij| + |ba_j|  (|cos| <= 1, LeakyReLU is 1-Lipschitz)
    for (int i = 0; i < GAC_NB; ++i) sacc += fabsf(wa[(size_t)i * GAC_E + j]);
    bad |= !(sacc <= 3e38f);
    mb = fmaxf(mb, sacc);
  }
  if (bad) mb = __int_as_float(0x7f800000);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mk = fmaxf(mk, __shfl_xor_sync(0xffffffffu, mk, o)); mu = fmaxf(mu, __shfl_xor_sync(0xffffffffu, mu, o));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
  }
  __shared__ unsigned int sh_max[3];                           // one global atomic per block and slot
  if (threadIdx.x < 3) sh_max[threadIdx.x] = 0u;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&sh_max[0], __float_as_uint(mk)); atomicMax(&sh_max[1], __float_as_uint(mu)); atomicMax(&sh_max[2], __float_as_uint(mb));
  }
  __syncthreads();
  if (threadIdx.x < 3 && sh_max[threadIdx.x]) atomicMax(slots + threadIdx.x, sh_max[threadIdx.x]);
}
__global__ void gru_scales_kernel(const unsigned int* __restrict__ slots, GruScales* __restrict__ out) {
  const float mk = __uint_as_float(slots[0]), mu = __uint_as_float(slots[1]), mb = __uint_as_float(slots[2]);
  GruScales r;
  r.ok = 0; r.s_ae = r.s_kx = r.s_u = r.inv_unit = 1.f; r.pad[0] = r.pad[1] = r.pad[2] = 0;
  if (mk > 0.f && mu > 0.f && mb > 0.f && mk < 1e30f && mu < 1e30f && mb < 1e30f) {
    int ek, eu, eb;
    frexpf(mk, &ek); frexpf(mu, &eu); frexpf(mb, &eb);         // x = f * 2^e, f in [0.5, 1)
    // U' = U * 2^(1-eu) in [1,2); unit = s_u.  kx' in [1,2) wants s_kx = 2^(1-ek), which fixes s_ae = 2^(ek-eu);
    // bound the scaled ae by 2^14 (fp16 max is 2^16): shift the excess onto kx.
    int ea = ek - eu;                                          // log2 s_ae
    if (eb + ea > 14) ea = 14 - eb;
    if (eb + ea < -8) ea = -8 - eb;
    const int esk = (1 - eu) - ea;                             // log2 s_kx
    const int kx_top = ek + esk;                               // scaled |kx| < 2^kx_top
    if (kx_top <= 15 && kx_top >= -8 && abs(ea) < 100 && abs(esk) < 100 && abs(eu) < 100) {
      r.s_ae = ldexpf(1.f, ea); r.s_kx = ldexpf(1.f, esk); r.s_u = ldexpf(1.f, 1 - eu); r.inv_unit = ldexpf(1.f, eu - 1);
      r.ok = 1;
    }
  }
  *out = r;
}

// Pre-split, pre-swizzled weights.  Slot (tile j, block kb2): kb2 < KB -> rows k of kx_a, gate order [h z r];
// kb2 >= KB -> rows k of U, gate order [z r h].  Tile row r = gate * nu + unit.  One thread per 16-byte chunk.
template <bool F16>
__global__ void gru_pack_kernel(const float* __restrict__ kxa, const float* __restrict__ U, unsigned char* __restrict__ out,
                                const GruScales* __restrict__ sc) {
  constexpr int KB = F16 ? GRU_KB16 : GRU_KB, BKE = F16 ? BK16 : BK, CE = F16 ? 8 : 4;   // elements per chunk
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_slot = 3 * GRU_NU * 8;
  if (idx >= (long long)GRU_NT * 2 * KB * per_slot) return;
  if (F16 && !sc->ok) return;
  const int c = idx % 8, r = (idx / 8) % (3 * GRU_NU);
  const int slot = idx / per_slot, kb2 = slot % (2 * KB), j = slot / (2 * KB);
  const int nu = min(GRU_NU, GAC_E - j * GRU_NU);
  if (r >= 3 * nu) return;
  const int half = kb2 >= KB, kb = kb2 - half * KB;
  const int gi = r / nu, u = j * GRU_NU + r % nu;
  const int gate = half ? gi : (gi + 2) % 3;                     // [z r h] for U, [h z r] for kx_a (z = 0, r = 1, h = 2)
  const float* W = half ? U : kxa;
  float x[CE];
#pragma unroll
  for (int e = 0; e < CE; ++e) {
    const int k = kb * BKE + CE * c + e;
    x[e] = k < GAC_E ? W[(size_t)k * GAC_G + gate * GAC_E + u] : 0.f;
  }
  unsigned char* base = out + (size_t)slot * GRU_BSLOT;
  const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
  if constexpr (F16) {
    const float sw = half ? sc->s_u : sc->s_kx;
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_f16x2(x[2 * e] * sw, x[2 * e + 1] * sw, h[e], l[e]);
    *reinterpret_cast<uint4*>(base + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(base + (size_t)3 * nu * 128 + off) = make_uint4(l[0], l[1], l[2], l[3]);
  } else {
    float h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_tf32(x[e], h[e], l[e]);
    *reinterpret_cast<float4*>(base + off) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(base + (size_t)3 * nu * 128 + off) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

template <bool F16>
__device__ __forceinline__ void gru_step_body(const GruStepArgs& g, const int bx) {
  constexpr int KB = F16 ? GRU_KB16 : GRU_KB, BKE = F16 ? BK16 : BK;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[GRU_STAGES], bar_empty[GRU_STAGES], bar_accum;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // 1-D grid, row-block major: the 7 tiles of a row block are neighbours in launch order, so the A rows (ae, h_prev)
  // come from DRAM once and from L2 six times (putting the narrow tiles last re-read all of A: measured 2x DRAM reads)
  const int jt = bx % GRU_NT;
  const int m0 = (bx / GRU_NT) * BM;
  const int m_end = g.dyn_rows ? min(g.rows, *g.dyn_rows) : g.rows;
  if (m0 >= m_end) return;                                       // dead tile (uniform per CTA)
  unsigned long long* dbg = (g.dbg && tid == 0) ? g.dbg + 8 + (size_t)bx * 8 : nullptr;
  if (dbg) {
    unsigned long long gt; unsigned sm;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
    dbg[0] = gt; dbg[1] = clock64(); dbg[7] = sm;
    if (bx == 0) g.dbg[0] = 0x600DC0DE600DC0DEull;
  }
  const int nu = min(GRU_NU, GAC_E - jt * GRU_NU);               // 64, or 16 for the last tile
  const int nkb = g.hprev ? 2 * KB : KB;
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
    const float s_ae = F16 ? g.sc->s_ae : 1.f;
    // 3xFP16 staging: a 64-wide K block is 16 float4 per thread: lane -> float4 (t & 15) of row (t >> 4) + 8 i, so a
    // warp instruction reads two full 256-byte row segments (every sector used).  Each float4 becomes 4 hi + 4 lo'
    // halves = two 8-byte stores.  The loads of the group's NEXT block are issued after the conversion of the current
    // one and before the wait for its smem stage, so part of their latency hides behind the hand-over.
    float x16[16][4];
    auto load16 = [&](int kb) {
      const int half = kb >= KB, k0 = (kb - half * KB) * BKE, live = min(BKE, GAC_E - k0);
      const int kq = t & 15;
      const float* src = (half ? g.hprev : g.ae) + (size_t)(m0 + (t >> 4)) * GAC_E + k0 + 4 * kq;
      const bool klive = kb < nkb && 4 * kq < live;               // live is a multiple of 16
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (klive && m0 + (t >> 4) + 8 * i < m_end) {
          const float4 v = *reinterpret_cast<const float4*>(src + (size_t)(8 * i) * GAC_E);
          x16[i][0] = v.x; x16[i][1] = v.y; x16[i][2] = v.z; x16[i][3] = v.w;
        } else {
          x16[i][0] = x16[i][1] = x16[i][2] = x16[i][3] = 0.f;
        }
      }
    };
    if constexpr (F16) load16(grp);
    for (int kb = grp; kb < nkb; kb += 2) {
      const int half = kb >= KB, k0 = (kb - half * KB) * BKE, live = min(BKE, GAC_E - k0);
      // registers of the staged tile: TF32 -> hi / lo fp32 of 8 chunks x 4; FP16 -> packed hi / lo halves of 16 float4
      float xa[8][4], la[8][4];
      uint32_t h16[16][2], l16[16][2];
      if constexpr (!F16) {
        load_operand<false, true, 8>(xa, half ? g.hprev : g.ae, GAC_E, m0, m_end, BM, k0, live, t);
        split_operand<8>(xa, la);     // waits for the loads, not for the stage
      } else {
        const float sc_a = half ? 1.f : s_ae;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          split_f16x2(x16[i][0] * sc_a, x16[i][1] * sc_a, h16[i][0], l16[i][0]);
          split_f16x2(x16[i][2] * sc_a, x16[i][3] * sc_a, h16[i][1], l16[i][1]);
        }
        load16(kb + 2);
      }
      const int s = kb % GRU_STAGES;
      const uint32_t ph = (uint32_t)(kb / GRU_STAGES) & 1u;
      mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
      const uint32_t sa = smem0 + (uint32_t)s * GRU_STAGE;
      if (t == 0) {
        const uint32_t bytes = 2u * b_lo_off;
        const unsigned char* srcp = (F16 ? g.wpk16 : g.wpk) + ((size_t)jt * 2 * KB + kb) * GRU_BSLOT;
        const uint32_t bar = smem_u32(&bar_full[s]);
        asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa + 2 * A_TILE_BYTES),
                     "l"(srcp), "r"(bytes), "r"(bar)
                     : "memory");
      }
      if constexpr (!F16) {
        store_split<false, 8>(xa, la, sa, sa + A_TILE_BYTES, BM, t);
      } else {
        const int kq = t & 15;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int r = (t >> 4) + 8 * i;
          const uint32_t off = (uint32_t)r * 128u + (uint32_t)(((kq >> 1) ^ (r & 7)) << 4) + (uint32_t)(kq & 1) * 8u;
          asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(sa + off), "r"(h16[i][0]), "r"(h16[i][1]));
          asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(sa + A_TILE_BYTES + off), "r"(l16[i][0]), "r"(l16[i][1]));
        }
      }
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
    const float inv_unit = F16 ? g.sc->inv_unit : 1.f;
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
        if constexpr (F16) {   // correction passes carry lo' = lo * 2^11; everything is in units of s_u
#pragma unroll
          for (int q = 0; q < 16; ++q) { v[q] = (v[q] + w[q] * (1.f / 2048.f)) * inv_unit; w[q] = 0.f; }
        }
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
      const uint32_t id3 = F16 ? make_idesc_f16(3 * nu) : make_idesc(3 * nu), id2 = F16 ? make_idesc_f16(2 * nu) : make_idesc(2 * nu),
                     id1 = F16 ? make_idesc_f16(nu) : make_idesc(nu);
      auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
        if constexpr (F16) mma_f16(d, a, b, id, acc); else mma_tf32(d, a, b, id, acc);
      };
      const uint32_t corr = tmem_d + 256u;
      for (int kb = 0; kb < nkb; ++kb) {
        const int half = kb >= KB, k0 = (kb - half * KB) * BKE, live = min(BKE, GAC_E - k0);
        const int s = kb % GRU_STAGES;
        const uint32_t ph = (uint32_t)(kb / GRU_STAGES) & 1u;
        mbar_wait(smem_u32(&bar_full[s]), ph);
        tc_fence_after();
        const uint32_t sa = smem0 + (uint32_t)s * GRU_STAGE;
        const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_TILE_BYTES);
        const uint64_t b_hi = make_desc(sa + 2 * A_TILE_BYTES), b_lo = make_desc(sa + 2 * A_TILE_BYTES + b_lo_off);
        const uint32_t dcol = half ? (uint32_t)nu : 0u;          // ae -> [xh z r], h_prev -> [z r hh]
        const int ksteps = F16 ? (live + 15) / 16 : (live + 7) / 8;   // 32 bytes of K per MMA either way
        for (int j = 0; j < ksteps; ++j) {
          const uint64_t adv = (uint64_t)(j * 2);
          if (half && kb == KB && j == 0) {
            // first touch of the hh columns: z, r keep accumulating, hh starts from zero
            const uint64_t bo = (uint64_t)((2u * (uint32_t)nu * 128u) >> 4);
            mma(corr + dcol, a_lo + adv, b_hi + adv, id2, 1u);
            mma(corr + dcol, a_hi + adv, b_lo + adv, id2, 1u);
            mma(tmem_d + dcol, a_hi + adv, b_hi + adv, id2, 1u);
            mma(corr + 3u * nu, a_lo + adv, b_hi + bo + adv, id1, 0u);
            mma(corr + 3u * nu, a_hi + adv, b_lo + bo + adv, id1, 1u);
            mma(tmem_d + 3u * nu, a_hi + adv, b_hi + bo + adv, id1, 0u);
            continue;
          }
          const uint32_t acc = (kb | j) ? 1u : 0u;
          mma(corr + dcol, a_lo + adv, b_hi + adv, id3, acc);
          mma(corr + dcol, a_hi + adv, b_lo + adv, id3, 1u);
          mma(tmem_d + dcol, a_hi + adv, b_hi + adv, id3, acc);
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

// one launch, variant chosen on the device (no host round trip, no second grid)
__global__ void __launch_bounds__(THREADS, 1) gru_step_tc_kernel(const GruStepArgs g) {
  if (g.skip_if) {
    // stand-by launch behind gru_step_v2_kernel: a small persistent grid, so that the usual case (v2 did the step) costs
    // one wave of CTAs that exit at once instead of thousands
    if (*g.skip_if) return;
    const int ntiles = GRU_NT * ((g.rows + BM - 1) / BM);
    for (int bx = blockIdx.x; bx < ntiles; bx += gridDim.x) {
      if (g.sc && g.sc->ok) gru_step_body<true>(g, bx);
      else gru_step_body<false>(g, bx);
      __syncthreads();
    }
    return;
  }
  if (g.sc && g.sc->ok) gru_step_body<true>(g, (int)blockIdx.x);
  else gru_step_body<false>(g, (int)blockIdx.x);
}

}  // namespace tc

// weights -> packed tiles of both variants (the 3xFP16 one exits at once when no valid scaling exists)
inline cudaError_t gru_pack(const float* kxa, const float* U, const float* wa, const float* ba, unsigned char* out32, unsigned char* out16,
                            tc::GruScales* sc, cudaStream_t st) {
  if (sc) {
    unsigned int* slots = reinterpret_cast<unsigned int*>(sc->pad);   // 3 scratch words inside the struct (rewritten by pass 2)
    cudaError_t e = cudaMemsetAsync(slots, 0, 3 * sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    tc::gru_amax_kernel<<<60, 256, 0, st>>>(kxa, U, wa, ba, slots);
    tc::gru_scales_kernel<<<1, 1, 0, st>>>(slots, sc);
    const long long t16 = (long long)tc::GRU_NT * 2 * tc::GRU_KB16 * 3 * tc::GRU_NU * 8;
    tc::gru_pack_kernel<true><<<(unsigned)cdiv64(t16, 256), 256, 0, st>>>(kxa, U, out16, sc);
  }
  const long long total = (long long)tc::GRU_NT * 2 * tc::GRU_KB * 3 * tc::GRU_NU * 8;
  tc::gru_pack_kernel<false><<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(kxa, U, out32, nullptr);
  return cudaGetLastError();
}

// g.wpk = 3xTF32 tiles; g.wpk16 / g.sc = 3xFP16 tiles and scales (or nullptr)
inline cudaError_t launch_gru_step(const tc::GruStepArgs& g, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc::gru_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GRU_SMEM);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  int nblk = tc::GRU_NT * cdiv(g.rows, tc::BM);
  if (g.skip_if && nblk > 148) nblk = 148;
  tc::gru_step_tc_kernel<<<nblk, tc::THREADS, tc::GRU_SMEM, st>>>(g);
  return cudaGetLastError();
}

}  // namespace gac
