// GRU step of the AIQN actor, second generation (sm_100a): a PERSISTENT, warp-specialised tcgen05 kernel whose
// epilogue overlaps the next tile's main loop and whose MMAs are 240 columns wide.
//
//   same arithmetic as gru_step_tc.cuh (Keras GRU cell, reset_after, gate order [z|r|h]; GAC/networks.py:166,217,259)
//
// What round 1's kernel left on the table (profiles/r1b_gru_step_summary.txt: 21.1k of a tile's 31.8k cycles in the MMA
// loop, the rest serial; N = 192 of 256 per instruction; a 16-unit remainder tile that costs a full main loop): the A
// operand (action embedding, h_prev) was staged through registers by the same eight warps that later ran the epilogue, so
// nothing could overlap, and TMEM was full (main + correction accumulator = 512 columns).  Here:
//   * BOTH operands arrive by cp.async.bulk.  The producers of the activations -- the embedding kernel for `ae`, this
//     kernel's own epilogue for `h` -- write, next to (or instead of) their fp32 output, the pre-split, pre-scaled,
//     pre-swizzled fp16 hi / lo tiles the MMA consumes ("a16" images: [row block][K block][hi | lo], 128 rows x 64 bytes,
//     K-major SWIZZLE_64B).  One producer lane issues two bulk copies per K block; no thread touches operand data.
//   * ONE accumulator per gate: hi*hi + lo*hi + hi*lo all add into the same TMEM columns with lo = x' - hi kept UNSCALED
//     (operands are scaled to ~2^12 so lo stays a normal fp16 for every element within 2^-10 of the largest).  The cost is
//     three accumulate-truncations per K step instead of one: measured max |h - fp64| 2.4e-6 against round 1's 0.9e-6
//     (tools/gru2_probe.py), inside the 1e-5 contract; the parity tests are the judge.
//   * THREE column groups per unit instead of four: [z | r | g].  z and r accumulate over both operand halves; g holds
//     ae * kx_h while the action-embedding half runs, is copied out by four drain warps the moment that half retires
//     (TMEM -> registers -> an L2-resident scratch row), and is then OVERWRITTEN by h_prev * U_h during the recurrent half
//     (reset_after needs the two parts of the candidate gate apart).  A tile is therefore 80 units x 3 = 240 columns:
//     every MMA is N = 240 (the tensor core needs N / 2 cycles per M = 128 kind::f16 instruction: 240 of the 256 columns it
//     could take), 400 units are exactly five tiles (no remainder tile), and TMEM holds TWO tiles, so the MMA lane fills one
//     while the epilogue warps work on the other.  7 x 150 -> 5 x 150 MMAs per 128-row block.
//   * 32-deep K blocks (64-byte rows): a stage is 46 KB, four stages fit beside the epilogue's staging buffer, so three
//     stages of operand data are in flight while one is consumed -- bulk-copy latency is off the critical path.
//   * persistent grid (one CTA per SM), static round-robin over (row block, unit tile) items: consecutive CTAs share a row
//     block (its A tiles are L2-hot).  (Tried and dropped: cp.async.bulk.prefetch.L2 of the next item's row block one tile
//     ahead -- 180.5 vs 176.8 us, profiles/r2e_gru2_prefetch_ab.txt: the operand stream is not waiting on HBM latency.)
//   * MMA issue: ONE instruction group per K block (mma_f16_kblock below).  Written instruction by instruction -- a predicate
//     and four 64-bit descriptor computations each -- the issuing lane needed ~168 cycles per MMA whatever N <= 256 and
//     whatever the cta_group (tools/mma_rate.cu, profiles/r2b_mma_rate.txt); the tensor core itself needs 120 at N = 240
//     (tools/mma_issue.cu, profiles/r2k_mma_issue.txt).
//   * what bounds it now (profiles/r2_final_gru2_timeline.txt, r2v_gru2_probe_bounds.txt): 18k of a tile's ~25k cycles are
//     MMA issue; the rest is the lane waiting for the epilogue to free an accumulator (its prefetched operand loads land
//     late behind the operand stream: 46 KB per 720 cycles = 64 B/clk per SM from L2) and for operands.  Removing those
//     stalls with dedicated TMEM reader warps cut the cycles and LOST time: the pool's B200s are power capped and the SM clock
//     fell with the stalls (profiles/r2z_gru2_reader_warps_probe.txt).
#pragma once
#include "gru_step_tc.cuh"

namespace gac {
namespace tc {

constexpr int G2_BK = 32;                                        // fp16 K elements per block = one 64-byte swizzle row
constexpr int G2_KB = (GAC_E + G2_BK - 1) / G2_BK;               // 13 K blocks per operand half (the last has 16 live columns)
constexpr int G2_NU = 80;                                        // hidden units per tile
constexpr int G2_NT = GAC_E / G2_NU;                             // 5 tiles, no remainder
static_assert(G2_NU * G2_NT == GAC_E && G2_NU % 16 == 0, "unit tiles must cover the 400 hidden units in 16-unit chunks");
constexpr uint32_t G2_A_TILE = (uint32_t)BM * 64u;               // 8 KB per hi / lo
constexpr uint32_t G2_A_SLOT = 2u * G2_A_TILE;                   // [hi | lo] of one (row block, K block)
constexpr uint32_t G2_B_TILE = 3u * G2_NU * 64u;                 // 15 KB per hi / lo: rows [z | r | h] x 80 units
constexpr uint32_t G2_B_SLOT = 2u * G2_B_TILE;
constexpr uint32_t G2_STAGE = G2_A_SLOT + G2_B_SLOT;             // 46 KB
constexpr int G2_STAGES = 4;
constexpr int G2_SS = 48;                                        // staging row stride in floats: [z16 | r16 | g16], quads XOR-rotated by (row >> 1) & 3:
                                                                 // conflict-free for phase 1 (8 consecutive rows, one quad) and phase 2 (2 rows x 4 quads)
constexpr int G2_DRAIN_WARPS = 8;                                // two per TMEM lane quadrant, 40 columns each
constexpr int G2_DS = 8;                                         // drain transposition row stride in floats; quad slot XOR (row >> 2) & 1
constexpr int G2_DRAIN_SMEM = G2_DRAIN_WARPS * 32 * G2_DS * 4;   // 8 KB: one 32 x 8 block per drain warp
constexpr int G2_SMEM = G2_STAGES * (int)G2_STAGE + BM * G2_SS * 4 + G2_DRAIN_SMEM + 4 * 3 * G2_NU * 4 + 1024;
// Warp roles, in warpgroups of four (setmaxnreg works per warpgroup): warps 0-15 epilogue (the gate math is a long dependent
// chain per thread: with two warps per scheduler its latencies were exposed and a tile's epilogue took ~29k cycles against a
// 23k-cycle main loop; four warps per scheduler hide them), warp 16 MMA issuer, warp 17 producer, warp 18 h_prev prefetcher,
// warp 19 idle, warps 20-27 drain (two per TMEM lane quadrant, so that a row's 80 columns are read in one batch of loads within
// the register budget).
constexpr int G2_EPI_WARPS = 16;
constexpr int G2_MMA_WARP = G2_EPI_WARPS, G2_PROD_WARP = G2_EPI_WARPS + 1;
constexpr int G2_DRAIN_WARP0 = G2_EPI_WARPS + 4;                 // warps 20..23 (lane quadrant = warp & 3)
constexpr int G2_THREADS = (G2_EPI_WARPS + 4 + G2_DRAIN_WARPS) * 32;   // 896: every role fits the 72 registers that leaves
// (ptxas 12.9 does not widen its allocation after setmaxnreg.inc -- tried: the drain region still spilled at the block-wide
// cap -- so every role is written to fit the 80 registers a 768-thread block gets.)
constexpr size_t G2_WPK_BYTES = (size_t)G2_NT * 2 * G2_KB * G2_B_SLOT;   // 4.0 MB
constexpr size_t G2_XH_BYTES_PER_CTA = 2u * BM * G2_NU * 4u;     // two tiles' ae * kx_h (one per accumulator buffer)
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
  // fp32 operands and weights for the in-kernel fallback (taken when the weights admit no fp16 scaling: sc->ok == 0)
  const float *ae32, *kxa32, *U32;
  float* xh_scratch;             // gridDim.x * G2_XH_BYTES_PER_CTA bytes (L2 resident): the candidate gate's input part between the halves
  int probe;                     // timing experiments (GAC_G2_PROBE bits, results are garbage): 1 no drain stores, 2 no TMEM loads in the
                                 // epilogue, 4 no global operand loads, 8 no global stores, 16 no gate math
  unsigned long long* dbg;       // optional per-CTA counters: [cycles, MMA-lane wait for operands, for a free accumulator, for the drain, tiles]
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

// Pre-split, pre-swizzled weights.  Slot (tile j, block kb2): kb2 < KB -> rows k of kx_a, kb2 >= KB -> rows k of U; tile row
// r = gate * 80 + unit in the natural gate order [z r h] for both; 64-byte rows (32 k).  One thread per 16-byte chunk.
__global__ void gru2_pack_kernel(const float* __restrict__ kxa, const float* __restrict__ U, unsigned char* __restrict__ out,
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
  unsigned char* base = out + (size_t)slot * G2_B_SLOT;
  const uint32_t off = sw64_off(r, c);
  *reinterpret_cast<uint4*>(base + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(base + G2_B_TILE + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// fp32 (rows, 400) -> a16 image, for activations whose producer does not write the image itself.  One thread per 16-byte
// chunk, ordered (K block, row, chunk) so that a warp writes 512 contiguous bytes and reads eight 128-byte row segments.
__global__ void a16_pack_kernel(const float* __restrict__ X, int ldx, int rows, const int* __restrict__ dyn_rows, int seg_rows,
                                const float* __restrict__ scale, unsigned char* __restrict__ out) {
  // rows come in segments of seg_rows (a multiple of 128; 0 = one segment) of which the first *dyn_rows are live
  const int nvalid = dyn_rows ? *dyn_rows : 0x7fffffff;
  const int n_rb = (rows + BM - 1) / BM;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_rb * G2_KB * BM * 4) return;
  const int c = idx & 3, r = (idx >> 2) & 127;
  const int t = idx >> 9, kb = t % G2_KB, rb = t / G2_KB;
  const int m = rb * BM + r, k0 = kb * G2_BK + 8 * c;
  const int in_seg = seg_rows > 0 ? m % seg_rows : m;
  if (seg_rows > 0 ? (rb * BM) % seg_rows >= nvalid : rb * BM >= nvalid) return;      // dead row block: never read
  const float s = *scale;
  uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
  if (m < rows && in_seg < nvalid && k0 < GAC_E) {                  // 400 = 50 x 8: a chunk is entirely inside or outside
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

// The step in plain fp32 for weights that admit no fp16 scaling (non-finite or absurd magnitudes, an all-zero action embedding):
// one (row, unit) per thread, 2 x 3 dot products of 400, accurate expf / tanhf.  Correct for any finite weights and slow by
// design (~6x the tensor path); it replaces a second kernel that used to be launched behind EVERY step just to find the flag
// clear and exit (12 launches and 8 weight-packing launches per train step).
__device__ __noinline__ void gru2_fallback(const Gru2Args& g, int m_end, bool has_h) {
  const long long total = (long long)m_end * GAC_E;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(e / GAC_E), u = (int)(e - (long long)m * GAC_E);
    const float* a = g.ae32 + (size_t)m * GAC_E;
    float xz = 0.f, xr = 0.f, xh = 0.f, hz = 0.f, hr = 0.f, hh = 0.f, hp = 0.f;
    for (int k = 0; k < GAC_E; ++k) {
      const float av = a[k];
      const float* w = g.kxa32 + (size_t)k * GAC_G + u;
      xz = fmaf(av, w[0], xz); xr = fmaf(av, w[GAC_E], xr); xh = fmaf(av, w[2 * GAC_E], xh);
    }
    if (has_h) {
      const float* h = g.hprev + (size_t)m * GAC_E;
      for (int k = 0; k < GAC_E; ++k) {
        const float hv = h[k];
        const float* w = g.U32 + (size_t)k * GAC_G + u;
        hz = fmaf(hv, w[0], hz); hr = fmaf(hv, w[GAC_E], hr); hh = fmaf(hv, w[2 * GAC_E], hh);
      }
      hp = h[u];
    }
    const float* sx = g.sx + (size_t)g.sidx[m] * GAC_G + u;
    const float z = 1.f / (1.f + expf(-((sx[0] + xz) + (hz + g.b1[u]))));
    const float r = 1.f / (1.f + expf(-((sx[GAC_E] + xr) + (hr + g.b1[GAC_E + u]))));
    const float hhb = hh + g.b1[2 * GAC_E + u];
    const float c = tanhf((sx[2 * GAC_E] + xh) + r * hhb);
    const size_t o = (size_t)m * GAC_E + u;
    g.hout[o] = z * hp + (1.f - z) * c;
    if (g.z) { g.z[o] = z; g.r[o] = r; g.c[o] = c; g.mhh[o] = hhb; }
  }
}

// One K block of the 3-pass product as ONE instruction group: hi*hi, lo*hi, hi*lo for both 16-element K steps.  The tensor
// core needs N / 2 cycles per M = 128 kind::f16 instruction (120 at N = 240), but the single issuing lane used to spend ~168:
// a predicate set-up and four 64-bit descriptor computations per instruction, interleaved with seven other warps of its
// scheduler (tools/mma_issue.cu, profiles/r2k_mma_issue.txt: the same loop stripped to this form runs at 120.1 cycles).
// Here the lane passes ONE descriptor (A hi of the stage); the other three differ from it by compile-time constants, the
// second K step by +32 bytes, and six MMAs share one predicate.  acc0: accumulate flag of the first instruction.
template <uint32_t A_LO, uint32_t B_HI, uint32_t B_LO>
__device__ __forceinline__ void mma_f16_kblock(uint32_t d, uint64_t a_hi, uint32_t idesc, uint32_t acc0) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b64 al, bh, bl, ah2, al2, bh2, bl2;\n"
      "setp.ne.b32 q, %3, 0;\n"
      "setp.eq.b32 p, 1, 1;\n"
      "add.u64 al, %1, %4;\n"
      "add.u64 bh, %1, %5;\n"
      "add.u64 bl, %1, %6;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, bh, %2, q;\n"
      "add.u64 ah2, %1, 2;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], al, bh, %2, p;\n"
      "add.u64 bh2, bh, 2;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, bl, %2, p;\n"
      "add.u64 al2, al, 2;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ah2, bh2, %2, p;\n"
      "add.u64 bl2, bl, 2;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], al2, bh2, %2, p;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ah2, bl2, %2, p;\n"
      "}\n" ::"r"(d), "l"(a_hi), "r"(idesc), "r"(acc0), "n"(A_LO), "n"(B_HI), "n"(B_LO)
      : "memory");
}
// the last K block of an operand half has 16 live columns: one K step
template <uint32_t A_LO, uint32_t B_HI, uint32_t B_LO>
__device__ __forceinline__ void mma_f16_kstep(uint32_t d, uint64_t a_hi, uint32_t idesc, uint32_t acc0) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b64 al, bh, bl;\n"
      "setp.ne.b32 q, %3, 0;\n"
      "setp.eq.b32 p, 1, 1;\n"
      "add.u64 al, %1, %4;\n"
      "add.u64 bh, %1, %5;\n"
      "add.u64 bl, %1, %6;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, bh, %2, q;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], al, bh, %2, p;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, bl, %2, p;\n"
      "}\n" ::"r"(d), "l"(a_hi), "r"(idesc), "r"(acc0), "n"(A_LO), "n"(B_HI), "n"(B_LO)
      : "memory");
}

template <bool DBG>
__global__ void __launch_bounds__(G2_THREADS, 1) gru_step_v2_kernel(const Gru2Args g) {
  // DBG = false: the counters and the timing-experiment switches are compiled out (they cost registers in every role)
  unsigned long long* const dbgp = DBG ? g.dbg : nullptr;
  const int probe = DBG ? g.probe : 0;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[G2_STAGES], bar_empty[G2_STAGES], bar_tfull[2], bar_tempty[2], bar_xfull[2], bar_xfree[2], bar_xsaved[2];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_end = g.dyn_rows ? min(g.rows, *g.dyn_rows) : g.rows;
  if (!g.sc->ok) {                                                 // no valid fp16 scaling of these weights: plain fp32, no tensor core
    gru2_fallback(g, m_end, g.hprev != nullptr);
    return;
  }
  const int n_items = ((m_end + BM - 1) / BM) * G2_NT;            // (row block, unit tile), row-block major
  const bool has_h = g.h_pk != nullptr;
  const int nkb = has_h ? 2 * G2_KB : G2_KB;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  float* stg = reinterpret_cast<float*>(smem_raw + (smem0 - smem_u32(smem_raw)) + G2_STAGES * G2_STAGE);
  float* dstg = stg + BM * G2_SS;                                  // drain warps' transposition blocks
  float* b1_base = dstg + G2_DRAIN_WARPS * 32 * G2_DS;                          // recurrent bias of the current tile, [z | r | h] x 80, one copy per epilogue group
  float* xh_cta = g.xh_scratch + (size_t)blockIdx.x * (G2_XH_BYTES_PER_CTA / 4);

  if (tid == 0) {
    for (int s = 0; s < G2_STAGES; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tfull[b]), 1); mbar_init(smem_u32(&bar_tempty[b]), G2_EPI_WARPS);
      mbar_init(smem_u32(&bar_xfull[b]), 1); mbar_init(smem_u32(&bar_xfree[b]), G2_DRAIN_WARPS); mbar_init(smem_u32(&bar_xsaved[b]), G2_DRAIN_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == G2_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (warp == G2_PROD_WARP) {
    // ------------------------------ producer: one lane, two bulk copies per K block ------------------------------
    if (lane == 0) {
      int it = 0;
      for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
        const int rb = i / G2_NT, jt = i - rb * G2_NT;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % G2_STAGES;
          const uint32_t ph = (uint32_t)(it / G2_STAGES) & 1u;
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
          const uint32_t sa = smem0 + (uint32_t)s * G2_STAGE, bar = smem_u32(&bar_full[s]);
          const int half = kb >= G2_KB, kbl = kb - half * G2_KB;
          const unsigned char* asrc = (half ? g.h_pk : g.ae_pk) + ((size_t)rb * G2_KB + kbl) * G2_A_SLOT;
          const unsigned char* bsrc = g.wpk + ((size_t)jt * 2 * G2_KB + kb) * G2_B_SLOT;
          if (probe & 64) { mbar_arrive(bar); continue; }            // probe 64: no operand copies (MMAs run on stale smem)
          mbar_expect_tx_arrive(bar, G2_STAGE);
          bulk_g2s(sa, asrc, G2_A_SLOT, bar);
          bulk_g2s(sa + G2_A_SLOT, bsrc, G2_B_SLOT, bar);
        }
      }
    }
    __syncwarp();
  } else if (warp == G2_MMA_WARP) {
    // ------------------------------ MMA issuer: one lane ------------------------------
    if (lane == 0) {
      long long t_wait_full = 0, t_wait_acc = 0, t_wait_x = 0;
      const long long t_begin = dbgp ? clock64() : 0;
      unsigned long long gt0 = 0;
      if (dbgp) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt0));
      constexpr uint32_t id3 = (1u << 4) | ((uint32_t)((3 * G2_NU) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);   // make_idesc_f16(240)
      constexpr uint32_t id2 = (1u << 4) | ((uint32_t)((2 * G2_NU) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      constexpr uint32_t id1 = (1u << 4) | ((uint32_t)(G2_NU >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint64_t desc0 = make_desc64(smem0);                   // A hi of stage 0
      int it = 0, t = 0;
      for (int i = blockIdx.x; i < n_items; i += gridDim.x, ++t) {
        const int b = t & 1;
        const uint32_t tph = (uint32_t)(t >> 1) & 1u;
        long long c0 = dbgp ? clock64() : 0;
        mbar_wait(smem_u32(&bar_tempty[b]), tph ^ 1u);             // the epilogue is done with this accumulator (two tiles ago)
        tc_fence_after();
        if (dbgp) t_wait_acc += clock64() - c0;
        const uint32_t d0 = tmem_d + (uint32_t)b * 256u;           // columns [z | r | g], 80 each
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % G2_STAGES;
          const uint32_t ph = (uint32_t)(it / G2_STAGES) & 1u;
          c0 = dbgp ? clock64() : 0;
          mbar_wait(smem_u32(&bar_full[s]), ph);
          tc_fence_after();
          if (dbgp) t_wait_full += clock64() - c0;
          // stages are contiguous: the descriptors of stage s are those of stage 0 plus s * (G2_STAGE >> 4) in the address field
          const uint64_t a_hi = desc0 + (uint64_t)((uint32_t)s * (G2_STAGE >> 4));
          constexpr uint32_t C_ALO = G2_A_TILE >> 4, C_BHI = G2_A_SLOT >> 4, C_BLO = (G2_A_SLOT + G2_B_TILE) >> 4;
          const bool last_of_half = kb == G2_KB - 1 || kb == 2 * G2_KB - 1;       // 400 = 12 x 32 + 16: one K step
          if (DBG && (probe & 32)) {
            // probe 32: no MMAs, the operand stream alone
          } else if (kb == G2_KB) {
            // the recurrent half starts: z, r keep accumulating; g switches from ae * kx_h (copied out by the drain warps)
            // to h_prev * U_h and starts from zero
            const uint64_t a_lo = a_hi + C_ALO, b_hi = a_hi + C_BHI, b_lo = a_hi + C_BLO;
            constexpr uint64_t bo = (uint64_t)((2u * G2_NU * 64u) >> 4);
            mma_f16(d0, a_hi, b_hi, id2, 1u);                      // z, r do not depend on the drain: ~3 MMA times of cover
            mma_f16(d0, a_lo, b_hi, id2, 1u);
            mma_f16(d0, a_hi, b_lo, id2, 1u);
            c0 = dbgp ? clock64() : 0;
            mbar_wait(smem_u32(&bar_xfree[b]), tph);
            tc_fence_after();
            if (dbgp) t_wait_x += clock64() - c0;
            mma_f16(d0 + 2u * G2_NU, a_hi, b_hi + bo, id1, 0u);
            mma_f16(d0 + 2u * G2_NU, a_lo, b_hi + bo, id1, 1u);
            mma_f16(d0 + 2u * G2_NU, a_hi, b_lo + bo, id1, 1u);
            mma_f16_kstep<C_ALO, C_BHI, C_BLO>(d0, a_hi + 2, id3, 1u);   // second K step of this block
          } else if (last_of_half) {
            mma_f16_kstep<C_ALO, C_BHI, C_BLO>(d0, a_hi, id3, 1u);
          } else {
            mma_f16_kblock<C_ALO, C_BHI, C_BLO>(d0, a_hi, id3, kb ? 1u : 0u);
          }
          mma_commit(smem_u32(&bar_empty[s]));
          if (has_h && kb == G2_KB - 1) mma_commit(smem_u32(&bar_xfull[b]));     // the action-embedding half has retired
        }
        mma_commit(smem_u32(&bar_tfull[b]));
      }
      if (dbgp) {
        unsigned long long* d = dbgp + 8 + (size_t)blockIdx.x * 16;
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt1));
        d[0] = (unsigned long long)(clock64() - t_begin); d[1] = (unsigned long long)t_wait_full; d[2] = (unsigned long long)t_wait_acc;
        d[3] = (unsigned long long)t_wait_x; d[4] = (unsigned long long)t; d[5] = gt1 - gt0;
        if (blockIdx.x == 0) { dbgp[0] = 0x600DC0DE600DC0E4ull; dbgp[1] = gridDim.x; }
      }
    }
    __syncwarp();
  } else if (warp == G2_PROD_WARP + 1) {
    // ------------------------------ h_prev prefetcher ------------------------------
    // The gate math reads h_prev (fp32, from HBM: the previous step wrote 52 MB of it) one 16-unit chunk ahead -- about one
    // barrier and one TMEM read of lead, far less than an HBM round trip under load, so every chunk of every tile used to
    // wait for it.  This warp pulls the NEXT tile's 128 x 80 block (three 128-byte lines per row) into L2 while the current
    // tile's MMAs run; it paces itself on the accumulator barriers (waiting on a phase consumes nothing).
    if (has_h) {
      int t = 0;
      for (int i = blockIdx.x; i < n_items; i += gridDim.x, ++t) {
        if (t > 0) mbar_wait(smem_u32(&bar_tfull[(t - 1) & 1]), (uint32_t)((t - 1) >> 1) & 1u);   // tile t - 1 has left the tensor core
        const int rb = i / G2_NT, jt = i - rb * G2_NT;
#pragma unroll
        for (int k = 0; k < 12; ++k) {                            // 128 rows x 3 lines = 384 = 12 per lane
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
    // ------------------------------ drain warps: g = ae * kx_h out of TMEM between the two halves ------------------------------
    if (has_h) {
      const int quarter = warp & 3;                               // the TMEM lane quadrant this warp may read
      const float inv_unit = g.sc->inv_unit;
      const bool stamp = dbgp && tid == G2_DRAIN_WARP0 * 32;
      long long dw = 0, dk = 0;
      int t = 0;
      for (int i = blockIdx.x; i < n_items; i += gridDim.x, ++t) {
        const int b = t & 1;
        const uint32_t tph = (uint32_t)(t >> 1) & 1u;
        long long c0 = stamp ? clock64() : 0;
        mbar_wait(smem_u32(&bar_xfull[b]), tph);
        tc_fence_after();
        long long c1 = stamp ? clock64() : 0;
        dw += c1 - c0;
        const int hsel = (warp - G2_DRAIN_WARP0) >> 2;             // which 40 of the 80 g columns
        const uint32_t d0 = tmem_d + (uint32_t)b * 256u + 2u * G2_NU + (uint32_t)(hsel * 40) + ((uint32_t)(quarter * 32) << 16);
        // 40 columns of this lane's row: three loads in flight, ONE wait, and the MMA lane is released; only then the values
        // go through a 32 x 8 block in smem to (row, quad) float4 so that every global store is four rows x 32 contiguous bytes
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
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_xfree[b]));       // the MMA lane may overwrite g now
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
        __syncwarp();                                              // orders the lanes' stores before lane 0's releasing arrive
        if (lane == 0) mbar_arrive(smem_u32(&bar_xsaved[b]));
        if (stamp) dk += clock64() - c1;
      }
      if (stamp) { unsigned long long* d = dbgp + 8 + (size_t)blockIdx.x * 16; d[12] = (unsigned long long)dw; d[13] = (unsigned long long)dk; }
    }
  } else {
    // ------------------------------ epilogue: sixteen warps, overlapped with the next tile's main loop ------------------------------
    const int quarter = warp & 3, sub = warp >> 2;                 // TMEM lane quadrant; phase-1 column group (0 z, 1 r, 2 g, 3 none)
    // Four independent quarter-tile groups: the four warps of a TMEM lane quadrant own its 32 rows.  Each group runs its own
    // phase 1 -> barrier -> phase 2 -> barrier cycle on its own staging rows and named barrier, so while one group waits for
    // its operand loads (the exposed latency of this epilogue: profiles/r2v_gru2_probe_bounds.txt) the others read TMEM or do
    // gate math, instead of all sixteen warps stalling in step.
    const int grp = quarter, gw = sub;                              // group; warp index within the group (0..3)
    const int gtid = gw * 32 + lane;                                // thread index within the group (0..127)
    const uint32_t gbar = 1u + (uint32_t)grp;
    float* b1_s = b1_base + grp * (3 * G2_NU);
#define G2_GROUP_SYNC() asm volatile("bar.sync %0, 128;" ::"r"(gbar) : "memory")
    const float inv_unit = g.sc->inv_unit, s_h = g.sc->s_h;
    auto tnh = [](float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); };
    // phase-2 ownership: ONE (row, unit quad) item per thread and chunk: row = 8 warp + (lane >> 2), quad = lane & 3.  128-bit
    // global accesses are coalesced per QUARTER warp: eight consecutive lanes = two rows x 64 contiguous bytes = four full
    // sectors.  (The first version put eight different rows into a quarter warp to make the staging reads conflict-free: every
    // global access then touched eight half-used sectors, and the operand loads alone cost 2.5k cycles per chunk.)
    const int prow = 32 * grp + gw * 8 + (lane >> 2), pj = lane & 3;
    const bool stamp = dbgp && tid == 0;
    long long e_wt = 0, e_wx = 0, e_p1 = 0, e_b1 = 0, e_p2 = 0, e_b2 = 0, e_ld = 0, ec = 0;
#define G2_STAMP(acc) if (stamp) { const long long now = clock64(); acc += now - ec; ec = now; }
    int t = 0;
    for (int i = blockIdx.x; i < n_items; i += gridDim.x, ++t) {
      const int rb = i / G2_NT, jt = i - rb * G2_NT;
      const int m0 = rb * BM, b = t & 1;
      if (stamp) ec = clock64();
      const uint32_t tph = (uint32_t)(t >> 1) & 1u;
      const int m = m0 + prow;
      const int sid = m < m_end ? g.sidx[m] : -1;
      if (gtid < 3 * G2_NU / 4)                                    // this tile's recurrent bias -> the group's smem copy
        *reinterpret_cast<float4*>(b1_s + 4 * gtid) = *reinterpret_cast<const float4*>(g.b1 + (gtid / (G2_NU / 4)) * GAC_E + jt * G2_NU + 4 * (gtid % (G2_NU / 4)));
      if (has_h) mbar_wait(smem_u32(&bar_xsaved[b]), tph);         // acquire: the drain warps' scratch rows are visible
      G2_STAMP(e_wx)
      // Operands of phase 2 that do not depend on the accumulators (gathered state projection, h_prev, the saved input part):
      // those of chunk q + 1 are issued inside phase 2 of chunk q, right after the registers they overwrite were consumed, and
      // fly during the barrier and the next phase 1.  (.cg: nothing here is reused through L1, which is 28 KB beside the smem.)
      float4 sz, sr, sh, hp, xh;
      auto load_ops = [&](int q) {
        hp = make_float4(0.f, 0.f, 0.f, 0.f); xh = hp;
        if (probe & 4) { sz = sr = sh = hp; return; }
        if (sid < 0) return;
        const int ul = q * 16 + 4 * pj, u = jt * G2_NU + ul;
        const float* sx = g.sx + (size_t)sid * GAC_G + u;
        // the per-state projection goes through L1: in the state-major half of the rows 64 consecutive rows share a state
        sz = __ldg(reinterpret_cast<const float4*>(sx)); sr = __ldg(reinterpret_cast<const float4*>(sx + GAC_E));
        sh = __ldg(reinterpret_cast<const float4*>(sx + 2 * GAC_E));
        if (has_h) {
          hp = __ldcg(reinterpret_cast<const float4*>(g.hprev + (size_t)m * GAC_E + u));
          xh = __ldcg(reinterpret_cast<const float4*>(xh_cta + ((size_t)b * BM + prow) * G2_NU + ul));
        }
      };
      load_ops(0);
      G2_STAMP(e_ld)
      G2_GROUP_SYNC();                                             // this tile's bias rows are in smem (phase 1 reads them)
      mbar_wait(smem_u32(&bar_tfull[b]), tph);
      tc_fence_after();
      G2_STAMP(e_wt)
      const uint32_t d0 = tmem_d + (uint32_t)b * 256u + ((uint32_t)(quarter * 32) << 16);
      for (int q = 0; q < G2_NU / 16; ++q) {
        const int ul = q * 16 + 4 * pj, u = jt * G2_NU + ul;       // this lane's four hidden units (within the tile / of 400)
        // phase 1: 16 units x [z r g] of 128 rows, TMEM -> staging: warp (quadrant, sub) moves 32 rows of column group `sub`
        if (sub < 3) {
          float* srow = stg + (size_t)(quarter * 32 + lane) * G2_SS + sub * 16;
          const int rot = (lane >> 1) & 3;                           // == (row >> 1) & 3: the row's quad rotation
          // recurrent bias of this column group, added here (one broadcast read per warp) instead of per item in phase 2;
          // on the first step g holds the INPUT part of the candidate gate, whose bias is already in sx
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
        G2_STAMP(e_p1)
        G2_GROUP_SYNC();
        G2_STAMP(e_b1)
        // phase 2: gate math and the state update, row-contiguous global traffic (no branch around the math: dead rows only
        // skip their stores)
        {
          const float* srow = stg + (size_t)prow * G2_SS + 4 * ((pj + (prow >> 1)) & 3);
          const float4 az = *reinterpret_cast<const float4*>(srow), ar = *reinterpret_cast<const float4*>(srow + 16),
                       ag = *reinterpret_cast<const float4*>(srow + 32);     // z, r (and g with a recurrent half) include the recurrent bias
          float4 bh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!has_h) bh = *reinterpret_cast<const float4*>(b1_s + 2 * G2_NU + ul);   // first step: h_prev * U_h = 0, the bias alone
          // with a recurrent half g holds h_prev * U_h and the input part comes from the scratch; on the first step g IS the input part
          const float4 axh = has_h ? xh : ag;
          const float4 ahh = has_h ? ag : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 ho, zo, ro, co, mo;
          if (probe & 16) { ho = az; zo = ar; ro = ag; co = axh; mo = ahh; } else {
          // sigmoid(a), sigmoid(b) share ONE reciprocal: with p = 1 + e^-a, q = 1 + e^-b, R = 1 / (p q): z = q R, r = p R.  The
          // exponents are clamped to +-30 (sigmoid is 0 or 1 to 1e-13 there), so p q <= 1.2e26 never overflows.
#define GRU2_LANE(f)                                                                              \
  {                                                                                               \
    const float hh = ahh.f + bh.f;                                                                \
    const float pa = 1.f + __expf(-fminf(fmaxf(sz.f + az.f, -30.f), 30.f));                       \
    const float pb = 1.f + __expf(-fminf(fmaxf(sr.f + ar.f, -30.f), 30.f));                       \
    const float R = __fdividef(1.f, pa * pb);                                                     \
    const float zz = pb * R, rr = pa * R;                                                         \
    const float cc = tnh((sh.f + axh.f) + rr * hh);                                               \
    ho.f = zz * hp.f + (1.f - zz) * cc; zo.f = zz; ro.f = rr; co.f = cc; mo.f = hh;               \
  }
          GRU2_LANE(x) GRU2_LANE(y) GRU2_LANE(z) GRU2_LANE(w)
#undef GRU2_LANE
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
          if (g.hout_pk) {                                         // the next step's A operand, already split / scaled / swizzled
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
        G2_STAMP(e_p2)
        G2_GROUP_SYNC();                                           // the staging rows are rewritten by the next chunk / tile
        G2_STAMP(e_b2)
      }
      // this tile's accumulator columns and scratch rows are free again (all their reads above have completed)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[b]));
    }
#undef G2_STAMP
#undef G2_GROUP_SYNC
    if (stamp) {
      unsigned long long* d = dbgp + 8 + (size_t)blockIdx.x * 16;
      d[6] = (unsigned long long)e_wt; d[7] = (unsigned long long)e_wx; d[8] = (unsigned long long)e_p1; d[9] = (unsigned long long)e_b1;
      d[10] = (unsigned long long)e_p2; d[11] = (unsigned long long)e_b2; d[14] = (unsigned long long)e_ld;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == G2_MMA_WARP) {
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
  const long long total = (long long)tc::G2_NT * 2 * tc::G2_KB * 3 * tc::G2_NU * 4;
  tc::gru2_pack_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(kxa, U, out, sc);
  return cudaGetLastError();
}
inline cudaError_t a16_pack(const float* X, int ldx, long long rows, const int* dyn_rows, int seg_rows, const float* scale, unsigned char* out,
                            cudaStream_t st) {
  const long long total = cdiv64(rows, tc::BM) * tc::G2_KB * tc::BM * 4;
  tc::a16_pack_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(X, ldx, (int)rows, dyn_rows, seg_rows, scale, out);
  return cudaGetLastError();
}
inline int gru2_grid() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    if (sms > 256) sms = 256;                                    // the scratch rows are sized for 256 CTAs (G2_MAX_GRID)
  }
  return sms;
}
inline cudaError_t launch_gru_step_v2(const tc::Gru2Args& g, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc::gru_step_v2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::G2_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gru_step_v2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::G2_SMEM);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int items = tc::G2_NT * cdiv(g.rows, tc::BM);
  static const int grid_env = getenv("GAC_G2_GRID") ? atoi(getenv("GAC_G2_GRID")) : 0;
  const int cap = grid_env > 0 && grid_env < gru2_grid() ? grid_env : gru2_grid();
  const int grid = items < cap ? items : cap;
  if (g.dbg || g.probe) tc::gru_step_v2_kernel<true><<<grid, tc::G2_THREADS, tc::G2_SMEM, st>>>(g);
  else tc::gru_step_v2_kernel<false><<<grid, tc::G2_THREADS, tc::G2_SMEM, st>>>(g);
  return cudaGetLastError();
}

}  // namespace gac
