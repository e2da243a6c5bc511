// tcgen05 (5th-gen tensor core) GEMM with fp32-accurate 3xTF32 splitting, sm_100a only.
//
//   C(128 x bn tile) = sum_k  A_hi*B_hi  +  sum_k (A_hi*B_lo + A_lo*B_hi)   (fp32 accumulate in TMEM)
//   hi = rna_tf32(x), lo = x - hi   (|lo| <= 2^-12 |x|; the dropped lo*lo term is ~2^-24)
// The tensor core truncates when it adds into the fp32 accumulator (measured: error grows with
// the number of MMAs per output, ~9e-6 at K = 1200 with one accumulator), so the two correction
// passes go to a SECOND TMEM accumulator (its magnitude, hence its truncation error, is 2^-12 of
// the main one) and the epilogue adds the two in fp32: 1 truncation per K-step instead of 3.
//
// Why: BASELINE.json asks for 1e-5 forwards, which single-pass TF32 (2^-11 input rounding)
// cannot meet.  Roofline for this engine = measured TF32 dense peak / 3.
//
// Structure (one 128 x bn output tile per CTA, bn a multiple of 16 <= 256):
//   warps 0-7  loaders (two groups of 4 warps alternate K blocks so one group's global-load
//              latency overlaps the other's smem stores), then epilogue.  They read the fp32
//              operands straight from their natural
//              row-major arena / activation layouts (either orientation: the transposition
//              needed by weight-gradient contractions happens here, in registers), split every
//              value into hi/lo and store both into K-major, 128B-swizzled smem tiles that
//              tcgen05.mma consumes directly -- no packed weight copies, no lo-twins in HBM, half
//              the L2->SM traffic of pre-split operands.
//   warp 8     allocates TMEM and issues the MMAs from one lane: per 32-wide K block, 4 K-steps
//              (UMMA_K = 8 for tf32) x 3 passes into one TMEM accumulator; tcgen05.commit frees
//              the smem stage / signals the epilogue through mbarriers.
//   epilogue   tcgen05.ld (32 lanes x 16 columns per warp instruction) -> fused epilogue of
//              gac_common.cuh (bias, gathered add, LeakyReLU, Hadamard, LeakyReLU', accumulate).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "gac_common.cuh"

namespace gac {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;                     // fp32 elements per K block = one 128-byte swizzle row
constexpr int LOADER_WARPS = 8;
constexpr int THREADS = (LOADER_WARPS + 1) * 32;
constexpr int MAX_STAGES = 4;
constexpr uint32_t A_TILE_BYTES = BM * 128;   // 16 KB
constexpr int SMEM_LIMIT = 227 * 1024 - 2048;   // dynamic part; static barriers + alignment slack stay below the 227 KB cap

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): start address >> 4 in
// bits [0,14), LBO (unused for swizzled K-major, set to 1) in [16,30), SBO = 1024 B (one 8-row
// group) in [32,46), version 1 at bit 46, layout type 2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32 instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (2 at bits 7-9 / 10-12),
// both operands K-major, N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// MN-major operands (bit 15: A, bit 16: B): the operand's M / N index is the contiguous one in shared memory.
// For 4-byte (tf32) elements the only MN-major layout the tensor core accepts is SWIZZLE_128B_BASE32B (layout type 1;
// cute: Swizzle<2,5,2> o ((4,8,m),(4,k)) : ((1,4,LBO),(32,SBO))): groups of 32 consecutive M/N elements (128 bytes) form
// one row of a 4-row (4 k) 512-byte atom whose 32-byte units are XOR-ed with the row index (address bits [5,7) ^= bits
// [7,9)); consecutive groups are LBO apart, consecutive 4-k atoms SBO apart.  This kernel stores a 32-deep K block of one
// group contiguously (32 k-rows x 128 B = 4096 B = LBO, SBO = 512 B), so one K step (8 k = 2 atoms) advances the start
// address by 1024 bytes.  A weight-gradient operand -- an activation matrix (batch rows, features) read as (features,
// batch) -- is copied into this layout row by row with 128-bit loads and stores: no transposition anywhere.
__device__ __forceinline__ uint32_t make_idesc_major(int n, bool a_mn, bool b_mn) {
  return make_idesc(n) | (a_mn ? 1u << 15 : 0u) | (b_mn ? 1u << 16 : 0u);
}
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc_f16(int n) {     // kind::f16: D = F32, A = B = F16 (format 0), K-major
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Consecutive K steps of the dual-accumulator 3-pass product as ONE instruction group (see gru_step_v2.cuh, mma_f16_kblock:
// the issuing lane, not the tensor core, limits a loop that rebuilds a predicate and the descriptor advances per
// instruction).  Per K step: corr (+)= lo * hi, corr += hi * lo, main (+)= hi * hi.  acc0: accumulate flag of the first step's
// first write to each accumulator.  K-major SWIZZLE_128B tiles advance by 32 bytes (2 >> 4 units) per 16-element step,
// the MN-major weight-gradient tiles by 16 k-rows x 128 bytes (128 units).
#define GAC_MMA_F16 "tcgen05.mma.cta_group::1.kind::f16 "
#define GAC_DUAL_FIRST GAC_MMA_F16 "[%1], %3, %4, %6, q;\n" GAC_MMA_F16 "[%1], %2, %5, %6, p;\n" GAC_MMA_F16 "[%0], %2, %4, %6, q;\n"
#define GAC_DUAL_NEXT(adv)                                                                                          \
  "add.u64 ah, %2, " adv ";\n add.u64 al, %3, " adv ";\n add.u64 bh, %4, " adv ";\n add.u64 bl, %5, " adv ";\n" \
  GAC_MMA_F16 "[%1], al, bh, %6, p;\n" GAC_MMA_F16 "[%1], ah, bl, %6, p;\n" GAC_MMA_F16 "[%0], ah, bh, %6, p;\n"
#define GAC_DUAL_ASM(body)                                                                                                         \
  asm volatile("{\n.reg .pred p, q;\n.reg .b64 ah, al, bh, bl;\nsetp.ne.b32 q, %7, 0;\nsetp.eq.b32 p, 1, 1;\n" body "}\n" ::"r"(d_main), \
               "r"(d_corr), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(acc0)                                       \
               : "memory")
template <int KSTEPS>
__device__ __forceinline__ void mma_f16_dual_k(uint32_t d_main, uint32_t d_corr, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                               uint32_t idesc, uint32_t acc0) {
  static_assert(KSTEPS >= 1 && KSTEPS <= 4, "one to four K steps per group");
  if constexpr (KSTEPS == 1) GAC_DUAL_ASM(GAC_DUAL_FIRST);
  else if constexpr (KSTEPS == 2) GAC_DUAL_ASM(GAC_DUAL_FIRST GAC_DUAL_NEXT("2"));
  else if constexpr (KSTEPS == 3) GAC_DUAL_ASM(GAC_DUAL_FIRST GAC_DUAL_NEXT("2") GAC_DUAL_NEXT("4"));
  else GAC_DUAL_ASM(GAC_DUAL_FIRST GAC_DUAL_NEXT("2") GAC_DUAL_NEXT("4") GAC_DUAL_NEXT("6"));
}
// single accumulator, unscaled lo (emb_gemm_tc.cuh): d (+)= hi * hi, d += lo * hi, d += hi * lo for four K steps of a K-major
// SWIZZLE_128B block (K = 64)
#define GAC_SGL_FIRST GAC_MMA_F16 "[%0], %1, %3, %5, q;\n" GAC_MMA_F16 "[%0], %2, %3, %5, p;\n" GAC_MMA_F16 "[%0], %1, %4, %5, p;\n"
#define GAC_SGL_NEXT(adv)                                                                                          \
  "add.u64 ah, %1, " adv ";\n add.u64 al, %2, " adv ";\n add.u64 bh, %3, " adv ";\n add.u64 bl, %4, " adv ";\n" \
  GAC_MMA_F16 "[%0], ah, bh, %5, p;\n" GAC_MMA_F16 "[%0], al, bh, %5, p;\n" GAC_MMA_F16 "[%0], ah, bl, %5, p;\n"
__device__ __forceinline__ void mma_f16_single_k4(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc, uint32_t acc0) {
  asm volatile("{\n.reg .pred p, q;\n.reg .b64 ah, al, bh, bl;\nsetp.ne.b32 q, %6, 0;\nsetp.eq.b32 p, 1, 1;\n" GAC_SGL_FIRST GAC_SGL_NEXT("2")
                   GAC_SGL_NEXT("4") GAC_SGL_NEXT("6") "}\n" ::"r"(d),
               "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(acc0)
               : "memory");
}
template <int KSTEPS>
__device__ __forceinline__ void mma_f16_dual_mn(uint32_t d_main, uint32_t d_corr, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                                uint32_t idesc, uint32_t acc0) {
  static_assert(KSTEPS >= 1 && KSTEPS <= 2, "one or two K steps per 32-deep block");
  if constexpr (KSTEPS == 1) GAC_DUAL_ASM(GAC_DUAL_FIRST);
  else GAC_DUAL_ASM(GAC_DUAL_FIRST GAC_DUAL_NEXT("128"));
}
// x (already scaled) -> fp16 hi and fp16 lo' = (x - hi) * 2^11, two values per 32-bit word
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);                   // one cvt.rn.f16x2.f32
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((x0 - hf.x) * 2048.f, (x1 - hf.y) * 2048.f);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// MN-major fp16 operand tile (standard SWIZZLE_128B: 64 M/N elements = 128 bytes per row, 8-row atoms, 16-byte chunks XOR
// row & 7): a 32-deep K block of one 64-element group is stored contiguously (32 k-rows x 128 B = 4096 B = LBO), 8-k atoms
// are SBO = 1024 B apart, and one K step (16 k) advances the start address by 2048 bytes.
__device__ __forceinline__ uint64_t make_desc_mn16(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// float4 of 4 consecutive M/N elements (already scaled) at k -> 4 hi + 4 lo' halves, 8 bytes each
template <int NCH>
__device__ __forceinline__ void store_mn16(const float (&x)[NCH][4], float scale, uint32_t s_hi, uint32_t s_lo, int rows, int t) {
  constexpr int CPR = NCH * 4;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int q = t + 128 * i, k = q / CPR, c4 = q % CPR;
    if (4 * c4 >= rows) continue;
    uint32_t h0, h1, l0, l1;
    split_f16x2(x[i][0] * scale, x[i][1] * scale, h0, l0);
    split_f16x2(x[i][2] * scale, x[i][3] * scale, h1, l1);
    const uint32_t off = (uint32_t)(c4 >> 4) * 4096u + (uint32_t)k * 128u + (uint32_t)(((((c4 & 15) >> 1) ^ (k & 7)) << 4) | ((c4 & 1) << 3));
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(s_hi + off), "r"(h0), "r"(h1));
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(s_lo + off), "r"(l0), "r"(l1));
  }
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}

// ---- K-block schedule ---------------------------------------------------------------------
// Non-dynamic GEMMs walk [kbeg, kend) in 32-wide blocks.  Weight-gradient GEMMs reduce over
// activation rows laid out step-major as nseg segments of seg rows of which only the first
// `nvalid` are live (M is data dependent and stays on the device): only LIVE blocks are
// enumerated and split-K partitions the live blocks, so dead rows cost neither loads nor MMAs.
struct KMap {
  int dyn, kbeg, kend, seg, bps, nvalid, ibeg, nkb;
};
__device__ __forceinline__ KMap make_kmap(const GemmArgs& g, int nvalid, int split) {
  KMap km;
  km.dyn = (g.dyn_rows && g.valid_on_k) ? 1 : 0;
  km.seg = g.seg_rows; km.nvalid = nvalid; km.kbeg = 0; km.kend = g.K; km.bps = 0; km.ibeg = 0;
  if (km.dyn) {
    const int nv = min(nvalid, g.seg_rows);
    km.nvalid = nv;
    km.bps = (nv + BK - 1) / BK;
    const int total = (g.K / g.seg_rows) * km.bps;
    const int per = g.splitk > 1 ? (total + g.splitk - 1) / g.splitk : total;
    km.ibeg = min(total, split * per);
    km.nkb = min(total, km.ibeg + per) - km.ibeg;
  } else {
    if (g.splitk > 1) {
      const int per = ((g.K + g.splitk - 1) / g.splitk + BK - 1) / BK * BK;
      km.kbeg = min(g.K, split * per);
      km.kend = min(g.K, km.kbeg + per);
    }
    km.nkb = (km.kend - km.kbeg + BK - 1) / BK;
  }
  return km;
}
__device__ __forceinline__ void kmap_get(const KMap& km, int kb, int& k0, int& live) {
  if (km.dyn) {
    const int i = km.ibeg + kb;
    const int sgm = i / km.bps, blk = i - sgm * km.bps;
    k0 = sgm * km.seg + blk * BK;
    live = min(BK, km.nvalid - blk * BK);
  } else {
    k0 = km.kbeg + kb * BK;
    live = min(BK, km.kend - k0);
  }
}

// ---- operand staging ------------------------------------------------------------------------
// One loader group (128 threads) stages a rows x 32 fp32 tile: global -> registers (all loads
// issued back to back so their latencies overlap), then registers -> hi/lo split -> smem.
//   TRANS = false: element (r, k) = src[(row0 + r) * ld + k0 + k]   (k contiguous in memory)
//   TRANS = true : element (r, k) = src[(k0 + k) * ld + row0 + r]   (r contiguous in memory)
// Thread t owns 16-byte chunks q = t + 128 i (4 consecutive k of one row); chunk c of row r
// lands at r*128 + ((c ^ (r & 7)) << 4): the 128-byte swizzle the UMMA descriptor declares.
constexpr int GROUP = 128;
// Chunk ownership of thread t (t in [0,128)), chosen so that no per-chunk division is needed and
// both global loads and swizzled smem stores are conflict-free / coalesced:
//   !TRANS: chunk i -> row (t >> 3) + 16 i, 16-byte column c = t & 7     (8 lanes = one 128-B row)
//    TRANS: chunk i -> row t + 128 (i >> 3), column c = i & 7            (32 lanes = 32 consecutive rows)
template <bool TRANS>
__device__ __forceinline__ void chunk_rc(int i, int t, int& r, int& c) {
  if (!TRANS) { r = (t >> 3) + 16 * i; c = t & 7; } else { r = t + GROUP * (i >> 3); c = i & 7; }
}
template <bool TRANS, bool VEC, int NCH>
__device__ __forceinline__ void load_operand(float (&x)[NCH][4], const float* __restrict__ src, int ld, int row0, int row_end,
                                             int rows, int k0, int live, int t) {
  if (!TRANS) {
    const int c = t & 7, kc = 4 * c;
    const float* p0 = src + (size_t)(row0 + (t >> 3)) * ld + k0 + kc;
    const bool full = VEC && (kc + 3 < live);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int r = (t >> 3) + 16 * i;
      const float* p = p0 + (size_t)(16 * i) * ld;
      x[i][0] = x[i][1] = x[i][2] = x[i][3] = 0.f;
      if (r < rows && row0 + r < row_end) {
        if (full) {
          const float4 v = *reinterpret_cast<const float4*>(p);
          x[i][0] = v.x; x[i][1] = v.y; x[i][2] = v.z; x[i][3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (kc + j < live) x[i][j] = p[j];
        }
      }
    }
  } else {
#pragma unroll
    for (int ii = 0; ii < NCH / 8; ++ii) {
      const int r = t + GROUP * ii;
      const bool ok = r < rows && row0 + r < row_end;
      const float* p = src + (size_t)k0 * ld + row0 + r;
      if (ok && live == BK) {
#pragma unroll
        for (int k = 0; k < BK; ++k) x[ii * 8 + (k >> 2)][k & 3] = p[(size_t)k * ld];
      } else {
#pragma unroll
        for (int k = 0; k < BK; ++k) x[ii * 8 + (k >> 2)][k & 3] = (ok && k < live) ? p[(size_t)k * ld] : 0.f;
      }
    }
  }
}
template <bool TRANS, int NCH>
__device__ __forceinline__ void store_operand(const float (&x)[NCH][4], uint32_t s_hi, uint32_t s_lo, int rows, int t) {
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    int r, c;
    chunk_rc<TRANS>(i, t, r, c);
    if (r >= rows) continue;
    float h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_tf32(x[i][j], h[j], l[j]);
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(s_hi + off), "f"(h[0]), "f"(h[1]), "f"(h[2]), "f"(h[3]));
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(s_lo + off), "f"(l[0]), "f"(l[1]), "f"(l[2]), "f"(l[3]));
  }
}

// Two-phase variant of store_operand: the hi/lo split happens in registers while the loader is still waiting for
// its smem stage, so that only the 128-bit stores remain on the stage's critical path (measured: the stage
// turn-around -- commit -> loader wake-up -> split + store -> arrive -> MMA wake-up -- is ~1,600 cycles, and with
// two stages per-K-block time is (T_mma + turn-around) / 2; the split is ~350 of those cycles).
template <int NCH>
__device__ __forceinline__ void split_operand(float (&x)[NCH][4], float (&l)[NCH][4]) {
#pragma unroll
  for (int i = 0; i < NCH; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { float h; split_tf32(x[i][j], h, l[i][j]); x[i][j] = h; }
}
template <bool TRANS, int NCH>
__device__ __forceinline__ void store_split(const float (&h)[NCH][4], const float (&l)[NCH][4], uint32_t s_hi, uint32_t s_lo, int rows, int t) {
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    int r, c;
    chunk_rc<TRANS>(i, t, r, c);
    if (r >= rows) continue;
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(s_hi + off), "f"(h[i][0]), "f"(h[i][1]), "f"(h[i][2]), "f"(h[i][3]));
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(s_lo + off), "f"(l[i][0]), "f"(l[i][1]), "f"(l[i][2]), "f"(l[i][3]));
  }
}

// ---- MN-major staging: element (r, k) = src[(k0 + k) * ld + row0 + r], r contiguous in memory ----
// 16-byte chunk q = t + 128 i of the (32 k) x (4 NCH chunks per k-row) tile: k = q / CPR, 4 rows starting at 4 (q % CPR).
// A warp reads one k-row segment of 512 contiguous bytes; the chunk lands at group (r / 32), row k, 32-byte unit
// ((r / 8) & 3) ^ (k & 3), half (r / 4) & 1.
template <int NCH>
__device__ __forceinline__ void load_mn(float (&x)[NCH][4], const float* __restrict__ src, int ld, int row0, int row_end, int rows,
                                        int k0, int live, int t) {
  constexpr int CPR = NCH * 4;                                  // chunks per k-row: 32 (128 rows) or 64 (256 rows)
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int q = t + GROUP * i, k = q / CPR, r4 = 4 * (q % CPR);
    x[i][0] = x[i][1] = x[i][2] = x[i][3] = 0.f;
    if (r4 < rows && row0 + r4 < row_end && k < live) {          // rows, row_end are multiples of 4 on this path
      const float4 v = *reinterpret_cast<const float4*>(src + (size_t)(k0 + k) * ld + row0 + r4);
      x[i][0] = v.x; x[i][1] = v.y; x[i][2] = v.z; x[i][3] = v.w;
    }
  }
}
template <int NCH>
__device__ __forceinline__ void store_mn(const float (&h)[NCH][4], const float (&l)[NCH][4], uint32_t s_hi, uint32_t s_lo, int rows, int t) {
  constexpr int CPR = NCH * 4;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int q = t + GROUP * i, k = q / CPR, c4 = q % CPR;
    if (4 * c4 >= rows) continue;
    const uint32_t off = (uint32_t)(c4 >> 3) * 4096u + (uint32_t)k * 128u + (uint32_t)(((((c4 & 7) >> 1) ^ (k & 3)) << 5) | ((c4 & 1) << 4));
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(s_hi + off), "f"(h[i][0]), "f"(h[i][1]), "f"(h[i][2]), "f"(h[i][3]));
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(s_lo + off), "f"(l[i][0]), "f"(l[i][1]), "f"(l[i][2]), "f"(l[i][3]));
  }
}

// TB means "B is stored (N, K) row-major" (g.transB); !TB is the natural Keras (K, N).
template <bool TA, bool TB, bool VEC, bool PACKED>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const GemmArgs g, int bn, int stages, int tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_accum;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * bn;
  const int nvalid = g.dyn_rows ? *g.dyn_rows : 0x7fffffff;
  const bool r_dyn = g.dyn_rows && !g.valid_on_k;
  if (r_dyn && (m0 % g.seg_rows) >= nvalid) return;   // dead tile (uniform per CTA)
  // rows of this tile that are live (forward / dgrad with a device-side row count)
  const int m_end = r_dyn ? min(g.M, m0 - (m0 % g.seg_rows) + nvalid) : g.M;

  const KMap km = make_kmap(g, nvalid, blockIdx.z);
  const int nkb = km.nkb;
  // optional per-CTA phase stamps (tools/wgrad_timeline.py): [globaltimer, start, setup, loaders done, accum ready, end, nkb, smid]
  unsigned long long* dbg = (g.dbg && tid == 0) ? g.dbg + 8 + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (dbg) {
    unsigned long long gt; unsigned sm;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
    dbg[0] = gt; dbg[1] = clock64(); dbg[6] = (unsigned long long)nkb; dbg[7] = sm;
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) { g.dbg[0] = 0x600DC0DE600DC0DFull; g.dbg[1] = (unsigned long long)gridDim.x * gridDim.y * gridDim.z; }
  }

  // operands whose M / N index is contiguous in memory (transposed A, (K, N) row-major B) are staged MN-major when
  // 128-bit accesses are possible; an MN-major B tile is padded to whole 32-column groups
  constexpr bool A_MN = TA && VEC, B_MN = !TB && VEC && !PACKED;
  const uint32_t b_tile = B_MN ? (uint32_t)((bn + 31) & ~31) * 128u : (uint32_t)bn * 128u;
  const uint32_t stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // 3xFP16 variant (weight gradients with bounded operands): fp16 tiles are half the bytes, so the same allocation holds
  // more stages; K blocks stay 32 deep (the live-block enumeration is shared), 2 K steps of 16 per block
  const bool f16 = A_MN && B_MN && g.f16 != nullptr && g.f16->ok != 0;
  const uint32_t a_tile16 = 2u * 4096u, b_tile16 = (uint32_t)((bn + 63) / 64) * 4096u, stage16 = 2 * a_tile16 + 2 * b_tile16;
  const int stages16 = min(MAX_STAGES, (int)(((uint32_t)stages * stage_bytes) / stage16));
  const int nstg = f16 ? stages16 : stages;

  if (tid == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(smem_u32(&bar_full[s]), LOADER_WARPS / 2); mbar_init(smem_u32(&bar_empty[s]), 1); }
    mbar_init(smem_u32(&bar_accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == LOADER_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols) : "memory");
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
    if constexpr (A_MN && B_MN) {
      if (f16) {
        const float s_a = g.f16->s_a, s_b = g.f16->s_b;
        for (int kb = grp; kb < nkb; kb += 2) {
          int k0, live;
          kmap_get(km, kb, k0, live);
          float xa[8][4], xb[16][4];
          load_mn<8>(xa, g.A, g.lda, m0, m_end, BM, k0, live, t);          // rows k >= live come back as zeros
          if (!g.Bpk16) load_mn<16>(xb, g.B, g.ldb, n0, g.N, bn, k0, live, t);
          const int s = kb % nstg;
          const uint32_t ph = (uint32_t)(kb / nstg) & 1u;
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
          const uint32_t sa = smem0 + (uint32_t)s * stage16;
          if (g.Bpk16) {
            // the X operand was pre-split into fp16 tiles by wg_pack_x_kernel: one bulk copy per K block, no register work
            if (t == 0) {
              const uint32_t bytes = 2u * b_tile16;
              const unsigned char* srcp = g.Bpk16 + ((size_t)(k0 / BK) * gridDim.x + blockIdx.x) * bytes;
              const uint32_t bar = smem_u32(&bar_full[s]);
              asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa + 2 * a_tile16),
                           "l"(srcp), "r"(bytes), "r"(bar)
                           : "memory");
            }
          } else {
            store_mn16<16>(xb, s_b, sa + 2 * a_tile16, sa + 2 * a_tile16 + b_tile16, bn, t);
          }
          store_mn16<8>(xa, s_a, sa, sa + a_tile16, BM, t);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
        }
      }
    }
    for (int kb = grp; kb < (f16 ? 0 : nkb); kb += 2) {
      int k0, live;
      kmap_get(km, kb, k0, live);
      constexpr int NB_CH = PACKED ? 8 : 16;
      float xa[8][4], xb[NB_CH][4];
      if constexpr (A_MN) load_mn<8>(xa, g.A, g.lda, m0, m_end, BM, k0, live, t);
      else load_operand<TA, VEC, 8>(xa, g.A, g.lda, m0, m_end, BM, k0, live, t);
      // B tile rows are output columns n; element (n, k) = B(k, n)
      if constexpr (B_MN) load_mn<NB_CH>(xb, g.B, g.ldb, n0, g.N, bn, k0, live, t);
      else if constexpr (!PACKED) load_operand<!TB, VEC, NB_CH>(xb, g.B, g.ldb, n0, g.N, bn, k0, live, t);
      float la[8][4];
      split_operand<8>(xa, la);     // waits for the A loads, not for the stage
      const int s = kb % stages;
      const uint32_t ph = (uint32_t)(kb / stages) & 1u;
      mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
      const uint32_t sa = smem0 + (uint32_t)s * stage_bytes;
      if constexpr (PACKED) {
        // weights: one elected thread pulls the pre-split, pre-swizzled hi|lo tile pair with a
        // single bulk async copy (UBLKCP) that completes on the stage's full barrier
        if (t == 0) {
          const uint32_t bytes = 2u * (uint32_t)bn * 128u;
          const unsigned char* srcp = g.Bpacked + ((size_t)blockIdx.x * (size_t)((g.K + BK - 1) / BK) + (size_t)(k0 / BK)) * bytes;
          const uint32_t bar = smem_u32(&bar_full[s]);
          asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa + 2 * A_TILE_BYTES),
                       "l"(srcp), "r"(bytes), "r"(bar)
                       : "memory");
        }
      }
      if constexpr (A_MN) store_mn<8>(xa, la, sa, sa + A_TILE_BYTES, BM, t);
      else store_split<TA, 8>(xa, la, sa, sa + A_TILE_BYTES, BM, t);
      if constexpr (B_MN) {
        float lb[NB_CH][4];
        split_operand<NB_CH>(xb, lb);
        store_mn<NB_CH>(xb, lb, sa + 2 * A_TILE_BYTES, sa + 2 * A_TILE_BYTES + b_tile, bn, t);
      } else if constexpr (!PACKED) {
        store_operand<!TB, NB_CH>(xb, sa + 2 * A_TILE_BYTES, sa + 2 * A_TILE_BYTES + b_tile, bn, t);
      }
      fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
    }
    // ------------------------------ epilogue ------------------------------
    if (dbg) dbg[3] = clock64();
    if (nkb > 0) {
      mbar_wait(smem_u32(&bar_accum), 0);
      tc_fence_after();
    }
    if (dbg) dbg[4] = clock64();
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may read
    const int m = m0 + quarter * 32 + lane;
    const bool m_ok = m < m_end;
    float* Cout = g.C + (g.splitk > 1 ? (size_t)blockIdx.z * g.split_stride : 0);
    const bool vec_st = !g.transC && (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(Cout) & 15) == 0) && (n0 % 4 == 0);
    // every epilogue operand that is present is 16-byte addressable per 16-column chunk?
    auto al16 = [](const void* p, int ld) { return p == nullptr || (((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 4 == 0)); };
    const bool fast_epi = al16(g.bias, 4) && al16(g.add, g.ldadd) && al16(g.mul, g.ldmul) && al16(g.gate, g.ldgate) &&
                          (!g.accumulate || g.transC || vec_st);
    for (int ci = warp >> 2; ci < bn / 16; ci += 2) {
      float v[16];
      if (nkb > 0) {
        float w[16];
        const uint32_t ta = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ci * 16);
        tmem_ld16(ta, v);
        tmem_ld16(ta + (uint32_t)bn, w);
        if (f16) {      // correction passes carry lo' = lo * 2^11; the accumulators are in units of s_a * s_b
          const float inv = g.f16->inv_unit;
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = (v[j] + w[j] * (1.f / 2048.f)) * inv;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += w[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
      if (!m_ok) continue;
      const int nb = n0 + ci * 16;
      const bool full = nb + 15 < g.N;
      if (g.splitk <= 1) {
        if (full && fast_epi) {
          // chunk-level epilogue: every option is one uniform branch + 128-bit loads
          if (g.bias) {
            const float4* b4 = reinterpret_cast<const float4*>(g.bias + nb);
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float4 b = b4[q]; v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w; }
          }
          if (g.add) {
            const float4* a4 = reinterpret_cast<const float4*>(g.add + (size_t)(g.addidx ? g.addidx[m] : m) * g.ldadd + nb);
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float4 b = a4[q]; v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w; }
          }
          if (g.act == ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = lrelu_f(v[j], g.alpha);
          } else if (g.act == ACT_LRELU2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = lrelu_f(lrelu_f(v[j], 0.01f), 0.2f);
          }
          if (g.mul) {
            const float4* m4 = reinterpret_cast<const float4*>(g.mul + (size_t)m * g.ldmul + nb);
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float4 b = m4[q]; v[4 * q] *= b.x; v[4 * q + 1] *= b.y; v[4 * q + 2] *= b.z; v[4 * q + 3] *= b.w; }
          }
          if (g.gate) {
            const float4* g4 = reinterpret_cast<const float4*>(g.gate + (size_t)m * g.ldgate + nb);
            const float ga = g.gate_alpha;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b = g4[q];
              v[4 * q] *= b.x > 0.f ? 1.f : ga; v[4 * q + 1] *= b.y > 0.f ? 1.f : ga;
              v[4 * q + 2] *= b.z > 0.f ? 1.f : ga; v[4 * q + 3] *= b.w > 0.f ? 1.f : ga;
            }
          }
          if (g.accumulate) {
            if (g.transC) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += Cout[(size_t)(nb + j) * g.ldc + m];
            } else {
              const float4* c4 = reinterpret_cast<const float4*>(Cout + (size_t)m * g.ldc + nb);
#pragma unroll
              for (int q = 0; q < 4; ++q) { const float4 b = c4[q]; v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w; }
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (nb + j < g.N) v[j] = gemm_epilogue(g, v[j], m, nb + j);
        }
      }
      if (g.transC) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (nb + j < g.N) Cout[(size_t)(nb + j) * g.ldc + m] = v[j];     // lanes = consecutive m: coalesced
      } else {
        float* crow = Cout + (size_t)m * g.ldc + nb;
        if (vec_st && full) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(crow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (nb + j < g.N) crow[j] = v[j];
        }
      }
    }
    if (dbg) dbg[5] = clock64();
  } else {
    // ------------------------------ MMA issuer (one lane) ------------------------------
    if (lane == 0 && nkb > 0 && f16) {
      const uint32_t idesc = make_idesc_f16(bn) | (1u << 15) | (1u << 16);   // both operands MN-major
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % nstg;
        const uint32_t ph = (uint32_t)(kb / nstg) & 1u;
        int k0, live;
        kmap_get(km, kb, k0, live);
        mbar_wait(smem_u32(&bar_full[s]), ph);
        tc_fence_after();
        const uint32_t sa = smem0 + (uint32_t)s * stage16, sb = sa + 2 * a_tile16;
        const uint64_t a_hi = make_desc_mn16(sa), a_lo = make_desc_mn16(sa + a_tile16);
        const uint64_t b_hi = make_desc_mn16(sb), b_lo = make_desc_mn16(sb + b_tile16);
        // two K steps of 16 k-rows (one when at most 16 rows of the block are live); corr accumulator at +bn
        if (live > 16) mma_f16_dual_mn<2>(tmem_d, tmem_d + (uint32_t)bn, a_hi, a_lo, b_hi, b_lo, idesc, kb ? 1u : 0u);
        else mma_f16_dual_mn<1>(tmem_d, tmem_d + (uint32_t)bn, a_hi, a_lo, b_hi, b_lo, idesc, kb ? 1u : 0u);
        mma_commit(smem_u32(&bar_empty[s]));
      }
      mma_commit(smem_u32(&bar_accum));
    } else if (lane == 0 && nkb > 0) {
      const uint32_t idesc = make_idesc_major(bn, A_MN, B_MN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (uint32_t)(kb / stages) & 1u;
        int k0, live;
        kmap_get(km, kb, k0, live);
        mbar_wait(smem_u32(&bar_full[s]), ph);
        tc_fence_after();
        const uint32_t sa = smem0 + (uint32_t)s * stage_bytes, sb = sa + 2 * A_TILE_BYTES;
        const uint64_t a_hi = A_MN ? make_desc_mn(sa) : make_desc(sa), a_lo = A_MN ? make_desc_mn(sa + A_TILE_BYTES) : make_desc(sa + A_TILE_BYTES);
        const uint64_t b_hi = B_MN ? make_desc_mn(sb) : make_desc(sb), b_lo = B_MN ? make_desc_mn(sb + b_tile) : make_desc(sb + b_tile);
        const int ksteps = (live + 7) / 8;
        for (int j = 0; j < ksteps; ++j) {
          // one K step = 8 tf32: +32 bytes along a K-major row, +8 k-rows (1024 bytes) in an MN-major tile (>>4 units)
          const uint64_t adva = (uint64_t)(A_MN ? j * 64 : j * 2), advb = (uint64_t)(B_MN ? j * 64 : j * 2);
          const uint32_t acc = (kb | j) ? 1u : 0u;
          // timing probe (tools/gemm_probe.py): 1 of the 3 passes only => how much of the tile time is MMA-bound
          if (g.probe == 2) { mma_tf32(tmem_d, a_hi + adva, b_hi + advb, idesc, acc); continue; }
          mma_tf32(tmem_d + (uint32_t)bn, a_lo + adva, b_hi + advb, idesc, acc);   // correction accumulator
          mma_tf32(tmem_d + (uint32_t)bn, a_hi + adva, b_lo + advb, idesc, 1u);
          mma_tf32(tmem_d, a_hi + adva, b_hi + advb, idesc, acc);                  // main accumulator
        }
        mma_commit(smem_u32(&bar_empty[s]));                 // stage reusable once these MMAs retire
      }
      mma_commit(smem_u32(&bar_accum));                      // accumulator complete
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == LOADER_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
  }
}

// Work items of the persistent kernel: (split, m tile, n tile), n fastest so that CTAs running
// concurrently share A row tiles in L2.
struct Item { int m0, n0, split; bool live; int m_end; KMap km; };
__device__ __forceinline__ Item decode_item(const GemmArgs& g, int item, int ntm, int ntn, int bn, int nvalid) {
  Item it;
  const int nt = item % ntn, rest = item / ntn;
  const int mt = rest % ntm;
  it.split = rest / ntm;
  it.m0 = mt * BM; it.n0 = nt * bn;
  const bool r_dyn = g.dyn_rows && !g.valid_on_k;
  it.live = !(r_dyn && (it.m0 % g.seg_rows) >= nvalid);
  it.m_end = r_dyn ? min(g.M, it.m0 - (it.m0 % g.seg_rows) + nvalid) : g.M;
  it.km = make_kmap(g, nvalid, it.split);
  return it;
}


// ---------------------------------------------------------------------------------------------
// Persistent variant for the packed-weight path (forward and dgrad GEMMs: ~70% of the step's
// tensor work).  grid = min(#tiles, #SMs); every role walks the same static tile list
// (tile = blockIdx.x + i * gridDim.x, n fastest).  Roles: warps 0-7 stage A (two groups alternate
// K blocks) and one elected thread per group bulk-copies the packed B tiles; warps 8-15 are
// dedicated epilogue warps (two per TMEM lane quarter, alternating 32-column groups); warp 16
// issues the MMAs.  The smem ring and its mbarrier phases run across tiles, so the loaders prefetch
// the next tile while the epilogue drains TMEM; the epilogue transposes each 32x32 block through
// a per-warp smem tile so that every global access of the fused epilogue (bias, gathered add,
// Hadamard, LeakyReLU', accumulate, store) is a full 128-byte line per row.
// ---------------------------------------------------------------------------------------------
constexpr int P_SPLIT_WARPS = 8;                 // 2 groups x 4 warps: smem -> smem hi/lo split of the TMA-loaded A tiles
constexpr int P_EPI_WARPS = 8;                   // two per TMEM lane quarter
constexpr int P_MMA_WARP = P_SPLIT_WARPS + P_EPI_WARPS;     // leader: MMA issue; peer: "my stage is full" relay
constexpr int P_TMA_WARP = P_MMA_WARP + 1;                  // one lane: A tiles by TMA, B halves by bulk copy
constexpr int P_THREADS = (P_SPLIT_WARPS + P_EPI_WARPS + 2) * 32;
constexpr int P_STG_BYTES = 32 * 36 * 4;          // per-warp epilogue staging tile, row pitch 36 floats (conflict free)

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// pair item -> this CTA's tile: the pair shares the n tile, CTA rank r takes m tile 2*mp + r
__device__ __forceinline__ Item decode_pair(const GemmArgs& g, int item, int rank, int ntm, int ntn, int bn, int nvalid) {
  Item it;
  const int nt = item % ntn, mp = item / ntn;
  it.split = 0;
  it.n0 = nt * bn;
  const bool r_dyn = g.dyn_rows && !g.valid_on_k;
  const int mA = (2 * mp) * BM, mB = (2 * mp + 1) * BM;
  // the pair runs if either tile has a live row (both CTAs must take the same decision)
  const bool liveA = !(r_dyn && (mA % g.seg_rows) >= nvalid), liveB = mB < g.M && !(r_dyn && (mB % g.seg_rows) >= nvalid);
  it.live = liveA || liveB;
  it.m0 = rank ? mB : mA;
  it.m_end = r_dyn ? min(g.M, it.m0 - (it.m0 % g.seg_rows) + nvalid) : g.M;
  it.km = make_kmap(g, nvalid, 0);
  return it;
}

// ---------------------------------------------------------------------------------------------
// Persistent cta_group::2 GEMM for the packed-weight path (forward and dgrad GEMMs).
//   * the two CTAs of a cluster form one 256 x bn MMA; each SM holds its own 128 rows of A and only
//     its half of the B tile, which lifts the SS-operand-fetch bound of cta_group::1;
//   * tiles are <= 128 wide so that TWO accumulator buffers (main + correction each) fit in the 512
//     TMEM columns: the epilogue of tile i overlaps the MMAs of tile i+1;
//   * A is loaded by TMA (cp.async.bulk.tensor, 128B hardware swizzle = the UMMA layout) into a raw
//     ring by one producer lane, so no global-load latency sits in a compute warp; two splitter groups
//     turn each raw tile into the hi / lo tiles smem -> smem (pure 16-byte-chunk map, the swizzle is
//     position preserving); B halves arrive by bulk copy from the packed image;
//   * the epilogue transposes 32x32 blocks through per-warp smem tiles: every global access of the
//     fused epilogue is a full 128-byte line per row.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t make_idesc2(int n) {   // M = 256 (m_dim = 16)
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit2_mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((unsigned short)3)
               : "memory");
}
// arrive on the barrier at the same offset in cluster rank 0 (local for the leader itself)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(bar));
  // default semantics (as CUTLASS's ClusterBarrier::arrive(cta_id)): with an explicit .release.cluster ptxas puts
  // MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of every arrive -- ~1.4k cycles per relayed stage, which is what starved the
  // pair kernels of operands (profiles/r2_pair_relay_probe.txt).  What the waiter consumes was written by bulk copies /
  // tcgen05 and is ordered by their own completion mechanisms, not by this thread's generic-proxy writes.
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}

__global__ void __launch_bounds__(P_THREADS, 1) gemm_tc_pair_kernel(const GemmArgs g, const __grid_constant__ CUtensorMap tmapA, int bn,
                                                                   int stages, int tmem_cols, int ntm, int ntn, int nitems) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_pfull[MAX_STAGES], bar_empty[MAX_STAGES], bar_afull[MAX_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t crank;                                  // rank in the 2-CTA cluster
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int nvalid = g.dyn_rows ? *g.dyn_rows : 0x7fffffff;
  const uint32_t hb = (uint32_t)(bn / 2) * 128u;      // this CTA's half of a B tile (bytes)
  const uint32_t stage_bytes = 2 * A_TILE_BYTES + 2 * hb;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg0 = smem0 + (uint32_t)stages * stage_bytes;                 // epilogue staging tiles

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), P_SPLIT_WARPS / 2);   // the splitter group's 4 warps (+ this CTA's B bytes)
      mbar_init(smem_u32(&bar_pfull[s]), 1);                  // leader only: "the peer's stage is full too"
      mbar_init(smem_u32(&bar_empty[s]), 1);                  // the pair's MMAs retired (multicast commit)
      mbar_init(smem_u32(&bar_afull[s]), 1);                  // producer's arrive.expect_tx + the A tile's TMA bytes
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tfull[b]), 1);
      mbar_init(smem_u32(&bar_tempty[b]), 2 * P_EPI_WARPS);   // leader: both CTAs' epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == P_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  // peers must not signal our barriers before they are initialised
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (warp < P_SPLIT_WARPS) {
    // ------------------------------ splitters: A_lo = x - trunc_tf32(x) ------------------------------
    // TMA lands the raw fp32 tile directly in the stage's A_hi slot: kind::tf32 ignores the low 13 mantissa
    // bits of its operands, so the raw tile IS the hi operand; only the lo tile is computed here
    // (smem -> smem, a pure 16-byte-chunk map because the 128B swizzle is position preserving).
    const int grp = warp >> 2, t = tid & (GROUP - 1);          // group g owns the K blocks q = g (mod 2)
    int gk = 0;
    for (int item = cluster_id; item < nitems; item += nclusters) {
      const Item it = decode_pair(g, item, (int)crank, ntm, ntn, bn, nvalid);
      if (!it.live) continue;
      const int nkb = it.km.nkb;
      for (int kb = 0; kb < nkb; ++kb) {
        const int q = gk + kb;
        if ((q & 1) != grp) continue;
        const int s = q % stages;
        mbar_wait(smem_u32(&bar_afull[s]), (uint32_t)(q / stages) & 1u);           // TMA landed the A tile of this stage
        const uint32_t sa = smem0 + (uint32_t)s * stage_bytes;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t off = (uint32_t)(t + GROUP * i) * 16u;
          float4 x;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(sa + off));
          const float l0 = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u), l1 = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
          const float l2 = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u), l3 = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sa + A_TILE_BYTES + off), "f"(l0), "f"(l1), "f"(l2), "f"(l3));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
      }
      gk += nkb;
    }
  } else if (warp < P_MMA_WARP) {
    // ------------------------------ epilogue warps ------------------------------
    const int ew = warp - P_SPLIT_WARPS;
    const int quarter = warp & 3, half = ew >> 2;
    const uint32_t stg = stg0 + (uint32_t)ew * P_STG_BYTES;
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    const int ngroups = (bn + 31) / 32;
    const int last_g = ngroups - 1 - half >= 0 ? ((ngroups - 1 - half) / 2) * 2 + half : -1;
    int ti = 0;
    for (int item = cluster_id; item < nitems; item += nclusters) {
      const Item it = decode_pair(g, item, (int)crank, ntm, ntn, bn, nvalid);
      if (!it.live) continue;
      mbar_wait(smem_u32(&bar_tfull[ti & 1]), (uint32_t)(ti >> 1) & 1u);
      tc_fence_after();
      if (g.dbg && blockIdx.x == 0 && ew == 0 && lane == 0 && ti < 16) g.dbg[ti * 8 + 4] = clock64();
      if (last_g < 0) {                             // narrow tile: this warp has no column group
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(smem_u32(&bar_tempty[ti & 1]));
      }
      const uint32_t tbuf = tmem_d + (uint32_t)((ti & 1) * 2 * bn);   // this tile's accumulator buffer
      for (int gi = half; gi < ngroups; gi += 2) {
        const int col0 = gi * 32, gw = min(32, bn - col0);
        float v[32];
        const bool estamp = g.dbg && blockIdx.x == 0 && ew == 0 && lane == 0 && ti == 2 && gi == half;
        unsigned long long e0 = 0, e1 = 0, e2 = 0;
        if (estamp) e0 = clock64();
        {
          float w[32];
          // TMEM columns of a buffer (N = 2*bn, CTA0's rows then CTA1's): [main n<bn/2 | corr n<bn/2 | main n>=bn/2 | corr n>=bn/2]
          const int hh = bn / 2;
          const int mcol = (col0 / hh) * bn + (col0 % hh);
          const uint32_t ta = tbuf + ((uint32_t)(quarter * 32) << 16) + (uint32_t)mcol;
          if (gw == 32) { tmem_ld32_nowait(ta, v); tmem_ld32_nowait(ta + (uint32_t)hh, w); }
          else { tmem_ld16_nowait(ta, v); tmem_ld16_nowait(ta + (uint32_t)hh, w); }
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < gw) v[j] += w[j];
        }
        if (estamp) { if (v[0] == 123.456f) e1 = 1; e1 += clock64(); }
        if (gi == last_g) {                          // this warp is done with the accumulator buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(smem_u32(&bar_tempty[ti & 1]));
        }
        // lane-per-row -> smem
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (q * 4 < gw)
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stg + (uint32_t)lane * 144u + (uint32_t)q * 16u), "f"(v[4 * q]),
                         "f"(v[4 * q + 1]), "f"(v[4 * q + 2]), "f"(v[4 * q + 3])
                         : "memory");
        __syncwarp();
        if (estamp) e2 = clock64();
        // row-contiguous: 8 lanes x 16 B = one 128-byte line per row, 4 rows per instruction
        const int nb = it.n0 + col0 + c4;
        if (c4 < gw && nb < g.N) {
          const bool colfull = nb + 3 < g.N;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = rsub + 4 * i;
            const int mm = it.m0 + quarter * 32 + r;
            if (mm >= it.m_end) continue;
            float4 x;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(stg + (uint32_t)r * 144u + (uint32_t)c4 * 4u));
            if (colfull) {
              if (g.bias) { const float4 b = *reinterpret_cast<const float4*>(g.bias + nb); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
              if (g.add) {
                const float4 b = *reinterpret_cast<const float4*>(g.add + (size_t)(g.addidx ? g.addidx[mm] : mm) * g.ldadd + nb);
                x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
              }
              if (g.act == ACT_LRELU) { x.x = lrelu_f(x.x, g.alpha); x.y = lrelu_f(x.y, g.alpha); x.z = lrelu_f(x.z, g.alpha); x.w = lrelu_f(x.w, g.alpha); }
              else if (g.act == ACT_LRELU2) {
                x.x = lrelu_f(lrelu_f(x.x, 0.01f), 0.2f); x.y = lrelu_f(lrelu_f(x.y, 0.01f), 0.2f);
                x.z = lrelu_f(lrelu_f(x.z, 0.01f), 0.2f); x.w = lrelu_f(lrelu_f(x.w, 0.01f), 0.2f);
              }
              if (g.mul) { const float4 b = *reinterpret_cast<const float4*>(g.mul + (size_t)mm * g.ldmul + nb); x.x *= b.x; x.y *= b.y; x.z *= b.z; x.w *= b.w; }
              if (g.gate) {
                const float4 b = *reinterpret_cast<const float4*>(g.gate + (size_t)mm * g.ldgate + nb);
                const float ga = g.gate_alpha;
                x.x *= b.x > 0.f ? 1.f : ga; x.y *= b.y > 0.f ? 1.f : ga; x.z *= b.z > 0.f ? 1.f : ga; x.w *= b.w > 0.f ? 1.f : ga;
              }
              if (g.accumulate) { const float4 b = *reinterpret_cast<const float4*>(g.C + (size_t)mm * g.ldc + nb); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
              *reinterpret_cast<float4*>(g.C + (size_t)mm * g.ldc + nb) = x;
            } else {
              const float* xe = &x.x;
              for (int e = 0; e < 4; ++e) if (nb + e < g.N) g.C[(size_t)mm * g.ldc + nb + e] = gemm_epilogue(g, xe[e], mm, nb + e);
            }
          }
        }
        __syncwarp();
        if (estamp) { g.dbg[300] = e0; g.dbg[301] = e1; g.dbg[302] = e2; g.dbg[303] = clock64(); }
      }
      if (g.dbg && blockIdx.x == 0 && ew == 0 && lane == 0 && ti < 16) g.dbg[ti * 8 + 5] = clock64();
      ++ti;
    }
  } else if (warp == P_MMA_WARP) {
    // ------------------------------ MMA warp ------------------------------
    if (lane == 0 && crank == 0) {
      // leader: one lane issues the pair's MMAs (M = 256 across the two CTAs)
      // each K step = 2 instructions of N = 2*bn against the CTA pair's contiguous [B_hi | B_lo] rows:
      //   A_hi x [B_hi | B_lo] -> [main | corr],   A_lo x [B_hi | B_lo] -> [main += lo*hi | corr += lo*lo (2^-24, harmless)]
      // (round 1 measured ~145 cycles per instruction "almost independent of N" here and chose two wide instructions over
      // three narrow ones; round 2's tools/mma_issue.cu shows that figure was this lane's own per-instruction overhead --
      // the tensor core needs N / 2 cycles -- see mma_f16_dual_k for the grouped form the default kernels now use.)
      // main sees 2 truncating adds per K step instead of 1.
      const uint32_t idesc = make_idesc2(2 * bn);
      int gk = 0, ti = 0;
      for (int item = cluster_id; item < nitems; item += nclusters) {
        const Item it = decode_pair(g, item, 0, ntm, ntn, bn, nvalid);
        if (!it.live) continue;
        const int nkb = it.km.nkb;
        unsigned long long t0 = 0, t1 = 0, t2 = 0;
        if (g.dbg && blockIdx.x == 0) t0 = clock64();
        mbar_wait(smem_u32(&bar_tempty[ti & 1]), ((uint32_t)(ti >> 1) & 1u) ^ 1u);   // both CTAs drained this buffer's previous tile
        tc_fence_after();
        if (g.dbg && blockIdx.x == 0) t1 = clock64();
        const uint32_t tbuf = tmem_d + (uint32_t)((ti & 1) * 2 * bn);
        for (int kb = 0; kb < nkb; ++kb) {
          const int q = gk + kb;
          const int s = q % stages;
          const uint32_t ph = (uint32_t)(q / stages) & 1u;
          int k0, live;
          kmap_get(it.km, kb, k0, live);
          mbar_wait(smem_u32(&bar_full[s]), ph);
          mbar_wait(smem_u32(&bar_pfull[s]), ph);
          tc_fence_after();
          if (g.dbg && blockIdx.x == 0 && kb == 0) t2 = clock64();
          const uint32_t sa = smem0 + (uint32_t)s * stage_bytes;
          const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_TILE_BYTES);
          const uint64_t b_all = make_desc(sa + 2 * A_TILE_BYTES);   // bn rows per CTA: [hi half | lo half]
          const int ksteps = (live + 7) / 8;
          for (int j = 0; j < ksteps; ++j) {
            const uint64_t adv = (uint64_t)(j * 2);
            const uint32_t acc = (kb | j) ? 1u : 0u;
            mma_tf32_2(tbuf, a_hi + adv, b_all + adv, idesc, acc);
            mma_tf32_2(tbuf, a_lo + adv, b_all + adv, idesc, 1u);
          }
          mma_commit2_mc(smem_u32(&bar_empty[s]));           // frees the stage in both CTAs
        }
        mma_commit2_mc(smem_u32(&bar_tfull[ti & 1]));        // accumulators complete in both CTAs
        if (g.dbg && blockIdx.x == 0 && ti < 16) {
          g.dbg[ti * 8 + 0] = t0; g.dbg[ti * 8 + 1] = t1; g.dbg[ti * 8 + 2] = t2; g.dbg[ti * 8 + 3] = clock64();
        }
        gk += nkb; ++ti;
      }
    } else if (lane == 0) {
      // peer: relay "my stage is full" to the leader's MMA thread
      int gk = 0;
      for (int item = cluster_id; item < nitems; item += nclusters) {
        const Item it = decode_pair(g, item, 1, ntm, ntn, bn, nvalid);
        if (!it.live) continue;
        const int nkb = it.km.nkb;
        for (int kb = 0; kb < nkb; ++kb) {
          const int q = gk + kb;
          const int s = q % stages;
          mbar_wait(smem_u32(&bar_full[s]), (uint32_t)(q / stages) & 1u);
          mbar_arrive_leader(smem_u32(&bar_pfull[s]));
        }
        gk += nkb;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ producer: A by TMA, B halves by bulk copy ------------------------------
    if (lane == 0) {
      int gk = 0;
      for (int item = cluster_id; item < nitems; item += nclusters) {
        const Item it = decode_pair(g, item, (int)crank, ntm, ntn, bn, nvalid);
        if (!it.live) continue;
        const int nkb = it.km.nkb;
        for (int kb = 0; kb < nkb; ++kb) {
          const int q = gk + kb;
          const int s = q % stages;
          int k0, live;
          kmap_get(it.km, kb, k0, live);
          // the stage is free once the pair's MMAs of its previous use retired
          mbar_wait(smem_u32(&bar_empty[s]), ((uint32_t)(q / stages) & 1u) ^ 1u);
          const uint32_t sa = smem0 + (uint32_t)s * stage_bytes;
          {
            // B halves: their expect_tx is posted before the A tile is requested, hence before the
            // splitters (who wait for the A tile) can arrive on full[s]
            const unsigned char* tile = g.Bpacked + ((size_t)(it.n0 / bn) * (size_t)((g.K + BK - 1) / BK) + (size_t)(k0 / BK)) * (size_t)(4u * hb);
            const uint32_t bar = smem_u32(&bar_full[s]);
            asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(2u * hb) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa + 2 * A_TILE_BYTES),
                         "l"(tile + (size_t)crank * hb), "r"(hb), "r"(bar)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa + 2 * A_TILE_BYTES + hb),
                         "l"(tile + 2u * (size_t)hb + (size_t)crank * hb), "r"(hb), "r"(bar)
                         : "memory");
          }
          {
            // A: 128 rows x 32 floats at (row m0, column k0) straight into the A_hi slot; OOB rows / columns are zero
            const uint32_t bar = smem_u32(&bar_afull[s]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(A_TILE_BYTES) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sa),
                         "l"(&tmapA), "r"(bar), "r"(k0), "r"(it.m0)
                         : "memory");
          }
        }
        gk += nkb;
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  // do not exit (or free TMEM) while the peer may still signal our barriers
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == P_MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
  }
}

// Pack a weight matrix B (K x N; element (k,n) = transB ? W[n*ld + k] : W[k*ld + n]) into the
// exact shared-memory image the MMA wants: for every (n tile j, K block kb) a contiguous
// [hi tile | lo tile] pair, each bn rows x 128 bytes, K-major, 128-byte swizzled, zero padded.
// Runs once per weight update (the weights change once per train step); every CTA of every GEMM
// that uses the matrix then fetches its B tiles with one bulk copy instead of staging them
// through registers.
__global__ void pack_b_kernel(const float* __restrict__ W, int ld, int transB, int K, int N, int bn, unsigned char* __restrict__ out) {
  const int nkb = (K + BK - 1) / BK, ntile = (N + bn - 1) / bn;
  const long long total = (long long)ntile * nkb * bn * 8;
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const int r = (int)(q % bn);
  const int c = (int)((q / bn) % 8);
  const int kb = (int)((q / ((long long)bn * 8)) % nkb);
  const int j = (int)(q / ((long long)bn * 8 * nkb));
  const int n = j * bn + r;
  float h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k = kb * BK + 4 * c + e;
    const float x = (n < N && k < K) ? (transB ? W[(size_t)n * ld + k] : W[(size_t)k * ld + n]) : 0.f;
    split_tf32(x, h[e], l[e]);
  }
  unsigned char* tile = out + ((size_t)j * nkb + kb) * (2u * (size_t)bn * 128u);
  const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
  *reinterpret_cast<float4*>(tile + off) = make_float4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<float4*>(tile + (size_t)bn * 128u + off) = make_float4(l[0], l[1], l[2], l[3]);
}

struct Plan { int bn, stages, tmem_cols; size_t smem; int stages_pers; size_t smem_pers; };
// cta_group::2 kernel: tiles <= 128 wide so that two accumulator buffers (main + correction each) fit in 512 TMEM columns
inline int pair_bn(int N) { return N <= 64 ? 64 : 128; }   // bn/2 must be a multiple of 32 (epilogue column groups)
inline bool al16h(const void* p, int ld) { return p == nullptr || (((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 4 == 0)); }
inline Plan plan_for(int N) {
  Plan p;
  const int ntiles = cdiv(N, 256);
  p.bn = round_up(cdiv(N, ntiles), 16);
  if (p.bn > 256) p.bn = 256;
  const int stage_bytes = 2 * (int)A_TILE_BYTES + 2 * p.bn * 128;
  p.stages = (SMEM_LIMIT - 1024) / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.bn) p.tmem_cols *= 2;    // main + correction accumulators
  p.smem = (size_t)p.stages * stage_bytes + 1024;
  p.stages_pers = (SMEM_LIMIT - 1024 - P_EPI_WARPS * P_STG_BYTES) / stage_bytes;
  if (p.stages_pers > MAX_STAGES) p.stages_pers = MAX_STAGES;
  p.smem_pers = (size_t)p.stages_pers * stage_bytes + 1024 + P_EPI_WARPS * P_STG_BYTES;
  return p;
}

}  // namespace tc

inline bool tc_gemm_supported(const GemmArgs& g) { return g.K >= 8 && g.N >= 32 && g.M >= 64 && g.act != ACT_TANH; }

template <bool TA, bool TB, bool VEC, bool PACKED>
inline cudaError_t tc_set_attr() {
  return cudaFuncSetAttribute(tc::gemm_tc_kernel<TA, TB, VEC, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT);
}
inline cudaError_t tc_gemm_init() {
  cudaError_t e;
#define GAC_TC_ATTR(a, b, c, d) if ((e = tc_set_attr<a, b, c, d>()) != cudaSuccess) return e;
  GAC_TC_ATTR(false, false, false, false) GAC_TC_ATTR(false, false, true, false) GAC_TC_ATTR(false, true, false, false)
  GAC_TC_ATTR(false, true, true, false) GAC_TC_ATTR(true, false, false, false) GAC_TC_ATTR(true, false, true, false)
  GAC_TC_ATTR(true, true, false, false) GAC_TC_ATTR(true, true, true, false)
  GAC_TC_ATTR(false, false, false, true) GAC_TC_ATTR(false, false, true, true)
#undef GAC_TC_ATTR
  if ((e = cudaFuncSetAttribute(tc::gemm_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT)) != cudaSuccess) return e;
  return cudaSuccess;
}

// 2-D tensor map of a row-major fp32 matrix (rows x cols, leading dimension ld) with a 32-column x
// 128-row box and the 128-byte swizzle the UMMA K-major layout expects; out-of-bounds reads are zero.
// cuTensorMapEncodeTiled is fetched from the driver at run time (no link-time dependency on libcuda).
inline bool tc_encode_tmap_2d(CUtensorMap* out, const float* base, int rows, int cols, int ld) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_fn>(p);
  }
  if (!fn) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)tc::BM};
  const cuuint32_t estr[2] = {1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The cta_group::2 persistent kernel (TMEM double buffered) passes the whole GPU suite but is, as of
// round 1, slower than the one-tile-per-CTA kernel: its register-staged A operand is bound by global
// load latency (profiles/r1_pair_kernel_timeline.txt).  Opt in with GAC_TC_PAIR=1.
inline bool tc_pair_enabled() {
  static const bool on = getenv("GAC_TC_PAIR") != nullptr && getenv("GAC_TC_PAIR")[0] == '1';
  return on;
}
inline size_t tc_packed_bytes(int K, int N) {
  const tc::Plan p = tc::plan_for(N);
  const size_t a = (size_t)cdiv(N, p.bn) * cdiv(K, tc::BK) * 2 * p.bn * 128;
  const int pb = tc::pair_bn(N);
  const size_t b = (size_t)cdiv(N, pb) * cdiv(K, tc::BK) * 2 * pb * 128;
  return a > b ? a : b;
}
// bn = 0: the one-tile-per-CTA kernel's tile width; otherwise the given width (cta_group::2 kernel)
inline cudaError_t tc_pack_b(const float* W, int ld, int transB, int K, int N, unsigned char* out, cudaStream_t st, int bn = 0) {
  if (bn == 0) bn = tc::plan_for(N).bn;
  const long long total = (long long)cdiv(N, bn) * cdiv(K, tc::BK) * bn * 8;
  tc::pack_b_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(W, ld, transB, K, N, bn, out);
  return cudaGetLastError();
}

// X operand of a weight-gradient GEMM (activations, (nseg * seg rows, ncols) fp32, the first *live rows of every segment
// live) -> MN-major fp16 hi / lo' tiles, one [hi | lo] slot per (32-row K block, n-tile): exactly the shared-memory image the
// 3xFP16 path of gemm_tc_kernel<1,0,1,0> stages for its B operand, so the GEMM fetches it with one bulk copy per K block.
// Dead K blocks are skipped; rows beyond `live` inside the last live block are zero.
namespace tc {
__global__ void wg_pack_x_kernel(const float* __restrict__ X, int ldx, int ncols, int seg, const int* __restrict__ live, const WgScales* __restrict__ sc,
                                 int bn, int ntiles, unsigned char* __restrict__ out) {
  const int blk = blockIdx.x, sg = blockIdx.y, j = blockIdx.z;
  const int nvalid = live ? min(*live, seg) : seg;
  if (blk * BK >= nvalid || !sc->ok) return;
  const float scale = sc->s_b;
  const uint32_t b_tile16 = (uint32_t)((bn + 63) / 64) * 4096u;
  unsigned char* slot = out + ((size_t)((size_t)sg * (seg / BK) + blk) * ntiles + j) * (2 * (size_t)b_tile16);
  const int n0 = j * bn, cpr = bn / 4;                          // float4 chunks per k-row of this tile
  for (int idx = threadIdx.x; idx < BK * cpr; idx += blockDim.x) {
    const int k = idx / cpr, c4 = idx - k * cpr, n = n0 + 4 * c4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (blk * BK + k < nvalid && n < ncols) v = *reinterpret_cast<const float4*>(X + ((size_t)sg * seg + blk * BK + k) * ldx + n);
    uint32_t h0, h1, l0, l1;
    split_f16x2(v.x * scale, v.y * scale, h0, l0);
    split_f16x2(v.z * scale, v.w * scale, h1, l1);
    const uint32_t off = (uint32_t)(c4 >> 4) * 4096u + (uint32_t)k * 128u + (uint32_t)(((((c4 & 15) >> 1) ^ (k & 7)) << 4) | ((c4 & 1) << 3));
    *reinterpret_cast<uint2*>(slot + off) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(slot + b_tile16 + off) = make_uint2(l0, l1);
  }
}
}  // namespace tc
inline size_t wg_pack_x_bytes(long long rows, int ncols) {
  const tc::Plan p = tc::plan_for(ncols);
  return (size_t)(rows / tc::BK) * cdiv(ncols, p.bn) * 2 * (size_t)((p.bn + 63) / 64) * 4096;
}
inline cudaError_t wg_pack_x(const float* X, int ldx, int ncols, int seg, int nseg, const int* live, const WgScales* sc, unsigned char* out,
                             cudaStream_t st) {
  const tc::Plan p = tc::plan_for(ncols);
  dim3 grid(seg / tc::BK, nseg, cdiv(ncols, p.bn));
  tc::wg_pack_x_kernel<<<grid, 256, 0, st>>>(X, ldx, ncols, seg, live, sc, p.bn, cdiv(ncols, p.bn), out);
  return cudaGetLastError();
}

inline cudaError_t launch_gemm_tc(const GemmArgs& g_in, cudaStream_t st) {
  GemmArgs g = g_in;
  static const int probe = getenv("GAC_TC_PROBE") ? atoi(getenv("GAC_TC_PROBE")) : 0;
  g.probe = probe;
  const tc::Plan p = tc::plan_for(g.N);
  dim3 grid(cdiv(g.N, p.bn), cdiv(g.M, tc::BM), g.splitk > 1 ? g.splitk : 1);
  // 16-byte vector accesses need aligned bases and leading dimensions
  // (MN-major staging reads float4 over 4 consecutive rows / columns: their count must be a multiple of 4)
  const bool veca = (g.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0) && (!g.transA || g.M % 4 == 0);
  const bool vec = veca && (g.ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0) && (g.transB || g.Bpacked || g.N % 4 == 0);
  // an MN-major B tile is padded to whole 32-column groups (hi and lo tile of every stage)
  const size_t stage_l = 2 * (size_t)tc::A_TILE_BYTES + 2 * (size_t)((p.bn + 31) & ~31) * 128;
  int stages_l = p.stages;
  while (stages_l > 2 && stages_l * stage_l + 1024 > (size_t)tc::SMEM_LIMIT) --stages_l;
  const size_t smem_l = stages_l * stage_l + 1024;
#define GAC_TC_LAUNCH(a, b, c, d) tc::gemm_tc_kernel<a, b, c, d><<<grid, tc::THREADS, smem_l, st>>>(g, p.bn, stages_l, p.tmem_cols)
  const bool epi_al = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && tc::al16h(g.bias, 4) &&
                      tc::al16h(g.add, g.ldadd) && tc::al16h(g.mul, g.ldmul) && tc::al16h(g.gate, g.ldgate);
  // Opt-in (GAC_TC_PERSISTENT=1): validated by the full GPU suite, measured no faster than the
  // one-tile-per-CTA kernel because both are bound by shared-memory bandwidth; it is
  // the base for the cta_group::2 version.
  if (g.Bpacked && g.Bpacked_bn > 0) {
    // packed for the cta_group::2 persistent kernel (A by TMA: needs 16-byte aligned rows)
    if (g.transA || g.transC || g.splitk > 1 || !epi_al || !veca) return cudaErrorInvalidValue;
    static int nsm2 = 0;
    if (nsm2 == 0) {
      int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm2, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm2 <= 0) nsm2 = 148;
    }
    CUtensorMap tmap;
    if (!tc_encode_tmap_2d(&tmap, g.A, g.M, g.K, g.lda)) return cudaErrorNotSupported;
    const int pbn = g.Bpacked_bn;
    const int ntn = cdiv(g.N, pbn), ntm = cdiv(g.M, tc::BM), npairs = ntn * cdiv(ntm, 2);
    const int nclus = npairs < nsm2 / 2 ? npairs : nsm2 / 2;
    const int pstage = 2 * (int)tc::A_TILE_BYTES + 2 * (pbn / 2) * 128;
    const int fixed = 1024 + tc::P_EPI_WARPS * tc::P_STG_BYTES;
    int pst = (tc::SMEM_LIMIT - fixed) / pstage;
    if (pst > tc::MAX_STAGES) pst = tc::MAX_STAGES;
    if (pst < 2) return cudaErrorInvalidValue;
    const size_t psmem = (size_t)pst * pstage + fixed;
    int tcols = 32;
    while (tcols < 4 * pbn) tcols *= 2;                       // 2 buffers x (main + correction)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * nclus); cfg.blockDim = dim3(tc::P_THREADS); cfg.dynamicSmemBytes = psmem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, tc::gemm_tc_pair_kernel, g, tmap, pbn, pst, tcols, ntm, ntn, npairs);
  }
  if (false) {
  } else if (g.Bpacked && !g.transA && g.splitk <= 1) {
    if (veca) GAC_TC_LAUNCH(false, false, true, true); else GAC_TC_LAUNCH(false, false, false, true);
  } else if (g.transA) {
    if (g.transB) { if (vec) GAC_TC_LAUNCH(true, true, true, false); else GAC_TC_LAUNCH(true, true, false, false); }
    else { if (vec) GAC_TC_LAUNCH(true, false, true, false); else GAC_TC_LAUNCH(true, false, false, false); }
  } else {
    if (g.transB) { if (vec) GAC_TC_LAUNCH(false, true, true, false); else GAC_TC_LAUNCH(false, true, false, false); }
    else { if (vec) GAC_TC_LAUNCH(false, false, true, false); else GAC_TC_LAUNCH(false, false, false, false); }
  }
#undef GAC_TC_LAUNCH
  return cudaGetLastError();
}

}  // namespace gac
