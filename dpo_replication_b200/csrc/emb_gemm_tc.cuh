// CosineBasisLinear as ONE tcgen05 kernel (sm_100a):  C[r, :] = act( cos(x_r * i * pi)_{i=1..64} * W + b ) (* mul[r, :])
// (networks.py:17-62: the action embedding, LeakyReLU(0.2) outside, and the noise embedding, multiplied by the GRU
// state right after -- networks.py:208-210, 219-221).
//
// As two kernels (cosine basis, then a K = 64 GEMM with 208-wide tiles) this path spent 50-65 us per 32768 rows on 6 us of
// tensor-core work: every row block was visited twice (two N tiles), the basis went through HBM, and the stores dominate.
// Here a CTA owns 128 complete rows: the 128 x 64 basis is computed in registers (never stored), W (64 x 400, pre-split fp16
// hi / lo) arrives by bulk copy, 12 MMAs (4 K steps x 3 passes, one instruction group) accumulate the columns in TMEM, and the
// epilogue applies bias / LeakyReLU / the Hadamard factor and writes the rows.
// Round 2: a CTA owns 128 rows x ONE column half (208 or 192 of the 400): 87 KB of smem and 256 TMEM columns instead of 136 KB
// and 512, so TWO CTAs share an SM and one's serial chain (weight copy -> basis -> MMAs -> TMEM reads -> stores) runs under the
// other's; the basis is computed twice per row block, which costs less than an SM idling through every latency of the chain.
//
// Arithmetic: 3xFP16 with ONE accumulator.  |cos| <= 1 is scaled by 2^13 and W by a power of two from its scanned maximum
// (emb_scale_kernel); lo = x' - hi is kept UNSCALED so that all three passes share a unit -- its absolute precision
// (fp16 subnormal spacing 2^-24 against operands of ~2^13) is 2^-37 relative, and with only 12 MMAs per column the
// accumulate truncation (<= 12 x 2^-24) needs no separate correction accumulator.  Non-finite / absurd weights clear the
// device flag and the CTA computes its tile with plain fp32 FMAs instead.
#pragma once
#include "gru_step_v2.cuh"

namespace gac {
namespace tc {

struct EmbScales { float s_w, inv_unit; int ok; unsigned int amax[2]; int pad[3]; };

struct EmbArgs {
  int rows;                      // M
  const int* dyn_rows; int seg_rows;   // optional device row count (rows come in segments of seg_rows, the first *dyn_rows live)
  const float* x; int incx;      // x_r = x[r * incx]; nullptr = 0 (first action dimension)
  // step-major rows of the training pass (xseg > 0): row r = d * xseg + i reads x[i * incx + d + xoff], 0 when d + xoff < 0
  // (the action embedding of step d sees the teacher's dimension d - 1, the noise embedding tau's dimension d)
  int xseg, xoff;
  const unsigned char* wpk;      // emb_pack_kernel output: [hi | lo], 400 rows (n) x 128 bytes (64 k), K-major SWIZZLE_128B
  const EmbScales* sc;
  const float* W;                // (64, 400) fp32, for the fallback
  const float* bias;             // (400)
  int act; float alpha;          // 1: LeakyReLU(alpha)
  const float* mul; int ldmul;   // optional (rows, 400) Hadamard factor applied last
  float* C; int ldc;             // (rows, 400)
  float* C2;                     // optional (with mul): C keeps the embedding itself and C2 (same ldc) takes the Hadamard product
                                 // (the training pass needs both: networks.py:219-221 and its backward)
  // optional a16 image (gru_step_v2.cuh) of the result, scaled by *c16_scale: written INSTEAD of C while *c16_ok != 0
  // (c16_also: next to C -- the training pass needs the fp32 embedding for its backward)
  unsigned char* C16; const float* c16_scale; const int* c16_ok; int c16_also;
};

constexpr uint32_t EMB_W_TILE = (uint32_t)GAC_E * 128u;          // 51,200 B per hi / lo
constexpr uint32_t EMB_A_TILE = (uint32_t)BM * 128u;             // 16 KB per hi / lo
constexpr int EMB_N0 = 208, EMB_N1 = GAC_E - EMB_N0;              // the two column halves (multiples of 16)
constexpr uint32_t EMB_WH_TILE = (uint32_t)EMB_N0 * 128u;        // smem of one half's hi (or lo) weight rows
constexpr int EMB_SMEM = 2 * EMB_WH_TILE + 2 * EMB_A_TILE + 1024;   // 87,040 B: two CTAs per SM

__global__ void emb_scale_kernel(EmbScales* __restrict__ sc) {
  const float mw = __uint_as_float(sc->amax[0]);
  sc->ok = 0; sc->s_w = 1.f; sc->inv_unit = 1.f;
  if (mw > 0.f && mw < 1e30f) {
    int ew;
    frexpf(mw, &ew);
    if (abs(ew) < 100) { sc->s_w = ldexpf(1.f, 13 - ew); sc->inv_unit = ldexpf(1.f, ew - 26); sc->ok = 1; }   // |W'| < 2^13, A' = cos * 2^13
  }
}
// W (64, 400) row-major -> fp16 hi | lo (unscaled lo), B-operand image: row n, 16-byte chunk (k / 8) ^ (n & 7)
__global__ void emb_pack_kernel(const float* __restrict__ W, const EmbScales* __restrict__ sc, unsigned char* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= GAC_E * 8 || !sc->ok) return;
  const int c = idx & 7, n = idx >> 3;
  const float s = sc->s_w;
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float x0 = W[(size_t)(8 * c + 2 * e) * GAC_E + n] * s, x1 = W[(size_t)(8 * c + 2 * e + 1) * GAC_E + n] * s;
    const __half2 hh = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    h[e] = *reinterpret_cast<const uint32_t*>(&hh); l[e] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  const uint32_t off = (uint32_t)n * 128u + (uint32_t)((c ^ (n & 7)) << 4);
  *reinterpret_cast<uint4*>(out + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(out + EMB_W_TILE + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ float emb_x(const EmbArgs& g, int m) {
  if (!g.x) return 0.f;
  if (g.xseg <= 0) return g.x[(size_t)m * g.incx];
  const int d = m / g.xseg, i = m - d * g.xseg;
  return d + g.xoff < 0 ? 0.f : g.x[(size_t)i * g.incx + d + g.xoff];
}

__global__ void __launch_bounds__(THREADS, 2) emb_gemm_tc_kernel(const EmbArgs g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_w, bar_a, bar_acc;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[GAC_E];                  // the epilogue used to fetch its bias from L2 once per 16-column chunk, in the
                                                                 // middle of a TMEM-read -> add -> store chain (profiles/r1c_dot_emb_summary.txt)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chalf = (int)blockIdx.x & 1, m0 = ((int)blockIdx.x >> 1) * BM;   // this CTA's column half and row block
  const int n0 = chalf ? EMB_N0 : 0, ncols = chalf ? EMB_N1 : EMB_N0;
  int m_end = g.rows;
  if (g.dyn_rows) {
    const int nvalid = *g.dyn_rows;
    if (g.seg_rows > 0) {
      const int in_seg = m0 % g.seg_rows;
      if (in_seg >= nvalid) return;
      m_end = min(g.rows, m0 - in_seg + nvalid);
    } else {
      m_end = min(g.rows, nvalid);
    }
  }
  if (m0 >= m_end) return;
  const bool ok = g.sc->ok != 0;
  const bool use16 = g.C16 != nullptr && *g.c16_ok != 0;
  const float s16 = use16 ? *g.c16_scale : 1.f;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sw_hi = smem0, sw_lo = smem0 + EMB_WH_TILE, sa_hi = smem0 + 2 * EMB_WH_TILE, sa_lo = sa_hi + EMB_A_TILE;

  if (!ok) {
    // fallback: plain fp32 (one thread per (row, column) pair, strided); correct for any finite weights
    for (int i = tid; i < BM * ncols; i += THREADS) {
      const int r = i / ncols, n = n0 + (i - r * ncols), m = m0 + r;
      if (m >= m_end) continue;
      const float xv = emb_x(g, m);
      float acc = 0.f;
      for (int k = 0; k < GAC_NB; ++k) acc = fmaf(cosf(__fmul_rn(xv, c_ipi[k])), g.W[(size_t)k * GAC_E + n], acc);
      float v = acc + g.bias[n];
      if (g.act) v = v > 0.f ? v : g.alpha * v;
      if (g.mul) {
        const float v2 = v * g.mul[(size_t)m * g.ldmul + n];
        if (g.C2) g.C2[(size_t)m * g.ldc + n] = v2; else v = v2;
      }
      if (use16) {
        const __half hh = __float2half_rn(v * s16), ll = __float2half_rn(v * s16 - __half2float(hh));
        unsigned char* dst = g.C16 + a16_off(m, n & ~3) + (n & 3) * 2;
        *reinterpret_cast<__half*>(dst) = hh; *reinterpret_cast<__half*>(dst + G2_A_TILE) = ll;
      }
      if (!use16 || g.c16_also) g.C[(size_t)m * g.ldc + n] = v;
    }
    return;
  }

  for (int i = tid; i < GAC_E; i += THREADS) bias_s[i] = g.bias[i];
  if (tid == 0) {
    mbar_init(smem_u32(&bar_w), 1); mbar_init(smem_u32(&bar_a), LOADER_WARPS); mbar_init(smem_u32(&bar_acc), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == LOADER_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (warp < LOADER_WARPS) {
    if (tid == 0) {   // this half's rows of the pre-split weight image: hi rows and lo rows, one bulk copy each
      const uint32_t bar = smem_u32(&bar_w), bytes = (uint32_t)ncols * 128u;
      // the barrier has one pending arrival: this thread's arrive carries the expected byte count
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sw_hi),
                   "l"(g.wpk + (size_t)n0 * 128), "r"(bytes), "r"(bar)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sw_lo),
                   "l"(g.wpk + EMB_W_TILE + (size_t)n0 * 128), "r"(bytes), "r"(bar)
                   : "memory");
    }
    // ---- the 128 x 64 cosine basis, straight into the A operand: thread = (row, half of the 64 k) ----
    {
      const int r = tid >> 1, kh = tid & 1, m = m0 + r;
      const float xv = m < m_end ? emb_x(g, m) : 0.f;
      const bool live = m < m_end;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {            // 16-byte chunks (8 k each) of this thread's half row
        const int c = kh * 4 + cc;
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = 8 * c + 2 * e;
          const float x0 = live ? cosf(__fmul_rn(xv, c_ipi[k])) * 8192.f : 0.f, x1 = live ? cosf(__fmul_rn(xv, c_ipi[k + 1])) * 8192.f : 0.f;
          const __half2 hh = __floats2half2_rn(x0, x1);
          const float2 hf = __half22float2(hh);
          const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
          h[e] = *reinterpret_cast<const uint32_t*>(&hh); l[e] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sa_hi + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]));
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sa_lo + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_a));
    }
    // ---- epilogue.  The accumulators are read lane-per-row from TMEM; writing (and reading the Hadamard factor) that way
    // would touch 32 different rows with 16 bytes each per instruction.  Instead the sums go through the now idle weight /
    // basis smem in two column halves (208 + 192), and (row, 4 columns) items make every global access row-contiguous,
    // with the Hadamard loads of four items in flight. ----
    mbar_wait(smem_u32(&bar_acc), 0);
    tc_fence_after();
    const int quarter = warp & 3, hf = warp >> 2;
    const float inv = g.sc->inv_unit, alpha = g.alpha;
    float* stg = reinterpret_cast<float*>(smem_raw + (smem0 - smem_u32(smem_raw)));
    const int nchunks = ncols / 16;                               // 13 or 12 sixteen-column chunks in this CTA's half
#pragma unroll 1
    for (int cbeg = 0; cbeg < nchunks; cbeg += 7) {               // staged through the idle operand smem at most 7 chunks (59 KB) at a time
      const int cend = min(cbeg + 7, nchunks);
      const int wcols = (cend - cbeg) * 16, SS = wcols + 4;       // (wcols + 4) / 4 is odd: conflict-free 128-bit row stores
      const int nmine = (cend - cbeg + 1 - hf) / 2;               // chunks cbeg + hf, cbeg + hf + 2, ...
      float* srow = stg + (size_t)(quarter * 32 + lane) * SS;
      for (int j = 0; j < nmine; ++j) {
        const int ci = cbeg + hf + 2 * j;
        float v[16];
        tmem_ld16(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ci * 16), v);
        const float4* b4 = reinterpret_cast<const float4*>(bias_s + n0 + ci * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 bb = b4[q];
          float4 o = make_float4(v[4 * q] * inv + bb.x, v[4 * q + 1] * inv + bb.y, v[4 * q + 2] * inv + bb.z, v[4 * q + 3] * inv + bb.w);
          if (g.act) { o.x = o.x > 0.f ? o.x : alpha * o.x; o.y = o.y > 0.f ? o.y : alpha * o.y; o.z = o.z > 0.f ? o.z : alpha * o.z; o.w = o.w > 0.f ? o.w : alpha * o.w; }
          *reinterpret_cast<float4*>(srow + (ci - cbeg) * 16 + 4 * q) = o;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int n4 = wcols >> 2, total = BM * n4, col0 = n0 + cbeg * 16;
      constexpr int IB = 4;
      for (int i0 = tid; i0 < total; i0 += IB * LOADER_WARPS * 32) {
        float4 ml[IB];
        int row[IB], col[IB];
        bool okk[IB];
#pragma unroll
        for (int u = 0; u < IB; ++u) {
          const int i = i0 + u * LOADER_WARPS * 32;
          row[u] = i / n4;
          col[u] = 4 * (i - row[u] * n4);
          okk[u] = i < total && m0 + row[u] < m_end;
          if (okk[u] && g.mul) ml[u] = *reinterpret_cast<const float4*>(g.mul + (size_t)(m0 + row[u]) * g.ldmul + col0 + col[u]);
        }
#pragma unroll
        for (int u = 0; u < IB; ++u) {
          if (!okk[u]) continue;
          float4 x = *reinterpret_cast<const float4*>(stg + (size_t)row[u] * SS + col[u]);
          if (g.mul) {
            const float4 x2 = make_float4(x.x * ml[u].x, x.y * ml[u].y, x.z * ml[u].z, x.w * ml[u].w);
            if (g.C2) *reinterpret_cast<float4*>(g.C2 + (size_t)(m0 + row[u]) * g.ldc + col0 + col[u]) = x2; else x = x2;
          }
          if (use16) {                                             // the GRU step's A operand: split, scaled, swizzled here
            uint32_t h0, l0, h1, l1;
            split_f16x2_u(x.x * s16, x.y * s16, h0, l0);
            split_f16x2_u(x.z * s16, x.w * s16, h1, l1);
            unsigned char* dst = g.C16 + a16_off(m0 + row[u], col0 + col[u]);
            *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(dst + G2_A_TILE) = make_uint2(l0, l1);
          }
          if (!use16 || g.c16_also) *reinterpret_cast<float4*>(g.C + (size_t)(m0 + row[u]) * g.ldc + col0 + col[u]) = x;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");             // the staging area is rewritten by the next sub-range
    }
  } else {
    if (lane == 0) {
      mbar_wait(smem_u32(&bar_w), 0);
      mbar_wait(smem_u32(&bar_a), 0);
      tc_fence_after();
      // K = 64 = 4 steps of 16, three passes each, as one instruction group (gemm_tc.cuh)
      mma_f16_single_k4(tmem_d, make_desc(sa_hi), make_desc(sa_lo), make_desc(sw_hi), make_desc(sw_lo), make_idesc_f16(ncols), 0u);
      mma_commit(smem_u32(&bar_acc));
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == LOADER_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256) : "memory");
  }
}

}  // namespace tc

inline cudaError_t emb_pack(const float* W, tc::EmbScales* sc, unsigned char* out, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(sc, 0, sizeof(tc::EmbScales), st);
  if (e != cudaSuccess) return e;
  tc::dot_amax_kernel<<<16, 256, 0, st>>>(W, GAC_E, GAC_NB, GAC_E, nullptr, nullptr, 0, 0, sc->amax);
  tc::emb_scale_kernel<<<1, 1, 0, st>>>(sc);
  tc::emb_pack_kernel<<<cdiv(GAC_E * 8, 256), 256, 0, st>>>(W, sc, out);
  return cudaGetLastError();
}
inline cudaError_t launch_emb_gemm(const tc::EmbArgs& g, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc::emb_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::EMB_SMEM);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  tc::emb_gemm_tc_kernel<<<2 * cdiv(g.rows, tc::BM), tc::THREADS, tc::EMB_SMEM, st>>>(g);
  return cudaGetLastError();
}

}  // namespace gac
