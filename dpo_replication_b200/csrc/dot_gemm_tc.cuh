// "Dense -> LeakyReLU -> Dense(1)" as ONE tcgen05 kernel (sm_100a): the hidden layer never reaches HBM.
//
//   part[row][tile] = sum_{n in tile} w2[n] * lrelu_alpha( (A * W)[row][n] + bias[n] )
//
// The reference ends every network with this pair: the actor head dense_layer_2(dense_layer_1(h * noise_embedding))
// (networks.py:171-174, 221) and the 400 -> 300 -> 1 tail of every FNN (networks.py:345-358).  Unfused, the hidden
// activations (rows x N fp32) are written by the GEMM epilogue -- at HBM write speed, overlapping no MMA work -- and read
// back by a row-dot kernel.  Here a CTA owns 128 rows x bn <= 256 hidden columns; its epilogue reads the accumulators
// lane-per-row straight from TMEM, applies bias + LeakyReLU, multiplies by the output weights and reduces along the row in
// registers: 4 bytes per row and tile leave the SM.  The (at most two) tile partials are added by the consumer in a fixed
// order (plus the output bias and the output nonlinearity).
//
// Mode 1 (store) turns the same pipeline into a plain GEMM C = A * W (+ LeakyReLU' gate, + accumulate) for the actor's
// input-gradient (dgrad) contractions; its A operand is a gradient tensor whose magnitude is tracked by the kernels that
// produce it (an atomicMax of |x| per block, order-independent), so the 3xFP16 variant applies there too.
//
// Same pipeline as gru_step_tc.cuh (two loader groups, packed weights by cp.async.bulk, dual TMEM accumulators) and the same
// two arithmetic variants: 3xTF32, or 3xFP16 when the caller can bound the A operand (DotScales.ok, see gru_step_tc.cuh).
#pragma once
#include "gru_step_tc.cuh"

namespace gac {
namespace tc {

struct DotScales {
  float s_a, s_w, inv_unit;    // A' = A * s_a, W' = W * s_w, accumulators in units of s_a * s_w
  int ok;
  unsigned int slots[4];       // [0] max |W|, [1] bound / max of |A| (float bit patterns), [2] log2 s_w, [3] weights usable
};

struct DotGemmArgs {
  int rows;                    // M
  const int* dyn_rows;         // optional device row count
  const float* A; int lda;     // (rows, K) row-major, 16-byte aligned rows
  int K, N, bn, ntiles;        // K multiple of 16; N hidden columns in ntiles tiles of bn (multiple of 16, <= 256)
  const unsigned char* wpk;    // dot_pack_kernel<false> tiles
  const unsigned char* wpk16;  // dot_pack_kernel<true> tiles (used when sc->ok)
  const DotScales* sc;         // nullptr: 3xTF32 only
  const float* bias;           // (N)
  const float* w2;             // (N) output weights
  float alpha;                 // LeakyReLU slope
  float* part;                 // (rows, ntiles)
  // mode 1: C (rows, N) = A * W, optionally times (gate > 0 ? 1 : gate_alpha), optionally added to C
  int mode, seg_rows;          // seg_rows > 0: rows come in segments of seg_rows of which the first *dyn_rows are live
  float* C; int ldc;
  const float* gate; int ldgate; float gate_alpha;
  int accumulate;
  int act_lrelu;               // mode 1: y = lrelu_alpha(acc + bias) before gate / accumulate (bias may be nullptr)
  unsigned long long* dbg;     // optional per-CTA phase stamps (GAC_DGRAD_TIMELINE, tools/wgrad_timeline.py --dgrad)
};

constexpr int DOT_STAGES = 2;
inline int dot_kblocks(int K, bool f16) { return f16 ? (K + BK16 - 1) / BK16 : (K + BK - 1) / BK; }
inline size_t dot_pack_bytes(int K, int bn, int ntiles, bool f16) { return (size_t)ntiles * dot_kblocks(K, f16) * 2 * bn * 128; }
inline int dot_smem_bytes(int bn) { return DOT_STAGES * (2 * (int)A_TILE_BYTES + 2 * bn * 128) + 1024 + 2 * bn * 4 + 256 * 4; }

// W is (K, N) row-major.  Slot (tile j, block kb) = [hi | lo], bn rows (hidden columns) x 128 bytes, 128B-swizzled.
// transW: the operand is W^T, i.e. element (k, n) = W[n * ldw + k] (dgrad against a Keras (in, out) kernel).
template <bool F16>
__global__ void dot_pack_kernel(const float* __restrict__ W, int ldw, int transW, int K, int N, int bn, int ntiles,
                                unsigned char* __restrict__ out, const DotScales* __restrict__ sc) {
  constexpr int BKE = F16 ? BK16 : BK, CE = F16 ? 8 : 4;
  const int KB = (K + BKE - 1) / BKE;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_slot = (long long)bn * 8;
  if (idx >= (long long)ntiles * KB * per_slot) return;
  if (F16 && !sc->slots[3]) return;
  const int c = idx % 8, r = (idx / 8) % bn;
  const int slot = idx / per_slot, kb = slot % KB, j = slot / KB;
  const int n = j * bn + r;
  float x[CE];
#pragma unroll
  for (int e = 0; e < CE; ++e) {
    const int k = kb * BKE + CE * c + e;
    x[e] = (k < K && n < N) ? (transW ? W[(size_t)n * ldw + k] : W[(size_t)k * ldw + n]) : 0.f;
  }
  unsigned char* base = out + (size_t)slot * (2 * (size_t)bn * 128);
  const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
  if constexpr (F16) {
    const float sw = ldexpf(1.f, (int)sc->slots[2]);
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_f16x2(x[2 * e] * sw, x[2 * e + 1] * sw, h[e], l[e]);
    *reinterpret_cast<uint4*>(base + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(base + (size_t)bn * 128 + off) = make_uint4(l[0], l[1], l[2], l[3]);
  } else {
    float h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_tf32(x[e], h[e], l[e]);
    *reinterpret_cast<float4*>(base + off) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(base + (size_t)bn * 128 + off) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// slots[0] = max |W| (n elements); slots[1] = max_j (sum_i |E_ij| + |e_j|): the bound of |cosine_embedding * E + e|
// (and of its product with a GRU state, |h| < 1).  Non-finite values poison slot 1 with +inf.  Slots zeroed before.
__global__ void dot_amax_kernel(const float* __restrict__ W, int ldw, int wr, int wc, const float* __restrict__ E,
                                const float* __restrict__ e, int nb, int ne, unsigned int* __restrict__ slots) {
  float mw = 0.f, mb = 0.f;
  bool bad = false;
  for (int r = blockIdx.x; r < wr; r += gridDim.x)
    for (int cc = threadIdx.x; cc < wc; cc += blockDim.x) {
      const float a = fabsf(W[(size_t)r * ldw + cc]);
      bad |= !(a <= 3e38f);
      mw = fmaxf(mw, a);
    }
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ne; j += gridDim.x * blockDim.x) {
    float s = fabsf(e[j]);
    for (int i = 0; i < nb; ++i) s += fabsf(E[(size_t)i * ne + j]);
    bad |= !(s <= 3e38f);
    mb = fmaxf(mb, s);
  }
  if (bad) mw = mb = __int_as_float(0x7f800000);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o)); mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o)); }
  // one global atomic per block and slot (a thousand same-address atomics serialise at L2 for longer than the scan takes)
  __shared__ unsigned int sh_max[2];
  if (threadIdx.x < 2) sh_max[threadIdx.x] = 0u;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { atomicMax(&sh_max[0], __float_as_uint(mw)); atomicMax(&sh_max[1], __float_as_uint(mb)); }
  __syncthreads();
  if (threadIdx.x < 2 && sh_max[threadIdx.x]) atomicMax(slots + threadIdx.x, sh_max[threadIdx.x]);
}
// weight half of the scales (at pack time): slots[2] = log2 s_w with |W * s_w| in [1, 2), slots[3] = usable
__global__ void dot_wscale_kernel(DotScales* __restrict__ sc) {
  const float mw = __uint_as_float(sc->slots[0]);
  int ew = 0, okw = 0;
  if (mw > 0.f && mw < 1e30f) { frexpf(mw, &ew); okw = abs(ew) < 100; }
  sc->slots[2] = (unsigned int)(1 - ew); sc->slots[3] = (unsigned int)okw;
  sc->s_w = ldexpf(1.f, 1 - ew);
  sc->ok = 0; sc->s_a = 1.f; sc->inv_unit = 1.f;
}
// operand half (just before the GEMM): the bound of |A| comes from slots[1] (weights-derived) or from a producer's
// running maximum (a_amax): |A * s_a| < 2^13
__global__ void dot_ascale_kernel(DotScales* __restrict__ sc, const unsigned int* __restrict__ a_amax) {
  const float mb = __uint_as_float(a_amax ? *a_amax : sc->slots[1]);
  sc->ok = 0; sc->s_a = 1.f; sc->inv_unit = 1.f;
  if (sc->slots[3] && mb > 0.f && mb < 1e30f) {
    int eb;
    frexpf(mb, &eb);
    if (abs(eb) < 100) {
      const int ea = 13 - eb, es = (int)sc->slots[2];
      sc->s_a = ldexpf(1.f, ea); sc->inv_unit = ldexpf(1.f, -(ea + es));
      sc->ok = 1;
    }
  }
}

template <bool F16>
__device__ __forceinline__ void dot_gemm_body(const DotGemmArgs& g) {
  constexpr int BKE = F16 ? BK16 : BK;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[DOT_STAGES], bar_empty[DOT_STAGES], bar_accum;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bn = g.bn, jt = (int)blockIdx.x % g.ntiles, m0 = ((int)blockIdx.x / g.ntiles) * BM, n0 = jt * bn;
  int m_end = g.rows;
  if (g.dyn_rows) {
    const int nvalid = *g.dyn_rows;
    if (g.seg_rows > 0) {                                       // segments of seg_rows (a multiple of 128) rows, the first nvalid live
      const int in_seg = m0 % g.seg_rows;
      if (in_seg >= nvalid) return;
      m_end = min(g.rows, m0 - in_seg + nvalid);
    } else {
      m_end = min(g.rows, nvalid);
    }
  }
  if (m0 >= m_end) return;
  const int nkb = (g.K + BKE - 1) / BKE;
  unsigned long long* dbg = (g.dbg && tid == 0) ? g.dbg + 8 + (size_t)blockIdx.x * 8 : nullptr;
  if (dbg) {
    unsigned long long gt; unsigned sm;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
    dbg[0] = gt; dbg[1] = clock64(); dbg[6] = (unsigned long long)nkb; dbg[7] = sm;
    if (blockIdx.x == 0) { g.dbg[0] = 0x600DC0DE600DC0DFull; g.dbg[1] = gridDim.x; }
  }
  const uint32_t stage_bytes = 2 * A_TILE_BYTES + 2 * (uint32_t)bn * 128u;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  float* bias_s = reinterpret_cast<float*>(smem_raw + (smem0 - smem_u32(smem_raw)) + DOT_STAGES * stage_bytes);
  float* w2_s = bias_s + bn;
  float* part_s = w2_s + bn;                                     // [2][128]

  if (tid == 0) {
    for (int s = 0; s < DOT_STAGES; ++s) { mbar_init(smem_u32(&bar_full[s]), LOADER_WARPS / 2); mbar_init(smem_u32(&bar_empty[s]), 1); }
    mbar_init(smem_u32(&bar_accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == LOADER_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (g.mode == 0 || g.bias)
    for (int i = tid; i < bn; i += THREADS) {
      const int n = n0 + i;
      bias_s[i] = (n < g.N && g.bias) ? g.bias[n] : 0.f;
      w2_s[i] = (n < g.N && g.w2) ? g.w2[n] : 0.f;               // padded columns contribute nothing
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  if (dbg) dbg[2] = clock64();

  if (warp < LOADER_WARPS) {
    // ------------------------------ loaders: two groups alternate K blocks ------------------------------
    const int grp = warp >> 2, t = tid & (GROUP - 1);
    const float s_a = F16 ? g.sc->s_a : 1.f;
    float x16[16][4];
    auto load16 = [&](int kb) {                                  // see gru_step_tc.cuh: sector-exact float4 mapping
      const int k0 = kb * BKE, live = min(BKE, g.K - k0);
      const int kq = t & 15;
      const float* src = g.A + (size_t)(m0 + (t >> 4)) * g.lda + k0 + 4 * kq;
      const bool klive = kb < nkb && 4 * kq < live;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (klive && m0 + (t >> 4) + 8 * i < m_end) {
          const float4 v = *reinterpret_cast<const float4*>(src + (size_t)(8 * i) * g.lda);
          x16[i][0] = v.x; x16[i][1] = v.y; x16[i][2] = v.z; x16[i][3] = v.w;
        } else {
          x16[i][0] = x16[i][1] = x16[i][2] = x16[i][3] = 0.f;
        }
      }
    };
    if constexpr (F16) load16(grp);
    for (int kb = grp; kb < nkb; kb += 2) {
      const int k0 = kb * BKE, live = min(BKE, g.K - k0);
      float xa[8][4], la[8][4];
      uint32_t h16[16][2], l16[16][2];
      if constexpr (!F16) {
        load_operand<false, true, 8>(xa, g.A, g.lda, m0, m_end, BM, k0, live, t);
        split_operand<8>(xa, la);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          split_f16x2(x16[i][0] * s_a, x16[i][1] * s_a, h16[i][0], l16[i][0]);
          split_f16x2(x16[i][2] * s_a, x16[i][3] * s_a, h16[i][1], l16[i][1]);
        }
        load16(kb + 2);
      }
      const int s = kb % DOT_STAGES;
      const uint32_t ph = (uint32_t)(kb / DOT_STAGES) & 1u;
      mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
      const uint32_t sa = smem0 + (uint32_t)s * stage_bytes;
      if (t == 0) {
        const uint32_t bytes = 2u * (uint32_t)bn * 128u;
        const unsigned char* srcp = (F16 ? g.wpk16 : g.wpk) + ((size_t)jt * nkb + kb) * bytes;
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
    // ------------------------------ epilogue: bias, LeakyReLU, dot with w2 along the row, from TMEM ------------------------------
    if (dbg) dbg[3] = clock64();
    mbar_wait(smem_u32(&bar_accum), 0);
    tc_fence_after();
    if (dbg) dbg[4] = clock64();
    const int quarter = warp & 3, hf = warp >> 2;
    const int nch = bn / 16, c0 = hf ? (nch + 1) / 2 : 0, c1 = hf ? nch : (nch + 1) / 2;
    const float inv_unit = F16 ? g.sc->inv_unit : 1.f, alpha = g.alpha;
    if (g.mode == 1 && (g.gate || g.accumulate)) {
      // store epilogue with per-element operands (gate / accumulate): lane-per-row global loads would serialise a
      // dependent load -> multiply -> store chain per 16-column chunk (measured 24.6k of a 59k-cycle tile).  Stage the
      // scaled sums through the idle pipeline smem instead and walk (row, 4 columns) items so that every global access is
      // row-contiguous and the operand loads of IB items are in flight together.
      const int SS = bn + 4;                                    // floats; (bn + 4) / 4 is odd => conflict-free 128-bit row stores
      float* stg = reinterpret_cast<float*>(smem_raw + (smem0 - smem_u32(smem_raw)));
      {
        float* srow = stg + (size_t)(quarter * 32 + lane) * SS;
        for (int ci = c0; ci < c1; ++ci) {
          float v[16], w[16];
          const uint32_t ta = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ci * 16);
          tmem_ld16(ta, v);
          tmem_ld16(ta + 256u, w);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            float y = F16 ? (v[q] + w[q] * (1.f / 2048.f)) * inv_unit : v[q] + w[q];
            if (g.bias) y += bias_s[ci * 16 + q];
            if (g.act_lrelu) y = y > 0.f ? y : alpha * y;
            v[q] = y;
          }
#pragma unroll
          for (int q = 0; q < 16; q += 4) *reinterpret_cast<float4*>(srow + ci * 16 + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int n4 = bn >> 2, total = BM * n4;
      constexpr int IB = 4;
      const float ga = g.gate_alpha;
      for (int i0 = tid; i0 < total; i0 += IB * LOADER_WARPS * 32) {
        float4 gt[IB], ac[IB];
        int row[IB], col[IB];
        bool okk[IB];
#pragma unroll
        for (int u = 0; u < IB; ++u) {
          const int i = i0 + u * LOADER_WARPS * 32;
          row[u] = i / n4;
          col[u] = 4 * (i - row[u] * n4);
          okk[u] = i < total && m0 + row[u] < m_end && n0 + col[u] < g.N;
          if (!okk[u]) continue;
          const size_t mrow = (size_t)(m0 + row[u]);
          if (g.gate) gt[u] = *reinterpret_cast<const float4*>(g.gate + mrow * g.ldgate + n0 + col[u]);
          if (g.accumulate) ac[u] = *reinterpret_cast<const float4*>(g.C + mrow * g.ldc + n0 + col[u]);
        }
#pragma unroll
        for (int u = 0; u < IB; ++u) {
          if (!okk[u]) continue;
          float4 x = *reinterpret_cast<const float4*>(stg + (size_t)row[u] * SS + col[u]);
          if (g.gate) { x.x *= gt[u].x > 0.f ? 1.f : ga; x.y *= gt[u].y > 0.f ? 1.f : ga; x.z *= gt[u].z > 0.f ? 1.f : ga; x.w *= gt[u].w > 0.f ? 1.f : ga; }
          if (g.accumulate) { x.x += ac[u].x; x.y += ac[u].y; x.z += ac[u].z; x.w += ac[u].w; }
          *reinterpret_cast<float4*>(g.C + (size_t)(m0 + row[u]) * g.ldc + n0 + col[u]) = x;
        }
      }
    } else     if (g.mode == 1) {
      // store epilogue: lane = row; 16-column chunks, 128-bit accesses (N, ldc, ldgate multiples of 4, aligned bases)
      const int m = m0 + quarter * 32 + lane;
      const bool m_ok = m < m_end;
      for (int ci = c0; ci < c1; ++ci) {
        float v[16], w[16];
        const uint32_t ta = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ci * 16);
        tmem_ld16(ta, v);
        tmem_ld16(ta + 256u, w);
        const int nb = n0 + ci * 16;
        if (!m_ok || nb >= g.N) continue;
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = F16 ? (v[q] + w[q] * (1.f / 2048.f)) * inv_unit : v[q] + w[q];
        if (g.bias) {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] += bias_s[ci * 16 + q];
        }
        if (g.act_lrelu) {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = v[q] > 0.f ? v[q] : alpha * v[q];
        }
        float* crow = g.C + (size_t)m * g.ldc + nb;
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
          if (nb + q >= g.N) break;
          float4 x = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
          if (g.gate) {
            const float4 b = *reinterpret_cast<const float4*>(g.gate + (size_t)m * g.ldgate + nb + q);
            const float ga = g.gate_alpha;
            x.x *= b.x > 0.f ? 1.f : ga; x.y *= b.y > 0.f ? 1.f : ga; x.z *= b.z > 0.f ? 1.f : ga; x.w *= b.w > 0.f ? 1.f : ga;
          }
          if (g.accumulate) {
            const float4 b = *reinterpret_cast<const float4*>(crow + q);
            x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
          }
          *reinterpret_cast<float4*>(crow + q) = x;
        }
      }
    } else {
    float acc = 0.f;
    for (int ci = c0; ci < c1; ++ci) {
      float v[16], w[16];
      const uint32_t ta = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ci * 16);
      tmem_ld16(ta, v);
      tmem_ld16(ta + 256u, w);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        float y = F16 ? (v[q] + w[q] * (1.f / 2048.f)) * inv_unit : v[q] + w[q];
        y += bias_s[ci * 16 + q];
        y = y > 0.f ? y : alpha * y;
        acc = fmaf(y, w2_s[ci * 16 + q], acc);
      }
    }
    part_s[hf * BM + quarter * 32 + lane] = acc;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (hf == 0) {
      const int row = quarter * 32 + lane;
      if (m0 + row < m_end) g.part[(size_t)(m0 + row) * g.ntiles + jt] = part_s[row] + part_s[BM + row];
    }
    }
    if (dbg) dbg[5] = clock64();
  } else {
    // ------------------------------ MMA issuer (one lane) ------------------------------
    if (lane == 0) {
      const uint32_t idesc = F16 ? make_idesc_f16(bn) : make_idesc(bn);
      const uint32_t corr = tmem_d + 256u;
      for (int kb = 0; kb < nkb; ++kb) {
        const int k0 = kb * BKE, live = min(BKE, g.K - k0);
        const int s = kb % DOT_STAGES;
        const uint32_t ph = (uint32_t)(kb / DOT_STAGES) & 1u;
        mbar_wait(smem_u32(&bar_full[s]), ph);
        tc_fence_after();
        const uint32_t sa = smem0 + (uint32_t)s * stage_bytes;
        const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_TILE_BYTES);
        const uint64_t b_hi = make_desc(sa + 2 * A_TILE_BYTES), b_lo = make_desc(sa + 2 * A_TILE_BYTES + (uint32_t)bn * 128u);
        const int ksteps = F16 ? (live + 15) / 16 : (live + 7) / 8;
        if constexpr (F16) {
          // one instruction group per K block (gemm_tc.cuh, mma_f16_dual_k)
          const uint32_t acc0 = kb ? 1u : 0u;
          if (ksteps == 4) mma_f16_dual_k<4>(tmem_d, corr, a_hi, a_lo, b_hi, b_lo, idesc, acc0);
          else if (ksteps == 3) mma_f16_dual_k<3>(tmem_d, corr, a_hi, a_lo, b_hi, b_lo, idesc, acc0);
          else if (ksteps == 2) mma_f16_dual_k<2>(tmem_d, corr, a_hi, a_lo, b_hi, b_lo, idesc, acc0);
          else mma_f16_dual_k<1>(tmem_d, corr, a_hi, a_lo, b_hi, b_lo, idesc, acc0);
        }
        for (int j = 0; j < (F16 ? 0 : ksteps); ++j) {
          const uint64_t adv = (uint64_t)(j * 2);
          const uint32_t acc = (kb | j) ? 1u : 0u;
          if constexpr (F16) {
          } else {
            mma_tf32(corr, a_lo + adv, b_hi + adv, idesc, acc);
            mma_tf32(corr, a_hi + adv, b_lo + adv, idesc, 1u);
            mma_tf32(tmem_d, a_hi + adv, b_hi + adv, idesc, acc);
          }
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

__global__ void __launch_bounds__(THREADS, 1) dot_gemm_tc_kernel(const DotGemmArgs g) {
  if (g.sc && g.sc->ok) dot_gemm_body<true>(g);
  else dot_gemm_body<false>(g);
}

}  // namespace tc

// weights -> packed tiles of both variants.  bound_E / bound_e (nb x ne, ne): the embedding whose weights bound the A operand
// (slots[1]); nullptr when the bound comes from a producer's running maximum at launch time.  sc == nullptr: 3xTF32 only.
inline cudaError_t dot_pack(const float* W, int ldw, int transW, int K, int N, int bn, int ntiles, const float* bound_E, const float* bound_e,
                            int nb, int ne, unsigned char* out32, unsigned char* out16, tc::DotScales* sc, cudaStream_t st) {
  if (sc) {
    cudaError_t e = cudaMemsetAsync(sc, 0, sizeof(tc::DotScales), st);
    if (e != cudaSuccess) return e;
    // |W| over the (K x N) or (N x K) matrix: rows of ldw floats, `cols` used
    const int wr = transW ? N : K, wc = transW ? K : N;
    tc::dot_amax_kernel<<<148, 256, 0, st>>>(W, ldw, wr, wc, bound_E, bound_e, nb, bound_E ? ne : 0, reinterpret_cast<unsigned int*>(sc) + 4);
    tc::dot_wscale_kernel<<<1, 1, 0, st>>>(sc);
    const long long t16 = (long long)ntiles * tc::dot_kblocks(K, true) * bn * 8;
    tc::dot_pack_kernel<true><<<(unsigned)cdiv64(t16, 256), 256, 0, st>>>(W, ldw, transW, K, N, bn, ntiles, out16, sc);
  }
  const long long t32 = (long long)ntiles * tc::dot_kblocks(K, false) * bn * 8;
  tc::dot_pack_kernel<false><<<(unsigned)cdiv64(t32, 256), 256, 0, st>>>(W, ldw, transW, K, N, bn, ntiles, out32, nullptr);
  return cudaGetLastError();
}
// operand half of the scales, right before the launch (a_amax == nullptr: the weights-derived bound in the struct)
inline cudaError_t dot_ascale(tc::DotScales* sc, const unsigned int* a_amax, cudaStream_t st) {
  tc::dot_ascale_kernel<<<1, 1, 0, st>>>(sc, a_amax);
  return cudaGetLastError();
}

inline cudaError_t launch_dot_gemm(const tc::DotGemmArgs& g, cudaStream_t st) {
  static int attr_bytes = 0;
  const int smem = tc::dot_smem_bytes(g.bn);
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(tc::dot_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_bytes = smem;
  }
  dim3 grid(g.ntiles * cdiv(g.rows, tc::BM));
  tc::dot_gemm_tc_kernel<<<grid, tc::THREADS, smem, st>>>(g);
  return cudaGetLastError();
}

}  // namespace gac
