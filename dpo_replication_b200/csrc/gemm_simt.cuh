// fp32 SIMT GEMM (FFMA) with the shared fused epilogue.  This is the parity reference
// engine (engine = 1) and the path for the small / oddly shaped contractions (K = S, S+A, 64,
// 300) that do not fill a tcgen05 tile.  128x128x8 tile, 256 threads, 8x8 register tile.
#pragma once
#include "gac_common.cuh"

namespace gac {

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmArgs g) {
  constexpr int BM = 128, BN = 128, BK = 8, PAD = 4;
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nvalid = g.dyn_rows ? *g.dyn_rows : 0x7fffffff;
  // seg_rows is a multiple of BM, so a tile never straddles a segment
  if (!g.valid_on_k && g.dyn_rows && (m0 % g.seg_rows) >= nvalid) return;

  int kbeg = 0, kend = g.K;
  if (g.splitk > 1) {
    const int per = ((g.K + g.splitk - 1) / g.splitk + BK - 1) / BK * BK;
    kbeg = min(g.K, (int)blockIdx.z * per);
    kend = min(g.K, kbeg + per);
  }

  float ra[4], rb[4];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int row, k;
      if (!TA) { row = (tid >> 3) + 32 * i; k = tid & 7; } else { row = tid & 127; k = (tid >> 7) + 2 * i; }
      const int m = m0 + row, kk = k0 + k;
      bool ok = m < g.M && kk < kend;
      if (ok && g.dyn_rows) ok = ((g.valid_on_k ? kk : m) % g.seg_rows) < nvalid;
      ra[i] = ok ? (TA ? g.A[(size_t)kk * g.lda + m] : g.A[(size_t)m * g.lda + kk]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int n, k;
      if (!TB) { n = tid & 127; k = (tid >> 7) + 2 * i; } else { n = (tid >> 3) + 32 * i; k = tid & 7; }
      const int nn = n0 + n, kk = k0 + k;
      bool ok = nn < g.N && kk < kend;
      if (ok && g.dyn_rows && g.valid_on_k) ok = (kk % g.seg_rows) < nvalid;
      rb[i] = ok ? (TB ? g.B[(size_t)nn * g.ldb + kk] : g.B[(size_t)kk * g.ldb + nn]) : 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int row, k;
      if (!TA) { row = (tid >> 3) + 32 * i; k = tid & 7; } else { row = tid & 127; k = (tid >> 7) + 2 * i; }
      As[buf][k][row] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int n, k;
      if (!TB) { n = tid & 127; k = (tid >> 7) + 2 * i; } else { n = (tid >> 3) + 32 * i; k = tid & 7; }
      Bs[buf][k][n] = rb[i];
    }
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  int buf = 0;
  if (kbeg < kend) {
    load_tiles(kbeg);
    store_tiles(0);
    __syncthreads();
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
      const bool more = k0 + BK < kend;
      if (more) load_tiles(k0 + BK);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (more) {
        store_tiles(buf ^ 1);
        __syncthreads();
        buf ^= 1;
      }
    }
  }

  float* Cout = g.C + (g.splitk > 1 ? (size_t)blockIdx.z * g.split_stride : 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
    if (!g.valid_on_k && g.dyn_rows && (m % g.seg_rows) >= nvalid) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      const size_t ci = g.transC ? (size_t)n * g.ldc + m : (size_t)m * g.ldc + n;
      if (g.splitk > 1) Cout[ci] = acc[i][j];
      else g.C[ci] = gemm_epilogue(g, acc[i][j], m, n);
    }
  }
}

// Small-M path (M <= 8: the acting path, one state per environment step -- agent.py:218-248): a 128x128 tile
// would run one serial K loop on a handful of CTAs.  Here a block owns 64 output columns, 8 thread rows split
// K, partial sums are folded through shared memory in fixed order; loads of B are coalesced along n.
template <bool TB>
__global__ void __launch_bounds__(512) gemm_smallm_kernel(const GemmArgs g) {
  __shared__ float red[8][8][65];
  const int tx = threadIdx.x, ty = threadIdx.y;          // tx: column in the block, ty: K slice
  const int n = blockIdx.x * 64 + tx;
  float acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.f;
  if (n < g.N) {
    for (int k = ty; k < g.K; k += 8) {
      const float b = TB ? g.B[(size_t)n * g.ldb + k] : g.B[(size_t)k * g.ldb + n];
#pragma unroll
      for (int m = 0; m < 8; ++m)
        if (m < g.M) acc[m] = fmaf(g.A[(size_t)m * g.lda + k], b, acc[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) red[ty][m][tx] = acc[m];
  __syncthreads();
  if (ty < g.M && n < g.N) {                             // thread row ty finishes output row m = ty
    const int m = ty;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][m][tx];
    g.C[(size_t)m * g.ldc + n] = gemm_epilogue(g, s, m, n);
  }
}

// dst[i] = (accumulate ? dst[i] : 0) + sum_z part[z*stride + i]   (fixed order => deterministic)
__global__ void reduce_splits_kernel(const float* __restrict__ part, long long stride, int nsplit, float* __restrict__ dst,
                                     long long n, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = accumulate ? dst[i] : 0.f;
  for (int z = 0; z < nsplit; ++z) s += part[(size_t)z * stride + i];
  dst[i] = s;
}

// split-K reduction for SMALL problems: part holds nsplit dense (M, N) partial products; the fused epilogue of the
// original GEMM (bias, activation, gates, accumulate, transposed store) is applied to the fixed-order sum.
__global__ void reduce_splits_epi_kernel(const float* __restrict__ part, int nsplit, const GemmArgs g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long mn = (long long)g.M * g.N;
  if (i >= mn) return;
  const int m = (int)(i / g.N), n = (int)(i - (long long)m * g.N);
  float s = 0.f;
  for (int z = 0; z < nsplit; ++z) s += part[(size_t)z * mn + i];
  const size_t ci = g.transC ? (size_t)n * g.ldc + m : (size_t)m * g.ldc + n;
  g.C[ci] = gemm_epilogue(g, s, m, n);
}

inline cudaError_t launch_gemm_simt(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 8 && !g.transA && !g.transC && g.splitk <= 1 && !g.dyn_rows) {
    dim3 blk(64, 8);
    if (g.transB) gemm_smallm_kernel<true><<<cdiv(g.N, 64), blk, 0, st>>>(g);
    else gemm_smallm_kernel<false><<<cdiv(g.N, 64), blk, 0, st>>>(g);
    return cudaGetLastError();
  }
  dim3 grid(cdiv(g.N, 128), cdiv(g.M, 128), g.splitk > 1 ? g.splitk : 1);
  if (g.transA) {
    if (g.transB) gemm_simt_kernel<true, true><<<grid, 256, 0, st>>>(g);
    else gemm_simt_kernel<true, false><<<grid, 256, 0, st>>>(g);
  } else {
    if (g.transB) gemm_simt_kernel<false, true><<<grid, 256, 0, st>>>(g);
    else gemm_simt_kernel<false, false><<<grid, 256, 0, st>>>(g);
  }
  return cudaGetLastError();
}

}  // namespace gac
