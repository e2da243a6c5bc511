// Elementwise / reduction / compaction kernels of the GAC step (all HBM- or L2-bound).
#pragma once
#include "gac_common.cuh"

namespace gac {

// i_pi[i] = fl32(fl32(i+1) * fl32(pi)): the reference builds the constant as
// np.arange(1,65,float32) * np.pi (GAC/networks.py:45) -- SURVEY.md F4.
__constant__ float c_ipi[GAC_NB];

struct RowValid {          // row r live iff (r % seg) < *dyn  (dyn == nullptr: always live)
  const int* dyn; int seg;
  __device__ __forceinline__ bool operator()(int r) const { return !dyn || (r % seg) < *dyn; }
};

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// fixed-order block sum (deterministic); result valid on thread 0
__device__ __forceinline__ float block_sum(float v, float* sh /* >= 32 floats */) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) for (int i = 0; i < (int)(blockDim.x + 31) / 32; ++i) t += sh[i];
  return t;
}

// ---- cosine basis: out[r][i] = cos(fl32(x[r*incx] * i_pi[i]))   networks.py:44-47 ----
// 4 basis functions per thread (one 128-bit store), 16 threads per row.
__global__ void cos_basis_kernel(const float* __restrict__ x, int incx, float* __restrict__ out, int rows, RowValid rv) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = idx >> 4, i = (idx & 15) * 4;
  if (r >= rows || !rv(r)) return;
  const float xv = x ? x[(size_t)r * incx] : 0.f;
  reinterpret_cast<float4*>(out)[idx] = make_float4(cosf(__fmul_rn(xv, c_ipi[i])), cosf(__fmul_rn(xv, c_ipi[i + 1])),
                                                    cosf(__fmul_rn(xv, c_ipi[i + 2])), cosf(__fmul_rn(xv, c_ipi[i + 3])));
}

// teacher-forced "shifted action": step d reads action d-1 (0 for d == 0)   networks.py:253-255
// out[(d*seg + r)][i] = cos(shift * i_pi[i]); second output for the noise embedding of tau[r][d].
__global__ void cos_basis_seq_kernel(const float* __restrict__ teacher, const float* __restrict__ taus, int A, int seg,
                                     float* __restrict__ ca, float* __restrict__ cn, const int* dyn) {
  const int nvalid = dyn ? *dyn : seg;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // 4 basis functions per thread
  const int i = (int)(idx & 15) * 4;
  const long long rs = idx >> 4;
  if (rs >= (long long)A * seg) return;
  const int d = rs / seg, r = rs % seg;
  if (r >= nvalid) return;
  const float prev = d == 0 ? 0.f : teacher[(size_t)r * A + d - 1], tau = taus[(size_t)r * A + d];
  reinterpret_cast<float4*>(ca)[idx] = make_float4(cosf(__fmul_rn(prev, c_ipi[i])), cosf(__fmul_rn(prev, c_ipi[i + 1])),
                                                   cosf(__fmul_rn(prev, c_ipi[i + 2])), cosf(__fmul_rn(prev, c_ipi[i + 3])));
  reinterpret_cast<float4*>(cn)[idx] = make_float4(cosf(__fmul_rn(tau, c_ipi[i])), cosf(__fmul_rn(tau, c_ipi[i + 1])),
                                                   cosf(__fmul_rn(tau, c_ipi[i + 2])), cosf(__fmul_rn(tau, c_ipi[i + 3])));
}

// ---- x = concat(states[sidx[r]], actions[r])   networks.py:385 ----
__global__ void concat_gather_kernel(const float* __restrict__ states, const int* __restrict__ sidx,
                                     const float* __restrict__ actions, int S, int A, int rows, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int W = S + A;
  if (idx >= (long long)rows * W) return;
  const int r = idx / W, c = idx % W;
  out[idx] = c < S ? states[(size_t)(sidx ? sidx[r] : r) * S + c] : actions[(size_t)r * A + (c - S)];
}

// ---- Critic.train input: concat(s, clip(a + (2u-1)*qn, -1, 1))   networks.py:411-417 ----
__global__ void critic_train_x_kernel(const float* __restrict__ s, const float* __restrict__ a, const float* __restrict__ u,
                                      float qn, int S, int A, int B, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int W = S + A;
  if (idx >= B * W) return;
  const int r = idx / W, c = idx % W;
  if (c < S) { out[idx] = s[r * S + c]; return; }
  const float nz = (u[r * A + c - S] * 2.f - 1.f) * qn;
  out[idx] = fminf(fmaxf(a[r * A + c - S] + nz, -1.f), 1.f);
}

// ---- GRU gates (Keras GRU, reset_after=True, gate order [z|r|h]) ----
// mx = sx_u[sidx[row]] (state half of the input projection + input bias, hoisted per unique
// state) + mxa[row] (action half); mh = mhr[row] + b1 (recurrent bias).  mhr == nullptr means
// h_prev == 0 (first step: 0*U is skipped, the bias is still added).
__global__ void gru_gates_fwd_kernel(const float* __restrict__ sx_u, const int* __restrict__ sidx,
                                     const float* __restrict__ mxa, const float* __restrict__ mhr,
                                     const float* __restrict__ b1, const float* __restrict__ hprev,
                                     float* __restrict__ hout, float* __restrict__ zs, float* __restrict__ rs_,
                                     float* __restrict__ cs, float* __restrict__ mhhs, int rows, RowValid rv) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * GAC_E) return;
  const int row = idx / GAC_E, j = idx % GAC_E;
  if (!rv(row)) return;
  const float* sx = sx_u + (size_t)sidx[row] * GAC_G;
  const float* mx = mxa + (size_t)row * GAC_G;
  const float xz = sx[j] + mx[j], xr = sx[GAC_E + j] + mx[GAC_E + j], xh = sx[2 * GAC_E + j] + mx[2 * GAC_E + j];
  float hz = b1[j], hr = b1[GAC_E + j], hh = b1[2 * GAC_E + j];
  float hp = 0.f;
  if (mhr) {
    const float* mh = mhr + (size_t)row * GAC_G;
    hz += mh[j]; hr += mh[GAC_E + j]; hh += mh[2 * GAC_E + j];
    hp = hprev[idx];
  }
  const float z = sigmoid_f(xz + hz), r = sigmoid_f(xr + hr);
  const float c = tanhf(xh + r * hh);
  hout[idx] = z * hp + (1.f - z) * c;
  if (zs) { zs[idx] = z; rs_[idx] = r; cs[idx] = c; mhhs[idx] = hh; }
}

// Running maximum of |x| of a tensor a kernel writes (the bound the 3xFP16 GEMMs scale their A operand with): one
// shared-memory atomicMax per thread, one global atomicMax per block; maxima are order-independent, so this is
// deterministic.  Non-negative floats order like their bit patterns; NaN / inf come out >= 0x7f800000 and disable FP16.
__device__ __forceinline__ void block_amax(float m, unsigned int* __restrict__ slot) {
  __shared__ unsigned int sh_amax;
  if (threadIdx.x == 0) sh_amax = 0u;
  __syncthreads();
  atomicMax(&sh_amax, __float_as_uint(fabsf(m)) & 0x7fffffffu);
  __syncthreads();
  if (threadIdx.x == 0 && sh_amax) atomicMax(slot, sh_amax);
}

// Scales of the 3xFP16 weight-gradient GEMM dW = X^T dY, right before its launch: |dY| from the producers' running maximum,
// |X| from a weights-derived bound (bound_x) or a constant (activations in [-1, 1]).  Scaled operands stay below 2^13.
__global__ void wg_scale_kernel(WgScales* __restrict__ sc, const unsigned int* __restrict__ amax_dy, const unsigned int* __restrict__ bound_x,
                                float bound_const) {
  const float ma = __uint_as_float(*amax_dy), mb = bound_x ? __uint_as_float(*bound_x) : bound_const;
  sc->ok = 0; sc->s_a = sc->s_b = sc->inv_unit = 1.f;
  if (ma > 0.f && mb > 0.f && ma < 1e30f && mb < 1e30f) {
    int ea, eb;
    frexpf(ma, &ea); frexpf(mb, &eb);
    if (abs(ea) < 100 && abs(eb) < 100) {
      sc->s_a = ldexpf(1.f, 13 - ea); sc->s_b = ldexpf(1.f, 13 - eb); sc->inv_unit = ldexpf(1.f, ea + eb - 26);
      sc->ok = 1;
    }
  }
}

// Producer kernels below also emit per-row-block column sums (bias gradients) so that no second
// pass over the activations is needed.  Layout: block = FB_COLS threads (one column each) x
// FB_ROWS rows; thread loops over the rows with 4-way ILP; partial sums go to
// part[(row_block) * ncols + col] and are reduced in fixed order by reduce_cols_kernel.
constexpr int FB_COLS = 100;   // 400 = 4 x 100 columns
constexpr int FB_ROWS = 64;

// backward of the gates for one step; dh = dh_next + dhu.  Writes dmx, dmh (rows,1200) and
// dh_carry = dh*z (the GEMM dmh*U^T is accumulated on top of it afterwards), plus the column
// sums of dmx / dmh of this row block (gradients of the GRU input / recurrent bias).
__global__ void __launch_bounds__(FB_COLS) gru_gates_bwd_kernel(const float* __restrict__ dh_next, const float* __restrict__ dhu,
                                     const float* __restrict__ zs, const float* __restrict__ rs_,
                                     const float* __restrict__ cs, const float* __restrict__ mhhs,
                                     const float* __restrict__ hprev, float* __restrict__ dmx, float* __restrict__ dmh,
                                     float* __restrict__ dh_carry, int rows, RowValid rv, float* __restrict__ part_mx,
                                     float* __restrict__ part_mh, unsigned int* __restrict__ amax /* [0] |dmx|, [1] |dmh| or null */) {
  const int j = blockIdx.x * FB_COLS + threadIdx.x;
  const int r0 = blockIdx.y * FB_ROWS;
  const int nvalid = rv.dyn ? min(*rv.dyn, rows) : rows;
  const int r1 = min(r0 + FB_ROWS, nvalid);
  float sz = 0.f, sr = 0.f, sc = 0.f, scr = 0.f, mx = 0.f;
#pragma unroll 4
  for (int row = r0; row < r1; ++row) {
    const size_t idx = (size_t)row * GAC_E + j;
    const float dh = (dh_next ? dh_next[idx] : 0.f) + dhu[idx];
    const float z = zs[idx], r = rs_[idx], c = cs[idx], hp = hprev ? hprev[idx] : 0.f;
    const float dz = dh * (hp - c), dc = dh * (1.f - z);
    const float dcp = dc * (1.f - c * c);
    const float dzp = dz * z * (1.f - z);
    const float drp = dcp * mhhs[idx] * r * (1.f - r);
    const size_t o = (size_t)row * GAC_G + j;
    dmx[o] = dzp; dmx[o + GAC_E] = drp; dmx[o + 2 * GAC_E] = dcp;
    dmh[o] = dzp; dmh[o + GAC_E] = drp; dmh[o + 2 * GAC_E] = dcp * r;
    dh_carry[idx] = dh * z;
    sz += dzp; sr += drp; sc += dcp; scr += dcp * r;
    mx = fmaxf(mx, fmaxf(fabsf(dzp), fmaxf(fabsf(drp), fabsf(dcp))));      // |dcp * r| <= |dcp|: one bound serves both tensors
    if (!(fabsf(dzp) + fabsf(drp) + fabsf(dcp) <= 3e38f)) mx = __int_as_float(0x7f800000);
  }
  float* pm = part_mx + (size_t)blockIdx.y * GAC_G + j;
  float* ph = part_mh + (size_t)blockIdx.y * GAC_G + j;
  pm[0] = sz; pm[GAC_E] = sr; pm[2 * GAC_E] = sc;
  ph[0] = sz; ph[GAC_E] = sr; ph[2 * GAC_E] = scr;
  if (amax) block_amax(mx, amax);
}

// dy1p[r][j] = dop[r] * w2[j] * lrelu'(y1[r][j]); partial sums: sum_r dy1p (d b1) and
// sum_r y1*dop (d w2).  Rows are step-major with segment length seg.
__global__ void __launch_bounds__(FB_COLS) head_bwd_kernel(const float* __restrict__ dop, const float* __restrict__ w2,
                                                           const float* __restrict__ y1, long long rows, RowValid rv,
                                                           float* __restrict__ dy1p, float* __restrict__ part_b1,
                                                           float* __restrict__ part_w2, unsigned int* __restrict__ amax /* |dy1p| or null */) {
  const int j = blockIdx.x * FB_COLS + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * FB_ROWS;
  long long r1 = min(rows, r0 + FB_ROWS);
  if (rv.dyn) { const long long seg0 = (r0 / rv.seg) * rv.seg; r1 = min(r1, seg0 + min(*rv.dyn, rv.seg)); }
  const float w = w2[j];
  float sb = 0.f, sw = 0.f, mx = 0.f;
#pragma unroll 4
  for (long long row = r0; row < r1; ++row) {
    const size_t idx = (size_t)row * GAC_E + j;
    const float y = y1[idx], d = dop[row];
    const float g = d * w * (y > 0.f ? 1.f : 0.01f);
    dy1p[idx] = g;
    sb += g; sw += y * d;
    mx = fmaxf(mx, fabsf(g));
    if (!(fabsf(g) <= 3e38f)) mx = __int_as_float(0x7f800000);
  }
  part_b1[(size_t)blockIdx.y * GAC_E + j] = sb;
  part_w2[(size_t)blockIdx.y * GAC_E + j] = sw;
  if (amax) block_amax(mx, amax);
}

// dne = du * h ; dhu = du * ne ; partial column sums of dne (gradient of the noise-embedding bias)
__global__ void __launch_bounds__(FB_COLS) hadamard_bwd_kernel(const float* __restrict__ du, const float* __restrict__ h,
                                                               const float* __restrict__ ne, long long rows, RowValid rv,
                                                               float* __restrict__ dne, float* __restrict__ dhu,
                                                               float* __restrict__ part_bn) {
  const int j = blockIdx.x * FB_COLS + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * FB_ROWS;
  long long r1 = min(rows, r0 + FB_ROWS);
  if (rv.dyn) { const long long seg0 = (r0 / rv.seg) * rv.seg; r1 = min(r1, seg0 + min(*rv.dyn, rv.seg)); }
  float sb = 0.f;
#pragma unroll 4
  for (long long row = r0; row < r1; ++row) {
    const size_t idx = (size_t)row * GAC_E + j;
    const float d = du[idx];
    const float a = d * h[idx];
    dne[idx] = a;
    dhu[idx] = d * ne[idx];
    sb += a;
  }
  part_bn[(size_t)blockIdx.y * GAC_E + j] = sb;
}

// dst[n] = (accumulate ? dst[n] : 0) + sum_{s < nsplit} part[s*N + n], fixed order.
// block (32, 32): 32 row lanes each sum every 32nd partial with 4-way ILP, then a fixed-order smem fold.
// gridDim.y > 1: slice y sums the partials [y * per, (y + 1) * per) into dst + y * N (a scratch row per slice; a second launch
// with gridDim.y == 1 folds the slices) -- with thousands of partial rows a one-slice launch keeps 13 SMs busy for 50 us.
constexpr int RC_LANES = 32;
__global__ void __launch_bounds__(32 * RC_LANES) reduce_cols_kernel(const float* __restrict__ part, int nsplit, int N,
                                                                   float* __restrict__ dst, int accumulate) {
  __shared__ float sh[RC_LANES][33];
  const int n = blockIdx.x * 32 + threadIdx.x, ly = threadIdx.y;
  if (gridDim.y > 1) {
    const int per = (nsplit + (int)gridDim.y - 1) / (int)gridDim.y, lo = (int)blockIdx.y * per;
    part += (size_t)lo * N; dst += (size_t)blockIdx.y * N;
    nsplit = max(0, min(per, nsplit - lo));
  }
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (n < N) {
    int s = ly;
    for (; s + 3 * RC_LANES < nsplit; s += 4 * RC_LANES) {
      s0 += part[(size_t)s * N + n]; s1 += part[(size_t)(s + RC_LANES) * N + n];
      s2 += part[(size_t)(s + 2 * RC_LANES) * N + n]; s3 += part[(size_t)(s + 3 * RC_LANES) * N + n];
    }
    for (; s < nsplit; s += RC_LANES) s0 += part[(size_t)s * N + n];
  }
  sh[ly][threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ly == 0 && n < N) {
    float t = accumulate ? dst[n] : 0.f;
#pragma unroll
    for (int k = 0; k < RC_LANES; ++k) t += sh[k][threadIdx.x];
    dst[n] = t;
  }
}

// ---- out[r] = f(dot(x[r,:n], w) + b): one warp per row.  mode 0: identity, 1: tanh. ----
// Twin variant takes a second (x2, w2, b2) and writes min (Critic.__call__, networks.py:388).
__global__ void rowdot_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ b,
                              const float* __restrict__ x2, const float* __restrict__ w2, const float* __restrict__ b2,
                              int n, int rows, int mode, float* __restrict__ out, int inc_out, float* __restrict__ out2,
                              int inc_out2, RowValid rv) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows || !rv(row)) return;
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s = fmaf(x[(size_t)row * ldx + j], w[j], s);
  s = warp_sum(s) + b[0];
  if (x2) {
    float t = 0.f;
    for (int j = lane; j < n; j += 32) t = fmaf(x2[(size_t)row * ldx + j], w2[j], t);
    t = warp_sum(t) + b2[0];
    s = fminf(s, t);
  }
  if (mode == 1) s = tanhf(s);
  if (lane == 0) {
    out[(size_t)row * inc_out] = s;
    if (out2) out2[(size_t)row * inc_out2] = s;
  }
}

// the training pass's head output for every step in one launch: blockIdx.y = step d, rows step-major (d * seg + i);
// out[i * A + d] = tanh(dot(x[d * seg + i], w) + b) (and the same to out2 when given)
__global__ void rowdot_steps_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ b, int n, int seg,
                                    int A, float* __restrict__ out, float* __restrict__ out2, RowValid rv) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, d = blockIdx.y;
  if (row >= seg || !rv(row)) return;
  const float* xr = x + ((size_t)d * seg + row) * ldx;
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s = fmaf(xr[j], w[j], s);
  s = tanhf(warp_sum(s) + b[0]);
  if (lane == 0) {
    out[(size_t)row * A + d] = s;
    if (out2) out2[(size_t)row * A + d] = s;
  }
}

// ---- consumer of dot_gemm_tc_kernel's tile partials for the actor head (networks.py:221): a = tanh(sum_t part[r][t] + b2),
// written to actions[r*inc_out]; when cosb != nullptr also the cosine basis of a (the next dimension's action
// embedding input, networks.py:44-47 / 208-210), 64 threads per row. ----
__global__ void head_finish_kernel(const float* __restrict__ part, int ntiles, const float* __restrict__ b2, float* __restrict__ out,
                                   int inc_out, float* __restrict__ cosb, int rows) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // 16 threads per row, 4 basis functions each
  const int r = idx >> 4, i = (idx & 15) * 4;
  if (r >= rows) return;
  float s = part[(size_t)r * ntiles];
  for (int t = 1; t < ntiles; ++t) s += part[(size_t)r * ntiles + t];
  const float a = tanhf(s + b2[0]);
  if (i == 0) out[(size_t)r * inc_out] = a;
  if (cosb)
    reinterpret_cast<float4*>(cosb)[idx] = make_float4(cosf(__fmul_rn(a, c_ipi[i])), cosf(__fmul_rn(a, c_ipi[i + 1])),
                                                       cosf(__fmul_rn(a, c_ipi[i + 2])), cosf(__fmul_rn(a, c_ipi[i + 3])));
}

// ---- first hidden layer of an FNN over concat(state, action) with the state half hoisted per unique state
// (networks.py:385-388): h0[r] = lrelu_0.01( sx0[sidx[r]] + a[r] * W0[S:] ), sx0 = s_u W0[:S] + b0 (n_states rows).
// One thread per (row, 4 columns): A FMAs per output instead of a (S + A)-deep GEMM over every candidate row. ----
__global__ void fnn_l1_gather_kernel(const float* __restrict__ sx0, const int* __restrict__ sidx, const float* __restrict__ actions,
                                     const float* __restrict__ w0a, int A, int H, int rows, float* __restrict__ h0) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int hq = H >> 2;
  if (idx >= (long long)rows * hq) return;
  const int r = (int)(idx / hq), c = (int)(idx - (long long)r * hq) * 4;
  float4 v = *reinterpret_cast<const float4*>(sx0 + (size_t)sidx[r] * H + c);
  const float* a = actions + (size_t)r * A;
  for (int i = 0; i < A; ++i) {
    const float ai = a[i];
    const float4 w = *reinterpret_cast<const float4*>(w0a + (size_t)i * H + c);
    v.x = fmaf(ai, w.x, v.x); v.y = fmaf(ai, w.y, v.y); v.z = fmaf(ai, w.z, v.z); v.w = fmaf(ai, w.w, v.w);
  }
  v.x = v.x > 0.f ? v.x : 0.01f * v.x; v.y = v.y > 0.f ? v.y : 0.01f * v.y;
  v.z = v.z > 0.f ? v.z : 0.01f * v.z; v.w = v.w > 0.f ? v.w : 0.01f * v.w;
  *reinterpret_cast<float4*>(h0 + (size_t)r * H + c) = v;
}

// ---- consumer of the tile partials of the twin critics' fused tails (networks.py:388): q = min over the two nets of
// (sum_t part[r][t] + b2) ----
__global__ void critic_finish_kernel(const float* __restrict__ part1, const float* __restrict__ part2, int ntiles, const float* __restrict__ b2a,
                                     const float* __restrict__ b2b, float* __restrict__ q, int rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float s1 = part1[(size_t)r * ntiles], s2 = part2[(size_t)r * ntiles];
  for (int t = 1; t < ntiles; ++t) { s1 += part1[(size_t)r * ntiles + t]; s2 += part2[(size_t)r * ntiles + t]; }
  q[r] = fminf(s1 + b2a[0], s2 + b2b[0]);
}

// ---- bound of the first hidden layer of an FNN over concat(state, action) inputs, |action| <= 1:
// |h0_j| <= sum_{i<S} |w0_ij| max_b |s_bi| + sum_{i>=S} |w0_ij| + |b0_j|   (LeakyReLU is 1-Lipschitz through 0).
// One block; result (float bits) max-ed into *slot, +inf on non-finite input. ----
__global__ void fnn_bound_kernel(const float* __restrict__ states, int n, int S, int W, const float* __restrict__ w0,
                                 const float* __restrict__ b0, int H, unsigned int* __restrict__ slot) {
  extern __shared__ float xmax[];               // W floats (as uint bit patterns while reducing)
  unsigned int* xbits = reinterpret_cast<unsigned int*>(xmax);
  for (int i = threadIdx.x; i < W; i += blockDim.x) xbits[i] = i < S ? 0u : __float_as_uint(1.f);
  __syncthreads();
  for (int e = threadIdx.x; e < n * S; e += blockDim.x) {
    const float a = fabsf(states[e]);
    atomicMax(&xbits[e % S], (a <= 3e38f) ? __float_as_uint(a) : 0x7f800000u);
  }
  __syncthreads();
  float mb = 0.f;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    float sacc = fabsf(b0[j]);
    for (int i = 0; i < W; ++i) sacc += fabsf(w0[(size_t)i * H + j]) * xmax[i];
    if (!(sacc <= 3e38f)) sacc = __int_as_float(0x7f800000);
    mb = fmaxf(mb, sacc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
  if ((threadIdx.x & 31) == 0) atomicMax(slot, __float_as_uint(mb));
}

// ---- agent.py:189-190: a = clip(a + 0.01*n, -1, 1) ----
__global__ void noise_clip_kernel(const float* __restrict__ a, const float* __restrict__ nrm, float scale, long long n,
                                  float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fminf(fmaxf(a[i] + nrm[i] * scale, -1.f), 1.f);
}

// ---- networks.py:484-485: v_critic[b] = mean_k q[b*K + k] (state-major rows) ----
__global__ void mean_k_kernel(const float* __restrict__ q, int K, int B, float* __restrict__ out) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += q[(size_t)b * K + k];
  s = warp_sum(s);
  if (lane == 0) out[b] = s / (float)K;
}

__global__ void gather_kernel(const float* __restrict__ src, const int* __restrict__ idx, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}

// ---- networks.py:415: yQ = r + gamma * V_tgt(s') * (1 - done) ----
__global__ void td_target_kernel(const float* __restrict__ r, const float* __restrict__ vnext, const float* __restrict__ done,
                                 float gamma, int B, float* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) y[i] = r[i] + gamma * vnext[i] * (1.f - done[i]);
}

// ---- mse over a size-1 axis, gradient of the SUM (F5): dout = 2(pred-y); loss = sum (pred-y)^2.
// single block (B is small); deterministic ----
__global__ void td_dout_kernel(const float* __restrict__ pred, const float* __restrict__ y, int B, float* __restrict__ dout,
                               float* __restrict__ loss) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float d = pred[i] - y[i];
    dout[i] = 2.f * d;
    acc += d * d;
  }
  const float t = block_sum(acc, sh);
  if (threadIdx.x == 0) *loss = t;
}

// ---- dx[r][j] = dout[r] * w[j] * (y[r][j] > 0 ? 1 : alpha): backward of a width-1 Dense
// through the LeakyReLU of the layer below (FNN layer 3, AIQN dense_layer_2). ----
__global__ void outer_gate_kernel(const float* __restrict__ dout, const float* __restrict__ w, const float* __restrict__ y,
                                  int n, long long rows, float alpha, float* __restrict__ dx, RowValid rv) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * n) return;
  const long long r = idx / n;
  const int j = idx % n;
  if (!rv((int)r)) return;
  dx[idx] = dout[r] * w[j] * (y[idx] > 0.f ? 1.f : alpha);
}

// ---- AIQN head backward, step-major: dop[d*seg+r] = dpred[r][d] * (1 - pred[r][d]^2) ----
__global__ void head_dop_kernel(const float* __restrict__ dpred, const float* __restrict__ pred, int A, int seg,
                                const int* dyn, float* __restrict__ dop) {
  const int nvalid = dyn ? *dyn : seg;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)A * seg) return;
  const int d = idx / seg, r = idx % seg;
  if (r >= nvalid) return;
  const float p = pred[(size_t)r * A + d];
  dop[idx] = dpred[(size_t)r * A + d] * (1.f - p * p);
}

// ---- out = a * b on live rows (u = ne * h, networks.py:220/:265) ----
// out = a * b over (rows, 400) activations, 128-bit accesses; dead rows (row % seg >= *dyn) are skipped
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, int seg, const int* dyn,
                           float* __restrict__ out) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // float4 index; GAC_E / 4 = 100 per row
  if (i4 * 4 >= n) return;
  if (dyn && (int)((i4 / (GAC_E / 4)) % seg) >= *dyn) return;
  const float4 x = reinterpret_cast<const float4*>(a)[i4], y = reinterpret_cast<const float4*>(b)[i4];
  reinterpret_cast<float4*>(out)[i4] = make_float4(x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w);
}

// row -> state index of the 3P candidate rows of one step: [0,P) state-major
// (networks.py:464-467), [P,3P) sample-major (agent.py:186)
__global__ void sidx3_kernel(int* __restrict__ s, int P, int B, int K) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= 3 * P) return;
  s[r] = r < P ? r / K : (r % P) % B;
}

// ---- column sums with optional row weights: part[blk][n] = sum_{r in blk} w[r] * X[r][n] ----
// (bias gradients, the width-1 layer's weight gradient).  Deterministic two-stage reduction.
constexpr int COLSUM_ROWS = 256;
// rpb rows per block (256 for the long time-parallel arrays, 16 for the B-row TD tensors: a single block
// walking 256 rows serially is latency bound at ~50 us)
__global__ void colsum_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ w, int N, long long rows,
                              RowValid rv, float* __restrict__ part, int rpb) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long r0 = (long long)blockIdx.y * rpb;
  long long r1 = min(rows, r0 + rpb);
  // a block of 256 rows never straddles a segment (seg is a multiple of 128 and of 256 or the
  // block is clipped below): live rows of this block are a prefix [r0, r_live)
  if (rv.dyn) {
    const long long seg0 = (r0 / rv.seg) * rv.seg;
    const long long seg_end = seg0 + rv.seg;
    const long long live_end = seg0 + min(*rv.dyn, rv.seg);
    if (r1 > seg_end) {           // straddling block (seg not a multiple of 256): fall back to per-row checks
      float s = 0.f;
      for (long long r = r0; r < r1; ++r)
        if (rv((int)r)) { const float x = X[(size_t)r * ldx + n]; s += w ? w[r] * x : x; }
      part[(size_t)blockIdx.y * N + n] = s;
      return;
    }
    r1 = min(r1, live_end);
  }
  // 4 independent accumulators: 4 loads in flight per thread (the serial version was latency bound)
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  long long r = r0;
  for (; r + 3 < r1; r += 4) {
    const float x0 = X[(size_t)r * ldx + n], x1 = X[(size_t)(r + 1) * ldx + n];
    const float x2 = X[(size_t)(r + 2) * ldx + n], x3 = X[(size_t)(r + 3) * ldx + n];
    if (w) { s0 += w[r] * x0; s1 += w[r + 1] * x1; s2 += w[r + 2] * x2; s3 += w[r + 3] * x3; }
    else { s0 += x0; s1 += x1; s2 += x2; s3 += x3; }
  }
  for (; r < r1; ++r) { const float x = X[(size_t)r * ldx + n]; s0 += w ? w[r] * x : x; }
  part[(size_t)blockIdx.y * N + n] = (s0 + s1) + (s2 + s3);
}

// ---- quantile Huber loss + gradient (networks.py:113-120; App. A.6), elementwise Huber ----
// dpred = |tau - 1[u>0]| * w[r] * clip(u,-1,1) * inv_norm ; loss partials per block.
__global__ void qhuber_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ taus,
                              const float* __restrict__ w, int A, int max_rows, const int* dyn, float inv_norm,
                              float* __restrict__ dpred, float* __restrict__ part) {
  __shared__ float sh[32];
  const int nvalid = dyn ? min(*dyn, max_rows) : max_rows;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (idx < (long long)nvalid * A) {
    const int r = idx / A;
    const float u = pred[idx] - target[idx];
    const float ind = u > 0.f ? 1.f : 0.f;
    const float au = fabsf(u), qq = fminf(au, 1.f);
    const float hub = 0.5f * qq * qq + (au - qq);
    const float coef = fabsf(taus[idx] - ind) * (w ? w[r] : 1.f);
    l = coef * hub * inv_norm;
    dpred[idx] = coef * fminf(fmaxf(u, -1.f), 1.f) * inv_norm;
  }
  const float t = block_sum(l, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

// single-block ordered sum of partials: out = (accumulate ? out : 0) + sum part[0..n)
__global__ void final_sum_kernel(const float* __restrict__ part, int n, float* __restrict__ out, int accumulate) {
  __shared__ float sh[32];
  // fixed assignment: thread t sums part[t], part[t+blockDim], ... then fixed-order block sum
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  const float t = block_sum(s, sh);
  if (threadIdx.x == 0) *out = (accumulate ? *out : 0.f) + t;
}

// ------------------------------------------------------------------------------------------
// Advantage filter: stable compaction of {i : q[i] > v[i]} (agent.py:207-215).
// 1024 candidates per block; ballot + popc inside a warp, smem scan across the 32 warps,
// a single-block scan across blocks.  NaN compares false, ties are excluded (strict >).
// ------------------------------------------------------------------------------------------
constexpr int CMP_BLOCK = 1024;
__global__ void compact_count_kernel(const float* __restrict__ q, const float* __restrict__ v, int n, int* __restrict__ blk_count) {
  __shared__ int sh[32];
  const int i = blockIdx.x * CMP_BLOCK + threadIdx.x;
  const bool f = i < n && q[i] > v[i];
  const unsigned b = __ballot_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = __popc(b);
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 32; ++w) t += sh[w];
    blk_count[blockIdx.x] = t;
  }
}
// exclusive scan of block counts (in place) and total -> *m_out
__global__ void compact_scan_kernel(int* __restrict__ blk_count, int nblk, int* __restrict__ m_out) {
  __shared__ int sh[1024];
  const int per = (nblk + 1023) / 1024;
  const int b0 = threadIdx.x * per, b1 = min(nblk, b0 + per);
  int s = 0;
  for (int b = b0; b < b1; ++b) s += blk_count[b];
  sh[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int t = 0; t < 1024; ++t) { const int c = sh[t]; sh[t] = run; run += c; }
    *m_out = run;
  }
  __syncthreads();
  int run = sh[threadIdx.x];
  for (int b = b0; b < b1; ++b) { const int c = blk_count[b]; blk_count[b] = run; run += c; }
}
__global__ void compact_scatter_kernel(const float* __restrict__ q, const float* __restrict__ v, int n,
                                       const int* __restrict__ blk_off, int64_t* __restrict__ idx_out,
                                       int* __restrict__ pos_out, float* __restrict__ adv_out,
                                       const float* __restrict__ actions, int A, float* __restrict__ actions_out,
                                       const int* __restrict__ sidx, int* __restrict__ sidx_out,
                                       float* __restrict__ blk_adv) {
  __shared__ int shc[32];
  __shared__ float shf[32];
  const int i = blockIdx.x * CMP_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float adv = 0.f;
  bool f = false;
  if (i < n) { const float qq = q[i], vv = v[i]; f = qq > vv; adv = qq - vv; }
  const unsigned b = __ballot_sync(0xffffffffu, f);
  if (lane == 0) shc[w] = __popc(b);
  __syncthreads();
  int woff = 0;
  for (int k = 0; k < w; ++k) woff += shc[k];
  const int pos = blk_off[blockIdx.x] + woff + __popc(b & ((1u << lane) - 1u));
  if (i < n && pos_out) pos_out[i] = pos;
  if (f) {
    idx_out[pos] = (int64_t)i;
    if (adv_out) adv_out[pos] = adv;
    if (sidx_out) sidx_out[pos] = sidx[i];
    if (actions_out) for (int d = 0; d < A; ++d) actions_out[(size_t)pos * A + d] = actions[(size_t)i * A + d];
  }
  const float t = block_sum(f ? adv : 0.f, shf);
  if (threadIdx.x == 0) blk_adv[blockIdx.x] = t;
}

// ---- per-unique-state reduction of d(sx): dsx_u[b][n] (+)= sum over live rows j of this chunk with sidx[j] == b of
// sum_d dmx[(d*seg + j)][n].  The rows are first grouped by state with a STABLE counting sort (the live list mixes a
// state-major and a sample-major tiling, so it is not sorted): per-block histograms, a scan over (state, block), and a
// scatter where a row's slot is base[block][state] + its rank among the block's earlier rows of the same state.
// Every state's rows end up in ascending row order, so the float summation order is fixed. ----
constexpr int SEG_BLOCK = 256;
__global__ void seg_hist_kernel(const int* __restrict__ sidx, int seg, const int* dyn, int nstates, int* __restrict__ hist) {
  extern __shared__ int sh_hist[];
  for (int i = threadIdx.x; i < nstates; i += blockDim.x) sh_hist[i] = 0;
  __syncthreads();
  const int j = blockIdx.x * SEG_BLOCK + threadIdx.x;
  const int nvalid = dyn ? min(*dyn, seg) : seg;
  if (j < nvalid) {
    const int b = sidx[j];
    if (b >= 0 && b < nstates) atomicAdd(&sh_hist[b], 1);       // integer: order-independent
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nstates; i += blockDim.x) hist[(size_t)blockIdx.x * nstates + i] = sh_hist[i];
}
// hist[blk][b] -> exclusive base of (blk, b) in the grouped list; info = [start[nstates] | end[nstates]]
__global__ void seg_scan_kernel(int* __restrict__ hist, int nblk, int nstates, int* __restrict__ info) {
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nstates; b0 += blockDim.x) {
    const int b = b0 + threadIdx.x;
    int tot = 0;
    if (b < nstates)
      for (int k = 0; k < nblk; ++k) { const int v = hist[(size_t)k * nstates + b]; hist[(size_t)k * nstates + b] = tot; tot += v; }
    // exclusive scan of tot over the states of this pass (blockDim.x <= 1024): warp scan + warp totals
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < w; ++k) woff += wsum[k];
    const int start = carry + woff + inc - tot;
    if (b < nstates) {
      info[b] = start; info[nstates + b] = start + tot;
      for (int k = 0; k < nblk; ++k) hist[(size_t)k * nstates + b] += start;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = start + tot;
    __syncthreads();
  }
}
__global__ void seg_scatter_kernel(const int* __restrict__ sidx, int seg, const int* dyn, int nstates, const int* __restrict__ base,
                                   int* __restrict__ perm) {
  __shared__ int sb[SEG_BLOCK];
  const int j = blockIdx.x * SEG_BLOCK + threadIdx.x;
  const int nvalid = dyn ? min(*dyn, seg) : seg;
  const int b = j < nvalid ? sidx[j] : -1;
  sb[threadIdx.x] = b;
  __syncthreads();
  if (b < 0 || b >= nstates) return;
  int rank = 0;
  for (int t = 0; t < (int)threadIdx.x; ++t) rank += sb[t] == b;
  perm[base[(size_t)blockIdx.x * nstates + b] + rank] = j;
}
__global__ void segment_sum_kernel(const float* __restrict__ dmx, const int* __restrict__ perm, const int* __restrict__ info, int A, int seg,
                                   int nstates, int accumulate, float* __restrict__ dsx_u) {
  const int b = blockIdx.x;
  const int n = blockIdx.y * 256 + threadIdx.x;     // 256 threads, 5 slabs cover 1200 (tail masked)
  const int j0 = info[b], j1 = info[nstates + b];
  __shared__ int rows[256];
  float s = (accumulate && n < GAC_G) ? dsx_u[(size_t)b * GAC_G + n] : 0.f;
  for (int base = j0; base < j1; base += 256) {
    const int cnt = min(256, j1 - base);
    __syncthreads();
    if ((int)threadIdx.x < cnt) rows[threadIdx.x] = perm[base + threadIdx.x];
    __syncthreads();
    if (n < GAC_G)
      for (int d = 0; d < A; ++d) {
        const float* p = dmx + (size_t)d * seg * GAC_G + n;
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
        int e = 0;
        for (; e + 3 < cnt; e += 4) {
          t0 += p[(size_t)rows[e] * GAC_G]; t1 += p[(size_t)rows[e + 1] * GAC_G];
          t2 += p[(size_t)rows[e + 2] * GAC_G]; t3 += p[(size_t)rows[e + 3] * GAC_G];
        }
        for (; e < cnt; ++e) t0 += p[(size_t)rows[e] * GAC_G];
        s += (t0 + t1) + (t2 + t3);
      }
  }
  if (n < GAC_G) dsx_u[(size_t)b * GAC_G + n] = s;
}

// ------------------------------------------------------------------------------------------
// Adam (Keras OptimizerV2 formula) fused with the Polyak target update
// (helpers.py:86) over the flat arenas; 128-bit accesses, 36 B/param of HBM traffic.
// ------------------------------------------------------------------------------------------
struct AdamNets {
  long long begin[GAC_NNETS + 1];
};
// one thread: advances step counters and derives alpha_t per net in fp32 like Keras does
__global__ void adam_prep_kernel(int* __restrict__ adam_t, const int* __restrict__ skip, float lr, float b1, float b2,
                                 float* __restrict__ alpha) {
  const int n = threadIdx.x;
  if (n >= GAC_NNETS) return;
  if (skip && skip[n]) { alpha[n] = 0.f; return; }
  const int t = adam_t[n] + 1;
  adam_t[n] = t;
  const float b1p = powf(b1, (float)t), b2p = powf(b2, (float)t);
  alpha[n] = lr * sqrtf(1.f - b2p) / (1.f - b1p);
}
__global__ void adam_polyak_kernel(float4* __restrict__ p, float4* __restrict__ tg, float4* __restrict__ m, float4* __restrict__ v,
                                   const float4* __restrict__ g, AdamNets nets, const float* __restrict__ alpha,
                                   const float* __restrict__ gscale, const int* __restrict__ skip, float omb1, float omb2,
                                   float eps, float one_minus_rho, float rho) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = i4 * 4;
  if (i >= nets.begin[GAC_NNETS]) return;
  int net = 0;
#pragma unroll
  for (int k = 1; k < GAC_NNETS; ++k) net += (i >= nets.begin[k]);
  float4 pp = p[i4];
  if (!(skip && skip[net])) {
    const float a = alpha[net], sc = gscale ? gscale[net] : 1.f;
    float4 gg = g[i4], mm = m[i4], vv = v[i4];
    float* pf = &pp.x; float* gf = &gg.x; float* mf = &mm.x; float* vf = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gf[k] * sc;
      mf[k] = mf[k] + (gk - mf[k]) * omb1;
      vf[k] = vf[k] + (gk * gk - vf[k]) * omb2;
      pf[k] = pf[k] - a * mf[k] / (sqrtf(vf[k]) + eps);
    }
    m[i4] = mm; v[i4] = vv; p[i4] = pp;
  }
  float4 tt = tg[i4];
  tt.x = __fadd_rn(__fmul_rn(tt.x, one_minus_rho), __fmul_rn(pp.x, rho));
  tt.y = __fadd_rn(__fmul_rn(tt.y, one_minus_rho), __fmul_rn(pp.y, rho));
  tt.z = __fadd_rn(__fmul_rn(tt.z, one_minus_rho), __fmul_rn(pp.z, rho));
  tt.w = __fadd_rn(__fmul_rn(tt.w, one_minus_rho), __fmul_rn(pp.w, rho));
  tg[i4] = tt;
}

// helpers.update alone: t <- fl(fl(t*(1-rho)) + fl(s*rho)); vector body + scalar tail
__global__ void polyak_kernel(float* __restrict__ t, const float* __restrict__ s, long long n, float one_minus_rho, float rho) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = i4 * 4;
  if (i + 3 < n) {
    float4 tt = reinterpret_cast<float4*>(t)[i4];
    const float4 ss = reinterpret_cast<const float4*>(s)[i4];
    tt.x = __fadd_rn(__fmul_rn(tt.x, one_minus_rho), __fmul_rn(ss.x, rho));
    tt.y = __fadd_rn(__fmul_rn(tt.y, one_minus_rho), __fmul_rn(ss.y, rho));
    tt.z = __fadd_rn(__fmul_rn(tt.z, one_minus_rho), __fmul_rn(ss.z, rho));
    tt.w = __fadd_rn(__fmul_rn(tt.w, one_minus_rho), __fmul_rn(ss.w, rho));
    reinterpret_cast<float4*>(t)[i4] = tt;
  } else {
    for (long long k = i; k < n; ++k) t[k] = __fadd_rn(__fmul_rn(t[k], one_minus_rho), __fmul_rn(s[k], rho));
  }
}
__global__ void polyak_scalar_kernel(float* __restrict__ t, const float* __restrict__ s, long long n, float one_minus_rho, float rho) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] = __fadd_rn(__fmul_rn(t[i], one_minus_rho), __fmul_rn(s[i], rho));
}

// ---- IQN actor (StochasticActor, networks.py:304-323) glue ----
__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, int rows, int W, RowValid rv,
                                   float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * W) return;
  const int r = i / W, c = i % W;
  if (!rv(r)) return;                  // dead rows carry uninitialised indices
  out[i] = src[(size_t)idx[r] * W + c];
}
// merge = se[sidx ? sidx[r] : r] * ne   (networks.py:319)
__global__ void iqn_merge_kernel(const float* __restrict__ se, const int* __restrict__ sidx, const float* __restrict__ ne, int rows,
                                 int D, RowValid rv, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * D) return;
  const int r = i / D, c = i % D;
  if (!rv(r)) return;
  out[i] = se[(size_t)(sidx ? sidx[r] : r) * D + c] * ne[i];
}
// dne = dmerge * se * lrelu'(ne) ; dse = dmerge * ne * lrelu'(se)    (both LeakyReLU(0.01); sign(y) = sign(pre))
__global__ void iqn_dmerge_kernel(const float* __restrict__ dmerge, const float* __restrict__ se, const float* __restrict__ ne,
                                  long long n, int D, RowValid rv, float* __restrict__ dne, float* __restrict__ dse) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!rv((int)(i / D))) return;
  const float d = dmerge[i], s = se[i], e = ne[i];
  dne[i] = d * s * (e > 0.f ? 1.f : 0.01f);
  dse[i] = d * e * (s > 0.f ? 1.f : 0.01f);
}
// dop = dpred * (1 - out^2)   (tanh output layer)
__global__ void tanh_bwd_kernel(const float* __restrict__ dpred, const float* __restrict__ out, long long n, int A, RowValid rv,
                                float* __restrict__ dop) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!rv((int)(i / A))) return;
  const float o = out[i];
  dop[i] = dpred[i] * (1.f - o * o);
}
__global__ void scale_int_kernel(const int* __restrict__ src, int mul, int cap, int* __restrict__ dst) {
  if (threadIdx.x == 0) { dst[0] = min(*src, cap); dst[1] = min(*src, cap) * mul; }
}

// ---- ReplayBuffer.sample_batch (helpers.py:65-72): one gather for the five transition arrays ----
__global__ void replay_gather_kernel(const float* __restrict__ obs1, const float* __restrict__ obs2, const float* __restrict__ acts,
                                     const float* __restrict__ rews, const float* __restrict__ done, const int* __restrict__ idx, int n,
                                     int S, int A, float* __restrict__ s, float* __restrict__ sp, float* __restrict__ a,
                                     float* __restrict__ r, float* __restrict__ it) {
  const int W = 2 * S + A + 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * W) return;
  const int row = i / W, c = i % W;
  const size_t src = (size_t)idx[row];
  if (c < S) s[(size_t)row * S + c] = obs1[src * S + c];
  else if (c < 2 * S) sp[(size_t)row * S + (c - S)] = obs2[src * S + (c - S)];
  else if (c < 2 * S + A) a[(size_t)row * A + (c - 2 * S)] = acts[src * A + (c - 2 * S)];
  else if (c == 2 * S + A) r[row] = rews[src];
  else it[row] = done[src];
}

// ---- step glue: chunk row counts and the actor gradient normalisation ----
// rows_c[c] = clamp(M - c*chunk, 0, chunk)
__global__ void chunk_rows_kernel(const int* __restrict__ m, int chunk, int nchunks, int* __restrict__ rows_c) {
  const int M = *m;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += gridDim.x * blockDim.x) {
    const long long left = (long long)M - (long long)c * chunk;      // c * chunk can pass 2^31 with small chunks
    rows_c[c] = (int)(left < 0 ? 0 : (left > chunk ? chunk : left));
  }
}
// stats (after the cross-rank SUM) -> gscale[4], skip[4].  IQNActor.train: weights = adv/sum(adv)
// (linear) or 1 (boltzmann, F6); loss = mean over M*A (networks.py:91,93,120); agent.py:158 skip.
__global__ void apply_prep_kernel(const float* __restrict__ stats, int A, int mode, float* __restrict__ gscale, int* __restrict__ skip) {
  if (threadIdx.x != 0) return;
  const float M = stats[5], sadv = stats[4];
  for (int n = 0; n < GAC_NNETS; ++n) { gscale[n] = 1.f; skip[n] = 0; }
  if (M < 0.5f) { skip[GAC_NET_ACTOR] = 1; gscale[GAC_NET_ACTOR] = 0.f; return; }
  const float denom = M * (float)A;
  gscale[GAC_NET_ACTOR] = mode == 0 ? 1.f / (denom * sadv) : 1.f / denom;
}
__global__ void stats_m_kernel(const int* __restrict__ m, float* __restrict__ stats) {
  if (threadIdx.x == 0) { stats[5] = (float)*m; stats[6] = 0.f; stats[7] = 0.f; }
}

}  // namespace gac
