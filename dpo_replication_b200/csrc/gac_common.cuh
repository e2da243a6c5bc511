// Shared declarations for libgac_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "gac_b200.h"

#define GAC_E 400    // embedding / GRU width          (reference GAC/networks.py:160,165,166,169)
#define GAC_G 1200   // 3 GRU gates [z|r|h]
#define GAC_NB 64    // cosine basis functions         (networks.py:145)
#define GAC_H1 400   // FNN hidden 1                   (networks.py:379)
#define GAC_H2 300   // FNN hidden 2

namespace gac {

// ------------------------------------------------------------------------------------------
// error plumbing (never throws across the ABI)
// ------------------------------------------------------------------------------------------
extern thread_local char g_err[512];
inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}
#define GAC_CUDA(expr)                                                                    \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) return gac::fail(GAC_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)
#define GAC_LAUNCH_CHECK(name)                                                            \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) return gac::fail(GAC_ERR_CUDA, "launch %s: %s", name, cudaGetErrorString(e__)); \
  } while (0)
#define GAC_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != 0) return rc__;  \
  } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return cdiv(a, b) * b; }

// ------------------------------------------------------------------------------------------
// GEMM description shared by the SIMT and the tcgen05 kernels
//   C(M,N) = epilogue( A(M,K) * B(K,N) )
// epilogue order: + bias[n]  + add[addidx?addidx[m]:m][n]  -> activation -> * mul[m][n]
//                 -> * (gate[m][n] > 0 ? 1 : gate_alpha)  -> (+ C_old if accumulate)
// ------------------------------------------------------------------------------------------
enum { ACT_NONE = 0, ACT_LRELU = 1, ACT_LRELU2 = 2 /* lrelu(0.01) then lrelu(0.2): networks.py:196 */, ACT_TANH = 3 };

struct WgScales { float s_a, s_b, inv_unit; int ok; };

struct GemmArgs {
  int M, N, K;
  const float* A; int lda; int transA;  // transA: A(m,k) = A[k*lda + m]
  const float* B; int ldb; int transB;  // transB: B(k,n) = B[n*ldb + k]
  const unsigned char* Bpacked;        // optional: B pre-split/pre-swizzled by tc::pack_b_kernel (weights)
  int Bpacked_bn;                      // tile width the packed image was built for (selects the kernel)
  float* C; int ldc; int transC;        // transC: C(m,n) stored at C[n*ldc + m]
  const float* bias;
  const float* add; int ldadd; const int* addidx;
  const float* mul; int ldmul;
  const float* gate; int ldgate; float gate_alpha;
  int act; float alpha;
  int accumulate;
  // dynamic validity of activation rows: row r is live iff (r % seg_rows) < *dyn_rows.
  // valid_on_k = 0: r indexes M (forward / dgrad); 1: r indexes K (wgrad reduction).
  const int* dyn_rows; int seg_rows; int valid_on_k;
  // split-K: blockIdx.z = split, raw partial (no epilogue) stored at C + z*split_stride
  int splitk; long long split_stride;
  unsigned long long* dbg;               // optional timeline probe (tools/gemm_probe.py --timeline)
  // 3xFP16 variant of the weight-gradient GEMM (both operands MN-major activations with known magnitude bounds):
  // device struct {s_a, s_b, inv_unit, ok}, nullptr = 3xTF32
  const struct WgScales* f16;
  const unsigned char* Bpk16;            // 3xFP16 weight gradients: the X operand pre-split into MN-major fp16 tiles (wg_pack_x_kernel)
  int probe;                             // timing experiments only (env GAC_TC_PROBE; results are garbage when set)
};

inline GemmArgs gemm_args(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.splitk = 1; g.seg_rows = 1 << 30; g.gate_alpha = 1.f;
  return g;
}

__device__ __forceinline__ float lrelu_f(float x, float a) { return x > 0.f ? x : a * x; }

__device__ __forceinline__ float gemm_epilogue(const GemmArgs& g, float acc, int m, int n) {
  float v = acc;
  if (g.bias) v += g.bias[n];
  if (g.add) v += g.add[(size_t)(g.addidx ? g.addidx[m] : m) * g.ldadd + n];
  if (g.act == ACT_LRELU) v = lrelu_f(v, g.alpha);
  else if (g.act == ACT_LRELU2) v = lrelu_f(lrelu_f(v, 0.01f), 0.2f);
  else if (g.act == ACT_TANH) v = tanhf(v);
  if (g.mul) v *= g.mul[(size_t)m * g.ldmul + n];
  if (g.gate) v *= (g.gate[(size_t)m * g.ldgate + n] > 0.f ? 1.f : g.gate_alpha);
  if (g.accumulate) v += g.transC ? g.C[(size_t)n * g.ldc + m] : g.C[(size_t)m * g.ldc + n];
  return v;
}

}  // namespace gac
