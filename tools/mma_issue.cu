// Is a tcgen05.mma instruction's ~168 cycles (tools/mma_rate.cu) the tensor core's time or the issuing thread's?  Same
// resident-operand loop, but the issue stream is stripped to the bone: descriptors precomputed, six MMAs per asm block
// sharing one predicate, no per-instruction address arithmetic.  If N = 64 now runs much faster than N = 256, the tensor
// core was not the limit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include tools/mma_issue.cu -o tools/mma_issue
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../dpo_replication_b200/csrc/gac_common.cuh"
#include "../dpo_replication_b200/csrc/gemm_tc.cuh"
#include "../dpo_replication_b200/csrc/gru_step_tc.cuh"
#include "../dpo_replication_b200/csrc/gru_step_v2.cuh"

using namespace gac;
using namespace gac::tc;

__device__ __forceinline__ void mma6(uint32_t d, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 ah2, al2, bh2, bl2;\n"
      "setp.eq.b32 p, 1, 1;\n"
      "add.u64 ah2, %1, 2; add.u64 al2, %2, 2; add.u64 bh2, %3, 2; add.u64 bl2, %4, 2;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %3, %5, p;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %3, %5, p;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %4, %5, p;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ah2, bh2, %5, p;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], al2, bh2, %5, p;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ah2, bl2, %5, p;\n"
      "}\n" ::"r"(d), "l"(ah), "l"(al), "l"(bh), "l"(bl), "r"(idesc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) mma_issue_kernel(int N, int iters, int commit_every, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_done, bar_dummy;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem0 - smem_u32(smem_raw));
  const uint32_t a_tile = 128u * 64u, b_tile = (uint32_t)N * 64u, stage = 2 * a_tile + 2 * b_tile;
  for (uint32_t i = tid; i < (4 * stage) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x14001400u + ((i * 2654435761u) & 0x03ff03ffu);
  if (tid == 0) { mbar_init(smem_u32(&bar_done), 1); mbar_init(smem_u32(&bar_dummy), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint64_t ah[4], al[4], bh[4], bl[4];
    for (int s = 0; s < 4; ++s) {
      const uint32_t sa = smem0 + (uint32_t)s * stage;
      ah[s] = make_desc64(sa); al[s] = make_desc64(sa + a_tile); bh[s] = make_desc64(sa + 2 * a_tile); bl[s] = make_desc64(sa + 2 * a_tile + b_tile);
    }
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      mma6(tmem_d, ah[0], al[0], bh[0], bl[0], idesc);
      if (commit_every) mma_commit(smem_u32(&bar_dummy));
      mma6(tmem_d + 256u, ah[1], al[1], bh[1], bl[1], idesc);
      if (commit_every) mma_commit(smem_u32(&bar_dummy));
      mma6(tmem_d, ah[2], al[2], bh[2], bl[2], idesc);
      if (commit_every) mma_commit(smem_u32(&bar_dummy));
      mma6(tmem_d + 256u, ah[3], al[3], bh[3], bl[3], idesc);
      if (commit_every) mma_commit(smem_u32(&bar_dummy));
    }
    mma_commit(smem_u32(&bar_done));
    mbar_wait(smem_u32(&bar_done), 0);
    const long long t1 = clock64();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
    out[blockIdx.x * 4] = (unsigned long long)(t1 - t0); out[blockIdx.x * 4 + 1] = (unsigned long long)iters * 6; out[blockIdx.x * 4 + 2] = g1 - g0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512) : "memory");
}

// ---- tcgen05.ld cost beside running MMAs: reader warps loop over LDTM (x16 / x32 / x64 columns) + wait while warp 0 issues
// the stripped MMA stream (N = 0: no MMAs).  Reports cycles per load per warp and the aggregate TMEM read rate.
template <int X>
__device__ __forceinline__ uint32_t ldtm(uint32_t taddr) {
  uint32_t r[X];
  if constexpr (X == 16) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
  } else if constexpr (X == 32) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                   "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
  } else {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
          "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]),
          "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]),
          "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < X; ++i) x ^= r[i];
  return x;
}

template <int X>
__global__ void __launch_bounds__(544, 1) ldtm_kernel(int N, int mma_iters, int ld_warps, int ld_iters, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_done, bar_dummy;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem0 - smem_u32(smem_raw));
  const int Nn = N ? N : 64;
  const uint32_t a_tile = 128u * 64u, b_tile = (uint32_t)Nn * 64u, stage = 2 * a_tile + 2 * b_tile;
  for (uint32_t i = tid; i < (4 * stage) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x14001400u + ((i * 2654435761u) & 0x03ff03ffu);
  if (tid == 0) { mbar_init(smem_u32(&bar_done), 1); mbar_init(smem_u32(&bar_dummy), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  if (warp == 0) {
    if (lane == 0 && N) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint64_t ah[4], al[4], bh[4], bl[4];
      for (int s = 0; s < 4; ++s) {
        const uint32_t sa = smem0 + (uint32_t)s * stage;
        ah[s] = make_desc64(sa); al[s] = make_desc64(sa + a_tile); bh[s] = make_desc64(sa + 2 * a_tile); bl[s] = make_desc64(sa + 2 * a_tile + b_tile);
      }
      const long long t0 = clock64();
      for (int i = 0; i < mma_iters; i += 4) {
        mma6(tmem_d, ah[0], al[0], bh[0], bl[0], idesc); mma_commit(smem_u32(&bar_dummy));
        mma6(tmem_d, ah[1], al[1], bh[1], bl[1], idesc); mma_commit(smem_u32(&bar_dummy));
        mma6(tmem_d, ah[2], al[2], bh[2], bl[2], idesc); mma_commit(smem_u32(&bar_dummy));
        mma6(tmem_d, ah[3], al[3], bh[3], bl[3], idesc); mma_commit(smem_u32(&bar_dummy));
      }
      mma_commit(smem_u32(&bar_done));
      mbar_wait(smem_u32(&bar_done), 0);
      out[blockIdx.x * 64] = (unsigned long long)(clock64() - t0); out[blockIdx.x * 64 + 1] = (unsigned long long)mma_iters * 6;
    }
    __syncwarp();
  } else if (warp <= ld_warps) {
    const int w = warp - 1, quarter = warp & 3;                    // a warp may only read the TMEM lanes of its quadrant
    const uint32_t ta = tmem_d + 256u + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(((w >> 2) * X) % 192);   // the buffer the MMAs do not write
    // let the MMAs get going
    const long long spin = clock64();
    while (clock64() - spin < 20000) {}
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < ld_iters; ++i) acc ^= ldtm<X>(ta);
    const long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x * 64 + 2 + w * 2] = (unsigned long long)(t1 - t0); out[blockIdx.x * 64 + 3 + w * 2] = acc; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512) : "memory");
}

template <int X>
static void run_ld(int N, int ld_warps, unsigned long long* out) {
  const int Nn = N ? N : 64, smem = 4 * (2 * 128 * 64 + 2 * Nn * 64) + 2048, ld_iters = 200, mma_iters = 600;
  cudaFuncSetAttribute(ldtm_kernel<X>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(out, 0, 148 * 64 * sizeof(unsigned long long));
  ldtm_kernel<X><<<148, 544, smem>>>(N, mma_iters, ld_warps, ld_iters, out);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("ldtm x%d N=%d: %s\n", X, N, cudaGetErrorString(err)); exit(1); }
  std::vector<unsigned long long> h(148 * 64);
  cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
  double cyc = 0, mc = 0, mn = 0;
  for (int b = 0; b < 148; ++b) { for (int w = 0; w < ld_warps; ++w) cyc += (double)h[b * 64 + 2 + 2 * w]; mc += (double)h[b * 64]; mn += (double)h[b * 64 + 1]; }
  const double per = cyc / (148.0 * ld_warps * ld_iters);
  printf("tcgen05.ld x%-2d, %2d reader warps, MMAs N=%3d: %7.1f cycles per load per warp, %6.1f B/clk/SM aggregate TMEM read", X, ld_warps, N, per,
         ld_warps * 32.0 * X * 4 / per);
  if (N) printf(", %6.1f cycles per MMA meanwhile", mc / mn);
  printf("\n");
}

int main() {
  unsigned long long* out;
  cudaMalloc(&out, 256 * 64 * sizeof(unsigned long long));
  if (getenv("MMA_ISSUE_LDTM")) {
    for (int N : {0, 240})
      for (int w : {4, 8, 12, 16}) { run_ld<16>(N, w, out); run_ld<32>(N, w, out); run_ld<64>(N, w, out); }
    return 0;
  }
  for (int commit : {0, 1})
    for (int N : {16, 64, 128, 192, 240, 256}) {
      const int smem = 4 * (2 * 128 * 64 + 2 * N * 64) + 2048;
      cudaFuncSetAttribute(mma_issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        mma_issue_kernel<<<148, 128, smem>>>(N, 1000, commit, out);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(err)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
      }
      std::vector<unsigned long long> h(148 * 4);
      cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
      double cyc = 0, n = 0, ns = 0;
      for (int i = 0; i < 148; ++i) { cyc += h[i * 4]; n += h[i * 4 + 1]; ns += h[i * 4 + 2]; }
      printf("N=%3d commit per stage=%d: %6.1f cycles per MMA, SM clock %4.0f MHz, %7.1f TFLOP/s chip-wide\n", N, commit, cyc / n, 1e3 * cyc / ns,
             2.0 * 128 * N * 16 * n / (ms * 1e-3) / 1e12);
    }
  return 0;
}
