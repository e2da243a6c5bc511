// tcgen05.mma issue-rate microbenchmark (sm_100a): how many SM cycles does one M = 128 (cta_group::1) or M = 256
// (cta_group::2) kind::f16 instruction with shared-memory operands take, as a function of N, of the operand swizzle
// (64-byte rows = 32-element K blocks, 128-byte rows = 64-element K blocks) and of what else uses the SM's shared-memory
// pipe (bulk copies landing in other stages, an epilogue-like ld/st.shared + tcgen05.ld loop)?  Every CTA of a full-chip grid
// issues the 3-pass pattern of the GRU step kernel (hi*hi, lo*hi, hi*lo per 16-element K step) on resident operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include tools/mma_rate.cu -o tools/mma_rate
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../dpo_replication_b200/csrc/gac_common.cuh"
#include "../dpo_replication_b200/csrc/gemm_tc.cuh"
#include "../dpo_replication_b200/csrc/gru_step_tc.cuh"
#include "../dpo_replication_b200/csrc/gru_step_v2.cuh"

using namespace gac;
using namespace gac::tc;

struct RateArgs {
  int N, sw, tiles, copies, epi, mn;   // mn: both operands MN-major (the weight-gradient layout: 64-element groups x 32 k-rows x 128 bytes)
  const unsigned char* src;        // >= 64 KB of global memory per CTA for the bulk copies
  unsigned long long* out;         // per CTA: [cycles, mmas, ns]
};

__device__ __forceinline__ void mma_f16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_cg2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((unsigned short)3)
               : "memory");
}

template <int CG>
__global__ void __launch_bounds__(384, 1) mma_rate_kernel(const RateArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_done, bar_dummy, bar_cp[2];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop_flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t crank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem0 - smem_u32(smem_raw));
  const int rowb = a.sw;                                        // bytes per operand row in a K block
  const int ncta = CG == 2 ? a.N / 2 : a.N;
  const uint32_t a_tile = a.mn ? 2u * 4096u : 128u * rowb, b_tile = a.mn ? (uint32_t)((ncta + 63) / 64) * 4096u : (uint32_t)ncta * rowb;
  const uint32_t stage = 2 * a_tile + 2 * b_tile;
  const int nstages = a.sw == 64 ? 4 : 1;
  const uint32_t cp_bytes = stage < 32768u ? stage : 32768u;
  const uint32_t cp_base = smem0 + nstages * stage;             // two landing buffers for the bulk copies
  const uint32_t epi_base = cp_base + 2 * cp_bytes;             // epilogue staging
  for (uint32_t i = tid; i < (nstages * stage) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x14001400u + ((i * 2654435761u) & 0x03ff03ffu);
  if (tid == 0) {
    mbar_init(smem_u32(&bar_done), 1); mbar_init(smem_u32(&bar_dummy), 1); mbar_init(smem_u32(&bar_cp[0]), 1); mbar_init(smem_u32(&bar_cp[1]), 1);
    stop_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  const int ksteps_per_stage = a.mn ? 2 : rowb / 32, ksteps = 50;   // one GRU tile: 2 x 25 K steps of 16 elements

  if (warp == 0) {
    if (lane == 0 && crank == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)((CG == 2 ? 256 : 128) >> 4) << 24) | (a.mn ? (1u << 15) | (1u << 16) : 0u);
      unsigned long long g0, g1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
      const long long t0 = clock64();
      long long n = 0;
      for (int t = 0; t < a.tiles; ++t) {
        const uint32_t d0 = tmem_d + (uint32_t)(t & 1) * 256u;
        for (int k = 0; k < ksteps; ++k) {
          const int s = (k / ksteps_per_stage) % nstages, j = k % ksteps_per_stage;
          const uint32_t sa = smem0 + (uint32_t)s * stage;
          const uint64_t adv = a.mn ? (uint64_t)(j * 128) : (uint64_t)(j * 2);
          auto mk = [&](uint32_t addr) { return a.mn ? make_desc_mn16(addr) : (a.sw == 64 ? make_desc64(addr) : make_desc(addr)); };
          const uint64_t ah = mk(sa) + adv, al = mk(sa + a_tile) + adv;
          const uint64_t bh = mk(sa + 2 * a_tile) + adv, bl = mk(sa + 2 * a_tile + b_tile) + adv;
          if (CG == 2) {
            mma_f16_cg2(d0, ah, bh, idesc, k ? 1u : 0u); mma_f16_cg2(d0, al, bh, idesc, 1u); mma_f16_cg2(d0, ah, bl, idesc, 1u);
          } else {
            mma_f16(d0, ah, bh, idesc, k ? 1u : 0u); mma_f16(d0, al, bh, idesc, 1u); mma_f16(d0, ah, bl, idesc, 1u);
          }
          n += 3;
          if (j == ksteps_per_stage - 1) { if (CG == 2) mma_commit_cg2(smem_u32(&bar_dummy)); else mma_commit(smem_u32(&bar_dummy)); }
        }
      }
      if (CG == 2) mma_commit_cg2(smem_u32(&bar_done)); else mma_commit(smem_u32(&bar_done));
      mbar_wait(smem_u32(&bar_done), 0);
      const long long t1 = clock64();
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
      stop_flag = 1;
      a.out[blockIdx.x * 4 + 0] = (unsigned long long)(t1 - t0); a.out[blockIdx.x * 4 + 1] = (unsigned long long)n; a.out[blockIdx.x * 4 + 2] = g1 - g0;
    } else if (lane == 0) {
      mbar_wait(smem_u32(&bar_done), 0);
      stop_flag = 1;
    }
    __syncwarp();
  } else if (warp == 1) {
    // bulk copies of one stage's bytes into two landing buffers, two in flight, until the MMAs are done
    if (lane == 0 && a.copies) {
      const unsigned char* src = a.src + (size_t)blockIdx.x * 65536;
      uint32_t ph[2] = {0, 0};
      for (int b = 0; b < 2; ++b) { mbar_expect_tx_arrive(smem_u32(&bar_cp[b]), cp_bytes); bulk_g2s(cp_base + b * cp_bytes, src + (size_t)b * 8192, cp_bytes, smem_u32(&bar_cp[b])); }
      int i = 0;
      while (!stop_flag) {
        const int b = i & 1;
        mbar_wait(smem_u32(&bar_cp[b]), ph[b]); ph[b] ^= 1u;
        mbar_expect_tx_arrive(smem_u32(&bar_cp[b]), cp_bytes);
        bulk_g2s(cp_base + b * cp_bytes, src + (size_t)((i * 4096) & 16383), cp_bytes, smem_u32(&bar_cp[b]));
        ++i;
      }
      for (int b = 0; b < 2; ++b) mbar_wait(smem_u32(&bar_cp[b]), ph[b]);
    }
    __syncwarp();
  } else if (warp >= 4 && a.epi) {
    // epilogue-like load on the shared-memory pipe: TMEM -> registers -> smem rows, then smem -> registers (float4)
    const int quarter = warp & 3, ew = warp - 4;
    float* stg = reinterpret_cast<float*>(sm + (epi_base - smem0)) + (size_t)(ew * 32 + lane) * 36;
    float acc = 0.f;
    while (!stop_flag) {
      float v[16];
      tmem_ld16(tmem_d + ((uint32_t)(quarter * 32) << 16) + 256u + (uint32_t)((ew >> 2) * 16), v);
#pragma unroll
      for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(stg + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
      __syncwarp();
#pragma unroll
      for (int e = 0; e < 16; e += 4) { const float4 x = *reinterpret_cast<const float4*>(stg + e); acc += x.x + x.y + x.z + x.w; }
      __syncwarp();
    }
    if (acc == 123.456f) a.out[0] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512) : "memory");
  }
}

static void run(int cg, int N, int sw, int copies, int epi, const unsigned char* src, unsigned long long* out_d, int mn = 0) {
  RateArgs a{N, sw, 40, copies, epi, mn, src, out_d};
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  sms &= ~1;
  const int ncta = cg == 2 ? N / 2 : N;
  const int stage = mn ? 2 * 8192 + 2 * ((ncta + 63) / 64) * 4096 : 2 * 128 * sw + 2 * ncta * sw, nstages = sw == 64 ? 4 : 1;
  const int smem = nstages * stage + 2 * (stage < 32768 ? stage : 32768) + 12 * 32 * 36 * 4 + 2048;
  if (smem > 227 * 1024) { printf("cg=%d N=%3d sw=%3d: needs %d bytes of smem, skipped\n", cg, N, sw, smem); return; }
  cudaMemset(out_d, 0, sms * 4 * sizeof(unsigned long long));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0.f;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    if (cg == 2) {
      cudaFuncSetAttribute(mma_rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(sms); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, mma_rate_kernel<2>, a);
    } else {
      cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      mma_rate_kernel<1><<<sms, 384, smem>>>(a);
    }
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("cg=%d N=%d sw=%d: %s\n", cg, N, sw, cudaGetErrorString(err)); exit(1); }
    cudaEventElapsedTime(&ms, e0, e1);
  }
  std::vector<unsigned long long> h(sms * 4);
  cudaMemcpy(h.data(), out_d, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  double cyc = 0, n = 0, ns = 0; int cnt = 0;
  for (int i = 0; i < sms; ++i) if (h[i * 4 + 1]) { cyc += (double)h[i * 4]; n += (double)h[i * 4 + 1]; ns += (double)h[i * 4 + 2]; ++cnt; }
  const double per = cyc / n, mhz = 1e3 * cyc / ns;
  const double macs = (cg == 2 ? 256.0 : 128.0) * N * 16;                 // per instruction
  const double tf = 2.0 * macs * n / (ms * 1e-3) / 1e12;
  printf("%s cg=%d N=%3d sw=%3d copies=%d epi=%d: %6.1f cycles per MMA (%d issuing CTAs), SM clock %4.0f MHz, %7.1f TFLOP/s fp16 chip-wide (%.0f MAC/clk/SM)\n", mn ? "MN-major" : " K-major", cg, N, sw,
         copies, epi, per, cnt, mhz, tf, macs / per / (cg == 2 ? 2 : 1));
}

int main() {
  unsigned char* src; unsigned long long* out;
  cudaMalloc(&src, (size_t)160 * 65536 + (1 << 20));
  cudaMemset(src, 0x14, (size_t)160 * 65536 + (1 << 20));
  cudaMalloc(&out, 256 * 4 * sizeof(unsigned long long));
  const int Ns[] = {64, 128, 192, 240, 256};
  if (getenv("MMA_RATE_MN")) {                                      // the weight-gradient operand layout against the K-major one
    for (int N : {64, 128, 208, 256}) { run(1, N, 64, 0, 0, src, out, 1); run(1, N, 128, 0, 0, src, out, 0); }
    run(1, 208, 64, 1, 0, src, out, 1);
    return 0;
  }
  for (int cg = 1; cg <= 2; ++cg)
    for (int sw : {64, 128})
      for (int N : Ns) run(cg, N, sw, 0, 0, src, out);
  for (int cg = 1; cg <= 2; ++cg)
    for (int sw : {64, 128}) {
      run(cg, 240, sw, 1, 0, src, out);
      run(cg, 240, sw, 0, 1, src, out);
      run(cg, 240, sw, 1, 1, src, out);
    }
  return 0;
}
