// tcgen05 3xTF32 GEMM engine (placeholder until the kernel lands; SIMT is used meanwhile).
#pragma once
#include "gac_common.cuh"
namespace gac {
inline bool tc_gemm_supported(const GemmArgs&) { return false; }
inline cudaError_t tc_gemm_init() { return cudaSuccess; }
inline cudaError_t launch_gemm_tc(const GemmArgs&, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace gac
