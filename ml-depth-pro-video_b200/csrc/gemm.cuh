// GEMM / implicit-GEMM convolution cores (see common.cuh GemmOp).
#pragma once
#include "common.cuh"

namespace dp {

// fp32 CUDA-core path (parity mode): all activations / weights are float.
void gemm_simt(const GemmOp& op, cudaStream_t stream);
// bf16 tcgen05 / TMEM / TMA path: activations / weights are bf16, fp32 accumulate.
void gemm_tc(const GemmOp& op, cudaStream_t stream);
void tmap_cache_clear();

inline void gemm(int prec, const GemmOp& op, cudaStream_t stream) {
  if (prec == BF16) gemm_tc(op, stream);
  else gemm_simt(op, stream);
}

}  // namespace dp
