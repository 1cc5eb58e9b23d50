// Attention cores (see attention.cu).
#pragma once
#include "common.cuh"

namespace dp {
// qkv (nseq*577, 3072) -> out (nseq*577, 1024)
void attention_f32(const float* qkv, float* out, int nseq, cudaStream_t s);
void attention_bf16(const bf16* qkv, bf16* out, int nseq, cudaStream_t s);      // mma.sync (legacy tensor path)
void attention_bf16_tc(const bf16* qkv, bf16* out, int nseq, cudaStream_t s);   // tcgen05 + TMEM + TMA
}  // namespace dp
