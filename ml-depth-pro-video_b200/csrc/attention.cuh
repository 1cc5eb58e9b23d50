// Attention cores (see attention.cu).
#pragma once
#include "common.cuh"

namespace dp {
// qkv (nseq*577, 3072) -> out (nseq*577, 1024)
void attention_f32(const float* qkv, float* out, int nseq, cudaStream_t s);
void attention_bf16(const bf16* qkv, bf16* out, int nseq, cudaStream_t s);      // mma.sync (legacy tensor path)
// tcgen05 + TMEM + TMA; reverse = 1 walks the (sequence, head) units from the last to the first (GemmOp::reverse)
void attention_bf16_tc(const bf16* qkv, bf16* out, int nseq, cudaStream_t s, int reverse = 0);
// Experiment switch of the tcgen05 kernel (process-wide; overrides DEPTHPRO_ATTN_EXP / DEPTHPRO_ATTN_PINGPONG):
// expv = kernel variant (list in attention_tc.cu), -1 = back to the default; pingpong 0/1.
void attention_tc_set_variant(int expv, int pingpong);
void attention_tc_set_sm_limit(int sms);   // experiment knob: at most this many CTAs (one per SM); 0 = all
// Debug counter: softmax warps that took the lazy-maximum rescale branch since the last reset (synchronises the device).
unsigned long long attention_tc_rescale_count(bool reset);
}  // namespace dp
