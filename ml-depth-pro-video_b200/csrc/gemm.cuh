// GEMM / implicit-GEMM convolution cores (see common.cuh GemmOp).
#pragma once
#include "common.cuh"

namespace dp {

// fp32 CUDA-core path (parity mode): all activations / weights are float.
void gemm_simt(const GemmOp& op, cudaStream_t stream);
// bf16 tcgen05 / TMEM / TMA path: activations / weights are bf16, fp32 accumulate.
void gemm_tc(const GemmOp& op, cudaStream_t stream);
void tmap_cache_clear();
void gemm_tc_set_res_prefetch(int on);
void gemm_tc_set_sm_limit(int sms);      // experiment knob: GEMM / conv launches use at most this many SMs (0 = all)
void gemm_tc_set_l2_persist(int mb);     // A/B switch: L2 set-aside (MB) that keeps the fp32 residual stream resident; 0 = off  // A/B switch of the fp32-residual forms' L2 prefetch (gemm_tc.cu)

inline void gemm(int prec, const GemmOp& op, cudaStream_t stream) {
  double rows = op.M;
  if (op.ngroups > 1) {
    rows = 0;
    for (int i = 0; i < op.ngroups; ++i) rows += op.grp[i].M;
  }
  const double flops = 2.0 * rows * static_cast<double>(op.N) * op.K;
  if (prec == BF16) {
    ProfScope ps(stream, op.a_mode == A_CONV3X3 ? KC_CONV_TC : KC_GEMM_TC, flops);
    gemm_tc(op, stream);
  } else {
    ProfScope ps(stream, KC_GEMM_SIMT, flops);
    gemm_simt(op, stream);
  }
}

}  // namespace dp
