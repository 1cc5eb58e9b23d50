// Phase timing of the tcgen05 flash-attention kernel's softmax warps (clock64 instrumentation compiled in with
// -DATTN_PROFILE).  Build + run: scripts/ubench/build_attn_prof.sh (links the engine's other objects).
#include <cstdio>
#include <vector>
#include "common.cuh"
#include "attention.cuh"
#include "kernels.cuh"
namespace dp { void attn_prof_read(unsigned long long* host10, bool reset); }
int main() {
  const int nseq = 37, SEQ = 577;
  const size_t nq = (size_t)nseq * SEQ * 3072, no = (size_t)nseq * SEQ * 1024;
  dp::bf16 *qkv, *out;
  cudaMalloc(&qkv, nq * 2); cudaMalloc(&out, no * 2);
  dp::fill_random_bf16(qkv, nq * 2, 12345u, nullptr);
  for (int i = 0; i < 3; ++i) dp::attention_bf16_tc(qkv, out, nseq, nullptr);
  static unsigned long long p[23 + 16 * 40];
  dp::attn_prof_read(p, true);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 20;
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) dp::attention_bf16_tc(qkv, out, nseq, nullptr);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  dp::attn_prof_read(p, false);
  const char* names[8] = {"wait S (s_full)", "TMEM load", "row maximum", "wait previous P V", "lazy max / rescale", "wait MUFU turn",
                          "exponentials + P stores", "hand-off (fences, arrive)"};
  const double blocks = (double)p[8];
  printf("attention 37 sequences: %.1f us per launch (instrumented); %.0f softmax blocks sampled\n", ms * 1e3 / iters, blocks);
  double tot = 0;
  for (int i = 0; i < 8; ++i) tot += p[i] / blocks;
  for (int i = 0; i < 8; ++i) printf("  %-28s %8.1f cycles per block  (%4.1f %%)\n", names[i], p[i] / blocks, 100.0 * p[i] / blocks / tot);
  printf("  %-28s %8.1f cycles per block, amortised (per query tile: %.0f)\n", "O read-out + store", p[9] / blocks, p[9] / blocks * 5);
  printf("  total per block %.1f cycles (+ read-out)\n", tot);
  printf("  wait S by key block j (cycles per tile):");
  for (int j = 0; j < 5; ++j) printf("  j=%d %.0f", j, p[10 + j] / (blocks / 5));
  printf("\n");
  const char* mn[7] = {"q_full", "k_full", "s_empty", "QK issue", "v_full + o_empty", "p_full", "PV issue"};
  const double mb = (double)p[22];
  printf("MMA warp, cycles per block (%.0f blocks):\n", mb);
  for (int i = 0; i < 7; ++i) printf("  %-20s %8.1f\n", mn[i], p[15 + i] / mb);
  // event trace of CTA 0 / stream 0 (last launch): cycles relative to the stream's first event
  const char* ev[15] = {"prod: K req", "prod: V req", "mma: K ready", "mma: s_empty ok", "mma: QK issued", "mma: V ready",
                        "mma: p_full ok", "mma: PV issued", "smx: at s_full", "smx: S ready", "smx: S loaded", "smx: PV(G-1) done",
                        "smx: MUFU turn", "smx: P published", "smx: at pv_done wait"};
  const long long* tr = reinterpret_cast<const long long*>(p + 23);
  long long t0 = tr[0 * 40 + 0];
  printf("event trace, CTA 0 stream 0, cycles since the first K request (columns = key block G):\n%-20s", "event");
  for (int G = 10; G < 22; ++G) printf("%8d", G);
  printf("\n");
  for (int e = 0; e < 15; ++e) {
    printf("%-20s", ev[e]);
    for (int G = 10; G < 22; ++G) printf("%8lld", tr[e * 40 + G] - t0);
    printf("\n");
  }
  return 0;
}
