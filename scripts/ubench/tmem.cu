// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput on one SM, alone and beside MUFU.EX2 work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ml-depth-pro-video_b200/csrc scripts/ubench/tmem.cu -o scripts/ubench/tmem
// Each participating warp reads 32 lanes x 32 columns x 4 B = 4 KB per tcgen05.ld.32x32b.x32; warp w may only
// touch TMEM lanes 32 (w % 4) .. 32 (w % 4) + 31.  MODE 0: back-to-back loads (4 in flight before a wait);
// MODE 1: each load followed by 32 ex2 on its registers (the attention softmax pattern); MODE 2: ex2 only.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace dp;

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int iters, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    ptx::tmem_alloc(&slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;  // 2 warps per quadrant: own 128 columns
  float acc = 0.f, a4[4] = {0.f, 0.f, 0.f, 0.f};
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    uint32_t r[4][32];
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (MODE != 2) {
#pragma unroll
        for (int c = 0; c < 4; ++c) ptx::tmem_ld32(base + c * 32, r[c]);
        ptx::tmem_ld_wait();
      }
      if (MODE == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) a4[c] += __uint_as_float(r[c][0]) + __uint_as_float(r[c][31]);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < 32; ++j) a4[j & 3] += ex2(MODE == 2 ? a4[(j + 1) & 3] * 1e-9f + j : __uint_as_float(r[c][j]) * 1e-30f);
      }
    }
    t1 = clock64();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (a4[0] + a4[1]) + (a4[2] + a4[3]);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(slot, 512);
  }
}

template <int MODE>
void run(const char* name, int nwarps, int blocks) {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8 * 1024);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) { k<MODE><<<blocks, 512>>>(out, cyc, iters, nwarps); cudaDeviceSynchronize(); }
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  const double per_iter = (double)h / iters;                      // cycles for nwarps x 16 KB (128 columns x 32 lanes x 4 B)
  const double bytes = MODE == 2 ? 0 : nwarps * 16384.0;
  printf("%-26s warps %2d blocks %3d: %8.1f cycles / 128-column pass  -> %6.1f B/clk/SM TMEM read, %5.2f ex2/clk/SM  (%s)\n", name,
         nwarps, blocks, per_iter, bytes / per_iter, MODE == 0 ? 0.0 : nwarps * 128 * 32 / per_iter, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int nw : {1, 4, 8}) run<0>("tcgen05.ld only", nw, 1);
  run<0>("tcgen05.ld only", 8, 148);
  for (int nw : {4, 8}) run<2>("ex2 only", nw, 1);
  for (int nw : {4, 8}) run<1>("tcgen05.ld + ex2", nw, 1);
  return 0;
}
