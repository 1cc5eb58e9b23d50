// Micro-benchmark: issue cost of MUFU.EX2 / F2FP / FMNMX3 per warp instruction on one SM sub-partition.
#include <cstdio>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[16];
  for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 0.001f + i * 0.01f;
  float acc = 0.f;
  unsigned pk = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) v[i] = ex2(v[i]);                                  // MUFU only
      if (MODE == 1) { v[i] = ex2(fmaf(v[i], 0.18f, -0.5f)); acc += v[i]; }  // FFMA + MUFU + FADD
      if (MODE == 2) { __nv_bfloat162 h = __floats2bfloat162_rn(v[i], v[(i + 1) & 15]); pk ^= *reinterpret_cast<unsigned*>(&h); v[i] += 1.0f; }  // F2FP + FADD
      if (MODE == 3) { v[i] = ex2(fmaf(v[i], 0.18f, -0.5f)); acc += v[i]; if (i & 1) { __nv_bfloat162 h = __floats2bfloat162_rn(v[i], v[i - 1]); pk ^= *reinterpret_cast<unsigned*>(&h); } }
      if (MODE == 4) v[i] = fmaxf(fmaxf(v[i], v[(i + 1) & 15]), acc);   // FMNMX3
    }
  }
  long long t1 = clock64();
  float s = acc + __uint_as_float(pk);
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, int threads) {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 1000;
  k<MODE><<<1, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
  k<MODE><<<1, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s threads %4d: %.2f cycles per unrolled element (16 per iter) per warp-slot\n", name, threads, (double)h / (iters * 16));
}
int main() {
  for (int th : {128, 256, 512}) {
    run<0>("MUFU.EX2", th);
    run<1>("FFMA+EX2+FADD", th);
    run<2>("F2FP+FADD", th);
    run<3>("FFMA+EX2+FADD+0.5 F2FP", th);
    run<4>("FMNMX3", th);
  }
  return 0;
}
