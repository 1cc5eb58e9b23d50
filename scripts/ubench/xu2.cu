// Micro-benchmark (round 2): does the packed half-precision MUFU.EX2 deliver two results per issue slot on B200?
// Cycles per warp instruction with 1 / 2 / 4 warps per SM sub-partition (128 / 256 / 512 threads, one CTA, one SM).
//   0 ex2.approx.ftz.f32            (reference: 8 cycles per warp instruction)
//   1 ex2.approx.f16x2              (two exponentials per instruction)
//   2 ex2.approx.ftz.bf16x2
//   3 the proposed softmax chain per PAIR of scores: fma.rn.f32x2 -> cvt.rn.f16x2.f32 -> ex2.approx.f16x2 -> add.f16x2
//   4 today's chain per PAIR: fma.rn.f32x2 -> 2 x ex2.f32 -> add.rn.f32x2 -> cvt.rn.bf16x2.f32
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2h2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned ex2b2(unsigned x) { unsigned y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned cvt_h2(unsigned long long v) {
  float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  unsigned r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ unsigned hadd2(unsigned a, unsigned b) { unsigned r; asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[16]; unsigned u[16]; unsigned long long w[16];
  for (int i = 0; i < 16; ++i) { v[i] = threadIdx.x * 0.001f + i * 0.01f; u[i] = 0x34003400u + threadIdx.x + i; w[i] = (unsigned long long)__float_as_uint(v[i]) << 32 | __float_as_uint(v[i] * 0.5f); }
  unsigned hacc = 0; unsigned long long facc = 0; unsigned pk = 0;
  const unsigned long long c2 = (unsigned long long)__float_as_uint(0.18f) << 32 | __float_as_uint(0.18f);
  const unsigned long long m2 = (unsigned long long)__float_as_uint(-0.5f) << 32 | __float_as_uint(-0.5f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) v[i] = ex2f(v[i]);
      if (MODE == 1) u[i] = ex2h2(u[i]);
      if (MODE == 2) u[i] = ex2b2(u[i]);
      if (MODE == 3) { const unsigned p = ex2h2(cvt_h2(fma2(w[i], c2, m2))); hacc = hadd2(hacc, p); pk ^= p; w[i] += 0x0000100000001000ull; }
      if (MODE == 4) {
        const unsigned long long x = fma2(w[i], c2, m2);
        float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x));
        a = ex2f(a); b = ex2f(b);
        unsigned long long p; asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a), "f"(b));
        facc = add2(facc, p);
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b); pk ^= *reinterpret_cast<unsigned*>(&h);
        w[i] += 0x0000100000001000ull;
      }
    }
  }
  long long t1 = clock64();
  float s = __uint_as_float(pk) + __uint_as_float(hacc) + __uint_as_float((unsigned)facc) + __uint_as_float((unsigned)(facc >> 32));
  for (int i = 0; i < 16; ++i) s += v[i] + __uint_as_float(u[i]) + __uint_as_float((unsigned)w[i]);
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
  printf("%-44s threads %4d: %.2f cycles per unrolled slot per warp\n", name, threads, (double)h / (iters * 16));
}
int main() {
  for (int th : {128, 256, 512}) {
    run<0>("MUFU.EX2 f32 (1 exp / instr)", th);
    run<1>("MUFU.EX2 f16x2 (2 exp / instr)", th);
    run<2>("MUFU.EX2 bf16x2 (2 exp / instr)", th);
    run<3>("chain/pair: FFMA2+F2FP+EX2.f16x2+HADD2", th);
    run<4>("chain/pair: FFMA2+2 EX2.f32+FADD2+F2FP", th);
  }
  // accuracy of the f16x2 path against exp2 in double, over the softmax argument range [-16, 0]
  return 0;
}
