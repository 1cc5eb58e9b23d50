// HBM access-pattern micro-benchmark for the proj / fc2 epilogue.
//
// The fp32-residual epilogue of gemm_tc_kernel updates x (M x 1024 fp32, row-major) in place and writes a bf16 copy:
// per 128 x 256 tile, warp (q, grp) walks four 32-column chunks and moves each as 128-byte row segments at a 4 KB
// row stride (bf16: 64-byte segments at 2 KB).  ncu: 218 MB per launch at 3.4 TB/s with DRAM 42 % busy.  Is that the
// access pattern?  This kernel reproduces ONLY the memory side of that epilogue (same tile walk, same lane mapping,
// residual values two chunks ahead in registers, no MMA) and compares layouts:
//   A  fp32 row-major RMW + bf16 row-major copy          (what the engine does today)
//   B  fp32 in 32x32 blocks (4 KB contiguous per chunk) + bf16 row-major copy
//   C  fp32 blocked + bf16 blocked (2 KB contiguous)     (upper bound; the next GEMM's TMA wants row-major bf16)
//   D  fp32 row-major RMW only                           (EPI_RES32 without the LayerNorm outputs)
//   E  fp32 blocked RMW only
// L2 is flushed before every timed launch.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rmw rmw.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));       \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int N = 1024, BM = 128, BN = 256;

// element offset of the float4 this lane moves for (tile row r, tile column c0 + 4 cq) of chunk (mt, nt, q, grp, c)
template <bool BLOCKED>
__device__ __forceinline__ long long off32(int M, int mt, int nt, int q, int grp, int c, int i, int lane) {
  const int rg = lane >> 3, cq = lane & 7;
  const int row = mt * BM + q * 32 + rg + 4 * i, col = nt * BN + grp * 128 + c * 32 + cq * 4;
  if (!BLOCKED) return static_cast<long long>(row) * N + col;
  // [M/32][N/32] blocks of 32 x 32: the warp's chunk is one contiguous 4 KB block, lane-linear inside
  const long long blk = static_cast<long long>(row >> 5) * (N / 32) + (col >> 5);
  return blk * 1024 + (row & 31) * 32 + (col & 31);
}

template <bool BLK32, bool BF16, bool BLK16>
__global__ void __launch_bounds__(256, 1) rmw_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ xb, int M) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, grp = warp >> 2;
  const int m_tiles = (M + BM - 1) / BM, tiles = m_tiles * (N / BN);
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int nt = t % (N / BN), mt = t / (N / BN);
    const int rows_left = M - (mt * BM + q * 32);
    float4 v[3][8];
    auto load = [&](int c, float4 (&r)[8]) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        r[i] = ((lane >> 3) + 4 * i) < rows_left ? *reinterpret_cast<const float4*>(x + off32<BLK32>(M, mt, nt, q, grp, c, i, lane))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    load(0, v[0]);
    load(1, v[1]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c + 2 < 4) load(c + 2, v[(c + 2) % 3]);
      float4(&r)[8] = v[c % 3];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (((lane >> 3) + 4 * i) >= rows_left) continue;
        float4 o = r[i];
        o.x = fmaf(o.x, 1.0001f, 0.5f), o.y = fmaf(o.y, 1.0001f, 0.5f), o.z = fmaf(o.z, 1.0001f, 0.5f), o.w = fmaf(o.w, 1.0001f, 0.5f);
        const long long o32 = off32<BLK32>(M, mt, nt, q, grp, c, i, lane);
        *reinterpret_cast<float4*>(x + o32) = o;
        if (BF16) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
          uint2 pk;
          pk.x = *reinterpret_cast<const unsigned*>(&lo), pk.y = *reinterpret_cast<const unsigned*>(&hi);
          *reinterpret_cast<uint2*>(xb + off32<BLK16>(M, mt, nt, q, grp, c, i, lane)) = pk;
        }
      }
    }
  }
}

// ---- the residual stream as a (hi, lo) pair of bf16 arrays (EPI_RES16_LN): both read and written in place ----------
//   W32: the epilogue's 32-column chunks, a lane moves 8 columns (16 B) of rows rg + 8 i  -> 64-byte row segments
//   W64: 64-column chunks, a lane moves 8 columns (16 B) of rows rg + 4 i                 -> 128-byte row segments
// DEPTH = chunks requested ahead of the one being processed.
template <int W, int DEPTH>
__global__ void __launch_bounds__(256, 1) pair_kernel(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int M) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, grp = warp >> 2;
  const int m_tiles = (M + BM - 1) / BM, tiles = m_tiles * (N / BN);
  constexpr int NCH = 128 / W;                       // chunks per warp-tile
  constexpr int LPR = W / 8;                         // lanes per row segment
  constexpr int RPI = 32 / LPR;                      // rows per instruction
  constexpr int NI = 32 / RPI;                       // instructions per chunk and array
  constexpr int NBUF = DEPTH + 1;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int nt = t % (N / BN), mt = t / (N / BN);
    const int rows_left = M - (mt * BM + q * 32);
    const int rg = lane / LPR, cq = lane % LPR;
    auto off = [&](int c, int i) {
      return static_cast<long long>(mt * BM + q * 32 + rg + RPI * i) * N + nt * BN + grp * 128 + c * W + cq * 8;
    };
    uint4 h[NBUF][NI], l[NBUF][NI];
    auto load = [&](int c, uint4 (&hh)[NI], uint4 (&ll)[NI]) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        if (rg + RPI * i < rows_left) {
          hh[i] = *reinterpret_cast<const uint4*>(hi + off(c, i));
          ll[i] = *reinterpret_cast<const uint4*>(lo + off(c, i));
        } else {
          hh[i] = ll[i] = make_uint4(0, 0, 0, 0);
        }
      }
    };
#pragma unroll
    for (int c = 0; c < DEPTH && c < NCH; ++c) load(c, h[c % NBUF], l[c % NBUF]);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (c + DEPTH < NCH) load(c + DEPTH, h[(c + DEPTH) % NBUF], l[(c + DEPTH) % NBUF]);
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        if (rg + RPI * i >= rows_left) continue;
        uint4 a = h[c % NBUF][i], b = l[c % NBUF][i];
        a.x += b.x, a.y ^= b.y, a.z += b.w, a.w ^= b.z;   // any dependence on both loads
        b.x ^= a.y, b.y += a.x, b.z ^= a.w, b.w += a.z;
        *reinterpret_cast<uint4*>(hi + off(c, i)) = a;
        *reinterpret_cast<uint4*>(lo + off(c, i)) = b;
      }
    }
  }
}
template <int W, int DEPTH>
static void run_pair(const char* name, __nv_bfloat16* hi, __nv_bfloat16* lo, int M, void* flush, size_t flush_bytes, int sms) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float tot = 0.f, best = 1e9f;
  const int iters = 12;
  for (int i = 0; i < iters + 2; ++i) {
    CK(cudaMemsetAsync(flush, i & 0xff, flush_bytes));
    CK(cudaEventRecord(a));
    pair_kernel<W, DEPTH><<<sms, 256>>>(hi, lo, M);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (i >= 2) tot += ms, best = ms < best ? ms : best;
  }
  CK(cudaGetLastError());
  const double mp = (M + 31) / 32 * 32;
  const double bytes = mp * N * 8.0;
  printf("%-58s %8.1f us mean %8.1f us best  %6.2f TB/s (mean)\n", name, tot / iters * 1e3, best * 1e3, bytes / (tot / iters * 1e-3) / 1e12);
}

template <bool BLK32, bool BF16, bool BLK16>
static void run(const char* name, float* x, __nv_bfloat16* xb, int M, void* flush, size_t flush_bytes, int sms) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float tot = 0.f, best = 1e9f;
  const int iters = 12;
  for (int i = 0; i < iters + 2; ++i) {
    CK(cudaMemsetAsync(flush, i & 0xff, flush_bytes));
    CK(cudaEventRecord(a));
    rmw_kernel<BLK32, BF16, BLK16><<<sms, 256>>>(x, xb, M);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (i >= 2) tot += ms, best = ms < best ? ms : best;
  }
  CK(cudaGetLastError());
  const double mp = (M + 31) / 32 * 32;
  const double bytes = mp * N * (8.0 + (BF16 ? 2.0 : 0.0));
  printf("%-58s %8.1f us mean %8.1f us best  %6.2f TB/s (mean)\n", name, tot / iters * 1e3, best * 1e3, bytes / (tot / iters * 1e-3) / 1e12);
}

int main() {
  const int M = 37 * 577;  // rows of the frame's residual stream
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t mp = (M + 127) / 128 * 128;
  float* x;
  __nv_bfloat16* xb;
  void* flush;
  const size_t flush_bytes = 256u << 20;
  CK(cudaMalloc(&x, mp * N * 4));
  CK(cudaMalloc(&xb, mp * N * 2));
  CK(cudaMalloc(&flush, flush_bytes));
  CK(cudaMemset(x, 0, mp * N * 4));
  CK(cudaMemset(xb, 0, mp * N * 2));
  printf("fp32 residual RMW (+ bf16 copy) of %d x %d, epilogue tile walk, %d CTAs x 8 warps, L2 flushed per launch\n", M, N, sms);
  run<false, true, false>("A fp32 row-major + bf16 row-major (today)", x, xb, M, flush, flush_bytes, sms);
  run<true, true, false>("B fp32 32x32-blocked + bf16 row-major", x, xb, M, flush, flush_bytes, sms);
  run<true, true, true>("C fp32 blocked + bf16 blocked", x, xb, M, flush, flush_bytes, sms);
  run<false, false, false>("D fp32 row-major only", x, xb, M, flush, flush_bytes, sms);
  run<true, false, false>("E fp32 blocked only", x, xb, M, flush, flush_bytes, sms);
  run<false, true, false>("A again", x, xb, M, flush, flush_bytes, sms);
  // pair stream: hi in xb, lo in the first half of x
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(x);
  printf("(hi, lo) pair stream, 8 B per element (%.0f MB)\n", (double)mp * N * 8 / 1e6);
  run_pair<32, 2>("P1 32-col chunks (64 B segments), 2 ahead (engine today)", xb, lo, M, flush, flush_bytes, sms);
  run_pair<32, 3>("P1b 32-col chunks, 3 ahead (whole warp-tile in flight)", xb, lo, M, flush, flush_bytes, sms);
  run_pair<64, 1>("P2 64-col chunks (128 B segments), 1 ahead", xb, lo, M, flush, flush_bytes, sms);
  run_pair<64, 2>("P3 64-col chunks, whole warp-tile in flight", xb, lo, M, flush, flush_bytes, sms);
  run_pair<128, 1>("P4 128-col chunks (256 B segments), whole tile in flight", xb, lo, M, flush, flush_bytes, sms);
  run_pair<32, 2>("P1 again", xb, lo, M, flush, flush_bytes, sms);
  return 0;
}
