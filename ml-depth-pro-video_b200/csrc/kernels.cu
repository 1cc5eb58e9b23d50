// HBM-bound kernels of the Depth Pro engine (see kernels.cuh).  Reference lines are cited per
// kernel; paths are relative to the reference repo root.
#include "kernels.cuh"

#include <cmath>
#include <cstdlib>

namespace dp {

namespace {

constexpr int IMG = 1536;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return h_to_f(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float v) { return f_to_h(v); }

inline int blocks_for(long long n, int threads) { return static_cast<int>((n + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------
// resize: src/depth_pro/depth_pro.py:125-132 (ToTensor, Normalize) + :273-279 (F.interpolate,
// bilinear, align_corners=False, no antialias).
// ------------------------------------------------------------------------------------------
// One thread = one output pixel, all three channels (the index arithmetic and the bilinear weights are
// shared); one block = 256 consecutive pixels of one output row: 32-bit arithmetic only, planar stores
// coalesced across the warp.  uint8 sources go through a 256-entry table of the exact
// ToTensor + Normalize values ((u / 255 - 0.5) * 2, each step rounded like ATen does).
template <int FMT>
__global__ void __launch_bounds__(256) resize_kernel(const void* __restrict__ src, int H, int W, float* __restrict__ x) {
  __shared__ float lut[256];
  if (FMT == 1) {
    const float t = __fdiv_rn(static_cast<float>(threadIdx.x), 255.f);  // ToTensor
    lut[threadIdx.x] = __fmul_rn(__fsub_rn(t, 0.5f), 2.f);              // Normalize(0.5, 0.5)
    __syncthreads();
  }
  const int ox = blockIdx.x * 256 + threadIdx.x, oy = blockIdx.y, b = blockIdx.z;
  int y0 = oy, x0 = ox, y1 = oy, x1 = ox;
  float ly1 = 0.f, lx1 = 0.f;
  const bool same = H == IMG && W == IMG;
  if (!same) {
    const float sh = static_cast<float>(H) / IMG, sw = static_cast<float>(W) / IMG;
    // ATen's CPU kernels (x86-64 AVX2 / AVX-512 builds) contract scale*(dst+0.5)-0.5 into one FMA;
    // the single rounding matters: the product is O(1e3), so an unfused multiply moves the source
    // index by up to ~5e-5 whenever in/out is not exactly representable.
    float fy = fmaf(sh, oy + 0.5f, -0.5f);
    float fx = fmaf(sw, ox + 0.5f, -0.5f);
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    y0 = y0 > H - 1 ? H - 1 : y0;
    x0 = x0 > W - 1 ? W - 1 : x0;
    y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    ly1 = fminf(fmaxf(fy - y0, 0.f), 1.f), lx1 = fminf(fmaxf(fx - x0, 0.f), 1.f);
  }
  const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  float* dst = x + (static_cast<size_t>(b) * 3 * IMG + oy) * IMG + ox;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float p00, p01, p10, p11;
    if (FMT == 1) {
      const uint8_t* im = reinterpret_cast<const uint8_t*>(src) + static_cast<size_t>(b) * H * W * 3 + c;
      const size_t r0 = static_cast<size_t>(y0) * W, r1 = static_cast<size_t>(y1) * W;
      p00 = lut[im[(r0 + x0) * 3]], p01 = lut[im[(r0 + x1) * 3]];
      p10 = lut[im[(r1 + x0) * 3]], p11 = lut[im[(r1 + x1) * 3]];
    } else {
      const float* im = reinterpret_cast<const float*>(src) + (static_cast<size_t>(b) * 3 + c) * H * W;
      const size_t r0 = static_cast<size_t>(y0) * W, r1 = static_cast<size_t>(y1) * W;
      p00 = im[r0 + x0], p01 = im[r0 + x1], p10 = im[r1 + x0], p11 = im[r1 + x1];
    }
    float v = p00;
    if (!same) {
      const float t0 = __fadd_rn(__fmul_rn(lx0, p00), __fmul_rn(lx1, p01));
      const float t1 = __fadd_rn(__fmul_rn(lx0, p10), __fmul_rn(lx1, p11));
      v = __fadd_rn(__fmul_rn(ly0, t0), __fmul_rn(ly1, t1));
    }
    dst[static_cast<size_t>(c) * IMG * IMG] = v;
  }
}

// Second generation (default since round 2; DEPTHPRO_HBM_V2=0 restores the kernel above): one block = ONE output row
// (384 threads x 4 consecutive pixels), so the row's source lines and vertical weights are computed once per block, every
// thread keeps four independent gather chains in flight and each channel leaves as one 128-bit store.  Same arithmetic,
// bit-identical output (test_resize_v2_is_bit_identical).
template <int FMT>
__global__ void __launch_bounds__(384) resize_kernel_v2(const void* __restrict__ src, int H, int W, float* __restrict__ x) {
  __shared__ float lut[256];
  if (FMT == 1) {
    if (threadIdx.x < 256) {
      const float t = __fdiv_rn(static_cast<float>(threadIdx.x), 255.f);
      lut[threadIdx.x] = __fmul_rn(__fsub_rn(t, 0.5f), 2.f);
    }
    __syncthreads();
  }
  const int ox0 = threadIdx.x * 4, oy = blockIdx.x, b = blockIdx.y;
  const bool same = H == IMG && W == IMG;
  int y0 = oy, y1 = oy;
  float ly1 = 0.f;
  const float sh = static_cast<float>(H) / IMG, sw = static_cast<float>(W) / IMG;
  if (!same) {
    float fy = fmaf(sh, oy + 0.5f, -0.5f);
    fy = fy < 0.f ? 0.f : fy;
    y0 = static_cast<int>(fy);
    y0 = y0 > H - 1 ? H - 1 : y0;
    y1 = y0 + (y0 < H - 1 ? 1 : 0);
    ly1 = fminf(fmaxf(fy - y0, 0.f), 1.f);
  }
  const float ly0 = 1.f - ly1;
  int x0[4], x1[4];
  float lx1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ox = ox0 + i;
    x0[i] = ox, x1[i] = ox, lx1[i] = 0.f;
    if (!same) {
      float fx = fmaf(sw, ox + 0.5f, -0.5f);
      fx = fx < 0.f ? 0.f : fx;
      x0[i] = static_cast<int>(fx);
      x0[i] = x0[i] > W - 1 ? W - 1 : x0[i];
      x1[i] = x0[i] + (x0[i] < W - 1 ? 1 : 0);
      lx1[i] = fminf(fmaxf(fx - x0[i], 0.f), 1.f);
    }
  }
  float* dst = x + (static_cast<size_t>(b) * 3 * IMG + oy) * IMG + ox0;
  const size_t r0 = static_cast<size_t>(y0) * W, r1 = static_cast<size_t>(y1) * W;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float p00, p01, p10, p11;
      if (FMT == 1) {
        const uint8_t* im = reinterpret_cast<const uint8_t*>(src) + static_cast<size_t>(b) * H * W * 3 + c;
        p00 = lut[im[(r0 + x0[i]) * 3]], p01 = lut[im[(r0 + x1[i]) * 3]];
        p10 = lut[im[(r1 + x0[i]) * 3]], p11 = lut[im[(r1 + x1[i]) * 3]];
      } else {
        const float* im = reinterpret_cast<const float*>(src) + (static_cast<size_t>(b) * 3 + c) * H * W;
        p00 = im[r0 + x0[i]], p01 = im[r0 + x1[i]], p10 = im[r1 + x0[i]], p11 = im[r1 + x1[i]];
      }
      v[i] = p00;
      if (!same) {
        const float lx0 = 1.f - lx1[i];
        const float t0 = __fadd_rn(__fmul_rn(lx0, p00), __fmul_rn(lx1[i], p01));
        const float t1 = __fadd_rn(__fmul_rn(lx0, p10), __fmul_rn(lx1[i], p11));
        v[i] = __fadd_rn(__fmul_rn(ly0, t0), __fmul_rn(ly1, t1));
      }
    }
    *reinterpret_cast<float4*>(dst + static_cast<size_t>(c) * IMG * IMG) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ------------------------------------------------------------------------------------------
// interpolation_mode = "bicubic" (depth_pro.py:247, 273-279, 288-291 pass the mode to F.interpolate): ATen's
// upsample_bicubic2d with align_corners=False -- source index scale * (dst + 0.5) - 0.5 WITHOUT the clamp at 0 the
// linear modes have, floor, lambda clamped to [0, 1], cubic-convolution coefficients with A = -0.75, the four taps per
// axis clamped to the image, rows first.  ("bilinear" and "bicubic" are the only modes F.interpolate accepts together
// with align_corners=False on 4-D input; every other mode raises ValueError in the reference and in the Python shim.)
// ------------------------------------------------------------------------------------------
struct CubicTaps {
  int i[4];
  float w[4];
};
__device__ __forceinline__ CubicTaps cubic_taps(float scale, int dst, int size) {
  constexpr float A = -0.75f;
  const float real = fmaf(scale, dst + 0.5f, -0.5f);
  int idx = static_cast<int>(floorf(real));
  idx = idx > size - 1 ? size - 1 : idx;
  const float t = fminf(fmaxf(real - idx, 0.f), 1.f);
  CubicTaps c;
  const float x0 = t + 1.f, x3 = (1.f - t) + 1.f, x2 = 1.f - t;
  c.w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c.w[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
  c.w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c.w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = idx + j - 1;
    c.i[j] = k < 0 ? 0 : (k > size - 1 ? size - 1 : k);
  }
  return c;
}

template <int FMT>
__global__ void __launch_bounds__(256) resize_bicubic_kernel(const void* __restrict__ src, int H, int W, float* __restrict__ x) {
  __shared__ float lut[256];
  if (FMT == 1) {
    const float t = __fdiv_rn(static_cast<float>(threadIdx.x), 255.f);
    lut[threadIdx.x] = __fmul_rn(__fsub_rn(t, 0.5f), 2.f);
    __syncthreads();
  }
  const int ox = blockIdx.x * 256 + threadIdx.x, oy = blockIdx.y, b = blockIdx.z;
  const CubicTaps ty = cubic_taps(static_cast<float>(H) / IMG, oy, H), tx = cubic_taps(static_cast<float>(W) / IMG, ox, W);
  float* dst = x + (static_cast<size_t>(b) * 3 * IMG + oy) * IMG + ox;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float row = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float p;
        if (FMT == 1) {
          const uint8_t* im = reinterpret_cast<const uint8_t*>(src) + static_cast<size_t>(b) * H * W * 3 + c;
          p = lut[im[(static_cast<size_t>(ty.i[j]) * W + tx.i[i]) * 3]];
        } else {
          const float* im = reinterpret_cast<const float*>(src) + (static_cast<size_t>(b) * 3 + c) * H * W;
          p = im[static_cast<size_t>(ty.i[j]) * W + tx.i[i]];
        }
        row = i == 0 ? __fmul_rn(tx.w[0], p) : __fadd_rn(row, __fmul_rn(tx.w[i], p));
      }
      acc = j == 0 ? __fmul_rn(ty.w[0], row) : __fadd_rn(acc, __fmul_rn(ty.w[j], row));
    }
    dst[static_cast<size_t>(c) * IMG * IMG] = acc;
  }
}

__global__ void __launch_bounds__(256) depth_epilogue_bicubic_kernel(const float* __restrict__ canon,
                                                                     const float* __restrict__ f_px, int H, int W,
                                                                     float* __restrict__ depth) {
  const int ox = blockIdx.x * 256 + threadIdx.x, oy = blockIdx.y, b = blockIdx.z;
  if (ox >= W) return;
  const float scale = static_cast<float>(W) / f_px[b];
  const float* src = canon + static_cast<size_t>(b) * IMG * IMG;
  const CubicTaps ty = cubic_taps(static_cast<float>(IMG) / H, oy, IMG), tx = cubic_taps(static_cast<float>(IMG) / W, ox, IMG);
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float row = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float p = src[ty.i[j] * IMG + tx.i[i]] * scale;
      row = i == 0 ? __fmul_rn(tx.w[0], p) : __fadd_rn(row, __fmul_rn(tx.w[i], p));
    }
    acc = j == 0 ? __fmul_rn(ty.w[0], row) : __fadd_rn(acc, __fmul_rn(ty.w[j], row));
  }
  depth[(static_cast<size_t>(b) * H + oy) * W + ox] = 1.0f / fminf(fmaxf(acc, 1e-4f), 1e4f);
}

// ------------------------------------------------------------------------------------------
// pyramid + split + im2col: src/depth_pro/network/encoder.py:151-188, 253-263 and the timm
// patch-embed conv (k16 s16) unrolled.  Level 1 / 2 are the closed forms of
// F.interpolate(scale_factor=0.5 / 0.25, bilinear): source index 2d+0.5 / 4d+1.5, weights 0.5.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float box4(float a, float b, float c, float d) {
  // 0.5*(0.5a + 0.5b) + 0.5*(0.5c + 0.5d): scalings by 0.5 are exact, two roundings as in ATen.
  return __fadd_rn(__fmul_rn(0.5f, __fadd_rn(__fmul_rn(0.5f, a), __fmul_rn(0.5f, b))),
                   __fmul_rn(0.5f, __fadd_rn(__fmul_rn(0.5f, c), __fmul_rn(0.5f, d))));
}

// 8 consecutive outputs (col is a multiple of 8): one 16-byte (bf16) or two 16-byte (fp32) stores
__device__ __forceinline__ void store8(float* dst, const float (&v)[8]) {
  reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* dst, const float (&v)[8]) {
  uint4 pk;
  bf16x2 t;
  t = f2_to_h2(v[0], v[1]), pk.x = *reinterpret_cast<uint32_t*>(&t);
  t = f2_to_h2(v[2], v[3]), pk.y = *reinterpret_cast<uint32_t*>(&t);
  t = f2_to_h2(v[4], v[5]), pk.z = *reinterpret_cast<uint32_t*>(&t);
  t = f2_to_h2(v[6], v[7]), pk.w = *reinterpret_cast<uint32_t*>(&t);
  *reinterpret_cast<uint4*>(dst) = pk;
}

template <typename T>
__global__ void split_im2col_kernel(const float* __restrict__ x, int B, T* __restrict__ A35, T* __restrict__ A1) {
  // one thread = 8 consecutive kx of one (row, c, ky)
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(B) * 35 * 576 * 96;
  if (idx >= total) return;
  const int col8 = static_cast<int>(idx % 96);
  const long long row = idx / 96;
  const int col = col8 * 8;
  const int c = col >> 8, ky = (col >> 4) & 15, kx0 = col & 15;
  const int tok = static_cast<int>(row % 576);
  const int patch = static_cast<int>((row / 576) % 35);
  const int b = static_cast<int>(row / (576 * 35));
  const int py = (tok / 24) * 16 + ky, px = (tok % 24) * 16 + kx0;
  const float* img = x + (static_cast<long long>(b) * 3 + c) * IMG * IMG;
  float v[8];
  if (patch < 25) {
    // X is a multiple of 8: two aligned 16-byte loads
    const int Y = (patch / 5) * 288 + py, X = (patch % 5) * 288 + px;
    const float4* p = reinterpret_cast<const float4*>(img + static_cast<long long>(Y) * IMG + X);
    const float4 a = p[0], c4 = p[1];
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = c4.x, v[5] = c4.y, v[6] = c4.z, v[7] = c4.w;
  } else if (patch < 34) {
    const int q = patch - 25;
    const int Y = (q / 3) * 192 + py, X = (q % 3) * 192 + px;
    const float4* p0 = reinterpret_cast<const float4*>(img + static_cast<long long>(2 * Y) * IMG + 2 * X);
    const float4* p1 = p0 + IMG / 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 r0 = p0[i], r1 = p1[i];
      v[2 * i] = box4(r0.x, r0.y, r1.x, r1.y);
      v[2 * i + 1] = box4(r0.z, r0.w, r1.z, r1.w);
    }
  } else {
    const float* p0 = img + static_cast<long long>(4 * py + 1) * IMG + 4 * px + 1;
    const float* p1 = p0 + IMG;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = box4(p0[4 * i], p0[4 * i + 1], p1[4 * i], p1[4 * i + 1]);
  }
  store8(A35 + row * 768 + col, v);
  if (A1 != nullptr && patch == 34) store8(A1 + (static_cast<long long>(b) * 576 + tok) * 768 + col, v);
}

__global__ void im2col_to_ref_kernel(const float* __restrict__ A35, int B, float* __restrict__ patches) {
  // patches index: (n, c, py, px), n in the reference order (encoder.py:186-188, 260-263)
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(B) * 35 * 3 * 384 * 384;
  if (idx >= total) return;
  const int px = static_cast<int>(idx % 384), py = static_cast<int>((idx / 384) % 384);
  const int c = static_cast<int>((idx / (384 * 384)) % 3);
  const int n = static_cast<int>(idx / (3LL * 384 * 384));
  int patch, b;
  if (n < 25 * B) patch = n / B, b = n % B;
  else if (n < 34 * B) patch = 25 + (n - 25 * B) / B, b = (n - 25 * B) % B;
  else patch = 34, b = n - 34 * B;
  const long long row = (static_cast<long long>(b) * 35 + patch) * 576 + (py / 16) * 24 + px / 16;
  patches[idx] = A35[row * 768 + c * 256 + (py % 16) * 16 + (px % 16)];
}

// ------------------------------------------------------------------------------------------
// tokens
// ------------------------------------------------------------------------------------------
__global__ void cls_rows_kernel(float* resid, const float* cls, const float* pos, int nseq) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nseq * 1024) return;
  const int n = idx & 1023, r = idx >> 10;
  resid[static_cast<long long>(r) * 577 * 1024 + n] = cls[n] + pos[n];
}

// dest row -> source row of the token matrix (encoder.py:190-231, merge + reshape_feature)
__device__ __forceinline__ long long map_row(const RowMap& m, long long d) {
  if (m.mode == 0) return d;
  const int S = m.S;
  const int x = static_cast<int>(d % S), y = static_cast<int>((d / S) % S);
  const long long b = d / (static_cast<long long>(S) * S);
  const int first = 24 - m.pad, stride = 24 - 2 * m.pad;
  int j = 0, i = 0;
  if (m.steps > 1) {
    j = y < first ? 0 : min(m.steps - 1, 1 + (y - first) / stride);
    i = x < first ? 0 : min(m.steps - 1, 1 + (x - first) / stride);
  }
  const int ty = y - j * stride, tx = x - i * stride;
  const int patch = m.patch_base + j * m.steps + i;
  return (m.seq_off + b * m.sb + static_cast<long long>(patch) * m.sp) * 577 + 1 + ty * 24 + tx;
}

// `in` is the fp32 token matrix, or -- when `in_hi` is set -- the stream's (hi, lo) pair of 16-bit arrays (x = hi + lo,
// common.cuh GemmOp::ln_xlo).
template <typename T>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ in, T* __restrict__ out,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        long long n_out, RowMap map, int ln, LnGroups grp,
                                                        const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo) {
  const long long d = blockIdx.x * 8LL + (threadIdx.x >> 5);
  asm volatile("griddepcontrol.wait;" ::: "memory");  // PDL: `in` comes from the previous kernel
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (d >= n_out) return;
  if (grp.n > 1) {
    const int gi = (d >= grp.end[0] ? 1 : 0) + (d >= grp.end[1] ? 1 : 0);
    w = grp.w[gi], bias = grp.b[gi];
  }
  const int lane = threadIdx.x & 31;
  float4 v[8];
  if (in_hi != nullptr) {
    const long long r = map_row(map, d) * 1024;
    const uint2* sh = reinterpret_cast<const uint2*>(in_hi + r);
    const uint2* sl = reinterpret_cast<const uint2*>(in_lo + r);
    uint2 h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = sh[lane + 32 * i], l[i] = sl[lane + 32 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 h0 = h2_to_f2(*reinterpret_cast<const bf16x2*>(&h[i].x)), h1 = h2_to_f2(*reinterpret_cast<const bf16x2*>(&h[i].y));
      const float2 l0 = h2_to_f2(*reinterpret_cast<const bf16x2*>(&l[i].x)), l1 = h2_to_f2(*reinterpret_cast<const bf16x2*>(&l[i].y));
      v[i] = make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
    }
  } else {
    const float4* src = reinterpret_cast<const float4*>(in + map_row(map, d) * 1024);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = src[lane + 32 * i];
  }
  if (ln) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / 1024.f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i].x -= mean, v[i].y -= mean, v[i].z -= mean, v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.f / sqrtf(q * (1.f / 1024.f) + 1e-6f);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 ww = w4[lane + 32 * i], bb = b4[lane + 32 * i];
      v[i].x = v[i].x * rstd * ww.x + bb.x;
      v[i].y = v[i].y * rstd * ww.y + bb.y;
      v[i].z = v[i].z * rstd * ww.z + bb.z;
      v[i].w = v[i].w * rstd * ww.w + bb.w;
    }
  }
  T* dst = out + d * 1024;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int e = (lane + 32 * i) * 4;
    if (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + e) = v[i];
    } else {
      bf16x2 lo = f2_to_h2(v[i].x, v[i].y), hi = f2_to_h2(v[i].z, v[i].w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(dst) + e) = pk;
    }
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm folded into the ViT GEMMs (common.cuh GemmOp::ln_stats)
// ------------------------------------------------------------------------------------------
// Entry of the folded chain (the residual stream right after patch embed): x fp32 -> raw bf16 copy +
// per-row (sum, sum of squares) in slot 0, the other slots zero.  Later layers get both from the
// proj / fc2 epilogues.
// With `xlo` the stream continues as a (hi, lo) pair: xb = hi = round16(x), xlo = round16(x - hi).
__global__ void __launch_bounds__(256) ln_stats_cast_kernel(const float* __restrict__ in, bf16* __restrict__ xb,
                                                            bf16* __restrict__ xlo, float* __restrict__ stats,
                                                            long long rows) {
  const long long d = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (d >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* src = reinterpret_cast<const float4*>(in + d * 1024);
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = src[lane + 32 * i];
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    q += fmaf(v[i].x, v[i].x, v[i].y * v[i].y) + fmaf(v[i].z, v[i].z, v[i].w * v[i].w);
    const bf16x2 lo = f2_to_h2(v[i].x, v[i].y), hi = f2_to_h2(v[i].z, v[i].w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&lo), pk.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(xb + d * 1024 + (lane + 32 * i) * 4) = pk;
    if (xlo != nullptr) {
      const float2 r0 = h2_to_f2(lo), r1 = h2_to_f2(hi);
      const bf16x2 e0 = f2_to_h2(v[i].x - r0.x, v[i].y - r0.y), e1 = f2_to_h2(v[i].z - r1.x, v[i].w - r1.y);
      uint2 pe;
      pe.x = *reinterpret_cast<const uint32_t*>(&e0), pe.y = *reinterpret_cast<const uint32_t*>(&e1);
      *reinterpret_cast<uint2*>(xlo + d * 1024 + (lane + 32 * i) * 4) = pe;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o), q += __shfl_xor_sync(0xffffffffu, q, o);
  if (lane < 2 * LN_SLOTS) stats[d * (2 * LN_SLOTS) + lane] = lane == 0 ? s : (lane == 1 ? q : 0.f);
}

// Weight fold, one block per output feature n (K = 1024):  wf[n,k] = bf16(g[k] * w[n,k]),
// c[n] = sum_k wf[n,k] (of the ROUNDED values: it cancels the mean of what the tensor core multiplies),
// d[n] = bias[n] + sum_k b_ln[k] * w[n,k].
__global__ void __launch_bounds__(256) ln_fold_kernel(const float* __restrict__ w, const float* __restrict__ g,
                                                      const float* __restrict__ b_ln, const float* __restrict__ bias,
                                                      bf16* __restrict__ wf, float* __restrict__ c, float* __restrict__ d,
                                                      int K) {
  const int n = blockIdx.x;
  float cs = 0.f, ds = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float wv = w[static_cast<long long>(n) * K + k];
    const bf16 r = f_to_h(g[k] * wv);
    wf[static_cast<long long>(n) * K + k] = r;
    cs += h_to_f(r);
    ds = fmaf(b_ln[k], wv, ds);
  }
  __shared__ float red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o), ds += __shfl_xor_sync(0xffffffffu, ds, o);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = cs, red[1][threadIdx.x >> 5] = ds;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) a += red[0][i], b += red[1][i];
    c[n] = a;
    d[n] = bias[n] + b;
  }
}

// Test helper: y[m,n] = (xb[m,n] - mean_m) * rstd_m from the producer's outputs (fp32 out).
__global__ void ln_apply_from_stats_kernel(const bf16* __restrict__ xb, const float* __restrict__ stats,
                                           float* __restrict__ y, long long rows) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows * 1024) return;
  const long long m = idx >> 10;
  float s = 0.f, q = 0.f;
  for (int i = 0; i < LN_SLOTS; ++i) s += stats[(m * LN_SLOTS + i) * 2], q += stats[(m * LN_SLOTS + i) * 2 + 1];
  const float mean = s * (1.f / 1024.f);
  const float var = fmaxf(q * (1.f / 1024.f) - mean * mean, 0.f);
  y[idx] = (h_to_f(xb[idx]) - mean) * (1.f / sqrtf(var + 1e-6f));
}

__global__ void merge_f32_kernel(const float* __restrict__ in, float* __restrict__ out, long long total, int C,
                                 RowMap map) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  const long long d = idx / C;
  out[idx] = in[map_row(map, d) * C + c];
}

// ------------------------------------------------------------------------------------------
// FOV head convolutions: src/depth_pro/network/fov.py:29-46, 78-82
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void conv_direct_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                   T* __restrict__ y, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                                   int Ho, int Wo, int relu, const T* __restrict__ add_tokens) {
  extern __shared__ float patch[];  // k*k*Cin
  const int o = blockIdx.x;         // output pixel (b, oy, ox)
  const int ox = o % Wo, oy = (o / Wo) % Ho, b = o / (Wo * Ho);
  const int nel = k * k * Cin;
  for (int e = threadIdx.x; e < nel; e += blockDim.x) {
    const int ci = e % Cin, t = e / Cin;
    const int yy = oy * stride - pad + t / k, xx = ox * stride - pad + t % k;
    patch[e] = (yy >= 0 && yy < H && xx >= 0 && xx < W)
                   ? to_f(x[((static_cast<long long>(b) * H + yy) * W + xx) * Cin + ci])
                   : 0.f;
  }
  __syncthreads();
  for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
    float acc = 0.f;
    for (int e = 0; e < nel; ++e) acc = fmaf(patch[e], w[static_cast<long long>(e) * Cout + co], acc);
    acc += bias[co];
    if (relu) acc = fmaxf(acc, 0.f);
    if (add_tokens) acc += to_f(add_tokens[static_cast<long long>(o) * Cout + co]);
    y[static_cast<long long>(o) * Cout + co] = from_f<T>(acc);
  }
}

template <typename T>
__global__ void im2col_kernel(const T* __restrict__ x, T* __restrict__ cols, int B, int H, int W, int C, int k,
                              int stride, int pad, int Ho, int Wo) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int K = k * k * C;
  const long long total = static_cast<long long>(B) * Ho * Wo * K;
  if (idx >= total) return;
  const int kk = static_cast<int>(idx % K);
  const long long m = idx / K;
  const int c = kk % C, tap = kk / C;
  const int ox = static_cast<int>(m % Wo), oy = static_cast<int>((m / Wo) % Ho);
  const long long b = m / (static_cast<long long>(Wo) * Ho);
  const int yy = oy * stride - pad + tap / k, xx = ox * stride - pad + tap % k;
  T v = from_f<T>(0.f);
  if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x[((b * H + yy) * W + xx) * C + c];
  cols[idx] = v;
}

template <typename T>
__global__ void fov_final_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                 float* __restrict__ fov) {
  __shared__ float red[8];
  const int b = blockIdx.x;
  float acc = 0.f;
  for (int e = threadIdx.x; e < 1152; e += blockDim.x) acc = fmaf(to_f(x[b * 1152 + e]), w[e], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) s += red[i];
    fov[b] = s + bias[0];
  }
}

// ------------------------------------------------------------------------------------------
// metric depth epilogue: src/depth_pro/depth_pro.py:282-298
// ------------------------------------------------------------------------------------------
__global__ void fpx_kernel(const float* fov_deg, const float* f_px_in, int W, float* f_px, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (f_px_in) {
    f_px[b] = f_px_in[b];
  } else {
    const float rad = fov_deg[b] * static_cast<float>(3.14159265358979323846 / 180.0);
    f_px[b] = (0.5f * W) / tanf(0.5f * rad);
  }
}

// one thread = one output pixel, one block = 256 consecutive pixels of one output row (32-bit arithmetic)
__global__ void __launch_bounds__(256) depth_epilogue_kernel(const float* __restrict__ canon, const float* __restrict__ f_px,
                                                             int H, int W, float* __restrict__ depth) {
  const int ox = blockIdx.x * 256 + threadIdx.x, oy = blockIdx.y, b = blockIdx.z;
  if (ox >= W) return;
  const float scale = static_cast<float>(W) / f_px[b];
  const float* src = canon + static_cast<size_t>(b) * IMG * IMG;
  float inv;
  if (H == IMG && W == IMG) {
    inv = src[oy * IMG + ox] * scale;
  } else {
    const float sh = static_cast<float>(IMG) / H, sw = static_cast<float>(IMG) / W;
    float fy = fmaf(sh, oy + 0.5f, -0.5f);  // see resize_kernel
    float fx = fmaf(sw, ox + 0.5f, -0.5f);
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = min(static_cast<int>(fy), IMG - 1), x0 = min(static_cast<int>(fx), IMG - 1);
    const int y1 = y0 + (y0 < IMG - 1 ? 1 : 0), x1 = x0 + (x0 < IMG - 1 ? 1 : 0);
    const float ly1 = fminf(fmaxf(fy - y0, 0.f), 1.f), lx1 = fminf(fmaxf(fx - x0, 0.f), 1.f);
    const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float p00 = src[y0 * IMG + x0] * scale, p01 = src[y0 * IMG + x1] * scale;
    const float p10 = src[y1 * IMG + x0] * scale, p11 = src[y1 * IMG + x1] * scale;
    const float t0 = __fadd_rn(__fmul_rn(lx0, p00), __fmul_rn(lx1, p01));
    const float t1 = __fadd_rn(__fmul_rn(lx0, p10), __fmul_rn(lx1, p11));
    inv = __fadd_rn(__fmul_rn(ly0, t0), __fmul_rn(ly1, t1));
  }
  depth[(static_cast<size_t>(b) * H + oy) * W + ox] = 1.0f / fminf(fmaxf(inv, 1e-4f), 1e4f);
}

// v2 (opt-in: DEPTHPRO_HBM_V2=1 or hbm_v2_set(); not the default until it has been measured): one thread = FOUR
// consecutive output pixels with exactly the per-pixel arithmetic above and one 128-bit store.  The v1 form launches
// H * ceil(W / 256) blocks of one dependent gather chain per thread (4K: 32 400 blocks = 27 waves of ~1.5 us, 43 us for
// 42 MB = 0.15 of the copy bandwidth); this one has a quarter of the blocks and four chains in flight per thread.
__global__ void __launch_bounds__(256) depth_epilogue_kernel_v2(const float* __restrict__ canon, const float* __restrict__ f_px,
                                                                int H, int W, float* __restrict__ depth) {
  const int ox0 = (blockIdx.x * 256 + threadIdx.x) * 4, oy = blockIdx.y, b = blockIdx.z;
  if (ox0 >= W) return;
  const float scale = static_cast<float>(W) / f_px[b];
  const float* src = canon + static_cast<size_t>(b) * IMG * IMG;
  const bool same = H == IMG && W == IMG;
  const float sh = static_cast<float>(IMG) / H, sw = static_cast<float>(IMG) / W;
  float fy = fmaf(sh, oy + 0.5f, -0.5f);
  fy = fy < 0.f ? 0.f : fy;
  const int y0 = min(static_cast<int>(fy), IMG - 1);
  const int y1 = y0 + (y0 < IMG - 1 ? 1 : 0);
  const float ly1 = fminf(fmaxf(fy - y0, 0.f), 1.f);
  const float ly0 = 1.f - ly1;
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ox = ox0 + i;
    float inv = 0.f;
    if (ox < W) {
      if (same) {
        inv = src[oy * IMG + ox] * scale;
      } else {
        float fx = fmaf(sw, ox + 0.5f, -0.5f);
        fx = fx < 0.f ? 0.f : fx;
        const int x0 = min(static_cast<int>(fx), IMG - 1);
        const int x1 = x0 + (x0 < IMG - 1 ? 1 : 0);
        const float lx1 = fminf(fmaxf(fx - x0, 0.f), 1.f);
        const float lx0 = 1.f - lx1;
        const float p00 = src[y0 * IMG + x0] * scale, p01 = src[y0 * IMG + x1] * scale;
        const float p10 = src[y1 * IMG + x0] * scale, p11 = src[y1 * IMG + x1] * scale;
        const float t0 = __fadd_rn(__fmul_rn(lx0, p00), __fmul_rn(lx1, p01));
        const float t1 = __fadd_rn(__fmul_rn(lx0, p10), __fmul_rn(lx1, p11));
        inv = __fadd_rn(__fmul_rn(ly0, t0), __fmul_rn(ly1, t1));
      }
    }
    v[i] = 1.0f / fminf(fmaxf(inv, 1e-4f), 1e4f);
  }
  float* d = depth + (static_cast<size_t>(b) * H + oy) * W + ox0;
  if (ox0 + 3 < W && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
    *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (ox0 + i < W) d[i] = v[i];
  }
}

// ------------------------------------------------------------------------------------------
// depth -> 3D unprojection with row-major stream compaction:
// img_to_normalized_pointcloud.py:819-856 (+ colours :1226)
// ------------------------------------------------------------------------------------------
constexpr int UNP_T = 256, UNP_PER = 4, UNP_BLK = UNP_T * UNP_PER;

__device__ __forceinline__ bool depth_valid(float d) { return !isnan(d) && d > 0.f; }

__global__ void unproject_count_kernel(const float* __restrict__ depth, long long n, int* __restrict__ counts) {
  const long long base = static_cast<long long>(blockIdx.x) * UNP_BLK;
  int c = 0;
#pragma unroll
  for (int it = 0; it < UNP_PER; ++it) {
    const long long p = base + it * UNP_T + threadIdx.x;
    c += (p < n && depth_valid(depth[p])) ? 1 : 0;
  }
  __shared__ int red[UNP_T / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int i = 0; i < UNP_T / 32; ++i) s += red[i];
    counts[blockIdx.x] = s;
  }
}

// exclusive scan of counts (in place -> offsets are written to `offsets`), single block
__global__ void unproject_scan_kernel(const int* __restrict__ counts, long long* __restrict__ offsets, int nblk,
                                      int64_t* __restrict__ n_valid) {
  __shared__ long long warp_tot[32];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + threadIdx.x;
    long long v = i < nblk ? counts[i] : 0;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      long long w = warp_tot[threadIdx.x];
      long long wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, wi, o);
        if (threadIdx.x >= o) wi += t;
      }
      warp_tot[threadIdx.x] = wi - w;  // exclusive warp offsets
    }
    __syncthreads();
    const long long excl = carry + warp_tot[threadIdx.x >> 5] + incl - v;
    if (i < nblk) offsets[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_valid = carry;
}

__global__ void unproject_write_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ rgb, int H, int W,
                                       const float* __restrict__ f_px, const long long* __restrict__ offsets,
                                       float* __restrict__ xyz, float* __restrict__ rgb_out,
                                       uint8_t* __restrict__ valid_mask) {
  __shared__ int warp_cnt[UNP_T / 32];
  __shared__ int running;
  const long long n = static_cast<long long>(H) * W;
  const long long base = static_cast<long long>(blockIdx.x) * UNP_BLK;
  const float f = f_px[0];
  const float cx = 0.5f * W, cy = 0.5f * H;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  const long long blk_off = offsets[blockIdx.x];
  for (int it = 0; it < UNP_PER; ++it) {
    const long long p = base + it * UNP_T + threadIdx.x;
    float d = 0.f;
    bool ok = false;
    if (p < n) {
      d = depth[p];
      ok = depth_valid(d);
      if (valid_mask) valid_mask[p] = ok ? 1 : 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int before = running;
    for (int i = 0; i < wid; ++i) before += warp_cnt[i];
    if (ok) {
      const long long dst = blk_off + before + __popc(bal & ((1u << lane) - 1));
      const unsigned pu = static_cast<unsigned>(p);  // n < 2^31 (checked on the host): 32-bit division
      const int v = static_cast<int>(pu / static_cast<unsigned>(W)), u = static_cast<int>(pu - static_cast<unsigned>(v) * W);
      // x = -(u - W/2) * z / f ; y = -(v - H/2) * z / f ; z = d
      xyz[dst * 3 + 0] = __fdiv_rn(__fmul_rn(-(static_cast<float>(u) - cx), d), f);
      xyz[dst * 3 + 1] = __fdiv_rn(__fmul_rn(-(static_cast<float>(v) - cy), d), f);
      xyz[dst * 3 + 2] = d;
      if (rgb && rgb_out) {
#pragma unroll
        for (int c = 0; c < 3; ++c) rgb_out[dst * 3 + c] = __fdiv_rn(static_cast<float>(rgb[p * 3 + c]), 255.f);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
      for (int i = 0; i < UNP_T / 32; ++i) s += warp_cnt[i];
      running += s;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// colourise / 16-bit export: generate_depth_maps.py:15-44, 128-143
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
// Pass 1 of 2: every block leaves ITS nan-min / nan-max (order-preserving keys) in mm[2 * block], no atomics and no
// initialisation launch; pass 2 (colorize_kernel) reduces the MINMAX_BLOCKS pairs in its prologue.
constexpr int MINMAX_BLOCKS = 592;  // 148 SMs x 4
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ d, long long n, unsigned* mm) {
  __shared__ unsigned slo[8], shi[8];
  unsigned lo = 0xffffffffu, hi = 0u;
  auto take = [&](float v) {
    if (!isnan(v)) {
      const unsigned o = f2ord(v);
      lo = min(lo, o), hi = max(hi, o);
    }
  };
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n4 = (reinterpret_cast<uintptr_t>(d) % 16 == 0) ? n / 4 : 0;  // 16-byte loads when aligned
  for (long long i = tid; i < n4; i += nth) {
    const float4 v = reinterpret_cast<const float4*>(d)[i];
    take(v.x), take(v.y), take(v.z), take(v.w);
  }
  for (long long i = n4 * 4 + tid; i < n; i += nth) take(d[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) slo[threadIdx.x >> 5] = lo, shi[threadIdx.x >> 5] = hi;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) lo = min(lo, slo[w]), hi = max(hi, shi[w]);
    mm[2 * blockIdx.x] = lo, mm[2 * blockIdx.x + 1] = hi;
  }
}
// one thread = 4 consecutive pixels: one 16-byte load, 12 bytes (three 32-bit words) or four uint16 out
// `lo_fix` / `hi_fix`: caller-supplied min_depth / max_depth (generate_depth_maps.py:15-31), NaN = take the image's own.
__global__ void __launch_bounds__(256) colorize_kernel(const float* __restrict__ d, long long n,
                                                       const unsigned* __restrict__ mm, const uint8_t* __restrict__ lut,
                                                       void* __restrict__ out, int vec, float lo_fix, float hi_fix) {
  __shared__ uint8_t slut[768];
  __shared__ unsigned slo[8], shi[8];
  __shared__ float s_lo, s_hi;
  if (lut)
    for (int i = threadIdx.x; i < 768; i += 256) slut[i] = lut[i];
  {
    unsigned lo = 0xffffffffu, hi = 0u;
    if (mm)
      for (int i = threadIdx.x; i < MINMAX_BLOCKS; i += 256) lo = min(lo, mm[2 * i]), hi = max(hi, mm[2 * i + 1]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) slo[threadIdx.x >> 5] = lo, shi[threadIdx.x >> 5] = hi;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) lo = min(lo, slo[w]), hi = max(hi, shi[w]);
      s_lo = isnan(lo_fix) ? ord2f(lo) : lo_fix;
      s_hi = isnan(hi_fix) ? ord2f(hi) : hi_fix;
    }
    __syncthreads();
  }
  const long long i0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4;
  if (i0 >= n) return;
  const float lo = s_lo, hi = s_hi;
  const float span = __fsub_rn(hi, lo);
  const int cnt = n - i0 < 4 ? static_cast<int>(n - i0) : 4;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (vec && cnt == 4) {
    const float4 t = *reinterpret_cast<const float4*>(d + i0);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  } else {
    for (int j = 0; j < cnt; ++j) v[j] = d[i0 + j];
  }
  if (lut) {
    uint8_t px[12];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float norm = __fdiv_rn(__fsub_rn(v[j], lo), span);
      int k = static_cast<int>(fminf(fmaxf(norm, 0.f), 1.f) * 256.f);
      k = k > 255 ? 255 : k;
      const bool bad = isnan(norm);
      px[3 * j] = bad ? 0 : slut[k * 3], px[3 * j + 1] = bad ? 0 : slut[k * 3 + 1], px[3 * j + 2] = bad ? 0 : slut[k * 3 + 2];
    }
    uint8_t* o = reinterpret_cast<uint8_t*>(out) + i0 * 3;
    if (vec && cnt == 4) {
      uint32_t w[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) w[q] = px[4 * q] | (px[4 * q + 1] << 8) | (px[4 * q + 2] << 16) | (px[4 * q + 3] << 24);
      uint32_t* o32 = reinterpret_cast<uint32_t*>(o);
      o32[0] = w[0], o32[1] = w[1], o32[2] = w[2];
    } else {
      for (int j = 0; j < 3 * cnt; ++j) o[j] = px[j];
    }
  } else {
    uint16_t* o = reinterpret_cast<uint16_t*>(out) + i0;
    for (int j = 0; j < cnt; ++j) {
      const float norm = __fdiv_rn(__fsub_rn(v[j], lo), span);
      o[j] = static_cast<uint16_t>(static_cast<int>(norm * 65535.f));
    }
  }
}

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = from_f<TO>(to_f(in[i]));
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int B, int H, int W, int C) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;  // over NCHW out
  const long long total = static_cast<long long>(B) * C * H * W;
  if (i >= total) return;
  const int x = static_cast<int>(i % W), y = static_cast<int>((i / W) % H);
  const int c = static_cast<int>((i / (static_cast<long long>(W) * H)) % C);
  const long long b = i / (static_cast<long long>(W) * H * C);
  out[i] = to_f(in[((b * H + y) * W + x) * C + c]);
}
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int C, int H, int W) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;  // over NHWC out
  const long long total = static_cast<long long>(B) * C * H * W;
  if (i >= total) return;
  const int c = static_cast<int>(i % C), x = static_cast<int>((i / C) % W);
  const int y = static_cast<int>((i / (static_cast<long long>(C) * W)) % H);
  const long long b = i / (static_cast<long long>(C) * W * H);
  out[i] = in[((b * C + c) * H + y) * W + x];
}
template <typename T>
__global__ void pack_ohwi_kernel(const float* __restrict__ w, T* __restrict__ out, int O, int I, int KH, int KW) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;  // over (o, ky, kx, ci)
  const long long total = static_cast<long long>(O) * I * KH * KW;
  if (i >= total) return;
  const int ci = static_cast<int>(i % I), kx = static_cast<int>((i / I) % KW);
  const int ky = static_cast<int>((i / (static_cast<long long>(I) * KW)) % KH);
  const long long o = i / (static_cast<long long>(I) * KW * KH);
  out[i] = from_f<T>(w[((o * I + ci) * KH + ky) * KW + kx]);
}
template <typename T>
__global__ void pack_convT_kernel(const float* __restrict__ w, T* __restrict__ out, int I, int O) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;  // over ((dy,dx,o), ci)
  const long long total = 4LL * I * O;
  if (i >= total) return;
  const int ci = static_cast<int>(i % I);
  const int o = static_cast<int>((i / I) % O);
  const int q = static_cast<int>(i / (static_cast<long long>(I) * O));
  out[i] = from_f<T>(w[((static_cast<long long>(ci) * O + o) * 2 + (q >> 1)) * 2 + (q & 1)]);
}
__global__ void pack_hwio_kernel(const float* __restrict__ w, float* __restrict__ out, int O, int I, int KH, int KW) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;  // over (ky, kx, ci, o)
  const long long total = static_cast<long long>(O) * I * KH * KW;
  if (i >= total) return;
  const int o = static_cast<int>(i % O), ci = static_cast<int>((i / O) % I);
  const int kx = static_cast<int>((i / (static_cast<long long>(O) * I)) % KW);
  const int ky = static_cast<int>(i / (static_cast<long long>(O) * I * KW));
  out[i] = w[((static_cast<long long>(o) * I + ci) * KH + ky) * KW + kx];
}

__global__ void compose_head_w_kernel(const float* __restrict__ w1, const float* __restrict__ w2, bf16* __restrict__ wc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over [128][1152]
  if (idx >= 128 * 1152) return;
  const int k = idx % 1152, n = idx / 1152;
  const int ci = k % 128, tap = k / 128, ty = tap / 3, tx = tap % 3;
  const int c2 = n % 32, par = n / 32, py = par >> 1, px = par & 1;
  float acc = 0.f;
  for (int ky = 0; ky < 3; ++ky) {
    const int uy = py + ky + 1;  // (py + ky - 1) + 2
    if ((uy >> 1) != ty) continue;
    const int dy = uy & 1;
    for (int kx = 0; kx < 3; ++kx) {
      const int ux = px + kx + 1;
      if ((ux >> 1) != tx) continue;
      const int dx = ux & 1;
      for (int c1 = 0; c1 < 128; ++c1)
        acc = fmaf(w1[((ci * 128 + c1) * 2 + dy) * 2 + dx], w2[((c2 * 128 + c1) * 3 + ky) * 3 + kx], acc);
    }
  }
  wc[idx] = f_to_h(acc);
}
__global__ void compose_head_b_kernel(const float* __restrict__ b1, const float* __restrict__ w2,
                                      const float* __restrict__ b2, float* __restrict__ cb) {
  const int c2 = threadIdx.x;  // 32 threads
  float full = b2[c2];
  for (int tap = 0; tap < 9; ++tap) {
    float acc = 0.f;
    for (int c1 = 0; c1 < 128; ++c1) acc = fmaf(b1[c1], w2[(c2 * 128 + c1) * 9 + tap], acc);
    cb[tap * 32 + c2] = acc;
    full += acc;
  }
  cb[9 * 32 + c2] = full;
}

__global__ void compose_deconv_kernel(const float* __restrict__ wd, const float* __restrict__ wo, bf16* __restrict__ wc,
                                      int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over [4*C][C]
  if (idx >= 4 * C * C) return;
  const int ci = idx % C, n = idx / C, cop = n % C, q = n / C, dy = q >> 1, dx = q & 1;
  float acc = 0.f;
  for (int co = 0; co < C; ++co) acc = fmaf(wo[cop * C + co], wd[((ci * C + co) * 2 + dy) * 2 + dx], acc);
  wc[idx] = f_to_h(acc);
}

// 1x1 conv (C -> C, weights wo [c][i], bias bo) followed by a 3x3 conv (C -> O, weights w3 OIHW [o][c][ky][kx], bias b3,
// zero padding) is ONE 3x3 conv:  wc[o][tap][i] = sum_c w3[o][c][tap] wo[c][i]  (O(HW)I, the implicit-GEMM order).
// The 1x1's bias reaches the output through every tap that lies INSIDE the image: cb[tap][o] = sum_c w3[o][c][tap] bo[c],
// cb[9][o] = b3[o] + all nine (interior pixels); border pixels subtract the taps that fall into the padding.
__global__ void compose_1x1_3x3_w_kernel(const float* __restrict__ wo, const float* __restrict__ w3, bf16* __restrict__ wc,
                                         int O, int C) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;  // over [O][9][C]
  if (idx >= static_cast<long long>(O) * 9 * C) return;
  const int i = static_cast<int>(idx % C), tap = static_cast<int>((idx / C) % 9), o = static_cast<int>(idx / (9LL * C));
  float acc = 0.f;
  for (int c = 0; c < C; ++c) acc = fmaf(w3[(static_cast<long long>(o) * C + c) * 9 + tap], wo[static_cast<long long>(c) * C + i], acc);
  wc[idx] = f_to_h(acc);
}
__global__ void compose_1x1_3x3_b_kernel(const float* __restrict__ bo, const float* __restrict__ w3,
                                         const float* __restrict__ b3, float* __restrict__ cb, int O, int C) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= O) return;
  float full = b3[o];
  for (int tap = 0; tap < 9; ++tap) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(w3[(static_cast<long long>(o) * C + c) * 9 + tap], bo[c], acc);
    cb[tap * O + o] = acc;
    full += acc;
  }
  cb[9 * O + o] = full;
}

}  // namespace

// ============================================================================ host wrappers
void compose_1x1_conv3x3(const float* wo, const float* bo, const float* w3, const float* b3, bf16* wc, float* cb, int O,
                         int C, cudaStream_t s) {
  compose_1x1_3x3_w_kernel<<<blocks_for(static_cast<long long>(O) * 9 * C, 256), 256, 0, s>>>(wo, w3, wc, O, C);
  DP_LAUNCH_CHECK();
  compose_1x1_3x3_b_kernel<<<blocks_for(O, 128), 128, 0, s>>>(bo, w3, b3, cb, O, C);
  DP_LAUNCH_CHECK();
}

void compose_deconv_1x1(const float* wd, const float* wo, bf16* wc, int C, cudaStream_t s) {
  compose_deconv_kernel<<<blocks_for(4LL * C * C, 256), 256, 0, s>>>(wd, wo, wc, C);
  DP_LAUNCH_CHECK();
}

void compose_head(const float* w1, const float* b1, const float* w2, const float* b2, bf16* wc, float* cb,
                  cudaStream_t s) {
  compose_head_w_kernel<<<blocks_for(128 * 1152, 128), 128, 0, s>>>(w1, w2, wc);
  DP_LAUNCH_CHECK();
  compose_head_b_kernel<<<1, 32, 0, s>>>(b1, w2, b2, cb);
  DP_LAUNCH_CHECK();
}

static std::atomic<int> g_hbm_v2{-1};
void hbm_v2_set(int on) {
  g_hbm_v2 = on != 0;
  bump_config_epoch();
}
static bool hbm_v2() {
  if (g_hbm_v2 < 0) {
    const char* e = getenv("DEPTHPRO_HBM_V2");
    g_hbm_v2 = e ? (atoi(e) != 0) : 1;  // default since round 2: bit-identical to v1 on B200 (test_depth_epilogue_v2_is_bit_identical)
  }
  return g_hbm_v2 != 0;
}

void resize_to_1536(const void* src, int src_fmt, int B, int H, int W, float* x, int interp, cudaStream_t s) {
  static_assert(IMG % 256 == 0, "one block = 256 pixels of a row");
  const dim3 grid(IMG / 256, IMG, B);
  if (interp == INTERP_BICUBIC && !(H == IMG && W == IMG)) {
    if (src_fmt == 1) resize_bicubic_kernel<1><<<grid, 256, 0, s>>>(src, H, W, x);
    else resize_bicubic_kernel<0><<<grid, 256, 0, s>>>(src, H, W, x);
  } else if (hbm_v2()) {
    static_assert(IMG == 384 * 4, "resize_kernel_v2: one block = one 1536-pixel row");
    if (src_fmt == 1) resize_kernel_v2<1><<<dim3(IMG, B), 384, 0, s>>>(src, H, W, x);
    else resize_kernel_v2<0><<<dim3(IMG, B), 384, 0, s>>>(src, H, W, x);
  } else {
    if (src_fmt == 1) resize_kernel<1><<<grid, 256, 0, s>>>(src, H, W, x);
    else resize_kernel<0><<<grid, 256, 0, s>>>(src, H, W, x);
  }
  DP_LAUNCH_CHECK();
}

template <typename T>
void split_im2col(const float* x, int B, T* A35, T* A1, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * 35 * 576 * 96;
  DP_CHECK(reinterpret_cast<uintptr_t>(x) % 16 == 0, "split: the frame buffer must be 16-byte aligned");
  split_im2col_kernel<T><<<blocks_for(total, 256), 256, 0, s>>>(x, B, A35, A1);
  DP_LAUNCH_CHECK();
}
template void split_im2col<float>(const float*, int, float*, float*, cudaStream_t);
template void split_im2col<bf16>(const float*, int, bf16*, bf16*, cudaStream_t);

void im2col_to_ref_patches(const float* A35, int B, float* patches, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * 35 * 3 * 384 * 384;
  im2col_to_ref_kernel<<<blocks_for(total, 256), 256, 0, s>>>(A35, B, patches);
  DP_LAUNCH_CHECK();
}

void write_cls_rows(float* resid, const float* cls, const float* pos, int nseq, cudaStream_t s) {
  cls_rows_kernel<<<blocks_for(static_cast<long long>(nseq) * 1024, 256), 256, 0, s>>>(resid, cls, pos, nseq);
  DP_LAUNCH_CHECK();
}

template <typename T>
void layernorm_rows(const float* in, T* out, const float* w, const float* b, long long n_out, RowMap map, int ln,
                    cudaStream_t s, const bf16* in_hi, const bf16* in_lo) {
  launch_pdl(layernorm_kernel<T>, dim3(blocks_for(n_out, 8)), dim3(256), 0, s, in, out, w, b, n_out, map, ln, LnGroups(),
             in_hi, in_lo);
  DP_LAUNCH_CHECK();
}
template <typename T>
void layernorm_rows_grouped(const float* in, T* out, const LnGroups& g, long long n_out, cudaStream_t s) {
  launch_pdl(layernorm_kernel<T>, dim3(blocks_for(n_out, 8)), dim3(256), 0, s, in, out, g.w[0], g.b[0], n_out, RowMap(), 1, g,
             static_cast<const bf16*>(nullptr), static_cast<const bf16*>(nullptr));
  DP_LAUNCH_CHECK();
}
template void layernorm_rows_grouped<float>(const float*, float*, const LnGroups&, long long, cudaStream_t);
template void layernorm_rows_grouped<bf16>(const float*, bf16*, const LnGroups&, long long, cudaStream_t);
template void layernorm_rows<float>(const float*, float*, const float*, const float*, long long, RowMap, int, cudaStream_t,
                                    const bf16*, const bf16*);
template void layernorm_rows<bf16>(const float*, bf16*, const float*, const float*, long long, RowMap, int, cudaStream_t,
                                   const bf16*, const bf16*);

// micro-benchmark operands: bf16 values uniform in [-1, 1) from a counter hash (an fp32 view of the same
// bytes is a float in that range too).  Zero-filled operands hide most of the tensor-core power.
__global__ void fill_random_bf16_kernel(uint16_t* __restrict__ p, long long n, unsigned seed) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  unsigned h = static_cast<unsigned>(i) * 2654435761u ^ seed;
  h ^= h >> 16, h *= 0x7feb352du, h ^= h >> 15, h *= 0x846ca68bu, h ^= h >> 16;
  const float v = static_cast<float>(h >> 8) * (2.f / 16777216.f) - 1.f;
  const bf16 b = f_to_h(v);
  p[i] = *reinterpret_cast<const uint16_t*>(&b);
}
void fill_random_bf16(void* p, size_t bytes, unsigned seed, cudaStream_t s) {
  const long long n = static_cast<long long>(bytes / 2);
  fill_random_bf16_kernel<<<blocks_for(n, 256), 256, 0, s>>>(reinterpret_cast<uint16_t*>(p), n, seed);
  DP_LAUNCH_CHECK();
}

void ln_stats_cast(const float* in, bf16* xb, float* stats, long long rows, cudaStream_t s, bf16* xlo) {
  ln_stats_cast_kernel<<<blocks_for(rows, 8), 256, 0, s>>>(in, xb, xlo, stats, rows);
  DP_LAUNCH_CHECK();
}
__global__ void pair_to_f32_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo, float* __restrict__ x, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) x[i] = h_to_f(hi[i]) + h_to_f(lo[i]);
}
void pair_to_f32(const bf16* hi, const bf16* lo, float* x, long long n, cudaStream_t s) {
  pair_to_f32_kernel<<<blocks_for(n, 256), 256, 0, s>>>(hi, lo, x, n);
  DP_LAUNCH_CHECK();
}
void ln_fold(const float* w, const float* g, const float* b_ln, const float* bias, bf16* wf, float* c, float* d, int N,
             int K, cudaStream_t s) {
  ln_fold_kernel<<<N, 256, 0, s>>>(w, g, b_ln, bias, wf, c, d, K);
  DP_LAUNCH_CHECK();
}
void ln_apply_from_stats(const bf16* xb, const float* stats, float* y, long long rows, cudaStream_t s) {
  ln_apply_from_stats_kernel<<<blocks_for(rows * 1024, 256), 256, 0, s>>>(xb, stats, y, rows);
  DP_LAUNCH_CHECK();
}

void merge_rows_f32(const float* in, float* out, int B, int C, RowMap map, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * map.S * map.S * C;
  merge_f32_kernel<<<blocks_for(total, 256), 256, 0, s>>>(in, out, total, C, map);
  DP_LAUNCH_CHECK();
}

template <typename T>
void conv_direct(const T* x, const float* w_hwio, const float* bias, T* y, int B, int H, int W, int Cin, int Cout,
                 int k, int stride, int pad, int relu, const T* add_tokens, cudaStream_t s) {
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const size_t smem = static_cast<size_t>(k) * k * Cin * sizeof(float);
  const int threads = Cout < 32 ? 32 : (Cout > 256 ? 256 : Cout);
  conv_direct_kernel<T><<<B * Ho * Wo, threads, smem, s>>>(x, w_hwio, bias, y, H, W, Cin, Cout, k, stride, pad, Ho, Wo,
                                                          relu, add_tokens);
  DP_LAUNCH_CHECK();
}
template void conv_direct<float>(const float*, const float*, const float*, float*, int, int, int, int, int, int, int,
                                 int, int, const float*, cudaStream_t);
template void conv_direct<bf16>(const bf16*, const float*, const float*, bf16*, int, int, int, int, int, int, int, int,
                                int, const bf16*, cudaStream_t);

template <typename T>
void im2col_nhwc(const T* x, T* cols, int B, int H, int W, int C, int k, int stride, int pad, cudaStream_t s) {
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * k * k * C;
  im2col_kernel<T><<<blocks_for(total, 256), 256, 0, s>>>(x, cols, B, H, W, C, k, stride, pad, Ho, Wo);
  DP_LAUNCH_CHECK();
}
template void im2col_nhwc<float>(const float*, float*, int, int, int, int, int, int, int, cudaStream_t);
template void im2col_nhwc<bf16>(const bf16*, bf16*, int, int, int, int, int, int, int, cudaStream_t);

template <typename T>
void fov_final(const T* x, const float* w_hwio, const float* bias, float* fov_deg, int B, cudaStream_t s) {
  fov_final_kernel<T><<<B, 256, 0, s>>>(x, w_hwio, bias, fov_deg);
  DP_LAUNCH_CHECK();
}
template void fov_final<float>(const float*, const float*, const float*, float*, int, cudaStream_t);
template void fov_final<bf16>(const bf16*, const float*, const float*, float*, int, cudaStream_t);

void compute_fpx(const float* fov_deg, const float* f_px_in, int W, float* f_px, int B, cudaStream_t s) {
  fpx_kernel<<<1, 64, 0, s>>>(fov_deg, f_px_in, W, f_px, B);
  DP_LAUNCH_CHECK();
}

void depth_epilogue(const float* canon, const float* f_px, int B, int H, int W, float* depth, int interp, cudaStream_t s) {
  DP_CHECK(H <= 65535 && B <= 65535, "depth epilogue: image too tall");
  if (interp == INTERP_BICUBIC && !(H == IMG && W == IMG))
    depth_epilogue_bicubic_kernel<<<dim3((W + 255) / 256, H, B), 256, 0, s>>>(canon, f_px, H, W, depth);
  else if (hbm_v2()) depth_epilogue_kernel_v2<<<dim3((W + 1023) / 1024, H, B), 256, 0, s>>>(canon, f_px, H, W, depth);
  else depth_epilogue_kernel<<<dim3((W + 255) / 256, H, B), 256, 0, s>>>(canon, f_px, H, W, depth);
  DP_LAUNCH_CHECK();
}

size_t unproject_scratch_ints(int H, int W) {
  const long long nblk = (static_cast<long long>(H) * W + UNP_BLK - 1) / UNP_BLK;
  return static_cast<size_t>(nblk) * 3 + 8;  // int counts + int64 offsets
}

void unproject(const float* depth, const uint8_t* rgb, int H, int W, const float* f_px, float* xyz, float* rgb_out,
               uint8_t* valid_mask, int64_t* n_valid, int* scratch, cudaStream_t s) {
  const long long n = static_cast<long long>(H) * W;
  DP_CHECK(n < (1LL << 31), "unproject: image too large");
  const int nblk = static_cast<int>((n + UNP_BLK - 1) / UNP_BLK);
  int* counts = scratch;
  long long* offsets = reinterpret_cast<long long*>(scratch + ((nblk + 1) & ~1));
  unproject_count_kernel<<<nblk, UNP_T, 0, s>>>(depth, n, counts);
  DP_LAUNCH_CHECK();
  unproject_scan_kernel<<<1, 1024, 0, s>>>(counts, offsets, nblk, n_valid);
  DP_LAUNCH_CHECK();
  unproject_write_kernel<<<nblk, UNP_T, 0, s>>>(depth, rgb, H, W, f_px, offsets, xyz, rgb_out, valid_mask);
  DP_LAUNCH_CHECK();
}

size_t colorize_scratch_bytes() { return MINMAX_BLOCKS * 2 * sizeof(unsigned); }
void colorize(const float* depth, int H, int W, const uint8_t* lut, void* out, float* minmax, float min_depth,
              float max_depth, cudaStream_t s) {
  const long long n = static_cast<long long>(H) * W;
  unsigned* mm = reinterpret_cast<unsigned*>(minmax);
  const bool need_scan = std::isnan(min_depth) || std::isnan(max_depth);   // both given: no reduction pass at all
  if (need_scan) {
    minmax_kernel<<<MINMAX_BLOCKS, 256, 0, s>>>(depth, n, mm);
    DP_LAUNCH_CHECK();
  }
  // 4 pixels per thread; the packed 32-bit stores need 16-byte aligned buffers (4 pixels = 12 output bytes)
  const int vec = reinterpret_cast<uintptr_t>(depth) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 4 == 0;
  colorize_kernel<<<blocks_for((n + 3) / 4, 256), 256, 0, s>>>(depth, n, need_scan ? mm : nullptr, lut, out, vec, min_depth,
                                                                max_depth);
  DP_LAUNCH_CHECK();
}

template <typename TI, typename TO>
void convert(const TI* in, TO* out, long long n, cudaStream_t s) {
  convert_kernel<TI, TO><<<blocks_for(n, 256), 256, 0, s>>>(in, out, n);
  DP_LAUNCH_CHECK();
}
template void convert<float, float>(const float*, float*, long long, cudaStream_t);
template void convert<float, bf16>(const float*, bf16*, long long, cudaStream_t);
template void convert<bf16, float>(const bf16*, float*, long long, cudaStream_t);

template <typename T>
void nhwc_to_nchw_f32(const T* in, float* out, int B, int H, int W, int C, cudaStream_t s) {
  nhwc_to_nchw_kernel<T><<<blocks_for(static_cast<long long>(B) * H * W * C, 256), 256, 0, s>>>(in, out, B, H, W, C);
  DP_LAUNCH_CHECK();
}
template void nhwc_to_nchw_f32<float>(const float*, float*, int, int, int, int, cudaStream_t);
template void nhwc_to_nchw_f32<bf16>(const bf16*, float*, int, int, int, int, cudaStream_t);

void nchw_to_nhwc_f32(const float* in, float* out, int B, int C, int H, int W, cudaStream_t s) {
  nchw_to_nhwc_kernel<<<blocks_for(static_cast<long long>(B) * H * W * C, 256), 256, 0, s>>>(in, out, B, C, H, W);
  DP_LAUNCH_CHECK();
}

template <typename T>
void pack_oihw_to_ohwi(const float* w, T* out, int O, int I, int KH, int KW, cudaStream_t s) {
  pack_ohwi_kernel<T><<<blocks_for(static_cast<long long>(O) * I * KH * KW, 256), 256, 0, s>>>(w, out, O, I, KH, KW);
  DP_LAUNCH_CHECK();
}
template void pack_oihw_to_ohwi<float>(const float*, float*, int, int, int, int, cudaStream_t);
template void pack_oihw_to_ohwi<bf16>(const float*, bf16*, int, int, int, int, cudaStream_t);

template <typename T>
void pack_convT_iohw(const float* w, T* out, int I, int O, cudaStream_t s) {
  pack_convT_kernel<T><<<blocks_for(4LL * I * O, 256), 256, 0, s>>>(w, out, I, O);
  DP_LAUNCH_CHECK();
}
template void pack_convT_iohw<float>(const float*, float*, int, int, cudaStream_t);
template void pack_convT_iohw<bf16>(const float*, bf16*, int, int, cudaStream_t);

void pack_oihw_to_hwio_f32(const float* w, float* out, int O, int I, int KH, int KW, cudaStream_t s) {
  pack_hwio_kernel<<<blocks_for(static_cast<long long>(O) * I * KH * KW, 256), 256, 0, s>>>(w, out, O, I, KH, KW);
  DP_LAUNCH_CHECK();
}

}  // namespace dp
