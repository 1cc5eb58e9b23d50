// bf16 GEMM / implicit-GEMM 3x3 convolution on the 5th-generation tensor cores (sm_100a).
//
//   D[M,N] = A[M,K] * W[N,K]^T   (fp32 accumulate in TMEM), fused epilogue (common.cuh GemmOp).
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer  : cp.async.bulk.tensor (128B swizzle) A and W k-blocks into a
//                               4-stage shared-memory ring, completion on mbarriers.
//                               A is either a 2D row-major matrix or — for 3x3 convolutions —
//                               a 4D NHWC tensor map: the 128-row tile is an 8x16 pixel block
//                               and every filter tap is one shifted box load whose
//                               out-of-bounds pixels TMA zero-fills (= the padding).
//   warp 1      MMA issuer    : one thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per k-block
//                               into a double-buffered TMEM accumulator; tcgen05.commit releases
//                               smem stages and publishes finished accumulators.
//   warp 2      TMEM allocator
//   warps 4-11  epilogue      : tcgen05.ld (thread = one output row, 32 columns per load),
//                               bias / GELU / ReLU / LayerScale / residual / pixel-shuffle
//                               scatter, vectorised global stores; overlaps the next tile's MMA.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include <mutex>

#include "gemm.cuh"
#include "ptx.cuh"

namespace dp {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
// smem ring: as many k-block stages as fit beside the 32 KB epilogue staging (at most 8).  A CTA pair
// (CL = 2, tcgen05 cta_group::2) keeps only HALF of the weight tile per CTA: 32 KB stages, 6 deep at
// BN = 256 instead of 4 x 48 KB -- the ring covers 2560 tensor-core cycles of load latency, not 1536.
// BN = 128 tiles put TWO 64-wide k-blocks into one stage (KSUB): an MMA of the narrow tile lasts 32 tensor-core cycles, a
// k-block 128 -- less than one full / empty barrier round trip of the issuing warp costs (ncu on head.0, N = 128: tensor
// pipe 57 % with the tensor core's shared-memory read pipe at 43 %: not a bandwidth limit).  Two k-blocks per hand-shake
// give the same 256 cycles per round trip as the BN = 256 tiles.
template <int BN, int CL, int EPI>
struct Cfg {
  static constexpr int B_ROWS = BN / CL;                      // weight rows held by one CTA
#ifndef DP_KSUB128
#define DP_KSUB128 2  // (-DDP_KSUB128=1 rebuilds the one-k-block-per-stage form for A/B runs)
#endif
  static constexpr int KSUB = BN == 128 ? DP_KSUB128 : 1;     // 64-wide k-blocks per ring stage
  static constexpr int STAGE = KSUB * (128 * 64 * 2 + B_ROWS * 64 * 2);
  static constexpr int STG_WARP = EPI == 3 ? 8192 : 4096;     // EPI_TMA2 stages x and relu(x) side by side
  static constexpr int FIT = (232448 - 1024 - 256 - 8 * STG_WARP) / STAGE;
  static constexpr int STAGES = FIT > 8 ? 8 : FIT;
};
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int TILE_W = 16, TILE_H = 8;  // conv: 128 rows = 8 x 16 output pixels

// m-tiles are scheduled in UNITS of CL tiles.  CL = 2 is a CTA pair on one TPC: the two CTAs hold two
// adjacent m-tiles and one half of the weight tile each, and the leader CTA issues ONE
// tcgen05.mma.cta_group::2 (M = 256) that reads both CTAs' shared memory and writes both CTAs' TMEM.
// A unit never straddles a group: every group is padded to a whole number of units.
struct TileGeom {
  int m_units, n_tiles, k_blocks;
  int tiles_x, tiles_y;  // conv only
  int unit_start[4];     // first unit of every group, unit_start[ngroups..3] = m_units
  int tiles_in_group[3]; // real m-tiles per group (tiles beyond are padding: loaded, not stored)
  int tma_out;           // 1: bf16 row-major output goes smem -> TMA store (full-line writes)
  int res_prefetch;      // 1: fp32-residual forms: the producer warp pulls each tile's residual rows into L2
  int reverse;           // 1: walk the tiles from the last m-unit to the first (GemmOp::reverse)
};
struct WeightMaps {
  CUtensorMap b[3];      // one weight tensor map per group
  CUtensorMap o[3];      // output tensor maps (TMA-store epilogue), one per group; o[0] for convs / ConvT
  CUtensorMap orelu;     // same geometry as o[0] over the ReLU'd twin (dual-store convs)
};

__device__ __forceinline__ int unit_group(const TileGeom& g, int mu) {
  return (mu >= g.unit_start[1] ? 1 : 0) + (mu >= g.unit_start[2] ? 1 : 0);
}

// GELU (exact-erf form) for the bf16 path: erf by Abramowitz-Stegun 7.1.28,
//   erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16,  |err| <= 3e-7 (far below bf16 output rounding),
// one MUFU rcp per element instead of the ~30-instruction erff() (7.1.26 needs rcp AND ex2).
// The fc1 epilogue is instruction-issue bound (ncu: 87 M warp instructions vs 23 M for the same GEMM
// without GELU), so two elements are processed per instruction with the packed fp32x2 FMA pipe.
struct F2 {
  unsigned long long u;
};
__device__ __forceinline__ F2 pack_f2(float a, float b) {
  F2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.u) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack_f2(F2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v.u)); }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
  F2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.u) : "l"(a.u), "l"(b.u), "l"(c.u));
  return r;
}
__device__ __forceinline__ F2 mul2(F2 a, F2 b) {
  F2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u));
  return r;
}
__device__ __forceinline__ F2 splat2(float a) { return pack_f2(a, a); }
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// gelu(x) = 0.5 x (1 + erf(x / sqrt 2)),  erf(z) = sign(z) (1 - s^-16),  s = 1 + a1 |z| + ... + a6 |z|^6
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const F2 x = pack_f2(x0, x1);
  F2 z;
  z.u = x.u & 0x7fffffff7fffffffull;                       // |x|
  z = mul2(z, splat2(0.70710678118654752440f));
  F2 s = fma2(z, splat2(0.0000430638f), splat2(0.0002765672f));
  s = fma2(s, z, splat2(0.0001520143f));
  s = fma2(s, z, splat2(0.0092705272f));
  s = fma2(s, z, splat2(0.0422820123f));
  s = fma2(s, z, splat2(0.0705230784f));
  s = fma2(s, z, splat2(1.0f));
  float s0, s1;
  unpack_f2(s, s0, s1);
  F2 r = pack_f2(rcp_approx(s0), rcp_approx(s1));          // s >= 1: r in (0, 1], r^16 underflows to 0 for large |x|
  r = mul2(r, r);
  r = mul2(r, r);
  r = mul2(r, r);
  r = mul2(r, r);
  F2 u = fma2(r, splat2(-1.0f), splat2(1.0f));             // |erf|
  u.u |= x.u & 0x8000000080000000ull;                      // copysign(|erf|, x)
  const F2 h = mul2(x, splat2(0.5f));
  const F2 g = fma2(h, u, h);
  unpack_f2(g, x0, x1);
}

template <typename T>
__device__ __forceinline__ void load32(const T* p, float (&v)[32]);
template <>
__device__ __forceinline__ void load32<float>(const float* p, float (&v)[32]) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 t = q[i];
    v[4 * i] = t.x, v[4 * i + 1] = t.y, v[4 * i + 2] = t.z, v[4 * i + 3] = t.w;
  }
}
template <>
__device__ __forceinline__ void load32<bf16>(const bf16* p, float (&v)[32]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 t = q[i];
    const bf16x2* h = reinterpret_cast<const bf16x2*>(&t);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = h2_to_f2(h[j]);
      v[8 * i + 2 * j] = f.x, v[8 * i + 2 * j + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void store32(float* p, const float (&v)[32]) {
  float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void store32(bf16* p, const float (&v)[32], bool relu) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 t;
    bf16x2* h = reinterpret_cast<bf16x2*>(&t);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = v[8 * i + 2 * j], b = v[8 * i + 2 * j + 1];
      if (relu) a = fmaxf(a, 0.f), b = fmaxf(b, 0.f);
      h[j] = f2_to_h2(a, b);
    }
    q[i] = t;
  }
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  bf16x2 h = f2_to_h2(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Epilogue for one output row `m` (linear row / NHWC pixel index) and 32 columns starting at n0.
// `m_local` is the row inside its group (patch-embed token placement), `resv` the residual values
// for these 32 columns, already loaded (and converted) by the caller so that the HBM round trip
// overlaps the accumulator wait.
__device__ __forceinline__ void epilogue_chunk(const GemmOp& op, const GemmGroup& gp, long long m, int n0,
                                               float (&v)[32], const float (&resv)[32]) {
  float t[32];
  if (gp.bias) {
    load32<float>(gp.bias + (op.bias_mod ? n0 % op.bias_mod : n0), t);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += t[j];
  }
  if (op.act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (op.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) gelu_erf2(v[j], v[j + 1]);
  }
  if (gp.gamma) {
    load32<float>(gp.gamma + n0, t);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= t[j];
  }

  long long off;  // element offset of (m, n0) in `out`
  if (op.out_mode == O_ROWMAJOR) {
    off = m * op.ldo + op.col_off + n0;
  } else if (op.out_mode == O_CONVT2X2) {
    const int q = n0 / op.cout, co = n0 - q * op.cout;
    const int x = static_cast<int>(m % op.W);
    const long long by = m / op.W;  // b*H + y
    const int y = static_cast<int>(by % op.H);
    const long long b = by / op.H;
    const long long orow = (b * 2 * op.H + 2 * y + (q >> 1)) * (2LL * op.W) + 2 * x + (q & 1);
    off = orow * op.ldo + op.col_off + co;
  } else if (op.out_mode == O_PATCH_EMBED) {
    const long long seq = m / 576;
    const int p = static_cast<int>(m - seq * 576);
    load32<float>(gp.pos + (1 + p) * static_cast<long long>(op.N) + n0, t);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += t[j];
    off = (seq * 577 + 1 + p) * op.ldo + n0;
  } else if (op.out_mode == O_HEAD_FUSED) {
    // chunk = one output parity (py, px) of coarse pixel m = (y, x); v = composed conv result.
    const int par = n0 >> 5, py = par >> 1, px = par & 1;
    const int x = static_cast<int>(m % op.W), y = static_cast<int>(m / op.W);
    const int Y = 2 * y + py, X = 2 * x + px, HH = 2 * op.H, WW = 2 * op.W;
    load32<float>(op.head_cb + 9 * 32, t);  // b2 + sum of all 9 tap terms (interior pixels)
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += t[j];
    if (Y == 0 || Y == HH - 1 || X == 0 || X == WW - 1) {
      // zero padding of the fine-resolution conv: taps that fall outside carry no head.1 bias
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = Y + ky - 1, xx = X + kx - 1;
          if (yy < 0 || yy >= HH || xx < 0 || xx >= WW) {
            load32<float>(op.head_cb + (ky * 3 + kx) * 32, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] -= t[j];
          }
        }
    }
    load32<float>(op.dot_w, t);
    float s = op.dot_b[0];
#pragma unroll
    for (int j = 0; j < 32; ++j) s = fmaf(fmaxf(v[j], 0.f), t[j], s);
    reinterpret_cast<float*>(op.out)[static_cast<long long>(Y) * WW + X] = fmaxf(s, 0.f);
    return;
  } else {  // O_DOT_RELU: (already bias + ReLU'd) 32-channel pixel -> 1 channel
    load32<float>(op.dot_w, t);
    float s = op.dot_b[0];
#pragma unroll
    for (int j = 0; j < 32; ++j) s = fmaf(v[j], t[j], s);
    reinterpret_cast<float*>(op.out)[m] = fmaxf(s, 0.f);
    return;
  }

  if (op.res) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += resv[j];
  }
  if (op.res2) {
    load32<bf16>(reinterpret_cast<const bf16*>(op.res2) + m * op.ldres + n0, t);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += t[j];
  }
  if (op.out) {
    if (op.out_f32)
      store32(reinterpret_cast<float*>(op.out) + off, v);
    else
      store32(reinterpret_cast<bf16*>(op.out) + off, v, false);
  }
  if (op.out_relu) store32(reinterpret_cast<bf16*>(op.out_relu) + off, v, true);
}

__device__ __forceinline__ void load_res(const GemmOp& op, long long m, int n0, float (&r)[32]) {
  const long long roff = m * op.ldres + n0;
  if (op.res_f32) load32<float>(reinterpret_cast<const float*>(op.res) + roff, r);
  else load32<bf16>(reinterpret_cast<const bf16*>(op.res) + roff, r);
}

__device__ __forceinline__ uint32_t relu2(uint32_t packed) {  // max(x, 0) on a bf16 pair
  bf16x2 x = *reinterpret_cast<bf16x2*>(&packed);
  x = __hmax2(x, f2_to_h2(0.f, 0.f));
  return *reinterpret_cast<uint32_t*>(&x);
}

// EPI_TMA chunk: 32 columns of this thread's row (bias already added) -> activation / LayerScale /
// residuals -> bf16 into the warp's swizzled slab.  The bf16 residual was requested one chunk ahead.
__device__ __forceinline__ void epi_tma_chunk(const GemmOp& op, const GemmGroup& gp, long long m, int n0, float (&v)[32],
                                              const uint4 (&resraw)[4], bool has_res, uint32_t slab_row, int half,
                                              int swz, uint32_t relu_row = 0) {
  if (op.act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (op.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) gelu_erf2(v[j], v[j + 1]);
  }
  float t[32];
  if (gp.gamma) {
    load32<float>(gp.gamma + n0, t);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= t[j];
  }
  if (has_res) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bf16x2* h = reinterpret_cast<const bf16x2*>(&resraw[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = h2_to_f2(h[j]);
        v[8 * i + 2 * j] += f.x, v[8 * i + 2 * j + 1] += f.y;
      }
    }
  }
  if (op.res2) {
    load32<bf16>(reinterpret_cast<const bf16*>(op.res2) + m * op.ldres + n0, t);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += t[j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t w0 = pack2(v[8 * i], v[8 * i + 1]), w1 = pack2(v[8 * i + 2], v[8 * i + 3]);
    const uint32_t w2 = pack2(v[8 * i + 4], v[8 * i + 5]), w3 = pack2(v[8 * i + 6], v[8 * i + 7]);
    ptx::sts_u4(slab_row + (((half * 4 + i) ^ swz) << 4), w0, w1, w2, w3);
    if (relu_row) ptx::sts_u4(relu_row + (((half * 4 + i) ^ swz) << 4), relu2(w0), relu2(w1), relu2(w2), relu2(w3));
  }
}

// ---------------------------------------------------------------------------------------------
// Column-per-lane epilogue for row-major outputs.  A warp's 32x32 fp32 accumulator chunk (thread =
// row, straight from tcgen05.ld) is transposed through a per-warp shared-memory buffer; afterwards
// lane l owns column n0+l, so every global access of the warp is ONE contiguous 64/128-byte row
// segment (1 LSU wavefront) instead of 32 scattered 16-byte pieces, and bias / gamma are one
// register each.  Kept deliberately small (rolled loops, no per-row mode switches): the first,
// fully general version of this path was instruction-fetch bound.
struct ColSlab {
  int nv;          // valid rows of the 32-row slab
  int wp;          // row -> pixel step: 16 for matrix rows (offset = rr), W for 8x16 conv tiles
  long long row0;  // first output row / pixel of the slab
  int b, y, x;     // conv: image, first row, first column of the slab
};

// proj / fc2: out = res + gamma * (acc + bias), fp32 in place.  The 32 residual values of this
// lane's column were prefetched into registers before the accumulator was ready.
// fp32 residual epilogue, 128-bit form.  The warp's 32x32 accumulator chunk is staged row-per-thread
// into a 4 KB smem tile (16-byte chunks XOR-swizzled by row, conflict-free both ways); afterwards lane
// (rg = lane >> 3, cq = lane & 7) owns columns [4cq, 4cq+4) of rows rg, rg+4, ..., rg+28, so every
// global access is a 128-byte row segment moved by LDG.128 / STG.128: 4x fewer LSU instructions than
// the 32-bit column-per-lane form (the epilogue of proj / fc2 was LSU-issue bound).
__device__ __forceinline__ void prefetch_res4(const float* __restrict__ res, long long ld, const ColSlab& cs, int n0,
                                              int lane, float4 (&rv)[8]) {
  const float* p = res + (cs.row0 + (lane >> 3)) * ld + n0 + (lane & 7) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    rv[i] = (4 * i + (lane >> 3)) < cs.nv ? *reinterpret_cast<const float4*>(p + 4 * i * ld) : make_float4(0.f, 0.f, 0.f, 0.f);
}
template <bool LN_OUT>
__device__ __forceinline__ void epi_rows4_resid32(float* __restrict__ out, long long ld, const ColSlab& cs, int n0,
                                                  uint32_t stg, int lane, const float4 b4, const float4 g4,
                                                  const float4 (&rv)[8], bf16* __restrict__ xb, float (&ls)[8],
                                                  float (&lq)[8]) {
  const int rg = lane >> 3, cq = lane & 7;
  const long long off = (cs.row0 + rg) * ld + n0 + cq * 4;
  float* p = out + off;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = 4 * i + rg;
    const float4 a = ptx::lds_v4(stg + row * 128 + ((cq ^ (row & 7)) << 4));
    float4 o;
    o.x = fmaf(a.x + b4.x, g4.x, rv[i].x);
    o.y = fmaf(a.y + b4.y, g4.y, rv[i].y);
    o.z = fmaf(a.z + b4.z, g4.z, rv[i].z);
    o.w = fmaf(a.w + b4.w, g4.w, rv[i].w);
    if (row < cs.nv) {
      *reinterpret_cast<float4*>(p + 4 * i * ld) = o;
      if constexpr (LN_OUT) {
        const bf16x2 lo = f2_to_h2(o.x, o.y), hi = f2_to_h2(o.z, o.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo), pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(xb + off + 4 * i * ld) = pk;
      }
    }
    if constexpr (LN_OUT) {
      ls[i] += (o.x + o.y) + (o.z + o.w);
      lq[i] += fmaf(o.x, o.x, o.y * o.y) + fmaf(o.z, o.z, o.w * o.w);
    }
  }
}

// EPI selects the epilogue form at compile time so that each instantiation carries only its own
// code and registers (one monolithic epilogue made every added feature slow the hot GEMMs down):
//   EPI_TMA   bf16 output (row-major, NHWC conv tile or ConvT pixel shuffle), single store:
//             row-per-thread math -> smem slab -> TMA store, software pipelined
//   EPI_TMA2  same with a second, ReLU'd copy of the output (x feeds the residual add, relu(x) the next
//             conv's TMA loads): two slabs per warp, one ring stage fewer
//   EPI_RES32 fp32 residual update in place (proj / fc2): smem transpose, column-per-lane, deep prefetch
//   EPI_MISC  everything else: ConvT scatter and dual (x, relu(x)) stores via TMA, direct stores for
//             patch-embed placement, fused dots, small / odd shapes
//   EPI_TMA_LN   EPI_TMA for a LayerNorm-folded GEMM (qkv / fc1): acc * rstd[m] - rstd[m] * mean[m] * c[n] + d[n]
//   EPI_RES32_LN EPI_RES32 that also emits bf16(x) and per-row partial (sum, sum of squares) for the next
//                LayerNorm-folded GEMM (common.cuh GemmOp::ln_stats)
//   EPI_RES16_LN the same over a residual stream stored as a (hi, lo) pair of 16-bit arrays, x = hi + lo with
//                hi = round16(x), lo = round16(x - hi) (GemmOp::ln_xlo): hi IS the next folded GEMM's operand, so the
//                separate copy disappears -- 4 B read + 4 B written per element instead of 4 + 6
constexpr int RES_PREFETCH_DEFAULT = 0;  // measured A/B pending: opt-in
enum { EPI_TMA = 0, EPI_RES32 = 1, EPI_MISC = 2, EPI_TMA2 = 3, EPI_TMA_LN = 4, EPI_RES32_LN = 5, EPI_RES16_LN = 6 };

template <int BN, int CL, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ WeightMaps tmW,
               const GemmOp op, const TileGeom g) {
  constexpr int STAGES = Cfg<BN, CL, EPI>::STAGES;
  constexpr int STG_WARP_BYTES = Cfg<BN, CL, EPI>::STG_WARP;
  constexpr bool PAIR = CL == 2;
  constexpr int KSUB = Cfg<BN, CL, EPI>::KSUB;
  constexpr uint32_t A_SUB = BM * BK * 2, B_SUB = (BN / CL) * BK * 2;  // one 64-wide k-block (B: this CTA's share)
  constexpr uint32_t A_BYTES = KSUB * A_SUB;        // per stage
  constexpr uint32_t B_BYTES = KSUB * B_SUB;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  constexpr uint32_t IDESC = ptx::umma_idesc_bf16(BM * CL, BN);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + 8 * STG_WARP_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const uint32_t stg_all = ptx::smem_u32(smem + STAGES * STAGE_BYTES);  // epilogue staging, 8 warps, 1024-B aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = g.m_units * g.n_tiles;
  const int crank = CL > 1 ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;  // cluster index / number of clusters
  // pair mode: the smem-full and accumulator-empty barriers that gate the MMA live in the leader CTA
  const uint32_t tempty_leader = PAIR ? ptx::mapa_u32(&tempty[0], 0) : ptx::smem_u32(&tempty[0]);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW.b[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull[a], 1);
      ptx::mbar_init(&tempty[a], 8 * CL);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    if (PAIR) {
      ptx::tmem_alloc_pair(tmem_slot, TMEM_COLS);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_slot, TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if (CL > 1) ptx::cluster_sync_all();  // peers' barriers must be initialised before any remote arrive
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above overlapped the previous kernel's tail; nothing below may run before its
  // writes are visible.  The next kernel may start placing CTAs as ours retire.
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();

  // register split: the four control warps need few registers, the epilogue warps hold a whole
  // tile's residual values in flight (384 x 168 = 128 x 40 + 256 x 232)
  if (warp < EPI_WARP0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp walks the loop (warp-uniform control flow: coordinates and descriptors stay in
    // uniform registers); one elected lane arms the barrier and issues the copies.
    {
      int s = 0;
      uint32_t ph = 0;
      for (int t = cid; t < total_units; t += ncl) {
        const int tt = g.reverse ? total_units - 1 - t : t;
        const int nt = tt % g.n_tiles, mu = tt / g.n_tiles;
        int b = 0, y0 = 0, x0 = 0, a_row = 0;
        const CUtensorMap* tmB = &tmW.b[0];
        if (op.a_mode == A_CONV3X3) {
          const int mt = mu * CL + crank;  // padding tiles land beyond the last image: TMA zero-fills
          const int per_img = g.tiles_x * g.tiles_y;
          b = mt / per_img;
          const int r = mt - b * per_img;
          y0 = (r / g.tiles_x) * TILE_H;
          x0 = (r % g.tiles_x) * TILE_W;
        } else {
          const int gi = unit_group(g, mu);
          const int lmt = (mu - g.unit_start[gi]) * CL + crank;
          a_row = static_cast<int>(op.grp[gi].a_row_off) + lmt * BM;
          tmB = &tmW.b[gi];
          if constexpr (EPI == EPI_RES32 || EPI == EPI_RES32_LN) {
            // The epilogue reads this CTA's 128 x BN fp32 residual tile as 128-byte row segments (one per
            // 32-column chunk): scattered at the DRAM page level, and proj / fc2 are bound by exactly that
            // traffic.  Request the tile's rows as whole BN*4-byte runs into L2 one main loop ahead of the
            // epilogue that consumes them.
            if (g.res_prefetch) {
              const int rows = op.grp[gi].M - lmt * BM;
              const float* rp = reinterpret_cast<const float*>(op.res) +
                                (op.grp[gi].o_row_off + static_cast<long long>(lmt) * BM) * op.ldres;
              if (op.N == op.ldres) {
                // full-width rows: the n-tiles of an m-unit run concurrently on neighbouring clusters, so the
                // CTA of n-tile 0 requests its 128 whole rows for all of them -- one contiguous block, 4 rows
                // (16 KB at N = 1024) per instruction (UBLKPF takes its address from a uniform register: the
                // per-lane requests are issued one after the other)
                const int r0 = 4 * lane, nr = rows - r0 < 4 ? rows - r0 : 4;
                if (nt == 0 && nr > 0) ptx::prefetch_l2_bulk(rp + static_cast<long long>(r0) * op.ldres, nr * op.ldres * 4);
              } else {
                const int cols = op.N - nt * BN < BN ? op.N - nt * BN : BN;
                for (int r = lane; r < BM && r < rows; r += 32)
                  ptx::prefetch_l2_bulk(rp + static_cast<long long>(r) * op.ldres + nt * BN, cols * 4);
              }
              __syncwarp();
            }
          }
        }
        int tap = 0, c0 = 0;  // conv: running (filter tap, channel offset) of the k-block
        for (int kb = 0; kb < g.k_blocks; kb += KSUB) {
          // an odd K leaves a half-filled last stage
          const int nsub = KSUB == 1 ? 1 : (g.k_blocks - kb < KSUB ? g.k_blocks - kb : KSUB);
          ptx::mbar_wait(&empty[s], ph ^ 1);
          const bool leader_lane = ptx::elect_one();
          if (leader_lane) {
            if constexpr (PAIR) {
              if (crank == 0) ptx::mbar_expect_tx(&full[s], 2 * nsub * (A_SUB + B_SUB));
            } else {
              ptx::mbar_expect_tx(&full[s], nsub * (A_SUB + B_SUB));
            }
          }
#pragma unroll
          for (int j = 0; j < KSUB; ++j) {
            if (j < nsub) {
              if (leader_lane) {
                if constexpr (PAIR) {
                  // both CTAs' bytes complete on the leader's barrier; the leader armed it for the pair
                  const uint32_t fbar = ptx::mapa_u32(&full[s], 0);
                  if (op.a_mode == A_CONV3X3) {
                    const int ky = tap / 3, kx = tap - ky * 3;
                    ptx::tma_load_4d_pair(sA + s * A_BYTES + j * A_SUB, &tmA, fbar, c0, x0 + kx - 1, y0 + ky - 1, b);
                  } else {
                    ptx::tma_load_2d_pair(sA + s * A_BYTES + j * A_SUB, &tmA, fbar, (kb + j) * BK, a_row);
                  }
                  ptx::tma_load_2d_pair(sB + s * B_BYTES + j * B_SUB, tmB, fbar, (kb + j) * BK, nt * BN + crank * (BN / 2));
                } else {
                  if (op.a_mode == A_CONV3X3) {
                    const int ky = tap / 3, kx = tap - ky * 3;
                    ptx::tma_load_4d(sA + s * A_BYTES + j * A_SUB, &tmA, &full[s], c0, x0 + kx - 1, y0 + ky - 1, b);
                  } else {
                    ptx::tma_load_2d(sA + s * A_BYTES + j * A_SUB, &tmA, &full[s], (kb + j) * BK, a_row);
                  }
                  ptx::tma_load_2d(sB + s * B_BYTES + j * B_SUB, tmB, &full[s], (kb + j) * BK, nt * BN);
                }
              }
              c0 += BK;
              if (c0 == op.C) c0 = 0, ++tap;
            }
          }
          __syncwarp();
          if (++s == STAGES) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1 && crank == 0) {
    // ------------------------------------------------------------ MMA issuer (pair: the leader CTA only)
    // Warp-converged loop, one elected lane issues.  With the loop under `if (lane == 0)` every
    // tcgen05.mma was wrapped in an elect / vote / R2UR sequence and the descriptor arithmetic ran on
    // the (slow, serial) uniform datapath: ~90 dependent instructions per k-block, as long as the
    // 512 tensor-core cycles they were supposed to hide behind.  Descriptors are now base + 2*k.
    {
      // 64-bit smem descriptor: lo = (addr >> 4) | LBO(1) << 16, hi = SBO(64) | version(1) << 14 | SW128(2) << 29
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = ((ptx::smem_u32(sA) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t b_lo0 = ((ptx::smem_u32(sB) & 0x3FFFF) >> 4) | (1u << 16);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = cid; t < total_units; t += ncl, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        ptx::mbar_wait(&tempty[acc], acc_ph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < g.k_blocks; kb += KSUB) {
          const int nsub = KSUB == 1 ? 1 : (g.k_blocks - kb < KSUB ? g.k_blocks - kb : KSUB);
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_lo = a_lo0 + s * (A_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + s * (B_BYTES >> 4);
          if (ptx::elect_one()) {
#pragma unroll
            for (int j = 0; j < KSUB; ++j) {
              if (j < nsub) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  // +32 B per K = 16 step inside a 64-wide k-block, + one sub-tile per k-block
                  const uint64_t da = (static_cast<uint64_t>(DESC_HI) << 32) | (a_lo + j * (A_SUB >> 4) + 2 * k);
                  const uint64_t db = (static_cast<uint64_t>(DESC_HI) << 32) | (b_lo + j * (B_SUB >> 4) + 2 * k);
                  if (PAIR) ptx::umma_bf16_pair(d_tmem, da, db, IDESC, (kb | j | k) != 0);
                  else ptx::umma_bf16(d_tmem, da, db, IDESC, (kb | j | k) != 0);
                }
              }
            }
            if (PAIR) {  // release the stage / publish the accumulator in both CTAs
              ptx::umma_commit_pair(&empty[s]);
              if (kb + KSUB >= g.k_blocks) ptx::umma_commit_pair(&tfull[acc]);
            } else {
              ptx::umma_commit(&empty[s]);
              if (kb + KSUB >= g.k_blocks) ptx::umma_commit(&tfull[acc]);
            }
          }
          __syncwarp();
          if (++s == STAGES) s = 0, ph ^= 1;
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int grp = (warp - EPI_WARP0) >> 2;  // column half
    constexpr int COLS_PER_GRP = (BN >= 64) ? BN / 2 : BN;
    const int row = q * 32 + lane;
    int it = 0;
    for (int t = cid; t < total_units; t += ncl, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const int tt = g.reverse ? total_units - 1 - t : t;
      const int nt = tt % g.n_tiles, mu = tt / g.n_tiles;
      long long m, m_a = 0;  // output row / A row of this thread
      bool valid;
      int gi = 0;
      if (op.a_mode == A_CONV3X3) {
        const int mt = mu * CL + crank;
        const int per_img = g.tiles_x * g.tiles_y;
        const int b = mt / per_img;
        const int r = mt - b * per_img;
        const int y = (r / g.tiles_x) * TILE_H + row / TILE_W;
        const int x = (r % g.tiles_x) * TILE_W + row % TILE_W;
        m = (static_cast<long long>(b) * op.H + y) * op.W + x;
        valid = mt < g.tiles_in_group[0];
      } else {
        gi = unit_group(g, mu);
        const int m_local = ((mu - g.unit_start[gi]) * CL + crank) * BM + row;
        valid = m_local < op.grp[gi].M;
        m = op.grp[gi].o_row_off + m_local;
        m_a = op.grp[gi].a_row_off + m_local;
      }
      const GemmGroup& gp = op.grp[gi];
      const bool active = (BN >= 64 || grp == 0);
      const int col0 = nt * BN + grp * COLS_PER_GRP;
      ColSlab cs;
      if (op.a_mode == A_CONV3X3) {
        const int mt = mu * CL + crank;
        const int per_img = g.tiles_x * g.tiles_y;
        const int b = mt / per_img;
        const int r = mt - b * per_img;
        cs.row0 = (static_cast<long long>(b) * op.H + (r / g.tiles_x) * TILE_H + 2 * q) * op.W + (r % g.tiles_x) * TILE_W;
        cs.wp = op.W;
        cs.nv = mt < g.tiles_in_group[0] ? 32 : 0;
        cs.b = b, cs.y = (r / g.tiles_x) * TILE_H + 2 * q, cs.x = (r % g.tiles_x) * TILE_W;
      } else {
        const int ml0 = ((mu - g.unit_start[gi]) * CL + crank) * BM + q * 32;
        const int left = op.grp[gi].M - ml0;
        cs.nv = left >= 32 ? 32 : (left > 0 ? left : 0);
        cs.row0 = op.grp[gi].o_row_off + ml0;
        cs.wp = 16;
        cs.b = cs.y = cs.x = 0;
      }
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + grp * COLS_PER_GRP;
      if (!active) {
        ptx::mbar_wait(&tfull[acc], acc_ph);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader + acc * 8);
        continue;
      }
      if constexpr (EPI == EPI_RES16_LN) {
        // ---- proj / fc2 over the (hi, lo) pair stream: x = hi + lo; x += gamma * (acc + bias); hi, lo and the row sums
        // of the new x go back in place.  Same smem transpose as the fp32 form below, but a lane owns EIGHT columns
        // (cq = lane & 3) of rows rg, rg + 8, rg + 16, rg + 24 (rg = lane >> 2), so that hi and lo move as 16-byte
        // pieces: the first version with 8-byte pieces doubled the number of memory requests in flight per warp and
        // ran proj at HALF the speed of the fp32 form (118 us vs 66 us) although it moved fewer bytes.
        static_assert(COLS_PER_GRP == 128, "LN partial sums are kept per 128-column slice");
        bf16* const xhi = reinterpret_cast<bf16*>(op.ln_xb);
        bf16* const xlo = reinterpret_cast<bf16*>(op.ln_xlo);
        const uint32_t tile = stg_all + (warp - EPI_WARP0) * STG_WARP_BYTES;
        const int rg = lane >> 2, cq = lane & 3;
        const long long ld = op.ldo;
        float ls[4] = {0.f, 0.f, 0.f, 0.f}, lq[4] = {0.f, 0.f, 0.f, 0.f};
        auto prefetch = [&](int cc, uint4 (&h)[4], uint4 (&l)[4]) {
          const long long o = (cs.row0 + rg) * ld + cc + cq * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (8 * i + rg < cs.nv) {
              h[i] = *reinterpret_cast<const uint4*>(xhi + o + 8 * i * ld);
              l[i] = *reinterpret_cast<const uint4*>(xlo + o + 8 * i * ld);
            } else {
              h[i] = l[i] = make_uint4(0u, 0u, 0u, 0u);
            }
          }
        };
        auto ld_bg = [&](int c, float4 (&b)[2], float4 (&gm)[2]) {
          const int nq = col0 + c + cq * 8;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            b[k] = gp.bias ? *reinterpret_cast<const float4*>(gp.bias + nq + 4 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
            gm[k] = gp.gamma ? *reinterpret_cast<const float4*>(gp.gamma + nq + 4 * k) : make_float4(1.f, 1.f, 1.f, 1.f);
          }
        };
        uint4 hv[4], lv[4], hn[4], ln1[4];
        prefetch(col0, hv, lv);
        prefetch(col0 + 32, hn, ln1);
        float4 b4[2], g4[2], b4n[2], g4n[2];
        ld_bg(0, b4, g4);
        ptx::mbar_wait(&tfull[acc], acc_ph);
        ptx::tc_fence_after();
        uint32_t r[32];
        ptx::tmem_ld32(t_acc, r);
#pragma unroll 1
        for (int c = 0; c < COLS_PER_GRP; c += 32) {
          uint4 hn2[4], ln2[4];
          if (c + 64 < COLS_PER_GRP) prefetch(col0 + c + 64, hn2, ln2);
          if (c + 32 < COLS_PER_GRP) ld_bg(c + 32, b4n, g4n);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            ptx::sts_v4(tile + lane * 128 + ((j ^ (lane & 7)) << 4), __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                        __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          if (c + 32 < COLS_PER_GRP) {
            ptx::tmem_ld32(t_acc + c + 32, r);
          } else {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader + acc * 8);
          }
          __syncwarp();
          const long long off = (cs.row0 + rg) * ld + col0 + c + cq * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = 8 * i + rg;
            const float4 a0 = ptx::lds_v4(tile + row * 128 + (((2 * cq) ^ (row & 7)) << 4));
            const float4 a1 = ptx::lds_v4(tile + row * 128 + (((2 * cq + 1) ^ (row & 7)) << 4));
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b4[0].x, b4[0].y, b4[0].z, b4[0].w, b4[1].x, b4[1].y, b4[1].z, b4[1].w};
            const float gv[8] = {g4[0].x, g4[0].y, g4[0].z, g4[0].w, g4[1].x, g4[1].y, g4[1].z, g4[1].w};
            const uint32_t hw[4] = {hv[i].x, hv[i].y, hv[i].z, hv[i].w}, lw[4] = {lv[i].x, lv[i].y, lv[i].z, lv[i].w};
            float o[8];
            uint32_t nh[4], nl[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 xh = h2_to_f2(*reinterpret_cast<const bf16x2*>(&hw[k]));
              const float2 xl = h2_to_f2(*reinterpret_cast<const bf16x2*>(&lw[k]));
              o[2 * k] = fmaf(av[2 * k] + bv[2 * k], gv[2 * k], xh.x + xl.x);
              o[2 * k + 1] = fmaf(av[2 * k + 1] + bv[2 * k + 1], gv[2 * k + 1], xh.y + xl.y);
              const bf16x2 h2 = f2_to_h2(o[2 * k], o[2 * k + 1]);
              const float2 rb = h2_to_f2(h2);
              const bf16x2 l2 = f2_to_h2(o[2 * k] - rb.x, o[2 * k + 1] - rb.y);
              nh[k] = *reinterpret_cast<const uint32_t*>(&h2), nl[k] = *reinterpret_cast<const uint32_t*>(&l2);
              ls[i] += o[2 * k] + o[2 * k + 1];
              lq[i] += fmaf(o[2 * k], o[2 * k], o[2 * k + 1] * o[2 * k + 1]);
            }
            if (row < cs.nv) {
              *reinterpret_cast<uint4*>(xhi + off + 8 * i * ld) = make_uint4(nh[0], nh[1], nh[2], nh[3]);
              *reinterpret_cast<uint4*>(xlo + off + 8 * i * ld) = make_uint4(nl[0], nl[1], nl[2], nl[3]);
            }
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) hv[i] = hn[i], lv[i] = ln1[i], hn[i] = hn2[i], ln1[i] = ln2[i];
#pragma unroll
          for (int k = 0; k < 2; ++k) b4[k] = b4n[k], g4[k] = g4n[k];
        }
        // the 4 lanes that share a row add up their 8 columns x 4 chunks; slot = this warp's 128-column slice
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float sv = ls[i], qv = lq[i];
          sv += __shfl_xor_sync(0xffffffffu, sv, 1), qv += __shfl_xor_sync(0xffffffffu, qv, 1);
          sv += __shfl_xor_sync(0xffffffffu, sv, 2), qv += __shfl_xor_sync(0xffffffffu, qv, 2);
          const int rr = 8 * i + rg;
          if (cq == 0 && rr < cs.nv)
            *reinterpret_cast<float2*>(op.ln_stats_out + ((cs.row0 + rr) * LN_SLOTS + (col0 >> 7)) * 2) = make_float2(sv, qv);
        }
      } else if constexpr (EPI == EPI_RES32 || EPI == EPI_RES32_LN) {
        constexpr bool LN_OUT = EPI == EPI_RES32_LN;
        float ls[8], lq[8];  // LN_OUT: this lane's partial row sums of rows (lane >> 3) + 4 i
#pragma unroll
        for (int i = 0; i < 8; ++i) ls[i] = lq[i] = 0.f;
        // ---- proj / fc2: out = res + gamma * (acc + bias), fp32 in place, column-per-lane through smem.
        // The residual values of TWO chunks are in flight ahead of the one being processed; the first
        // two are requested before the accumulator is even ready.
        const float* resp = reinterpret_cast<const float*>(op.res);
        const uint32_t tile = stg_all + (warp - EPI_WARP0) * STG_WARP_BYTES;
        float4 resv[8], resn[8];
        prefetch_res4(resp, op.ldres, cs, col0, lane, resv);
        if (COLS_PER_GRP > 32) prefetch_res4(resp, op.ldres, cs, col0 + 32, lane, resn);
        auto ld_bg = [&](int c, float4& b4, float4& g4) {
          const int nq = col0 + c + (lane & 7) * 4;
          b4 = gp.bias ? *reinterpret_cast<const float4*>(gp.bias + nq) : make_float4(0.f, 0.f, 0.f, 0.f);
          g4 = gp.gamma ? *reinterpret_cast<const float4*>(gp.gamma + nq) : make_float4(1.f, 1.f, 1.f, 1.f);
        };
        float4 b4, g4, b4n, g4n;
        ld_bg(0, b4, g4);
        ptx::mbar_wait(&tfull[acc], acc_ph);
        ptx::tc_fence_after();
        uint32_t r[32];
        ptx::tmem_ld32(t_acc, r);
#pragma unroll 1
        for (int c = 0; c < COLS_PER_GRP; c += 32) {
          float4 resn2[8];
          if (c + 64 < COLS_PER_GRP) prefetch_res4(resp, op.ldres, cs, col0 + c + 64, lane, resn2);
          if (c + 32 < COLS_PER_GRP) ld_bg(c + 32, b4n, g4n);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            ptx::sts_v4(tile + lane * 128 + ((j ^ (lane & 7)) << 4), __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                        __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          if (c + 32 < COLS_PER_GRP) {
            ptx::tmem_ld32(t_acc + c + 32, r);  // next chunk flies during this chunk's read-modify-write
          } else {  // accumulator is out of TMEM: release it, stores overlap the next tile
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader + acc * 8);
          }
          __syncwarp();
          epi_rows4_resid32<LN_OUT>(reinterpret_cast<float*>(op.out), op.ldo, cs, col0 + c, tile, lane, b4, g4, resv,
                                    reinterpret_cast<bf16*>(op.ln_xb), ls, lq);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) resv[j] = resn[j], resn[j] = resn2[j];
          b4 = b4n, g4 = g4n;
        }
        if constexpr (LN_OUT) {
          // the 8 lanes that share a row add up their 16 columns x 4 chunks; slot = this warp's 128-column slice
          static_assert(COLS_PER_GRP == 128, "LN partial sums are kept per 128-column slice");
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float sv = ls[i], qv = lq[i];
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              sv += __shfl_xor_sync(0xffffffffu, sv, o);
              qv += __shfl_xor_sync(0xffffffffu, qv, o);
            }
            const int rr = 4 * i + (lane >> 3);
            if ((lane & 7) == 0 && rr < cs.nv)
              *reinterpret_cast<float2*>(op.ln_stats_out + ((cs.row0 + rr) * LN_SLOTS + (col0 >> 7)) * 2) = make_float2(sv, qv);
          }
        }
      } else if constexpr (EPI == EPI_TMA || EPI == EPI_TMA2 || EPI == EPI_TMA_LN) {
        constexpr bool LN_IN = EPI == EPI_TMA_LN;
        // ---- row-per-thread math, bf16 output through a swizzled smem slab + TMA store.  Software
        // pipelined: as soon as chunk c has been copied out of the TMEM-load registers (bias add), the
        // TMEM load, bias and residual of chunk c+1 are issued and fly during the math of chunk c.
        static_assert(COLS_PER_GRP % 64 == 0, "EPI_TMA works on 64-column slabs");
        const uint32_t slab = stg_all + (warp - EPI_WARP0) * STG_WARP_BYTES;
        const uint32_t slab_relu = EPI == EPI_TMA2 ? slab + 4096 : 0;
        const bool has_res = !LN_IN && op.res != nullptr && valid;
        const bf16* resp = reinterpret_cast<const bf16*>(op.res) + m * op.ldres + col0;
        uint32_t r[32];
        float bias[32];
        float lnc[LN_IN ? 32 : 1];
        float ln_rs = 0.f, ln_nm = 0.f;  // rstd and -rstd * mean of this thread's row
        if constexpr (LN_IN) {
          if (valid) {
            const float4* sp = reinterpret_cast<const float4*>(op.ln_stats + m_a * (2 * LN_SLOTS));
            const float4 p0 = sp[0], p1 = sp[1], p2 = sp[2], p3 = sp[3];  // (s0 q0 s1 q1) (s2 q2 s3 q3) ...
            const float sum = ((p0.x + p0.z) + (p1.x + p1.z)) + ((p2.x + p2.z) + (p3.x + p3.z));
            const float sq = ((p0.y + p0.w) + (p1.y + p1.w)) + ((p2.y + p2.w) + (p3.y + p3.w));
            const float mean = sum * (1.f / 1024.f);
            const float var = fmaxf(fmaf(-mean, mean, sq * (1.f / 1024.f)), 0.f);
            ln_rs = 1.f / sqrtf(var + 1e-6f);
            ln_nm = -ln_rs * mean;
          }
        }
        uint4 resc[4], resn[4];
        auto prefetch = [&](int c) {  // bias + residual of the chunk at column offset c
          if (gp.bias) load32<float>(gp.bias + (op.bias_mod ? (col0 + c) % op.bias_mod : col0 + c), bias);
          if constexpr (LN_IN) load32<float>(gp.ln_c + col0 + c, lnc);
          if (has_res) {
#pragma unroll
            for (int i = 0; i < 4; ++i) resn[i] = reinterpret_cast<const uint4*>(resp + c)[i];
          }
        };
        prefetch(0);  // in flight while the MMA of this tile finishes
        ptx::mbar_wait(&tfull[acc], acc_ph);
        ptx::tc_fence_after();
        ptx::tmem_ld32(t_acc, r);
#pragma unroll 1
        for (int c = 0; c < COLS_PER_GRP; c += 64) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int cc = c + half * 32;
            const bool more = cc + 32 < COLS_PER_GRP;
            float v[32];
            ptx::tmem_ld_wait();
            if constexpr (LN_IN) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), ln_rs, fmaf(ln_nm, lnc[j], bias[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + (gp.bias ? bias[j] : 0.f);
              if (op.border_cb != nullptr && valid) {
                // composed 1x1 -> conv3x3: the 1x1's bias does not exist in the zero padding (GemmOp::border_cb)
                const int px = static_cast<int>(m % op.W), py = static_cast<int>((m / op.W) % op.H);
                if (px == 0 || px == op.W - 1 || py == 0 || py == op.H - 1) {
                  for (int ky = 0; ky < 3; ++ky)
                    for (int kx = 0; kx < 3; ++kx) {
                      const int yy = py + ky - 1, xx = px + kx - 1;
                      if (yy < 0 || yy >= op.H || xx < 0 || xx >= op.W) {
                        float t[32];
                        load32<float>(op.border_cb + (ky * 3 + kx) * op.N + col0 + cc, t);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] -= t[j];
                      }
                    }
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) resc[i] = resn[i];
            if (more) {
              ptx::tmem_ld32(t_acc + cc + 32, r);
              prefetch(cc + 32);
            } else {  // accumulator is out of TMEM: release it, the rest overlaps the next tile's MMA
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader + acc * 8);
            }
            if (half == 0) {
              if (lane == 0) ptx::tma_store_wait_read();  // the previous slab store has drained the buffer
              __syncwarp();
            }
            epi_tma_chunk(op, gp, m, col0 + cc, v, resc, has_res, slab + lane * 128, half, lane & 7,
                          EPI == EPI_TMA2 ? slab_relu + lane * 128 : 0);
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0 && cs.nv > 0) {
            auto issue = [&](const CUtensorMap* tm, uint32_t src) {
              const int ccol = op.col_off + col0 + c;
              if (op.out_mode == O_CONVT2X2) {
                // n-chunk -> (parity, channel); the slab is 32 consecutive input pixels of one image row
                const int nn = col0 + c, cq = nn / op.cout;
                ptx::tma_store_5d(tm, src, op.col_off + nn - cq * op.cout, cq & 1, static_cast<int>(cs.row0 % op.W),
                                  cq >> 1, static_cast<int>(cs.row0 / op.W));
              } else if (op.a_mode == A_CONV3X3) {
                ptx::tma_store_4d(tm, src, ccol, cs.x, cs.y, cs.b);
              } else {
                ptx::tma_store_2d(tm, src, ccol, static_cast<int>(cs.row0));
              }
            };
            issue(op.out_mode == O_ROWMAJOR && op.a_mode != A_CONV3X3 ? &tmW.o[gi] : &tmW.o[0], slab);
            if (EPI == EPI_TMA2) issue(&tmW.orelu, slab_relu);
            ptx::tma_store_commit();
          }
        }
      } else {
        // ---- EPI_MISC: row-per-thread math, direct global stores (patch-embed token placement, fused
        // dots, fp32 / odd-shaped outputs, outputs whose geometry the TMA-store forms do not cover)
        const bool has_res = op.res != nullptr && valid;
        float resv[32];
        if (has_res) load_res(op, m, col0, resv);  // in flight while the MMA of this tile finishes
        ptx::mbar_wait(&tfull[acc], acc_ph);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < COLS_PER_GRP; c += 32) {
          uint32_t r[32];
          ptx::tmem_ld32(t_acc + c, r);
          const bool more = c + 32 < COLS_PER_GRP;
          float resn[32];
          if (has_res && more) load_res(op, m, col0 + c + 32, resn);
          ptx::tmem_ld_wait();
          if (!more) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader + acc * 8);
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (valid) epilogue_chunk(op, gp, m, col0 + c, v, resv);
          if (has_res && more) {
#pragma unroll
            for (int j = 0; j < 32; ++j) resv[j] = resn[j];
          }
        }
      }
    }
  }

  if (warp >= EPI_WARP0 && lane == 0) ptx::tma_store_wait_read();  // smem must outlive the bulk stores
  __syncwarp();
  ptx::tc_fence_before();
  if (CL > 1) ptx::cluster_sync_all();  // no CTA may exit while the pair's MMA can still read its smem / write its TMEM
  else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    DP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    DP_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct TmapKey {
  const void* ptr;
  uint64_t d0, d1, d2, d3, d4, s1, s2, s3, s4;
  uint32_t b0, b1, b2, b3, b4;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && d3 == o.d3 && d4 == o.d4 && s1 == o.s1 &&
           s2 == o.s2 && s3 == o.s3 && s4 == o.s4 && b0 == o.b0 && b1 == o.b1 && b2 == o.b2 && b3 == o.b3 && b4 == o.b4;
  }
};
struct TmapHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    for (uint64_t v : {k.d0, k.d1, k.d2, k.d3, k.d4, k.s1, k.s2, k.s3, k.s4, (uint64_t)k.b0, (uint64_t)k.b1, (uint64_t)k.b2,
                       (uint64_t)k.b3})
      h = h * 1000003u ^ v;
    return h;
  }
};
std::unordered_map<TmapKey, CUtensorMap, TmapHash>& tmap_cache() {
  static std::unordered_map<TmapKey, CUtensorMap, TmapHash> c;
  return c;
}
// The cache is process-wide (device pointers are unique across GPUs under UVA) and shared by every engine and host
// thread: lookups and inserts take this mutex, and a descriptor is handed out BY VALUE, so a concurrent
// tmap_cache_clear() (engine destruction, test entry points) can never pull it from under a launch in flight.
std::mutex& tmap_mutex() {
  static std::mutex m;
  return m;
}

CUtensorMap get_tmap(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
  TmapKey key{ptr, dims[0], dims[1], rank > 2 ? dims[2] : 0, rank > 3 ? dims[3] : 0, rank > 4 ? dims[4] : 0,
              strides_bytes[0], rank > 2 ? strides_bytes[1] : 0, rank > 3 ? strides_bytes[2] : 0,
              rank > 4 ? strides_bytes[3] : 0, box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
              rank > 4 ? box[4] : 0};
  std::lock_guard<std::mutex> lk(tmap_mutex());
  auto& cache = tmap_cache();
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  CUtensorMap tm;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) gd[i] = dims[i], bx[i] = box[i];
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  DP_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base must be 16-byte aligned");
  CUresult r = encode_fn()(&tm, DP_TMAP_ELEM, rank, const_cast<void*>(ptr), gd, gs, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DP_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return cache.emplace(key, tm).first->second;
}

int num_sms() {
  static const int n = [] {   // every GPU of a B200 box has the same SM count
    int dev, v;
    DP_CUDA(cudaGetDevice(&dev));
    DP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    return v;
  }();
  return n;
}

// L2 persistence of the fp32 residual stream (opt-in: DEPTHPRO_L2_PERSIST_MB=<MB>, or gemm_tc_set_l2_persist()).
// The set-aside is device-wide (cudaLimitPersistingL2CacheSize) and comes out of every other kernel's L2.
static std::atomic<long long> g_l2_persist_req{-1};   // requested MB; -1 = read the environment on first use
static std::atomic<size_t> g_l2_carve{0}, g_l2_window{0};
static std::atomic<bool> g_l2_applied{false};
static std::mutex g_l2_mu;
static void set_l2_persist_impl(int mb) {
  g_l2_persist_req = mb < 0 ? 0 : mb;
  g_l2_applied = false;
}
static void l2_apply() {
  std::lock_guard<std::mutex> lk(g_l2_mu);
  if (g_l2_applied) return;
  if (g_l2_persist_req < 0) {
    const char* e = getenv("DEPTHPRO_L2_PERSIST_MB");
    g_l2_persist_req = e ? atoll(e) : 0;
  }
  static bool ever_on = false;
  if (g_l2_persist_req == 0 && !ever_on) {  // default path: never touch the device limits
    g_l2_carve = 0, g_l2_applied = true;
    return;
  }
  ever_on = true;
  int dev = 0, max_persist = 0, max_window = 0;
  DP_CUDA(cudaGetDevice(&dev));
  DP_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
  DP_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
  size_t want = static_cast<size_t>(g_l2_persist_req) << 20;
  if (want > static_cast<size_t>(max_persist)) want = static_cast<size_t>(max_persist);
  DP_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
  if (want == 0) DP_CUDA(cudaCtxResetPersistingL2Cache());
  g_l2_carve = want, g_l2_window = static_cast<size_t>(max_window);
  g_l2_applied = true;
  if (getenv("DEPTHPRO_VERBOSE"))
    fprintf(stderr, "[depthpro] L2 persisting set-aside %zu MB (device max %d MB, window max %d MB)\n", want >> 20,
            max_persist >> 20, max_window >> 20);
}
static size_t l2_persist_bytes() {
  if (!g_l2_applied) l2_apply();
  return g_l2_carve;
}
static size_t l2_window_max() { return g_l2_window; }

static std::atomic<int> g_sm_limit{0};

template <int BN, int CL, int EPI>
void launch(const GemmOp& op, const TileGeom& g, const CUtensorMap& tmA, const WeightMaps& tmW,
            cudaStream_t stream) {
  using C = Cfg<BN, CL, EPI>;
  constexpr size_t SMEM = C::STAGES * C::STAGE + 1024 /*align*/ + 256 /*barriers*/ + 8 * C::STG_WARP;
  static std::atomic<unsigned long long> configured{0};
  if (first_use_on_device(configured)) {
    DP_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, CL, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  const int units = g.m_units * g.n_tiles;
  const int lim = g_sm_limit.load();  // experiment knob (gemm_tc_set_sm_limit): run on at most this many SMs
  const int max_clusters = (lim > 0 && lim < num_sms() ? lim : num_sms()) / CL;
  const int clusters = units < max_clusters ? units : max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[3];
  int na = 0;
  if constexpr (EPI == EPI_RES32 || EPI == EPI_RES32_LN) {
    // Keep the fp32 residual stream resident in L2 across the block's kernels: accesses of THIS launch that fall
    // into x's window are marked persisting (proj / fc2 are bound by x's HBM round trip, scripts/ubench/rmw.cu).
    const size_t carve = l2_persist_bytes();
    if (carve) {
      long long rows = 0;
      for (int i = 0; i < op.ngroups; ++i) rows = std::max<long long>(rows, op.grp[i].o_row_off + op.grp[i].M);
      size_t bytes = static_cast<size_t>(rows) * op.ldres * 4;
      if (bytes > l2_window_max()) bytes = l2_window_max();
      attr[na].id = cudaLaunchAttributeAccessPolicyWindow;
      attr[na].val.accessPolicyWindow.base_ptr = const_cast<void*>(op.res);
      attr[na].val.accessPolicyWindow.num_bytes = bytes;
      attr[na].val.accessPolicyWindow.hitRatio = bytes <= carve ? 1.0f : static_cast<float>(static_cast<double>(carve) / bytes);
      attr[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      ++na;
    }
  }
  if (CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CL, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DP_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CL, EPI>, tmA, tmW, op, g));
  count_launch();
}

}  // namespace

void tmap_cache_clear() {
  std::lock_guard<std::mutex> lk(tmap_mutex());
  tmap_cache().clear();
}

// cached bf16 tensor map of any rank <= 5 (128B swizzle) for the other translation units
CUtensorMap get_tmap_bf16(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box) {
  return get_tmap(ptr, rank, dims, strides_bytes, box);
}

// 2D row-major bf16 matrix (rows x cols, leading dimension ld_elems), 128B-swizzled box.
CUtensorMap get_tmap_2d_bf16(const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_cols,
                             uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t str[1] = {ld_elems * 2};
  const uint32_t box[2] = {box_cols, box_rows};
  return get_tmap(ptr, 2, dims, str, box);
}

// Residual L2 prefetch switch: DEPTHPRO_RES_PREFETCH=0/1 in the environment, or gemm_tc_set_res_prefetch() (A/B runs).
static std::atomic<int> g_res_prefetch{-1};
void gemm_tc_set_res_prefetch(int on) {
  g_res_prefetch = on != 0;
  bump_config_epoch();
}
void gemm_tc_set_sm_limit(int sms) {
  g_sm_limit = sms;
  bump_config_epoch();
}
void gemm_tc_set_l2_persist(int mb) {
  set_l2_persist_impl(mb);
  bump_config_epoch();
}
static bool res_prefetch_enabled() {
  if (g_res_prefetch < 0) {
    const char* e = getenv("DEPTHPRO_RES_PREFETCH");
    g_res_prefetch = e ? (atoi(e) != 0) : RES_PREFETCH_DEFAULT;
  }
  return g_res_prefetch != 0;
}

void gemm_tc(const GemmOp& op_in, cudaStream_t stream) {
  GemmOp op = op_in;
  op.finish();
  DP_CHECK(op.K % BK == 0, "gemm_tc: K must be a multiple of 64");
  int bn = 0;
  if (op.N % 256 == 0) bn = 256;
  else if (op.N % 128 == 0) bn = 128;
  else if (op.N % 64 == 0) bn = 64;
  else if (op.N == 32) bn = 32;
  DP_CHECK(bn != 0, "gemm_tc: unsupported N");
  if (op.out_mode == O_DOT_RELU) DP_CHECK(op.N == 32, "O_DOT_RELU needs N == 32");
  if (op.out_mode == O_HEAD_FUSED) DP_CHECK(op.N == 128 && op.B == 1 && op.a_mode == A_CONV3X3, "bad fused head");
  if (op.out_mode == O_CONVT2X2) DP_CHECK(op.cout % 32 == 0 && op.N == 4 * op.cout, "bad ConvT shape");
  DP_CHECK(op.ngroups >= 1 && op.ngroups <= 3, "gemm_tc: 1..3 groups");

  TileGeom g{};
  g.reverse = op.reverse ? 1 : 0;
  g.n_tiles = op.N / bn;
  g.k_blocks = op.K / BK;
  // CTA pairs (cluster of 2 along M) share every weight tile through TMA multicast: the GEMMs are
  // bound by L2->SM bandwidth (48 KB per 128x256x64 k-block), the pair cuts that to 32 KB per CTA.
  int m_tiles_total = 0;
  if (op.a_mode == A_CONV3X3) m_tiles_total = op.B * (op.W / TILE_W) * (op.H / TILE_H);
  else
    for (int i = 0; i < op.ngroups; ++i) m_tiles_total += (op.grp[i].M + BM - 1) / BM;
  static const bool no_cluster = getenv("DEPTHPRO_NO_CLUSTER") != nullptr;  // debugging switch
  const int cl = (!no_cluster && bn >= 128 && m_tiles_total * g.n_tiles >= num_sms()) ? 2 : 1;

  CUtensorMap tmA_v;
  const CUtensorMap* tmA = &tmA_v;
  if (op.a_mode == A_CONV3X3) {
    DP_CHECK(op.ngroups == 1, "conv3x3 launches are not grouped");
    DP_CHECK(op.C % BK == 0 && op.K == 9 * op.C, "conv3x3: C must be a multiple of 64");
    DP_CHECK(op.W % TILE_W == 0 && op.H % TILE_H == 0, "conv3x3: H, W must be multiples of 8, 16");
    g.tiles_x = op.W / TILE_W;
    g.tiles_y = op.H / TILE_H;
    g.tiles_in_group[0] = m_tiles_total;
    g.m_units = (m_tiles_total + cl - 1) / cl;
    g.unit_start[0] = 0, g.unit_start[1] = g.unit_start[2] = g.unit_start[3] = g.m_units;
    const uint64_t dims[4] = {(uint64_t)op.C, (uint64_t)op.W, (uint64_t)op.H, (uint64_t)op.B};
    const uint64_t str[3] = {(uint64_t)op.C * 2, (uint64_t)op.W * op.C * 2, (uint64_t)op.H * op.W * op.C * 2};
    const uint32_t box[4] = {BK, TILE_W, TILE_H, 1};
    tmA_v = get_tmap(op.A, 4, dims, str, box);
  } else {
    int mu = 0;
    for (int i = 0; i < 3; ++i) {
      g.unit_start[i] = mu;
      g.tiles_in_group[i] = 0;
      if (i < op.ngroups) {
        g.tiles_in_group[i] = (op.grp[i].M + BM - 1) / BM;
        mu += (g.tiles_in_group[i] + cl - 1) / cl;
      }
    }
    g.unit_start[3] = mu;
    for (int i = op.ngroups; i < 3; ++i) g.unit_start[i] = mu;
    g.m_units = mu;
    // one map over every addressable row: a group's last tile may read rows of the next group
    // (rows are independent; the epilogue masks them), tiles past the end are zero-filled by TMA
    const uint64_t dims[2] = {(uint64_t)op.K, (uint64_t)op.a_rows};
    const uint64_t str[1] = {(uint64_t)op.lda * 2};
    const uint32_t box[2] = {BK, BM};
    tmA_v = get_tmap(op.A, 2, dims, str, box);
  }
  WeightMaps tmW;
  const uint64_t wd[2] = {(uint64_t)op.K, (uint64_t)op.N};
  const uint64_t ws[1] = {(uint64_t)op.K * 2};
  const uint32_t wb[2] = {BK, (uint32_t)(bn / cl)};  // each CTA of a cluster loads bn / cl weight rows
  for (int i = 0; i < 3; ++i) tmW.b[i] = get_tmap(op.grp[i < op.ngroups ? i : 0].Wt, 2, wd, ws, wb);
  // TMA-store epilogue for plain bf16 row-major outputs (no ReLU'd twin): rows leave the SM as full
  // 128-byte lines instead of 32 scattered 16-byte pieces per store instruction
  static const bool no_tma_out = getenv("DEPTHPRO_NO_TMA_STORE") != nullptr;  // debugging switch
  const bool rowmajor_ok = op.out_mode == O_ROWMAJOR && !op.out_f32 && op.out != nullptr;
  // ConvT: a 32-row slab must be 32 consecutive input pixels of ONE image row, a 64-column chunk must
  // stay inside one (dy, dx) parity
  const bool convt_ok = op.out_mode == O_CONVT2X2 && op.out != nullptr && op.W % 32 == 0 && op.cout % 64 == 0 &&
                        op.a_mode == A_ROWMAJOR && op.ngroups == 1 && op.M % 32 == 0;
  g.tma_out = (!no_tma_out && (rowmajor_ok || convt_ok) && bn >= 128 && op.ldo % 8 == 0 && (op.col_off % 64) == 0) ? 1 : 0;
  for (int i = 0; i < 3; ++i) tmW.o[i] = tmW.b[0];
  tmW.orelu = tmW.b[0];
  if (g.tma_out) {
    auto make = [&](const void* base, int group) -> CUtensorMap {
      if (op.out_mode == O_CONVT2X2) {
        // output (b, 2y+dy, 2x+dx, c) viewed as [c][dx][x][dy][b*H+y]
        const uint64_t ld = (uint64_t)op.ldo * 2;
        const uint64_t od[5] = {(uint64_t)op.ldo, 2, (uint64_t)op.W, 2, (uint64_t)op.B * op.H};
        const uint64_t os[4] = {ld, 2 * ld, 2 * (uint64_t)op.W * ld, 4 * (uint64_t)op.W * ld};
        const uint32_t ob[5] = {64, 1, 32, 1, 1};
        return get_tmap(base, 5, od, os, ob);
      }
      if (op.a_mode == A_CONV3X3) {
        const uint64_t od[4] = {(uint64_t)op.ldo, (uint64_t)op.W, (uint64_t)op.H, (uint64_t)op.B};
        const uint64_t os[3] = {(uint64_t)op.ldo * 2, (uint64_t)op.W * op.ldo * 2, (uint64_t)op.H * op.W * op.ldo * 2};
        const uint32_t ob[4] = {64, TILE_W, 2, 1};
        return get_tmap(base, 4, od, os, ob);
      }
      // rows past a group's end are out of bounds for ITS map, so a tail tile never touches the next group
      const uint64_t od[2] = {(uint64_t)op.ldo, (uint64_t)(op.grp[group].o_row_off + op.grp[group].M)};
      const uint64_t os[1] = {(uint64_t)op.ldo * 2};
      const uint32_t ob[2] = {64, 32};
      return get_tmap(base, 2, od, os, ob);
    };
    for (int i = 0; i < op.ngroups; ++i) tmW.o[i] = make(op.out, i);
    if (op.out_relu) tmW.orelu = make(op.out_relu, 0);
  }

  const bool resid32 = op.out_mode == O_ROWMAJOR && op.res != nullptr && op.res_f32 && op.out_f32 && op.res2 == nullptr &&
                       op.out_relu == nullptr && op.act == ACT_NONE && op.a_mode == A_ROWMAJOR && op.res == op.out &&
                       op.ldres == op.ldo && op.col_off == 0 && bn >= 128;
  if (op.out_mode == O_CONVT2X2) DP_CHECK(op.res == nullptr && op.res2 == nullptr, "ConvT epilogue has no residual");
  if (op.border_cb != nullptr)
    DP_CHECK(op.a_mode == A_CONV3X3 && g.tma_out && op.out_relu == nullptr && !(op.res && op.res_f32) && op.act == ACT_NONE,
             "border_cb: conv3x3 with the single-store TMA epilogue only");
  const bool tma_epi = g.tma_out && !(op.res && op.res_f32);
  const int epi = resid32 ? EPI_RES32 : (tma_epi ? (op.out_relu ? EPI_TMA2 : EPI_TMA) : EPI_MISC);
  g.tma_out = tma_epi ? 1 : 0;
  g.res_prefetch = (resid32 && res_prefetch_enabled() && op.ldres % 4 == 0 && op.N % 4 == 0 &&
                    reinterpret_cast<uintptr_t>(op.res) % 16 == 0) ? 1 : 0;
  if (op.ln_xlo != nullptr) {
    // residual stream as a (hi, lo) pair of 16-bit arrays, updated in place; hi doubles as the LayerNorm-folded operand
    DP_CHECK(bn == 256 && op.a_mode == A_ROWMAJOR && op.out_mode == O_ROWMAJOR && op.N == 1024 && op.ldo == 1024 &&
                 op.ln_xb != nullptr && op.ln_stats_out != nullptr && op.ln_stats == nullptr && op.res == nullptr &&
                 op.res2 == nullptr && op.out == nullptr && op.out_relu == nullptr && op.act == ACT_NONE && op.col_off == 0,
             "pair-residual producer: N = ldo = 1024, ln_xb + ln_xlo + ln_stats_out, no other output");
    g.tma_out = 0, g.res_prefetch = 0;
    if (cl == 2) launch<256, 2, EPI_RES16_LN>(op, g, *tmA, tmW, stream);
    else launch<256, 1, EPI_RES16_LN>(op, g, *tmA, tmW, stream);
    return;
  }
  if (op.ln_stats != nullptr || op.ln_xb != nullptr) {
    // LayerNorm-folded forms exist for the ViT shapes only (BN = 256)
    DP_CHECK(bn == 256 && op.a_mode == A_ROWMAJOR, "LN-folded GEMM: N must be a multiple of 256");
    if (op.ln_stats != nullptr) {
      DP_CHECK(epi == EPI_TMA && op.K == 1024 && op.res == nullptr && op.res2 == nullptr && op.gamma == nullptr,
               "LN-folded consumer: bf16 row-major output, K = 1024, no residual");
      for (int i = 0; i < op.ngroups; ++i)
        DP_CHECK(op.grp[i].ln_c != nullptr && op.grp[i].bias != nullptr && op.grp[i].gamma == nullptr, "LN-folded consumer: ln_c / bias per group");
      if (cl == 2) launch<256, 2, EPI_TMA_LN>(op, g, *tmA, tmW, stream);
      else launch<256, 1, EPI_TMA_LN>(op, g, *tmA, tmW, stream);
    } else {
      DP_CHECK(epi == EPI_RES32 && op.N == 1024 && op.ldo == 1024 && op.ln_stats_out != nullptr,
               "LN-folded producer: fp32 residual form with N = 1024");
      if (cl == 2) launch<256, 2, EPI_RES32_LN>(op, g, *tmA, tmW, stream);
      else launch<256, 1, EPI_RES32_LN>(op, g, *tmA, tmW, stream);
    }
    return;
  }
#define DP_LAUNCH(BN_, CL_)                                                          \
  do {                                                                               \
    if (epi == EPI_RES32) launch<BN_, CL_, EPI_RES32>(op, g, *tmA, tmW, stream);     \
    else if (epi == EPI_TMA) launch<BN_, CL_, EPI_TMA>(op, g, *tmA, tmW, stream);    \
    else if (epi == EPI_TMA2) launch<BN_, CL_, EPI_TMA2>(op, g, *tmA, tmW, stream);  \
    else launch<BN_, CL_, EPI_MISC>(op, g, *tmA, tmW, stream);                       \
  } while (0)
  if (cl == 2) {
    if (bn == 256) DP_LAUNCH(256, 2);
    else DP_LAUNCH(128, 2);
  } else if (bn == 256) DP_LAUNCH(256, 1);
  else if (bn == 128) DP_LAUNCH(128, 1);
  else if (bn == 64) launch<64, 1, EPI_MISC>(op, g, *tmA, tmW, stream);
  else launch<32, 1, EPI_MISC>(op, g, *tmA, tmW, stream);
#undef DP_LAUNCH
}

}  // namespace dp
