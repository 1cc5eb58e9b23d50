// Softmax attention of the DINOv2 ViT-L blocks: 16 heads x 64, 577 tokens, scale 1/8, no mask
// (timm Attention -> F.scaled_dot_product_attention; wired at
// src/depth_pro/network/vit_factory.py:97-110).  qkv rows are tokens, columns
// [q(16x64) | k(16x64) | v(16x64)]; output rows are tokens, columns head-major (16x64).
//
//   attention_f32   : parity mode, CUDA-core fp32, online softmax, thread = one query row.
//   attention_bf16  : flash-style, bf16 tensor-core MMA (m16n8k16, fp32 accumulate), 128 query
//                     rows per CTA, 64-key K/V tiles double-buffered with cp.async, 128B-row
//                     XOR-swizzled shared memory read with ldmatrix.
#include "attention.cuh"

namespace dp {
namespace {

constexpr int SEQ = 577, HD = 64, NH = 16, LDQ = 3 * NH * HD, LDO = NH * HD;

// ------------------------------------------------------------------------------- fp32
constexpr int F_QT = 128, F_KT = 32;

__global__ void __launch_bounds__(F_QT) attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out) {
  __shared__ __align__(16) float sK[F_KT][HD];
  __shared__ __align__(16) float sV[F_KT][HD];
  const int seq = blockIdx.x / NH, h = blockIdx.x % NH;
  const int qi = blockIdx.y * F_QT + threadIdx.x;
  const bool qok = qi < SEQ;
  const float* base = qkv + static_cast<long long>(seq) * SEQ * LDQ + h * HD;

  float q[HD], o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  if (qok) {
    const float4* qp = reinterpret_cast<const float4*>(base + static_cast<long long>(qi) * LDQ);
#pragma unroll
    for (int d = 0; d < HD / 4; ++d) {
      const float4 t = qp[d];
      q[4 * d] = t.x, q[4 * d + 1] = t.y, q[4 * d + 2] = t.z, q[4 * d + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] = 0.f;
  }
  float mx = -INFINITY, l = 0.f;

  for (int k0 = 0; k0 < SEQ; k0 += F_KT) {
    __syncthreads();
    for (int e = threadIdx.x; e < F_KT * HD / 4; e += F_QT) {
      const int r = e / (HD / 4), c = e % (HD / 4);
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (k0 + r < SEQ) {
        const float* row = base + static_cast<long long>(k0 + r) * LDQ;
        kk = reinterpret_cast<const float4*>(row + NH * HD)[c];
        vv = reinterpret_cast<const float4*>(row + 2 * NH * HD)[c];
      }
      reinterpret_cast<float4*>(&sK[r][0])[c] = kk;
      reinterpret_cast<float4*>(&sV[r][0])[c] = vv;
    }
    __syncthreads();
    const int nk = min(F_KT, SEQ - k0);
    float s[F_KT];
    float cmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < F_KT; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) a = fmaf(q[d], sK[j][d], a);
      s[j] = j < nk ? a * 0.125f : -INFINITY;
      cmax = fmaxf(cmax, s[j]);
    }
    const float mnew = fmaxf(mx, cmax);
    const float corr = expf(mx - mnew);
    l *= corr;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] *= corr;
#pragma unroll
    for (int j = 0; j < F_KT; ++j) {
      const float p = expf(s[j] - mnew);
      l += p;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] = fmaf(p, sV[j][d], o[d]);
    }
    mx = mnew;
  }
  if (qok) {
    const float inv = 1.f / l;
    float4* op = reinterpret_cast<float4*>(out + (static_cast<long long>(seq) * SEQ + qi) * LDO + h * HD);
#pragma unroll
    for (int d = 0; d < HD / 4; ++d)
      op[d] = make_float4(o[4 * d] * inv, o[4 * d + 1] * inv, o[4 * d + 2] * inv, o[4 * d + 3] * inv);
  }
}

// ------------------------------------------------------------------------------- bf16
constexpr int B_QT = 128, B_KT = 64, B_THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 16-byte chunk `c` (0..7) of 128-byte row `r`, XOR-swizzled
__device__ __forceinline__ uint32_t sw_off(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32." DP_MMA_SYNC_TYPE "." DP_MMA_SYNC_TYPE ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  bf16x2 h = f2_to_h2(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(B_THREADS) attention_bf16_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                       // 128 x 128 B
  uint8_t* sK = smem + B_QT * 128;          // 2 x 64 x 128 B
  uint8_t* sV = sK + 2 * B_KT * 128;        // 2 x 64 x 128 B
  const int seq = blockIdx.x / NH, h = blockIdx.x % NH;
  const int q0 = blockIdx.y * B_QT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* base = qkv + static_cast<long long>(seq) * SEQ * LDQ + h * HD;

  // ---- async loads: Q tile, then K/V tile 0
  for (int e = threadIdx.x; e < B_QT * 8; e += B_THREADS) {
    const int r = e >> 3, c = e & 7;
    const bool ok = q0 + r < SEQ;
    cp_async16(smem_u32(sQ) + sw_off(r, c), base + static_cast<long long>(ok ? q0 + r : 0) * LDQ + c * 8, ok);
  }
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * B_KT;
    for (int e = threadIdx.x; e < B_KT * 8; e += B_THREADS) {
      const int r = e >> 3, c = e & 7;
      const bool ok = k0 + r < SEQ;
      const bf16* row = base + static_cast<long long>(ok ? k0 + r : 0) * LDQ + c * 8;
      cp_async16(smem_u32(sK) + buf * B_KT * 128 + sw_off(r, c), row + NH * HD, ok);
      cp_async16(smem_u32(sV) + buf * B_KT * 128 + sw_off(r, c), row + 2 * NH * HD, ok);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  constexpr int NT = (SEQ + B_KT - 1) / B_KT;  // 10 key tiles
  const float sl2 = 0.125f * 1.4426950408889634f;  // softmax scale * log2(e)
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  uint32_t qf[4][4];

  for (int t = 0; t < NT; ++t) {
    const int buf = t & 1;
    if (t + 1 < NT) {
      load_kv(t + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
      // Q fragments for this warp's 16 rows: 4 k-steps of 16
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = ks * 2 + (lane >> 4);
        ldsm_x4(smem_u32(sQ) + sw_off(r, c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    // ---- S = Q K^T (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const int r = nt * 8 + (lane & 7);
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {  // two ldmatrix.x4, each covers k = kh*32 .. +32
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(sK) + buf * B_KT * 128 + sw_off(r, kh * 4 + (lane >> 3)), b0, b1, b2, b3);
        mma_bf16(s[nt], qf[kh * 2], b0, b1);
        mma_bf16(s[nt], qf[kh * 2 + 1], b2, b3);
      }
    }
    // ---- mask keys beyond the sequence (last tile only)
    if (t == NT - 1) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = t * B_KT + nt * 8 + (lane & 3) * 2;
        if (key >= SEQ) s[nt][0] = s[nt][2] = -INFINITY;
        if (key + 1 >= SEQ) s[nt][1] = s[nt][3] = -INFINITY;
      }
    }
    // ---- online softmax (rows g and g+8 of this warp's 16)
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = exp2f((m0 - mx0) * sl2), c1 = exp2f((m1 - mx1) * sl2);
    m0 = mx0, m1 = mx1;
    const float ms0 = mx0 * sl2, ms1 = mx1 * sl2;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(fmaf(s[nt][0], sl2, -ms0)), p1 = exp2f(fmaf(s[nt][1], sl2, -ms0));
      const float p2 = exp2f(fmaf(s[nt][2], sl2, -ms1)), p3 = exp2f(fmaf(s[nt][3], sl2, -ms1));
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2] = pack_bf16(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      o[dt][0] *= c0, o[dt][1] *= c0;
      o[dt][2] *= c1, o[dt][3] *= c1;
    }
    // ---- O += P V : k = keys (4 steps of 16), n = head dim (8 tiles of 8)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int r = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int dp2 = 0; dp2 < 4; ++dp2) {  // two d-tiles per ldmatrix.x4.trans
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_u32(sV) + buf * B_KT * 128 + sw_off(r, dp2 * 2 + (lane >> 4)), b0, b1, b2, b3);
        mma_bf16(o[dp2 * 2], pf[ks], b0, b1);
        mma_bf16(o[dp2 * 2 + 1], pf[ks], b2, b3);
      }
    }
    __syncthreads();  // all warps done with `buf` before it is refilled at t+2
  }

  // ---- finalise: divide by the row sums (quad-reduced) and store bf16
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  bf16* obase = out + static_cast<long long>(seq) * SEQ * LDO + h * HD + (lane & 3) * 2;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    if (r0 < SEQ)
      *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(r0) * LDO + dt * 8) = pack_bf16(o[dt][0] * i0, o[dt][1] * i0);
    if (r1 < SEQ)
      *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(r1) * LDO + dt * 8) = pack_bf16(o[dt][2] * i1, o[dt][3] * i1);
  }
}

}  // namespace

void attention_f32(const float* qkv, float* out, int nseq, cudaStream_t s) {
  dim3 grid(nseq * NH, (SEQ + F_QT - 1) / F_QT);
  attention_f32_kernel<<<grid, F_QT, 0, s>>>(qkv, out);
  DP_LAUNCH_CHECK();
}

void attention_bf16(const bf16* qkv, bf16* out, int nseq, cudaStream_t s) {
  constexpr int SMEM = B_QT * 128 + 4 * B_KT * 128;  // 48 KB
  static std::atomic<unsigned long long> configured{0};
  if (first_use_on_device(configured)) {
    DP_CUDA(cudaFuncSetAttribute(attention_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  }
  dim3 grid(nseq * NH, (SEQ + B_QT - 1) / B_QT);
  attention_bf16_kernel<<<grid, B_THREADS, SMEM, s>>>(qkv, out);
  DP_LAUNCH_CHECK();
}

}  // namespace dp
