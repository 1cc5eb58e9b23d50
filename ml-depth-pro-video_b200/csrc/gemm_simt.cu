// fp32 GEMM / implicit-GEMM 3x3 convolution on the CUDA cores — the parity-mode core.
// True fp32 FMA (no TF32): the fp32 mode has to match the CPU reference to 1e-4 through
// 24 transformer blocks + ~40 convolutions.  128x128x16 tiles, 256 threads, 8x8 micro-tiles,
// register-prefetched double buffering.  Not the performance path (that is gemm_tc.cu).
#include "common.cuh"
#include "gemm.cuh"

namespace dp {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

struct GroupTiles {
  int mt_start[4];
};

__global__ void __launch_bounds__(NT) gemm_simt_kernel(const GemmOp op, const GroupTiles gt) {
  __shared__ __align__(16) float sA[2][BK][BM + 4];
  __shared__ __align__(16) float sB[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int mt = blockIdx.y;
  const int gi = (mt >= gt.mt_start[1] ? 1 : 0) + (mt >= gt.mt_start[2] ? 1 : 0);
  const GemmGroup& gp = op.grp[gi];
  const long long m0 = static_cast<long long>(mt - gt.mt_start[gi]) * BM;  // row inside the group
  const int Mg = gp.M;
  const int n0 = blockIdx.x * BN;
  const float* __restrict__ A = reinterpret_cast<const float*>(op.A);
  const float* __restrict__ Wt = reinterpret_cast<const float*>(gp.Wt);

  // loader mapping: thread -> (row lr, 8 consecutive k starting at lk)
  const int lr = tid & 127, lk = (tid >> 7) * 8;
  const long long am = gp.a_row_off + m0 + lr;
  const bool a_ok = m0 + lr < Mg;
  int py = 0, px = 0;
  long long pb = 0;
  if (op.a_mode == A_CONV3X3 && a_ok) {
    px = static_cast<int>(am % op.W);
    const long long by = am / op.W;
    py = static_cast<int>(by % op.H);
    pb = by / op.H;
  }
  const int bn = n0 + lr;
  const bool b_ok = bn < op.N;

  float4 ra[2], rb[2];
  auto fetch = [&](int k0) {
    ra[0] = ra[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    rb[0] = rb[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a_ok) {
      const float* src = nullptr;
      if (op.a_mode == A_CONV3X3) {
        const int k = k0 + lk;
        const int tap = k / op.C, c = k - tap * op.C;
        const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
        if (yy >= 0 && yy < op.H && xx >= 0 && xx < op.W)
          src = A + ((pb * op.H + yy) * op.W + xx) * op.C + c;
      } else {
        src = A + am * op.lda + k0 + lk;
      }
      if (src) {
        ra[0] = *reinterpret_cast<const float4*>(src);
        ra[1] = *reinterpret_cast<const float4*>(src + 4);
      }
    }
    if (b_ok) {
      const float* src = Wt + static_cast<long long>(bn) * op.K + k0 + lk;
      rb[0] = *reinterpret_cast<const float4*>(src);
      rb[1] = *reinterpret_cast<const float4*>(src + 4);
    }
  };
  auto stash = [&](int buf) {
    const float a[8] = {ra[0].x, ra[0].y, ra[0].z, ra[0].w, ra[1].x, ra[1].y, ra[1].z, ra[1].w};
    const float b[8] = {rb[0].x, rb[0].y, rb[0].z, rb[0].w, rb[1].x, rb[1].y, rb[1].z, rb[1].w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sA[buf][lk + i][lr] = a[i];
      sB[buf][lk + i][lr] = b[i];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  fetch(0);
  stash(0);
  __syncthreads();
  const int nk = op.K / BK;
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) fetch((kb + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sA[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sA[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sB[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sB[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue: rows {ty*4+i, 64+ty*4+i}, cols {tx*4+j, 64+tx*4+j}
  float* out = reinterpret_cast<float*>(op.out);
  float* out_relu = reinterpret_cast<float*>(op.out_relu);
  const float* res = reinterpret_cast<const float*>(op.res);
  const float* res2 = reinterpret_cast<const float*>(op.res2);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long ml = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    const bool mok = ml < Mg;
    const long long m = gp.o_row_off + ml;
    float dot = 0.f;
    long long orow = 0;
    int p = 0;
    if (mok && op.out_mode == O_CONVT2X2) {
      // decode once per row; (dy,dx) added per column below
      orow = m;
    } else if (mok && op.out_mode == O_PATCH_EMBED) {
      const long long patch = m / 576;
      p = static_cast<int>(m - patch * 576);
      orow = patch * 577 + 1 + p;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (!mok || n >= op.N) continue;
      float v = acc[i][j];
      if (gp.bias) v += gp.bias[op.bias_mod ? n % op.bias_mod : n];
      if (op.act == ACT_RELU) v = fmaxf(v, 0.f);
      else if (op.act == ACT_GELU) v = gelu_erf(v);
      if (gp.gamma) v *= gp.gamma[n];
      long long off;
      if (op.out_mode == O_ROWMAJOR) {
        off = m * op.ldo + op.col_off + n;
      } else if (op.out_mode == O_CONVT2X2) {
        const int q = n / op.cout, co = n - q * op.cout;
        const int x = static_cast<int>(orow % op.W);
        const long long by = orow / op.W;
        const int y = static_cast<int>(by % op.H);
        const long long b = by / op.H;
        off = ((b * 2 * op.H + 2 * y + (q >> 1)) * (2LL * op.W) + 2 * x + (q & 1)) * op.ldo + op.col_off + co;
      } else if (op.out_mode == O_PATCH_EMBED) {
        v += gp.pos[(1 + p) * static_cast<long long>(op.N) + n];
        off = orow * op.ldo + n;
      } else {
        dot = fmaf(v, op.dot_w[n], dot);
        continue;
      }
      if (res) v += res[m * op.ldres + n];
      if (res2) v += res2[m * op.ldres + n];
      if (out) out[off] = v;
      if (out_relu) out_relu[off] = fmaxf(v, 0.f);
    }
    if (op.out_mode == O_DOT_RELU) {
      // N == 32: columns live in lanes tx = 0..7 (4 each) of this half-warp
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      if (mok && tx == 0) out[m] = fmaxf(dot + op.dot_b[0], 0.f);
    }
  }
}

}  // namespace

void gemm_simt(const GemmOp& op_in, cudaStream_t stream) {
  GemmOp op = op_in;
  op.finish();
  DP_CHECK(op.K % BK == 0, "gemm_simt: K must be a multiple of 16");
  if (op.a_mode == A_CONV3X3) DP_CHECK(op.C % BK == 0 && op.K == 9 * op.C && op.ngroups == 1, "gemm_simt: bad conv shape");
  else DP_CHECK(op.lda % 4 == 0, "gemm_simt: lda must be a multiple of 4");
  if (op.out_mode == O_DOT_RELU) DP_CHECK(op.N == 32, "O_DOT_RELU needs N == 32");
  GroupTiles gt{};
  int mt = 0;
  for (int i = 0; i < 3; ++i) {
    gt.mt_start[i] = mt;
    if (i < op.ngroups) mt += (op.grp[i].M + BM - 1) / BM;
  }
  for (int i = op.ngroups; i < 4; ++i) gt.mt_start[i] = mt;
  dim3 grid((op.N + BN - 1) / BN, mt);
  gemm_simt_kernel<<<grid, NT, 0, stream>>>(op, gt);
  DP_LAUNCH_CHECK();
}

}  // namespace dp
