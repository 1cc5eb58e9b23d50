// Shared declarations of the Depth Pro B200 engine (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <stdexcept>
#include <string>

namespace dp {

// The engine's 16-bit storage type.  The library is built TWICE from the same sources (build.py): the default flavour
// stores activations / weights as bfloat16 (libdepthpro_b200.so), -DDP_ACT_FP16 as IEEE half (libdepthpro_b200_fp16.so,
// what `precision=torch.half` selects: the reference's model.half(), depth_pro.py:122-123).  Accumulation, the ViT
// residual stream, LayerNorm / softmax statistics and the metric-depth epilogue are fp32 in both.  The type keeps the name
// `bf16` ("the 16-bit activation type") throughout the sources; everything flavour-specific is in this block: the two
// conversions, the tcgen05 operand-format bits, the mma.sync operand type and the TMA element type.
#ifdef DP_ACT_FP16
typedef __half bf16;
typedef __half2 bf16x2;
#define DP_ACT_NAME "fp16"
#define DP_MMA_SYNC_TYPE "f16"
#define DP_UMMA_AB_FORMAT 0u  // tcgen05 kind::f16 instruction descriptor: A / B format 0 = F16
#define DP_TMAP_ELEM CU_TENSOR_MAP_DATA_TYPE_FLOAT16
// float -> half conversions SATURATE to +-65504 on the device (F2FP.SATFINITE, one instruction like the plain form): fp16
// has no headroom above 65504, and an activation outlier must not turn into inf and poison a whole row of a GEMM.
__host__ __device__ __forceinline__ bf16x2 f2_to_h2(float a, float b) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));  // d = {hi: %1, lo: %2}
  return *reinterpret_cast<bf16x2*>(&r);
#else
  return __floats2half2_rn(a, b);
#endif
}
__host__ __device__ __forceinline__ float2 h2_to_f2(bf16x2 v) { return __half22float2(v); }
__host__ __device__ __forceinline__ bf16 f_to_h(float v) {
#ifdef __CUDA_ARCH__
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return *reinterpret_cast<bf16*>(&r);
#else
  return __float2half_rn(v);
#endif
}
__host__ __device__ __forceinline__ float h_to_f(bf16 v) { return __half2float(v); }
#else
typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf16x2;
#define DP_ACT_NAME "bf16"
#define DP_MMA_SYNC_TYPE "bf16"
#define DP_UMMA_AB_FORMAT 1u  // A / B format 1 = BF16
#define DP_TMAP_ELEM CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
__host__ __device__ __forceinline__ bf16x2 f2_to_h2(float a, float b) { return __floats2bfloat162_rn(a, b); }
__host__ __device__ __forceinline__ float2 h2_to_f2(bf16x2 v) { return __bfloat1622float2(v); }
__host__ __device__ __forceinline__ bf16 f_to_h(float v) { return __float2bfloat16_rn(v); }
__host__ __device__ __forceinline__ float h_to_f(bf16 v) { return __bfloat162float(v); }
#endif

struct Error : std::runtime_error {
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};

#define DP_CHECK(cond, msg)                                                        \
  do {                                                                             \
    if (!(cond)) throw ::dp::Error(std::string(msg) + " [" #cond "] at " __FILE__ ":" + \
                                   std::to_string(__LINE__));                      \
  } while (0)

#define DP_CUDA(expr)                                                              \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess)                                                         \
      throw ::dp::Error(std::string("CUDA error: ") + cudaGetErrorString(_e) +     \
                        " in " #expr " at " __FILE__ ":" + std::to_string(__LINE__)); \
  } while (0)

// Launch counter (reported by dp_launch_count, used for bench.py's gpu_launches): process-wide, atomic -- engines on
// several host threads (one per GPU, or several per GPU) may launch concurrently.
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Bumped by every process-wide A/B switch (attention variant, residual prefetch, L2 set-aside, HBM kernel generation):
// an engine re-captures its CUDA graphs when the epoch it captured under is no longer current.
extern std::atomic<unsigned> g_config_epoch;
inline void bump_config_epoch() { g_config_epoch.fetch_add(1); }

// NVTX ranges per stage of a frame (SURVEY.md §5 "tracing"): visible in Nsight Systems / ncu --nvtx, free when no tool
// is attached (nvtx3 is header-only and resolves its injection library lazily).
struct NvtxRange {
  explicit NvtxRange(const char* name);
  ~NvtxRange();
};

// ---- optional per-launch profiling with CUDA events (bench.py roofline numbers) ----------
enum KernelClass {
  KC_GEMM_TC = 0,   // tcgen05 GEMM (dense contraction), work = flops
  KC_CONV_TC = 1,   // tcgen05 implicit-GEMM 3x3 conv, work = flops
  KC_ATTENTION = 2, // work = flops (4*N^2*d per head)
  KC_LAYERNORM = 3, // work = bytes
  KC_GEMM_SIMT = 4, // fp32 CUDA-core GEMM / conv, work = flops
  KC_COUNT = 5
};
void prof_begin(cudaStream_t s, int cls, double work);  // no-ops unless profiling is enabled
void prof_end(cudaStream_t s);
struct ProfScope {
  cudaStream_t s;
  ProfScope(cudaStream_t s_, int cls, double work) : s(s_) { prof_begin(s, cls, work); }
  ~ProfScope() { prof_end(s); }
};
void prof_enable(bool on);
bool prof_active();
// sums since prof_enable(true); synchronises the device
void prof_collect(double* ms_by_class, double* work_by_class, long long* launches_by_class);

// Programmatic dependent launch (PDL): a kernel launched with this attribute may start its CTAs --
// barrier init, TMEM allocation, tensor-map prefetch -- while the previous kernel of the stream is
// still draining; it executes `griddepcontrol.wait` before it touches any global memory the
// previous kernel may have written.  DEPTHPRO_PDL=0 switches it off (without the attribute the
// griddepcontrol instructions are no-ops).  Round 1 measured no gain on eagerly launched frames under the power cap;
// with the frame replayed as a CUDA graph (programmatic edges between consecutive kernel nodes) it is worth +0.7 %
// frames/s and became the default late in round 2.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  DP_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

// cudaFuncSetAttribute (opt-in dynamic shared memory) applies to the CURRENT device only.  One process per GPU is the
// intended deployment, but two engines on two GPUs in one process must work too: every launcher keeps a bit per device.
inline bool first_use_on_device(std::atomic<unsigned long long>& mask) {
  int dev = 0;
  DP_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  return (mask.fetch_or(bit) & bit) == 0;
}

#define DP_LAUNCH_CHECK()                \
  do {                                   \
    ::dp::count_launch();                \
    DP_CUDA(cudaGetLastError());         \
  } while (0)

enum Prec { FP32 = 0, BF16 = 1 };
constexpr int LN_SLOTS = 8;  // partial (sum, sum of squares) pairs per row: 1024 columns / 128 per epilogue warp
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

// How the A operand rows of a GEMM are addressed.
enum AMode {
  A_ROWMAJOR = 0,  // A[m, k] at A + m*lda + k
  A_CONV3X3 = 1,   // implicit 3x3 / pad 1 / stride 1 over NHWC (B,H,W,C); k = (ky*3+kx)*C + c
};

// Where a GEMM output element (m, n) goes.
enum OutMode {
  O_ROWMAJOR = 0,    // out[m*ldo + col_off + n]
  O_CONVT2X2 = 1,    // ConvTranspose2d k2 s2: n = (dy*2+dx)*Cout + co, m = (b,y,x) on an HxW grid
                     //   -> out[((b*2H + 2y+dy)*2W + 2x+dx)*ldo + col_off + co]
  O_PATCH_EMBED = 2, // m = patch*576 + p -> row patch*577 + 1 + p, plus pos_embed[1+p][n]
  O_DOT_RELU = 3,    // head.2 + ReLU + head.4 (1x1, 32->1) + ReLU: out[m] (fp32), N must be 32
  O_HEAD_FUSED = 4,  // head.1 (ConvT 2x2) o head.2 (conv3x3) pre-composed into one 3x3 conv over the
                     //   768^2 map with N = 4 parities x 32; per parity: + border-aware bias, ReLU,
                     //   head.4 dot (32->1), ReLU -> out[(2y+py)*2W + 2x+px] (fp32)
};

// One GEMM / implicit-GEMM convolution launch, shared by the fp32 SIMT and bf16 tcgen05 cores.
// Activations are `prec`-typed (float or bf16) unless stated; weights Wt are [N, K] K-major.
//
// Grouped launches: up to 3 row-groups stacked along M share one launch but use different
// weights (the patch / image / fov ViT-L encoders run layer by layer as ONE GEMM each).  Group g
// covers A rows [a_row_off, a_row_off + M) and output rows [o_row_off, o_row_off + M); tiles never
// straddle a group.  With ngroups == 1 the group fields mirror the scalar fields below.
struct GemmGroup {
  int M = 0;
  long long a_row_off = 0, o_row_off = 0;
  const void* Wt = nullptr;
  const float* bias = nullptr;
  const float* gamma = nullptr;
  const float* pos = nullptr;
  const float* ln_c = nullptr;  // LayerNorm-folded GEMM: column sums of the (g * W) weights, see GemmOp::ln_stats
};

struct GemmOp {
  int M = 0, N = 0, K = 0;
  // A
  const void* A = nullptr;
  int a_mode = A_ROWMAJOR;
  int lda = 0;
  long long a_rows = 0;             // rows addressable behind A (grouped launches); 0 -> M
  int B = 1, H = 0, W = 0, C = 0;  // conv geometry (input grid); also the grid for O_CONVT2X2
  // W
  const void* Wt = nullptr;
  // epilogue:  v = acc + bias[n]; v = act(v); v = v*gamma[n]; v += res; v += res2
  const float* bias = nullptr;
  int bias_mod = 0;             // if > 0, bias index is n % bias_mod (ConvT)
  const float* gamma = nullptr;
  int act = ACT_NONE;
  const void* res = nullptr;    // residual, same (row, n) addressing as out in O_ROWMAJOR
  int res_f32 = 0;              // residual dtype: 1 = float, 0 = activation dtype
  const void* res2 = nullptr;   // second addend (activation dtype)
  int ldres = 0;
  // output
  void* out = nullptr;
  int out_f32 = 0;              // 1 = float output regardless of prec
  void* out_relu = nullptr;     // optional second store of relu(v) (activation dtype)
  int out_mode = O_ROWMAJOR;
  int ldo = 0, col_off = 0;
  int cout = 0;                 // O_CONVT2X2: Cout
  const float* pos = nullptr;   // O_PATCH_EMBED: pos_embed (577, N) fp32
  const float* dot_w = nullptr; // O_DOT_RELU: head.4 weight (32), dot_b: bias (1)
  const float* dot_b = nullptr;
  const float* head_cb = nullptr;  // O_HEAD_FUSED: [9][32] per-tap bias terms, then [32] full bias
  // conv3x3 with a composed 1x1 in front (bf16 TMA-store epilogue): [9][N] per-tap share of the 1x1's bias; `bias` holds
  // the interior value (all nine included), pixels on the image border subtract the taps that fall into the zero padding
  const float* border_cb = nullptr;
  // LayerNorm folded into the surrounding GEMMs (bf16 tcgen05 core only).  With x the fp32 residual
  // stream, LN(x) W^T + b  =  rstd * (x (g*W)^T)  -  rstd * mean * colsum(g*W)  +  (W b_ln + b):
  //  * producer (the fp32-residual form, proj / fc2, N == 1024): besides x += gamma * (acc + bias) it
  //    stores bf16(x) to `ln_xb` (ld = ldo) and per-row partial sums (sum x, sum x^2) of its 128-column
  //    slice to `ln_stats_out[row][LN_SLOTS][2]` -- no atomics, the consumer adds the slots in order;
  //  * consumer (bf16 output through the TMA-store epilogue, qkv / fc1, K == 1024): A is that raw bf16
  //    copy, Wt = bf16(g * W), grp.bias = W b_ln + b, grp.ln_c = colsum(Wt); `ln_stats` are the
  //    producer's partial sums of the A rows.
  const float* ln_stats = nullptr;
  void* ln_xb = nullptr;
  float* ln_stats_out = nullptr;
  //  * pair-residual producer (`ln_xlo` set; res / out unused): the stream itself lives in 16-bit storage as
  //    x = hi + lo, hi = round16(x) in `ln_xb` (the consumer's operand, as above), lo = round16(x - hi) in `ln_xlo`,
  //    both updated in place -- 8 B of traffic per element instead of 10.
  void* ln_xlo = nullptr;
  // Tile order.  The L2 (126 MB) still holds the END of what the previous kernel wrote: a consumer that walks its
  // m-units from the last to the first finds its first operands there instead of in HBM (Engine::run_vits alternates
  // the direction from kernel to kernel).  Results do not depend on the order.
  int reverse = 0;
  // groups
  int ngroups = 1;
  GemmGroup grp[3];

  // fill grp[0] from the scalar fields (single-group launches)
  void finish() {
    if (ngroups == 1 && grp[0].M == 0) {
      grp[0].M = M, grp[0].Wt = Wt, grp[0].bias = bias, grp[0].gamma = gamma, grp[0].pos = pos;
    }
    if (a_rows == 0) a_rows = M;
  }
};

}  // namespace dp
