/*
 * depthpro_b200.h — C-ABI of the B200-native Depth Pro inference engine.
 *
 * The reference (tdj28/ml-depth-pro-video) has no FFI layer: its hot path is reached only
 * through Python (`depth_pro.create_model_and_transforms`, `DepthPro.infer`, `depth_to_3d`).
 * This header is the boundary a binding for that path targets; the drop-in Python package
 * `ml-depth-pro-video_b200/depth_pro` calls it through ctypes.  Every entry point names the
 * reference interface it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - plain C types only; pointers are device pointers unless the name says `host`;
 *   - every function returns 0 on success, non-zero on error; `dp_last_error()` returns the
 *     message of the calling thread's last failure (Python shim raises RuntimeError);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), no hidden syncs
 *     except in create / finalize / destroy and the `*_host` helpers;
 *   - one engine per GPU; a handle is not thread-safe; the engine owns weights + workspace,
 *     the caller owns inputs and outputs;
 *   - there is no CPU fallback: on a machine without an sm_100 device `dp_engine_create`
 *     fails.
 */
#ifndef DEPTHPRO_B200_H
#define DEPTHPRO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dp_engine dp_engine;

/* precision of the compute path */
enum { DP_PREC_FP32 = 0, DP_PREC_BF16 = 1 };
/* source image formats accepted by dp_preprocess / dp_infer */
enum {
  DP_SRC_F32_CHW = 0, /* float32 (B,3,H,W) already normalised to [-1,1]  (output of the
                         reference `transform`, src/depth_pro/depth_pro.py:125-132) */
  DP_SRC_U8_HWC = 1   /* uint8 (B,H,W,3) as returned by `load_rgb` (src/depth_pro/utils.py:47);
                         ToTensor + Normalize(0.5,0.5) is fused into the resize kernel */
};

const char* dp_last_error(void);
int dp_version(void);
/* The library exists in two flavours built from the same sources: libdepthpro_b200.so stores 16-bit activations and
 * weights as bfloat16 ("bf16"), libdepthpro_b200_fp16.so as IEEE half ("fp16": what the reference's model.half() means,
 * src/depth_pro/depth_pro.py:122-123).  DP_PREC_BF16 below selects "the 16-bit mode" of whichever flavour is loaded;
 * accumulation, the ViT residual stream and all statistics are fp32 in both. */
const char* dp_act_dtype(void);

/* src/depth_pro/depth_pro.py:72-123  create_model_and_transforms (model construction). */
int dp_engine_create(int device, int precision, int max_batch, dp_engine** out);
/* Non-default configurations of src/depth_pro/depth_pro.py:26-36, 100-108 (DepthProConfig.use_fov_head /
 * fov_encoder_preset) and src/depth_pro/network/fov.py:29-56.  fov_mode 2 = the default (FOV head with its own ViT-L
 * encoder; dp_engine_create), 1 = FOV head WITHOUT encoder (fov_encoder_preset=None: state_dict keys fov.head.{0,2,4,6},
 * four convs on the low-resolution decoder feature), 0 = use_fov_head=False (no fov.* keys; dp_forward leaves fov_deg
 * untouched -- it may be NULL -- and dp_infer fails unless f_px_host is given, where the reference fails on
 * fov_deg=None, depth_pro.py:282-283). */
enum { DP_FOV_NONE = 0, DP_FOV_HEAD_ONLY = 1, DP_FOV_ENCODER = 2 };
int dp_engine_create_ex(int device, int precision, int max_batch, int fov_mode, dp_engine** out);
int dp_engine_destroy(dp_engine* e);

/* src/depth_pro/depth_pro.py:134-149  (load_state_dict): hand over ONE tensor of the
 * reference-format state_dict (fp32, contiguous, reference key name and shape).  `on_device`
 * != 0 means `data` is a device pointer on the engine's GPU, else host memory.  The engine
 * repacks into its kernel-native layout (bf16 / fp32, K-major, conv taps outermost). */
int dp_engine_set_weight(dp_engine* e, const char* name, const void* data,
                         const int64_t* shape, int ndim, int on_device);
/* bf16 engines fold each ViT block's norm1 / norm2 into attn.qkv / mlp.fc1 at finalize and drop the
 * originals: to change any of those tensors later, hand over the block's norm and qkv / fc1 weight + bias
 * again before the next dp_engine_finalize (the Python shim always re-sends every parameter). */
/* Number of reference tensors still missing (0 = complete). */
int dp_engine_missing_weights(dp_engine* e);
/* Allocate workspace, build TMA descriptors.  Must follow the last set_weight. */
int dp_engine_finalize(dp_engine* e);

/* src/depth_pro/depth_pro.py:125-132 (transform) + :273-279 (F.interpolate to 1536^2,
 * bilinear, align_corners=False).  img -> x_1536 float32 (B,3,1536,1536). */
int dp_preprocess(dp_engine* e, const void* img, int B, int H, int W, int src_fmt,
                  float* x_1536, void* stream);

/* The same with DepthPro.infer's `interpolation_mode` (depth_pro.py:247, 273-279): DP_INTERP_BILINEAR or
 * DP_INTERP_BICUBIC (ATen upsample_bicubic2d, align_corners=False, A = -0.75) -- the only two modes F.interpolate accepts
 * with align_corners=False on 4-D input; the reference raises ValueError for every other mode, and so does the shim. */
enum { DP_INTERP_BILINEAR = 0, DP_INTERP_BICUBIC = 1 };
int dp_preprocess_ex(dp_engine* e, const void* img, int B, int H, int W, int src_fmt, int interp_mode,
                     float* x_1536, void* stream);

/* src/depth_pro/network/encoder.py:151-188, 253-263  (_create_pyramid + split + cat):
 * x_1536 (B,3,1536,1536) -> patches float32 (35*B,3,384,384) in the REFERENCE order
 * (25*B level-0 patches, then 9*B level-1, then B level-2; patch-major, batch-minor). */
int dp_split(dp_engine* e, const float* x_1536, int B, float* patches, void* stream);

/* src/depth_pro/network/encoder.py:190-231 (reshape_feature + merge): tokens float32
 * (steps*steps*B, 577, C) in reference order -> merged float32 (B, C, S, S), S = 96 for
 * steps 5 / padding 3, 48 for steps 3 / padding 6. */
int dp_merge(dp_engine* e, const float* tokens, int B, int steps, int padding, int C,
             float* merged, void* stream);

/* src/depth_pro/depth_pro.py:218-241  DepthPro.forward: x (B,3,1536,1536) ->
 * canonical inverse depth (B,1,1536,1536) and fov_deg (B). */
int dp_forward(dp_engine* e, const float* x_1536, int B, float* canon_inv_depth,
               float* fov_deg, void* stream);

/* src/depth_pro/depth_pro.py:243-298  DepthPro.infer.  `f_px_host` (host, B floats) may be
 * NULL -> focal length estimated from the FOV head.  depth_out float32 (B,H,W), f_px_out
 * float32 (B) on device. */
int dp_infer(dp_engine* e, const void* img, int B, int H, int W, int src_fmt,
             const float* f_px_host, float* depth_out, float* f_px_out, void* stream);

/* dp_infer with `interpolation_mode` for BOTH resizes (depth_pro.py:273-279 and :288-291). */
int dp_infer_ex(dp_engine* e, const void* img, int B, int H, int W, int src_fmt, int interp_mode,
                const float* f_px_host, float* depth_out, float* f_px_out, void* stream);

/* Same call with HOST buffers: H2D copy, dp_infer, D2H copy, stream sync.  This is the
 * call generate_depth_maps.py:113-121 (transform -> infer -> .cpu().numpy()) amounts to. */
int dp_infer_host(dp_engine* e, const void* img_host, int B, int H, int W, int src_fmt,
                  const float* f_px_host, float* depth_out_host, float* f_px_out_host);

/* img_to_normalized_pointcloud.py:819-856  depth_to_3d (+ colours, :1226).  Row-major stream
 * compaction by valid = !isnan(d) && d > 0.  xyz float32 (N,3); rgb_out float32 (N,3) =
 * rgb/255 if rgb != NULL; valid_mask uint8 (H,W) optional; n_valid int64 on device.
 * `f_px_dev` is a device pointer to the focal length (so it chains after dp_infer). */
int dp_unproject(dp_engine* e, const float* depth, const uint8_t* rgb, int H, int W,
                 const float* f_px_dev, float* xyz, float* rgb_out, uint8_t* valid_mask,
                 int64_t* n_valid, void* stream);

/* generate_depth_maps.py:15-44, 128-143  colorize_depth / 16-bit export.  `lut` = 256x3 uint8
 * colour table (device) -> rgb_out uint8 (H,W,3); if lut == NULL writes uint16 (H,W)
 * normalised depth to `out` instead. */
int dp_colorize(dp_engine* e, const float* depth, int H, int W, const uint8_t* lut,
                void* out, void* stream);
/* colorize_depth(depth, min_depth, max_depth, cmap) with caller-supplied range ends (generate_depth_maps.py:15-31):
 * NaN = "None" = the image's own nanmin / nanmax.  With both ends given the reduction pass is skipped (one launch). */
int dp_colorize_range(dp_engine* e, const float* depth, int H, int W, const uint8_t* lut,
                      void* out, float min_depth, float max_depth, void* stream);

/* img_to_normalized_pointcloud.py:880-975  normalize_point_cloud_to_ground, on the points dp_unproject
 * produced: xyz float32 (n,3) on the device, updated IN PLACE.  `normal3` (host, 3 doubles) and `d` are the
 * fitted ground plane a x + b y + c z + d = 0 (the fit itself, :376-816, stays on the CPU).  Distances to the
 * plane, Rodrigues rotation of the normal onto +y, shift so that the 2nd percentile of the near-plane
 * heights is y = 0, ground points below 0 -> 0, other points below -0.1 -> -0.1; arithmetic in double,
 * np.percentile reproduced exactly (radix select).  `counters` (device, 6 x uint64, may be NULL) receives
 * [0] points within 5 cm of the plane, [1] points set to y = 0, [2] points limited to -0.1. */
int dp_ground_normalize(dp_engine* e, float* xyz, int64_t n, const double* normal3, double d,
                        uint64_t* counters, void* stream);
/* img_to_normalized_pointcloud.py:977-1118  grid_based_ground_adjustment: grid_size x grid_size cells over
 * the XZ bounding box; every cell with >= 10 points and >= 5 points below 0.2 whose `percentile`-th
 * percentile of those low heights exceeds 0.01 is lowered by it (fully below 0.1, linearly fading to 0 at
 * 1.5), clamped at y = 0.  counters[3] points lowered, [4] cells with >= 10 points, [5] cells adjusted. */
int dp_ground_grid_adjust(dp_engine* e, float* xyz, int64_t n, int grid_size, double percentile,
                          uint64_t* counters, void* stream);

/* Parity taps: copy a named stage tensor of the LAST dp_forward as float32 in the
 * reference's layout (NCHW / (n,577,C)).  Returns the element count in *numel. */
int dp_tap(dp_engine* e, const char* stage, float* out, int64_t capacity, int64_t* numel,
           void* stream);

/* Unit-test / microbenchmark entry for the GEMM cores:  C[M,N] = A[M,K] * W[N,K]^T (+bias),
 * fp32 row-major in and out; `backend` 0 = fp32 CUDA-core, 1 = bf16 tcgen05 (inputs are
 * rounded to bf16 on the fly).  `act`: low byte = activation (0 none, 1 ReLU, 2 GELU); for the
 * bf16 backend the high bits select the epilogue form under test: 0x100 bf16 output through the
 * TMA-store epilogue (returned as fp32), 0x200 fp32 residual form  C += bias * (acc + bias)  in
 * place (LayerScale gamma := bias), 0x400 ConvTranspose k2 s2 pixel shuffle (M = S*S pixels,
 * N = 4*Cout; C is the (2S, 2S, Cout) map), 0x800 with 0x100 / 0x400: dual store, C = ReLU twin;
 * 0x1000 with 0x100, K = 1024: LayerNorm folded into the GEMM (timm Block: norm1 -> qkv, norm2 -> fc1),
 * C = LN(A; g, b_ln, eps 1e-6) * W^T + bias with g[k] = 1 + 0.25 sin(0.37 k), b_ln[k] = 0.1 cos(0.11 k);
 * 0x2000 with 0x200, N = 1024: the residual form also emits bf16(x) and per-row partial sums for the
 * next folded GEMM; C is (2M, N) and rows [M, 2M) return (x - mean) / sqrt(var + 1e-6) rebuilt from them. */
int dp_gemm_test(dp_engine* e, int backend, const float* A, const float* Wt, const float* bias,
                 float* C, int M, int N, int K, int act, void* stream);
/* Same for 3x3 / pad 1 / stride 1 convolution over NHWC fp32 (B,H,W,Cin) with OIHW weights.
 * `backend` high bits (bf16 only): 0x100 bf16 output through the TMA-store epilogue, 0x800 (with
 * 0x100) dual store, y = ReLU twin. */
int dp_conv3x3_test(dp_engine* e, int backend, const float* x_nhwc, const float* w_oihw,
                    const float* bias, float* y_nhwc, int B, int H, int W, int Cin, int Cout,
                    void* stream);
/* Attention core: qkv fp32 (n,577,3072) -> out fp32 (n,577,1024), 16 heads x 64.  `backend` low byte: 0 fp32
 * CUDA-core, 1 bf16 tcgen05, 2 bf16 mma.sync.  For backend 1, bits 8-15 = 1 + kernel variant (0 scalar exp2 chain +
 * strict MUFU ping-pong, 5 packed fp32x2 chain + early hand-over of the MUFU turn, 12 = 5 with P through TMEM, 13 = 12 with
 * 25 % of the exponentials on the FMA pipe, 15-18 = 12 / 13 with the P V MMA split in two 64-key halves; the default is
 * the fastest measured one, csrc/attention_tc.cu), 0xFF = back to the default, and bit 16 = no MUFU ping-pong; the choice
 * is process-wide and sticky (same switch as DEPTHPRO_ATTN_EXP / DEPTHPRO_ATTN_PINGPONG). */
int dp_attention_test(dp_engine* e, int backend, const float* qkv, float* out, int n,
                      void* stream);

/* Micro-benchmark of one bf16 core on scratch buffers; mean ms per launch over `iters` launches.
 * kind 0 GEMM+bias, 1 GEMM+bias+GELU, 2 GEMM+bias*gamma+fp32 residual (in place), 3 conv3x3 on an
 * MxM map (Cin=K, Cout=N), 4 attention over M sequences, 5 LayerNorm over M rows.  Kinds 6-10 time
 * the HBM-bound kernels either side of the network on an M x N (H x W) image, L2 flushed before
 * every launch: 6 uint8 HWC -> fused transform + resize -> fp32 3x1536^2, 7 pyramid + 35-patch split +
 * im2col of one 1536^2 frame, 8 depth epilogue 1536^2 -> M x N, 9 unprojection + colours, 10 colourise.
 * Kinds 11 / 12 / 13: kinds 0 / 1 / 2 in their LayerNorm-folded forms.  A/B bits (process-wide, sticky):
 * kind | 0x100 / 0x200 switches the fp32-residual forms' L2 prefetch on / off (DEPTHPRO_RES_PREFETCH), kind | 0x400 /
 * 0x800 their persisting-L2 window on (set-aside = `iters >> 16` MB, 0 = 96; `iters` is taken modulo 65536) / off
 * (DEPTHPRO_L2_PERSIST_MB), kind | 0x1000 / 0x2000 the opt-in second-generation HBM kernels on / off
 * (DEPTHPRO_HBM_V2); for kind 4,
 * N = 1 + variant + 64 * (no ping-pong) selects the attention variant (0 leaves it unchanged). */
int dp_kernel_bench(dp_engine* e, int kind, int M, int N, int K, int iters, float* ms_out);

/* Per-launch CUDA-event profiling of the hot kernels (used by bench.py for the roofline line).
 * Classes: 0 tcgen05 GEMM, 1 tcgen05 conv3x3, 2 attention, 3 LayerNorm, 4 fp32 CUDA-core GEMM.
 * `work` is algorithmic FLOPs (classes 0,1,2,4) or bytes (class 3).  Arrays hold 5 entries.
 * dp_profile_collect synchronises the device and sums everything since dp_profile_enable(e,1). */
#define DP_NUM_KERNEL_CLASSES 5
int dp_profile_enable(dp_engine* e, int on);
int dp_profile_collect(dp_engine* e, double* ms_by_class, double* work_by_class,
                       int64_t* launches_by_class);

/* Debug counters of the kernels (tests only; synchronises the device).  id 0: softmax warps of the tcgen05 attention
 * kernel that took the lazy-maximum RESCALE branch since the last reset (timm SDPA semantics are unchanged by it; the
 * counter proves a test input really exercised it).  Returns the value and resets it when `reset` != 0; -1 on error. */
int64_t dp_debug_counter(dp_engine* e, int id, int reset);

/* Kernel launches issued by this engine since creation (bench.py's `gpu_launches`). */
int64_t dp_launch_count(dp_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* DEPTHPRO_B200_H */
