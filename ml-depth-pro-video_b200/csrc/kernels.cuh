// HBM-bound kernels of the Depth Pro engine: resize, pyramid/split/im2col gather, LayerNorm
// (+ merge gather), small direct convolutions, metric-depth epilogue, unprojection, colourise,
// weight repacking and layout conversion helpers.  Declarations only; see kernels.cu.
#pragma once
#include "common.cuh"

namespace dp {

// ---- preprocessing ------------------------------------------------------------------
// src (u8 HWC or f32 CHW, B images of HxW) -> x (B,3,1536,1536) f32, bilinear align_corners=False.
// `interp`: INTERP_BILINEAR or INTERP_BICUBIC (the two modes F.interpolate accepts with align_corners=False).
enum { INTERP_BILINEAR = 0, INTERP_BICUBIC = 1 };
void resize_to_1536(const void* src, int src_fmt, int B, int H, int W, float* x, int interp, cudaStream_t s);

// x (B,3,1536,1536) f32 -> patch-embed im2col rows.  Frame-major order: row =
// ((b*35 + patch)*576 + ty*24 + tx), col = c*256 + ky*16 + kx.  patch 0..24 = 5x5 windows of the
// full image (stride 288), 25..33 = 3x3 windows of the 2x box-mean image (stride 192),
// 34 = the 4x-downsampled image.  `A35` gets all 35 patches, `A1` (optional) only patch 34.
template <typename T>
void split_im2col(const float* x, int B, T* A35, T* A1, cudaStream_t s);
// im2col rows (fp32) -> patches (35B,3,384,384) in the reference's patch-major / batch-minor order.
void im2col_to_ref_patches(const float* A35, int B, float* patches, cudaStream_t s);

// ---- tokens ---------------------------------------------------------------------------
// resid[(r*577)*1024 + n] = cls[n] + pos[n] for every sequence r.
void write_cls_rows(float* resid, const float* cls, const float* pos, int nseq, cudaStream_t s);

struct RowMap {
  int mode = 0;       // 0 identity, 1 merge gather (dest pixel -> source token)
  int S = 0;          // dest grid side (96 / 48 / 24)
  int steps = 1;      // patches per side
  int pad = 0;        // tokens cropped on interior edges
  int patch_base = 0; // first patch of this pyramid level inside a frame's 35
  int sb = 35, sp = 1;  // source sequence index = seq_off + b*sb + patch*sp
  int seq_off = 0;
};
// LayerNorm over C=1024 (eps 1e-6) of gathered rows; `ln` == 0 copies/converts only.
// in: fp32 rows of width 1024; out: T rows (n_out x 1024).
// Grouped launches: rows [0, end[0]) use (w[0], b[0]), [end[0], end[1]) use (w[1], b[1]), ...
struct LnGroups {
  int n = 1;
  long long end[3] = {0, 0, 0};
  const float* w[3] = {nullptr, nullptr, nullptr};
  const float* b[3] = {nullptr, nullptr, nullptr};
};
template <typename T>
void layernorm_rows(const float* in, T* out, const float* w, const float* b, long long n_out,
                    RowMap map, int ln, cudaStream_t s, const bf16* in_hi = nullptr, const bf16* in_lo = nullptr);
template <typename T>
void layernorm_rows_grouped(const float* in, T* out, const LnGroups& g, long long n_out, cudaStream_t s);
// generic-width gather without LN (used by dp_merge): in (nseq,577,C) f32 -> out (B,S,S,C) f32
void merge_rows_f32(const float* in, float* out, int B, int C, RowMap map, cudaStream_t s);

// ---- small direct convolution (FOV head) ------------------------------------------------
// NHWC in (B,H,W,Cin) of type T, weights HWIO fp32 (k,k,Cin,Cout), out NHWC (B,Ho,Wo,Cout) of T.
// v = conv + bias; relu optional; then + add_tokens (same NHWC layout as y) if not null.
template <typename T>
void conv_direct(const T* x, const float* w_hwio, const float* bias, T* y, int B, int H, int W, int Cin,
                 int Cout, int k, int stride, int pad, int relu, const T* add_tokens, cudaStream_t s);
// im2col for small strided convolutions: NHWC (B,H,W,C) -> rows (B*Ho*Wo, k*k*C), tap-major
// (column = (ky*k + kx)*C + c, matching the O(HW)I weight packing); out-of-image taps are zero.
template <typename T>
void im2col_nhwc(const T* x, T* cols, int B, int H, int W, int C, int k, int stride, int pad, cudaStream_t s);
// final 6x6 valid conv over (B,6,6,32) -> fov_deg[B] (fp32)
template <typename T>
void fov_final(const T* x, const float* w_hwio, const float* bias, float* fov_deg, int B, cudaStream_t s);

// micro-benchmark operands: bf16 uniform in [-1, 1)
void fill_random_bf16(void* p, size_t bytes, unsigned seed, cudaStream_t s);

// ---- LayerNorm folded into the ViT GEMMs (common.cuh GemmOp::ln_stats) -------------------------
// x fp32 (rows, 1024) -> raw bf16 copy + per-row partial sums [rows][LN_SLOTS][2] (slot 0 filled)
// with `xlo`: also the low half of the (hi, lo) pair form of the stream, xlo = round16(x - xb)
void ln_stats_cast(const float* in, bf16* xb, float* stats, long long rows, cudaStream_t s, bf16* xlo = nullptr);
// x = hi + lo (test helper of the pair form)
void pair_to_f32(const bf16* hi, const bf16* lo, float* x, long long n, cudaStream_t s);
// wf = bf16(g * w) [N,K], c = colsum(wf), d = bias + w b_ln   (w fp32 [N,K], K = 1024)
void ln_fold(const float* w, const float* g, const float* b_ln, const float* bias, bf16* wf, float* c, float* d, int N,
             int K, cudaStream_t s);
// test helper: (xb - mean) * rstd from the producer epilogue's outputs
void ln_apply_from_stats(const bf16* xb, const float* stats, float* y, long long rows, cudaStream_t s);

// ---- metric depth epilogue ----------------------------------------------------------------
// f_px[b] = f_px_in ? f_px_in[b] : 0.5*W / tan(0.5*deg2rad(fov_deg[b]))   (depth_pro.py:282-283)
void compute_fpx(const float* fov_deg, const float* f_px_in, int W, float* f_px, int B, cudaStream_t s);
// depth[b,y,x] = 1 / clamp(resize(canon * (W / f_px[b]))[y,x], 1e-4, 1e4)    (depth_pro.py:285-293)
void depth_epilogue(const float* canon, const float* f_px, int B, int H, int W, float* depth, int interp, cudaStream_t s);
void hbm_v2_set(int on);  // A/B switch of the opt-in second-generation HBM kernels (DEPTHPRO_HBM_V2)

// ---- video add-on ---------------------------------------------------------------------------
void unproject(const float* depth, const uint8_t* rgb, int H, int W, const float* f_px, float* xyz,
               float* rgb_out, uint8_t* valid_mask, int64_t* n_valid, int* scratch, cudaStream_t s);
size_t unproject_scratch_ints(int H, int W);
// `minmax`: >= colorize_scratch_bytes() of device scratch; min_depth / max_depth: NaN = the image's own nan-min / nan-max
size_t colorize_scratch_bytes();
void colorize(const float* depth, int H, int W, const uint8_t* lut, void* out, float* minmax, float min_depth,
              float max_depth, cudaStream_t s);

// ---- ground normalisation of a point cloud (ground.cu; img_to_normalized_pointcloud.py:880-1118) -------
// xyz float32 (n,3) in place; `scratch` >= ground_scratch_bytes(n, grid_size) bytes; `counters` (device, 6 x
// uint64, optional): ground points, set to y=0, limited to -0.1, lowered by the grid pass, cells with >= 10
// points, cells adjusted
size_t ground_scratch_bytes(long long n, int grid_size);
void ground_normalize(float* xyz, long long n, const double normal[3], double d, void* scratch,
                      unsigned long long* counters, cudaStream_t s);
void ground_grid_adjust(float* xyz, long long n, int grid_size, double percentile, void* scratch,
                        unsigned long long* counters, cudaStream_t s);

// ---- layout / dtype helpers ---------------------------------------------------------------
template <typename TI, typename TO>
void convert(const TI* in, TO* out, long long n, cudaStream_t s);
// NHWC (B,H,W,C) of T -> NCHW fp32
template <typename T>
void nhwc_to_nchw_f32(const T* in, float* out, int B, int H, int W, int C, cudaStream_t s);
void nchw_to_nhwc_f32(const float* in, float* out, int B, int C, int H, int W, cudaStream_t s);
// weight repacks (fp32 source in PyTorch layout)
template <typename T>
void pack_oihw_to_ohwi(const float* w, T* out, int O, int I, int KH, int KW, cudaStream_t s);  // conv
template <typename T>
void pack_convT_iohw(const float* w, T* out, int I, int O, cudaStream_t s);  // (I,O,2,2) -> ((dy,dx,o), i)
// head.1 (ConvT 128->128 k2 s2, weight (ci,c1,2,2) + bias b1) followed by head.2 (conv3x3 128->32,
// weight (c2,c1,3,3) + bias b2) composed in fp32 into one 3x3 conv over the coarse map:
//   wc[(py*2+px)*32 + c2][(ty*3+tx)*128 + ci]  (bf16, [128][1152]);
//   cb[tap][c2] = sum_c1 b1[c1]*w2[c2,c1,tap]  (9x32), cb[9][c2] = b2[c2] + sum_tap cb[tap][c2].
void compose_head(const float* w1, const float* b1, const float* w2, const float* b2, bf16* wc, float* cb,
                  cudaStream_t s);
// FeatureFusionBlock2d tail (decoder.py:176-178): ConvTranspose2d k2 s2 (no bias, weight (ci,co,2,2))
// followed by a 1x1 conv (weight (co',co)) composed in fp32 into one ConvT in the GEMM layout
//   wc[(dy*2+dx)*C + co'][ci] = sum_co wo[co'][co] * wd[ci][co][dy][dx]      (C x C channels)
// wc (O, 9, C) bf16 = conv3x3 (w3 OIHW, b3) o conv1x1 (wo [c][i], bo); cb [10][O]: per-tap share of bo, then the interior bias
void compose_1x1_conv3x3(const float* wo, const float* bo, const float* w3, const float* b3, bf16* wc, float* cb, int O,
                         int C, cudaStream_t s);
void compose_deconv_1x1(const float* wd, const float* wo, bf16* wc, int C, cudaStream_t s);
void pack_oihw_to_hwio_f32(const float* w, float* out, int O, int I, int KH, int KW, cudaStream_t s);

}  // namespace dp
