// Depth Pro engine: owns packed weights + workspace of one GPU and runs the forward graph
// (src/depth_pro/depth_pro.py:218-298 of the reference) as a fixed sequence of kernel launches.
#pragma once

#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace dp {

struct VitWeights {
  const float *cls, *pos, *pe_b, *norm_w, *norm_b;
  const void* pe_w;
  struct Block {
    const float *n1w, *n1b, *qkv_b, *proj_b, *g1, *n2w, *n2b, *fc1_b, *fc2_b, *g2;
    const void *qkv_w, *proj_w, *fc1_w, *fc2_w;
    // LayerNorm-folded form (bf16 mode): qkv_w / fc1_w then hold bf16(g * W), qkv_b / fc1_b hold W b_ln + b
    const float *qkv_c = nullptr, *fc1_c = nullptr;  // column sums of the folded weights
  } blk[24];
};

struct Packed {
  void* ptr = nullptr;
  size_t bytes = 0;
};

class Engine {
 public:
  // fov_mode (depth_pro.py:100-108, fov.py:29-55): 0 = use_fov_head=False (no FOV network: forward yields no fov_deg and
  // infer needs a caller-supplied f_px), 1 = FOV head without its own encoder (fov_encoder_preset=None: four convs on the
  // low-resolution decoder feature), 2 = the default (ViT-L fov encoder + Linear + downsample + head)
  Engine(int device, int prec, int max_batch, int fov_mode = 2);
  ~Engine();

  void set_weight(const std::string& name, const void* data, const int64_t* shape, int ndim, bool on_device);
  int missing_weights() const;
  void finalize();

  void preprocess(const void* img, int B, int H, int W, int src_fmt, float* x, int interp, cudaStream_t s);
  void split(const float* x, int B, float* patches, cudaStream_t s);
  void merge(const float* tokens, int B, int steps, int padding, int C, float* merged, cudaStream_t s);
  void forward(const float* x, int B, float* canon, float* fov_deg, cudaStream_t s);
  void infer(const void* img, int B, int H, int W, int src_fmt, const float* f_px_host, float* depth, float* f_px_out,
             int interp, cudaStream_t s);
  void infer_host(const void* img_host, int B, int H, int W, int src_fmt, const float* f_px_host, float* depth_host,
                  float* f_px_out_host);
  void unproject(const float* depth, const uint8_t* rgb, int H, int W, const float* f_px_dev, float* xyz, float* rgb_out,
                 uint8_t* valid_mask, int64_t* n_valid, cudaStream_t s);
  void colorize(const float* depth, int H, int W, const uint8_t* lut, void* out, float min_depth, float max_depth,
                cudaStream_t s);
  // img_to_normalized_pointcloud.py:880-1118 on the GPU (ground.cu); `counters` = 6 x uint64 on the device or null
  void ground_normalize(float* xyz, int64_t n, const double* normal3, double d, uint64_t* counters, cudaStream_t s);
  void ground_grid_adjust(float* xyz, int64_t n, int grid_size, double percentile, uint64_t* counters, cudaStream_t s);
  int64_t tap(const std::string& stage, float* out, int64_t capacity, cudaStream_t s);

  void gemm_test(int backend, const float* A, const float* Wt, const float* bias, float* C, int M, int N, int K, int act,
                 cudaStream_t s);
  void conv3x3_test(int backend, const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Cin,
                    int Cout, cudaStream_t s);
  void attention_test(int backend, const float* qkv, float* out, int n, cudaStream_t s);
  float kernel_bench(int kind, int M, int N, int K, int iters);

  int device() const { return device_; }
  int prec() const { return prec_; }

 private:
  template <typename T>
  void forward_impl(const float* x, int B, float* canon, float* fov_deg, cudaStream_t s);
  // xbuf_ -> canon_, fov_ (the buffers dp_infer uses): replays a captured CUDA graph of the whole forward pass when
  // graphs are enabled and one exists for this batch size, captures one on the second call, runs eagerly otherwise
  void forward_cached(int B, cudaStream_t s);
  void drop_graphs();
  template <typename T>
  void run_vits(int B, cudaStream_t s);
  template <typename T>
  void decode_frame(int f, float* canon, float* fov_deg, cudaStream_t s);
  template <typename T>
  int64_t tap_impl(const std::string& stage, float* out, int64_t capacity, cudaStream_t s);

  void* alloc(size_t bytes);
  const void* W(const std::string& name) const;   // packed weight (activation dtype)
  const float* F(const std::string& name) const;  // packed fp32 tensor
  VitWeights vit_weights(const std::string& prefix) const;
  size_t esz() const { return prec_ == BF16 ? 2 : 4; }

  int device_, prec_, max_batch_;
  int fov_mode_ = 2;
  int n_enc_ = 3;   // ViT encoders run as one grouped batch: patch, image (+ fov when fov_mode_ == 2)
  int seqs_per_frame() const { return 35 + n_enc_ - 1; }
  bool finalized_ = false;
  bool weights_changed_ = false;  // a set_weight since the last finalize
  bool attn_legacy_ = false;
  std::map<std::string, std::vector<int64_t>> manifest_;
  std::unordered_map<std::string, Packed> packed_;
  std::unordered_map<std::string, Packed> raw_;  // fp32 originals needed for weight composition
  std::vector<void*> allocs_;
  float* stage_ = nullptr;  // staging buffer for host -> device weight upload
  size_t stage_bytes_ = 0;

  VitWeights vit_patch_, vit_image_, vit_fov_;

  // ---- workspace (activation dtype unless noted)
  float* xbuf_ = nullptr;      // (max_batch,3,1536,1536) f32
  float* canon_ = nullptr;     // (max_batch,1536,1536) f32
  float* fov_ = nullptr;       // (max_batch) f32
  float* fpx_ = nullptr;       // (max_batch) f32
  float* fpx_in_ = nullptr;    // (max_batch) f32
  void* A35_;                  // im2col rows: 35*B patch sequences, then B copies of patch 34 (image/fov encoders)
  float* resid_;               // residual stream f32, 37 sequences per frame (35 patch + image + fov)
  void *xn_, *qkv_, *attn_, *hid_;
  void *lat0m_, *lat1m_, *x0m_, *x1m_, *x2m_, *globm_, *fovtok_;  // merged maps (max_batch frames)
  // per-frame decoder workspace
  void *u0a_, *u0b_, *u0c_, *enc0_, *enc0r_, *u1a_, *u1b_, *enc1_, *u2a_, *enc2_, *u3a_, *enc3_, *u4a_, *cat_, *enc4_;
  void *lowres_, *lowres_r_, *x1_, *x1r_, *t_, *x_, *xr_, *x2_, *y_, *feat_[5];
  void *c1_, *c1r_, *c2_, *c2r_, *c3_, *c3r_;  // decoder.convs.{1,2,3} outputs (+ ReLU twins), one set per level
  void *h0_, *h1_;
  void* head_wc_ = nullptr;    // composed head.1 o head.2 weights (bf16 mode)
  float* head_cb_ = nullptr;
  void* head0_wc_ = nullptr;   // composed fusions.0.out_conv o head.0 weights (bf16 mode), [128][9][256]
  float* head0_cb_ = nullptr;  // [9][128] per-tap share of out_conv's bias, then [128] the interior bias
  bool head0_fused_ = false;   // DEPTHPRO_HEAD0_FUSE=0 disables
  bool serpentine_dec_ = false;  // ... and the decoder's main chain (DEPTHPRO_SERPENTINE=1: ViT only)
  bool serpentine_ = false;    // ViT kernels alternate their tile direction (GemmOp::reverse); DEPTHPRO_SERPENTINE=0 disables
  void *fovlin_, *fov_a_, *fov_b_, *fov_c_, *fovcol_;
  int last_B_ = 0;
  // host-call staging
  void* himg_ = nullptr;
  size_t himg_bytes_ = 0;
  float* hdepth_ = nullptr;
  size_t hdepth_bytes_ = 0;
  bool res_pair_ = false;      // bf16 mode + folded LayerNorm: residual stream as a (hi, lo) 16-bit pair (DEPTHPRO_RES_PAIR=0 disables)
  void* xlo_ = nullptr;        // ... its low half (the high half is xn_)
  bool ln_fuse_ = false;       // bf16 mode: LayerNorm folded into qkv / fc1 (DEPTHPRO_LN_FUSE=0 disables)
  float* ln_stats_ = nullptr;  // (tokens, LN_SLOTS, 2) partial row sums of the residual stream
  float* colorize_mm_ = nullptr;
  void* ground_scratch_ = nullptr;
  size_t ground_scratch_bytes_ = 0;
  void* ground_scratch(int64_t n, int grid_size);
  int* unproject_scratch_ = nullptr;
  size_t unproject_scratch_ints_ = 0;
  cudaStream_t host_stream_ = nullptr;

  // ---- concurrency inside a frame (decode_frame): the five encoder project + upsample branches, the three decoder skip
  // convs and the FOV head are independent of the main fusion chain and are latency-bound small launches (18-80 CTAs
  // each); they run on side streams forked from / joined into the caller's stream with events.
  static constexpr int NSIDE = 5;
  cudaStream_t side_[NSIDE] = {};
  cudaEvent_t ev_fork_ = nullptr, ev_lowres_ = nullptr, ev_side_[NSIDE] = {};
  bool multi_stream_ = true;     // DEPTHPRO_STREAMS=0: everything on the caller's stream (round-1 behaviour)

  // ---- CUDA graphs: one instantiated graph of the forward pass per batch size (DEPTHPRO_GRAPH=0 disables)
  struct FrameGraph {
    cudaGraphExec_t exec = nullptr;
    long long launches = 0;      // kernel launches recorded while capturing (added to the launch counter per replay)
    unsigned epoch = 0;
    int eager_calls = 0;
  };
  std::map<int, FrameGraph> graphs_;
  bool use_graph_ = true;
  cudaStream_t cap_stream_ = nullptr;
};

}  // namespace dp
