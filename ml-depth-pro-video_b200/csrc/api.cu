// extern "C" boundary of the engine (include/depthpro_b200.h).  No C++ or torch types cross it.
#include <cmath>
#include <string>

#include "../../include/depthpro_b200.h"
#include "attention.cuh"
#include "engine.cuh"

namespace {
thread_local std::string g_err;

template <typename F>
int guard(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  } catch (...) {
    g_err = "unknown error";
    return 2;
  }
}
inline dp::Engine* E(dp_engine* e) {
  if (!e) throw dp::Error("null engine handle");
  return reinterpret_cast<dp::Engine*>(e);
}
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
}  // namespace

extern "C" {

const char* dp_last_error(void) { return g_err.c_str(); }
int dp_version(void) { return 200; }
const char* dp_act_dtype(void) { return DP_ACT_NAME; }

int dp_engine_create(int device, int precision, int max_batch, dp_engine** out) {
  return guard([&] {
    if (!out) throw dp::Error("null out pointer");
    *out = reinterpret_cast<dp_engine*>(new dp::Engine(device, precision, max_batch));
  });
}
int dp_engine_create_ex(int device, int precision, int max_batch, int fov_mode, dp_engine** out) {
  return guard([&] {
    if (!out) throw dp::Error("null out pointer");
    *out = reinterpret_cast<dp_engine*>(new dp::Engine(device, precision, max_batch, fov_mode));
  });
}
int dp_engine_destroy(dp_engine* e) {
  return guard([&] { delete reinterpret_cast<dp::Engine*>(e); });
}
int dp_engine_set_weight(dp_engine* e, const char* name, const void* data, const int64_t* shape, int ndim, int on_device) {
  return guard([&] { E(e)->set_weight(name, data, shape, ndim, on_device != 0); });
}
int dp_engine_missing_weights(dp_engine* e) {
  int n = -1;
  guard([&] { n = E(e)->missing_weights(); });
  return n;
}
int dp_engine_finalize(dp_engine* e) {
  return guard([&] { E(e)->finalize(); });
}
int dp_preprocess(dp_engine* e, const void* img, int B, int H, int W, int src_fmt, float* x_1536, void* stream) {
  return guard([&] { E(e)->preprocess(img, B, H, W, src_fmt, x_1536, 0, S(stream)); });
}
int dp_preprocess_ex(dp_engine* e, const void* img, int B, int H, int W, int src_fmt, int interp_mode, float* x_1536,
                     void* stream) {
  return guard([&] { E(e)->preprocess(img, B, H, W, src_fmt, x_1536, interp_mode, S(stream)); });
}
int dp_split(dp_engine* e, const float* x_1536, int B, float* patches, void* stream) {
  return guard([&] { E(e)->split(x_1536, B, patches, S(stream)); });
}
int dp_merge(dp_engine* e, const float* tokens, int B, int steps, int padding, int C, float* merged, void* stream) {
  return guard([&] { E(e)->merge(tokens, B, steps, padding, C, merged, S(stream)); });
}
int dp_forward(dp_engine* e, const float* x_1536, int B, float* canon_inv_depth, float* fov_deg, void* stream) {
  return guard([&] { E(e)->forward(x_1536, B, canon_inv_depth, fov_deg, S(stream)); });
}
int dp_infer(dp_engine* e, const void* img, int B, int H, int W, int src_fmt, const float* f_px_host, float* depth_out,
             float* f_px_out, void* stream) {
  return guard([&] { E(e)->infer(img, B, H, W, src_fmt, f_px_host, depth_out, f_px_out, 0, S(stream)); });
}
int dp_infer_ex(dp_engine* e, const void* img, int B, int H, int W, int src_fmt, int interp_mode, const float* f_px_host,
                float* depth_out, float* f_px_out, void* stream) {
  return guard([&] { E(e)->infer(img, B, H, W, src_fmt, f_px_host, depth_out, f_px_out, interp_mode, S(stream)); });
}
int dp_infer_host(dp_engine* e, const void* img_host, int B, int H, int W, int src_fmt, const float* f_px_host,
                  float* depth_out_host, float* f_px_out_host) {
  return guard([&] { E(e)->infer_host(img_host, B, H, W, src_fmt, f_px_host, depth_out_host, f_px_out_host); });
}
int dp_unproject(dp_engine* e, const float* depth, const uint8_t* rgb, int H, int W, const float* f_px_dev, float* xyz,
                 float* rgb_out, uint8_t* valid_mask, int64_t* n_valid, void* stream) {
  return guard([&] { E(e)->unproject(depth, rgb, H, W, f_px_dev, xyz, rgb_out, valid_mask, n_valid, S(stream)); });
}
int dp_colorize(dp_engine* e, const float* depth, int H, int W, const uint8_t* lut, void* out, void* stream) {
  return guard([&] { E(e)->colorize(depth, H, W, lut, out, NAN, NAN, S(stream)); });
}
int dp_colorize_range(dp_engine* e, const float* depth, int H, int W, const uint8_t* lut, void* out, float min_depth,
                      float max_depth, void* stream) {
  return guard([&] { E(e)->colorize(depth, H, W, lut, out, min_depth, max_depth, S(stream)); });
}
int dp_ground_normalize(dp_engine* e, float* xyz, int64_t n, const double* normal3, double d, uint64_t* counters,
                        void* stream) {
  return guard([&] { E(e)->ground_normalize(xyz, n, normal3, d, counters, S(stream)); });
}
int dp_ground_grid_adjust(dp_engine* e, float* xyz, int64_t n, int grid_size, double percentile, uint64_t* counters,
                          void* stream) {
  return guard([&] { E(e)->ground_grid_adjust(xyz, n, grid_size, percentile, counters, S(stream)); });
}
int dp_tap(dp_engine* e, const char* stage, float* out, int64_t capacity, int64_t* numel, void* stream) {
  return guard([&] {
    const int64_t n = E(e)->tap(stage, out, capacity, S(stream));
    if (numel) *numel = n;
  });
}
int dp_gemm_test(dp_engine* e, int backend, const float* A, const float* Wt, const float* bias, float* C, int M, int N,
                 int K, int act, void* stream) {
  return guard([&] { E(e)->gemm_test(backend, A, Wt, bias, C, M, N, K, act, S(stream)); });
}
int dp_conv3x3_test(dp_engine* e, int backend, const float* x_nhwc, const float* w_oihw, const float* bias, float* y_nhwc,
                    int B, int H, int W, int Cin, int Cout, void* stream) {
  return guard([&] { E(e)->conv3x3_test(backend, x_nhwc, w_oihw, bias, y_nhwc, B, H, W, Cin, Cout, S(stream)); });
}
int dp_attention_test(dp_engine* e, int backend, const float* qkv, float* out, int n, void* stream) {
  return guard([&] { E(e)->attention_test(backend, qkv, out, n, S(stream)); });
}
int dp_kernel_bench(dp_engine* e, int kind, int M, int N, int K, int iters, float* ms_out) {
  return guard([&] { *ms_out = E(e)->kernel_bench(kind, M, N, K, iters); });
}
int dp_profile_enable(dp_engine* e, int on) {
  (void)e;
  return guard([&] { dp::prof_enable(on != 0); });
}
int dp_profile_collect(dp_engine* e, double* ms_by_class, double* work_by_class, int64_t* launches_by_class) {
  (void)e;
  return guard([&] {
    long long l[dp::KC_COUNT];
    dp::prof_collect(ms_by_class, work_by_class, l);
    for (int i = 0; i < dp::KC_COUNT; ++i) launches_by_class[i] = l[i];
  });
}
int64_t dp_debug_counter(dp_engine* e, int id, int reset) {
  int64_t v = -1;
  guard([&] {
    E(e);
    if (id != 0) throw dp::Error("dp_debug_counter: unknown id");
    v = static_cast<int64_t>(dp::attention_tc_rescale_count(reset != 0));
  });
  return v;
}
int64_t dp_launch_count(dp_engine* e) {
  (void)e;
  return dp::g_launches;
}

}  // extern "C"
