// Depth Pro engine implementation (see engine.cuh).  Reference call stack followed:
//   DepthPro.infer            src/depth_pro/depth_pro.py:243-298
//   DepthPro.forward          src/depth_pro/depth_pro.py:218-241
//   DepthProEncoder.forward   src/depth_pro/network/encoder.py:233-332
//   timm forward_features     (third party; wired at network/vit_factory.py:97-110)
//   MultiresConvDecoder       src/depth_pro/network/decoder.py:74-93, 166-180
//   FOVNetwork.forward        src/depth_pro/network/fov.py:56-82
#include "engine.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <type_traits>

#include <nvtx3/nvToolsExt.h>

#include "attention.cuh"
#include "gemm.cuh"
#include "kernels.cuh"

namespace dp {

std::atomic<int64_t> g_launches{0};
std::atomic<unsigned> g_config_epoch{0};

NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

bool pdl_enabled() {
  static const bool on = [] {
    // on by default since the frame is replayed as a CUDA graph (round 2: +0.7 % frames/s); "0" disables
    const char* e = getenv("DEPTHPRO_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

// ---------------------------------------------------------------------------- profiler
namespace {
struct ProfRec {
  cudaEvent_t a, b;
  int cls;
  double work;
  cudaStream_t s = nullptr;
  bool closed = false;
};
// The per-launch profiler is process-wide (bench.py switches it on for a replay of the timed steps): one mutex guards
// the records and the event pool, the on/off flag is atomic so the disabled fast path takes no lock.
std::atomic<bool> g_prof_on{false};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_event_pool;
cudaEvent_t new_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  DP_CUDA(cudaEventCreate(&e));
  return e;
}
}  // namespace

bool prof_active() { return g_prof_on.load(); }
void prof_enable(bool on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) g_event_pool.push_back(r.a), g_event_pool.push_back(r.b);
  g_prof.clear();
  g_prof_on = on;
}
void prof_begin(cudaStream_t s, int cls, double work) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{new_event(), new_event(), cls, work, s, false};
  DP_CUDA(cudaEventRecord(r.a, s));
  g_prof.push_back(r);
}
void prof_end(cudaStream_t s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  // the record this scope opened is the last one of ITS stream (scopes of one stream nest strictly; profiling several
  // engines at once interleaves records, so match on the stream)
  for (auto it = g_prof.rbegin(); it != g_prof.rend(); ++it)
    if (it->s == s && !it->closed) {
      DP_CUDA(cudaEventRecord(it->b, s));
      it->closed = true;
      return;
    }
}
void prof_collect(double* ms, double* work, long long* launches) {
  DP_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < KC_COUNT; ++i) ms[i] = 0, work[i] = 0, launches[i] = 0;
  for (auto& r : g_prof) {
    if (!r.closed) continue;
    float t = 0.f;
    DP_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.cls] += t, work[r.cls] += r.work, launches[r.cls] += 1;
  }
}

namespace {

constexpr int IMG = 1536, EMB = 1024, SEQ = 577;

void add_vit(std::map<std::string, std::vector<int64_t>>& m, const std::string& p) {
  m[p + "cls_token"] = {1, 1, EMB};
  m[p + "pos_embed"] = {1, SEQ, EMB};
  m[p + "patch_embed.proj.weight"] = {EMB, 3, 16, 16};
  m[p + "patch_embed.proj.bias"] = {EMB};
  for (int i = 0; i < 24; ++i) {
    const std::string b = p + "blocks." + std::to_string(i) + ".";
    m[b + "norm1.weight"] = {EMB};
    m[b + "norm1.bias"] = {EMB};
    m[b + "attn.qkv.weight"] = {3 * EMB, EMB};
    m[b + "attn.qkv.bias"] = {3 * EMB};
    m[b + "attn.proj.weight"] = {EMB, EMB};
    m[b + "attn.proj.bias"] = {EMB};
    m[b + "ls1.gamma"] = {EMB};
    m[b + "norm2.weight"] = {EMB};
    m[b + "norm2.bias"] = {EMB};
    m[b + "mlp.fc1.weight"] = {4 * EMB, EMB};
    m[b + "mlp.fc1.bias"] = {4 * EMB};
    m[b + "mlp.fc2.weight"] = {EMB, 4 * EMB};
    m[b + "mlp.fc2.bias"] = {EMB};
    m[b + "ls2.gamma"] = {EMB};
  }
  m[p + "norm.weight"] = {EMB};
  m[p + "norm.bias"] = {EMB};
}

// name -> shape of the reference state_dict (pinned by tests/golden/state_dict_manifest.json)
std::map<std::string, std::vector<int64_t>> build_manifest(int fov_mode) {
  std::map<std::string, std::vector<int64_t>> m;
  add_vit(m, "encoder.patch_encoder.");
  add_vit(m, "encoder.image_encoder.");
  if (fov_mode == 2) add_vit(m, "fov.encoder.0.");
  m["encoder.upsample_latent0.0.weight"] = {256, EMB, 1, 1};
  for (int i = 1; i <= 3; ++i) m["encoder.upsample_latent0." + std::to_string(i) + ".weight"] = {256, 256, 2, 2};
  m["encoder.upsample_latent1.0.weight"] = {256, EMB, 1, 1};
  for (int i = 1; i <= 2; ++i) m["encoder.upsample_latent1." + std::to_string(i) + ".weight"] = {256, 256, 2, 2};
  const int dims[3] = {512, 1024, 1024};
  for (int i = 0; i < 3; ++i) {
    const std::string n = "encoder.upsample" + std::to_string(i);
    m[n + ".0.weight"] = {dims[i], EMB, 1, 1};
    m[n + ".1.weight"] = {dims[i], dims[i], 2, 2};
  }
  m["encoder.upsample_lowres.weight"] = {EMB, 1024, 2, 2};
  m["encoder.upsample_lowres.bias"] = {1024};
  m["encoder.fuse_lowres.weight"] = {1024, 2048, 1, 1};
  m["encoder.fuse_lowres.bias"] = {1024};
  const int cd[5] = {0, 256, 512, 1024, 1024};
  for (int i = 1; i <= 4; ++i) m["decoder.convs." + std::to_string(i) + ".weight"] = {256, cd[i], 3, 3};
  for (int f = 0; f < 5; ++f) {
    const std::string p = "decoder.fusions." + std::to_string(f) + ".";
    for (const char* rn : {"resnet1", "resnet2"})
      for (const char* c : {"1", "3"}) {
        m[p + rn + ".residual." + c + ".weight"] = {256, 256, 3, 3};
        m[p + rn + ".residual." + c + ".bias"] = {256};
      }
    if (f != 0) m[p + "deconv.weight"] = {256, 256, 2, 2};
    m[p + "out_conv.weight"] = {256, 256, 1, 1};
    m[p + "out_conv.bias"] = {256};
  }
  m["head.0.weight"] = {128, 256, 3, 3};
  m["head.0.bias"] = {128};
  m["head.1.weight"] = {128, 128, 2, 2};
  m["head.1.bias"] = {128};
  m["head.2.weight"] = {32, 128, 3, 3};
  m["head.2.bias"] = {32};
  m["head.4.weight"] = {1, 32, 1, 1};
  m["head.4.bias"] = {1};
  if (fov_mode == 2) {  // fov.py:47-54: encoder + Linear, downsample = fov_head0, head = the remaining three convs
    m["fov.encoder.1.weight"] = {128, EMB};
    m["fov.encoder.1.bias"] = {128};
    m["fov.downsample.0.weight"] = {128, 256, 3, 3};
    m["fov.downsample.0.bias"] = {128};
    m["fov.head.0.weight"] = {64, 128, 3, 3};
    m["fov.head.0.bias"] = {64};
    m["fov.head.2.weight"] = {32, 64, 3, 3};
    m["fov.head.2.bias"] = {32};
    m["fov.head.4.weight"] = {1, 32, 6, 6};
    m["fov.head.4.bias"] = {1};
  } else if (fov_mode == 1) {  // fov.py:55-56: head = fov_head0 + fov_head (Sequential indices 0, 2, 4, 6)
    m["fov.head.0.weight"] = {128, 256, 3, 3};
    m["fov.head.0.bias"] = {128};
    m["fov.head.2.weight"] = {64, 128, 3, 3};
    m["fov.head.2.bias"] = {64};
    m["fov.head.4.weight"] = {32, 64, 3, 3};
    m["fov.head.4.bias"] = {32};
    m["fov.head.6.weight"] = {1, 32, 6, 6};
    m["fov.head.6.bias"] = {1};
  }
  return m;
}

bool ends_with(const std::string& s, const std::string& suf) {
  return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}
bool starts_with(const std::string& s, const std::string& pre) { return s.compare(0, pre.size(), pre) == 0; }

// ConvTranspose2d weights are (Cin, Cout, 2, 2)
bool is_convT(const std::string& n) {
  if (n == "encoder.upsample_lowres.weight" || n == "head.1.weight") return true;
  if (ends_with(n, "deconv.weight")) return true;
  if (starts_with(n, "encoder.upsample") && ends_with(n, ".weight")) {
    const size_t e = n.size() - 7;              // position of ".weight"
    const size_t d = n.rfind('.', e - 1);       // dot before the layer index
    const std::string idx = n.substr(d + 1, e - d - 1);
    return !idx.empty() && isdigit(idx[0]) && std::stoi(idx) >= 1;
  }
  return false;
}

}  // namespace

// ============================================================================ lifecycle
Engine::Engine(int device, int prec, int max_batch, int fov_mode)
    : device_(device), prec_(prec), max_batch_(max_batch), fov_mode_(fov_mode), n_enc_(fov_mode == 2 ? 3 : 2) {
  DP_CHECK(prec == FP32 || prec == BF16, "precision must be 0 (fp32) or 1 (bf16)");
  DP_CHECK(fov_mode >= 0 && fov_mode <= 2, "fov_mode must be 0 (no FOV head), 1 (head without encoder) or 2 (default)");
  DP_CHECK(max_batch >= 1 && max_batch <= 64, "max_batch out of range");
  int count = 0;
  DP_CUDA(cudaGetDeviceCount(&count));
  DP_CHECK(device >= 0 && device < count, "no such CUDA device (this engine has no CPU fallback)");
  DP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  DP_CUDA(cudaGetDeviceProperties(&prop, device));
  DP_CHECK(prop.major == 10, "depthpro_b200 is built for sm_100a (Blackwell B200) only; found sm_" +
                                 std::to_string(prop.major) + std::to_string(prop.minor));
  manifest_ = build_manifest(fov_mode_);
  const char* a = getenv("DEPTHPRO_ATTN");  // debugging switch: "mma" selects the mma.sync kernel
  attn_legacy_ = a != nullptr && std::string(a) == "mma";
  const char* lf = getenv("DEPTHPRO_LN_FUSE");  // debugging switch: "0" keeps the stand-alone LayerNorm launches
  ln_fuse_ = prec_ == BF16 && !(lf != nullptr && lf[0] == '0');
  // the ViT residual stream as a (hi, lo) pair of 16-bit arrays (common.cuh GemmOp::ln_xlo) instead of fp32 + a 16-bit
  // copy; needs the folded LayerNorm.  "0" keeps the fp32 stream (A/B, debugging)
  // "0": every kernel walks its tiles first to last; "1": alternating direction in the ViT only; default: ViT + decoder
  const char* sp = getenv("DEPTHPRO_SERPENTINE");
  serpentine_ = prec_ == BF16 && !(sp != nullptr && sp[0] == '0');
  serpentine_dec_ = serpentine_ && !(sp != nullptr && sp[0] == '1');
  const char* hf = getenv("DEPTHPRO_HEAD0_FUSE");  // "0": keep fusions.0.out_conv as its own 1x1 launch (A/B, debugging)
  head0_fused_ = prec_ == BF16 && !(hf != nullptr && hf[0] == '0');
  const char* rp = getenv("DEPTHPRO_RES_PAIR");
  res_pair_ = ln_fuse_ && !(rp != nullptr && rp[0] == '0');
  DP_CUDA(cudaStreamCreateWithFlags(&host_stream_, cudaStreamNonBlocking));
  const char* ms = getenv("DEPTHPRO_STREAMS");  // "0": no side streams inside a frame
  multi_stream_ = !(ms != nullptr && ms[0] == '0');
  const char* gr = getenv("DEPTHPRO_GRAPH");    // "0": no CUDA-graph replay of the forward pass
  use_graph_ = !(gr != nullptr && gr[0] == '0');
  DP_CUDA(cudaStreamCreateWithFlags(&cap_stream_, cudaStreamNonBlocking));
  for (int i = 0; i < NSIDE; ++i) {
    DP_CUDA(cudaStreamCreateWithFlags(&side_[i], cudaStreamNonBlocking));
    DP_CUDA(cudaEventCreateWithFlags(&ev_side_[i], cudaEventDisableTiming));
  }
  DP_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
  DP_CUDA(cudaEventCreateWithFlags(&ev_lowres_, cudaEventDisableTiming));
}

void Engine::drop_graphs() {
  for (auto& kv : graphs_)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  graphs_.clear();
}

Engine::~Engine() {
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  drop_graphs();
  for (int i = 0; i < NSIDE; ++i) {
    if (side_[i]) cudaStreamDestroy(side_[i]);
    if (ev_side_[i]) cudaEventDestroy(ev_side_[i]);
  }
  if (ev_fork_) cudaEventDestroy(ev_fork_);
  if (ev_lowres_) cudaEventDestroy(ev_lowres_);
  if (cap_stream_) cudaStreamDestroy(cap_stream_);
  for (void* p : allocs_) cudaFree(p);
  for (auto& kv : packed_) cudaFree(kv.second.ptr);
  for (auto& kv : raw_) cudaFree(kv.second.ptr);
  if (stage_) cudaFree(stage_);
  if (himg_) cudaFree(himg_);
  if (hdepth_) cudaFree(hdepth_);
  if (host_stream_) cudaStreamDestroy(host_stream_);
  tmap_cache_clear();
}

void* Engine::alloc(size_t bytes) {
  void* p = nullptr;
  DP_CUDA(cudaMalloc(&p, (bytes + 255) & ~size_t(255)));
  allocs_.push_back(p);
  return p;
}

void Engine::set_weight(const std::string& name, const void* data, const int64_t* shape, int ndim, bool on_device) {
  DP_CUDA(cudaSetDevice(device_));
  auto it = manifest_.find(name);
  if (it == manifest_.end()) throw Error("unexpected key in state_dict: " + name);
  weights_changed_ = true;
  const auto& ms = it->second;
  DP_CHECK(static_cast<int>(ms.size()) == ndim, "rank mismatch for " + name);
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    DP_CHECK(ms[i] == shape[i], "shape mismatch for " + name);
    n *= static_cast<size_t>(shape[i]);
  }
  cudaStream_t s = nullptr;
  const float* src = reinterpret_cast<const float*>(data);
  if (!on_device) {
    if (stage_bytes_ < n * 4) {
      if (stage_) DP_CUDA(cudaFree(stage_));
      DP_CUDA(cudaMalloc(&stage_, n * 4));
      stage_bytes_ = n * 4;
    }
    DP_CUDA(cudaMemcpy(stage_, data, n * 4, cudaMemcpyHostToDevice));
    src = stage_;
  }
  const bool bf = prec_ == BF16;
  const bool fusion_tail = starts_with(name, "decoder.fusions.") && name[16] != '0' &&
                           (ends_with(name, ".deconv.weight") || ends_with(name, ".out_conv.weight"));
  // folded at finalize with the block's norm1 / norm2 affine, from the fp32 originals
  const bool ln_folded = ln_fuse_ && (ends_with(name, ".attn.qkv.weight") || ends_with(name, ".mlp.fc1.weight") ||
                                      ends_with(name, ".attn.qkv.bias") || ends_with(name, ".mlp.fc1.bias"));
  // fusions.0.out_conv (1x1) is composed into head.0 (conv3x3) at finalize
  const bool head0_chain = name == "head.0.weight" || name == "head.0.bias" || name == "decoder.fusions.0.out_conv.weight" ||
                           name == "decoder.fusions.0.out_conv.bias";
  if (bf && (name == "head.1.weight" || name == "head.1.bias" || name == "head.2.weight" || name == "head.2.bias" ||
             fusion_tail || ln_folded || head0_chain)) {
    Packed& rw = raw_[name];
    if (rw.bytes != n * 4) {
      if (rw.ptr) DP_CUDA(cudaFree(rw.ptr));
      DP_CUDA(cudaMalloc(&rw.ptr, n * 4));
      rw.bytes = n * 4;
    }
    DP_CUDA(cudaMemcpy(rw.ptr, src, n * 4, cudaMemcpyDeviceToDevice));
  }
  Packed& pk = packed_[name];
  auto ensure = [&](size_t bytes) {
    if (pk.bytes != bytes) {
      if (pk.ptr) DP_CUDA(cudaFree(pk.ptr));
      DP_CUDA(cudaMalloc(&pk.ptr, bytes));
      pk.bytes = bytes;
    }
  };
  // the 6x6 valid conv that ends the FOV head, reduced by fov_final (fp32 HWIO)
  const bool fov_conv = name == (fov_mode_ == 1 ? "fov.head.6.weight" : "fov.head.4.weight") && fov_mode_ != 0;
  if (ndim == 4 && is_convT(name)) {
    ensure(n * esz());
    if (bf) pack_convT_iohw<bf16>(src, reinterpret_cast<bf16*>(pk.ptr), (int)shape[0], (int)shape[1], s);
    else pack_convT_iohw<float>(src, reinterpret_cast<float*>(pk.ptr), (int)shape[0], (int)shape[1], s);
  } else if (fov_conv) {
    ensure(n * 4);
    pack_oihw_to_hwio_f32(src, reinterpret_cast<float*>(pk.ptr), (int)shape[0], (int)shape[1], (int)shape[2], (int)shape[3], s);
  } else if (ndim == 4 && shape[2] == 3 && name != "head.4.weight") {
    ensure(n * esz());
    if (bf) pack_oihw_to_ohwi<bf16>(src, reinterpret_cast<bf16*>(pk.ptr), (int)shape[0], (int)shape[1], 3, 3, s);
    else pack_oihw_to_ohwi<float>(src, reinterpret_cast<float*>(pk.ptr), (int)shape[0], (int)shape[1], 3, 3, s);
  } else if ((ndim == 2 || ndim == 4) && name != "head.4.weight") {
    // Linear (N,K); 1x1 conv (O,I,1,1); patch-embed conv (O, 3*16*16): already [N, K] K-major
    ensure(n * esz());
    if (bf) convert<float, bf16>(src, reinterpret_cast<bf16*>(pk.ptr), (long long)n, s);
    else convert<float, float>(src, reinterpret_cast<float*>(pk.ptr), (long long)n, s);
  } else {
    ensure(n * 4);
    convert<float, float>(src, reinterpret_cast<float*>(pk.ptr), (long long)n, s);
  }
  DP_CUDA(cudaStreamSynchronize(s));
}

int Engine::missing_weights() const {
  int miss = 0;
  for (auto& kv : manifest_)
    if (!packed_.count(kv.first)) ++miss;
  return miss;
}

const void* Engine::W(const std::string& name) const {
  auto it = packed_.find(name);
  if (it == packed_.end()) throw Error("missing key in state_dict: " + name);
  return it->second.ptr;
}
const float* Engine::F(const std::string& name) const { return reinterpret_cast<const float*>(W(name)); }

VitWeights Engine::vit_weights(const std::string& p) const {
  VitWeights w;
  w.cls = F(p + "cls_token");
  w.pos = F(p + "pos_embed");
  w.pe_w = W(p + "patch_embed.proj.weight");
  w.pe_b = F(p + "patch_embed.proj.bias");
  w.norm_w = F(p + "norm.weight");
  w.norm_b = F(p + "norm.bias");
  for (int i = 0; i < 24; ++i) {
    const std::string b = p + "blocks." + std::to_string(i) + ".";
    auto& k = w.blk[i];
    k.n1w = F(b + "norm1.weight"), k.n1b = F(b + "norm1.bias");
    k.qkv_w = W(b + "attn.qkv.weight"), k.qkv_b = F(b + "attn.qkv.bias");
    k.proj_w = W(b + "attn.proj.weight"), k.proj_b = F(b + "attn.proj.bias");
    k.g1 = F(b + "ls1.gamma");
    k.n2w = F(b + "norm2.weight"), k.n2b = F(b + "norm2.bias");
    k.fc1_w = W(b + "mlp.fc1.weight"), k.fc1_b = F(b + "mlp.fc1.bias");
    k.fc2_w = W(b + "mlp.fc2.weight"), k.fc2_b = F(b + "mlp.fc2.bias");
    k.g2 = F(b + "ls2.gamma");
    if (ln_fuse_) {
      k.qkv_b = F(b + "attn.qkv.ln_d"), k.qkv_c = F(b + "attn.qkv.ln_c");
      k.fc1_b = F(b + "mlp.fc1.ln_d"), k.fc1_c = F(b + "mlp.fc1.ln_c");
    }
  }
  return w;
}

void Engine::finalize() {
  DP_CUDA(cudaSetDevice(device_));
  if (finalized_ && !weights_changed_) return;  // nothing was handed over since the last finalize: no-op
  DP_CUDA(cudaDeviceSynchronize());
  drop_graphs();  // captured launches hold weight pointers and composed weights
  if (missing_weights() != 0) {
    for (auto& kv : manifest_)
      if (!packed_.count(kv.first)) throw Error("missing key in state_dict: " + kv.first);
  }
  if (ln_fuse_) {
    // LN(x) W^T + b = rstd (x (g*W)^T) - rstd mean colsum(g*W) + (W b_ln + b): fold norm1 into qkv and
    // norm2 into fc1, in fp32 from the fp32 originals, rounded to bf16 once.  The fp32 copies are dropped.
    auto slot = [&](const std::string& k, size_t bytes) {
      Packed& pk = packed_[k];
      if (pk.bytes != bytes) {
        if (pk.ptr) DP_CUDA(cudaFree(pk.ptr));
        DP_CUDA(cudaMalloc(&pk.ptr, bytes));
        pk.bytes = bytes;
      }
      return pk.ptr;
    };
    for (const char* enc : {"encoder.patch_encoder.", "encoder.image_encoder.", "fov.encoder.0."}) {
      if (fov_mode_ != 2 && std::string(enc) == "fov.encoder.0.") continue;
      for (int i = 0; i < 24; ++i) {
        const std::string b = std::string(enc) + "blocks." + std::to_string(i) + ".";
        const std::pair<const char*, const char*> pairs[2] = {{"attn.qkv", "norm1"}, {"mlp.fc1", "norm2"}};
        for (auto& pr : pairs) {
          const std::string lin = b + pr.first, nrm = b + pr.second;
          auto rw = raw_.find(lin + ".weight"), rb = raw_.find(lin + ".bias");
          if (rw == raw_.end() || rb == raw_.end())
            throw Error("LayerNorm folding needs " + lin + ".weight / .bias set again together with " + nrm +
                        ".weight / .bias before dp_engine_finalize (the folded copy replaced the original)");
          const int N = static_cast<int>(rb->second.bytes / 4);
          ln_fold((const float*)rw->second.ptr, F(nrm + ".weight"), F(nrm + ".bias"), (const float*)rb->second.ptr,
                  (bf16*)packed_.at(lin + ".weight").ptr, (float*)slot(lin + ".ln_c", N * 4), (float*)slot(lin + ".ln_d", N * 4),
                  N, EMB, nullptr);
        }
      }
    }
    DP_CUDA(cudaStreamSynchronize(nullptr));
    for (auto it = raw_.begin(); it != raw_.end();) {
      if (ends_with(it->first, ".attn.qkv.weight") || ends_with(it->first, ".mlp.fc1.weight") ||
          ends_with(it->first, ".attn.qkv.bias") || ends_with(it->first, ".mlp.fc1.bias")) {
        cudaFree(it->second.ptr);
        it = raw_.erase(it);
      } else {
        ++it;
      }
    }
  }
  vit_patch_ = vit_weights("encoder.patch_encoder.");
  vit_image_ = vit_weights("encoder.image_encoder.");
  if (fov_mode_ == 2) vit_fov_ = vit_weights("fov.encoder.0.");
  if (prec_ == BF16) {
    // exact-linear fusion head.1 o head.2 (no nonlinearity in between, depth_pro.py:182-201),
    // composed in fp32 from the fp32 originals, then rounded to bf16 once
    if (!head_wc_) head_wc_ = alloc(128 * 1152 * 2), head_cb_ = (float*)alloc(10 * 32 * 4);
    auto R = [&](const char* k) { return reinterpret_cast<const float*>(raw_.at(k).ptr); };
    compose_head(R("head.1.weight"), R("head.1.bias"), R("head.2.weight"), R("head.2.bias"), (bf16*)head_wc_, head_cb_,
                 nullptr);
    // exact-linear fusion fusions.0.out_conv (1x1 + bias, decoder.py:178 with deconv=False) o head.0 (conv3x3,
    // depth_pro.py:183-185): the decoder's 768^2 x 256 output map is never written or re-read (604 MB per frame)
    if (!head0_wc_) head0_wc_ = alloc(128 * 9 * 256 * 2), head0_cb_ = (float*)alloc(10 * 128 * 4);
    compose_1x1_conv3x3(R("decoder.fusions.0.out_conv.weight"), R("decoder.fusions.0.out_conv.bias"), R("head.0.weight"),
                        R("head.0.bias"), (bf16*)head0_wc_, head0_cb_, 128, 256, nullptr);
    // exact-linear fusion deconv (no bias) o out_conv (1x1 + bias) of fusion blocks 1-4 (decoder.py:176-178)
    for (int i = 1; i <= 4; ++i) {
      const std::string p = "decoder.fusions." + std::to_string(i) + ".";
      Packed& pk = packed_[p + "deconv_out.weight"];
      if (!pk.ptr) {
        DP_CUDA(cudaMalloc(&pk.ptr, 4 * 256 * 256 * 2));
        pk.bytes = 4 * 256 * 256 * 2;
      }
      compose_deconv_1x1(R((p + "deconv.weight").c_str()), R((p + "out_conv.weight").c_str()), (bf16*)pk.ptr, 256, nullptr);
    }
    DP_CUDA(cudaStreamSynchronize(nullptr));
  }
  weights_changed_ = false;
  if (finalized_) return;  // workspace already allocated; weights re-bound above

  const size_t e = esz();
  const size_t MB = static_cast<size_t>(max_batch_);
  const size_t T = MB * seqs_per_frame() * SEQ;  // 35 patch + 1 image (+ 1 fov) sequences per frame
  xbuf_ = (float*)alloc(MB * 3 * IMG * IMG * 4);
  canon_ = (float*)alloc(MB * IMG * IMG * 4);
  fov_ = (float*)alloc(MB * 4);
  fpx_ = (float*)alloc(MB * 4);
  fpx_in_ = (float*)alloc(MB * 4);
  A35_ = alloc(MB * 36 * 576 * 768 * e);
  resid_ = (float*)alloc(T * EMB * 4);
  xn_ = alloc(T * EMB * e);
  if (ln_fuse_) ln_stats_ = (float*)alloc(T * LN_SLOTS * 2 * 4);
  if (res_pair_) xlo_ = alloc(T * EMB * e);
  qkv_ = alloc(T * 3 * EMB * e);
  attn_ = alloc(T * EMB * e);
  hid_ = alloc(T * 4 * EMB * e);
  lat0m_ = alloc(MB * 96 * 96 * EMB * e);
  lat1m_ = alloc(MB * 96 * 96 * EMB * e);
  x0m_ = alloc(MB * 96 * 96 * EMB * e);
  x1m_ = alloc(MB * 48 * 48 * EMB * e);
  x2m_ = alloc(MB * 24 * 24 * EMB * e);
  globm_ = alloc(MB * 24 * 24 * EMB * e);
  fovtok_ = alloc(MB * 24 * 24 * EMB * e);
  // per-frame decoder workspace
  const size_t P96 = 96 * 96, P192 = 192 * 192, P384 = 384 * 384, P768 = 768 * 768, P48 = 48 * 48, P24 = 24 * 24;
  u0a_ = alloc(P96 * 256 * e), u0b_ = alloc(P192 * 256 * e), u0c_ = alloc(P384 * 256 * e);
  enc0_ = alloc(P768 * 256 * e), enc0r_ = alloc(P768 * 256 * e);
  u1a_ = alloc(P96 * 256 * e), u1b_ = alloc(P192 * 256 * e), enc1_ = alloc(P384 * 256 * e);
  u2a_ = alloc(P96 * 512 * e), enc2_ = alloc(P192 * 512 * e);
  u3a_ = alloc(P48 * 1024 * e), enc3_ = alloc(P96 * 1024 * e);
  u4a_ = alloc(P24 * 1024 * e), cat_ = alloc(P48 * 2048 * e), enc4_ = alloc(P48 * 1024 * e);
  lowres_ = alloc(P48 * 256 * e), lowres_r_ = alloc(P48 * 256 * e);
  x1_ = nullptr, x1r_ = nullptr, t_ = alloc(P768 * 256 * e);
  c1_ = alloc(P384 * 256 * e), c1r_ = alloc(P384 * 256 * e), c2_ = alloc(P192 * 256 * e), c2r_ = alloc(P192 * 256 * e);
  c3_ = alloc(P96 * 256 * e), c3r_ = alloc(P96 * 256 * e);
  x_ = alloc(P768 * 256 * e), xr_ = alloc(P768 * 256 * e), x2_ = alloc(P768 * 256 * e), y_ = alloc(P768 * 256 * e);
  const size_t fs[5] = {P768, P768, P384, P192, P96};  // feat_[i] = output of fusion i
  for (int i = 0; i < 5; ++i) feat_[i] = alloc(fs[i] * 256 * e);
  h0_ = alloc(P768 * 128 * e);
  h1_ = prec_ == BF16 ? nullptr : alloc(static_cast<size_t>(IMG) * IMG * 128 * e);  // bf16 mode fuses head.1 into head.2
  fovcol_ = alloc(P24 * 2304 * e);
  fovlin_ = alloc(P24 * 128 * e), fov_a_ = alloc(P24 * 128 * e), fov_b_ = alloc(12 * 12 * 64 * e),
  fov_c_ = alloc(6 * 6 * 32 * e);
  colorize_mm_ = (float*)alloc(colorize_scratch_bytes());
  finalized_ = true;
}

// ============================================================================ small entry points
void Engine::preprocess(const void* img, int B, int H, int W, int src_fmt, float* x, int interp, cudaStream_t s) {
  DP_CHECK(B >= 1 && H >= 1 && W >= 1, "bad image shape");
  DP_CHECK(src_fmt == 0 || src_fmt == 1, "bad src_fmt");
  DP_CHECK(interp == INTERP_BILINEAR || interp == INTERP_BICUBIC, "interpolation mode must be 0 (bilinear) or 1 (bicubic)");
  resize_to_1536(img, src_fmt, B, H, W, x, interp, s);
}

void Engine::split(const float* x, int B, float* patches, cudaStream_t s) {
  float* tmp = nullptr;
  const size_t n = static_cast<size_t>(B) * 35 * 576 * 768;
  DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&tmp), n * 4, s));
  split_im2col<float>(x, B, tmp, nullptr, s);
  im2col_to_ref_patches(tmp, B, patches, s);
  DP_CUDA(cudaFreeAsync(tmp, s));
}

void Engine::merge(const float* tokens, int B, int steps, int padding, int C, float* merged, cudaStream_t s) {
  DP_CHECK(steps >= 1 && padding >= 0 && 2 * padding < 24, "bad merge geometry");
  RowMap m;
  m.mode = 1, m.steps = steps, m.pad = padding, m.patch_base = 0;
  m.S = steps == 1 ? 24 : (24 - padding) * 2 + (24 - 2 * padding) * (steps - 2);
  m.sb = 1, m.sp = B;  // reference order: patch-major, batch-minor (encoder.py:200)
  float* tmp = nullptr;
  const size_t n = static_cast<size_t>(B) * m.S * m.S * C;
  DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&tmp), n * 4, s));
  // gather works on rows of width C with 577 tokens per sequence
  merge_rows_f32(tokens, tmp, B, C, m, s);
  nhwc_to_nchw_f32<float>(tmp, merged, B, m.S, m.S, C, s);
  DP_CUDA(cudaFreeAsync(tmp, s));
}

// ============================================================================ ViT
// The three DINOv2 ViT-L/16 encoders (patch: 35*B sequences, image: B, fov: B) have identical
// structure and different weights, so they run layer by layer as ONE grouped launch per op:
// rows [0, 35B*577) use the patch-encoder weights, the next B*577 the image encoder's, the last
// B*577 the fov encoder's.  The small encoders ride along for +5.7% tiles instead of 2 x 170
// latency-bound launches.
template <typename T>
void Engine::run_vits(int B, cudaStream_t s) {
  const VitWeights* vw[3] = {&vit_patch_, &vit_image_, &vit_fov_};
  const int NG = n_enc_;  // 3 encoders, or 2 without the fov encoder
  const int nseq_g[3] = {35 * B, B, B};
  const int nseq = seqs_per_frame() * B;
  const long long M = static_cast<long long>(nseq) * SEQ;
  T* xn = (T*)xn_;
  T* qkv = (T*)qkv_;
  T* attn = (T*)attn_;
  T* hid = (T*)hid_;
  float* resid = resid_;

  auto grouped = [&](GemmOp& op, int rows_per_seq, bool a_shared_small) {
    op.ngroups = NG;
    long long off = 0, aoff = 0;
    for (int g = 0; g < NG; ++g) {
      op.grp[g].M = nseq_g[g] * rows_per_seq;
      op.grp[g].o_row_off = off;
      op.grp[g].a_row_off = aoff;
      off += op.grp[g].M;
      // patch embed: the image and fov encoders read the SAME im2col rows (patch 34 of each frame)
      if (!(a_shared_small && g == 1)) aoff += op.grp[g].M;
    }
    op.M = static_cast<int>(off);
    op.a_rows = a_shared_small ? static_cast<long long>(36) * B * rows_per_seq : off;
  };

  NvtxRange r_vit("vit (3 encoders, grouped)");
  {  // patch embed (timm PatchEmbed conv k16 s16 as GEMM) + cls + pos_embed
    NvtxRange r("vit.patch_embed");
    GemmOp op;
    op.N = EMB, op.K = 768, op.A = A35_, op.lda = 768;
    op.out = resid, op.out_f32 = 1, op.out_mode = O_PATCH_EMBED, op.ldo = EMB;
    grouped(op, 576, true);
    for (int g = 0; g < NG; ++g) op.grp[g].Wt = vw[g]->pe_w, op.grp[g].bias = vw[g]->pe_b, op.grp[g].pos = vw[g]->pos;
    gemm(prec_, op, s);
    int seq0 = 0;
    for (int g = 0; g < NG; ++g) {
      write_cls_rows(resid + static_cast<long long>(seq0) * SEQ * EMB, vw[g]->cls, vw[g]->pos, nseq_g[g], s);
      seq0 += nseq_g[g];
    }
  }
  LnGroups lg;
  lg.n = NG;
  lg.end[0] = 35LL * B * SEQ, lg.end[1] = 36LL * B * SEQ, lg.end[2] = M;
  // bf16 mode: norm1 / norm2 are folded into qkv / fc1 (common.cuh GemmOp::ln_stats); xn holds the RAW
  // bf16 residual stream, written with its row statistics by this one launch for layer 0 and by the
  // proj / fc2 epilogues from then on
  const bool fuse = ln_fuse_ && std::is_same<T, bf16>::value;
  // pair form: from here on the stream is (xn, xlo) = (hi, lo); the fp32 array only holds the patch-embed output
  const bool pair = fuse && res_pair_;
  bf16* xlo = pair ? (bf16*)xlo_ : nullptr;
  if (fuse) ln_stats_cast(resid, (bf16*)xn, ln_stats_, M, s, xlo);
  // Serpentine tile order (GemmOp::reverse): every kernel of the chain walks its rows in the direction opposite to its
  // producer's, so that it starts on what the L2 still holds.  The first consumer follows an ascending elementwise pass.
  int rev = serpentine_ ? 1 : 0;
  auto next_dir = [&]() {
    const int r = rev;
    if (serpentine_) rev ^= 1;
    return r;
  };
  for (int i = 0; i < 24; ++i) {
    NvtxRange r_blk("vit.block");
    if (!fuse) {
      ProfScope ps(s, KC_LAYERNORM, static_cast<double>(M) * EMB * (4 + sizeof(T)));
      for (int g = 0; g < NG; ++g) lg.w[g] = vw[g]->blk[i].n1w, lg.b[g] = vw[g]->blk[i].n1b;
      layernorm_rows_grouped<T>(resid, xn, lg, M, s);
    }
    {
      GemmOp op;
      op.N = 3 * EMB, op.K = EMB, op.A = xn, op.lda = EMB, op.out = qkv, op.ldo = 3 * EMB;
      op.reverse = next_dir();
      grouped(op, SEQ, false);
      for (int g = 0; g < NG; ++g) op.grp[g].Wt = vw[g]->blk[i].qkv_w, op.grp[g].bias = vw[g]->blk[i].qkv_b;
      if (fuse) {
        op.ln_stats = ln_stats_;
        for (int g = 0; g < NG; ++g) op.grp[g].ln_c = vw[g]->blk[i].qkv_c;
      }
      gemm(prec_, op, s);
    }
    {
      ProfScope ps(s, KC_ATTENTION, 4.0 * SEQ * SEQ * 64 * 16 * nseq);
      if (prec_ == BF16) {
        if (attn_legacy_) attention_bf16((const bf16*)qkv, (bf16*)attn, nseq, s);
        else attention_bf16_tc((const bf16*)qkv, (bf16*)attn, nseq, s, next_dir());
      } else {
        attention_f32((const float*)qkv, (float*)attn, nseq, s);
      }
    }
    {
      GemmOp op;
      op.N = EMB, op.K = EMB, op.A = attn, op.lda = EMB, op.ldo = EMB;
      op.reverse = next_dir();
      if (!pair) op.res = resid, op.res_f32 = 1, op.ldres = EMB, op.out = resid, op.out_f32 = 1;
      grouped(op, SEQ, false);
      for (int g = 0; g < NG; ++g)
        op.grp[g].Wt = vw[g]->blk[i].proj_w, op.grp[g].bias = vw[g]->blk[i].proj_b, op.grp[g].gamma = vw[g]->blk[i].g1;
      if (fuse) op.ln_xb = xn, op.ln_stats_out = ln_stats_, op.ln_xlo = xlo;
      gemm(prec_, op, s);
    }
    if (!fuse) {
      ProfScope ps(s, KC_LAYERNORM, static_cast<double>(M) * EMB * (4 + sizeof(T)));
      for (int g = 0; g < NG; ++g) lg.w[g] = vw[g]->blk[i].n2w, lg.b[g] = vw[g]->blk[i].n2b;
      layernorm_rows_grouped<T>(resid, xn, lg, M, s);
    }
    {
      GemmOp op;
      op.N = 4 * EMB, op.K = EMB, op.A = xn, op.lda = EMB, op.act = ACT_GELU, op.out = hid, op.ldo = 4 * EMB;
      op.reverse = next_dir();
      grouped(op, SEQ, false);
      for (int g = 0; g < NG; ++g) op.grp[g].Wt = vw[g]->blk[i].fc1_w, op.grp[g].bias = vw[g]->blk[i].fc1_b;
      if (fuse) {
        op.ln_stats = ln_stats_;
        for (int g = 0; g < NG; ++g) op.grp[g].ln_c = vw[g]->blk[i].fc1_c;
      }
      gemm(prec_, op, s);
    }
    {
      GemmOp op;
      op.N = EMB, op.K = 4 * EMB, op.A = hid, op.lda = 4 * EMB, op.ldo = EMB;
      op.reverse = next_dir();
      if (!pair) op.res = resid, op.res_f32 = 1, op.ldres = EMB, op.out = resid, op.out_f32 = 1;
      grouped(op, SEQ, false);
      for (int g = 0; g < NG; ++g)
        op.grp[g].Wt = vw[g]->blk[i].fc2_w, op.grp[g].bias = vw[g]->blk[i].fc2_b, op.grp[g].gamma = vw[g]->blk[i].g2;
      // (the pair form has no fp32 output: the last block writes (hi, lo) + unused row sums like every other)
      if (fuse && (i < 23 || pair)) op.ln_xb = xn, op.ln_stats_out = ln_stats_, op.ln_xlo = xlo;
      gemm(prec_, op, s);
    }
    if (i == 5 || i == 11) {
      // forward hooks on blocks 5 / 11 of the PATCH encoder (encoder.py:133-144): pre-norm residual
      // stream of the 25 level-0 patches, cls dropped, merged with padding 3 (encoder.py:268-289)
      RowMap m;
      m.mode = 1, m.S = 96, m.steps = 5, m.pad = 3, m.patch_base = 0;
      layernorm_rows<T>(resid, (T*)(i == 5 ? lat0m_ : lat1m_), nullptr, nullptr, (long long)B * 96 * 96, m, 0, s,
                        pair ? (const bf16*)xn : nullptr, xlo);
    }
  }
}

// ============================================================================ forward
template <typename T>
void Engine::forward_impl(const float* x, int B, float* canon, float* fov_deg, cudaStream_t s) {
  {
    NvtxRange r("pyramid+split+im2col");
    split_im2col<T>(x, B, (T*)A35_, (T*)A35_ + static_cast<size_t>(B) * 35 * 576 * 768, s);
  }
  run_vits<T>(B, s);
  {  // final norms fused with the merges (encoder.py:267-305), one launch per consumer
    NvtxRange r("final norm + merge");
    const bool pair = res_pair_ && std::is_same<T, bf16>::value;  // the stream ended as (xn, xlo) = (hi, lo)
    const bf16 *hi = pair ? (const bf16*)xn_ : nullptr, *lo = pair ? (const bf16*)xlo_ : nullptr;
    RowMap m;
    m.mode = 1, m.S = 96, m.steps = 5, m.pad = 3, m.patch_base = 0;
    layernorm_rows<T>(resid_, (T*)x0m_, vit_patch_.norm_w, vit_patch_.norm_b, (long long)B * 96 * 96, m, 1, s, hi, lo);
    m.S = 48, m.steps = 3, m.pad = 6, m.patch_base = 25;
    layernorm_rows<T>(resid_, (T*)x1m_, vit_patch_.norm_w, vit_patch_.norm_b, (long long)B * 48 * 48, m, 1, s, hi, lo);
    m.S = 24, m.steps = 1, m.pad = 0, m.patch_base = 34;
    layernorm_rows<T>(resid_, (T*)x2m_, vit_patch_.norm_w, vit_patch_.norm_b, (long long)B * 24 * 24, m, 1, s, hi, lo);
    // image encoder (encoder.py:308-311) and fov encoder (fov.py:70-77): sequences 35B.. and 36B..
    RowMap ms;
    ms.mode = 1, ms.S = 24, ms.steps = 1, ms.pad = 0, ms.patch_base = 0, ms.sb = 1, ms.sp = 1;
    ms.seq_off = 35 * B;
    layernorm_rows<T>(resid_, (T*)globm_, vit_image_.norm_w, vit_image_.norm_b, (long long)B * 24 * 24, ms, 1, s, hi, lo);
    if (fov_mode_ == 2) {
      ms.seq_off = 36 * B;
      layernorm_rows<T>(resid_, (T*)fovtok_, vit_fov_.norm_w, vit_fov_.norm_b, (long long)B * 24 * 24, ms, 1, s, hi, lo);
    }
  }
  NvtxRange r_dec("decode (encoder upsample, decoder, head, fov)");
  for (int f = 0; f < B; ++f)
    decode_frame<T>(f, canon + static_cast<size_t>(f) * IMG * IMG, fov_deg ? fov_deg + f : nullptr, s);
  last_B_ = B;
}

template <typename T>
void Engine::decode_frame(int f, float* canon, float* fov_deg, cudaStream_t s) {
  auto frame = [&](void* base, size_t elems_per_frame) { return (T*)base + static_cast<size_t>(f) * elems_per_frame; };
  const T* lat0m = frame(lat0m_, 96 * 96 * EMB);
  const T* lat1m = frame(lat1m_, 96 * 96 * EMB);
  const T* x0m = frame(x0m_, 96 * 96 * EMB);
  const T* x1m = frame(x1m_, 48 * 48 * EMB);
  const T* x2m = frame(x2m_, 24 * 24 * EMB);
  const T* globm = frame(globm_, 24 * 24 * EMB);
  const T* fovtok = frame(fovtok_, 24 * 24 * EMB);

  // side streams: fork from s here, join into s where a result is consumed (and all of them before returning)
  // (not while bench.py's per-launch profiler is on: its CUDA-event brackets must time one kernel at a time)
  const bool ms = multi_stream_ && !prof_active();
  cudaStream_t sA = s, sB = s, sC = s, sD = s, sF = s;
  if (ms) {
    sA = side_[0], sB = side_[1], sC = side_[2], sD = side_[3], sF = side_[4];
    DP_CUDA(cudaEventRecord(ev_fork_, s));
    for (int i = 0; i < 4; ++i) DP_CUDA(cudaStreamWaitEvent(side_[i], ev_fork_, 0));
  }
  auto done = [&](int i) {  // the side stream's work so far is a dependency of whatever s runs next
    if (ms) {
      DP_CUDA(cudaEventRecord(ev_side_[i], side_[i]));
      DP_CUDA(cudaStreamWaitEvent(s, ev_side_[i], 0));
    }
  };
  // Serpentine tile order along the main chain (GemmOp::reverse, see run_vits): each launch on `s` walks its tiles in
  // the direction opposite to the previous one's.
  int dec_dir = 0;
  auto main_dir = [&](cudaStream_t st) {
    if (st != s || !serpentine_dec_) return 0;
    dec_dir ^= 1;
    return dec_dir;
  };
  auto conv1x1 = [&](cudaStream_t st, const void* in, int S, int Cin, const std::string& wname, int Cout, void* out,
                     const float* bias) {
    GemmOp op;
    op.reverse = main_dir(st);
    op.M = S * S, op.N = Cout, op.K = Cin, op.A = in, op.lda = Cin, op.Wt = W(wname), op.bias = bias;
    op.out = out, op.ldo = Cout;
    gemm(prec_, op, st);
  };
  // ConvTranspose2d k2 s2 on an SxS grid: GEMM with N = 4*Cout + pixel-shuffle scatter
  auto convT = [&](cudaStream_t st, const void* in, int S, int Cin, const std::string& wname, int Cout, void* out, int ldo,
                   int col_off, const float* bias, void* out_relu) {
    GemmOp op;
    op.reverse = main_dir(st);
    op.M = S * S, op.N = 4 * Cout, op.K = Cin, op.A = in, op.lda = Cin, op.Wt = W(wname);
    op.bias = bias, op.bias_mod = bias ? Cout : 0;
    op.B = 1, op.H = S, op.W = S, op.cout = Cout, op.out_mode = O_CONVT2X2;
    op.out = out, op.out_relu = out_relu, op.ldo = ldo, op.col_off = col_off;
    gemm(prec_, op, st);
  };
  auto conv3x3 = [&](cudaStream_t st, const void* in, int S, int Cin, const std::string& wname, int Cout, const float* bias,
                     int act, const void* res, const void* res2, void* out, void* out_relu) {
    GemmOp op;
    op.reverse = main_dir(st);
    op.M = S * S, op.N = Cout, op.K = 9 * Cin, op.A = in, op.a_mode = A_CONV3X3, op.B = 1, op.H = S, op.W = S, op.C = Cin;
    op.Wt = W(wname), op.bias = bias, op.act = act, op.res = res, op.res2 = res2, op.ldres = Cout;
    op.out = out, op.out_relu = out_relu, op.ldo = Cout;
    gemm(prec_, op, st);
  };

  // ---- encoder.py:314-324: project + upsample every level; decoder.py:80-88: the skip convs of levels 1-3
  // (convs[0] is Identity).  Branch i ends in what fusion i consumes.
  {
    NvtxRange r("encoder.upsample");
    conv1x1(sA, lat0m, 96, EMB, "encoder.upsample_latent0.0.weight", 256, u0a_, nullptr);
    convT(sA, u0a_, 96, 256, "encoder.upsample_latent0.1.weight", 256, u0b_, 256, 0, nullptr, nullptr);
    convT(sA, u0b_, 192, 256, "encoder.upsample_latent0.2.weight", 256, u0c_, 256, 0, nullptr, nullptr);
    convT(sA, u0c_, 384, 256, "encoder.upsample_latent0.3.weight", 256, enc0_, 256, 0, nullptr, enc0r_);
    conv1x1(sB, lat1m, 96, EMB, "encoder.upsample_latent1.0.weight", 256, u1a_, nullptr);
    convT(sB, u1a_, 96, 256, "encoder.upsample_latent1.1.weight", 256, u1b_, 256, 0, nullptr, nullptr);
    convT(sB, u1b_, 192, 256, "encoder.upsample_latent1.2.weight", 256, enc1_, 256, 0, nullptr, nullptr);
    conv3x3(sB, enc1_, 384, 256, "decoder.convs.1.weight", 256, nullptr, ACT_NONE, nullptr, nullptr, c1_, c1r_);
    conv1x1(sC, x0m, 96, EMB, "encoder.upsample0.0.weight", 512, u2a_, nullptr);
    convT(sC, u2a_, 96, 512, "encoder.upsample0.1.weight", 512, enc2_, 512, 0, nullptr, nullptr);
    conv3x3(sC, enc2_, 192, 512, "decoder.convs.2.weight", 256, nullptr, ACT_NONE, nullptr, nullptr, c2_, c2r_);
    conv1x1(sD, x1m, 48, EMB, "encoder.upsample1.0.weight", 1024, u3a_, nullptr);
    convT(sD, u3a_, 48, 1024, "encoder.upsample1.1.weight", 1024, enc3_, 1024, 0, nullptr, nullptr);
    conv3x3(sD, enc3_, 96, 1024, "decoder.convs.3.weight", 256, nullptr, ACT_NONE, nullptr, nullptr, c3_, c3r_);
    conv1x1(s, x2m, 24, EMB, "encoder.upsample2.0.weight", 1024, u4a_, nullptr);
    // torch.cat((x2_features, x_global_features), dim=1): both ConvT outputs land in one 2048-wide map
    convT(s, u4a_, 24, 1024, "encoder.upsample2.1.weight", 1024, cat_, 2048, 0, nullptr, nullptr);
    convT(s, globm, 24, EMB, "encoder.upsample_lowres.weight", 1024, cat_, 2048, 1024, F("encoder.upsample_lowres.bias"), nullptr);
    conv1x1(s, cat_, 48, 2048, "encoder.fuse_lowres.weight", 1024, enc4_, F("encoder.fuse_lowres.bias"));
  }

  // ---- decoder.py:74-93
  conv3x3(s, enc4_, 48, 1024, "decoder.convs.4.weight", 256, nullptr, ACT_NONE, nullptr, nullptr, lowres_, lowres_r_);
  if (ms && fov_mode_ != 0) {  // the FOV head only needs the low-resolution feature: it runs beside the decoder
    DP_CUDA(cudaEventRecord(ev_lowres_, s));
    DP_CUDA(cudaStreamWaitEvent(sF, ev_lowres_, 0));
  }
  // ---- FOV head (fov.py:56-82), on its own stream beside the decoder
  if (fov_mode_ != 0) {
    NvtxRange r("fov.head");
    cudaStream_t st = sF;
    if (fov_mode_ == 2) {
      GemmOp op;  // Linear 1024 -> 128 on the 576 non-cls tokens (cls is dropped at fov.py:77)
      op.M = 576, op.N = 128, op.K = EMB, op.A = fovtok, op.lda = EMB, op.Wt = W("fov.encoder.1.weight");
      op.bias = F("fov.encoder.1.bias"), op.out = fovlin_, op.ldo = 128;
      gemm(prec_, op, st);
    }
    // the three stride-2 3x3 convs (24^2, 12^2, 6^2 outputs) as im2col + tensor-core GEMM
    auto conv_s2 = [&](const void* in, int S, int Cin, const std::string& name, int Cout, const void* addend, void* out) {
      im2col_nhwc<T>((const T*)in, (T*)fovcol_, 1, S, S, Cin, 3, 2, 1, st);
      GemmOp op;
      op.M = (S / 2) * (S / 2), op.N = Cout, op.K = 9 * Cin, op.A = fovcol_, op.lda = 9 * Cin, op.Wt = W(name + ".weight");
      op.bias = F(name + ".bias"), op.act = ACT_RELU, op.res = addend, op.ldres = Cout, op.out = out, op.ldo = Cout;
      gemm(prec_, op, st);
    };
    if (fov_mode_ == 2) {
      conv_s2(lowres_, 48, 256, "fov.downsample.0", 128, fovlin_, fov_a_);  // relu(conv) + tokens (fov.py:78-79)
      conv_s2(fov_a_, 24, 128, "fov.head.0", 64, nullptr, fov_b_);
      conv_s2(fov_b_, 12, 64, "fov.head.2", 32, nullptr, fov_c_);
      fov_final<T>((const T*)fov_c_, F("fov.head.4.weight"), F("fov.head.4.bias"), fov_deg, 1, st);
    } else {  // fov_encoder_preset=None: head(lowres_feature) with head = fov_head0 + fov_head (fov.py:55-56, 80-82)
      conv_s2(lowres_, 48, 256, "fov.head.0", 128, nullptr, fov_a_);
      conv_s2(fov_a_, 24, 128, "fov.head.2", 64, nullptr, fov_b_);
      conv_s2(fov_b_, 12, 64, "fov.head.4", 32, nullptr, fov_c_);
      fov_final<T>((const T*)fov_c_, F("fov.head.6.weight"), F("fov.head.6.bias"), fov_deg, 1, st);
    }
  }
  const int S_of[5] = {768, 384, 192, 96, 48};
  const void* skip[5] = {enc0_, c1_, c2_, c3_, nullptr};      // x1 of fusion i (decoder.py:170): convs[i](encodings[i])
  const void* skip_r[5] = {enc0r_, c1r_, c2r_, c3r_, nullptr};
  const int branch_of[5] = {0, 1, 2, 3, -1};
  for (int i = 4; i >= 0; --i) {
    NvtxRange r("decoder.fusion");
    const int S = S_of[i];
    const std::string p = "decoder.fusions." + std::to_string(i) + ".";
    const void *xin, *xin_r;
    if (i == 4) {
      xin = lowres_, xin_r = lowres_r_;  // fusions[-1](features): no skip input, resnet1 unused (decoder.py:89)
    } else {
      done(branch_of[i]);
      const void *a = skip[i], *a_r = skip_r[i];
      // x = x0 + resnet1(x1)   (decoder.py:170-172, 111-118)
      conv3x3(s, a_r, S, 256, p + "resnet1.residual.1.weight", 256, F(p + "resnet1.residual.1.bias"), ACT_RELU, nullptr,
              nullptr, t_, nullptr);
      conv3x3(s, t_, S, 256, p + "resnet1.residual.3.weight", 256, F(p + "resnet1.residual.3.bias"), ACT_NONE, a,
              feat_[i + 1], x_, xr_);
      xin = x_, xin_r = xr_;
    }
    // x = resnet2(x)
    conv3x3(s, xin_r, S, 256, p + "resnet2.residual.1.weight", 256, F(p + "resnet2.residual.1.bias"), ACT_RELU, nullptr,
            nullptr, t_, nullptr);
    conv3x3(s, t_, S, 256, p + "resnet2.residual.3.weight", 256, F(p + "resnet2.residual.3.bias"), ACT_NONE, xin, nullptr,
            x2_, nullptr);
    if (i != 0 && prec_ == BF16) {
      // deconv o out_conv pre-composed into one ConvT (+bias): one GEMM and one full-resolution
      // round trip less per level
      convT(s, x2_, S, 256, p + "deconv_out.weight", 256, feat_[i], 256, 0, F(p + "out_conv.bias"), nullptr);
    } else if (i != 0) {
      convT(s, x2_, S, 256, p + "deconv.weight", 256, y_, 256, 0, nullptr, nullptr);
      conv1x1(s, y_, 2 * S, 256, p + "out_conv.weight", 256, feat_[i], F(p + "out_conv.bias"));
    } else if (prec_ != BF16 || !head0_fused_) {
      conv1x1(s, x2_, S, 256, p + "out_conv.weight", 256, feat_[0], F(p + "out_conv.bias"));
    }  // bf16: composed into head.0 below; tap("decoder_out") evaluates it on demand
  }

  // ---- depth head (depth_pro.py:182-204)
  NvtxRange r_head("head+fov");
  if (prec_ == BF16 && head0_fused_) {
    GemmOp op;
    op.M = 768 * 768, op.N = 128, op.K = 9 * 256, op.A = x2_, op.a_mode = A_CONV3X3, op.B = 1, op.H = 768, op.W = 768, op.C = 256;
    op.Wt = head0_wc_, op.bias = head0_cb_ + 9 * 128, op.border_cb = head0_cb_, op.out = h0_, op.ldo = 128;
    op.reverse = main_dir(s);
    gemm(prec_, op, s);
  } else {
    conv3x3(s, feat_[0], 768, 256, "head.0.weight", 128, F("head.0.bias"), ACT_NONE, nullptr, nullptr, h0_, nullptr);
  }
  if (prec_ == BF16) {
    // head.1 (ConvT) + head.2 (conv3x3) + ReLU + head.4 (1x1) + ReLU as ONE conv over the 768^2 map:
    // the 604 MB 128x1536^2 intermediate is never materialised and 77 GF of ConvT work disappears
    GemmOp op;
    op.M = 768 * 768, op.N = 128, op.K = 9 * 128, op.A = h0_, op.a_mode = A_CONV3X3, op.B = 1, op.H = 768, op.W = 768, op.C = 128;
    op.Wt = head_wc_, op.head_cb = head_cb_, op.out = canon, op.out_f32 = 1, op.out_mode = O_HEAD_FUSED;
    op.dot_w = F("head.4.weight"), op.dot_b = F("head.4.bias");
    op.reverse = main_dir(s);
    gemm(prec_, op, s);
  } else {
    convT(s, h0_, 768, 128, "head.1.weight", 128, h1_, 128, 0, F("head.1.bias"), nullptr);
    GemmOp op;
    op.M = IMG * IMG, op.N = 32, op.K = 9 * 128, op.A = h1_, op.a_mode = A_CONV3X3, op.B = 1, op.H = IMG, op.W = IMG, op.C = 128;
    op.Wt = W("head.2.weight"), op.bias = F("head.2.bias"), op.act = ACT_RELU;
    op.out = canon, op.out_f32 = 1, op.out_mode = O_DOT_RELU, op.dot_w = F("head.4.weight"), op.dot_b = F("head.4.bias");
    gemm(prec_, op, s);
  }

  if (fov_mode_ != 0) done(4);  // join the FOV head's stream
}

void Engine::forward(const float* x, int B, float* canon, float* fov_deg, cudaStream_t s) {
  DP_CHECK(finalized_, "dp_engine_finalize has not been called");
  DP_CHECK(B >= 1 && B <= max_batch_, "batch exceeds max_batch");
  DP_CUDA(cudaSetDevice(device_));
  if (prec_ == BF16) forward_impl<bf16>(x, B, canon, fov_deg, s);
  else forward_impl<float>(x, B, canon, fov_deg, s);
}

void Engine::forward_cached(int B, cudaStream_t s) {
  // Profiling needs per-launch events on the caller's stream; fp32 parity mode has no use for graphs.
  if (!use_graph_ || prec_ != BF16 || prof_active()) {
    forward(xbuf_, B, canon_, fov_, s);
    return;
  }
  DP_CUDA(cudaSetDevice(device_));
  FrameGraph& g = graphs_[B];
  const unsigned epoch = g_config_epoch.load();
  if (g.exec && g.epoch != epoch) {  // an A/B switch changed which kernels a frame launches
    DP_CUDA(cudaGraphExecDestroy(g.exec));
    g = FrameGraph();
  }
  if (!g.exec) {
    // first call: eager (every launcher's one-time setup -- cudaFuncSetAttribute, tensor-map encoding -- happens here);
    // second call: the same launch sequence is captured on an engine-owned stream (PyTorch's default stream is the
    // legacy stream, which cannot be captured) and instantiated; from then on it is replayed on the caller's stream.
    if (g.eager_calls++ == 0) {
      forward(xbuf_, B, canon_, fov_, s);
      return;
    }
    cudaGraph_t graph = nullptr;
    const long long l0 = g_launches.load();
    DP_CUDA(cudaStreamBeginCapture(cap_stream_, cudaStreamCaptureModeThreadLocal));
    try {
      forward(xbuf_, B, canon_, fov_, cap_stream_);
    } catch (...) {
      cudaStreamEndCapture(cap_stream_, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    DP_CUDA(cudaStreamEndCapture(cap_stream_, &graph));
    g.launches = g_launches.load() - l0;
    g_launches.fetch_sub(g.launches);  // nothing ran yet; every replay adds them
    cudaError_t e = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      g.exec = nullptr;
      throw Error(std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(e));
    }
    g.epoch = epoch;
  }
  DP_CUDA(cudaGraphLaunch(g.exec, s));
  count_launch(static_cast<int>(g.launches));
  last_B_ = B;
}

void Engine::infer(const void* img, int B, int H, int W, int src_fmt, const float* f_px_host, float* depth,
                   float* f_px_out, int interp, cudaStream_t s) {
  DP_CHECK(finalized_, "dp_engine_finalize has not been called");
  DP_CHECK(B >= 1 && B <= max_batch_, "batch exceeds max_batch");
  DP_CHECK(fov_mode_ != 0 || f_px_host != nullptr,
           "this model has no FOV head (use_fov_head=False): infer needs f_px (the reference fails on fov_deg=None here)");
  DP_CUDA(cudaSetDevice(device_));
  NvtxRange r("infer");
  if (src_fmt == 0 && H == IMG && W == IMG) {
    // already at network resolution: straight into the engine's input buffer (one 28 MB device copy per frame, so that
    // the captured graph always reads the same address)
    if (img != xbuf_)
      DP_CUDA(cudaMemcpyAsync(xbuf_, img, static_cast<size_t>(B) * 3 * IMG * IMG * 4, cudaMemcpyDeviceToDevice, s));
  } else {
    NvtxRange rp("preprocess (transform + resize)");
    preprocess(img, B, H, W, src_fmt, xbuf_, interp, s);
  }
  forward_cached(B, s);
  NvtxRange re("metric depth epilogue");
  const float* fin = nullptr;
  if (f_px_host) {
    DP_CUDA(cudaMemcpyAsync(fpx_in_, f_px_host, B * 4, cudaMemcpyHostToDevice, s));
    fin = fpx_in_;
  }
  compute_fpx(fov_, fin, W, f_px_out ? f_px_out : fpx_, B, s);
  depth_epilogue(canon_, f_px_out ? f_px_out : fpx_, B, H, W, depth, interp, s);
}

void Engine::infer_host(const void* img_host, int B, int H, int W, int src_fmt, const float* f_px_host, float* depth_host,
                        float* f_px_out_host) {
  DP_CUDA(cudaSetDevice(device_));
  const size_t in_bytes = static_cast<size_t>(B) * H * W * 3 * (src_fmt == 1 ? 1 : 4);
  const size_t out_bytes = static_cast<size_t>(B) * H * W * 4;
  if (himg_bytes_ < in_bytes) {
    if (himg_) DP_CUDA(cudaFree(himg_));
    DP_CUDA(cudaMalloc(&himg_, in_bytes));
    himg_bytes_ = in_bytes;
  }
  if (hdepth_bytes_ < out_bytes) {
    if (hdepth_) DP_CUDA(cudaFree(hdepth_));
    DP_CUDA(cudaMalloc(reinterpret_cast<void**>(&hdepth_), out_bytes));
    hdepth_bytes_ = out_bytes;
  }
  // a synchronous helper on the engine's own stream: work a caller enqueued earlier on ANOTHER stream may still be
  // using the (single) workspace, so wait for the device first
  DP_CUDA(cudaDeviceSynchronize());
  cudaStream_t s = host_stream_;
  DP_CUDA(cudaMemcpyAsync(himg_, img_host, in_bytes, cudaMemcpyHostToDevice, s));
  infer(himg_, B, H, W, src_fmt, f_px_host, hdepth_, fpx_, INTERP_BILINEAR, s);
  DP_CUDA(cudaMemcpyAsync(depth_host, hdepth_, out_bytes, cudaMemcpyDeviceToHost, s));
  if (f_px_out_host) DP_CUDA(cudaMemcpyAsync(f_px_out_host, fpx_, B * 4, cudaMemcpyDeviceToHost, s));
  DP_CUDA(cudaStreamSynchronize(s));
}

void Engine::unproject(const float* depth, const uint8_t* rgb, int H, int W, const float* f_px_dev, float* xyz,
                       float* rgb_out, uint8_t* valid_mask, int64_t* n_valid, cudaStream_t s) {
  DP_CUDA(cudaSetDevice(device_));
  const size_t need = unproject_scratch_ints(H, W);
  if (unproject_scratch_ints_ < need) {
    unproject_scratch_ = (int*)alloc(need * 4);
    unproject_scratch_ints_ = need;
  }
  dp::unproject(depth, rgb, H, W, f_px_dev, xyz, rgb_out, valid_mask, n_valid, unproject_scratch_, s);
}

void* Engine::ground_scratch(int64_t n, int grid_size) {
  const size_t need = ground_scratch_bytes(n, grid_size);
  if (ground_scratch_bytes_ < need) {  // grows monotonically; earlier blocks stay owned by the engine
    ground_scratch_ = alloc(need);
    ground_scratch_bytes_ = need;
  }
  return ground_scratch_;
}
void Engine::ground_normalize(float* xyz, int64_t n, const double* normal3, double d, uint64_t* counters, cudaStream_t s) {
  DP_CUDA(cudaSetDevice(device_));
  if (n == 0) return;
  DP_CHECK(xyz != nullptr && normal3 != nullptr && n > 0, "ground_normalize: bad arguments");
  dp::ground_normalize(xyz, n, normal3, d, ground_scratch(n, 1), reinterpret_cast<unsigned long long*>(counters), s);
}
void Engine::ground_grid_adjust(float* xyz, int64_t n, int grid_size, double percentile, uint64_t* counters, cudaStream_t s) {
  DP_CUDA(cudaSetDevice(device_));
  if (n == 0) return;
  DP_CHECK(xyz != nullptr && n > 0, "ground_grid_adjust: bad arguments");
  DP_CHECK(percentile >= 0.0 && percentile <= 100.0, "ground_grid_adjust: percentile outside [0, 100]");
  dp::ground_grid_adjust(xyz, n, grid_size, percentile, ground_scratch(n, grid_size),
                         reinterpret_cast<unsigned long long*>(counters), s);
}

void Engine::colorize(const float* depth, int H, int W, const uint8_t* lut, void* out, float min_depth, float max_depth,
                      cudaStream_t s) {
  DP_CUDA(cudaSetDevice(device_));
  if (!colorize_mm_) colorize_mm_ = (float*)alloc(colorize_scratch_bytes());
  dp::colorize(depth, H, W, lut, out, colorize_mm_, min_depth, max_depth, s);
}

// ============================================================================ taps
template <typename T>
int64_t Engine::tap_impl(const std::string& stage, float* out, int64_t capacity, cudaStream_t s) {
  struct Map { const char* name; void* ptr; int S, C; bool batched; };
  const Map maps[] = {
      {"lat0_merged", lat0m_, 96, EMB, true}, {"lat1_merged", lat1m_, 96, EMB, true}, {"x0_merged", x0m_, 96, EMB, true},
      {"x1_merged", x1m_, 48, EMB, true},     {"x2_tokens", x2m_, 24, EMB, true},     {"global_tokens", globm_, 24, EMB, true},
      {"enc0", enc0_, 768, 256, false},       {"enc1", enc1_, 384, 256, false},       {"enc2", enc2_, 192, 512, false},
      {"enc3", enc3_, 96, 1024, false},       {"enc4", enc4_, 48, 1024, false},       {"lowres", lowres_, 48, 256, false},
      {"decoder_out", feat_[0], 768, 256, false},
  };
  for (const Map& m : maps) {
    if (stage == m.name) {
      const int B = m.batched ? last_B_ : 1;
      const int64_t n = static_cast<int64_t>(B) * m.S * m.S * m.C;
      DP_CHECK(n <= capacity, "tap buffer too small");
      nhwc_to_nchw_f32<T>((const T*)m.ptr, out, B, m.S, m.S, m.C, s);
      return n;
    }
  }
  throw Error("unknown tap stage: " + stage);
}

int64_t Engine::tap(const std::string& stage, float* out, int64_t capacity, cudaStream_t s) {
  DP_CHECK(finalized_ && last_B_ > 0, "dp_tap needs a previous dp_forward");
  DP_CUDA(cudaSetDevice(device_));
  if (stage == "decoder_out" && prec_ == BF16 && head0_fused_) {
    // bf16 mode never materialises the decoder's output map (fusions.0.out_conv is composed into head.0): evaluate the
    // 1x1 on demand from resnet2's output of the last decoded frame
    GemmOp op;
    const std::string p = "decoder.fusions.0.out_conv.";
    op.M = 768 * 768, op.N = 256, op.K = 256, op.A = x2_, op.lda = 256, op.Wt = W(p + "weight"), op.bias = F(p + "bias");
    op.out = feat_[0], op.ldo = 256;
    gemm(prec_, op, s);
  }
  return prec_ == BF16 ? tap_impl<bf16>(stage, out, capacity, s) : tap_impl<float>(stage, out, capacity, s);
}

// ============================================================================ unit-test entries
namespace {
template <typename T>
T* to_dev(const float* src, size_t n, cudaStream_t s) {
  T* p = nullptr;
  DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&p), n * sizeof(T), s));
  convert<float, T>(src, p, (long long)n, s);
  return p;
}
}  // namespace

void Engine::gemm_test(int backend, const float* A, const float* Wt, const float* bias, float* C, int M, int N, int K,
                       int act_flags, cudaStream_t s) {
  DP_CUDA(cudaSetDevice(device_));
  // act_flags: low byte = activation; bf16 backend only: 0x100 bf16 output (TMA-store epilogue), 0x200 fp32
  // residual form C += bias * (acc + bias) in place (gamma := bias), 0x400 ConvT k2 s2 pixel shuffle of an
  // S x S map (M = S*S, N = 4*Cout) into a (2S, 2S, Cout) bf16 map, 0x800 with 0x100/0x400: return the ReLU twin
  const int act = act_flags & 0xff, flags = act_flags & ~0xff;
  GemmOp op;
  op.M = M, op.N = N, op.K = K, op.lda = K, op.bias = bias, op.act = act, op.out = C, op.out_f32 = 1, op.ldo = N;
  if (backend == BF16) {
    bf16* a = to_dev<bf16>(A, (size_t)M * K, s);
    bf16* w = to_dev<bf16>(Wt, (size_t)N * K, s);
    op.A = a, op.Wt = w;
    bf16 *ob = nullptr, *orelu = nullptr;
    const size_t n_out = (size_t)M * N;
    // 0x1000 LayerNorm-folded consumer (with 0x100): C = LN(A; g, b_ln) W^T + bias, g[k] = 1 + 0.25 sin(0.37 k),
    //        b_ln[k] = 0.1 cos(0.11 k), computed from the RAW bf16 A + row statistics + folded weights;
    // 0x2000 LayerNorm-emitting producer (with 0x200, N = 1024): C is (2M, N); rows [M, 2M) receive
    //        (x - mean) * rstd rebuilt from the epilogue's bf16 copy and partial sums of the updated x
    float *ln_stats = nullptr, *ln_c = nullptr, *ln_d = nullptr, *ln_g = nullptr, *ln_b = nullptr;
    bf16* ln_xb = nullptr;
    if (flags & 0x1000) {
      DP_CHECK((flags & 0x100) && K == 1024 && bias != nullptr, "LN consumer test: 0x100, K = 1024, bias");
      std::vector<float> hg(K), hb(K);
      for (int k = 0; k < K; ++k) hg[k] = 1.f + 0.25f * std::sin(0.37f * k), hb[k] = 0.1f * std::cos(0.11f * k);
      DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ln_g), K * 4, s));
      DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ln_b), K * 4, s));
      DP_CUDA(cudaMemcpyAsync(ln_g, hg.data(), K * 4, cudaMemcpyHostToDevice, s));
      DP_CUDA(cudaMemcpyAsync(ln_b, hb.data(), K * 4, cudaMemcpyHostToDevice, s));
      DP_CUDA(cudaStreamSynchronize(s));  // the host vectors go out of scope
      DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ln_stats), (size_t)M * LN_SLOTS * 8, s));
      DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ln_c), N * 4, s));
      DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ln_d), N * 4, s));
      ln_stats_cast(A, a, ln_stats, M, s);
      ln_fold(Wt, ln_g, ln_b, bias, w, ln_c, ln_d, N, K, s);
      op.ln_stats = ln_stats;
      op.grp[0].M = M, op.grp[0].Wt = w, op.grp[0].bias = ln_d, op.grp[0].ln_c = ln_c;
      op.bias = ln_d;
    }
    // 0x4000 (with 0x200 | 0x2000): the producer in its PAIR form -- C is split into (hi, lo) 16-bit arrays first,
    //        updated in place by the epilogue, and rebuilt as hi + lo afterwards
    bf16* ln_xlo = nullptr;
    if (flags & 0x2000) {
      DP_CHECK((flags & 0x200) && N == 1024, "LN producer test: 0x200, N = 1024");
      DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ln_stats), (size_t)M * LN_SLOTS * 8, s));
      DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ln_xb), n_out * 2, s));
      op.ln_xb = ln_xb, op.ln_stats_out = ln_stats;
      if (flags & 0x4000) {
        DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ln_xlo), n_out * 2, s));
        ln_stats_cast(C, ln_xb, ln_stats, M, s, ln_xlo);
        op.ln_xlo = ln_xlo;
      }
    }
    if (flags & 0x4000) {
      DP_CHECK(flags & 0x2000, "pair producer test: 0x200 | 0x2000 | 0x4000");
      op.gamma = bias, op.out = nullptr, op.out_f32 = 0;
    } else if (flags & 0x200) {
      op.gamma = bias, op.res = C, op.res_f32 = 1, op.ldres = N;
    } else if (flags & (0x100 | 0x400)) {
      DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ob), n_out * 2, s));
      op.out = ob, op.out_f32 = 0;
      if (flags & 0x800) {
        DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&orelu), n_out * 2, s));
        op.out_relu = orelu;
      }
      if (flags & 0x400) {
        int S = 1;
        while (S * S < M) ++S;
        DP_CHECK(S * S == M && N % 4 == 0, "ConvT test: M must be a square, N = 4*Cout");
        op.B = 1, op.H = S, op.W = S, op.cout = N / 4, op.out_mode = O_CONVT2X2, op.ldo = N / 4;
        op.bias_mod = bias ? N / 4 : 0;
      }
    }
    gemm_tc(op, s);
    if (ob) convert<bf16, float>((flags & 0x800) ? orelu : ob, C, (long long)n_out, s);
    if (flags & 0x4000) pair_to_f32(ln_xb, ln_xlo, C, (long long)n_out, s);
    if (flags & 0x2000) ln_apply_from_stats(ln_xb, ln_stats, C + n_out, M, s);
    for (void* p : {(void*)ln_stats, (void*)ln_c, (void*)ln_d, (void*)ln_g, (void*)ln_b, (void*)ln_xb, (void*)ln_xlo})
      if (p) DP_CUDA(cudaFreeAsync(p, s));
    DP_CUDA(cudaFreeAsync(a, s));
    DP_CUDA(cudaFreeAsync(w, s));
    if (ob) DP_CUDA(cudaFreeAsync(ob, s));
    if (orelu) DP_CUDA(cudaFreeAsync(orelu, s));
    DP_CUDA(cudaStreamSynchronize(s));
    tmap_cache_clear();  // the temporaries' tensor maps must not be reused
  } else {
    DP_CHECK(flags == 0, "gemm_test: epilogue flags need the bf16 backend");
    op.A = A, op.Wt = Wt;
    gemm_simt(op, s);
  }
}
void Engine::conv3x3_test(int backend_flags, const float* x, const float* w, const float* bias, float* y, int B, int H,
                          int W_, int Cin, int Cout, cudaStream_t s) {
  DP_CUDA(cudaSetDevice(device_));
  // backend_flags: low byte = backend; bf16 only: 0x100 bf16 NHWC output through the TMA-store epilogue,
  // 0x800 (with 0x100) dual store, y receives the ReLU twin
  const int backend = backend_flags & 0xff, flags = backend_flags & ~0xff;
  GemmOp op;
  op.M = B * H * W_, op.N = Cout, op.K = 9 * Cin, op.a_mode = A_CONV3X3, op.B = B, op.H = H, op.W = W_, op.C = Cin;
  op.bias = bias, op.out = y, op.out_f32 = 1, op.ldo = Cout;
  const size_t nw = (size_t)Cout * Cin * 9;
  if (backend == BF16) {
    bf16* a = to_dev<bf16>(x, (size_t)B * H * W_ * Cin, s);
    bf16* wp = nullptr;
    DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&wp), nw * 2, s));
    pack_oihw_to_ohwi<bf16>(w, wp, Cout, Cin, 3, 3, s);
    op.A = a, op.Wt = wp;
    bf16 *ob = nullptr, *orelu = nullptr;
    const size_t n_out = (size_t)B * H * W_ * Cout;
    if (flags & 0x100) {
      DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ob), n_out * 2, s));
      op.out = ob, op.out_f32 = 0;
      if (flags & 0x800) {
        DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&orelu), n_out * 2, s));
        op.out_relu = orelu;
      }
    }
    gemm_tc(op, s);
    if (ob) convert<bf16, float>((flags & 0x800) ? orelu : ob, y, (long long)n_out, s);
    DP_CUDA(cudaFreeAsync(a, s));
    DP_CUDA(cudaFreeAsync(wp, s));
    if (ob) DP_CUDA(cudaFreeAsync(ob, s));
    if (orelu) DP_CUDA(cudaFreeAsync(orelu, s));
    DP_CUDA(cudaStreamSynchronize(s));
    tmap_cache_clear();
  } else {
    float* wp = nullptr;
    DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&wp), nw * 4, s));
    pack_oihw_to_ohwi<float>(w, wp, Cout, Cin, 3, 3, s);
    op.A = x, op.Wt = wp;
    gemm_simt(op, s);
    DP_CUDA(cudaFreeAsync(wp, s));
  }
}

void Engine::attention_test(int backend, const float* qkv, float* out, int n, cudaStream_t s) {
  DP_CUDA(cudaSetDevice(device_));
  if (backend >= BF16) {
    const size_t nq = (size_t)n * SEQ * 3 * EMB, no = (size_t)n * SEQ * EMB;
    bf16* q = to_dev<bf16>(qkv, nq, s);
    bf16* o = nullptr;
    DP_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&o), no * 2, s));
    // backend bits 8-15: tcgen05 kernel variant + 1 (0xFF = back to the default), bit 16: no MUFU ping-pong
    if (((backend >> 8) & 0xFF) == 0xFF) attention_tc_set_variant(-1, 1);
    else if ((backend >> 8) & 0xFF) attention_tc_set_variant(((backend >> 8) & 0xFF) - 1, !((backend >> 16) & 1));
    if ((backend & 0xFF) == 2) attention_bf16(q, o, n, s);
    else attention_bf16_tc(q, o, n, s);
    convert<bf16, float>(o, out, (long long)no, s);
    DP_CUDA(cudaStreamSynchronize(s));
    tmap_cache_clear();
    DP_CUDA(cudaFreeAsync(q, s));
    DP_CUDA(cudaFreeAsync(o, s));
  } else {
    attention_f32(qkv, out, n, s);
  }
}

}  // namespace dp

namespace dp {
// Micro-benchmark of the bf16 cores on scratch buffers (values are irrelevant, zeros):
//  kind 0: GEMM + bias -> bf16          (qkv)      kind 1: GEMM + bias + GELU -> bf16 (fc1)
//  kind 2: GEMM + bias, *gamma, + fp32 residual in place (proj / fc2)
//  kind 3: conv3x3 on an MxM map, Cin = K, Cout = N, + bias + ReLU -> bf16
//  kind 4: attention over M sequences       kind 5: LayerNorm over M rows
//  kinds 6-10: the HBM-bound kernels either side of the network (resize, split, depth epilogue,
//  unprojection, colourise) on an M x N image
// Returns the mean milliseconds per launch over `iters` launches (CUDA events).
float Engine::kernel_bench(int kind, int M, int N, int K, int iters) {
  DP_CUDA(cudaSetDevice(device_));
  // A/B bits (process-wide, sticky): 0x100 / 0x200 switch the fp32-residual L2 prefetch on / off
  if (kind & 0x100) gemm_tc_set_res_prefetch(1);
  if (kind & 0x200) gemm_tc_set_res_prefetch(0);
  // 0x400: L2 persistence of the fp32 residual stream on, set-aside = `iters >> 16` MB (0 -> 96); 0x800: off
  if (kind & 0x400) gemm_tc_set_l2_persist((iters >> 16) ? (iters >> 16) : 96);
  if (kind & 0x800) gemm_tc_set_l2_persist(0);
  // 0x4000 / 0x8000 (sticky): GEMM / attention launches use at most `iters >> 16` SMs (0 = all) -- the SM-partition
  // experiment of DESIGN.md §9.5
  if (kind & 0x4000) gemm_tc_set_sm_limit(iters >> 16);
  if (kind & 0x8000) attention_tc_set_sm_limit(iters >> 16);
  if (kind & 0x1000) hbm_v2_set(1);  // 0x1000 / 0x2000: opt-in second-generation HBM kernels on / off
  if (kind & 0x2000) hbm_v2_set(0);
  iters &= 0xFFFF;
  kind &= 0xFF;
  cudaStream_t s = nullptr;
  auto dalloc = [&](size_t bytes) {
    void* p = nullptr;
    DP_CUDA(cudaMalloc(&p, bytes));
    DP_CUDA(cudaMemset(p, 0, bytes));
    return p;
  };
  std::vector<void*> bufs;
  // Tensor-core operands are filled with random bf16 values unless DEPTHPRO_BENCH_DATA=zeros: measured on
  // B200, a GEMM on zero-filled operands runs at the full 1965 MHz (913 W), the same kernel on real data is
  // power-capped to ~1.5 GHz -- zeros overstate the sustained rate by ~40 %.
  static const bool zeros = [] { const char* e = getenv("DEPTHPRO_BENCH_DATA"); return e && std::string(e) == "zeros"; }();
  const bool tensor_kind = kind <= 4 || (kind >= 11 && kind <= 14) || kind == 20 || kind == 21;
  auto B = [&](size_t bytes) {
    void* p = dalloc(bytes);
    if (tensor_kind && !zeros && bytes >= 2) fill_random_bf16(p, bytes, 0x9e3779b9u + static_cast<unsigned>(bufs.size()), s);
    bufs.push_back(p);
    return p;
  };
  GemmOp op;
  std::function<void()> run;
  if (kind <= 2 || (kind >= 11 && kind <= 14)) {
    // 11 / 12 / 13: kinds 0 / 1 / 2 in their LayerNorm-folded forms (consumer, consumer + GELU, producer);
    // 14: the producer over a (hi, lo) 16-bit pair residual stream
    op.M = M, op.N = N, op.K = K, op.lda = K;
    op.A = B((size_t)M * K * 2), op.Wt = B((size_t)N * K * 2), op.bias = (float*)B((size_t)N * 4);
    if (kind == 14) {
      op.gamma = (float*)B((size_t)N * 4), op.ldo = N;
      op.ln_xb = B((size_t)M * N * 2), op.ln_xlo = B((size_t)M * N * 2), op.ln_stats_out = (float*)B((size_t)M * LN_SLOTS * 8);
    } else if (kind == 2 || kind == 13) {
      float* r = (float*)B((size_t)M * N * 4);
      op.gamma = (float*)B((size_t)N * 4), op.res = r, op.res_f32 = 1, op.ldres = N, op.out = r, op.out_f32 = 1, op.ldo = N;
      if (kind == 13) op.ln_xb = B((size_t)M * N * 2), op.ln_stats_out = (float*)B((size_t)M * LN_SLOTS * 8);
    } else {
      op.out = B((size_t)M * N * 2), op.ldo = N, op.act = (kind == 1 || kind == 12) ? ACT_GELU : ACT_NONE;
      if (kind >= 11) {
        op.ln_stats = (float*)B((size_t)M * LN_SLOTS * 8);
        op.grp[0].M = M, op.grp[0].Wt = op.Wt, op.grp[0].bias = op.bias, op.grp[0].ln_c = (float*)B((size_t)N * 4);
      }
    }
    run = [&] { gemm_tc(op, s); };
  } else if (kind == 3) {
    op.M = M * M, op.N = N, op.K = 9 * K, op.a_mode = A_CONV3X3, op.B = 1, op.H = M, op.W = M, op.C = K;
    op.A = B((size_t)M * M * K * 2), op.Wt = B((size_t)N * 9 * K * 2), op.bias = (float*)B((size_t)N * 4);
    op.act = ACT_RELU, op.out = B((size_t)M * M * N * 2), op.ldo = N;
    run = [&] { gemm_tc(op, s); };
  } else if (kind == 4) {
    bf16* q = (bf16*)B((size_t)M * SEQ * 3 * EMB * 2);
    bf16* o = (bf16*)B((size_t)M * SEQ * EMB * 2);
    if (N > 0) attention_tc_set_variant((N - 1) & 63, !(((N - 1) >> 6) & 1));  // N = 1 + variant + 64 * no-ping-pong
    run = [&, q, o] { attention_bf16_tc(q, o, M, s); };
  } else if (kind == 5) {
    float* x = (float*)B((size_t)M * EMB * 4);
    bf16* y = (bf16*)B((size_t)M * EMB * 2);
    float* w = (float*)B(EMB * 4);
    run = [&, x, y, w] { layernorm_rows<bf16>(x, y, w, w, M, RowMap(), 1, s); };
  } else if (kind == 6) {  // uint8 HWC (M x N) -> fused transform + bilinear resize -> fp32 3x1536^2
    uint8_t* src = (uint8_t*)B((size_t)M * N * 3);
    float* x = (float*)B((size_t)3 * IMG * IMG * 4);
    run = [&, src, x] { resize_to_1536(src, 1, 1, M, N, x, INTERP_BILINEAR, s); };
  } else if (kind == 7) {  // pyramid + 35-patch split + 16x16 im2col, fp32 frame -> bf16 A operands
    float* x = (float*)B((size_t)3 * IMG * IMG * 4);
    bf16* a35 = (bf16*)B((size_t)35 * 576 * 768 * 2);
    bf16* a1 = (bf16*)B((size_t)576 * 768 * 2);
    run = [&, x, a35, a1] { split_im2col<bf16>(x, 1, a35, a1, s); };
  } else if (kind == 8 || kind == 9 || kind == 10) {
    // 8: canonical inverse depth 1536^2 -> metric depth (M x N);  9: depth (M x N) + rgb -> xyz + colours;
    // 10: depth (M x N) -> colour-mapped uint8 RGB
    std::vector<float> ones((size_t)IMG * IMG > (size_t)M * N ? (size_t)IMG * IMG : (size_t)M * N, 2.0f);
    float* canon = (float*)B((size_t)IMG * IMG * 4);
    float* depth = (float*)B((size_t)M * N * 4);
    float* f = (float*)B(4);
    DP_CUDA(cudaMemcpy(canon, ones.data(), (size_t)IMG * IMG * 4, cudaMemcpyHostToDevice));
    DP_CUDA(cudaMemcpy(depth, ones.data(), (size_t)M * N * 4, cudaMemcpyHostToDevice));
    const float fpx = 1000.f;
    DP_CUDA(cudaMemcpy(f, &fpx, 4, cudaMemcpyHostToDevice));
    if (kind == 8) {
      run = [&, canon, f, depth] { depth_epilogue(canon, f, 1, M, N, depth, INTERP_BILINEAR, s); };
    } else if (kind == 9) {
      uint8_t* rgb = (uint8_t*)B((size_t)M * N * 3);
      float* xyz = (float*)B((size_t)M * N * 12);
      float* col = (float*)B((size_t)M * N * 12);
      int64_t* nv = (int64_t*)B(8);
      int* scratch = (int*)B(unproject_scratch_ints(M, N) * sizeof(int));
      run = [&, depth, rgb, f, xyz, col, nv, scratch] { dp::unproject(depth, rgb, M, N, f, xyz, col, nullptr, nv, scratch, s); };
    } else {
      uint8_t* lut = (uint8_t*)B(768);
      uint8_t* out = (uint8_t*)B((size_t)M * N * 3);
      float* mm = (float*)B(colorize_scratch_bytes());
      run = [&, depth, lut, out, mm] { dp::colorize(depth, M, N, lut, out, mm, NAN, NAN, s); };
    }
  } else if (kind == 20 || kind == 21) {
    // SM-partition experiment: the GEMMs of one ViT layer for M rows (qkv, proj, fc1 + GELU, fc2; plain epilogues) and the
    // attention of N sequences.  kind 20: one after the other on one stream; kind 21: the GEMM chain on one stream, the
    // attention on a second one (so they can share the chip under the SM limits set with 0x4000 / 0x8000).
    static cudaStream_t s2 = nullptr;
    static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    if (!s2) {
      DP_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
      DP_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      DP_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    static cudaStream_t s1 = nullptr;
    if (!s1) DP_CUDA(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    s = s1;
    bf16* x = (bf16*)B((size_t)M * EMB * 2);
    bf16* q = (bf16*)B((size_t)M * 3 * EMB * 2);
    bf16* hdn = (bf16*)B((size_t)M * 4 * EMB * 2);
    bf16* aq = (bf16*)B((size_t)N * SEQ * 3 * EMB * 2);
    bf16* ao = (bf16*)B((size_t)N * SEQ * EMB * 2);
    void* w_qkv = B((size_t)3 * EMB * EMB * 2);
    void* w_proj = B((size_t)EMB * EMB * 2);
    void* w_fc1 = B((size_t)4 * EMB * EMB * 2);
    void* w_fc2 = B((size_t)4 * EMB * EMB * 2);
    float* bias = (float*)B((size_t)4 * EMB * 4);
    auto g1 = [=](const void* A, int Kk, const void* Wt, int Nn, void* out, int act) {
      GemmOp o;
      o.M = M, o.N = Nn, o.K = Kk, o.lda = Kk, o.A = A, o.Wt = Wt, o.bias = bias, o.act = act, o.out = out, o.ldo = Nn;
      gemm_tc(o, s1);
    };
    const bool conc = kind == 21;
    run = [=] {
      if (conc) {
        DP_CUDA(cudaEventRecord(ev_fork, s1));
        DP_CUDA(cudaStreamWaitEvent(s2, ev_fork, 0));
        attention_bf16_tc(aq, ao, N, s2);
        DP_CUDA(cudaEventRecord(ev_join, s2));
      } else {
        attention_bf16_tc(aq, ao, N, s1);
      }
      g1(x, EMB, w_qkv, 3 * EMB, q, ACT_NONE);
      g1(x, EMB, w_proj, EMB, x, ACT_NONE);
      g1(x, EMB, w_fc1, 4 * EMB, hdn, ACT_GELU);
      g1(hdn, 4 * EMB, w_fc2, EMB, x, ACT_NONE);
      if (conc) DP_CUDA(cudaStreamWaitEvent(s1, ev_join, 0));
    };
  } else {
    DP_CHECK(false, "kernel_bench: unknown kind");
  }
  for (int i = 0; i < 3; ++i) run();
  cudaEvent_t a, b;
  DP_CUDA(cudaEventCreate(&a));
  DP_CUDA(cudaEventCreate(&b));
  float ms = 0.f;
  if (kind >= 5 && kind <= 10) {
    // HBM-bound kernels with working sets near the 126 MB L2: overwrite a 256 MB buffer before every
    // timed launch so inputs come from HBM, and time each launch on its own
    const size_t FLUSH = 256u << 20;
    void* fl = B(FLUSH);
    for (int i = 0; i < iters; ++i) {
      DP_CUDA(cudaMemsetAsync(fl, i & 0xff, FLUSH, s));
      DP_CUDA(cudaEventRecord(a, s));
      run();
      DP_CUDA(cudaEventRecord(b, s));
      DP_CUDA(cudaEventSynchronize(b));
      float t = 0.f;
      DP_CUDA(cudaEventElapsedTime(&t, a, b));
      ms += t;
    }
  } else {
    DP_CUDA(cudaEventRecord(a, s));
    for (int i = 0; i < iters; ++i) run();
    DP_CUDA(cudaEventRecord(b, s));
    DP_CUDA(cudaEventSynchronize(b));
    DP_CUDA(cudaEventElapsedTime(&ms, a, b));
  }
  cudaEventDestroy(a), cudaEventDestroy(b);
  for (void* p : bufs) cudaFree(p);
  tmap_cache_clear();
  return ms / iters;
}
}  // namespace dp
