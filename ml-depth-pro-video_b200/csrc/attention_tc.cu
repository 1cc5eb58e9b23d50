// Flash attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Problem: 16 heads x 64, 577 tokens per sequence, softmax(Q K^T / 8) V, no mask (timm Attention ->
// F.scaled_dot_product_attention, wired at src/depth_pro/network/vit_factory.py:97-110).
// qkv is (nseq*577, 3072) bf16 with columns [q | k | v], each 16 heads x 64; out is (nseq*577, 1024).
//
// One persistent CTA per SM loops over work items (sequence, head, PAIR of 128-row query tiles); the
// two query tiles of an item are two independent streams A / B that share every K / V block.  Keys
// are processed in 5 blocks of 128 (640 >= 577, the tail is masked).  Warp roles:
//   warp 0      TMA producer : Q tiles (once per item) and K/V blocks (3-deep ring), 128B swizzle.
//   warp 1      MMA issuer   : per stream S = Q K^T (tcgen05.mma M128 N128 K16 x4, A/B K-major) and
//                              O_j = P V (M128 N64 K16 x8, A = P K-major from smem, B = V MN-major
//                              straight from the TMA tile); S_A, S_B, O_A[2], O_B[2] fill the 512 TMEM
//                              columns.
//   warp 2      TMEM allocator
//   warps 4-7   softmax of stream A, warps 8-11 softmax of stream B: one thread per query row:
//                              tcgen05.ld the 128 scores, online softmax in fp32 (exp2, log2e folded into
//                              the scale), P as bf16 into the swizzled smem tile the next MMA reads, O_j
//                              folded from TMEM into registers with the running-max correction, final
//                              1/l and bf16 store.  The two groups are NOT synchronised with each other,
//                              so on every SM sub-partition one warp's TMEM loads overlap the other's
//                              MUFU exp2 work (a single group serialises LDTM -> MUFU -> STS per block
//                              and leaves the tensor pipe idle 80% of the time).
// Register budget: 384 threads x 168 would not hold 128 scores + 64 outputs per softmax thread, so
// the control warps drop to 40 registers and the softmax warps grow to 232 (setmaxnreg).
#include "attention.cuh"
#include "ptx.cuh"

namespace dp {
namespace {

constexpr int SEQ = 577, HD = 64, NH = 16, LDQ = 3 * NH * HD, LDO = NH * HD;
constexpr int QT = 128;                         // query rows per stream
constexpr int KB = 128;                         // keys per block
constexpr int NB = (SEQ + KB - 1) / KB;         // 5 key blocks
constexpr int NQT = (SEQ + QT - 1) / QT;        // 5 query tiles
constexpr int NPAIR = (NQT + 1) / 2;            // 3 items per (sequence, head): tiles (0,1) (2,3) (4,-)
constexpr int KV_STAGES = 3;
constexpr int THREADS = 384;
constexpr uint32_t TILE_BYTES = 128 * 128;      // 128 rows x 64 bf16 = 16 KB
constexpr uint32_t SMEM_BYTES = TILE_BYTES * (2 + 2 * KV_STAGES + 4) + 1024 + 256;

// idesc: D=f32, A=B=bf16, A K-major; B K-major (QK^T) or MN-major (PV: bit 16)
constexpr uint32_t IDESC_QK = ptx::umma_idesc_bf16(128, 128);
constexpr uint32_t IDESC_PV = ptx::umma_idesc_bf16(128, 64) | (1u << 16);
// the last key block holds 577 - 512 = 65 keys: its Q K^T runs at N = 80 and its P V at K = 80
constexpr int LAST_KEYS = SEQ - (NB - 1) * KB;                 // 65
constexpr int LAST_N = (LAST_KEYS + 15) / 16 * 16;            // 80
constexpr uint32_t IDESC_QK_LAST = ptx::umma_idesc_bf16(128, LAST_N);
// smem descriptor high word: SBO = 1024 B, version 1, SWIZZLE_128B (same for K-major and MN-major tiles)
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);

__device__ __forceinline__ uint64_t desc(uint32_t lo) { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; }
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16); }

__device__ __forceinline__ float ex2(float x) {  // MUFU ex2.approx: 2 ulp, -inf -> 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, int nseq) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                               // 2 tiles (stream A, B)
  uint8_t* sK = sQ + 2 * TILE_BYTES;                // KV_STAGES tiles
  uint8_t* sV = sK + KV_STAGES * TILE_BYTES;        // KV_STAGES tiles
  uint8_t* sP = sV + KV_STAGES * TILE_BYTES;        // 2 streams x 2 k-chunk tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * TILE_BYTES);
  uint64_t* q_full = bars;                          // 1
  uint64_t* q_empty = bars + 1;                     // 1
  uint64_t* kv_full = bars + 2;                     // KV_STAGES
  uint64_t* kv_empty = kv_full + KV_STAGES;         // KV_STAGES
  uint64_t* s_full = kv_empty + KV_STAGES;          // per stream
  uint64_t* s_empty = s_full + 2;                   // per stream, 128 arrivals
  uint64_t* p_full = s_empty + 2;                   // per stream, 128 arrivals
  uint64_t* pv_done = p_full + 2;                   // per stream
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = nseq * NH * NPAIR;

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&tmQKV);
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(q_empty, 1);
    for (int i = 0; i < KV_STAGES; ++i) ptx::mbar_init(&kv_full[i], 1), ptx::mbar_init(&kv_empty[i], 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&s_empty[i], 128);
      ptx::mbar_init(&p_full[i], 128);
      ptx::mbar_init(&pv_done[i], 1);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S_A [0,128), S_B [128,256), O_A[2] [256,384), O_B[2] [384,512)

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      for (int item = blockIdx.x, it = 0; item < n_items; item += gridDim.x, ++it) {
        const int pr = item % NPAIR, h = (item / NPAIR) % NH, seq = item / (NPAIR * NH);
        const bool hasB = 2 * pr + 1 < NQT;
        const int row0 = seq * SEQ;
        ptx::mbar_wait(q_empty, (it & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(q_full, hasB ? 2 * TILE_BYTES : TILE_BYTES);
          ptx::tma_load_2d(sQ, &tmQKV, q_full, h * HD, row0 + 2 * pr * QT);
          if (hasB) ptx::tma_load_2d(sQ + TILE_BYTES, &tmQKV, q_full, h * HD, row0 + (2 * pr + 1) * QT);
        }
        __syncwarp();
        for (int j = 0; j < NB; ++j) {
          const int gb = it * NB + j, st = gb % KV_STAGES;
          ptx::mbar_wait(&kv_empty[st], ((gb / KV_STAGES) & 1) ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&kv_full[st], 2 * TILE_BYTES);
            ptx::tma_load_2d(sK + st * TILE_BYTES, &tmQKV, &kv_full[st], NH * HD + h * HD, row0 + j * KB);
            ptx::tma_load_2d(sV + st * TILE_BYTES, &tmQKV, &kv_full[st], 2 * NH * HD + h * HD, row0 + j * KB);
          }
          __syncwarp();
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer
      // all tiles sit at compile-time offsets from the (1024-B aligned) smem base: one live register
      const uint32_t q_lo = desc_lo(ptx::smem_u32(sQ));
      const uint32_t k_lo = q_lo + (2 * TILE_BYTES >> 4);
      const uint32_t v_lo = k_lo + (KV_STAGES * TILE_BYTES >> 4);
      const uint32_t p_lo = v_lo + (KV_STAGES * TILE_BYTES >> 4);
      int nB = 0;  // items so far in which stream B was active (its barrier phases advance only then)
      for (int item = blockIdx.x, it = 0; item < n_items; item += gridDim.x, ++it) {
        const bool hasB = 2 * (item % NPAIR) + 1 < NQT;
        ptx::mbar_wait(q_full, it & 1);
        // P V of key block jj for both streams, then release that K/V stage
        auto issue_pv = [&](int jj) {
          const int st = (it * NB + jj) % KV_STAGES;
          for (int sidx = 0; sidx < (hasB ? 2 : 1); ++sidx) {
            const int gbs = (sidx == 0 ? it : nB) * NB + jj;
            ptx::mbar_wait(&p_full[sidx], gbs & 1);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              const uint32_t d = tmem_base + 256 + sidx * 128 + (jj & 1) * 64;
              const int nks = jj == NB - 1 ? LAST_N / 16 : KB / 16;
#pragma unroll
              for (int ks = 0; ks < KB / 16; ++ks) {
                if (ks >= nks) break;
                // A = P: k-chunk tile (ks >> 2), +32 B per K=16 step; B = V (MN-major): +16 keys = 2048 B
                const uint64_t da = desc(p_lo + sidx * (2 * TILE_BYTES >> 4) + (ks >> 2) * (TILE_BYTES >> 4) + (ks & 3) * 2);
                const uint64_t db = desc(v_lo + st * (TILE_BYTES >> 4) + ks * (2048 >> 4));
                ptx::umma_bf16(d, da, db, IDESC_PV, ks != 0);
              }
              ptx::umma_commit(&pv_done[sidx]);
            }
            __syncwarp();
          }
          if (ptx::elect_one()) ptx::umma_commit(&kv_empty[st]);
          __syncwarp();
        };
        for (int j = 0; j < NB; ++j) {
          const int st = (it * NB + j) % KV_STAGES;
          ptx::mbar_wait(&kv_full[st], ((it * NB + j) / KV_STAGES) & 1);
          for (int sidx = 0; sidx < (hasB ? 2 : 1); ++sidx) {
            const int gbs = (sidx == 0 ? it : nB) * NB + j;
            ptx::mbar_wait(&s_empty[sidx], (gbs & 1) ^ 1);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < HD / 16; ++ks)
                ptx::umma_bf16(tmem_base + sidx * 128, desc(q_lo + sidx * (TILE_BYTES >> 4) + ks * 2),
                               desc(k_lo + st * (TILE_BYTES >> 4) + ks * 2), j == NB - 1 ? IDESC_QK_LAST : IDESC_QK,
                               ks != 0);
              ptx::umma_commit(&s_full[sidx]);
            }
            __syncwarp();
          }
          if (j == NB - 1) {
            if (ptx::elect_one()) ptx::umma_commit(q_empty);  // Q tiles are free once the last QK^T retires
            __syncwarp();
          }
          if (j > 0) issue_pv(j - 1);
        }
        issue_pv(NB - 1);
        if (hasB) ++nB;
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ------------------------------------------------------------------ softmax / accumulate
    const int q = warp & 3;             // TMEM lane quadrant
    const int sidx = (warp - 4) >> 2;   // stream: 0 = A, 1 = B
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + sidx * 128, tO = tmem_base + 256 + sidx * 128;
    uint8_t* prow = sP + sidx * 2 * TILE_BYTES + row * 128;
    const float c = 0.125f * 1.4426950408889634f;  // softmax scale * log2(e)
    int ns = 0;                          // items this stream has processed (barrier phase counter)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int pr = item % NPAIR, h = (item / NPAIR) % NH, seq = item / (NPAIR * NH);
      const int qt = 2 * pr + sidx;
      if (qt >= NQT) continue;           // stream B idles on the odd last tile
      const int q_row = qt * QT + row;
      float o[HD];
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] = 0.f;
      float m_run = -INFINITY, m_acc = -INFINITY, l = 0.f;

      // fold O_jj (computed with P relative to max `mj`) into the register accumulator
      auto acc_o = [&](int jj, float mj) {
        uint32_t r0[32], r1[32];
        ptx::tmem_ld32(tO + (jj & 1) * 64 + lane_addr, r0);
        ptx::tmem_ld32(tO + (jj & 1) * 64 + 32 + lane_addr, r1);
        ptx::tmem_ld_wait();
        const float alpha = ex2((m_acc - mj) * c);
        m_acc = mj;
#pragma unroll
        for (int d = 0; d < 32; ++d) {
          o[d] = fmaf(o[d], alpha, __uint_as_float(r0[d]));
          o[32 + d] = fmaf(o[32 + d], alpha, __uint_as_float(r1[d]));
        }
      };

      for (int j = 0; j < NB; ++j) {
        const int gbs = ns * NB + j;
        ptx::mbar_wait(&s_full[sidx], gbs & 1);
        ptx::tc_fence_after();
        uint32_t sr[4][32];
        const bool last = j == NB - 1;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          if (!(last && ch * 32 >= LAST_N)) ptx::tmem_ld32(tS + ch * 32 + lane_addr, sr[ch]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&s_empty[sidx]);  // S is in registers: the next Q K^T may overwrite it
        const int key0 = j * KB;
        const float m_prev = m_run;
        float mx = m_run;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            float v = __uint_as_float(sr[ch][e]);
            if (last && key0 + ch * 32 + e >= SEQ) v = -INFINITY;  // also covers the columns that were not loaded
            sr[ch][e] = __float_as_uint(v);
            mx = fmaxf(mx, v);
          }
        const float alpha = ex2((m_run - mx) * c);
        m_run = mx;
        const float mc = mx * c;
        // P is single-buffered per stream: the previous block's P V must have consumed it
        if (j > 0) ptx::mbar_wait(&pv_done[sidx], (gbs - 1) & 1);
        float rs = 0.f;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k8 = (ch * 4 + i) * 8;            // first key of this 8-key group inside the block
            if (last && k8 >= LAST_N) continue;          // beyond the K = 80 the last P V reads
            uint32_t pk[4] = {0u, 0u, 0u, 0u};
            if (!(last && k8 >= LAST_KEYS)) {            // fully masked groups are zeros without MUFU work
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                const float p0 = ex2(fmaf(__uint_as_float(sr[ch][i * 8 + 2 * w]), c, -mc));
                const float p1 = ex2(fmaf(__uint_as_float(sr[ch][i * 8 + 2 * w + 1]), c, -mc));
                rs += p0 + p1;
                pk[w] = pack_bf16(p0, p1);
              }
            }
            const int chunk = (ch & 1) * 4 + i;
            uint4* dst = reinterpret_cast<uint4*>(prow + (ch >> 1) * TILE_BYTES + ((chunk ^ (row & 7)) << 4));
            *dst = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        l = fmaf(l, alpha, rs);
        ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        ptx::tc_fence_before();
        ptx::mbar_arrive(&p_full[sidx]);
        // O_{j-1} (other TMEM buffer than the P V just enabled) -> registers
        if (j > 0) {
          ptx::tc_fence_after();
          acc_o(j - 1, m_prev);
        }
      }
      ptx::mbar_wait(&pv_done[sidx], (ns * NB + NB - 1) & 1);
      ptx::tc_fence_after();
      acc_o(NB - 1, m_run);
      ptx::tc_fence_before();
      if (q_row < SEQ) {
        const float inv = 1.f / l;
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<long long>(seq) * SEQ + q_row) * LDO + h * HD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 t;
          t.x = pack_bf16(o[8 * i] * inv, o[8 * i + 1] * inv);
          t.y = pack_bf16(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
          t.z = pack_bf16(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
          t.w = pack_bf16(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
          dst[i] = t;
        }
      }
      ++ns;
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

const CUtensorMap& get_tmap_2d_bf16(const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_cols,
                                    uint32_t box_rows);  // gemm_tc.cu

void attention_bf16_tc(const bf16* qkv, bf16* out, int nseq, cudaStream_t s) {
  static bool configured = false;
  static int sms = 0;
  if (!configured) {
    DP_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    int dev;
    DP_CUDA(cudaGetDevice(&dev));
    DP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    configured = true;
  }
  const CUtensorMap& tm = get_tmap_2d_bf16(qkv, LDQ, static_cast<uint64_t>(nseq) * SEQ, LDQ, 64, 128);
  const int items = nseq * NH * NPAIR;
  attention_tc_kernel<<<items < sms ? items : sms, THREADS, SMEM_BYTES, s>>>(tm, out, nseq);
  DP_LAUNCH_CHECK();
}

}  // namespace dp
