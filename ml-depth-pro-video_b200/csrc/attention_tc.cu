// Flash attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Problem: 16 heads x 64, 577 tokens per sequence, softmax(Q K^T / 8) V, no mask (timm Attention ->
// F.scaled_dot_product_attention, wired at src/depth_pro/network/vit_factory.py:97-110).
// qkv is (nseq*577, 3072) bf16 with columns [q | k | v], each 16 heads x 64; out is (nseq*577, 1024).
//
// One persistent CTA per SM runs TWO INDEPENDENT STREAMS.  A stream owns whole (sequence, head) units
// -- 5 query tiles of 128 rows x 5 key blocks of 128 (the 65-key tail runs at N = K = 80 and is masked)
// -- and has its own warps, shared-memory rings, TMEM columns and mbarriers, so the streams drift out of
// phase and one stream's MUFU exp2 work overlaps the other's TMEM loads, maxima and barrier waits:
//   warp 0 / 2     TMA producer : Q tile, K ring (2 deep) and V ring (2 deep), 128B swizzle.  K and V
//                                 have separate barriers: a K stage is recycled as soon as its Q K^T has
//                                 retired, two key blocks before it is needed again.
//   warp 1 / 3     MMA issuer   : S = Q K^T (tcgen05.mma M128 N128 K16 x4) and O += P V (M128 N64 K16 x8,
//                                 A = P K-major from smem, B = V MN-major straight from the TMA tile).  O
//                                 ACCUMULATES IN TMEM over the 5 key blocks of a query tile.
//   warps 4-7 / 8-11  softmax   : one thread per query row: tcgen05.ld the 128 scores, row maximum, exp2
//                                 (log2e folded into the scale), P as bf16 into the swizzled smem tile the
//                                 P V MMA reads.  The running maximum is LAZY: O (in TMEM) and l are only
//                                 rescaled when some row's maximum grew by more than 2^8 -- otherwise the
//                                 old reference maximum stays in use (P <= 256, exact after the final 1/l);
//                                 a rescale is a TMEM load / multiply / store of the warp's 32 x 64 slice.
//                                 After the last block: O from TMEM, 1/l, bf16 store.
// 296 stream slots x 2 units = 592 = 37 sequences x 16 heads: a frame's attention is perfectly balanced.
// Register budget: the control warps drop to 40 registers, the softmax warps grow to 232 (setmaxnreg).
#include <atomic>
#include <cstdlib>
#include <type_traits>

#include "attention.cuh"
#include "ptx.cuh"

namespace dp {
#ifdef ATTN_PROFILE
// Phase timing of the softmax warps (scripts/ubench/attn_prof.cu): cycles per phase summed over quadrant-0
// softmax warps of every stream, [8] = number of blocks counted.
__device__ unsigned long long g_attn_prof[10];
__device__ unsigned long long g_attn_prof_sj[5];    // softmax: cycles waiting for S, by key block j
__device__ unsigned long long g_attn_prof_mma[8];   // MMA warp: q_full, k_full, s_empty, QK issue, v_full+o_empty, p_full, PV issue, #blocks
// Event trace of ONE stream (CTA 0, stream 0): clock64 of every hand-shake of its first TRACE_BLOCKS key blocks, so the
// actual ordering of producer / MMA / softmax events can be read off (scripts/ubench/attn_prof.cu prints it).
constexpr int TRACE_BLOCKS = 40, TRACE_EVENTS = 16;
__device__ long long g_attn_trace[TRACE_EVENTS][TRACE_BLOCKS];
#define TRACE(ev, G)                                                                         \
  do {                                                                                       \
    if (blockIdx.x == 0 && sidx == 0 && (G) < TRACE_BLOCKS) g_attn_trace[ev][G] = clock64(); \
  } while (0)
#define PROF_DECL long long prof_acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; long long prof_last = clock64();
#define PROF_T(i)                          \
  {                                        \
    const long long _t = clock64();        \
    prof_acc[i] += _t - prof_last;         \
    prof_last = _t;                        \
  }
#else
#define PROF_DECL
#define PROF_T(i)
#define TRACE(ev, G)
#endif
namespace {

constexpr int SEQ = 577, HD = 64, NH = 16, LDQ = 3 * NH * HD, LDO = NH * HD;
constexpr int QT = 128;                         // query rows per tile
constexpr int KB = 128;                         // keys per block
constexpr int NB = (SEQ + KB - 1) / KB;         // 5 key blocks
constexpr int NQT = (SEQ + QT - 1) / QT;        // 5 query tiles
constexpr int RING = 2;                         // K ring depth = V ring depth
constexpr int THREADS = 384;
constexpr uint32_t TILE_BYTES = 128 * 128;      // 128 rows x 64 bf16 = 16 KB
constexpr int STREAM_TILES = 1 + 2 * RING + 2;  // Q, K ring, V ring, P (two 64-key halves)
constexpr uint32_t SMEM_BYTES = 2 * STREAM_TILES * TILE_BYTES + 1024 + 512;
constexpr int NBAR = 7 + 4 * RING;              // mbarriers per stream

// idesc: D=f32, A=B=bf16, A K-major; B K-major (QK^T) or MN-major (PV: bit 16)
constexpr uint32_t IDESC_QK = ptx::umma_idesc_bf16(128, 128);
constexpr uint32_t IDESC_PV = ptx::umma_idesc_bf16(128, 64) | (1u << 16);
// the last key block holds 577 - 512 = 65 keys: its Q K^T runs at N = 80 and its P V at K = 80
constexpr int LAST_KEYS = SEQ - (NB - 1) * KB;                 // 65
constexpr int LAST_N = (LAST_KEYS + 15) / 16 * 16;            // 80
constexpr uint32_t IDESC_QK_LAST = ptx::umma_idesc_bf16(128, LAST_N);
// smem descriptor high word: SBO = 1024 B, version 1, SWIZZLE_128B (same for K-major and MN-major tiles)
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr float RESCALE_LOG2 = 8.0f;            // lazy maximum: tolerate P up to 2^8

__device__ __forceinline__ uint64_t desc(uint32_t lo) { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; }
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16); }

__device__ __forceinline__ float ex2(float x) {  // MUFU ex2.approx: 2 ulp, -inf -> 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  bf16x2 h = f2_to_h2(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- exp2 variants of the softmax loop (template parameter EXPV, selected by DEPTHPRO_ATTN_EXP) -------------
// For d = 64 the kernel's only hard bound is the MUFU (1024 cycles per 128x128 block against 512 of MMA,
// DESIGN.md §9).  EXPV >= 1 processes the scores as fp32x2 pairs (fma.rn.f32x2 / add.rn.f32x2: half the issue
// slots of the scalar chain); EXPV >= 2 additionally evaluates a compile-time share of the pairs on the FMA pipe
// instead of the MUFU: round-to-nearest split x = n + f by the 1.5 * 2^23 magic add, degree-3 minimax polynomial
// of 2^f on [-0.5, 0.5] (relative error 7.5e-5: 50x below the bf16 rounding of P), exponent patched in with
// an integer add (bits(t) << 23 keeps exactly the low bits of n).  x is clamped at -125 so the exponent field
// cannot wrap; 2^-125 contributes nothing to l.
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
__device__ __forceinline__ F2 add2(F2 a, F2 b) {
  F2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u));
  return r;
}
__device__ __forceinline__ void exp2_poly2(F2 x, float& p0, float& p1) {
  constexpr float MAGIC = 12582912.f;  // 1.5 * 2^23: x + MAGIC rounds x to the nearest integer in the low mantissa bits
  float x0, x1;
  unpack_f2(x, x0, x1);
  const F2 xc = pack_f2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const F2 t = add2(xc, pack_f2(MAGIC, MAGIC));
  const F2 n = add2(t, pack_f2(-MAGIC, -MAGIC));
  const F2 f = fma2(n, pack_f2(-1.f, -1.f), xc);  // x - n, exact, in [-0.5, 0.5]
  F2 p = fma2(f, pack_f2(0.0551716685295105f, 0.0551716685295105f), pack_f2(0.2426111251115799f, 0.2426111251115799f));
  p = fma2(p, f, pack_f2(0.6932609677314758f, 0.6932609677314758f));
  p = fma2(p, f, pack_f2(0.9999280571937561f, 0.9999280571937561f));
  float q0, q1, t0, t1;
  unpack_f2(p, q0, q1);
  unpack_f2(t, t0, t1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}
// ---- kernel variants (template parameter EXPV, selected by DEPTHPRO_ATTN_EXP) -------------------
//    0  scalar chain, every exponential on the MUFU, strict MUFU ping-pong, P through shared memory (round-1 v5 kernel)
//    5  packed fp32x2 chain, MUFU turn handed to the other stream after 12 of a block's 16 eight-key groups (round-1 default)
//   12  = 5 with P handed to the P V MMA through TMEM (tcgen05.st + A-from-TMEM MMA) instead of shared memory
//   13  = 12 with 25 % of the exponentials as a degree-3 polynomial on the FMA pipe
// Measured and deleted from the library in round 2 (profiles/r1_v6_attention_*, profiles/r2_attention_variants.json):
// 1-4 (polynomial shares without hand-over), 6-11 (other hand-over points), 14 (second Q buffer + cross-tile Q K^T
// issue: 87.1 us against 84.4 for variant 12) and 15-18 (P V split in two 64-key halves with their own p_full /
// pv_done barriers, so that P V of keys 0..63 runs under the exponentials of keys 64..127: 94.7-96.4 us against 80.6
// for variant 13 -- the extra barrier waits inside the unrolled exp loop cost more than the P V latency they hide).
template <int EXPV>
__device__ __forceinline__ constexpr bool poly_pair(int g8, int w) {
  return EXPV == 13 && w == 3;
}
template <int EXPV>
__device__ __forceinline__ constexpr bool packed_chain() { return EXPV != 0; }
template <int EXPV>
__device__ __forceinline__ constexpr bool p_in_tmem() { return EXPV >= 12; }
template <int EXPV>
__device__ __forceinline__ constexpr int arrive_at() { return EXPV == 0 ? 16 : 12; }
// Two more round-2 experiments, measured and removed again (profiles/r2_attention_variants.json,
// r2_attention_phase_profile_*.log): variant 12 with every exponential replaced by one FMUL (wrong results on purpose)
// still takes 71.4 us per layer against 82.0 us for variant 13 -- the MUFU is NOT what bounds this kernel, the load /
// MMA-issue / barrier pipeline is (the MMA warp needs ~450 + ~320 cycles to get the 4 + 8 tcgen05.mma of a block through
// the tensor core's queue while the other stream's MMAs interleave, and waits ~430 + ~290 cycles for K / V); and a producer
// that requests K (G + 1) before it waits for V (G)'s stage changes nothing (82.0 us).  An OPTIMISTIC running maximum (no
// row-maximum pre-pass in blocks 1-4 of a tile: exponentials against the running reference, block maximum tracked inside
// the exp loop, block redone if the threshold trips; bit-identical results) was slower too: 85.7 us against 83.1 us.
// A second form of it (variant 20, profiles/r2_attention_variant20_optimistic.*) used the scale invariance of floating
// point instead: NO maximum at all in blocks 1-4 (any reference works as long as nothing overflows), checked after the
// fact on the row sum the block needs anyway (sum < 2^60, else the block is redone against its true maximum).  Correct
// (all parity cases, incl. scores growing by 2^36 per block), no spills, 250 cycles less work per block -- and 84.96 us
// against 82.84: the phase profile shows the exponential phase growing by exactly what the pre-pass lost (1231 -> 1384
// cycles per block, total 2728 -> 2736).  The period of a stream is NOT set by the softmax warp's own instruction chain.
// So were a SPLIT-COLUMN softmax (two warps per row: 16 softmax warps, each 64 of a block's 128 columns, half-row maxima
// exchanged per block through shared memory + a 64-thread named barrier, row sums combined per tile; correct at the first
// run, 101.2 us with / 108.9 us without the polynomial against 81.8 us) and ONE mbarrier arrival per softmax warp instead
// of one per thread (84.0-85.2 us).  ncu (profiles/r2_ncu_full_vit_block_raw.csv): L2 -> SM traffic 0.53 GB per launch =
// 6.1 TB/s at lts__throughput 27 % (the qkv GEMM pulls 11.7 TB/s), so the K / V re-reads are not the bound either.

// Debug counter (dp_debug_counter): how many times a softmax warp took the lazy-maximum RESCALE branch.
__device__ unsigned long long g_attn_rescales;

template <int EXPV>
__global__ void __launch_bounds__(THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmOut, int nseq,
                    int pingpong, int reverse) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sidx = warp < 4 ? (warp >> 1) : ((warp - 4) >> 2);  // stream of this warp
  // per-stream shared memory: [Q | K ring | V ring | P lo, P hi]
  uint8_t* sQ = smem + sidx * STREAM_TILES * TILE_BYTES;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + RING * TILE_BYTES;
  uint8_t* sP = sV + RING * TILE_BYTES;
  uint64_t* bars_all = reinterpret_cast<uint64_t*>(smem + 2 * STREAM_TILES * TILE_BYTES);
  uint64_t* bars = bars_all + sidx * NBAR;
  uint64_t* q_full = bars;                  // 1 (TMA tx)
  uint64_t* q_empty = bars + 1;             // 1 (commit after the tile's last Q K^T)
  uint64_t* s_full = bars + 2;              // 1 (commit)
  uint64_t* s_empty = bars + 3;             // 128 softmax threads
  uint64_t* p_full = bars + 4;              // 128 softmax threads
  uint64_t* pv_done = bars + 5;             // 1 (commit)
  uint64_t* o_empty = bars + 6;             // 128 softmax threads: the tile's O has been read out of TMEM
  uint64_t* k_full = bars + 7;              // RING
  uint64_t* k_empty = k_full + RING;        // RING
  uint64_t* v_full = k_empty + RING;        // RING
  uint64_t* v_empty = v_full + RING;        // RING
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_all + 2 * NBAR);

  const int n_units = nseq * NH;
  const int slot = blockIdx.x + sidx * gridDim.x, n_slots = 2 * gridDim.x;

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&tmQKV);
  if ((warp == 1 || warp == 3) && lane == 0) {
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(q_empty, 1);
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(s_empty, 128);
    ptx::mbar_init(p_full, 128);
    ptx::mbar_init(pv_done, 1);
    ptx::mbar_init(o_empty, 128);
    for (int i = 0; i < RING; ++i) {
      ptx::mbar_init(&k_full[i], 1), ptx::mbar_init(&k_empty[i], 1);
      ptx::mbar_init(&v_full[i], 1), ptx::mbar_init(&v_empty[i], 1);
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
  // TMEM columns of stream s: S at [256 s, 256 s + 128), O at [256 s + 128, 256 s + 192)
  const uint32_t tmem_base = *tmem_slot + sidx * 256;
  ptx::pdl_wait();  // PDL: the qkv GEMM's output is visible from here on
  ptx::pdl_launch_dependents();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if ((warp & 1) == 0) {
      // ---------------------------------------------------------------- TMA producer
      int T = 0;  // query tiles this stream has started
      for (int uu = slot; uu < n_units; uu += n_slots) {
        const int u = reverse ? n_units - 1 - uu : uu;  // (sequence, head) units from the last to the first
        const int h = u % NH, row0 = (u / NH) * SEQ;
        for (int qt = 0; qt < NQT; ++qt, ++T) {
          ptx::mbar_wait(q_empty, (T & 1) ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(q_full, TILE_BYTES);
            ptx::tma_load_2d(sQ, &tmQKV, q_full, h * HD, row0 + qt * QT);
          }
          __syncwarp();
          for (int j = 0; j < NB; ++j) {
            const int G = T * NB + j, st = G % RING;
            const uint32_t ph = ((G / RING) & 1) ^ 1;
            ptx::mbar_wait(&k_empty[st], ph);
            if (lane == 0) TRACE(0, G);
            if (ptx::elect_one()) {
              ptx::mbar_expect_tx(&k_full[st], TILE_BYTES);
              ptx::tma_load_2d(sK + st * TILE_BYTES, &tmQKV, &k_full[st], NH * HD + h * HD, row0 + j * KB);
            }
            __syncwarp();
            ptx::mbar_wait(&v_empty[st], ph);
            if (lane == 0) TRACE(1, G);
            if (ptx::elect_one()) {
              ptx::mbar_expect_tx(&v_full[st], TILE_BYTES);
              ptx::tma_load_2d(sV + st * TILE_BYTES, &tmQKV, &v_full[st], 2 * NH * HD + h * HD, row0 + j * KB);
            }
            __syncwarp();
          }
        }
      }
    } else {
      // ---------------------------------------------------------------- MMA issuer
      const uint32_t q_lo = desc_lo(ptx::smem_u32(sQ));
      const uint32_t k_lo = q_lo + (TILE_BYTES >> 4);
      const uint32_t v_lo = k_lo + (RING * TILE_BYTES >> 4);
      const uint32_t p_lo = v_lo + (RING * TILE_BYTES >> 4);
      const uint32_t tS = tmem_base, tO = tmem_base + 128;
      // O += P V of key block jj of tile Tt, then release that V stage
#ifdef ATTN_PROFILE
      long long mp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      long long mlast = clock64();
#define MPROF(i) { const long long _t = clock64(); mp[i] += _t - mlast; mlast = _t; }
#else
#define MPROF(i)
#endif
      auto issue_pv = [&](int Tt, int jj) {
            const int G = Tt * NB + jj, st = G % RING;
            ptx::mbar_wait(&v_full[st], (G / RING) & 1);
            if (jj == 0) ptx::mbar_wait(o_empty, (Tt & 1) ^ 1);  // the previous tile's O has been read out
            MPROF(4)
            if (lane == 0) TRACE(5, G);
            const int nks = jj == NB - 1 ? LAST_N / 16 : KB / 16;
            ptx::mbar_wait(p_full, G & 1);
            MPROF(5)
            if (lane == 0) TRACE(6, G);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < KB / 16; ++ks) {
                if (ks >= nks) break;
                // A = P: 64-key half (ks >> 2), +32 B per K=16 step; B = V (MN-major): +16 keys = 2048 B
                const uint64_t da = desc(p_lo + (ks >> 2) * (TILE_BYTES >> 4) + (ks & 3) * 2);
                const uint64_t db = desc(v_lo + st * (TILE_BYTES >> 4) + ks * (2048 >> 4));
                if constexpr (p_in_tmem<EXPV>()) ptx::umma_bf16_ts(tO, tmem_base + 192 + ks * 8, db, IDESC_PV, (jj | ks) != 0);
                else ptx::umma_bf16(tO, da, db, IDESC_PV, (jj | ks) != 0);
              }
              ptx::umma_commit(pv_done);
              ptx::umma_commit(&v_empty[st]);
            }
            __syncwarp();
            MPROF(6)
            if (lane == 0) TRACE(7, G);
      };
      // (Issuing the last P V of tile T after the first Q K^T of tile T+1 was measured and changed nothing: the
      // single-buffered Q tile and the first K block of the next tile arrive too late for it to matter.)
      {
      int T = 0;
      for (int u = slot; u < n_units; u += n_slots) {
        for (int qt = 0; qt < NQT; ++qt, ++T) {
          ptx::mbar_wait(q_full, T & 1);
          MPROF(0)
          for (int j = 0; j < NB; ++j) {
            const int G = T * NB + j, st = G % RING;
            ptx::mbar_wait(&k_full[st], (G / RING) & 1);
            MPROF(1)
            if (lane == 0) TRACE(2, G);
            ptx::mbar_wait(s_empty, (G & 1) ^ 1);
            MPROF(2)
            if (lane == 0) TRACE(3, G);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < HD / 16; ++ks)
                ptx::umma_bf16(tS, desc(q_lo + ks * 2), desc(k_lo + st * (TILE_BYTES >> 4) + ks * 2),
                               j == NB - 1 ? IDESC_QK_LAST : IDESC_QK, ks != 0);
              ptx::umma_commit(s_full);
              ptx::umma_commit(&k_empty[st]);
              if (j == NB - 1) ptx::umma_commit(q_empty);  // the Q tile is free once its last Q K^T retires
            }
            __syncwarp();
            MPROF(3)
            if (lane == 0) TRACE(4, G);
#ifdef ATTN_PROFILE
            mp[7] += 1;
#endif
            if (j > 0) issue_pv(T, j - 1);
          }
          issue_pv(T, NB - 1);
        }
      }
      }
#ifdef ATTN_PROFILE
      if (lane == 0)
        for (int i = 0; i < 8; ++i) atomicAdd(&g_attn_prof_mma[i], static_cast<unsigned long long>(mp[i]));
#endif
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ------------------------------------------------------------------ softmax
    const int q = warp & 3;             // TMEM lane quadrant
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr, tO = tmem_base + 128 + lane_addr, tP = tmem_base + 192 + lane_addr;
    const uint32_t prow = ptx::smem_u32(sP) + row * 128;
    const float c = 0.125f * 1.4426950408889634f;  // softmax scale * log2(e)
    // MUFU ping-pong.  The softmax warps of quadrant q of both streams sit on the same SM sub-partition
    // and share its MUFU unit.  Left alone the two streams phase-LOCK: when their exp2 phases collide
    // both slow down, finish together and then also do their TMEM loads / maxima / barrier waits
    // together, leaving the MUFU idle half of the time (ncu: XU pipe 46 %).  A pair of named barriers
    // per quadrant makes the exp2 phases alternate strictly, so one stream's exponentials always run
    // under the other's non-MUFU work.  Only the first min(blocks A, blocks B) blocks take part.
    auto units_of = [&](int sl) { return sl < n_units ? (n_units - sl + n_slots - 1) / n_slots : 0; };
    const int u_a = units_of(blockIdx.x), u_b = units_of(blockIdx.x + gridDim.x);
    const int n_pp = pingpong ? (u_a < u_b ? u_a : u_b) * NQT * NB : 0;   // blocks that ping-pong
    const int bar_mine = 1 + q + 4 * sidx, bar_peer = 1 + q + 4 * (sidx ^ 1);
    int T = 0;
    PROF_DECL
    for (int uu = slot; uu < n_units; uu += n_slots) {
      const int u = reverse ? n_units - 1 - uu : uu;
      const int h = u % NH, seq = u / NH;
      for (int qt = 0; qt < NQT; ++qt, ++T) {
        float m_ref = -INFINITY;  // reference maximum the exponentials (and O, l) are relative to
        float l = 0.f;

        auto block = [&](int j, auto last_tag) {
          constexpr bool LAST = decltype(last_tag)::value;
          constexpr int NCH = LAST ? (LAST_N + 31) / 32 : 4;  // 32-column chunks to load
          constexpr int NKEY = LAST ? LAST_KEYS : KB;         // valid keys in this block
          constexpr int NG = (LAST ? LAST_N : KB) / 8;        // 8-key groups the P V MMA reads
          const int G = T * NB + j;
          if (q == 0 && lane == 0) TRACE(8, G);
          ptx::mbar_wait(s_full, G & 1);
          ptx::tc_fence_after();
          if (q == 0 && lane == 0) TRACE(9, G);
#ifdef ATTN_PROFILE
          if (q == 0 && lane == 0) atomicAdd(&g_attn_prof_sj[j], static_cast<unsigned long long>(clock64() - prof_last));
#endif
          PROF_T(0)  // waiting for S
          uint32_t sr[NCH][32];
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) ptx::tmem_ld32(tS + ch * 32, sr[ch]);
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
          ptx::mbar_arrive(s_empty);  // S is in registers: the next Q K^T may overwrite it
          PROF_T(1)  // TMEM load
          if (q == 0 && lane == 0) TRACE(10, G);
          // row maximum over the valid keys, four independent chains
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int k = 0; k < NKEY; ++k) mx4[k & 3] = fmaxf(mx4[k & 3], __uint_as_float(sr[k >> 5][k & 31]));
          const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
          PROF_T(2)  // row maximum
          // P and O are single-buffered: the previous block's P V must have retired
          if (q == 0 && lane == 0) TRACE(14, G);
          if (j > 0) {
            ptx::mbar_wait(pv_done, (G - 1) & 1);
            ptx::tc_fence_after();
          }
          PROF_T(3)  // waiting for the previous P V
          if (q == 0 && lane == 0) TRACE(11, G);
          if (j == 0) {
            m_ref = mx;
          } else if (__any_sync(0xffffffffu, (mx - m_ref) * c > RESCALE_LOG2)) {
            if (lane == 0) atomicAdd(&g_attn_rescales, 1ull);
            const float m_new = fmaxf(m_ref, mx);
            const float alpha = ex2((m_ref - m_new) * c);
            m_ref = m_new;
            l *= alpha;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t o[32];
              ptx::tmem_ld32(tO + hh * 32, o);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int d = 0; d < 32; ++d) o[d] = __float_as_uint(__uint_as_float(o[d]) * alpha);
              ptx::tmem_st32(tO + hh * 32, o);
            }
            ptx::tmem_st_wait();
          }
          // my turn on the MUFU: stream A leads, B follows A's block G, A's block G follows B's block G-1
          PROF_T(4)  // lazy-maximum check / rescale
          if (G < n_pp && (sidx == 1 || G > 0)) asm volatile("bar.sync %0, 64;" ::"r"(bar_mine) : "memory");
          PROF_T(5)  // waiting for the MUFU turn
          if (q == 0 && lane == 0) TRACE(12, G);
          if (j == 0) {  // the previous tile's output store must have drained this warp's P rows
            if (lane == 0) ptx::tma_store_wait_read();
            __syncwarp();
          }
          const float mc = m_ref * c;
          constexpr int ARR = arrive_at<EXPV>() >= 16 ? 16 : (LAST ? arrive_at<EXPV>() * NG / 16 : arrive_at<EXPV>());
          const bool pp_arrive = G < n_pp && (sidx == 0 || G < n_pp - 1);
          if constexpr (packed_chain<EXPV>()) {
            const F2 c2 = pack_f2(c, c), nmc2 = pack_f2(-mc, -mc);
            F2 rs2[2] = {pack_f2(0.f, 0.f), pack_f2(0.f, 0.f)};
            [[maybe_unused]] uint32_t pacc[32];  // p_in_tmem: the packed pairs of 64 keys, stored with one tcgen05.st
#pragma unroll
            for (int g8 = 0; g8 < NG; ++g8) {
              if (ARR < 16 && g8 == ARR && pp_arrive) asm volatile("bar.arrive %0, 64;" ::"r"(bar_peer) : "memory");
              uint32_t pk[4] = {0u, 0u, 0u, 0u};
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                const int k0 = g8 * 8 + 2 * w;  // even: k0 and k0 + 1 sit in the same 32-column chunk
                if (k0 < NKEY) {
                  const F2 x = fma2(pack_f2(__uint_as_float(sr[k0 >> 5][k0 & 31]), __uint_as_float(sr[k0 >> 5][(k0 & 31) + 1])),
                                    c2, nmc2);
                  float p0, p1;
                  if (poly_pair<EXPV>(g8, w)) {
                    exp2_poly2(x, p0, p1);
                  } else {
                    float x0, x1;
                    unpack_f2(x, x0, x1);
                    p0 = ex2(x0), p1 = ex2(x1);
                  }
                  if (k0 + 1 >= NKEY) p1 = 0.f;
                  rs2[w & 1] = add2(rs2[w & 1], pack_f2(p0, p1));
                  pk[w] = pack_bf16(p0, p1);
                }
              }
              if constexpr (p_in_tmem<EXPV>()) {
                // column c of the stream's P region holds keys (2c, 2c + 1) of this thread's row
#pragma unroll
                for (int w = 0; w < 4; ++w) pacc[(g8 & 7) * 4 + w] = pk[w];
                if ((g8 & 7) == 7) ptx::tmem_st32(tP + (g8 >> 3) * 32, pacc);
                if (LAST && g8 == NG - 1) {  // 80-key tail: keys 64..79 -> columns 32..39
                  uint32_t tail[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) tail[i] = pacc[i];
                  ptx::tmem_st8(tP + 32, tail);
                }
              } else {
                ptx::sts_u4(prow + (g8 >> 3) * TILE_BYTES + (((g8 & 7) ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
              }
            }
            if constexpr (p_in_tmem<EXPV>()) ptx::tmem_st_wait();
            float a0, a1, b0, b1;
            unpack_f2(rs2[0], a0, a1);
            unpack_f2(rs2[1], b0, b1);
            l += (a0 + a1) + (b0 + b1);
          } else {
          float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int g8 = 0; g8 < NG; ++g8) {
            if (ARR < 16 && g8 == ARR && pp_arrive) asm volatile("bar.arrive %0, 64;" ::"r"(bar_peer) : "memory");
            uint32_t pk[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const int k0 = g8 * 8 + 2 * w;
              float p0 = 0.f, p1 = 0.f;
              if (k0 < NKEY) p0 = ex2(fmaf(__uint_as_float(sr[k0 >> 5][k0 & 31]), c, -mc));
              if (k0 + 1 < NKEY) p1 = ex2(fmaf(__uint_as_float(sr[(k0 + 1) >> 5][(k0 + 1) & 31]), c, -mc));
              if (k0 < NKEY) rs4[w] += p0 + p1;
              pk[w] = pack_bf16(p0, p1);
            }
            // 64-key half (g8 >> 3), 16-byte chunk (g8 & 7) of this row, 128B swizzle
            ptx::sts_u4(prow + (g8 >> 3) * TILE_BYTES + (((g8 & 7) ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
          }
          l += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
          }
          PROF_T(6)  // exponentials + P stores
          if (ARR >= 16 && pp_arrive) asm volatile("bar.arrive %0, 64;" ::"r"(bar_peer) : "memory");
          if constexpr (!p_in_tmem<EXPV>()) ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
          ptx::tc_fence_before();
          ptx::mbar_arrive(p_full);
          PROF_T(7)  // hand-off
          if (q == 0 && lane == 0) TRACE(13, G);
#ifdef ATTN_PROFILE
          prof_acc[8] += 1;
#endif
        };
        for (int j = 0; j < NB - 1; ++j) block(j, std::false_type{});
        block(NB - 1, std::true_type{});

        // the tile's O: TMEM -> registers -> 1/l -> bf16
        ptx::mbar_wait(pv_done, (T * NB + NB - 1) & 1);
        ptx::tc_fence_after();
        uint32_t o0[32], o1[32];
        ptx::tmem_ld32(tO, o0);
        ptx::tmem_ld32(tO + 32, o1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(o_empty);
        // 1/l, bf16, then out through the (now idle) low P tile: 32 rows x 128 B per warp in the 128B-swizzle
        // layout, one TMA store per warp.  tmOut is (sequence, token, channel): rows >= 577 are clipped.
        const float inv = 1.f / l;
        const uint32_t orow = ptx::smem_u32(sP) + row * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          ptx::sts_u4(orow + ((i ^ (row & 7)) << 4),
                      pack_bf16(__uint_as_float(o0[8 * i]) * inv, __uint_as_float(o0[8 * i + 1]) * inv),
                      pack_bf16(__uint_as_float(o0[8 * i + 2]) * inv, __uint_as_float(o0[8 * i + 3]) * inv),
                      pack_bf16(__uint_as_float(o0[8 * i + 4]) * inv, __uint_as_float(o0[8 * i + 5]) * inv),
                      pack_bf16(__uint_as_float(o0[8 * i + 6]) * inv, __uint_as_float(o0[8 * i + 7]) * inv));
#pragma unroll
        for (int i = 0; i < 4; ++i)
          ptx::sts_u4(orow + (((4 + i) ^ (row & 7)) << 4),
                      pack_bf16(__uint_as_float(o1[8 * i]) * inv, __uint_as_float(o1[8 * i + 1]) * inv),
                      pack_bf16(__uint_as_float(o1[8 * i + 2]) * inv, __uint_as_float(o1[8 * i + 3]) * inv),
                      pack_bf16(__uint_as_float(o1[8 * i + 4]) * inv, __uint_as_float(o1[8 * i + 5]) * inv),
                      pack_bf16(__uint_as_float(o1[8 * i + 6]) * inv, __uint_as_float(o1[8 * i + 7]) * inv));
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0 && qt * QT + q * 32 < SEQ) {
          ptx::tma_store_3d(&tmOut, ptx::smem_u32(sP) + q * 4096, h * HD, qt * QT + q * 32, seq);
          ptx::tma_store_commit();
        }
#ifdef ATTN_PROFILE
        {  // O read-out, normalisation, store: accounted separately
          const long long _t = clock64();
          if (q == 0 && lane == 0) atomicAdd(&g_attn_prof[9], static_cast<unsigned long long>(_t - prof_last));
          prof_last = _t;
        }
#endif
      }
    }
#ifdef ATTN_PROFILE
    if (q == 0 && lane == 0)
      for (int i = 0; i < 9; ++i) atomicAdd(&g_attn_prof[i], static_cast<unsigned long long>(prof_acc[i]));
#endif
  }

  if (warp >= 4 && lane == 0) ptx::tma_store_wait_read();  // smem must outlive the bulk stores
  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(*tmem_slot, 512);
  }
}

}  // namespace

CUtensorMap get_tmap_bf16(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box);  // gemm_tc.cu (cached, thread-safe, returned by value)
CUtensorMap get_tmap_2d_bf16(const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_cols,
                             uint32_t box_rows);  // gemm_tc.cu

#ifdef ATTN_PROFILE
void attn_prof_read(unsigned long long* host10, bool reset) {
  DP_CUDA(cudaDeviceSynchronize());
  DP_CUDA(cudaMemcpyFromSymbol(host10, g_attn_prof, sizeof(unsigned long long) * 10));
  DP_CUDA(cudaMemcpyFromSymbol(host10 + 10, g_attn_prof_sj, sizeof(unsigned long long) * 5));
  DP_CUDA(cudaMemcpyFromSymbol(host10 + 15, g_attn_prof_mma, sizeof(unsigned long long) * 8));
  DP_CUDA(cudaMemcpyFromSymbol(host10 + 23, g_attn_trace, sizeof(long long) * TRACE_EVENTS * TRACE_BLOCKS));
  if (reset) {
    unsigned long long z[10] = {0};
    DP_CUDA(cudaMemcpyToSymbol(g_attn_prof, z, sizeof(z)));
    DP_CUDA(cudaMemcpyToSymbol(g_attn_prof_sj, z, sizeof(unsigned long long) * 5));
    DP_CUDA(cudaMemcpyToSymbol(g_attn_prof_mma, z, sizeof(unsigned long long) * 8));
  }
}
#endif

unsigned long long attention_tc_rescale_count(bool reset) {
  unsigned long long v = 0;
  DP_CUDA(cudaDeviceSynchronize());
  DP_CUDA(cudaMemcpyFromSymbol(&v, g_attn_rescales, sizeof(v)));
  if (reset) {
    const unsigned long long z = 0;
    DP_CUDA(cudaMemcpyToSymbol(g_attn_rescales, &z, sizeof(z)));
  }
  return v;
}

template <int EXPV>
static void launch_attention(const CUtensorMap& tm, const CUtensorMap& tmo, int nseq, int ctas, int pingpong, int reverse,
                             cudaStream_t s) {
  constexpr uint32_t SMEM = SMEM_BYTES;
  static_assert(SMEM <= 232448, "more than 227 KB of shared memory");
  static std::atomic<unsigned long long> configured{0};
  if (first_use_on_device(configured)) {
    DP_CUDA(cudaFuncSetAttribute(attention_tc_kernel<EXPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  launch_pdl(attention_tc_kernel<EXPV>, dim3(ctas), dim3(THREADS), SMEM, s, tm, tmo, nseq, pingpong, reverse);
}

// DEPTHPRO_ATTN_EXP selects a variant (list at the top of this file); DEPTHPRO_ATTN_PINGPONG = 0 lets the two streams'
// exp phases overlap freely (slower).
static std::atomic<int> g_expv{-1}, g_pingpong{1}, g_attn_sm_limit{0};
void attention_tc_set_sm_limit(int sms) {  // experiment knob: at most this many CTAs (= SMs); 0 = all
  g_attn_sm_limit = sms;
  bump_config_epoch();
}

constexpr int ATTN_EXP_DEFAULT = 13;  // measured on B200 (profiles/r2_attention_variants.json)
static bool variant_compiled(int v) { return v == 0 || v == 5 || v == 12 || v == 13; }

void attention_tc_set_variant(int expv, int pingpong) {
  bump_config_epoch();
  if (expv < 0) {  // back to the environment's / built-in default on the next launch
    g_expv = -1;
    return;
  }
  if (!variant_compiled(expv)) throw std::runtime_error("attention variant not compiled in (0, 5, 12, 13)");
  g_expv = expv, g_pingpong = pingpong != 0;
}

void attention_bf16_tc(const bf16* qkv, bf16* out, int nseq, cudaStream_t s, int reverse) {
  static const int sms = [] {
    int dev, v;
    DP_CUDA(cudaGetDevice(&dev));
    DP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    return v;
  }();
  if (g_expv < 0) {  // first call and no explicit choice: the environment, else the default
    const char* e = getenv("DEPTHPRO_ATTN_EXP");
    const char* p = getenv("DEPTHPRO_ATTN_PINGPONG");
    attention_tc_set_variant(e && atoi(e) >= 0 ? atoi(e) : ATTN_EXP_DEFAULT, p ? atoi(p) : 1);
  }
  const int expv = g_expv, pingpong = g_pingpong;
  const CUtensorMap tm = get_tmap_2d_bf16(qkv, LDQ, static_cast<uint64_t>(nseq) * SEQ, LDQ, 64, 128);
  const uint64_t od[3] = {(uint64_t)LDO, (uint64_t)SEQ, (uint64_t)nseq};
  const uint64_t os[2] = {(uint64_t)LDO * 2, (uint64_t)SEQ * LDO * 2};
  const uint32_t ob[3] = {64, 32, 1};
  const CUtensorMap tmo = get_tmap_bf16(out, 3, od, os, ob);
  int ctas = (nseq * NH + 1) / 2;  // two streams per CTA, one (sequence, head) unit at a time each
  if (ctas > sms) ctas = sms;
  const int lim = g_attn_sm_limit.load();
  if (lim > 0 && ctas > lim) ctas = lim;
  switch (expv) {
    case 0: launch_attention<0>(tm, tmo, nseq, ctas, pingpong, reverse, s); break;
    case 5: launch_attention<5>(tm, tmo, nseq, ctas, pingpong, reverse, s); break;
    case 12: launch_attention<12>(tm, tmo, nseq, ctas, pingpong, reverse, s); break;
    case 13: launch_attention<13>(tm, tmo, nseq, ctas, pingpong, reverse, s); break;
    default: throw std::runtime_error("attention variant not compiled in");
  }
  DP_LAUNCH_CHECK();
}

}  // namespace dp
