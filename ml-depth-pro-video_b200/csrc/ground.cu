// Ground normalisation of an unprojected point cloud, the step right after
// depth_to_3d in the reference:
//   normalize_point_cloud_to_ground   img_to_normalized_pointcloud.py:880-975
//   grid_based_ground_adjustment      img_to_normalized_pointcloud.py:977-1118
// Points are float32 (N,3), updated in place; all arithmetic is done in double like the reference's
// numpy code (it works on float64 points), so thresholds and order statistics select the same elements.
// The plane FIT (RANSAC / L-BFGS on a 50 k sample, :376-816) stays on the CPU and hands over (normal, d).
//
// The only non-trivial piece is np.percentile on a data-dependent subset -- globally (2nd percentile of the
// near-plane heights) and per XZ grid cell (5th percentile of the low points of each of 20 x 20 cells).
// It is computed EXACTLY, without sorting, by a segmented radix select on order-preserving 32-bit keys:
// three histogram passes (11 + 11 + 10 bits) narrow every cell's k-th smallest key, one more pass finds the
// next larger key when the (k+1)-th order statistic is not a duplicate, and numpy's linear interpolation
// (lib/_function_base_impl.py _lerp) finishes in double.  Since round 2 each entry point is ONE cooperative
// kernel (grid-wide barriers between the phases) working on a compacted copy of the percentile's population.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace dp {
namespace {

constexpr int SEL_BINS = 2048;
constexpr int T = 256;

struct SelCell {
  unsigned count_all;  // points of the cell
  unsigned count;      // points of the cell that satisfy the predicate (the percentile's population)
  unsigned k;          // rank of the lower order statistic (0-based)
  unsigned rank;       // residual rank inside the current prefix
  unsigned prefix;     // key bits fixed so far
  unsigned less;       // population elements strictly smaller than the current prefix range
  unsigned eq;         // multiplicity of the k-th key (after the last pass)
  unsigned next;       // smallest key greater than the k-th key (0xffffffff = none)
  int active;
  int pad;
  double gamma;        // interpolation weight between the k-th and (k+1)-th order statistics
  double value;        // the percentile
};

__device__ __forceinline__ unsigned f2key(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ---- second generation (round 2): ONE cooperative launch per entry point ---------------------------------------------
// Round 1 ran 12 (normalise) / 16 (grid adjustment) launches, each re-reading all n points (24 B of xyz + dist per point
// and pass): 197 / 308 us for a 1080p cloud, 0.04 / 0.017 of the copy bandwidth.  Now each entry point is one persistent
// kernel launched with cudaLaunchCooperativeKernel; its phases are separated by grid-wide barriers:
//   * the first sweep over the points does the geometry AND compacts the percentile's population into a dense array of
//     32-bit order-preserving keys (+ cell ids), so the three radix-select passes and the next-key pass read 4-6 bytes per
//     POPULATION element instead of 24 bytes per point;
//   * the plane distances survive as one flag byte per point instead of a double;
//   * the single-block / per-cell scans between the passes are done by the blocks themselves after a barrier.
// The arithmetic (double, same operation order) and therefore every output bit is unchanged: the golden and oracle tests
// of round 1 run against it as they are.  Data another block wrote in an earlier phase is read with __ldcg (L2): L1 is not
// coherent across the grid barrier.
struct SelParams {
  double q;
  unsigned min_all, min_pop;
};

// numpy percentile, method 'linear': virtual index = n q + (1 - q) - 1, k = floor, gamma = fraction
__device__ __forceinline__ void sel_begin(SelCell& s, double q, unsigned min_all, unsigned min_pop) {
  s.active = (s.count_all >= min_all && s.count >= min_pop && s.count > 0) ? 1 : 0;
  s.prefix = 0, s.less = 0, s.eq = 0, s.next = 0xffffffffu, s.value = 0.0;
  if (!s.active) return;
  const double nq = static_cast<double>(s.count);
  double v = __dadd_rn(__dadd_rn(__dmul_rn(nq, q), __dadd_rn(1.0, __dmul_rn(q, -1.0))), -1.0);
  if (v < 0.0) v = 0.0;
  if (v > nq - 1.0) v = nq - 1.0;
  const double fl = floor(v);
  s.k = static_cast<unsigned>(fl);
  s.rank = s.k;
  s.gamma = v - fl;
}
// State in GLOBAL memory (LOCAL = false): every field other blocks may have updated since this SM last touched the cell is
// read through L2.  LOCAL = true: the block's own copy in shared memory.
template <typename V>
__device__ __forceinline__ V ld_state(const V* p, bool local) { return local ? *p : __ldcg(p); }
template <bool LOCAL = false>
__device__ __forceinline__ void sel_finish(SelCell& s) {
  if (!ld_state(&s.active, LOCAL)) return;
  const unsigned prefix = ld_state(&s.prefix, LOCAL), less = ld_state(&s.less, LOCAL), eq = ld_state(&s.eq, LOCAL),
                 next = ld_state(&s.next, LOCAL), k = ld_state(&s.k, LOCAL), count = ld_state(&s.count, LOCAL);
  const double a = static_cast<double>(key2f(prefix));
  double b = a;
  // ranks less .. less + eq - 1 hold the k-th key; the (k+1)-th is a duplicate unless k is the last of them
  if (k + 1 >= less + eq && k + 1 < count && next != 0xffffffffu) b = static_cast<double>(key2f(next));
  // numpy _lerp: a + (b - a) t, evaluated from the other end for t >= 0.5
  const double diff = __dadd_rn(b, -a), t = ld_state(&s.gamma, LOCAL);
  double r = __dadd_rn(a, __dmul_rn(diff, t));
  if (t >= 0.5) r = __dadd_rn(b, -__dmul_rn(diff, __dadd_rn(1.0, -t)));
  if (t == 0.0) r = a;  // numpy's gamma == 0 short cut keeps the lower statistic exactly
  s.value = r;
}
// One block: the digit of cell `s` whose cumulative count crosses the residual rank; clears the histogram row (LOCAL: the
// state is the block's shared-memory copy, the histogram is shared by all blocks and left alone).
// All T threads of the block call it (contains __syncthreads); `part` is T + 2 words of shared memory.
template <bool LOCAL = false>
__device__ void sel_scan_block(SelCell* sp, unsigned* h, int shift, int bits, int last, unsigned* part) {
  const int nb = 1 << bits, per = SEL_BINS / T;  // 8 consecutive bins per thread
  unsigned loc[SEL_BINS / T];
  unsigned sum = 0;
#pragma unroll
  for (int j = 0; j < per; ++j) {
    const int b = threadIdx.x * per + j;
    loc[j] = b < nb ? __ldcg(&h[b]) : 0u;
    sum += loc[j];
    if (!LOCAL) h[b] = 0;
  }
  const int active = ld_state(&sp->active, LOCAL);
  const unsigned rank = ld_state(&sp->rank, LOCAL);
  __syncthreads();  // `part` may still be in use by the previous call
  part[threadIdx.x] = sum;
  __syncthreads();
  if (!active) return;
  if (threadIdx.x == 0) {
    unsigned acc = 0;
    int t = 0;
    for (; t < T; ++t) {
      if (acc + part[t] > rank) break;
      acc += part[t];
    }
    part[T] = static_cast<unsigned>(t < T ? t : T - 1);  // thread that owns the crossing
    part[T + 1] = acc;
  }
  __syncthreads();
  if (threadIdx.x == static_cast<int>(part[T])) {
    unsigned acc = part[T + 1];
    int j = 0;
    for (; j < per - 1; ++j) {
      if (acc + loc[j] > rank) break;
      acc += loc[j];
    }
    const unsigned dgt = static_cast<unsigned>(threadIdx.x * per + j);
    sp->prefix = ld_state(&sp->prefix, LOCAL) | (dgt << shift);
    sp->less = ld_state(&sp->less, LOCAL) + acc;
    sp->rank = rank - acc;
    if (last) sp->eq = loc[j];
  }
}

// Append `key` (+ cell) to THIS BLOCK's segment of the compact population arrays.  Every block owns the same points in
// every phase, so its population lives in a private segment `seg` = [blockIdx.x * cap, ...) and is counted in shared
// memory: no global atomic per element (one same-address global atomic per warp was what round 1's count pass, and the
// first version of this kernel, spent most of their time on).
__device__ __forceinline__ void pop_append(bool in, unsigned key, unsigned short cell, unsigned* s_count,
                                           unsigned* __restrict__ seg_keys, unsigned short* __restrict__ seg_cell) {
  const unsigned b = __ballot_sync(0xffffffffu, in);
  if (!b) return;
  const int lane = threadIdx.x & 31, leader = __ffs(b) - 1;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(s_count, static_cast<unsigned>(__popc(b)));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (!in) return;
  const unsigned pos = base + __popc(b & ((1u << lane) - 1u));
  seg_keys[pos] = key;
  if (seg_cell) seg_cell[pos] = cell;
}

struct GroundXf {
  double r[9];    // rotation (identity when the plane normal is within acos(0.99) of +-y)
  double n[3];    // unit normal
  double d;
  double shift;   // y offset of the rotated plane: -d / (R normal).y, 0 without rotation
};


struct NormArgs {
  float* xyz;
  long long n;
  GroundXf xf;
  uint8_t* flags;             // bit 0: |dist| < 0.1 (population of the percentile), bit 1: |dist| < 0.05 ("ground")
  unsigned* keys;             // compact population: order-preserving keys of the rotated heights, one segment per block
  long long seg;              // segment capacity (points a block owns)
  unsigned* pop_count;
  unsigned* next_key;         // smallest population key above the k-th one (starts at 0xffffffff)
  SelCell* st;                // one cell
  unsigned* hist;             // 3 x SEL_BINS: one histogram per radix pass
  unsigned long long* counters;
};

// normalize_point_cloud_to_ground (:880-975) in one cooperative launch
__global__ void __launch_bounds__(T) ground_normalize_kernel(const NormArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ unsigned sh[SEL_BINS];
  __shared__ unsigned part[T + 2];
  __shared__ unsigned s_pop;                 // this block's population count (kept across the phases)
  const long long stride = static_cast<long long>(gridDim.x) * T;
  const GroundXf& xf = a.xf;
  unsigned* seg_keys = a.keys + blockIdx.x * a.seg;
  if (threadIdx.x == 0) s_pop = 0;
  __syncthreads();
  // ---- phase 0: distances to the plane, rotation, plane to y = const (:900-945); population -> compact keys
  for (long long i0 = blockIdx.x * static_cast<long long>(T); i0 < a.n; i0 += stride) {
    const long long i = i0 + threadIdx.x;
    bool in = false;
    unsigned key = 0;
    if (i < a.n) {
      const double x = a.xyz[3 * i], y = a.xyz[3 * i + 1], z = a.xyz[3 * i + 2];
      // np.dot(points, normal) + d
      const double dist = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, xf.n[0]), __dmul_rn(y, xf.n[1])), __dmul_rn(z, xf.n[2])), xf.d);
      const double rx = __dadd_rn(__dadd_rn(__dmul_rn(xf.r[0], x), __dmul_rn(xf.r[1], y)), __dmul_rn(xf.r[2], z));
      const double ry = __dadd_rn(__dadd_rn(__dmul_rn(xf.r[3], x), __dmul_rn(xf.r[4], y)), __dmul_rn(xf.r[5], z));
      const double rz = __dadd_rn(__dadd_rn(__dmul_rn(xf.r[6], x), __dmul_rn(xf.r[7], y)), __dmul_rn(xf.r[8], z));
      const float yn = static_cast<float>(__dadd_rn(ry, -xf.shift));
      a.xyz[3 * i] = static_cast<float>(rx);
      a.xyz[3 * i + 1] = yn;
      a.xyz[3 * i + 2] = static_cast<float>(rz);
      in = fabs(dist) < 0.1;
      a.flags[i] = static_cast<uint8_t>((in ? 1 : 0) | (fabs(dist) < 0.05 ? 2 : 0));
      key = f2key(yn);
    }
    pop_append(in, key, 0, &s_pop, seg_keys, nullptr);
  }
  __syncthreads();
  const unsigned mine = s_pop;
  if (threadIdx.x == 0 && mine) atomicAdd(a.pop_count, mine);
  grid.sync();
  // ---- phase 1: exact 2nd percentile of the population by radix select.  Every block keeps ITS OWN copy of the (single)
  // cell's select state in shared memory and scans the shared histogram of a pass itself (one histogram per pass, never
  // cleared): one grid barrier per pass instead of two.
  const unsigned M = __ldcg(a.pop_count);
  __shared__ SelCell cs;
  if (threadIdx.x == 0) {
    cs = SelCell();
    cs.count_all = M, cs.count = M;
    sel_begin(cs, 0.02, 0, 1);
  }
  __syncthreads();
  const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
  unsigned fixed = 0;
  if (M > 0) {
    for (int p = 0; p < 3; ++p) {
      for (int b = threadIdx.x; b < SEL_BINS; b += T) sh[b] = 0;
      __syncthreads();
      const unsigned prefix = cs.prefix, dmask = (1u << bits[p]) - 1u;
      for (unsigned i = threadIdx.x; i < mine; i += T) {
        const unsigned key = seg_keys[i];
        if ((key & fixed) == prefix) atomicAdd(&sh[(key >> shifts[p]) & dmask], 1u);
      }
      __syncthreads();
      unsigned* hist = a.hist + p * SEL_BINS;
      for (int b = threadIdx.x; b < SEL_BINS; b += T)
        if (sh[b]) atomicAdd(&hist[b], sh[b]);
      grid.sync();
      sel_scan_block<true>(&cs, hist, shifts[p], bits[p], p == 2, part);
      __syncthreads();
      fixed |= dmask << shifts[p];
    }
    // smallest population key greater than the k-th key (needed when the (k+1)-th statistic is not a duplicate)
    const unsigned kth = cs.prefix;
    unsigned nx = 0xffffffffu;
    for (unsigned i = threadIdx.x; i < mine; i += T) {
      const unsigned key = seg_keys[i];
      if (key > kth && key < nx) nx = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nx = min(nx, __shfl_xor_sync(0xffffffffu, nx, o));
    if ((threadIdx.x & 31) == 0 && nx != 0xffffffffu) atomicMin(a.next_key, nx);
    grid.sync();
    if (threadIdx.x == 0) {
      cs.next = __ldcg(a.next_key);
      sel_finish<true>(cs);
      if (blockIdx.x == 0) *a.st = cs;   // for the record (tests / debugging read the state back)
    }
    __syncthreads();
  }
  // ---- phase 2: ground level shift + clamps (:947-972)
  const bool apply = M > 10;                       // "if len(ground_y_values) > 10"
  const double value = cs.value;
  unsigned n_ground = 0, n_zero = 0, n_floor = 0;  // the reference prints these three counts; kept for the caller's log
  for (long long i0 = blockIdx.x * static_cast<long long>(T); i0 < a.n; i0 += stride) {
    const long long i = i0 + threadIdx.x;
    bool ground = false, to_zero = false, to_floor = false;
    if (i < a.n) {
      double y = a.xyz[3 * i + 1];
      if (apply) y = __dadd_rn(y, -value);
      ground = (a.flags[i] & 2) != 0;
      to_zero = y < 0.0 && ground;
      if (to_zero) y = 0.0;
      to_floor = y < -0.1 && !ground;
      if (to_floor) y = -0.1;
      a.xyz[3 * i + 1] = static_cast<float>(y);
    }
    n_ground += ground, n_zero += to_zero, n_floor += to_floor;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_ground += __shfl_xor_sync(0xffffffffu, n_ground, o);
    n_zero += __shfl_xor_sync(0xffffffffu, n_zero, o);
    n_floor += __shfl_xor_sync(0xffffffffu, n_floor, o);
  }
  if ((threadIdx.x & 31) == 0) {   // one atomic per warp and counter for the whole kernel
    if (n_ground) atomicAdd(&a.counters[0], static_cast<unsigned long long>(n_ground));
    if (n_zero) atomicAdd(&a.counters[1], static_cast<unsigned long long>(n_zero));
    if (n_floor) atomicAdd(&a.counters[2], static_cast<unsigned long long>(n_floor));
  }
}

// ---- grid adjustment ------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long dkey(double v) {  // order-preserving key of a double
  const unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(v));
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key2d(unsigned long long k) {
  return __longlong_as_double(static_cast<long long>((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}

// np.digitize(v, edges) - 1 clipped to [0, g - 1]; digitize = number of edges <= v (edges increasing)
__device__ __forceinline__ int bin_of(double v, const double* __restrict__ e, int g) {
  int lo = 0, hi = g + 1;  // first index with e[idx] > v
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (e[mid] <= v) lo = mid + 1;
    else hi = mid;
  }
  int bin = lo - 1;
  bin = bin < 0 ? 0 : bin;
  return bin > g - 1 ? g - 1 : bin;
}

constexpr int PRIV_CELLS = 2048;   // grids up to 45 x 45 keep their per-cell counters in shared memory (16 KB)
struct GridArgs {
  float* xyz;
  long long n;
  int g;                      // grid_size
  double q;                   // percentile / 100
  unsigned short* cells;      // cell of every point
  unsigned* keys;             // compact population (heights y < 0.2): keys ...
  unsigned short* kcell;      // ... and their cells, one segment per block
  long long seg;
  unsigned* pop_count;
  SelCell* st;                // g * g cells
  unsigned* hist;             // g * g * SEL_BINS
  unsigned long long* bounds; // x min, z min, x max, z max (order-preserving keys of doubles)
  double* edges;              // 2 * (g + 1)
  unsigned long long* counters;
};

// grid_based_ground_adjustment (:977-1118) in one cooperative launch
__global__ void __launch_bounds__(T) ground_grid_kernel(const GridArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double se[];            // 2 * (g + 1) edges, then (small grids) 2 * ncells per-block cell counters
  __shared__ unsigned part[T + 2];
  __shared__ unsigned s_pop;
  __shared__ unsigned long long s_box[4];
  const int g = a.g, ncells = g * g;
  const long long stride = static_cast<long long>(gridDim.x) * T;
  const long long first = blockIdx.x * static_cast<long long>(T) + threadIdx.x;
  unsigned* seg_keys = a.keys + blockIdx.x * a.seg;
  unsigned short* seg_cell = a.kcell + blockIdx.x * a.seg;
  const bool priv = ncells <= PRIV_CELLS;   // per-block counters in shared memory instead of one global atomic per point
  unsigned* s_cnt = reinterpret_cast<unsigned*>(se + 2 * (g + 1));
  if (threadIdx.x == 0) s_pop = 0, s_box[0] = s_box[1] = 0xffffffffffffffffull, s_box[2] = s_box[3] = 0ull;
  if (priv)
    for (int c = threadIdx.x; c < 2 * ncells; c += T) s_cnt[c] = 0;
  __syncthreads();
  // ---- phase 0: XZ bounding box (:1009-1018)
  {
    unsigned long long lo_x = 0xffffffffffffffffull, hi_x = 0, lo_z = lo_x, hi_z = 0;
    for (long long i = first; i < a.n; i += stride) {
      const unsigned long long kx = dkey(static_cast<double>(a.xyz[3 * i])), kz = dkey(static_cast<double>(a.xyz[3 * i + 2]));
      lo_x = kx < lo_x ? kx : lo_x, hi_x = kx > hi_x ? kx : hi_x;
      lo_z = kz < lo_z ? kz : lo_z, hi_z = kz > hi_z ? kz : hi_z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long t;
      t = __shfl_xor_sync(0xffffffffu, lo_x, o), lo_x = t < lo_x ? t : lo_x;
      t = __shfl_xor_sync(0xffffffffu, hi_x, o), hi_x = t > hi_x ? t : hi_x;
      t = __shfl_xor_sync(0xffffffffu, lo_z, o), lo_z = t < lo_z ? t : lo_z;
      t = __shfl_xor_sync(0xffffffffu, hi_z, o), hi_z = t > hi_z ? t : hi_z;
    }
    if ((threadIdx.x & 31) == 0)
      atomicMin(&s_box[0], lo_x), atomicMin(&s_box[1], lo_z), atomicMax(&s_box[2], hi_x), atomicMax(&s_box[3], hi_z);
    __syncthreads();
    if (threadIdx.x == 0)   // one global atomic per block and bound
      atomicMin(&a.bounds[0], s_box[0]), atomicMin(&a.bounds[1], s_box[1]), atomicMax(&a.bounds[2], s_box[2]),
          atomicMax(&a.bounds[3], s_box[3]);
  }
  grid.sync();
  // ---- phase 1: np.linspace edges (every block its own copy), cell of every point, per-cell counts, population
  for (int i = threadIdx.x; i <= g; i += T) {
    for (int ax = 0; ax < 2; ++ax) {
      const double lo = key2d(__ldcg(&a.bounds[ax])), hi = key2d(__ldcg(&a.bounds[2 + ax]));
      const double step = __ddiv_rn(__dadd_rn(hi, -lo), static_cast<double>(g));
      const double e = i == g ? hi : __dadd_rn(__dmul_rn(static_cast<double>(i), step), lo);
      se[ax * (g + 1) + i] = e;
      if (blockIdx.x == 0) a.edges[ax * (g + 1) + i] = e;
    }
  }
  __syncthreads();
  for (long long i0 = blockIdx.x * static_cast<long long>(T); i0 < a.n; i0 += stride) {
    const long long i = i0 + threadIdx.x;
    bool in = false;
    unsigned key = 0;
    unsigned short c = 0;
    if (i < a.n) {
      const int bx = bin_of(static_cast<double>(a.xyz[3 * i]), se, g), bz = bin_of(static_cast<double>(a.xyz[3 * i + 2]), se + g + 1, g);
      c = static_cast<unsigned short>(bx * g + bz);
      a.cells[i] = c;
      const float y = a.xyz[3 * i + 1];
      in = static_cast<double>(y) < 0.2;                              // the cell's low points (:1058)
      key = f2key(y);
      if (priv) {
        atomicAdd(&s_cnt[2 * c], 1u);
        if (in) atomicAdd(&s_cnt[2 * c + 1], 1u);
      } else {
        atomicAdd(&a.st[c].count_all, 1u);
        if (in) atomicAdd(&a.st[c].count, 1u);
      }
    }
    pop_append(in, key, c, &s_pop, seg_keys, seg_cell);
  }
  __syncthreads();
  if (priv)
    for (int c = threadIdx.x; c < ncells; c += T) {
      if (s_cnt[2 * c]) atomicAdd(&a.st[c].count_all, s_cnt[2 * c]);
      if (s_cnt[2 * c + 1]) atomicAdd(&a.st[c].count, s_cnt[2 * c + 1]);
    }
  const unsigned mine = s_pop;
  grid.sync();
  // cells with >= 10 points and >= 5 low points (:1046, :1061); np.percentile(lowest_y, percentile)
  for (long long c = first; c < ncells; c += stride) sel_begin(a.st[c], a.q, 10, 5);
  grid.sync();
  // Small grids keep a per-block copy of the cells' select state in shared memory (the 8 bytes per cell the counters
  // used), refreshed after every grid barrier: per-element __ldcg reads of a few hundred hot lines were the bottleneck.
  unsigned* s_pref = s_cnt;                                             // [ncells] key bits fixed so far
  uint8_t* s_act = reinterpret_cast<uint8_t*>(s_cnt + ncells);          // [ncells] cell takes part in the selection
  unsigned* s_next = s_cnt + ncells;                                    // [ncells] (next-key pass) block-local minimum
  double* s_val = reinterpret_cast<double*>(s_cnt);                     // [ncells] (apply pass) amount to lower by, 0 = none
  auto load_state = [&]() {
    __syncthreads();
    if (priv)
      for (int c = threadIdx.x; c < ncells; c += T) s_pref[c] = __ldcg(&a.st[c].prefix), s_act[c] = __ldcg(&a.st[c].active) != 0;
    __syncthreads();
  };
  const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
  unsigned fixed = 0;
  for (int p = 0; p < 3; ++p) {
    const unsigned dmask = (1u << bits[p]) - 1u;
    load_state();
    // consecutive points of an unprojected frame are image neighbours: same cell, similar height, so most lanes of a
    // warp hit the SAME (cell, digit) bin -- lanes with equal bins elect one to add their count (same-address global
    // atomics serialise in L2; one per lane was what this pass spent its time on)
    for (unsigned base = threadIdx.x & ~31u; base < mine; base += T) {
      const unsigned i = base + (threadIdx.x & 31);
      unsigned tag = 0xffffffffu;
      if (i < mine) {
        const unsigned key = seg_keys[i];
        const unsigned c = seg_cell[i];
        const bool act = priv ? s_act[c] != 0 : __ldcg(&a.st[c].active) != 0;
        const unsigned pref = priv ? s_pref[c] : __ldcg(&a.st[c].prefix);
        if (act && (key & fixed) == pref) tag = c * SEL_BINS + ((key >> shifts[p]) & dmask);
      }
      const unsigned same = __match_any_sync(0xffffffffu, tag);
      if (tag != 0xffffffffu && (threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&a.hist[tag], static_cast<unsigned>(__popc(same)));
    }
    grid.sync();
    for (int c = blockIdx.x; c < ncells; c += gridDim.x)
      sel_scan_block(&a.st[c], a.hist + static_cast<size_t>(c) * SEL_BINS, shifts[p], bits[p], p == 2, part);
    grid.sync();
    fixed |= dmask << shifts[p];
  }
  // smallest population key greater than the k-th key, per cell
  __syncthreads();
  if (priv)
    for (int c = threadIdx.x; c < ncells; c += T)
      s_pref[c] = __ldcg(&a.st[c].active) ? __ldcg(&a.st[c].prefix) : 0xffffffffu, s_next[c] = 0xffffffffu;   // inactive: nothing is greater
  __syncthreads();
  for (unsigned i = threadIdx.x; i < mine; i += T) {
    const unsigned key = seg_keys[i];
    const int c = seg_cell[i];
    if (priv) {
      if (key > s_pref[c] && key < s_next[c]) atomicMin(&s_next[c], key);
    } else {
      if (!__ldcg(&a.st[c].active)) continue;
      if (key > __ldcg(&a.st[c].prefix) && key < __ldcg(&a.st[c].next)) atomicMin(&a.st[c].next, key);
    }
  }
  __syncthreads();
  if (priv)
    for (int c = threadIdx.x; c < ncells; c += T)
      if (s_next[c] != 0xffffffffu) atomicMin(&a.st[c].next, s_next[c]);
  grid.sync();
  for (long long c = first; c < ncells; c += stride) {
    sel_finish(a.st[c]);
    if (__ldcg(&a.st[c].count_all) >= 10) atomicAdd(&a.counters[4], 1ull);                    // cells with sufficient points
    if (__ldcg(&a.st[c].active) && a.st[c].value > 0.01) atomicAdd(&a.counters[5], 1ull);    // cells requiring adjustment
  }
  grid.sync();
  // ---- per-cell lowering with the height-graded factor, clamp at y = 0   (:1067-1106)
  __syncthreads();
  if (priv)
    for (int c = threadIdx.x; c < ncells; c += T) s_val[c] = __ldcg(&a.st[c].active) ? __ldcg(&a.st[c].value) : 0.0;
  __syncthreads();
  unsigned lowered = 0;
  for (long long i = first; i < a.n; i += stride) {
    const int c = a.cells[i];
    double p;
    if (priv) {
      p = s_val[c];
    } else {
      if (!__ldcg(&a.st[c].active)) continue;
      p = __ldcg(&a.st[c].value);
    }
    if (!(p > 0.01)) continue;
    const double y = a.xyz[3 * i + 1];
    double adj = 0.0;
    if (y < 0.1) adj = p;
    else if (y < 1.5) adj = __dmul_rn(p, __dadd_rn(1.0, -__ddiv_rn(__dadd_rn(y, -0.1), 1.4)));
    double out = __dadd_rn(y, -adj);
    if (out < 0.0) out = 0.0;
    a.xyz[3 * i + 1] = static_cast<float>(out);
    if (adj > 0.0) ++lowered;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lowered += __shfl_xor_sync(0xffffffffu, lowered, o);
  if ((threadIdx.x & 31) == 0 && lowered) atomicAdd(&a.counters[3], static_cast<unsigned long long>(lowered));
}

// persistent grid of a cooperative launch: every block must be resident
template <typename K>
int coop_blocks(K kernel, size_t smem, long long n) {
  int dev = 0, sms = 0, per_sm = 0;
  DP_CUDA(cudaGetDevice(&dev));
  DP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, T, smem));
  DP_CHECK(per_sm >= 1, "ground kernels: no resident block fits");
  long long want = (n + T - 1) / T;
  const long long cap = static_cast<long long>(sms) * (per_sm > 4 ? 4 : per_sm);
  want = want < 1 ? 1 : (want > cap ? cap : want);
  return static_cast<int>(want);
}

}  // namespace

namespace {
inline size_t up256(size_t v) { return (v + 255) & ~size_t(255); }
// per-block population segments: blocks * ceil(n / (blocks * T)) * T <= n + blocks * T entries, blocks <= 4 per SM
constexpr size_t SEG_SLACK = 256 * 4 * T;
struct Scratch {
  uint8_t* flags;             // normalise: one flag byte per point  |  grid: uint16 cell per point (same storage)
  unsigned short* cells;
  unsigned* keys;             // compact population keys (at most n)
  unsigned short* kcell;      // their cells (grid adjustment)
  SelCell* st;
  unsigned* hist;
  unsigned* pop_count;        // misc block: +0 population counter, +64 bounds (4 x u64), +128 spare counters (6 x u64)
  unsigned long long* bounds;
  unsigned long long* spare;
  double* edges;
};
Scratch carve(void* base, long long n, int grid_size) {
  const size_t cells = static_cast<size_t>(grid_size) * grid_size;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  Scratch sc;
  sc.flags = p, sc.cells = reinterpret_cast<unsigned short*>(p), p += up256(static_cast<size_t>(n) * 2);
  sc.keys = reinterpret_cast<unsigned*>(p), p += up256((static_cast<size_t>(n) + SEG_SLACK) * 4);
  sc.kcell = reinterpret_cast<unsigned short*>(p), p += up256((static_cast<size_t>(n) + SEG_SLACK) * 2);
  sc.st = reinterpret_cast<SelCell*>(p), p += up256(cells * sizeof(SelCell));
  sc.hist = reinterpret_cast<unsigned*>(p), p += (cells < 3 ? 3 : cells) * SEL_BINS * 4;
  sc.pop_count = reinterpret_cast<unsigned*>(p);
  sc.bounds = reinterpret_cast<unsigned long long*>(p + 64);
  sc.spare = reinterpret_cast<unsigned long long*>(p + 128);
  p += 256;
  sc.edges = reinterpret_cast<double*>(p);
  return sc;
}
}  // namespace

size_t ground_scratch_bytes(long long n, int grid_size) {
  const size_t cells = static_cast<size_t>(grid_size) * grid_size;
  return up256(static_cast<size_t>(n) * 2) + up256((static_cast<size_t>(n) + SEG_SLACK) * 4) +
         up256((static_cast<size_t>(n) + SEG_SLACK) * 2) +
         up256(cells * sizeof(SelCell)) + (cells < 3 ? 3 : cells) * SEL_BINS * 4 + 256 + up256(2 * (static_cast<size_t>(grid_size) + 1) * 8) + 256;
}

// counters (device, 6 x uint64): [0] ground points (|dist| < 0.05), [1] set to y = 0, [2] limited to -0.1,
// [3] points lowered by the grid pass, [4] cells with >= 10 points, [5] cells adjusted
void ground_normalize(float* xyz, long long n, const double normal[3], double d, void* scratch, unsigned long long* counters,
                      cudaStream_t s) {
  if (n <= 0) return;
  Scratch sc = carve(scratch, n, 1);
  // host side of :900-941 -- 3 x 3 algebra on the plane parameters, in double like numpy
  GroundXf xf;
  const double nn = std::sqrt(normal[0] * normal[0] + normal[1] * normal[1] + normal[2] * normal[2]);
  DP_CHECK(nn > 0.0, "ground_normalize: zero normal");
  for (int i = 0; i < 3; ++i) xf.n[i] = normal[i] / nn;   // point_plane_distances normalises (:873)
  xf.d = d;
  for (int i = 0; i < 9; ++i) xf.r[i] = (i % 4 == 0) ? 1.0 : 0.0;
  xf.shift = 0.0;
  // "if np.abs(np.dot(from_vec, to_vec)) > 0.99" uses the normal AS GIVEN (:912)
  if (!(std::fabs(normal[1]) > 0.99)) {
    const double f[3] = {normal[0] / nn, normal[1] / nn, normal[2] / nn};
    double ax[3] = {f[1] * 0.0 - f[2] * 1.0, f[2] * 0.0 - f[0] * 0.0, f[0] * 1.0 - f[1] * 0.0};  // cross(f, [0,1,0])
    const double an = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    for (double& a : ax) a /= an;
    double c = f[1];
    c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
    const double ang = std::acos(c), sn = std::sin(ang), oc = 1.0 - std::cos(ang);
    const double K[9] = {0, -ax[2], ax[1], ax[2], 0, -ax[0], -ax[1], ax[0], 0};
    double K2[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) K2[3 * i + j] = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
    for (int i = 0; i < 9; ++i) xf.r[i] = ((i % 4 == 0) ? 1.0 : 0.0) + sn * K[i] + oc * K2[i];
    const double rn_y = xf.r[3] * normal[0] + xf.r[4] * normal[1] + xf.r[5] * normal[2];  // (R @ normal)[1], raw normal
    xf.shift = -d / rn_y;
  }
  DP_CHECK(n < (1ll << 32), "ground_normalize: more than 2^32 points");
  // scratch tail: [3 histograms | misc block]; one memset clears both, a second sets the next-key word (misc + 8) to ~0
  DP_CUDA(cudaMemsetAsync(sc.hist, 0, 3 * SEL_BINS * 4 + 256, s));
  DP_CUDA(cudaMemsetAsync(sc.pop_count + 2, 0xff, 4, s));
  if (counters) DP_CUDA(cudaMemsetAsync(counters, 0, 3 * 8, s));
  NormArgs a;
  a.xyz = xyz, a.n = n, a.xf = xf, a.flags = sc.flags, a.keys = sc.keys, a.pop_count = sc.pop_count, a.st = sc.st;
  a.next_key = sc.pop_count + 2;
  a.hist = sc.hist, a.counters = counters ? counters : sc.spare;
  const int blocks = coop_blocks(ground_normalize_kernel, 0, n);
  a.seg = (n + static_cast<long long>(blocks) * T - 1) / (static_cast<long long>(blocks) * T) * T;
  DP_CHECK(static_cast<size_t>(blocks) * a.seg <= static_cast<size_t>(n) + SEG_SLACK, "ground_normalize: segment plan");
  void* args[] = {&a};
  DP_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(ground_normalize_kernel), dim3(blocks), dim3(T), args, 0, s));
  DP_LAUNCH_CHECK();
}

void ground_grid_adjust(float* xyz, long long n, int grid_size, double percentile, void* scratch, unsigned long long* counters,
                        cudaStream_t s) {
  if (n <= 0) return;
  DP_CHECK(grid_size >= 1 && grid_size <= 255, "grid_size must be in 1..255");
  DP_CHECK(n < (1ll << 32), "ground_grid_adjust: more than 2^32 points");
  const int ncells = grid_size * grid_size;
  Scratch sc = carve(scratch, n, grid_size);
  DP_CUDA(cudaMemsetAsync(sc.st, 0, ncells * sizeof(SelCell), s));
  DP_CUDA(cudaMemsetAsync(sc.hist, 0, static_cast<size_t>(ncells) * SEL_BINS * 4, s));
  DP_CUDA(cudaMemsetAsync(sc.pop_count, 0, 256, s));
  DP_CUDA(cudaMemsetAsync(sc.bounds, 0xff, 16, s));   // x min, z min start at the largest key; x max, z max at 0
  if (counters) DP_CUDA(cudaMemsetAsync(counters + 3, 0, 3 * 8, s));
  GridArgs a;
  a.xyz = xyz, a.n = n, a.g = grid_size, a.q = percentile / 100.0, a.cells = sc.cells, a.keys = sc.keys, a.kcell = sc.kcell;
  a.pop_count = sc.pop_count, a.st = sc.st, a.hist = sc.hist, a.bounds = sc.bounds, a.edges = sc.edges;
  a.counters = counters ? counters : sc.spare;
  const size_t smem = 2 * (static_cast<size_t>(grid_size) + 1) * sizeof(double) + (ncells <= PRIV_CELLS ? 2 * ncells * 4 : 0);
  const int blocks = coop_blocks(ground_grid_kernel, smem, n);
  a.seg = (n + static_cast<long long>(blocks) * T - 1) / (static_cast<long long>(blocks) * T) * T;
  DP_CHECK(static_cast<size_t>(blocks) * a.seg <= static_cast<size_t>(n) + SEG_SLACK, "ground_grid_adjust: segment plan");
  void* args[] = {&a};
  DP_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(ground_grid_kernel), dim3(blocks), dim3(T), args, smem, s));
  DP_LAUNCH_CHECK();
}

}  // namespace dp
