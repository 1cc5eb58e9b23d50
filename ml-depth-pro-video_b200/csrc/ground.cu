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
// (lib/_function_base_impl.py _lerp) finishes in double.
#include "common.cuh"
#include "kernels.cuh"

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

// MODE 0: one cell; population = heights y of the points with |dist| < 0.1     (:947)
// MODE 1: cell = cells[i]; population = heights y < 0.2 of each cell            (:1058)
template <int MODE>
__device__ __forceinline__ bool in_population(const float* __restrict__ xyz, const double* __restrict__ dist,
                                              const unsigned short* __restrict__ cells, long long i, int& cell, float& y) {
  y = xyz[3 * i + 1];
  if (MODE == 0) {
    cell = 0;
    return fabs(dist[i]) < 0.1;
  }
  cell = cells[i];
  return static_cast<double>(y) < 0.2;
}

template <int MODE>
__global__ void __launch_bounds__(T) sel_count_kernel(const float* __restrict__ xyz, const double* __restrict__ dist,
                                                      const unsigned short* __restrict__ cells, long long n, SelCell* st) {
  for (long long i = blockIdx.x * static_cast<long long>(T) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * T) {
    int c;
    float y;
    const bool in = in_population<MODE>(xyz, dist, cells, i, c, y);
    if (MODE == 1) {
      atomicAdd(&st[c].count_all, 1u);
      if (in) atomicAdd(&st[c].count, 1u);
    } else {
      // one cell: warp-aggregate before touching the single counter
      const unsigned b = __ballot_sync(__activemask(), in);
      if (in && (threadIdx.x & 31) == __ffs(b) - 1) atomicAdd(&st[0].count, static_cast<unsigned>(__popc(b)));
    }
  }
}

// numpy percentile, method 'linear': virtual index = n q + (1 - q) - 1, k = floor, gamma = fraction
__global__ void sel_begin_kernel(SelCell* st, int ncells, double q, unsigned min_all, unsigned min_pop) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  SelCell& s = st[c];
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

// histogram of the next digit over the elements whose fixed bits match the cell's prefix
template <int MODE>
__global__ void __launch_bounds__(T) sel_hist_kernel(const float* __restrict__ xyz, const double* __restrict__ dist,
                                                     const unsigned short* __restrict__ cells, long long n,
                                                     const SelCell* __restrict__ st, unsigned* __restrict__ hist, int shift,
                                                     int bits, unsigned fixed_mask) {
  __shared__ unsigned sh[MODE == 0 ? SEL_BINS : 1];
  if (MODE == 0) {
    for (int b = threadIdx.x; b < SEL_BINS; b += T) sh[b] = 0;
    __syncthreads();
  }
  const unsigned dmask = (1u << bits) - 1u;
  for (long long i = blockIdx.x * static_cast<long long>(T) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * T) {
    int c;
    float y;
    if (!in_population<MODE>(xyz, dist, cells, i, c, y)) continue;
    if (!st[c].active) continue;
    const unsigned key = f2key(y);
    if ((key & fixed_mask) != st[c].prefix) continue;
    const unsigned dgt = (key >> shift) & dmask;
    if (MODE == 0) atomicAdd(&sh[dgt], 1u);
    else atomicAdd(&hist[static_cast<size_t>(c) * SEL_BINS + dgt], 1u);
  }
  if (MODE == 0) {
    __syncthreads();
    for (int b = threadIdx.x; b < SEL_BINS; b += T)
      if (sh[b]) atomicAdd(&hist[b], sh[b]);
  }
}

// one block per cell: find the digit whose cumulative count crosses the residual rank, clear the histogram
__global__ void __launch_bounds__(T) sel_scan_kernel(SelCell* st, unsigned* hist, int shift, int bits, int last) {
  const int c = blockIdx.x;
  SelCell& s = st[c];
  unsigned* h = hist + static_cast<size_t>(c) * SEL_BINS;
  __shared__ unsigned part[T];
  const int nb = 1 << bits, per = SEL_BINS / T;  // 8 consecutive bins per thread
  unsigned loc[SEL_BINS / T];
  unsigned sum = 0;
#pragma unroll
  for (int j = 0; j < per; ++j) {
    const int b = threadIdx.x * per + j;
    loc[j] = b < nb ? h[b] : 0u;
    sum += loc[j];
    h[b] = 0;
  }
  part[threadIdx.x] = sum;
  __syncthreads();
  if (!s.active) return;
  if (threadIdx.x == 0) {
    unsigned acc = 0;
    int t = 0;
    for (; t < T; ++t) {
      if (acc + part[t] > s.rank) break;
      acc += part[t];
    }
    part[0] = static_cast<unsigned>(t < T ? t : T - 1);  // thread that owns the crossing
    part[1] = acc;
  }
  __syncthreads();
  if (threadIdx.x == static_cast<int>(part[0])) {
    unsigned acc = part[1];
    int j = 0;
    for (; j < per - 1; ++j) {
      if (acc + loc[j] > s.rank) break;
      acc += loc[j];
    }
    const unsigned dgt = static_cast<unsigned>(threadIdx.x * per + j);
    s.prefix |= dgt << shift;
    s.less += acc;
    s.rank -= acc;
    if (last) s.eq = loc[j];
  }
}

// smallest population key greater than the k-th key (needed when the (k+1)-th statistic is not a duplicate)
template <int MODE>
__global__ void __launch_bounds__(T) sel_next_kernel(const float* __restrict__ xyz, const double* __restrict__ dist,
                                                     const unsigned short* __restrict__ cells, long long n, SelCell* st) {
  for (long long i = blockIdx.x * static_cast<long long>(T) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * T) {
    int c;
    float y;
    if (!in_population<MODE>(xyz, dist, cells, i, c, y)) continue;
    if (!st[c].active) continue;
    const unsigned key = f2key(y);
    if (key > st[c].prefix && key < st[c].next) atomicMin(&st[c].next, key);
  }
}

__global__ void sel_finish_kernel(SelCell* st, int ncells) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  SelCell& s = st[c];
  if (!s.active) return;
  const double a = static_cast<double>(key2f(s.prefix));
  double b = a;
  // ranks less .. less + eq - 1 hold the k-th key; the (k+1)-th is a duplicate unless k is the last of them
  if (s.k + 1 >= s.less + s.eq && s.k + 1 < s.count && s.next != 0xffffffffu) b = static_cast<double>(key2f(s.next));
  // numpy _lerp: a + (b - a) t, evaluated from the other end for t >= 0.5
  const double diff = __dadd_rn(b, -a), t = s.gamma;
  double r = __dadd_rn(a, __dmul_rn(diff, t));
  if (t >= 0.5) r = __dadd_rn(b, -__dmul_rn(diff, __dadd_rn(1.0, -t)));
  if (t == 0.0) r = a;  // numpy's gamma == 0 short cut keeps the lower statistic exactly
  s.value = r;
}

struct GroundXf {
  double r[9];    // rotation (identity when the plane normal is within acos(0.99) of +-y)
  double n[3];    // unit normal
  double d;
  double shift;   // y offset of the rotated plane: -d / (R normal).y, 0 without rotation
};

// distances to the plane, rotation, plane to y = const   (:900-945)
__global__ void __launch_bounds__(T) ground_transform_kernel(float* __restrict__ xyz, double* __restrict__ dist, long long n,
                                                             const GroundXf xf) {
  const long long i = blockIdx.x * static_cast<long long>(T) + threadIdx.x;
  if (i >= n) return;
  const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
  // np.dot(points, normal) + d
  dist[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, xf.n[0]), __dmul_rn(y, xf.n[1])), __dmul_rn(z, xf.n[2])), xf.d);
  const double rx = __dadd_rn(__dadd_rn(__dmul_rn(xf.r[0], x), __dmul_rn(xf.r[1], y)), __dmul_rn(xf.r[2], z));
  const double ry = __dadd_rn(__dadd_rn(__dmul_rn(xf.r[3], x), __dmul_rn(xf.r[4], y)), __dmul_rn(xf.r[5], z));
  const double rz = __dadd_rn(__dadd_rn(__dmul_rn(xf.r[6], x), __dmul_rn(xf.r[7], y)), __dmul_rn(xf.r[8], z));
  xyz[3 * i] = static_cast<float>(rx);
  xyz[3 * i + 1] = static_cast<float>(__dadd_rn(ry, -xf.shift));
  xyz[3 * i + 2] = static_cast<float>(rz);
}

// ground level shift + clamps   (:947-972)
__global__ void __launch_bounds__(T) ground_clamp_kernel(float* __restrict__ xyz, const double* __restrict__ dist, long long n,
                                                         const SelCell* __restrict__ st, unsigned long long* __restrict__ counters) {
  const long long i = blockIdx.x * static_cast<long long>(T) + threadIdx.x;
  if (i >= n) return;
  double y = xyz[3 * i + 1];
  if (st[0].count > 10) y = __dadd_rn(y, -st[0].value);   // "if len(ground_y_values) > 10"
  const bool ground = fabs(dist[i]) < 0.05;
  const bool to_zero = y < 0.0 && ground;
  if (to_zero) y = 0.0;
  const bool to_floor = y < -0.1 && !ground;
  if (to_floor) y = -0.1;
  xyz[3 * i + 1] = static_cast<float>(y);
  // the reference prints these three counts; kept for the caller's log
  const unsigned bg = __ballot_sync(__activemask(), ground), bz = __ballot_sync(__activemask(), to_zero),
                 bf = __ballot_sync(__activemask(), to_floor);
  if ((threadIdx.x & 31) == 0) {
    if (bg) atomicAdd(&counters[0], static_cast<unsigned long long>(__popc(bg)));
    if (bz) atomicAdd(&counters[1], static_cast<unsigned long long>(__popc(bz)));
    if (bf) atomicAdd(&counters[2], static_cast<unsigned long long>(__popc(bf)));
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

__global__ void bounds_init_kernel(unsigned long long* b) {
  b[0] = b[2] = 0xffffffffffffffffull;  // x min, z min
  b[1] = b[3] = 0ull;                   // x max, z max
}
__global__ void __launch_bounds__(T) bounds_kernel(const float* __restrict__ xyz, long long n, unsigned long long* b) {
  unsigned long long lo_x = 0xffffffffffffffffull, hi_x = 0, lo_z = lo_x, hi_z = 0;
  for (long long i = blockIdx.x * static_cast<long long>(T) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * T) {
    const unsigned long long kx = dkey(static_cast<double>(xyz[3 * i])), kz = dkey(static_cast<double>(xyz[3 * i + 2]));
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
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&b[0], lo_x), atomicMax(&b[1], hi_x), atomicMin(&b[2], lo_z), atomicMax(&b[3], hi_z);
  }
}
// np.linspace(lo, hi, g + 1): step = (hi - lo) / g, edge[i] = i * step + lo, last edge = hi exactly
__global__ void edges_kernel(const unsigned long long* __restrict__ b, double* __restrict__ edges, int g) {
  const int i = threadIdx.x;
  if (i > g) return;
  for (int a = 0; a < 2; ++a) {
    const double lo = key2d(b[2 * a]), hi = key2d(b[2 * a + 1]);
    const double step = __ddiv_rn(__dadd_rn(hi, -lo), static_cast<double>(g));
    edges[a * (g + 1) + i] = i == g ? hi : __dadd_rn(__dmul_rn(static_cast<double>(i), step), lo);
  }
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
__global__ void __launch_bounds__(T) cells_kernel(const float* __restrict__ xyz, long long n, const double* __restrict__ edges,
                                                  int g, unsigned short* __restrict__ cells) {
  extern __shared__ double se[];
  for (int i = threadIdx.x; i < 2 * (g + 1); i += T) se[i] = edges[i];
  __syncthreads();
  const long long i = blockIdx.x * static_cast<long long>(T) + threadIdx.x;
  if (i >= n) return;
  const int bx = bin_of(static_cast<double>(xyz[3 * i]), se, g), bz = bin_of(static_cast<double>(xyz[3 * i + 2]), se + g + 1, g);
  cells[i] = static_cast<unsigned short>(bx * g + bz);
}
// per-cell lowering with the height-graded factor, clamp at y = 0   (:1067-1106)
__global__ void __launch_bounds__(T) grid_apply_kernel(float* __restrict__ xyz, long long n,
                                                       const unsigned short* __restrict__ cells,
                                                       const SelCell* __restrict__ st, unsigned long long* __restrict__ counters) {
  const long long i = blockIdx.x * static_cast<long long>(T) + threadIdx.x;
  if (i >= n) return;
  const SelCell& s = st[cells[i]];
  if (!s.active || !(s.value > 0.01)) return;
  const double y = xyz[3 * i + 1], p = s.value;
  double adj = 0.0;
  if (y < 0.1) adj = p;
  else if (y < 1.5) adj = __dmul_rn(p, __dadd_rn(1.0, -__ddiv_rn(__dadd_rn(y, -0.1), 1.4)));
  double out = __dadd_rn(y, -adj);
  if (out < 0.0) out = 0.0;
  xyz[3 * i + 1] = static_cast<float>(out);
  if (adj > 0.0) atomicAdd(&counters[3], 1ull);
}
__global__ void grid_stats_kernel(const SelCell* __restrict__ st, int ncells, unsigned long long* __restrict__ counters) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  if (st[c].count_all >= 10) atomicAdd(&counters[4], 1ull);          // cells with sufficient points
  if (st[c].active && st[c].value > 0.01) atomicAdd(&counters[5], 1ull);  // cells requiring adjustment
}

int grid_blocks(long long n) {
  const long long b = (n + T - 1) / T;
  return static_cast<int>(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

template <int MODE>
void select_percentile(const float* xyz, const double* dist, const unsigned short* cells, long long n, SelCell* st,
                       unsigned* hist, int ncells, double q, unsigned min_all, unsigned min_pop, cudaStream_t s) {
  const int gb = grid_blocks(n);
  sel_count_kernel<MODE><<<gb, T, 0, s>>>(xyz, dist, cells, n, st);
  DP_LAUNCH_CHECK();
  sel_begin_kernel<<<(ncells + 127) / 128, 128, 0, s>>>(st, ncells, q, min_all, min_pop);
  DP_LAUNCH_CHECK();
  const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
  unsigned fixed = 0;
  for (int p = 0; p < 3; ++p) {
    sel_hist_kernel<MODE><<<gb, T, 0, s>>>(xyz, dist, cells, n, st, hist, shifts[p], bits[p], fixed);
    DP_LAUNCH_CHECK();
    sel_scan_kernel<<<ncells, T, 0, s>>>(st, hist, shifts[p], bits[p], p == 2);
    DP_LAUNCH_CHECK();
    fixed |= ((1u << bits[p]) - 1u) << shifts[p];
  }
  sel_next_kernel<MODE><<<gb, T, 0, s>>>(xyz, dist, cells, n, st);
  DP_LAUNCH_CHECK();
  sel_finish_kernel<<<(ncells + 127) / 128, 128, 0, s>>>(st, ncells);
  DP_LAUNCH_CHECK();
}

}  // namespace

size_t ground_scratch_bytes(long long n, int grid_size) {
  const size_t cells = static_cast<size_t>(grid_size) * grid_size;
  size_t b = 0;
  b += (static_cast<size_t>(n) * 8 + 255) & ~size_t(255);                  // dist (double)
  b += (static_cast<size_t>(n) * 2 + 255) & ~size_t(255);                  // cell index
  b += (cells * sizeof(SelCell) + 255) & ~size_t(255);                     // select state
  b += cells * SEL_BINS * 4;                                               // histograms
  b += 256 + (2 * (static_cast<size_t>(grid_size) + 1) * 8 + 255) / 256 * 256;  // bounds, edges
  return b + 256;
}

namespace {
struct Scratch {
  double* dist;
  unsigned short* cells;
  SelCell* st;
  unsigned* hist;
  unsigned long long* bounds;
  double* edges;
};
Scratch carve(void* base, long long n, int grid_size) {
  const size_t cells = static_cast<size_t>(grid_size) * grid_size;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  Scratch sc;
  sc.dist = reinterpret_cast<double*>(p), p += (static_cast<size_t>(n) * 8 + 255) & ~size_t(255);
  sc.cells = reinterpret_cast<unsigned short*>(p), p += (static_cast<size_t>(n) * 2 + 255) & ~size_t(255);
  sc.st = reinterpret_cast<SelCell*>(p), p += (cells * sizeof(SelCell) + 255) & ~size_t(255);
  sc.hist = reinterpret_cast<unsigned*>(p), p += cells * SEL_BINS * 4;
  sc.bounds = reinterpret_cast<unsigned long long*>(p), p += 256;
  sc.edges = reinterpret_cast<double*>(p);
  return sc;
}
}  // namespace

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
  DP_CUDA(cudaMemsetAsync(sc.st, 0, sizeof(SelCell), s));
  DP_CUDA(cudaMemsetAsync(sc.hist, 0, SEL_BINS * 4, s));
  if (counters) DP_CUDA(cudaMemsetAsync(counters, 0, 3 * 8, s));
  const int nb = static_cast<int>((n + T - 1) / T);
  ground_transform_kernel<<<nb, T, 0, s>>>(xyz, sc.dist, n, xf);
  DP_LAUNCH_CHECK();
  select_percentile<0>(xyz, sc.dist, nullptr, n, sc.st, sc.hist, 1, 0.02, 0, 1, s);
  ground_clamp_kernel<<<nb, T, 0, s>>>(xyz, sc.dist, n, sc.st, counters ? counters : sc.bounds + 8);
  DP_LAUNCH_CHECK();
}

void ground_grid_adjust(float* xyz, long long n, int grid_size, double percentile, void* scratch, unsigned long long* counters,
                        cudaStream_t s) {
  if (n <= 0) return;
  DP_CHECK(grid_size >= 1 && grid_size <= 255, "grid_size must be in 1..255");
  const int ncells = grid_size * grid_size;
  Scratch sc = carve(scratch, n, grid_size);
  DP_CUDA(cudaMemsetAsync(sc.st, 0, ncells * sizeof(SelCell), s));
  DP_CUDA(cudaMemsetAsync(sc.hist, 0, static_cast<size_t>(ncells) * SEL_BINS * 4, s));
  if (counters) DP_CUDA(cudaMemsetAsync(counters + 3, 0, 3 * 8, s));
  const int nb = static_cast<int>((n + T - 1) / T);
  bounds_init_kernel<<<1, 1, 0, s>>>(sc.bounds);
  DP_LAUNCH_CHECK();
  bounds_kernel<<<grid_blocks(n), T, 0, s>>>(xyz, n, sc.bounds);
  DP_LAUNCH_CHECK();
  edges_kernel<<<1, 256, 0, s>>>(sc.bounds, sc.edges, grid_size);
  DP_LAUNCH_CHECK();
  cells_kernel<<<nb, T, 2 * (grid_size + 1) * sizeof(double), s>>>(xyz, n, sc.edges, grid_size, sc.cells);
  DP_LAUNCH_CHECK();
  // cells with >= 10 points and >= 5 low points (:1046, :1061); np.percentile(lowest_y, percentile)
  select_percentile<1>(xyz, nullptr, sc.cells, n, sc.st, sc.hist, ncells, percentile / 100.0, 10, 5, s);
  unsigned long long* ctr = counters ? counters : sc.bounds + 8;
  grid_apply_kernel<<<nb, T, 0, s>>>(xyz, n, sc.cells, sc.st, ctr);
  DP_LAUNCH_CHECK();
  grid_stats_kernel<<<(ncells + 127) / 128, 128, 0, s>>>(sc.st, ncells, ctr);
  DP_LAUNCH_CHECK();
}

}  // namespace dp
