// libvet_b200.so -- C ABI (include/vet_b200.h) over the sm_100a kernels.
// Host side: configuration validation, table construction (lattice, per-axis
// direction tables, cell->tile LUTs, FOV weight columns), scratch management and
// kernel launches.  No CPU compute path: every stage runs on the device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "vet_b200.h"
#include "vet_common.cuh"
#include "vet_stream.cuh"
#include "vet_stream_tma.cuh"
#include "vet_tables.cuh"
#include "vet_transition.cuh"
#include "vet_transition2.cuh"
#include "vet_transition3.cuh"
#include "vet_vectors.cuh"
#include "vet_whist.cuh"
#include "vet_whist_i8.cuh"
#include "vet_naive.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define VET_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return fail(VET_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

struct TileSet {
  int n = 0;  // tile_count as configured
  int T = 0;  // number of lattice points
  std::vector<double> h_centres;  // [T,3]
  std::vector<double> h_unit;     // [T,3] centres / ||centre||
  double* d_unit = nullptr;       // same on the device
  uint16_t* d_lut = nullptr;      // [C]
  uint8_t* d_lut8 = nullptr;      // [C] same table in bytes when T <= 255 (halves the shared-memory LUT)
  std::vector<uint16_t> h_lut;
  uint32_t* d_col_ptr = nullptr;  // [T+1]
  uint32_t* d_cell_idx = nullptr;
  double* d_w_val = nullptr;
  uint64_t nnz = 0;
  // grouped dense weight blocks for the batched weighted histogram (vet_whist.cuh)
  int G = 0;
  int32_t* d_group_tiles = nullptr;    // [G,8]
  uint32_t* d_group_chunk0 = nullptr;  // [G+1]
  double* d_chunks = nullptr;          // [nchunks][TG][Q][kChunkUnits]
  uint32_t* d_units = nullptr;         // [nchunks*kChunkUnits + pad]
  uint32_t nchunks = 0;
  std::vector<uint32_t> h_group_chunk0;  // host copy (item costs of the schedule)
  // per-CTA item schedule of k_whist, cached for the last frame count it was built for
  uint32_t* d_sched = nullptr;
  int64_t sched_F = -1;
  int sched_blocks = 0, sched_max_items = 0;
  double* d_hist = nullptr;            // [frames,T] scratch rows (grown on demand)
  size_t hist_bytes = 0;
  // int8 tensor-core path (vet_whist_i8.cuh), built on first use
  bool i8_built = false;
  int i8_blocks = 0;                   // N blocks of 48 tiles
  uint8_t* d_w8 = nullptr;             // [i8_blocks*240, kp] weight slices, row nb*240 + s*48 + j
  int2* d_kb_range = nullptr;          // [i8_blocks] K-block range of every N block
  CUtensorMap tm_w;
};

}  // namespace

struct vet_handle {
  int device = 0;
  int W = 0, H = 0;
  int64_t C = 0;
  int Cpad = 0;  // C rounded up to a multiple of 4: row pitch (in cells) of the per-frame cell histogram
  int K = 0;
  double fov = 120.0, pf = 2.0, max_d = 0.0;
  int use_weight = 1;
  // latitude/longitude grid tiling (NaiveSpatialEntropyAnalyzer): one tile set of grid codes
  bool naive = false;
  int naive_w = 0, naive_h = 0;
  int norm_T0 = 0;      // (180/h)*(360/w): the tile count the entropy is normalised by (EU:409)
  int norm_always = 0;  // config.use_weight_distribution (EU:443)
  int sm_count = 0;
  size_t smem_optin = 0;
  int maxT = 0;
  std::vector<TileSet> ts;
  double *d_cosT = nullptr, *d_sinT = nullptr, *d_sinP = nullptr, *d_cosP = nullptr;
  double* d_cellvec = nullptr;  // [C,3]
  uint32_t* d_flags = nullptr;
  // scratch (grown on demand)
  uint32_t* d_cnt = nullptr;
  size_t cnt_bytes = 0;
  uint32_t* d_nvalid = nullptr;  // [frames] present users per frame
  size_t nvalid_bytes = 0;
  uint32_t* d_work = nullptr;    // work counters of the dynamic schedulers
  uint32_t* d_lut_packed = nullptr;  // [C] byte k = tile under tile count k (K <= 4 and every T <= 255), else null
  uint32_t* d_ihist = nullptr;   // [frames, sum T_k] integer tile histograms (direct unweighted path)
  size_t ihist_bytes = 0;
  int sumT = 0;
  bool direct_only = false;      // video too large for the cell tables: packed input goes decode -> vectors path
  bool global_tables = false;    // cell grid too large for shared memory but small enough for per-cell tables in
                                 // global memory: k_stream_global + the usual table-regime epilogues
  uint16_t* d_identity = nullptr;  // [maxT] identity LUT (vectors path feeds tile indices to k_transition)
  void* d_vscratch[3] = {nullptr, nullptr, nullptr};  // idx[F,U] i32, per_k[K,F] f64, vec[F,U,3] f64
  size_t vscratch_bytes[3] = {0, 0, 0};
  void* d_cells = nullptr;
  size_t cells_bytes = 0;
  uint32_t* d_tables = nullptr;
  size_t tables_words = 0;
  uint32_t* d_pairs = nullptr;  // [CTAs, U] packed (prev, cur) tiles of the frame pair in flight (k_transition2)
  size_t pairs_bytes = 0;
  // int8 tensor-core weighted histogram: count byte planes [3][plane_rows][kp] (rows of planes 1, 2 are zero
  // unless marked in d_dirty), per-frame-block flags hi1/hi2 [2][plane_rows/128]
  uint8_t* d_planes = nullptr;
  int64_t plane_rows = 0;        // row capacity of the allocation (multiple of 128)
  uint8_t* d_dirty = nullptr;    // [plane_rows]
  uint32_t* d_i8flags = nullptr; // [2][plane_rows/128]
  CUtensorMap tm_cnt;
  bool planes_from_stream = false;  // the last launch_stream wrote the planes of its batch itself
  bool i8_attr_set = false;
  uint32_t* d_redo = nullptr;   // [rows] frame pairs the two-pass transition kernel left to k_transition2
  size_t redo_bytes = 0;
  double* d_trk = nullptr;      // [K, rows] per-tile-count transition entropies when the caller wants none
  size_t trk_bytes = 0;
  uint32_t tables_cap = 0;  // slot count the tables are currently laid out (and cleared) for
  int tables_blocks = 0;    // number of per-CTA tables cleared for that layout
  // host-buffer path
  void* d_in[2] = {nullptr, nullptr};
  size_t in_bytes = 0;
  void* d_hout[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // entropy, per_k, hist0, assign x2
  size_t hout_bytes[5] = {0, 0, 0, 0, 0};
  cudaStream_t s_copy = nullptr, s_exec = nullptr, s_out = nullptr;
  int64_t launches = 0;
  // optional per-kernel timing (vet_profile_*): CUDA events recorded around each launch
  bool profiling = false;
  struct Span {
    int kernel;
    cudaEvent_t a, b;
  };
  std::vector<Span> spans;
};

namespace {

// ---- host-side table construction (libm; used when the caller passes no tables) ----

// numpy's remainder for doubles (npy_divmod): result takes the sign of the divisor.
double np_mod(double a, double b) {
  double m = std::fmod(a, b);
  if (m != 0.0) {
    if ((b < 0) != (m < 0)) m += b;
  } else {
    m = std::copysign(0.0, b);
  }
  return m;
}
double np_radians(double x) { return x * (M_PI / 180.0); }
double np_round6(double v) { return std::rint(v * 1e6) / 1e6; }
// CPython round(x, 1): correctly rounded decimal (round-half-even on the exact
// binary value) -- glibc's printf does exactly that.
double py_round1(double x) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.1f", x);
  return strtod(buf, nullptr);
}

// Vector.from_spherical, DT:204-216.
void from_spherical(double lon, double lat, double* out) {
  const double theta = np_radians(lon), phi = np_radians(90 - lat);
  out[0] = np_round6(std::sin(phi) * std::cos(theta));
  out[1] = np_round6(std::sin(phi) * std::sin(theta));
  out[2] = np_round6(std::cos(phi));
}

// generate_fibonacci_lattice, DU:40-54.
std::vector<double> make_lattice(int n) {
  const double phi = (1 + std::sqrt(5.0)) / 2;
  const int N = n / 2;
  std::vector<double> c((size_t)(2 * N + 1) * 3);
  for (int i = -N; i <= N; ++i) {
    const double lat = std::asin(2.0 * i / (2 * N + 1)) * 180 / M_PI;
    double lon = np_mod((double)i, phi) * 360 / phi;
    lon = np_mod(lon + 180, 360.0) - 180;
    from_spherical(lon, lat, &c[(size_t)(i + N) * 3]);
  }
  return c;
}

// pixel_to_spherical + rounding + wrap quirk, DU:283-284, 390-397.
void make_axis_tables(int W, int H, std::vector<double>& lon, std::vector<double>& lat) {
  lon.resize(W + 1);
  lat.resize(H + 1);
  for (int px = 0; px <= W; ++px) {
    double v = ((double)px / W) * 360 - 180;
    v = py_round1(v);
    if (v <= -180) v = np_mod(v + 360, 360.0) - 180;
    lon[px] = v;
  }
  for (int py = 0; py <= H; ++py) {
    double v = 90 - ((double)py / H) * 180;
    v = py_round1(v);
    if (v <= -90) v = np_mod(v + 180, 180.0) - 90;
    lat[py] = v;
  }
}

template <typename T>
int upload(T** dptr, const T* host, size_t count) {
  VET_CUDA(cudaMalloc((void**)dptr, std::max<size_t>(count, 1) * sizeof(T)));
  if (count) VET_CUDA(cudaMemcpy(*dptr, host, count * sizeof(T), cudaMemcpyHostToDevice));
  return VET_OK;
}

int grow(void** ptr, size_t* have, size_t want) {
  if (*have >= want) return VET_OK;
  if (*ptr) VET_CUDA(cudaFree(*ptr));
  *ptr = nullptr;
  *have = 0;
  VET_CUDA(cudaMalloc(ptr, want));
  *have = want;
  return VET_OK;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct LaunchTimer {  // records an event pair around one kernel launch when profiling is on
  vet_handle* h;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  int kernel;
  LaunchTimer(vet_handle* h_, int kernel_, cudaStream_t st_) : h(h_), st(st_), kernel(kernel_) {
    h->launches++;
    if (h->profiling && cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess) cudaEventRecord(a, st);
  }
  ~LaunchTimer() {
    if (a && b) {
      cudaEventRecord(b, st);
      h->spans.push_back({kernel, a, b});
    }
  }
};

int launch_transition(vet_handle* h, vet::TransitionArgs& a, int64_t rows, int64_t U, int Tmax, cudaStream_t st);
int spatial_direct(vet_handle* h, const void* packed, int dtype, int64_t F, int64_t U, double* entropy, double* per_k,
                   double* hist0, uint16_t* assign0, cudaStream_t st);
int transition_direct(vet_handle* h, const void* packed, int dtype, int64_t F, int64_t U, double* entropy, double* per_k,
                      int32_t* prev_count0, uint16_t* pairs0, int mode, cudaStream_t st);

constexpr size_t kStaticSmemSlack = 1024;
constexpr int kMaxT = 16384;
constexpr int64_t kGlobalTableCells = 262144;  // largest cell grid of the global-table regime (tables scale with C*T)
constexpr int64_t kGlobalLutCells = (int64_t)1 << 24;  // the same for unweighted handles (only LUTs: 2 B x C per tile count)

// tensor-core weighted histogram (defined with launch_whist_i8 below)
bool use_whist_i8(const vet_handle* h, int64_t F, int64_t U);
int ensure_planes(vet_handle* h, int64_t F, cudaStream_t st);
int64_t i8_kp(const vet_handle* h);
uint32_t* i8_hi1(vet_handle* h);
uint32_t* i8_hi2(vet_handle* h);

// Blocking of k_whist: "wide" = 8 tiles x 8 frames per warp, "tall" = 4 tiles x 16 frames.
// 0 = wide (default), 1 = tall, 2 = quad
int whist_shape() {
  static const int shape = [] {
    const char* e = getenv("VET_WHIST_SHAPE");
    if (e && std::string(e) == "tall") return 1;  // measured: wide 0.49 ms, tall 0.67 ms on configs[2]
    if (e && std::string(e) == "quad") return 2;
    return 0;
  }();
  return shape;
}

// Longest-processing-time schedule of the (frame block, group) items of k_whist over the
// CTAs (items differ in size: a group's cost is its number of weight chunks); each CTA's
// list is then put in frame-block-major order for L2 locality.
int build_whist_schedule(TileSet& t, int64_t fblocks, int blocks) {
  struct Item {
    uint32_t id, cost;
  };
  std::vector<Item> items;
  items.reserve((size_t)fblocks * t.G);
  for (int64_t fb = 0; fb < fblocks; ++fb)
    for (int g = 0; g < t.G; ++g)
      items.push_back({(uint32_t)(fb * t.G + g), t.h_group_chunk0[g + 1] - t.h_group_chunk0[g] + 2});
  std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.cost > b.cost; });
  std::vector<std::vector<uint32_t>> lists(blocks);
  std::vector<std::pair<uint64_t, int>> load(blocks);  // min-heap on (load, cta)
  for (int b = 0; b < blocks; ++b) load[b] = {0, b};
  auto cmp = [](const std::pair<uint64_t, int>& a, const std::pair<uint64_t, int>& b) { return a > b; };
  std::make_heap(load.begin(), load.end(), cmp);
  for (const Item& it : items) {
    std::pop_heap(load.begin(), load.end(), cmp);
    auto& top = load.back();
    lists[top.second].push_back(it.id);
    top.first += it.cost;
    std::push_heap(load.begin(), load.end(), cmp);
  }
  size_t max_items = 1;
  for (auto& l : lists) {
    std::sort(l.begin(), l.end());
    max_items = std::max(max_items, l.size());
  }
  std::vector<uint32_t> flat((size_t)blocks * max_items, 0xFFFFFFFFu);
  for (int b = 0; b < blocks; ++b) std::copy(lists[b].begin(), lists[b].end(), flat.begin() + (size_t)b * max_items);
  if (t.d_sched) cudaFree(t.d_sched);
  t.d_sched = nullptr;
  if (int rc = upload(&t.d_sched, flat.data(), flat.size())) return rc;
  t.sched_blocks = blocks;
  t.sched_max_items = (int)max_items;
  return VET_OK;
}

// Clusters the tiles into groups of TG spatial neighbours and lays every group's
// weights out as dense [cells][kTG] blocks over the union of the members' supports
// (see vet_whist.cuh).  Values are the device-computed ones of the column table.
template <typename S>
int build_weight_groups(vet_handle* h, TileSet& t, const std::vector<uint32_t>& col_ptr, const std::vector<double>& unit) {
  constexpr int kTG = S::TG, kQ = S::Q, kChunkCells = S::kChunkCells;
  const int T = t.T;
  std::vector<uint32_t> cell_idx(std::max<uint64_t>(t.nnz, 1));
  std::vector<double> w_val(std::max<uint64_t>(t.nnz, 1));
  if (t.nnz) {
    VET_CUDA(cudaMemcpy(cell_idx.data(), t.d_cell_idx, t.nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    VET_CUDA(cudaMemcpy(w_val.data(), t.d_w_val, t.nnz * sizeof(double), cudaMemcpyDeviceToHost));
  }
  // greedy proximity clustering: seed = lowest unassigned tile, members = its nearest unassigned tiles
  std::vector<char> used(T, 0);
  std::vector<int32_t> group_tiles;
  std::vector<std::pair<double, int>> cand;
  for (int seed = 0; seed < T; ++seed) {
    if (used[seed]) continue;
    cand.clear();
    for (int j = 0; j < T; ++j)
      if (!used[j]) {
        const double d = unit[3 * seed] * unit[3 * j] + unit[3 * seed + 1] * unit[3 * j + 1] + unit[3 * seed + 2] * unit[3 * j + 2];
        cand.emplace_back(-d, j);
      }
    const size_t take = std::min<size_t>(kTG, cand.size());
    std::partial_sort(cand.begin(), cand.begin() + take, cand.end());
    for (int m = 0; m < kTG; ++m) {
      if ((size_t)m < take) {
        group_tiles.push_back(cand[m].second);
        used[cand[m].second] = 1;
      } else {
        group_tiles.push_back(-1);
      }
    }
  }
  const int G = (int)(group_tiles.size() / kTG);
  std::vector<uint32_t> chunk0(G + 1, 0);
  std::vector<double> chunks;       // [nchunks][TG][Q][kChunkUnits]
  std::vector<uint32_t> units_all;  // [nchunks][kChunkUnits] first cell of each load unit
  const int64_t n_units_total = (h->Cpad + kQ - 1) / kQ;
  std::vector<int32_t> slot(n_units_total, -1);
  std::vector<uint32_t> units;
  const size_t chunk_doubles = (size_t)kChunkCells * kTG;
  for (int g = 0; g < G; ++g) {
    units.clear();
    for (int m = 0; m < kTG; ++m) {
      const int tile = group_tiles[g * kTG + m];
      if (tile < 0) continue;
      for (uint32_t j = col_ptr[tile]; j < col_ptr[tile + 1]; ++j) {
        const uint32_t u = cell_idx[j] / kQ;
        if (slot[u] < 0) {
          slot[u] = 0;
          units.push_back(u);
        }
      }
    }
    std::sort(units.begin(), units.end());
    for (size_t i = 0; i < units.size(); ++i) slot[units[i]] = (int32_t)i;
    const uint32_t nch = (uint32_t)std::max<size_t>(1, (units.size() + S::kChunkUnits - 1) / S::kChunkUnits);
    chunk0[g] = (uint32_t)(chunks.size() / chunk_doubles);
    const size_t base = chunks.size();
    chunks.resize(base + (size_t)nch * chunk_doubles, 0.0);             // zero weights for padding
    units_all.resize((size_t)(chunk0[g] + nch) * S::kChunkUnits, 0);  // padding units point at cell 0
    for (size_t i = 0; i < units.size(); ++i) units_all[(size_t)chunk0[g] * S::kChunkUnits + i] = units[i] * kQ;
    for (int m = 0; m < kTG; ++m) {
      const int tile = group_tiles[g * kTG + m];
      if (tile < 0) continue;
      for (uint32_t j = col_ptr[tile]; j < col_ptr[tile + 1]; ++j) {
        const size_t i = (size_t)slot[cell_idx[j] / kQ];
        const int q = (int)(cell_idx[j] % kQ);
        double* ch = chunks.data() + base + (i / S::kChunkUnits) * chunk_doubles;
        ch[(m * kQ + q) * S::kChunkUnits + i % S::kChunkUnits] = w_val[j];
      }
    }
    for (uint32_t u : units) slot[u] = -1;
  }
  chunk0[G] = (uint32_t)(chunks.size() / chunk_doubles);
  units_all.resize((size_t)chunk0[G] * S::kChunkUnits + vet::kUnitPad, 0);
  t.G = G;
  t.nchunks = chunk0[G];
  if (int rc = upload(&t.d_group_tiles, group_tiles.data(), group_tiles.size())) return rc;
  if (int rc = upload(&t.d_group_chunk0, chunk0.data(), chunk0.size())) return rc;
  t.h_group_chunk0 = chunk0;
  if (int rc = upload(&t.d_chunks, chunks.data(), chunks.size())) return rc;
  if (int rc = upload(&t.d_units, units_all.data(), units_all.size())) return rc;
  return VET_OK;
}

// unit tile centres c/||c|| (EU:59), uploaded as t.d_unit and kept in t.h_unit
int build_unit_centres(TileSet& t) {
  const int T = t.T;
  t.h_unit.resize((size_t)T * 3);
  for (int i = 0; i < T; ++i) {
    const double x = t.h_centres[3 * i], y = t.h_centres[3 * i + 1], z = t.h_centres[3 * i + 2];
    // np.linalg.norm == sqrt(dot(x,x)), ddot as an FMA chain
    const double nrm = std::sqrt(std::fma(z, z, std::fma(y, y, x * x)));
    if (!(nrm > 0)) return fail(VET_ERR_INVALID_ARG, "Vector cannot have zero length (tile %d)", i);
    t.h_unit[3 * i] = x / nrm;
    t.h_unit[3 * i + 1] = y / nrm;
    t.h_unit[3 * i + 2] = z / nrm;
  }
  return upload(&t.d_unit, t.h_unit.data(), t.h_unit.size());
}

int build_tile_set(vet_handle* h, TileSet& t) {
  const int T = t.T;
  if (int rc = build_unit_centres(t)) return rc;
  const std::vector<double>& unit = t.h_unit;
  VET_CUDA(cudaMalloc((void**)&t.d_lut, (size_t)h->C * sizeof(uint16_t) + 16));  // readable in 16 B units
  const size_t smem = (size_t)T * 3 * sizeof(double);
  VET_CUDA(cudaFuncSetAttribute(vet::k_nearest<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = 256;
  const int64_t rounds = (h->C + threads / 4 - 1) / (threads / 4);
  const int blocks = (int)std::min<int64_t>(rounds, (int64_t)h->sm_count * 8);
  vet::k_nearest<uint16_t><<<blocks, threads, smem>>>(h->d_cellvec, h->C, t.d_unit, T, t.d_lut);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  t.h_lut.resize(h->C);
  VET_CUDA(cudaMemcpy(t.h_lut.data(), t.d_lut, (size_t)h->C * sizeof(uint16_t), cudaMemcpyDeviceToHost));
  if (T <= 255) {
    std::vector<uint8_t> l8(h->C + 16, 0);  // padded: the kernels copy it in 16 B units
    for (int64_t c = 0; c < h->C; ++c) l8[c] = (uint8_t)t.h_lut[c];
    if (int rc = upload(&t.d_lut8, l8.data(), l8.size())) return rc;
  }
  if (h->use_weight) {
    uint32_t* d_count = nullptr;
    VET_CUDA(cudaMalloc((void**)&d_count, (size_t)T * sizeof(uint32_t)));
    vet::k_weight_columns<false><<<T, 256>>>(h->d_cellvec, (int)h->C, t.d_unit, T, h->max_d, h->pf, d_count, nullptr,
                                             nullptr, nullptr);
    h->launches++;
    VET_CUDA(cudaGetLastError());
    std::vector<uint32_t> count(T), ptr(T + 1, 0);
    VET_CUDA(cudaMemcpy(count.data(), d_count, (size_t)T * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    VET_CUDA(cudaFree(d_count));
    uint64_t nnz = 0;
    for (int i = 0; i < T; ++i) {
      ptr[i] = (uint32_t)nnz;
      nnz += count[i];
    }
    if (nnz >= 0xFFFFFFFFull) return fail(VET_ERR_UNSUPPORTED, "weight table too large (%llu entries)", (unsigned long long)nnz);
    ptr[T] = (uint32_t)nnz;
    t.nnz = nnz;
    if (int rc = upload(&t.d_col_ptr, ptr.data(), ptr.size())) return rc;
    VET_CUDA(cudaMalloc((void**)&t.d_cell_idx, std::max<uint64_t>(nnz, 1) * sizeof(uint32_t)));
    VET_CUDA(cudaMalloc((void**)&t.d_w_val, std::max<uint64_t>(nnz, 1) * sizeof(double)));
    vet::k_weight_columns<true><<<T, 256>>>(h->d_cellvec, (int)h->C, t.d_unit, T, h->max_d, h->pf, nullptr, t.d_col_ptr,
                                            t.d_cell_idx, t.d_w_val);
    h->launches++;
    VET_CUDA(cudaGetLastError());
    const int shape = whist_shape();
    if (int rc = shape == 1   ? build_weight_groups<vet::WhistTall>(h, t, ptr, unit)
                 : shape == 2 ? build_weight_groups<vet::WhistQuad>(h, t, ptr, unit)
                              : build_weight_groups<vet::WhistWide>(h, t, ptr, unit))
      return rc;
  }
  return VET_OK;
}

// Grid tiling: cell -> code LUT from the per-axis degree tables (find_naive_tile_index, EU:378-381).
int build_naive_tile_set(vet_handle* h, TileSet& t, const std::vector<double>& lon, const std::vector<double>& lat) {
  const int nlat1 = 180 / h->naive_h + 1;
  std::vector<int> li(h->W + 1), la(h->H + 1);
  for (int px = 0; px <= h->W; ++px) li[px] = (int)((lon[px] + 180) / h->naive_w);
  for (int py = 0; py <= h->H; ++py) la[py] = (int)((lat[py] + 90) / h->naive_h);
  t.h_lut.resize(h->C);
  for (int py = 0; py <= h->H; ++py)
    for (int px = 0; px <= h->W; ++px) t.h_lut[(size_t)py * (h->W + 1) + px] = (uint16_t)(li[px] * nlat1 + la[py]);
  std::vector<uint16_t> l16(h->C + 8, 0);
  std::copy(t.h_lut.begin(), t.h_lut.end(), l16.begin());
  if (int rc = upload(&t.d_lut, l16.data(), l16.size())) return rc;
  if (t.T <= 255) {
    std::vector<uint8_t> l8(h->C + 16, 0);
    for (int64_t c = 0; c < h->C; ++c) l8[c] = (uint8_t)t.h_lut[c];
    if (int rc = upload(&t.d_lut8, l8.data(), l8.size())) return rc;
  }
  return VET_OK;
}

void free_tile_set(TileSet& t) {
  cudaFree(t.d_unit);
  cudaFree(t.d_lut);
  cudaFree(t.d_lut8);
  cudaFree(t.d_col_ptr);
  cudaFree(t.d_cell_idx);
  cudaFree(t.d_w_val);
  cudaFree(t.d_group_tiles);
  cudaFree(t.d_group_chunk0);
  cudaFree(t.d_chunks);
  cudaFree(t.d_units);
  cudaFree(t.d_sched);
  cudaFree(t.d_hist);
  cudaFree(t.d_w8);
  cudaFree(t.d_kb_range);
}

size_t stream_smem_bytes(const vet_handle* h) { return (size_t)h->Cpad * 4 + (size_t)h->C * 2 + 16; }
size_t stream_tma_smem_bytes(const vet_handle* h, bool lut8) {
  return (size_t)vet::kStages * vet::kStageBytes + (size_t)h->Cpad * 4 + (size_t)h->C * (lut8 ? 1 : 2) + 32;
}
size_t epilogue_smem_bytes(const vet_handle* h) {
  return (size_t)h->Cpad * 4 + (size_t)h->maxT * 8 + (size_t)h->maxT * 4 + 16;
}
bool use_tma_stream(const vet_handle* h, const void* packed) {
  static const bool force_simple = [] {
    const char* e = getenv("VET_STREAM_IMPL");
    return e && std::string(e) == "simple";
  }();
  if (force_simple) return false;
  if (((uintptr_t)packed & 15) != 0) return false;  // bulk copies need a 16 B aligned tensor base
  const bool lut8 = h->ts[0].d_lut8 != nullptr;
  return stream_tma_smem_bytes(h, lut8) + kStaticSmemSlack <= h->smem_optin;
}

// bytes of the cell-histogram scratch for batches of fb frames (none where no kernel of the handle uses it)
size_t cnt_scratch_bytes(const vet_handle* h, int64_t fb) {
  if (h->global_tables && !h->use_weight) return 16;
  return (size_t)(fb + vet::kWhRowPad) * h->Cpad * 4;
}

// frames per batch so that the per-frame cell histogram scratch stays bounded
int64_t frames_per_batch(const vet_handle* h, int64_t F, int64_t U, bool need_cells) {
  const size_t budget = (size_t)1 << 30;  // 1 GiB of scratch
  size_t per_frame = (size_t)h->Cpad * 4;
  if (h->global_tables && !h->use_weight) per_frame = (size_t)h->sumT * 4;  // tile histograms only, no cell histogram
  if (need_cells) per_frame += (size_t)U * (h->C <= 65535 ? 2 : 4);
  int64_t fb = (int64_t)std::max<size_t>(2, budget / std::max<size_t>(per_frame, 1));
  return std::min<int64_t>(F, fb);
}

int launch_stream(vet_handle* h, const void* packed, int dtype, int64_t F, int64_t U, uint16_t* assign0, bool cells,
                  cudaStream_t st) {
  vet::StreamArgs a{};
  a.packed = packed;
  a.F = F;
  a.U = U;
  a.W = h->W;
  a.H = h->H;
  a.C = (int)h->C;
  a.lut0 = h->ts[0].d_lut;
  a.assign0 = assign0;
  a.cell16 = (cells && h->C <= 65535) ? (uint16_t*)h->d_cells : nullptr;
  a.cell32 = (cells && h->C > 65535) ? (int32_t*)h->d_cells : nullptr;
  a.cnt = h->d_cnt;
  a.nvalid = h->d_nvalid;
  a.flags = h->d_flags;
  // enough work items to balance the SMs, chunks no smaller than 32k users
  const int64_t want_items = (int64_t)h->sm_count * 24;
  int64_t cpf = std::min<int64_t>((U + 32767) / 32768, (want_items + F - 1) / F);
  cpf = std::max<int64_t>(cpf, 1);
  a.chunk_users = (U + cpf - 1) / cpf;
  a.chunks_per_frame = (int)((U + a.chunk_users - 1) / std::max<int64_t>(a.chunk_users, 1));
  if (a.chunks_per_frame < 1) a.chunks_per_frame = 1;
  a.cpad = h->Cpad;
  if (a.chunks_per_frame > 1 && !h->global_tables) {
    VET_CUDA(cudaMemsetAsync(h->d_cnt, 0, (size_t)F * h->Cpad * 4, st));
    VET_CUDA(cudaMemsetAsync(h->d_nvalid, 0, (size_t)F * 4, st));
  }
  h->planes_from_stream = false;
  if (h->global_tables) {
    vet::StreamGlobalArgs G{};
    G.s = a;
    G.s.chunks_per_frame = 1;
    G.K = h->K;
    G.sumT = h->sumT;
    int off = 0;
    for (int k = 0; k < h->K; ++k) {
      G.hist_off[k] = off;
      off += h->ts[k].T;
      G.lut[k] = h->ts[k].d_lut;
    }
    VET_CUDA(cudaMemsetAsync(h->d_nvalid, 0, (size_t)F * 4, st));
    if (h->use_weight) {
      VET_CUDA(cudaMemsetAsync(h->d_cnt, 0, (size_t)F * h->Cpad * 4, st));
    } else {  // unweighted: tile histograms directly, no cell histogram
      G.s.cnt = nullptr;
      if (int rc = grow((void**)&h->d_ihist, &h->ihist_bytes, (size_t)F * h->sumT * 4)) return rc;
      VET_CUDA(cudaMemsetAsync(h->d_ihist, 0, (size_t)F * h->sumT * 4, st));
      G.ihist = h->d_ihist;
    }
    const int gblocks = (int)std::min<int64_t>((F * U + 255) / 256, (int64_t)h->sm_count * 16);
    LaunchTimer lt(h, VET_KERNEL_STREAM, st);
    if (dtype == VET_F32) vet::k_stream_global<float><<<gblocks, 256, 0, st>>>(G);
    else vet::k_stream_global<double><<<gblocks, 256, 0, st>>>(G);
    VET_CUDA(cudaGetLastError());
    return VET_OK;
  }
  const int64_t items = F * a.chunks_per_frame;
  const int blocks = (int)std::min<int64_t>(items, h->sm_count);
  if (use_tma_stream(h, packed)) {
    const bool lut8 = h->ts[0].d_lut8 != nullptr;
    vet::StreamTmaArgs A{};
    A.s = a;
    A.lut0_typed = lut8 ? (const void*)h->ts[0].d_lut8 : (const void*)h->ts[0].d_lut;
    A.total_bytes = F * U * 3 * (int64_t)(dtype == VET_F32 ? 4 : 8);
    A.cpad = h->Cpad;
    if (h->use_weight && a.chunks_per_frame == 1 && use_whist_i8(h, F, U)) {
      // frames of one chunk: the kernel writes the byte planes of the tensor-core epilogue instead of `cnt`
      if (int rc = ensure_planes(h, F, st)) return rc;
      VET_CUDA(cudaMemsetAsync(h->d_i8flags, 0, (size_t)2 * (h->plane_rows / vet::kI8M) * 4, st));
      A.planes = h->d_planes;
      A.kp = (int)i8_kp(h);
      A.plane_stride = h->plane_rows * (int64_t)A.kp;
      A.dirty = h->d_dirty;
      A.hi1 = i8_hi1(h);
      A.hi2 = i8_hi2(h);
      h->planes_from_stream = true;
    }
    const size_t smem = stream_tma_smem_bytes(h, lut8);
    LaunchTimer lt(h, VET_KERNEL_STREAM, st);
    const dim3 grid(blocks), block(vet::kStreamThreads);
#define VET_LAUNCH_STREAM(TIN, TLUT, ASSIGN, CELLS) vet::k_stream_tma<TIN, TLUT, ASSIGN, CELLS><<<grid, block, smem, st>>>(A)
    const int cmode = a.cell16 ? 1 : (a.cell32 ? 2 : 0);
#define VET_LAUNCH_ASSIGN(TIN, CELLS)                              \
  do {                                                             \
    if (lut8) VET_LAUNCH_STREAM(TIN, uint8_t, true, CELLS);        \
    else VET_LAUNCH_STREAM(TIN, uint16_t, true, CELLS);            \
  } while (0)
    if (assign0) {  // spatial stage (cmode 0) or both analyzers in one pass (cell ids as well)
      if (dtype == VET_F32) {
        if (cmode == 0) VET_LAUNCH_ASSIGN(float, 0);
        else if (cmode == 1) VET_LAUNCH_ASSIGN(float, 1);
        else VET_LAUNCH_ASSIGN(float, 2);
      } else {
        if (cmode == 0) VET_LAUNCH_ASSIGN(double, 0);
        else if (cmode == 1) VET_LAUNCH_ASSIGN(double, 1);
        else VET_LAUNCH_ASSIGN(double, 2);
      }
    } else if (dtype == VET_F32) {
      if (cmode == 0) VET_LAUNCH_STREAM(float, uint8_t, false, 0);
      else if (cmode == 1) VET_LAUNCH_STREAM(float, uint8_t, false, 1);
      else VET_LAUNCH_STREAM(float, uint8_t, false, 2);
    } else {
      if (cmode == 0) VET_LAUNCH_STREAM(double, uint8_t, false, 0);
      else if (cmode == 1) VET_LAUNCH_STREAM(double, uint8_t, false, 1);
      else VET_LAUNCH_STREAM(double, uint8_t, false, 2);
    }
#undef VET_LAUNCH_ASSIGN
#undef VET_LAUNCH_STREAM
  } else {
    const size_t smem = stream_smem_bytes(h);
    LaunchTimer lt(h, VET_KERNEL_STREAM, st);
    if (dtype == VET_F32)
      vet::k_stream_simple<float><<<blocks, 1024, smem, st>>>(a);
    else
      vet::k_stream_simple<double><<<blocks, 1024, smem, st>>>(a);
  }
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

// ---- direct unweighted path: per-sample tile lookups, tile histograms, no cell histogram ----
struct TilesPlan {
  bool ok = false;
  vet::StreamTilesArgs A{};
  size_t smem = 0;
};

TilesPlan plan_tiles(const vet_handle* h, const void* packed, int64_t U) {
  TilesPlan p;
  if (h->use_weight || h->direct_only || h->global_tables) return p;
  if (((uintptr_t)packed & 15) != 0) return p;
  static const bool disabled = [] {
    const char* e = getenv("VET_STREAM_IMPL");
    return e && (std::string(e) == "simple" || std::string(e) == "cells");
  }();
  if (disabled) return p;
  // worth it when there is a single tile count or the frames are small against the cell grid
  if (!(h->K == 1 || U < 2 * h->C)) return p;
  int off = 0, hoff = 0, soff = 0;
  for (int k = 0; k < h->K; ++k) {
    const TileSet& t = h->ts[k];
    p.A.T[k] = t.T;
    p.A.hist_off[k] = hoff;
    hoff += t.T;
    // Interleaved copies of small histograms were measured SLOWER on B200 (configs[1]: 0.20 vs 0.17 ms):
    // the ATOMS.POPC.INC path already absorbs same-address increments, so one copy is used.
    const int rs = 0;
    p.A.rep_shift[k] = rs;
    p.A.shist_off[k] = soff;
    soff += t.T << rs;
    p.A.lut_wide[k] = t.d_lut8 ? 0 : 1;
    p.A.lut[k] = t.d_lut8 ? (const void*)t.d_lut8 : (const void*)t.d_lut;
    p.A.lut_off[k] = off;
    off += (int)((h->C * (t.d_lut8 ? 1 : 2) + 15) & ~(int64_t)15);
  }
  p.A.K = h->K;
  p.A.sumT = hoff;
  p.A.shist_words = soff;
  p.A.lut_packed = h->d_lut_packed;
  if (h->d_lut_packed) off = (int)(((size_t)h->C * 4 + 15) & ~(size_t)15);
  p.A.lut_bytes = off;
  p.smem = (size_t)vet::kStages * vet::kStageBytes + (size_t)((soff + 3) & ~3) * 4 + off + 16;
  p.ok = p.smem + kStaticSmemSlack <= h->smem_optin;
  return p;
}

int launch_stream_tiles(vet_handle* h, TilesPlan& p, const void* packed, int dtype, int64_t F, int64_t U, uint16_t* assign0,
                        cudaStream_t st) {
  vet::StreamArgs a{};
  a.packed = packed;
  a.F = F;
  a.U = U;
  a.W = h->W;
  a.H = h->H;
  a.C = (int)h->C;
  a.assign0 = assign0;
  a.nvalid = h->d_nvalid;
  a.flags = h->d_flags;
  const int64_t want_items = (int64_t)h->sm_count * 24;
  int64_t cpf = std::min<int64_t>((U + 32767) / 32768, (want_items + F - 1) / F);
  cpf = std::max<int64_t>(cpf, 1);
  a.chunk_users = (U + cpf - 1) / cpf;
  a.chunks_per_frame = (int)std::max<int64_t>(1, (U + a.chunk_users - 1) / std::max<int64_t>(a.chunk_users, 1));
  if (int rc = grow((void**)&h->d_ihist, &h->ihist_bytes, (size_t)F * p.A.sumT * 4)) return rc;
  if (a.chunks_per_frame > 1) {
    VET_CUDA(cudaMemsetAsync(h->d_ihist, 0, (size_t)F * p.A.sumT * 4, st));
    VET_CUDA(cudaMemsetAsync(h->d_nvalid, 0, (size_t)F * 4, st));
  }
  p.A.s = a;
  p.A.total_bytes = F * U * 3 * (int64_t)(dtype == VET_F32 ? 4 : 8);
  p.A.ihist = h->d_ihist;
  const int blocks = (int)std::min<int64_t>(F * a.chunks_per_frame, h->sm_count);
  const int sm = (int)(h->smem_optin - kStaticSmemSlack);
  LaunchTimer lt(h, VET_KERNEL_STREAM, st);
#define VET_LAUNCH_TILES(TIN, ASSIGN, KP)                                                                              \
  do {                                                                                                                  \
    VET_CUDA(cudaFuncSetAttribute(vet::k_stream_tiles<TIN, ASSIGN, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)); \
    vet::k_stream_tiles<TIN, ASSIGN, KP><<<blocks, vet::kStreamThreads, p.smem, st>>>(p.A);                             \
  } while (0)
#define VET_LAUNCH_TILES_K(TIN, ASSIGN)                        \
  do {                                                         \
    switch (h->d_lut_packed ? h->K : 0) {                      \
      case 1: VET_LAUNCH_TILES(TIN, ASSIGN, 1); break;         \
      case 2: VET_LAUNCH_TILES(TIN, ASSIGN, 2); break;         \
      case 3: VET_LAUNCH_TILES(TIN, ASSIGN, 3); break;         \
      case 4: VET_LAUNCH_TILES(TIN, ASSIGN, 4); break;         \
      default: VET_LAUNCH_TILES(TIN, ASSIGN, 0); break;        \
    }                                                          \
  } while (0)
  if (dtype == VET_F32) {
    if (assign0) VET_LAUNCH_TILES_K(float, true);
    else VET_LAUNCH_TILES_K(float, false);
  } else {
    if (assign0) VET_LAUNCH_TILES_K(double, true);
    else VET_LAUNCH_TILES_K(double, false);
  }
#undef VET_LAUNCH_TILES_K
#undef VET_LAUNCH_TILES
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

int launch_tiles_epilogue(vet_handle* h, const TilesPlan& p, int64_t F, double* entropy, double* per_k, int64_t per_k_stride,
                          double* hist0, cudaStream_t st) {
  vet::EntropyRowsArgs e{};
  e.F = F;
  e.K = h->K;
  for (int k = 0; k < h->K; ++k) {
    e.T[k] = h->ts[k].T;
    e.ioff[k] = p.A.hist_off[k];
  }
  e.ihist = h->d_ihist;
  e.istride = p.A.sumT;
  e.hist0_out = hist0;
  e.nvalid = h->d_nvalid;
  e.use_weight = 0;
  e.norm_always = h->norm_always;
  e.norm_T0 = h->norm_T0;
  e.entropy = entropy;
  e.per_k = per_k;
  e.per_k_stride = per_k_stride;
  e.flags = h->d_flags;
  LaunchTimer lt(h, VET_KERNEL_EPILOGUE, st);
  const int kw = h->K >= 8 ? 8 : (h->K >= 4 ? 4 : (h->K >= 2 ? 2 : 1));  // warps per frame
  const int groups = 8 / kw;
  const int blocks = (int)std::min<int64_t>((F + groups - 1) / groups, (int64_t)h->sm_count * 8);
  vet::k_entropy_rows<<<blocks, 256, 0, st>>>(e, kw);
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

// k_whist for tile count k over F frames of the cell histogram `cnt` -> hist[F,T_k]
int launch_whist(vet_handle* h, int k, int64_t F, const uint32_t* cnt, double* hist, cudaStream_t st) {
  TileSet& t = h->ts[k];
  const int shape = whist_shape();
  const int frames_per_cta = shape == 1 ? vet::WhistTall::kFramesPerCta : vet::WhistWide::kFramesPerCta;
  const size_t wh_smem = (size_t)vet::kWhStages * (shape == 1   ? vet::WhistTall::kChunkBytes
                                                   : shape == 2 ? vet::WhistQuad::kChunkBytes
                                                                : vet::WhistWide::kChunkBytes);
  const int64_t fblocks = (F + frames_per_cta - 1) / frames_per_cta;
  const int blocks = (int)std::min<int64_t>(fblocks * t.G, h->sm_count);
  if (t.sched_F != F || t.sched_blocks != blocks) {
    VET_CUDA(cudaStreamSynchronize(st));  // the previous schedule may still be in use
    if (int rc = build_whist_schedule(t, fblocks, blocks)) return rc;
    t.sched_F = F;
  }
  vet::WhistArgs a{};
  a.cnt = cnt;
  a.F = F;
  a.cpad = h->Cpad;
  a.T = t.T;
  a.G = t.G;
  a.group_tiles = t.d_group_tiles;
  a.group_chunk0 = t.d_group_chunk0;
  a.chunks = reinterpret_cast<const unsigned char*>(t.d_chunks);
  a.units = t.d_units;
  a.hist = hist;
  a.cta_items = t.d_sched;
  a.max_items = t.sched_max_items;
  {
    LaunchTimer lt(h, VET_KERNEL_EPILOGUE, st);
    if (shape == 1)
      vet::k_whist<vet::WhistTall><<<blocks, vet::kWhThreads, wh_smem, st>>>(a);
    else if (shape == 2)
      vet::k_whist<vet::WhistQuad><<<blocks, vet::kWhThreads, wh_smem, st>>>(a);
    else
      vet::k_whist<vet::WhistWide><<<blocks, vet::kWhThreads, wh_smem, st>>>(a);
  }
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}


// ---- int8 tensor-core weighted histogram (vet_whist_i8.cuh) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static const EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// tensor map of a row-major uint8 matrix [rows, kp] read in boxes of {128 bytes, box_rows} with the 128-byte swizzle
int make_u8_map(CUtensorMap* m, const void* base, uint64_t kp, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(VET_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
  const cuuint64_t dims[2] = {kp, rows};
  const cuuint64_t strides[1] = {kp};
  const cuuint32_t box[2] = {(cuuint32_t)vet::kI8BK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VET_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return VET_OK;
}

int64_t i8_kp(const vet_handle* h) { return (h->C + vet::kI8BK - 1) / vet::kI8BK * vet::kI8BK; }

// 0 = heuristic, 1 = always the FP64 kernel, 2 = the tensor-core kernel whenever it applies
int whist_impl() {
  const char* e = getenv("VET_WHIST_IMPL");
  if (e && std::string(e) == "fp64") return 1;
  if (e && std::string(e) == "i8") return 2;
  return 0;
}

bool use_whist_i8(const vet_handle* h, int64_t F, int64_t U) {
  const int impl = whist_impl();
  if (impl == 1 || !encode_tiled_fn()) return false;
  if (U * 255 >= ((int64_t)1 << 31)) return false;  // int32 accumulators: D <= 255 * sum(count plane) <= 255 U
  if ((size_t)vet::kI8SmemBytes + kStaticSmemSlack > h->smem_optin) return false;
  if (impl == 2) return true;
  // A CTA of the tensor-core kernel takes ~70 us whatever the batch (it walks all cells of 128 frames x 48
  // tiles); the FP64 kernel costs ~0.14 us per frame at 201 tiles.  From a few hundred frames on the
  // tensor cores win (measured: equal at 450 frames, 0.07 vs 0.50 ms at 3600).
  return F >= 512;
}

// Quantised weight slices of tile set t: row nb*240 + s*48 + j of [i8_blocks*240, kp] holds slice s of
// rint(w(cell, nb*48+j) * 2^39) for every cell; plus the K-block range of every N block.
int build_i8_tables(vet_handle* h, TileSet& t) {
  if (t.i8_built) return VET_OK;
  const int T = t.T;
  const int64_t kp = i8_kp(h);
  std::vector<uint32_t> col_ptr(T + 1), cell_idx(std::max<uint64_t>(t.nnz, 1));
  std::vector<double> w_val(std::max<uint64_t>(t.nnz, 1));
  VET_CUDA(cudaMemcpy(col_ptr.data(), t.d_col_ptr, (size_t)(T + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (t.nnz) {
    VET_CUDA(cudaMemcpy(cell_idx.data(), t.d_cell_idx, t.nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    VET_CUDA(cudaMemcpy(w_val.data(), t.d_w_val, t.nnz * sizeof(double), cudaMemcpyDeviceToHost));
  }
  const int nblk = (T + vet::kI8TilesPerBlock - 1) / vet::kI8TilesPerBlock;
  std::vector<uint8_t> w8((size_t)nblk * vet::kI8N * kp, 0);
  std::vector<int2> range(nblk);
  for (int nb = 0; nb < nblk; ++nb) {
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    for (int j = 0; j < vet::kI8TilesPerBlock; ++j) {
      const int tile = nb * vet::kI8TilesPerBlock + j;
      if (tile >= T) break;
      for (uint32_t e = col_ptr[tile]; e < col_ptr[tile + 1]; ++e) {
        const uint32_t c = cell_idx[e];
        const uint64_t q = (uint64_t)std::llrint(std::ldexp(w_val[e], vet::kI8FracBits));
        if (!q) continue;
        lo = std::min(lo, c);
        hi = std::max(hi, c);
        for (int s = 0; s < vet::kI8Slices; ++s)
          w8[((size_t)nb * vet::kI8N + (size_t)s * vet::kI8TilesPerBlock + j) * kp + c] = (uint8_t)(q >> (8 * s));
      }
    }
    if (lo > hi) lo = hi = 0;  // no weight at all: one K block of zeros
    range[nb] = make_int2((int)(lo / vet::kI8BK), (int)(hi / vet::kI8BK) + 1);
  }
  if (int rc = upload(&t.d_w8, w8.data(), w8.size())) return rc;
  if (int rc = upload(&t.d_kb_range, range.data(), range.size())) return rc;
  if (int rc = make_u8_map(&t.tm_w, t.d_w8, (uint64_t)kp, (uint64_t)nblk * vet::kI8N, vet::kI8N)) return rc;
  t.i8_blocks = nblk;
  t.i8_built = true;
  return VET_OK;
}

// Scratch of the tensor-core path for batches of up to F frames: byte planes, dirty marks, flags and the
// tensor map over the planes.  Planes 1 and 2 start out zero (the invariant the dirty marks protect).
int ensure_planes(vet_handle* h, int64_t F, cudaStream_t st) {
  const int64_t kp = i8_kp(h);
  const int64_t rows = (F + vet::kI8M - 1) / vet::kI8M * vet::kI8M;
  if (rows <= h->plane_rows) return VET_OK;
  VET_CUDA(cudaStreamSynchronize(st));
  cudaFree(h->d_planes);
  cudaFree(h->d_dirty);
  cudaFree(h->d_i8flags);
  h->d_planes = nullptr;
  h->d_dirty = nullptr;
  h->d_i8flags = nullptr;
  h->plane_rows = 0;
  VET_CUDA(cudaMalloc((void**)&h->d_planes, (size_t)3 * rows * kp));
  VET_CUDA(cudaMalloc((void**)&h->d_dirty, (size_t)rows));
  VET_CUDA(cudaMalloc((void**)&h->d_i8flags, (size_t)2 * (rows / vet::kI8M) * 4));
  VET_CUDA(cudaMemsetAsync(h->d_planes, 0, (size_t)3 * rows * kp, st));
  VET_CUDA(cudaMemsetAsync(h->d_dirty, 0, (size_t)rows, st));
  if (int rc = make_u8_map(&h->tm_cnt, h->d_planes, (uint64_t)kp, (uint64_t)3 * rows, vet::kI8M)) return rc;
  h->plane_rows = rows;
  return VET_OK;
}
uint32_t* i8_hi1(vet_handle* h) { return h->d_i8flags; }
uint32_t* i8_hi2(vet_handle* h) { return h->d_i8flags + h->plane_rows / vet::kI8M; }

// cell histogram rows -> byte planes, for batches whose frames the streaming kernel split into chunks
int launch_cnt_planes(vet_handle* h, int64_t F, const uint32_t* cnt, cudaStream_t st) {
  const int64_t kp = i8_kp(h);
  if (int rc = ensure_planes(h, F, st)) return rc;
  VET_CUDA(cudaMemsetAsync(h->d_i8flags, 0, (size_t)2 * (h->plane_rows / vet::kI8M) * 4, st));
  vet::CntPlanesArgs a{};
  a.cnt = cnt;
  a.F = F;
  a.cpad = h->Cpad;
  a.kp = (int)kp;
  a.plane_stride = h->plane_rows * kp;
  a.planes = h->d_planes;
  a.dirty = h->d_dirty;
  a.hi1 = i8_hi1(h);
  a.hi2 = i8_hi2(h);
  LaunchTimer lt(h, VET_KERNEL_EPILOGUE, st);
  vet::k_cnt_planes<<<h->sm_count * 8, 256, 0, st>>>(a);
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

// k_whist_i8 for tile count k over the planes of the current batch -> hist[F,T_k].  Pass 1: count bits
// [0,16) (plane 1 only for frame blocks flagged hi1); pass 2, CTAs of frame blocks flagged hi2 only:
// bits [16,24), added to the result.
int launch_whist_i8(vet_handle* h, int k, int64_t F, double* hist, cudaStream_t st) {
  TileSet& t = h->ts[k];
  if (int rc = build_i8_tables(h, t)) return rc;
  if (!h->i8_attr_set) {  // function attributes are per device: once per handle
    VET_CUDA(cudaFuncSetAttribute(vet::k_whist_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, vet::kI8SmemBytes));
    h->i8_attr_set = true;
  }
  const int fblocks = (int)((F + vet::kI8M - 1) / vet::kI8M);
  vet::WhistI8Args a{};
  a.F = F;
  a.T = t.T;
  a.n_blocks = t.i8_blocks;
  a.kb_range = t.d_kb_range;
  a.hist = hist;
  const int grid = fblocks * t.i8_blocks;
  for (int pass = 0; pass < 2; ++pass) {
    a.row_a = pass == 0 ? 0 : (int)(2 * h->plane_rows);
    a.row_b = (int)h->plane_rows;
    a.flag_b = pass == 0 ? i8_hi1(h) : nullptr;
    a.run_if = pass == 0 ? nullptr : i8_hi2(h);
    a.shift = pass == 0 ? 0 : 16;
    a.accumulate = pass;
    LaunchTimer lt(h, VET_KERNEL_EPILOGUE, st);
    vet::k_whist_i8<<<grid, vet::kI8Threads, vet::kI8SmemBytes, st>>>(h->tm_cnt, t.tm_w, a);
  }
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

// weighted histograms' scratch rows of tile count k (k == 0 may write straight into the caller's hist0)
int whist_rows(vet_handle* h, int k, int64_t F, double* hist0, double** out) {
  TileSet& t = h->ts[k];
  if (k == 0 && hist0) {
    *out = hist0;
    return VET_OK;
  }
  if (int rc = grow((void**)&t.d_hist, &t.hist_bytes, (size_t)F * t.T * 8)) return rc;
  *out = t.d_hist;
  return VET_OK;
}

int launch_weighted_rows(vet_handle* h, int64_t F, double* const* hists, const uint32_t* nvalid, double* entropy, double* per_k,
                         int64_t per_k_stride, cudaStream_t st) {
  vet::EntropyRowsArgs e{};
  e.F = F;
  e.K = h->K;
  for (int k = 0; k < h->K; ++k) {
    e.T[k] = h->ts[k].T;
    e.hist[k] = hists[k];
  }
  e.nvalid = nvalid;
  e.use_weight = 1;
  e.entropy = entropy;
  e.per_k = per_k;
  e.per_k_stride = per_k_stride;
  e.flags = h->d_flags;
  LaunchTimer lt(h, VET_KERNEL_EPILOGUE, st);
  const int kw = h->K >= 8 ? 8 : (h->K >= 4 ? 4 : (h->K >= 2 ? 2 : 1));  // warps per frame
  const int groups = 8 / kw;
  const int blocks = (int)std::min<int64_t>((F + groups - 1) / groups, (int64_t)h->sm_count * 8);
  vet::k_entropy_rows<<<blocks, 256, 0, st>>>(e, kw);
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

int launch_weighted_epilogue(vet_handle* h, int64_t F, int64_t U, double* entropy, double* per_k, int64_t per_k_stride,
                             double* hist0, cudaStream_t st) {
  double* hists[vet::kMaxTileCounts];
  const bool i8 = use_whist_i8(h, F, U);
  if (i8 && !h->planes_from_stream)
    if (int rc = launch_cnt_planes(h, F, h->d_cnt, st)) return rc;
  for (int k = 0; k < h->K; ++k) {
    if (int rc = whist_rows(h, k, F, hist0, &hists[k])) return rc;
    if (int rc = i8 ? launch_whist_i8(h, k, F, hists[k], st) : launch_whist(h, k, F, h->d_cnt, hists[k], st)) return rc;
  }
  return launch_weighted_rows(h, F, hists, h->d_nvalid, entropy, per_k, per_k_stride, st);
}

int launch_epilogue(vet_handle* h, int64_t F, int64_t U, double* entropy, double* per_k, int64_t per_k_stride, double* hist0,
                    cudaStream_t st) {
  if (h->use_weight) return launch_weighted_epilogue(h, F, U, entropy, per_k, per_k_stride, hist0, st);
  if (h->global_tables) {  // k_stream_global already made the tile histograms
    TilesPlan tp;
    int off = 0;
    for (int k = 0; k < h->K; ++k) {
      tp.A.hist_off[k] = off;
      off += h->ts[k].T;
    }
    tp.A.sumT = off;
    return launch_tiles_epilogue(h, tp, F, entropy, per_k, per_k_stride, hist0, st);
  }
  vet::EpilogueArgs a{};
  a.cnt = h->d_cnt;
  a.F = F;
  a.C = (int)h->C;
  a.cpad = h->Cpad;
  a.K = h->K;
  a.use_weight = h->use_weight;
  a.norm_always = h->norm_always;
  a.norm_T0 = h->norm_T0;
  for (int k = 0; k < h->K; ++k) {
    a.ts[k].T = h->ts[k].T;
    a.ts[k].lut = h->ts[k].d_lut;
    a.ts[k].col_ptr = h->ts[k].d_col_ptr;
    a.ts[k].cell_idx = h->ts[k].d_cell_idx;
    a.ts[k].w_val = h->ts[k].d_w_val;
  }
  a.entropy = entropy;
  a.per_k = per_k;
  a.per_k_stride = per_k_stride;
  a.hist0 = hist0;
  a.flags = h->d_flags;
  const int blocks = (int)std::min<int64_t>(F, (int64_t)h->sm_count * 2);
  {
    LaunchTimer lt(h, VET_KERNEL_EPILOGUE, st);
    vet::k_epilogue<<<blocks, 512, epilogue_smem_bytes(h), st>>>(a, h->maxT);
  }
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

}  // namespace

// ================================ C ABI ==========================================

extern "C" const char* vet_last_error(void) { return g_err.c_str(); }
extern "C" const char* vet_version(void) { return "vet_b200 0.1 (sm_100a)"; }

extern "C" int vet_create(vet_handle** out, const vet_config* cfg) {
  if (!out || !cfg) return fail(VET_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  // CFG:62-67
  if (cfg->video_width <= 0 || cfg->video_height <= 0) return fail(VET_ERR_INVALID_ARG, "Video dimensions must be positive");
  const bool naive = cfg->naive_tile_width != 0 || cfg->naive_tile_height != 0;
  if (naive) {
    // EU:410-417 (negative sizes -- the -1 placeholders of CFG:104-105 -- are rejected here)
    if (cfg->naive_tile_width <= 0 || cfg->naive_tile_height <= 0) return fail(VET_ERR_INVALID_ARG, "No tile dimensions provided");
    if (180 % cfg->naive_tile_height != 0) return fail(VET_ERR_INVALID_ARG, "Tile height must divide 180!");
    if (360 % cfg->naive_tile_width != 0) return fail(VET_ERR_INVALID_ARG, "Tile width must divide 360!");
  } else {
    if (cfg->num_tile_counts <= 0 || !cfg->tile_counts) return fail(VET_ERR_INVALID_ARG, "Must specify at least one tile count");
    for (int k = 0; k < cfg->num_tile_counts; ++k)
      if (cfg->tile_counts[k] <= 0) return fail(VET_ERR_INVALID_ARG, "Tile counts must be positive");
  }
  // DU:239
  if (cfg->video_width % 2 || cfg->video_height % 2) return fail(VET_ERR_INVALID_ARG, "Video dimensions must be even numbers");
  // EU:35-38
  if (!naive && !(cfg->fov_angle > 0 && cfg->fov_angle <= 360)) return fail(VET_ERR_INVALID_ARG, "FOV angle must be between 0 and 360 degrees");
  if (!naive && !(cfg->power_factor > 0)) return fail(VET_ERR_INVALID_ARG, "Power factor must be positive");
  if (!naive && cfg->num_tile_counts > vet::kMaxTileCounts)
    return fail(VET_ERR_UNSUPPORTED, "at most %d tile counts per handle", vet::kMaxTileCounts);

  int ndev = 0;
  VET_CUDA(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(VET_ERR_INVALID_ARG, "no such CUDA device %d", cfg->device);
  DeviceGuard guard(cfg->device);
  if (!guard.ok) return fail(VET_ERR_CUDA, "cudaSetDevice(%d) failed", cfg->device);

  vet_handle* h = new (std::nothrow) vet_handle();
  if (!h) return fail(VET_ERR_NOMEM, "out of host memory");
  struct Cleanup {
    vet_handle* h;
    bool armed = true;
    ~Cleanup() {
      if (armed) vet_destroy(h);
    }
  } cleanup{h};

  h->device = cfg->device;
  h->W = cfg->video_width;
  h->H = cfg->video_height;
  h->C = (int64_t)(h->W + 1) * (h->H + 1);
  h->Cpad = (int)((h->C + 3) & ~(int64_t)3);
  h->K = naive ? 1 : cfg->num_tile_counts;
  h->fov = naive ? 120.0 : cfg->fov_angle;
  h->pf = naive ? 2.0 : cfg->power_factor;
  h->use_weight = (!naive && cfg->use_weight_distribution) ? 1 : 0;
  h->naive = naive;
  if (naive) {
    h->naive_w = cfg->naive_tile_width;
    h->naive_h = cfg->naive_tile_height;
    h->norm_T0 = (180 / h->naive_h) * (360 / h->naive_w);
    h->norm_always = cfg->use_weight_distribution ? 1 : 0;
  }
  h->max_d = np_radians(h->fov / 2.0);  // EU:124
  cudaDeviceProp prop;
  VET_CUDA(cudaGetDeviceProperties(&prop, h->device));
  h->sm_count = prop.multiProcessorCount;
  h->smem_optin = prop.sharedMemPerBlockOptin;

  h->ts.resize(h->K);
  for (int k = 0; k < h->K; ++k) {
    TileSet& t = h->ts[k];
    if (naive) {
      t.n = 0;
      t.T = (360 / h->naive_w + 1) * (180 / h->naive_h + 1);  // grid codes incl. the closed upper edges
      if (t.T > kMaxT) return fail(VET_ERR_UNSUPPORTED, "%dx%d degree tiles give %d grid codes; at most %d supported", h->naive_w, h->naive_h, t.T, kMaxT);
      h->maxT = t.T;
      h->sumT = t.T;
      continue;
    }
    t.n = cfg->tile_counts[k];
    t.T = 2 * (t.n / 2) + 1;  // DU:43-45
    if (cfg->num_tiles) {
      if (!cfg->centres || !cfg->centres[k] || cfg->num_tiles[k] <= 0)
        return fail(VET_ERR_INVALID_ARG, "No tile centers provided");  // EU:170-171
      t.T = cfg->num_tiles[k];
    }
    if (t.T > kMaxT) return fail(VET_ERR_UNSUPPORTED, "tile_count %d gives %d tiles; at most %d supported", t.n, t.T, kMaxT);
    h->maxT = std::max(h->maxT, t.T);
    h->sumT += t.T;
    if (cfg->centres && cfg->centres[k])
      t.h_centres.assign(cfg->centres[k], cfg->centres[k] + (size_t)t.T * 3);
    else
      t.h_centres = make_lattice(t.n);
  }
  // table regime: the per-frame cell histogram (u32) and the LUT must fit in shared memory;
  // larger videos use the direct per-sample path (decode -> vectors)
  h->direct_only = stream_smem_bytes(h) + kStaticSmemSlack > h->smem_optin ||
                   epilogue_smem_bytes(h) + kStaticSmemSlack > h->smem_optin;
  // weighted handles need per-cell weight tables (C x T): bounded; unweighted ones only the cell -> tile LUTs
  if (h->direct_only && (h->C <= kGlobalTableCells || (!h->use_weight && h->C <= kGlobalLutCells))) {
    const char* e = getenv("VET_REGIME");  // "direct" pins the per-sample path for A/B runs and tests
    if (!(e && std::string(e) == "direct")) {
      h->direct_only = false;
      h->global_tables = true;
    }
  }
  if (h->C >= ((int64_t)1 << 31)) return fail(VET_ERR_UNSUPPORTED, "video %dx%d has too many cells", h->W, h->H);
  if (naive && h->direct_only) return fail(VET_ERR_UNSUPPORTED, "video %dx%d is too large for the grid-tiling tables", h->W, h->H);

  std::vector<double> lon, lat;
  if (cfg->lon_by_px && cfg->lat_by_py) {
    lon.assign(cfg->lon_by_px, cfg->lon_by_px + h->W + 1);
    lat.assign(cfg->lat_by_py, cfg->lat_by_py + h->H + 1);
  } else {
    make_axis_tables(h->W, h->H, lon, lat);
  }
  for (double v : lon)
    if (!(v >= -180 && v <= 180)) return fail(VET_ERR_INVALID_ARG, "Longitude must be between -180 and 180 degrees");  // DT:80-81
  for (double v : lat)
    if (!(v >= -90 && v <= 90)) return fail(VET_ERR_INVALID_ARG, "Latitude must be between -90 and 90 degrees");  // DT:82-83
  std::vector<double> cosT(h->W + 1), sinT(h->W + 1), sinP(h->H + 1), cosP(h->H + 1);
  for (int px = 0; px <= h->W; ++px) {
    const double th = np_radians(lon[px]);  // DT:204
    cosT[px] = std::cos(th);
    sinT[px] = std::sin(th);
  }
  for (int py = 0; py <= h->H; ++py) {
    const double ph = np_radians(90 - lat[py]);  // DT:205
    sinP[py] = std::sin(ph);
    cosP[py] = std::cos(ph);
  }
  if (int rc = upload(&h->d_cosT, cosT.data(), cosT.size())) return rc;
  if (int rc = upload(&h->d_sinT, sinT.data(), sinT.size())) return rc;
  if (int rc = upload(&h->d_sinP, sinP.data(), sinP.size())) return rc;
  if (int rc = upload(&h->d_cosP, cosP.data(), cosP.size())) return rc;
  VET_CUDA(cudaMalloc((void**)&h->d_flags, sizeof(uint32_t)));
  VET_CUDA(cudaMemset(h->d_flags, 0, sizeof(uint32_t)));
  VET_CUDA(cudaMalloc((void**)&h->d_work, sizeof(uint32_t) * vet::kMaxTileCounts));
  {
    std::vector<uint16_t> ident(h->maxT);
    for (int i = 0; i < h->maxT; ++i) ident[i] = (uint16_t)i;
    if (int rc = upload(&h->d_identity, ident.data(), ident.size())) return rc;
  }
  if (h->direct_only) {
    for (int k = 0; k < h->K; ++k)
      if (int rc = build_unit_centres(h->ts[k])) return rc;
  } else {
    VET_CUDA(cudaMalloc((void**)&h->d_cellvec, (size_t)h->C * 3 * sizeof(double)));
    VET_CUDA(cudaFuncSetAttribute(vet::k_whist<vet::WhistWide>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  vet::kWhStages * vet::WhistWide::kChunkBytes));
    VET_CUDA(cudaFuncSetAttribute(vet::k_whist<vet::WhistTall>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  vet::kWhStages * vet::WhistTall::kChunkBytes));
    VET_CUDA(cudaFuncSetAttribute(vet::k_whist<vet::WhistQuad>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  vet::kWhStages * vet::WhistQuad::kChunkBytes));
    vet::k_cell_vectors<<<std::min<int64_t>((h->C + 255) / 256, 1024), 256>>>(h->d_cosT, h->d_sinT, h->d_sinP, h->d_cosP,
                                                                               h->W, h->H, h->d_cellvec);
    h->launches++;
    VET_CUDA(cudaGetLastError());
    for (int k = 0; k < h->K; ++k)
      if (int rc = naive ? build_naive_tile_set(h, h->ts[k], lon, lat) : build_tile_set(h, h->ts[k])) return rc;
    if (h->K <= 4 && h->maxT <= 255) {
      std::vector<uint32_t> packed_lut(h->C + 4, 0);
      for (int k = 0; k < h->K; ++k)
        for (int64_t c = 0; c < h->C; ++c) packed_lut[c] |= (uint32_t)h->ts[k].h_lut[c] << (8 * k);
      if (int rc = upload(&h->d_lut_packed, packed_lut.data(), packed_lut.size())) return rc;
    }
    // The attribute is per function, not per handle: always allow the device maximum so that
    // handles of different configurations can coexist.
    const size_t sm = h->smem_optin - kStaticSmemSlack;
    VET_CUDA(cudaFuncSetAttribute(vet::k_stream_simple<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    VET_CUDA(cudaFuncSetAttribute(vet::k_stream_simple<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    VET_CUDA(cudaFuncSetAttribute(vet::k_epilogue, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    {
#define VET_SMEM_ATTR(...) VET_CUDA(cudaFuncSetAttribute(vet::k_stream_tma<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm))
      VET_SMEM_ATTR(float, uint8_t, true, 0);
      VET_SMEM_ATTR(float, uint16_t, true, 0);
      VET_SMEM_ATTR(double, uint8_t, true, 0);
      VET_SMEM_ATTR(double, uint16_t, true, 0);
      VET_SMEM_ATTR(float, uint8_t, true, 1);
      VET_SMEM_ATTR(float, uint16_t, true, 1);
      VET_SMEM_ATTR(double, uint8_t, true, 1);
      VET_SMEM_ATTR(double, uint16_t, true, 1);
      VET_SMEM_ATTR(float, uint8_t, true, 2);
      VET_SMEM_ATTR(float, uint16_t, true, 2);
      VET_SMEM_ATTR(double, uint8_t, true, 2);
      VET_SMEM_ATTR(double, uint16_t, true, 2);
      VET_SMEM_ATTR(float, uint8_t, false, 0);
      VET_SMEM_ATTR(float, uint8_t, false, 1);
      VET_SMEM_ATTR(float, uint8_t, false, 2);
      VET_SMEM_ATTR(double, uint8_t, false, 0);
      VET_SMEM_ATTR(double, uint8_t, false, 1);
      VET_SMEM_ATTR(double, uint8_t, false, 2);
#undef VET_SMEM_ATTR
    }
  }
  VET_CUDA(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
  VET_CUDA(cudaStreamCreateWithFlags(&h->s_exec, cudaStreamNonBlocking));
  VET_CUDA(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  VET_CUDA(cudaDeviceSynchronize());
  cleanup.armed = false;
  *out = h;
  return VET_OK;
}

extern "C" int vet_destroy(vet_handle* h) {
  if (!h) return VET_OK;
  DeviceGuard guard(h->device);
  for (auto& t : h->ts) free_tile_set(t);
  for (auto& s : h->spans) {
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  cudaFree(h->d_cosT);
  cudaFree(h->d_sinT);
  cudaFree(h->d_sinP);
  cudaFree(h->d_cosP);
  cudaFree(h->d_cellvec);
  cudaFree(h->d_flags);
  cudaFree(h->d_cnt);
  cudaFree(h->d_nvalid);
  cudaFree(h->d_work);
  cudaFree(h->d_cells);
  cudaFree(h->d_identity);
  cudaFree(h->d_ihist);
  cudaFree(h->d_lut_packed);
  for (void* p : h->d_vscratch) cudaFree(p);
  cudaFree(h->d_tables);
  cudaFree(h->d_pairs);
  cudaFree(h->d_redo);
  cudaFree(h->d_trk);
  cudaFree(h->d_planes);
  cudaFree(h->d_dirty);
  cudaFree(h->d_i8flags);
  cudaFree(h->d_in[0]);
  cudaFree(h->d_in[1]);
  for (void* p : h->d_hout) cudaFree(p);
  if (h->s_copy) cudaStreamDestroy(h->s_copy);
  if (h->s_exec) cudaStreamDestroy(h->s_exec);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  delete h;
  return VET_OK;
}

extern "C" int vet_num_tiles(const vet_handle* h, int k) {
  if (!h || k < 0 || k >= h->K) return fail(VET_ERR_INVALID_ARG, "bad tile-count index");
  return h->ts[k].T;
}
extern "C" int64_t vet_num_cells(const vet_handle* h) { return h ? h->C : fail(VET_ERR_INVALID_ARG, "null handle"); }
extern "C" int64_t vet_launch_count(const vet_handle* h) { return h ? h->launches : 0; }

extern "C" int vet_lattice(const vet_handle* h, int k, double* centres_host) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_lattice: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || !centres_host || k < 0 || k >= h->K) return fail(VET_ERR_INVALID_ARG, "bad argument");
  std::memcpy(centres_host, h->ts[k].h_centres.data(), h->ts[k].h_centres.size() * sizeof(double));
  return VET_OK;
}

extern "C" int vet_cell_lut(const vet_handle* h, int k, uint16_t* lut_host) {
  if (!h || !lut_host || k < 0 || k >= h->K) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (h->direct_only) return fail(VET_ERR_UNSUPPORTED, "no cell tables for a %dx%d video (direct per-sample mode)", h->W, h->H);
  std::memcpy(lut_host, h->ts[k].h_lut.data(), h->ts[k].h_lut.size() * sizeof(uint16_t));
  return VET_OK;
}

extern "C" int vet_decode(vet_handle* h, const void* packed_dev, int dtype, int64_t n, double* vec_dev, int32_t* cell_dev,
                          void* stream) {
  if (!h || (!packed_dev && n > 0) || n < 0 || (dtype != VET_F32 && dtype != VET_F64))
    return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (n == 0) return VET_OK;
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16);
  if (dtype == VET_F32)
    vet::k_decode<float><<<blocks, 256, 0, st>>>((const float*)packed_dev, n, h->W, h->H, h->d_cosT, h->d_sinT, h->d_sinP,
                                                 h->d_cosP, vec_dev, cell_dev, h->d_flags);
  else
    vet::k_decode<double><<<blocks, 256, 0, st>>>((const double*)packed_dev, n, h->W, h->H, h->d_cosT, h->d_sinT,
                                                  h->d_sinP, h->d_cosP, vec_dev, cell_dev, h->d_flags);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

extern "C" int vet_nearest_tile(vet_handle* h, int k, const double* vec_dev, int64_t n, int32_t* idx_dev, void* stream) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_nearest_tile: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || k < 0 || k >= h->K || n < 0 || (n > 0 && (!vec_dev || !idx_dev))) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (n == 0) return VET_OK;
  DeviceGuard guard(h->device);
  const TileSet& t = h->ts[k];
  const size_t smem = (size_t)t.T * 3 * sizeof(double);
  VET_CUDA(cudaFuncSetAttribute(vet::k_nearest<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = 256;
  const int64_t rounds = (n + threads / 4 - 1) / (threads / 4);
  const int blocks = (int)std::min<int64_t>(rounds, (int64_t)h->sm_count * 8);
  vet::k_nearest<int32_t><<<blocks, threads, smem, (cudaStream_t)stream>>>(vec_dev, n, t.d_unit, t.T, idx_dev);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

extern "C" int vet_tile_weights(vet_handle* h, int k, const double* vec_dev, int64_t n, double* w_dev, void* stream) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_tile_weights: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || k < 0 || k >= h->K || n < 0 || (n > 0 && (!vec_dev || !w_dev))) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (n == 0) return VET_OK;
  DeviceGuard guard(h->device);
  const TileSet& t = h->ts[k];
  const size_t smem = (size_t)t.T * 3 * sizeof(double);
  VET_CUDA(cudaFuncSetAttribute(vet::k_tile_weights, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = (int)std::min<int64_t>((n + 7) / 8, (int64_t)h->sm_count * 8);
  vet::k_tile_weights<<<blocks, 256, smem, (cudaStream_t)stream>>>(vec_dev, n, t.d_unit, t.T, h->max_d, h->pf, h->use_weight,
                                                                   w_dev);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

extern "C" int vet_spatial(vet_handle* h, const void* packed_dev, int dtype, int64_t F, int64_t U, double* entropy_dev,
                           double* per_k_dev, double* hist0_dev, uint16_t* assign0_dev, void* stream) {
  if (!h || F < 0 || U < 0 || (dtype != VET_F32 && dtype != VET_F64)) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (F == 0) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty vector dictionary");  // EU:168-169
  if (!packed_dev || !entropy_dev) return fail(VET_ERR_INVALID_ARG, "null buffer");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->direct_only) return spatial_direct(h, packed_dev, dtype, F, U, entropy_dev, per_k_dev, hist0_dev, assign0_dev, st);
  const int64_t fb = frames_per_batch(h, F, U, false);
  if (int rc = grow((void**)&h->d_cnt, &h->cnt_bytes, cnt_scratch_bytes(h, fb))) return rc;
  if (int rc = grow((void**)&h->d_nvalid, &h->nvalid_bytes, (size_t)fb * 4)) return rc;
  const size_t esz = dtype == VET_F32 ? 4 : 8;
  const int T0 = h->ts[0].T;
  for (int64_t f0 = 0; f0 < F; f0 += fb) {
    const int64_t nf = std::min(fb, F - f0);
    const char* in = (const char*)packed_dev + (size_t)f0 * U * 3 * esz;
    TilesPlan tp = plan_tiles(h, in, U);
    if (tp.ok) {
      if (int rc = launch_stream_tiles(h, tp, in, dtype, nf, U, assign0_dev ? assign0_dev + f0 * U : nullptr, st)) return rc;
      if (int rc = launch_tiles_epilogue(h, tp, nf, entropy_dev + f0, per_k_dev ? per_k_dev + f0 : nullptr, F,
                                         hist0_dev ? hist0_dev + f0 * T0 : nullptr, st))
        return rc;
      continue;
    }
    if (int rc = launch_stream(h, in, dtype, nf, U, assign0_dev ? assign0_dev + f0 * U : nullptr, false, st)) return rc;
    if (int rc = launch_epilogue(h, nf, U, entropy_dev + f0, per_k_dev ? per_k_dev + f0 : nullptr, F,
                                 hist0_dev ? hist0_dev + f0 * T0 : nullptr, st))
      return rc;
  }
  return VET_OK;
}

extern "C" int vet_transition(vet_handle* h, const void* packed_dev, int dtype, int64_t F, int64_t U, double* entropy_dev,
                              double* per_k_dev, int32_t* prev_count0_dev, uint16_t* pairs0_dev, int mode, void* stream) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_transition: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || F < 0 || U < 0 || (dtype != VET_F32 && dtype != VET_F64)) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (mode != VET_TRANSITION_LITERAL && mode != VET_TRANSITION_TEXTBOOK) return fail(VET_ERR_INVALID_ARG, "bad mode");
  if (F <= 1) return VET_OK;  // TA:143-146: the first frame yields no row
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty vector dictionary");  // EU:239-240
  if (!packed_dev || !entropy_dev) return fail(VET_ERR_INVALID_ARG, "null buffer");
  if (U >= 0xFFFFFFFFll) return fail(VET_ERR_UNSUPPORTED, "too many users");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->direct_only)
    return transition_direct(h, packed_dev, dtype, F, U, entropy_dev, per_k_dev, prev_count0_dev, pairs0_dev, mode, st);
  const int64_t fb = std::max<int64_t>(2, frames_per_batch(h, F, U, true));
  const size_t csz = h->C <= 65535 ? 2 : 4;
  if (int rc = grow((void**)&h->d_cnt, &h->cnt_bytes, cnt_scratch_bytes(h, fb))) return rc;
  if (int rc = grow((void**)&h->d_nvalid, &h->nvalid_bytes, (size_t)fb * 4)) return rc;
  if (int rc = grow(&h->d_cells, &h->cells_bytes, (size_t)fb * U * csz)) return rc;
  const size_t esz = dtype == VET_F32 ? 4 : 8;
  const int T0 = h->ts[0].T;
  // batches overlap by one frame (the halo frame of SURVEY 8e)
  for (int64_t f0 = 0; f0 < F - 1; f0 += fb - 1) {
    const int64_t nf = std::min(fb, F - f0);
    const char* in = (const char*)packed_dev + (size_t)f0 * U * 3 * esz;
    // one tile count: the streaming kernel writes the tile ids themselves (its own LUT lookup) into the
    // scratch and the transition kernels take them through the identity table -- no lookups per pair
    const bool tiles_direct = h->K == 1;
    if (int rc = launch_stream(h, in, dtype, nf, U, tiles_direct ? (uint16_t*)h->d_cells : nullptr, !tiles_direct, st)) return rc;
    vet::TransitionArgs a{};
    a.cell16 = (csz == 2 || tiles_direct) ? (const uint16_t*)h->d_cells : nullptr;
    a.cell32 = (csz == 4 && !tiles_direct) ? (const int32_t*)h->d_cells : nullptr;
    a.F = nf;
    a.U = U;
    a.K = h->K;
    for (int k = 0; k < h->K; ++k) {
      a.T[k] = h->ts[k].T;
      a.lut[k] = tiles_direct ? h->d_identity : h->ts[k].d_lut;
    }
    a.entropy = entropy_dev + f0;
    a.per_k = per_k_dev ? per_k_dev + f0 : nullptr;
    a.per_k_stride = F - 1;
    a.prev_count0 = prev_count0_dev ? prev_count0_dev + f0 * T0 : nullptr;
    a.pairs0 = pairs0_dev ? pairs0_dev + f0 * U * 2 : nullptr;
    a.mode = mode;
    a.flags = h->d_flags;
    if (int rc = launch_transition(h, a, nf - 1, U, h->maxT, st)) return rc;
    if (nf == F - f0) break;
  }
  return VET_OK;
}

extern "C" int vet_analyze(vet_handle* h, const void* packed_dev, int dtype, int64_t F, int64_t U, double* sp_entropy_dev,
                           double* sp_per_k_dev, double* hist0_dev, uint16_t* assign0_dev, double* tr_entropy_dev,
                           double* tr_per_k_dev, int32_t* prev_count0_dev, uint16_t* pairs0_dev, int mode, void* stream) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_analyze: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || F < 0 || U < 0 || (dtype != VET_F32 && dtype != VET_F64)) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (mode != VET_TRANSITION_LITERAL && mode != VET_TRANSITION_TEXTBOOK) return fail(VET_ERR_INVALID_ARG, "bad mode");
  if (F == 0) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty vector dictionary");
  if (!packed_dev || !sp_entropy_dev || (F > 1 && !tr_entropy_dev)) return fail(VET_ERR_INVALID_ARG, "null buffer");
  if (U >= 0xFFFFFFFFll) return fail(VET_ERR_UNSUPPORTED, "too many users");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->direct_only || F == 1) {  // no shared pass to gain: run the two stages one after the other
    if (int rc = vet_spatial(h, packed_dev, dtype, F, U, sp_entropy_dev, sp_per_k_dev, hist0_dev, assign0_dev, stream)) return rc;
    return vet_transition(h, packed_dev, dtype, F, U, tr_entropy_dev, tr_per_k_dev, prev_count0_dev, pairs0_dev, mode, stream);
  }
  const int64_t fb = std::max<int64_t>(2, frames_per_batch(h, F, U, true));
  const size_t csz = h->C <= 65535 ? 2 : 4;
  if (int rc = grow((void**)&h->d_cnt, &h->cnt_bytes, cnt_scratch_bytes(h, fb))) return rc;
  if (int rc = grow((void**)&h->d_nvalid, &h->nvalid_bytes, (size_t)fb * 4)) return rc;
  if (int rc = grow(&h->d_cells, &h->cells_bytes, (size_t)fb * U * csz)) return rc;
  // the streaming kernel always writes assignments here (its LUT copy is what selects the fused variant)
  uint16_t* assign = assign0_dev;
  if (!assign) {
    if (int rc = grow(&h->d_vscratch[0], &h->vscratch_bytes[0], (size_t)fb * U * 2)) return rc;
  }
  const size_t esz = dtype == VET_F32 ? 4 : 8;
  const int T0 = h->ts[0].T;
  for (int64_t f0 = 0; f0 < F; f0 += fb - 1) {  // batches overlap by the halo frame of the transition stage
    const int64_t nf = std::min(fb, F - f0);
    const char* in = (const char*)packed_dev + (size_t)f0 * U * 3 * esz;
    uint16_t* asg = assign ? assign + f0 * U : (uint16_t*)h->d_vscratch[0];
    const bool tiles_direct = h->K == 1;  // the assignments double as the transition stage's input (identity table)
    if (int rc = launch_stream(h, in, dtype, nf, U, asg, !tiles_direct, st)) return rc;
    if (int rc = launch_epilogue(h, nf, U, sp_entropy_dev + f0, sp_per_k_dev ? sp_per_k_dev + f0 : nullptr, F,
                                 hist0_dev ? hist0_dev + f0 * T0 : nullptr, st))
      return rc;
    if (nf >= 2) {
      vet::TransitionArgs a{};
      a.cell16 = tiles_direct ? asg : (csz == 2 ? (const uint16_t*)h->d_cells : nullptr);
      a.cell32 = (csz == 4 && !tiles_direct) ? (const int32_t*)h->d_cells : nullptr;
      a.F = nf;
      a.U = U;
      a.K = h->K;
      for (int k = 0; k < h->K; ++k) {
        a.T[k] = h->ts[k].T;
        a.lut[k] = tiles_direct ? h->d_identity : h->ts[k].d_lut;
      }
      a.entropy = tr_entropy_dev + f0;
      a.per_k = tr_per_k_dev ? tr_per_k_dev + f0 : nullptr;
      a.per_k_stride = F - 1;
      a.prev_count0 = prev_count0_dev ? prev_count0_dev + f0 * T0 : nullptr;
      a.pairs0 = pairs0_dev ? pairs0_dev + f0 * U * 2 : nullptr;
      a.mode = mode;
      a.flags = h->d_flags;
      if (int rc = launch_transition(h, a, nf - 1, U, h->maxT, st)) return rc;
    }
    if (nf == F - f0) break;
  }
  return VET_OK;
}

namespace {

int launch_nearest_i32(vet_handle* h, int k, const double* vec, int64_t n, int32_t* idx, cudaStream_t st) {
  const TileSet& t = h->ts[k];
  const size_t smem = (size_t)t.T * 3 * sizeof(double);
  VET_CUDA(cudaFuncSetAttribute(vet::k_nearest<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = 256;
  const int64_t rounds = (n + threads / 4 - 1) / (threads / 4);
  const int blocks = (int)std::min<int64_t>(rounds, (int64_t)h->sm_count * 8);
  vet::k_nearest<int32_t><<<blocks, threads, smem, st>>>(vec, n, t.d_unit, t.T, idx);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

int spatial_vectors_impl(vet_handle* h, const double* vec, int64_t F, int64_t U, double* entropy, double* per_k,
                         int64_t per_k_stride, double* hist0, uint16_t* assign0, cudaStream_t st) {
  const int64_t n = F * U;
  const bool need_idx = !h->use_weight || assign0;
  if (need_idx)
    if (int rc = grow(&h->d_vscratch[0], &h->vscratch_bytes[0], (size_t)n * 4)) return rc;
  double* pk = per_k;
  int64_t pk_stride = per_k_stride;
  if (!pk) {
    if (int rc = grow(&h->d_vscratch[1], &h->vscratch_bytes[1], (size_t)h->K * F * 8)) return rc;
    pk = (double*)h->d_vscratch[1];
    pk_stride = F;
  }
  int32_t* idx = (int32_t*)h->d_vscratch[0];
  for (int k = 0; k < h->K; ++k) {
    const TileSet& t = h->ts[k];
    if (!h->use_weight || (k == 0 && assign0)) {
      if (int rc = launch_nearest_i32(h, k, vec, n, idx, st)) return rc;
      if (k == 0 && assign0) {
        vet::k_idx_to_u16<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 8), 256, 0, st>>>(idx, n, assign0);
        h->launches++;
      }
    }
    vet::VecSpatialArgs a{};
    a.vec = vec;
    a.F = F;
    a.U = U;
    a.unit = t.d_unit;
    a.T = t.T;
    a.max_d = h->max_d;
    a.pf = h->pf;
    a.use_weight = h->use_weight;
    a.idx = idx;
    a.per_k = pk + k * pk_stride;
    a.hist = (k == 0) ? hist0 : nullptr;
    a.flags = h->d_flags;
    const size_t smem = (size_t)t.T * 12 + (size_t)vet::kVecChunk * 24 + 16;
    VET_CUDA(cudaFuncSetAttribute(vet::k_spatial_vectors, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vet::k_spatial_vectors<<<(int)std::min<int64_t>(F, (int64_t)h->sm_count * 4), 256, smem, st>>>(a);
    h->launches++;
    VET_CUDA(cudaGetLastError());
  }
  vet::k_average_rows<<<(int)std::min<int64_t>((F + 255) / 256, 1024), 256, 0, st>>>(pk, h->K, F, pk_stride, entropy);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

int transition_vectors_impl(vet_handle* h, const double* vec, int64_t F, int64_t U, double* entropy, double* per_k,
                            int64_t per_k_stride, int32_t* prev_count0, uint16_t* pairs0, int mode, cudaStream_t st) {
  const int64_t n = F * U;
  if (int rc = grow(&h->d_vscratch[0], &h->vscratch_bytes[0], (size_t)n * 4)) return rc;
  double* pk = per_k;
  int64_t pk_stride = per_k_stride;
  if (!pk) {
    if (int rc = grow(&h->d_vscratch[1], &h->vscratch_bytes[1], (size_t)h->K * (F - 1) * 8)) return rc;
    pk = (double*)h->d_vscratch[1];
    pk_stride = F - 1;
  }
  int32_t* idx = (int32_t*)h->d_vscratch[0];
  for (int k = 0; k < h->K; ++k) {
    if (int rc = launch_nearest_i32(h, k, vec, n, idx, st)) return rc;
    vet::TransitionArgs a{};
    a.cell32 = idx;  // tile indices play the role of cell ids, mapped through the identity LUT
    a.F = F;
    a.U = U;
    a.K = 1;
    a.T[0] = h->ts[k].T;
    a.lut[0] = h->d_identity;
    a.entropy = pk + k * pk_stride;  // K == 1: the "mean" is the tile count's own entropy
    a.per_k = nullptr;
    a.prev_count0 = (k == 0) ? prev_count0 : nullptr;
    a.pairs0 = (k == 0) ? pairs0 : nullptr;
    a.mode = mode;
    a.flags = h->d_flags;
    if (int rc = launch_transition(h, a, F - 1, U, h->ts[k].T, st)) return rc;
  }
  vet::k_average_rows<<<(int)std::min<int64_t>((F - 1 + 255) / 256, 1024), 256, 0, st>>>(pk, h->K, F - 1, pk_stride, entropy);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

// decode a frame batch of packed samples into the vector scratch (direct mode)
int decode_batch(vet_handle* h, const void* packed, int dtype, int64_t n, cudaStream_t st) {
  if (int rc = grow(&h->d_vscratch[2], &h->vscratch_bytes[2], (size_t)n * 24)) return rc;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16);
  if (dtype == VET_F32)
    vet::k_decode<float><<<blocks, 256, 0, st>>>((const float*)packed, n, h->W, h->H, h->d_cosT, h->d_sinT, h->d_sinP,
                                                 h->d_cosP, (double*)h->d_vscratch[2], nullptr, h->d_flags);
  else
    vet::k_decode<double><<<blocks, 256, 0, st>>>((const double*)packed, n, h->W, h->H, h->d_cosT, h->d_sinT, h->d_sinP,
                                                  h->d_cosP, (double*)h->d_vscratch[2], nullptr, h->d_flags);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

int spatial_direct(vet_handle* h, const void* packed, int dtype, int64_t F, int64_t U, double* entropy, double* per_k,
                   double* hist0, uint16_t* assign0, cudaStream_t st) {
  const size_t esz = dtype == VET_F32 ? 4 : 8;
  const int T0 = h->ts[0].T;
  const int64_t fb = std::min<int64_t>(F, std::max<int64_t>(1, (int64_t)(((size_t)1 << 30) / ((size_t)U * 24))));
  for (int64_t f0 = 0; f0 < F; f0 += fb) {
    const int64_t nf = std::min(fb, F - f0);
    if (int rc = decode_batch(h, (const char*)packed + (size_t)f0 * U * 3 * esz, dtype, nf * U, st)) return rc;
    if (int rc = spatial_vectors_impl(h, (const double*)h->d_vscratch[2], nf, U, entropy + f0, per_k ? per_k + f0 : nullptr, F,
                                      hist0 ? hist0 + f0 * T0 : nullptr, assign0 ? assign0 + f0 * U : nullptr, st))
      return rc;
  }
  return VET_OK;
}

int transition_direct(vet_handle* h, const void* packed, int dtype, int64_t F, int64_t U, double* entropy, double* per_k,
                      int32_t* prev_count0, uint16_t* pairs0, int mode, cudaStream_t st) {
  const size_t esz = dtype == VET_F32 ? 4 : 8;
  const int T0 = h->ts[0].T;
  const int64_t fb = std::min<int64_t>(F, std::max<int64_t>(2, (int64_t)(((size_t)1 << 30) / ((size_t)U * 24))));
  for (int64_t f0 = 0; f0 < F - 1; f0 += fb - 1) {  // batches overlap by the halo frame
    const int64_t nf = std::min(fb, F - f0);
    if (int rc = decode_batch(h, (const char*)packed + (size_t)f0 * U * 3 * esz, dtype, nf * U, st)) return rc;
    if (int rc = transition_vectors_impl(h, (const double*)h->d_vscratch[2], nf, U, entropy + f0,
                                         per_k ? per_k + f0 : nullptr, F - 1, prev_count0 ? prev_count0 + f0 * T0 : nullptr,
                                         pairs0 ? pairs0 + f0 * U * 2 : nullptr, mode, st))
      return rc;
    if (nf == F - f0) break;
  }
  return VET_OK;
}

}  // namespace

extern "C" int vet_angular_distances(vet_handle* h, int k, const double* vec_dev, int64_t n, double* d_dev, void* stream) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_angular_distances: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || k < 0 || k >= h->K || n < 0 || (n > 0 && (!vec_dev || !d_dev))) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (n == 0) return VET_OK;
  DeviceGuard guard(h->device);
  const TileSet& t = h->ts[k];
  const int blocks = (int)std::min<int64_t>((n + 7) / 8, (int64_t)h->sm_count * 8);
  vet::k_angular_distances<<<blocks, 256, 0, (cudaStream_t)stream>>>(vec_dev, n, t.d_unit, t.T, d_dev);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

extern "C" int vet_spatial_vectors(vet_handle* h, const double* vec_dev, int64_t F, int64_t U, double* entropy_dev,
                                   double* per_k_dev, double* hist0_dev, uint16_t* assign0_dev, void* stream) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_spatial_vectors: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || F < 0 || U < 0) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (F == 0) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty vector dictionary");  // EU:168-169
  if (!vec_dev || !entropy_dev) return fail(VET_ERR_INVALID_ARG, "null buffer");
  DeviceGuard guard(h->device);
  return spatial_vectors_impl(h, vec_dev, F, U, entropy_dev, per_k_dev, F, hist0_dev, assign0_dev, (cudaStream_t)stream);
}

extern "C" int vet_transition_vectors(vet_handle* h, const double* vec_dev, int64_t F, int64_t U, double* entropy_dev,
                                      double* per_k_dev, int32_t* prev_count0_dev, uint16_t* pairs0_dev, int mode,
                                      void* stream) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_transition_vectors: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || F < 0 || U < 0) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (mode != VET_TRANSITION_LITERAL && mode != VET_TRANSITION_TEXTBOOK) return fail(VET_ERR_INVALID_ARG, "bad mode");
  if (F <= 1) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty vector dictionary");  // EU:239-240
  if (!vec_dev || !entropy_dev) return fail(VET_ERR_INVALID_ARG, "null buffer");
  if (U >= 0xFFFFFFFFll) return fail(VET_ERR_UNSUPPORTED, "too many users");
  DeviceGuard guard(h->device);
  return transition_vectors_impl(h, vec_dev, F, U, entropy_dev, per_k_dev, F - 1, prev_count0_dev, pairs0_dev, mode,
                                 (cudaStream_t)stream);
}

namespace {

// Sizes the (prev,cur) pair tables, picks the shared- or global-memory variant and launches
// k_transition for `rows` frame pairs.
// k_transition2 over all tile counts of `a` (dense table / shared-memory hash / global table per tile count);
// only_rows != null restricts it to the flagged rows.
int launch_transition2(vet_handle* h, const vet::TransitionArgs& a, int64_t U, int Tmax, int blocks, size_t tile_bytes,
                       const uint32_t* only_rows, cudaStream_t st) {
  {
    // fast paths: dense T*T table or shared-memory hash per tile count, global table as the in-kernel fallback
    vet::Transition2Args A2{};
    A2.t = a;
    A2.only_rows = only_rows;
    const size_t budget = h->smem_optin - kStaticSmemSlack;
    size_t table_words = 0;
    for (int k = 0; k < a.K; ++k) {
      const size_t dense_words = (size_t)a.T[k] * a.T[k];
      if (tile_bytes + dense_words * 4 + 64 <= budget) {
        A2.mode[k] = vet::kTrDense;
        table_words = std::max(table_words, dense_words);
      } else if (tile_bytes + (size_t)3 * vet::kHashSlots * 4 + 64 <= budget) {
        A2.mode[k] = vet::kTrHash;
        table_words = std::max(table_words, (size_t)3 * vet::kHashSlots);
      } else {
        A2.mode[k] = vet::kTrGlobal;
      }
    }
    // LUT staging area after the table area, for the tile counts whose LUT still fits
    const size_t lut_off = (tile_bytes + table_words * 4 + 64 + 15) & ~(size_t)15;
    size_t lut_area = 0;
    for (int k = 0; k < a.K; ++k) {
      // a.lut[k] is one of the handle's uint16 LUTs (or the identity table of the vectors path)
      const uint8_t* l8 = nullptr;
      for (int j = 0; j < h->K; ++j)
        if (h->ts[j].d_lut == a.lut[k]) l8 = h->ts[j].d_lut8;
      const bool is_cell_lut = a.lut[k] != h->d_identity;
      const size_t bytes = (((size_t)h->C * (l8 ? 1 : 2)) + 15) & ~(size_t)15;
      A2.lut8[k] = l8;
      A2.lut_smem[k] = (is_cell_lut && lut_off + bytes <= budget) ? 1 : 0;
      if (A2.lut_smem[k]) lut_area = std::max(lut_area, bytes);
    }
    A2.lut_area_off = (int)lut_off;
    const size_t smem2 = lut_off + lut_area;
    // packed (prev | cur << 16) per user, one row per CTA
    if (int rc = grow((void**)&h->d_pairs, &h->pairs_bytes, (size_t)blocks * U * 4)) return rc;
    A2.pair_scratch = h->d_pairs;
    A2.t.C = (int)h->C;
    VET_CUDA(cudaFuncSetAttribute(vet::k_transition2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    LaunchTimer lt(h, VET_KERNEL_TRANSITION, st);
    vet::k_transition2<<<blocks, vet::kTrThreads, smem2, st>>>(A2, Tmax);
  }
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

int launch_transition(vet_handle* h, vet::TransitionArgs& a, int64_t rows, int64_t U, int Tmax, cudaStream_t st) {
  // capacity >= 2 x the most distinct (prev,cur) pairs a frame pair can hold
  const uint64_t max_pairs = std::min<uint64_t>((uint64_t)U, (uint64_t)Tmax * Tmax);
  uint32_t cap = 1024;
  while ((uint64_t)cap < 2 * max_pairs) cap <<= 1;
  const size_t tile_bytes = (size_t)Tmax * (8 + 4 * 4);
  const size_t smem_tab = tile_bytes + (size_t)cap * 16 + 64;
  const bool in_smem = smem_tab + kStaticSmemSlack <= h->smem_optin;
  const int blocks = (int)std::min<int64_t>(rows, (int64_t)h->sm_count * (in_smem ? 1 : 2));
  if (!in_smem) {
    const size_t words = (size_t)blocks * 4 * cap;
    if (h->tables_words < words || h->tables_cap != cap) {
      if (h->tables_words < words) {
        if (h->d_tables) VET_CUDA(cudaFree(h->d_tables));
        h->d_tables = nullptr;
        h->tables_words = 0;
        VET_CUDA(cudaMalloc((void**)&h->d_tables, words * 4));
        h->tables_words = words;
      }
      // keys/firsts = 0xFFFFFFFF, counts = 0 (list needs no initial value).  Done once per layout:
      // the kernels reset every slot they touch, so the tables stay clean between calls.
      for (int b = 0; b < blocks; ++b) {
        VET_CUDA(cudaMemsetAsync(h->d_tables + (size_t)b * 4 * cap, 0xFF, (size_t)cap * 8, st));
        VET_CUDA(cudaMemsetAsync(h->d_tables + (size_t)b * 4 * cap + 2 * (size_t)cap, 0, (size_t)cap * 4, st));
      }
      h->tables_cap = cap;
      h->tables_blocks = blocks;
    } else if (h->tables_blocks < blocks) {
      for (int b = h->tables_blocks; b < blocks; ++b) {
        VET_CUDA(cudaMemsetAsync(h->d_tables + (size_t)b * 4 * cap, 0xFF, (size_t)cap * 8, st));
        VET_CUDA(cudaMemsetAsync(h->d_tables + (size_t)b * 4 * cap + 2 * (size_t)cap, 0, (size_t)cap * 4, st));
      }
      h->tables_blocks = blocks;
    }
  }
  a.cap = cap;
  a.g_tables = h->d_tables;
  // VET_TRANSITION_IMPL = v1 | v2 pins an older kernel generation (A/B runs, tests); read at every call
  const bool force_v1 = [] {
    const char* e = getenv("VET_TRANSITION_IMPL");
    return e && std::string(e) == "v1";
  }();
  const bool force_v2 = [] {
    const char* e = getenv("VET_TRANSITION_IMPL");
    return e && std::string(e) == "v2";
  }();
  // two-pass kernel, one launch per tile count; rows it cannot hold (hash overflow) are flagged in d_redo
  // and recomputed by k_transition2 below
  bool redo_only = false;
  if (a.mode == VET_TRANSITION_LITERAL && !in_smem && !force_v1 && !force_v2 && a.cell16 && U < ((int64_t)1 << 31)) {
    const size_t budget = h->smem_optin - kStaticSmemSlack;
    struct Plan {
      int mode, lw;
      size_t tab_off, lut_off, smem;
      const void* lut;
    } plan[vet::kMaxTileCounts];
    bool ok = true;
    for (int k = 0; k < a.K && ok; ++k) {
      const size_t T = (size_t)a.T[k];
      const uint8_t* l8 = nullptr;
      bool is_cell_lut = false;
      for (int j = 0; j < h->K; ++j)
        if (h->ts[j].d_lut == a.lut[k]) {
          l8 = h->ts[j].d_lut8;
          is_cell_lut = true;
        }
      const bool identity = a.lut[k] == h->d_identity;  // the input rows hold tile ids already
      if (!is_cell_lut && !identity) {
        ok = false;
        break;
      }
      Plan& pl = plan[k];
      pl.tab_off = (T * vet::kT3TileBytes + 15) & ~(size_t)15;
      const size_t dense = T * vet::t3_row_stride((uint32_t)T) * 4, hash = (size_t)2 * vet::kT3Slots * 4;
      size_t tab;
      if (pl.tab_off + dense <= budget) {
        pl.mode = vet::kT3Dense;
        tab = dense;
      } else if (pl.tab_off + hash <= budget) {
        pl.mode = vet::kT3Hash;  // rows with more distinct pairs than the table holds are redone by k_transition2
        tab = hash;
      } else {
        pl.mode = -1;  // this tile count goes to k_transition2 (global pair tables)
        continue;
      }
      pl.lut_off = (pl.tab_off + tab + 15) & ~(size_t)15;
      const size_t lut_bytes = (((size_t)h->C * (l8 ? 1 : 2)) + 15) & ~(size_t)15;
      if (identity) {
        pl.lw = vet::kLutIdentity;
        pl.lut = nullptr;
        pl.smem = pl.lut_off;
      } else if (pl.lut_off + lut_bytes <= budget) {
        pl.lw = l8 ? vet::kLutS8 : vet::kLutS16;
        pl.lut = l8 ? (const void*)l8 : (const void*)a.lut[k];
        pl.smem = pl.lut_off + lut_bytes;
      } else {
        pl.lw = vet::kLutG16;
        pl.lut = a.lut[k];
        pl.smem = pl.lut_off;
      }
    }
    if (ok) {
      const int blocks3 = (int)std::min<int64_t>(rows, h->sm_count);
      if (int rc = grow((void**)&h->d_pairs, &h->pairs_bytes, (size_t)std::max(blocks, blocks3) * U * 4)) return rc;
      if (int rc = grow((void**)&h->d_redo, &h->redo_bytes, (size_t)rows * 4)) return rc;
      VET_CUDA(cudaMemsetAsync(h->d_redo, 0, (size_t)rows * 4, st));
      double* per_k = a.per_k;
      int64_t stride = a.per_k_stride;
      if (a.K > 1 && !per_k) {
        if (int rc = grow((void**)&h->d_trk, &h->trk_bytes, (size_t)a.K * rows * 8)) return rc;
        per_k = h->d_trk;
        stride = rows;
      }
      bool any_hash = false;
      for (int k = 0; k < a.K; ++k) {
        const Plan& pl = plan[k];
        double* out_k = a.K == 1 ? a.entropy : per_k + k * stride;
        if (pl.mode < 0) {
          vet::TransitionArgs a1 = a;
          a1.K = 1;
          a1.T[0] = a.T[k];
          a1.lut[0] = a.lut[k];
          a1.entropy = out_k;
          a1.per_k = nullptr;
          a1.prev_count0 = k == 0 ? a.prev_count0 : nullptr;
          a1.pairs0 = k == 0 ? a.pairs0 : nullptr;
          if (int rc = launch_transition2(h, a1, U, Tmax, blocks, tile_bytes, nullptr, st)) return rc;
          continue;
        }
        any_hash = any_hash || pl.mode == vet::kT3Hash;
        vet::Transition3Args A3{};
        A3.cell16 = a.cell16;
        A3.F = a.F;
        A3.U = (uint32_t)U;
        A3.T = a.T[k];
        A3.C = (int)h->C;
        A3.lut_src = pl.lut;
        A3.tab_off = (int)pl.tab_off;
        A3.lut_off = (int)pl.lut_off;
        A3.out = out_k;
        A3.prev_count0 = k == 0 ? a.prev_count0 : nullptr;
        A3.pairs0 = k == 0 ? a.pairs0 : nullptr;
        A3.pair_scratch = h->d_pairs;
        A3.redo = h->d_redo;
        A3.flags = a.flags;
        LaunchTimer lt(h, VET_KERNEL_TRANSITION, st);
#define VET_T3(MODE, LW)                                                                                              \
  do {                                                                                                                \
    VET_CUDA(cudaFuncSetAttribute(vet::k_transition3<MODE, LW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget)); \
    vet::k_transition3<MODE, LW><<<blocks3, vet::kT3Threads, pl.smem, st>>>(A3);                                      \
  } while (0)
        if (pl.mode == vet::kT3Dense) {
          if (pl.lw == vet::kLutS8) VET_T3(vet::kT3Dense, vet::kLutS8);
          else if (pl.lw == vet::kLutS16) VET_T3(vet::kT3Dense, vet::kLutS16);
          else if (pl.lw == vet::kLutIdentity) VET_T3(vet::kT3Dense, vet::kLutIdentity);
          else VET_T3(vet::kT3Dense, vet::kLutG16);
        } else {
          if (pl.lw == vet::kLutS8) VET_T3(vet::kT3Hash, vet::kLutS8);
          else if (pl.lw == vet::kLutS16) VET_T3(vet::kT3Hash, vet::kLutS16);
          else if (pl.lw == vet::kLutIdentity) VET_T3(vet::kT3Hash, vet::kLutIdentity);
          else VET_T3(vet::kT3Hash, vet::kLutG16);
        }
#undef VET_T3
        VET_CUDA(cudaGetLastError());
      }
      if (a.K == 1 && a.per_k) {
        VET_CUDA(cudaMemcpyAsync(a.per_k, a.entropy, (size_t)rows * 8, cudaMemcpyDeviceToDevice, st));
      } else if (a.K > 1) {
        LaunchTimer lt(h, VET_KERNEL_TRANSITION, st);
        vet::k_mean_rows<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(per_k, stride, a.K, rows, a.entropy);
        VET_CUDA(cudaGetLastError());
      }
      if (!any_hash) return VET_OK;  // nothing can have been left over
      redo_only = true;
    }
  }
  if (a.mode == VET_TRANSITION_LITERAL && !in_smem && !force_v1) {
    if (int rc = launch_transition2(h, a, U, Tmax, blocks, tile_bytes, redo_only ? h->d_redo : nullptr, st)) return rc;
  } else if (in_smem) {
    VET_CUDA(cudaFuncSetAttribute(vet::k_transition<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tab));
    LaunchTimer lt(h, VET_KERNEL_TRANSITION, st);
    vet::k_transition<true><<<blocks, 512, smem_tab, st>>>(a, Tmax);
  } else {
    VET_CUDA(cudaFuncSetAttribute(vet::k_transition<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(tile_bytes + 64)));
    LaunchTimer lt(h, VET_KERNEL_TRANSITION, st);
    vet::k_transition<false><<<blocks, 512, tile_bytes + 64, st>>>(a, Tmax);
  }
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

}  // namespace

extern "C" int vet_naive_points(vet_handle* h, const double* lonlat_dev, int64_t F, int64_t U, int32_t tile_width,
                                int32_t tile_height, int32_t use_weight_distribution, double* entropy_dev,
                                int32_t* lon_idx_dev, int32_t* lat_idx_dev, void* stream) {
  if (!h || F < 0 || U < 0) return fail(VET_ERR_INVALID_ARG, "bad argument");
  // EU:404-417
  if (tile_width <= 0 || tile_height <= 0) return fail(VET_ERR_INVALID_ARG, "No tile dimensions provided");
  if (180 % tile_height != 0) return fail(VET_ERR_INVALID_ARG, "Tile height must divide 180!");
  if (360 % tile_width != 0) return fail(VET_ERR_INVALID_ARG, "Tile width must divide 360!");
  if (F == 0) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty radial points dictionary");
  if (!lonlat_dev || !entropy_dev) return fail(VET_ERR_INVALID_ARG, "null buffer");
  DeviceGuard guard(h->device);
  vet::NaivePointsArgs a{};
  a.lonlat = lonlat_dev;
  a.F = F;
  a.U = U;
  a.tile_width = tile_width;
  a.tile_height = tile_height;
  a.nlat1 = 180 / tile_height + 1;
  a.ncodes = (360 / tile_width + 1) * a.nlat1;
  a.num_tiles = (180 / tile_height) * (360 / tile_width);
  a.norm_always = use_weight_distribution ? 1 : 0;
  a.entropy = entropy_dev;
  a.lon_idx = lon_idx_dev;
  a.lat_idx = lat_idx_dev;
  a.flags = h->d_flags;
  const size_t smem = (size_t)a.ncodes * 4;
  if (smem + kStaticSmemSlack > h->smem_optin)
    return fail(VET_ERR_UNSUPPORTED, "%dx%d degree tiles give %d grid codes; too many for one frame's shared-memory histogram", tile_width, tile_height, a.ncodes);
  VET_CUDA(cudaFuncSetAttribute(vet::k_naive_points, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(h->smem_optin - kStaticSmemSlack)));
  h->launches++;
  vet::k_naive_points<<<(int)std::min<int64_t>(F, (int64_t)h->sm_count * 8), 256, smem, (cudaStream_t)stream>>>(a);
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

extern "C" int vet_poll_flags(vet_handle* h, void* stream, uint32_t* flags) {
  if (!h || !flags) return fail(VET_ERR_INVALID_ARG, "null argument");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  VET_CUDA(cudaMemcpyAsync(flags, h->d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  VET_CUDA(cudaMemsetAsync(h->d_flags, 0, sizeof(uint32_t), st));
  VET_CUDA(cudaStreamSynchronize(st));
  return VET_OK;
}

extern "C" int vet_profile_enable(vet_handle* h, int on) {
  if (!h) return fail(VET_ERR_INVALID_ARG, "null handle");
  DeviceGuard guard(h->device);
  for (auto& s : h->spans) {
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  h->spans.clear();
  h->profiling = on != 0;
  return VET_OK;
}

extern "C" int vet_profile_read(vet_handle* h, double* ms_by_kernel, int64_t* launches_by_kernel) {
  if (!h || !ms_by_kernel || !launches_by_kernel) return fail(VET_ERR_INVALID_ARG, "null argument");
  DeviceGuard guard(h->device);
  for (int i = 0; i < VET_KERNEL_COUNT; ++i) {
    ms_by_kernel[i] = 0.0;
    launches_by_kernel[i] = 0;
  }
  for (auto& s : h->spans) {
    VET_CUDA(cudaEventSynchronize(s.b));
    float ms = 0.f;
    VET_CUDA(cudaEventElapsedTime(&ms, s.a, s.b));
    ms_by_kernel[s.kernel] += ms;
    launches_by_kernel[s.kernel] += 1;
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  h->spans.clear();
  return VET_OK;
}

// ---- host-buffer variants ---------------------------------------------------------

namespace {

// frames per host batch: about 256 MiB of packed input per copy
int64_t host_batch_frames(int64_t F, int64_t U, size_t esz) {
  const size_t per_frame = (size_t)U * 3 * esz;
  return std::min<int64_t>(F, std::max<int64_t>(2, (int64_t)(((size_t)256 << 20) / std::max<size_t>(per_frame, 1))));
}

}  // namespace

extern "C" int vet_spatial_host(vet_handle* h, const void* packed_host, int dtype, int64_t F, int64_t U,
                                double* entropy_host, double* per_k_host, double* hist0_host, uint16_t* assign0_host) {
  if (!h || F < 0 || U < 0 || (dtype != VET_F32 && dtype != VET_F64)) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (F == 0) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty vector dictionary");
  if (!packed_host || !entropy_host) return fail(VET_ERR_INVALID_ARG, "null buffer");
  DeviceGuard guard(h->device);
  const size_t esz = dtype == VET_F32 ? 4 : 8;
  if (h->direct_only) {  // large-video mode: plain upload, direct kernels, download
    const int T0d = h->ts[0].T;
    void* d_in = nullptr;
    double *d_e = nullptr, *d_p = nullptr, *d_h = nullptr;
    uint16_t* d_a = nullptr;
    VET_CUDA(cudaMalloc(&d_in, (size_t)F * U * 3 * esz));
    VET_CUDA(cudaMalloc((void**)&d_e, (size_t)F * 8));
    if (per_k_host) VET_CUDA(cudaMalloc((void**)&d_p, (size_t)F * h->K * 8));
    if (hist0_host) VET_CUDA(cudaMalloc((void**)&d_h, (size_t)F * T0d * 8));
    if (assign0_host) VET_CUDA(cudaMalloc((void**)&d_a, (size_t)F * U * 2));
    cudaMemcpyAsync(d_in, packed_host, (size_t)F * U * 3 * esz, cudaMemcpyHostToDevice, h->s_exec);
    int rc = spatial_direct(h, d_in, dtype, F, U, d_e, d_p, d_h, d_a, h->s_exec);
    if (rc == VET_OK) {
      cudaMemcpyAsync(entropy_host, d_e, (size_t)F * 8, cudaMemcpyDeviceToHost, h->s_exec);
      if (per_k_host) cudaMemcpyAsync(per_k_host, d_p, (size_t)F * h->K * 8, cudaMemcpyDeviceToHost, h->s_exec);
      if (hist0_host) cudaMemcpyAsync(hist0_host, d_h, (size_t)F * T0d * 8, cudaMemcpyDeviceToHost, h->s_exec);
      if (assign0_host) cudaMemcpyAsync(assign0_host, d_a, (size_t)F * U * 2, cudaMemcpyDeviceToHost, h->s_exec);
    }
    cudaError_t e = cudaStreamSynchronize(h->s_exec);
    cudaFree(d_in);
    cudaFree(d_e);
    cudaFree(d_p);
    cudaFree(d_h);
    cudaFree(d_a);
    if (rc != VET_OK) return rc;
    if (e != cudaSuccess) return fail(VET_ERR_CUDA, "host-buffer pipeline failed: %s", cudaGetErrorString(e));
    return VET_OK;
  }
  const int64_t fb = host_batch_frames(F, U, esz);
  const size_t in_bytes = (size_t)fb * U * 3 * esz;
  if (h->in_bytes < in_bytes) {
    for (int i = 0; i < 2; ++i) {
      if (h->d_in[i]) VET_CUDA(cudaFree(h->d_in[i]));
      h->d_in[i] = nullptr;
    }
    h->in_bytes = 0;
    for (int i = 0; i < 2; ++i) VET_CUDA(cudaMalloc(&h->d_in[i], in_bytes));
    h->in_bytes = in_bytes;
  }
  const int T0 = h->ts[0].T;
  // device-side result buffers are kept in the handle and only grown (cudaMalloc/cudaFree synchronise)
  if (int rc = grow(&h->d_hout[0], &h->hout_bytes[0], (size_t)F * 8)) return rc;
  if (per_k_host)
    if (int rc = grow(&h->d_hout[1], &h->hout_bytes[1], (size_t)F * h->K * 8)) return rc;
  if (hist0_host)
    if (int rc = grow(&h->d_hout[2], &h->hout_bytes[2], (size_t)F * T0 * 8)) return rc;
  if (assign0_host)
    for (int i = 0; i < 2; ++i)
      if (int rc = grow(&h->d_hout[3 + i], &h->hout_bytes[3 + i], (size_t)fb * U * 2)) return rc;
  double* d_ent = (double*)h->d_hout[0];
  double* d_perk = per_k_host ? (double*)h->d_hout[1] : nullptr;
  double* d_hist = hist0_host ? (double*)h->d_hout[2] : nullptr;
  uint16_t* d_assign[2] = {assign0_host ? (uint16_t*)h->d_hout[3] : nullptr, assign0_host ? (uint16_t*)h->d_hout[4] : nullptr};
  // Three streams: copy-in, execute, copy-out.  Batch b+1 is uploaded while batch b runs and
  // batch b-1's assignments are downloaded (PCIe is full duplex).
  cudaEvent_t in_done[2], exec_done[2], out_done[2];
  for (int i = 0; i < 2; ++i) {
    VET_CUDA(cudaEventCreateWithFlags(&in_done[i], cudaEventDisableTiming));
    VET_CUDA(cudaEventCreateWithFlags(&exec_done[i], cudaEventDisableTiming));
    VET_CUDA(cudaEventCreateWithFlags(&out_done[i], cudaEventDisableTiming));
  }
  int rc = VET_OK;
  int b = 0;
  for (int64_t f0 = 0; f0 < F && rc == VET_OK; f0 += fb, b ^= 1) {
    const int64_t nf = std::min(fb, F - f0);
    cudaStreamWaitEvent(h->s_copy, exec_done[b], 0);  // input buffer b was last read two batches ago
    cudaMemcpyAsync(h->d_in[b], (const char*)packed_host + (size_t)f0 * U * 3 * esz, (size_t)nf * U * 3 * esz,
                    cudaMemcpyHostToDevice, h->s_copy);
    cudaEventRecord(in_done[b], h->s_copy);
    cudaStreamWaitEvent(h->s_exec, in_done[b], 0);
    cudaStreamWaitEvent(h->s_exec, out_done[b], 0);  // assignment buffer b must have been downloaded
    const int64_t fbs = frames_per_batch(h, nf, U, false);
    rc = grow((void**)&h->d_cnt, &h->cnt_bytes, cnt_scratch_bytes(h, fbs));
    if (rc == VET_OK) rc = grow((void**)&h->d_nvalid, &h->nvalid_bytes, (size_t)fbs * 4);
    for (int64_t g0 = 0; g0 < nf && rc == VET_OK; g0 += fbs) {
      const int64_t ng = std::min(fbs, nf - g0);
      const char* in = (const char*)h->d_in[b] + (size_t)g0 * U * 3 * esz;
      TilesPlan tp = plan_tiles(h, in, U);
      if (tp.ok) {
        rc = launch_stream_tiles(h, tp, in, dtype, ng, U, d_assign[b] ? d_assign[b] + g0 * U : nullptr, h->s_exec);
        if (rc == VET_OK)
          rc = launch_tiles_epilogue(h, tp, ng, d_ent + f0 + g0, d_perk ? d_perk + f0 + g0 : nullptr, F,
                                     d_hist ? d_hist + (f0 + g0) * T0 : nullptr, h->s_exec);
        continue;
      }
      rc = launch_stream(h, in, dtype, ng, U, d_assign[b] ? d_assign[b] + g0 * U : nullptr, false, h->s_exec);
      if (rc == VET_OK)
        rc = launch_epilogue(h, ng, U, d_ent + f0 + g0, d_perk ? d_perk + f0 + g0 : nullptr, F,
                             d_hist ? d_hist + (f0 + g0) * T0 : nullptr, h->s_exec);
    }
    cudaEventRecord(exec_done[b], h->s_exec);
    if (rc == VET_OK && assign0_host) {
      cudaStreamWaitEvent(h->s_out, exec_done[b], 0);
      cudaMemcpyAsync(assign0_host + f0 * U, d_assign[b], (size_t)nf * U * 2, cudaMemcpyDeviceToHost, h->s_out);
      cudaEventRecord(out_done[b], h->s_out);
    }
  }
  if (rc == VET_OK) {
    cudaMemcpyAsync(entropy_host, d_ent, (size_t)F * 8, cudaMemcpyDeviceToHost, h->s_exec);
    if (per_k_host) cudaMemcpyAsync(per_k_host, d_perk, (size_t)F * h->K * 8, cudaMemcpyDeviceToHost, h->s_exec);
    if (hist0_host) cudaMemcpyAsync(hist0_host, d_hist, (size_t)F * T0 * 8, cudaMemcpyDeviceToHost, h->s_exec);
  }
  cudaError_t e1 = cudaStreamSynchronize(h->s_copy), e2 = cudaStreamSynchronize(h->s_exec),
              e3 = cudaStreamSynchronize(h->s_out);
  for (int i = 0; i < 2; ++i) {
    cudaEventDestroy(in_done[i]);
    cudaEventDestroy(exec_done[i]);
    cudaEventDestroy(out_done[i]);
  }
  if (rc != VET_OK) return rc;
  for (cudaError_t e : {e1, e2, e3})
    if (e != cudaSuccess) return fail(VET_ERR_CUDA, "host-buffer pipeline failed: %s", cudaGetErrorString(e));
  return VET_OK;
}

extern "C" int vet_transition_host(vet_handle* h, const void* packed_host, int dtype, int64_t F, int64_t U,
                                   double* entropy_host, double* per_k_host, int32_t* prev_count0_host,
                                   uint16_t* pairs0_host, int mode) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_transition_host: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || F < 0 || U < 0 || (dtype != VET_F32 && dtype != VET_F64)) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (F <= 1) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty vector dictionary");
  if (!packed_host || !entropy_host) return fail(VET_ERR_INVALID_ARG, "null buffer");
  DeviceGuard guard(h->device);
  const size_t esz = dtype == VET_F32 ? 4 : 8;
  const int T0 = h->ts[0].T;
  void* d_in = nullptr;
  double *d_ent = nullptr, *d_perk = nullptr;
  int32_t* d_pc = nullptr;
  uint16_t* d_pairs = nullptr;
  const size_t in_bytes = (size_t)F * U * 3 * esz;
  VET_CUDA(cudaMalloc(&d_in, in_bytes));
  VET_CUDA(cudaMalloc((void**)&d_ent, (size_t)(F - 1) * 8));
  if (per_k_host) VET_CUDA(cudaMalloc((void**)&d_perk, (size_t)(F - 1) * h->K * 8));
  if (prev_count0_host) VET_CUDA(cudaMalloc((void**)&d_pc, (size_t)(F - 1) * T0 * 4));
  if (pairs0_host) VET_CUDA(cudaMalloc((void**)&d_pairs, (size_t)(F - 1) * U * 4));
  cudaMemcpyAsync(d_in, packed_host, in_bytes, cudaMemcpyHostToDevice, h->s_exec);
  int rc = vet_transition(h, d_in, dtype, F, U, d_ent, d_perk, d_pc, d_pairs, mode, h->s_exec);
  if (rc == VET_OK) {
    cudaMemcpyAsync(entropy_host, d_ent, (size_t)(F - 1) * 8, cudaMemcpyDeviceToHost, h->s_exec);
    if (per_k_host) cudaMemcpyAsync(per_k_host, d_perk, (size_t)(F - 1) * h->K * 8, cudaMemcpyDeviceToHost, h->s_exec);
    if (prev_count0_host)
      cudaMemcpyAsync(prev_count0_host, d_pc, (size_t)(F - 1) * T0 * 4, cudaMemcpyDeviceToHost, h->s_exec);
    if (pairs0_host) cudaMemcpyAsync(pairs0_host, d_pairs, (size_t)(F - 1) * U * 4, cudaMemcpyDeviceToHost, h->s_exec);
  }
  cudaError_t e = cudaStreamSynchronize(h->s_exec);
  cudaFree(d_in);
  cudaFree(d_ent);
  cudaFree(d_perk);
  cudaFree(d_pc);
  cudaFree(d_pairs);
  if (rc != VET_OK) return rc;
  if (e != cudaSuccess) return fail(VET_ERR_CUDA, "host-buffer pipeline failed: %s", cudaGetErrorString(e));
  return VET_OK;
}
