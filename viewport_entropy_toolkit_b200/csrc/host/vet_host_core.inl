// Host core of libvet_b200: error reporting, per-tile-count tables (TileSet), the handle, libm/numpy-compatible
// table helpers, small CUDA utilities and the forward declarations shared by the other fragments.
// Textual fragment of vet_b200.cu (one translation unit).
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
  // one-pass transition kernel (vet_transition4.cuh)
  uint16_t* d_drank = nullptr;    // [bands][2 win + 2] rank (as rank * T * 4) of the tile-index delta c - p of a transition,
                                  // per band of 2^band_shift consecutive previous tiles (0xFFFF = unranked)
  int* d_rank_delta = nullptr;    // [bands][16] delta of every rank (rank 0 = staying in the tile)
  int band_shift = 0, win = 0, bands = 0;
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
  bool i8_ok = false;                  // weight quantisation within the entropy tolerance (build_i8_tables)
  double i8_rho = 0.0;
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
  void* d_i8acc = nullptr;       // split-K slices of k_whist_i8's accumulators
  size_t i8acc_bytes = 0;
  CUtensorMap tm_cnt;
  bool planes_from_stream = false;  // the last launch_stream wrote the planes of its batch itself
  bool i8_attr_set = false;
  int64_t call_frames = 0;       // frames of the API call in progress: its batches all take the same weighted kernel
  uint32_t* d_redo = nullptr;   // [rows] frame pairs the two-pass transition kernel left to k_transition2
  size_t redo_bytes = 0;
  void* d_t4 = nullptr;         // one-pass transition kernel: per tile count the count and the flags of the pairs it
  size_t t4_bytes = 0;          // leaves to the two-pass kernels
  uint16_t* d_rows = nullptr;   // [frames, U] tile ids of one tile count, relabelled from the cell ids (several tile counts)
  size_t rows_bytes = 0;
  double* d_trk = nullptr;      // [K, rows] per-tile-count transition entropies when the caller wants none
  size_t trk_bytes = 0;
  uint32_t tables_cap = 0;  // slot count the tables are currently laid out (and cleared) for
  int tables_blocks = 0;    // number of per-CTA tables cleared for that layout
  // host-buffer path
  void* d_in[2] = {nullptr, nullptr};
  size_t in_bytes = 0;
  void* d_in2[2] = {nullptr, nullptr};  // staging of the two-column host layout (VET_OPT_HOST_LAYOUT), widened into d_in
  size_t in2_bytes = 0;
  void* d_hout[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // entropy, per_k, hist0, assign x2
  size_t hout_bytes[5] = {0, 0, 0, 0, 0};
  void* d_hout2[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // transition rows: entropy, per_k, prev_count0, pairs x2
  size_t hout2_bytes[5] = {0, 0, 0, 0, 0};
  cudaStream_t s_copy = nullptr, s_exec = nullptr, s_out = nullptr;
  cudaStream_t s_side = nullptr;                     // vet_analyze: spatial epilogue beside the transition kernel
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_pipe[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // host-buffer pipelines: in / exec / out done x 2
  struct T3cOcc {
    int lw, S;
    size_t smem;
    int n;
  };
  // CUDA graphs of whole API calls (run_graphed): key = everything the launch sequence depends on
  struct GraphSlot {
    int api = 0, dtype = 0, mode = 0, seen = 0;
    int64_t F = 0, U = 0;
    const void* ptr[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t st = nullptr;
    uint64_t epoch = 0;       // scratch allocations at capture time: a reallocation invalidates the graph
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;     // kernel launches of one replay
    uint64_t used = 0;
  };
  std::vector<GraphSlot> graphs;
  uint64_t graph_clock = 0;
  int64_t graph_replays = 0;
  std::vector<T3cOcc> t3c_occ;  // co-resident clusters of k_transition3c per (LUT variant, cluster size, shared memory)
  int64_t launches = 0;
  int opt[VET_OPT_COUNT] = {0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0};  // vet_set_option (defaults: cluster tail auto, analyze overlap on)
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

uint64_t g_scratch_epoch = 0;  // bumped by every scratch reallocation (captured graphs hold the old addresses)

int grow(void** ptr, size_t* have, size_t want) {
  if (*have >= want) return VET_OK;
  ++g_scratch_epoch;
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
void drop_graphs(vet_handle* h);
int spatial_direct(vet_handle* h, const void* packed, int dtype, int64_t F, int64_t U, double* entropy, double* per_k,
                   double* hist0, uint16_t* assign0, cudaStream_t st);
int transition_direct(vet_handle* h, const void* packed, int dtype, int64_t F, int64_t U, double* entropy, double* per_k,
                      int32_t* prev_count0, uint16_t* pairs0, int mode, cudaStream_t st);

constexpr size_t kStaticSmemSlack = 1024;
constexpr int kMaxT = 16384;
constexpr int64_t kGlobalTableCells = 262144;  // largest cell grid of the global-table regime (tables scale with C*T)
constexpr int64_t kGlobalLutCells = (int64_t)1 << 24;  // the same for unweighted handles (only LUTs: 2 B x C per tile count)

// tensor-core weighted histogram (defined with launch_whist_i8 below)
constexpr int64_t kI8MinFrames = VET_I8_MIN_FRAMES;
bool use_whist_i8(const vet_handle* h, int64_t F, int64_t U);
int build_i8_tables(vet_handle* h, TileSet& t);
int ensure_planes(vet_handle* h, int64_t F, cudaStream_t st);
int64_t i8_kp(const vet_handle* h);
uint32_t* i8_hi1(vet_handle* h);
uint32_t* i8_hi2(vet_handle* h);

}  // namespace
