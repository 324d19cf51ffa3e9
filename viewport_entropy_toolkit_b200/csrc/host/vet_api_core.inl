// C ABI: handle lifetime, table getters, decode / nearest / weights, vet_spatial, vet_transition, vet_analyze.
// Textual fragment of vet_b200.cu.

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
  if (cfg->regime != VET_REGIME_AUTO && cfg->regime != VET_REGIME_DIRECT) return fail(VET_ERR_INVALID_ARG, "bad regime");
  h->direct_only = stream_smem_bytes(h) + kStaticSmemSlack > h->smem_optin ||
                   epilogue_smem_bytes(h) + kStaticSmemSlack > h->smem_optin;
  // weighted handles need per-cell weight tables (C x T): bounded; unweighted ones only the cell -> tile LUTs
  if (cfg->regime == VET_REGIME_DIRECT && !naive) {
    h->direct_only = true;  // pinned: per-sample evaluation without cell tables
  } else if (h->direct_only) {
    // Weighted handles keep per-cell weight columns: about C x sum(T_k) x (share of the sphere inside fov/2) entries of
    // 12 bytes, plus the dense blocks / quantised slices made from them.  Up to 262,144 cells always; beyond that (a
    // 1920x1080 video has 2.08 M cells) while the estimate stays under 2^27 entries (~1.6 GB of columns) -- e.g. 201
    // tiles at fov = 120 on 1920x1080: 105 M; the reference's default five tile counts there: 741 M -> direct regime.
    const double cap_share = 0.5 * (1.0 - std::cos(h->max_d));
    const double est_entries = (double)h->C * (double)h->sumT * std::min(1.0, cap_share * 1.15 + 0.01);
    const bool weights_fit = h->C <= kGlobalTableCells || (h->C <= kGlobalLutCells && est_entries <= (double)((int64_t)1 << 27));
    if ((h->use_weight && weights_fit) || (!h->use_weight && h->C <= kGlobalLutCells)) {
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
  VET_CUDA(cudaStreamCreateWithFlags(&h->s_side, cudaStreamNonBlocking));
  for (auto& e : h->ev_pipe) VET_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  VET_CUDA(cudaDeviceSynchronize());
  cleanup.armed = false;
  *out = h;
  return VET_OK;
}

extern "C" int vet_destroy(vet_handle* h) {
  if (!h) return VET_OK;
  DeviceGuard guard(h->device);
  for (auto& t : h->ts) free_tile_set(t);
  drop_graphs(h);
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
  cudaFree(h->d_t4);
  cudaFree(h->d_rows);
  cudaFree(h->d_trk);
  cudaFree(h->d_planes);
  cudaFree(h->d_dirty);
  cudaFree(h->d_i8flags);
  cudaFree(h->d_i8acc);
  cudaFree(h->d_in[0]);
  cudaFree(h->d_in[1]);
  cudaFree(h->d_in2[0]);
  cudaFree(h->d_in2[1]);
  for (void* p : h->d_hout) cudaFree(p);
  for (void* p : h->d_hout2) cudaFree(p);
  for (auto e : h->ev_pipe)
    if (e) cudaEventDestroy(e);
  if (h->s_side) cudaStreamDestroy(h->s_side);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
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
extern "C" int64_t vet_graph_replays(const vet_handle* h) { return h ? h->graph_replays : 0; }

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

namespace {

void drop_graphs(vet_handle* h) {
  for (auto& g : h->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  h->graphs.clear();
}

// Runs `body` (the launch sequence of one API call on stream st) -- eagerly the first two times a call with this key is
// seen, captured into a CUDA graph the third time, replayed from then on: one graph launch instead of the 10-20 kernel,
// memset and event calls of a step, whose gaps are ~3 % of the configs[2] step.  The key holds every buffer address,
// size and option the sequence depends on plus the scratch epoch; the legacy default stream cannot be captured and
// profiling wants per-kernel events, so both stay eager.  Any failure of the capture falls back to the eager path.
template <typename Body>
int run_graphed(vet_handle* h, vet_handle::GraphSlot key, cudaStream_t st, Body&& body) {
  const bool eligible = h->opt[VET_OPT_CUDA_GRAPH] != 0 && !h->profiling && st != nullptr && st != cudaStreamLegacy &&
                        st != cudaStreamPerThread;
  if (!eligible) return body();
  key.st = st;
  vet_handle::GraphSlot* slot = nullptr;
  for (auto& g : h->graphs) {
    bool same = g.api == key.api && g.dtype == key.dtype && g.mode == key.mode && g.F == key.F && g.U == key.U && g.st == key.st;
    for (int i = 0; i < 10 && same; ++i) same = g.ptr[i] == key.ptr[i];
    if (same) slot = &g;
  }
  if (slot && slot->exec && slot->epoch == g_scratch_epoch) {
    slot->used = ++h->graph_clock;
    h->launches += slot->launches;
    h->graph_replays++;
    VET_CUDA(cudaGraphLaunch(slot->exec, st));
    return VET_OK;
  }
  if (!slot) {
    if (h->graphs.size() >= 8) {  // drop the least recently used
      size_t lru = 0;
      for (size_t i = 1; i < h->graphs.size(); ++i)
        if (h->graphs[i].used < h->graphs[lru].used) lru = i;
      if (h->graphs[lru].exec) cudaGraphExecDestroy(h->graphs[lru].exec);
      h->graphs.erase(h->graphs.begin() + (long)lru);
    }
    key.seen = 0;
    h->graphs.push_back(key);
    slot = &h->graphs.back();
  }
  slot->used = ++h->graph_clock;
  if (slot->exec) {  // stale: scratch was reallocated since the capture
    cudaGraphExecDestroy(slot->exec);
    slot->exec = nullptr;
    slot->seen = 0;
  }
  if (++slot->seen < 3) return body();  // tables, scratch and lazy state settle in the eager calls
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return body();  // the caller is capturing this stream itself
  }
  const uint64_t epoch0 = g_scratch_epoch;
  const int64_t launches0 = h->launches;
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    cudaGetLastError();
    slot->seen = -1000000;  // not capturable here: stay eager
    return body();
  }
  const int rc = body();
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(st, &graph);
  cudaGraphExec_t exec = nullptr;
  if (rc == VET_OK && ce == cudaSuccess && graph && epoch0 == g_scratch_epoch &&
      cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess && exec) {
    cudaGraphDestroy(graph);
    slot->exec = exec;
    slot->epoch = g_scratch_epoch;
    slot->launches = h->launches - launches0;
    h->graph_replays++;
    VET_CUDA(cudaGraphLaunch(exec, st));
    return VET_OK;
  }
  cudaGetLastError();
  if (graph) cudaGraphDestroy(graph);
  if (exec) cudaGraphExecDestroy(exec);
  h->launches = launches0;
  slot->seen = -1000000;
  if (rc != VET_OK) return rc;
  return body();  // nothing ran during the failed capture
}

// SpatialEntropyAnalyzer.compute_entropy on F resident frames (table regimes); per_k rows `per_k_stride` apart.
int spatial_core(vet_handle* h, const void* packed_dev, int dtype, int64_t F, int64_t U, double* entropy_dev,
                 double* per_k_dev, int64_t per_k_stride, double* hist0_dev, uint16_t* assign0_dev, cudaStream_t st) {
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
      if (int rc = launch_tiles_epilogue(h, tp, nf, entropy_dev + f0, per_k_dev ? per_k_dev + f0 : nullptr, per_k_stride,
                                         hist0_dev ? hist0_dev + f0 * T0 : nullptr, st))
        return rc;
      continue;
    }
    if (int rc = launch_stream(h, in, dtype, nf, U, assign0_dev ? assign0_dev + f0 * U : nullptr, false, st)) return rc;
    if (int rc = launch_epilogue(h, nf, U, entropy_dev + f0, per_k_dev ? per_k_dev + f0 : nullptr, per_k_stride,
                                 hist0_dev ? hist0_dev + f0 * T0 : nullptr, st))
      return rc;
  }
  return VET_OK;
}

}  // namespace

extern "C" int vet_spatial(vet_handle* h, const void* packed_dev, int dtype, int64_t F, int64_t U, double* entropy_dev,
                           double* per_k_dev, double* hist0_dev, uint16_t* assign0_dev, void* stream) {
  if (!h || F < 0 || U < 0 || (dtype != VET_F32 && dtype != VET_F64)) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (F == 0) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty vector dictionary");  // EU:168-169
  if (!packed_dev || !entropy_dev) return fail(VET_ERR_INVALID_ARG, "null buffer");
  DeviceGuard guard(h->device);
  h->call_frames = F;
  cudaStream_t st = (cudaStream_t)stream;
  if (h->direct_only) return spatial_direct(h, packed_dev, dtype, F, U, entropy_dev, per_k_dev, hist0_dev, assign0_dev, st);
  vet_handle::GraphSlot key;
  key.api = 1;
  key.dtype = dtype;
  key.F = F;
  key.U = U;
  key.ptr[0] = packed_dev;
  key.ptr[1] = entropy_dev;
  key.ptr[2] = per_k_dev;
  key.ptr[3] = hist0_dev;
  key.ptr[4] = assign0_dev;
  return run_graphed(h, key, st, [&] { return spatial_core(h, packed_dev, dtype, F, U, entropy_dev, per_k_dev, F, hist0_dev, assign0_dev, st); });
}

namespace {

// TransitionEntropyAnalyzer.compute_entropy on F resident frames (table regimes).  per_k rows are `per_k_stride` apart:
// a caller that walks a longer video in pieces (the host-buffer pipeline) passes the row count of the whole video.
int transition_core(vet_handle* h, const void* packed_dev, int dtype, int64_t F, int64_t U, double* entropy_dev,
                    double* per_k_dev, int64_t per_k_stride, int32_t* prev_count0_dev, uint16_t* pairs0_dev, int mode,
                    cudaStream_t st) {
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
    a.per_k_stride = per_k_stride;
    a.prev_count0 = prev_count0_dev ? prev_count0_dev + f0 * T0 : nullptr;
    a.pairs0 = pairs0_dev ? pairs0_dev + f0 * U * 2 : nullptr;
    a.mode = mode;
    a.flags = h->d_flags;
    a.nvalid = h->d_nvalid;  // of this batch, written by the streaming kernel above
    if (int rc = launch_transition(h, a, nf - 1, U, h->maxT, st)) return rc;
    if (nf == F - f0) break;
  }
  return VET_OK;
}

}  // namespace

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
  vet_handle::GraphSlot key;
  key.api = 2;
  key.dtype = dtype;
  key.mode = mode;
  key.F = F;
  key.U = U;
  key.ptr[0] = packed_dev;
  key.ptr[1] = entropy_dev;
  key.ptr[2] = per_k_dev;
  key.ptr[3] = prev_count0_dev;
  key.ptr[4] = pairs0_dev;
  return run_graphed(h, key, st, [&] {
    return transition_core(h, packed_dev, dtype, F, U, entropy_dev, per_k_dev, F - 1, prev_count0_dev, pairs0_dev, mode, st);
  });
}

namespace {

// Both analyzers on F resident frames with one read of the input (table regimes, F >= 2); per_k rows sp_stride /
// tr_stride apart.
int analyze_core(vet_handle* h, const void* packed_dev, int dtype, int64_t F, int64_t U, double* sp_entropy_dev,
                 double* sp_per_k_dev, int64_t sp_stride, double* hist0_dev, uint16_t* assign0_dev, double* tr_entropy_dev,
                 double* tr_per_k_dev, int64_t tr_stride, int32_t* prev_count0_dev, uint16_t* pairs0_dev, int mode,
                 cudaStream_t st) {
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
    // The two consumers of the streaming kernel's outputs are independent: the transition kernel goes first on
    // the caller's stream (its persistent CTAs take every SM), the spatial epilogue on a side stream fills the SMs
    // that fall idle in the transition kernel's last, partial round.  VET_OPT_ANALYZE_OVERLAP=0 runs them in sequence.
    const bool side = h->opt[VET_OPT_ANALYZE_OVERLAP] && nf >= 2;
    cudaStream_t se = side ? h->s_side : st;
    if (side) {
      if (!h->ev_fork) {
        VET_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        VET_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
      }
      VET_CUDA(cudaEventRecord(h->ev_fork, st));
      VET_CUDA(cudaStreamWaitEvent(h->s_side, h->ev_fork, 0));
    } else {
      if (int rc = launch_epilogue(h, nf, U, sp_entropy_dev + f0, sp_per_k_dev ? sp_per_k_dev + f0 : nullptr, sp_stride,
                                   hist0_dev ? hist0_dev + f0 * T0 : nullptr, st))
        return rc;
    }
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
      a.per_k_stride = tr_stride;
      a.prev_count0 = prev_count0_dev ? prev_count0_dev + f0 * T0 : nullptr;
      a.pairs0 = pairs0_dev ? pairs0_dev + f0 * U * 2 : nullptr;
      a.mode = mode;
      a.flags = h->d_flags;
      a.nvalid = h->d_nvalid;  // of this batch, written by the streaming kernel above
      if (int rc = launch_transition(h, a, nf - 1, U, h->maxT, st)) return rc;
    }
    if (side) {
      if (int rc = launch_epilogue(h, nf, U, sp_entropy_dev + f0, sp_per_k_dev ? sp_per_k_dev + f0 : nullptr, sp_stride,
                                   hist0_dev ? hist0_dev + f0 * T0 : nullptr, se))
        return rc;
      VET_CUDA(cudaEventRecord(h->ev_join, se));
      VET_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
    }
    if (nf == F - f0) break;
  }
  return VET_OK;
}

}  // namespace

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
  h->call_frames = F;
  cudaStream_t st = (cudaStream_t)stream;
  if (h->direct_only || F == 1) {  // no shared pass to gain: run the two stages one after the other
    if (int rc = vet_spatial(h, packed_dev, dtype, F, U, sp_entropy_dev, sp_per_k_dev, hist0_dev, assign0_dev, stream)) return rc;
    return vet_transition(h, packed_dev, dtype, F, U, tr_entropy_dev, tr_per_k_dev, prev_count0_dev, pairs0_dev, mode, stream);
  }
  vet_handle::GraphSlot key;
  key.api = 3;
  key.dtype = dtype;
  key.mode = mode;
  key.F = F;
  key.U = U;
  key.ptr[0] = packed_dev;
  key.ptr[1] = sp_entropy_dev;
  key.ptr[2] = sp_per_k_dev;
  key.ptr[3] = hist0_dev;
  key.ptr[4] = assign0_dev;
  key.ptr[5] = tr_entropy_dev;
  key.ptr[6] = tr_per_k_dev;
  key.ptr[7] = prev_count0_dev;
  key.ptr[8] = pairs0_dev;
  return run_graphed(h, key, st, [&] {
    return analyze_core(h, packed_dev, dtype, F, U, sp_entropy_dev, sp_per_k_dev, F, hist0_dev, assign0_dev, tr_entropy_dev,
                        tr_per_k_dev, F - 1, prev_count0_dev, pairs0_dev, mode, st);
  });
}
