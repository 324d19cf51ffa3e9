// Launchers of the spatial stage: streaming kernels (TMA / simple / global-table / direct tile histograms),
// the weighted histogram on the tensor cores (planes, weight slices, tensor maps) or the FP64 pipe, entropy rows.
// Textual fragment of vet_b200.cu.
namespace {


size_t stream_smem_bytes(const vet_handle* h) { return (size_t)h->Cpad * 4 + (size_t)h->C * 2 + 16; }
size_t stream_tma_smem_bytes(const vet_handle* h, bool lut8) {
  return (size_t)vet::kStages * vet::kStageBytes + (size_t)h->Cpad * 4 + (size_t)h->C * (lut8 ? 1 : 2) + 32;
}
size_t epilogue_smem_bytes(const vet_handle* h) {
  return (size_t)h->Cpad * 4 + (size_t)h->maxT * 8 + (size_t)h->maxT * 4 + 16;
}
bool use_tma_stream(const vet_handle* h, const void* packed) {
  if (h->opt[VET_OPT_STREAM_KERNEL] == 1) return false;
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

// Grid of a persistent kernel whose equal-cost items are dealt round-robin: the fewest CTAs that need the same
// number of rounds as one CTA per SM would.
int balanced_grid(int64_t items, int sm_count) {
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(items, sm_count));
  const int64_t rounds = (items + blocks - 1) / blocks;
  return (int)((items + rounds - 1) / rounds);
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
    // weighted, and the grid fits shared memory as 16-bit counters: privatised histogram per (frame, chunk)
    const size_t smem16 = (size_t)((h->Cpad + 1) / 2) * 4;
    const bool no16 = h->opt[VET_OPT_STREAM_KERNEL] == 3;
    if (h->use_weight && !no16 && smem16 + kStaticSmemSlack <= h->smem_optin) {
      vet::StreamArgs b = a;
      const int64_t cpf = std::max<int64_t>((U + 65534) / 65535, std::min<int64_t>((U + 16383) / 16384, ((int64_t)h->sm_count * 8 + F - 1) / F));
      b.chunk_users = (U + cpf - 1) / cpf;
      b.chunks_per_frame = (int)((U + b.chunk_users - 1) / std::max<int64_t>(b.chunk_users, 1));
      if (b.chunks_per_frame > 1) {
        VET_CUDA(cudaMemsetAsync(h->d_cnt, 0, (size_t)F * h->Cpad * 4, st));
        VET_CUDA(cudaMemsetAsync(h->d_nvalid, 0, (size_t)F * 4, st));
      }
      const int fblocks = balanced_grid(F * b.chunks_per_frame, h->sm_count);
      LaunchTimer lt(h, VET_KERNEL_STREAM, st);
      if (dtype == VET_F32) {
        VET_CUDA(cudaFuncSetAttribute(vet::k_stream_frame16<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
        vet::k_stream_frame16<float><<<fblocks, 1024, smem16, st>>>(b);
      } else {
        VET_CUDA(cudaFuncSetAttribute(vet::k_stream_frame16<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
        vet::k_stream_frame16<double><<<fblocks, 1024, smem16, st>>>(b);
      }
      VET_CUDA(cudaGetLastError());
      return VET_OK;
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
  // Items are dealt round-robin and cost the same, so the kernel takes ceil(items / CTAs) rounds: use the fewest
  // CTAs that still need that many rounds (3600 frames: 144 CTAs x 25 instead of 148 CTAs of which 100 do only 24).
  // Measured on configs[2]: 0.790 -> 0.782 ms (97.6 -> 98.6 % of the HBM peak).
  const int blocks = balanced_grid(items, h->sm_count);
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
  if (h->opt[VET_OPT_STREAM_KERNEL] == 1 || h->opt[VET_OPT_STREAM_KERNEL] == 2) return p;
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
  p.smem = (size_t)vet::kStages * vet::kStageBytes + (size_t)2 * ((soff + 3) & ~3) * 4 + off + 16;  // two copies of the histograms
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
  const int blocks = balanced_grid(F * a.chunks_per_frame, h->sm_count);
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

// Row schedule of k_entropy_rows (see EntropyRowsPlan): G frames per block, their G x K rows dealt longest-first to the
// least loaded of the 8 warps; the G with the best balance wins (ties: fewer frames per block, more blocks).
vet::EntropyRowsPlan plan_entropy_rows(const vet_handle* h, int64_t F) {
  vet::EntropyRowsPlan best{};
  double best_cost = 0.0;
  for (int G : {1, 2, 4, 8}) {
    if (G > 1 && (G * h->K > 8 * 16 || F < (int64_t)G * h->sm_count)) continue;
    vet::EntropyRowsPlan pl{};
    pl.G = G;
    std::vector<std::pair<int, int>> rows;  // (cost, g * 16 + k)
    for (int g = 0; g < G; ++g)
      for (int k = 0; k < h->K; ++k) rows.push_back({(h->ts[k].T + 31) / 32 + 1, g * 16 + k});
    std::stable_sort(rows.begin(), rows.end(), [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.first > y.first; });
    int load[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool ok = true;
    for (const auto& r : rows) {
      int w = 0;
      for (int j = 1; j < 8; ++j)
        if (load[j] < load[w]) w = j;
      if (pl.nrow[w] >= 16) {
        ok = false;
        break;
      }
      pl.row_g[w][pl.nrow[w]] = (unsigned char)(r.second / 16);
      pl.row_k[w][pl.nrow[w]] = (unsigned char)(r.second % 16);
      pl.nrow[w]++;
      load[w] += r.first;
    }
    if (!ok) continue;
    const double cost = (double)*std::max_element(load, load + 8) / G;  // block time per frame
    if (best.G == 0 || cost < best_cost - 1e-9) {
      best = pl;
      best_cost = cost;
    }
  }
  return best;
}

// Entropies of F frames from their histogram rows.  Several tile counts: k_entropy_frames (one frame per block and step,
// its rows spread over the 8 warps: vet_whist.cuh).  One tile count: k_entropy_rows, whose 8 warps take one row each of 8
// frames -- nothing to balance there, and one round of blocks instead of three at 3600 frames (headline step: 0.862 ms
// against 0.868 with k_entropy_frames); also when the rows of a frame do not fit 40 KB of shared memory.
int launch_entropy_kernel(vet_handle* h, vet::EntropyRowsArgs& e, int64_t F, cudaStream_t st) {
  vet::EntropyFramesPlan fp{};
  for (int k = 0; k < e.K; ++k) {
    const int units = (e.T[k] + 31) / 32;
    fp.uoff[k + 1] = fp.uoff[k] + units;
    fp.soff[k + 1] = fp.soff[k] + units * 32;
  }
  fp.nunits = fp.uoff[e.K];
  const size_t smem = (size_t)fp.soff[e.K] * 16;
  if (e.K >= 2 && smem <= 40 * 1024) {  // 48 KB without opt-in, static shared memory included
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(F, (int64_t)h->sm_count * 8));
    vet::k_entropy_frames<<<blocks, 256, smem, st>>>(e, fp);
  } else {
    const vet::EntropyRowsPlan pl = plan_entropy_rows(h, F);
    const int blocks = (int)std::min<int64_t>((F + pl.G - 1) / pl.G, (int64_t)h->sm_count * 8);
    vet::k_entropy_rows<<<blocks, 256, 0, st>>>(e, pl);
  }
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
  return launch_entropy_kernel(h, e, F, st);
}

// k_whist for tile count k over F frames of the cell histogram `cnt` -> hist[F,T_k]
int launch_whist(vet_handle* h, int k, int64_t F, const uint32_t* cnt, double* hist, cudaStream_t st) {
  TileSet& t = h->ts[k];
  const int frames_per_cta = vet::WhistWide::kFramesPerCta;
  const size_t wh_smem = (size_t)vet::kWhStages * vet::WhistWide::kChunkBytes;
  const int64_t fblocks = (F + frames_per_cta - 1) / frames_per_cta;
  const int blocks = (int)std::min<int64_t>(fblocks * t.G, h->sm_count);
  if (t.sched_F != F || t.sched_blocks != blocks) {
    ++g_scratch_epoch;                    // captured graphs of other frame counts read the schedule this one replaces
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

bool use_whist_i8(const vet_handle* h, int64_t F, int64_t U) {
  // 0 = heuristic, 1 = always the FP64 kernel, 2 = the tensor-core kernel whenever it applies
  const int impl = h->opt[VET_OPT_WEIGHTED_KERNEL];
  if (impl == 1 || !encode_tiled_fn()) return false;
  if (U * 255 >= ((int64_t)1 << 31)) return false;  // int32 accumulators: D <= 255 * sum(count plane) <= 255 U
  if ((size_t)vet::kI8SmemBytes + kStaticSmemSlack > h->smem_optin) return false;
  if (impl == 2) return true;
  // A CTA of the tensor-core kernel takes ~70 us over all cells of 128 frames x 48 tiles, split over up to 8 CTAs
  // when the batch has few such tiles (launch_whist_i8); the FP64 kernel costs ~0.14 us per frame at 201 tiles.
  // From a few hundred frames on the tensor cores win (0.07 vs 0.50 ms at 3600 frames).
  // decided per API call, not per internal batch, so that a short last batch does not switch kernels
  if (std::max(F, h->call_frames) < kI8MinFrames) return false;
  // and only where the weight quantisation keeps the entropies inside the stated tolerance (i8_ok, build_i8_tables)
  vet_handle* hm = const_cast<vet_handle*>(h);
  for (auto& t : hm->ts) {
    if (build_i8_tables(hm, t) != VET_OK) return false;
    if (!t.i8_ok) return false;
  }
  return true;
}

// Quantised weight slices of tile set t: row nb*240 + s*48 + j of [i8_blocks*240, kp] holds slice s of
// q = rint(w(cell, nb*48+j) * 2^39) for every cell (q = 1 for a positive weight below half a quantum, so that the
// support of a histogram row -- the tiles with d < fov/2, EU:133 -- is the FP64 kernel's); plus the K-block range of
// every N block.
// Error of the quantisation, |dw| < 2^-39 =: delta per (cell, tile) entry: a user in cell c adds s_c = sum_t w(c,t) to
// the total of its frame and at most n_c delta (n_c tiles in its FOV) of error, so sum_t |dh_t| / total <= rho :=
// max_c n_c delta / s_c whatever the users do.  With dH/dh_t = -(log2 p_t + H) / total and p_t >= 2^-39 / (U T):
// |dH| <= rho (39 + log2(U T) + log2 T) <= 77 rho, i.e. |dHn| <= 77 rho / log2 T on the normalised entropy (EU:201-209).
// The automatic dispatch takes the tensor-core kernel only when that is <= 2.5e-10 (i8_ok): relative 1e-9 for every
// frame with Hn >= 0.25 -- fov=90, pf=2, 201 tiles: rho = 1.1e-11, bound 1.1e-10; measured differences are ~5e-14.
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
        uint64_t q = (uint64_t)std::llrint(std::ldexp(w_val[e], vet::kI8FracBits));
        if (!q && w_val[e] > 0.0) q = 1;  // keeps the tile in the support of the row
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
  {
    std::vector<double> s_c(h->C, 0.0);
    std::vector<uint32_t> n_c(h->C, 0);
    for (int tile = 0; tile < T; ++tile)
      for (uint32_t e = col_ptr[tile]; e < col_ptr[tile + 1]; ++e)
        if (w_val[e] > 0.0) {
          s_c[cell_idx[e]] += w_val[e];
          n_c[cell_idx[e]]++;
        }
    double rho = 0.0;
    const double delta = std::ldexp(1.0, -vet::kI8FracBits);
    for (int64_t c = 0; c < h->C; ++c)
      if (n_c[c]) rho = std::max(rho, n_c[c] * delta / s_c[c]);
    t.i8_rho = rho;
    t.i8_ok = 77.0 * rho / std::log2((double)std::max(T, 2)) <= 2.5e-10;
  }
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
  ++g_scratch_epoch;
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
  // few output tiles (20 at 450 frames x 201 tiles, each ~70 us of TMA + MMA steps over every cell): split the cells
  // over up to 8 CTAs per tile, at least 8 K blocks each
  const int tiles = fblocks * t.i8_blocks;
  const int kblocks = (int)(i8_kp(h) / vet::kI8BK);
  const int ksplit = std::max(1, std::min({vet::kI8MaxSplit, h->sm_count / std::max(tiles, 1), kblocks / 8}));
  if (ksplit > 1) {
    if (int rc = grow((void**)&h->d_i8acc, &h->i8acc_bytes, (size_t)tiles * ksplit * (2 * vet::kI8N * vet::kI8M) * 4)) return rc;
    a.part = (int*)h->d_i8acc;
  }
  a.ksplit = ksplit;
  const int grid = tiles * ksplit;
  for (int pass = 0; pass < 2; ++pass) {
    a.row_a = pass == 0 ? 0 : (int)(2 * h->plane_rows);
    a.row_b = (int)h->plane_rows;
    a.flag_b = pass == 0 ? i8_hi1(h) : nullptr;
    a.run_if = pass == 0 ? nullptr : i8_hi2(h);
    a.shift = pass == 0 ? 0 : 16;
    a.accumulate = pass;
    LaunchTimer lt(h, VET_KERNEL_EPILOGUE, st);
    vet::k_whist_i8<<<grid, vet::kI8Threads, vet::kI8SmemBytes, st>>>(h->tm_cnt, t.tm_w, a);
    if (ksplit > 1) vet::k_whist_i8_finish<<<dim3(vet::kI8TilesPerBlock, tiles), vet::kI8M, 0, st>>>(a);
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
  return launch_entropy_kernel(h, e, F, st);
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
