// Launchers of the transition stage: k_transition3 per tile count, k_transition2 / k_transition behind it.
// Textual fragment of vet_b200.cu.
namespace {

// Sizes the (prev,cur) pair tables, picks the shared- or global-memory variant and launches
// k_transition for `rows` frame pairs.
// Global (prev,cur) pair tables of k_transition2's global mode and of k_transition<false>: [blocks][4][cap] words, allocated
// and laid out only when one of those kernels is about to run (620 MB at 201 tiles and two CTAs per SM, ~10 GB at 1001).
int ensure_global_tables(vet_handle* h, int blocks, uint32_t cap, cudaStream_t st) {
  const size_t words = (size_t)blocks * 4 * cap;
  if (h->tables_words < words || h->tables_cap != cap) {
    ++g_scratch_epoch;  // new layout of the global pair tables: captured graphs were laid out for the old one
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
    ++g_scratch_epoch;
    for (int b = h->tables_blocks; b < blocks; ++b) {
      VET_CUDA(cudaMemsetAsync(h->d_tables + (size_t)b * 4 * cap, 0xFF, (size_t)cap * 8, st));
      VET_CUDA(cudaMemsetAsync(h->d_tables + (size_t)b * 4 * cap + 2 * (size_t)cap, 0, (size_t)cap * 4, st));
    }
    h->tables_blocks = blocks;
  }
  return VET_OK;
}

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
    bool any_global = false;
    for (int k = 0; k < a.K; ++k) any_global = any_global || A2.mode[k] != vet::kTrDense;  // the shared hash falls back to it
    if (any_global) {
      if (int rc = ensure_global_tables(h, blocks, a.cap, st)) return rc;
      A2.t.g_tables = h->d_tables;
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

// k_transition3c: `rows` clusters of S CTAs, one frame pair each.
template <int LW>
int launch_t3c(const vet::Transition3Args& A, int rows, int S, size_t smem, cudaStream_t st, int* max_clusters) {
  auto* kern = vet::k_transition3c<LW>;
  VET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(rows * S));
  cfg.blockDim = dim3(vet::kT3Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)S;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (max_clusters) {  // query only: how many clusters of this shape are co-resident
    cfg.gridDim = dim3((unsigned)S);
    if (cudaOccupancyMaxActiveClusters(max_clusters, kern, &cfg) != cudaSuccess) {
      cudaGetLastError();
      *max_clusters = 0;
    }
    return VET_OK;
  }
  VET_CUDA(cudaLaunchKernelEx(&cfg, kern, A));
  return VET_OK;
}

int launch_transition3c(int lw, const vet::Transition3Args& A, int rows, int S, size_t smem, cudaStream_t st,
                        int* max_clusters = nullptr) {
  switch (lw) {
    case vet::kLutS8: return launch_t3c<vet::kLutS8>(A, rows, S, smem, st, max_clusters);
    case vet::kLutS16: return launch_t3c<vet::kLutS16>(A, rows, S, smem, st, max_clusters);
    case vet::kLutIdentity: return launch_t3c<vet::kLutIdentity>(A, rows, S, smem, st, max_clusters);
    default: return launch_t3c<vet::kLutG16>(A, rows, S, smem, st, max_clusters);
  }
}

// Co-resident clusters of k_transition3c<lw> with S CTAs and `smem` bytes each (cached per handle).
int t3c_max_clusters(vet_handle* h, int lw, int S, size_t smem) {
  for (const auto& c : h->t3c_occ)
    if (c.lw == lw && c.S == S && c.smem == smem) return c.n;
  int n = 0;
  vet::Transition3Args none{};
  if (launch_transition3c(lw, none, 1, S, smem, nullptr, &n) != VET_OK) n = 0;
  h->t3c_occ.push_back({lw, S, smem, n});
  return n;
}

// k_transition4 (one pass over rows of tile ids: [3][16][T] tables + list of unranked users): shared-memory layout
struct Plan4 {
  bool ok = false;
  int threads = 512, ctas_per_sm = 1;
  size_t tab_off = 0, key_off = 0, term_off = 0, rdelta_off = 0, drank_off = 0, smem = 0;
};
Plan4 plan_transition4(const vet_handle* h, const TileSet& ts, int64_t rows) {
  Plan4 p;
  if (!ts.d_drank) return p;
  const size_t T = (size_t)ts.T;
  const size_t budget = h->smem_optin - kStaticSmemSlack;
  p.tab_off = 0;
  p.key_off = (size_t)3 * (vet::kT4Ranks * T + 4) * 4;  // three [16][T] tables with a dummy entry (+ padding) each
  p.term_off = p.key_off + (size_t)vet::kT4OvfCap * 8;
  p.rdelta_off = p.term_off + T * 8;
  p.drank_off = (p.rdelta_off + (size_t)ts.bands * vet::kT4Ranks * 4 + 15) & ~(size_t)15;
  p.smem = (p.drank_off + (size_t)ts.bands * (2 * ts.win + 2) * 2 + 15) & ~(size_t)15;
  if (p.smem > budget) return p;
  p.ok = true;
  // 512 threads when several CTAs fit an SM (228 KB of shared memory; 64 registers per thread: 1024 threads), else
  // one CTA of 1024
  // With few pairs per SM the grid must not depend on where the CTAs land (two CTAs with two pairs each on one SM,
  // two with one pair each on the next): one CTA per SM then, every SM takes the same number of pairs and the rest
  // goes to clusters.  With many pairs two smaller CTAs overlap each other's per-pair serial parts.
  const int fit = (int)(((size_t)228 * 1024) / (p.smem + 1024));
  if (fit >= 2 && rows >= (int64_t)8 * h->sm_count) {
    p.threads = 512;
    p.ctas_per_sm = 2;
  } else {
    p.threads = 1024;  // (768 threads with 85 registers each, no spills, measured slower: 0.60 vs 0.57 ms on configs[4])
    p.ctas_per_sm = 1;
  }
  return p;
}

// Co-resident clusters of k_transition4<true> with S CTAs of `threads` threads and `smem` bytes each (cached per handle).
int t4c_max_clusters(vet_handle* h, int S, int threads, size_t smem) {
  for (const auto& c : h->t3c_occ)
    if (c.lw == -4 - threads && c.S == S && c.smem == smem) return c.n;
  int n = 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)S);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)S;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (cudaOccupancyMaxActiveClusters(&n, vet::k_transition4<true>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  h->t3c_occ.push_back({-4 - threads, S, smem, n});
  return n;
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
  a.cap = cap;
  a.g_tables = nullptr;  // the global pair tables are only made for the kernels that use them (ensure_global_tables)
  // VET_OPT_TRANSITION_KERNEL pins k_transition (1) or k_transition2 (2): the parity suite runs them against the two-pass kernels
  const bool force_v1 = h->opt[VET_OPT_TRANSITION_KERNEL] == 1;
  const bool force_v2 = h->opt[VET_OPT_TRANSITION_KERNEL] == 2;
  // two-pass kernel, one launch per tile count; rows it cannot hold (hash overflow) are flagged in d_redo
  // and recomputed by k_transition2 below
  // VET_OPT_CLUSTER_TAIL: 0 keeps every frame pair on k_transition3, 2 takes the cluster kernel whenever it can
  // run, also where it does not pay (small frames; tests)
  const int cluster_tail = h->opt[VET_OPT_CLUSTER_TAIL];
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
      pl.tab_off = (T * vet::kT3TileBytes + 16 + 15) & ~(size_t)15;
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
      // VET_OPT_T3_PAIR_SCRATCH keeps the pair scratch with the identity table too (variant compared by the tests)
      const bool keep_scratch = h->opt[VET_OPT_T3_PAIR_SCRATCH] != 0;
      bool need_scratch = keep_scratch;  // packed (prev | cur << 16) per user, one row per CTA: not with the identity table
      for (int k = 0; k < a.K; ++k) need_scratch = need_scratch || (plan[k].mode >= 0 && plan[k].lw != vet::kLutIdentity);
      if (need_scratch)
        if (int rc = grow((void**)&h->d_pairs, &h->pairs_bytes, (size_t)std::max(blocks, blocks3) * U * 4)) return rc;
      if (int rc = grow((void**)&h->d_redo, &h->redo_bytes, (size_t)rows * 4)) return rc;
      VET_CUDA(cudaMemsetAsync(h->d_redo, 0, (size_t)rows * 4, st));
      // one-pass kernel in front (VET_OPT_TRANSITION_KERNEL auto): per tile count a counter and the flags of the pairs
      // it leaves to the two-pass kernels
      const bool use_t4 = h->opt[VET_OPT_TRANSITION_KERNEL] == 0;
      const size_t t4_zero = (size_t)a.K * 16 + (size_t)a.K * rows * 4;
      if (use_t4) {
        if (int rc = grow(&h->d_t4, &h->t4_bytes, t4_zero)) return rc;
        VET_CUDA(cudaMemsetAsync(h->d_t4, 0, t4_zero, st));
      }
      double* per_k = a.per_k;
      int64_t stride = a.per_k_stride;
      if (a.K > 1 && !per_k) {
        if (int rc = grow((void**)&h->d_trk, &h->trk_bytes, (size_t)a.K * rows * 8)) return rc;
        per_k = h->d_trk;
        stride = rows;
      }
      // one-pass kernel plans of every tile count, and the tile-id rows they read where the rows hold cell ids (several
      // tile counts): relabelled here, up to kRelabelGroup tile counts per pass over the cell ids
      Plan4 plan4[vet::kMaxTileCounts];
      const uint16_t* rows_of[vet::kMaxTileCounts] = {};
      {
        int need[vet::kMaxTileCounts], nneed = 0;
        for (int k = 0; k < a.K; ++k) {
          const TileSet& tsk = h->ts[std::min(k, h->K - 1)];
          plan4[k] = (use_t4 && plan[k].mode >= 0 && k < h->K && tsk.T == a.T[k]) ? plan_transition4(h, tsk, rows) : Plan4{};
          if (plan4[k].ok && plan[k].lw != vet::kLutIdentity && a.cell16) need[nneed++] = k;
        }
        const int64_t n = a.F * U;
        const size_t slot = ((size_t)n * 2 + 16 + 15) & ~(size_t)15;
        const int cp = (int)((h->C + 1 + 7) & ~(int64_t)7);
        const int group = (int)std::min<size_t>(vet::kRelabelGroup, (h->smem_optin - kStaticSmemSlack) / ((size_t)cp * 2));
        if (nneed)
          if (int rc = grow((void**)&h->d_rows, &h->rows_bytes, slot * (group >= 2 ? nneed : 1))) return rc;
        for (int i0 = 0; group >= 2 && i0 < nneed; i0 += group) {
          vet::RelabelArgs R{};
          R.cells = a.cell16;
          R.G = std::min(group, nneed - i0);
          R.C = (int)h->C;
          R.cp = cp;
          R.n = n;
          for (int g = 0; g < R.G; ++g) {
            R.lut[g] = a.lut[need[i0 + g]];
            R.tiles[g] = (uint16_t*)((char*)h->d_rows + slot * (i0 + g));
            rows_of[need[i0 + g]] = R.tiles[g];
          }
          const size_t smem = (size_t)R.G * cp * 2;
          VET_CUDA(cudaFuncSetAttribute(vet::k_relabel_rows_group, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          LaunchTimer lt(h, VET_KERNEL_TRANSITION, st);
          vet::k_relabel_rows_group<<<(unsigned)std::min<int64_t>((n / 8 + vet::kRelabelThreads - 1) / vet::kRelabelThreads + 1, h->sm_count),
                                      vet::kRelabelThreads, smem, st>>>(R);
          VET_CUDA(cudaGetLastError());
        }
      }
      bool any_hash = false;
      for (int k = 0; k < a.K; ++k) {
        Plan pl = plan[k];
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
        A3.pair_scratch = (pl.lw == vet::kLutIdentity && !keep_scratch) ? nullptr : h->d_pairs;
        A3.redo = h->d_redo;
        A3.flags = a.flags;
        A3.nvalid = h->opt[VET_OPT_T3_ASSUME_MISSING] ? nullptr : a.nvalid;  // option: always test for missing users
        A3.ush = 3;  // granularity of the "early" bound bytes of the dense pass 2: (U - 1) >> ush <= 254
        while (((uint64_t)(U - 1) >> A3.ush) > 254) ++A3.ush;
        bool t4_done = false;
        // the k-th tile set of the handle (cell LUT or, with tile ids in the rows, the identity table over it)
        const TileSet& tsk = h->ts[std::min(k, h->K - 1)];
        const Plan4 p4 = plan4[k];
        if (p4.ok) {
          if (pl.lw != vet::kLutIdentity) {
            // several tile counts: the rows hold cell ids -> tile ids of this tile count first (one lookup per sample
            // instead of two per user and pair inside the kernels)
            if (!rows_of[k]) {  // tables of two tile counts do not fit shared memory together: one pass per tile count
              const int64_t n = a.F * U;
              LaunchTimer lt(h, VET_KERNEL_TRANSITION, st);
              vet::k_relabel_rows<<<(unsigned)std::min<int64_t>((n / 8 + 255) / 256 + 1, (int64_t)h->sm_count * 16), 256, 0, st>>>(
                  a.cell16, a.lut[k], n, h->d_rows);
              VET_CUDA(cudaGetLastError());
              rows_of[k] = h->d_rows;
            }
            A3.cell16 = rows_of[k];
            A3.lut_src = nullptr;
            A3.pair_scratch = keep_scratch ? h->d_pairs : nullptr;
            pl.lw = vet::kLutIdentity;
            pl.lut = nullptr;
            pl.smem = pl.lut_off;
          }
          if (A3.pairs0) {
            const int64_t n = rows * U;
            LaunchTimer lt(h, VET_KERNEL_TRANSITION, st);
            vet::k_pairs_from_rows<<<(unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16), 256, 0, st>>>(
                A3.cell16, a.F, U, reinterpret_cast<uint32_t*>(A3.pairs0));
            VET_CUDA(cudaGetLastError());
            A3.pairs0 = nullptr;
          }
          char* base = (char*)h->d_t4;
          uint32_t* count4 = (uint32_t*)(base + (size_t)k * 16);
          uint32_t* redo4 = (uint32_t*)(base + (size_t)a.K * 16 + (size_t)k * rows * 4);
          vet::Transition4Args A4{};
          A4.t = A3;
          A4.t.tab_off = (int)p4.tab_off;
          A4.t.pair_scratch = nullptr;
          A4.t.redo = redo4;
          A4.drank = tsk.d_drank;
          A4.rank_delta = tsk.d_rank_delta;
          A4.band_shift = tsk.band_shift;
          A4.win = tsk.win;
          A4.bands = tsk.bands;
          A4.rdelta_off = (int)p4.rdelta_off;
          A4.key_off = (int)p4.key_off;
          A4.drank_off = (int)p4.drank_off;
          A4.term_off = (int)p4.term_off;
          A4.redo_count = count4;
          A4.ovf_cap = h->opt[VET_OPT_T4_LIST_CAP] > 0 ? h->opt[VET_OPT_T4_LIST_CAP] : vet::kT4OvfCap;
          // The rows % SMs pairs left after the full rounds of one pair per SM go to clusters (the users of a pair split
          // over S CTAs, tables merged through distributed shared memory) instead of a last round on a few SMs -- from
          // ~130k users saved per CTA on (VET_OPT_CLUSTER_TAIL: 0 never, 2 whenever it can run).
          int64_t tail4 = 0;
          int S4 = 0;
          VET_CUDA(cudaFuncSetAttribute(vet::k_transition4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p4.smem));
          VET_CUDA(cudaFuncSetAttribute(vet::k_transition4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p4.smem));
          if (cluster_tail) {
            const int64_t rem = rows % h->sm_count;
            // clusters of 16 (non-portable size, where the occupancy query grants them) when 16 shares of a frame still
            // hold 32k users: 5 pairs of 1M users 77 -> 70 us
            const int Smax = U >= (int64_t)16 * 32768 ? 16 : 8;
            if (Smax == 16) VET_CUDA(cudaFuncSetAttribute(vet::k_transition4<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            for (int S = Smax; S >= 2 && rem > 0 && !S4; S >>= 1)
              if (U >= (int64_t)S * p4.threads * 8 && (cluster_tail == 2 || U * (S - 1) >= (int64_t)131072 * S) &&
                  rem <= t4c_max_clusters(h, S, p4.threads, p4.smem))
                S4 = S;
            if (S4) tail4 = rem;
          }
          const int64_t main4 = rows - tail4;
          if (main4 > 0) {
            vet::Transition4Args AM = A4;
            AM.t.F = main4 + 1;
            const int blocks4 = (int)std::min<int64_t>(main4, (int64_t)h->sm_count * p4.ctas_per_sm);
            LaunchTimer lt(h, VET_KERNEL_TRANSITION, st);
            vet::k_transition4<false><<<blocks4, p4.threads, p4.smem, st>>>(AM);
            VET_CUDA(cudaGetLastError());
          }
          if (tail4 > 0) {
            vet::Transition4Args AT = A4;
            AT.t.nvalid = A4.t.nvalid ? A4.t.nvalid + main4 : nullptr;
            AT.t.cell16 = A4.t.cell16 + main4 * U;
            AT.t.F = tail4 + 1;
            AT.t.out = A4.t.out + main4;
            AT.t.prev_count0 = A4.t.prev_count0 ? A4.t.prev_count0 + main4 * A4.t.T : nullptr;
            AT.t.redo = A4.t.redo + main4;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)(tail4 * S4));
            cfg.blockDim = dim3((unsigned)p4.threads);
            cfg.dynamicSmemBytes = p4.smem;
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)S4;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            LaunchTimer lt(h, VET_KERNEL_TRANSITION_TAIL, st);
            VET_CUDA(cudaLaunchKernelEx(&cfg, vet::k_transition4<true>, AT));
          }
          // the two-pass kernel below only takes the pairs the one-pass kernel flagged (its CTAs leave at once when none)
          A3.only_rows = redo4;
          A3.only_count = count4;
          t4_done = true;
        }
        // dense tables: the rows % SMs pairs left after the full rounds go to k_transition3c, one pair per
        // cluster of S CTAs (users split across the cluster) instead of one more, mostly idle, round
        int64_t tail_rows = 0;
        int tail_S = 0;
        if (pl.mode == vet::kT3Dense && cluster_tail && !t4_done) {
          const int64_t rem = rows % h->sm_count;
          // a pair costs ~0.33 ns per user on one CTA; the cluster barriers, the merge of the tables and the extra
          // launch ~20-40 us: worth it from ~130k users saved per CTA (measured neutral to slower at 100k users)
          for (int S = 8; S >= 2 && rem > 0 && !tail_S; S >>= 1)
            if (U >= (int64_t)S * vet::kT3Threads * 8 && (cluster_tail == 2 || U * (S - 1) >= (int64_t)131072 * S) &&
                rem <= t3c_max_clusters(h, pl.lw, S, pl.smem))
              tail_S = S;
          if (tail_S) tail_rows = rem;
        }
        if (tail_rows) {
          vet::Transition3Args AT = A3;
          const int64_t r0 = rows - tail_rows;
          AT.nvalid = A3.nvalid ? A3.nvalid + r0 : nullptr;
          AT.cell16 = A3.cell16 + r0 * U;
          AT.F = tail_rows + 1;
          AT.out = A3.out + r0;
          AT.prev_count0 = A3.prev_count0 ? A3.prev_count0 + r0 * A3.T : nullptr;
          AT.pairs0 = A3.pairs0 ? A3.pairs0 + r0 * U * 2 : nullptr;
          A3.F = r0 + 1;
          LaunchTimer lt(h, VET_KERNEL_TRANSITION_TAIL, st);
          if (int rc = launch_transition3c(pl.lw, AT, (int)tail_rows, tail_S, pl.smem, st)) return rc;
          if (r0 == 0) continue;
        }
        const int blocks3k = (int)std::min<int64_t>(A3.F - 1, h->sm_count);
        LaunchTimer lt(h, t4_done ? VET_KERNEL_TRANSITION_TAIL : VET_KERNEL_TRANSITION, st);
#define VET_T3(MODE, LW)                                                                                              \
  do {                                                                                                                \
    VET_CUDA(cudaFuncSetAttribute(vet::k_transition3<MODE, LW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget)); \
    vet::k_transition3<MODE, LW><<<blocks3k, vet::kT3Threads, pl.smem, st>>>(A3);                                     \
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
    if (int rc = ensure_global_tables(h, blocks, cap, st)) return rc;
    a.g_tables = h->d_tables;
    VET_CUDA(cudaFuncSetAttribute(vet::k_transition<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(tile_bytes + 64)));
    LaunchTimer lt(h, VET_KERNEL_TRANSITION, st);
    vet::k_transition<false><<<blocks, 512, tile_bytes + 64, st>>>(a, Tmax);
  }
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

}  // namespace
