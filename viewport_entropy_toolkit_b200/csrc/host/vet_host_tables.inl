// Table construction at handle creation: lattices -> unit centres, cell -> tile LUTs (k_nearest), FOV weight
// columns and the grouped dense blocks of the FP64 weighted histogram, the grid-code table of the naive tiling.
// Textual fragment of vet_b200.cu.
namespace {

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

// Ranking of the tile-index deltas of a transition for k_transition4 (vet_transition4.cuh).  A user moves a few cells
// per frame, so its (prev, cur) tiles are the tiles of two nearby cells.  Over all cell pairs up to 4 cells apart,
// weighted by a Gaussian of their distance, the deltas lut[b] - lut[a] are scored per BAND of 2^s consecutive
// previous tiles (tiles are numbered by latitude, and the index offsets of a tile's lattice neighbours change slowly
// with latitude); the 15 heaviest deltas of a band get the ranks 1..15, rank 0 = staying in the tile.  The ranking
// decides which transitions take the table path of the kernel -- all others go through its list of unranked users,
// so no result depends on it.  (Measured on random-walk trajectories: 8-80 unranked users per 100k.)
int build_delta_ranks(vet_handle* h, TileSet& t) {
  const int T = t.T, W1 = h->W + 1, H1 = h->H + 1;
  if (h->global_tables || h->direct_only || 15 * T * 4 >= 0xFFFF || T < 2) return VET_OK;  // rank * T * 4 must fit 16 bits
  int s = 0;
  while (((T + (1 << s) - 1) >> s) > 32) ++s;
  const int bands = (T + (1 << s) - 1) >> s;
  const int win = std::min(T - 1, 160), L = 2 * win + 1;
  std::vector<double> score((size_t)bands * L, 0.0);
  for (int py = 0; py < H1; ++py)
    for (int px = 0; px < W1; ++px) {
      const int c0 = t.h_lut[(size_t)py * W1 + px];
      for (int dy = -4; dy <= 4; ++dy)
        for (int dx = -4; dx <= 4; ++dx) {
          const int nx = px + dx, ny = py + dy;
          if ((dx == 0 && dy == 0) || nx < 0 || ny < 0 || nx >= W1 || ny >= H1) continue;
          const int d = (int)t.h_lut[(size_t)ny * W1 + nx] - c0;
          if (d && d >= -win && d <= win) score[(size_t)(c0 >> s) * L + (d + win)] += std::exp(-(double)(dx * dx + dy * dy) / 4.5);
        }
    }
  // per band a row of 2 win + 2 entries: byte offset rank * T * 4 of the rank's row in a [16][T] table, 0xFFFF = no rank
  // (always so in the last entry, where the kernel sends every delta outside the window)
  const int Lp = L + 1;
  std::vector<uint16_t> drank((size_t)bands * Lp + 16, 0xFFFF);
  std::vector<int> rdelta((size_t)bands * 16, 0);
  for (int b = 0; b < bands; ++b) {
    std::vector<int> order;
    for (int i = 0; i < L; ++i)
      if (score[(size_t)b * L + i] > 0.0) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return score[(size_t)b * L + x] > score[(size_t)b * L + y]; });
    drank[(size_t)b * Lp + win] = 0;
    for (int k = 1; k < 16 && k - 1 < (int)order.size(); ++k) {
      drank[(size_t)b * Lp + order[k - 1]] = (uint16_t)(k * T * 4);
      rdelta[(size_t)b * 16 + k] = order[k - 1] - win;
    }
  }
  t.band_shift = s;
  t.win = win;
  t.bands = bands;
  if (int rc = upload(&t.d_drank, drank.data(), drank.size())) return rc;
  return upload(&t.d_rank_delta, rdelta.data(), rdelta.size());
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
  if (int rc = build_delta_ranks(h, t)) return rc;
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
    if (int rc = build_weight_groups<vet::WhistWide>(h, t, ptr, unit)) return rc;
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
  cudaFree(t.d_drank);
  cudaFree(t.d_rank_delta);
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

}  // namespace
