// Streaming kernels: packed samples -> per-frame cell histogram (+ tile assignment,
// + cell ids), and the per-frame epilogue cell histogram -> tile histograms ->
// normalised entropy (north_star stages 1-3 in the table regime: every sample is
// one of (W+1)(H+1) cells, so all per-tile work is done per CELL, not per sample).
#pragma once
#include "vet_common.cuh"

namespace vet {

struct StreamArgs {
  const void* packed;   // [F,U,3]
  int64_t F, U;
  int W, H, C;
  const uint16_t* lut0; // [C] nearest tile of tile_counts[0]
  uint16_t* assign0;    // [F,U] or null
  uint16_t* cell16;     // [F,U] or null (cell ids for the transition stage, C <= 65535)
  int32_t* cell32;      // [F,U] or null (same, any C)
  uint32_t* nvalid;     // [F] present users per frame (pre-zeroed when chunks_per_frame > 1)
  uint32_t* cnt;        // [F,cpad] per-frame cell histogram
  int cpad;             // row pitch of cnt in cells (C rounded up to a multiple of 4)
  int chunks_per_frame; // >1: cnt is pre-zeroed and flushed with atomics
  int64_t chunk_users;
  uint32_t* flags;
};

// Baseline streaming kernel: one (frame, user-chunk) per block iteration, one
// sample per thread iteration, shared-memory privatised cell histogram.
template <typename TIN>
__global__ void __launch_bounds__(1024, 1) k_stream_simple(StreamArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw);
  uint16_t* s_lut = reinterpret_cast<uint16_t*>(s_hist + a.cpad);
  const TIN* __restrict__ packed = static_cast<const TIN*>(a.packed);
  const float Wf = (float)a.W, Hf = (float)a.H;
  const bool want_assign = a.assign0 != nullptr;
  if (want_assign)
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) s_lut[c] = a.lut0[c];
  const int64_t items = a.F * a.chunks_per_frame;
  uint32_t bad = 0;
  __shared__ uint32_t s_nvalid;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t f = item / a.chunks_per_frame;
    const int64_t u0 = (item % a.chunks_per_frame) * a.chunk_users;
    const int64_t u1 = min(a.U, u0 + a.chunk_users);
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) s_hist[c] = 0u;
    if (threadIdx.x == 0) s_nvalid = 0u;
    __syncthreads();
    const int64_t base = f * a.U;
    uint32_t nv = 0;
    for (int64_t u = u0 + threadIdx.x; u < u1; u += blockDim.x) {
      const TIN mu = packed[3 * (base + u) + 1];
      const TIN mv = packed[3 * (base + u) + 2];
      int cell;
      const int st = decode_cell(mu, mv, Wf, Hf, a.W, a.H, cell);
      nv += (st == kOk);
      if (st == kOk) atomicAdd(&s_hist[cell], 1u);
      if (st == kOutOfRange) bad = 1;
      if (want_assign) a.assign0[base + u] = (st == kOk) ? s_lut[cell] : (uint16_t)VET_MISSING;
      if (a.cell16) a.cell16[base + u] = (st == kOk) ? (uint16_t)cell : (uint16_t)0xFFFF;
      if (a.cell32) a.cell32[base + u] = cell;
    }
    nv = __reduce_add_sync(kFull, nv);
    if ((threadIdx.x & 31) == 0 && nv) atomicAdd(&s_nvalid, nv);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (a.chunks_per_frame == 1) a.nvalid[f] = s_nvalid;
      else if (s_nvalid) atomicAdd(&a.nvalid[f], s_nvalid);
    }
    uint32_t* __restrict__ row = a.cnt + f * (int64_t)a.cpad;
    if (a.chunks_per_frame == 1) {
      for (int c = threadIdx.x; c < a.C; c += blockDim.x) row[c] = s_hist[c];
    } else {
      for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
        const uint32_t v = s_hist[c];
        if (v) atomicAdd(&row[c], v);
      }
    }
    __syncthreads();
  }
  if (bad) atomicOr(a.flags, (uint32_t)VET_FLAG_OUT_OF_RANGE);
}

// Streaming kernel of the GLOBAL-TABLE regime: videos whose cell grid does not fit the shared-memory
// histogram + LUT of the kernels above (e.g. the 200x400 of the reference's README: 80,601 cells) but
// whose per-cell tables are still small.  Same per-cell formulation, the tables just stay in global
// memory / L2: one RED into the frame's cell histogram (weighted handles; the tensor-core / FP64
// weighted histogram then runs on it unchanged) or one RED per tile count into the frame's tile
// histograms (unweighted handles), one LUT gather for the assignment.  Bound by L2 atomics, not by HBM.
struct StreamGlobalArgs {
  StreamArgs s;                       // cnt may be null (unweighted handles)
  int K;
  int sumT;
  int hist_off[16];                   // offset of tile count k inside an ihist row
  const uint16_t* lut[16];            // [C] per tile count
  uint32_t* ihist;                    // [F, sumT] (pre-zeroed) or null
};

template <typename TIN>
__global__ void __launch_bounds__(256) k_stream_global(StreamGlobalArgs A) {
  const StreamArgs& a = A.s;
  const TIN* __restrict__ packed = static_cast<const TIN*>(a.packed);
  const float Wf = (float)a.W, Hf = (float)a.H;
  const int64_t n = a.F * a.U;
  const int lane = threadIdx.x & 31;
  uint32_t bad = 0;
  // whole warps walk consecutive samples, so a warp usually sits inside one frame
  const int64_t nround = (n + 31) & ~(int64_t)31;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nround; i += (int64_t)gridDim.x * blockDim.x) {
    const bool in = i < n;
    const int64_t f = in ? i / a.U : -1;
    int cell = -1;
    int st = kMissing;
    if (in) st = decode_cell(packed[3 * i + 1], packed[3 * i + 2], Wf, Hf, a.W, a.H, cell);
    const bool ok = st == kOk;
    if (st == kOutOfRange) bad = 1;
    if (ok) {
      if (a.cnt) atomicAdd(&a.cnt[f * (int64_t)a.cpad + cell], 1u);
      const uint32_t t0 = __ldg(A.lut[0] + cell);
      if (a.assign0) a.assign0[i] = (uint16_t)t0;
      if (A.ihist) {
        uint32_t* __restrict__ row = A.ihist + f * (int64_t)A.sumT;
        atomicAdd(&row[A.hist_off[0] + t0], 1u);
        for (int k = 1; k < A.K; ++k) atomicAdd(&row[A.hist_off[k] + __ldg(A.lut[k] + cell)], 1u);
      }
    } else if (in && a.assign0) {
      a.assign0[i] = (uint16_t)VET_MISSING;
    }
    if (in && a.cell16) a.cell16[i] = ok ? (uint16_t)cell : (uint16_t)0xFFFF;
    if (in && a.cell32) a.cell32[i] = ok ? cell : -1;
    // present users per frame: one atomic per warp when the warp is inside one frame
    const int64_t f0 = __shfl_sync(kFull, f, 0);
    if (__all_sync(kFull, f == f0 || !in)) {
      const uint32_t c = __popc(__ballot_sync(kFull, ok));
      if (lane == 0 && c) atomicAdd(&a.nvalid[f0], c);
    } else if (ok) {
      atomicAdd(&a.nvalid[f], 1u);
    }
  }
  if (bad) atomicOr(a.flags, (uint32_t)VET_FLAG_OUT_OF_RANGE);
}

// Weighted handles of the global-table regime whose cell grid fits shared memory as 16-BIT counters
// (up to ~110k cells: the README's 200x400 video): one CTA per (frame, chunk of <= 65535 users) keeps a
// privatised histogram of packed uint16 pairs -- a sample adds 1 or 65536 to its pair's word, and a chunk
// cannot overflow a half -- instead of one L2 RED per sample.  The flush writes the uint32 row of `cnt`
// (plain stores when the frame is one chunk, RED into the pre-zeroed row otherwise).
template <typename TIN>
__global__ void __launch_bounds__(1024, 1) k_stream_frame16(StreamArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_h = reinterpret_cast<uint32_t*>(smem_raw);
  __shared__ uint32_t s_nvalid;
  const TIN* __restrict__ packed = static_cast<const TIN*>(a.packed);
  const float Wf = (float)a.W, Hf = (float)a.H;
  const int words = (a.cpad + 1) >> 1;
  const int64_t items = a.F * a.chunks_per_frame;
  uint32_t bad = 0;
  for (int c = threadIdx.x; c < words; c += blockDim.x) s_h[c] = 0u;
  if (threadIdx.x == 0) s_nvalid = 0u;
  __syncthreads();
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t f = item / a.chunks_per_frame;
    const int64_t u0 = (item % a.chunks_per_frame) * a.chunk_users;
    const int64_t u1 = min(a.U, u0 + a.chunk_users);
    const int64_t base = f * a.U;
    uint32_t nv = 0;
    for (int64_t u = u0 + threadIdx.x; u < u1; u += blockDim.x) {
      const TIN mu = packed[3 * (base + u) + 1];
      const TIN mv = packed[3 * (base + u) + 2];
      int cell;
      const int st = decode_cell(mu, mv, Wf, Hf, a.W, a.H, cell);
      const bool ok = st == kOk;
      if (st == kOutOfRange) bad = 1;
      if (ok) {
        atomicAdd(&s_h[cell >> 1], (cell & 1) ? 0x10000u : 1u);
        ++nv;
      }
      if (a.assign0) a.assign0[base + u] = ok ? __ldg(a.lut0 + cell) : (uint16_t)VET_MISSING;
      if (a.cell16) a.cell16[base + u] = ok ? (uint16_t)cell : (uint16_t)0xFFFF;
      if (a.cell32) a.cell32[base + u] = ok ? cell : -1;
    }
    nv = __reduce_add_sync(kFull, nv);
    if ((threadIdx.x & 31) == 0 && nv) atomicAdd(&s_nvalid, nv);
    __syncthreads();
    uint32_t* __restrict__ row = a.cnt + f * (int64_t)a.cpad;  // cpad is a multiple of 4: pairs never straddle rows
    if (a.chunks_per_frame == 1) {
      for (int c = threadIdx.x; c < words; c += blockDim.x) {
        const uint32_t v = s_h[c];
        s_h[c] = 0u;
        *reinterpret_cast<uint2*>(row + 2 * c) = make_uint2(v & 0xFFFFu, v >> 16);
      }
      if (threadIdx.x == 0) a.nvalid[f] = s_nvalid;
    } else {
      for (int c = threadIdx.x; c < words; c += blockDim.x) {
        const uint32_t v = s_h[c];
        if (v) {
          s_h[c] = 0u;
          if (v & 0xFFFFu) atomicAdd(&row[2 * c], v & 0xFFFFu);
          if (v >> 16) atomicAdd(&row[2 * c + 1], v >> 16);
        }
      }
      if (threadIdx.x == 0 && s_nvalid) atomicAdd(&a.nvalid[f], s_nvalid);
    }
    __syncthreads();
    if (threadIdx.x == 0) s_nvalid = 0u;
    __syncthreads();
  }
  if (bad) atomicOr(a.flags, (uint32_t)VET_FLAG_OUT_OF_RANGE);
}

struct TileSetDev {
  int T;
  const uint16_t* lut;       // [C]
  const uint32_t* col_ptr;   // [T+1]  (weighted only)
  const uint32_t* cell_idx;  // [nnz]
  const double* w_val;       // [nnz]
};

constexpr int kMaxTileCounts = 16;

struct EpilogueArgs {
  const uint32_t* cnt;  // [F,cpad]
  int64_t F;
  int C;
  int cpad;
  int K;
  int use_weight;
  int norm_always;  // normalise by the tile count even for few users (naive tiling with use_weight_distribution)
  int norm_T0;      // tile count of the normalisation for tile set 0 when it differs from its T (naive tiling), else 0
  TileSetDev ts[kMaxTileCounts];
  double* entropy;  // [F]
  double* per_k;    // [K, per_k_stride] or null
  int64_t per_k_stride;
  double* hist0;    // [F,T0] or null
  uint32_t* flags;
};

// Per-frame epilogue.  One block per frame (grid-stride).  The frame's cell
// histogram is staged in shared memory; for every tile count the tile histogram is
//   unweighted: hist[t] = sum of cnt[cell] over cells whose nearest tile is t
//               (integer, exact, order-free)                              EU:139-142
//   weighted:   hist[t] = sum_cell cnt[cell] * w(cell,t) as a gather over tile t's
//               column of the weight table, one warp per tile, fixed summation
//               order (deterministic)                                     EU:130-138,190-192
// then EU:195-209 gives the normalised entropy; SA:156 averages over tile counts.
__global__ void __launch_bounds__(512, 1) k_epilogue(EpilogueArgs a, int maxT) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(smem_raw);
  double* s_hist = reinterpret_cast<double*>(smem_raw + (size_t)a.cpad * 4);
  uint32_t* s_ihist = reinterpret_cast<uint32_t*>(s_hist + maxT);
  __shared__ double s_red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int64_t f = blockIdx.x; f < a.F; f += gridDim.x) {
    const uint32_t* __restrict__ row = a.cnt + f * (int64_t)a.cpad;
    unsigned long long nloc = 0;
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      const uint32_t v = row[c];
      s_cnt[c] = v;
      nloc += v;
    }
    const double n_valid = block_sum((double)nloc, s_red);  // exact: integers < 2^53
    if (n_valid == 0.0 && threadIdx.x == 0) atomicOr(a.flags, (uint32_t)VET_FLAG_EMPTY_FRAME);
    double esum = 0.0;
    for (int k = 0; k < a.K; ++k) {
      const TileSetDev ts = a.ts[k];
      const int T = ts.T;
      double total;
      if (!a.use_weight) {
        for (int t = threadIdx.x; t < T; t += blockDim.x) s_ihist[t] = 0u;
        __syncthreads();
        for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
          const uint32_t v = s_cnt[c];
          if (v) atomicAdd(&s_ihist[ts.lut[c]], v);
        }
        __syncthreads();
        for (int t = threadIdx.x; t < T; t += blockDim.x) s_hist[t] = (double)s_ihist[t];
        total = n_valid;
        __syncthreads();
      } else {
        for (int t = wid; t < T; t += nw) {
          const uint32_t j0 = ts.col_ptr[t], j1 = ts.col_ptr[t + 1];
          double acc = 0.0;
          for (uint32_t j = j0 + lane; j < j1; j += 32) {
            const uint32_t v = s_cnt[ts.cell_idx[j]];
            acc = fma((double)v, ts.w_val[j], acc);
          }
          acc = warp_sum(acc);
          if (lane == 0) s_hist[t] = acc;
        }
        __syncthreads();
        double part = 0.0;
        for (int t = threadIdx.x; t < T; t += blockDim.x) part += s_hist[t];
        total = block_sum(part, s_red);
      }
      double e = normalized_entropy(s_hist, T, total, a.use_weight != 0 || a.norm_always != 0, s_red, k == 0 ? a.norm_T0 : 0);
      if (n_valid == 0.0) e = __longlong_as_double(0x7ff8000000000000LL);
      if (threadIdx.x == 0 && a.per_k) a.per_k[k * a.per_k_stride + f] = e;
      if (k == 0 && a.hist0)
        for (int t = threadIdx.x; t < T; t += blockDim.x) a.hist0[f * (int64_t)T + t] = s_hist[t];
      esum += e;  // SA:151: sequential accumulation over tile counts
      __syncthreads();
    }
    if (threadIdx.x == 0) a.entropy[f] = esum / (double)a.K;  // SA:156
    __syncthreads();
  }
}

// (2dmu, 2dmv) records of the two-column host layout -> the (time, 2dmu, 2dmv) records the streaming kernels read;
// the time column is never read by any kernel and is set to zero.
template <typename T>
__global__ void k_widen_records(const T* __restrict__ uv, int64_t n, T* __restrict__ packed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    packed[3 * i] = (T)0;
    packed[3 * i + 1] = uv[2 * i];
    packed[3 * i + 2] = uv[2 * i + 1];
  }
}

}  // namespace vet
