// FOV-weighted tile histograms for many frames at once (north_star stage 3, weighted).
//
//   HIST[f, t] = sum_cell CNT[f, cell] * w(cell, t)            (EU:130-138, 190-192)
//
// CNT is the per-frame cell histogram written by the streaming kernel; w is the
// (cell, tile) FOV weight, non-zero for ~15-25 % of the pairs.  A per-frame gather
// re-reads the whole weight table for every frame and is L2-bandwidth bound (first
// profile: 2.0 ms, 13 TB/s of L2 traffic).  Here the contraction is blocked like a
// GEMM on the FP64 pipe instead:
//
//   * tiles are clustered into groups of 8 spatial neighbours; a group owns the UNION
//     of its members' supports as a sorted cell list and a dense [cells][8] weight
//     block (zeros where a member does not see the cell); cells are taken as aligned
//     pairs so that counts load as 8-byte words; weights are chunked 128 cells at a time;
//   * a CTA of 8 warps works on (frame block of 64, group):
//     lane 0 of warp 0 streams the group's chunks into shared memory with cp.async.bulk
//     (three stages, mbarrier hand-off); every consumer warp owns 8 frames;
//   * lane = cell pair: per step a lane reads its cells' 8 weights each from shared
//     memory (conflict-free) and the pair's counts in each of the warp's 8 frames from
//     global (L2-resident, near-coalesced, prefetched two steps ahead in registers) and
//     issues 2x8x8 DFMAs into register accumulators: one 4-byte load per 8 DFMAs;
//   * at the end the 64 accumulators are reduced across the warp with a transposing
//     butterfly (62 shuffles instead of 320) in a fixed order -> deterministic sums.
//
// No tensor cores: the contraction runs on counts and weights that must stay fp64 for
// the 1e-9 entropy tolerance, and the fp64 tensor path offers no throughput over DFMA.
#pragma once
#include "vet_common.cuh"
#include "vet_stream_tma.cuh"

namespace vet {

#ifndef VET_WHIST_WARPS
#define VET_WHIST_WARPS 8
#endif
constexpr int kWhWarps = VET_WHIST_WARPS;        // warps per CTA, two per SM sub-partition; lane 0 of warp 0 doubles as the
                                   // weight-chunk producer (256 threads -> 255 registers each)
constexpr int kWhStages = 3;
constexpr int kWhThreads = kWhWarps * 32;
constexpr int kWhRowPad = 128;     // frames per CTA of the tallest shape: the count scratch has this many spare rows
constexpr int kUnitPad = 2 * 8 * 32;  // the unit list is readable this far past its end (prefetch runs ahead; depth <= 8)

// Blocking of the contraction: TG tiles per group x FW frames per warp (TG*FW = 64 register
// accumulators), Q cells per load unit.
// DEPTH = depth of the register ring: counts are fetched DEPTH-1 warp steps ahead; a staged
// weight chunk holds DEPTH warp steps (the ring is unrolled over them, so its rotation costs no moves).
template <int TG_, int FW_, int Q_, int DEPTH_>
struct WhistShape {
  static constexpr int TG = TG_, FW = FW_, Q = Q_, kDepth = DEPTH_;
  static constexpr int kChunkUnits = 32 * kDepth;  // load units per staged weight chunk
  static constexpr int kChunkCells = kChunkUnits * Q;
  static constexpr int kChunkBytes = kChunkCells * TG * 8;  // weights [TG][Q][kChunkUnits] f64
  static constexpr int kFramesPerCta = kWhWarps * FW;
  static_assert(TG * FW == 64, "the warp reduction below is written for 64 accumulators");
};
// 8 tiles x 8 frames, counts as aligned pairs, 3 steps of prefetch.  (Measured on configs[2] against 4 tiles x 16 frames
// and against aligned quads: 0.49 ms vs 0.67 / 0.52 ms; those shapes are gone.)
using WhistWide = WhistShape<8, 8, 2, 4>;

struct WhistArgs {
  const uint32_t* cnt;     // [F + kWhRowPad, cpad]: readable past F (rows of a partial last frame block)
  int64_t F;
  int cpad;
  int T;
  int G;                         // tile groups
  const int32_t* group_tiles;    // [G,TG] tile index or -1
  const uint32_t* group_chunk0;  // [G+1] first chunk of each group
  const unsigned char* chunks;   // [nchunks][kChunkBytes]
  const uint32_t* units;         // [nchunks*kChunkUnits + kUnitPad] first cell of every load unit (multiple of Q)
  double* hist;                  // [F,T]
  const uint32_t* cta_items;     // [gridDim.x, max_items] (frame block * G + group) per CTA, 0xFFFFFFFF = none:
  int max_items;                 // host-side longest-processing-time schedule (items differ in size)
};

// exact u32 -> f64.  VET_CVT_MAGIC: 2^52 + v has v in its low mantissa bits, one DADD on
// the FP64 pipe plus two register moves; otherwise I2F.F64.U32 on the conversion unit,
// which runs beside the FP64 pipe.
__device__ __forceinline__ double u32_to_f64(uint32_t v) {
#if defined(VET_CVT_MAGIC)
  return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
#else
  return (double)v;
#endif
}

// One step of the transposing warp reduction: lanes with (lane & O) keep the upper LIVE
// values and send the lower ones to their partner, the others the reverse.
template <int O, int LIVE>
__device__ __forceinline__ void butterfly_step(double* flat, int lane) {
  const bool upper = (lane & O) != 0;
#pragma unroll
  for (int i = 0; i < LIVE; ++i) {
    const double keep = upper ? flat[i + LIVE] : flat[i];
    const double send = upper ? flat[i] : flat[i + LIVE];
    flat[i] = keep + __shfl_xor_sync(kFull, send, O);
  }
}

template <int Q>
struct CountUnit;
template <>
struct CountUnit<1> {
  uint32_t v;
  __device__ __forceinline__ void load(const uint32_t* p) { v = __ldg(p); }
  __device__ __forceinline__ uint32_t get(int) const { return v; }
};
template <>
struct CountUnit<4> {
  uint4 v;
  __device__ __forceinline__ void load(const uint32_t* p) { v = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ uint32_t get(int j) const { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }
};
template <>
struct CountUnit<2> {
  uint2 v;
  __device__ __forceinline__ void load(const uint32_t* p) { v = __ldg(reinterpret_cast<const uint2*>(p)); }
  __device__ __forceinline__ uint32_t get(int j) const { return j == 0 ? v.x : v.y; }
};

// (Skipping the half of a tile group that cannot see a run of cells -- 18 % fewer DFMAs on
// configs[2] -- was measured SLOWER on B200, 0.57 vs 0.49 ms, per step or per chunk: the
// halved DFMA-per-load ratio and the extra code paths cost more than the saved math.)

// DFMAs of one warp step restricted to tiles [T0, T1) of the group.
template <typename S, int T0, int T1>
__device__ __forceinline__ void whist_step(const double* sW, int s, int lane, const CountUnit<S::Q> (&use)[S::FW],
                                           double (&acc)[S::TG][S::FW]) {
#pragma unroll
  for (int j = 0; j < S::Q; ++j) {
    double w[S::TG];
#pragma unroll
    for (int t = T0; t < T1; ++t) w[t] = sW[(t * S::Q + j) * S::kChunkUnits + s * 32 + lane];
#pragma unroll
    for (int r = 0; r < S::FW; ++r) {
      const double v = u32_to_f64(use[r].get(j));
#pragma unroll
      for (int t = T0; t < T1; ++t) acc[t][r] = fma(v, w[t], acc[t][r]);
    }
  }
}

template <typename S>
__global__ void __launch_bounds__(kWhThreads, 1) k_whist(WhistArgs a) {
  constexpr int TG = S::TG, FW = S::FW, Q = S::Q, kD = S::kDepth;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long s_full[kWhStages], s_empty[kWhStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWhStages; ++i) {
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&s_empty[i]), kWhWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // Items (frame block, group) come from a host-made per-CTA list balanced by size; within a
  // CTA they are frame-block major so that CTAs running together share CNT rows in L2.  No
  // CTA-wide barrier after this point: warps only meet through the weight-chunk mbarriers.
  uint32_t n = 0;  // chunk sequence number (stage = n % kWhStages), identical in every warp
  for (int it = 0; it < a.max_items; ++it) {
    const uint32_t item = a.cta_items[(size_t)blockIdx.x * a.max_items + it];
    if (item == 0xFFFFFFFFu) break;
    const int g = (int)(item % (uint32_t)a.G);
    const int64_t fb = item / (uint32_t)a.G;
    const uint32_t c0 = a.group_chunk0[g], c1 = a.group_chunk0[g + 1];
    // weight chunk `c` of this item goes to stage (n0 + c - c0) % kWhStages; issued by lane 0 of warp 0
    const uint32_t n0 = n;
    auto produce = [&](uint32_t c) {
      if (warp == 0 && lane == 0 && c < c1) {
        const uint32_t m = n0 + (c - c0);
        const int stage = m % kWhStages;
        mbar_wait(smem_u32(&s_empty[stage]), ((m / kWhStages) & 1u) ^ 1u);
        const uint32_t bar = smem_u32(&s_full[stage]);
        mbar_expect_tx(bar, S::kChunkBytes);
        bulk_g2s(smem_u32(smem_raw + stage * S::kChunkBytes), a.chunks + (size_t)c * S::kChunkBytes, S::kChunkBytes, bar);
      }
    };
#pragma unroll
    for (int i = 0; i < kWhStages - 1; ++i) produce(c0 + i);

    const int64_t f0 = fb * S::kFramesPerCta + warp * FW;
    // first row of this warp's frames; rows past F exist in the scratch (zero-cost padding, results discarded)
    const uint32_t* __restrict__ row0 = a.cnt + f0 * (int64_t)a.cpad;
    double acc[TG][FW];
#pragma unroll
    for (int t = 0; t < TG; ++t)
#pragma unroll
      for (int r = 0; r < FW; ++r) acc[t][r] = 0.0;

    // Register pipeline over warp steps (32 load units each).  While step S runs its DFMAs,
    // the counts of steps S+1 .. S+S::kDepth-1 are in flight in the ring and the unit indices of
    // steps up to S+2(S::kDepth-1) in uring, so L2/DRAM latency is covered by S::kDepth-1 steps of
    // FP64 work without staging the counts in shared memory.
    const uint32_t* __restrict__ up = a.units + (size_t)c0 * S::kChunkUnits + lane;
    CountUnit<Q> ring[S::kDepth][FW];
    uint32_t uring[S::kDepth];
#pragma unroll
    for (int i = 0; i < S::kDepth - 1; ++i) {
      const uint32_t* p = row0 + __ldg(up + 32 * i);
#pragma unroll
      for (int r = 0; r < FW; ++r, p += a.cpad) ring[i][r].load(p);
    }
#pragma unroll
    for (int x = S::kDepth - 1; x < 2 * S::kDepth - 2; ++x) uring[x % S::kDepth] = __ldg(up + 32 * x);
    up += 32 * (2 * S::kDepth - 2);

    for (uint32_t c = c0; c < c1; ++c, ++n) {
      produce(c + kWhStages - 1);
      const int stage = n % kWhStages;
      mbar_wait(smem_u32(&s_full[stage]), (n / kWhStages) & 1u);
      const double* sW = reinterpret_cast<const double*>(smem_raw + stage * S::kChunkBytes);
#pragma unroll
      for (int s = 0; s < S::kDepth; ++s) {
        // issue the count loads of step S+S::kDepth-1 and the unit load of step S+2(S::kDepth-1)
        const uint32_t* p = row0 + uring[(s + S::kDepth - 1) % S::kDepth];
#pragma unroll
        for (int r = 0; r < FW; ++r, p += a.cpad) ring[(s + S::kDepth - 1) % S::kDepth][r].load(p);
        uring[(s + S::kDepth - 2) % S::kDepth] = __ldg(up);
        up += 32;
        whist_step<S, 0, TG>(sW, s, lane, ring[s], acc);  // step S out of ring[s]
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_empty[stage]));
    }

    // transposing butterfly: after the step with offset o a lane keeps the half of its
    // live accumulators selected by (lane & o); 64 values -> 2 per lane
    double* flat = &acc[0][0];
    butterfly_step<16, 32>(flat, lane);
    butterfly_step<8, 16>(flat, lane);
    butterfly_step<4, 8>(flat, lane);
    butterfly_step<2, 4>(flat, lane);
    butterfly_step<1, 2>(flat, lane);
    int base = 0;
    {
      int span = 64;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        span >>= 1;
        if (lane & o) base += span;
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int id = base + j;  // = t * FW + r
      const int t = id / FW, r = id % FW;
      const int tile = a.group_tiles[g * TG + t];
      const int64_t f = f0 + r;
      if (tile >= 0 && f < a.F) a.hist[f * (int64_t)a.T + tile] = flat[j];
    }
  }
}

// Per-frame normalised entropy from finished histogram rows (EU:195-209) and the
// average over tile counts (SA:151-156).  One warp per frame.
struct EntropyRowsArgs {
  int64_t F;
  int K;
  int T[kMaxTileCounts];
  const double* hist[kMaxTileCounts];  // [F,T_k] weighted histograms (fp64), or null when ihist is used
  const uint32_t* ihist;               // [F,istride] integer tile histograms of the direct unweighted path
  int64_t istride;
  int ioff[kMaxTileCounts];            // offset of tile count k inside an ihist row
  double* hist0_out;                   // [F,T_0] (ihist mode only): tile_counts[0] histogram as float64, or null
  const uint32_t* nvalid;              // [F] present users per frame
  int use_weight;
  int norm_always;   // naive tiling with use_weight_distribution: always normalise by the tile count
  int norm_T0;       // naive tiling: tile count of the normalisation for tile set 0 (0 = its T)
  double* entropy;   // [F]
  double* per_k;     // [K, stride] or null
  int64_t per_k_stride;
  uint32_t* flags;
};

// Row schedule of k_entropy_rows: a block of 8 warps takes G consecutive frames; their G x K histogram rows are dealt
// to the warps by the host (longest row first onto the least loaded warp), because a row costs one fp64 division and
// one log2 per 32 tiles of latency -- with one warp per tile count the warp of the 201-tile row worked seven times
// longer than the one of the 21-tile row and the block waited for it (ncu: barrier stalls 3.9 per issue).
struct EntropyRowsPlan {
  int G;                        // frames per block (1, 2, 4 or 8)
  unsigned char nrow[8];        // rows of every warp
  unsigned char row_g[8][16];   // frame of the row inside the block's group
  unsigned char row_k[8][16];   // tile count of the row
};

__global__ void __launch_bounds__(256) k_entropy_rows(EntropyRowsArgs a, EntropyRowsPlan pl) {
  __shared__ double s_e[8][kMaxTileCounts];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;   // warp in block
  for (int64_t f0 = (int64_t)blockIdx.x * pl.G; f0 < a.F; f0 += (int64_t)gridDim.x * pl.G) {
    for (int i = 0; i < pl.nrow[wib]; ++i) {
      const int g = pl.row_g[wib][i], k = pl.row_k[wib][i];
      const int64_t f = f0 + g;
      if (f >= a.F) continue;
      const uint32_t nv = __ldcg(a.nvalid + f);
      const int T = a.T[k];
      const double* __restrict__ row = a.ihist ? nullptr : a.hist[k] + f * (int64_t)T;
      const uint32_t* __restrict__ irow = a.ihist ? a.ihist + f * a.istride + a.ioff[k] : nullptr;
      double part = 0.0;
      if (a.use_weight)
        for (int t = lane; t < T; t += 32) part += irow ? (double)__ldcg(irow + t) : row[t];
      const double total = a.use_weight ? warp_sum(part) : (double)nv;
      double acc = 0.0;
      for (int t = lane; t < T; t += 32) {
        const double w = irow ? (double)__ldcg(irow + t) : row[t];
        if (irow && k == 0 && a.hist0_out) a.hist0_out[f * (int64_t)T + t] = w;
        if (w > 0.0) {
          const double p = w / total;
          acc = fma(-p, log2(p), acc);
        }
      }
      const double Hs = warp_sum(acc);
      const double nt = (double)((k == 0 && a.norm_T0 > 0) ? a.norm_T0 : T);
      const double nn = (a.use_weight || a.norm_always || total > nt) ? nt : total;
      const double mp = 1.0 / nn;
      const double mx = -nn * mp * log2(mp);
      double e = Hs / mx;
      if (nv == 0) e = __longlong_as_double(0x7ff8000000000000LL);
      if (lane == 0) {
        if (a.per_k) a.per_k[k * a.per_k_stride + f] = e;
        s_e[g][k] = e;
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < pl.G && f0 + threadIdx.x < a.F) {   // SA:151-156: the average in tile-count order
      const int64_t f = f0 + threadIdx.x;
      if (__ldcg(a.nvalid + f) == 0) atomicOr(a.flags, (uint32_t)VET_FLAG_EMPTY_FRAME);
      double esum = 0.0;
      for (int k = 0; k < a.K; ++k) esum += s_e[threadIdx.x][k];
      a.entropy[f] = esum / (double)a.K;
    }
    __syncthreads();
  }
}

// One frame per block and step, the div + log2 of a frame spread over all 8 warps.  k_entropy_rows gives a warp whole
// rows: a 201-tile row is seven dependent div + log2 rounds of one warp (~3.5 us), a group of 8 frames x 4 tile counts
// takes a block 14 us however many blocks run -- 15 us behind a 60 us streaming kernel on configs[1].  Here unit (k, j)
// = tiles 32 j .. 32 j + 31 of tile count k goes to warp u mod 8, which leaves p and log2 p in shared memory; then the
// warp of row k adds the products in k_entropy_rows' order (lane l: tiles l, l + 32, ...; butterfly), so the bits are
// the same.  Dynamic shared memory: 2 x soff[K] doubles.  configs[1]: 15 -> 10 us, the step 0.077 -> 0.072 ms.
// Measured and dropped: this kernel launched as a programmatic dependent of k_stream_tiles (griddepcontrol), waiting
// per frame on completion counters and working beside the streaming CTAs -- one or two blocks fit next to a streaming
// CTA, a frame takes such a block ~4 us, and the step grew to 0.094 ms.
struct EntropyFramesPlan {
  int nunits;
  int uoff[kMaxTileCounts + 1];  // first unit of every tile count
  int soff[kMaxTileCounts + 1];  // first shared-memory slot of every tile count (32 per unit)
};

__global__ void __launch_bounds__(256) k_entropy_frames(EntropyRowsArgs a, EntropyFramesPlan pl) {
  extern __shared__ __align__(16) double s_pl[];
  double* s_p = s_pl;
  double* s_l = s_pl + pl.soff[a.K];
  __shared__ double s_tot[kMaxTileCounts], s_e[kMaxTileCounts];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int64_t f = blockIdx.x; f < a.F; f += gridDim.x) {
    const uint32_t nv = __ldcg(a.nvalid + f);
    if (a.use_weight) {  // totals of the weighted rows (the unweighted total is the number of present users)
      for (int k = wib; k < a.K; k += 8) {
        const int T = a.T[k];
        const double* __restrict__ row = a.hist[k] + f * (int64_t)T;
        double part = 0.0;
        for (int t = lane; t < T; t += 32) part += row[t];
        part = warp_sum(part);
        if (lane == 0) s_tot[k] = part;
      }
      __syncthreads();
    }
    for (int u = wib; u < pl.nunits; u += 8) {
      int k = 0;
      while (u >= pl.uoff[k + 1]) ++k;
      const int j = u - pl.uoff[k], t = 32 * j + lane, T = a.T[k];
      const double total = a.use_weight ? s_tot[k] : (double)nv;
      double w = 0.0;
      if (t < T) {
        w = a.ihist ? (double)__ldcg(a.ihist + f * a.istride + a.ioff[k] + t) : a.hist[k][f * (int64_t)T + t];
        if (a.ihist && k == 0 && a.hist0_out) a.hist0_out[f * (int64_t)T + t] = w;
      }
      double p = 0.0, l = 0.0;
      if (w > 0.0) {
        p = w / total;
        l = log2(p);
      }
      s_p[pl.soff[k] + t] = p;
      s_l[pl.soff[k] + t] = l;
    }
    __syncthreads();
    for (int k = wib; k < a.K; k += 8) {
      const int T = a.T[k], nj = pl.uoff[k + 1] - pl.uoff[k];
      const double total = a.use_weight ? s_tot[k] : (double)nv;
      double acc = 0.0;
      for (int j = 0; j < nj; ++j) acc = fma(-s_p[pl.soff[k] + 32 * j + lane], s_l[pl.soff[k] + 32 * j + lane], acc);
      const double Hs = warp_sum(acc);
      const double nt = (double)((k == 0 && a.norm_T0 > 0) ? a.norm_T0 : T);
      const double nn = (a.use_weight || a.norm_always || total > nt) ? nt : total;
      const double mp = 1.0 / nn;
      const double mx = -nn * mp * log2(mp);
      double e = Hs / mx;
      if (nv == 0) e = __longlong_as_double(0x7ff8000000000000LL);
      if (lane == 0) {
        if (a.per_k) a.per_k[k * a.per_k_stride + f] = e;
        s_e[k] = e;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // SA:151-156: the average in tile-count order
      if (nv == 0) atomicOr(a.flags, (uint32_t)VET_FLAG_EMPTY_FRAME);
      double esum = 0.0;
      for (int k = 0; k < a.K; ++k) esum += s_e[k];
      a.entropy[f] = esum / (double)a.K;
    }
  }
}

}  // namespace vet
