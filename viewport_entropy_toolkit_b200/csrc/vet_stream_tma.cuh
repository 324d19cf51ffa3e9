// Streaming kernel, Blackwell version: a warp-specialised producer/consumer pipeline.
//
//   producer (warp 0, one elected lane)  cp.async.bulk global -> shared, 24 KB tiles of
//                                        the packed (time,2dmu,2dmv) stream, completion
//                                        signalled on an mbarrier (complete_tx::bytes)
//   consumers (warps 1..NC)              wait on the tile's "full" barrier, decode their
//                                        samples from shared memory (the stride-3-word reads
//                                        are bank-conflict free; the conflicts ncu reports for
//                                        this kernel -- 56 % of its shared-memory wavefronts --
//                                        are the histogram atomics below), bump the privatised
//                                        per-frame cell histogram with shared-memory
//                                        atomics, look the tile index up in the
//                                        shared-memory LUT and store it, then release the
//                                        stage on its "empty" barrier
//
// One CTA per SM.  A work item is (frame, user-chunk); the producer runs ahead across
// item boundaries, so HBM stays busy while the consumers flush a finished histogram.
// Global loads cost no LSU issue slots and no registers; the bytes in flight per SM are
// kStages x 24 KB, enough to cover HBM latency at the measured 6.5 TB/s.
#pragma once
#include "vet_common.cuh"
#include "vet_stream.cuh"

namespace vet {

#ifndef VET_STREAM_STAGES
#define VET_STREAM_STAGES 4
#endif
#ifndef VET_STREAM_CWARPS
#define VET_STREAM_CWARPS 16
#endif
#ifndef VET_STREAM_MINBLOCKS
#define VET_STREAM_MINBLOCKS 1
#endif
constexpr int kStages = VET_STREAM_STAGES;
constexpr int kTileBytes = 24576;              // payload per stage: 2048 fp32 samples / 1024 fp64 samples
constexpr int kStageBytes = kTileBytes + 32;   // + 16 B alignment head and tail
constexpr int kConsumerWarps = VET_STREAM_CWARPS;
constexpr int kStreamThreads = (kConsumerWarps + 1) * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared (SASS: UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void consumer_sync() {  // named barrier 1: consumer warps only
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
}

struct StreamTmaArgs {
  StreamArgs s;
  const void* lut0_typed;   // uint8_t[C] or uint16_t[C]
  int64_t total_bytes;      // bytes of the packed tensor of this call (for the alignment clip)
  int cpad;                 // C rounded up to a multiple of 4 (row pitch of cnt, in cells)
  // Optional (frames of one chunk only): write the frame's cell histogram as BYTE PLANES for the
  // tensor-core weighted histogram (vet_whist_i8.cuh) instead of the uint32 row of `cnt`.
  //   plane p, row f = bits [8p, 8p+8) of every count, kp bytes per row (zero padded)
  // Planes 1 and 2 are almost always zero: their rows are only touched when the frame has such
  // counts or when the row is marked dirty (non-zero from an earlier call), so the usual cost is
  // kp bytes per frame instead of 4 cpad.
  uint8_t* planes;          // [3][plane_rows][kp] or null
  int64_t plane_stride;     // bytes between planes
  int kp;                   // bytes per plane row (multiple of 128)
  uint8_t* dirty;           // [plane_rows] bit p-1: row f of plane p may be non-zero
  uint32_t* hi1;            // [plane_rows / 128] frame block has counts >= 256
  uint32_t* hi2;            // [plane_rows / 128] frame block has counts >= 65536
};

__device__ __forceinline__ uint32_t pack_bytes(uint4 v, int shift) {
  return ((v.x >> shift) & 0xFFu) | (((v.y >> shift) & 0xFFu) << 8) | (((v.z >> shift) & 0xFFu) << 16) |
         (((v.w >> shift) & 0xFFu) << 24);
}

// Slow-path classification of a sample that failed the fast [0,1] bit test.
template <typename T>
__device__ __forceinline__ int classify_slow(T mu, T mv) {
  if (mu != mu || mv != mv) return kMissing;
  if (mu < (T)0 || mu > (T)1 || mv < (T)0 || mv > (T)1) return kOutOfRange;
  return kOk;  // -0.0
}
// v in [+0, 1] as one unsigned compare on the bit pattern (false for NaN, negatives and -0.0,
// which take the slow path)
__device__ __forceinline__ bool unit_range_fast(float v) { return __float_as_uint(v) <= 0x3F800000u; }
__device__ __forceinline__ bool unit_range_fast(double v) {
  return (unsigned long long)__double_as_longlong(v) <= 0x3FF0000000000000ull;
}
// NaN (= a missing sample) by its bit pattern: a frame with missing users then needs no branch into classify_slow --
// with 20 % of the samples missing every warp step used to diverge into it (configs[2]: 1.12 ms against 0.86).
__device__ __forceinline__ bool is_nan_bits(float v) { return (__float_as_uint(v) & 0x7FFFFFFFu) > 0x7F800000u; }
__device__ __forceinline__ bool is_nan_bits(double v) {
  return ((unsigned long long)__double_as_longlong(v) & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull;
}

// Cooperative 16-byte copy global -> shared of `bytes` (rounded up to 16; both buffers are
// allocated with that padding).  Element-wise copies of the LUTs showed up as 30 % of the
// short direct-unweighted kernel (dependent byte loads in the prologue).
__device__ __forceinline__ void copy_to_smem16(void* dst, const void* __restrict__ src, int bytes) {
  const int n16 = (bytes + 15) >> 4;
  const uint4* __restrict__ s = static_cast<const uint4*>(src);
  uint4* d = static_cast<uint4*>(dst);
  for (int i = threadIdx.x; i < n16; i += blockDim.x) d[i] = __ldg(s + i);
}

// Producer side of the streaming pipeline (one elected lane): walks the CTA's work items
// and fills the stage ring with 16 B-aligned supersets of each 24 KB tile.
template <typename TIN>
__device__ __forceinline__ void stream_producer(const StreamArgs& a, int64_t total_bytes, unsigned char* s_stage,
                                                unsigned long long* s_full, unsigned long long* s_empty) {
  constexpr int kSampleBytes = 3 * (int)sizeof(TIN);
  const int64_t items = a.F * a.chunks_per_frame;
  const unsigned char* __restrict__ gbase = static_cast<const unsigned char*>(a.packed);
  const int64_t total16 = total_bytes & ~(int64_t)15;  // bulk copies never read past this
  uint32_t n = 0;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t f = item / a.chunks_per_frame;
    const int64_t u0 = (item % a.chunks_per_frame) * a.chunk_users;
    const int64_t u1 = min(a.U, u0 + a.chunk_users);
    int64_t b = (f * a.U + u0) * kSampleBytes;
    const int64_t bend = (f * a.U + u1) * kSampleBytes;
    for (; b < bend; b += kTileBytes, ++n) {
      const int stage = n % kStages;
      mbar_wait(smem_u32(&s_empty[stage]), ((n / kStages) & 1u) ^ 1u);
      const int64_t a0 = b & ~(int64_t)15;
      const int64_t a1 = min((min(b + kTileBytes, bend) + 15) & ~(int64_t)15, total16);
      const uint32_t bytes = a1 > a0 ? (uint32_t)(a1 - a0) : 0u;
      const uint32_t bar = smem_u32(&s_full[stage]);
      mbar_expect_tx(bar, bytes);
      if (bytes) bulk_g2s(smem_u32(s_stage + stage * kStageBytes), gbase + a0, bytes, bar);
    }
  }
}

// Loads this thread's kPerThread samples of the current tile from the stage buffer
// (fast path: a full tile that was fetched completely; else bounds-checked, with the last
// bytes of a tensor whose size is not a multiple of 16 read straight from global memory).
template <typename TIN, int kPerThread>
__device__ __forceinline__ void load_tile_samples(const TIN* sbuf, const unsigned char* gbase, int64_t scur, int64_t b0,
                                                  int nsamp, bool full_tile, int64_t total16, int ctid, TIN (&mu)[kPerThread],
                                                  TIN (&mv)[kPerThread]) {
  constexpr int kSampleBytes = 3 * (int)sizeof(TIN);
  if (full_tile) {
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const int s = ctid + j * (kConsumerWarps * 32);
      mu[j] = sbuf[3 * s + 1];
      mv[j] = sbuf[3 * s + 2];
    }
  } else if (b0 + (int64_t)nsamp * kSampleBytes <= total16) {
    // a partial tile (the last one of a frame or chunk) that was fetched completely: one 32-bit test per sample
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const int s = ctid + j * (kConsumerWarps * 32);
      const bool in = s < nsamp;
      mu[j] = in ? sbuf[3 * s + 1] : (TIN)0;
      mv[j] = in ? sbuf[3 * s + 2] : (TIN)0;
    }
  } else {
    const int64_t a1 = min((b0 + (int64_t)nsamp * kSampleBytes + 15) & ~(int64_t)15, total16);
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const int s = ctid + j * (kConsumerWarps * 32);
      mu[j] = (TIN)0;
      mv[j] = (TIN)0;
      if (s < nsamp) {
        if (b0 + (int64_t)(s + 1) * kSampleBytes <= a1) {
          mu[j] = sbuf[3 * s + 1];
          mv[j] = sbuf[3 * s + 2];
        } else {
          const TIN* p = reinterpret_cast<const TIN*>(gbase) + 3 * (scur + s);
          mu[j] = p[1];
          mv[j] = p[2];
        }
      }
    }
  }
}

// CELLS: 0 = no cell-id output, 1 = uint16 cell ids, 2 = int32 cell ids (transition stage input)
template <typename TIN, typename TLUT, bool ASSIGN, int CELLS>
__global__ void __launch_bounds__(kStreamThreads, VET_STREAM_MINBLOCKS) k_stream_tma(StreamTmaArgs A) {
  constexpr int kTileSamples = kTileBytes / (3 * (int)sizeof(TIN));
  constexpr int kPerThread = kTileSamples / (kConsumerWarps * 32);
  constexpr int kSampleBytes = 3 * (int)sizeof(TIN);
  static_assert(kTileSamples % (kConsumerWarps * 32) == 0, "tile must split evenly over the consumer threads");
  const StreamArgs& a = A.s;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* s_stage = smem_raw;
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw + kStages * kStageBytes);
  TLUT* s_lut = reinterpret_cast<TLUT*>(s_hist + A.cpad);
  __shared__ __align__(8) unsigned long long s_full[kStages], s_empty[kStages];
  __shared__ uint32_t s_nvalid, s_hibits;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&s_empty[i]), kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int c = threadIdx.x; c < A.cpad; c += blockDim.x) s_hist[c] = 0u;
  if (threadIdx.x == 0) {
    s_nvalid = 0u;
    s_hibits = 0u;
  }
  if (ASSIGN) copy_to_smem16(s_lut, A.lut0_typed, a.C * (int)sizeof(TLUT));
  __syncthreads();

  const int64_t items = a.F * a.chunks_per_frame;
  const unsigned char* __restrict__ gbase = static_cast<const unsigned char*>(a.packed);
  const int64_t total16 = A.total_bytes & ~(int64_t)15;  // bulk copies never read past this

  if (warp == 0) {
    if (lane == 0) stream_producer<TIN>(a, A.total_bytes, s_stage, s_full, s_empty);
    return;
  }

  // ===================== consumers =====================
  const int ctid = threadIdx.x - 32;  // 0 .. kConsumerWarps*32-1
  const float Wf = (float)a.W, Hf = (float)a.H;
  const int W1 = a.W + 1;
  uint32_t n = 0;
  uint32_t bad = 0;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t f = item / a.chunks_per_frame;
    const int64_t u0 = (item % a.chunks_per_frame) * a.chunk_users;
    const int64_t u1 = min(a.U, u0 + a.chunk_users);
    int64_t scur = f * a.U + u0;
    const int64_t send = f * a.U + u1;
    uint32_t nv = 0;
    for (; scur < send; scur += kTileSamples, ++n) {
      const int stage = n % kStages;
      const int64_t b0 = scur * kSampleBytes;
      const int nsamp = (int)min((int64_t)kTileSamples, send - scur);
      const bool full_tile = (nsamp == kTileSamples) && (b0 + kTileBytes <= total16);
      mbar_wait(smem_u32(&s_full[stage]), (n / kStages) & 1u);
      const TIN* sbuf = reinterpret_cast<const TIN*>(s_stage + stage * kStageBytes + (int)(b0 & 15));
      uint16_t* __restrict__ out_assign = ASSIGN ? a.assign0 + scur : nullptr;
      uint16_t* __restrict__ out_c16 = CELLS == 1 ? a.cell16 + scur : nullptr;
      int32_t* __restrict__ out_c32 = CELLS == 2 ? a.cell32 + scur : nullptr;
      TIN mu[kPerThread], mv[kPerThread];
      load_tile_samples<TIN, kPerThread>(sbuf, gbase, scur, b0, nsamp, full_tile, total16, ctid, mu, mv);
#pragma unroll
      for (int j = 0; j < kPerThread; ++j) {
        const int s = ctid + j * (kConsumerWarps * 32);
        if (full_tile || s < nsamp) {
          bool ok = unit_range_fast(mu[j]) && unit_range_fast(mv[j]);
          if (!ok && !is_nan_bits(mu[j]) && !is_nan_bits(mv[j])) {  // not in [+0, 1] and not missing: -0.0 or out of range
            const int st = classify_slow(mu[j], mv[j]);
            ok = st == kOk;
            if (st == kOutOfRange) bad = 1;
          }
          int cell = 0;
          if (ok) {
            cell = pixel_of(mv[j], Hf, a.H) * W1 + pixel_of(mu[j], Wf, a.W);
            atomicAdd(&s_hist[cell], 1u);
            ++nv;
          }
          if (ASSIGN) out_assign[s] = ok ? (uint16_t)s_lut[cell] : (uint16_t)VET_MISSING;
          if (CELLS == 1) out_c16[s] = ok ? (uint16_t)cell : (uint16_t)0xFFFF;
          if (CELLS == 2) out_c32[s] = ok ? cell : -1;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_empty[stage]));
    }
    // item finished: flush the privatised histogram and clear it
    nv = __reduce_add_sync(kFull, nv);
    if (lane == 0 && nv) atomicAdd(&s_nvalid, nv);
    consumer_sync();
    if (ctid == 0) {
      if (a.chunks_per_frame == 1) a.nvalid[f] = s_nvalid;
      else if (s_nvalid) atomicAdd(&a.nvalid[f], s_nvalid);
      s_nvalid = 0u;
    }
    uint4* __restrict__ row = reinterpret_cast<uint4*>(a.cnt + f * (int64_t)A.cpad);
    uint4* s_hist4 = reinterpret_cast<uint4*>(s_hist);
    const int n4 = A.cpad >> 2;
    if (A.planes && a.chunks_per_frame == 1) {
      const uint32_t was = A.dirty[f];  // uniform over the CTA
      uint32_t* __restrict__ r0 = reinterpret_cast<uint32_t*>(A.planes + f * (int64_t)A.kp);
      uint32_t* __restrict__ r1 = reinterpret_cast<uint32_t*>(A.planes + A.plane_stride + f * (int64_t)A.kp);
      uint32_t* __restrict__ r2 = reinterpret_cast<uint32_t*>(A.planes + 2 * A.plane_stride + f * (int64_t)A.kp);
      uint32_t hi = 0;
      for (int c = ctid; c < (A.kp >> 2); c += kConsumerWarps * 32) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (c < n4) {
          v = s_hist4[c];
          s_hist4[c] = make_uint4(0u, 0u, 0u, 0u);
        }
        r0[c] = pack_bytes(v, 0);
        const uint32_t w1 = pack_bytes(v, 8), w2 = pack_bytes(v, 16);
        if ((was & 1u) || w1) r1[c] = w1;
        if ((was & 2u) || w2) r2[c] = w2;
        hi |= (w1 ? 1u : 0u) | (w2 ? 2u : 0u);
      }
      if (hi) atomicOr(&s_hibits, hi);
      consumer_sync();
      if (ctid == 0) {
        const uint32_t now = s_hibits;
        s_hibits = 0u;
        if (now != was) A.dirty[f] = (uint8_t)now;
        if (now & 1u) atomicOr(&A.hi1[f >> 7], 1u);
        if (now & 2u) atomicOr(&A.hi2[f >> 7], 1u);
      }
      continue;
    }
    if (a.chunks_per_frame == 1) {
      for (int c = ctid; c < n4; c += kConsumerWarps * 32) {
        row[c] = s_hist4[c];
        s_hist4[c] = make_uint4(0u, 0u, 0u, 0u);
      }
    } else {
      uint32_t* __restrict__ row1 = a.cnt + f * (int64_t)A.cpad;
      for (int c = ctid; c < a.C; c += kConsumerWarps * 32) {
        const uint32_t v = s_hist[c];
        if (v) {
          atomicAdd(&row1[c], v);
          s_hist[c] = 0u;
        }
      }
    }
    consumer_sync();
  }
  if (bad) atomicOr(a.flags, (uint32_t)VET_FLAG_OUT_OF_RANGE);
}

// ---------------------------------------------------------------------------------------
// Direct unweighted variant: per-sample tile lookups for EVERY tile count and privatised
// per-frame TILE histograms (a few KB) instead of the 81 KB cell histogram.  Used when the
// frames are small against the cell grid (U < ~2C) or there is a single tile count: no cell
// histogram is written or re-read, the frame result is sum(T_k) integers.
struct StreamTilesArgs {
  StreamArgs s;
  int64_t total_bytes;
  int K;
  int sumT;                            // sum of T_k
  int T[kMaxTileCounts];
  int hist_off[kMaxTileCounts];        // offset of tile count k inside a frame's histogram row
  int shist_off[kMaxTileCounts];       // offset of tile count k inside the shared-memory histogram area
  int rep_shift[kMaxTileCounts];       // log2 of the number of interleaved copies of histogram k in shared memory:
                                       // small histograms are replicated per lane group to cut same-address conflicts
  int shist_words;                     // total words of the shared-memory histogram area
  int lut_off[kMaxTileCounts];         // byte offset of LUT k inside the shared-memory LUT area
  int lut_wide[kMaxTileCounts];        // 1 = uint16 entries, 0 = uint8
  const void* lut[kMaxTileCounts];     // global LUTs (uint8 when T <= 255 else uint16)
  int lut_bytes;                       // total bytes of the shared-memory LUT area
  const uint32_t* lut_packed;          // [C] byte k = tile of tile count k (K <= 4, every T <= 255), or null
  uint32_t* ihist;                     // [F, sumT] integer tile histograms (pre-zeroed when chunks_per_frame > 1)
};

// KP > 0: packed variant for KP <= 4 tile counts that all fit a byte: ONE 32-bit shared-memory load
// yields the sample's tile under every tile count and the per-count loop is unrolled (the generic
// loop spends ~25 instructions per sample and tile count on parameter loads and address math).
template <typename TIN, bool ASSIGN, int KP>
__global__ void __launch_bounds__(kStreamThreads, 1) k_stream_tiles(StreamTilesArgs A) {
  constexpr int kTileSamples = kTileBytes / (3 * (int)sizeof(TIN));
  constexpr int kPerThread = kTileSamples / (kConsumerWarps * 32);
  constexpr int kSampleBytes = 3 * (int)sizeof(TIN);
  const StreamArgs& a = A.s;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* s_stage = smem_raw;
  // two copies of the tile histograms: the frame being counted and the one being flushed (one barrier per frame)
  uint32_t* s_hist_base = reinterpret_cast<uint32_t*>(smem_raw + kStages * kStageBytes);  // [2][hw]
  const int hw = (A.shist_words + 3) & ~3;
  unsigned char* s_lut = reinterpret_cast<unsigned char*>(s_hist_base + 2 * hw);
  __shared__ __align__(8) unsigned long long s_full[kStages], s_empty[kStages];
  __shared__ uint32_t s_nvalid[2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&s_empty[i]), kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_nvalid[0] = 0u;
    s_nvalid[1] = 0u;
  }
  for (int t = threadIdx.x; t < 2 * hw; t += blockDim.x) s_hist_base[t] = 0u;
  if (KP > 0) {
    copy_to_smem16(s_lut, A.lut_packed, a.C * 4);
  } else {
    for (int k = 0; k < A.K; ++k) copy_to_smem16(s_lut + A.lut_off[k], A.lut[k], a.C * (A.lut_wide[k] ? 2 : 1));
  }
  __syncthreads();

  if (warp == 0) {
    if (lane == 0) stream_producer<TIN>(a, A.total_bytes, s_stage, s_full, s_empty);
    return;
  }

  const int64_t items = a.F * a.chunks_per_frame;
  const unsigned char* __restrict__ gbase = static_cast<const unsigned char*>(a.packed);
  const int64_t total16 = A.total_bytes & ~(int64_t)15;
  const int ctid = threadIdx.x - 32;
  const float Wf = (float)a.W, Hf = (float)a.H;
  const int W1 = a.W + 1;
  uint32_t* hk[KP > 0 ? KP : 1];  // packed variant: base of every tile count's histogram, in registers
#pragma unroll
  for (int k = 0; k < (KP > 0 ? KP : 1); ++k) hk[k] = s_hist_base + A.shist_off[k];
  uint32_t hka[KP > 0 ? KP : 1];  // the same as 32-bit shared-space addresses (fast path)
#pragma unroll
  for (int k = 0; k < (KP > 0 ? KP : 1); ++k) hka[k] = smem_u32(hk[k]);
  const uint32_t* s_lut32 = reinterpret_cast<const uint32_t*>(s_lut);
  uint32_t n = 0;
  uint32_t bad = 0;
  uint32_t buf = 0;  // histogram copy of the current item
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x, buf ^= 1u) {
    uint32_t* s_hist = s_hist_base + buf * hw;
    const uint32_t hboff = buf * (uint32_t)hw * 4u;
    const int64_t f = item / a.chunks_per_frame;
    const int64_t u0 = (item % a.chunks_per_frame) * a.chunk_users;
    const int64_t u1 = min(a.U, u0 + a.chunk_users);
    int64_t scur = f * a.U + u0;
    const int64_t send = f * a.U + u1;
    uint32_t nv = 0;
    for (; scur < send; scur += kTileSamples, ++n) {
      const int stage = n % kStages;
      const int64_t b0 = scur * kSampleBytes;
      const int nsamp = (int)min((int64_t)kTileSamples, send - scur);
      const bool full_tile = (nsamp == kTileSamples) && (b0 + kTileBytes <= total16);
      mbar_wait(smem_u32(&s_full[stage]), (n / kStages) & 1u);
      const TIN* sbuf = reinterpret_cast<const TIN*>(s_stage + stage * kStageBytes + (int)(b0 & 15));
      uint16_t* __restrict__ out_assign = ASSIGN ? a.assign0 + scur : nullptr;
      TIN mu[kPerThread], mv[kPerThread];
      load_tile_samples<TIN, kPerThread>(sbuf, gbase, scur, b0, nsamp, full_tile, total16, ctid, mu, mv);
      // Fast path (packed LUT, a full tile, every coordinate of this thread in [0, 1]): the samples in phases -- cells
      // and LUT words, then the K increments per sample on 32-bit shared addresses, then the assignments -- without
      // the per-sample bounds and range branches of the general loop below (ncu: 75 instructions per sample there,
      // the kernel issue-bound at 42 % of the HBM peak on configs[1]).
      bool done = false;
      if (KP > 0 && full_tile) {
        bool allok = true;
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) allok = allok && unit_range_fast(mu[j]) && unit_range_fast(mv[j]);
        if (allok) {
          uint32_t w[kPerThread];
#pragma unroll
          for (int j = 0; j < kPerThread; ++j) w[j] = s_lut32[pixel_of(mv[j], Hf, a.H) * W1 + pixel_of(mu[j], Wf, a.W)];
#pragma unroll
          for (int j = 0; j < kPerThread; ++j)
#pragma unroll
            for (int k = 0; k < (KP > 0 ? KP : 1); ++k)
              asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hka[k] + hboff + (__byte_perm(w[j], 0u, 0x4440u + k) << 2)) : "memory");
          if (ASSIGN) {
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) out_assign[ctid + j * (kConsumerWarps * 32)] = (uint16_t)(w[j] & 0xFFu);
          }
          nv += kPerThread;
          done = true;
        }
      }
      if (!done)
#pragma unroll
      for (int j = 0; j < kPerThread; ++j) {
        const int s = ctid + j * (kConsumerWarps * 32);
        if (full_tile || s < nsamp) {
          bool ok = unit_range_fast(mu[j]) && unit_range_fast(mv[j]);
          if (!ok && !is_nan_bits(mu[j]) && !is_nan_bits(mv[j])) {  // not in [+0, 1] and not missing: -0.0 or out of range
            const int st = classify_slow(mu[j], mv[j]);
            ok = st == kOk;
            if (st == kOutOfRange) bad = 1;
          }
          uint32_t t0 = VET_MISSING;
          if (ok) {
            const int cell = pixel_of(mv[j], Hf, a.H) * W1 + pixel_of(mu[j], Wf, a.W);
            ++nv;
            if (KP > 0) {
              const uint32_t w = s_lut32[cell];
              t0 = w & 0xFFu;
#pragma unroll
              for (int k = 0; k < KP; ++k) atomicAdd(hk[k] + buf * hw + ((w >> (8 * k)) & 0xFFu), 1u);
            } else
            for (int k = 0; k < A.K; ++k) {
              const unsigned char* l = s_lut + A.lut_off[k];
              const uint32_t t = A.lut_wide[k] ? (uint32_t) reinterpret_cast<const uint16_t*>(l)[cell] : (uint32_t)l[cell];
              const int rs = A.rep_shift[k];
              atomicAdd(&s_hist[A.shist_off[k] + (int)(t << rs) + (lane & ((1 << rs) - 1))], 1u);
              if (k == 0) t0 = t;
            }
          }
          if (ASSIGN) out_assign[s] = (uint16_t)t0;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_empty[stage]));
    }
    nv = __reduce_add_sync(kFull, nv);
    if (lane == 0 && nv) atomicAdd(&s_nvalid[buf], nv);
    consumer_sync();
    if (ctid == 0) {
      if (a.chunks_per_frame == 1) a.nvalid[f] = s_nvalid[buf];
      else if (s_nvalid[buf]) atomicAdd(&a.nvalid[f], s_nvalid[buf]);
      s_nvalid[buf] = 0u;
    }
    uint32_t* __restrict__ row = A.ihist + f * (int64_t)A.sumT;
    for (int k = 0; k < A.K; ++k) {
      const int rs = A.rep_shift[k];
      for (int t = ctid; t < A.T[k]; t += kConsumerWarps * 32) {
        uint32_t v = 0;
        for (int r = 0; r < (1 << rs); ++r) {
          v += s_hist[A.shist_off[k] + (t << rs) + r];
          s_hist[A.shist_off[k] + (t << rs) + r] = 0u;
        }
        if (a.chunks_per_frame == 1) row[A.hist_off[k] + t] = v;
        else if (v) atomicAdd(&row[A.hist_off[k] + t], v);
      }
    }
    // no second barrier: the next item counts into the other copy, and its own barrier orders this flush before
    // the copy is used again two items later
  }
  if (bad) atomicOr(a.flags, (uint32_t)VET_FLAG_OUT_OF_RANGE);
}

}  // namespace vet
