// Stage 4, fast paths (literal mode).  Same closed form as vet_transition.cuh, with the
// (prev,cur) bookkeeping kept entirely in shared memory:
//
//   DENSE   T*T first-index table (u32) when it fits (T <= ~200): one atomicMin per user,
//           no hashing, rows scanned by warps;
//   HASH    open-addressing table of kHashSlots (prev,cur) keys for larger T.  Real
//           trajectories touch a few neighbours per tile (<= ~7T pairs); if a frame pair has
//           more distinct pairs than the table holds, the CTA falls back to its global table
//           for that pair (same code as k_transition), so any input stays exact.
//
// One CTA per frame pair; tile counts are processed one after the other, the cell ids of
// the two frames are re-read from L2 for every pass (3 passes per tile count).
#pragma once
#include "vet_transition.cuh"

namespace vet {

constexpr int kTrThreads = 1024;
constexpr uint32_t kHashSlots = 16384;          // shared-memory pair table (keys + firsts + list = 192 KB)
constexpr uint32_t kHashLimit = kHashSlots * 3 / 4;
constexpr int kMaxProbes = 128;

enum : int { kTrDense = 0, kTrHash = 1, kTrGlobal = 2 };

struct Transition2Args {
  TransitionArgs t;               // shared fields; t.cap / t.g_tables describe the per-CTA global fallback table
  int mode[kMaxTileCounts];       // kTrDense / kTrHash / kTrGlobal per tile count
  int lut_smem[kMaxTileCounts];   // 1: the tile count's LUT is staged in shared memory (uint8 when T <= 255 else uint16)
  const uint8_t* lut8[kMaxTileCounts];  // uint8 LUTs (null when T > 255)
  int lut_area_off;               // byte offset of the LUT staging area inside dynamic shared memory
  uint32_t* pair_scratch;         // [gridDim.x, U] (prev | cur << 16) of the current frame pair and tile count
  const uint32_t* only_rows;      // optional [F-1]: process only the flagged rows (left over by k_transition3)
};

// per previous tile: users, first user, distinct keys, count of the latest key, latest key
struct TileArrays {
  unsigned long long* latest;
  uint32_t *m, *first, *kcnt, *wcnt;
};

__device__ __forceinline__ void clear_tiles(const TileArrays& ta, int T) {
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    ta.latest[t] = 0ull;
    ta.m[t] = 0u;
    ta.first[t] = kEmpty;
    ta.kcnt[t] = 0u;
    ta.wcnt[t] = 0u;
  }
}

// Tile lookup of one cell id: shared-memory LUT (uint8 / uint16) when it was staged, else the
// global uint16 LUT.  A 32-lane gather in a 20-40 KB global table touches ~30 cache lines (one
// L1 wavefront each); in shared memory it costs its bank-conflict degree (3-4).
struct LutView {
  const uint8_t* s8;
  const uint16_t* s16;
  const uint16_t* g16;
  __device__ __forceinline__ uint32_t operator()(int cell) const {
    if (s8) return s8[cell];
    if (s16) return s16[cell];
    return g16[cell];
  }
};

constexpr int kUnr = 8;
constexpr uint32_t kNoPair = 0xFFFFFFFFu;

// pass 1: looks every common user's (prev, cur) tiles up ONCE, stores them packed for the later
// passes, counts users per previous tile and finds each tile's first user (EU:271-276, 289-292).
// kUnr users per thread: all cell ids first, then all lookups, then the atomics, so 2*kUnr
// independent loads are in flight per thread.
__device__ __forceinline__ void pass_count(const TransitionArgs& a, const LutView& lut, int64_t prow, int64_t crow,
                                           const TileArrays& ta, uint32_t* __restrict__ pairs, bool write_pairs0) {
  for (int64_t u0 = 0; u0 < a.U; u0 += (int64_t)blockDim.x * kUnr) {
    int cp[kUnr], cc[kUnr];
#pragma unroll
    for (int j = 0; j < kUnr; ++j) {
      const int64_t u = u0 + threadIdx.x + (int64_t)j * blockDim.x;
      cp[j] = -1;
      cc[j] = -1;
      if (u < a.U) {
        cp[j] = load_cell(a, prow + u);
        cc[j] = load_cell(a, crow + u);
      }
    }
    uint32_t pc[kUnr];
#pragma unroll
    for (int j = 0; j < kUnr; ++j) pc[j] = (cp[j] >= 0 && cc[j] >= 0) ? (lut(cp[j]) | (lut(cc[j]) << 16)) : kNoPair;
#pragma unroll
    for (int j = 0; j < kUnr; ++j) {
      const int64_t u = u0 + threadIdx.x + (int64_t)j * blockDim.x;
      if (u >= a.U) continue;
      pairs[u] = pc[j];
      if (pc[j] != kNoPair) {
        const uint32_t p = pc[j] & 0xFFFFu;
        atomicAdd(&ta.m[p], 1u);
        // minima only decrease: a plain (broadcast) load filters out almost every atomic, since user
        // indices arrive in roughly increasing order
        if ((uint32_t)u < *(volatile uint32_t*)&ta.first[p]) atomicMin(&ta.first[p], (uint32_t)u);
      }
      if (write_pairs0) *reinterpret_cast<uint32_t*>(a.pairs0 + 2 * (prow + u)) = pc[j];  // (prev, cur) or (0xFFFF, 0xFFFF)
    }
  }
}

// later passes: the packed pairs of the common users, kUnr per thread
template <typename F>
__device__ __forceinline__ void for_pairs(const TransitionArgs& a, const uint32_t* __restrict__ pairs, F&& fn) {
  for (int64_t u0 = 0; u0 < a.U; u0 += (int64_t)blockDim.x * kUnr) {
    uint32_t pc[kUnr];
#pragma unroll
    for (int j = 0; j < kUnr; ++j) {
      const int64_t u = u0 + threadIdx.x + (int64_t)j * blockDim.x;
      pc[j] = u < a.U ? __ldcg(pairs + u) : kNoPair;
    }
#pragma unroll
    for (int j = 0; j < kUnr; ++j)
      if (pc[j] != kNoPair) fn((uint32_t)(u0 + threadIdx.x + (int64_t)j * blockDim.x), pc[j] & 0xFFFFu, pc[j] >> 16);
  }
}

// pass 3: occurrences of the latest key among the non-first users (EU:308, stale weight)
__device__ __forceinline__ void pass_latest(const TransitionArgs& a, const uint32_t* __restrict__ pairs, const TileArrays& ta) {
  for_pairs(a, pairs, [&](uint32_t u, uint32_t p, uint32_t c) {
    if (ta.first[p] != u && (uint32_t)ta.latest[p] == c) atomicAdd(&ta.wcnt[p], 1u);
  });
}

// EU:297-330 from the per-tile arrays; returns the normalised entropy (all threads)
__device__ __forceinline__ double literal_entropy(const TileArrays& ta, int T, double total, double* red) {
  double acc = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const uint32_t m = ta.m[t];
    if (m == 0u) continue;
    const double Kp = 1.0 + (double)ta.kcnt[t];
    const double wp = (m == 1u) ? 1.0 : (double)ta.wcnt[t];
    const double tp = wp / (double)m;
    acc += -((double)m / total) * (Kp * (tp * log2(tp)));
  }
  const double Hs = block_sum(acc, red);
  const double n = (total > (double)T) ? (double)T : total;
  const double q = 1.0 / n;
  return Hs / (n * -q * log2(q));
}

__global__ void __launch_bounds__(kTrThreads, 1) k_transition2(Transition2Args A, int maxT) {
  const TransitionArgs& a = A.t;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileArrays ta;
  ta.latest = reinterpret_cast<unsigned long long*>(smem_raw);
  ta.m = reinterpret_cast<uint32_t*>(ta.latest + maxT);
  ta.first = ta.m + maxT;
  ta.kcnt = ta.first + maxT;
  ta.wcnt = ta.kcnt + maxT;
  uint32_t* s_tab = ta.wcnt + maxT;  // dense T*T table, or hash keys | firsts | list
  __shared__ double s_red[32];
  __shared__ uint32_t s_used, s_overflow;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  // the table area starts in its "empty" state; every path leaves it empty again
  bool any_dense = false, any_hash = false;
  int denseT = 0;
  for (int k = 0; k < a.K; ++k) {
    if (A.mode[k] == kTrDense) {
      any_dense = true;
      denseT = max(denseT, a.T[k]);
    }
    if (A.mode[k] == kTrHash) any_hash = true;
  }
  {
    const uint32_t words = max(any_dense ? (uint32_t)(denseT * denseT) : 0u, any_hash ? 3u * kHashSlots : 0u);
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) s_tab[i] = kEmpty;
  }
  if (threadIdx.x == 0) {
    s_used = 0u;
    s_overflow = 0u;
  }
  __syncthreads();

  PairTable gtb;  // global fallback table of this CTA
  {
    uint32_t* base = a.g_tables + (size_t)blockIdx.x * 4 * a.cap;
    gtb.keys = base;
    gtb.firsts = base + a.cap;
    gtb.counts = base + 2 * (size_t)a.cap;
    gtb.list = base + 3 * (size_t)a.cap;
    gtb.mask = a.cap - 1;
  }
  PairTable stb;  // shared-memory hash table
  stb.keys = s_tab;
  stb.firsts = s_tab + kHashSlots;
  stb.counts = nullptr;
  stb.list = s_tab + 2 * kHashSlots;
  stb.mask = kHashSlots - 1;

  uint32_t* __restrict__ pairs = A.pair_scratch + (size_t)blockIdx.x * a.U;
  unsigned char* s_lutarea = smem_raw + A.lut_area_off;
  int staged_k = -1;
  for (int64_t r = blockIdx.x; r < a.F - 1; r += gridDim.x) {
    if (A.only_rows && A.only_rows[r] == 0u) continue;  // uniform over the CTA
    const int64_t prow = r * a.U, crow = (r + 1) * a.U;
    double esum = 0.0;
    for (int k = 0; k < a.K; ++k) {
      const int T = a.T[k];
      LutView lut{nullptr, nullptr, a.lut[k]};
      if (A.lut_smem[k]) {
        const bool narrow = A.lut8[k] != nullptr;
        if (staged_k != k) {  // with one tile count the LUT is staged once for the whole launch
          const int bytes = a.C * (narrow ? 1 : 2);
          const uint4* __restrict__ src = reinterpret_cast<const uint4*>(narrow ? (const void*)A.lut8[k] : (const void*)a.lut[k]);
          uint4* dst = reinterpret_cast<uint4*>(s_lutarea);
          for (int i = threadIdx.x; i < (bytes + 15) / 16; i += blockDim.x) dst[i] = __ldg(src + i);
          staged_k = k;
        }
        if (narrow) lut.s8 = s_lutarea;
        else lut.s16 = reinterpret_cast<const uint16_t*>(s_lutarea);
      }
      clear_tiles(ta, T);
      __syncthreads();
      pass_count(a, lut, prow, crow, ta, pairs, k == 0 && a.pairs0 != nullptr);
      __syncthreads();
      unsigned long long tloc = 0;
      for (int t = threadIdx.x; t < T; t += blockDim.x) tloc += ta.m[t];
      const double total = block_sum((double)tloc, s_red);

      int mode = A.mode[k];
      if (mode == kTrDense) {
        // pass 2: smallest non-first user index of every (prev,cur)
        for_pairs(a, pairs, [&](uint32_t u, uint32_t p, uint32_t c) {
          if (ta.first[p] != u) {
            uint32_t* s = &s_tab[p * (uint32_t)T + c];
            if (u < *(volatile uint32_t*)s) atomicMin(s, u);
          }
        });
        __syncthreads();
        // one warp per previous tile: distinct keys and the key seen first latest; rows reset on the way
        for (int p = wid; p < T; p += nw) {
          uint32_t cnt = 0;
          unsigned long long best = 0ull;
          for (int c = lane; c < T; c += 32) {
            const uint32_t fu = s_tab[p * T + c];
            if (fu != kEmpty) {
              ++cnt;
              best = max(best, ((unsigned long long)fu << 32) | (uint32_t)c);
              s_tab[p * T + c] = kEmpty;
            }
          }
          cnt = __reduce_add_sync(kFull, cnt);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(kFull, best, o));
          if (lane == 0) {
            ta.kcnt[p] = cnt;
            ta.latest[p] = best;
          }
        }
        __syncthreads();
      } else {
        if (mode == kTrHash) {
          for_pairs(a, pairs, [&](uint32_t u, uint32_t p, uint32_t c) {
            if (ta.first[p] == u || *(volatile uint32_t*)&s_overflow) return;
            const uint32_t key = p * (uint32_t)T + c;
            uint32_t slot = (key * 2654435761u) & stb.mask;
            for (int probe = 0;; ++probe) {
              // most users repeat a key that is already in the table: a plain load finds it without
              // the serialised same-address CAS, and the first-user minimum rarely needs its atomic
              uint32_t old = *(volatile uint32_t*)&stb.keys[slot];
              if (old == kEmpty) {
                old = atomicCAS(&stb.keys[slot], kEmpty, key);
                if (old == kEmpty) {
                  const uint32_t pos = atomicAdd(&s_used, 1u);
                  if (pos < kHashLimit) stb.list[pos] = slot;
                  else s_overflow = 1u;
                  old = key;
                }
              }
              if (old == key) {
                if (u < *(volatile uint32_t*)&stb.firsts[slot]) atomicMin(&stb.firsts[slot], u);
                break;
              }
              if (probe >= kMaxProbes) {
                s_overflow = 1u;
                break;
              }
              slot = (slot + 1) & stb.mask;
            }
          });
          __syncthreads();
          if (s_overflow) {  // too many distinct pairs for shared memory: wipe it and redo the pair in the global table
            for (uint32_t i = threadIdx.x; i < 3u * kHashSlots; i += blockDim.x) s_tab[i] = kEmpty;
            __syncthreads();
            if (threadIdx.x == 0) {
              s_used = 0u;
              s_overflow = 0u;
            }
            __syncthreads();
            mode = kTrGlobal;
          }
        }
        if (mode == kTrGlobal) {
          for_pairs(a, pairs, [&](uint32_t u, uint32_t p, uint32_t c) {
            if (ta.first[p] != u) pair_insert<false>(gtb, p * (uint32_t)T + c, u, &s_used, false);
          });
          __syncthreads();
        }
        // per previous tile: number of distinct keys and the latest-first-seen key; slots reset on the way
        const uint32_t used = s_used;
        for (uint32_t i = threadIdx.x; i < used; i += blockDim.x) {
          uint32_t slot, key, fu;
          if (mode == kTrHash) {
            slot = stb.list[i];
            key = stb.keys[slot];
            fu = stb.firsts[slot];
            stb.keys[slot] = kEmpty;
            stb.firsts[slot] = kEmpty;
            stb.list[i] = kEmpty;  // the area is shared with the dense table of other tile counts
          } else {
            slot = tb_load<false>(&gtb.list[i]);
            key = tb_load<false>(&gtb.keys[slot]);
            fu = tb_load<false>(&gtb.firsts[slot]);
            tb_store<false>(&gtb.keys[slot], kEmpty);
            tb_store<false>(&gtb.firsts[slot], kEmpty);
          }
          const uint32_t p = key / (uint32_t)T, c = key % (uint32_t)T;
          atomicAdd(&ta.kcnt[p], 1u);
          atomicMax(&ta.latest[p], ((unsigned long long)fu << 32) | c);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_used = 0u;
      }
      pass_latest(a, pairs, ta);
      __syncthreads();
      double e = literal_entropy(ta, T, total, s_red);
      if (total == 0.0) {
        e = qnan;
        if (threadIdx.x == 0) atomicOr(a.flags, (uint32_t)VET_FLAG_NO_COMMON_USER);
      }
      if (threadIdx.x == 0 && a.per_k) a.per_k[k * a.per_k_stride + r] = e;
      if (k == 0 && a.prev_count0)
        for (int t = threadIdx.x; t < T; t += blockDim.x) a.prev_count0[r * (int64_t)T + t] = (int32_t)ta.m[t];
      esum += e;
      __syncthreads();
    }
    if (threadIdx.x == 0) a.entropy[r] = esum / (double)a.K;  // TA:160
  }
}

}  // namespace vet
