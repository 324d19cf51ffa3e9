// Stage 4: tile-to-tile transitions of consecutive frames and the transition
// entropy, with the reference's literal (order-dependent) bookkeeping.
//
// compute_transition_entropy (EU:259-330), users in packed-row order.  For a
// previous tile p let m_p be its user count.  The reference's inner dict is
// keyed by `int` for the FIRST user of p and by `Vector` for later users
// (EU:280-287), and its entropy loop reuses the weight of the last-inserted key
// (EU:312-315).  In closed form, per p:
//   K_p = 1 + #distinct cur tiles among the non-first users of p
//   w_p = 1 if m_p == 1 else the number of non-first users whose cur tile is the
//         one that appears LATEST for the first time among them
//   H   = sum_p (m_p/total) * -(K_p * q_p log2 q_p),  q_p = w_p/m_p
// Order dependence enters only through minima of user indices, which atomicMin
// resolves exactly, so the result is deterministic and bit-exact in its counts.
#pragma once
#include "vet_common.cuh"

namespace vet {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;

struct PairTable {      // open-addressing table of (prev,cur) pairs of one frame pair
  uint32_t* keys;       // [cap] p*T+c, kEmpty when free
  uint32_t* firsts;     // [cap] smallest user index inserted under the key
  uint32_t* counts;     // [cap] occurrences (textbook mode only)
  uint32_t* list;       // [cap] occupied slots in claim order
  uint32_t mask;        // cap-1 (cap is a power of two)
};

// Table words are touched with atomics (resolved in L2 for the global variant) and
// with the plain accesses below; the global variant must bypass L1 so that a line
// cached by an earlier frame pair is never read stale.
template <bool SMEM>
__device__ __forceinline__ uint32_t tb_load(const uint32_t* p) {
  if (SMEM) return *p;
  return __ldcg(p);
}
template <bool SMEM>
__device__ __forceinline__ void tb_store(uint32_t* p, uint32_t v) {
  if (SMEM) *p = v;
  else __stcg(p, v);
}

template <bool SMEM>
__device__ __forceinline__ uint32_t pair_insert(const PairTable& tb, uint32_t key, uint32_t user, uint32_t* n_used,
                                                bool count) {
  uint32_t slot = (key * 2654435761u) & tb.mask;
  while (true) {
    const uint32_t old = atomicCAS(&tb.keys[slot], kEmpty, key);
    if (old == kEmpty) tb_store<SMEM>(&tb.list[atomicAdd(n_used, 1u)], slot);
    if (old == kEmpty || old == key) {
      atomicMin(&tb.firsts[slot], user);
      if (count) atomicAdd(&tb.counts[slot], 1u);
      return slot;
    }
    slot = (slot + 1) & tb.mask;
  }
}

struct TransitionArgs {
  const uint16_t* cell16;  // [F,U] cell ids (0xFFFF = missing) or null
  const int32_t* cell32;   // [F,U] (-1 = missing) or null
  int64_t F, U;            // rows r in [0, F-1) pair frames (r, r+1)
  int C;                   // cells per LUT (k_transition2 stages LUTs in shared memory)
  int K;
  int T[kMaxTileCounts];
  const uint16_t* lut[kMaxTileCounts];
  double* entropy;         // [F-1]
  double* per_k;           // [K, per_k_stride] or null
  int64_t per_k_stride;
  int32_t* prev_count0;    // [F-1,T0] or null
  uint16_t* pairs0;        // [F-1,U,2] or null
  int mode;
  uint32_t* flags;
  const uint32_t* nvalid;  // [F] present users per frame from the streaming kernel of the same batch, or null
  // pair tables: in shared memory (SMEM variant), else one global region per block
  uint32_t cap;            // slots per table (power of two)
  uint32_t* g_tables;      // [gridDim.x, 4, cap] for the global variant (keys pre-set to kEmpty, firsts to kEmpty, counts to 0)
};

__device__ __forceinline__ int load_cell(const TransitionArgs& a, int64_t idx) {
  if (a.cell16) {
    const uint16_t v = a.cell16[idx];
    return v == 0xFFFF ? -1 : (int)v;
  }
  return a.cell32[idx];
}

template <bool SMEM>
__global__ void __launch_bounds__(512, 1) k_transition(TransitionArgs a, int maxT) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // per-tile arrays
  unsigned long long* s_latest = reinterpret_cast<unsigned long long*>(smem_raw);  // (first'<<32)|cur
  uint32_t* s_m = reinterpret_cast<uint32_t*>(s_latest + maxT);
  uint32_t* s_first = s_m + maxT;
  uint32_t* s_kcnt = s_first + maxT;
  uint32_t* s_wcnt = s_kcnt + maxT;
  uint32_t* s_tab = s_wcnt + maxT;
  __shared__ double s_red[32];
  __shared__ uint32_t s_used;
  PairTable tb;
  uint32_t* base = SMEM ? s_tab : a.g_tables + (size_t)blockIdx.x * 4 * a.cap;
  tb.keys = base;
  tb.firsts = base + a.cap;
  tb.counts = base + 2 * (size_t)a.cap;
  tb.list = base + 3 * (size_t)a.cap;
  tb.mask = a.cap - 1;
  if (SMEM) {
    for (uint32_t i = threadIdx.x; i < a.cap; i += blockDim.x) {
      tb.keys[i] = kEmpty;
      tb.firsts[i] = kEmpty;
      tb.counts[i] = 0u;
    }
  }
  if (threadIdx.x == 0) s_used = 0u;
  __syncthreads();
  const bool textbook = a.mode == VET_TRANSITION_TEXTBOOK;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  for (int64_t r = blockIdx.x; r < a.F - 1; r += gridDim.x) {
    const int64_t prow = r * a.U, crow = (r + 1) * a.U;
    double esum = 0.0;
    for (int k = 0; k < a.K; ++k) {
      const int T = a.T[k];
      const uint16_t* __restrict__ lut = a.lut[k];
      for (int t = threadIdx.x; t < T; t += blockDim.x) {
        s_latest[t] = 0ull;
        s_m[t] = 0u;
        s_first[t] = kEmpty;
        s_kcnt[t] = 0u;
        s_wcnt[t] = 0u;
      }
      __syncthreads();
      // pass 1: users per previous tile and the first user of each (EU:271-276, 289-292)
      for (int64_t u = threadIdx.x; u < a.U; u += blockDim.x) {
        const int cp = load_cell(a, prow + u), cc = load_cell(a, crow + u);
        const bool ok = (cp >= 0) && (cc >= 0);  // present in both frames (EU:260)
        uint16_t p = VET_MISSING, c = VET_MISSING;
        if (ok) {
          p = lut[cp];
          c = lut[cc];
          atomicAdd(&s_m[p], 1u);
          atomicMin(&s_first[p], (uint32_t)u);
        }
        if (k == 0 && a.pairs0) {
          a.pairs0[2 * (prow + u) + 0] = p;
          a.pairs0[2 * (prow + u) + 1] = c;
        }
      }
      __syncthreads();
      // pass 2: distinct (p,c) among the non-first users (all users in textbook mode)
      for (int64_t u = threadIdx.x; u < a.U; u += blockDim.x) {
        const int cp = load_cell(a, prow + u), cc = load_cell(a, crow + u);
        if (cp < 0 || cc < 0) continue;
        const uint32_t p = lut[cp], c = lut[cc];
        if (!textbook && s_first[p] == (uint32_t)u) continue;
        pair_insert<SMEM>(tb, p * (uint32_t)T + c, (uint32_t)u, &s_used, textbook);
      }
      __syncthreads();
      // per previous tile: number of distinct keys and the latest-first-seen key
      const uint32_t used = s_used;
      double tb_acc = 0.0;
      unsigned long long tloc = 0;
      for (int t = threadIdx.x; t < T; t += blockDim.x) tloc += s_m[t];
      const double total = block_sum((double)tloc, s_red);
      for (uint32_t i = threadIdx.x; i < used; i += blockDim.x) {
        const uint32_t slot = tb_load<SMEM>(&tb.list[i]);
        const uint32_t key = tb_load<SMEM>(&tb.keys[slot]), fu = tb_load<SMEM>(&tb.firsts[slot]);
        const uint32_t p = key / (uint32_t)T, c = key % (uint32_t)T;
        if (textbook) {
          const double cnt = (double)tb_load<SMEM>(&tb.counts[slot]);
          tb_acc -= (cnt / total) * log2(cnt / (double)s_m[p]);
          tb_store<SMEM>(&tb.counts[slot], 0u);
        } else {
          atomicAdd(&s_kcnt[p], 1u);
          atomicMax(&s_latest[p], ((unsigned long long)fu << 32) | c);
        }
        tb_store<SMEM>(&tb.keys[slot], kEmpty);
        tb_store<SMEM>(&tb.firsts[slot], kEmpty);
      }
      __syncthreads();
      if (threadIdx.x == 0) s_used = 0u;
      double Hs;
      if (textbook) {
        Hs = block_sum(tb_acc, s_red);
      } else {
        // pass 3: occurrences of the latest key among the non-first users (EU:308, stale weight)
        for (int64_t u = threadIdx.x; u < a.U; u += blockDim.x) {
          const int cp = load_cell(a, prow + u), cc = load_cell(a, crow + u);
          if (cp < 0 || cc < 0) continue;
          const uint32_t p = lut[cp], c = lut[cc];
          if (s_first[p] == (uint32_t)u) continue;
          if ((uint32_t)s_latest[p] == c) atomicAdd(&s_wcnt[p], 1u);
        }
        __syncthreads();
        double acc = 0.0;
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
          const uint32_t m = s_m[t];
          if (m == 0u) continue;
          const double Kp = 1.0 + (double)s_kcnt[t];
          const double wp = (m == 1u) ? 1.0 : (double)s_wcnt[t];
          const double tp = wp / (double)m;                 // EU:313
          const double cell = Kp * (tp * log2(tp));         // EU:312-315, K_p equal addends
          acc += -((double)m / total) * cell;               // EU:301,317-318
        }
        Hs = block_sum(acc, s_red);
      }
      // EU:321-330
      const double n = (total > (double)T) ? (double)T : total;
      const double q = 1.0 / n;
      const double mx = n * -q * log2(q);
      double e = Hs / mx;
      if (total == 0.0) {
        e = qnan;
        if (threadIdx.x == 0) atomicOr(a.flags, (uint32_t)VET_FLAG_NO_COMMON_USER);
      }
      if (threadIdx.x == 0 && a.per_k) a.per_k[k * a.per_k_stride + r] = e;
      if (k == 0 && a.prev_count0)
        for (int t = threadIdx.x; t < T; t += blockDim.x) a.prev_count0[r * (int64_t)T + t] = (int32_t)s_m[t];
      esum += e;
      __syncthreads();
    }
    if (threadIdx.x == 0) a.entropy[r] = esum / (double)a.K;  // TA:160
  }
}

}  // namespace vet
