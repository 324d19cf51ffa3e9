// Stage 4, two-pass kernel (literal mode, one tile count per launch).
//
// Same closed form as vet_transition.cuh, reorganised so that the users of a frame
// pair are walked TWICE instead of three times and every per-user step is a handful of
// instructions.  For a previous tile p with users S_p (m_p of them, first user f_p = min S_p) the
// reference's bookkeeping needs, over the NON-first users N_p = S_p \ {f_p}:
//     K_p = 1 + #distinct cur tiles in N_p
//     w_p = occurrences in N_p of the cur tile whose first appearance in N_p is the latest.
// Pass 1 builds A[p][c] = min{u in S_p : cur(u) = c} over ALL users (one pre-checked atomicMin per
// user) and m_p.  From row p of A:  D_p = #entries, (f_p, c_f) = the minimum entry (the first user
// and its cur tile), (x_p, l'_p) = the maximum entry among c != c_f.  Excluding f_p only changes
// the column c_f, so pass 2 collects, per p:  cnt_cf = #{u : cur = c_f},  other = #{u : cur != c_f}
// (m_p = cnt_cf + other),  cnt_l = #{u : cur = l'_p}  and  early = [some u != f_p with cur = c_f comes
// before x_p].  Then
//     distinct(N_p) = D_p - 1 + [cnt_cf >= 2]
//     latest key    = c_f if cnt_cf >= 2 and (no l'_p or not early) else l'_p
//     w_p           = cnt_cf - 1 resp. cnt_l            (m_p == 1: w_p = 1)
// Order enters only through minima of user indices -> deterministic, counts bit-exact.
//
//   DENSE  A is a T x T table in shared memory (T <= ~220), rows scanned by warps;
//   HASH   A is an open-addressing table of 16384 (p,c) keys in shared memory; a frame pair with
//          more distinct pairs (or a probe chain > 128) is flagged in `redo` and recomputed by
//          k_transition2 (global tables), so any input stays exact.
//
// One CTA (1024 threads) per frame pair; users are taken 8 per thread with 128-bit loads when
// U % 8 == 0, the tile LUT is staged in shared memory when it fits.
#pragma once
#include <type_traits>

#include "vet_transition2.cuh"

namespace vet {

constexpr int kT3Threads = 1024;
constexpr uint32_t kT3Slots = 16384;
constexpr uint32_t kT3Limit = kT3Slots * 3 / 4;
constexpr int kT3Probes = 128;
constexpr int kT3TileBytes = 2 * 8 + 6 * 4 + 2;  // per-tile arrays below (the byte array needs T + 3 <= 2 T)
constexpr uint32_t kNoTile = 0x3FFFu;        // 14-bit tile fields of s_cfl
constexpr uint32_t kEarlyBit = 0x80000000u;

enum : int { kT3Dense = 0, kT3Hash = 1 };
enum : int { kLutS8 = 0, kLutS16 = 1, kLutG16 = 2, kLutIdentity = 3 };  // kLutIdentity: the input already holds tile ids

struct Transition3Args {
  const uint16_t* cell16;  // [F,U] cell ids (or tile ids with kLutIdentity), 0xFFFF = missing
  int64_t F;
  uint32_t U;
  int T;
  int C;
  const void* lut_src;     // uint8[C] (kLutS8) or uint16[C]
  int tab_off;             // byte offset of the pair table inside dynamic shared memory
  int lut_off;             // byte offset of the staged LUT
  double* out;             // [F-1] normalised entropy of this tile count
  int32_t* prev_count0;    // [F-1,T] or null
  uint16_t* pairs0;        // [F-1,U,2] or null
  uint32_t* pair_scratch;  // [gridDim.x, U]; null with kLutIdentity: pass 2 rebuilds the pairs from the rows
  uint32_t* redo;          // [F-1] rows to be recomputed by k_transition2 (HASH overflow)
  uint32_t* flags;
  const uint32_t* nvalid;  // [F] present users per frame (streaming kernel), or null: pairs of two complete frames skip the missing-user tests
  int ush;                 // >= 3 with (U - 1) >> ush <= 254: granularity of the per-tile "early" bound bytes (dense pass 2)
  const uint32_t* only_rows;   // optional [F-1]: process only the flagged pairs (left over by the fused pass, vet_stream_trans.cuh)
  const uint32_t* only_count;  // with only_rows: number of flagged pairs (0: the CTAs leave at once)
};

template <int LW>
struct Lut3 {
  const void* p;
  __device__ __forceinline__ uint32_t operator()(uint32_t cell) const {
    if (LW == kLutIdentity) return cell;
    if (LW == kLutS8) return static_cast<const uint8_t*>(p)[cell];
    if (LW == kLutS16) return static_cast<const uint16_t*>(p)[cell];
    return __ldg(static_cast<const uint16_t*>(p) + cell);
  }
};

// Row pitch of the dense table: even, so that the diagonal (p, p) -- the pair of most users -- walks all
// 32 banks (index p * (pitch + 1), pitch + 1 odd) instead of 16.
__host__ __device__ __forceinline__ uint32_t t3_row_stride(uint32_t T) { return T + (T & 1u); }

__device__ __forceinline__ uint32_t lds_u32(const uint32_t* p) { return *(const volatile uint32_t*)p; }

template <int MODE>
__device__ __forceinline__ void t3_update(uint32_t* tab, uint32_t* diag, uint32_t T, uint32_t p, uint32_t c, uint32_t u,
                                          uint32_t* s_used, uint32_t* s_overflow) {
  // minima only decrease and users arrive in roughly increasing order: a plain (broadcast) load
  // filters out almost every atomic
  if (MODE == kT3Dense) {
    uint32_t* s = tab + p * t3_row_stride(T) + c;
    if (u < lds_u32(s)) atomicMin(s, u);
  } else if (p == c) {
    // staying inside the tile is by far the most frequent pair: it has its own dense slot
    if (u < lds_u32(diag + p)) atomicMin(diag + p, u);
  } else {
    uint32_t* keys = tab;
    uint32_t* firsts = tab + kT3Slots;
    const uint32_t key = p * T + c;
    uint32_t slot = (key * 2654435761u) >> 18;  // top 14 bits
    // single exit: the warp reconverges after the loop; once the table has overflowed the row is
    // recomputed elsewhere, so further inserts are skipped
    bool done = lds_u32(s_overflow) != 0u;
#pragma unroll 1
    for (int probe = 0; !done; ++probe) {
      uint32_t old = lds_u32(keys + slot);
      if (old == kEmpty) {
        old = atomicCAS(keys + slot, kEmpty, key);
        if (old == kEmpty) {
          if (atomicAdd(s_used, 1u) >= kT3Limit) *s_overflow = 1u;
          old = key;
        }
      }
      if (old == key) {
        if (u < lds_u32(firsts + slot)) atomicMin(firsts + slot, u);
        done = true;
      } else if (probe >= kT3Probes) {
        *s_overflow = 1u;
        done = true;
      } else {
        slot = (slot + 1) & (kT3Slots - 1);
      }
    }
  }
}

template <int MODE, int LW>
__global__ void __launch_bounds__(kT3Threads, 1) k_transition3(Transition3Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t T = (uint32_t)a.T;
  unsigned long long* s_rmin = reinterpret_cast<unsigned long long*>(smem_raw);  // row minimum (f_p<<32)|c_f, ~0 = empty row
  unsigned long long* s_rmax = s_rmin + T;                                        // ((x_p+1)<<32)|l'_p, 0 = none
  uint32_t* s_other = reinterpret_cast<uint32_t*>(s_rmax + T);  // users with cur != c_f  (m_p = cnt_cf + other)
  uint32_t* s_d = s_other + T;    // distinct cur tiles of the row
  uint32_t* s_cfl = s_d + T;      // c_f | l'_p << 14 (kNoTile = none) | "early" << 31
  uint32_t* s_cf = s_cfl + T;     // cnt_cf
  uint32_t* s_cl = s_cf + T;      // cnt_l
  uint32_t* s_diag = s_cl + T;    // HASH: A[p][p]
  uint8_t* s_ub = reinterpret_cast<uint8_t*>(s_diag + T);  // DENSE: only users with (u >> ush) < s_ub[p] can still set "early"
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem_raw + a.tab_off);
  __shared__ double s_red[32];
  __shared__ uint32_t s_used, s_overflow, s_valid;
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  if (a.only_rows && __ldg(a.only_count) == 0u) return;  // nothing was left over
  {
    const uint32_t words = MODE == kT3Dense ? T * t3_row_stride(T) : 2u * kT3Slots;
    for (uint32_t i = tid; i < words; i += kT3Threads) s_tab[i] = kEmpty;
  }
  Lut3<LW> lut{a.lut_src};
  if (LW == kLutS8 || LW == kLutS16) {
    const int bytes = a.C * (LW == kLutS8 ? 1 : 2);
    const uint4* __restrict__ src = static_cast<const uint4*>(a.lut_src);
    uint4* dst = reinterpret_cast<uint4*>(smem_raw + a.lut_off);
    for (int i = tid; i < (bytes + 15) / 16; i += kT3Threads) dst[i] = __ldg(src + i);
    lut.p = smem_raw + a.lut_off;
  }
  if (tid == 0) {
    s_used = 0u;
    s_overflow = 0u;
    s_valid = 0u;
  }
  // Tile ids straight from the streaming kernel (identity LUT): the packed pair of a user is just its two row
  // entries, so pass 2 re-reads the rows (4 B/user, like the scratch) and pass 1 skips the 4 B/user scratch write.
  const bool scratch = !(LW == kLutIdentity && a.pair_scratch == nullptr);
  uint32_t* __restrict__ pairs = scratch ? a.pair_scratch + (size_t)blockIdx.x * a.U : nullptr;
  const uint32_t U = a.U;
  const bool vec = (U & 7u) == 0u;

  for (int64_t r = blockIdx.x; r < a.F - 1; r += gridDim.x) {
    if (a.only_rows && __ldg(a.only_rows + r) == 0u) continue;  // uniform over the CTA
    const uint16_t* __restrict__ prow = a.cell16 + r * (int64_t)U;
    const uint16_t* __restrict__ crow = prow + U;
    uint32_t* __restrict__ p0row = a.pairs0 ? reinterpret_cast<uint32_t*>(a.pairs0) + r * (int64_t)U : nullptr;
    for (uint32_t t = tid; t < T; t += kT3Threads) {
      s_rmin[t] = ~0ull;
      s_rmax[t] = 0ull;
      s_other[t] = 0u;
      s_d[t] = 0u;
      s_cfl[t] = kNoTile | (kNoTile << 14);
      s_cf[t] = 0u;
      s_cl[t] = 0u;
      if (MODE == kT3Hash) s_diag[t] = kEmpty;
      if (MODE == kT3Dense) s_ub[t] = (uint8_t)0;
    }
    __syncthreads();

    // ---- pass 1: (prev, cur) tiles of every common user, m_p, A[p][c] ----
    uint32_t nvalid = 0;
    // both frames complete (no missing user): the 0xFFFF tests of both passes drop out (block-uniform)
    const bool full = a.nvalid && __ldg(a.nvalid + r) == U && __ldg(a.nvalid + r + 1) == U;
    auto pass1_vec = [&](auto full_c) {
      constexpr bool FULL = decltype(full_c)::value;
      // One step = 8 users per thread.  Block-uniform trip count: every lane reaches the __syncwarp that
      // re-joins the warp after the (divergent) table updates.
      auto step = [&](const uint32_t u0, const uint4 vp, const uint4 vc) {
        uint32_t pending = 0u;
        uint32_t pcs[8];
        if (u0 < U) {
          const uint32_t wp[4] = {vp.x, vp.y, vp.z, vp.w}, wc[4] = {vc.x, vc.y, vc.z, vc.w};
          uint32_t pc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t cp = (wp[j >> 1] >> (16 * (j & 1))) & 0xFFFFu, cc = (wc[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
            const bool ok = FULL || (cp != 0xFFFFu && cc != 0xFFFFu);
            pc[j] = ok ? (lut(ok ? cp : 0u) | (lut(ok ? cc : 0u) << 16)) : kNoPair;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) pcs[j] = pc[j];
          if (scratch) {
            *reinterpret_cast<uint4*>(pairs + u0) = make_uint4(pc[0], pc[1], pc[2], pc[3]);
            *reinterpret_cast<uint4*>(pairs + u0 + 4) = make_uint4(pc[4], pc[5], pc[6], pc[7]);
          }
          if (p0row) {
            *reinterpret_cast<uint4*>(p0row + u0) = make_uint4(pc[0], pc[1], pc[2], pc[3]);
            *reinterpret_cast<uint4*>(p0row + u0 + 4) = make_uint4(pc[4], pc[5], pc[6], pc[7]);
          }
          if (MODE == kT3Dense) {
            // all 8 table entries are requested before the first one is compared (ncu: short-scoreboard stalls on
            // the load-compare chains were the top stall reason at 32 warps per SM)
            uint32_t slot[8], cur[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const bool ok = FULL || pc[j] != kNoPair;
              slot[j] = ok ? (pc[j] & 0xFFFFu) * t3_row_stride(T) + (pc[j] >> 16) : 0u;
              cur[j] = ok ? lds_u32(s_tab + slot[j]) : 0u;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (!FULL && pc[j] != kNoPair) ++nvalid;
              if (u0 + j < cur[j]) atomicMin(s_tab + slot[j], u0 + j);
            }
          } else {
            // users that stay in their tile (the majority) take the short dense path together; the others are
            // queued in a bit mask and go through the hash table in a compacted loop below
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (FULL || pc[j] != kNoPair) {
                const uint32_t p = pc[j] & 0xFFFFu, c = pc[j] >> 16;
                ++nvalid;
                if (p == c) t3_update<MODE>(s_tab, s_diag, T, p, c, u0 + j, &s_used, &s_overflow);
                else pending |= 1u << j;
              }
          }
        }
        if (MODE == kT3Hash) {
          while (__any_sync(kFull, pending != 0u)) {
            if (pending) {
              const int j = __ffs(pending) - 1;
              pending &= pending - 1u;
              uint32_t v = pcs[0];
#pragma unroll
              for (int q = 1; q < 8; ++q) v = j == q ? pcs[q] : v;
              t3_update<MODE>(s_tab, s_diag, T, v & 0xFFFFu, v >> 16, u0 + j, &s_used, &s_overflow);
            }
            __syncwarp();
          }
        }
        __syncwarp();
      };
      // The rows of 1M-user frames come from DRAM (ncu: 43 % of the stall samples sat on the first use of the
      // loads when they were issued in the step that consumes them).  kT3Ring register buffers, the loop unrolled
      // over them so that no buffer is copied (a rotating copy waits for the youngest load): the rows of step
      // i + kT3Ring are requested when step i starts.  prefetch.global.L2 further ahead was measured slower.
      constexpr uint32_t kStep = kT3Threads * 8u;
      constexpr int kT3Ring = MODE == kT3Dense ? 3 : 2;
      uint4 bp[kT3Ring], bc[kT3Ring];
#pragma unroll
      for (int j = 0; j < kT3Ring; ++j) {
        bp[j] = bc[j] = make_uint4(0u, 0u, 0u, 0u);
        if (tid * 8u + j * kStep < U) {
          bp[j] = __ldg(reinterpret_cast<const uint4*>(prow + tid * 8u + j * kStep));
          bc[j] = __ldg(reinterpret_cast<const uint4*>(crow + tid * 8u + j * kStep));
        }
      }
      for (uint32_t base = 0; base < U; base += kT3Ring * kStep) {
#pragma unroll
        for (int j = 0; j < kT3Ring; ++j) {
          if (base + j * kStep < U) {  // block-uniform
            const uint32_t u0 = base + j * kStep + tid * 8u;
            const uint4 vp = bp[j], vc = bc[j];
            if (u0 + kT3Ring * kStep < U) {
              bp[j] = __ldg(reinterpret_cast<const uint4*>(prow + u0 + kT3Ring * kStep));
              bc[j] = __ldg(reinterpret_cast<const uint4*>(crow + u0 + kT3Ring * kStep));
            }
            step(u0, vp, vc);
          }
        }
      }
    };
    if (vec) {
      if (full) pass1_vec(std::true_type{});
      else pass1_vec(std::false_type{});
    } else {
      for (uint32_t base = 0; base < U; base += kT3Threads) {
        const uint32_t u = base + tid;
        const uint32_t cp = u < U ? prow[u] : 0xFFFFu, cc = u < U ? crow[u] : 0xFFFFu;
        const bool ok = cp != 0xFFFFu && cc != 0xFFFFu;
        const uint32_t pc = ok ? (lut(ok ? cp : 0u) | (lut(ok ? cc : 0u) << 16)) : kNoPair;
        if (u < U) {
          if (scratch) pairs[u] = pc;
          if (p0row) p0row[u] = pc;
        }
        if (ok) {
          const uint32_t p = pc & 0xFFFFu, c = pc >> 16;
          ++nvalid;
          t3_update<MODE>(s_tab, s_diag, T, p, c, u, &s_used, &s_overflow);
        }
        __syncwarp();
      }
    }
    nvalid = __reduce_add_sync(kFull, nvalid);
    if (lane == 0 && nvalid) atomicAdd(&s_valid, nvalid);
    __syncthreads();
    // complete frames walked 8 users at a time (dense table): every user is a common user, nothing was counted
    const double total = (MODE == kT3Dense && vec && full) ? (double)U : (double)s_valid;

    if (MODE == kT3Hash && s_overflow) {
      // too many distinct pairs for the shared-memory table: wipe it, leave the row to k_transition2
      for (uint32_t i = tid; i < 2u * kT3Slots; i += kT3Threads) s_tab[i] = kEmpty;
      __syncthreads();
      if (tid == 0) {
        a.redo[r] = 1u;
        s_used = 0u;
        s_overflow = 0u;
        s_valid = 0u;
      }
      __syncthreads();
      continue;
    }

    // ---- rows of A: D_p, (f_p, c_f), (x_p, l'_p); the table is left empty ----
    if (MODE == kT3Dense) {
      for (uint32_t p = wid; p < T; p += kT3Threads / 32) {
        uint32_t* row = s_tab + p * t3_row_stride(T);
        uint32_t cnt = 0;
        unsigned long long best = ~0ull;
        for (uint32_t c = lane; c < T; c += 32) {
          const uint32_t v = row[c];
          if (v != kEmpty) {
            ++cnt;
            best = min(best, ((unsigned long long)v << 32) | c);
          }
        }
        cnt = __reduce_add_sync(kFull, cnt);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(kFull, best, o));
        if (best == ~0ull) continue;  // no user came from tile p (warp-uniform)
        const uint32_t cf = (uint32_t)best;
        unsigned long long top = 0ull;
        for (uint32_t c = lane; c < T; c += 32) {
          const uint32_t v = row[c];
          if (v != kEmpty) {
            if (c != cf) top = max(top, ((unsigned long long)(v + 1u) << 32) | c);
            row[c] = 0u;  // from here on the entry COUNTS the users of (p, c): pass 2
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) top = max(top, __shfl_xor_sync(kFull, top, o));
        if (lane == 0) {
          s_d[p] = cnt;
          s_rmin[p] = best;
          s_rmax[p] = top;
          s_cfl[p] = cf | ((top ? (uint32_t)top & kNoTile : kNoTile) << 14);
          // "early" can only be set by a user below x_p (and only if there is an l'_p at all)
          s_ub[p] = top ? (uint8_t)min((((uint32_t)(top >> 32) - 1u) >> a.ush) + 1u, 255u) : (uint8_t)0;
        }
      }
      __syncthreads();
    } else {
      uint32_t* keys = s_tab;
      uint32_t* firsts = s_tab + kT3Slots;
      for (uint32_t p = tid; p < T; p += kT3Threads) {
        const uint32_t dg = s_diag[p];
        if (dg != kEmpty) {
          s_rmin[p] = ((unsigned long long)dg << 32) | p;
          s_d[p] = 1u;
        }
      }
      __syncthreads();
      for (uint32_t s = tid; s < kT3Slots; s += kT3Threads) {
        const uint32_t key = keys[s];
        if (key != kEmpty) {
          const uint32_t p = key / T, c = key - p * T;
          atomicMin(&s_rmin[p], ((unsigned long long)firsts[s] << 32) | c);
          atomicAdd(&s_d[p], 1u);
        }
      }
      __syncthreads();
      for (uint32_t p = tid; p < T; p += kT3Threads) {
        const uint32_t dg = s_diag[p];
        if (dg != kEmpty && p != (uint32_t)s_rmin[p]) s_rmax[p] = ((unsigned long long)(dg + 1u) << 32) | p;
      }
      __syncthreads();
      for (uint32_t s = tid; s < kT3Slots; s += kT3Threads) {
        const uint32_t key = keys[s];
        if (key != kEmpty) {
          const uint32_t p = key / T, c = key - p * T;
          if (c != (uint32_t)s_rmin[p]) atomicMax(&s_rmax[p], ((unsigned long long)(firsts[s] + 1u) << 32) | c);
          keys[s] = kEmpty;
          firsts[s] = kEmpty;
        }
      }
      __syncthreads();
      for (uint32_t p = tid; p < T; p += kT3Threads) {
        const unsigned long long best = s_rmin[p], top = s_rmax[p];
        if (best != ~0ull) s_cfl[p] = ((uint32_t)best & kNoTile) | ((top ? (uint32_t)top & kNoTile : kNoTile) << 14);
      }
      __syncthreads();
    }

    // ---- pass 2: users per previous tile split into cur == c_f / other, occurrences of l'_p, and whether a
    // non-first user of (p, c_f) comes before the first user of l'_p ("early": then l'_p is the latest key).
    // One 32-bit load and one increment per user; the early test stops once its bit is set.
    auto second_pass = [&](auto full_c, uint32_t u, uint32_t pc) {
      if (!decltype(full_c)::value && pc == kNoPair) return;
      const uint32_t p = pc & 0xFFFFu, c = pc >> 16;
      const uint32_t w = lds_u32(&s_cfl[p]);
      if (c == (w & kNoTile)) {
        atomicAdd(&s_cf[p], 1u);
        if (!(w & kEarlyBit) && ((w >> 14) & kNoTile) != kNoTile) {
          const uint32_t f = reinterpret_cast<const uint32_t*>(s_rmin + p)[1];
          const uint32_t x = reinterpret_cast<const uint32_t*>(s_rmax + p)[1] - 1u;
          if (u != f && u < x) atomicOr(&s_cfl[p], kEarlyBit);
        }
      } else {
        atomicAdd(&s_other[p], 1u);
        if (c == ((w >> 14) & kNoTile)) atomicAdd(&s_cl[p], 1u);
      }
    };
    auto pass2 = [&](auto full_c) {
      constexpr bool FULL = decltype(full_c)::value;
      if (!scratch) {
        // identity LUT: pairs rebuilt from the two rows
        auto from_rows = [&](uint32_t cp, uint32_t cc) { return (FULL || (cp != 0xFFFFu && cc != 0xFFFFu)) ? (cp | (cc << 16)) : kNoPair; };
        if (vec) {
          uint4 np = make_uint4(0u, 0u, 0u, 0u), nc = np;
          if (tid * 8u < U) {
            np = __ldg(reinterpret_cast<const uint4*>(prow + tid * 8u));
            nc = __ldg(reinterpret_cast<const uint4*>(crow + tid * 8u));
          }
          for (uint32_t u0 = tid * 8u; u0 < U; u0 += kT3Threads * 8u) {
            const uint4 vp = np, vc = nc;
            if (u0 + kT3Threads * 8u < U) {
              np = __ldg(reinterpret_cast<const uint4*>(prow + u0 + kT3Threads * 8u));
              nc = __ldg(reinterpret_cast<const uint4*>(crow + u0 + kT3Threads * 8u));
            }
            const uint32_t wp[4] = {vp.x, vp.y, vp.z, vp.w}, wc[4] = {vc.x, vc.y, vc.z, vc.w};
#pragma unroll
            for (int j = 0; j < 8; ++j)
              second_pass(full_c, u0 + j, from_rows((wp[j >> 1] >> (16 * (j & 1))) & 0xFFFFu, (wc[j >> 1] >> (16 * (j & 1))) & 0xFFFFu));
          }
        } else {
          for (uint32_t u = tid; u < U; u += kT3Threads) second_pass(full_c, u, from_rows(prow[u], crow[u]));
        }
      } else if (vec) {
        uint4 n0 = make_uint4(0u, 0u, 0u, 0u), n1 = n0;  // prefetched like in pass 1
        if (tid * 8u < U) {
          n0 = __ldcg(reinterpret_cast<const uint4*>(pairs + tid * 8u));
          n1 = __ldcg(reinterpret_cast<const uint4*>(pairs + tid * 8u + 4));
        }
        for (uint32_t u0 = tid * 8u; u0 < U; u0 += kT3Threads * 8u) {
          const uint4 a0 = n0, a1 = n1;
          if (u0 + kT3Threads * 8u < U) {
            n0 = __ldcg(reinterpret_cast<const uint4*>(pairs + u0 + kT3Threads * 8u));
            n1 = __ldcg(reinterpret_cast<const uint4*>(pairs + u0 + kT3Threads * 8u + 4));
          }
          second_pass(full_c, u0, a0.x);
          second_pass(full_c, u0 + 1, a0.y);
          second_pass(full_c, u0 + 2, a0.z);
          second_pass(full_c, u0 + 3, a0.w);
          second_pass(full_c, u0 + 4, a1.x);
          second_pass(full_c, u0 + 5, a1.y);
          second_pass(full_c, u0 + 6, a1.z);
          second_pass(full_c, u0 + 7, a1.w);
        }
      } else {
        for (uint32_t u = tid; u < U; u += kT3Threads) second_pass(full_c, u, __ldcg(pairs + u));
      }
    };
    // DENSE: the table is free again after the rows were read, so pass 2 simply counts the users of every (p, c) in
    // it -- one increment per user, no per-tile lookup, no branch on the pair -- and the counts the closed form needs
    // (cnt_cf, m_p, cnt_l) are read off row p afterwards.  "early" (a non-first user of (p, c_f) before x_p) keeps its
    // exact test, but behind one byte load: s_ub[p] bounds the users that can still set it, and drops to 0 once it is set.
    // the increment; returns the bound byte of the row (0 for a missing user)
    auto count_pair = [&](auto full_c, uint32_t pc) -> uint32_t {
      if (!decltype(full_c)::value && pc == kNoPair) return 0u;
      const uint32_t p = pc & 0xFFFFu, c = pc >> 16;
      atomicAdd(s_tab + p * t3_row_stride(T) + c, 1u);
      return (uint32_t)s_ub[p];
    };
    auto early_test = [&](uint32_t u, uint32_t pc) {   // rare: u may still be the user that sets "early" for its row
      const uint32_t p = pc & 0xFFFFu, c = pc >> 16;
      const uint32_t w = lds_u32(&s_cfl[p]);
      if (c == (w & kNoTile) && !(w & kEarlyBit)) {
        const uint32_t f = reinterpret_cast<const uint32_t*>(s_rmin + p)[1];
        const uint32_t x = reinterpret_cast<const uint32_t*>(s_rmax + p)[1] - 1u;
        if (u != f && u < x) {
          atomicOr(&s_cfl[p], kEarlyBit);
          s_ub[p] = (uint8_t)0;
        }
      }
    };
    auto count_user = [&](auto full_c, uint32_t ushifted, uint32_t u, uint32_t pc) {
      if (ushifted < count_pair(full_c, pc)) early_test(u, pc);
    };
    // 8 users: all increments and bound-byte loads first, then the (rare) early tests
    auto count_users8 = [&](auto full_c, uint32_t us, uint32_t u0, const uint32_t (&w)[8]) {
      uint32_t ub[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) ub[j] = count_pair(full_c, w[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (us < ub[j]) early_test(u0 + j, w[j]);
    };
    auto pass2_dense = [&](auto full_c) {
      constexpr bool FULL = decltype(full_c)::value;
      const uint32_t ush = (uint32_t)a.ush;
      if (!scratch) {
        auto from_rows = [&](uint32_t cp, uint32_t cc) { return (FULL || (cp != 0xFFFFu && cc != 0xFFFFu)) ? (cp | (cc << 16)) : kNoPair; };
        if (vec) {
          uint4 np = make_uint4(0u, 0u, 0u, 0u), nc = np;
          if (tid * 8u < U) {
            np = __ldg(reinterpret_cast<const uint4*>(prow + tid * 8u));
            nc = __ldg(reinterpret_cast<const uint4*>(crow + tid * 8u));
          }
          for (uint32_t u0 = tid * 8u; u0 < U; u0 += kT3Threads * 8u) {
            const uint4 vp = np, vc = nc;
            if (u0 + kT3Threads * 8u < U) {
              np = __ldg(reinterpret_cast<const uint4*>(prow + u0 + kT3Threads * 8u));
              nc = __ldg(reinterpret_cast<const uint4*>(crow + u0 + kT3Threads * 8u));
            }
            const uint32_t wp[4] = {vp.x, vp.y, vp.z, vp.w}, wc[4] = {vc.x, vc.y, vc.z, vc.w};
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              w[j] = from_rows((wp[j >> 1] >> (16 * (j & 1))) & 0xFFFFu, (wc[j >> 1] >> (16 * (j & 1))) & 0xFFFFu);
            count_users8(full_c, u0 >> ush, u0, w);  // ush >= 3 and u0 % 8 == 0: one shifted index for the 8 users
          }
        } else {
          for (uint32_t u = tid; u < U; u += kT3Threads) count_user(full_c, u >> ush, u, from_rows(prow[u], crow[u]));
        }
      } else if (vec) {
        uint4 n0 = make_uint4(0u, 0u, 0u, 0u), n1 = n0;
        if (tid * 8u < U) {
          n0 = __ldcg(reinterpret_cast<const uint4*>(pairs + tid * 8u));
          n1 = __ldcg(reinterpret_cast<const uint4*>(pairs + tid * 8u + 4));
        }
        for (uint32_t u0 = tid * 8u; u0 < U; u0 += kT3Threads * 8u) {
          const uint4 a0 = n0, a1 = n1;
          if (u0 + kT3Threads * 8u < U) {
            n0 = __ldcg(reinterpret_cast<const uint4*>(pairs + u0 + kT3Threads * 8u));
            n1 = __ldcg(reinterpret_cast<const uint4*>(pairs + u0 + kT3Threads * 8u + 4));
          }
          const uint32_t w[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          count_users8(full_c, u0 >> ush, u0, w);
        }
      } else {
        for (uint32_t u = tid; u < U; u += kT3Threads) count_user(full_c, u >> ush, u, __ldcg(pairs + u));
      }
    };
    if (MODE == kT3Dense) {
      if (full) pass2_dense(std::true_type{});
      else pass2_dense(std::false_type{});
      __syncthreads();
      // counts of row p: m_p = sum, cnt_cf = N[p][c_f], cnt_l = N[p][l'_p]; the table is left empty for the next pair
      for (uint32_t p = wid; p < T; p += kT3Threads / 32) {
        if (s_rmin[p] == ~0ull) continue;  // warp-uniform
        uint32_t* row = s_tab + p * t3_row_stride(T);
        const uint32_t w = s_cfl[p];
        const uint32_t cf = w & kNoTile, cl = (w >> 14) & kNoTile;
        uint32_t sum = 0, ncf = 0, ncl = 0;
        for (uint32_t c = lane; c < T; c += 32) {
          const uint32_t v = row[c];
          if (v != kEmpty) {
            sum += v;
            if (c == cf) ncf = v;
            if (c == cl) ncl = v;
            row[c] = kEmpty;
          }
        }
        sum = __reduce_add_sync(kFull, sum);
        ncf = __reduce_add_sync(kFull, ncf);
        ncl = __reduce_add_sync(kFull, ncl);
        if (lane == 0) {
          s_cf[p] = ncf;
          s_other[p] = sum - ncf;
          s_cl[p] = ncl;
        }
      }
    } else {
      if (full) pass2(std::true_type{});
      else pass2(std::false_type{});
    }
    __syncthreads();

    // ---- EU:297-330 ----
    double acc = 0.0;
    for (uint32_t p = tid; p < T; p += kT3Threads) {
      const uint32_t ncf = s_cf[p];
      const uint32_t m = ncf + s_other[p];
      if (m == 0u) continue;
      const bool has2 = ncf >= 2u;
      const double Kp = 1.0 + (double)(s_d[p] - 1u + (has2 ? 1u : 0u));
      double wp = 1.0;
      if (m > 1u) {
        const unsigned long long top = s_rmax[p];
        const bool cf_latest = has2 && (top == 0ull || !(s_cfl[p] & kEarlyBit));
        wp = cf_latest ? (double)(ncf - 1u) : (double)s_cl[p];
      }
      const double tp = wp / (double)m;
      acc += -((double)m / total) * (Kp * (tp * log2(tp)));
    }
    const double Hs = block_sum(acc, s_red);
    const double n = (total > (double)T) ? (double)T : total;
    const double q = 1.0 / n;
    double e = Hs / (n * -q * log2(q));
    if (total == 0.0) {
      e = qnan;
      if (tid == 0) atomicOr(a.flags, (uint32_t)VET_FLAG_NO_COMMON_USER);
    }
    if (tid == 0) {
      a.out[r] = e;
      s_valid = 0u;
      s_used = 0u;
    }
    if (a.prev_count0)
      for (uint32_t t = tid; t < T; t += kT3Threads) a.prev_count0[r * (int64_t)T + t] = (int32_t)(s_cf[t] + s_other[t]);
    __syncthreads();
  }
}

// entropy[r] = (sum_k per_k[k][r]) / K in tile-count order (TA:160), for the per-tile-count launches
__global__ void k_mean_rows(const double* __restrict__ per_k, int64_t stride, int K, int64_t R, double* __restrict__ entropy) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= R) return;
  double s = 0.0;
  for (int k = 0; k < K; ++k) s += per_k[k * stride + r];
  entropy[r] = s / (double)K;
}

}  // namespace vet
