// Stage 4, two-pass kernel for the TAIL of a launch: one frame pair per thread-block CLUSTER.
//
// k_transition3 gives a frame pair to one CTA, so `rows` pairs on B SMs cost ceil(rows / B) rounds: 449 pairs
// of 1M users on 148 SMs run 4 rounds for 3.03 rounds of work.  Everything the closed form of
// vet_transition3.cuh takes from the users is a minimum of user indices (A[p][c]), a count (cnt_cf, other,
// cnt_l) or an OR (early) -- all of them merge across slices of the users.  Here the S CTAs of a cluster
// (S = 2, 4 or 8) split the users of ONE pair in interleaved chunks:
//   pass 1   every CTA fills its own dense T x T table A_j from its chunks;
//   rows     row p of A = min_j A_j[p][.] is read through distributed shared memory by CTA (p mod S), which
//            broadcasts D_p, (f_p, c_f), (x_p, l'_p) into the per-tile arrays of all S CTAs and leaves the
//            tables empty;
//   pass 2   every CTA counts its own users against its copy of those arrays;
//   entropy  CTA 0 adds the S partial counts (OR of the early bits) and finishes like k_transition3 -- same
//            integers, same summation order, bit-identical output.
// The host launches it for the rows % B pairs left over after the full rounds of k_transition3 (dense tables
// only: T <= ~220), with the largest S whose clusters are all co-resident.
#pragma once
#include <cooperative_groups.h>

#include "vet_transition3.cuh"

namespace vet {

namespace cg = cooperative_groups;

constexpr int kT3cCols = 8;  // merged entries of a row per lane: T <= 256 (the dense table stops near 220)

template <int LW>
__global__ void __launch_bounds__(kT3Threads, 1) k_transition3c(Transition3Args a) {
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t S = cluster.num_blocks(), rank = cluster.block_rank();
  const uint32_t cid = blockIdx.x / S, nclusters = gridDim.x / S;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t T = (uint32_t)a.T;
  unsigned long long* s_rmin = reinterpret_cast<unsigned long long*>(smem_raw);  // as in k_transition3
  unsigned long long* s_rmax = s_rmin + T;
  uint32_t* s_other = reinterpret_cast<uint32_t*>(s_rmax + T);
  uint32_t* s_d = s_other + T;
  uint32_t* s_cfl = s_d + T;
  uint32_t* s_cf = s_cfl + T;
  uint32_t* s_cl = s_cf + T;
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem_raw + a.tab_off);
  __shared__ double s_red[32];
  __shared__ uint32_t s_valid;
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  const uint32_t pitch = t3_row_stride(T);

  for (uint32_t i = tid; i < T * pitch; i += kT3Threads) s_tab[i] = kEmpty;
  Lut3<LW> lut{a.lut_src};
  if (LW == kLutS8 || LW == kLutS16) {
    const int bytes = a.C * (LW == kLutS8 ? 1 : 2);
    const uint4* __restrict__ src = static_cast<const uint4*>(a.lut_src);
    uint4* dst = reinterpret_cast<uint4*>(smem_raw + a.lut_off);
    for (int i = tid; i < (bytes + 15) / 16; i += kT3Threads) dst[i] = __ldg(src + i);
    lut.p = smem_raw + a.lut_off;
  }
  if (tid == 0) s_valid = 0u;
  // identity LUT without a scratch: pass 2 rebuilds the pairs from the two tile-id rows (see k_transition3)
  const bool scratch = !(LW == kLutIdentity && a.pair_scratch == nullptr);
  uint32_t* __restrict__ pairs = scratch ? a.pair_scratch + (size_t)cid * a.U : nullptr;  // one row per cluster, chunks owned by their CTA
  const uint32_t U = a.U;
  const bool vec = (U & 7u) == 0u;
  const uint32_t chunk = vec ? kT3Threads * 8u : kT3Threads;  // users per CTA step; chunk i belongs to CTA i % S

  for (int64_t r = cid; r < a.F - 1; r += nclusters) {  // same trip count in every CTA of the cluster
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
    }
    __syncthreads();

    // ---- pass 1 over this CTA's chunks ----
    uint32_t nvalid = 0;
    // both frames complete (no missing user): the 0xFFFF tests of both passes drop out (block- and cluster-uniform)
    const bool full = a.nvalid && __ldg(a.nvalid + r) == U && __ldg(a.nvalid + r + 1) == U;
    auto pass1 = [&](auto full_c) {
      constexpr bool FULL = decltype(full_c)::value;
      uint4 nvp = make_uint4(0u, 0u, 0u, 0u), nvc = nvp;  // loads of the next step, issued before the updates of this one
      if (vec && (uint64_t)rank * chunk + tid * 8u < U) {
        nvp = __ldg(reinterpret_cast<const uint4*>(prow + rank * chunk + tid * 8u));
        nvc = __ldg(reinterpret_cast<const uint4*>(crow + rank * chunk + tid * 8u));
      }
      for (uint64_t base = (uint64_t)rank * chunk; base < U; base += (uint64_t)S * chunk) {
        if (vec) {
          const uint32_t u0 = (uint32_t)base + tid * 8u;
          const uint4 vp = nvp, vc = nvc;
          if ((uint64_t)u0 + (uint64_t)S * chunk < U) {
            nvp = __ldg(reinterpret_cast<const uint4*>(prow + u0 + S * chunk));
            nvc = __ldg(reinterpret_cast<const uint4*>(crow + u0 + S * chunk));
          }
          if (u0 < U) {
            const uint32_t wp[4] = {vp.x, vp.y, vp.z, vp.w}, wc[4] = {vc.x, vc.y, vc.z, vc.w};
            uint32_t pc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t cp = (wp[j >> 1] >> (16 * (j & 1))) & 0xFFFFu, cc = (wc[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
              const bool ok = FULL || (cp != 0xFFFFu && cc != 0xFFFFu);
              pc[j] = ok ? (lut(ok ? cp : 0u) | (lut(ok ? cc : 0u) << 16)) : kNoPair;
            }
            if (scratch) {
              *reinterpret_cast<uint4*>(pairs + u0) = make_uint4(pc[0], pc[1], pc[2], pc[3]);
              *reinterpret_cast<uint4*>(pairs + u0 + 4) = make_uint4(pc[4], pc[5], pc[6], pc[7]);
            }
            if (p0row) {
              *reinterpret_cast<uint4*>(p0row + u0) = make_uint4(pc[0], pc[1], pc[2], pc[3]);
              *reinterpret_cast<uint4*>(p0row + u0 + 4) = make_uint4(pc[4], pc[5], pc[6], pc[7]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (FULL || pc[j] != kNoPair) {
                ++nvalid;
                t3_update<kT3Dense>(s_tab, nullptr, T, pc[j] & 0xFFFFu, pc[j] >> 16, u0 + j, nullptr, nullptr);
              }
          }
        } else {
          const uint32_t u = (uint32_t)base + tid;
          const uint32_t cp = u < U ? prow[u] : 0xFFFFu, cc = u < U ? crow[u] : 0xFFFFu;
          const bool ok = cp != 0xFFFFu && cc != 0xFFFFu;
          const uint32_t pc = ok ? (lut(ok ? cp : 0u) | (lut(ok ? cc : 0u) << 16)) : kNoPair;
          if (u < U) {
            if (scratch) pairs[u] = pc;
            if (p0row) p0row[u] = pc;
          }
          if (ok) {
            ++nvalid;
            t3_update<kT3Dense>(s_tab, nullptr, T, pc & 0xFFFFu, pc >> 16, u, nullptr, nullptr);
          }
        }
        __syncwarp();
      }
    };
    if (full) pass1(std::true_type{});
    else pass1(std::false_type{});
    nvalid = __reduce_add_sync(kFull, nvalid);
    if (lane == 0 && nvalid) atomicAdd(&s_valid, nvalid);
    cluster.sync();  // every table of the cluster is complete

    // ---- rows p = rank (mod S) of the merged table; results go to all S CTAs, the tables are left empty ----
    for (uint32_t p = rank + S * wid; p < T; p += S * (kT3Threads / 32)) {
      uint32_t v[kT3cCols];
      uint32_t cnt = 0;
      unsigned long long best = ~0ull;
#pragma unroll
      for (int i = 0; i < kT3cCols; ++i) {
        const uint32_t c = lane + 32u * i;
        uint32_t m = kEmpty;
        if (c < T)
          for (uint32_t j = 0; j < S; ++j) {
            uint32_t* e = cluster.map_shared_rank(s_tab, j) + p * pitch + c;
            const uint32_t x = *e;
            if (x != kEmpty) {
              m = min(m, x);
              *e = kEmpty;
            }
          }
        v[i] = m;
        if (m != kEmpty) {
          ++cnt;
          best = min(best, ((unsigned long long)m << 32) | c);
        }
      }
      cnt = __reduce_add_sync(kFull, cnt);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(kFull, best, o));
      if (best == ~0ull) continue;  // no user came from tile p (warp-uniform)
      const uint32_t cf = (uint32_t)best;
      unsigned long long top = 0ull;
#pragma unroll
      for (int i = 0; i < kT3cCols; ++i) {
        const uint32_t c = lane + 32u * i;
        if (v[i] != kEmpty && c != cf) top = max(top, ((unsigned long long)(v[i] + 1u) << 32) | c);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) top = max(top, __shfl_xor_sync(kFull, top, o));
      if (lane < S) {
        cluster.map_shared_rank(s_d, lane)[p] = cnt;
        cluster.map_shared_rank(s_rmin, lane)[p] = best;
        cluster.map_shared_rank(s_rmax, lane)[p] = top;
        cluster.map_shared_rank(s_cfl, lane)[p] = cf | ((top ? (uint32_t)top & kNoTile : kNoTile) << 14);
      }
    }
    cluster.sync();

    // ---- pass 2 over this CTA's chunks (see k_transition3) ----
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
      // without the scratch the two row loads stand in for the two halves of the 8 packed pairs
      const uint4* __restrict__ src0 = scratch ? nullptr : reinterpret_cast<const uint4*>(prow);
      const uint4* __restrict__ src1 = scratch ? nullptr : reinterpret_cast<const uint4*>(crow);
      auto load0 = [&](uint32_t u0) { return scratch ? __ldcg(reinterpret_cast<const uint4*>(pairs + u0)) : __ldg(src0 + (u0 >> 3)); };
      auto load1 = [&](uint32_t u0) { return scratch ? __ldcg(reinterpret_cast<const uint4*>(pairs + u0 + 4)) : __ldg(src1 + (u0 >> 3)); };
      auto from_rows = [&](uint32_t cp, uint32_t cc) { return (FULL || (cp != 0xFFFFu && cc != 0xFFFFu)) ? (cp | (cc << 16)) : kNoPair; };
      uint4 n0 = make_uint4(0u, 0u, 0u, 0u), n1 = n0;
      if (vec && (uint64_t)rank * chunk + tid * 8u < U) {
        n0 = load0(rank * chunk + tid * 8u);
        n1 = load1(rank * chunk + tid * 8u);
      }
      for (uint64_t base = (uint64_t)rank * chunk; base < U; base += (uint64_t)S * chunk) {
        if (vec) {
          const uint32_t u0 = (uint32_t)base + tid * 8u;
          const uint4 a0 = n0, a1 = n1;
          if ((uint64_t)u0 + (uint64_t)S * chunk < U) {
            n0 = load0(u0 + S * chunk);
            n1 = load1(u0 + S * chunk);
          }
          if (u0 < U) {
            if (scratch) {
              const uint32_t w[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
              for (int q = 0; q < 8; ++q) second_pass(full_c, u0 + q, w[q]);
            } else {
              const uint32_t wp[4] = {a0.x, a0.y, a0.z, a0.w}, wc[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
              for (int q = 0; q < 8; ++q)
                second_pass(full_c, u0 + q, from_rows((wp[q >> 1] >> (16 * (q & 1))) & 0xFFFFu, (wc[q >> 1] >> (16 * (q & 1))) & 0xFFFFu));
            }
          }
        } else {
          const uint32_t u = (uint32_t)base + tid;
          if (u < U) second_pass(full_c, u, scratch ? __ldcg(pairs + u) : from_rows(prow[u], crow[u]));
        }
      }
    };
    if (full) pass2(std::true_type{});
    else pass2(std::false_type{});
    cluster.sync();  // every CTA's partial counts are complete

    // ---- EU:297-330, CTA 0 over the partial counts of the cluster ----
    if (rank == 0) {
      uint32_t tot = 0;
      for (uint32_t j = 0; j < S; ++j) tot += *cluster.map_shared_rank(&s_valid, j);
      const double total = (double)tot;
      double acc = 0.0;
      for (uint32_t p = tid; p < T; p += kT3Threads) {
        uint32_t ncf = 0, oth = 0, ncl = 0, early = 0;
        for (uint32_t j = 0; j < S; ++j) {
          ncf += cluster.map_shared_rank(s_cf, j)[p];
          oth += cluster.map_shared_rank(s_other, j)[p];
          ncl += cluster.map_shared_rank(s_cl, j)[p];
          early |= cluster.map_shared_rank(s_cfl, j)[p] & kEarlyBit;
        }
        const uint32_t m = ncf + oth;
        if (a.prev_count0) a.prev_count0[r * (int64_t)T + p] = (int32_t)m;
        if (m == 0u) continue;
        const bool has2 = ncf >= 2u;
        const double Kp = 1.0 + (double)(s_d[p] - 1u + (has2 ? 1u : 0u));
        double wp = 1.0;
        if (m > 1u) {
          const unsigned long long top = s_rmax[p];
          const bool cf_latest = has2 && (top == 0ull || !early);
          wp = cf_latest ? (double)(ncf - 1u) : (double)ncl;
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
      if (tid == 0) a.out[r] = e;
    }
    cluster.sync();  // CTA 0 has read everything: the per-tile arrays may be reset
    if (tid == 0) s_valid = 0u;
  }
}

}  // namespace vet
