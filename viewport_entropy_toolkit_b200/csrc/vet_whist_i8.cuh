// FOV-weighted tile histograms on the 5th-generation tensor cores (north_star stage 3, weighted).
//
//   HIST[f, t] = sum_cell CNT[f, cell] * w(cell, t)            (EU:130-138, 190-192)
//
// CNT holds INTEGERS (users per cell and frame), so the contraction can be done exactly in
// integer arithmetic (the "Ozaki" splitting used to emulate fp64 GEMMs on int8 units):
//
//   * a weight is quantised once, at table-build time, to 39 fractional bits,
//     q = rint(w * 2^39) in [0, 2^39] (w = 1.0 occurs: a cell that coincides with a tile
//     centre), and cut into five 8-bit slices q = sum_s q_s 2^(8s); a count is cut into three
//     8-bit planes c = c_0 + 2^8 c_1 + 2^16 c_2 (planes 1 and 2 are zero for almost every frame
//     and are skipped per frame block through device-side flags; plane 2 -- 65536 users in one
//     cell -- takes a second launch that adds to the result);
//   * D_p,s[f, t] = sum_cell c_p[f, cell] * q_s[cell, t] is an unsigned 8-bit GEMM with int32
//     accumulation -- exact: sum_cell c_p <= U, so D <= 255 U < 2^31 (U < 8.4 M, checked on the host);
//   * HIST = 2^-39 * sum_s 2^(8s) (D_0,s + 2^8 D_1,s + 2^16 D_2,s), evaluated in fp64 from the
//     exact integers in a fixed order.  The only deviation from the FP64 kernel is the weight
//     quantisation: |dw| <= 2^-40 = 9.1e-13 per (cell, tile), i.e. <= 1e-11 relative on a
//     histogram entry whose mean weight is >= 0.1 (tolerance of the path: 1e-9, tests/).
//
// GEMM shape: M = frames (128 per CTA = the 128 TMEM lanes), N = 240 = 5 slices x 48 tiles
// (column s*48 + j of N block nb is slice s of tile nb*48 + j), K = cells.  Tiles are in
// lattice order = sorted by latitude (DU:46), cells are row-major by latitude, so the support
// of an N block is one contiguous run of 128-cell K blocks: everything outside it is skipped.
//
// One CTA per (frame block, N block), 192 threads, warp-specialised:
//   warp 0, one lane   TMA producer: cp.async.bulk.tensor (128-byte swizzle) of the count planes
//                      [128 x 128 B each] and the weight slices [240 x 128 B] into a ring of 4 stages
//                      (3 when the frame block needs the second count plane);
//   warp 1, one lane   tcgen05.mma.kind::i8 (M128 N240 K32, u8 x u8 -> s32) into TMEM: accumulator
//                      of plane 0 in columns [0,240), of plane 1 in [256,496); tcgen05.commit
//                      releases the stage / signals the epilogue;
//   warps 2..5         epilogue: tcgen05.ld (lane = frame), recombine the slices in fp64, store the
//                      histogram row segment.
#pragma once
#include <cuda.h>

#include "vet_common.cuh"
#include "vet_stream_tma.cuh"

namespace vet {

constexpr int kI8Slices = 5;
constexpr int kI8FracBits = 39;
constexpr int kI8TilesPerBlock = 48;
constexpr int kI8N = kI8Slices * kI8TilesPerBlock;  // 240
constexpr int kI8M = 128;
constexpr int kI8BK = 128;  // cells (= bytes) per pipeline stage: one swizzle row
constexpr int kI8Stages = 3;      // with both count planes (62 KB per stage)
constexpr int kI8StagesOne = 4;   // with plane 0 only (46 KB per stage): the usual case
constexpr int kI8ABytes = kI8M * kI8BK;                    // 16 KB per count plane
constexpr int kI8BBytes = kI8N * kI8BK;                    // 30 KB
constexpr int kI8StageBytes = 2 * kI8ABytes + kI8BBytes;   // 62 KB, multiple of 1024
constexpr int kI8Threads = 192;
constexpr int kI8MaxSplit = 8;          // CTAs that may share one output tile (split over the cells)
constexpr int kI8TmemCols = 512;
constexpr int kI8SmemBytes = kI8Stages * kI8StageBytes + 1024;  // + slack to align the ring to 1024 B
static_assert(kI8StageBytes % 1024 == 0 && (kI8ABytes + kI8BBytes) % 1024 == 0, "stages must keep the 1024 B swizzle alignment");
static_assert(kI8StagesOne * (kI8ABytes + kI8BBytes) <= kI8Stages * kI8StageBytes, "both layouts share one ring");

// ---- PTX wrappers (layouts: PTX ISA "tcgen05" / CUTLASS cute/arch/mma_sm100_desc.hpp) ----
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// arrives on the mbarrier once every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, unsigned 8-bit operands, int32 accumulators
__device__ __forceinline__ void tc_mma_i8(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp gets lane (taddr.lane + i)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major operand tile stored as 128-byte rows with the
// 128-byte swizzle (what the TMA box {128 B, rows} writes): 8-row groups are 1024 B apart (SBO),
// LBO is unused for swizzled K-major layouts (1), descriptor version 1 (sm_100), layout type 2.
__device__ __forceinline__ uint64_t i8_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor: dense, no saturation, D = s32 (2 @ bit 4), A = B = unsigned 8-bit (0 @ bits 7, 10),
// both K-major (0 @ bits 15, 16), N >> 3 @ bit 17, M >> 4 @ bit 24.
constexpr uint32_t kI8InstrDesc = (2u << 4) | ((uint32_t)(kI8N >> 3) << 17) | ((uint32_t)(kI8M >> 4) << 24);

struct WhistI8Args {
  int64_t F;
  int T;
  int n_blocks;               // N blocks of 48 tiles
  int row_a;                  // first row, in the plane tensor, of the plane accumulated in TMEM columns [0,240)
  int row_b;                  // same for the plane accumulated in columns [256,496), used where flag_b is set
  const uint32_t* flag_b;     // [frame blocks] or null: the frame block needs the second plane
  const uint32_t* run_if;     // [frame blocks] or null: CTAs of a frame block whose flag is zero do nothing
  int shift;                  // the planes hold count bits [shift, shift+16)
  int accumulate;             // add to hist instead of storing (second pass, planes of higher bits)
  const int2* kb_range;       // [n_blocks] first / past-the-end 128-cell K block with a non-zero weight
  double* hist;               // [F, T]
  // split-K (few frame blocks: 20 output tiles at 450 frames x 201 tiles, each walking every cell for ~70 us): ksplit
  // CTAs share one output tile, each stores its int32 accumulators in its own slice of `part` and k_whist_i8_finish
  // adds the slices (integers: exact whatever the split) and turns the sums into the histogram rows
  int ksplit;                 // 1 = one CTA per output tile, rows straight from TMEM
  int* part;                  // [frame blocks * n_blocks][ksplit][2 planes][240 columns][128 frames]
};

__global__ void __launch_bounds__(kI8Threads, 1)
k_whist_i8(const __grid_constant__ CUtensorMap tm_cnt, const __grid_constant__ CUtensorMap tm_w, WhistI8Args a) {
  const int tile_id = blockIdx.x / a.ksplit;  // the CTAs of one output tile are neighbours: they read the same planes
  const int sp = blockIdx.x % a.ksplit;
  const int nb = tile_id % a.n_blocks;
  const int mb = tile_id / a.n_blocks;
  if (a.run_if && a.run_if[mb] == 0u) return;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) unsigned long long s_full[kI8StagesOne], s_empty[kI8StagesOne], s_accum;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  int2 kr = a.kb_range[nb];
  if (a.ksplit > 1) {  // this CTA's share of the K blocks (possibly none)
    const int per = (kr.y - kr.x + a.ksplit - 1) / a.ksplit;
    kr.x = min(kr.y, kr.x + sp * per);
    kr.y = min(kr.y, kr.x + per);
  }
  const bool two = a.flag_b && a.flag_b[mb] != 0u;

  // the ring holds 3 stages of {plane 0, plane 1, weights} or 4 stages of {plane 0, weights}
  const uint32_t nstages = two ? kI8Stages : kI8StagesOne;
  const uint32_t stage_bytes = two ? (uint32_t)kI8StageBytes : (uint32_t)(kI8ABytes + kI8BBytes);
  const uint32_t w_off = two ? 2u * kI8ABytes : (uint32_t)kI8ABytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kI8StagesOne; ++i) {
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&s_empty[i]), 1);
    }
    mbar_init(smem_u32(&s_accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tc_alloc(smem_u32(&s_tmem), kI8TmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t stage = 0, phase = 0;
      for (int kb = kr.x; kb < kr.y; ++kb) {
        mbar_wait(smem_u32(&s_empty[stage]), phase ^ 1u);
        const uint32_t bar = smem_u32(&s_full[stage]);
        const uint32_t base = ring + stage * stage_bytes;
        mbar_expect_tx(bar, stage_bytes);
        tma_load_2d(base, &tm_cnt, bar, kb * kI8BK, a.row_a + mb * kI8M);
        if (two) tma_load_2d(base + kI8ABytes, &tm_cnt, bar, kb * kI8BK, a.row_b + mb * kI8M);
        tma_load_2d(base + w_off, &tm_w, bar, kb * kI8BK, nb * kI8N);
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      uint32_t stage = 0, phase = 0;
      for (int kb = kr.x; kb < kr.y; ++kb) {
        mbar_wait(smem_u32(&s_full[stage]), phase);
        tc_fence_after();
        const uint32_t base = ring + stage * stage_bytes;
#pragma unroll
        for (int k = 0; k < kI8BK / 32; ++k) {  // one instruction covers 32 cells
          const uint32_t acc = (kb > kr.x || k > 0) ? 1u : 0u;
          const uint64_t bdesc = i8_smem_desc(base + w_off + k * 32);
          tc_mma_i8(tmem, i8_smem_desc(base + k * 32), bdesc, kI8InstrDesc, acc);
          if (two) tc_mma_i8(tmem + 256, i8_smem_desc(base + kI8ABytes + k * 32), bdesc, kI8InstrDesc, acc);
        }
        tc_commit(smem_u32(&s_empty[stage]));  // stage is free once these MMAs have read it
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      tc_commit(smem_u32(&s_accum));
    }
    __syncwarp();
  } else {
    // ===== epilogue: warp w may read TMEM lanes [32 (w % 4), +32) =====
    const int q = warp & 3;
    const int64_t f = (int64_t)mb * kI8M + q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    if (a.ksplit > 1) {
      // column-major slice: the 32 frames of a warp are one 128-byte store
      int* __restrict__ part = a.part + ((int64_t)tile_id * a.ksplit + sp) * (2 * kI8N * kI8M) + q * 32 + lane;
      const bool have = kr.y > kr.x;  // no K block for this CTA: its slice is zero
      if (have) {
        mbar_wait(smem_u32(&s_accum), 0u);
        tc_fence_after();
      }
#pragma unroll 1
      for (int c0 = 0; c0 < kI8N; c0 += 16) {
        uint32_t r0[16], r1[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) r0[j] = r1[j] = 0u;
        if (have) {
          tc_ld16(lane_base + c0, r0);
          if (two) tc_ld16(lane_base + 256 + c0, r1);
          tc_wait_ld();
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          __stcg(part + (c0 + j) * kI8M, (int)r0[j]);
          if (two) __stcg(part + (kI8N + c0 + j) * kI8M, (int)r1[j]);
        }
      }
    } else {
      mbar_wait(smem_u32(&s_accum), 0u);
      tc_fence_after();
#pragma unroll 1
      for (int j0 = 0; j0 < kI8TilesPerBlock; j0 += 16) {
        double v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.0;
#pragma unroll
        for (int s = kI8Slices - 1; s >= 0; --s) {  // most significant slice first
          uint32_t r0[16], r1[16];
          tc_ld16(lane_base + s * kI8TilesPerBlock + j0, r0);
          if (two) tc_ld16(lane_base + 256 + s * kI8TilesPerBlock + j0, r1);
          tc_wait_ld();
          const double scale = __longlong_as_double((long long)(1023 + 8 * s - kI8FracBits + a.shift) << 52);  // 2^(8s-39+shift)
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            long long x = (long long)(int)r0[j];
            if (two) x += (long long)(int)r1[j] << 8;
            v[j] = fma((double)x, scale, v[j]);
          }
        }
        if (f < a.F) {
          double* __restrict__ row = a.hist + f * (int64_t)a.T;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int tile = nb * kI8TilesPerBlock + j0 + j;
            if (tile < a.T) row[tile] = a.accumulate ? row[tile] + v[j] : v[j];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tc_dealloc(tmem, kI8TmemCols);
}

// Second half of the split-K launch: thread (frame, tile) adds the ksplit slices of its 5 (10) accumulators and runs
// the same fma chain as the one-CTA epilogue above, so the rows carry the same bits however the cells were split.
// Grid: (48 tile columns, output tiles), 128 threads = the frames of the block.
__global__ void __launch_bounds__(kI8M) k_whist_i8_finish(WhistI8Args a) {
  const int tile_id = blockIdx.y;
  const int nb = tile_id % a.n_blocks;
  const int mb = tile_id / a.n_blocks;
  if (a.run_if && a.run_if[mb] == 0u) return;
  const int j = blockIdx.x;
  const int tile = nb * kI8TilesPerBlock + j;
  const int64_t f = (int64_t)mb * kI8M + threadIdx.x;
  if (tile >= a.T || f >= a.F) return;
  const bool two = a.flag_b && a.flag_b[mb] != 0u;
  const int* __restrict__ part = a.part + (int64_t)tile_id * a.ksplit * (2 * kI8N * kI8M) + threadIdx.x;
  // every load of the thread in flight at once: 5 slices x up to kI8MaxSplit parts (x 2 planes)
  int x0[kI8Slices], x1[kI8Slices];
#pragma unroll
  for (int s = 0; s < kI8Slices; ++s) x0[s] = x1[s] = 0;
#pragma unroll
  for (int sp = 0; sp < kI8MaxSplit; ++sp) {
    if (sp < a.ksplit) {
#pragma unroll
      for (int s = 0; s < kI8Slices; ++s) {
        const int* p = part + (int64_t)sp * (2 * kI8N * kI8M) + (s * kI8TilesPerBlock + j) * kI8M;
        x0[s] += __ldcg(p);
        if (two) x1[s] += __ldcg(p + kI8N * kI8M);
      }
    }
  }
  double v = 0.0;
#pragma unroll
  for (int s = kI8Slices - 1; s >= 0; --s) {
    long long x = (long long)x0[s];
    if (two) x += (long long)x1[s] << 8;
    const double scale = __longlong_as_double((long long)(1023 + 8 * s - kI8FracBits + a.shift) << 52);
    v = fma((double)x, scale, v);
  }
  double* __restrict__ row = a.hist + f * (int64_t)a.T;
  row[tile] = a.accumulate ? row[tile] + v : v;
}

// Cell histogram rows (uint32) -> the byte planes of the tensor-core kernel, for frames that the
// streaming kernel accumulated in several chunks (it then adds into `cnt` with RED and cannot emit the
// planes itself).  All three planes of every row are written; the rows are marked dirty.
struct CntPlanesArgs {
  const uint32_t* cnt;  // [F, cpad]
  int64_t F;
  int cpad;
  int kp;               // cells per plane row, multiple of 128
  int64_t plane_stride; // bytes between planes
  uint8_t* planes;
  uint8_t* dirty;       // [rows]
  uint32_t* hi1;        // [frame blocks], zeroed by the caller
  uint32_t* hi2;
};

__global__ void __launch_bounds__(256) k_cnt_planes(CntPlanesArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int upr = a.kp >> 7;  // 128-cell units per row
  const int64_t units = a.F * upr;
  uint32_t* __restrict__ p0 = reinterpret_cast<uint32_t*>(a.planes);
  uint32_t* __restrict__ p1 = reinterpret_cast<uint32_t*>(a.planes + a.plane_stride);
  uint32_t* __restrict__ p2 = reinterpret_cast<uint32_t*>(a.planes + 2 * a.plane_stride);
  constexpr int kUnroll = 4;
  for (int64_t u0 = warp * kUnroll; u0 < units; u0 += nwarps * kUnroll) {
    uint4 v[kUnroll];
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      const int64_t u = u0 + i;
      v[i] = make_uint4(0u, 0u, 0u, 0u);
      if (u < units) {
        const int64_t f = u / upr;
        const int c = (int)(u % upr) * 128 + lane * 4;
        if (c < a.cpad) v[i] = __ldg(reinterpret_cast<const uint4*>(a.cnt + f * (int64_t)a.cpad + c));
      }
    }
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      const int64_t u = u0 + i;
      if (u < units) {
        const int64_t f = u / upr;
        const int64_t o = (f * (int64_t)a.kp + (u % upr) * 128) / 4 + lane;
        const uint32_t w1 = pack_bytes(v[i], 8), w2 = pack_bytes(v[i], 16);
        p0[o] = pack_bytes(v[i], 0);
        p1[o] = w1;
        p2[o] = w2;
        // a few flags for millions of words: look before the atomic (a stale zero only costs one more atomic)
        if (w1 && *reinterpret_cast<volatile uint32_t*>(&a.hi1[f >> 7]) == 0u) atomicOr(&a.hi1[f >> 7], 1u);
        if (w2 && *reinterpret_cast<volatile uint32_t*>(&a.hi2[f >> 7]) == 0u) atomicOr(&a.hi2[f >> 7], 1u);
        if (u % upr == 0 && lane == 0) a.dirty[f] = 3;
      }
    }
  }
}

}  // namespace vet
