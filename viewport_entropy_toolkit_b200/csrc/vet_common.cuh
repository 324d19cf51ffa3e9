// Shared device helpers for the viewport-entropy kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vet_b200.h"

namespace vet {

constexpr int kWarp = 32;
constexpr uint32_t kFull = 0xffffffffu;

// ---- decode of one coordinate -------------------------------------------------
// Reference: normalize_to_pixel, DU:256-261: trunc(float64(v) * dim), values outside
// [0,1] rejected.  For float32 input the fp64 product is exact (24-bit x <=24-bit
// mantissas), so trunc(v*dim) == floor of the exact product; a round-down fp32
// multiply can only land in [floor(P), P] because floor(P) <= 2^24 is representable,
// hence floor(fl_rd(P)) == floor(P) without touching the fp64 pipe.
__device__ __forceinline__ int pixel_of(float v, float dimf, int /*dim*/) {
  return __float2int_rz(__fmul_rd(v, dimf));
}
__device__ __forceinline__ int pixel_of(double v, float /*dimf*/, int dim) {
  return __double2int_rz(__dmul_rn(v, (double)dim));
}

template <typename T>
struct Sample {
  T mu, mv;
};

// status of a decoded sample
enum : int { kOk = 0, kMissing = 1, kOutOfRange = 2 };

template <typename T>
__device__ __forceinline__ int decode_cell(T mu, T mv, float Wf, float Hf, int W, int H, int& cell) {
  if (mu != mu || mv != mv) {  // NaN: dropna at DU:314 / None at SA:131-135
    cell = -1;
    return kMissing;
  }
  if (mu < (T)0 || mu > (T)1 || mv < (T)0 || mv > (T)1) {  // DU:256-257
    cell = -1;
    return kOutOfRange;
  }
  const int px = pixel_of(mu, Wf, W);
  const int py = pixel_of(mv, Hf, H);
  cell = py * (W + 1) + px;
  return kOk;
}

// ---- Vector.from_spherical from the per-axis tables ----------------------------
// DT:208-216: x = sin(phi)cos(theta), y = sin(phi)sin(theta), z = cos(phi), each
// rounded with numpy's round(., 6) == rint(v * 1e6) / 1e6.  The trig factors are
// host-made (numpy or libm); the products, rint and the IEEE division are
// bit-identical on the device.  No FMA contraction: every op is explicit.
__device__ __forceinline__ double round6(double v) {
  return __ddiv_rn(rint(__dmul_rn(v, 1e6)), 1e6);
}
__device__ __forceinline__ void cell_vector(const double* __restrict__ cosT, const double* __restrict__ sinT,
                                            const double* __restrict__ sinP, const double* __restrict__ cosP,
                                            int px, int py, double& x, double& y, double& z) {
  const double sp = sinP[py];
  x = round6(__dmul_rn(sp, cosT[px]));
  y = round6(__dmul_rn(sp, sinT[px]));
  z = round6(cosP[py]);
}

// ---- vector_angle_distance pieces (EU:54-64) ------------------------------------
// np.linalg.norm / np.dot on 3 elements evaluate as acc=x0*y0; acc=fma(x1,y1,acc);
// acc=fma(x2,y2,acc) on the reference host (OpenBLAS ddot).
__device__ __forceinline__ double dot3(double ax, double ay, double az, double bx, double by, double bz) {
  double acc = __dmul_rn(ax, bx);
  acc = __fma_rn(ay, by, acc);
  acc = __fma_rn(az, bz, acc);
  return acc;
}
__device__ __forceinline__ void normalize3(double& x, double& y, double& z) {
  const double n = sqrt(dot3(x, y, z, x, y, z));  // IEEE sqrt
  x = __ddiv_rn(x, n);
  y = __ddiv_rn(y, n);
  z = __ddiv_rn(z, n);
}
__device__ __forceinline__ double clip1(double d) { return fmin(fmax(d, -1.0), 1.0); }

// FOV weight of one (vector, tile) pair, EU:124,133-136.  Returns 0 when the tile
// is outside fov/2 (the reference's dict then has no entry).
__device__ __forceinline__ double fov_weight(double dot, double max_d, double pf) {
  const double d = acos(clip1(dot));
  if (!(d < max_d)) return 0.0;
  const double r = __ddiv_rn(__dsub_rn(max_d, d), max_d);
  if (pf == 2.0) return __dmul_rn(r, r);
  if (pf == 1.0) return r;
  return pow(r, pf);
}

// ---- reductions -----------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// Deterministic block sum (fixed tree); `red` holds >= 32 doubles of shared memory.
// Every thread receives the result.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double r = (lane < nw) ? red[lane] : 0.0;
  r = warp_sum(r);
  return r;
}

// Normalised Shannon entropy of one histogram row, EU:195-209.
//   H = -sum_{t: hist>0} p log2 p,  p = hist/total
//   n = T if (weighted or total > T) else total;  Hn = H / (-n * (1/n) * log2(1/n))
// Called by all threads of the block; hist in shared memory.
// The sum runs on warp 0 alone, lane = t mod 32 and a butterfly at the end: the same order as k_entropy_rows, so a
// histogram gives the same bits whichever kernel finishes it (the unweighted cell and tile paths are interchangeable).
// norm_T: the tile count of the normalisation when it differs from the histogram length (naive
// lat/lon tiling: codes of the closed upper edges exist beyond num_tiles, EU:409,443-448); 0 = T.
__device__ __forceinline__ double normalized_entropy(const double* hist, int T, double total, bool by_tiles_always,
                                                     double* red, int norm_T = 0) {
  __syncthreads();
  if (threadIdx.x < 32) {
    double acc = 0.0;
    for (int t = threadIdx.x; t < T; t += 32) {
      const double w = hist[t];
      if (w > 0.0) {
        const double p = w / total;
        acc = fma(-p, log2(p), acc);  // spelled out: k_entropy_frames adds the same products from shared memory
      }
    }
    acc = warp_sum(acc);
    if (threadIdx.x == 0) red[0] = acc;
  }
  __syncthreads();
  const double Hs = red[0];
  const double nt = (double)(norm_T > 0 ? norm_T : T);
  const double n = (by_tiles_always || total > nt) ? nt : total;
  const double mp = 1.0 / n;
  const double mx = -n * mp * log2(mp);
  return Hs / mx;
}

}  // namespace vet
