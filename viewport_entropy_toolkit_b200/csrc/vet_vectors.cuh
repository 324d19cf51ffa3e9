// Direct (per-sample) kernels on arbitrary direction vectors: the functional API of the
// reference (compute_spatial_entropy / compute_transition_entropy on dicts of Vector,
// EU:147-332) and arbitrary tile-centre lists.  No cell tables are involved: every
// (user, tile) pair is evaluated in fp64 exactly like EU:41-67, and the per-tile sums
// are accumulated in USER ORDER like EU:190-192, so they match the reference's
// summation order.  O(U*T) per frame: for small frames and for validation, not the
// streaming path.
#pragma once
#include "vet_common.cuh"
#include "vet_stream.cuh"

namespace vet {

struct VecSpatialArgs {
  const double* vec;     // [F,U,3], NaN = missing user
  int64_t F, U;
  const double* unit;    // [T,3] unit tile centres
  int T;
  double max_d, pf;
  int use_weight;
  const int32_t* idx;    // [F,U] nearest tile (unweighted mode), -1 = missing
  double* per_k;         // [F] entropy of this tile count
  double* hist;          // [F,T] or null
  uint32_t* flags;
};

constexpr int kVecChunk = 128;

// One block per frame.  Weighted: thread t owns tile t (t, t+blockDim, ...) and walks the
// users in order.  Unweighted: integer histogram of the nearest-tile indices.
__global__ void __launch_bounds__(256) k_spatial_vectors(VecSpatialArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_hist = reinterpret_cast<double*>(smem_raw);             // [T]
  double* s_user = s_hist + a.T;                                    // [kVecChunk,3] normalised users
  uint32_t* s_ihist = reinterpret_cast<uint32_t*>(s_user + 3 * kVecChunk);  // [T]
  __shared__ double s_red[32];
  __shared__ uint32_t s_nv;
  for (int64_t f = blockIdx.x; f < a.F; f += gridDim.x) {
    if (threadIdx.x == 0) s_nv = 0u;
    for (int t = threadIdx.x; t < a.T; t += blockDim.x) {
      s_hist[t] = 0.0;
      s_ihist[t] = 0u;
    }
    __syncthreads();
    double total;
    if (a.use_weight) {
      for (int64_t u0 = 0; u0 < a.U; u0 += kVecChunk) {
        const int nu = (int)min((int64_t)kVecChunk, a.U - u0);
        uint32_t present = 0;
        for (int i = threadIdx.x; i < nu; i += blockDim.x) {
          const double* p = a.vec + ((f * a.U + u0 + i) * 3);
          double x = p[0], y = p[1], z = p[2];
          const bool ok = !(x != x || y != y || z != z);
          if (ok) normalize3(x, y, z);
          s_user[3 * i] = ok ? x : __longlong_as_double(0x7ff8000000000000LL);
          s_user[3 * i + 1] = y;
          s_user[3 * i + 2] = z;
          present += ok;
        }
        if (present) atomicAdd(&s_nv, present);
        __syncthreads();
        for (int t = threadIdx.x; t < a.T; t += blockDim.x) {
          const double bx = a.unit[3 * t], by = a.unit[3 * t + 1], bz = a.unit[3 * t + 2];
          double acc = s_hist[t];
          for (int i = 0; i < nu; ++i) {
            const double ax = s_user[3 * i];
            if (ax != ax) continue;  // missing user
            acc += fov_weight(dot3(ax, s_user[3 * i + 1], s_user[3 * i + 2], bx, by, bz), a.max_d, a.pf);  // EU:190-191
          }
          s_hist[t] = acc;
        }
        __syncthreads();
      }
      double part = 0.0;
      for (int t = threadIdx.x; t < a.T; t += blockDim.x) part += s_hist[t];
      total = block_sum(part, s_red);
    } else {
      uint32_t present = 0;
      for (int64_t u = threadIdx.x; u < a.U; u += blockDim.x) {
        const int t = a.idx[f * a.U + u];
        if (t >= 0) {
          atomicAdd(&s_ihist[t], 1u);
          ++present;
        }
      }
      if (present) atomicAdd(&s_nv, present);
      __syncthreads();
      for (int t = threadIdx.x; t < a.T; t += blockDim.x) s_hist[t] = (double)s_ihist[t];
      __syncthreads();
      total = (double)s_nv;
    }
    const uint32_t nv = s_nv;
    double e = normalized_entropy(s_hist, a.T, total, a.use_weight != 0, s_red);
    if (nv == 0) {
      e = __longlong_as_double(0x7ff8000000000000LL);
      if (threadIdx.x == 0) atomicOr(a.flags, (uint32_t)VET_FLAG_EMPTY_FRAME);
    }
    if (threadIdx.x == 0) a.per_k[f] = e;
    if (a.hist)
      for (int t = threadIdx.x; t < a.T; t += blockDim.x) a.hist[f * (int64_t)a.T + t] = s_hist[t];
    __syncthreads();
  }
}

// entropy[r] = (((0 + e_0) + e_1) + ...) / K   (SA:151-156 / TA:157-160)
__global__ void k_average_rows(const double* __restrict__ per_k, int K, int64_t n, int64_t stride,
                               double* __restrict__ out) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += per_k[k * stride + r];
    out[r] = s / (double)K;
  }
}

__global__ void k_idx_to_u16(const int32_t* __restrict__ idx, int64_t n, uint16_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = idx[i] < 0 ? (uint16_t)VET_MISSING : (uint16_t)idx[i];
}

// angular distances [n,T] (find_angular_distances, EU:70-87): arccos(clip(dot)), one warp per vector
__global__ void k_angular_distances(const double* __restrict__ vec, int64_t n, const double* __restrict__ unit, int T,
                                    double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    double ax = vec[3 * i], ay = vec[3 * i + 1], az = vec[3 * i + 2];
    normalize3(ax, ay, az);
    for (int t = lane; t < T; t += 32)
      out[i * T + t] = acos(clip1(dot3(ax, ay, az, unit[3 * t], unit[3 * t + 1], unit[3 * t + 2])));
  }
}

// vector_angle_distance (EU:41-67) for n independent pairs: both operands re-normalised, arccos(clip(dot))
__global__ void k_pair_angles(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double ax = a[3 * i], ay = a[3 * i + 1], az = a[3 * i + 2];
    double bx = b[3 * i], by = b[3 * i + 1], bz = b[3 * i + 2];
    normalize3(ax, ay, az);
    normalize3(bx, by, bz);
    out[i] = acos(clip1(dot3(ax, ay, az, bx, by, bz)));
  }
}

}  // namespace vet
