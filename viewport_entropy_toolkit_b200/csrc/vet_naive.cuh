// Latitude/longitude grid tiling of the reference's NaiveSpatialEntropyAnalyzer for ARBITRARY
// RadialPoints (compute_naive_spatial_entropy, EU:362-453; find_naive_tile_index, EU:360-381).
// Packed (time, 2dmu, 2dmv) input does not come through here: for it the grid tiling is just
// another cell -> tile LUT of the streaming kernels (vet_create with naive_tile_width > 0).
#pragma once
#include "vet_common.cuh"

namespace vet {

struct NaivePointsArgs {
  const double* lonlat;  // [F,U,2] degrees, NaN = absent
  int64_t F, U;
  int tile_width, tile_height;
  int nlat1;       // 180/tile_height + 1 latitude codes (the closed upper edge lat = 90 has its own)
  int ncodes;      // (360/tile_width + 1) * nlat1
  int num_tiles;   // (180/tile_height) * (360/tile_width), EU:409
  int norm_always; // config.use_weight_distribution, EU:443
  double* entropy;   // [F]
  int32_t* lon_idx;  // [F,U] or null
  int32_t* lat_idx;  // [F,U] or null
  uint32_t* flags;
};

// One CTA per frame; dynamic shared memory: ncodes counters.
__global__ void __launch_bounds__(256) k_naive_points(NaivePointsArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(smem_raw);
  __shared__ double s_red[32];
  __shared__ uint32_t s_n;
  for (int64_t f = blockIdx.x; f < a.F; f += gridDim.x) {
    for (int t = threadIdx.x; t < a.ncodes; t += blockDim.x) s_cnt[t] = 0u;
    if (threadIdx.x == 0) s_n = 0u;
    __syncthreads();
    uint32_t nv = 0, bad = 0;
    for (int64_t u = threadIdx.x; u < a.U; u += blockDim.x) {
      const double lon = a.lonlat[(f * a.U + u) * 2], lat = a.lonlat[(f * a.U + u) * 2 + 1];
      int li = -1, la = -1;
      if (lon == lon && lat == lat) {
        if (lon < -180.0 || lon > 180.0 || lat < -90.0 || lat > 90.0) {  // RadialPoint validation, DT:78-83
          bad = 1;
        } else {
          li = __double2int_rz(__ddiv_rn(__dadd_rn(lon, 180.0), (double)a.tile_width));   // EU:378
          la = __double2int_rz(__ddiv_rn(__dadd_rn(lat, 90.0), (double)a.tile_height));   // EU:379
          atomicAdd(&s_cnt[li * a.nlat1 + la], 1u);
          ++nv;
        }
      }
      if (a.lon_idx) a.lon_idx[f * a.U + u] = li;
      if (a.lat_idx) a.lat_idx[f * a.U + u] = la;
    }
    nv = __reduce_add_sync(kFull, nv);
    if ((threadIdx.x & 31) == 0 && nv) atomicAdd(&s_n, nv);
    if (bad) atomicOr(a.flags, (uint32_t)VET_FLAG_OUT_OF_RANGE);
    __syncthreads();
    const double total = (double)s_n;
    double acc = 0.0;
    for (int t = threadIdx.x; t < a.ncodes; t += blockDim.x) {
      const uint32_t c = s_cnt[t];
      if (c) {
        const double p = (double)c / total;  // EU:437-439
        acc -= p * log2(p);
      }
    }
    const double Hs = block_sum(acc, s_red);
    const double nt = (double)a.num_tiles;
    const double n = (a.norm_always || total > nt) ? nt : total;  // EU:443-448
    const double mp = 1.0 / n;
    double e = Hs / (-n * mp * log2(mp));
    if (total == 0.0) {
      e = __longlong_as_double(0x7ff8000000000000LL);
      if (threadIdx.x == 0) atomicOr(a.flags, (uint32_t)VET_FLAG_EMPTY_FRAME);  // EU:404-405
    }
    if (threadIdx.x == 0) a.entropy[f] = e;
    __syncthreads();
  }
}

}  // namespace vet
