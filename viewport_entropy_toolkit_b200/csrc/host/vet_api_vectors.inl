// C ABI on ARBITRARY direction vectors (the reference's free functions) and the direct per-sample regime.
// Textual fragment of vet_b200.cu.
namespace {

int launch_nearest_i32(vet_handle* h, int k, const double* vec, int64_t n, int32_t* idx, cudaStream_t st) {
  const TileSet& t = h->ts[k];
  const size_t smem = (size_t)t.T * 3 * sizeof(double);
  VET_CUDA(cudaFuncSetAttribute(vet::k_nearest<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = 256;
  const int64_t rounds = (n + threads / 4 - 1) / (threads / 4);
  const int blocks = (int)std::min<int64_t>(rounds, (int64_t)h->sm_count * 8);
  vet::k_nearest<int32_t><<<blocks, threads, smem, st>>>(vec, n, t.d_unit, t.T, idx);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

int spatial_vectors_impl(vet_handle* h, const double* vec, int64_t F, int64_t U, double* entropy, double* per_k,
                         int64_t per_k_stride, double* hist0, uint16_t* assign0, cudaStream_t st) {
  const int64_t n = F * U;
  const bool need_idx = !h->use_weight || assign0;
  if (need_idx)
    if (int rc = grow(&h->d_vscratch[0], &h->vscratch_bytes[0], (size_t)n * 4)) return rc;
  double* pk = per_k;
  int64_t pk_stride = per_k_stride;
  if (!pk) {
    if (int rc = grow(&h->d_vscratch[1], &h->vscratch_bytes[1], (size_t)h->K * F * 8)) return rc;
    pk = (double*)h->d_vscratch[1];
    pk_stride = F;
  }
  int32_t* idx = (int32_t*)h->d_vscratch[0];
  for (int k = 0; k < h->K; ++k) {
    const TileSet& t = h->ts[k];
    if (!h->use_weight || (k == 0 && assign0)) {
      if (int rc = launch_nearest_i32(h, k, vec, n, idx, st)) return rc;
      if (k == 0 && assign0) {
        vet::k_idx_to_u16<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 8), 256, 0, st>>>(idx, n, assign0);
        h->launches++;
      }
    }
    vet::VecSpatialArgs a{};
    a.vec = vec;
    a.F = F;
    a.U = U;
    a.unit = t.d_unit;
    a.T = t.T;
    a.max_d = h->max_d;
    a.pf = h->pf;
    a.use_weight = h->use_weight;
    a.idx = idx;
    a.per_k = pk + k * pk_stride;
    a.hist = (k == 0) ? hist0 : nullptr;
    a.flags = h->d_flags;
    const size_t smem = (size_t)t.T * 12 + (size_t)vet::kVecChunk * 24 + 16;
    VET_CUDA(cudaFuncSetAttribute(vet::k_spatial_vectors, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vet::k_spatial_vectors<<<(int)std::min<int64_t>(F, (int64_t)h->sm_count * 4), 256, smem, st>>>(a);
    h->launches++;
    VET_CUDA(cudaGetLastError());
  }
  vet::k_average_rows<<<(int)std::min<int64_t>((F + 255) / 256, 1024), 256, 0, st>>>(pk, h->K, F, pk_stride, entropy);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

int transition_vectors_impl(vet_handle* h, const double* vec, int64_t F, int64_t U, double* entropy, double* per_k,
                            int64_t per_k_stride, int32_t* prev_count0, uint16_t* pairs0, int mode, cudaStream_t st) {
  const int64_t n = F * U;
  if (int rc = grow(&h->d_vscratch[0], &h->vscratch_bytes[0], (size_t)n * 4)) return rc;
  double* pk = per_k;
  int64_t pk_stride = per_k_stride;
  if (!pk) {
    if (int rc = grow(&h->d_vscratch[1], &h->vscratch_bytes[1], (size_t)h->K * (F - 1) * 8)) return rc;
    pk = (double*)h->d_vscratch[1];
    pk_stride = F - 1;
  }
  int32_t* idx = (int32_t*)h->d_vscratch[0];
  for (int k = 0; k < h->K; ++k) {
    if (int rc = launch_nearest_i32(h, k, vec, n, idx, st)) return rc;
    vet::TransitionArgs a{};
    a.cell32 = idx;  // tile indices play the role of cell ids, mapped through the identity LUT
    a.F = F;
    a.U = U;
    a.K = 1;
    a.T[0] = h->ts[k].T;
    a.lut[0] = h->d_identity;
    a.entropy = pk + k * pk_stride;  // K == 1: the "mean" is the tile count's own entropy
    a.per_k = nullptr;
    a.prev_count0 = (k == 0) ? prev_count0 : nullptr;
    a.pairs0 = (k == 0) ? pairs0 : nullptr;
    a.mode = mode;
    a.flags = h->d_flags;
    if (int rc = launch_transition(h, a, F - 1, U, h->ts[k].T, st)) return rc;
  }
  vet::k_average_rows<<<(int)std::min<int64_t>((F - 1 + 255) / 256, 1024), 256, 0, st>>>(pk, h->K, F - 1, pk_stride, entropy);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

// decode a frame batch of packed samples into the vector scratch (direct mode)
int decode_batch(vet_handle* h, const void* packed, int dtype, int64_t n, cudaStream_t st) {
  if (int rc = grow(&h->d_vscratch[2], &h->vscratch_bytes[2], (size_t)n * 24)) return rc;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16);
  if (dtype == VET_F32)
    vet::k_decode<float><<<blocks, 256, 0, st>>>((const float*)packed, n, h->W, h->H, h->d_cosT, h->d_sinT, h->d_sinP,
                                                 h->d_cosP, (double*)h->d_vscratch[2], nullptr, h->d_flags);
  else
    vet::k_decode<double><<<blocks, 256, 0, st>>>((const double*)packed, n, h->W, h->H, h->d_cosT, h->d_sinT, h->d_sinP,
                                                  h->d_cosP, (double*)h->d_vscratch[2], nullptr, h->d_flags);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

int spatial_direct(vet_handle* h, const void* packed, int dtype, int64_t F, int64_t U, double* entropy, double* per_k,
                   double* hist0, uint16_t* assign0, cudaStream_t st) {
  const size_t esz = dtype == VET_F32 ? 4 : 8;
  const int T0 = h->ts[0].T;
  const int64_t fb = std::min<int64_t>(F, std::max<int64_t>(1, (int64_t)(((size_t)1 << 30) / ((size_t)U * 24))));
  for (int64_t f0 = 0; f0 < F; f0 += fb) {
    const int64_t nf = std::min(fb, F - f0);
    if (int rc = decode_batch(h, (const char*)packed + (size_t)f0 * U * 3 * esz, dtype, nf * U, st)) return rc;
    if (int rc = spatial_vectors_impl(h, (const double*)h->d_vscratch[2], nf, U, entropy + f0, per_k ? per_k + f0 : nullptr, F,
                                      hist0 ? hist0 + f0 * T0 : nullptr, assign0 ? assign0 + f0 * U : nullptr, st))
      return rc;
  }
  return VET_OK;
}

int transition_direct(vet_handle* h, const void* packed, int dtype, int64_t F, int64_t U, double* entropy, double* per_k,
                      int32_t* prev_count0, uint16_t* pairs0, int mode, cudaStream_t st) {
  const size_t esz = dtype == VET_F32 ? 4 : 8;
  const int T0 = h->ts[0].T;
  const int64_t fb = std::min<int64_t>(F, std::max<int64_t>(2, (int64_t)(((size_t)1 << 30) / ((size_t)U * 24))));
  for (int64_t f0 = 0; f0 < F - 1; f0 += fb - 1) {  // batches overlap by the halo frame
    const int64_t nf = std::min(fb, F - f0);
    if (int rc = decode_batch(h, (const char*)packed + (size_t)f0 * U * 3 * esz, dtype, nf * U, st)) return rc;
    if (int rc = transition_vectors_impl(h, (const double*)h->d_vscratch[2], nf, U, entropy + f0,
                                         per_k ? per_k + f0 : nullptr, F - 1, prev_count0 ? prev_count0 + f0 * T0 : nullptr,
                                         pairs0 ? pairs0 + f0 * U * 2 : nullptr, mode, st))
      return rc;
    if (nf == F - f0) break;
  }
  return VET_OK;
}

}  // namespace

extern "C" int vet_angular_distances(vet_handle* h, int k, const double* vec_dev, int64_t n, double* d_dev, void* stream) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_angular_distances: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || k < 0 || k >= h->K || n < 0 || (n > 0 && (!vec_dev || !d_dev))) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (n == 0) return VET_OK;
  DeviceGuard guard(h->device);
  const TileSet& t = h->ts[k];
  const int blocks = (int)std::min<int64_t>((n + 7) / 8, (int64_t)h->sm_count * 8);
  vet::k_angular_distances<<<blocks, 256, 0, (cudaStream_t)stream>>>(vec_dev, n, t.d_unit, t.T, d_dev);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

extern "C" int vet_vector_angles(vet_handle* h, const double* a_dev, const double* b_dev, int64_t n, double* d_dev, void* stream) {
  if (!h || n < 0 || (n > 0 && (!a_dev || !b_dev || !d_dev))) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (n == 0) return VET_OK;
  DeviceGuard guard(h->device);
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 8);
  vet::k_pair_angles<<<blocks, 256, 0, (cudaStream_t)stream>>>(a_dev, b_dev, n, d_dev);
  h->launches++;
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

extern "C" int vet_spatial_vectors(vet_handle* h, const double* vec_dev, int64_t F, int64_t U, double* entropy_dev,
                                   double* per_k_dev, double* hist0_dev, uint16_t* assign0_dev, void* stream) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_spatial_vectors: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || F < 0 || U < 0) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (F == 0) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty vector dictionary");  // EU:168-169
  if (!vec_dev || !entropy_dev) return fail(VET_ERR_INVALID_ARG, "null buffer");
  DeviceGuard guard(h->device);
  return spatial_vectors_impl(h, vec_dev, F, U, entropy_dev, per_k_dev, F, hist0_dev, assign0_dev, (cudaStream_t)stream);
}

extern "C" int vet_transition_vectors(vet_handle* h, const double* vec_dev, int64_t F, int64_t U, double* entropy_dev,
                                      double* per_k_dev, int32_t* prev_count0_dev, uint16_t* pairs0_dev, int mode,
                                      void* stream) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_transition_vectors: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || F < 0 || U < 0) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (mode != VET_TRANSITION_LITERAL && mode != VET_TRANSITION_TEXTBOOK) return fail(VET_ERR_INVALID_ARG, "bad mode");
  if (F <= 1) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty vector dictionary");  // EU:239-240
  if (!vec_dev || !entropy_dev) return fail(VET_ERR_INVALID_ARG, "null buffer");
  if (U >= 0xFFFFFFFFll) return fail(VET_ERR_UNSUPPORTED, "too many users");
  DeviceGuard guard(h->device);
  return transition_vectors_impl(h, vec_dev, F, U, entropy_dev, per_k_dev, F - 1, prev_count0_dev, pairs0_dev, mode,
                                 (cudaStream_t)stream);
}
