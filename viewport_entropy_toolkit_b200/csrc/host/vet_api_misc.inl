// C ABI: vet_naive_points, flag word, per-kernel profiling.
// Textual fragment of vet_b200.cu.

extern "C" int vet_naive_points(vet_handle* h, const double* lonlat_dev, int64_t F, int64_t U, int32_t tile_width,
                                int32_t tile_height, int32_t use_weight_distribution, double* entropy_dev,
                                int32_t* lon_idx_dev, int32_t* lat_idx_dev, void* stream) {
  if (!h || F < 0 || U < 0) return fail(VET_ERR_INVALID_ARG, "bad argument");
  // EU:404-417
  if (tile_width <= 0 || tile_height <= 0) return fail(VET_ERR_INVALID_ARG, "No tile dimensions provided");
  if (180 % tile_height != 0) return fail(VET_ERR_INVALID_ARG, "Tile height must divide 180!");
  if (360 % tile_width != 0) return fail(VET_ERR_INVALID_ARG, "Tile width must divide 360!");
  if (F == 0) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty radial points dictionary");
  if (!lonlat_dev || !entropy_dev) return fail(VET_ERR_INVALID_ARG, "null buffer");
  DeviceGuard guard(h->device);
  vet::NaivePointsArgs a{};
  a.lonlat = lonlat_dev;
  a.F = F;
  a.U = U;
  a.tile_width = tile_width;
  a.tile_height = tile_height;
  a.nlat1 = 180 / tile_height + 1;
  a.ncodes = (360 / tile_width + 1) * a.nlat1;
  a.num_tiles = (180 / tile_height) * (360 / tile_width);
  a.norm_always = use_weight_distribution ? 1 : 0;
  a.entropy = entropy_dev;
  a.lon_idx = lon_idx_dev;
  a.lat_idx = lat_idx_dev;
  a.flags = h->d_flags;
  const size_t smem = (size_t)a.ncodes * 4;
  if (smem + kStaticSmemSlack > h->smem_optin)
    return fail(VET_ERR_UNSUPPORTED, "%dx%d degree tiles give %d grid codes; too many for one frame's shared-memory histogram", tile_width, tile_height, a.ncodes);
  VET_CUDA(cudaFuncSetAttribute(vet::k_naive_points, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(h->smem_optin - kStaticSmemSlack)));
  h->launches++;
  vet::k_naive_points<<<(int)std::min<int64_t>(F, (int64_t)h->sm_count * 8), 256, smem, (cudaStream_t)stream>>>(a);
  VET_CUDA(cudaGetLastError());
  return VET_OK;
}

extern "C" int vet_set_option(vet_handle* h, int option, int value) {
  if (!h || option < 0 || option >= VET_OPT_COUNT) return fail(VET_ERR_INVALID_ARG, "bad option");
  static const int max_value[VET_OPT_COUNT] = {2, 3, 3, 2, 1, 1, 1, 1 << 30, 1024, 1, 1};
  if (value < 0 || value > max_value[option]) return fail(VET_ERR_INVALID_ARG, "option %d: value %d out of range", option, value);
  h->opt[option] = value;
  drop_graphs(h);  // captured launch sequences were chosen under the old options
  return VET_OK;
}

extern "C" int vet_get_option(const vet_handle* h, int option, int* value) {
  if (!h || !value || option < 0 || option >= VET_OPT_COUNT) return fail(VET_ERR_INVALID_ARG, "bad option");
  *value = h->opt[option];
  return VET_OK;
}

extern "C" int vet_poll_flags(vet_handle* h, void* stream, uint32_t* flags) {
  if (!h || !flags) return fail(VET_ERR_INVALID_ARG, "null argument");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  VET_CUDA(cudaMemcpyAsync(flags, h->d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  VET_CUDA(cudaMemsetAsync(h->d_flags, 0, sizeof(uint32_t), st));
  VET_CUDA(cudaStreamSynchronize(st));
  return VET_OK;
}

extern "C" int vet_profile_enable(vet_handle* h, int on) {
  if (!h) return fail(VET_ERR_INVALID_ARG, "null handle");
  DeviceGuard guard(h->device);
  for (auto& s : h->spans) {
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  h->spans.clear();
  h->profiling = on != 0;
  return VET_OK;
}

extern "C" int vet_profile_read(vet_handle* h, double* ms_by_kernel, int64_t* launches_by_kernel) {
  if (!h || !ms_by_kernel || !launches_by_kernel) return fail(VET_ERR_INVALID_ARG, "null argument");
  DeviceGuard guard(h->device);
  for (int i = 0; i < VET_KERNEL_COUNT; ++i) {
    ms_by_kernel[i] = 0.0;
    launches_by_kernel[i] = 0;
  }
  for (auto& s : h->spans) {
    VET_CUDA(cudaEventSynchronize(s.b));
    float ms = 0.f;
    VET_CUDA(cudaEventElapsedTime(&ms, s.a, s.b));
    ms_by_kernel[s.kernel] += ms;
    launches_by_kernel[s.kernel] += 1;
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  h->spans.clear();
  return VET_OK;
}
