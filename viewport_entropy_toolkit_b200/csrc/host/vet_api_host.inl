// C ABI: host-buffer variants (pinned staging, copies overlapped with the kernels).
// Textual fragment of vet_b200.cu.
// ---- host-buffer variants ---------------------------------------------------------

namespace {

// frames per host batch: about 256 MiB of packed input per copy; weighted handles take 512 frames when that
// stays under 1 GiB, so that the host path runs the same tensor-core weighted histogram as the device path
// (use_whist_i8: from 512 frames per batch) and returns the same bits
int64_t host_batch_frames(const vet_handle* h, int64_t F, int64_t U, size_t esz) {
  const size_t per_frame = std::max<size_t>((size_t)U * 3 * esz, 1);
  int64_t fb = std::max<int64_t>(2, (int64_t)(((size_t)256 << 20) / per_frame));
  if (h->use_weight && fb < 512 && (size_t)512 * per_frame <= ((size_t)1 << 30)) fb = 512;
  return std::min<int64_t>(F, fb);
}

}  // namespace

extern "C" int vet_spatial_host(vet_handle* h, const void* packed_host, int dtype, int64_t F, int64_t U,
                                double* entropy_host, double* per_k_host, double* hist0_host, uint16_t* assign0_host) {
  if (!h || F < 0 || U < 0 || (dtype != VET_F32 && dtype != VET_F64)) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (F == 0) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty vector dictionary");
  if (!packed_host || !entropy_host) return fail(VET_ERR_INVALID_ARG, "null buffer");
  DeviceGuard guard(h->device);
  h->call_frames = F;
  const size_t esz = dtype == VET_F32 ? 4 : 8;
  if (h->direct_only) {  // large-video mode: plain upload, direct kernels, download
    const int T0d = h->ts[0].T;
    void* d_in = nullptr;
    double *d_e = nullptr, *d_p = nullptr, *d_h = nullptr;
    uint16_t* d_a = nullptr;
    VET_CUDA(cudaMalloc(&d_in, (size_t)F * U * 3 * esz));
    VET_CUDA(cudaMalloc((void**)&d_e, (size_t)F * 8));
    if (per_k_host) VET_CUDA(cudaMalloc((void**)&d_p, (size_t)F * h->K * 8));
    if (hist0_host) VET_CUDA(cudaMalloc((void**)&d_h, (size_t)F * T0d * 8));
    if (assign0_host) VET_CUDA(cudaMalloc((void**)&d_a, (size_t)F * U * 2));
    cudaMemcpyAsync(d_in, packed_host, (size_t)F * U * 3 * esz, cudaMemcpyHostToDevice, h->s_exec);
    int rc = spatial_direct(h, d_in, dtype, F, U, d_e, d_p, d_h, d_a, h->s_exec);
    if (rc == VET_OK) {
      cudaMemcpyAsync(entropy_host, d_e, (size_t)F * 8, cudaMemcpyDeviceToHost, h->s_exec);
      if (per_k_host) cudaMemcpyAsync(per_k_host, d_p, (size_t)F * h->K * 8, cudaMemcpyDeviceToHost, h->s_exec);
      if (hist0_host) cudaMemcpyAsync(hist0_host, d_h, (size_t)F * T0d * 8, cudaMemcpyDeviceToHost, h->s_exec);
      if (assign0_host) cudaMemcpyAsync(assign0_host, d_a, (size_t)F * U * 2, cudaMemcpyDeviceToHost, h->s_exec);
    }
    cudaError_t e = cudaStreamSynchronize(h->s_exec);
    cudaFree(d_in);
    cudaFree(d_e);
    cudaFree(d_p);
    cudaFree(d_h);
    cudaFree(d_a);
    if (rc != VET_OK) return rc;
    if (e != cudaSuccess) return fail(VET_ERR_CUDA, "host-buffer pipeline failed: %s", cudaGetErrorString(e));
    return VET_OK;
  }
  const int64_t fb = host_batch_frames(h, F, U, esz);
  const size_t in_bytes = (size_t)fb * U * 3 * esz;
  if (h->in_bytes < in_bytes) {
    for (int i = 0; i < 2; ++i) {
      if (h->d_in[i]) VET_CUDA(cudaFree(h->d_in[i]));
      h->d_in[i] = nullptr;
    }
    h->in_bytes = 0;
    for (int i = 0; i < 2; ++i) VET_CUDA(cudaMalloc(&h->d_in[i], in_bytes));
    h->in_bytes = in_bytes;
  }
  const int T0 = h->ts[0].T;
  // device-side result buffers are kept in the handle and only grown (cudaMalloc/cudaFree synchronise)
  if (int rc = grow(&h->d_hout[0], &h->hout_bytes[0], (size_t)F * 8)) return rc;
  if (per_k_host)
    if (int rc = grow(&h->d_hout[1], &h->hout_bytes[1], (size_t)F * h->K * 8)) return rc;
  if (hist0_host)
    if (int rc = grow(&h->d_hout[2], &h->hout_bytes[2], (size_t)F * T0 * 8)) return rc;
  if (assign0_host)
    for (int i = 0; i < 2; ++i)
      if (int rc = grow(&h->d_hout[3 + i], &h->hout_bytes[3 + i], (size_t)fb * U * 2)) return rc;
  double* d_ent = (double*)h->d_hout[0];
  double* d_perk = per_k_host ? (double*)h->d_hout[1] : nullptr;
  double* d_hist = hist0_host ? (double*)h->d_hout[2] : nullptr;
  uint16_t* d_assign[2] = {assign0_host ? (uint16_t*)h->d_hout[3] : nullptr, assign0_host ? (uint16_t*)h->d_hout[4] : nullptr};
  // Three streams: copy-in, execute, copy-out.  Batch b+1 is uploaded while batch b runs and
  // batch b-1's assignments are downloaded (PCIe is full duplex).
  cudaEvent_t in_done[2], exec_done[2], out_done[2];
  for (int i = 0; i < 2; ++i) {
    VET_CUDA(cudaEventCreateWithFlags(&in_done[i], cudaEventDisableTiming));
    VET_CUDA(cudaEventCreateWithFlags(&exec_done[i], cudaEventDisableTiming));
    VET_CUDA(cudaEventCreateWithFlags(&out_done[i], cudaEventDisableTiming));
  }
  int rc = VET_OK;
  int b = 0;
  for (int64_t f0 = 0; f0 < F && rc == VET_OK; f0 += fb, b ^= 1) {
    const int64_t nf = std::min(fb, F - f0);
    cudaStreamWaitEvent(h->s_copy, exec_done[b], 0);  // input buffer b was last read two batches ago
    cudaMemcpyAsync(h->d_in[b], (const char*)packed_host + (size_t)f0 * U * 3 * esz, (size_t)nf * U * 3 * esz,
                    cudaMemcpyHostToDevice, h->s_copy);
    cudaEventRecord(in_done[b], h->s_copy);
    cudaStreamWaitEvent(h->s_exec, in_done[b], 0);
    cudaStreamWaitEvent(h->s_exec, out_done[b], 0);  // assignment buffer b must have been downloaded
    const int64_t fbs = frames_per_batch(h, nf, U, false);
    rc = grow((void**)&h->d_cnt, &h->cnt_bytes, cnt_scratch_bytes(h, fbs));
    if (rc == VET_OK) rc = grow((void**)&h->d_nvalid, &h->nvalid_bytes, (size_t)fbs * 4);
    for (int64_t g0 = 0; g0 < nf && rc == VET_OK; g0 += fbs) {
      const int64_t ng = std::min(fbs, nf - g0);
      const char* in = (const char*)h->d_in[b] + (size_t)g0 * U * 3 * esz;
      TilesPlan tp = plan_tiles(h, in, U);
      if (tp.ok) {
        rc = launch_stream_tiles(h, tp, in, dtype, ng, U, d_assign[b] ? d_assign[b] + g0 * U : nullptr, h->s_exec);
        if (rc == VET_OK)
          rc = launch_tiles_epilogue(h, tp, ng, d_ent + f0 + g0, d_perk ? d_perk + f0 + g0 : nullptr, F,
                                     d_hist ? d_hist + (f0 + g0) * T0 : nullptr, h->s_exec);
        continue;
      }
      rc = launch_stream(h, in, dtype, ng, U, d_assign[b] ? d_assign[b] + g0 * U : nullptr, false, h->s_exec);
      if (rc == VET_OK)
        rc = launch_epilogue(h, ng, U, d_ent + f0 + g0, d_perk ? d_perk + f0 + g0 : nullptr, F,
                             d_hist ? d_hist + (f0 + g0) * T0 : nullptr, h->s_exec);
    }
    cudaEventRecord(exec_done[b], h->s_exec);
    if (rc == VET_OK && assign0_host) {
      cudaStreamWaitEvent(h->s_out, exec_done[b], 0);
      cudaMemcpyAsync(assign0_host + f0 * U, d_assign[b], (size_t)nf * U * 2, cudaMemcpyDeviceToHost, h->s_out);
      cudaEventRecord(out_done[b], h->s_out);
    }
  }
  if (rc == VET_OK) {
    cudaMemcpyAsync(entropy_host, d_ent, (size_t)F * 8, cudaMemcpyDeviceToHost, h->s_exec);
    if (per_k_host) cudaMemcpyAsync(per_k_host, d_perk, (size_t)F * h->K * 8, cudaMemcpyDeviceToHost, h->s_exec);
    if (hist0_host) cudaMemcpyAsync(hist0_host, d_hist, (size_t)F * T0 * 8, cudaMemcpyDeviceToHost, h->s_exec);
  }
  cudaError_t e1 = cudaStreamSynchronize(h->s_copy), e2 = cudaStreamSynchronize(h->s_exec),
              e3 = cudaStreamSynchronize(h->s_out);
  for (int i = 0; i < 2; ++i) {
    cudaEventDestroy(in_done[i]);
    cudaEventDestroy(exec_done[i]);
    cudaEventDestroy(out_done[i]);
  }
  if (rc != VET_OK) return rc;
  for (cudaError_t e : {e1, e2, e3})
    if (e != cudaSuccess) return fail(VET_ERR_CUDA, "host-buffer pipeline failed: %s", cudaGetErrorString(e));
  return VET_OK;
}

extern "C" int vet_transition_host(vet_handle* h, const void* packed_host, int dtype, int64_t F, int64_t U,
                                   double* entropy_host, double* per_k_host, int32_t* prev_count0_host,
                                   uint16_t* pairs0_host, int mode) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_transition_host: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || F < 0 || U < 0 || (dtype != VET_F32 && dtype != VET_F64)) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (F <= 1) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty vector dictionary");
  if (!packed_host || !entropy_host) return fail(VET_ERR_INVALID_ARG, "null buffer");
  DeviceGuard guard(h->device);
  const size_t esz = dtype == VET_F32 ? 4 : 8;
  const int T0 = h->ts[0].T;
  void* d_in = nullptr;
  double *d_ent = nullptr, *d_perk = nullptr;
  int32_t* d_pc = nullptr;
  uint16_t* d_pairs = nullptr;
  const size_t in_bytes = (size_t)F * U * 3 * esz;
  VET_CUDA(cudaMalloc(&d_in, in_bytes));
  VET_CUDA(cudaMalloc((void**)&d_ent, (size_t)(F - 1) * 8));
  if (per_k_host) VET_CUDA(cudaMalloc((void**)&d_perk, (size_t)(F - 1) * h->K * 8));
  if (prev_count0_host) VET_CUDA(cudaMalloc((void**)&d_pc, (size_t)(F - 1) * T0 * 4));
  if (pairs0_host) VET_CUDA(cudaMalloc((void**)&d_pairs, (size_t)(F - 1) * U * 4));
  cudaMemcpyAsync(d_in, packed_host, in_bytes, cudaMemcpyHostToDevice, h->s_exec);
  int rc = vet_transition(h, d_in, dtype, F, U, d_ent, d_perk, d_pc, d_pairs, mode, h->s_exec);
  if (rc == VET_OK) {
    cudaMemcpyAsync(entropy_host, d_ent, (size_t)(F - 1) * 8, cudaMemcpyDeviceToHost, h->s_exec);
    if (per_k_host) cudaMemcpyAsync(per_k_host, d_perk, (size_t)(F - 1) * h->K * 8, cudaMemcpyDeviceToHost, h->s_exec);
    if (prev_count0_host)
      cudaMemcpyAsync(prev_count0_host, d_pc, (size_t)(F - 1) * T0 * 4, cudaMemcpyDeviceToHost, h->s_exec);
    if (pairs0_host) cudaMemcpyAsync(pairs0_host, d_pairs, (size_t)(F - 1) * U * 4, cudaMemcpyDeviceToHost, h->s_exec);
  }
  cudaError_t e = cudaStreamSynchronize(h->s_exec);
  cudaFree(d_in);
  cudaFree(d_ent);
  cudaFree(d_perk);
  cudaFree(d_pc);
  cudaFree(d_pairs);
  if (rc != VET_OK) return rc;
  if (e != cudaSuccess) return fail(VET_ERR_CUDA, "host-buffer pipeline failed: %s", cudaGetErrorString(e));
  return VET_OK;
}
