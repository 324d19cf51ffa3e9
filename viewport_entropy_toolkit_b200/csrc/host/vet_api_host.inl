// C ABI: host-buffer variants (pinned staging, copies overlapped with the kernels).
// Textual fragment of vet_b200.cu.
// ---- host-buffer variants ---------------------------------------------------------

namespace {

// frames per host batch: about 256 MiB of packed input per copy; weighted handles take 512 frames when that
// stays under 1 GiB, so that the host path runs the same tensor-core weighted histogram as the device path
// (use_whist_i8: decided per call from VET_I8_MIN_FRAMES frames on) and returns the same bits
int64_t host_batch_frames(const vet_handle* h, int64_t F, int64_t U, size_t esz) {
  if (h->opt[VET_OPT_HOST_BATCH_FRAMES] > 0) return std::min<int64_t>(F, std::max(2, h->opt[VET_OPT_HOST_BATCH_FRAMES]));
  const size_t per_frame = std::max<size_t>((size_t)U * 3 * esz, 1);
  int64_t fb = std::max<int64_t>(2, (int64_t)(((size_t)256 << 20) / per_frame));
  if (h->use_weight && fb < 512 && (size_t)512 * per_frame <= ((size_t)1 << 30)) fb = 512;
  return std::min<int64_t>(F, fb);
}

// First CUDA error of a pipeline: later calls are skipped, the error is reported after the streams have drained.
struct PipeStatus {
  int rc = VET_OK;
  bool ok() const { return rc == VET_OK; }
  void cuda(cudaError_t e, const char* what) {
    if (rc == VET_OK && e != cudaSuccess) rc = fail(VET_ERR_CUDA, "host-buffer pipeline: %s failed: %s", what, cudaGetErrorString(e));
  }
  void call(int r) {
    if (rc == VET_OK && r != VET_OK) rc = r;
  }
};
#define VET_PIPE(ps, expr) (ps).cuda((ps).ok() ? (expr) : cudaSuccess, #expr)

// Waits for the three pipeline streams (also after an error: nothing may still be using the caller's buffers).
int drain_pipeline(vet_handle* h, PipeStatus& ps) {
  const cudaError_t e1 = cudaStreamSynchronize(h->s_copy), e2 = cudaStreamSynchronize(h->s_exec),
                    e3 = cudaStreamSynchronize(h->s_out);
  if (!ps.ok()) return ps.rc;
  for (cudaError_t e : {e1, e2, e3})
    if (e != cudaSuccess) return fail(VET_ERR_CUDA, "host-buffer pipeline failed: %s", cudaGetErrorString(e));
  return VET_OK;
}

// Temporaries of the monolithic paths (large-video direct regime): freed on every exit.
struct DeviceTemps {
  std::vector<void*> ptrs;
  ~DeviceTemps() {
    for (void* p : ptrs) cudaFree(p);
  }
  int alloc(void** out, size_t bytes) {
    *out = nullptr;
    VET_CUDA(cudaMalloc(out, std::max<size_t>(bytes, 1)));
    ptrs.push_back(*out);
    return VET_OK;
  }
};

// Upload of `nf` frames of host records into `dst` (device, three columns) on the copy stream.  Two-column host layout
// (VET_OPT_HOST_LAYOUT): the (2dmu, 2dmv) records go to the staging buffer `b` and are widened on the EXECUTE stream after
// the copy; *widen is set so that the caller launches it behind its in_done wait.
struct Upload {
  const void* src = nullptr;  // staging buffer to widen from, or null
  void* dst = nullptr;
  int64_t n = 0;              // records
};
int grow_inputs2(vet_handle* h, size_t bytes) {
  if (h->in2_bytes >= bytes) return VET_OK;
  for (int i = 0; i < 2; ++i) {
    if (h->d_in2[i]) VET_CUDA(cudaFree(h->d_in2[i]));
    h->d_in2[i] = nullptr;
  }
  h->in2_bytes = 0;
  for (int i = 0; i < 2; ++i) VET_CUDA(cudaMalloc(&h->d_in2[i], bytes));
  h->in2_bytes = bytes;
  return VET_OK;
}
void upload_frames(vet_handle* h, PipeStatus& ps, const void* host, int rec, size_t esz, int64_t f0, int64_t nf, int64_t U, int b,
                   void* dst, Upload* widen) {
  *widen = Upload{};
  const size_t rec_bytes = (size_t)rec * esz;
  const char* src = (const char*)host + (size_t)f0 * U * rec_bytes;
  if (rec == 3) {
    VET_PIPE(ps, cudaMemcpyAsync(dst, src, (size_t)nf * U * rec_bytes, cudaMemcpyHostToDevice, h->s_copy));
    return;
  }
  VET_PIPE(ps, cudaMemcpyAsync(h->d_in2[b], src, (size_t)nf * U * rec_bytes, cudaMemcpyHostToDevice, h->s_copy));
  widen->src = h->d_in2[b];
  widen->dst = dst;
  widen->n = nf * U;
}
void widen_records(vet_handle* h, PipeStatus& ps, const Upload& w, int dtype, cudaStream_t st) {
  if (!w.src || !ps.ok()) return;
  const int blocks = (int)std::min<int64_t>((w.n + 255) / 256, (int64_t)h->sm_count * 16);
  h->launches++;
  if (dtype == VET_F32) vet::k_widen_records<float><<<blocks, 256, 0, st>>>((const float*)w.src, w.n, (float*)w.dst);
  else vet::k_widen_records<double><<<blocks, 256, 0, st>>>((const double*)w.src, w.n, (double*)w.dst);
  ps.cuda(cudaGetLastError(), "k_widen_records");
}

// Large-video (direct) regime: plain upload, direct kernels, download.  O(users x tiles) per frame on the device: the
// copies are not what bounds it.
int host_direct(vet_handle* h, const void* packed_host, int dtype, int64_t F, int64_t U, bool spatial, bool transition,
                double* sp_entropy, double* sp_per_k, double* hist0, uint16_t* assign0, double* tr_entropy, double* tr_per_k,
                int32_t* prev_count0, uint16_t* pairs0, int mode) {
  const size_t esz = dtype == VET_F32 ? 4 : 8;
  const int T0 = h->ts[0].T;
  const int64_t R = std::max<int64_t>(F - 1, 0);
  DeviceTemps t;
  void *d_in, *d_e = nullptr, *d_p = nullptr, *d_h = nullptr, *d_a = nullptr, *d_te = nullptr, *d_tp = nullptr, *d_pc = nullptr,
             *d_pr = nullptr;
  if (int rc = t.alloc(&d_in, (size_t)F * U * 3 * esz)) return rc;
  if (spatial) {
    if (int rc = t.alloc(&d_e, (size_t)F * 8)) return rc;
    if (sp_per_k)
      if (int rc = t.alloc(&d_p, (size_t)F * h->K * 8)) return rc;
    if (hist0)
      if (int rc = t.alloc(&d_h, (size_t)F * T0 * 8)) return rc;
    if (assign0)
      if (int rc = t.alloc(&d_a, (size_t)F * U * 2)) return rc;
  }
  if (transition && R > 0) {
    if (int rc = t.alloc(&d_te, (size_t)R * 8)) return rc;
    if (tr_per_k)
      if (int rc = t.alloc(&d_tp, (size_t)R * h->K * 8)) return rc;
    if (prev_count0)
      if (int rc = t.alloc(&d_pc, (size_t)R * T0 * 4)) return rc;
    if (pairs0)
      if (int rc = t.alloc(&d_pr, (size_t)R * U * 4)) return rc;
  }
  PipeStatus ps;
  cudaStream_t st = h->s_exec;
  if (h->opt[VET_OPT_HOST_LAYOUT]) {  // two-column host records: widened on the device
    void* d_uv = nullptr;
    if (int rc = t.alloc(&d_uv, (size_t)F * U * 2 * esz)) return rc;
    VET_PIPE(ps, cudaMemcpyAsync(d_uv, packed_host, (size_t)F * U * 2 * esz, cudaMemcpyHostToDevice, st));
    Upload w;
    w.src = d_uv;
    w.dst = d_in;
    w.n = F * U;
    widen_records(h, ps, w, dtype, st);
  } else {
    VET_PIPE(ps, cudaMemcpyAsync(d_in, packed_host, (size_t)F * U * 3 * esz, cudaMemcpyHostToDevice, st));
  }
  if (ps.ok() && spatial)
    ps.call(spatial_direct(h, d_in, dtype, F, U, (double*)d_e, (double*)d_p, (double*)d_h, (uint16_t*)d_a, st));
  if (ps.ok() && transition && R > 0)
    ps.call(transition_direct(h, d_in, dtype, F, U, (double*)d_te, (double*)d_tp, (int32_t*)d_pc, (uint16_t*)d_pr, mode, st));
  if (spatial) {
    VET_PIPE(ps, cudaMemcpyAsync(sp_entropy, d_e, (size_t)F * 8, cudaMemcpyDeviceToHost, st));
    if (sp_per_k) VET_PIPE(ps, cudaMemcpyAsync(sp_per_k, d_p, (size_t)F * h->K * 8, cudaMemcpyDeviceToHost, st));
    if (hist0) VET_PIPE(ps, cudaMemcpyAsync(hist0, d_h, (size_t)F * T0 * 8, cudaMemcpyDeviceToHost, st));
    if (assign0) VET_PIPE(ps, cudaMemcpyAsync(assign0, d_a, (size_t)F * U * 2, cudaMemcpyDeviceToHost, st));
  }
  if (transition && R > 0) {
    VET_PIPE(ps, cudaMemcpyAsync(tr_entropy, d_te, (size_t)R * 8, cudaMemcpyDeviceToHost, st));
    if (tr_per_k) VET_PIPE(ps, cudaMemcpyAsync(tr_per_k, d_tp, (size_t)R * h->K * 8, cudaMemcpyDeviceToHost, st));
    if (prev_count0) VET_PIPE(ps, cudaMemcpyAsync(prev_count0, d_pc, (size_t)R * T0 * 4, cudaMemcpyDeviceToHost, st));
    if (pairs0) VET_PIPE(ps, cudaMemcpyAsync(pairs0, d_pr, (size_t)R * U * 4, cudaMemcpyDeviceToHost, st));
  }
  return drain_pipeline(h, ps);
}

// Staging buffers of the pipelines: two input buffers of `frames` frames each (kept in the handle, only grown).
int grow_inputs(vet_handle* h, size_t bytes) {
  if (h->in_bytes >= bytes) return VET_OK;
  for (int i = 0; i < 2; ++i) {
    if (h->d_in[i]) VET_CUDA(cudaFree(h->d_in[i]));
    h->d_in[i] = nullptr;
  }
  h->in_bytes = 0;
  for (int i = 0; i < 2; ++i) VET_CUDA(cudaMalloc(&h->d_in[i], bytes));
  h->in_bytes = bytes;
  return VET_OK;
}

// TransitionEntropyAnalyzer (spatial == false) or both analyzers (spatial == true) on a HOST tensor, in frame batches
// on three streams: batch b+1 is uploaded while batch b runs and the per-user results of batch b-1 are downloaded.
// Transition row r needs frames r and r+1: the last frame of a batch is carried over ON THE DEVICE as the first
// ("halo") frame of the next one (one device-to-device copy of a frame; nothing is uploaded twice).
int host_pipeline_halo(vet_handle* h, const void* packed_host, int dtype, int64_t F, int64_t U, bool spatial,
                       double* sp_entropy, double* sp_per_k, double* hist0, uint16_t* assign0, double* tr_entropy,
                       double* tr_per_k, int32_t* prev_count0, uint16_t* pairs0, int mode) {
  const size_t esz = dtype == VET_F32 ? 4 : 8;
  const size_t frame_bytes = (size_t)U * 3 * esz;
  const int T0 = h->ts[0].T;
  const int64_t R = F - 1;
  const int64_t fb = host_batch_frames(h, F, U, esz);  // NEW frames per batch; the buffers hold one more (the halo)
  if (int rc = grow_inputs(h, (size_t)(fb + 1) * frame_bytes)) return rc;
  const int rec = h->opt[VET_OPT_HOST_LAYOUT] ? 2 : 3;
  if (rec == 2)
    if (int rc = grow_inputs2(h, (size_t)fb * U * 2 * esz)) return rc;
  // device-side result rows are kept in the handle and only grown (cudaMalloc / cudaFree synchronise)
  if (spatial) {
    if (int rc = grow(&h->d_hout[0], &h->hout_bytes[0], (size_t)F * 8)) return rc;
    if (sp_per_k)
      if (int rc = grow(&h->d_hout[1], &h->hout_bytes[1], (size_t)F * h->K * 8)) return rc;
    if (hist0)
      if (int rc = grow(&h->d_hout[2], &h->hout_bytes[2], (size_t)F * T0 * 8)) return rc;
    if (assign0)
      for (int i = 0; i < 2; ++i)
        if (int rc = grow(&h->d_hout[3 + i], &h->hout_bytes[3 + i], (size_t)(fb + 1) * U * 2)) return rc;
  }
  if (int rc = grow(&h->d_hout2[0], &h->hout2_bytes[0], (size_t)std::max<int64_t>(R, 1) * 8)) return rc;
  if (tr_per_k)
    if (int rc = grow(&h->d_hout2[1], &h->hout2_bytes[1], (size_t)std::max<int64_t>(R, 1) * h->K * 8)) return rc;
  if (prev_count0)
    if (int rc = grow(&h->d_hout2[2], &h->hout2_bytes[2], (size_t)std::max<int64_t>(R, 1) * T0 * 4)) return rc;
  if (pairs0)
    for (int i = 0; i < 2; ++i)
      if (int rc = grow(&h->d_hout2[3 + i], &h->hout2_bytes[3 + i], (size_t)fb * U * 4)) return rc;
  double* d_ent = (double*)h->d_hout[0];
  double* d_perk = (spatial && sp_per_k) ? (double*)h->d_hout[1] : nullptr;
  double* d_hist = (spatial && hist0) ? (double*)h->d_hout[2] : nullptr;
  double* d_tent = (double*)h->d_hout2[0];
  double* d_tperk = tr_per_k ? (double*)h->d_hout2[1] : nullptr;
  int32_t* d_pc = prev_count0 ? (int32_t*)h->d_hout2[2] : nullptr;
  cudaEvent_t* in_done = h->ev_pipe;
  cudaEvent_t* exec_done = h->ev_pipe + 2;
  cudaEvent_t* out_done = h->ev_pipe + 4;
  PipeStatus ps;
  // events left recorded by an earlier call are complete (every call drains its streams): waiting on them is a no-op
  int b = 0;
  for (int64_t f0 = 0; f0 < F && ps.ok(); f0 += fb, b ^= 1) {
    const int64_t nf = std::min(fb, F - f0);      // new frames [f0, f0 + nf), uploaded behind the halo slot
    const bool halo = f0 > 0;
    char* buf = (char*)h->d_in[b];
    VET_PIPE(ps, cudaStreamWaitEvent(h->s_copy, exec_done[b], 0));  // buffer b was last read two batches ago
    Upload widen;
    upload_frames(h, ps, packed_host, rec, esz, f0, nf, U, b, buf + frame_bytes, &widen);
    VET_PIPE(ps, cudaEventRecord(in_done[b], h->s_copy));
    VET_PIPE(ps, cudaStreamWaitEvent(h->s_exec, in_done[b], 0));
    VET_PIPE(ps, cudaStreamWaitEvent(h->s_exec, out_done[b], 0));  // per-user result buffers b must have been downloaded
    widen_records(h, ps, widen, dtype, h->s_exec);
    // frames on the device for this batch: [f0 - halo, f0 + nf) at buf + (halo ? 0 : 1 frame); the halo frame itself
    // was put into slot 0 of this buffer by the previous batch (below)
    const char* in = buf + (halo ? 0 : frame_bytes);
    const int64_t nd = nf + (halo ? 1 : 0);
    const int64_t fd = f0 - (halo ? 1 : 0);  // first frame on the device = first transition row of the batch
    uint16_t* d_asg = (spatial && assign0) ? (uint16_t*)h->d_hout[3 + b] : nullptr;
    uint16_t* d_pairs = pairs0 ? (uint16_t*)h->d_hout2[3 + b] : nullptr;
    if (ps.ok()) {
      if (spatial && nd >= 2)
        ps.call(analyze_core(h, in, dtype, nd, U, d_ent + fd, d_perk ? d_perk + fd : nullptr, F, d_hist ? d_hist + fd * T0 : nullptr,
                             d_asg, d_tent + fd, d_tperk ? d_tperk + fd : nullptr, R, d_pc ? d_pc + fd * T0 : nullptr, d_pairs,
                             mode, h->s_exec));
      else if (spatial)  // a single frame in the whole call
        ps.call(spatial_core(h, in, dtype, nd, U, d_ent + fd, d_perk ? d_perk + fd : nullptr, F, d_hist ? d_hist + fd * T0 : nullptr,
                             d_asg, h->s_exec));
      else if (nd >= 2)
        ps.call(transition_core(h, in, dtype, nd, U, d_tent + fd, d_tperk ? d_tperk + fd : nullptr, R,
                                d_pc ? d_pc + fd * T0 : nullptr, d_pairs, mode, h->s_exec));
    }
    if (f0 + nf < F)  // the last frame of this batch is the halo of the next one: slot 0 of the other buffer
      VET_PIPE(ps, cudaMemcpyAsync(h->d_in[b ^ 1], in + (size_t)(nd - 1) * frame_bytes, frame_bytes, cudaMemcpyDeviceToDevice,
                                   h->s_exec));
    VET_PIPE(ps, cudaEventRecord(exec_done[b], h->s_exec));
    if (d_asg || d_pairs) {
      VET_PIPE(ps, cudaStreamWaitEvent(h->s_out, exec_done[b], 0));
      if (d_asg)  // the halo frame's assignments went out with the previous batch
        VET_PIPE(ps, cudaMemcpyAsync(assign0 + f0 * U, d_asg + (halo ? U : 0), (size_t)nf * U * 2, cudaMemcpyDeviceToHost, h->s_out));
      if (d_pairs && nd >= 2)
        VET_PIPE(ps, cudaMemcpyAsync(pairs0 + fd * U * 2, d_pairs, (size_t)(nd - 1) * U * 4, cudaMemcpyDeviceToHost, h->s_out));
      VET_PIPE(ps, cudaEventRecord(out_done[b], h->s_out));
    }
  }
  if (spatial) {
    VET_PIPE(ps, cudaMemcpyAsync(sp_entropy, d_ent, (size_t)F * 8, cudaMemcpyDeviceToHost, h->s_exec));
    if (d_perk) VET_PIPE(ps, cudaMemcpyAsync(sp_per_k, d_perk, (size_t)F * h->K * 8, cudaMemcpyDeviceToHost, h->s_exec));
    if (d_hist) VET_PIPE(ps, cudaMemcpyAsync(hist0, d_hist, (size_t)F * T0 * 8, cudaMemcpyDeviceToHost, h->s_exec));
  }
  if (R > 0) {
    VET_PIPE(ps, cudaMemcpyAsync(tr_entropy, d_tent, (size_t)R * 8, cudaMemcpyDeviceToHost, h->s_exec));
    if (d_tperk) VET_PIPE(ps, cudaMemcpyAsync(tr_per_k, d_tperk, (size_t)R * h->K * 8, cudaMemcpyDeviceToHost, h->s_exec));
    if (d_pc) VET_PIPE(ps, cudaMemcpyAsync(prev_count0, d_pc, (size_t)R * T0 * 4, cudaMemcpyDeviceToHost, h->s_exec));
  }
  return drain_pipeline(h, ps);
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
  if (h->direct_only)
    return host_direct(h, packed_host, dtype, F, U, true, false, entropy_host, per_k_host, hist0_host, assign0_host, nullptr,
                       nullptr, nullptr, nullptr, VET_TRANSITION_LITERAL);
  const int64_t fb = host_batch_frames(h, F, U, esz);
  if (int rc = grow_inputs(h, (size_t)fb * U * 3 * esz)) return rc;
  const int rec = h->opt[VET_OPT_HOST_LAYOUT] ? 2 : 3;
  if (rec == 2)
    if (int rc = grow_inputs2(h, (size_t)fb * U * 2 * esz)) return rc;
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
  cudaEvent_t* in_done = h->ev_pipe;
  cudaEvent_t* exec_done = h->ev_pipe + 2;
  cudaEvent_t* out_done = h->ev_pipe + 4;
  PipeStatus ps;
  int b = 0;
  for (int64_t f0 = 0; f0 < F && ps.ok(); f0 += fb, b ^= 1) {
    const int64_t nf = std::min(fb, F - f0);
    VET_PIPE(ps, cudaStreamWaitEvent(h->s_copy, exec_done[b], 0));  // input buffer b was last read two batches ago
    Upload widen;
    upload_frames(h, ps, packed_host, rec, esz, f0, nf, U, b, h->d_in[b], &widen);
    VET_PIPE(ps, cudaEventRecord(in_done[b], h->s_copy));
    VET_PIPE(ps, cudaStreamWaitEvent(h->s_exec, in_done[b], 0));
    VET_PIPE(ps, cudaStreamWaitEvent(h->s_exec, out_done[b], 0));  // assignment buffer b must have been downloaded
    widen_records(h, ps, widen, dtype, h->s_exec);
    if (ps.ok())
      ps.call(spatial_core(h, h->d_in[b], dtype, nf, U, d_ent + f0, d_perk ? d_perk + f0 : nullptr, F,
                           d_hist ? d_hist + f0 * T0 : nullptr, d_assign[b], h->s_exec));
    VET_PIPE(ps, cudaEventRecord(exec_done[b], h->s_exec));
    if (assign0_host) {
      VET_PIPE(ps, cudaStreamWaitEvent(h->s_out, exec_done[b], 0));
      VET_PIPE(ps, cudaMemcpyAsync(assign0_host + f0 * U, d_assign[b], (size_t)nf * U * 2, cudaMemcpyDeviceToHost, h->s_out));
      VET_PIPE(ps, cudaEventRecord(out_done[b], h->s_out));
    }
  }
  VET_PIPE(ps, cudaMemcpyAsync(entropy_host, d_ent, (size_t)F * 8, cudaMemcpyDeviceToHost, h->s_exec));
  if (per_k_host) VET_PIPE(ps, cudaMemcpyAsync(per_k_host, d_perk, (size_t)F * h->K * 8, cudaMemcpyDeviceToHost, h->s_exec));
  if (hist0_host) VET_PIPE(ps, cudaMemcpyAsync(hist0_host, d_hist, (size_t)F * T0 * 8, cudaMemcpyDeviceToHost, h->s_exec));
  return drain_pipeline(h, ps);
}

extern "C" int vet_transition_host(vet_handle* h, const void* packed_host, int dtype, int64_t F, int64_t U,
                                   double* entropy_host, double* per_k_host, int32_t* prev_count0_host,
                                   uint16_t* pairs0_host, int mode) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_transition_host: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || F < 0 || U < 0 || (dtype != VET_F32 && dtype != VET_F64)) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (mode != VET_TRANSITION_LITERAL && mode != VET_TRANSITION_TEXTBOOK) return fail(VET_ERR_INVALID_ARG, "bad mode");
  if (F <= 1) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty vector dictionary");
  if (!packed_host || !entropy_host) return fail(VET_ERR_INVALID_ARG, "null buffer");
  if (U >= 0xFFFFFFFFll) return fail(VET_ERR_UNSUPPORTED, "too many users");
  DeviceGuard guard(h->device);
  if (h->direct_only)
    return host_direct(h, packed_host, dtype, F, U, false, true, nullptr, nullptr, nullptr, nullptr, entropy_host, per_k_host,
                       prev_count0_host, pairs0_host, mode);
  return host_pipeline_halo(h, packed_host, dtype, F, U, false, nullptr, nullptr, nullptr, nullptr, entropy_host, per_k_host,
                            prev_count0_host, pairs0_host, mode);
}

extern "C" int vet_analyze_host(vet_handle* h, const void* packed_host, int dtype, int64_t F, int64_t U,
                                double* sp_entropy_host, double* sp_per_k_host, double* hist0_host, uint16_t* assign0_host,
                                double* tr_entropy_host, double* tr_per_k_host, int32_t* prev_count0_host,
                                uint16_t* pairs0_host, int mode) {
  if (h && h->naive) return fail(VET_ERR_UNSUPPORTED, "vet_analyze_host: not available for the latitude/longitude grid tiling (the reference has no such path)");
  if (!h || F < 0 || U < 0 || (dtype != VET_F32 && dtype != VET_F64)) return fail(VET_ERR_INVALID_ARG, "bad argument");
  if (mode != VET_TRANSITION_LITERAL && mode != VET_TRANSITION_TEXTBOOK) return fail(VET_ERR_INVALID_ARG, "bad mode");
  if (F == 0) return VET_OK;
  if (U == 0) return fail(VET_ERR_INVALID_ARG, "Empty vector dictionary");
  if (!packed_host || !sp_entropy_host || (F > 1 && !tr_entropy_host)) return fail(VET_ERR_INVALID_ARG, "null buffer");
  if (U >= 0xFFFFFFFFll) return fail(VET_ERR_UNSUPPORTED, "too many users");
  DeviceGuard guard(h->device);
  h->call_frames = F;
  if (h->direct_only)
    return host_direct(h, packed_host, dtype, F, U, true, true, sp_entropy_host, sp_per_k_host, hist0_host, assign0_host,
                       tr_entropy_host, tr_per_k_host, prev_count0_host, pairs0_host, mode);
  return host_pipeline_halo(h, packed_host, dtype, F, U, true, sp_entropy_host, sp_per_k_host, hist0_host, assign0_host,
                            tr_entropy_host, tr_per_k_host, prev_count0_host, pairs0_host, mode);
}
