// cvaegan_b200 - generation, filtering and single-network inference (cvae_gan.py:339-378).
// Rows are independent in eval mode (BatchNorm running stats, no dropout, no power iteration), so the
// row stream is processed in workspace-sized chunks and shards across GPUs with no collective.
#include "engine.cuh"

namespace cvg {

#ifndef CVG_LAUNCH_CHECK
#define CVG_LAUNCH_CHECK()                 \
  do {                                     \
    CVG_CUDA(cudaGetLastError());          \
    e.launches++;                          \
  } while (0)
#endif

static int fill_z(Engine& e, const float* z, int rows, uint64_t seed, uint64_t global_row0, cudaStream_t st) {
  FillArgs f;
  f.njobs = 1;
  f.M = rows;
  f.ld = e.ws.ld;
  f.seed = seed;
  f.counter = 0;
  f.row_base = global_row0;
  f.keep_prob = 1.f;
  f.job[0].out = e.ws.z;
  f.job[0].injected = z;
  f.job[0].kind = 0;
  f.job[0].nfeat = e.Z;
  f.job[0].npass = 1;
  f.job[0].stream = RS_GEN;
  long long t = (long long)((e.Z + 3) / 4) * rows;
  int blocks = (int)((t + 255) / 256);
  if (blocks > 8 * e.num_sms) blocks = 8 * e.num_sms;
  fill_noise_kernel<<<dim3(blocks, 1), 256, 0, st>>>(f);
  CVG_LAUNCH_CHECK();
  return 0;
}

int generate(Engine& e, int label, int64_t n, const float* z, uint64_t seed, uint64_t row_offset, int train_mode,
             float* x_out, cudaStream_t st) {
  if (!e.ws_base) CVG_FAIL("workspace not bound");
  if (label < 0 || label >= e.K) CVG_FAIL("label out of range");
  if (train_mode && n > e.ws.rows_cap) CVG_FAIL("train-mode generation needs n <= max_batch (batch statistics)");
  if (train_mode && n < 2) CVG_FAIL("BatchNorm in train mode needs more than 1 row");
  if (!train_mode && e.use_tc) return tc_generate(e, label, n, z, seed, row_offset, x_out, st);
  const int cap = e.ws.rows_cap;
  for (int64_t done = 0; done < n; done += cap) {
    const int rows = (int)((n - done) < cap ? (n - done) : cap);
    CVG_TRY(fill_z(e, z ? z + done * e.Z : nullptr, rows, seed, row_offset + (uint64_t)done, st));
    if (train_mode) CVG_CUDA(cudaMemsetAsync(e.ws.acc, 0, e.ws.acc_bytes, st));
    CVG_TRY(fwd_generator(e, 1, train_mode != 0, false, label, rows, (float)rows, true, st));
    to_row_major_kernel<<<(rows + 127) / 128, 128, 0, st>>>(e.ws.g_out, rows, e.F, e.ws.ld, x_out + done * e.F);
    CVG_LAUNCH_CHECK();
  }
  return 0;
}

int generate_filter(Engine& e, int label, int64_t n, float thr, const float* z, uint64_t seed, uint64_t row_offset,
                    float* x_out, int64_t* idx_out, int64_t capacity, unsigned long long* count_out, float* logits_out,
                    uint8_t* keep_out, cudaStream_t st) {
  if (!e.ws_base) CVG_FAIL("workspace not bound");
  if (label < 0 || label >= e.K) CVG_FAIL("label out of range");
  if (e.use_tc)
    return tc_generate_filter(e, label, n, thr, z, seed, row_offset, x_out, idx_out, capacity, count_out, logits_out, keep_out, st);
  const int cap = e.ws.rows_cap;
  for (int64_t done = 0; done < n; done += cap) {
    const int rows = (int)((n - done) < cap ? (n - done) : cap);
    CVG_TRY(fill_z(e, z ? z + done * e.Z : nullptr, rows, seed, row_offset + (uint64_t)done, st));
    CVG_TRY(fwd_generator(e, 1, false, false, label, rows, (float)rows, true, st));
    CVG_TRY(fwd_classifier(e, e.ws.g_out, 0, 1, false, rows, st));
    filter_compact_kernel<true><<<(rows + 255) / 256, 256, 0, st>>>(
        e.ws.g_out, e.ws.c_logit, rows, e.ws.ld, e.F, e.K, label, thr, row_offset + (uint64_t)done, x_out,
        (long long*)idx_out, capacity, count_out, logits_out ? logits_out + done * e.K : nullptr,
        keep_out ? keep_out + done : nullptr);
    CVG_LAUNCH_CHECK();
  }
  return 0;
}

int classifier_forward(Engine& e, const float* x, int64_t n, float* logits_out, cudaStream_t st) {
  if (!e.ws_base) CVG_FAIL("workspace not bound");
  if (e.use_tc) return tc_classifier_forward(e, x, n, logits_out, st);
  const int cap = e.ws.rows_cap;
  for (int64_t done = 0; done < n; done += cap) {
    const int rows = (int)((n - done) < cap ? (n - done) : cap);
    to_feature_major_kernel<<<(rows + 127) / 128, 128, 0, st>>>(x + done * e.F, nullptr, rows, e.F, e.ws.ld, e.ws.xT);
    CVG_LAUNCH_CHECK();
    CVG_TRY(fwd_classifier(e, e.ws.xT, 0, 1, false, rows, st));
    to_row_major_kernel<<<(rows + 127) / 128, 128, 0, st>>>(e.ws.c_logit, rows, e.K, e.ws.ld, logits_out + done * e.K);
    CVG_LAUNCH_CHECK();
  }
  return 0;
}

int encoder_forward(Engine& e, const float* x, int label, int64_t n, float* mu_out, float* lv_out, cudaStream_t st) {
  if (!e.ws_base) CVG_FAIL("workspace not bound");
  if (label < 0 || label >= e.K) CVG_FAIL("label out of range");
  if (e.use_tc) return tc_encoder_forward(e, x, label, n, mu_out, lv_out, st);
  const int cap = e.ws.rows_cap;
  for (int64_t done = 0; done < n; done += cap) {
    const int rows = (int)((n - done) < cap ? (n - done) : cap);
    to_feature_major_kernel<<<(rows + 127) / 128, 128, 0, st>>>(x + done * e.F, nullptr, rows, e.F, e.ws.ld, e.ws.xT);
    CVG_LAUNCH_CHECK();
    CVG_TRY(fwd_encoder(e, false, label, rows, (float)rows, true, st));
    to_row_major_kernel<<<(rows + 127) / 128, 128, 0, st>>>(e.ws.e_ml, rows, e.Z, e.ws.ld, mu_out + done * e.Z);
    CVG_LAUNCH_CHECK();
    to_row_major_kernel<<<(rows + 127) / 128, 128, 0, st>>>(e.ws.e_ml + (size_t)e.Z * e.ws.ld, rows, e.Z, e.ws.ld,
                                                           lv_out + done * e.Z);
    CVG_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace cvg
