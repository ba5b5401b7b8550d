/*
 * cvaegan_b200 - C ABI of the B200-native (sm_100a) CVAE-GAN hot path.
 *
 * Drop-in boundary for /root/reference/src/cvae_gan.py (class CVAEGAN) and
 * /root/reference/src/models/cvae_gan_models.py.  The reference has no FFI of its own (it is pure
 * Python over torch ops); these are the entry points a Python host binds with ctypes
 * (cvae_gan_b200/_lib.py; the reference-side stub is shown in INTEGRATION.md).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types.
 *   - every function returns 0 on success, non-zero on failure; cvg_last_error() gives the message.
 *   - the library never allocates or frees tensor memory: the host owns parameters, gradients, Adam
 *     moments, BatchNorm running stats, spectral-norm u/v, the workspace and all outputs.  The
 *     library borrows raw DEVICE pointers.
 *   - all work is enqueued on the cudaStream_t passed as `stream` (a void* here) and is
 *     asynchronous; calls on one handle must be serialised by the caller.
 *   - no CPU fallback and no backend dispatch: cvg_create fails on anything but an sm_100 device.
 *   - user-facing matrices are row-major float32 [rows, features] like the reference's tensors;
 *     labels are one int per call because a reference batch carries ONE label
 *     (cvae_gan.py:109,136,166).
 */
#ifndef CVAEGAN_B200_H
#define CVAEGAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVG_ABI_VERSION 2

/* network ids (cvae_gan.py:19-39) */
enum { CVG_NET_ENCODER = 0, CVG_NET_GENERATOR = 1, CVG_NET_DISCRIMINATOR = 2, CVG_NET_CLASSIFIER = 3, CVG_NUM_NETS = 4 };

/* flags for the step functions */
enum {
  CVG_STEP_NO_UPDATE = 1,  /* compute losses + gradients, skip Adam (gradients stay in the grad buffers) */
  CVG_STEP_LOCAL_BN  = 2,  /* data parallel only: per-rank BatchNorm statistics (DEVIATES from the reference) */
  CVG_VISIT_LAMBDA_ZERO = 4, /* cvg_visit: the caller guarantees lambda_class == 0 (epochs < 200): skip the classifier backward */
  CVG_STEP_PRIOR_ONLY = 8,  /* cvg_step_g / cvg_visit: the sibling trainer CGAN's generator step (/root/reference/src/cgan.py:138-178):
                             * x_fake = G(z_prior) only, no encoder / reconstruction / KL, Adam on the generator alone; x_real is
                             * not read (may be null) and a visit draws no batch for these steps; loss_out = {0, 0, adv, class} */
  CVG_STEP_CVAE = 16        /* cvg_step_g / cvg_visit: the sibling trainer CVAE's encoder/generator step (/root/reference/src/cvae.py:117-166):
                             * x_recon = G(E(x)) only, no critic and no z_prior pass; the classification term is taken on x_recon;
                             * Adam on encoder + generator; loss_out = {recon, kl, 0, class}.  A CVAE label visit is
                             * cvg_visit(d_loop = 0, c_loop, g_loop, flags | CVG_STEP_CVAE) (cvae.py:86-166).  Mutually exclusive
                             * with CVG_STEP_PRIOR_ONLY. */
};

/* Mirrors /root/reference/src/config/gan_config.py:1-21 plus the torch defaults the models rely on. */
typedef struct CvgConfig {
  int32_t feature_num;   /* datasets.feature_num (cvae_gan.py:15) */
  int32_t label_num;     /* datasets.label_num   (cvae_gan.py:16) */
  int32_t z_size;        /* gan_config.z_size = 128 */
  int32_t max_batch;     /* largest per-rank batch the workspace is sized for */
  int32_t world_size;    /* data-parallel ranks (1 = single GPU) */
  int32_t rank;
  float lambda_recon, lambda_kl, lambda_adv; /* cvae_gan_config; lambda_class is passed per step */
  float g_lr, d_lr, c_lr;                    /* 2e-4, 2e-4, 1e-4 */
  float adam_beta1, adam_beta2, adam_eps;    /* 0.5, 0.999, 1e-8 (cvae_gan.py:75-97) */
  float bn_momentum, bn_eps, ln_eps, sn_eps; /* 0.1, 1e-5, 1e-5, 1e-12 */
  float lrelu_slope, dropout_p;              /* 0.2, 0.3 */
  int32_t hidden[3];     /* {0,0,0}: the reference's widths (cvae_gan_models.py:16-18,85-87,173-175,257-259: max(256,in) /
                          * max(128,in/2) / max(64,in/4) | 64).  Otherwise the three hidden widths of ALL four networks - the
                          * widened model of BASELINE.json configs[4] (1024,512,256), which the reference's hard-coded widths
                          * cannot express; each a multiple of 64, <= 1024 (classifier LayerNorm width hidden[1] <= 512). */
  int32_t unconditional; /* 0: encoder, generator and critic take the one-hot label beside their input (cvae_gan_models.py:47-58,
                          * 136-150, 215-233).  1: the sibling trainer VAE-GAN's networks (/root/reference/src/models/
                          * vae_gan_models.py:8-154): the same layers WITHOUT the label columns - first Linear of E and D over F
                          * inputs, of G over z_size; `label` arguments of the calls are then ignored (pass 0). */
} CvgConfig;

typedef struct CvgHandle CvgHandle;

/* One entry per tensor of a network, named with the reference's state_dict key. */
typedef struct CvgTensorDesc {
  char key[96];       /* e.g. "main_model.0.weight", "discriminator_network.3.parametrizations.weight.original" */
  int32_t kind;       /* 0 = parameter (offset into the params/grads/adam buffers), 1 = float state buffer */
  int32_t ndim;
  int64_t shape[2];
  int64_t offset;     /* in floats, 16-byte aligned */
} CvgTensorDesc;

/* Optional injected randomness (parity runs).  Any pointer may be NULL -> in-kernel Philox is used
 * for that tensor.  All are DEVICE pointers, row-major, shapes in reference layout:
 *   z        [B, Z]  prior noise (cvae_gan.py:114,140,173)
 *   eps      [B, Z]  reparameterisation noise (cvae_gan_models.py:68); step_g only
 *   d_mask1  [P, B, H1], d_mask2 [P, B, H2]  critic dropout KEEP masks (uint8 0/1), P = 2 for step_d
 *            (real pass, then fake pass), P = 1 for step_g
 *   c_mask1, c_mask2  same for the classifier (P = 2 for step_c, 1 for step_g)                     */
typedef struct CvgNoise {
  const float* z;
  const float* eps;
  const uint8_t* d_mask1;
  const uint8_t* d_mask2;
  const uint8_t* c_mask1;
  const uint8_t* c_mask2;
} CvgNoise;

const char* cvg_last_error(void);
int cvg_abi_version(void);
int cvg_config_bytes(void);   /* sizeof(CvgConfig) as this library was compiled: a binding checks its struct against it */

/* cvae_gan.py:12-56 (CVAEGAN.__init__): sizes the four networks.  Fails unless the current CUDA device is sm_100. */
int cvg_create(const CvgConfig* cfg, CvgHandle** out);
void cvg_destroy(CvgHandle* h);

/* Layout queries: floats needed for a net's parameter buffer (params / grads / adam m / adam v all
 * have this size; the grad buffer needs CVG_GRAD_TAIL extra floats) and for its float state buffer
 * (BatchNorm running_mean/var for E and G; spectral-norm _u/_v for D). */
#define CVG_GRAD_TAIL 16
int cvg_net_sizes(const CvgHandle* h, int net, int64_t* n_param_floats, int64_t* n_state_floats);
int cvg_tensor_table(const CvgHandle* h, int net, CvgTensorDesc* out, int32_t capacity, int32_t* count);
int64_t cvg_workspace_bytes(const CvgHandle* h);

/* Borrow the host-owned device buffers. grads has n_param_floats + CVG_GRAD_TAIL floats. */
int cvg_bind_net(CvgHandle* h, int net, float* params, float* grads, float* adam_m, float* adam_v, float* state);
int cvg_bind_workspace(CvgHandle* h, void* workspace, int64_t bytes, void* stream);

/* Adam step counters (torch.optim.Adam state['step']); one per network. */
int cvg_set_adam_step(CvgHandle* h, int net, int64_t t);
int64_t cvg_get_adam_step(const CvgHandle* h, int net);

/* Data parallel (one process per GPU).  Rank 0 calls cvg_comm_unique_id, the host broadcasts the 128
 * bytes (torch.distributed), every rank calls cvg_comm_init.  The library then all-reduces BatchNorm
 * batch moments between layers and the flat gradient buffers before Adam with NCCL on `stream`. */
int cvg_comm_unique_id(void* out128);
int cvg_comm_init(CvgHandle* h, const void* id128, int rank, int world_size);

/* Optional, after cvg_comm_init on a single node with NVLink peer access: the latency-bound exchanges (BatchNorm
 * moments, flat gradients) then use a one-shot all-reduce over peer memory instead of NCCL.  Every rank calls
 * cvg_nvl_local_handle (creates its staging buffer, returns a 64-byte CUDA IPC handle), the host all-gathers the
 * handles in rank order (world_size x 64 bytes), every rank calls cvg_nvl_attach, then the host runs a barrier. */
int cvg_nvl_local_handle(CvgHandle* h, void* out64);
int cvg_nvl_attach(CvgHandle* h, const void* handles);
/* Back to NCCL for the exchanges (all ranks together, e.g. when one rank could not map a peer). */
int cvg_nvl_disable(CvgHandle* h);

/* The three optimiser steps of one label visit.
 *   x_real   [B, F] row-major device pointer: the batch _get_target_samples returned (cvae_gan.py:108)
 *   label    the visit's target label
 *   noise    NULL or injected randomness (see CvgNoise)
 *   seed / counter   Philox key and per-step counter for everything not injected; rows are keyed by
 *            GLOBAL row index (rank * B + row) so results do not depend on the GPU count
 *   loss_out device float[4]:
 *            step_d -> {d_loss, mean D(real), mean D(fake), 0}            (cvae_gan.py:119-126)
 *            step_c -> {c_loss, CE real, CE fake, 0}                      (cvae_gan.py:147-154)
 *            step_g -> {recon, kl, adv, class}                            (cvae_gan.py:184-194)      */
int cvg_step_d(CvgHandle* h, const float* x_real, int label, int B, const CvgNoise* noise, uint64_t seed,
               uint64_t counter, int flags, float* loss_out, void* stream);
int cvg_step_c(CvgHandle* h, const float* x_real, int label, int B, const CvgNoise* noise, uint64_t seed,
               uint64_t counter, int flags, float* loss_out, void* stream);
int cvg_step_g(CvgHandle* h, const float* x_real, int label, int B, const CvgNoise* noise, uint64_t seed,
               uint64_t counter, float lambda_class_now, int flags, float* loss_out, void* stream);

/* One optimiser step of the DOWNSTREAM classifier fine-tuning (/root/reference/src/classifier.py:24-45, used by
 * scripts/train_cvae_gan.py:143-165 on the augmented set): logits = C(x) in train mode, loss = mean
 * cross_entropy(logits, labels) with per-row labels (device int64[B], values in [0, label_num)), backward, Adam with the
 * caller's hyper-parameters (classifier_config.lr = 1e-3, torch defaults 0.9 / 0.999 / 1e-8) on the classifier's
 * Adam state (reset it with cvg_set_adam_step + zeroed moment buffers for a fresh optimiser).  Only c_mask1 / c_mask2
 * ([1, B, H]) of `noise` are used.  loss_out: device float[4] = {loss, loss, 0, 0}. */
int cvg_step_classifier(CvgHandle* h, const float* x, const int64_t* labels, int B, const CvgNoise* noise, uint64_t seed,
                        uint64_t counter, float lr, float beta1, float beta2, float eps, int flags, float* loss_out,
                        void* stream);

/* One label visit of the training loop (cvae_gan.py:102-216): d_loop critic steps, c_loop classifier steps and
 * g_loop encoder/generator steps, each on a freshly drawn batch - either drawn on the device from class_rows
 * [n_rows, F] (_get_target_samples) or taken from x_batches [d_loop+c_loop+g_loop, B_local, F] when that is
 * non-NULL.  Philox seed/counter, the Adam step counts and lambda_class are read from a device-resident control
 * block (cvg_ctl_set), nothing dynamic is passed by value: the call can be captured in a CUDA graph (stream
 * capture on `stream`) once per (label, lambda_class != 0) and replayed.  loss_out: device float
 * [d_loop+c_loop+g_loop][4], one row per step in the layout of the step functions. */
int cvg_visit(CvgHandle* h, int label, int B_local, int64_t B_global, const float* class_rows, int64_t n_rows,
              const float* x_batches, int d_loop, int c_loop, int g_loop, int flags, float* loss_out, void* stream);

/* Writes the control block on `stream`: Philox key + step counter (if set_rng) and the epoch's lambda_class
 * (if set_lambda).  The step functions below do this themselves from their arguments. */
int cvg_ctl_set(CvgHandle* h, uint64_t seed, uint64_t counter, int set_rng, float lambda_class, int set_lambda, void* stream);

/* torch.optim.Adam.step for the networks in net_mask (bit i = net i), using the bound grad buffers.
 * The step functions call this themselves unless CVG_STEP_NO_UPDATE is set. */
int cvg_adam(CvgHandle* h, int net_mask, void* stream);

/* _get_target_samples (cvae_gan.py:247-260) on the device: draws row indices of a class of n rows for
 * a GLOBAL batch of B_global rows (n < B_global: with replacement; n == B_global: identity;
 * n > B_global: distinct rows via a keyed Feistel permutation of [0, n)) and gathers draws
 * [draw_offset, draw_offset + B_local) into x_out [B_local, F].  A data-parallel rank passes
 * draw_offset = rank * B_local, so the union over ranks is the single-GPU draw.
 * idx_out (device int64[B_local]) may be NULL. */
int cvg_sample_rows(CvgHandle* h, const float* class_rows, int64_t n, int64_t B_global, int64_t draw_offset,
                    int B_local, uint64_t seed, uint64_t counter, float* x_out, int64_t* idx_out, void* stream);

/* generate_samples (cvae_gan.py:339-345): x_out[n, F] = G(z, onehot(label)) with G in eval mode
 * (train_mode = 0: BatchNorm running stats) or train mode (batch stats over these n rows, running
 * stats updated - what the reference does if called before fit()).  z is NULL (Philox keyed by
 * row_offset + row) or an injected [n, Z] device tensor. */
int cvg_generate(CvgHandle* h, int label, int64_t n, const float* z, uint64_t seed, uint64_t row_offset,
                 int train_mode, float* x_out, void* stream);

/* Fused generate -> classify -> threshold -> compact (cvae_gan.py:357-371 without the chunk-of-10
 * loop): for rows r in [0, n) of the noise stream (global row = row_offset + r) computes
 * x = G_eval(z_r), logits = C_eval(x), keeps rows with max softmax > thr AND argmax == label, and
 * appends them to x_out (capacity rows) with their global row index in idx_out.  *count_out (device
 * uint64, NOT reset by this call) is advanced by the number of accepted rows; rows beyond capacity
 * are counted but not written.  Output order is unspecified; sort by idx_out for stream order.
 * logits_out ([n, K]) and keep_out (uint8[n]) may be NULL. */
int cvg_generate_filter(CvgHandle* h, int label, int64_t n, float thr, const float* z, uint64_t seed,
                        uint64_t row_offset, float* x_out, int64_t* idx_out, int64_t capacity,
                        unsigned long long* count_out, float* logits_out, uint8_t* keep_out, void* stream);

/* The filter decision alone (cvae_gan.py:366-370) on caller-supplied logits [n, K]:
 * keep[i] = max softmax(logits[i]) > thr && argmax == label (first maximal index on ties). */
int cvg_filter_logits(const float* logits, int64_t n, int K, int label, float thr, uint8_t* keep_out, void* stream);

/* Standalone memory-bound filter over materialised tensors: keep decision from logits [n, K] and
 * compaction of the matching rows of x [n, F] (+ their indices) - the kernel measured against the
 * HBM roofline (4F + 4K + a(4F+8) bytes per row). */
int cvg_filter_compact(const float* x, const float* logits, int64_t n, int F, int K, int label, float thr,
                       uint64_t row_offset, float* x_out, int64_t* idx_out, int64_t capacity,
                       unsigned long long* count_out, void* stream);

/* Forward passes of single networks on caller data (row-major in/out), used by the nn.Module-like
 * wrappers (`gan.classifier(x)`, cvae_gan.py:362; classifier.py:37,57) and by the parity tests.
 *   cvg_classifier_forward: logits[n, K] = C(x) in eval mode.
 *   cvg_encoder_forward:    mu[n, Z], logvar[n, Z] = E(x, label) in eval mode. */
int cvg_classifier_forward(CvgHandle* h, const float* x, int64_t n, float* logits_out, void* stream);
int cvg_encoder_forward(CvgHandle* h, const float* x, int label, int64_t n, float* mu_out, float* logvar_out, void* stream);

/* Host-only integer logic of the reference's chunk-of-10 / patience-20 loop (cvae_gan.py:350-376):
 * given the keep mask of the row stream, how many rows the loop consumes and accepts. */
int cvg_patience_scan(const uint8_t* keep_host, int64_t n, int64_t num, int chunk, int patience,
                      int64_t* rows_consumed, int64_t* rows_accepted);

/* Test hook: copies one feature-major workspace matrix (after a step) to dst as row-major [rows, features].
 * name: "xT", "z", "g_h0".."g_h2", "g_out", "g_dy0".."g_dy2", "g_dout", "e_h0".."e_h2", "e_ml", "e_dy0".."e_dy2",
 * "e_dml", "d_a0".."d_a2", "d_s", "d_g0".."d_g2", "c_a1", "c_h2", "c_a2", "c_a3", "c_logit", "c_dlogit",
 * "c_g0".."c_g2", "dx".  *features_out receives the feature count.  dst may be NULL to query it. */
int cvg_debug_read(CvgHandle* h, const char* name, int pass, int rows, float* dst, int* features_out, void* stream);

/* Development hook: when dev_counters (device int64[32]) is non-NULL the tensor-core chain kernel adds its per-role
 * cycle counts to it ([0] issuer waiting for activations, [1] for weights, [2] for free stages, [3] issuer total,
 * [4] epilogue waiting for accumulators, [5] input generation, [6] epilogue total, [7] MMA issue, [8..15] issue per layer, [16] TMEM loads (64-row kernel) or plane_free waits (128-row kernel), [17] fences + arrive, [18..25] epilogue per layer).  NULL switches it off. */
int cvg_debug_tc_counters(CvgHandle* h, long long* dev_counters);

/* Measurement hook for bench.py: when enabled, every GEMM launch is bracketed by CUDA events on its
 * stream.  cvg_profile_read synchronises and returns, per kernel class (0 = forward GEMM, 1 = input-
 * gradient GEMM, 2 = weight-gradient GEMM), the launches, the ALGORITHMIC flops (2*M*N*K per launch)
 * and the summed device time in ms since the last enable.  Not used while `value` is being timed. */
int cvg_profile_enable(CvgHandle* h, int enable);
int cvg_profile_read(CvgHandle* h, int kernel_class, int64_t* launches, double* flops, double* ms);

/* Number of kernels this handle has launched since creation (bench.py reports it as gpu_launches). */
int64_t cvg_launch_count(const CvgHandle* h);

/* Training executor switches (tests, A/B runs; the defaults come from the environment at cvg_create).
 *   "train_mode"  0 = stand-alone FP32-FMA layer kernels (default),
 *                 1 = step-program kernel: one persistent tcgen05 kernel per optimiser step / label visit (CVG_TRAIN_MODE=mk)
 *   "mk_max_ops"  truncate every recorded program after this many ops (-1 = off; bisecting)
 *   "mk_allbar"   grid barrier before every op       "mk_coop"  cooperative launch on / off
 *   "hoist"       1 (default; CVG_HOIST): a label visit runs the generator forward of all its critic / classifier steps once, up front
 *   "streams"     1 (default; CVG_STREAMS): independent kernels of a step run on two event-joined side streams
 *   "fuse_stats"  1 (default; CVG_FUSE_STATS): data parallel BatchNorm sums are pushed / polled inside the GEMM kernels
 * cvg_debug_get: "train_mode", "mk_supported", "mk_last_nops" (ops of the last program incl. the finish op).
 * cvg_debug_mk_cycles: per-op cycle counts of the last program as seen by CTA 0 (needs CVG_MK_DBG=1). */
int cvg_debug_set(CvgHandle* h, const char* key, int value);
int cvg_debug_get(const CvgHandle* h, const char* key, int* value);
int cvg_debug_mk_cycles(CvgHandle* h, long long* out, int capacity, int* count);

#ifdef __cplusplus
}
#endif
#endif /* CVAEGAN_B200_H */
