// cvaegan_b200 - fused eval-mode MLP chains on the sm_100a tensor cores (tcgen05.mma kind::tf32, 3xTF32 split).
//
//   generate -> classify -> threshold -> compact in ONE pass  (cvae_gan.py:339-378, cvae_gan_models.py:89-160, 261-292)
//   generate only, classifier only, encoder only              (cvae_gan.py:339-345, 362; classifier.py:37,57)
//
// Rows are independent in eval mode (BatchNorm running stats, no dropout, no power iteration), so a persistent CTA
// takes 64-row tiles of the row stream through the whole chain without touching HBM in between:
//
//   * orientation: D[feature][row] = W[feature][k] * act[row][k]  -> MMA M = output features (128 or 64 per tile),
//     N = 64 batch rows, K = input features.  Both operands are K-major, no-swizzle core-matrix layouts (tc05.cuh):
//       weights  (A): pre-split hi/lo chunks of 16 k, written once per call by tc_prep_kernel in exactly the
//                     shared-memory order, streamed by 1-D bulk copies (TMA engine) through a 4-stage mbarrier ring;
//       activations (B): written by the epilogue threads (thread = feature) as hi/lo planes, row pitch padded so the
//                     4-byte scattered stores are bank-conflict free.
//   * accumulators: fp32 in TMEM (feature = lane, row = column); the epilogue reads them with tcgen05.ld, applies
//     bias / BatchNorm(eval) / LayerNorm / activation and writes the next layer's operand in place.
//   * warp roles: warps 0-7 = epilogue (two warps per 32-lane TMEM quadrant, 32 of the 64 rows each), warp 8 = MMA
//     issuer, warp 9 = weight producer (bulk copies).
//   * z comes from Philox keyed by the GLOBAL row index (same stream as fill_noise_kernel), the filter decision is
//     filter_decide() (bit-exact torch softmax semantics), accepted rows are compacted with one atomic per tile.
//
// Three tile variants share the layer tables, the prep kernel and the host code below (tc_mode()):
//   tc_eval_kernel      (this file)       64-row tiles, planes hold a whole 256-wide layer; any supported shape
//   tc_eval128_kernel   (eval_tc128.cuh)  128-row tiles, extra accumulators in TMEM, split-K hand-off; the default
//   tc_eval_pp_kernel   (eval_tcpp.cuh)   two 64-row tiles in flight (experimental, CVG_TC_MODE=pp)
#include "engine.cuh"
#include "tc05.cuh"

namespace cvg {

#ifndef CVG_LAUNCH_CHECK
#define CVG_LAUNCH_CHECK()                 \
  do {                                     \
    CVG_CUDA(cudaGetLastError());          \
    e.launches++;                          \
  } while (0)
#endif

using namespace tc;

#ifndef CVG_TC_DEFAULT_MODE
#define CVG_TC_DEFAULT_MODE 1
#endif
constexpr int TC_MAX_LAYERS = 8;
constexpr int TC_ROWS = 64;                          // batch rows per tile (MMA N)
constexpr int TC_KC = 16;                            // contraction values per weight chunk
constexpr int TC_STAGES = 4;
constexpr int TC_MAXK = 256;                         // widest layer input
constexpr int TC_LBO_B = TC_ROWS * 16 + 16;          // bytes between 4-k groups of the activation operand (padded)
constexpr int TC_BBYTES = (TC_MAXK / 4) * TC_LBO_B;  // one plane (hi or lo)
constexpr int TC_STAGE_BYTES = 2 * TC_KC * 128 * 4;  // hi + lo chunk of a 128-row weight tile
constexpr int TC_MAXF = 64;                          // widest generator output kept for compaction
constexpr int TC_EPI_WARPS = 8;                       // two warps per TMEM lane quadrant, 32 of the 64 rows each
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
constexpr int TC_THREADS = TC_EPI_THREADS + 64;      // + MMA issuer warp + weight producer warp
constexpr int TC_LN_PITCH = TC_ROWS + 1;
constexpr int TC_XS_PITCH = TC_ROWS + 1;            // generator outputs [F][pitch] kept for the compaction

enum { TEPI_BN_LRELU = 0, TEPI_RELU = 1, TEPI_LN_RELU = 2, TEPI_SIGMOID_X = 3, TEPI_LOGITS = 4, TEPI_OUT = 5 };

struct TcLayer {
  int K;           // contraction length, multiple of 8 (zero padded)
  int N;           // real output features
  int M;           // MMA M: 128 or 64
  int n_mtiles, n_kchunks;
  int epi;
  int Npad;        // const-table pitch
  int c_off;       // floats into the const table: [4][Npad]
  long long w_off; // floats into the prepped weight buffer
};

struct TcPrepLayer {
  const float* W; int ldw; int Kreal;
  const float* bias; const float* wlabel; int ldwl;       // bias' = bias + wlabel[n * ldwl]
  const float* gamma; const float* beta; const float* rmean; const float* rvar;
};

struct TcEvalArgs {
  TcLayer L[TC_MAX_LAYERS];
  int nl;
  const float* wprep;
  const float* consts;
  int in_kind;          // 0 = Philox z, 1 = injected z [n][in_feat], 2 = x rows [n][in_feat]
  const float* in;
  int in_feat;
  long long n;
  unsigned long long seed, row_offset;
  float* x_all;         // [n][F] generator output of every row, or null
  float* out_plain;     // TEPI_OUT: features [0, out_split) -> out_plain[n][out_ld], the rest -> out_plain2[n][out_ld]
  float* out_plain2;
  int out_ld, out_split;
  int do_filter, F, Kc, label;
  float thr;
  float* x_out; long long* idx_out; long long capacity; unsigned long long* count;
  float* logits_out; uint8_t* keep_out;
  float slope, ln_eps;
  long long* dbg;       // optional cycle counters (development)
};

struct TcPrepArgs {
  TcLayer L[TC_MAX_LAYERS];
  TcPrepLayer P[TC_MAX_LAYERS];
  int nl;
  float* wprep;
  float* consts;
  float bn_eps;
};

// ------------------------------------------------------------------------------------------------
// weight / constant preparation: one launch per call, grid (blocks, nl)
//   chunk (mtile, kc) of a layer = [hi: kc_len/4 x M x 4 floats][lo: same], element (row, k) of the chunk at
//   (k / 4) * (M * 4) + row * 4 + k % 4  floats  (K-major core matrices: SBO = 128 B, LBO = M * 16 B)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tc_prep_kernel(const TcPrepArgs a) {
  const TcLayer L = a.L[blockIdx.y];
  const TcPrepLayer P = a.P[blockIdx.y];
  const long long total = (long long)L.n_mtiles * L.M * L.K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // enumerate in destination order within a plane so that stores coalesce
    const int per_mt = L.M * L.K;
    const int mt = (int)(i / per_mt);
    int r = (int)(i % per_mt);
    const int kc = r / (TC_KC * L.M);
    const int kc_len = min(TC_KC, L.K - kc * TC_KC);
    r -= kc * TC_KC * L.M;
    const int k4 = r / (L.M * 4), row = (r / 4) % L.M, kk = r % 4;
    const int k = kc * TC_KC + k4 * 4 + kk, n = mt * L.M + row;
    float w = 0.f;
    if (n < L.N && k < P.Kreal) w = P.W[(size_t)n * P.ldw + k];
    float hi, lo;
    split_tf32(w, hi, lo);
    float* chunk = a.wprep + L.w_off + (size_t)mt * per_mt * 2 + (size_t)kc * TC_KC * L.M * 2;
    chunk[r] = hi;
    chunk[(size_t)kc_len * L.M + r] = lo;
  }
  if (blockIdx.x == 0) {
    float* c = a.consts + L.c_off;
    for (int n = threadIdx.x; n < L.Npad; n += blockDim.x) {
      float c0 = 0.f, c1 = 0.f, c2 = 1.f, c3 = 0.f;
      if (n < L.N) {
        c0 = P.bias ? P.bias[n] : 0.f;
        if (P.wlabel) c0 += P.wlabel[(size_t)n * P.ldwl];
        if (L.epi == TEPI_BN_LRELU) {
          c1 = P.rmean[n];
          c2 = P.gamma[n] * (1.0f / sqrtf(P.rvar[n] + a.bn_eps));
          c3 = P.beta[n];
        } else if (L.epi == TEPI_LN_RELU) {
          c2 = P.gamma[n];
          c3 = P.beta[n];
        }
      }
      c[n] = c0; c[L.Npad + n] = c1; c[2 * L.Npad + n] = c2; c[3 * L.Npad + n] = c3;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the fused chain
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

struct TcSmem {
  uint64_t full[TC_STAGES], empty[TC_STAGES], acc_full, act_ready;
  uint32_t tmem_slot;
  int warp_cnt[2];
  unsigned long long base;
};

__device__ __forceinline__ void b_store(uint8_t* b_hi, uint8_t* b_lo, uint32_t off, float y) {
  float hi, lo;
  split_tf32(y, hi, lo);
  *reinterpret_cast<float*>(b_hi + off) = hi;
  *reinterpret_cast<float*>(b_lo + off) = lo;
}

#define TC_CLK(var) \
  if (prof) var = clock64();

__global__ void __launch_bounds__(TC_THREADS, 1) tc_eval_kernel(const __grid_constant__ TcEvalArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* b_hi = smem;
  uint8_t* b_lo = smem + TC_BBYTES;
  uint8_t* stages = smem + 2 * TC_BBYTES;
  float* xs = reinterpret_cast<float*>(stages + TC_STAGES * TC_STAGE_BYTES);   // [TC_MAXF][64] generator outputs
  float* lg = xs + TC_MAXF * TC_XS_PITCH;                                         // [32][64] logits
  float* red = lg + FILTER_MAXK * TC_ROWS;                                      // [10][64] LayerNorm scratch
  TcSmem* S = reinterpret_cast<TcSmem*>(red + 10 * TC_ROWS);
  float* ln_scratch = reinterpret_cast<float*>(b_hi + (TC_MAXK / 8) * TC_LBO_B);  // upper half of the hi plane

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const long long ntiles = (a.n + TC_ROWS - 1) / TC_ROWS;
  const bool prof = a.dbg != nullptr;

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&S->full[s], 1); mbar_init(&S->empty[s], 1); }
    mbar_init(&S->acc_full, 1);
    mbar_init(&S->act_ready, TC_EPI_THREADS);
    fence_mbar_init();
  }
  if (warp == TC_EPI_WARPS) tmem_alloc(&S->tmem_slot, 128);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, S->tmem_slot, 0);   // warp-uniform for the compiler

  if (warp == TC_EPI_WARPS) {
    // ===================== MMA issuer =====================
    long long t_act = 0, t_full = 0, t_empty = 0, t_all = 0, c0 = 0, c1 = 0, t_issue = 0, t_lay[TC_MAX_LAYERS] = {0, 0, 0, 0, 0, 0, 0, 0};
    TC_CLK(t_all);
    long long my_tiles = 0;
    if ((long long)blockIdx.x < ntiles) my_tiles = (ntiles - 1 - blockIdx.x) / gridDim.x + 1;
    unsigned long long g = 0;
    unsigned long long n_act = 0;
    bool peeked = false;
    const uint32_t b_hi_a = smem_u32(b_hi), b_lo_a = smem_u32(b_lo), st_a = smem_u32(stages);
    for (long long t = 0; t < my_tiles; ++t) {
      for (int l = 0; l < a.nl; ++l) {
        const TcLayer& Lr = a.L[l];
        TC_CLK(c0);
        mbar_wait(&S->act_ready, (uint32_t)(n_act & 1));
        TC_CLK(c1);
        t_act += c1 - c0;
        ++n_act;
        tc_fence_after_sync();
        const uint32_t idesc = idesc_tf32(Lr.M, TC_ROWS, 0, 0);
        const uint32_t a_lbo = (uint32_t)Lr.M * 16u;
        for (int mt = 0; mt < Lr.n_mtiles; ++mt) {
          for (int kc = 0; kc < Lr.n_kchunks; ++kc) {
            const int s = (int)(g % TC_STAGES);
            TC_CLK(c0);
            if (!peeked) mbar_wait(&S->full[s], (uint32_t)((g / TC_STAGES) & 1));
            TC_CLK(c1);
            t_full += c1 - c0;
            tc_fence_after_sync();
            // probe the NEXT chunk's barrier now: the probe's latency hides behind the MMA issue below
            const bool peek_next = mbar_test_wait(&S->full[(g + 1) % TC_STAGES], (uint32_t)(((g + 1) / TC_STAGES) & 1));
            TC_CLK(c0);
            if (elect_one()) {
              const int kc_len = min(TC_KC, Lr.K - kc * TC_KC);
              const uint32_t a_hi = st_a + (uint32_t)s * TC_STAGE_BYTES;
              const uint32_t a_lo = a_hi + (uint32_t)kc_len * Lr.M * 4u;
              const uint32_t boff = (uint32_t)(kc * TC_KC / 4) * TC_LBO_B;
              uint64_t dah = smem_desc(a_hi, a_lbo, 128), dal = smem_desc(a_lo, a_lbo, 128);
              uint64_t dbh = smem_desc(b_hi_a + boff, TC_LBO_B, 128), dbl = smem_desc(b_lo_a + boff, TC_LBO_B, 128);
              const uint32_t d = tmem + (uint32_t)mt * TC_ROWS;
              const uint64_t a_step = (uint64_t)((2 * a_lbo) >> 4), b_step = (uint64_t)((2 * TC_LBO_B) >> 4);
              for (int ks = 0; ks < kc_len / 8; ++ks) {
                mma_tf32(d, dal, dbh, idesc, !(kc == 0 && ks == 0));   // small terms first
                mma_tf32(d, dah, dbl, idesc, true);
                mma_tf32(d, dah, dbh, idesc, true);
                dah += a_step; dal += a_step; dbh += b_step; dbl += b_step;
              }
              mma_commit(&S->empty[s]);     // frees the weight stage for the producer warp
              if (mt == Lr.n_mtiles - 1 && kc == Lr.n_kchunks - 1) mma_commit(&S->acc_full);
            }
            __syncwarp();
            TC_CLK(c1);
            t_issue += c1 - c0;
            t_lay[l] += c1 - c0;
            peeked = peek_next;
            ++g;
          }
        }
      }
    }
    if (prof && lane == 0) {
      t_all = clock64() - t_all;
      atomicAdd((unsigned long long*)a.dbg + 0, (unsigned long long)t_act);
      atomicAdd((unsigned long long*)a.dbg + 1, (unsigned long long)t_full);
      atomicAdd((unsigned long long*)a.dbg + 2, (unsigned long long)t_empty);
      atomicAdd((unsigned long long*)a.dbg + 3, (unsigned long long)t_all);
      atomicAdd((unsigned long long*)a.dbg + 7, (unsigned long long)t_issue);
      for (int l = 0; l < a.nl; ++l) atomicAdd((unsigned long long*)a.dbg + 8 + l, (unsigned long long)t_lay[l]);
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ===================== weight producer: the chunk sequence of a tile is static, it just runs ahead =====================
    int cpt = 0;
    for (int l = 0; l < a.nl; ++l) cpt += a.L[l].n_mtiles * a.L[l].n_kchunks;
    long long my_tiles = 0;
    if ((long long)blockIdx.x < ntiles) my_tiles = (ntiles - 1 - blockIdx.x) / gridDim.x + 1;
    const unsigned long long total = (unsigned long long)my_tiles * cpt;
    int ll = 0, lmt = 0, lkc = 0;
    for (unsigned long long gl = 0; gl < total; ++gl) {
      const int s = (int)(gl % TC_STAGES);
      if (gl >= TC_STAGES) mbar_wait(&S->empty[s], (uint32_t)(((gl / TC_STAGES) - 1) & 1));
      const TcLayer& Lr = a.L[ll];
      const int kc_len = min(TC_KC, Lr.K - lkc * TC_KC);
      const uint32_t bytes = 2u * kc_len * Lr.M * 4u;
      const float* src = a.wprep + Lr.w_off + (size_t)lmt * Lr.K * Lr.M * 2 + (size_t)lkc * TC_KC * Lr.M * 2;
      if (elect_one()) {
        mbar_arrive_expect_tx(&S->full[s], bytes);
        bulk_g2s(stages + (size_t)s * TC_STAGE_BYTES, src, bytes, &S->full[s]);
      }
      __syncwarp();
      if (++lkc == Lr.n_kchunks) { lkc = 0; if (++lmt == Lr.n_mtiles) { lmt = 0; if (++ll == a.nl) ll = 0; } }
    }
  } else {
    // ============== epilogue warps: thread = (output feature = TMEM lane, half of the 64 rows) ==============
    long long t_acc = 0, t_in = 0, t_all = 0, c0 = 0, c1 = 0, t_ld = 0, t_fence = 0, t_epi_l[TC_MAX_LAYERS] = {0, 0, 0, 0, 0, 0, 0, 0};
    TC_CLK(t_all);
    unsigned long long n_acc = 0;
    const int in_groups = a.L[0].K / 4;
    const int q = warp & 3, h = warp >> 2;      // TMEM lane quadrant, column half
    const int mbase = h * 32;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long row0 = tile * TC_ROWS;
      const int nrows = (int)min((long long)TC_ROWS, a.n - row0);
      TC_CLK(c0);
      // ---- layer-0 operand: z (Philox or injected) or x rows, 4 consecutive features of one row per item ----
      for (int i = tid; i < TC_ROWS * in_groups; i += TC_EPI_THREADS) {
        const int m = i % TC_ROWS, fg = i / TC_ROWS;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < nrows) {
          if (a.in_kind == 0) {
            const U4 r = philox_at(a.seed, 0, RS_GEN, 0, a.row_offset + (uint64_t)(row0 + m), (uint32_t)fg);
            box_muller(r.x, r.y, v[0], v[1]);
            box_muller(r.z, r.w, v[2], v[3]);
          } else {
            const float* src = a.in + (size_t)(row0 + m) * a.in_feat + fg * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (fg * 4 + j < a.in_feat) v[j] = src[j];
          }
        }
        float4 hi, lo;
        split_tf32(v[0], hi.x, lo.x); split_tf32(v[1], hi.y, lo.y); split_tf32(v[2], hi.z, lo.z); split_tf32(v[3], hi.w, lo.w);
        const uint32_t off = (uint32_t)fg * TC_LBO_B + (uint32_t)m * 16;
        *reinterpret_cast<float4*>(b_hi + off) = hi;
        *reinterpret_cast<float4*>(b_lo + off) = lo;
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&S->act_ready);
      TC_CLK(c1);
      t_in += c1 - c0;

      for (int l = 0; l < a.nl; ++l) {
        const TcLayer& Lr = a.L[l];
        const int next_K = (l + 1 < a.nl) ? a.L[l + 1].K : 0;
        const int f_local = (Lr.M == 128) ? (q * 32 + lane) : (q * 16 + lane);
        const bool lane_ok = (Lr.M == 128) || (lane < 16);
        const float* cst = a.consts + Lr.c_off;
        // per-feature constants of both m-tiles are fetched while the MMAs are still running
        float cpre[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int f = mt * Lr.M + f_local;
          const bool v = mt < Lr.n_mtiles && lane_ok && f < Lr.N;
#pragma unroll
          for (int j = 0; j < 4; ++j) cpre[mt][j] = v ? __ldg(cst + j * Lr.Npad + f) : 0.f;
        }
        TC_CLK(c0);
        mbar_wait(&S->acc_full, (uint32_t)(n_acc & 1));
        TC_CLK(c1);
        t_acc += c1 - c0;
        ++n_acc;
        tc_fence_after_sync();
        long long e0 = 0, e1 = 0;
        TC_CLK(e0);
        for (int mt = 0; mt < Lr.n_mtiles; ++mt) {
          float v[32];
          const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * TC_ROWS + mbase);
          TC_CLK(c0);
          tmem_ld32(taddr, v);
          tmem_wait_ld();
          TC_CLK(c1);
          t_ld += c1 - c0;
          const int f = mt * Lr.M + f_local;
          const bool valid = lane_ok && f < Lr.N;
          const float c0f = cpre[mt & 1][0], c1f = cpre[mt & 1][1], c2f = cpre[mt & 1][2], c3f = cpre[mt & 1][3];
          const uint32_t boff = (uint32_t)(f >> 2) * TC_LBO_B + (uint32_t)(f & 3) * 4 + (uint32_t)mbase * 16;
          if (Lr.epi == TEPI_BN_LRELU) {
            if (valid) {
#pragma unroll
              for (int m = 0; m < 32; ++m) {
                float y = fmaf((v[m] + c0f) - c1f, c2f, c3f);
                y = y > 0.f ? y : y * a.slope;
                b_store(b_hi, b_lo, boff + m * 16, y);
              }
            }
          } else if (Lr.epi == TEPI_RELU) {
            if (valid) {
#pragma unroll
              for (int m = 0; m < 32; ++m) b_store(b_hi, b_lo, boff + m * 16, fmaxf(v[m] + c0f, 0.f));
            }
          } else if (Lr.epi == TEPI_SIGMOID_X) {
            // few output features (F): the owning lanes only park the pre-activations, then ALL epilogue threads share
            // the sigmoid / split / stores (otherwise F lanes of one warp serialise 64 rows each)
            if (valid) {
#pragma unroll
              for (int m = 0; m < 32; ++m) xs[f * TC_XS_PITCH + mbase + m] = v[m] + c0f;
            }
            named_bar(1, TC_EPI_THREADS);
            for (int i = tid; i < Lr.N * TC_ROWS; i += TC_EPI_THREADS) {
              const int ff = i / TC_ROWS, m = i - ff * TC_ROWS;
              const float y = 1.0f / (1.0f + expf(-xs[ff * TC_XS_PITCH + m]));
              xs[ff * TC_XS_PITCH + m] = y;
              if (next_K) b_store(b_hi, b_lo, (uint32_t)(ff >> 2) * TC_LBO_B + (uint32_t)(ff & 3) * 4 + (uint32_t)m * 16, y);
            }
            if (a.x_all) {
              named_bar(1, TC_EPI_THREADS);
              for (int i = tid; i < nrows * Lr.N; i += TC_EPI_THREADS) {      // row-major order: coalesced stores
                const int m = i / Lr.N, ff = i - m * Lr.N;
                a.x_all[(size_t)(row0 + m) * a.F + ff] = xs[ff * TC_XS_PITCH + m];
              }
            }
          } else if (Lr.epi == TEPI_LOGITS) {
            if (valid) {
#pragma unroll
              for (int m = 0; m < 32; ++m) lg[f * TC_ROWS + mbase + m] = v[m] + c0f;
            }
          } else if (Lr.epi == TEPI_OUT) {
            if (valid) {
#pragma unroll
              for (int m = 0; m < 32; ++m)
                if (mbase + m < nrows) {
                  if (f < a.out_split) a.out_plain[(size_t)(row0 + mbase + m) * a.out_ld + f] = v[m] + c0f;
                  else a.out_plain2[(size_t)(row0 + mbase + m) * a.out_ld + (f - a.out_split)] = v[m] + c0f;
                }
            }
          } else {   // TEPI_LN_RELU (one 128-feature tile): LayerNorm over the features of each row, two-pass variance
#pragma unroll
            for (int m = 0; m < 32; ++m) {
              v[m] += c0f;
              if (valid) ln_scratch[f * TC_LN_PITCH + mbase + m] = v[m];
            }
            named_bar(1, TC_EPI_THREADS);
            const int m_r = tid & 63, part = tid >> 6, nq = Lr.N / 4;
            float s = 0.f;
            {
              float s4[4] = {0.f, 0.f, 0.f, 0.f};           // independent partial sums: the loads pipeline
              const float* col = ln_scratch + (size_t)part * nq * TC_LN_PITCH + m_r;
#pragma unroll 4
              for (int j = 0; j < nq; j += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) s4[u] += col[(j + u) * TC_LN_PITCH];
              }
              s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            }
            red[part * TC_ROWS + m_r] = s;
            named_bar(1, TC_EPI_THREADS);
            const float mean = ((red[m_r] + red[TC_ROWS + m_r]) + (red[2 * TC_ROWS + m_r] + red[3 * TC_ROWS + m_r])) / (float)Lr.N;
            float q2 = 0.f;
            {
              float q4[4] = {0.f, 0.f, 0.f, 0.f};
              const float* col = ln_scratch + (size_t)part * nq * TC_LN_PITCH + m_r;
#pragma unroll 4
              for (int j = 0; j < nq; j += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float dlt = col[(j + u) * TC_LN_PITCH] - mean;
                  q4[u] = fmaf(dlt, dlt, q4[u]);
                }
              }
              q2 = (q4[0] + q4[1]) + (q4[2] + q4[3]);
            }
            red[(4 + part) * TC_ROWS + m_r] = q2;
            named_bar(1, TC_EPI_THREADS);
            if (part == 0) {
              const float var =
                  ((red[4 * TC_ROWS + m_r] + red[5 * TC_ROWS + m_r]) + (red[6 * TC_ROWS + m_r] + red[7 * TC_ROWS + m_r])) / (float)Lr.N;
              red[8 * TC_ROWS + m_r] = mean;
              red[9 * TC_ROWS + m_r] = 1.0f / sqrtf(var + a.ln_eps);
            }
            named_bar(1, TC_EPI_THREADS);
            if (valid) {
#pragma unroll
              for (int m = 0; m < 32; ++m) {
                const float nrm = (v[m] - red[8 * TC_ROWS + mbase + m]) * red[9 * TC_ROWS + mbase + m] * c2f + c3f;
                b_store(b_hi, b_lo, boff + m * 16, fmaxf(nrm, 0.f));
              }
            }
          }
          // zero padding of the next layer's contraction range (N not a multiple of 8)
          if (lane_ok && f >= Lr.N && f < next_K) {
#pragma unroll
            for (int m = 0; m < 32; ++m) b_store(b_hi, b_lo, boff + m * 16, 0.f);
          }
        }
        TC_CLK(c0);
        tc_fence_before_sync();
        if (l + 1 < a.nl) {
          fence_proxy_async_smem();
          mbar_arrive(&S->act_ready);
        }
        TC_CLK(e1);
        t_fence += e1 - c0;
        t_epi_l[l] += e1 - e0;
      }

      // ---- filter decision + compaction (cvae_gan.py:366-370): thread = row ----
      if (a.do_filter) {
        named_bar(1, TC_EPI_THREADS);
        if (tid < TC_ROWS) {
          const int m = tid;
          bool keep = false;
          if (m < nrows) {
            keep = filter_decide([&](int k) { return lg[k * TC_ROWS + m]; }, a.Kc, a.label, a.thr);
            if (a.keep_out) a.keep_out[row0 + m] = keep ? 1 : 0;
            if (a.logits_out)
              for (int k = 0; k < a.Kc; ++k) a.logits_out[(size_t)(row0 + m) * a.Kc + k] = lg[k * TC_ROWS + m];
          }
          const unsigned bal = __ballot_sync(0xffffffffu, keep);
          if (lane == 0) S->warp_cnt[warp] = __popc(bal);
          named_bar(2, TC_ROWS);
          if (tid == 0) {
            const int tot = S->warp_cnt[0] + S->warp_cnt[1];
            S->base = tot ? atomicAdd(a.count, (unsigned long long)tot) : 0ull;
          }
          named_bar(2, TC_ROWS);
          if (keep) {
            const long long pos = (long long)S->base + (warp ? S->warp_cnt[0] : 0) + __popc(bal & ((1u << lane) - 1u));
            if (pos < a.capacity) {
              for (int f = 0; f < a.F; ++f) a.x_out[pos * a.F + f] = xs[f * TC_XS_PITCH + m];
              if (a.idx_out) a.idx_out[pos] = (long long)(a.row_offset + (unsigned long long)(row0 + m));
            }
          }
        }
      }
      named_bar(1, TC_EPI_THREADS);   // xs / lg / red are reused by the next tile
    }
    if (prof && tid == 0) {
      t_all = clock64() - t_all;
      atomicAdd((unsigned long long*)a.dbg + 4, (unsigned long long)t_acc);
      atomicAdd((unsigned long long*)a.dbg + 5, (unsigned long long)t_in);
      atomicAdd((unsigned long long*)a.dbg + 6, (unsigned long long)t_all);
      atomicAdd((unsigned long long*)a.dbg + 16, (unsigned long long)t_ld);
      atomicAdd((unsigned long long*)a.dbg + 17, (unsigned long long)t_fence);
      for (int l = 0; l < a.nl; ++l) atomicAdd((unsigned long long*)a.dbg + 18 + l, (unsigned long long)t_epi_l[l]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == TC_EPI_WARPS) tmem_dealloc(tmem, 128);
}

}  // namespace cvg

#include "eval_tc128.cuh"
#include "eval_tcpp.cuh"

namespace cvg {

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static size_t tc_eval_smem() {
  return 2 * (size_t)TC_BBYTES + (size_t)TC_STAGES * TC_STAGE_BYTES +
         sizeof(float) * ((size_t)TC_MAXF * TC_XS_PITCH + (size_t)FILTER_MAXK * TC_ROWS + 10 * TC_ROWS) + sizeof(TcSmem) + 64;
}

void tc_set_kernel_attributes() {
  cudaFuncSetAttribute(tc_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_eval_smem());
  cudaFuncSetAttribute(tc_eval128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_eval128_smem());
  cudaFuncSetAttribute(tc_eval_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_eval_pp_smem());
}

// Kernel variant: 0 = 64-row tiles (any supported shape), 1 = 128-row tiles, 2 = two 64-row tiles in flight (ping-pong).
// 1 and 2 need F <= 32 and K <= 16 for their side buffers.  CVG_TC_MODE = 64 | 128 | pp overrides (read per call:
// tests toggle it); CVG_TC_ROWS64=1 is the older spelling of CVG_TC_MODE=64.
static int tc_mode(const Engine& e) {
  const bool small = e.F <= TC128_MAXF && e.K <= TC128_MAXKC;
  const char* r64 = getenv("CVG_TC_ROWS64");
  const char* m = getenv("CVG_TC_MODE");
  if ((r64 && r64[0] == '1') || !small) return 0;
  if (m && !strcmp(m, "64")) return 0;
  if (m && !strcmp(m, "128")) return 1;
  if (m && !strcmp(m, "pp")) return 2;
  return CVG_TC_DEFAULT_MODE;
}

// Which networks the tensor-core chain supports: every hidden width a multiple of 64 and <= 256.
bool tc_supported(const Engine& e) {
  auto ok = [](const int* h) { return h[0] <= 256 && h[1] <= 256 && h[2] <= 256 && h[0] % 64 == 0 && h[1] % 64 == 0 && h[2] % 64 == 0; };
  if (e.Kc != e.K) return false;      // unconditional networks (VAE-GAN sibling): the chains fold a label column into the bias
  return ok(e.eh) && ok(e.gh) && ok(e.dh) && ok(e.ch) && e.F <= TC_MAXF && e.Z % 8 == 0 && e.Z <= TC128_MAXK && 2 * e.Z <= 256 &&
         e.K <= FILTER_MAXK && e.ch[1] % 2 == 0;
}

int64_t tc_prep_floats(const Engine& e) {
  // per layer n_mtiles * M * Kpad * 2; generous bound: 4 layers x (256 x 256 x 2) per chain of up to 8 layers
  (void)e;
  return (int64_t)TC_MAX_LAYERS * 256 * 256 * 2;
}
int64_t tc_const_floats() { return (int64_t)TC_MAX_LAYERS * 4 * 256; }

struct ChainBuilder {
  const Engine& e;
  TcPrepArgs p;
  long long w_off = 0;
  int c_off = 0;
  explicit ChainBuilder(const Engine& en) : e(en) { p.nl = 0; }
  // net layer `l`; Kreal = columns of W that are contracted (one-hot label column folded into the bias)
  void add(int net, int l, int Kreal, int label_col, int epi) {
    const LinearP& lp = e.lay[net].lin[l];
    TcLayer& L = p.L[p.nl];
    TcPrepLayer& P = p.P[p.nl];
    L.K = (Kreal + 7) & ~7;
    L.N = lp.out;
    L.M = lp.out > 64 ? 128 : 64;
    L.n_mtiles = (lp.out + L.M - 1) / L.M;
    L.n_kchunks = (L.K + TC_KC - 1) / TC_KC;
    L.epi = epi;
    L.Npad = (lp.out + 7) & ~7;
    L.c_off = c_off;
    L.w_off = w_off;
    c_off += 4 * L.Npad;
    w_off += (long long)L.n_mtiles * L.M * L.K * 2;
    P.W = e.P(net, lp.w);
    P.ldw = lp.in;
    P.Kreal = Kreal;
    P.bias = e.P(net, lp.b);
    P.wlabel = label_col >= 0 ? e.P(net, lp.w) + label_col : nullptr;
    P.ldwl = lp.in;
    P.gamma = lp.gamma >= 0 ? e.P(net, lp.gamma) : nullptr;
    P.beta = lp.beta >= 0 ? e.P(net, lp.beta) : nullptr;
    P.rmean = lp.rmean >= 0 ? e.S(net, lp.rmean) : nullptr;
    P.rvar = lp.rvar >= 0 ? e.S(net, lp.rvar) : nullptr;
    ++p.nl;
  }
};

static int tc_run(Engine& e, ChainBuilder& cb, TcEvalArgs& a, cudaStream_t st) {
  if (!e.ws.tc_w || !e.ws.tc_c) CVG_FAIL("workspace lacks the tensor-core staging buffers");
  if (cb.w_off > tc_prep_floats(e) || cb.c_off > tc_const_floats()) CVG_FAIL("tensor-core chain exceeds its staging buffers");
  cb.p.wprep = e.ws.tc_w;
  cb.p.consts = e.ws.tc_c;
  cb.p.bn_eps = e.cfg.bn_eps;
  tc_prep_kernel<<<dim3(32, cb.p.nl), 256, 0, st>>>(cb.p);
  CVG_LAUNCH_CHECK();
  a.nl = cb.p.nl;
  for (int i = 0; i < a.nl; ++i) a.L[i] = cb.p.L[i];
  a.wprep = e.ws.tc_w;
  a.consts = e.ws.tc_c;
  a.slope = e.cfg.lrelu_slope;
  a.ln_eps = e.cfg.ln_eps;
  a.F = e.F;
  a.Kc = e.K;
  a.dbg = e.tc_dbg;
  const int mode = tc_mode(e);
  if (mode == 2) {
    const long long npairs = ((a.n + TCPP_ROWS - 1) / TCPP_ROWS + 1) / 2;
    const int grid = (int)(npairs < e.num_sms ? npairs : e.num_sms);
    tc_eval_pp_kernel<<<grid, TC_THREADS, tc_eval_pp_smem(), st>>>(a);
  } else if (mode == 1) {
    const long long ntiles = (a.n + TC128_ROWS - 1) / TC128_ROWS;
    const int grid = (int)(ntiles < e.num_sms ? ntiles : e.num_sms);
    tc_eval128_kernel<<<grid, TC_THREADS, tc_eval128_smem(), st>>>(a);
  } else {
    const long long ntiles = (a.n + TC_ROWS - 1) / TC_ROWS;
    const int grid = (int)(ntiles < e.num_sms ? ntiles : e.num_sms);
    tc_eval_kernel<<<grid, TC_THREADS, tc_eval_smem(), st>>>(a);
  }
  CVG_LAUNCH_CHECK();
  return 0;
}

static void add_generator(ChainBuilder& cb, int label) {
  const Engine& e = cb.e;
  cb.add(CVG_NET_GENERATOR, 0, e.Z, e.Z + label, TEPI_BN_LRELU);
  cb.add(CVG_NET_GENERATOR, 1, e.gh[0], -1, TEPI_BN_LRELU);
  cb.add(CVG_NET_GENERATOR, 2, e.gh[1], -1, TEPI_BN_LRELU);
  cb.add(CVG_NET_GENERATOR, 3, e.gh[2], -1, TEPI_SIGMOID_X);
}
static void add_classifier(ChainBuilder& cb) {
  const Engine& e = cb.e;
  cb.add(CVG_NET_CLASSIFIER, 0, e.F, -1, TEPI_RELU);
  cb.add(CVG_NET_CLASSIFIER, 1, e.ch[0], -1, TEPI_LN_RELU);
  cb.add(CVG_NET_CLASSIFIER, 2, e.ch[1], -1, TEPI_RELU);
  cb.add(CVG_NET_CLASSIFIER, 3, e.ch[2], -1, TEPI_LOGITS);
}

static TcEvalArgs blank_args() {
  TcEvalArgs a;
  memset(&a, 0, sizeof(a));
  return a;
}

// generate_samples in eval mode (cvae_gan.py:339-345)
int tc_generate(Engine& e, int label, int64_t n, const float* z, uint64_t seed, uint64_t row_offset, float* x_out, cudaStream_t st) {
  if (n == 0) return 0;
  ChainBuilder cb(e);
  add_generator(cb, label);
  TcEvalArgs a = blank_args();
  a.in_kind = z ? 1 : 0;
  a.in = z;
  a.in_feat = e.Z;
  a.n = n;
  a.seed = seed;
  a.row_offset = row_offset;
  a.x_all = x_out;
  return tc_run(e, cb, a, st);
}

// generate -> classify -> threshold -> compact (cvae_gan.py:357-371)
int tc_generate_filter(Engine& e, int label, int64_t n, float thr, const float* z, uint64_t seed, uint64_t row_offset, float* x_out,
                       int64_t* idx_out, int64_t capacity, unsigned long long* count_out, float* logits_out, uint8_t* keep_out,
                       cudaStream_t st) {
  if (n == 0) return 0;
  ChainBuilder cb(e);
  add_generator(cb, label);
  add_classifier(cb);
  TcEvalArgs a = blank_args();
  a.in_kind = z ? 1 : 0;
  a.in = z;
  a.in_feat = e.Z;
  a.n = n;
  a.seed = seed;
  a.row_offset = row_offset;
  a.do_filter = 1;
  a.label = label;
  a.thr = thr;
  a.x_out = x_out;
  a.idx_out = (long long*)idx_out;
  a.capacity = capacity;
  a.count = count_out;
  a.logits_out = logits_out;
  a.keep_out = keep_out;
  return tc_run(e, cb, a, st);
}

// logits = C(x) in eval mode (cvae_gan.py:362, classifier.py:37,57)
int tc_classifier_forward(Engine& e, const float* x, int64_t n, float* logits_out, cudaStream_t st) {
  if (n == 0) return 0;
  ChainBuilder cb(e);
  add_classifier(cb);
  cb.p.L[cb.p.nl - 1].epi = TEPI_OUT;
  TcEvalArgs a = blank_args();
  a.in_kind = 2;
  a.in = x;
  a.in_feat = e.F;
  a.n = n;
  a.out_plain = logits_out;
  a.out_ld = e.K;
  a.out_split = e.K;
  return tc_run(e, cb, a, st);
}

// mu, logvar = E(x, label) in eval mode (cvae_gan_models.py:47-64)
int tc_encoder_forward(Engine& e, const float* x, int label, int64_t n, float* mu_out, float* lv_out, cudaStream_t st) {
  if (n == 0) return 0;
  ChainBuilder cb(e);
  cb.add(CVG_NET_ENCODER, 0, e.F, e.F + label, TEPI_BN_LRELU);
  cb.add(CVG_NET_ENCODER, 1, e.eh[0], -1, TEPI_BN_LRELU);
  cb.add(CVG_NET_ENCODER, 2, e.eh[1], -1, TEPI_BN_LRELU);
  cb.add(CVG_NET_ENCODER, 3, e.eh[2], -1, TEPI_OUT);
  TcEvalArgs a = blank_args();
  a.in_kind = 2;
  a.in = x;
  a.in_feat = e.F;
  a.n = n;
  a.out_plain = mu_out;
  a.out_plain2 = lv_out;
  a.out_ld = e.Z;
  a.out_split = e.Z;
  return tc_run(e, cb, a, st);
}

}  // namespace cvg
