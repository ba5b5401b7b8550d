// cvaegan_b200 - ping-pong variant of the fused eval chain (included by eval_tc.cu): TWO 64-row tiles in flight per CTA.
//
// In tc_eval_kernel / tc_eval128_kernel the MMAs of a layer and the epilogue that turns its accumulator into the next
// layer's operand are serialised (one tile, one set of activation planes).  Here every CTA owns two tile slots with
// their own planes (64 rows x 128 features x hi/lo = 65 KB each, the K = 256 layers are consumed in K-halves with the
// m-tile accumulators parked in TMEM exactly like eval_tc128.cuh) and their own TMEM accumulators (3 x 64 columns):
// while the eight epilogue warps work on slot A's layer l, the issuer runs slot B's MMAs of the same layer, then they
// swap.  Both roles walk the same static sequence of (slot, layer, half) items; per-slot mbarriers carry the
// dependencies (act_ready: operand written, acc_full: accumulator complete, plane_free: K-half-0 MMAs done).
// Each layer's weights are streamed once per slot (same L2 traffic per row as the 64-row kernel).
#pragma once

namespace cvg {

constexpr int TCPP_ROWS = 64;
constexpr int TCPP_MAXK = 128;
constexpr int TCPP_LBO_B = TCPP_ROWS * 16 + 16;
constexpr int TCPP_BBYTES = (TCPP_MAXK / 4) * TCPP_LBO_B;     // one plane (hi or lo) of one slot
constexpr int TCPP_MAXF = 32;
constexpr int TCPP_MAXKC = 16;
constexpr int TCPP_XS_PITCH = TCPP_ROWS + 1;
constexpr int TCPP_ACC_COLS = 192;                            // TMEM columns per slot: three 64-column accumulators

struct TcppSmem {
  uint64_t full[TC_STAGES], empty[TC_STAGES];
  uint64_t acc_full[2], act_ready[2], plane_free[2];
  uint32_t tmem_slot;
  int warp_cnt[4];
  unsigned long long base;
};

// sum over the 32 lanes of a warp of 32 per-lane values; on return lane l holds the sum of value l
__device__ __forceinline__ float warp_transpose_sum32(const float* v, int lane) {
  float r16[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float keep = (lane & 16) ? v[16 + i] : v[i];
    const float send = (lane & 16) ? v[i] : v[16 + i];
    r16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float r8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float keep = (lane & 8) ? r16[8 + i] : r16[i];
    const float send = (lane & 8) ? r16[i] : r16[8 + i];
    r8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float r4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = (lane & 4) ? r8[4 + i] : r8[i];
    const float send = (lane & 4) ? r8[i] : r8[4 + i];
    r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  float r2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = (lane & 2) ? r4[2 + i] : r4[i];
    const float send = (lane & 2) ? r4[i] : r4[2 + i];
    r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  const float keep = (lane & 1) ? r2[1] : r2[0];
  const float send = (lane & 1) ? r2[0] : r2[1];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

__device__ __forceinline__ uint32_t bpp_off(int f, int m) {
  return (uint32_t)(f >> 2) * TCPP_LBO_B + (uint32_t)(f & 3) * 4 + (uint32_t)m * 16;
}

__global__ void __launch_bounds__(TC_THREADS, 1) tc_eval_pp_kernel(const __grid_constant__ TcEvalArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* planes = smem;                                       // [slot][hi | lo][TCPP_BBYTES]
  uint8_t* stages = smem + 4 * TCPP_BBYTES;
  float* xs_all = reinterpret_cast<float*>(stages + TC_STAGES * TC_STAGE_BYTES);   // [slot][TCPP_MAXF][65]
  float* lg_all = xs_all + 2 * TCPP_MAXF * TCPP_XS_PITCH;                           // [slot][16][64]
  float* red_all = lg_all + 2 * TCPP_MAXKC * TCPP_ROWS;                             // [slot][10][64]
  TcppSmem* S = reinterpret_cast<TcppSmem*>(red_all + 2 * 10 * TCPP_ROWS);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const long long ntiles = (a.n + TCPP_ROWS - 1) / TCPP_ROWS;
  const long long npairs = (ntiles + 1) / 2;

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&S->full[s], 1); mbar_init(&S->empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&S->acc_full[s], 1);
      mbar_init(&S->plane_free[s], 1);
      mbar_init(&S->act_ready[s], TC_EPI_THREADS);
    }
    fence_mbar_init();
  }
  if (warp == TC_EPI_WARPS) tmem_alloc(&S->tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, S->tmem_slot, 0);

  long long my_pairs = 0;
  if ((long long)blockIdx.x < npairs) my_pairs = (npairs - 1 - blockIdx.x) / gridDim.x + 1;
  const int half_chunks = TCPP_MAXK / TC_KC;

  if (warp == TC_EPI_WARPS) {
    // ===================== MMA issuer =====================
    unsigned long long g = 0, n_act[2] = {0, 0};
    bool peeked = false;
    const uint32_t planes_a = smem_u32(planes), st_a = smem_u32(stages);
    // chunks [kc0, kc1) of every m-tile of layer Lr for slot `sl`
    auto issue_range = [&](const TcLayer& Lr, int sl, int kc0, int kc1, uint32_t acc0, bool last_of_layer) {
      const uint32_t idesc = idesc_tf32(Lr.M, TCPP_ROWS, 0, 0);
      const uint32_t a_lbo = (uint32_t)Lr.M * 16u;
      const uint32_t b_hi_a = planes_a + (uint32_t)sl * 2u * TCPP_BBYTES, b_lo_a = b_hi_a + TCPP_BBYTES;
      for (int mt = 0; mt < Lr.n_mtiles; ++mt) {
        for (int kc = kc0; kc < kc1; ++kc) {
          const int s = (int)(g % TC_STAGES);
          if (!peeked) mbar_wait(&S->full[s], (uint32_t)((g / TC_STAGES) & 1));
          tc_fence_after_sync();
          const bool peek_next = mbar_test_wait(&S->full[(g + 1) % TC_STAGES], (uint32_t)(((g + 1) / TC_STAGES) & 1));
          if (elect_one()) {
            const int kc_len = min(TC_KC, Lr.K - kc * TC_KC);
            const uint32_t a_hi = st_a + (uint32_t)s * TC_STAGE_BYTES;
            const uint32_t a_lo = a_hi + (uint32_t)kc_len * Lr.M * 4u;
            const uint32_t boff = (uint32_t)((kc % half_chunks) * TC_KC / 4) * TCPP_LBO_B;
            uint64_t dah = smem_desc(a_hi, a_lbo, 128), dal = smem_desc(a_lo, a_lbo, 128);
            uint64_t dbh = smem_desc(b_hi_a + boff, TCPP_LBO_B, 128), dbl = smem_desc(b_lo_a + boff, TCPP_LBO_B, 128);
            const uint32_t d = tmem + (uint32_t)sl * TCPP_ACC_COLS + acc0 + (uint32_t)mt * TCPP_ROWS;
            const uint64_t a_step = (uint64_t)((2 * a_lbo) >> 4), b_step = (uint64_t)((2 * TCPP_LBO_B) >> 4);
            for (int ks = 0; ks < kc_len / 8; ++ks) {
              mma_tf32(d, dal, dbh, idesc, !(kc == 0 && ks == 0));
              mma_tf32(d, dah, dbl, idesc, true);
              mma_tf32(d, dah, dbh, idesc, true);
              dah += a_step; dal += a_step; dbh += b_step; dbl += b_step;
            }
            mma_commit(&S->empty[s]);
            if (mt == Lr.n_mtiles - 1 && kc == kc1 - 1) mma_commit(last_of_layer ? &S->acc_full[sl] : &S->plane_free[sl]);
          }
          __syncwarp();
          peeked = peek_next;
          ++g;
        }
      }
    };
    for (long long t = 0; t < my_pairs; ++t) {
      for (int l = 0; l < a.nl; ++l) {
        const TcLayer& Lr = a.L[l];
        const bool split_in = Lr.K > TCPP_MAXK;
        if (!split_in) {
          for (int sl = 0; sl < 2; ++sl) {
            mbar_wait(&S->act_ready[sl], (uint32_t)(n_act[sl] & 1));
            ++n_act[sl];
            tc_fence_after_sync();
            issue_range(Lr, sl, 0, Lr.n_kchunks, 0u, true);
          }
        } else {
          for (int hf = 0; hf < 2; ++hf)
            for (int sl = 0; sl < 2; ++sl) {
              mbar_wait(&S->act_ready[sl], (uint32_t)(n_act[sl] & 1));
              ++n_act[sl];
              tc_fence_after_sync();
              issue_range(Lr, sl, hf * half_chunks, hf ? Lr.n_kchunks : half_chunks, 128u, hf == 1);
            }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ===================== weight producer: same item order as the issuer =====================
    unsigned long long gl = 0;
    auto load_range = [&](const TcLayer& Lr, int kc0, int kc1) {
      for (int mt = 0; mt < Lr.n_mtiles; ++mt)
        for (int kc = kc0; kc < kc1; ++kc) {
          const int s = (int)(gl % TC_STAGES);
          if (gl >= TC_STAGES) mbar_wait(&S->empty[s], (uint32_t)(((gl / TC_STAGES) - 1) & 1));
          const int kc_len = min(TC_KC, Lr.K - kc * TC_KC);
          const uint32_t bytes = 2u * kc_len * Lr.M * 4u;
          const float* src = a.wprep + Lr.w_off + (size_t)mt * Lr.K * Lr.M * 2 + (size_t)kc * TC_KC * Lr.M * 2;
          if (elect_one()) {
            mbar_arrive_expect_tx(&S->full[s], bytes);
            bulk_g2s(stages + (size_t)s * TC_STAGE_BYTES, src, bytes, &S->full[s]);
          }
          __syncwarp();
          ++gl;
        }
    };
    for (long long t = 0; t < my_pairs; ++t)
      for (int l = 0; l < a.nl; ++l) {
        const TcLayer& Lr = a.L[l];
        if (Lr.K <= TCPP_MAXK) {
          load_range(Lr, 0, Lr.n_kchunks);
          load_range(Lr, 0, Lr.n_kchunks);
        } else {
          load_range(Lr, 0, half_chunks);
          load_range(Lr, 0, half_chunks);
          load_range(Lr, half_chunks, Lr.n_kchunks);
          load_range(Lr, half_chunks, Lr.n_kchunks);
        }
      }
  } else {
    // ============== epilogue warps: thread = (output feature = TMEM lane, 32 of the 64 rows of a slot) ==============
    unsigned long long n_acc[2] = {0, 0}, n_pf[2] = {0, 0};
    const int in_groups = a.L[0].K / 4;
    const int q = warp & 3, h = warp >> 2;
    const int mbase = h * 32;
    for (long long pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
      long long row0s[2];
      int nrows_s[2];
      for (int sl = 0; sl < 2; ++sl) {
        row0s[sl] = (2 * pr + sl) * TCPP_ROWS;
        const long long left = a.n - row0s[sl];
        nrows_s[sl] = (int)(left < 0 ? 0 : (left > TCPP_ROWS ? TCPP_ROWS : left));
      }
      // ---- layer-0 operands of both slots ----
      for (int sl = 0; sl < 2; ++sl) {
        uint8_t* b_hi = planes + (size_t)sl * 2 * TCPP_BBYTES;
        uint8_t* b_lo = b_hi + TCPP_BBYTES;
        const long long row0 = row0s[sl];
        const int nrows = nrows_s[sl];
        for (int i = tid; i < TCPP_ROWS * in_groups; i += TC_EPI_THREADS) {
          const int m = i % TCPP_ROWS, fg = i / TCPP_ROWS;
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
          const uint32_t off = (uint32_t)fg * TCPP_LBO_B + (uint32_t)m * 16;
          *reinterpret_cast<float4*>(b_hi + off) = hi;
          *reinterpret_cast<float4*>(b_lo + off) = lo;
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(&S->act_ready[sl]);
      }

      for (int l = 0; l < a.nl; ++l) {
        const TcLayer& Lr = a.L[l];
        const int next_K = (l + 1 < a.nl) ? a.L[l + 1].K : 0;
        const bool split_out = next_K > TCPP_MAXK;
        const bool split_in = Lr.K > TCPP_MAXK;
        const uint32_t acc0 = split_in ? 128u : 0u;
        const int f_local = (Lr.M == 128) ? (q * 32 + lane) : (q * 16 + lane);
        const bool lane_ok = (Lr.M == 128) || (lane < 16);
        const float* cst = a.consts + Lr.c_off;
        float cpre[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int f = mt * Lr.M + f_local;
          const bool vv = mt < Lr.n_mtiles && lane_ok && f < Lr.N;
#pragma unroll
          for (int j = 0; j < 4; ++j) cpre[mt][j] = vv ? __ldg(cst + j * Lr.Npad + f) : 0.f;
        }
        for (int mt = 0; mt < Lr.n_mtiles; ++mt) {
          for (int sl = 0; sl < 2; ++sl) {
            uint8_t* b_hi = planes + (size_t)sl * 2 * TCPP_BBYTES;
            uint8_t* b_lo = b_hi + TCPP_BBYTES;
            float* xs = xs_all + (size_t)sl * TCPP_MAXF * TCPP_XS_PITCH;
            float* lg = lg_all + (size_t)sl * TCPP_MAXKC * TCPP_ROWS;
            float* red = red_all + (size_t)sl * 10 * TCPP_ROWS;
            const long long row0 = row0s[sl];
            const int nrows = nrows_s[sl];
            if (mt == 0) {
              mbar_wait(&S->acc_full[sl], (uint32_t)(n_acc[sl] & 1));
              ++n_acc[sl];
            } else if (split_out) {
              mbar_wait(&S->plane_free[sl], (uint32_t)(n_pf[sl] & 1));
              ++n_pf[sl];
            }
            tc_fence_after_sync();
            float v[32];
            const uint32_t taddr = tmem + (uint32_t)sl * TCPP_ACC_COLS + acc0 + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * TCPP_ROWS + mbase);
            tmem_ld32(taddr, v);
            tmem_wait_ld();
            const int f = mt * Lr.M + f_local;
            const int fp = split_out ? f_local : f;
            const bool valid = lane_ok && f < Lr.N;
            const float c0f = cpre[mt & 1][0], c1f = cpre[mt & 1][1], c2f = cpre[mt & 1][2], c3f = cpre[mt & 1][3];
            const uint32_t boff = bpp_off(fp, mbase);
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
              if (valid) {
#pragma unroll
                for (int m = 0; m < 32; ++m) xs[f * TCPP_XS_PITCH + mbase + m] = v[m] + c0f;
              }
              named_bar(1, TC_EPI_THREADS);
              for (int i = tid; i < Lr.N * TCPP_ROWS; i += TC_EPI_THREADS) {
                const int ff = i / TCPP_ROWS, m = i - ff * TCPP_ROWS;
                const float y = 1.0f / (1.0f + expf(-xs[ff * TCPP_XS_PITCH + m]));
                xs[ff * TCPP_XS_PITCH + m] = y;
                if (next_K) b_store(b_hi, b_lo, bpp_off(ff, m), y);
              }
              if (a.x_all) {
                named_bar(1, TC_EPI_THREADS);
                for (int i = tid; i < nrows * Lr.N; i += TC_EPI_THREADS) {
                  const int m = i / Lr.N, ff = i - m * Lr.N;
                  a.x_all[(size_t)(row0 + m) * a.F + ff] = xs[ff * TCPP_XS_PITCH + m];
                }
              }
            } else if (Lr.epi == TEPI_LOGITS) {
              if (valid) {
#pragma unroll
                for (int m = 0; m < 32; ++m) lg[f * TCPP_ROWS + mbase + m] = v[m] + c0f;
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
            } else {   // TEPI_LN_RELU
#pragma unroll
              for (int m = 0; m < 32; ++m) v[m] = valid ? v[m] + c0f : 0.f;
              float s0 = warp_transpose_sum32(v, lane);
              red[q * TCPP_ROWS + mbase + lane] = s0;
              named_bar(1, TC_EPI_THREADS);
              if (tid < TCPP_ROWS)
                red[8 * TCPP_ROWS + tid] = ((red[tid] + red[TCPP_ROWS + tid]) + (red[2 * TCPP_ROWS + tid] + red[3 * TCPP_ROWS + tid])) / (float)Lr.N;
              named_bar(1, TC_EPI_THREADS);
              {
                float d[32];
#pragma unroll
                for (int m = 0; m < 32; ++m) {
                  const float dl = valid ? v[m] - red[8 * TCPP_ROWS + mbase + m] : 0.f;
                  d[m] = dl * dl;
                }
                s0 = warp_transpose_sum32(d, lane);
              }
              red[(4 + q) * TCPP_ROWS + mbase + lane] = s0;
              named_bar(1, TC_EPI_THREADS);
              if (tid < TCPP_ROWS) {
                const float var = ((red[4 * TCPP_ROWS + tid] + red[5 * TCPP_ROWS + tid]) + (red[6 * TCPP_ROWS + tid] + red[7 * TCPP_ROWS + tid])) / (float)Lr.N;
                red[9 * TCPP_ROWS + tid] = 1.0f / sqrtf(var + a.ln_eps);
              }
              named_bar(1, TC_EPI_THREADS);
              if (valid) {
#pragma unroll
                for (int m = 0; m < 32; ++m) {
                  const float nrm = (v[m] - red[8 * TCPP_ROWS + mbase + m]) * red[9 * TCPP_ROWS + mbase + m] * c2f + c3f;
                  b_store(b_hi, b_lo, boff + m * 16, fmaxf(nrm, 0.f));
                }
              }
            }
            if (lane_ok && f >= Lr.N && f < next_K && !split_out) {
#pragma unroll
              for (int m = 0; m < 32; ++m) b_store(b_hi, b_lo, boff + m * 16, 0.f);
            }
            // hand the operand (or operand half) of this slot to the issuer
            tc_fence_before_sync();
            if (l + 1 < a.nl && (split_out || mt == Lr.n_mtiles - 1)) {
              fence_proxy_async_smem();
              mbar_arrive(&S->act_ready[sl]);
            }
          }
        }
      }

      // ---- filter decision + compaction for both slots together: thread = (slot, row) ----
      if (a.do_filter) {
        named_bar(1, TC_EPI_THREADS);
        if (tid < 2 * TCPP_ROWS) {
          const int sl = tid >> 6, m = tid & 63;
          const float* lg = lg_all + (size_t)sl * TCPP_MAXKC * TCPP_ROWS;
          const float* xs = xs_all + (size_t)sl * TCPP_MAXF * TCPP_XS_PITCH;
          const long long row0 = row0s[sl];
          bool keep = false;
          if (m < nrows_s[sl]) {
            keep = filter_decide([&](int k) { return lg[k * TCPP_ROWS + m]; }, a.Kc, a.label, a.thr);
            if (a.keep_out) a.keep_out[row0 + m] = keep ? 1 : 0;
            if (a.logits_out)
              for (int k = 0; k < a.Kc; ++k) a.logits_out[(size_t)(row0 + m) * a.Kc + k] = lg[k * TCPP_ROWS + m];
          }
          const unsigned bal = __ballot_sync(0xffffffffu, keep);
          if (lane == 0) S->warp_cnt[warp] = __popc(bal);
          named_bar(2, 2 * TCPP_ROWS);
          if (tid == 0) {
            int tot = 0;
            for (int k = 0; k < 4; ++k) { const int c = S->warp_cnt[k]; S->warp_cnt[k] = tot; tot += c; }
            S->base = tot ? atomicAdd(a.count, (unsigned long long)tot) : 0ull;
          }
          named_bar(2, 2 * TCPP_ROWS);
          if (keep) {
            const long long pos = (long long)S->base + S->warp_cnt[warp] + __popc(bal & ((1u << lane) - 1u));
            if (pos < a.capacity) {
              for (int f = 0; f < a.F; ++f) a.x_out[pos * a.F + f] = xs[f * TCPP_XS_PITCH + m];
              if (a.idx_out) a.idx_out[pos] = (long long)(a.row_offset + (unsigned long long)(row0 + m));
            }
          }
        }
      }
      named_bar(1, TC_EPI_THREADS);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == TC_EPI_WARPS) tmem_dealloc(tmem, 512);
}

inline size_t tc_eval_pp_smem() {
  return 4 * (size_t)TCPP_BBYTES + (size_t)TC_STAGES * TC_STAGE_BYTES +
         sizeof(float) * 2 * ((size_t)TCPP_MAXF * TCPP_XS_PITCH + (size_t)TCPP_MAXKC * TCPP_ROWS + 10 * TCPP_ROWS) +
         sizeof(TcppSmem) + 64;
}

}  // namespace cvg
