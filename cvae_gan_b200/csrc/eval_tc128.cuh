// cvaegan_b200 - 128-row-tile variant of the fused eval chain (included by eval_tc.cu).
//
// One M=128 tf32 MMA costs the same ~64 cycles for N = 64 and N = 128 batch rows (it is bound by the weight operand
// read), so 128-row tiles halve the tensor-pipe time, the weight traffic and the barrier hand-offs per row.  hi+lo
// activation planes of a 256-wide layer for 128 rows (266 KB) do not fit in shared memory, so the planes hold at most
// 128 features and the extra buffering moves to TENSOR MEMORY:
//   * a 256-wide layer (two 128-feature m-tiles) keeps BOTH accumulators in TMEM (columns [0,128) and [128,256));
//   * its consumer (K = 256) runs in two K-halves: epilogue(m-tile 0) -> planes -> MMAs of K-half 0 into a THIRD
//     accumulator (columns [256,384)) -> `plane_free` -> epilogue(m-tile 1) -> planes -> MMAs of K-half 1.
// LayerNorm has no spare plane to transpose through: every thread owns one feature x 64 rows, so the per-row moments
// are taken with a transposing butterfly (62 shuffles reduce 64 values across the 32 lanes of a warp; lane l ends up
// with rows 2l, 2l+1) and a 4-way combine of the lane quadrants in shared memory; two passes like torch.
#pragma once

namespace cvg {

constexpr int TC128_ROWS = 128;
constexpr int TC128_MAXK = 128;                               // features the planes hold at a time
constexpr int TC128_LBO_B = TC128_ROWS * 16 + 16;             // padded pitch of a 4-k group
constexpr int TC128_BBYTES = (TC128_MAXK / 4) * TC128_LBO_B;  // one plane (hi or lo)
constexpr int TC128_MAXF = 32;
constexpr int TC128_MAXKC = 16;                               // classes
constexpr int TC128_XS_PITCH = TC128_ROWS + 1;

struct Tc128Smem {
  uint64_t full[TC_STAGES], empty[TC_STAGES], acc_full, act_ready, plane_free;
  uint32_t tmem_slot;
  int warp_cnt[4];
  unsigned long long base;
};

// sum over the 32 lanes of a warp of 64 per-lane values; on return lane l holds the sums of values 2l and 2l+1
__device__ __forceinline__ void warp_transpose_sum64(const float* v, int lane, float& s0, float& s1) {
  float r32[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float keep = (lane & 16) ? v[32 + i] : v[i];
    const float send = (lane & 16) ? v[i] : v[32 + i];
    r32[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float r16[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float keep = (lane & 8) ? r32[16 + i] : r32[i];
    const float send = (lane & 8) ? r32[i] : r32[16 + i];
    r16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float r8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float keep = (lane & 4) ? r16[8 + i] : r16[i];
    const float send = (lane & 4) ? r16[i] : r16[8 + i];
    r8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  float r4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = (lane & 2) ? r8[4 + i] : r8[i];
    const float send = (lane & 2) ? r8[i] : r8[4 + i];
    r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const float k0 = (lane & 1) ? r4[2] : r4[0], k1 = (lane & 1) ? r4[3] : r4[1];
    const float n0 = (lane & 1) ? r4[0] : r4[2], n1 = (lane & 1) ? r4[1] : r4[3];
    s0 = k0 + __shfl_xor_sync(0xffffffffu, n0, 1);
    s1 = k1 + __shfl_xor_sync(0xffffffffu, n1, 1);
  }
}

__device__ __forceinline__ uint32_t b128_off(int f, int m) {
  return (uint32_t)(f >> 2) * TC128_LBO_B + (uint32_t)(f & 3) * 4 + (uint32_t)m * 16;
}

__global__ void __launch_bounds__(TC_THREADS, 1) tc_eval128_kernel(const __grid_constant__ TcEvalArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* b_hi = smem;
  uint8_t* b_lo = smem + TC128_BBYTES;
  uint8_t* stages = smem + 2 * TC128_BBYTES;
  float* xs = reinterpret_cast<float*>(stages + TC_STAGES * TC_STAGE_BYTES);   // [TC128_MAXF][129] generator outputs
  float* lg = xs + TC128_MAXF * TC128_XS_PITCH;                                 // [16][128] logits
  float* red = lg + TC128_MAXKC * TC128_ROWS;                                   // [10][128] LayerNorm scratch
  Tc128Smem* S = reinterpret_cast<Tc128Smem*>(red + 10 * TC128_ROWS);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const long long ntiles = (a.n + TC128_ROWS - 1) / TC128_ROWS;

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&S->full[s], 1); mbar_init(&S->empty[s], 1); }
    mbar_init(&S->acc_full, 1);
    mbar_init(&S->plane_free, 1);
    mbar_init(&S->act_ready, TC_EPI_THREADS);
    fence_mbar_init();
  }
  if (warp == TC_EPI_WARPS) tmem_alloc(&S->tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, S->tmem_slot, 0);

  long long my_tiles = 0;
  if ((long long)blockIdx.x < ntiles) my_tiles = (ntiles - 1 - blockIdx.x) / gridDim.x + 1;

  if (warp == TC_EPI_WARPS) {
    // ===================== MMA issuer =====================
    unsigned long long g = 0, n_act = 0;
    bool peeked = false;
    const bool prof = a.dbg != nullptr;
    long long t_act = 0, t_full = 0, t_issue = 0, t_all = 0, c0 = 0, c1 = 0;
    TC_CLK(t_all);
    const uint32_t b_hi_a = smem_u32(b_hi), b_lo_a = smem_u32(b_lo), st_a = smem_u32(stages);
    for (long long t = 0; t < my_tiles; ++t) {
      for (int l = 0; l < a.nl; ++l) {
        const TcLayer& Lr = a.L[l];
        const bool split_in = Lr.K > TC128_MAXK;                 // consumed in two K-halves of 128
        const uint32_t idesc = idesc_tf32(Lr.M, TC128_ROWS, 0, 0);
        const uint32_t a_lbo = (uint32_t)Lr.M * 16u;
        const uint32_t acc0 = split_in ? 256u : 0u;               // third accumulator for a split consumer
        const int half_chunks = TC128_MAXK / TC_KC;
        TC_CLK(c0);
        mbar_wait(&S->act_ready, (uint32_t)(n_act & 1));
        TC_CLK(c1);
        t_act += c1 - c0;
        ++n_act;
        tc_fence_after_sync();
        for (int mt = 0; mt < Lr.n_mtiles; ++mt) {
          for (int kc = 0; kc < Lr.n_kchunks; ++kc) {
            if (split_in && kc == half_chunks) {
              // planes may be overwritten once the K-half-0 MMAs have completed; then wait for K-half 1
              if (elect_one()) mma_commit(&S->plane_free);
              __syncwarp();
              TC_CLK(c0);
              mbar_wait(&S->act_ready, (uint32_t)(n_act & 1));
              TC_CLK(c1);
              t_act += c1 - c0;
              ++n_act;
              tc_fence_after_sync();
            }
            const int s = (int)(g % TC_STAGES);
            TC_CLK(c0);
            if (!peeked) mbar_wait(&S->full[s], (uint32_t)((g / TC_STAGES) & 1));
            TC_CLK(c1);
            t_full += c1 - c0;
            tc_fence_after_sync();
            const bool peek_next = mbar_test_wait(&S->full[(g + 1) % TC_STAGES], (uint32_t)(((g + 1) / TC_STAGES) & 1));
            TC_CLK(c0);
            if (elect_one()) {
              const int kc_len = min(TC_KC, Lr.K - kc * TC_KC);
              const uint32_t a_hi = st_a + (uint32_t)s * TC_STAGE_BYTES;
              const uint32_t a_lo = a_hi + (uint32_t)kc_len * Lr.M * 4u;
              const uint32_t boff = (uint32_t)((kc % half_chunks) * TC_KC / 4) * TC128_LBO_B;
              uint64_t dah = smem_desc(a_hi, a_lbo, 128), dal = smem_desc(a_lo, a_lbo, 128);
              uint64_t dbh = smem_desc(b_hi_a + boff, TC128_LBO_B, 128), dbl = smem_desc(b_lo_a + boff, TC128_LBO_B, 128);
              const uint32_t d = tmem + acc0 + (uint32_t)mt * TC128_ROWS;
              const uint64_t a_step = (uint64_t)((2 * a_lbo) >> 4), b_step = (uint64_t)((2 * TC128_LBO_B) >> 4);
              for (int ks = 0; ks < kc_len / 8; ++ks) {
                mma_tf32(d, dal, dbh, idesc, !(kc == 0 && ks == 0));
                mma_tf32(d, dah, dbl, idesc, true);
                mma_tf32(d, dah, dbh, idesc, true);
                dah += a_step; dal += a_step; dbh += b_step; dbl += b_step;
              }
              mma_commit(&S->empty[s]);
              if (mt == Lr.n_mtiles - 1 && kc == Lr.n_kchunks - 1) mma_commit(&S->acc_full);
            }
            __syncwarp();
            TC_CLK(c1);
            t_issue += c1 - c0;
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
      atomicAdd((unsigned long long*)a.dbg + 3, (unsigned long long)t_all);
      atomicAdd((unsigned long long*)a.dbg + 7, (unsigned long long)t_issue);
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ===================== weight producer =====================
    int cpt = 0;
    for (int l = 0; l < a.nl; ++l) cpt += a.L[l].n_mtiles * a.L[l].n_kchunks;
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
    // ============== epilogue warps: thread = (output feature = TMEM lane, 64 of the 128 rows) ==============
    unsigned long long n_acc = 0, n_pf = 0;
    const bool prof = a.dbg != nullptr;
    long long t_acc = 0, t_in = 0, t_all = 0, t_pf = 0, c0 = 0, c1 = 0, e0 = 0, e1 = 0, t_epi_l[TC_MAX_LAYERS] = {0, 0, 0, 0, 0, 0, 0, 0};
    TC_CLK(t_all);
    const int in_groups = a.L[0].K / 4;
    const int q = warp & 3, h = warp >> 2;
    const int mbase = h * 64;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long row0 = tile * TC128_ROWS;
      const int nrows = (int)min((long long)TC128_ROWS, a.n - row0);
      TC_CLK(c0);
      for (int i = tid; i < TC128_ROWS * in_groups; i += TC_EPI_THREADS) {
        const int m = i % TC128_ROWS, fg = i / TC128_ROWS;
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
        const uint32_t off = (uint32_t)fg * TC128_LBO_B + (uint32_t)m * 16;
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
        const bool split_out = next_K > TC128_MAXK;        // the consumer takes this layer's two m-tiles as K-halves
        const bool split_in = Lr.K > TC128_MAXK;
        const uint32_t acc0 = split_in ? 256u : 0u;
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
        TC_CLK(c0);
        mbar_wait(&S->acc_full, (uint32_t)(n_acc & 1));
        TC_CLK(c1);
        t_acc += c1 - c0;
        ++n_acc;
        tc_fence_after_sync();
        TC_CLK(e0);
        for (int mt = 0; mt < Lr.n_mtiles; ++mt) {
          if (split_out && mt == 1) {
            // m-tile 0 went to the planes; the consumer's K-half-0 MMAs must finish before they are overwritten
            tc_fence_before_sync();
            fence_proxy_async_smem();
            mbar_arrive(&S->act_ready);
            TC_CLK(c0);
            mbar_wait(&S->plane_free, (uint32_t)(n_pf & 1));
            TC_CLK(c1);
            t_pf += c1 - c0;
            ++n_pf;
            tc_fence_after_sync();
          }
          float v[64];
          const uint32_t taddr = tmem + acc0 + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * TC128_ROWS + mbase);
          tmem_ld32(taddr, v);
          tmem_ld32(taddr + 32, v + 32);
          tmem_wait_ld();
          const int f = mt * Lr.M + f_local;                 // feature index of the layer output
          const int fp = split_out ? f_local : f;            // its position in the planes (K-half local when split)
          const bool valid = lane_ok && f < Lr.N;
          const float c0f = cpre[mt & 1][0], c1f = cpre[mt & 1][1], c2f = cpre[mt & 1][2], c3f = cpre[mt & 1][3];
          const uint32_t boff = b128_off(fp, mbase);
          if (Lr.epi == TEPI_BN_LRELU) {
            if (valid) {
#pragma unroll
              for (int m = 0; m < 64; ++m) {
                float y = fmaf((v[m] + c0f) - c1f, c2f, c3f);
                y = y > 0.f ? y : y * a.slope;
                b_store(b_hi, b_lo, boff + m * 16, y);
              }
            }
          } else if (Lr.epi == TEPI_RELU) {
            if (valid) {
#pragma unroll
              for (int m = 0; m < 64; ++m) b_store(b_hi, b_lo, boff + m * 16, fmaxf(v[m] + c0f, 0.f));
            }
          } else if (Lr.epi == TEPI_SIGMOID_X) {
            if (valid) {
#pragma unroll
              for (int m = 0; m < 64; ++m) xs[f * TC128_XS_PITCH + mbase + m] = v[m] + c0f;
            }
            named_bar(1, TC_EPI_THREADS);
            for (int i = tid; i < Lr.N * TC128_ROWS; i += TC_EPI_THREADS) {
              const int ff = i / TC128_ROWS, m = i - ff * TC128_ROWS;
              const float y = 1.0f / (1.0f + expf(-xs[ff * TC128_XS_PITCH + m]));
              xs[ff * TC128_XS_PITCH + m] = y;
              if (next_K) b_store(b_hi, b_lo, b128_off(ff, m), y);
            }
            if (a.x_all) {
              named_bar(1, TC_EPI_THREADS);
              for (int i = tid; i < nrows * Lr.N; i += TC_EPI_THREADS) {
                const int m = i / Lr.N, ff = i - m * Lr.N;
                a.x_all[(size_t)(row0 + m) * a.F + ff] = xs[ff * TC128_XS_PITCH + m];
              }
            }
          } else if (Lr.epi == TEPI_LOGITS) {
            if (valid) {
#pragma unroll
              for (int m = 0; m < 64; ++m) lg[f * TC128_ROWS + mbase + m] = v[m] + c0f;
            }
          } else if (Lr.epi == TEPI_OUT) {
            if (valid) {
#pragma unroll
              for (int m = 0; m < 64; ++m)
                if (mbase + m < nrows) {
                  if (f < a.out_split) a.out_plain[(size_t)(row0 + mbase + m) * a.out_ld + f] = v[m] + c0f;
                  else a.out_plain2[(size_t)(row0 + mbase + m) * a.out_ld + (f - a.out_split)] = v[m] + c0f;
                }
            }
          } else {   // TEPI_LN_RELU: one 128-feature tile; per-row moments by transposing butterflies + quadrant combine
#pragma unroll
            for (int m = 0; m < 64; ++m) v[m] = valid ? v[m] + c0f : 0.f;
            float s0, s1;
            warp_transpose_sum64(v, lane, s0, s1);
            red[q * TC128_ROWS + mbase + 2 * lane] = s0;
            red[q * TC128_ROWS + mbase + 2 * lane + 1] = s1;
            named_bar(1, TC_EPI_THREADS);
            if (tid < TC128_ROWS)
              red[8 * TC128_ROWS + tid] = ((red[tid] + red[TC128_ROWS + tid]) + (red[2 * TC128_ROWS + tid] + red[3 * TC128_ROWS + tid])) / (float)Lr.N;
            named_bar(1, TC_EPI_THREADS);
            {
              float d[64];
#pragma unroll
              for (int m = 0; m < 64; ++m) {
                const float dl = valid ? v[m] - red[8 * TC128_ROWS + mbase + m] : 0.f;
                d[m] = dl * dl;
              }
              warp_transpose_sum64(d, lane, s0, s1);
            }
            red[(4 + q) * TC128_ROWS + mbase + 2 * lane] = s0;
            red[(4 + q) * TC128_ROWS + mbase + 2 * lane + 1] = s1;
            named_bar(1, TC_EPI_THREADS);
            if (tid < TC128_ROWS) {
              const float var = ((red[4 * TC128_ROWS + tid] + red[5 * TC128_ROWS + tid]) + (red[6 * TC128_ROWS + tid] + red[7 * TC128_ROWS + tid])) / (float)Lr.N;
              red[9 * TC128_ROWS + tid] = 1.0f / sqrtf(var + a.ln_eps);
            }
            named_bar(1, TC_EPI_THREADS);
            if (valid) {
#pragma unroll
              for (int m = 0; m < 64; ++m) {
                const float nrm = (v[m] - red[8 * TC128_ROWS + mbase + m]) * red[9 * TC128_ROWS + mbase + m] * c2f + c3f;
                b_store(b_hi, b_lo, boff + m * 16, fmaxf(nrm, 0.f));
              }
            }
          }
          if (lane_ok && f >= Lr.N && f < next_K && !split_out) {   // zero padding of the next contraction range
#pragma unroll
            for (int m = 0; m < 64; ++m) b_store(b_hi, b_lo, boff + m * 16, 0.f);
          }
        }
        tc_fence_before_sync();
        if (l + 1 < a.nl) {
          fence_proxy_async_smem();
          mbar_arrive(&S->act_ready);
        }
        TC_CLK(e1);
        t_epi_l[l] += e1 - e0;
      }

      if (a.do_filter) {
        named_bar(1, TC_EPI_THREADS);
        if (tid < TC128_ROWS) {
          const int m = tid;
          bool keep = false;
          if (m < nrows) {
            keep = filter_decide([&](int k) { return lg[k * TC128_ROWS + m]; }, a.Kc, a.label, a.thr);
            if (a.keep_out) a.keep_out[row0 + m] = keep ? 1 : 0;
            if (a.logits_out)
              for (int k = 0; k < a.Kc; ++k) a.logits_out[(size_t)(row0 + m) * a.Kc + k] = lg[k * TC128_ROWS + m];
          }
          const unsigned bal = __ballot_sync(0xffffffffu, keep);
          if (lane == 0) S->warp_cnt[warp] = __popc(bal);
          named_bar(2, TC128_ROWS);
          if (tid == 0) {
            int tot = 0;
            for (int k = 0; k < 4; ++k) { const int c = S->warp_cnt[k]; S->warp_cnt[k] = tot; tot += c; }
            S->base = tot ? atomicAdd(a.count, (unsigned long long)tot) : 0ull;
          }
          named_bar(2, TC128_ROWS);
          if (keep) {
            const long long pos = (long long)S->base + S->warp_cnt[warp] + __popc(bal & ((1u << lane) - 1u));
            if (pos < a.capacity) {
              for (int f = 0; f < a.F; ++f) a.x_out[pos * a.F + f] = xs[f * TC128_XS_PITCH + m];
              if (a.idx_out) a.idx_out[pos] = (long long)(a.row_offset + (unsigned long long)(row0 + m));
            }
          }
        }
      }
      named_bar(1, TC_EPI_THREADS);
    }
    if (prof && tid == 0) {
      t_all = clock64() - t_all;
      atomicAdd((unsigned long long*)a.dbg + 4, (unsigned long long)t_acc);
      atomicAdd((unsigned long long*)a.dbg + 5, (unsigned long long)t_in);
      atomicAdd((unsigned long long*)a.dbg + 6, (unsigned long long)t_all);
      atomicAdd((unsigned long long*)a.dbg + 16, (unsigned long long)t_pf);
      for (int l = 0; l < a.nl; ++l) atomicAdd((unsigned long long*)a.dbg + 18 + l, (unsigned long long)t_epi_l[l]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == TC_EPI_WARPS) tmem_dealloc(tmem, 512);
}

inline size_t tc_eval128_smem() {
  return 2 * (size_t)TC128_BBYTES + (size_t)TC_STAGES * TC_STAGE_BYTES +
         sizeof(float) * ((size_t)TC128_MAXF * TC128_XS_PITCH + (size_t)TC128_MAXKC * TC128_ROWS + 10 * TC128_ROWS) +
         sizeof(Tc128Smem) + 64;
}

}  // namespace cvg
