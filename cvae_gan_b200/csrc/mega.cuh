// cvaegan_b200 - the training "step program" kernel (sm_100a): ONE persistent cooperative kernel executes a whole
// optimiser step - or a whole label visit of 13 steps - of cvae_gan.py:104-216.
//
//   * The host records the step as a PROGRAM of ops (train.cu emits the same GemmArgs / DwArgs / LnArgs ... records
//     it used to launch one kernel each for); one CTA per SM walks the program, takes the work items of every op
//     round-robin, and a grid barrier replaces each kernel boundary (~1.5 us instead of a launch + drain + ramp).
//   * Every GEMM (forward, input-gradient, weight-gradient) runs on the 5th-generation tensor cores:
//     tcgen05.mma kind::tf32 with the 3xTF32 split (fp32-level accuracy), accumulators in TMEM, both operands staged
//     into shared memory in the K-major no-swizzle core-matrix layout of tc05.cuh by all 512 threads - the operand
//     transforms (BatchNorm + LeakyReLU of the producing layer, BatchNorm backward) are applied in registers on the
//     way, so normalised activations never exist in memory.  Orientation: D[feature][row] (forward / input gradient:
//     MMA M = 128 output features, N = 64 or 128 batch rows) and D[out][in] (weight gradient: M = 128, N = in-features,
//     K = batch rows).  A two-stage ring lets the asynchronous MMAs of chunk c overlap the staging of chunk c + 1.
//   * Weight gradients are DETERMINISTIC: every (layer tile, row slice) writes its partial to a scratch slot and a
//     reduce op sums the slots in a fixed order (no float atomics).
//   * Data parallel: the BatchNorm-moment and gradient exchanges are ops of the same program (LL packets over NVLink
//     peer memory, comm_nvl.cuh) - no extra launches.
//
// Activations stay in the feature-major workspace of gemm.cuh ([features][ld], L2 resident), so every op has the
// same inputs and outputs as the stand-alone kernel it replaces: the FFMA kernels remain as the A/B reference
// (CVG_TRAIN_MODE=ffma) and tests compare the two paths buffer by buffer.
#pragma once
#include "engine.cuh"
#include "tc05.cuh"
#include <type_traits>

namespace cvg {
namespace mk {
using namespace tc;

constexpr int THREADS = 320;                  // 8 worker warps + MMA issuer warp + weight producer warp
constexpr int WORKERS = 256;
constexpr int ISSUER_WARP = 8;
constexpr int PRODUCER_WARP = 9;
constexpr int KC = 32;                        // contraction values per pipeline stage
constexpr int KG = KC / 4;                    // groups of 4 contraction values (one 16-byte core-matrix row)
constexpr int LBO_A = 128 * 16;               // bytes between k-groups of a pre-split weight chunk (written by TMA)
constexpr int LBO_AP = 128 * 16 + 16;         // ... of an A operand staged by threads (weight gradient), padded: conflict-free stores
constexpr int A_PLANE = KG * LBO_AP;
constexpr int WPLANE_FLOATS = KG * LBO_A / 4;      // one plane (hi or lo) of a prepped weight chunk: 128 rows x 32 k
constexpr int CHUNK_FLOATS = 2 * WPLANE_FLOATS;
constexpr int B_MAXN = 256;                   // widest MMA N
constexpr int LBO_B_MAX = B_MAXN * 16 + 16;
constexpr int B_PLANE = KG * LBO_B_MAX;
constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;   // hi + lo planes of both operands
constexpr int NSTAGE = 2;
// Shared-memory pool (2 * STAGE_BYTES).  Weight-gradient items use it as two stages {P hi, P lo, Q hi, Q lo}; forward /
// input-gradient items as 3 weight slots (TMA), 2 operand slots (hi + lo planes of up to 128 batch rows) and a ring of
// raw activation rows (TMA).  The epilogue's transpose scratch sits inside operand slot 1 in both maps.
constexpr int MN_A_SLOT = 2 * KG * 128 * 16;                   // 32768
constexpr int MN_B_OFF = 3 * MN_A_SLOT;                        // 98304
constexpr int MN_B_PLANE = KG * (128 * 16 + 16);               // 16512
constexpr int MN_B_SLOT = 2 * MN_B_PLANE;                      // 33024
constexpr int MN_RAW_OFF = MN_B_OFF + 2 * MN_B_SLOT;           // 164352
constexpr int TP_OFF = 131840;                                 // 8 warps x 32 x 20 floats
constexpr int TP_BYTES = 8 * 32 * 20 * 4;
constexpr int CS_FLOATS = 2560;               // per-feature constants of the operand transforms / epilogues (C <= 256)
constexpr int RED_DOUBLES = 1024;
constexpr int TMEM_COLS = 256;
constexpr int OP_BYTES = 512;
constexpr int MAX_C = 256;                    // widest BatchNorm layer the program kernel handles

enum {
  K_MN = 1, K_DW, K_DWRED, K_FILL, K_STAGE, K_SN_POWER, K_SN_DOT, K_SN_GRAD, K_LN_FWD, K_LN_BWD, K_CE, K_SEED, K_ZERO,
  K_PACK, K_UNPACK, K_ADAM, K_CTL_SET, K_FINISH, K_NVL_F32, K_NVL_F64, K_REPARAM, K_PREP
};

struct alignas(16) OpRec {
  int kind;
  int bar_before;     // grid barrier before this op (it reads what earlier ops wrote)
  int items;          // independent work items
  int first;          // offset of the op's items within its phase (rotates the CTA assignment)
  int aux[4];
  unsigned char payload[OP_BYTES - 32];
};
static_assert(sizeof(OpRec) == OP_BYTES, "op record size");

// ---- payloads that have no stand-alone kernel argument struct -------------------------------------------------------
struct DwRedArgs {
  const float* part; const float* bpart;     // [nz][N][Kp], [nz][N]
  int nz, nsplit, npass;
  int N, K, Kp;
  float* dW; long long sdW; int ldw, wcol0;
  float* db; int label_col;
  float* dgamma; float* dbeta; const double* bstats; long long sb; int C; int add_affine;
};
struct StageArgs {            // _get_target_samples (optional) + transpose to the feature-major workspace
  const float* src;           // row-major [.][F]: the batch itself, or the class table when sampling
  long long n_rows;           // > 0: draw rows from the class table (cvae_gan.py:247-260)
  long long B_global, draw_offset;
  int M, F, ld;
  const StepCtl* ctl; unsigned long long counter_off;
  float* xT;
};
struct ZeroArgs { void* p; long long bytes; };
struct PackArgs { const double* acc; float* tail; };
struct UnpackArgs { const float* tail; float* out; int kind; float Bg, F; float* tail_w; };
struct CtlSetArgs { StepCtl* ctl; unsigned long long seed, counter; int set_rng; float lambda_class; int set_lambda; };
struct FinishArgs { StepCtl* ctl; unsigned long long dcounter; int adam_inc[4]; int n_exchanges; };
struct NvlArgs { void* data; long long seg_len, seg_stride; int nseg; int exchange; };   // exchange: index within the program
struct ReparamArgs { const float* mu; const float* lv; const float* eps; float* out; int M, ld, Z; };
struct AdamOp { AdamArgs a; int t_off[2]; };

struct Params {
  const OpRec* ops;
  int nops;
  unsigned int* bar_counter;     // zeroed by the host before every launch
  NvlDev nvl;                    // world == 0: no peer exchanges in this program
  long long* dbg;                // optional per-op cycle counters [nops] (CTA 0)
  long long* prof;               // optional section counters of the GEMM items (CTA 0, development)
};

// mbarriers.  Slot ids (bit positions of Pipe::ub): A 0-2, B 3-4, RAW 5-8.
struct Ctrl {
  uint64_t a_full[3], a_free[3];       // pre-split weight chunks (TMA) / their MMAs have completed
  uint64_t b_full[2], done[2];         // activation operand staged by the workers / its MMAs have completed
  uint64_t raw_full[4], raw_free[4];   // raw activation rows (TMA) / the workers have consumed them
  unsigned int seq;                    // GEMM items of this CTA whose MMAs have all completed (gates the producer: pool reuse)
  uint32_t tmem_slot;
  uint32_t pad;
  unsigned long long nvl_epoch0;
};

constexpr size_t SMEM_BYTES = (size_t)NSTAGE * STAGE_BYTES + CS_FLOATS * 4 + RED_DOUBLES * 8 + OP_BYTES + sizeof(Ctrl) + 64;
constexpr int MN_RAW_BYTES = NSTAGE * STAGE_BYTES - MN_RAW_OFF;   // 33280
static_assert(MN_RAW_BYTES >= 32768, "raw ring");
static_assert(TP_OFF >= MN_B_OFF + MN_B_SLOT && TP_OFF + TP_BYTES <= MN_RAW_OFF, "transpose scratch inside operand slot 1");
static_assert(TP_OFF >= STAGE_BYTES + 2 * A_PLANE && TP_OFF + TP_BYTES <= STAGE_BYTES + 2 * A_PLANE + B_PLANE, "transpose scratch inside Q-hi of stage 1");

// Per-CTA pipeline state.  Everything here lives in registers (all users are force-inlined): with ~215 KB of shared
// memory the L1 has ~10 KB left, so a single spilled word costs an L2 round trip.
//
// Warp roles inside a GEMM item (the tcgen05 issue loop blocks for as long as the tensor pipe is busy, so it must not
// sit in a warp that also stages operands):
//   warps 0-7  workers : stage the activation operand(s) (global -> registers -> transform -> hi/lo split -> shared memory),
//                        then run the epilogue (TMEM -> registers -> shared-memory transpose -> coalesced global I/O)
//   warp 8     issuer  : waits for a stage to be full, issues its MMAs, commits them to the stage's "done" barrier
//   warp 9     producer: streams the pre-split weight chunks of forward / input-gradient GEMMs with 1-D bulk copies (TMA)
// Stage hand-off is mbarrier-only (full[s]: 256 worker arrivals + the producer's expect_tx; done[s]: tcgen05.commit).
struct Pipe {
  uint8_t* stages;
  float* cs;
  double* red;
  Ctrl* S;
  uint32_t tmem;
  uint32_t ub;          // per slot id: bit i = uses so far mod 2, bit 16 + i = used at least once (every role counts identically)
  uint32_t nitem;       // GEMM items this CTA has started
  int tid, warp, lane;
  long long* prof;
};
#define MK_T(var) if (c.prof) var = clock64();
#define MK_ACC(slot, t0, t1) if (c.prof && c.tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(c.prof) + (slot), (unsigned long long)((t1) - (t0)));

__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---- loads of data that other CTAs produced earlier in the SAME launch: L2 only (L1 is not coherent) ---------------
__device__ __forceinline__ float ldg1(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ double ldgd(const double* p) { return __ldcg(p); }

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of the (cooperative, co-resident) grid; a CTA that never arrives makes the others trap instead of hanging
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target) {
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    while (ld_acquire_gpu(counter) < target) {
      if (clock64() - t0 > 20000000000ll) __trap();
    }
    __threadfence();
  }
  __syncthreads();
}

// ---- BatchNorm constants (same arithmetic as gemm.cuh, statistics read through L2); worker threads only -------------
__device__ __forceinline__ void mk_bn_mean_rstd(const BnRef& bn, int pass, int c, float Bg, float eps, float& mean, float& rstd) {
  float var;
  if (bn.eval) {
    mean = ldg1(bn.rmean + c);
    var = ldg1(bn.rvar + c);
  } else {
    const double* s = bn.fstats + (long long)pass * bn.sf;
    const double inv = 1.0 / (double)Bg;
    const double m = ldgd(s + c) * inv;
    double v = ldgd(s + bn.C + c) * inv - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
  }
  rstd = 1.0f / sqrtf(var + eps);
}

template <int KIND>
__device__ __forceinline__ void mk_operand_consts(const Operand& o, int pass, float Bg, float eps, float* cs, int tid) {
  if (KIND == OP_BN_ACT) {
    const int C = o.bn.C;
    for (int c = tid; c < C; c += WORKERS) {
      float mean, rstd;
      mk_bn_mean_rstd(o.bn, pass, c, Bg, eps, mean, rstd);
      cs[c] = ldg1(o.bn.gamma + c) * rstd;
      cs[C + c] = ldg1(o.bn.beta + c);
      cs[2 * C + c] = mean;
    }
  } else if (KIND == OP_BN_BWD) {
    const int C = o.bn.C;
    const double* bs = o.bn.bstats + (long long)pass * o.bn.sb;
    for (int c = tid; c < C; c += WORKERS) {
      float mean, rstd;
      mk_bn_mean_rstd(o.bn, pass, c, Bg, eps, mean, rstd);
      cs[c] = ldg1(o.bn.gamma + c) * rstd;
      cs[C + c] = (float)(ldgd(bs + c) / (double)Bg);
      cs[2 * C + c] = (float)(ldgd(bs + C + c) / (double)Bg);
      cs[3 * C + c] = mean;
      cs[4 * C + c] = rstd;
    }
  }
}

__device__ __forceinline__ void mk_bn_update_running(const BnRef& bn, int npass, float Bg, float momentum, int tid) {
  for (int c = tid; c < bn.C; c += WORKERS) {
    float rm = ldg1(bn.rmean + c), rv = ldg1(bn.rvar + c);
    for (int p = 0; p < npass; ++p) {
      const double* s = bn.fstats + (long long)p * bn.sf;
      const double m = ldgd(s + c) / (double)Bg;
      double v = ldgd(s + bn.C + c) / (double)Bg - m * m;
      if (v < 0.0) v = 0.0;
      const double unb = v * ((double)Bg / ((double)Bg - 1.0));
      rm = (1.0f - momentum) * rm + momentum * (float)m;
      rv = (1.0f - momentum) * rv + momentum * (float)unb;
    }
    bn.rmean[c] = rm;
    bn.rvar[c] = rv;
  }
}

// transform of ONE operand value of feature row r (the constants of the row come from shared memory)
template <int KIND>
__device__ __forceinline__ float mk_xform(float a, float b, const float* cs, int C, int r, float slope, float cst) {
  if (KIND == OP_BN_ACT) return act_lrelu(fmaf(a - cs[2 * C + r], cs[r], cs[C + r]), slope);
  if (KIND == OP_BN_BWD) return cs[r] * (a - cs[C + r] - (b - cs[3 * C + r]) * cs[4 * C + r] * cs[2 * C + r]);
  if (KIND == OP_CONST) return cst;
  return a;
}
// ... of FOUR consecutive feature rows r .. r + 3 (r % 4 == 0) of one batch row: the constants come as 16-byte loads.
// Values beyond the contraction length were loaded as zeros; BatchNorm kinds always have R == C (a multiple of 4).
template <int KIND>
__device__ __forceinline__ float4 mk_xform4(const float* a, const float* b, const float* cs, int C, int r, int R, float slope, float cst) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r >= R) return v;
  if (KIND == OP_BN_ACT) {
    const float4 sc = *reinterpret_cast<const float4*>(cs + r), sh = *reinterpret_cast<const float4*>(cs + C + r);
    const float4 mn = *reinterpret_cast<const float4*>(cs + 2 * C + r);
    v.x = act_lrelu(fmaf(a[0] - mn.x, sc.x, sh.x), slope);
    v.y = act_lrelu(fmaf(a[1] - mn.y, sc.y, sh.y), slope);
    v.z = act_lrelu(fmaf(a[2] - mn.z, sc.z, sh.z), slope);
    v.w = act_lrelu(fmaf(a[3] - mn.w, sc.w, sh.w), slope);
  } else if (KIND == OP_BN_BWD) {
    const float4 c1 = *reinterpret_cast<const float4*>(cs + r), c2 = *reinterpret_cast<const float4*>(cs + C + r);
    const float4 c3 = *reinterpret_cast<const float4*>(cs + 2 * C + r), mn = *reinterpret_cast<const float4*>(cs + 3 * C + r);
    const float4 rs = *reinterpret_cast<const float4*>(cs + 4 * C + r);
    v.x = c1.x * (a[0] - c2.x - (b[0] - mn.x) * rs.x * c3.x);
    v.y = c1.y * (a[1] - c2.y - (b[1] - mn.y) * rs.y * c3.y);
    v.z = c1.z * (a[2] - c2.z - (b[2] - mn.z) * rs.z * c3.z);
    v.w = c1.w * (a[3] - c2.w - (b[3] - mn.w) * rs.w * c3.w);
  } else if (KIND == OP_CONST) {
    v.x = cst;
    v.y = r + 1 < R ? cst : 0.f;
    v.z = r + 2 < R ? cst : 0.f;
    v.w = r + 3 < R ? cst : 0.f;
  } else {
    v = make_float4(a[0], a[1], a[2], a[3]);
  }
  return v;
}

// hi / lo planes of one operand value group: 16-byte stores
__device__ __forceinline__ void st_split4(uint8_t* hi_plane, uint8_t* lo_plane, uint32_t off, float4 v) {
  float4 h, l;
  split_tf32(v.x, h.x, l.x);
  split_tf32(v.y, h.y, l.y);
  split_tf32(v.z, h.z, l.z);
  split_tf32(v.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi_plane + off) = h;
  *reinterpret_cast<float4*>(lo_plane + off) = l;
}

// ---- slot bookkeeping: every role that touches a slot calls use_begin() once per chunk, so the mbarrier parities agree
// without communication.  Returns the parity on which THIS use completes; `first` = the slot has never been used.
enum { ID_A = 0, ID_B = 3, ID_RAW = 5 };
__device__ __forceinline__ uint32_t use_begin(Pipe& c, int id, bool& first) {
  const uint32_t par = (c.ub >> id) & 1u;
  first = ((c.ub >> (16 + id)) & 1u) == 0u;
  c.ub = (c.ub ^ (1u << id)) | (1u << (16 + id));
  return par;
}
// the release ("free" / "done") barrier of the PREVIOUS use of a slot, before it is overwritten
__device__ __forceinline__ void wait_prev_release(uint64_t* bar, uint32_t par, bool first) {
  if (!first) mbar_wait(bar, par ^ 1u);
}
// workers, end of an item: the MMAs of the newest use of operand slot s have completed (accumulator final)
__device__ __forceinline__ void wait_last_done(const Pipe& c, int s) {
  if ((c.ub >> (16 + ID_B + s)) & 1u) mbar_wait(&c.S->done[s], ((c.ub >> (ID_B + s)) & 1u) ^ 1u);
}

// issuer warp: MMAs of one staged chunk - nks steps of 8 contraction values, three tf32 MMAs each (small terms first);
// commits to `bar1` (and `bar2`).  Descriptor inputs are broadcast from lane 0 (uniform registers in the issue loop).
__device__ __forceinline__ void issue_mmas(const Pipe& c, uint32_t a_addr, uint32_t a_lo_off, uint32_t lbo_a, uint32_t b_addr,
                                           uint32_t b_lo_off, uint32_t lbo_b, int nks, int n_mma, bool first_chunk, uint64_t* bar1,
                                           uint64_t* bar2) {
  tc_fence_after_sync();
  const uint32_t aa = __shfl_sync(0xffffffffu, a_addr, 0), ab = __shfl_sync(0xffffffffu, b_addr, 0);
  const uint32_t alo = __shfl_sync(0xffffffffu, a_lo_off, 0), blo = __shfl_sync(0xffffffffu, b_lo_off, 0);
  const uint32_t lba = __shfl_sync(0xffffffffu, lbo_a, 0), lbb = __shfl_sync(0xffffffffu, lbo_b, 0);
  const uint32_t nm = __shfl_sync(0xffffffffu, (uint32_t)n_mma, 0);
  const uint32_t tm = __shfl_sync(0xffffffffu, c.tmem, 0);
  const int steps = __shfl_sync(0xffffffffu, nks, 0);
  const uint32_t fresh = __shfl_sync(0xffffffffu, first_chunk ? 1u : 0u, 0);
  if (elect_one()) {
    uint64_t dah = smem_desc(aa, lba, 128), dal = smem_desc(aa + alo, lba, 128);
    uint64_t dbh = smem_desc(ab, lbb, 128), dbl = smem_desc(ab + blo, lbb, 128);
    const uint32_t idesc = idesc_tf32(128, (int)nm, 0, 0);
    const uint64_t a_step = (uint64_t)((2 * lba) >> 4), b_step = (uint64_t)((2 * lbb) >> 4);
    for (int ks = 0; ks < steps; ++ks) {
      mma_tf32(tm, dal, dbh, idesc, !(fresh && ks == 0));
      mma_tf32(tm, dah, dbl, idesc, true);
      mma_tf32(tm, dah, dbh, idesc, true);
      dah += a_step; dal += a_step; dbh += b_step; dbl += b_step;
    }
    mma_commit(bar1);
    if (bar2) mma_commit(bar2);
  }
  __syncwarp();
}

// ---- epilogue transpose: a worker warp turns its TMEM block (lane = feature, 16 consecutive columns) into
// (feature = 8 fi + lane / 4, 4 consecutive columns at 4 (lane % 4)) so that every global access of the warp touches
// 64 contiguous bytes per feature instead of 16.  Scratch: 32 x 20 floats per warp (B planes of stage 1, idle now).
constexpr int TP_PITCH = 20;
__device__ __forceinline__ float* tp_scratch(const Pipe& c) {
  return reinterpret_cast<float*>(c.stages + TP_OFF) + c.warp * (32 * TP_PITCH);
}
__device__ __forceinline__ void tp_write(float* sc, int lane, const float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(sc + lane * TP_PITCH + i * 4) = make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
}
__device__ __forceinline__ float4 tp_read(const float* sc, int lane, int fi) {
  return *reinterpret_cast<const float4*>(sc + (fi * 8 + (lane >> 2)) * TP_PITCH + (lane & 3) * 4);
}

// ======================================================================================================================
// forward / input-gradient GEMM item:  D[n][m] = sum_r A[n][r] * B[m][r]
//   A = pre-split weights (prep op), streamed by the producer warp;  B = operand g.a (feature-major [r][m] in memory):
//   raw rows streamed by the producer warp (TMA), transformed + hi/lo split by the workers; m = batch rows of this tile.
// Kinds are run-time switches and the loops are real loops on purpose: a step touches every code path once or twice, so
// the instruction cache sees the code cold - compact code beats unrolled code here.
// ======================================================================================================================
struct MnGeo {
  int M, ld, N, R, pass, m0, n0, mt, nchunks, nt_shift, ak, raw_chunk, rs;
};

__device__ __forceinline__ float4 mk_xform4_rt(int ak, const float* a, const float* b, const float* cs, int C, int r, int R, float slope,
                                               float cst) {
  if (ak == OP_BN_ACT) return mk_xform4<OP_BN_ACT>(a, b, cs, C, r, R, slope, cst);
  if (ak == OP_BN_BWD) return mk_xform4<OP_BN_BWD>(a, b, cs, C, r, R, slope, cst);
  if (ak == OP_CONST) return mk_xform4<OP_CONST>(a, b, cs, C, r, R, slope, cst);
  return mk_xform4<OP_PLAIN>(a, b, cs, C, r, R, slope, cst);
}

// epilogue of one kind: thread -> features n0 + 32 q + 8 fi + lane / 4 (fi = 0..3), 4 batch rows at column 4 (lane % 4) of
// every 16-column block.  The memory operands of block cb + 1 are requested before block cb is processed.
template <int EK>
__device__ __forceinline__ void mn_epilogue(Pipe& c, const GemmArgs& g, const MnGeo& geo, const float* cs_e, float scale) {
  const int M = geo.M, ld = geo.ld, N = geo.N, pass = geo.pass, m0 = geo.m0, n0 = geo.n0;
  const int Nt = 1 << geo.nt_shift;
  const int q = c.warp & 3, cgp = c.warp >> 2;
  const int fl0 = q * 32 + (c.lane >> 2), rg4 = (c.lane & 3) * 4;
  const int colw = Nt >> 1, ncb = Nt >> 5;
  const float slope = g.slope, keep_inv = g.keep_inv;
  constexpr bool need_pv = EK == EP_DBN || EK == EP_DACT || EK == EP_STORE;
  constexpr bool need_mk = EK == EP_LINEAR || EK == EP_DACT;
  const uint8_t* mask_p = (need_mk && g.mask) ? g.mask + (long long)pass * g.smask : nullptr;
  const float* prev_p = (EK == EP_DBN || EK == EP_DACT) ? g.prev + (long long)pass * g.sprev : nullptr;
  float* Yp = g.Y + (long long)pass * g.sY;
  const bool acc_y = EK == EP_STORE && g.accumulate;

  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  float e_sc[4], e_sh[4], e_mean[4], e_rstd[4];
#pragma unroll
  for (int fi = 0; fi < 4; ++fi) {
    const int n = n0 + fl0 + fi * 8;
    if (EK == EP_LINEAR && n < N) {
      bias[fi] = g.bias ? ldg1(g.bias + n) : 0.f;
      if (g.wlabel) bias[fi] += scale * ldg1(g.wlabel + (size_t)n * g.ldwl);
    }
    if (EK == EP_DBN) {
      const int C = g.prev_bn.C;
      const int nn = min(n, C - 1);
      e_sc[fi] = cs_e[nn]; e_sh[fi] = cs_e[C + nn]; e_mean[fi] = cs_e[2 * C + nn]; e_rstd[fi] = cs_e[3 * C + nn];
    }
  }
  float4 pv[4], pvn[4];
  uchar4 mkv[4], mkn[4];
  auto fetch = [&](int cb, float4* pvx, uchar4* mkx) {
    const int m = m0 + cgp * colw + cb * 16 + rg4;
#pragma unroll
    for (int fi = 0; fi < 4; ++fi) {
      const int n = n0 + fl0 + fi * 8;
      pvx[fi] = make_float4(0.f, 0.f, 0.f, 0.f);
      mkx[fi] = make_uchar4(1, 1, 1, 1);
      if (n < N && m < M) {
        const size_t off = (size_t)n * ld + m;
        if (need_pv && prev_p) pvx[fi] = ldg4(prev_p + off);
        if (need_mk && mask_p) mkx[fi] = __ldcg(reinterpret_cast<const uchar4*>(mask_p + off));
        if (EK == EP_STORE && acc_y) pvx[fi] = ldg4(Yp + off);
      }
    }
  };
  long long e0 = 0, e1 = 0;
  MK_T(e0);
  fetch(0, pv, mkv);                               // the first two blocks: in flight while the last MMAs complete
  if (ncb > 1) fetch(1, pvn, mkn);
  MK_T(e1);
  MK_ACC(10, e0, e1);
  wait_last_done(c, (geo.nchunks - 1) & 1);        // tcgen05.commit covers every MMA issued before it
  tc_fence_after_sync();
  MK_T(e0);
  MK_ACC(11, e1, e0);

  double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
  double tot = 0.0, klsum = 0.0;
  float* tps = tp_scratch(c);
#pragma unroll 1
  for (int cb = 0; cb < ncb; ++cb) {
    const int col = cgp * colw + cb * 16;
    {
      float v[16];
      tmem_ld16(c.tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v);
      tmem_wait_ld();
      tp_write(tps, c.lane, v);
    }
    __syncwarp();
    MK_T(e1);
    MK_ACC(12, e0, e1);
    const int m = m0 + col + rg4;
#pragma unroll
    for (int fi = 0; fi < 4; ++fi) {
      const int n = n0 + fl0 + fi * 8;
      if (n >= N || m >= M) continue;
      const float4 yy = tp_read(tps, c.lane, fi);
      float y[4] = {yy.x, yy.y, yy.z, yy.w};
      const bool rowv[4] = {true, m + 1 < M, m + 2 < M, m + 3 < M};
      const size_t off = (size_t)n * ld + m;
      if (EK == EP_LINEAR) {
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = fmaf(y[i], scale, bias[fi]);
        if (g.kl_acc) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (rowv[i]) klsum += (n < g.kl_split) ? 0.5 * (double)y[i] * (double)y[i] : -0.5 * (1.0 + (double)y[i] - (double)expf(y[i]));
        }
        if (g.ostats) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (rowv[i]) { s1[fi] += (double)y[i]; s2[fi] += (double)y[i] * (double)y[i]; }
        }
        if (g.act == ACT_LRELU) {
#pragma unroll
          for (int i = 0; i < 4; ++i) y[i] = act_lrelu(y[i], slope);
        } else if (g.act == ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 4; ++i) y[i] = fmaxf(y[i], 0.f);
        } else if (g.act == ACT_SIGMOID) {
#pragma unroll
          for (int i = 0; i < 4; ++i) y[i] = 1.0f / (1.0f + expf(-y[i]));
        }
        if (mask_p) {
          y[0] = mkv[fi].x ? y[0] * keep_inv : 0.f;
          y[1] = mkv[fi].y ? y[1] * keep_inv : 0.f;
          y[2] = mkv[fi].z ? y[2] * keep_inv : 0.f;
          y[3] = mkv[fi].w ? y[3] * keep_inv : 0.f;
        }
        if (g.osum) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (rowv[i]) tot += (double)y[i];
        }
        st4(Yp + off, make_float4(y[0], y[1], y[2], y[3]));
      } else if (EK == EP_DBN) {
        const float hh[4] = {pv[fi].x, pv[fi].y, pv[fi].z, pv[fi].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float pre = fmaf(hh[i] - e_mean[fi], e_sc[fi], e_sh[fi]);
          const float dy = rowv[i] ? (pre > 0.f ? y[i] : y[i] * slope) : 0.f;
          y[i] = dy;
          s1[fi] += (double)dy;
          s2[fi] += (double)(dy * ((hh[i] - e_mean[fi]) * e_rstd[fi]));
        }
        st4(Yp + off, make_float4(y[0], y[1], y[2], y[3]));
      } else if (EK == EP_DACT) {
        const float aa[4] = {pv[fi].x, pv[fi].y, pv[fi].z, pv[fi].w};
        float keep[4] = {1.f, 1.f, 1.f, 1.f};
        if (mask_p) {
          keep[0] = mkv[fi].x ? keep_inv : 0.f;
          keep[1] = mkv[fi].y ? keep_inv : 0.f;
          keep[2] = mkv[fi].z ? keep_inv : 0.f;
          keep[3] = mkv[fi].w ? keep_inv : 0.f;
        }
        const float neg = (g.act == ACT_LRELU) ? slope : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float d = y[i] * scale * keep[i];
          y[i] = aa[i] > 0.f ? d : d * neg;
        }
        st4(Yp + off, make_float4(y[0], y[1], y[2], y[3]));
      } else if (EK == EP_STORE) {
        float4 o = make_float4(y[0] * scale, y[1] * scale, y[2] * scale, y[3] * scale);
        if (acc_y) { o.x += pv[fi].x; o.y += pv[fi].y; o.z += pv[fi].z; o.w += pv[fi].w; }
        st4(Yp + off, o);
      } else if (EK == EP_REPARAM_BWD) {
        const float4 mu = ldg4(g.mu + off), lv = ldg4(g.lv + off), ee4 = ldg4(g.eps + off);
        const float mm[4] = {mu.x, mu.y, mu.z, mu.w}, ll[4] = {lv.x, lv.y, lv.z, lv.w}, ee[4] = {ee4.x, ee4.y, ee4.z, ee4.w};
        float dmu[4], dlv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          dmu[i] = rowv[i] ? y[i] + g.kl_coef * mm[i] : 0.f;
          dlv[i] = rowv[i] ? y[i] * ee[i] * 0.5f * expf(0.5f * ll[i]) + g.kl_coef * 0.5f * (expf(ll[i]) - 1.0f) : 0.f;
        }
        st4(g.Y + off, make_float4(dmu[0], dmu[1], dmu[2], dmu[3]));
        st4(g.Y + (size_t)(N + n) * ld + m, make_float4(dlv[0], dlv[1], dlv[2], dlv[3]));
      }
    }
    __syncwarp();
#pragma unroll
    for (int fi = 0; fi < 4; ++fi) { pv[fi] = pvn[fi]; mkv[fi] = mkn[fi]; }
    if (cb + 2 < ncb) fetch(cb + 2, pvn, mkn);     // two blocks ahead
    MK_T(e0);
    MK_ACC(13, e1, e0);
  }
  tc_fence_before_sync();
  if ((EK == EP_LINEAR || EK == EP_DBN) && g.ostats) {
    // per feature: the 4 column lanes of a warp by shuffle, the two column groups through shared memory, then one
    // atomic pair per (feature, tile)
#pragma unroll
    for (int fi = 0; fi < 4; ++fi) {
      s1[fi] += __shfl_xor_sync(0xffffffffu, s1[fi], 1);
      s1[fi] += __shfl_xor_sync(0xffffffffu, s1[fi], 2);
      s2[fi] += __shfl_xor_sync(0xffffffffu, s2[fi], 1);
      s2[fi] += __shfl_xor_sync(0xffffffffu, s2[fi], 2);
    }
    if (cgp == 1 && (c.lane & 3) == 0) {
#pragma unroll
      for (int fi = 0; fi < 4; ++fi) {
        c.red[(fl0 + fi * 8) * 2] = s1[fi];
        c.red[(fl0 + fi * 8) * 2 + 1] = s2[fi];
      }
    }
    worker_bar();
    if (cgp == 0 && (c.lane & 3) == 0) {
      double* st = g.ostats + (long long)pass * g.sostats;
#pragma unroll
      for (int fi = 0; fi < 4; ++fi) {
        const int n = n0 + fl0 + fi * 8;
        if (n < N) {
          atomicAdd(st + n, s1[fi] + c.red[(fl0 + fi * 8) * 2]);
          atomicAdd(st + N + n, s2[fi] + c.red[(fl0 + fi * 8) * 2 + 1]);
        }
      }
    }
  }
  if (EK == EP_LINEAR && (g.osum || g.kl_acc)) {
    // one atomic per tile (the sums of the 8 worker warps meet in shared memory first)
    tot = warp_sum_d(tot);
    klsum = warp_sum_d(klsum);
    worker_bar();
    if (c.lane == 0) { c.red[512 + c.warp] = tot; c.red[528 + c.warp] = klsum; }
    worker_bar();
    if (c.tid == 0) {
      double t = 0.0, kk = 0.0;
      for (int w = 0; w < WORKERS / 32; ++w) { t += c.red[512 + w]; kk += c.red[528 + w]; }
      if (g.osum && t != 0.0) atomicAdd(g.osum + pass, t);
      if (g.kl_acc && kk != 0.0) atomicAdd(g.kl_acc, kk);
    }
  }
  MK_T(e1);
  MK_ACC(14, e0, e1);
}

__device__ __forceinline__ void mn_item(Pipe& c, const GemmArgs& g, const float* wprep, const int nt_shift, const int item) {
  MnGeo geo;
  geo.M = g.M; geo.ld = g.ld; geo.N = g.N;
  geo.R = min(g.R, g.a.rows);
  geo.nt_shift = nt_shift;
  geo.ak = g.a.kind;
  const int Nt = 1 << nt_shift;
  {
    const int ntm = (geo.M + Nt - 1) >> nt_shift, nmt = (geo.N + 127) >> 7;
    const int rt = item % ntm, t2 = item / ntm;
    geo.mt = t2 % nmt;
    geo.pass = g.only_pass >= 0 ? g.only_pass : t2 / nmt;
    geo.m0 = rt << nt_shift;
    geo.n0 = geo.mt << 7;
  }
  geo.nchunks = (geo.R + KC - 1) / KC;
  const int raw_planes = geo.ak == OP_BN_BWD ? 2 : 1;
  geo.raw_chunk = raw_planes * KC * Nt * 4;                                   // bytes of raw rows per chunk: 8 / 16 / 32 KB
  geo.rs = geo.ak == OP_CONST ? 1 : min(4, 32768 / geo.raw_chunk);            // raw ring slots: 4 / 2 / 1
  const int M = geo.M, ld = geo.ld, R = geo.R, pass = geo.pass, m0 = geo.m0, nchunks = geo.nchunks, ak = geo.ak, RS = geo.rs;
  const uint32_t lbo_b = (uint32_t)Nt * 16u + 16u;
  Ctrl* S = c.S;
  const uint32_t my_item = c.nitem++;

  // ---------------------------------------------------- producer ----------------------------------------------------
  // weights run 3 chunks ahead of the MMAs, raw activation rows RS chunks ahead of the workers
  if (c.warp == PRODUCER_WARP) {
    // every MMA of the previous GEMM items of this CTA has completed (a weight-gradient item uses the whole pool)
    for (uint32_t spin = 0; *reinterpret_cast<volatile unsigned int*>(&S->seq) < my_item; ++spin)
      if (spin > (1u << 28)) __trap();
    const float* wsrc = wprep + (size_t)geo.mt * nchunks * CHUNK_FLOATS;
    const int row_bytes = min(Nt, ld - m0) * 4;
    const float* rsrc = (ak == OP_CONST) ? nullptr : g.a.p + (long long)pass * g.a.sp + m0;
    const float* hsrc = (ak == OP_BN_BWD) ? g.a.h + (long long)pass * g.a.sh + m0 : nullptr;
    auto load_weights = [&](int ch) {
      const int sl = ch % 3;
      bool first;
      const uint32_t par = use_begin(c, ID_A + sl, first);
      wait_prev_release(&S->a_free[sl], par, first);
      if (elect_one()) {
        const int nk4 = 2 * ((min(KC, R - ch * KC) + 7) >> 3);
        uint8_t* dst = c.stages + (size_t)sl * MN_A_SLOT;
        const uint32_t bytes = (uint32_t)nk4 * LBO_A;
        mbar_arrive_expect_tx(&S->a_full[sl], 2 * bytes);
        bulk_g2s(dst, wsrc + (size_t)ch * CHUNK_FLOATS, bytes, &S->a_full[sl]);
        bulk_g2s(dst + MN_A_SLOT / 2, wsrc + (size_t)ch * CHUNK_FLOATS + WPLANE_FLOATS, bytes, &S->a_full[sl]);
      }
      __syncwarp();
    };
    auto load_raw = [&](int ch) {
      if (ak == OP_CONST) return;
      const int sl = ch % RS;
      bool first;
      const uint32_t par = use_begin(c, ID_RAW + sl, first);
      wait_prev_release(&S->raw_free[sl], par, first);
      const int r0 = ch * KC;
      const int nr = min(KC, R - r0);
      uint8_t* dst = c.stages + MN_RAW_OFF + (size_t)sl * geo.raw_chunk;
      if (c.lane == 0) mbar_arrive_expect_tx(&S->raw_full[sl], (uint32_t)(raw_planes * nr * row_bytes));
      __syncwarp();
      if (c.lane < nr) {      // one feature row per lane: Nt consecutive batch rows = one contiguous run in the feature-major workspace
        bulk_g2s(dst + (size_t)c.lane * Nt * 4, rsrc + (size_t)(r0 + c.lane) * ld, (uint32_t)row_bytes, &S->raw_full[sl]);
        if (ak == OP_BN_BWD) bulk_g2s(dst + (size_t)(KC + c.lane) * Nt * 4, hsrc + (size_t)(r0 + c.lane) * ld, (uint32_t)row_bytes, &S->raw_full[sl]);
      }
      __syncwarp();
    };
    for (int ch = 0; ch < min(RS, nchunks); ++ch) load_raw(ch);
    for (int ch = 0; ch < min(3, nchunks); ++ch) load_weights(ch);
    for (int ch = 0; ch < nchunks; ++ch) {
      if (ch + RS < nchunks) load_raw(ch + RS);
      if (ch + 3 < nchunks) load_weights(ch + 3);
    }
    return;
  }
  // ----------------------------------------------------- issuer -----------------------------------------------------
  if (c.warp == ISSUER_WARP) {
    const uint32_t pool = smem_u32(c.stages);
    for (int ch = 0; ch < nchunks; ++ch) {
      const int sa = ch % 3, sb = ch & 1;
      bool f1, f2;
      const uint32_t pa = use_begin(c, ID_A + sa, f1);
      const uint32_t pbb = use_begin(c, ID_B + sb, f2);
      const int nks = (min(KC, R - ch * KC) + 7) >> 3;
      mbar_wait(&S->a_full[sa], pa);
      mbar_wait(&S->b_full[sb], pbb);
      issue_mmas(c, pool + (uint32_t)sa * MN_A_SLOT, MN_A_SLOT / 2, LBO_A, pool + MN_B_OFF + (uint32_t)sb * MN_B_SLOT, MN_B_PLANE, lbo_b, nks,
                 Nt, ch == 0, &S->done[sb], &S->a_free[sa]);
    }
    return;
  }
  // ----------------------------------------------------- workers ----------------------------------------------------
  const int tid = c.tid;
  const float slope = g.slope;
  float* cs_a = c.cs;
  const int Ca = (ak == OP_BN_ACT || ak == OP_BN_BWD) ? g.a.bn.C : 0;
  float* cs_e = c.cs + (ak == OP_BN_ACT ? 3 * Ca : (ak == OP_BN_BWD ? 5 * Ca : 0));
  long long q0 = 0, q1 = 0, q2 = 0, q3 = 0;
  MK_T(q0);

  // per-feature constants
  if (ak == OP_BN_ACT) mk_operand_consts<OP_BN_ACT>(g.a, pass, g.Bg, g.bn_eps, cs_a, tid);
  else if (ak == OP_BN_BWD) mk_operand_consts<OP_BN_BWD>(g.a, pass, g.Bg, g.bn_eps, cs_a, tid);
  if (g.ekind == EP_DBN) {
    const int C = g.prev_bn.C;
    for (int ch = tid; ch < C; ch += WORKERS) {
      float mean, rstd;
      mk_bn_mean_rstd(g.prev_bn, pass, ch, g.Bg, g.bn_eps, mean, rstd);
      cs_e[ch] = ldg1(g.prev_bn.gamma + ch) * rstd;
      cs_e[C + ch] = ldg1(g.prev_bn.beta + ch);
      cs_e[2 * C + ch] = mean;
      cs_e[3 * C + ch] = rstd;
    }
  }
  if (ak == OP_BN_ACT && g.a.bn.update_running && !g.a.bn.eval && item == 0) mk_bn_update_running(g.a.bn, g.npass, g.Bg, g.momentum, tid);
  float scale = 1.0f;
  if (g.scale) scale = ldg1(g.scale + pass);
  worker_bar();
  MK_T(q1);
  MK_ACC(0, q0, q1);

  // operand staging: thread -> batch row ml0, k-groups k40 + j * k4step (j < nb)
  const int ml0 = tid & (Nt - 1), k40 = tid >> nt_shift;
  const int k4step = WORKERS >> nt_shift, nb = Nt >> 5;
  const bool row_ok = m0 + ml0 < M;
#pragma unroll 1
  for (int ch = 0; ch < nchunks; ++ch) {
    const int sb = ch & 1, sr = ch % RS;
    const int r0 = ch * KC;
    const int nk4 = 2 * ((min(KC, R - r0) + 7) >> 3);
    bool first;
    MK_T(q2);
    const float* raw = reinterpret_cast<const float*>(c.stages + MN_RAW_OFF + (size_t)sr * geo.raw_chunk);
    if (ak != OP_CONST) {
      const uint32_t pr = use_begin(c, ID_RAW + sr, first);
      mbar_wait(&S->raw_full[sr], pr);                    // the raw rows of this chunk have landed
    }
    const uint32_t pbb = use_begin(c, ID_B + sb, first);
    wait_prev_release(&S->done[sb], pbb, first);          // the MMAs that read this operand slot two chunks ago are done
    MK_T(q3);
    MK_ACC(1, q2, q3);
    uint8_t* base = c.stages + MN_B_OFF + (size_t)sb * MN_B_SLOT;
#pragma unroll 1
    for (int j = 0; j < nb; ++j) {
      const int k4 = k40 + j * k4step;
      if (k4 < nk4) {
        float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
        if (ak != OP_CONST) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (r0 + k4 * 4 + i < R) {
              a[i] = raw[(k4 * 4 + i) * Nt + ml0];
              if (ak == OP_BN_BWD) b[i] = raw[(KC + k4 * 4 + i) * Nt + ml0];
            }
          }
        }
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_ok) v = mk_xform4_rt(ak, a, b, cs_a, Ca, r0 + k4 * 4, R, slope, g.a.cst);
        st_split4(base, base + MN_B_PLANE, (uint32_t)k4 * lbo_b + (uint32_t)ml0 * 16u, v);
      }
    }
    MK_T(q2);
    MK_ACC(2, q3, q2);
    fence_proxy_async_smem();
    mbar_arrive(&S->b_full[sb]);
    if (ak != OP_CONST) mbar_arrive(&S->raw_free[sr]);
    MK_T(q3);
    MK_ACC(3, q2, q3);
  }

  MK_T(q2);
  switch (g.ekind) {
    case EP_LINEAR: mn_epilogue<EP_LINEAR>(c, g, geo, cs_e, scale); break;
    case EP_DBN: mn_epilogue<EP_DBN>(c, g, geo, cs_e, scale); break;
    case EP_DACT: mn_epilogue<EP_DACT>(c, g, geo, cs_e, scale); break;
    case EP_STORE: mn_epilogue<EP_STORE>(c, g, geo, cs_e, scale); break;
    default: mn_epilogue<EP_REPARAM_BWD>(c, g, geo, cs_e, scale); break;
  }
  if (tid == 0) *reinterpret_cast<volatile unsigned int*>(&S->seq) = my_item + 1u;
  if (c.prof && c.tid == 0) { atomicAdd(reinterpret_cast<unsigned long long*>(c.prof) + 8, (unsigned long long)nchunks); atomicAdd(reinterpret_cast<unsigned long long*>(c.prof) + 9, 1ull); }
  worker_bar();     // TMEM drained, constants / reduction scratch / transpose scratch free for the next item
  MK_T(q3);
  MK_ACC(6, q2, q3);
}

// ======================================================================================================================
// weight-gradient GEMM item:  part[z][n][k] = sum_{m in slice} P[n][m] * Q[k][m]      (+ bias partial sum_m P[n][m])
// Both operands are activations: the workers stage them (thread = feature x 4-row group); the producer idles.
// ======================================================================================================================
struct DwScratch { float* part; float* bpart; int nsplit; int kp; };

__device__ __forceinline__ float mk_xform_rt(int kind, float a, float b, const float* cs, int C, int r, float slope, float cst) {
  if (kind == OP_BN_ACT) return mk_xform<OP_BN_ACT>(a, b, cs, C, r, slope, cst);
  if (kind == OP_BN_BWD) return mk_xform<OP_BN_BWD>(a, b, cs, C, r, slope, cst);
  if (kind == OP_CONST) return cst;
  return a;
}

template <int PK, int QK>
__device__ __forceinline__ void dw_item(Pipe& c, const DwArgs& g, const DwScratch& sc, const int item) {
  const int tid = c.tid;
  const int M = g.M, ld = g.ld, N = g.N, K = g.K;
  const int nnt = (N + 127) >> 7;
  const int nt = item % nnt, z = item / nnt;
  const int pass = z / sc.nsplit, split = z % sc.nsplit;
  const int n0 = nt << 7;
  const int mbeg = split * g.rows_per_cta;
  const int mend = min(M, mbeg + g.rows_per_cta);
  const int npad = (K + 15) & ~15;
  const uint32_t lbo_b = (uint32_t)npad * 16u + 16u;
  const int nchunks = mend > mbeg ? (mend - mbeg + KC - 1) / KC : 0;
  Ctrl* S = c.S;
  const uint32_t my_item = c.nitem++;
  if (c.warp == PRODUCER_WARP) return;         // both operands are activations: nothing to stream
  if (c.warp == ISSUER_WARP) {
    const uint32_t pool = smem_u32(c.stages);
    for (int ch = 0; ch < nchunks; ++ch) {
      const int sb = ch & 1;
      bool first;
      const uint32_t pbb = use_begin(c, ID_B + sb, first);
      const int nks = (min(KC, mend - (mbeg + ch * KC)) + 7) >> 3;
      mbar_wait(&S->b_full[sb], pbb);
      const uint32_t st = pool + (uint32_t)sb * STAGE_BYTES;
      issue_mmas(c, st, A_PLANE, LBO_AP, st + 2 * A_PLANE, B_PLANE, lbo_b, nks, npad, ch == 0, &S->done[sb], nullptr);
    }
    return;
  }

  const float slope = g.slope;
  float* cs_p = c.cs;
  const int Cp = (PK == OP_BN_BWD) ? g.p.bn.C : 0;
  const int Cq = (QK == OP_BN_ACT) ? g.q.bn.C : 0;
  float* cs_q = c.cs + 5 * Cp;
  const int prows = min(N, g.p.rows), qrows = min(K, g.q.rows);

  // thread -> (feature rows f0 + 32 j, row group rgp): 8 lanes cover the 32 batch rows of a chunk for one feature
  const int rgp = tid & 7, f0 = tid >> 3;
  const size_t step32 = (size_t)32 * ld;
  const float* pp0 = (PK == OP_CONST) ? nullptr : g.p.p + (long long)pass * g.p.sp + (size_t)(n0 + f0) * ld + rgp * 4;
  const float* ph0 = (PK == OP_BN_BWD) ? g.p.h + (long long)pass * g.p.sh + (size_t)(n0 + f0) * ld + rgp * 4 : nullptr;
  const float* pq0 = g.q.p + (long long)pass * g.q.sp + (size_t)f0 * ld + rgp * 4;
  const int nq = (npad + 31) >> 5;    // Q feature groups of 32 this thread stages (K <= 256 -> at most 8)

  if (PK == OP_BN_BWD) mk_operand_consts<OP_BN_BWD>(g.p, pass, g.Bg, g.bn_eps, cs_p, tid);
  if (QK == OP_BN_ACT) mk_operand_consts<OP_BN_ACT>(g.q, pass, g.Bg, g.bn_eps, cs_q, tid);
  worker_bar();

  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int ch = 0; ch < nchunks; ++ch) {
    const int sb = ch & 1;
    const int mb = mbeg + ch * KC;
    const int m = mb + rgp * 4;
    const bool rows_ok = m < mend;
    float4 rp[4], rph[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      rp[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      rph[j] = rp[j];
      if (PK != OP_CONST && n0 + f0 + j * 32 < prows && rows_ok) {
        rp[j] = ldg4(pp0 + j * step32 + mb);
        if (PK == OP_BN_BWD) rph[j] = ldg4(ph0 + j * step32 + mb);
      }
    }
    float4 rq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      rq[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < nq && f0 + j * 32 < qrows && rows_ok) rq[j] = ldg4(pq0 + j * step32 + mb);
    }
    bool first;
    const uint32_t pbb = use_begin(c, ID_B + sb, first);
    wait_prev_release(&S->done[sb], pbb, first);
    uint8_t* base = c.stages + (size_t)sb * STAGE_BYTES;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int f = n0 + f0 + j * 32;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < prows && rows_ok) {
        v.x = mk_xform<PK>(rp[j].x, rph[j].x, cs_p, Cp, f, slope, g.p.cst);
        v.y = (m + 1 < mend) ? mk_xform<PK>(rp[j].y, rph[j].y, cs_p, Cp, f, slope, g.p.cst) : 0.f;
        v.z = (m + 2 < mend) ? mk_xform<PK>(rp[j].z, rph[j].z, cs_p, Cp, f, slope, g.p.cst) : 0.f;
        v.w = (m + 3 < mend) ? mk_xform<PK>(rp[j].w, rph[j].w, cs_p, Cp, f, slope, g.p.cst) : 0.f;
      }
      bsum[j] += (v.x + v.y) + (v.z + v.w);
      st_split4(base, base + A_PLANE, (uint32_t)rgp * LBO_AP + (uint32_t)(f0 + j * 32) * 16u, v);
    }
#pragma unroll 1
    for (int jq = 0; jq < nq; jq += 4) {
      float4 rn[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {        // the next sub-pass is requested before this one is transformed
        rn[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (jq + 4 + j < nq && f0 + (jq + 4 + j) * 32 < qrows && rows_ok) rn[j] = ldg4(pq0 + (jq + 4 + j) * step32 + mb);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int f = f0 + (jq + j) * 32;
        if (jq + j < nq && f < npad) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (f < qrows && rows_ok) {
            v.x = mk_xform<QK>(rq[j].x, 0.f, cs_q, Cq, f, slope, 0.f);
            v.y = (m + 1 < mend) ? mk_xform<QK>(rq[j].y, 0.f, cs_q, Cq, f, slope, 0.f) : 0.f;
            v.z = (m + 2 < mend) ? mk_xform<QK>(rq[j].z, 0.f, cs_q, Cq, f, slope, 0.f) : 0.f;
            v.w = (m + 3 < mend) ? mk_xform<QK>(rq[j].w, 0.f, cs_q, Cq, f, slope, 0.f) : 0.f;
          }
          st_split4(base + 2 * A_PLANE, base + 2 * A_PLANE + B_PLANE, (uint32_t)rgp * lbo_b + (uint32_t)f * 16u, v);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) rq[j] = rn[j];
    }
    fence_proxy_async_smem();
    mbar_arrive(&S->b_full[sb]);
  }
  if (nchunks > 0) wait_last_done(c, (nchunks - 1) & 1);
  tc_fence_after_sync();
  if (tid == 0) *reinterpret_cast<volatile unsigned int*>(&S->seq) = my_item + 1u;

  // ---- epilogue: the partial tile goes to its scratch slot (row pitch kp): 16 columns per TMEM load, transposed in shared
  // memory so that a warp writes 64 contiguous bytes per row
  const int q = c.warp & 3, cgp = c.warp >> 2;
  const int fl0 = q * 32 + (c.lane >> 2), rg4 = (c.lane & 3) * 4;
  float* tps = tp_scratch(c);
#pragma unroll 1
  for (int b = cgp; b < npad / 16; b += 2) {
    float v[16];
    if (nchunks > 0) {
      tmem_ld16(c.tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 16), v);
      tmem_wait_ld();
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
    }
    tp_write(tps, c.lane, v);
    __syncwarp();
#pragma unroll
    for (int fi = 0; fi < 4; ++fi) {
      const int n = n0 + fl0 + fi * 8;
      if (n < N) st4(sc.part + ((size_t)z * N + n) * sc.kp + b * 16 + rg4, tp_read(tps, c.lane, fi));
    }
    __syncwarp();
  }
  // bias partial: the 8 lanes that staged the 8 row groups of a feature
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float b = bsum[j];
    b += __shfl_xor_sync(0xffffffffu, b, 1);
    b += __shfl_xor_sync(0xffffffffu, b, 2);
    b += __shfl_xor_sync(0xffffffffu, b, 4);
    const int f = n0 + f0 + j * 32;
    if (rgp == 0 && f < N) sc.bpart[(size_t)z * N + f] = b;
  }
  tc_fence_before_sync();
  worker_bar();
}

// ---- weight preparation: hi / lo planes of every 128 x 32 chunk of a GEMM's A operand, in shared-memory order ----------
//   chunk (mt, kc) = [hi: 8 k-groups x 128 rows x 4][lo: same]; element (row, k) at (k / 4) * 512 + row * 4 + k % 4 floats
struct PrepEntry {
  const float* W;
  int ldw, wcol0;
  int wt;          // 1: A[n][r] = W[n][r] (forward)   0: A[n][r] = W[r][n] (input gradient)
  int R, N;        // contraction length, A rows
  long long off;   // floats into the prepped buffer
};
constexpr int PREP_MAX = 10;
struct PrepArgs {
  PrepEntry e[PREP_MAX];
  int n;
  float* wprep;
};
__device__ __forceinline__ int prep_chunks(const PrepEntry& e) { return ((e.N + 127) >> 7) * ((e.R + KC - 1) / KC); }

__device__ void prep_item(const PrepArgs& a, int item) {
  int ei = 0, ch = item;
  while (ei < a.n && ch >= prep_chunks(a.e[ei])) { ch -= prep_chunks(a.e[ei]); ++ei; }
  if (ei >= a.n) return;
  const PrepEntry e = a.e[ei];
  const int nkc = (e.R + KC - 1) / KC;
  const int mt = ch / nkc, kc = ch % nkc;
  float* dst = a.wprep + e.off + (size_t)ch * CHUNK_FLOATS;
  for (int i = threadIdx.x; i < WPLANE_FLOATS; i += THREADS) {
    const int k4 = i >> 9, row = (i >> 2) & 127, kk = i & 3;
    const int n = mt * 128 + row, r = kc * KC + k4 * 4 + kk;
    float w = 0.f;
    if (n < e.N && r < e.R) w = e.wt ? ldg1(e.W + (size_t)n * e.ldw + e.wcol0 + r) : ldg1(e.W + (size_t)r * e.ldw + e.wcol0 + n);
    float hi, lo;
    split_tf32(w, hi, lo);
    dst[i] = hi;
    dst[WPLANE_FLOATS + i] = lo;
  }
}

// sums the row-slice partials in a fixed order and adds them to the gradient buffer (deterministic)
__device__ void dwred_item(const DwRedArgs& a, int item) {
  const int per_item = 4 * THREADS;
  const int ndest = a.sdW != 0 ? a.npass : 1;
  const int zper = a.sdW != 0 ? a.nsplit : a.nz;
  const long long nel = (long long)a.N * a.K;
  for (int d = 0; d < ndest; ++d) {
    float* dW = a.dW + (long long)d * a.sdW;
    for (long long e = (long long)item * per_item + threadIdx.x; e < min(nel, (long long)(item + 1) * per_item); e += THREADS) {
      const int n = (int)(e / a.K), k = (int)(e % a.K);
      float s = 0.f;
      for (int z = d * zper; z < (d + 1) * zper; ++z) s += ldg1(a.part + ((size_t)z * a.N + n) * a.Kp + k);
      float* dst = dW + (size_t)n * a.ldw + a.wcol0 + k;
      *dst = ldg1(dst) + s;
    }
    if (item == 0 && a.label_col >= 0) {
      for (int n = threadIdx.x; n < a.N; n += THREADS) {
        float s = 0.f;
        for (int z = d * zper; z < (d + 1) * zper; ++z) s += ldg1(a.bpart + (size_t)z * a.N + n);
        float* dst = dW + (size_t)n * a.ldw + a.label_col;
        *dst = ldg1(dst) + s;
      }
    }
  }
  if (item == 0) {
    if (a.db) {
      for (int n = threadIdx.x; n < a.N; n += THREADS) {
        float s = 0.f;
        for (int z = 0; z < a.nz; ++z) s += ldg1(a.bpart + (size_t)z * a.N + n);
        a.db[n] = ldg1(a.db + n) + s;
      }
    }
    if (a.add_affine && a.dgamma) {
      for (int p = 0; p < a.npass; ++p) {
        const double* bs = a.bstats + (long long)p * a.sb;
        for (int ch = threadIdx.x; ch < a.C; ch += THREADS) {
          a.dbeta[ch] = ldg1(a.dbeta + ch) + (float)ldgd(bs + ch);
          a.dgamma[ch] = ldg1(a.dgamma + ch) + (float)ldgd(bs + a.C + ch);
        }
      }
    }
  }
}

// ======================================================================================================================
// row-wise / element-wise ops: the bodies of misc_kernels.cuh re-cut for 512-thread CTAs and grid-strided work items
// ======================================================================================================================
constexpr int MK_LN_F = 16;      // LayerNorm features per thread (C <= 128)
constexpr int FILL_VB = 64;       // work items per fill job
constexpr int ADAM_VB = 96;       // work items per Adam segment
constexpr int NVL_VB = 8;         // work items per exchange (each thread pushes / polls its own elements)

__device__ void fill_item(const FillArgs& a, int item) {
  const FillJob j = a.job[item / FILL_VB];
  const int vb = item % FILL_VB;
  const int ngroups = (j.nfeat + 3) >> 2;
  const long long total = (long long)j.npass * ngroups * a.M;
  const uint32_t keep_thr = (uint32_t)((double)a.keep_prob * 4294967296.0);
  const uint64_t seed = a.ctl ? __ldcg(&a.ctl->seed) : a.seed;
  const uint64_t counter = a.ctl ? __ldcg(&a.ctl->counter) + a.counter_off : a.counter;
  // (batch row, feature group, pass) of an element; all counts fit 32 bits (rows <= max_batch, groups <= 256, passes <= HOIST_MAX)
  const unsigned uM = (unsigned)a.M, uMG = uM * (unsigned)ngroups;
  for (unsigned t = (unsigned)vb * THREADS + threadIdx.x; t < (unsigned)total; t += (unsigned)FILL_VB * THREADS) {
    const int pass = (int)(t / uMG);
    const unsigned rem = t - (unsigned)pass * uMG;
    const int fg = (int)(rem / uM);
    const int m = (int)(rem - (unsigned)fg * uM);
    float vals[4] = {0.f, 0.f, 0.f, 0.f};
    uint8_t bits[4] = {0, 0, 0, 0};
    if (j.injected) {
      for (int i = 0; i < 4; ++i) {
        const int f = fg * 4 + i;
        if (f < j.nfeat) {
          const size_t src = ((size_t)pass * a.M + m) * j.nfeat + f;
          if (j.kind == 0) vals[i] = ((const float*)j.injected)[src];
          else bits[i] = ((const uint8_t*)j.injected)[src] ? 1 : 0;
        }
      }
    } else {
      const U4 r = j.cstep ? philox_at(seed, counter + (uint64_t)pass * j.cstep, (uint32_t)j.stream, 0u, a.row_base + (uint64_t)m, (uint32_t)fg)
                           : philox_at(seed, counter, (uint32_t)j.stream, (uint32_t)pass, a.row_base + (uint64_t)m, (uint32_t)fg);
      if (j.kind == 0) {
        box_muller(r.x, r.y, vals[0], vals[1]);
        box_muller(r.z, r.w, vals[2], vals[3]);
      } else {
        bits[0] = r.x < keep_thr; bits[1] = r.y < keep_thr; bits[2] = r.z < keep_thr; bits[3] = r.w < keep_thr;
      }
    }
    for (int i = 0; i < 4; ++i) {
      const int f = fg * 4 + i;
      if (f < j.nfeat) {
        const size_t dst = ((size_t)pass * j.nfeat + f) * a.ld + m;
        if (j.kind == 0) ((float*)j.out)[dst] = vals[i];
        else ((uint8_t*)j.out)[dst] = bits[i];
      }
    }
  }
}

// cvae_gan.py:247-260 (optional) + transpose of the batch to the feature-major workspace: thread = batch row
__device__ void stage_item(const StageArgs& a, int item) {
  const int il = item * THREADS + threadIdx.x;
  if (il >= a.M) return;
  long long r = il;
  if (a.n_rows > 0) {
    const uint64_t seed = __ldcg(&a.ctl->seed);
    const uint64_t counter = __ldcg(&a.ctl->counter) + a.counter_off;
    const long long n = a.n_rows, B = a.B_global;
    const long long i = a.draw_offset + il;
    if (n == B) {
      r = i;
    } else if (n < B) {
      const U4 u = philox_at(seed, counter, RS_SAMPLE, 0, (uint64_t)i, 0);
      const uint64_t w = ((uint64_t)u.x << 32) | u.y;
      r = (long long)(w % (uint64_t)n);
    } else {
      int bits = 1;
      while ((1ll << bits) < n) ++bits;
      const int half = (bits + 1) >> 1;
      uint32_t keys[6];
      const U4 k0 = philox_at(seed, counter, RS_SAMPLE, 1, 0, 0), k1 = philox_at(seed, counter, RS_SAMPLE, 1, 1, 0);
      keys[0] = k0.x; keys[1] = k0.y; keys[2] = k0.z; keys[3] = k0.w; keys[4] = k1.x; keys[5] = k1.y;
      uint64_t x = (uint64_t)i;
      do { x = feistel(x, half, keys); } while (x >= (uint64_t)n);
      r = (long long)x;
    }
  }
  for (int f = 0; f < a.F; ++f) a.xT[(size_t)f * a.ld + il] = a.src[(size_t)r * a.F + f];
}

__device__ void reparam_item(const ReparamArgs& a, int item) {
  const long long idx = (long long)item * THREADS + threadIdx.x;
  if (idx >= (long long)a.Z * a.ld) return;
  const int m = (int)(idx % a.ld);
  float v = 0.f;
  if (m < a.M) v = ldg1(a.mu + idx) + ldg1(a.eps + idx) * expf(0.5f * ldg1(a.lv + idx));
  a.out[idx] = v;
}

// LayerNorm forward: 64 batch rows per item; warp = (row half, feature group), lane = row
__device__ void ln_fwd_item(const LnArgs& g, int item, float* red /* [8][32] */) {
  const int nrb = (g.M + 31) / 32;
  const int pass = item / nrb, rb = item % nrb;
  const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) & 7;
  const bool worker = threadIdx.x < WORKERS;
  const int m = rb * 32 + lane;
  const bool valid = worker && m < g.M;
  const int fpt = (g.C + 7) / 8;
  const int c0 = w * fpt;
  const float* h = g.h + (long long)pass * g.sh + (valid ? m : 0);
  float* rd = red;
  float v[MK_LN_F];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MK_LN_F; ++i) {
    const int cc = c0 + i;
    v[i] = (i < fpt && cc < g.C && valid) ? ldg1(h + (size_t)cc * g.ld) : 0.f;
    s += v[i];
  }
  if (worker) rd[w * 32 + lane] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) tot += rd[k * 32 + lane];
  const float mean = tot / (float)g.C;
  __syncthreads();
  float qv = 0.f;
#pragma unroll
  for (int i = 0; i < MK_LN_F; ++i) {
    const int cc = c0 + i;
    if (i < fpt && cc < g.C) { const float d = v[i] - mean; qv = fmaf(d, d, qv); }
  }
  if (worker) rd[w * 32 + lane] = qv;
  __syncthreads();
  tot = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) tot += rd[k * 32 + lane];
  const float rstd = 1.0f / sqrtf(tot / (float)g.C + g.eps);
  if (valid) {
    float* a = g.a + (long long)pass * g.sa + m;
    const uint8_t* mk = g.mask ? g.mask + (long long)pass * g.smask + m : nullptr;
#pragma unroll
    for (int i = 0; i < MK_LN_F; ++i) {
      const int cc = c0 + i;
      if (i < fpt && cc < g.C) {
        float nv = (v[i] - mean) * rstd * ldg1(g.g + cc) + ldg1(g.b + cc);
        nv = fmaxf(nv, 0.f);
        if (mk) nv = __ldcg(mk + (size_t)cc * g.ld) ? nv * g.keep_inv : 0.f;
        a[(size_t)cc * g.ld] = nv;
      }
    }
    if (g.rs && w == 0) {
      float* rs = g.rs + (long long)pass * g.srs;
      rs[m] = mean;
      rs[g.ld + m] = rstd;
    }
  }
  __syncthreads();
}

// LayerNorm backward; the affine gradients of an item go to a scratch slot (summed in order by the last... see host)
__device__ void ln_bwd_item(const LnBwdArgs& g, int item, float* red /* [2][8][32] */) {
  const int nrb = (g.M + 31) / 32;
  const int pass = item / nrb, rb = item % nrb;
  const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) & 7;
  const bool worker = threadIdx.x < WORKERS;
  const int m = rb * 32 + lane;
  const bool valid = worker && m < g.M;
  const int mm = valid ? m : 0;
  const int fpt = (g.C + 7) / 8;
  const int c0 = w * fpt;
  const float* h = g.h + (long long)pass * g.sh + mm;
  float* dn = g.dn + (long long)pass * g.sdn + mm;
  const float* rs = g.rs + (long long)pass * g.srs;
  const float mean = ldg1(rs + mm), rstd = ldg1(rs + g.ld + mm);
  float* r1 = red;
  float* r2 = r1 + 256;
  float xh[MK_LN_F], d[MK_LN_F];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < MK_LN_F; ++i) {
    const int cc = c0 + i;
    const bool on = i < fpt && cc < g.C && valid;
    xh[i] = on ? (ldg1(h + (size_t)cc * g.ld) - mean) * rstd : 0.f;
    d[i] = on ? ldg1(dn + (size_t)cc * g.ld) : 0.f;
    const float dx = on ? d[i] * ldg1(g.g + cc) : 0.f;
    s1 += dx;
    s2 = fmaf(dx, xh[i], s2);
  }
  if (worker) {
    r1[w * 32 + lane] = s1;
    r2[w * 32 + lane] = s2;
  }
  __syncthreads();
  s1 = 0.f; s2 = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { s1 += r1[k * 32 + lane]; s2 += r2[k * 32 + lane]; }
  s1 /= (float)g.C;
  s2 /= (float)g.C;
#pragma unroll
  for (int i = 0; i < MK_LN_F; ++i) {
    const int cc = c0 + i;
    const bool on = worker && i < fpt && cc < g.C;       // warp-uniform
    if (!on) continue;
    if (valid) dn[(size_t)cc * g.ld] = rstd * (d[i] * ldg1(g.g + cc) - s1 - xh[i] * s2);
    if (g.dg) {
      const float a = warp_sum(d[i] * xh[i]), b = warp_sum(d[i]);
      if (lane == 0) {
        atomicAdd(g.dg + cc, a);
        atomicAdd(g.db + cc, b);
      }
    }
  }
  __syncthreads();
}

__device__ void ce_item(const CeArgs& g, int item) {
  const int nrb = (g.M + THREADS - 1) / THREADS;
  const int pass = item / nrb;
  const int m = (item % nrb) * THREADS + threadIdx.x;
  double nll = 0.0;
  if (m < g.M) {
    const float* l = g.logits + (long long)pass * g.sl + m;
    float mx = -INFINITY;
    for (int k = 0; k < g.K; ++k) mx = fmaxf(mx, ldg1(l + (size_t)k * g.ld));
    float s = 0.f;
    for (int k = 0; k < g.K; ++k) s += expf(ldg1(l + (size_t)k * g.ld) - mx);
    const float lse = logf(s);
    const int tgt = g.labels ? (int)g.labels[m] : g.label;
    nll = -(double)(ldg1(l + (size_t)tgt * g.ld) - mx - lse);
    float* d = g.dlogits + (long long)pass * g.sd + m;
    const float coef = g.ctl ? g.coef * __ldcg(&g.ctl->lambda_class) : g.coef;
    for (int k = 0; k < g.K; ++k) {
      const float p = expf(ldg1(l + (size_t)k * g.ld) - mx - lse);
      d[(size_t)k * g.ld] = (p - (k == tgt ? 1.f : 0.f)) * coef;
    }
  }
  nll = warp_sum_d(nll);
  if ((threadIdx.x & 31) == 0 && nll != 0.0) atomicAdd(g.loss + pass, nll);
}

__device__ void seed_item(const SeedArgs& g, int item) {
  const int nblk = (int)(((long long)g.F * g.ld + THREADS - 1) / THREADS);
  const int pass = item / nblk;
  const int idx = (item % nblk) * THREADS + threadIdx.x;
  const int f = idx / g.ld, m = idx % g.ld;
  double sq = 0.0;
  if (f < g.F) {
    const size_t off = (size_t)f * g.ld + m;
    float d = 0.f;
    if (m < g.M) {
      const float o = ldg1(g.out + (long long)pass * g.sout + off);
      float dout;
      if (pass == 0) {
        const float diff = o - ldg1(g.x + off);
        sq = (double)diff * (double)diff;
        dout = g.coef_recon * 2.0f * diff;
      } else {
        dout = ldg1(g.dx + off);
      }
      d = dout * (1.0f - o) * o;
    }
    g.dpre[(long long)pass * g.sdpre + off] = d;
  }
  if (pass == 0) {
    sq = warp_sum_d(sq);
    if ((threadIdx.x & 31) == 0 && sq != 0.0) atomicAdd(g.recon_acc, sq);
  }
}

// spectral norm power iteration of one critic layer (item = layer), `npass` consecutive forwards.  W is copied to shared
// memory once (at most 128 x 256 floats = 128 KB of the pool); the three mat-vecs of every pass then run from there.
// sm: W at [0, rows * cols), then u, v, t (SN_MAXDIM each) and THREADS partials
__device__ void sn_power_item(const SnArgs& g, int item, float* sm, double* red) {
  const SnLayer L = g.L[item];
  const int n_el = L.rows * L.cols;
  float* Ws = sm;
  float* su = sm + ((n_el + 3) & ~3);
  float* sv = su + SN_MAXDIM;
  float* stt = sv + SN_MAXDIM;
  float* part = stt + SN_MAXDIM;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = THREADS >> 5;
  for (int i = tid; i < n_el; i += THREADS) Ws[i] = ldg1(L.W + i);
  for (int i = tid; i < L.rows; i += THREADS) su[i] = ldg1(L.u + i);
  for (int i = tid; i < L.cols; i += THREADS) sv[i] = ldg1(L.v + i);
  __syncthreads();
  int cp = 1;
  while (cp < L.cols) cp <<= 1;
  if (cp > 256) cp = 256;
  const int ng = 256 / cp, kq = tid % cp, gq = tid / cp;     // threads >= 256 idle in the column-parallel product
  for (int p = 0; p < g.npass; ++p) {
    if (g.do_power) {
      for (int n = w; n < L.rows; n += nw) {
        float s = 0.f;
        for (int k = lane; k < L.cols; k += 32) s = fmaf(Ws[n * L.cols + k], sv[k], s);
        s = warp_sum(s);
        if (lane == 0) stt[n] = s;
      }
      __syncthreads();
      double q = 0.0;
      for (int i = tid; i < L.rows; i += THREADS) q += (double)stt[i] * stt[i];
      q = block_sum_d(q, red);
      float nrm = fmaxf((float)sqrt(q), g.eps);
      for (int i = tid; i < L.rows; i += THREADS) su[i] = stt[i] / nrm;
      __syncthreads();
      for (int k0 = 0; k0 < L.cols; k0 += cp) {
        const int k = k0 + kq;
        float s = 0.f;
        if (tid < 256 && k < L.cols)
          for (int n = gq; n < L.rows; n += ng) s = fmaf(Ws[n * L.cols + k], su[n], s);
        if (tid < 256) part[tid] = s;
        __syncthreads();
        if (tid < 256 && gq == 0 && k < L.cols) {
          float t = 0.f;
          for (int j = 0; j < ng; ++j) t += part[j * cp + kq];
          stt[k] = t;
        }
        __syncthreads();
      }
      q = 0.0;
      for (int i = tid; i < L.cols; i += THREADS) q += (double)stt[i] * stt[i];
      q = block_sum_d(q, red);
      nrm = fmaxf((float)sqrt(q), g.eps);
      for (int i = tid; i < L.cols; i += THREADS) sv[i] = stt[i] / nrm;
      __syncthreads();
    }
    double sg = 0.0;
    for (int n = w; n < L.rows; n += nw) {
      float s = 0.f;
      for (int k = lane; k < L.cols; k += 32) s = fmaf(Ws[n * L.cols + k], sv[k], s);
      s = warp_sum(s);
      if (lane == 0) sg += (double)s * su[n];
    }
    sg = block_sum_d(sg, red);
    if (tid == 0) {
      g.sigma[p * 4 + item] = (float)sg;
      g.inv_sigma[item * 2 + p] = 1.0f / (float)sg;
    }
    float* us = g.u_snap + (long long)p * g.ssnap + L.snap_off;
    float* vs = g.v_snap + (long long)p * g.ssnap + L.snap_off;
    for (int i = tid; i < L.rows; i += THREADS) us[i] = su[i];
    for (int i = tid; i < L.cols; i += THREADS) vs[i] = sv[i];
    __syncthreads();
  }
  if (g.do_power) {
    for (int i = tid; i < L.rows; i += THREADS) L.u[i] = su[i];
    for (int i = tid; i < L.cols; i += THREADS) L.v[i] = sv[i];
  }
  __syncthreads();
}

constexpr int SN_DOT_VB = 8;
__device__ void sn_dot_item(const SnGradArgs& g, double* dots, int item, double* red) {
  const int vb = item % SN_DOT_VB, l = (item / SN_DOT_VB) % 4, p = item / (SN_DOT_VB * 4);
  const SnLayer L = g.L[l];
  const int n_el = L.rows * L.cols;
  const float* G = g.Gp + (long long)p * g.sG + g.w_off[l];
  double d = 0.0;
  for (int i = vb * THREADS + threadIdx.x; i < n_el; i += SN_DOT_VB * THREADS) d += (double)ldg1(G + i) * (double)ldg1(L.W + i);
  d = block_sum_d(d, red);
  if (threadIdx.x == 0 && d != 0.0) atomicAdd(dots + l * 2 + p, d);
  __syncthreads();
}

constexpr int SN_GRAD_VB = 16;
__device__ void sn_grad_item(const SnGradArgs& g, const double* dots, int item) {
  const int vb = item % SN_GRAD_VB, l = item / SN_GRAD_VB;
  const SnLayer L = g.L[l];
  const int n_el = L.rows * L.cols;
  for (int i = vb * THREADS + threadIdx.x; i < n_el; i += SN_GRAD_VB * THREADS) {
    const int n = i / L.cols, k = i % L.cols;
    float acc = 0.f;
    for (int p = 0; p < g.npass; ++p) {
      const float is = ldg1(g.inv_sigma + l * 2 + p);
      const float* G = g.Gp + (long long)p * g.sG + g.w_off[l];
      const float u = ldg1(g.u_snap + (long long)p * g.ssnap + L.snap_off + n);
      const float v = ldg1(g.v_snap + (long long)p * g.ssnap + L.snap_off + k);
      acc += ldg1(G + i) * is - (float)(ldgd(dots + l * 2 + p) * (double)is * (double)is) * u * v;
    }
    float* dst = g.grad + g.w_off[l] + i;
    *dst = ldg1(dst) + acc;
  }
  if (l == 3 && vb == 0 && threadIdx.x == 0 && g.last_bias_grad) *g.last_bias_grad = ldg1(g.last_bias_grad) + g.last_bias_value;
}

__device__ void adam_item(const AdamOp& op, int item) {
  __shared__ float sh[2];
  const AdamArgs& a = op.a;
  const int si = item / ADAM_VB, vb = item % ADAM_VB;
  const AdamSeg s = a.seg[si];
  __syncthreads();
  if (threadIdx.x == 0) {
    const double t = (double)(__ldcg(s.t_prev) + 1 + op.t_off[si]);
    sh[0] = (float)((double)s.lr / (1.0 - pow((double)a.b1, t)));
    sh[1] = (float)sqrt(1.0 - pow((double)a.b2, t));
  }
  __syncthreads();
  const float step_size = sh[0], bc2_sqrt = sh[1];
  const float w = 1.0f - a.b1;
  for (long long i = (long long)vb * THREADS + threadIdx.x; i < s.n; i += (long long)ADAM_VB * THREADS) {
    const float g = ldg1(s.g + i);
    float m = ldg1(s.m + i), v = ldg1(s.v + i);
    m = (w < 0.5f) ? m + w * (g - m) : g - (g - m) * (1.0f - w);
    v = v * a.b2 + (1.0f - a.b2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + a.eps;
    s.p[i] = ldg1(s.p + i) - step_size * (m / denom);
    s.m[i] = m;
    s.v[i] = v;
    if (a.clear_grad) s.g[i] = 0.f;
  }
}

__device__ void zero_item(const ZeroArgs& a, int item) {
  const long long n16 = a.bytes >> 4;
  uint4* p = reinterpret_cast<uint4*>(a.p);
  const long long per = (long long)THREADS * 8;
  for (long long i = (long long)item * per + threadIdx.x; i < min(n16, (long long)(item + 1) * per); i += THREADS) p[i] = make_uint4(0u, 0u, 0u, 0u);
  if (item == 0 && threadIdx.x < (a.bytes & 15)) reinterpret_cast<unsigned char*>(a.p)[(n16 << 4) + threadIdx.x] = 0;
}

// one exchange of the LL all-reduce of comm_nvl.cuh; ep = exchange number on every rank
template <typename T>
__device__ void nvl_item(const NvlDev& d, const NvlArgs& a, unsigned long long ep, int item) {
  nvl_exchange<T>(d, reinterpret_cast<T*>(a.data), a.seg_len, a.seg_stride, a.nseg, ep, (long long)item * THREADS + threadIdx.x,
                  (long long)NVL_VB * THREADS);
}

// ======================================================================================================================
// the program interpreter
// ======================================================================================================================
template <typename T>
__device__ __forceinline__ const T& payload(const OpRec* op) { return *reinterpret_cast<const T*>(op->payload); }

__global__ void __launch_bounds__(THREADS, 1) step_program_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Pipe c;
  c.stages = smem;
  c.cs = reinterpret_cast<float*>(smem + (size_t)NSTAGE * STAGE_BYTES);
  c.red = reinterpret_cast<double*>(c.cs + CS_FLOATS);
  OpRec* sop = reinterpret_cast<OpRec*>(c.red + RED_DOUBLES);
  Ctrl* S = reinterpret_cast<Ctrl*>(sop + 1);
  c.tid = threadIdx.x;
  c.warp = __shfl_sync(0xffffffffu, c.tid >> 5, 0);
  c.lane = c.tid & 31;
  c.ub = 0u;
  c.nitem = 0u;
  c.S = S;
  c.prof = blockIdx.x == 0 ? P.prof : nullptr;
  if (c.tid == 0) {
    for (int i2 = 0; i2 < 3; ++i2) { mbar_init(&S->a_full[i2], 1); mbar_init(&S->a_free[i2], 1); }
    for (int i2 = 0; i2 < 2; ++i2) { mbar_init(&S->b_full[i2], WORKERS); mbar_init(&S->done[i2], 1); }
    for (int i2 = 0; i2 < 4; ++i2) { mbar_init(&S->raw_full[i2], 1); mbar_init(&S->raw_free[i2], WORKERS); }
    fence_mbar_init();
    S->seq = 0u;
    S->nvl_epoch0 = P.nvl.world > 1 ? *reinterpret_cast<volatile unsigned long long*>(P.nvl.epoch) : 0ull;
  }
  if (c.warp == 0) tmem_alloc(&S->tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  c.tmem = __shfl_sync(0xffffffffu, S->tmem_slot, 0);

  const unsigned long long nvl_epoch0 = S->nvl_epoch0;
  unsigned int target = 0;
  const int G = (int)gridDim.x;
  long long t_op = 0;

  for (int oi = 0; oi < P.nops; ++oi) {
    __syncthreads();
    if (c.tid < OP_BYTES / 4) reinterpret_cast<uint32_t*>(sop)[c.tid] = reinterpret_cast<const uint32_t*>(P.ops + oi)[c.tid];
    __syncthreads();
    if (sop->bar_before) grid_barrier(P.bar_counter, target);
    if (P.dbg && blockIdx.x == 0 && c.tid == 0) { t_op = clock64(); P.dbg[1024 + oi] = t_op; }
    const int items = sop->items;
    const int i0 = (((int)blockIdx.x - sop->first) % G + G) % G;
    switch (sop->kind) {
      case K_MN: {
        const GemmArgs& g = payload<GemmArgs>(sop);
        const float* wp;
        memcpy(&wp, sop->payload + GEMM_ARGS_OP_BYTES, sizeof(float*));
        for (int it = i0; it < items; it += G) mn_item(c, g, wp, sop->aux[1] == 128 ? 7 : 6, it);
      } break;
      case K_DW: {
        const DwArgs& g = payload<DwArgs>(sop);
        DwScratch sc;
        memcpy(&sc.part, sop->payload + sizeof(DwArgs), sizeof(float*));
        memcpy(&sc.bpart, sop->payload + sizeof(DwArgs) + sizeof(float*), sizeof(float*));
        sc.nsplit = sop->aux[0];
        sc.kp = sop->aux[1];
        const int pk = g.p.kind, qk = g.q.kind;
        if (pk == OP_PLAIN && qk == OP_PLAIN) { for (int it = i0; it < items; it += G) dw_item<OP_PLAIN, OP_PLAIN>(c, g, sc, it); }
        else if (pk == OP_CONST && qk == OP_PLAIN) { for (int it = i0; it < items; it += G) dw_item<OP_CONST, OP_PLAIN>(c, g, sc, it); }
        else if (pk == OP_PLAIN && qk == OP_BN_ACT) { for (int it = i0; it < items; it += G) dw_item<OP_PLAIN, OP_BN_ACT>(c, g, sc, it); }
        else if (pk == OP_BN_BWD && qk == OP_BN_ACT) { for (int it = i0; it < items; it += G) dw_item<OP_BN_BWD, OP_BN_ACT>(c, g, sc, it); }
        else { for (int it = i0; it < items; it += G) dw_item<OP_BN_BWD, OP_PLAIN>(c, g, sc, it); }
      } break;
      case K_PREP: {
        const PrepArgs& a = payload<PrepArgs>(sop);
        for (int it = i0; it < items; it += G) prep_item(a, it);
      } break;
      case K_DWRED: {
        const DwRedArgs& a = payload<DwRedArgs>(sop);
        for (int it = i0; it < items; it += G) dwred_item(a, it);
      } break;
      case K_FILL: {
        const FillArgs& a = payload<FillArgs>(sop);
        for (int it = i0; it < items; it += G) fill_item(a, it);
      } break;
      case K_STAGE: {
        const StageArgs& a = payload<StageArgs>(sop);
        for (int it = i0; it < items; it += G) stage_item(a, it);
      } break;
      case K_REPARAM: {
        const ReparamArgs& a = payload<ReparamArgs>(sop);
        for (int it = i0; it < items; it += G) reparam_item(a, it);
      } break;
      case K_SN_POWER: {
        const SnArgs& a = payload<SnArgs>(sop);
        for (int it = i0; it < items; it += G) sn_power_item(a, it, reinterpret_cast<float*>(c.stages), c.red);
      } break;
      case K_SN_DOT: {
        const SnGradArgs& a = payload<SnGradArgs>(sop);
        double* dots;
        memcpy(&dots, sop->payload + sizeof(SnGradArgs), sizeof(double*));
        for (int it = i0; it < items; it += G) sn_dot_item(a, dots, it, c.red);
      } break;
      case K_SN_GRAD: {
        const SnGradArgs& a = payload<SnGradArgs>(sop);
        double* dots;
        memcpy(&dots, sop->payload + sizeof(SnGradArgs), sizeof(double*));
        for (int it = i0; it < items; it += G) sn_grad_item(a, dots, it);
      } break;
      case K_LN_FWD: {
        const LnArgs& a = payload<LnArgs>(sop);
        for (int it = i0; it < items; it += G) ln_fwd_item(a, it, c.cs);
      } break;
      case K_LN_BWD: {
        const LnBwdArgs& a = payload<LnBwdArgs>(sop);
        for (int it = i0; it < items; it += G) ln_bwd_item(a, it, c.cs);
      } break;
      case K_CE: {
        const CeArgs& a = payload<CeArgs>(sop);
        for (int it = i0; it < items; it += G) ce_item(a, it);
      } break;
      case K_SEED: {
        const SeedArgs& a = payload<SeedArgs>(sop);
        for (int it = i0; it < items; it += G) seed_item(a, it);
      } break;
      case K_ZERO: {
        const ZeroArgs& a = payload<ZeroArgs>(sop);
        for (int it = i0; it < items; it += G) zero_item(a, it);
      } break;
      case K_ADAM: {
        const AdamOp& a = payload<AdamOp>(sop);
        for (int it = i0; it < items; it += G) adam_item(a, it);
      } break;
      case K_PACK: {
        const PackArgs& a = payload<PackArgs>(sop);
        if (i0 == 0 && c.tid < L_COUNT) a.tail[c.tid] = (float)ldgd(a.acc + c.tid);
      } break;
      case K_UNPACK: {
        const UnpackArgs& a = payload<UnpackArgs>(sop);
        if (i0 == 0) {
          if (c.tid == 0) {
            const float* tail = a.tail;
            float* out = a.out;
            const float Bg = a.Bg;
            if (a.kind == 0) {
              const float r = ldg1(tail + L_DREAL) / Bg, f = ldg1(tail + L_DFAKE) / Bg;
              out[0] = -r + f; out[1] = r; out[2] = f; out[3] = 0.f;
            } else if (a.kind == 1) {
              const float r = ldg1(tail + L_CE0) / Bg, f = ldg1(tail + L_CE1) / Bg;
              out[0] = r + f; out[1] = r; out[2] = f; out[3] = 0.f;
            } else {
              out[0] = ldg1(tail + L_RECON) / (Bg * a.F);
              out[1] = ldg1(tail + L_KL) / Bg;
              out[2] = -ldg1(tail + L_DFAKE) / Bg;
              out[3] = ldg1(tail + L_CE0) / Bg;
            }
          }
          __syncthreads();
          if (c.tid < CVG_GRAD_TAIL) a.tail_w[c.tid] = 0.f;
        }
      } break;
      case K_CTL_SET: {
        const CtlSetArgs& a = payload<CtlSetArgs>(sop);
        if (i0 == 0 && c.tid == 0) {
          if (a.set_rng) { a.ctl->seed = a.seed; a.ctl->counter = a.counter; }
          if (a.set_lambda) a.ctl->lambda_class = a.lambda_class;
        }
      } break;
      case K_FINISH: {
        const FinishArgs& a = payload<FinishArgs>(sop);
        if (i0 == 0 && c.tid == 0) {
          a.ctl->counter = __ldcg(&a.ctl->counter) + a.dcounter;
          for (int n = 0; n < 4; ++n) a.ctl->adam_t[n] = __ldcg(&a.ctl->adam_t[n]) + a.adam_inc[n];
          if (P.nvl.world > 1 && a.n_exchanges > 0)
            *reinterpret_cast<volatile unsigned long long*>(P.nvl.epoch) = nvl_epoch0 + (unsigned long long)a.n_exchanges;
        }
      } break;
      case K_NVL_F32: {
        const NvlArgs& a = payload<NvlArgs>(sop);
        for (int it = i0; it < items; it += G) nvl_item<float>(P.nvl, a, nvl_epoch0 + 1ull + (unsigned long long)a.exchange, it);
      } break;
      case K_NVL_F64: {
        const NvlArgs& a = payload<NvlArgs>(sop);
        for (int it = i0; it < items; it += G) nvl_item<double>(P.nvl, a, nvl_epoch0 + 1ull + (unsigned long long)a.exchange, it);
      } break;
      default:
        break;
    }
    if (P.dbg && blockIdx.x == 0 && c.tid == 0) P.dbg[oi] = clock64() - t_op;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (c.warp == 0) tmem_dealloc(c.tmem, TMEM_COLS);
}

static_assert(GEMM_ARGS_OP_BYTES + sizeof(float*) <= OP_BYTES - 32, "GemmArgs does not fit an op record");
static_assert(sizeof(PrepArgs) <= OP_BYTES - 32, "PrepArgs does not fit an op record");
static_assert(sizeof(DwArgs) + 2 * sizeof(float*) <= OP_BYTES - 32, "DwArgs does not fit an op record");
static_assert(sizeof(FillArgs) <= OP_BYTES - 32, "FillArgs does not fit an op record");
static_assert(sizeof(SnArgs) <= OP_BYTES - 32, "SnArgs does not fit an op record");
static_assert(sizeof(SnGradArgs) + sizeof(double*) <= OP_BYTES - 32, "SnGradArgs does not fit an op record");
static_assert(sizeof(AdamOp) <= OP_BYTES - 32, "AdamOp does not fit an op record");

}  // namespace mk
}  // namespace cvg
