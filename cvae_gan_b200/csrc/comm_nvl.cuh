// cvaegan_b200 - small-message all-reduce over NVLink peer memory (one process per GPU, one node).
//
// The training path exchanges 66 BatchNorm moment vectors and 13 flat gradient buffers per label visit (SURVEY.md
// 8e).  They are latency bound: an NCCL all-reduce of a few KB costs ~30 us at 8 GPUs, which is a third of the visit.
// This is a one-shot all-reduce written for that regime:
//   * every rank owns a staging buffer [parity 2][source rank][slot] + flags, shared with its peers through CUDA IPC;
//   * exchange number `ep`: each thread stores its elements into slot[ep & 1][rank] of EVERY peer (plain stores to
//     peer-mapped addresses travel over NVLink / NVSwitch) as LL packets {4-byte word, ep} in aligned 8-byte stores;
//   * it then polls its own slots of all sources until every packet carries `ep`, and sums the `world` values IN RANK
//     ORDER, so every rank computes bit-identical sums (the replicated parameters stay bit-identical);
//   * parity double buffering is enough: a rank can only start exchange ep + 2 after it finished ep + 1, which needed
//     every peer's ep + 1 flags, which a peer writes after it has finished reading the slots of ep.
// The exchange number lives in device memory (advanced by the last CTA of each launch), so the launches are
// CUDA-graph capturable.  A peer that never arrives makes the wait trap after ~60 s instead of hanging the GPU
// (ranks legitimately drift by seconds around graph instantiation and host-side work).
#pragma once
#include "common.cuh"

namespace cvg {

constexpr int NVL_MAX_WORLD = 8;
constexpr int NVL_MAX_CTAS = 64;
constexpr int NVL_THREADS = 256;

struct NvlDev {
  int world, rank;
  unsigned char* peer[NVL_MAX_WORLD];   // base of every rank's staging buffer (peer[rank] = local)
  unsigned long long slot_bytes;        // payload capacity of one (parity, source) slot
  unsigned long long* epoch;            // local: number of completed exchanges
  unsigned int* done;                   // local: CTAs of the running launch that finished
};

struct NvlPending {                     // batch sums pushed by their producer, not yet taken up by a reader
  const double* stats;                  // slot of the first pushed pass
  int npass, pass0, channel;
};

struct NvlState {
  bool on = false;
  void* local = nullptr;
  size_t total_bytes = 0;
  NvlDev dev{};
  bool opened[NVL_MAX_WORLD] = {};
  // Channel 1: a second staging area with its own exchange numbering inside the same allocation, for the GEMM chain that
  // runs on a side stream beside the caller's (exchange numbers are per stream order, so two streams need two channels).
  // Channel 2: the gradient reduction + Adam of a step, which a label visit runs on side stream 0 beside the next step's head.
  NvlDev dev1{}, dev2{};
  NvlDev* dev_d[2] = {nullptr, nullptr};   // device copies (channels 0, 1) for the kernels that fold the exchange in
  bool fuse = true;                     // CVG_FUSE_STATS=0: BatchNorm sums go through nvl_allreduce_kernel launches
  bool via_lead = false;                // folded exchange: one CTA per pass polls for the launch (gemm.cuh bn_publish)
  NvlPending pending[8];
  int n_pending = 0;
};

__host__ __device__ inline unsigned long long nvl_slot_off(const NvlDev& d, int parity, int src) {
  return ((unsigned long long)parity * d.world + src) * d.slot_bytes;
}
__host__ __device__ inline unsigned long long nvl_flag_off(const NvlDev& d, int parity, int src, int cta) {
  return 2ull * d.world * d.slot_bytes + (((unsigned long long)parity * d.world + src) * NVL_MAX_CTAS + cta) * sizeof(unsigned int);
}
inline size_t nvl_total_bytes(int world, size_t slot_bytes) {
  return 2 * (size_t)world * slot_bytes + 2 * (size_t)world * NVL_MAX_CTAS * sizeof(unsigned int) + 256;
}

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// LL ("low latency") packets: every 4-byte payload word travels in an aligned 8-byte store {word, exchange number}.
// Aligned 8-byte stores are single-copy atomic, so the receiver polls the packet itself: no fence, no separate flag,
// one fabric traversal per exchange.
__device__ __forceinline__ void ll_store(unsigned char* p, unsigned int word, unsigned int ep) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(word), "r"(ep) : "memory");
}
__device__ __forceinline__ void ll_load(const unsigned char* p, unsigned int& word, unsigned int& ep) {
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(word), "=r"(ep) : "l"(p) : "memory");
}

// One exchange, the part every participating thread runs: elements first, first + step, ... of a vector of `nseg`
// segments of `seg_len` elements, `seg_stride` apart (BatchNorm moments live in per-pass slots of 2 * STAT_C doubles of
// which only the first 2 C entries are used; a contiguous vector is one segment).  U elements are in flight per thread:
// their loads (push: the local values; poll: all `world` packets of each) are issued before the first one is used, so a
// round of the loop costs one memory round trip whatever the world size.
template <typename T>
__device__ __forceinline__ void nvl_exchange(const NvlDev& d, T* __restrict__ data, long long seg_len, long long seg_stride,
                                             int nseg, unsigned long long ep, long long first, long long step) {
  constexpr int W = sizeof(T) / 4;
  constexpr int U = 4 / W;                      // 4 floats or 2 doubles
  const unsigned int ep32 = (unsigned int)ep;
  const int par = (int)(ep & 1ull);
  const int world = d.world;
  const long long n = seg_len * nseg;
  // 1. my elements -> slot[par][rank] of every peer (self included), packed (no segment gaps)
  for (long long e0 = first; e0 < n; e0 += U * step) {
    T x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long e = e0 + u * step;
      if (e < n) {
        const long long sg = e / seg_len, off = e - sg * seg_len;
        x[u] = __ldcg(data + sg * seg_stride + off);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long e = e0 + u * step;
      if (e < n) {
        unsigned int w[W];
        memcpy(w, &x[u], sizeof(T));
        for (int p = 0; p < world; ++p) {
          unsigned char* dst = d.peer[p] + nvl_slot_off(d, par, d.rank) + (unsigned long long)e * (W * 8);
#pragma unroll
          for (int k = 0; k < W; ++k) ll_store(dst + k * 8, w[k], ep32);
        }
      }
    }
  }
  // 2. poll the packets of every source and sum IN RANK ORDER (identical on every rank)
  const unsigned char* base = d.peer[d.rank];
  const long long t0 = clock64();
  for (long long e0 = first; e0 < n; e0 += U * step) {
    unsigned int w[U][NVL_MAX_WORLD][W];
    for (;;) {
      bool ok = true;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = e0 + u * step;
        if (e < n) {
#pragma unroll
          for (int q = 0; q < NVL_MAX_WORLD; ++q) {
            if (q < world) {
              const unsigned char* src = base + nvl_slot_off(d, par, q) + (unsigned long long)e * (W * 8);
#pragma unroll
              for (int k = 0; k < W; ++k) {
                unsigned int f;
                ll_load(src + k * 8, w[u][q][k], f);
                ok = ok && (f == ep32);
              }
            }
          }
        }
      }
      if (ok) break;
      if (clock64() - t0 > 120000000000ll) __trap();   // ~60 s: a peer never arrived
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long e = e0 + u * step;
      if (e < n) {
        T s = 0;
#pragma unroll
        for (int q = 0; q < NVL_MAX_WORLD; ++q) {
          if (q < world) {
            T x;
            memcpy(&x, w[u][q], sizeof(T));
            s += x;
          }
        }
        const long long sg = e / seg_len, off = e - sg * seg_len;
        data[sg * seg_stride + off] = s;
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(NVL_THREADS) nvl_allreduce_kernel(const NvlDev d, T* __restrict__ data, long long seg_len,
                                                                    long long seg_stride, int nseg) {
  __shared__ unsigned long long ep_s;
  const int tid = threadIdx.x;
  if (tid == 0) ep_s = *reinterpret_cast<volatile unsigned long long*>(d.epoch) + 1ull;
  __syncthreads();
  const unsigned long long ep = ep_s;
  nvl_exchange<T>(d, data, seg_len, seg_stride, nseg, ep, (long long)blockIdx.x * NVL_THREADS + tid,
                  (long long)gridDim.x * NVL_THREADS);
  // 3. the last CTA of the launch completes the exchange
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(d.done, 1u);
    if (prev == gridDim.x - 1) {
      *d.done = 0u;
      __threadfence();
      *reinterpret_cast<volatile unsigned long long*>(d.epoch) = ep;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The same exchange folded into the GEMM kernels either side of it (BatchNorm batch sums; gemm.cuh):
//   * the LAST CTA of the kernel that accumulated the local sums (found with the `done` counter) reads the finished
//     sums and stores them as LL packets into slot[ep & 1][rank] of every peer, then records `ep` next to the sums
//     and advances the epoch - no separate launch, and the packets fly while the next kernel is being launched;
//   * every CTA of the FIRST kernel that uses those sums polls its own slots for the features it needs and adds the
//     ranks' values in rank order (bit-identical on all ranks and to nvl_allreduce_kernel); one CTA writes the
//     global sums back for the later readers of the same statistics.
// The parity argument above holds unchanged: a rank's last CTA pushes ep + 1 only after all of its CTAs finished
// polling ep.
// ------------------------------------------------------------------------------------------------
// Two elements at once (a feature's pair of sums): all 4 * world loads are issued before the first flag is looked at,
// so a poll that finds its packets costs one L2 round trip whatever the world size.
__device__ __forceinline__ void nvl_poll_f64x2(const NvlDev& d, unsigned int ep32, int par, long long e0, long long e1,
                                               double& out0, double& out1) {
  const unsigned char* base = d.peer[d.rank];
  const int world = d.world;
  const unsigned long long off0 = (unsigned long long)e0 * 16, off1 = (unsigned long long)e1 * 16;
  unsigned int w0[NVL_MAX_WORLD][2], w1[NVL_MAX_WORLD][2];
  const long long t0 = clock64();
  for (;;) {
    bool ok = true;
#pragma unroll
    for (int q = 0; q < NVL_MAX_WORLD; ++q) {
      if (q < world) {
        const unsigned char* src = base + nvl_slot_off(d, par, q);
        unsigned int f0, f1, f2, f3;
        ll_load(src + off0, w0[q][0], f0);
        ll_load(src + off0 + 8, w0[q][1], f1);
        ll_load(src + off1, w1[q][0], f2);
        ll_load(src + off1 + 8, w1[q][1], f3);
        ok = ok && f0 == ep32 && f1 == ep32 && f2 == ep32 && f3 == ep32;
      }
    }
    if (ok) break;
    if (clock64() - t0 > 120000000000ll) __trap();   // ~60 s: a peer never arrived
  }
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int q = 0; q < NVL_MAX_WORLD; ++q) {
    if (q < world) {
      s0 += __hiloint2double((int)w0[q][1], (int)w0[q][0]);
      s1 += __hiloint2double((int)w1[q][1], (int)w1[q][0]);
    }
  }
  out0 = s0;
  out1 = s1;
}

// Called by EVERY thread of EVERY CTA as the last statement of a kernel that accumulated `npass` slots of 2 C doubles
// (`stride` doubles apart) into `stats` with atomics.  stats[stride - 1] receives the exchange number for the reader.
// `stats` = slot of the first pass of the launch.
__device__ __forceinline__ void nvl_push_stats_tail(const NvlDev& d, double* stats, long long stride, int npass, int C,
                                                    unsigned int total_ctas) {
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(d.done, 1u) == total_ctas - 1u) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const unsigned long long ep = *reinterpret_cast<volatile unsigned long long*>(d.epoch) + 1ull;
  const unsigned int ep32 = (unsigned int)ep;
  const int par = (int)(ep & 1ull);
  const int n = npass * 2 * C;
  for (int e0 = threadIdx.x; e0 < n; e0 += 4 * blockDim.x) {     // loads first: one round trip for up to 4 elements
    double x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e = e0 + k * blockDim.x;
      if (e < n) {
        const int ps = e / (2 * C), idx = e - ps * 2 * C;
        x[k] = __ldcg(stats + (long long)ps * stride + idx);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e = e0 + k * blockDim.x;
      if (e < n) {
        const unsigned int lo = (unsigned int)__double2loint(x[k]), hi = (unsigned int)__double2hiint(x[k]);
        for (int p = 0; p < d.world; ++p) {
          unsigned char* dst = d.peer[p] + nvl_slot_off(d, par, d.rank) + (unsigned long long)e * 16;
          ll_store(dst, lo, ep32);
          ll_store(dst + 8, hi, ep32);
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *d.done = 0u;
    stats[stride - 1] = (double)ep;
    __threadfence();
    *reinterpret_cast<volatile unsigned long long*>(d.epoch) = ep;
  }
}

}  // namespace cvg
