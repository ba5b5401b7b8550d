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

struct NvlState {
  bool on = false;
  void* local = nullptr;
  size_t total_bytes = 0;
  NvlDev dev{};
  bool opened[NVL_MAX_WORLD] = {};
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

// The vector is `nseg` segments of `seg_len` elements, `seg_stride` apart (BatchNorm moments live in per-pass slots of
// 2 * STAT_C doubles of which only the first 2 C entries are used); a contiguous vector is one segment.
template <typename T>
__global__ void __launch_bounds__(NVL_THREADS) nvl_allreduce_kernel(const NvlDev d, T* __restrict__ data, long long seg_len,
                                                                    long long seg_stride, int nseg) {
  constexpr int W = sizeof(T) / 4;
  __shared__ unsigned long long ep_s;
  const int tid = threadIdx.x;
  if (tid == 0) ep_s = *reinterpret_cast<volatile unsigned long long*>(d.epoch) + 1ull;
  __syncthreads();
  const unsigned long long ep = ep_s;
  const unsigned int ep32 = (unsigned int)ep;
  const int par = (int)(ep & 1ull);
  const long long n = seg_len * nseg;
  const long long stride = (long long)gridDim.x * NVL_THREADS;
  // 1. my elements -> slot[par][rank] of every peer (self included), packed (no segment gaps)
  for (long long e = (long long)blockIdx.x * NVL_THREADS + tid; e < n; e += stride) {
    const long long sg = e / seg_len, off = e - sg * seg_len;
    const T x = data[sg * seg_stride + off];
    unsigned int w[W];
    memcpy(w, &x, sizeof(T));
    for (int p = 0; p < d.world; ++p) {
      unsigned char* dst = d.peer[p] + nvl_slot_off(d, par, d.rank) + (unsigned long long)e * (W * 8);
#pragma unroll
      for (int k = 0; k < W; ++k) ll_store(dst + k * 8, w[k], ep32);
    }
  }
  // 2. poll the packets of every source and sum IN RANK ORDER (identical on every rank)
  const unsigned char* base = d.peer[d.rank];
  const long long t0 = clock64();
  for (long long e = (long long)blockIdx.x * NVL_THREADS + tid; e < n; e += stride) {
    unsigned int w[NVL_MAX_WORLD][W];
    unsigned int pending = (1u << d.world) - 1u;
    while (pending) {
#pragma unroll
      for (int q = 0; q < NVL_MAX_WORLD; ++q) {
        if (q < d.world && ((pending >> q) & 1u)) {
          const unsigned char* src = base + nvl_slot_off(d, par, q) + (unsigned long long)e * (W * 8);
          bool ok = true;
#pragma unroll
          for (int k = 0; k < W; ++k) {
            unsigned int f;
            ll_load(src + k * 8, w[q][k], f);
            ok &= (f == ep32);
          }
          if (ok) pending &= ~(1u << q);
        }
      }
      if (pending && clock64() - t0 > 120000000000ll) __trap();   // ~60 s: a peer never arrived
    }
    T s = 0;
#pragma unroll
    for (int q = 0; q < NVL_MAX_WORLD; ++q) {
      if (q < d.world) {
        T x;
        memcpy(&x, w[q], sizeof(T));
        s += x;
      }
    }
    const long long sg = e / seg_len, off = e - sg * seg_len;
    data[sg * seg_stride + off] = s;
  }
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

}  // namespace cvg
