// cvaegan_b200 - small-message all-reduce over NVLink peer memory (one process per GPU, one node).
//
// The training path exchanges 66 BatchNorm moment vectors and 13 flat gradient buffers per label visit (SURVEY.md
// 8e).  They are latency bound: an NCCL all-reduce of a few KB costs ~30 us at 8 GPUs, which is a third of the visit.
// This is a one-shot all-reduce written for that regime:
//   * every rank owns a staging buffer [parity 2][source rank][slot] + flags, shared with its peers through CUDA IPC;
//   * exchange number `ep`: each CTA stores its chunk of the local vector into slot[ep & 1][rank] of EVERY peer
//     (plain stores to peer-mapped addresses travel over NVLink / NVSwitch), fences, then stores the flag
//     (ep & 1, rank, cta) = ep at every peer;
//   * it then waits until its own flags of all sources carry `ep`, and sums the `world` slots IN RANK ORDER, so every
//     rank computes bit-identical sums (the replicated parameters stay bit-identical);
//   * parity double buffering is enough: a rank can only start exchange ep + 2 after it finished ep + 1, which needed
//     every peer's ep + 1 flags, which a peer writes after it has finished reading the slots of ep.
// The exchange number lives in device memory (advanced by the last CTA of each launch), so the launches are
// CUDA-graph capturable.  A peer that never arrives makes the wait trap after ~4 s instead of hanging the GPU.
#pragma once
#include "common.cuh"

namespace cvg {

constexpr int NVL_MAX_WORLD = 8;
constexpr int NVL_MAX_CTAS = 32;
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

template <typename T>
__global__ void __launch_bounds__(NVL_THREADS) nvl_allreduce_kernel(const NvlDev d, T* __restrict__ data, long long n) {
  __shared__ unsigned long long ep_s;
  const int tid = threadIdx.x;
  if (tid == 0) ep_s = *reinterpret_cast<volatile unsigned long long*>(d.epoch) + 1ull;
  __syncthreads();
  const unsigned long long ep = ep_s;
  const int par = (int)(ep & 1ull);
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long i0 = (long long)blockIdx.x * per, i1 = min(n, i0 + per);
  // 1. my chunk -> slot[par][rank] of every peer (self included)
  for (int p = 0; p < d.world; ++p) {
    T* dst = reinterpret_cast<T*>(d.peer[p] + nvl_slot_off(d, par, d.rank));
    for (long long i = i0 + tid; i < i1; i += NVL_THREADS) dst[i] = data[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish, 3. wait for every source
  if (tid < d.world) {
    st_release_sys(reinterpret_cast<unsigned int*>(d.peer[tid] + nvl_flag_off(d, par, d.rank, blockIdx.x)), (unsigned int)ep);
    const unsigned int* f = reinterpret_cast<const unsigned int*>(d.peer[d.rank] + nvl_flag_off(d, par, tid, blockIdx.x));
    const long long t0 = clock64();
    while (ld_acquire_sys(f) != (unsigned int)ep) {
      if (clock64() - t0 > 8000000000ll) __trap();   // ~4 s: a peer never arrived
    }
  }
  __syncthreads();
  __threadfence_system();
  // 4. reduce in rank order (identical on every rank)
  const unsigned char* base = d.peer[d.rank];
  for (long long i = i0 + tid; i < i1; i += NVL_THREADS) {
    T s = 0;
    for (int q = 0; q < d.world; ++q) s += __ldcv(reinterpret_cast<const T*>(base + nvl_slot_off(d, par, q)) + i);
    data[i] = s;
  }
  // 5. the last CTA of the launch completes the exchange
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
