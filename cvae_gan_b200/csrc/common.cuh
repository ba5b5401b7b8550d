// cvaegan_b200 - shared device/host helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/cvaegan_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "cvaegan_b200 is written for sm_100a (B200) only"
#endif

namespace cvg {

// ------------------------------------------------------------------------------------------------
// error handling (never throw across the C ABI)
// ------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
#define CVG_FAIL(msg)            \
  do {                           \
    ::cvg::set_error(msg);       \
    return 1;                    \
  } while (0)
#define CVG_CUDA(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::cvg::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                       std::to_string(__LINE__) + ")");                                           \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)
#define CVG_TRY(expr)        \
  do {                       \
    int _r = (expr);         \
    if (_r != 0) return _r;  \
  } while (0)

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the 16 lanes that share (lane >> 4)
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double half_warp_sum_d(double v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of doubles (blockDim.x multiple of 32, <= 1024); result valid on every thread
__device__ __forceinline__ double block_sum_d(double v, double* smem32) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum_d(v);
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  double r = 0.0;
  int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) r += smem32[i];
  return r;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al. 2011).  Keyed by the user seed; the counter
// carries (global row, feature group, stream id, step counter) so every draw is a pure function of
// its coordinates: results do not depend on launch geometry or on the number of GPUs.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

struct U4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
    U4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// stream ids (which random tensor of a step a draw belongs to)
enum {
  RS_Z = 1,
  RS_EPS = 2,
  RS_DMASK1 = 3,
  RS_DMASK2 = 4,
  RS_CMASK1 = 5,
  RS_CMASK2 = 6,
  RS_SAMPLE = 7,
  RS_GEN = 8
};

__host__ __device__ __forceinline__ U4 philox_at(uint64_t seed, uint64_t counter, uint32_t stream, uint32_t pass,
                                                 uint64_t grow, uint32_t fgroup) {
  U4 c;
  c.x = (uint32_t)grow;
  c.y = (uint32_t)(grow >> 32) ^ ((uint32_t)(counter >> 32) * 0x9E3779B1u);
  c.z = (stream << 24) | (pass << 20) | (fgroup & 0xFFFFFu);
  c.w = (uint32_t)counter;
  return philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// two uniforms -> two standard normals (Box-Muller; u1 in (0,1])
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;  // (a+1)/2^32
  float u2 = (float)b * 2.3283064365386963e-10f;
  if (u1 > 1.0f) u1 = 1.0f;
  float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// ------------------------------------------------------------------------------------------------
// network layout (host side)
// ------------------------------------------------------------------------------------------------
struct LinearP {
  int out = 0, in = 0;       // weight [out][in], reference layout
  int64_t w = 0, b = 0;      // offsets (floats) into the net's parameter buffer
  int64_t gamma = -1, beta = -1;     // BatchNorm / LayerNorm affine that follows (or -1)
  int64_t rmean = -1, rvar = -1;     // running stats in the state buffer (BN nets)
  int64_t u = -1, v = -1;            // spectral-norm vectors in the state buffer (critic)
};

struct NetLayout {
  int nlin = 0;
  LinearP lin[5];
  int64_t n_param = 0;  // floats, padded
  int64_t n_state = 0;
  std::vector<CvgTensorDesc> table;
};

struct NetBuffers {
  float* params = nullptr;
  float* grads = nullptr;
  float* m = nullptr;
  float* v = nullptr;
  float* state = nullptr;
};

}  // namespace cvg
