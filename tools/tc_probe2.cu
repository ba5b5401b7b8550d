// tools/tc_probe2.cu - impulse-response probe: which (row, k) does tcgen05.mma read from a given byte offset of an
// MN-major no-swizzle operand?  One-hot operand in shared memory, known K-major partner, one MMA per offset.
// Development tool only.
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../cvae_gan_b200/csrc/tc05.cuh"
using namespace cvg::tc;

// which: 0 = probe B (MN-major), A K-major known ; 1 = probe A (MN-major), B K-major known
__global__ void __launch_bounds__(128) impulse_kernel(int which, uint32_t lbo, uint32_t sbo, int layout_type, int span_bytes,
                                                      int N, int* out /*[span/4][2]*/) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int found[2];
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* known = smem;                // K-major operand, 128 rows x 8 k:  [k/4][row][k%4]  (SBO 128, LBO rows*16)
  uint8_t* probe = smem + 8192;         // region under test
  for (int i = tid; i < 8192 / 4; i += 128) reinterpret_cast<float*>(known)[i] = 0.f;
  for (int i = tid; i < span_bytes / 4 + 4096; i += 128) reinterpret_cast<float*>(probe)[i] = 0.f;
  __syncthreads();
  // known operand: row 0 holds (k + 1), every other row r holds 0 except a marker so that rows are distinguishable
  if (tid < 8) *reinterpret_cast<float*>(known + (tid / 4) * (128 * 16) + 0 * 16 + (tid % 4) * 4) = (float)(tid + 1);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  uint32_t parity = 0;
  const uint64_t dk = smem_desc(smem_u32(known), 128 * 16, 128);
  uint64_t dp = smem_desc(smem_u32(probe), lbo, sbo) | ((uint64_t)layout_type << 61);
  for (int x = 0; x < span_bytes; x += 4) {
    if (tid == 0) {
      if (x) *reinterpret_cast<float*>(probe + x - 4) = 0.f;
      *reinterpret_cast<float*>(probe + x) = 1.0f;
      found[0] = -1;
      found[1] = -1;
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      fence_proxy_async_smem();
      if (which == 0) mma_tf32(tmem, dk, dp, idesc_tf32(128, N, 0, 1), false);   // D[i][j] = sum_k known[i][k] * probe[k][j]
      else mma_tf32(tmem, dp, dk, idesc_tf32(128, N, 1, 0), false);              // D[i][j] = sum_k probe[i][k] * known[j][k]
      mma_commit(&bar);
    }
    mbar_wait(&bar, parity);
    parity ^= 1;
    tc_fence_after_sync();
    if (which == 0) {
      if (warp == 0) {
        for (int c0 = 0; c0 < N; c0 += 32) {
          float v[32];
          tmem_ld32(tmem + c0, v);
          tmem_wait_ld();
          if (tid == 0)
            for (int j = 0; j < 32; ++j)
              if (v[j] != 0.f) { found[0] = (int)(v[j] + 0.5f) - 1; found[1] = c0 + j; }
        }
      }
    } else {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
      tmem_wait_ld();
      if (v[0] != 0.f) { found[0] = (int)(v[0] + 0.5f) - 1; found[1] = tid; }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      out[(x / 4) * 2 + 0] = found[0];
      out[(x / 4) * 2 + 1] = found[1];
    }
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
  const int span = 4096;
  int* dout;
  cudaMalloc(&dout, span / 4 * 2 * sizeof(int));
  std::vector<int> h(span / 4 * 2);
  cudaFuncSetAttribute(impulse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  struct Cfg { int which; uint32_t lbo, sbo; int lt; int N; };
  const Cfg cfgs[] = {{0, 128, 1024, 0, 64}, {0, 1024, 128, 0, 64}, {0, 256, 2048, 0, 64}, {0, 2048, 256, 0, 64},
                      {1, 128, 1024, 0, 64}, {1, 1024, 128, 0, 64},
                      {0, 1024, 128, 6, 64},  /* 32B swizzle */
                      {0, 2048, 1024, 2, 64}, /* 128B swizzle */
                      {0, 2048, 1024, 1, 64}  /* 128B base 32B */};
  for (const Cfg& c : cfgs) {
    cudaMemset(dout, 0xFF, h.size() * 4);
    impulse_kernel<<<1, 128, 8192 + span + 16384 + 1024>>>(c.which, c.lbo, c.sbo, c.lt, span, c.N, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost);
    printf("== which=%d (%s MN-major) lbo=%u sbo=%u layout_type=%d : byte offset -> (k, %s)\n", c.which, c.which ? "A" : "B", c.lbo,
           c.sbo, c.lt, c.which ? "row" : "col");
    for (int x = 0; x < span / 4; ++x) {
      if (h[2 * x] >= 0) printf(" %4d:(k%d,%d)", x * 4, h[2 * x], h[2 * x + 1]);
      if (x % 8 == 7) printf("\n");
    }
    printf("\n");
  }
  return 0;
}
