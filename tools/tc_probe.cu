// tools/tc_probe.cu - bring-up probe for cvae_gan_b200/csrc/tc05.cuh on a real B200.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/tc_probe tools/tc_probe.cu && tools/tc_probe
//
// One CTA computes D[M][N] = A[M][K] * B[K][N] with tcgen05.mma kind::tf32 from no-swizzle shared-memory
// descriptors, for every operand-major combination the library uses, and prints the error against an fp64
// reference for the single-pass and the 3xTF32 variants.  It also reports which TMEM lane holds which row
// (M = 64 and M = 128) and the issue cost of back-to-back MMAs for several N.  Development tool only.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../cvae_gan_b200/csrc/tc05.cuh"

using namespace cvg::tc;

struct Case {
  int M, N, K;
  int a_mn, b_mn;      // 0 = K-major, 1 = MN-major
  int a_var, b_var;    // which of the two legal core-matrix orders
  int swap_a, swap_b;  // pass (LBO, SBO) swapped (tests the field interpretation)
  int nsplit;          // 1 = plain tf32, 3 = 3xTF32
};

struct OpLayout {
  uint32_t lbo, sbo, kstep, bytes;
};

__host__ __device__ inline OpLayout op_layout(int rows, int K, int mn_major, int var) {
  OpLayout o;
  if (!mn_major) {
    if (var == 0) { o.sbo = 128; o.lbo = rows * 16; } else if (var == 2) { o.sbo = 128; o.lbo = rows * 16 + 16; } else if (var == 3) { o.sbo = 128; o.lbo = rows * 16 + 128; } else { o.lbo = 128; o.sbo = (K / 4) * 128; }
    o.kstep = 2 * o.lbo;
    o.bytes = (var != 1) ? (K / 4) * o.lbo : (rows / 8) * o.sbo;
  } else {
    if (var == 0) { o.lbo = 128; o.sbo = (K / 8) * 128 + 16; } else { o.sbo = 128; o.lbo = (rows / 4) * 128; }
    o.kstep = o.lbo;
    o.bytes = (var == 0) ? (rows / 4) * o.sbo : (K / 8) * o.lbo;
  }
  return o;
}
__host__ __device__ inline uint32_t op_offset(const OpLayout& o, int mn_major, int row, int k) {
  if (!mn_major) return (row / 8) * o.sbo + (row % 8) * 16 + (k / 4) * o.lbo + (k % 4) * 4;
  return (row / 4) * o.sbo + (row % 4) * 4 + (k / 8) * o.lbo + (k % 8) * 16;
}

// A: [M][K] row-major in global, B: [K][N] row-major in global (= feature-major activations)
__global__ void __launch_bounds__(128) probe_kernel(Case c, const float* A, const float* B, float* D /*[128][N] raw TMEM dump*/,
                                                    long long* cycles, int repeat, int commit_every = 0, int two_acc = 0) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const OpLayout la = op_layout(c.M, c.K, c.a_mn, c.a_var), lb = op_layout(c.N, c.K, c.b_mn, c.b_var);
  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + ((la.bytes + 127) & ~127u);
  uint8_t* b_hi = a_lo + ((la.bytes + 127) & ~127u);
  uint8_t* b_lo = b_hi + ((lb.bytes + 127) & ~127u);

  for (int i = tid; i < c.M * c.K; i += 128) {
    const int row = i / c.K, k = i % c.K;
    float hi, lo;
    split_tf32(A[i], hi, lo);
    if (c.nsplit == 1) hi = A[i];
    *reinterpret_cast<float*>(a_hi + op_offset(la, c.a_mn, row, k)) = hi;
    *reinterpret_cast<float*>(a_lo + op_offset(la, c.a_mn, row, k)) = lo;
  }
  for (int i = tid; i < c.K * c.N; i += 128) {
    const int k = i / c.N, n = i % c.N;
    float hi, lo;
    split_tf32(B[i], hi, lo);
    if (c.nsplit == 1) hi = B[i];
    *reinterpret_cast<float*>(b_hi + op_offset(lb, c.b_mn, n, k)) = hi;
    *reinterpret_cast<float*>(b_lo + op_offset(lb, c.b_mn, n, k)) = lo;
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar2, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(c.M, c.N, c.a_mn, c.b_mn);
    const uint32_t a_l = c.swap_a ? la.sbo : la.lbo, a_s = c.swap_a ? la.lbo : la.sbo;
    const uint32_t b_l = c.swap_b ? lb.sbo : lb.lbo, b_s = c.swap_b ? lb.lbo : lb.sbo;
    const uint64_t dah0 = smem_desc(smem_u32(a_hi), a_l, a_s), dal0 = smem_desc(smem_u32(a_lo), a_l, a_s);
    const uint64_t dbh0 = smem_desc(smem_u32(b_hi), b_l, b_s), dbl0 = smem_desc(smem_u32(b_lo), b_l, b_s);
    const uint32_t ka = la.kstep >> 4, kb = lb.kstep >> 4;
    const int ksteps = c.K / 8;
    t0 = clock64();
    for (int r = 0; r < repeat; ++r) {
      bool acc = false;
#pragma unroll 4
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t dah = dah0 + (uint64_t)(ks * ka), dal = dal0 + (uint64_t)(ks * ka);
        const uint64_t dbh = dbh0 + (uint64_t)(ks * kb), dbl = dbl0 + (uint64_t)(ks * kb);
        const uint32_t dd = tmem + ((two_acc && (ks & 1)) ? 256u - (uint32_t)c.N : 0u);
        if (c.nsplit == 3) {
          mma_tf32(dd, dal, dbh, idesc, acc);
          mma_tf32(dd, dah, dbl, idesc, true);
          acc = true;
        }
        mma_tf32(dd, dah, dbh, idesc, acc);
        acc = true;
        if (commit_every && ((ks + 1) % commit_every) == 0) mma_commit(&bar2);
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (tid == 0) {
    t1 = clock64();
    cycles[0] = t1 - t0;
  }
  tc_fence_after_sync();
  // raw dump: thread t reads TMEM lane t (its warp's 32-lane slice), all N columns
  for (int c0 = 0; c0 < c.N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j)
      if (c0 + j < c.N) D[(size_t)tid * c.N + c0 + j] = v[j];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("device %s sm_%d%d\n", p.name, p.major, p.minor);
  const int MAXM = 128, MAXN = 256, MAXK = 64;
  std::vector<float> hA(MAXM * MAXK), hB(MAXK * MAXN), hD(128 * MAXN);
  float *dA, *dB, *dD;
  long long* dcyc;
  cudaMalloc(&dA, hA.size() * 4);
  cudaMalloc(&dB, hB.size() * 4);
  cudaMalloc(&dD, hD.size() * 4);
  cudaMalloc(&dcyc, 8);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);

  std::vector<Case> cases;
  // forward: A = W (K-major), B = activations (MN-major)
  for (int av = 0; av < 2; ++av)
    for (int bv = 0; bv < 2; ++bv) cases.push_back({128, 64, 32, 0, 1, av, bv, 0, 0, 1});
  for (int av = 0; av < 2; ++av)
    for (int bv = 0; bv < 2; ++bv) cases.push_back({128, 64, 32, 0, 1, av, bv, 0, 0, 3});
  // dW: both K-major
  for (int av = 0; av < 2; ++av)
    for (int bv = 0; bv < 2; ++bv) cases.push_back({128, 64, 32, 0, 0, av, bv, 0, 0, 3});
  // dX: A = W^T (MN-major), B MN-major
  for (int av = 0; av < 2; ++av) cases.push_back({128, 64, 32, 1, 1, av, 0, 0, 0, 3});
  // other shapes
  cases.push_back({128, 128, 64, 0, 1, 0, 0, 0, 0, 3});
  cases.push_back({128, 256, 16, 0, 1, 0, 0, 0, 0, 3});
  cases.push_back({128, 16, 16, 0, 1, 0, 0, 0, 0, 3});
  cases.push_back({64, 64, 32, 0, 0, 0, 0, 0, 0, 3});
  cases.push_back({64, 8, 32, 0, 0, 0, 0, 0, 0, 3});
  cases.push_back({128, 256, 64, 0, 0, 0, 0, 0, 0, 3});
  cases.push_back({128, 16, 16, 0, 0, 0, 0, 0, 0, 3});
  cases.push_back({128, 64, 64, 0, 0, 0, 1, 0, 0, 3});
  cases.push_back({128, 64, 64, 0, 0, 0, 2, 0, 0, 3});
  cases.push_back({128, 64, 64, 0, 0, 0, 3, 0, 0, 3});
  cases.push_back({64, 64, 64, 0, 0, 0, 2, 0, 0, 3});
  // field-interpretation checks (LBO/SBO swapped) go last: a wrong descriptor may fault
  std::vector<Case> swapped;
  swapped.push_back({128, 64, 32, 0, 1, 0, 0, 1, 0, 1});
  swapped.push_back({128, 64, 32, 0, 1, 1, 0, 1, 0, 1});
  swapped.push_back({128, 64, 32, 0, 1, 0, 0, 0, 1, 1});
  swapped.push_back({128, 64, 32, 0, 1, 0, 1, 0, 1, 1});
  swapped.push_back({128, 64, 32, 1, 1, 0, 0, 1, 0, 1});

  int ci = 0;
  auto run_cases = [&](const std::vector<Case>& list) -> int {
  for (const Case& c : list) {
    srand(1234 + ci);
    for (auto& v : hA) v = (float)frand();
    for (auto& v : hB) v = (float)frand();
    cudaMemcpy(dA, hA.data(), (size_t)c.M * c.K * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), (size_t)c.K * c.N * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xFF, hD.size() * 4);
    const OpLayout la = op_layout(c.M, c.K, c.a_mn, c.a_var), lb = op_layout(c.N, c.K, c.b_mn, c.b_var);
    const size_t smem = 2 * ((la.bytes + 127) & ~127u) + 2 * ((lb.bytes + 127) & ~127u) + 256;
    probe_kernel<<<1, 128, smem>>>(c, dA, dB, dD, dcyc, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("case %d: CUDA error %s\n", ci, cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(hD.data(), dD, (size_t)128 * c.N * 4, cudaMemcpyDeviceToHost);
    // reference, A given row-major [M][K] (for a_mn the SAME logical matrix, only the smem order differs)
    double maxerr = 0, maxref = 0;
    int bad_lane_map = 0;
    for (int i = 0; i < c.M; ++i)
      for (int j = 0; j < c.N; ++j) {
        double r = 0;
        for (int k = 0; k < c.K; ++k) r += (double)hA[i * c.K + k] * (double)hB[k * c.N + j];
        const double got = hD[(size_t)i * c.N + j];
        maxerr = fmax(maxerr, fabs(got - r));
        maxref = fmax(maxref, fabs(r));
      }
    printf("case %2d M%3d N%3d K%2d a_mn%d b_mn%d a_var%d b_var%d swap_a%d swap_b%d split%d : max_err %.3e (ref max %.2f) %s\n", ci,
           c.M, c.N, c.K, c.a_mn, c.b_mn, c.a_var, c.b_var, c.swap_a, c.swap_b, c.nsplit, maxerr, maxref,
           maxerr < 1e-2 ? (maxerr < 2e-5 ? "EXACT" : "ok-tf32") : "WRONG");
    if (c.M == 64 && maxerr > 1e-2) {
      // find where rows live: for rows 0, 1, 15, 16, 31, 32, 63 search all lanes
      const int probe_rows[7] = {0, 1, 15, 16, 31, 32, 63};
      for (int pr = 0; pr < 7; ++pr) {
        const int i = probe_rows[pr];
        double r = 0;
        for (int k = 0; k < c.K; ++k) r += (double)hA[i * c.K + k] * (double)hB[k * c.N + 0];
        for (int lane = 0; lane < 128; ++lane)
          for (int col = 0; col < c.N; ++col)
            if (fabs(hD[(size_t)lane * c.N + col] - r) < 1e-4 * fmax(1.0, fabs(r))) printf("   row %d col 0 found at lane %d col %d\n", i, lane, col);
      }
    }
    (void)bad_lane_map;
    ++ci;
  }
  return 0;
  };
  if (run_cases(cases)) return 1;

  // issue-rate: back-to-back MMAs, K = 64 (8 k-steps) x repeat
  for (int N : {16, 32, 64, 128, 256}) {
    for (int bvar : {0, 1, 2, 3}) {
      const int split = 3;
      Case c{128, N, 64, 0, 0, 0, bvar, 0, 0, split};
      const OpLayout la = op_layout(c.M, c.K, 0, 0), lb = op_layout(c.N, c.K, 0, bvar);
      const size_t smem = 2 * ((la.bytes + 127) & ~127u) + 2 * ((lb.bytes + 127) & ~127u) + 256;
      const int repeat = 64;
      probe_kernel<<<1, 128, smem>>>(c, dA, dB, dD, dcyc, repeat);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("timing N=%d: CUDA error %s\n", N, cudaGetErrorString(e));
        return 1;
      }
      long long cyc;
      cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
      const int nmma = repeat * 8 * split;
      printf("timing M128 N%3d b_var%d: %lld cycles for %d MMAs = %.1f cyc/MMA (floor %d)\n", N, bvar, cyc, nmma,
             (double)cyc / nmma, 128 * N / 256);
    }
  }
  for (int ce : {0, 1, 2, 4, 8}) {
    for (int two : {0, 1}) {
      Case c{128, 64, 64, 0, 0, 0, 0, 0, 0, 3};
      const OpLayout la = op_layout(c.M, c.K, 0, 0), lb = op_layout(c.N, c.K, 0, 0);
      const size_t smem = 2 * ((la.bytes + 127) & ~127u) + 2 * ((lb.bytes + 127) & ~127u) + 256;
      probe_kernel<<<1, 128, smem>>>(c, dA, dB, dD, dcyc, 64, ce, two);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("commit test: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      long long cyc;
      cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
      printf("commit every %d k-steps (3 MMAs each), two_acc %d: %.1f cyc/MMA\n", ce, two, (double)cyc / (64 * 8 * 3));
    }
  }
  if (run_cases(swapped)) return 1;
  printf("probe done\n");
  return 0;
}
