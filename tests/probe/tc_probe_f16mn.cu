// Development probe (test infrastructure): tcgen05.mma kind::f16 with MN-major (transposed) shared-memory operands
// and M=64 accumulators — what a fused attention backward needs to contract over the TMEM-lane index:
//   * MN-major no-swizzle operand, 16-bit elements: core matrix = 8 k-rows x 16 B (8 MN-contiguous halfs);
//     element (mn, k) at  (mn&7)*2 + (mn>>3)*SBO + (k&7)*16 + (k>>3)*LBO  bytes; instruction-descriptor bits 15 (A) / 16 (B);
//   * M=64, cta_group::1: which TMEM lanes hold D row m, and whether the other lanes stay untouched
//     (expected from the CUTLASS fragment layout: row m -> lane (m&15) + 32*(m>>4), lanes 16..31 of each quarter free);
//   * issue cost of a chain of such SS MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -Ivaesne-dev_b200/csrc tests/probe/tc_probe_f16mn.cu -o tests/probe/tc_probe_f16mn
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "tc_common.cuh"
using namespace vaesne::tc;

__host__ __device__ inline int mn_off(int mn, int k, int lbo, int sbo) { return (mn & 7) * 2 + (mn >> 3) * sbo + (k & 7) * 16 + (k >> 3) * lbo; }

constexpr int KT = 128;       // contraction length (keys)
constexpr int A_SBO = 2048, A_LBO = 128, B_SBO = 2048, B_LBO = 128;

// mode: M (64|128), N (8|16), lane_off (0|16) of the D address
__global__ void __launch_bounds__(128) probe(const __half* A, const __half* B, float* D, int M, int N, int lane_off, long long* clk, int reps, int indep) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* sA = raw;                      // 16 m-groups x 2048 B = 32 KB
  unsigned char* sB = raw + 32768;              // up to 16 n-groups x 2048 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(raw + 32768 + 32768);
  uint32_t* tmem_s = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * KT; i += 128) { const int m = i / KT, k = i % KT; *reinterpret_cast<__half*>(sA + mn_off(m, k, A_LBO, A_SBO)) = A[i]; }
  for (int i = tid; i < 16 * KT; i += 128) { const int n = i / KT, k = i % KT; *reinterpret_cast<__half*>(sB + mn_off(n, k, B_LBO, B_SBO)) = B[i]; }
  if (tid == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  if (warp == 0) tmem_alloc<512>(tmem_s);
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tb = *tmem_s, tl = tb + ((uint32_t)(warp * 32) << 16);
  { uint32_t v[16]; for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(-777.f); tmem_st16(tl, v); tmem_st16(tl + 16, v); tmem_wait_st(); }
  fence_before(); __syncthreads();
  uint32_t ph = 0;
  const uint32_t id = idesc_f16_mn(M, N, true, true);
  const uint32_t dst = tb + ((uint32_t)lane_off << 16);
  if (tid == 0) {
    fence_after();
    for (int t = 0; t < KT / 16; ++t)
      mma_ss_f16(dst, smem_desc(smem_u32(sA) + t * 256, A_LBO, A_SBO), smem_desc(smem_u32(sB) + t * 256, B_LBO, B_SBO), id, t > 0);
    commit(bar);
  }
  mbar_wait(bar, ph); ph ^= 1; fence_after();
  { uint32_t v[16]; tmem_ld16(tl, v); tmem_wait_ld(); for (int i = 0; i < 16; ++i) D[tid * 16 + i] = __uint_as_float(v[i]); }
  fence_before(); __syncthreads();
  // timing: reps chains of 8 MMAs into columns 16.. (accumulating), one commit at the end
  if (tid == 0) {
    fence_after();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int t = 0; t < KT / 16; ++t)
        mma_ss_f16(dst + 256 + ((indep && (t & 1)) ? (16u << 16) : 0u), smem_desc(smem_u32(sA) + t * 256, A_LBO, A_SBO), smem_desc(smem_u32(sB) + t * 256, B_LBO, B_SBO), id, 1);
    commit(bar);
    const long long t1 = clock64();
    mbar_wait(bar, ph);
    const long long t2 = clock64();
    clk[0] = t1 - t0; clk[1] = t2 - t0;
  }
  __syncthreads();
  fence_before(); __syncthreads();
  if (warp == 0) { fence_after(); tmem_dealloc<512>(tb); }
}

int main() {
  std::vector<__half> A(128 * KT), B(16 * KT);
  srand(5);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (auto& x : A) x = __float2half(rnd());
  for (auto& x : B) x = __float2half(rnd());
  __half *dA, *dB; float* dD; long long* dC;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * 16 * 4); cudaMalloc(&dC, 16);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = 32768 + 32768 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<double> want(128 * 16);
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
    double s = 0; for (int k = 0; k < KT; ++k) s += (double)__half2float(A[m * KT + k]) * __half2float(B[n * KT + k]);
    want[m * 16 + n] = s;
  }
  const int cfg[][4] = {{128, 16, 0, 0}, {128, 8, 0, 0}, {64, 16, 0, 0}, {64, 8, 0, 0}, {64, 16, 16, 0}, {64, 8, 16, 0}, {64, 8, 0, 1}, {128, 64, 0, 0}, {128, 96, 0, 0}, {128, 128, 0, 0}, {128, 32, 0, 0}};
  for (auto& c : cfg) {
    const int M = c[0], N = c[1], lo = c[2], reps = 64;
    probe<<<1, 128, smem>>>(dA, dB, dD, M, N, lo, dC, reps, c[3]);
    cudaError_t e = cudaDeviceSynchronize();
    printf("M=%d N=%d lane_off=%d two-accumulators=%d: %s\n", M, N, lo, c[3], cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> D(128 * 16); long long clk[2];
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(clk, dC, 16, cudaMemcpyDeviceToHost);
    // which row does each lane hold?  (match on column 0..N-1)
    int touched = 0, matched = 0; double worst = 0;
    std::vector<int> lane_row(128, -2);
    for (int l = 0; l < 128; ++l) {
      if (D[l * 16] == -777.f) { lane_row[l] = -1; continue; }
      ++touched;
      for (int m = 0; m < 128; ++m) {
        double err = 0; for (int n = 0; n < N; ++n) err = fmax(err, fabs(D[l * 16 + n] - want[m * 16 + n]));
        if (err < 2e-2) { lane_row[l] = m; ++matched; worst = fmax(worst, err); break; }
      }
    }
    printf("  lanes touched %d, matched to a row %d, worst |err| %.3e ; untouched cols N..15 on lane0: %s\n", touched, matched, worst,
           (N == 16 || D[N] == -777.f) ? "yes" : "NO");
    printf("  lane->row:");
    for (int l = 0; l < 128; ++l) { if (l % 32 == 0) printf("\n   "); printf(" %d", lane_row[l]); }
    printf("\n  issue %.1f clk/MMA, complete %.1f clk/MMA (chain of %d)\n", (double)clk[0] / (reps * 8), (double)clk[1] / (reps * 8), reps * 8);
  }
  return 0;
}
