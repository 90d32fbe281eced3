// Development probe (test infrastructure): conventions of tcgen05.mma kind::f16 (fp16 operands, fp32 accumulate)
// needed by the attention kernels' second product:
//   * shared-memory K-major no-swizzle operand with 16-bit elements: core matrix = 8 rows x 16 B = 8 rows x 8 halfs,
//     K = 16 per MMA = two core matrices LBO apart, 8-row groups SBO apart;
//   * A operand in TMEM: row m in lane m, K elements 2c and 2c+1 packed in 32-bit column c (low half = even k);
//   * instruction descriptor: c_format F32 (bit 4), a/b format F16 (0), N>>3 at bit 17, M>>4 at bit 24;
//   * a kind::f16 MMA may accumulate onto a D tile written by a kind::tf32 MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -Ivaesne-dev_b200/csrc tests/probe/tc_probe_f16.cu -o tests/probe/tc_probe_f16
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "tc_common.cuh"
using namespace vaesne::tc;

__device__ __forceinline__ uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_ss_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_f16(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
// element (row, k) of a K-major fp16 operand, in halfs: groups of 8 rows (SBO), chunks of 8 k (LBO), 8 halfs per row
__host__ __device__ inline int hoff(int row, int k, int lbo_h, int sbo_h) { return (row >> 3) * sbo_h + (k >> 3) * lbo_h + (row & 7) * 8 + (k & 7); }

__global__ void __launch_bounds__(128) probe(const __half* A, const __half* B, const float* C, float* D1, float* D2, float* D3) {
  __shared__ __align__(128) __half sA[128 * 16];
  __shared__ __align__(128) __half sB[16 * 16];
  __shared__ __align__(128) float sC[16 * 8];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  // A: 16 row groups, each [2 k-chunks][8 rows][8 halfs]: LBO = 64 halfs (128 B), SBO = 128 halfs (256 B)
  for (int i = tid; i < 128 * 16; i += 128) { const int r = i / 16, k = i % 16; sA[hoff(r, k, 64, 128)] = A[i]; }
  for (int i = tid; i < 16 * 16; i += 128) { const int r = i / 16, k = i % 16; sB[hoff(r, k, 64, 128)] = B[i]; }
  for (int i = tid; i < 16 * 8; i += 128) { const int r = i / 8, k = i % 8; sC[kmaj_off(r, k)] = C[i]; }
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  if (warp == 0) tmem_alloc<128>(&tmem_s);
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tb = tmem_s, tl = tb + ((uint32_t)(warp * 32) << 16);
  uint32_t ph = 0;
  // test 1: SS  D1[128x16] = A . B^T
  if (tid == 0) {
    mma_ss_f16(tb, smem_desc(smem_u32(sA), 128, 256), smem_desc(smem_u32(sB), 128, 256), idesc_f16(128, 16), 0);
    commit(&bar);
  }
  mbar_wait(&bar, ph); ph ^= 1; fence_after();
  { uint32_t v[16]; tmem_ld16(tl, v); tmem_wait_ld(); for (int i = 0; i < 16; ++i) D1[tid * 16 + i] = __uint_as_float(v[i]); }
  // test 2: TS  A packed in TMEM columns 32..39 (k = 2c low half, 2c+1 high half)
  {
    uint32_t p[8];
    for (int c = 0; c < 8; ++c) {
      const __half2 h = __halves2half2(A[tid * 16 + 2 * c], A[tid * 16 + 2 * c + 1]);
      p[c] = *reinterpret_cast<const uint32_t*>(&h);
    }
    tmem_st8(tl + 32, p); tmem_wait_st();
  }
  fence_before(); __syncthreads();
  if (tid == 0) {
    fence_after();
    mma_ts_f16(tb + 64, tb + 32, smem_desc(smem_u32(sB), 128, 256), idesc_f16(128, 16), 0);
    commit(&bar);
  }
  mbar_wait(&bar, ph); ph ^= 1; fence_after();
  { uint32_t v[16]; tmem_ld16(tl + 64, v); tmem_wait_ld(); for (int i = 0; i < 16; ++i) D2[tid * 16 + i] = __uint_as_float(v[i]); }
  // test 3: mixed kinds on one accumulator: D3 = tf32(P8 . C^T) + f16(A . B^T), P8 = first 8 columns of A as fp32 in TMEM
  {
    uint32_t p[8];
    for (int c = 0; c < 8; ++c) p[c] = __float_as_uint(__half2float(A[tid * 16 + c]));
    tmem_st8(tl + 48, p); tmem_wait_st();
  }
  fence_before(); __syncthreads();
  if (tid == 0) {
    fence_after();
    mma_ts(tb + 96, tb + 48, smem_desc(smem_u32(sC), 128, 256), idesc_tf32(128, 16), 0);
    mma_ts_f16(tb + 96, tb + 32, smem_desc(smem_u32(sB), 128, 256), idesc_f16(128, 16), 1);
    commit(&bar);
  }
  mbar_wait(&bar, ph); ph ^= 1; fence_after();
  { uint32_t v[16]; tmem_ld16(tl + 96, v); tmem_wait_ld(); for (int i = 0; i < 16; ++i) D3[tid * 16 + i] = __uint_as_float(v[i]); }
  fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tb);
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  std::vector<__half> A(128 * 16), B(16 * 16);
  std::vector<float> C(16 * 8);
  srand(3);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (auto& x : A) x = __float2half(rnd()); for (auto& x : B) x = __float2half(rnd()); for (auto& x : C) x = rnd();
  __half *dA, *dB; float *dC, *d1, *d2, *d3;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dC, C.size() * 4);
  cudaMalloc(&d1, 128 * 16 * 4); cudaMalloc(&d2, 128 * 16 * 4); cudaMalloc(&d3, 128 * 16 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dC, C.data(), C.size() * 4, cudaMemcpyHostToDevice);
  probe<<<1, 128>>>(dA, dB, dC, d1, d2, d3);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> D1(128 * 16), D2(128 * 16), D3(128 * 16);
  cudaMemcpy(D1.data(), d1, D1.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(D2.data(), d2, D2.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D3.data(), d3, D3.size() * 4, cudaMemcpyDeviceToHost);
  double e1 = 0, e2 = 0, e3 = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
    double s = 0, t = 0;
    for (int k = 0; k < 16; ++k) s += (double)__half2float(A[m * 16 + k]) * __half2float(B[n * 16 + k]);
    for (int k = 0; k < 8; ++k) t += (double)__half2float(A[m * 16 + k]) * tf32_trunc(C[n * 8 + k]);
    e1 = fmax(e1, fabs(D1[m * 16 + n] - s)); e2 = fmax(e2, fabs(D2[m * 16 + n] - s)); e3 = fmax(e3, fabs(D3[m * 16 + n] - (s + t)));
  }
  printf("test1 SS f16        max|err| %.3e\n", e1);
  printf("test2 TS f16 packed max|err| %.3e\n", e2);
  printf("test3 tf32 + f16 on one accumulator max|err| %.3e\n", e3);
  printf("D1[0..3] = %f %f %f %f\n", D1[0], D1[1], D1[2], D1[3]);
  return 0;
}
