// Development probe: decode how tcgen05 (kind::tf32, no swizzle) addresses an MN-major B operand.
// smem tile holds its own float index; A (TMEM) is one-hot in k, so D[m][n] = index of the float used as B[n][k = m % 8].
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../../vaesne-dev_b200/csrc/tc_common.cuh"
using namespace vaesne::tc;

__global__ void __launch_bounds__(128) probe(float* D, uint32_t lbo, uint32_t sbo, int bmn) {
  __shared__ __align__(128) float sB[2048];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < 2048; i += 128) sB[i] = (float)i;
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  if (warp == 0) tmem_alloc<64>(&tslot);
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tb = tslot, lane_base = (uint32_t)(warp * 32) << 16;
  uint32_t p[8];
  for (int k = 0; k < 8; ++k) p[k] = __float_as_uint(k == (tid % 8) ? 1.f : 0.f);
  tmem_st8(tb + lane_base + 0, p); tmem_wait_st();
  fence_before(); __syncthreads();
  if (tid == 0) {
    fence_after();
    uint32_t id = idesc_tf32(128, 16) | (bmn ? (1u << 16) : 0u);
    mma_ts(tb + 16, tb + 0, smem_desc(smem_u32(sB), lbo, sbo), id, 0);
    commit(&bar);
  }
  mbar_wait(&bar, 0); fence_after();
  uint32_t o[16];
  tmem_ld16(tb + lane_base + 16, o); tmem_wait_ld();
  for (int i = 0; i < 16; ++i) D[tid * 16 + i] = __uint_as_float(o[i]);
  fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tb);
}

int main() {
  float* d; cudaMalloc(&d, 128 * 16 * 4);
  std::vector<float> D(128 * 16);
  const uint32_t cfg[][3] = {{128, 256, 0}, {256, 128, 1}, {128, 256, 1}, {512, 128, 1}, {128, 512, 1}, {1024, 64, 1}, {64, 1024, 1}};
  for (auto& c : cfg) {
    probe<<<1, 128>>>(d, c[0], c[1], (int)c[2]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("lbo=%u sbo=%u bmn=%u: %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), d, D.size() * 4, cudaMemcpyDeviceToHost);
    printf("lbo=%u sbo=%u b_mn=%u : float index used for B[n][k]\n", c[0], c[1], c[2]);
    for (int k = 0; k < 8; ++k) { printf("  k=%d:", k); for (int n = 0; n < 16; ++n) printf(" %5.0f", D[k * 16 + n]); printf("\n"); }
  }
  return 0;
}
