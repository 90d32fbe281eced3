// Development probe (test infrastructure): validates the tcgen05 / TMEM conventions the attention
// kernel relies on — no-swizzle K-major shared-memory descriptors for kind::tf32, the instruction
// descriptor, D / A(TMEM) layouts, tcgen05.ld/st 32x32b, commit -> mbarrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 tests/probe/tc_probe.cu -o /tmp/tc_probe && /tmp/tc_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
  return d;                   // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t make_idesc_tf32_bmn(int M, int N) { return make_idesc_tf32(M, N) | (1u << 16); }   // B is MN-major
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
                 "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                  "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// canonical no-swizzle K-major tile: element (row, k) of a [rows x 8] fp32 tile
__host__ __device__ inline int kmaj_off(int row, int k) { return (row / 8) * 64 + (k / 4) * 32 + (row % 8) * 4 + (k % 4); }   // in floats; SBO=256B, LBO=128B

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, const float* P, const float* Vt, float* D1, float* D2, float* D3, float* D4) {
  __shared__ __align__(128) float sA[128 * 8];
  __shared__ __align__(128) float sB[128 * 8];
  __shared__ __align__(128) float sV[16 * 8];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < 128 * 8; i += 128) { int r = i / 8, k = i % 8; sA[kmaj_off(r, k)] = A[i]; sB[kmaj_off(r, k)] = B[i]; }
  for (int i = tid; i < 16 * 8; i += 128) { int r = i / 8, k = i % 8; sV[kmaj_off(r, k)] = Vt[i]; }
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  // ---- test 1: SS, D1[128x128] = A[128x8] . B[128x8]^T -----------------------------------------
  if (tid == 0) {
    mma_ss(tb + 0, make_desc(smem_u32(sA), 128, 256), make_desc(smem_u32(sB), 128, 256), make_idesc_tf32(128, 128), 0);
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < 128; c += 32) {
    float v[32];
    ld32(tb + lane_base + c, v);
    for (int i = 0; i < 32; ++i) D1[tid * 128 + c + i] = v[i];
  }
  // ---- test 2: accumulate a second product on top (acc = 1) with A and B swapped roles of data --------
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mma_ss(tb + 0, make_desc(smem_u32(sB), 128, 256), make_desc(smem_u32(sA), 128, 256), make_idesc_tf32(128, 128), 1);
    commit(&bar);
  }
  mbar_wait(&bar, 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < 128; c += 32) {
    float v[32];
    ld32(tb + lane_base + c, v);
    for (int i = 0; i < 32; ++i) D2[tid * 128 + c + i] = v[i];
  }
  // ---- test 3: TS, D3[128x16] = P[128x8](TMEM, cols 128..135) . Vt[16x8]^T into cols 160..175 ---------
  {
    float p[8];
    for (int k = 0; k < 8; ++k) p[k] = P[tid * 8 + k];
    st8(tb + lane_base + 128, p);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mma_ts(tb + 160, tb + 128, make_desc(smem_u32(sV), 128, 256), make_idesc_tf32(128, 16), 0);
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    float v[32];
    ld32(tb + lane_base + 160, v);
    for (int i = 0; i < 16; ++i) D3[tid * 16 + i] = v[i];
  }
  // ---- test 4: TS with B = rows 8..15 of the K-major tile sB viewed MN-major: D4[m][d] = sum_r P[m][r] * B[8+r][d] ----
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // MN-major no-swizzle: 16-byte unit = 4 consecutive MN (d); 8 consecutive K (rows) 16 B apart; SBO = 128 B between
    // MN groups of 4 (d/4), LBO = 256 B between K groups of 8
    mma_ts(tb + 192, tb + 128, make_desc(smem_u32(sB) + 256, 256, 128), make_idesc_tf32_bmn(128, 16), 0);
    commit(&bar);
  }
  mbar_wait(&bar, 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    float v[32];
    ld32(tb + lane_base + 192, v);
    for (int i = 0; i < 16; ++i) D4[tid * 16 + i] = v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "n"(256));
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  std::vector<float> A(128 * 8), B(128 * 8), P(128 * 8), Vt(16 * 8);
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (auto& x : A) x = rnd(); for (auto& x : B) x = rnd(); for (auto& x : P) x = rnd(); for (auto& x : Vt) x = rnd();
  float *dA, *dB, *dP, *dV, *d1, *d2, *d3, *d4;
  cudaMalloc(&d4, 128 * 16 * 4);
  cudaMalloc(&dA, 4096); cudaMalloc(&dB, 4096); cudaMalloc(&dP, 4096); cudaMalloc(&dV, 512);
  cudaMalloc(&d1, 128 * 128 * 4); cudaMalloc(&d2, 128 * 128 * 4); cudaMalloc(&d3, 128 * 16 * 4);
  cudaMemcpy(dA, A.data(), 4096, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), 4096, cudaMemcpyHostToDevice);
  cudaMemcpy(dP, P.data(), 4096, cudaMemcpyHostToDevice); cudaMemcpy(dV, Vt.data(), 512, cudaMemcpyHostToDevice);
  probe<<<1, 128>>>(dA, dB, dP, dV, d1, d2, d3, d4);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> D1(128 * 128), D2(128 * 128), D3(128 * 16);
  cudaMemcpy(D1.data(), d1, D1.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(D2.data(), d2, D2.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D3.data(), d3, D3.size() * 4, cudaMemcpyDeviceToHost);
  double e1 = 0, e1t = 0, e2 = 0, e3 = 0, e3t = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
    double s = 0, st = 0, s2 = 0;
    for (int k = 0; k < 8; ++k) { s += (double)A[m * 8 + k] * B[n * 8 + k]; st += (double)tf32_trunc(A[m * 8 + k]) * tf32_trunc(B[n * 8 + k]); s2 += (double)B[m * 8 + k] * A[n * 8 + k]; }
    e1 = fmax(e1, fabs(D1[m * 128 + n] - s)); e1t = fmax(e1t, fabs(D1[m * 128 + n] - st)); e2 = fmax(e2, fabs(D2[m * 128 + n] - (s + s2)));
  }
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
    double s = 0, st = 0;
    for (int k = 0; k < 8; ++k) { s += (double)P[m * 8 + k] * Vt[n * 8 + k]; st += (double)tf32_trunc(P[m * 8 + k]) * tf32_trunc(Vt[n * 8 + k]); }
    e3 = fmax(e3, fabs(D3[m * 16 + n] - s)); e3t = fmax(e3t, fabs(D3[m * 16 + n] - st));
  }
  printf("test1 SS   max|err| vs exact %.3e  vs tf32-truncated inputs %.3e\n", e1, e1t);
  printf("test2 acc  max|err| vs exact %.3e\n", e2);
  printf("test3 TS   max|err| vs exact %.3e  vs tf32-truncated inputs %.3e\n", e3, e3t);
  {
    std::vector<float> D4(128 * 16);
    cudaMemcpy(D4.data(), d4, D4.size() * 4, cudaMemcpyDeviceToHost);
    double e4 = 0;
    for (int m = 0; m < 128; ++m) for (int d = 0; d < 8; ++d) {
      double st = 0;
      for (int r = 0; r < 8; ++r) st += (double)tf32_trunc(P[m * 8 + r]) * tf32_trunc(B[(8 + r) * 8 + d]);
      e4 = fmax(e4, fabs(D4[m * 16 + d] - st));
    }
    printf("test4 TS, B MN-major view of a K-major tile: max|err| vs tf32-truncated inputs %.3e (cols 8..15 ignored)\n", e4);
  }
  printf("D1[0..3]= %f %f %f %f ; D3[0..3]= %f %f %f %f\n", D1[0], D1[1], D1[2], D1[3], D3[0], D3[1], D3[2], D3[3]);
  return 0;
}
