// Development probe (test infrastructure): per-SM rates that bound the tcgen05 attention kernels —
// TMEM load/store bandwidth, MUFU.EX2 rate, tcgen05.mma issue cost and pipe time by shape, commit->mbarrier
// round trip.  One CTA of 288 threads (the kernel's geometry), timed with clock64 inside the kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Ivaesne-dev_b200/csrc tests/probe/tc_rates.cu -o tests/probe/tc_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace vaesne::tc;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int NT = 288;
__global__ void __launch_bounds__(NT, 1) rates(long long* out, float* sink, int reps) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* tiles = (float*)smem;                       // 64 KB of operand tiles (content irrelevant)
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 16384; i += NT) tiles[i] = (float)(i & 7) * 0.125f;
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 8) tmem_alloc<512>(&tmem_s);
  fence_async_smem(); fence_before(); __syncthreads(); fence_after();
  const uint32_t tb = tmem_s;
  const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 128;
  float acc = 0.f;
  long long t0, t1;
  int slot = 0;
  auto sync = [&]() { fence_before(); __syncthreads(); fence_after(); };
  auto rec = [&](long long dt) { if (tid == 0) out[slot] = dt; ++slot; };

  // 0: tcgen05.ld, 8 warps, 128 columns each, per rep
  for (int nw = 8; nw >= 4; nw -= 4) {
    sync(); t0 = clock64();
    if (warp < nw) {
      for (int r = 0; r < reps; ++r) {
        uint32_t v[128];
        tmem_ld32(tl, v); tmem_ld32(tl + 32, v + 32); tmem_ld32(tl + 64, v + 64); tmem_ld32(tl + 96, v + 96); tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 128; c += 8) acc += __uint_as_float(v[c]);
      }
    }
    sync(); t1 = clock64(); rec(t1 - t0);
  }
  // 2: tcgen05.st, 8 warps then 4
  for (int nw = 8; nw >= 4; nw -= 4) {
    sync(); t0 = clock64();
    if (warp < nw) {
      uint32_t v[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] = (uint32_t)(c + tid);
      for (int r = 0; r < reps; ++r) {
        tmem_st32(tl, v); tmem_st32(tl + 32, v); tmem_st32(tl + 64, v); tmem_st32(tl + 96, v); tmem_wait_st();
        v[r & 31] += 1;
      }
    }
    sync(); t1 = clock64(); rec(t1 - t0);
  }
  // 4: MUFU ex2, 8 warps / 4 warps: 128 independent exps per rep per thread
  for (int nw = 8; nw >= 4; nw -= 4) {
    sync(); t0 = clock64();
    if (warp < nw) {
      float x[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) x[c] = -0.001f * (float)(c + lane);
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int c = 0; c < 32; ++c) x[c] = ex2(x[c]) - 1.0f;
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) acc += x[c];
    }
    sync(); t1 = clock64(); rec(t1 - t0);
  }
  // cvt.rn.f16x2.f32 (F2FP) throughput, 8 warps: 64 packs per rep per thread
  {
    sync(); t0 = clock64();
    if (warp < 8) {
      float x[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) x[c] = 0.001f * (float)(c + lane);
      uint32_t accu = 0;
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            uint32_t h; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x[c + 1]), "f"(x[c]));
            accu ^= h; x[c] += 1.0f;
          }
      }
      acc += (float)accu;
    }
    sync(); t1 = clock64(); rec(t1 - t0);
  }
  // 6..: MMA issue + completion. one thread issues `reps` MMAs then commits; records issue-only and total clocks
  const uint64_t dA = smem_desc(smem_u32(tiles), 128, 256), dB = smem_desc(smem_u32(tiles) + 8192, 128, 256);
  const int shapes[5] = {128, 64, 32, 16, 8};
  uint32_t ph = 0;
  for (int ts = 0; ts < 2; ++ts) {
    for (int si = 0; si < 5; ++si) {
      const int N = shapes[si];
      if (N < 16) continue;
      sync();
      if (warp == 8 && elect_one()) {
        const uint32_t id = idesc_tf32(128, N);
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
          if (ts) mma_ts(tb + 256, tb + (uint32_t)(r & 15) * 8, dB, id, 1);
          else mma_ss(tb + 256, dA, dB, id, 1);
        }
        long long ti = clock64();
        commit(&bar[0]);
        mbar_wait(&bar[0], ph);
        t1 = clock64();
        out[slot] = ti - t0; out[slot + 1] = t1 - t0;
      }
      slot += 2; ph ^= 1u;
    }
  }
  // realistic per-tile chain of the backward key pass: 6 first-product MMAs (N=32) + 8 accumulation MMAs (N=16), per-tile
  // commit; (a) alone, (b) while the 8 other warps stream tcgen05.ld/st (TMEM port contention)
  for (int load = 0; load < 4; ++load) {        // 0: none, 1: ld+st, 2: ld only, 3: st only
    sync(); t0 = clock64();
    if (warp == 8) {
      if (elect_one()) {
        const long long m0 = clock64();
        const uint32_t id32 = idesc_tf32(128, 32), id16 = idesc_tf32(128, 16);
        for (int r = 0; r < reps; ++r) {
          for (int q = 0; q < 6; ++q) mma_ts(tb + 384 + (q / 3) * 32, tb + 448 + (q % 3) * 8, dB, id32, (q % 3) ? 1u : 0u);
          for (int q = 0; q < 8; ++q) mma_ts(tb + 480 + (q & 1) * 16, tb + 256 + q * 8, dB, id16, 1u);
          if (r >= 2) mbar_wait(&bar[2 + (r & 1)], (uint32_t)((((r - 2) >> 1) + load * (reps / 2)) & 1));     // two tiles in flight
          commit(&bar[2 + (r & 1)]);
        }
        for (int r = reps - 2; r < reps; ++r) mbar_wait(&bar[2 + (r & 1)], (uint32_t)(((r >> 1) + load * (reps / 2)) & 1));
        out[40 + load] = clock64() - m0;
      }
      __syncwarp();
    } else if (load) {
      const long long w0 = clock64();
      uint32_t v[64];
#pragma unroll
      for (int c = 0; c < 64; ++c) v[c] = (uint32_t)(c + tid);
      for (int r = 0; r < reps * 2; ++r) {
        if (load != 3) { tmem_ld32(tl, v); tmem_ld32(tl + 32, v + 32); tmem_wait_ld(); }
#pragma unroll
        for (int c = 0; c < 64; ++c) v[c] += 0x1000u;
        if (load != 2) { tmem_st32(tl + 64, v); tmem_st32(tl + 96, v + 32); tmem_wait_st(); }
        acc += __uint_as_float(v[r & 63]);
      }
      if (tid == 0) out[44 + load] = clock64() - w0;
    }
    sync(); t1 = clock64(); rec(t1 - t0);
  }
  // dW-style SS MMAs (lin_tc bwd): M=128, N=64, K=8, transposed operands with 144-byte chunk stride (LBO) / 4608-byte groups
  {
    sync();
    if (warp == 8 && elect_one()) {
      const uint32_t id = idesc_tf32(128, 64);
      const long long m0 = clock64();
      for (int r = 0; r < reps; ++r)
        mma_ss(tb + 256, smem_desc(smem_u32(tiles) + (r & 15) * 288, 144, 4608), smem_desc(smem_u32(tiles) + 18432 + (r & 15) * 288, 144, 4608), id, 1);
      commit(&bar[3]);
      mbar_wait(&bar[3], 0);
      out[50] = clock64() - m0;
      const long long m1 = clock64();
      for (int r = 0; r < reps; ++r)
        mma_ss(tb + 256, smem_desc(smem_u32(tiles) + (r & 15) * 256, 128, 4096), smem_desc(smem_u32(tiles) + 16384 + (r & 15) * 256, 128, 4096), id, 1);
      commit(&bar[3]);
      mbar_wait(&bar[3], 1);
      out[51] = clock64() - m1;
    }
    __syncwarp();
  }
  // commit -> wait round trip with a single tiny MMA
  sync();
  if (warp == 8 && elect_one()) {
    t0 = clock64();
    mma_ss(tb + 256, dA, dB, idesc_tf32(128, 16), 0);
    commit(&bar[1]);
    mbar_wait(&bar[1], 0);
    t1 = clock64();
    out[slot] = t1 - t0;
  }
  ++slot;
  sync();
  if (tid == 0) out[63] = slot;
  sink[tid] = acc;
  if (warp == 8) tmem_dealloc<512>(tb);
}

int main() {
  long long* d; float* s;
  cudaMalloc(&d, 64 * 8); cudaMalloc(&s, NT * 4); cudaMemset(d, 0, 64 * 8);
  const int reps = 64;
  cudaFuncSetAttribute(rates, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  for (int pass = 0; pass < 2; ++pass) {
    rates<<<1, NT, 131072>>>(d, s, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 1; }
  }
  long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int i = 0;
  for (int nw = 8; nw >= 4; nw -= 4, ++i) printf("tcgen05.ld %d warps: %lld clk  -> %.1f B/clk/SM\n", nw, h[i], (double)reps * nw * 32 * 128 * 4 / h[i]);
  for (int nw = 8; nw >= 4; nw -= 4, ++i) printf("tcgen05.st %d warps: %lld clk  -> %.1f B/clk/SM\n", nw, h[i], (double)reps * nw * 32 * 128 * 4 / h[i]);
  for (int nw = 8; nw >= 4; nw -= 4, ++i) printf("ex2+fadd   %d warps: %lld clk  -> %.2f ex2/clk/SM\n", nw, h[i], (double)reps * nw * 32 * 128 / h[i]);
  printf("cvt.rn.f16x2.f32 (+xor+fadd) 8 warps: %lld clk -> %.2f packs/clk/SM\n", h[i], (double)reps * 8 * 32 * 64 / h[i]); ++i;
  const int shapes[5] = {128, 64, 32, 16, 8};
  for (int ts = 0; ts < 2; ++ts) for (int si = 0; si < 4; ++si, i += 2)
    printf("mma.%s M128 N%-3d K8 x%d: issue %.1f clk/mma, issue+complete %.1f clk/mma\n", ts ? "ts" : "ss", shapes[si], reps, (double)h[i] / reps, (double)h[i + 1] / reps);
  const char* lname[4] = {"no TMEM traffic", "8 warps tcgen05.ld+st", "8 warps tcgen05.ld", "8 warps tcgen05.st"};
  for (int l = 0; l < 4; ++l, ++i)
    printf("key-pass tile chain (6 x N32 + 8 x N16 MMAs + commit) with %-22s: issuer %.1f clk/tile ; streaming warps %.1f clk per 64-col ld/st round\n",
           lname[l], (double)h[40 + l] / reps, (double)h[44 + l] / (2 * reps));
  printf("dW-style mma.ss M128 N64 K8, LBO 144 / SBO 4608: %.1f clk/mma ; LBO 128 / SBO 4096: %.1f clk/mma\n", (double)h[50] / reps, (double)h[51] / reps);
  printf("single mma + commit + wait round trip: %lld clk\n", h[i]);
  return 0;
}
