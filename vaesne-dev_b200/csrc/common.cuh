// Shared device/host helpers for the VAESNe B200 kernels.
//
// The portable kernels in this directory are written against a tiny launch/shared-memory
// macro layer so that the test-suite can also compile the very same sources with g++
// against tests/emu (a CUDA execution-model emulator used only by CPU tests).  The
// product build is always nvcc -gencode arch=compute_100a,code=sm_100a.
#pragma once
#ifdef VAESNE_EMU
#include "emu_cuda.h"
#define VLAUNCH(kfn, grid, block, smem, stream, ...) \
  emu::launch(grid, block, smem, [=]() { kfn(__VA_ARGS__); })
#define VDYNSMEM(type, name) type* name = reinterpret_cast<type*>(emu::dyn_smem())
#define VSET_SMEM(kfn, bytes) (void)0
#else
#include <cuda_runtime.h>
#define VLAUNCH(kfn, grid, block, smem, stream, ...) kfn<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define VDYNSMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw[]; \
  type* name = reinterpret_cast<type*>(name##_raw)
#define VSET_SMEM(kfn, bytes) \
  cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#endif
#include <stdint.h>
#include <math.h>

namespace vaesne {

// ---- error plumbing (C-ABI: int status + thread-local message) -------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);
enum : int { V_OK = 0, V_EBADSHAPE = -1, V_EUNSUPPORTED = -2, V_EALIGN = -3, V_ECUDA = -4, V_ENULL = -5 };

#define V_REQUIRE(cond, code, ...) do { if (!(cond)) { vaesne::set_error(__VA_ARGS__); return (code); } } while (0)

constexpr int kD = 32;      // model_dim supported by the fused kernels (all reference scripts use 32)
constexpr int kH = 4;       // heads
constexpr int kDh = 8;      // head_dim
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ---- counter-based dropout RNG ---------------------------------------------------------
// keep(i) for element i of dropout stream `stream` under seed (s0,s1): one 32-bit hash
// serves two consecutive elements (16-bit thresholds, p quantised to 1/65536).
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
  return h;
}
__device__ __forceinline__ uint32_t hash_ctr(uint32_t s0, uint32_t s1, uint32_t stream, uint64_t ctr) {
  uint32_t lo = (uint32_t)ctr, hi = (uint32_t)(ctr >> 32);
  uint32_t h = mix32(lo * 0x9E3779B1U + s0);
  h = mix32(h ^ (hi * 0x85EBCA77U + s1));
  h = mix32(h + stream * 0xC2B2AE3DU);
  return h;
}
struct DropCfg {
  uint32_t s0, s1, stream;
  uint32_t thresh;   // drop if rnd16 < thresh ; thresh = round(p*65536)
  float scale;       // 1/(1-thresh/65536)
  bool on;
};
__device__ __forceinline__ DropCfg make_drop(float p, const uint64_t* seed, uint32_t stream) {
  DropCfg d;
  d.on = (p > 0.f) && (seed != nullptr);
  d.s0 = d.s1 = 0; d.stream = stream; d.thresh = 0; d.scale = 1.f;
  if (d.on) {
    uint64_t s = *seed;
    d.s0 = (uint32_t)s; d.s1 = (uint32_t)(s >> 32);
    float t = p * 65536.f + 0.5f;
    d.thresh = t > 65535.f ? 65535u : (uint32_t)t;
    d.scale = 1.f / (1.f - (float)d.thresh * (1.f / 65536.f));
  }
  return d;
}
// multiplier (0 or scale) for element index `idx`
__device__ __forceinline__ float drop_mult(const DropCfg& d, uint64_t idx) {
  uint32_t r = hash_ctr(d.s0, d.s1, d.stream, idx >> 1);
  uint32_t r16 = (idx & 1) ? (r >> 16) : (r & 0xffffu);
  return r16 < d.thresh ? 0.f : d.scale;
}

// Row-wise variant for the 32 features of one token (LayerNorm-input dropout): one counter hash per token, then one
// mix32 per feature pair — 4x fewer integer instructions than drop_mult per element.
__device__ __forceinline__ uint32_t drop_row_hash(const DropCfg& d, uint64_t row) { return hash_ctr(d.s0, d.s1, d.stream, row); }
__device__ __forceinline__ float drop_mult_row(const DropCfg& d, uint32_t rh, int j) {
  const uint32_t r = mix32(rh + (uint32_t)(j >> 1) * 0x9E3779B1U);
  const uint32_t r16 = (j & 1) ? (r >> 16) : (r & 0xffffu);
  return r16 < d.thresh ? 0.f : d.scale;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
// Same derivative with erf from Abramowitz & Stegun 7.1.26 (|error| < 1.5e-7), sharing its exponential with the
// density term: two MUFU operations and ~12 FMA-pipe instructions per element instead of erff + expf (~45).  Used where
// the derivative is the bulk of a kernel's arithmetic (the pipelined linear backward).
__device__ __forceinline__ float gelu_erf_grad_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  const float e = __expf(-0.5f * x * x);
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erf_abs = fmaf(-p * t, e, 1.f);
  const float cdf = fmaf(0.5f, copysignf(erf_abs, x), 0.5f);
  return fmaf(x * 0.3989422804014327f, e, cdf);
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

}  // namespace vaesne
