// Cross attention onto a handful of context tokens (Lk <= 8): the decoders' 982 / 60 query tokens attend to the 4-5
// latent (+ phase) tokens (PhotometricLayers.py:65-67, SpectraLayers.py:59-62; nn.MultiheadAttention core as in attn.cu).
// With so few keys there is nothing to tile: the step is a pure stream over the query tokens (q, O, dO, dq: 128 B per
// token each), so the kernels are laid out for coalescing — lane = (token, head) reads its contiguous 32 bytes, a warp
// reads 1 KB — with K/V of the batch row in shared memory and all per-key state in registers.
//   forward : one (token, head) per thread.
//   backward: one CTA per batch row walks its tokens 32 at a time; dK/dV (Lk x 8 per head) accumulate in registers
//             across the whole row and are reduced once (shuffles, then shared memory) — no atomics, no second pass.
// Dropout uses the same counter indexing as the general kernels (attn.cu), so either path regenerates the same mask.
#include "common.cuh"
#include "vaesne_b200.h"
#include "attn_args.cuh"
#include <stdlib.h>

namespace vaesne {

constexpr float kSScale = 0.35355339059327373f;     // sqrt(1/8)

__device__ __forceinline__ void ld8s(float* d, const float* p) {
  if (((uintptr_t)p & 15) == 0) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) d[c] = p[c];
  }
}
__device__ __forceinline__ void st8s(float* p, const float* d) {
  if (((uintptr_t)p & 15) == 0) {
    reinterpret_cast<float4*>(p)[0] = make_float4(d[0], d[1], d[2], d[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(d[4], d[5], d[6], d[7]);
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) p[c] = d[c];
  }
}

template <int MK>
__device__ __forceinline__ void stage_kv_small(const AttnArgs& a, int n, float (*sK)[32], float (*sV)[32], float* sB, int tid, int nthreads) {
  for (int i = tid; i < MK * 32; i += nthreads) {
    const int j = i >> 5, c = i & 31;
    const bool in = j < a.Lk;
    sK[j][c] = in ? a.k[((long long)n * a.Lk + j) * a.ldk + c] : 0.f;
    sV[j][c] = in ? a.v[((long long)n * a.Lk + j) * a.ldv + c] : 0.f;
  }
  for (int j = tid; j < MK; j += nthreads) {
    float b = 0.f;
    if (j >= a.Lk) b = -INFINITY;
    else if (a.mask && j < a.mask_len && a.mask[(long long)(n % a.mask_rows) * a.mask_len + j]) b = -INFINITY;
    sB[j] = b;
  }
}

template <int MK>
__global__ void __launch_bounds__(256) attn_small_fwd_kernel(AttnArgs a) {
  __shared__ __align__(16) float sK[MK][32];
  __shared__ __align__(16) float sV[MK][32];
  __shared__ float sB[MK];
  const int tid = threadIdx.x, n = blockIdx.y;
  const int i = blockIdx.x * 64 + (tid >> 2), h = tid & 3;
  stage_kv_small<MK>(a, n, sK, sV, sB, tid, 256);
  __syncthreads();
  if (i >= a.Lq) return;
  float q[8];
  ld8s(q, a.q + ((long long)n * a.Lq + i) * a.ldq + h * 8);
#pragma unroll
  for (int c = 0; c < 8; ++c) q[c] *= kSScale * kLog2e;
  float sc[MK], m = -INFINITY;
#pragma unroll
  for (int j = 0; j < MK; ++j) {
    float s = sB[j];
#pragma unroll
    for (int c = 0; c < 8; ++c) s = fmaf(q[c], sK[j][h * 8 + c], s);
    sc[j] = s; m = fmaxf(m, s);
  }
  const DropCfg dc = make_drop(a.p_drop, a.seed, a.stream_id);
  const uint64_t drow = ((uint64_t)(n * kH + h) * a.Lq + i) * (uint64_t)a.Lk;
  float l = 0.f, o[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) o[c] = 0.f;
#pragma unroll
  for (int j = 0; j < MK; ++j) {
    const float p = (sc[j] == -INFINITY) ? 0.f : exp2f(sc[j] - m);
    l += p;
    float pd = p;
    if (dc.on && j < a.Lk) pd *= drop_mult(dc, drow + (uint64_t)j);
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = fmaf(pd, sV[j][h * 8 + c], o[c]);
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int c = 0; c < 8; ++c) o[c] *= inv;
  st8s(a.O + ((long long)n * a.Lq + i) * a.ldo + h * 8, o);
  a.LSE[((long long)n * kH + h) * a.Lq + i] = (m + log2f(l)) * kLn2;
}

template <int MK>
__global__ void __launch_bounds__(128) attn_small_bwd_kernel(AttnArgs a) {
  __shared__ __align__(16) float sK[MK][32];
  __shared__ __align__(16) float sV[MK][32];
  __shared__ float sB[MK];
  __shared__ float sAcc[4][2][MK][32];       // per warp: dK, dV of the row (head h at columns h*8..)
  const int tid = threadIdx.x, n = blockIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = tid & 3;
  stage_kv_small<MK>(a, n, sK, sV, sB, tid, 128);
  __syncthreads();
  const DropCfg dc = make_drop(a.p_drop, a.seed, a.stream_id);
  float dk[MK][8], dv[MK][8];
#pragma unroll
  for (int j = 0; j < MK; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) { dk[j][c] = 0.f; dv[j][c] = 0.f; }

  for (int i0 = 0; i0 < a.Lq; i0 += 32) {
    const int i = i0 + (tid >> 2);
    if (i >= a.Lq) continue;
    float q[8], g[8], o[8], dq[8];
    ld8s(q, a.q + ((long long)n * a.Lq + i) * a.ldq + h * 8);
    ld8s(g, a.dO + ((long long)n * a.Lq + i) * a.lddo + h * 8);
    ld8s(o, a.O + ((long long)n * a.Lq + i) * a.ldo + h * 8);
    const float lse2 = a.LSE[((long long)n * kH + h) * a.Lq + i] * kLog2e;
    float delta = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) { q[c] *= kSScale * kLog2e; delta = fmaf(g[c], o[c], delta); dq[c] = 0.f; }
    const uint64_t drow = ((uint64_t)(n * kH + h) * a.Lq + i) * (uint64_t)a.Lk;
#pragma unroll
    for (int j = 0; j < MK; ++j) {
      float s = sB[j], dp = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) { s = fmaf(q[c], sK[j][h * 8 + c], s); dp = fmaf(g[c], sV[j][h * 8 + c], dp); }
      const float p = exp2f(s - lse2);             // -inf bias -> 0
      float dm = 1.f;
      if (dc.on && j < a.Lk) dm = drop_mult(dc, drow + (uint64_t)j);
      const float pd = p * dm;
      const float ds = p * (dp * dm - delta);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        dq[c] = fmaf(ds, sK[j][h * 8 + c], dq[c]);
        dk[j][c] = fmaf(ds, q[c], dk[j][c]);
        dv[j][c] = fmaf(pd, g[c], dv[j][c]);
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) dq[c] *= kSScale;
    st8s(a.dq + ((long long)n * a.Lq + i) * a.lddq + h * 8, dq);
  }
  // reduce over the 8 tokens of the warp (lanes with equal head), then over the 4 warps
#pragma unroll
  for (int j = 0; j < MK; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float x = dk[j][c], y = dv[j][c];
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) { x += __shfl_xor_sync(0xffffffffu, x, off); y += __shfl_xor_sync(0xffffffffu, y, off); }
      if (lane < 4) { sAcc[warp][0][j][h * 8 + c] = x * kLn2; sAcc[warp][1][j][h * 8 + c] = y; }     // q carried log2(e)
    }
  __syncthreads();
  for (int idx = tid; idx < a.Lk * 32; idx += 128) {
    const int j = idx >> 5, c = idx & 31;
    a.dk[((long long)n * a.Lk + j) * a.lddk + c] = sAcc[0][0][j][c] + sAcc[1][0][j][c] + sAcc[2][0][j][c] + sAcc[3][0][j][c];
    a.dv[((long long)n * a.Lk + j) * a.lddv + c] = sAcc[0][1][j][c] + sAcc[1][1][j][c] + sAcc[2][1][j][c] + sAcc[3][1][j][c];
  }
}

bool attn_small_eligible(const AttnArgs& a) {
  static const bool off = [] { const char* e = getenv("VAESNE_NO_SMALL_ATTN"); return e && e[0] && e[0] != '0'; }();
  return !off && a.Lk <= 8 && a.Lq >= 32 && a.N <= 65535;
}
int attn_small_fwd(const AttnArgs& a, cudaStream_t st) {
  const dim3 grid((a.Lq + 63) / 64, a.N), block(256);
  if (a.Lk <= 5) { auto kf = attn_small_fwd_kernel<5>; VLAUNCH(kf, grid, block, 0, st, a); }
  else { auto kf = attn_small_fwd_kernel<8>; VLAUNCH(kf, grid, block, 0, st, a); }
  return check_launch("attn_small_fwd");
}
int attn_small_bwd(const AttnArgs& a, cudaStream_t st) {
  if (a.Lk <= 5) { auto kf = attn_small_bwd_kernel<5>; VLAUNCH(kf, dim3(a.N), dim3(128), 0, st, a); }
  else { auto kf = attn_small_bwd_kernel<8>; VLAUNCH(kf, dim3(a.N), dim3(128), 0, st, a); }
  return check_launch("attn_small_bwd");
}

}  // namespace vaesne
