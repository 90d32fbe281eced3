// Attention over short and medium sequences (Lq, Lk <= 255): the photometry stacks' 60-point light curves — decoder
// self-attention on K*M*B rows of 60 tokens, the encoder's 60 x 60 and 8 x 60 blocks (PhotometricLayers.py:48-67,117-143; core
// as in attn.cu) — and anything up to where the tcgen05 kernels take over (256 tokens): real light curves with more than 64
// points, short spectra.  Everything one batch row needs (K, V, and in the backward Q, dO, lse, delta) fits in shared memory
// (34 KB at 64 tokens, 137 KB at 256), so ONE CTA serves a batch row with all four heads: global memory is touched once,
// coalesced, and the score matrix never exists.  Three instantiations (64 / 128 / 256 tokens, 4 threads per token).
//   forward : thread = (query, head); two sweeps over the keys in shared memory (row max, then exp / sum / PV).
//   backward: the same CTA runs a query-major sweep (dQ) and a key-major sweep (dK, dV) over the staged operands — each
//             gradient row is owned by one thread, so there are no atomics and no second launch.
// Dropout uses the counter indexing of the general kernels (attn.cu), so the paths regenerate identical masks.
#include "common.cuh"
#include "vaesne_b200.h"
#include "attn_args.cuh"
#include <stdlib.h>

namespace vaesne {

constexpr int kMidMax = 256;                        // largest instantiation (exclusive upper bound of the window is 256)
constexpr float kMScale = 0.35355339059327373f;     // sqrt(1/8)

__device__ __forceinline__ void ld8m(float* d, const float* p) {
  if (((uintptr_t)p & 15) == 0) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) d[c] = p[c];
  }
}
__device__ __forceinline__ void st8m(float* p, const float* d) {
  if (((uintptr_t)p & 15) == 0) {
    reinterpret_cast<float4*>(p)[0] = make_float4(d[0], d[1], d[2], d[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(d[4], d[5], d[6], d[7]);
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) p[c] = d[c];
  }
}
__device__ __forceinline__ void lds8(float* d, const float* p) {      // 32-byte aligned shared row segment
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
__device__ __forceinline__ float dot8(const float* a, const float* b, float acc) {
#pragma unroll
  for (int c = 0; c < 8; ++c) acc = fmaf(a[c], b[c], acc);
  return acc;
}

// rows of a [L x 32] operand -> shared memory (one (token, head) slice per thread), optionally scaled
__device__ __forceinline__ void stage_rows(float (*dst)[32], const float* src, long long ld, int L, int tid, float scale) {
  const int i = tid >> 2, h = tid & 3;
  float x[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) x[c] = 0.f;
  if (i < L) ld8m(x, src + (long long)i * ld + h * 8);
#pragma unroll
  for (int c = 0; c < 8; ++c) x[c] *= scale;
  reinterpret_cast<float4*>(&dst[i][h * 8])[0] = make_float4(x[0], x[1], x[2], x[3]);
  reinterpret_cast<float4*>(&dst[i][h * 8])[1] = make_float4(x[4], x[5], x[6], x[7]);
}
template <int ML>
__device__ __forceinline__ void stage_bias(const AttnArgs& a, int n, float* sB, int tid) {
  if (tid < ML) {
    float b = 0.f;
    if (tid >= a.Lk) b = -INFINITY;
    else if (a.mask && tid < a.mask_len && a.mask[(long long)(n % a.mask_rows) * a.mask_len + tid]) b = -INFINITY;
    sB[tid] = b;
  }
}

template <int ML>
__global__ void __launch_bounds__(ML * 4) attn_mid_fwd_kernel(AttnArgs a) {
  VDYNSMEM(float, smem);
  float (*sK)[32] = reinterpret_cast<float (*)[32]>(smem);
  float (*sV)[32] = reinterpret_cast<float (*)[32]>(smem + ML * 32);
  float* sB = smem + 2 * ML * 32;
  const int tid = threadIdx.x, n = blockIdx.x;
  const int i = tid >> 2, h = tid & 3;
  stage_rows(sK, a.k + (long long)n * a.Lk * a.ldk, a.ldk, a.Lk, tid, 1.f);
  stage_rows(sV, a.v + (long long)n * a.Lk * a.ldv, a.ldv, a.Lk, tid, 1.f);
  stage_bias<ML>(a, n, sB, tid);
  float q[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) q[c] = 0.f;
  if (i < a.Lq) ld8m(q, a.q + ((long long)n * a.Lq + i) * a.ldq + h * 8);
#pragma unroll
  for (int c = 0; c < 8; ++c) q[c] *= kMScale * kLog2e;
  __syncthreads();
  if (i >= a.Lq) return;
  float m = -INFINITY;
#pragma unroll 4
  for (int j = 0; j < a.Lk; ++j) {
    float kk[8];
    lds8(kk, &sK[j][h * 8]);
    m = fmaxf(m, dot8(q, kk, sB[j]));
  }
  const DropCfg dc = make_drop(a.p_drop, a.seed, a.stream_id);
  const uint64_t drow = ((uint64_t)(n * kH + h) * a.Lq + i) * (uint64_t)a.Lk;
  float l = 0.f, o[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) o[c] = 0.f;
#pragma unroll 4
  for (int j = 0; j < a.Lk; ++j) {
    float kk[8], vv[8];
    lds8(kk, &sK[j][h * 8]);
    lds8(vv, &sV[j][h * 8]);
    const float s = dot8(q, kk, sB[j]);
    const float p = (s == -INFINITY) ? 0.f : exp2f(s - m);
    l += p;
    const float pd = dc.on ? p * drop_mult(dc, drow + (uint64_t)j) : p;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = fmaf(pd, vv[c], o[c]);
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int c = 0; c < 8; ++c) o[c] *= inv;
  st8m(a.O + ((long long)n * a.Lq + i) * a.ldo + h * 8, o);
  a.LSE[((long long)n * kH + h) * a.Lq + i] = (m + log2f(l)) * kLn2;
}

template <int ML>
__global__ void __launch_bounds__(ML * 4) attn_mid_bwd_kernel(AttnArgs a) {
  VDYNSMEM(float, smem);
  float (*sQ)[32] = reinterpret_cast<float (*)[32]>(smem);                      // scaled by sqrt(1/8) * log2(e)
  float (*sK)[32] = reinterpret_cast<float (*)[32]>(smem + ML * 32);
  float (*sV)[32] = reinterpret_cast<float (*)[32]>(smem + 2 * ML * 32);
  float (*sG)[32] = reinterpret_cast<float (*)[32]>(smem + 3 * ML * 32);       // dO
  float (*sL)[ML] = reinterpret_cast<float (*)[ML]>(smem + 4 * ML * 32);       // lse (log2 units)
  float (*sD)[ML] = reinterpret_cast<float (*)[ML]>(smem + 4 * ML * 32 + kH * ML);   // delta = rowsum(dO * O)
  float* sB = smem + 4 * ML * 32 + 2 * kH * ML;
  const int tid = threadIdx.x, n = blockIdx.x;
  const int i = tid >> 2, h = tid & 3;
  const long long nh = (long long)n * kH + h;
  stage_rows(sK, a.k + (long long)n * a.Lk * a.ldk, a.ldk, a.Lk, tid, 1.f);
  stage_rows(sV, a.v + (long long)n * a.Lk * a.ldv, a.ldv, a.Lk, tid, 1.f);
  stage_rows(sQ, a.q + (long long)n * a.Lq * a.ldq, a.ldq, a.Lq, tid, kMScale * kLog2e);
  stage_bias<ML>(a, n, sB, tid);
  float g[8], lse2 = INFINITY, delta = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) g[c] = 0.f;
  if (i < a.Lq) {
    float o[8];
    ld8m(g, a.dO + ((long long)n * a.Lq + i) * a.lddo + h * 8);
    ld8m(o, a.O + ((long long)n * a.Lq + i) * a.ldo + h * 8);
    delta = dot8(g, o, 0.f);
    lse2 = a.LSE[nh * a.Lq + i] * kLog2e;
  }
  reinterpret_cast<float4*>(&sG[i][h * 8])[0] = make_float4(g[0], g[1], g[2], g[3]);
  reinterpret_cast<float4*>(&sG[i][h * 8])[1] = make_float4(g[4], g[5], g[6], g[7]);
  sL[h][i] = lse2; sD[h][i] = delta;
  __syncthreads();
  const DropCfg dc = make_drop(a.p_drop, a.seed, a.stream_id);
  // ---- query-major sweep: dQ_i = scale * sum_j dS_ij K_j --------------------------------------------------------------
  if (i < a.Lq) {
    float q[8], dq[8];
    lds8(q, &sQ[i][h * 8]);
#pragma unroll
    for (int c = 0; c < 8; ++c) dq[c] = 0.f;
    const uint64_t drow = ((uint64_t)nh * a.Lq + i) * (uint64_t)a.Lk;
#pragma unroll 4
    for (int j = 0; j < a.Lk; ++j) {
      float kk[8], vv[8];
      lds8(kk, &sK[j][h * 8]);
      lds8(vv, &sV[j][h * 8]);
      const float p = exp2f(dot8(q, kk, sB[j]) - lse2);       // -inf bias -> 0
      const float dm = dc.on ? drop_mult(dc, drow + (uint64_t)j) : 1.f;
      const float ds = p * (dot8(g, vv, 0.f) * dm - delta);
#pragma unroll
      for (int c = 0; c < 8; ++c) dq[c] = fmaf(ds, kk[c], dq[c]);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) dq[c] *= kMScale;
    st8m(a.dq + ((long long)n * a.Lq + i) * a.lddq + h * 8, dq);
  }
  // ---- key-major sweep: dV_j = sum_i Pd_ij dO_i ; dK_j = scale * sum_i dS_ij Q_i ---------------------------------------
  const int j = i;                       // this thread's key
  if (j < a.Lk) {
    float kk[8], vv[8], dk[8], dv[8];
    lds8(kk, &sK[j][h * 8]);
    lds8(vv, &sV[j][h * 8]);
    const float bias = sB[j];
#pragma unroll
    for (int c = 0; c < 8; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
#pragma unroll 2
    for (int r = 0; r < a.Lq; ++r) {
      float q[8], gg[8];
      lds8(q, &sQ[r][h * 8]);
      lds8(gg, &sG[r][h * 8]);
      const float p = exp2f(dot8(q, kk, bias) - sL[h][r]);
      const float dm = dc.on ? drop_mult(dc, ((uint64_t)nh * a.Lq + r) * (uint64_t)a.Lk + (uint64_t)j) : 1.f;
      const float pd = p * dm;
      const float ds = p * (dot8(gg, vv, 0.f) * dm - sD[h][r]);
#pragma unroll
      for (int c = 0; c < 8; ++c) { dk[c] = fmaf(ds, q[c], dk[c]); dv[c] = fmaf(pd, gg[c], dv[c]); }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) dk[c] *= kLn2;        // Q carried log2(e)
    st8m(a.dk + ((long long)n * a.Lk + j) * a.lddk + h * 8, dk);
    st8m(a.dv + ((long long)n * a.Lk + j) * a.lddv + h * 8, dv);
  }
}

bool attn_mid_eligible(const AttnArgs& a) {
  static const bool off = [] { const char* e = getenv("VAESNE_NO_MID_ATTN"); return e && e[0] && e[0] != '0'; }();
  return !off && a.Lk > 8 && a.Lk < kMidMax && a.Lq < kMidMax && a.Lq >= 1;
}
template <int ML>
static int mid_fwd(const AttnArgs& a, cudaStream_t st) {
  const size_t smem = sizeof(float) * (2 * ML * 32 + ML);
  auto k = attn_mid_fwd_kernel<ML>;
  VSET_SMEM(k, smem);
  VLAUNCH(k, dim3(a.N), dim3(ML * 4), smem, st, a);
  return check_launch("attn_mid_fwd");
}
template <int ML>
static int mid_bwd(const AttnArgs& a, cudaStream_t st) {
  const size_t smem = sizeof(float) * (4 * ML * 32 + 2 * kH * ML + ML);
  auto k = attn_mid_bwd_kernel<ML>;
  VSET_SMEM(k, smem);
  VLAUNCH(k, dim3(a.N), dim3(ML * 4), smem, st, a);
  return check_launch("attn_mid_bwd");
}
int attn_mid_fwd(const AttnArgs& a, cudaStream_t st) {
  const int L = a.Lq > a.Lk ? a.Lq : a.Lk;
  return L <= 64 ? mid_fwd<64>(a, st) : (L <= 128 ? mid_fwd<128>(a, st) : mid_fwd<256>(a, st));
}
int attn_mid_bwd(const AttnArgs& a, cudaStream_t st) {
  const int L = a.Lq > a.Lk ? a.Lq : a.Lk;
  return L <= 64 ? mid_bwd<64>(a, st) : (L <= 128 ? mid_bwd<128>(a, st) : mid_bwd<256>(a, st));
}

}  // namespace vaesne
