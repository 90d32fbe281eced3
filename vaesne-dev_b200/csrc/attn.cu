// Masked multi-head attention (head_dim 8, 4 heads), flash style, forward + backward.
//
// Restates nn.MultiheadAttention as called by the reference's TransformerBlock
// (/root/reference/package/VAESNe/util_layers.py:289,297,301; torch semantics in
// torch/nn/functional.py multi_head_attention_forward): q scaled by 1/sqrt(dh) before QK^T,
// boolean key-padding mask applied as additive -inf *inside* the kernel (never materialised),
// softmax over keys, dropout on the probabilities, PV.  The [Lq, Lk] probability matrix only
// ever exists as per-thread registers; the backward recomputes it from the saved log-sum-exp.
//
// This is the general-shape kernel (any Lq, Lk; used for cross-attention with 4-5 or 60-983
// keys, 8-query encoder attention and 60x60 photometry self-attention).  The 982/983-token
// self-attention hot spot has a dedicated tcgen05 kernel (attn_tc.cu) that takes over when the
// shape qualifies; this kernel is also its reference in the GPU tests.
//
// Mapping: one query (fwd, dQ) or one key (dK/dV) per LANES consecutive lanes; K/V (resp.
// Q/dO/lse/delta) tiles of 128 rows are staged in shared memory and read as broadcasts.
#include "common.cuh"
#include "vaesne_b200.h"
#include "attn_args.cuh"

namespace vaesne {

constexpr int AT = 128;    // threads per CTA and rows per shared-memory tile

__device__ __forceinline__ void ld8(float* d, const float* p) {
  if (((uintptr_t)p & 15) == 0) {
    float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) d[c] = p[c];
  }
}
__device__ __forceinline__ void st8(float* p, const float* d) {
  if (((uintptr_t)p & 15) == 0) {
    reinterpret_cast<float4*>(p)[0] = make_float4(d[0], d[1], d[2], d[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(d[4], d[5], d[6], d[7]);
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) p[c] = d[c];
  }
}
__device__ __forceinline__ float dot8(const float* a, const float* b) {
  float s = a[0] * b[0];
#pragma unroll
  for (int c = 1; c < 8; ++c) s = fmaf(a[c], b[c], s);
  return s;
}
__device__ __forceinline__ float key_bias(const AttnArgs& a, int n, int j) {
  if (j >= a.Lk) return -INFINITY;
  if (a.mask && j < a.mask_len && a.mask[(long long)(n % a.mask_rows) * a.mask_len + j]) return -INFINITY;
  return 0.f;
}

// ------------------------------------------------------------------------------------------
template <int LANES>
__global__ void __launch_bounds__(AT) attn_fwd_kernel(AttnArgs a) {
  __shared__ __align__(16) float sK[AT][8];
  __shared__ __align__(16) float sV[AT][8];
  __shared__ __align__(16) float sB[AT];
  const int tid = threadIdx.x, sub = tid % LANES;
  const int h = blockIdx.y, n = blockIdx.z;
  const int i = blockIdx.x * (AT / LANES) + tid / LANES;
  const bool valid = i < a.Lq;
  const float qscale = 0.35355339059327373f * kLog2e;    // sqrt(1/8), exponent in base 2
  float q[8], o[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { q[c] = 0.f; o[c] = 0.f; }
  if (valid) {
    ld8(q, a.q + ((long long)n * a.Lq + i) * a.ldq + h * 8);
#pragma unroll
    for (int c = 0; c < 8; ++c) q[c] *= qscale;
  }
  float m = -INFINITY, l = 0.f;
  DropCfg dc = make_drop(a.p_drop, a.seed, a.stream_id);
  const uint64_t drow = ((uint64_t)(n * kH + h) * a.Lq + (valid ? i : 0)) * (uint64_t)a.Lk;

  for (int j0 = 0; j0 < a.Lk; j0 += AT) {
    __syncthreads();
    {
      const int j = j0 + tid;
      if (j < a.Lk) {
        ld8(sK[tid], a.k + ((long long)n * a.Lk + j) * a.ldk + h * 8);
        ld8(sV[tid], a.v + ((long long)n * a.Lk + j) * a.ldv + h * 8);
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) { sK[tid][c] = 0.f; sV[tid][c] = 0.f; }
      }
      sB[tid] = key_bias(a, n, j);
    }
    __syncthreads();
    const int nk = min(AT, a.Lk - j0);
    for (int jj = sub; jj < nk; jj += LANES) {
      const float s = dot8(q, sK[jj]) + sB[jj];
      if (s > m) {
        const float alpha = exp2f(m - s);     // m = -inf -> 0
        l *= alpha;
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] *= alpha;
        m = s;
      }
      const float p = (s == -INFINITY) ? 0.f : exp2f(s - m);
      l += p;
      float pd = p;
      if (dc.on) pd *= drop_mult(dc, drow + (uint64_t)(j0 + jj));
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = fmaf(pd, sV[jj][c], o[c]);
    }
  }
  if (LANES > 1) {
#pragma unroll
    for (int off = LANES / 2; off >= 1; off >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, off);
      const float l2 = __shfl_xor_sync(0xffffffffu, l, off);
      const float mn = fmaxf(m, m2);
      const float a1 = (m == -INFINITY) ? 0.f : exp2f(m - mn);
      const float a2 = (m2 == -INFINITY) ? 0.f : exp2f(m2 - mn);
      l = l * a1 + l2 * a2;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float o2 = __shfl_xor_sync(0xffffffffu, o[c], off);
        o[c] = o[c] * a1 + o2 * a2;
      }
      m = mn;
    }
  }
  if (valid && sub == 0) {
    const float inv = 1.f / l;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] *= inv;
    st8(a.O + ((long long)n * a.Lq + i) * a.ldo + h * 8, o);
    a.LSE[((long long)n * kH + h) * a.Lq + i] = (m + log2f(l)) * kLn2;
  }
}

// ------------------------------------------------------------------------------------------
template <int LANES>
__global__ void __launch_bounds__(AT) attn_bwd_dq_kernel(AttnArgs a) {
  __shared__ __align__(16) float sK[AT][8];
  __shared__ __align__(16) float sV[AT][8];
  __shared__ __align__(16) float sB[AT];
  const int tid = threadIdx.x, sub = tid % LANES;
  const int h = blockIdx.y, n = blockIdx.z;
  const int i = blockIdx.x * (AT / LANES) + tid / LANES;
  const bool valid = i < a.Lq;
  const float scale = 0.35355339059327373f;
  float q[8], dO[8], dq[8];
  float lse2 = INFINITY, delta = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) { q[c] = 0.f; dO[c] = 0.f; dq[c] = 0.f; }
  if (valid) {
    float o[8];
    ld8(q, a.q + ((long long)n * a.Lq + i) * a.ldq + h * 8);
    ld8(dO, a.dO + ((long long)n * a.Lq + i) * a.lddo + h * 8);
    ld8(o, a.O + ((long long)n * a.Lq + i) * a.ldo + h * 8);
#pragma unroll
    for (int c = 0; c < 8; ++c) q[c] *= scale * kLog2e;
    delta = dot8(dO, o);
    lse2 = a.LSE[((long long)n * kH + h) * a.Lq + i] * kLog2e;
  }
  DropCfg dc = make_drop(a.p_drop, a.seed, a.stream_id);
  const uint64_t drow = ((uint64_t)(n * kH + h) * a.Lq + (valid ? i : 0)) * (uint64_t)a.Lk;

  for (int j0 = 0; j0 < a.Lk; j0 += AT) {
    __syncthreads();
    {
      const int j = j0 + tid;
      if (j < a.Lk) {
        ld8(sK[tid], a.k + ((long long)n * a.Lk + j) * a.ldk + h * 8);
        ld8(sV[tid], a.v + ((long long)n * a.Lk + j) * a.ldv + h * 8);
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) { sK[tid][c] = 0.f; sV[tid][c] = 0.f; }
      }
      sB[tid] = key_bias(a, n, j);
    }
    __syncthreads();
    const int nk = min(AT, a.Lk - j0);
    for (int jj = sub; jj < nk; jj += LANES) {
      const float s = dot8(q, sK[jj]) + sB[jj];
      const float p = exp2f(s - lse2);          // -inf bias or +inf lse2 -> 0
      float dp = dot8(dO, sV[jj]);
      if (dc.on) dp *= drop_mult(dc, drow + (uint64_t)(j0 + jj));
      const float ds = p * (dp - delta);
#pragma unroll
      for (int c = 0; c < 8; ++c) dq[c] = fmaf(ds, sK[jj][c], dq[c]);
    }
  }
  if (LANES > 1) {
#pragma unroll
    for (int off = LANES / 2; off >= 1; off >>= 1) {
#pragma unroll
      for (int c = 0; c < 8; ++c) dq[c] += __shfl_xor_sync(0xffffffffu, dq[c], off);
    }
  }
  if (valid && sub == 0) {
#pragma unroll
    for (int c = 0; c < 8; ++c) dq[c] *= scale;
    st8(a.dq + ((long long)n * a.Lq + i) * a.lddq + h * 8, dq);
    a.delta[((long long)n * kH + h) * a.Lq + i] = delta;
  }
}

// ------------------------------------------------------------------------------------------
template <int LANES>
__global__ void __launch_bounds__(AT) attn_bwd_dkv_kernel(AttnArgs a) {
  __shared__ __align__(16) float sQ[AT][8];
  __shared__ __align__(16) float sDO[AT][8];
  __shared__ __align__(16) float sL[AT];
  __shared__ __align__(16) float sD[AT];
  const int tid = threadIdx.x, sub = tid % LANES;
  const int h = blockIdx.y, n = blockIdx.z;
  const int j = blockIdx.x * (AT / LANES) + tid / LANES;
  const bool valid = j < a.Lk;
  const float scale = 0.35355339059327373f;
  float k[8], v[8], dk[8], dv[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { k[c] = 0.f; v[c] = 0.f; dk[c] = 0.f; dv[c] = 0.f; }
  float bias = -INFINITY;
  if (valid) {
    ld8(k, a.k + ((long long)n * a.Lk + j) * a.ldk + h * 8);
    ld8(v, a.v + ((long long)n * a.Lk + j) * a.ldv + h * 8);
    bias = key_bias(a, n, j);
  }
  DropCfg dc = make_drop(a.p_drop, a.seed, a.stream_id);
  const uint64_t dbase = (uint64_t)(n * kH + h) * a.Lq;

  for (int i0 = 0; i0 < a.Lq; i0 += AT) {
    __syncthreads();
    {
      const int i = i0 + tid;
      if (i < a.Lq) {
        ld8(sQ[tid], a.q + ((long long)n * a.Lq + i) * a.ldq + h * 8);
        ld8(sDO[tid], a.dO + ((long long)n * a.Lq + i) * a.lddo + h * 8);
#pragma unroll
        for (int c = 0; c < 8; ++c) sQ[tid][c] *= scale * kLog2e;
        sL[tid] = a.LSE[((long long)n * kH + h) * a.Lq + i] * kLog2e;
        sD[tid] = a.delta[((long long)n * kH + h) * a.Lq + i];
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) { sQ[tid][c] = 0.f; sDO[tid][c] = 0.f; }
        sL[tid] = INFINITY; sD[tid] = 0.f;
      }
    }
    __syncthreads();
    const int nq = min(AT, a.Lq - i0);
    for (int ii = sub; ii < nq; ii += LANES) {
      const float s = dot8(sQ[ii], k) + bias;
      const float p = exp2f(s - sL[ii]);
      float dm = 1.f;
      if (dc.on) dm = drop_mult(dc, (dbase + (uint64_t)(i0 + ii)) * (uint64_t)a.Lk + (uint64_t)(valid ? j : 0));
      const float pd = p * dm;
      const float dp = dot8(sDO[ii], v) * dm;
      const float ds = p * (dp - sD[ii]);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        dv[c] = fmaf(pd, sDO[ii][c], dv[c]);
        dk[c] = fmaf(ds, sQ[ii][c], dk[c]);
      }
    }
  }
  if (LANES > 1) {
#pragma unroll
    for (int off = LANES / 2; off >= 1; off >>= 1) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        dk[c] += __shfl_xor_sync(0xffffffffu, dk[c], off);
        dv[c] += __shfl_xor_sync(0xffffffffu, dv[c], off);
      }
    }
  }
  if (valid && sub == 0) {
#pragma unroll
    for (int c = 0; c < 8; ++c) dk[c] *= kLn2;     // sQ carried log2(e)
    st8(a.dk + ((long long)n * a.Lk + j) * a.lddk + h * 8, dk);
    st8(a.dv + ((long long)n * a.Lk + j) * a.lddv + h * 8, dv);
  }
}

static int check_common(const AttnArgs& a, const char* what) {
  V_REQUIRE(a.q && a.k && a.v, V_ENULL, "%s: null q/k/v", what);
  V_REQUIRE(a.N >= 0 && a.Lq >= 1 && a.Lk >= 1, V_EBADSHAPE, "%s: bad shape N=%d Lq=%d Lk=%d", what, a.N, a.Lq, a.Lk);
  V_REQUIRE(a.N <= 65535, V_EBADSHAPE, "%s: N=%d exceeds grid.z (split the batch)", what, a.N);
  V_REQUIRE(a.mask == nullptr || (a.mask_rows >= 1 && a.mask_len >= 0), V_EBADSHAPE, "%s: bad mask geometry", what);
  return V_OK;
}

}  // namespace vaesne

using namespace vaesne;

extern "C" int vaesne_attn_fwd_ex(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv,
                               int N, int Lq, int Lk, const unsigned char* mask, int mask_rows, int mask_len,
                               float p_drop, const uint64_t* seed, uint32_t stream_id,
                               float* O, long long ldo, float* LSE, int flags, void* stream) {
  AttnArgs a{};
  a.q = q; a.ldq = ldq; a.k = k; a.ldk = ldk; a.v = v; a.ldv = ldv; a.N = N; a.Lq = Lq; a.Lk = Lk;
  a.mask = mask; a.mask_rows = mask_rows; a.mask_len = mask_len; a.p_drop = p_drop; a.seed = seed; a.stream_id = stream_id;
  a.O = O; a.ldo = ldo; a.LSE = LSE; a.flags = flags;
  if (N == 0) return V_OK;                  // an empty batch has no buffers: torch hands out null pointers for it
  int rc = check_common(a, "attn_fwd"); if (rc) return rc;
#ifndef VAESNE_EMU
  V_REQUIRE(flags == 0 || attn_tc_eligible(a), V_EUNSUPPORTED, "attn_fwd: key-block calls (flags) are served by the tcgen05 kernels only (96 <= Lq, Lk <= 1024)");
#else
  V_REQUIRE(flags == 0, V_EUNSUPPORTED, "attn_fwd: key-block calls (flags) are served by the tcgen05 kernels only");
#endif
  V_REQUIRE(O && LSE, V_ENULL, "attn_fwd: null O/LSE");
  cudaStream_t st = (cudaStream_t)stream;
#ifndef VAESNE_EMU
  // the forward and backward of one attention call must pick the same path (their dropout masks differ)
  if (attn_tc_eligible(a) && (a.p_drop == 0.f || attn_tc_has_bwd())) return attn_tc_fwd(a, st);
#endif
  if (attn_small_eligible(a)) return attn_small_fwd(a, st);
  if (attn_mid_eligible(a)) return attn_mid_fwd(a, st);
  dim3 block(AT);
  if (Lq <= 32) { dim3 grid((Lq + 3) / 4, kH, N); auto kf = attn_fwd_kernel<32>; VLAUNCH(kf, grid, block, 0, st, a); }
  else { dim3 grid((Lq + AT - 1) / AT, kH, N); auto kf = attn_fwd_kernel<1>; VLAUNCH(kf, grid, block, 0, st, a); }
  return check_launch("attn_fwd");
}

extern "C" int vaesne_attn_fwd(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv,
                               int N, int Lq, int Lk, const unsigned char* mask, int mask_rows, int mask_len,
                               float p_drop, const uint64_t* seed, uint32_t stream_id,
                               float* O, long long ldo, float* LSE, void* stream) {
  return vaesne_attn_fwd_ex(q, ldq, k, ldk, v, ldv, N, Lq, Lk, mask, mask_rows, mask_len, p_drop, seed, stream_id, O, ldo, LSE, 0, stream);
}

extern "C" int vaesne_attn_bwd_ex(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv,
                               int N, int Lq, int Lk, const unsigned char* mask, int mask_rows, int mask_len,
                               float p_drop, const uint64_t* seed, uint32_t stream_id,
                               const float* O, long long ldo, const float* LSE, const float* dO, long long lddo,
                               float* delta_ws, float* dq, long long lddq, float* dk, long long lddk, float* dv, long long lddv,
                               int flags, void* stream) {
  AttnArgs a{};
  a.q = q; a.ldq = ldq; a.k = k; a.ldk = ldk; a.v = v; a.ldv = ldv; a.N = N; a.Lq = Lq; a.Lk = Lk;
  a.mask = mask; a.mask_rows = mask_rows; a.mask_len = mask_len; a.p_drop = p_drop; a.seed = seed; a.stream_id = stream_id;
  a.O = const_cast<float*>(O); a.ldo = ldo; a.LSE = const_cast<float*>(LSE); a.dO = dO; a.lddo = lddo; a.delta = delta_ws;
  a.dq = dq; a.lddq = lddq; a.dk = dk; a.lddk = lddk; a.dv = dv; a.lddv = lddv; a.flags = flags;
  if (N == 0) return V_OK;
  int rc = check_common(a, "attn_bwd"); if (rc) return rc;
#ifndef VAESNE_EMU
  V_REQUIRE(flags == 0 || attn_tc_eligible(a), V_EUNSUPPORTED, "attn_bwd: key-block calls (flags) are served by the tcgen05 kernels only (96 <= Lq, Lk <= 1024)");
#else
  V_REQUIRE(flags == 0, V_EUNSUPPORTED, "attn_bwd: key-block calls (flags) are served by the tcgen05 kernels only");
#endif
  V_REQUIRE(O && LSE && dO && delta_ws && dq && dk && dv, V_ENULL, "attn_bwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
#ifndef VAESNE_EMU
  if (attn_tc_eligible(a) && attn_tc_has_bwd()) return attn_tc_bwd(a, st);
#endif
  if (attn_small_eligible(a)) return attn_small_bwd(a, st);
  if (attn_mid_eligible(a)) return attn_mid_bwd(a, st);
  dim3 block(AT);
  if (Lq <= 32) { dim3 grid((Lq + 3) / 4, kH, N); auto kf = attn_bwd_dq_kernel<32>; VLAUNCH(kf, grid, block, 0, st, a); }
  else { dim3 grid((Lq + AT - 1) / AT, kH, N); auto kf = attn_bwd_dq_kernel<1>; VLAUNCH(kf, grid, block, 0, st, a); }
  rc = check_launch("attn_bwd_dq"); if (rc) return rc;
  if (Lk <= 32) { dim3 grid((Lk + 3) / 4, kH, N); auto kf = attn_bwd_dkv_kernel<32>; VLAUNCH(kf, grid, block, 0, st, a); }
  else { dim3 grid((Lk + AT - 1) / AT, kH, N); auto kf = attn_bwd_dkv_kernel<1>; VLAUNCH(kf, grid, block, 0, st, a); }
  return check_launch("attn_bwd_dkv");
}

extern "C" int vaesne_attn_bwd(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv,
                               int N, int Lq, int Lk, const unsigned char* mask, int mask_rows, int mask_len,
                               float p_drop, const uint64_t* seed, uint32_t stream_id,
                               const float* O, long long ldo, const float* LSE, const float* dO, long long lddo,
                               float* delta_ws, float* dq, long long lddq, float* dk, long long lddk, float* dv, long long lddv,
                               void* stream) {
  return vaesne_attn_bwd_ex(q, ldq, k, ldk, v, ldv, N, Lq, Lk, mask, mask_rows, mask_len, p_drop, seed, stream_id, O, ldo, LSE, dO, lddo,
                            delta_ws, dq, lddq, dk, lddk, dv, lddv, 0, stream);
}
