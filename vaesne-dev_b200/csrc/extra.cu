// Contrastive objective and device-side data augmentation.
//
//   * negInfoNCE (losses.py:98-110): F.normalize(z, dim=-1) -> logits = z1 z2^T / tau -> symmetric cross-entropy against the
//     diagonal.  The reference runs it as five ATen / cuBLAS launches; here: row normalisation, one warp per logits row for the
//     cross-entropy terms (the B x B matrix is never materialised: proj_dim is 8, a dot product is 8 FMAs), one combine kernel,
//     and the matching backward kernels.  The row form takes a label offset and separate row / column sets, so the
//     data-parallel variant (global negatives: columns = the gathered projections of all ranks) uses the same kernels.
//   * augmentation of the training scripts (cannon/test_photospectra.py:45-47,75-78, cannon/ZTF_photospect.py:46-66): repeat
//     the set `copies` times, add Gaussian noise per element and / or one Gaussian shift per row, OR random masking — one
//     kernel with a counter-based generator (common.cuh hash_ctr + Box-Muller), no host round trip.
#include "common.cuh"
#include "vaesne_b200.h"

namespace vaesne {

static inline int ew_grid2(long long n, int block) {
  long long g = (n + block - 1) / block;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

// ---- F.normalize(x, p=2, dim=-1, eps): y = x / max(||x||, eps) ---------------------------------------------------------
__global__ void l2norm_fwd_kernel(const float* x, int B, int P, float eps, float* y, float* inv) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < B; r += gridDim.x * blockDim.x) {
    float ss = 0.f;
    for (int c = 0; c < P; ++c) { const float v = x[(long long)r * P + c]; ss = fmaf(v, v, ss); }
    const float nrm = sqrtf(ss);
    const float iv = 1.f / fmaxf(nrm, eps);
    for (int c = 0; c < P; ++c) y[(long long)r * P + c] = x[(long long)r * P + c] * iv;
    inv[r] = nrm > eps ? iv : -iv;       // sign flags the clamped branch (the norm does not depend on x there)
  }
}
// dx = inv * (dy - y (y . dy))   (clamped rows: dx = dy * inv); ACCUMULATES into dx when acc != 0
__global__ void l2norm_bwd_kernel(const float* y, const float* inv, const float* dy, int B, int P, float* dx, int acc) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < B; r += gridDim.x * blockDim.x) {
    const float iv = inv[r];
    float dot = 0.f;
    if (iv > 0.f) for (int c = 0; c < P; ++c) dot = fmaf(y[(long long)r * P + c], dy[(long long)r * P + c], dot);
    for (int c = 0; c < P; ++c) {
      const long long o = (long long)r * P + c;
      const float g = fabsf(iv) * (dy[o] - (iv > 0.f ? y[o] * dot : 0.f));
      dx[o] = acc ? dx[o] + g : g;
    }
  }
}

constexpr int kMaxP = 64;
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per row i of A: lse[i] = logsumexp_j (inv_tau * a_i . b_j), loss[i] = lse[i] - inv_tau * a_i . b_{i + off}
__global__ void __launch_bounds__(128) ce_rows_fwd_kernel(const float* A, int n, const float* Bm, int m, int P, float inv_tau, int off,
                                                         float* lse, float* loss) {
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += gridDim.x * (blockDim.x >> 5)) {
    float a[kMaxP];
    for (int c = 0; c < P; ++c) a[c] = A[(long long)i * P + c] * inv_tau;
    float mx = -INFINITY;
    for (int j = lane; j < m; j += 32) {
      float d = 0.f;
      for (int c = 0; c < P; ++c) d = fmaf(a[c], Bm[(long long)j * P + c], d);
      mx = fmaxf(mx, d);
    }
    mx = warp_max(mx);
    float sum = 0.f, diag = 0.f;
    for (int j = lane; j < m; j += 32) {
      float d = 0.f;
      for (int c = 0; c < P; ++c) d = fmaf(a[c], Bm[(long long)j * P + c], d);
      sum += expf(d - mx);
      if (j == i + off) diag = d;
    }
    sum = warp_sum(sum); diag = warp_sum(diag);
    if (lane == 0) { const float l = mx + logf(sum); lse[i] = l; loss[i] = l - diag; }
  }
}
// dA[i] = w * inv_tau * sum_j (softmax_ij - [j == i + off]) b_j          (one warp per row i; overwrites or accumulates)
__global__ void __launch_bounds__(128) ce_rows_bwd_a_kernel(const float* A, int n, const float* Bm, int m, int P, float inv_tau, int off,
                                                           const float* lse, float w, const float* gptr, float* dA, int acc) {
  const int lane = threadIdx.x & 31;
  const float ww = w * (gptr ? *gptr : 1.f) * inv_tau;
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += gridDim.x * (blockDim.x >> 5)) {
    float a[kMaxP], g[kMaxP];
    for (int c = 0; c < P; ++c) { a[c] = A[(long long)i * P + c] * inv_tau; g[c] = 0.f; }
    const float l = lse[i];
    for (int j = lane; j < m; j += 32) {
      float d = 0.f;
      for (int c = 0; c < P; ++c) d = fmaf(a[c], Bm[(long long)j * P + c], d);
      const float p = expf(d - l) - (j == i + off ? 1.f : 0.f);
      for (int c = 0; c < P; ++c) g[c] = fmaf(p, Bm[(long long)j * P + c], g[c]);
    }
    for (int c = 0; c < P; ++c) {
      const float s = warp_sum(g[c]);
      if (lane == 0) { float* o = dA + (long long)i * P + c; *o = acc ? *o + ww * s : ww * s; }
    }
  }
}
// dB[j] = w * inv_tau * sum_i (softmax_ij - [j == i + off]) a_i          (one warp per column j)
__global__ void __launch_bounds__(128) ce_rows_bwd_b_kernel(const float* A, int n, const float* Bm, int m, int P, float inv_tau, int off,
                                                           const float* lse, float w, const float* gptr, float* dB, int acc) {
  const int lane = threadIdx.x & 31;
  const float ww = w * (gptr ? *gptr : 1.f) * inv_tau;
  for (int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < m; j += gridDim.x * (blockDim.x >> 5)) {
    float b[kMaxP], g[kMaxP];
    for (int c = 0; c < P; ++c) { b[c] = Bm[(long long)j * P + c] * inv_tau; g[c] = 0.f; }
    for (int i = lane; i < n; i += 32) {
      float d = 0.f;
      for (int c = 0; c < P; ++c) d = fmaf(b[c], A[(long long)i * P + c], d);
      const float p = expf(d - lse[i]) - (j == i + off ? 1.f : 0.f);
      for (int c = 0; c < P; ++c) g[c] = fmaf(p, A[(long long)i * P + c], g[c]);
    }
    for (int c = 0; c < P; ++c) {
      const float s = warp_sum(g[c]);
      if (lane == 0) { float* o = dB + (long long)j * P + c; *o = acc ? *o + ww * s : ww * s; }
    }
  }
}
// out = scale * (sum a [+ sum b])        (single CTA; n is a batch size)
__global__ void __launch_bounds__(256) sum_scale_kernel(const float* a, const float* b, int n, float scale, float* out) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += a[i] + (b ? b[i] : 0.f);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < 8; ++w) t += red[w]; *out = scale * t; }
}

// ---- augmentation --------------------------------------------------------------------------------------------------
// Row r of the output reads source row r % B.  x_out = x + sigma_elem * N(0,1) [per element] + sigma_row * N(0,1) [one draw per
// output row]; mask_out = mask | (U < mask_p).  Normals by Box-Muller on two 24-bit uniforms of a counter hash; every
// (stream, element) pair has its own counter, so the result does not depend on the launch geometry.
__device__ __forceinline__ float u01(uint32_t h) { return ((float)(h >> 8) + 0.5f) * (1.f / 16777216.f); }
__device__ __forceinline__ float normal_from(uint32_t h1, uint32_t h2) {
  return sqrtf(-2.f * logf(u01(h1))) * cosf(6.283185307179586f * u01(h2));
}
__global__ void augment_kernel(const float* x, const unsigned char* mask, long long R, long long B, int L, float sigma_elem,
                               float sigma_row, float mask_p, const unsigned long long* seed, uint32_t stream, float* xo, unsigned char* mo) {
  const unsigned long long sd = *seed;
  const uint32_t s0 = (uint32_t)sd, s1 = (uint32_t)(sd >> 32);
  const long long total = R * L;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / L; const int l = (int)(i - r * L);
    const long long src = (r % B) * L + l;
    if (xo) {
      float v = x[src];
      if (sigma_elem != 0.f) v += sigma_elem * normal_from(hash_ctr(s0, s1, stream, (uint64_t)i * 2), hash_ctr(s0, s1, stream, (uint64_t)i * 2 + 1));
      if (sigma_row != 0.f) v += sigma_row * normal_from(hash_ctr(s1, s0, stream ^ 0x68bc21ebu, (uint64_t)r * 2), hash_ctr(s1, s0, stream ^ 0x68bc21ebu, (uint64_t)r * 2 + 1));
      xo[i] = v;
    }
    if (mo) {
      unsigned char mk = mask ? mask[src] : 0;
      if (mask_p > 0.f && u01(hash_ctr(s0, s1, stream ^ 0x2545f491u, (uint64_t)i)) < mask_p) mk = 1;
      mo[i] = mk;
    }
  }
}

// ---- long attention as key blocks: O = sum_b exp(LSE_b - LSE) O_b, LSE = logsumexp_b LSE_b ---------------------------------
struct CombineParts { const float* O[8]; const float* L[8]; };
__global__ void attn_combine_kernel(CombineParts p, int nparts, long long N, int Lb, int Lq, int q0, float* O, long long ldo, float* LSE) {
  const long long total = N * Lb * 4;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int h = (int)(t & 3);
    const long long ni = t >> 2;
    const long long n = ni / Lb; const int i = (int)(ni - n * Lb);
    float l[8], m = -INFINITY;
    for (int b = 0; b < nparts; ++b) { l[b] = p.L[b][(n * 4 + h) * Lb + i]; m = fmaxf(m, l[b]); }
    float sum = 0.f, acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    for (int b = 0; b < nparts; ++b) {
      if (l[b] == -INFINITY) continue;                       // a block with every key masked contributes nothing
      const float w = expf(l[b] - m);
      sum += w;
      const float* o = p.O[b] + (n * Lb + i) * 32 + h * 8;
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = fmaf(w, o[c], acc[c]);
    }
    float* dst = O + (n * Lq + q0 + i) * ldo + h * 8;
    const float inv = 1.f / sum;                              // every block empty: 0 / 0 = NaN, as the unblocked kernels give
#pragma unroll
    for (int c = 0; c < 8; ++c) dst[c] = acc[c] * inv;
    LSE[(n * 4 + h) * Lq + q0 + i] = m + logf(sum);
  }
}

}  // namespace vaesne
using namespace vaesne;

extern "C" int vaesne_attn_combine(const float* const* O_parts, const float* const* LSE_parts, int nparts, long long N, int Lb, int Lq, int q0,
                                   float* O, long long ldo, float* LSE, void* stream) {
  V_REQUIRE(O_parts && LSE_parts && O && LSE, V_ENULL, "attn_combine: null argument");
  V_REQUIRE(nparts >= 1 && nparts <= 8, V_EUNSUPPORTED, "attn_combine: 1..8 key blocks (got %d)", nparts);
  if (N * Lb == 0) return V_OK;
  CombineParts p{};
  for (int b = 0; b < nparts; ++b) { p.O[b] = O_parts[b]; p.L[b] = LSE_parts[b]; }
  auto k = attn_combine_kernel;
  VLAUNCH(k, dim3(ew_grid2(N * Lb * 4, 256)), dim3(256), 0, (cudaStream_t)stream, p, nparts, N, Lb, Lq, q0, O, ldo, LSE);
  return check_launch("attn_combine");
}

extern "C" int vaesne_l2norm_fwd(const float* x, int B, int P, float eps, float* y, float* inv_norm, void* stream) {
  V_REQUIRE(x && y && inv_norm, V_ENULL, "l2norm_fwd: null argument");
  V_REQUIRE(P >= 1, V_EBADSHAPE, "l2norm_fwd: bad shape");
  if (B == 0) return V_OK;
  auto k = l2norm_fwd_kernel;
  VLAUNCH(k, dim3(ew_grid2(B, 128)), dim3(128), 0, (cudaStream_t)stream, x, B, P, eps, y, inv_norm);
  return check_launch("l2norm_fwd");
}
extern "C" int vaesne_l2norm_bwd(const float* y, const float* inv_norm, const float* dy, int B, int P, float* dx, int accumulate, void* stream) {
  V_REQUIRE(y && inv_norm && dy && dx, V_ENULL, "l2norm_bwd: null argument");
  if (B == 0) return V_OK;
  auto k = l2norm_bwd_kernel;
  VLAUNCH(k, dim3(ew_grid2(B, 128)), dim3(128), 0, (cudaStream_t)stream, y, inv_norm, dy, B, P, dx, accumulate);
  return check_launch("l2norm_bwd");
}
extern "C" int vaesne_ce_rows_fwd(const float* A, int n, const float* Bm, int m, int P, float inv_tau, int label_off, float* lse, float* loss, void* stream) {
  V_REQUIRE(A && Bm && lse && loss, V_ENULL, "ce_rows_fwd: null argument");
  V_REQUIRE(P >= 1 && P <= kMaxP && m >= 1, V_EUNSUPPORTED, "ce_rows_fwd: projection width 1..%d (got %d), at least one column", kMaxP, P);
  if (n == 0) return V_OK;
  auto k = ce_rows_fwd_kernel;
  VLAUNCH(k, dim3(ew_grid2(n, 4)), dim3(128), 0, (cudaStream_t)stream, A, n, Bm, m, P, inv_tau, label_off, lse, loss);
  return check_launch("ce_rows_fwd");
}
extern "C" int vaesne_ce_rows_bwd(const float* A, int n, const float* Bm, int m, int P, float inv_tau, int label_off, const float* lse,
                                  float w, const float* gptr, float* dA, int dA_acc, float* dB, int dB_acc, void* stream) {
  V_REQUIRE(A && Bm && lse, V_ENULL, "ce_rows_bwd: null argument");
  V_REQUIRE(P >= 1 && P <= kMaxP, V_EUNSUPPORTED, "ce_rows_bwd: projection width 1..%d (got %d)", kMaxP, P);
  if (n == 0 || m == 0) return V_OK;
  if (dA) {
    auto k = ce_rows_bwd_a_kernel;
    VLAUNCH(k, dim3(ew_grid2(n, 4)), dim3(128), 0, (cudaStream_t)stream, A, n, Bm, m, P, inv_tau, label_off, lse, w, gptr, dA, dA_acc);
    int rc = check_launch("ce_rows_bwd_a"); if (rc) return rc;
  }
  if (dB) {
    auto k = ce_rows_bwd_b_kernel;
    VLAUNCH(k, dim3(ew_grid2(m, 4)), dim3(128), 0, (cudaStream_t)stream, A, n, Bm, m, P, inv_tau, label_off, lse, w, gptr, dB, dB_acc);
    return check_launch("ce_rows_bwd_b");
  }
  return V_OK;
}
extern "C" int vaesne_sum_scale(const float* a, const float* b, int n, float scale, float* out, void* stream) {
  V_REQUIRE(a && out, V_ENULL, "sum_scale: null argument");
  auto k = sum_scale_kernel;
  VLAUNCH(k, dim3(1), dim3(256), 0, (cudaStream_t)stream, a, b, n, scale, out);
  return check_launch("sum_scale");
}
extern "C" int vaesne_augment(const float* x, const unsigned char* mask, long long R, long long B, int L, float sigma_elem, float sigma_row,
                              float mask_p, const unsigned long long* seed, uint32_t stream_id, float* x_out, unsigned char* mask_out, void* stream) {
  V_REQUIRE(seed && (x_out || mask_out) && (!x_out || x), V_ENULL, "augment: null argument");
  V_REQUIRE(B >= 1 && L >= 1 && R >= 0, V_EBADSHAPE, "augment: bad shape");
  if (R == 0) return V_OK;
  auto k = augment_kernel;
  VLAUNCH(k, dim3(ew_grid2(R * L, 256)), dim3(256), 0, (cudaStream_t)stream, x, mask, R, B, L, sigma_elem, sigma_row, mask_p, seed, stream_id, x_out, mask_out);
  return check_launch("augment");
}
