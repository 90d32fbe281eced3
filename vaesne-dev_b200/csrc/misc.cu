// Small bandwidth-bound helpers of the VAESNe step: sinusoidal features, band-embedding
// gather / scatter, row expansion (K-sample and mixture-source replication), strided block
// copies (token concatenation) and the fused flat AdamW update.
//
// Reference call sites: sinusoid features util_layers.py:125-129,142-146; nn.Embedding gather
// PhotometricLayers.py:61,129; K-expansion PhotometricVAE.py:191-197 / SpectraVAE.py:189-194;
// token concatenation SpectraLayers.py:59,128; AdamW (torch.optim, decoupled weight decay) as
// constructed in cannon/test_photospectra.py:135.
#include "common.cuh"
#include "vaesne_b200.h"

namespace vaesne {

// out[t, j] = sin(x[t]*div[j]), out[t, nf+j] = cos(x[t]*div[j])
__global__ void sincos_kernel(const float* x, long long T, const float* div, int nf, float* out, long long ld) {
  const long long total = T * nf;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / nf; const int j = (int)(i - t * nf);
    const float a = x[t] * div[j];
    out[t * ld + j] = sinf(a);
    out[t * ld + nf + j] = cosf(a);
  }
}

// out[t, 0:32] (+)= table[idx[t], 0:32]
__global__ void gather_kernel(const long long* idx, long long T, const float* table, int nrows, float* out, long long ld, int acc) {
  const long long total = T * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i >> 5; const int c = (int)(i & 31);
    long long r = idx[t];
    r = r < 0 ? 0 : (r >= nrows ? nrows - 1 : r);
    const float v = table[r * 32 + c];
    float* p = out + t * ld + c;
    *p = acc ? (*p + v) : v;
  }
}

// dtable[idx[t]] += dout[t]; per-CTA partial table in shared memory (nrows <= 64)
__global__ void __launch_bounds__(256) scatter_kernel(const long long* idx, long long T, const float* dout, long long ld, float* dtable, int nrows) {
  __shared__ float sT[64 * 32];
  for (int i = threadIdx.x; i < nrows * 32; i += blockDim.x) sT[i] = 0.f;
  __syncthreads();
  const long long total = T * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i >> 5; const int c = (int)(i & 31);
    long long r = idx[t];
    r = r < 0 ? 0 : (r >= nrows ? nrows - 1 : r);
    atomicAdd(&sT[r * 32 + c], dout[t * ld + c]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nrows * 32; i += blockDim.x) atomicAdd(&dtable[i], sT[i]);
}

// dst[(c*Bs + b), :] = src[b, :]   (row r of the expanded batch reads source row r % Bs)
__global__ void expand_kernel(const float* src, long long row_elems, long long Bs, int copies, float* dst) {
  const long long total = row_elems * Bs * copies;
  const long long per = row_elems * Bs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i % per];
}
// 16-byte variants (per % 4 == 0, aligned pointers): every source vector is read once and written `copies` times — no
// per-element 64-bit modulo, four times fewer memory instructions
__global__ void expand4_kernel(const float4* src, long long per4, int copies, float4* dst) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    for (int c = 0; c < copies; ++c) dst[(long long)c * per4 + i] = v;
  }
}
__global__ void expand4_bwd_kernel(const float4* ddst, long long per4, int copies, float4* dsrc, int acc) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per4; i += (long long)gridDim.x * blockDim.x) {
    float4 s = acc ? dsrc[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < copies; ++c) {
      const float4 v = ddst[(long long)c * per4 + i];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    dsrc[i] = s;
  }
}
// dsrc[b, :] (+)= sum_c ddst[(c*Bs + b), :]
__global__ void expand_bwd_kernel(const float* ddst, long long row_elems, long long Bs, int copies, float* dsrc, int acc) {
  const long long per = row_elems * Bs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < copies; ++c) s += ddst[(long long)c * per + i];
    dsrc[i] = acc ? (dsrc[i] + s) : s;
  }
}

// dst[g*dgs + r*drs + c] (+)= src[g*sgs + r*srs + c]
__global__ void copy3d_kernel(const float* src, long long sgs, long long srs, float* dst, long long dgs, long long drs,
                              long long G, long long R, long long C, int acc) {
  const long long total = G * R * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long g = i / (R * C); const long long rem = i - g * R * C; const long long r = rem / C; const long long c = rem - r * C;
    const float v = src[g * sgs + r * srs + c];
    float* p = dst + g * dgs + r * drs + c;
    *p = acc ? (*p + v) : v;
  }
}

// torch.optim.AdamW semantics (decoupled decay, bias correction, eps outside the sqrt correction as torch does)
__global__ void adamw_kernel(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
                             float eps, float wd, const int* step_ptr, float grad_scale) {
  const float step = (float)(*step_ptr);
  const float bc1 = 1.f - powf(b1, step);
  const float bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

__global__ void step_inc_kernel(int* step, unsigned long long* seed) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (step) *step += 1;
    if (seed) *seed = *seed * 6364136223846793005ULL + 1442695040888963407ULL;
  }
}

// cell = lcg(cell); out = cell   (one dropout seed per forward call, replayable inside a CUDA graph)
__global__ void seed_next_kernel(unsigned long long* cell, unsigned long long* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long s = *cell * 6364136223846793005ULL + 1442695040888963407ULL;
    *cell = s;
    *out = s;
  }
}

static inline int ew_grid(long long total, int block) {
  long long g = (total + block - 1) / block;
  if (g > 148 * 16) g = 148 * 16;
  return g < 1 ? 1 : (int)g;
}

}  // namespace vaesne

using namespace vaesne;

extern "C" int vaesne_sincos_feat(const float* x, long long T, const float* div, int nf, float* out, long long ld, void* stream) {
  V_REQUIRE(x && div && out, V_ENULL, "sincos_feat: null argument");
  V_REQUIRE(T >= 0 && nf >= 1 && ld >= 2 * nf, V_EBADSHAPE, "sincos_feat: bad shape");
  if (T == 0) return V_OK;
  auto k = sincos_kernel;
  VLAUNCH(k, dim3(ew_grid(T * nf, 256)), dim3(256), 0, (cudaStream_t)stream, x, T, div, nf, out, ld);
  return check_launch("sincos_feat");
}

extern "C" int vaesne_gather_rows(const long long* idx, long long T, const float* table, int nrows, float* out, long long ld, int accumulate, void* stream) {
  V_REQUIRE(idx && table && out, V_ENULL, "gather_rows: null argument");
  V_REQUIRE(nrows >= 1 && ld >= 32, V_EBADSHAPE, "gather_rows: bad shape");
  if (T == 0) return V_OK;
  auto k = gather_kernel;
  VLAUNCH(k, dim3(ew_grid(T * 32, 256)), dim3(256), 0, (cudaStream_t)stream, idx, T, table, nrows, out, ld, accumulate);
  return check_launch("gather_rows");
}

extern "C" int vaesne_scatter_rows(const long long* idx, long long T, const float* dout, long long ld, float* dtable, int nrows, void* stream) {
  V_REQUIRE(idx && dout && dtable, V_ENULL, "scatter_rows: null argument");
  V_REQUIRE(nrows >= 1 && nrows <= 64, V_EUNSUPPORTED, "scatter_rows: at most 64 embedding rows (got %d)", nrows);
  if (T == 0) return V_OK;
  auto k = scatter_kernel;
  long long g = (T * 32 + 256 * 64 - 1) / (256 * 64);
  if (g > 148 * 2) g = 148 * 2;
  if (g < 1) g = 1;
  VLAUNCH(k, dim3((int)g), dim3(256), 0, (cudaStream_t)stream, idx, T, dout, ld, dtable, nrows);
  return check_launch("scatter_rows");
}

extern "C" int vaesne_expand_rows(const float* src, long long row_elems, long long Bs, int copies, float* dst, void* stream) {
  V_REQUIRE(src && dst, V_ENULL, "expand_rows: null argument");
  if (row_elems * Bs * copies == 0) return V_OK;
  const long long per = row_elems * Bs;
  if ((per & 3) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
    auto k4 = expand4_kernel;
    VLAUNCH(k4, dim3(ew_grid(per / 4, 256)), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(src), per / 4, copies,
            reinterpret_cast<float4*>(dst));
    return check_launch("expand_rows");
  }
  auto k = expand_kernel;
  VLAUNCH(k, dim3(ew_grid(row_elems * Bs * copies, 256)), dim3(256), 0, (cudaStream_t)stream, src, row_elems, Bs, copies, dst);
  return check_launch("expand_rows");
}

extern "C" int vaesne_expand_rows_bwd(const float* ddst, long long row_elems, long long Bs, int copies, float* dsrc, int accumulate, void* stream) {
  V_REQUIRE(ddst && dsrc, V_ENULL, "expand_rows_bwd: null argument");
  if (row_elems * Bs == 0) return V_OK;
  const long long per = row_elems * Bs;
  if ((per & 3) == 0 && (((uintptr_t)ddst | (uintptr_t)dsrc) & 15) == 0) {
    auto k4 = expand4_bwd_kernel;
    VLAUNCH(k4, dim3(ew_grid(per / 4, 256)), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(ddst), per / 4, copies,
            reinterpret_cast<float4*>(dsrc), accumulate);
    return check_launch("expand_rows_bwd");
  }
  auto k = expand_bwd_kernel;
  VLAUNCH(k, dim3(ew_grid(row_elems * Bs, 256)), dim3(256), 0, (cudaStream_t)stream, ddst, row_elems, Bs, copies, dsrc, accumulate);
  return check_launch("expand_rows_bwd");
}

extern "C" int vaesne_copy3d(const float* src, long long sgs, long long srs, float* dst, long long dgs, long long drs,
                             long long G, long long R, long long C, int accumulate, void* stream) {
  V_REQUIRE(src && dst, V_ENULL, "copy3d: null argument");
  if (G * R * C == 0) return V_OK;
  auto k = copy3d_kernel;
  VLAUNCH(k, dim3(ew_grid(G * R * C, 256)), dim3(256), 0, (cudaStream_t)stream, src, sgs, srs, dst, dgs, drs, G, R, C, accumulate);
  return check_launch("copy3d");
}

extern "C" int vaesne_adamw_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                                 float eps, float weight_decay, const int* step, float grad_scale, void* stream) {
  V_REQUIRE(p && g && m && v && step, V_ENULL, "adamw_flat: null argument");
  if (n == 0) return V_OK;
  auto k = adamw_kernel;
  VLAUNCH(k, dim3(ew_grid(n, 256)), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale);
  return check_launch("adamw_flat");
}

extern "C" int vaesne_step_advance(int* step, unsigned long long* seed, void* stream) {
  auto k = step_inc_kernel;
  VLAUNCH(k, dim3(1), dim3(32), 0, (cudaStream_t)stream, step, seed);
  return check_launch("step_advance");
}

extern "C" int vaesne_seed_next(unsigned long long* cell, unsigned long long* out, void* stream) {
  V_REQUIRE(cell && out, V_ENULL, "seed_next: null argument");
  auto k = seed_next_kernel;
  VLAUNCH(k, dim3(1), dim3(32), 0, (cudaStream_t)stream, cell, out);
  return check_launch("seed_next");
}
