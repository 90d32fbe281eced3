// Posterior heads, reparameterised sampling, mixture-of-experts latent terms, likelihoods and the
// IW-ELBO / ELBO reductions, forward and backward.
//
// Restates (paths under /root/reference/package/VAESNe):
//   posterior heads  mu = tokens[:T], scale = softplus(tokens[T:])   PhotometricVAE.py:53-54, SpectraVAE.py:48-49
//   rsample          Laplace: z = mu - s*sign(u)*log1p(-|u|), Normal: z = mu + s*eps  (torch/distributions)
//   _m_iwae          lw = log p(z) + sum_d scaling_d * log p(x_d|z) - log mean_m q_m(z)   losses.py:47-62
//   m_iwae           sum_b ( logsumexp_{M*K} lw - log(M*K) )                               losses.py:78-93
//   elbo             mean_{K,B}( sum_L lpx*scaling - sum_{T,Z} KL(q||p) )                  losses.py:16-24
//   likelihood scale 1 + 1e8*mask (photometry) / 1 + 1e10*mask (spectra), evaluated in fp32 by the host
//                    and passed as `scale_masked`                    PhotometricVAE.py:91-93, SpectraVAE.py:84-86
// family codes: 0 = Laplace, 1 = Normal.
#include "common.cuh"
#include "vaesne_b200.h"

namespace vaesne {

constexpr int MAXM = 4;

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // F.softplus(beta=1, threshold=20)
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float sgn(float x) { return (x > 0.f) - (x < 0.f); }

__device__ __forceinline__ float logp(int fam, float x, float mu, float s) {
  const float d = x - mu;
  if (fam == 0) return -logf(2.f * s) - fabsf(d) / s;
  return -(d * d) / (2.f * s * s) - logf(s) - 0.9189385332046727f;
}
// derivatives of log p wrt (x, s); d/dmu = -d/dx
__device__ __forceinline__ void dlogp(int fam, float x, float mu, float s, float& dx, float& ds) {
  const float d = x - mu;
  if (fam == 0) { dx = -sgn(d) / s; ds = -1.f / s + fabsf(d) / (s * s); }
  else { dx = -d / (s * s); ds = -1.f / s + d * d / (s * s * s); }
}
__device__ __forceinline__ float noise_eps(int fam, float u) {
  return fam == 0 ? -sgn(u) * log1pf(-fabsf(u)) : u;
}

struct LatentArgs {
  int M, K, B, T, Z;
  const float* bott[MAXM];     // [B, 2T, Z]
  const float* noise[MAXM];    // [K, B, T, Z]
  int fam_post[MAXM];
  int fam_prior; const float* pz_mu; const float* pz_s;    // [T, Z]
  float* z;                    // [M, K, B, T, Z]
  float* mu[MAXM]; float* s[MAXM];   // [B, T, Z]
  float* lat;                  // [M, K, B]   log p(z) - log mean_m q_m(z)     (nullable)
  float* pi;                   // [M, K, B, M] softmax_m of the expert log-densities (nullable)
  // backward
  const float* dz; const float* dlat;       // [M,K,B,T,Z], [M,K,B] (nullable)
  const float* dmu_ext[MAXM]; const float* ds_ext[MAXM];   // extra grads on mu / s (nullable)
  float kl_coef;               // d objective / d KL(q_m||p)[b,t,z]  (0 when unused; elbo: -g/B)
  float* dbott[MAXM];
};

// one thread per (r, k, b)
__global__ void latent_fwd_kernel(LatentArgs a) {
  const int TZ = a.T * a.Z;
  const long long total = (long long)a.M * a.K * a.B;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx % a.B); const int k = (int)((idx / a.B) % a.K); const int r = (int)(idx / ((long long)a.B * a.K));
    float lpz = 0.f; float lq[MAXM];
#pragma unroll
    for (int m = 0; m < MAXM; ++m) lq[m] = 0.f;
    for (int e = 0; e < TZ; ++e) {
      const int t = e / a.Z, zz = e - t * a.Z;
      const float mu = a.bott[r][((long long)b * 2 * a.T + t) * a.Z + zz];
      const float s = softplus_f(a.bott[r][((long long)b * 2 * a.T + a.T + t) * a.Z + zz]);
      const float u = a.noise[r][(((long long)k * a.B + b) * TZ) + e];
      const float z = mu + s * noise_eps(a.fam_post[r], u);
      a.z[(((long long)(r * a.K + k) * a.B + b) * TZ) + e] = z;
      if (k == 0) { a.mu[r][(long long)b * TZ + e] = mu; a.s[r][(long long)b * TZ + e] = s; }
      if (a.lat) {
        lpz += logp(a.fam_prior, z, a.pz_mu[e], a.pz_s[e]);
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
          if (m < a.M) {
            const float mum = (m == r) ? mu : a.bott[m][((long long)b * 2 * a.T + t) * a.Z + zz];
            const float sm = (m == r) ? s : softplus_f(a.bott[m][((long long)b * 2 * a.T + a.T + t) * a.Z + zz]);
            lq[m] += logp(a.fam_post[m], z, mum, sm);
          }
        }
      }
    }
    if (a.lat) {
      float mx = -INFINITY;
#pragma unroll
      for (int m = 0; m < MAXM; ++m) if (m < a.M) mx = fmaxf(mx, lq[m]);
      float se = 0.f;
#pragma unroll
      for (int m = 0; m < MAXM; ++m) if (m < a.M) se += expf(lq[m] - mx);
      const float lse = mx + logf(se);
      a.lat[idx] = lpz - (lse - logf((float)a.M));
      if (a.pi) {
#pragma unroll
        for (int m = 0; m < MAXM; ++m) if (m < a.M) a.pi[idx * a.M + m] = expf(lq[m] - lse);
      }
    }
  }
}

// one thread per (b, t, z)
__global__ void latent_bwd_kernel(LatentArgs a) {
  const int TZ = a.T * a.Z;
  const long long total = (long long)a.B * TZ;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(idx % TZ); const int b = (int)(idx / TZ);
    const int t = e / a.Z, zz = e - t * a.Z;
    float mu[MAXM], s[MAXM], raw[MAXM], dmu[MAXM], ds[MAXM];
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
      mu[m] = 0.f; s[m] = 1.f; raw[m] = 0.f; dmu[m] = 0.f; ds[m] = 0.f;
      if (m < a.M) {
        mu[m] = a.bott[m][((long long)b * 2 * a.T + t) * a.Z + zz];
        raw[m] = a.bott[m][((long long)b * 2 * a.T + a.T + t) * a.Z + zz];
        s[m] = softplus_f(raw[m]);
        if (a.dmu_ext[m]) dmu[m] = a.dmu_ext[m][(long long)b * TZ + e];
        if (a.ds_ext[m]) ds[m] = a.ds_ext[m][(long long)b * TZ + e];
      }
    }
    const float pmu = a.pz_mu ? a.pz_mu[e] : 0.f, ps = a.pz_s ? a.pz_s[e] : 1.f;
#pragma unroll
    for (int r = 0; r < MAXM; ++r) {
      if (r < a.M) {
        for (int k = 0; k < a.K; ++k) {
          const long long row = (long long)(r * a.K + k) * a.B + b;
          const float u = a.noise[r][(((long long)k * a.B + b) * TZ) + e];
          const float eps = noise_eps(a.fam_post[r], u);
          const float z = mu[r] + s[r] * eps;
          float gz = a.dz ? a.dz[row * TZ + e] : 0.f;
          if (a.dlat) {
            const float g = a.dlat[row];
            float dx, dsx;
            dlogp(a.fam_prior, z, pmu, ps, dx, dsx);
            gz += g * dx;
#pragma unroll
            for (int m = 0; m < MAXM; ++m) {
              if (m < a.M) {
                const float c = -g * a.pi[row * a.M + m];
                dlogp(a.fam_post[m], z, mu[m], s[m], dx, dsx);
                gz += c * dx; dmu[m] -= c * dx; ds[m] += c * dsx;
              }
            }
          }
          dmu[r] += gz; ds[r] += gz * eps;
        }
      }
    }
    if (a.kl_coef != 0.f) {
#pragma unroll
      for (int m = 0; m < MAXM; ++m) {
        if (m < a.M) {
          const float d = mu[m] - pmu;
          float kmu, ks;
          if (a.fam_post[m] == 0) {
            const float ex = expf(-fabsf(d) / s[m]);
            kmu = sgn(d) / ps * (1.f - ex);
            ks = -1.f / s[m] + ex / ps * (1.f + fabsf(d) / s[m]);
          } else {
            kmu = d / (ps * ps);
            ks = s[m] / (ps * ps) - 1.f / s[m];
          }
          dmu[m] += a.kl_coef * kmu; ds[m] += a.kl_coef * ks;
        }
      }
    }
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
      if (m < a.M && a.dbott[m]) {
        a.dbott[m][((long long)b * 2 * a.T + t) * a.Z + zz] = dmu[m];
        const float sp = raw[m] > 20.f ? 1.f : sigmoid_f(raw[m]);
        a.dbott[m][((long long)b * 2 * a.T + a.T + t) * a.Z + zz] = ds[m] * sp;
      }
    }
  }
}

// kld[b] = sum_{t,z} KL(q(mu,s) || p)   — one thread per b
__global__ void kl_fwd_kernel(const float* mu, const float* s, int fam, const float* pz_mu, const float* pz_s, int B, int TZ, float* kld) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int e = 0; e < TZ; ++e) {
      const float mq = mu[(long long)b * TZ + e], sq = s[(long long)b * TZ + e], mp = pz_mu[e], sp = pz_s[e];
      if (fam == 0) {
        const float r = sq / sp, d = fabsf(mq - mp);
        acc += -logf(r) + d / sp + r * expf(-d / sq) - 1.f;
      } else {
        const float vr = (sq / sp) * (sq / sp), t1 = ((mq - mp) / sp) * ((mq - mp) / sp);
        acc += 0.5f * (vr + t1 - 1.f - logf(vr));
      }
    }
    kld[b] = acc;
  }
}

// (dmu, ds)[b,e] = coef * (*gptr) * d KL(q||p)[b,e] / d(mu, s)
__global__ void kl_bwd_kernel(const float* mu, const float* s, int fam, const float* pz_mu, const float* pz_s, long long n, int TZ,
                              float coef, const float* gptr, float* dmu, float* ds) {
  if (gptr) coef *= *gptr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % TZ);
    const float sq = s[i], sp = pz_s[e], d = mu[i] - pz_mu[e];
    float kmu, ks;
    if (fam == 0) {
      const float ex = expf(-fabsf(d) / sq);
      kmu = sgn(d) / sp * (1.f - ex);
      ks = -1.f / sq + ex / sp * (1.f + fabsf(d) / sq);
    } else {
      kmu = d / (sp * sp);
      ks = sq / (sp * sp) - 1.f / sq;
    }
    dmu[i] = coef * kmu; ds[i] = coef * ks;
  }
}

// dst[i] = mult * (*gptr) * src[i]
__global__ void scale_kernel(const float* src, long long n, float mult, const float* gptr, float* dst) {
  if (gptr) mult *= *gptr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = mult * src[i];
}

// lpx[r, b] (+)= scaling * sum_l log p(x[b,l] | loc[r,b,l], scale(mask[b,l]))   — one warp per (r,b)
__global__ void __launch_bounds__(256) loglik_fwd_kernel(const float* loc, const float* x, const unsigned char* mask, int R, int B, int L,
                                                         int fam, float scale_masked, float scaling, float* lpx, int acc) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp; row < (long long)R * B; row += nwarps) {
    const int b = (int)(row % B);
    float s = 0.f;
    for (int l = lane; l < L; l += 32) {
      const float sc = (mask && mask[(long long)b * L + l]) ? scale_masked : 1.f;
      s += logp(fam, x[(long long)b * L + l], loc[row * L + l], sc);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) lpx[row] = acc ? (lpx[row] + s * scaling) : s * scaling;
  }
}

// dloc[r,b,l] = gscale * coef[r,b] * scaling * d log p / d loc
__global__ void loglik_bwd_kernel(const float* loc, const float* x, const unsigned char* mask, int R, int B, int L, int fam,
                                  float scale_masked, float scaling, const float* coef, float gscale, const float* gptr, float* dloc) {
  if (gptr) gscale *= *gptr;
  const long long total = (long long)R * B * L;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / L; const int l = (int)(i - row * L); const int b = (int)(row % B);
    const float sc = (mask && mask[(long long)b * L + l]) ? scale_masked : 1.f;
    const float d = x[(long long)b * L + l] - loc[i];
    const float dl = fam == 0 ? sgn(d) / sc : d / (sc * sc);
    const float c = coef ? coef[row] : 1.f;
    dloc[i] = gscale * c * scaling * dl;
  }
}

// IWAE combine: lw = lat + lpx ; obj = sum_b (LSE_r lw[r,b] - log R) ; w[r,b] = softmax_r. Single CTA (deterministic).
__global__ void __launch_bounds__(256) iwae_kernel(const float* lat, const float* lpx, int R, int B, float* w, float* lw_out, float* obj) {
  __shared__ float red[256];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float mx = -INFINITY;
    for (int r = 0; r < R; ++r) {
      const float v = lpx[(long long)r * B + b] + (lat ? lat[(long long)r * B + b] : 0.f);
      if (lw_out) lw_out[(long long)r * B + b] = v;
      mx = fmaxf(mx, v);
    }
    float se = 0.f;
    for (int r = 0; r < R; ++r) se += expf(lpx[(long long)r * B + b] + (lat ? lat[(long long)r * B + b] : 0.f) - mx);
    const float lse = mx + logf(se);
    for (int r = 0; r < R; ++r) w[(long long)r * B + b] = expf(lpx[(long long)r * B + b] + (lat ? lat[(long long)r * B + b] : 0.f) - lse);
    acc += lse - logf((float)R);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int sft = 128; sft >= 1; sft >>= 1) {
    if ((int)threadIdx.x < sft) red[threadIdx.x] += red[threadIdx.x + sft];
    __syncthreads();
  }
  if (threadIdx.x == 0) *obj = red[0];
}

// ELBO combine: obj = (1/(K*B)) sum_{k,b} lpx[k,b] - (1/B) sum_b kld[b].  Single CTA.
__global__ void __launch_bounds__(256) elbo_kernel(const float* lpx, const float* kld, int K, int B, float* obj) {
  __shared__ float red[256];
  float acc = 0.f;
  for (int i = threadIdx.x; i < K * B; i += blockDim.x) acc += lpx[i] / (float)(K * B);
  for (int b = threadIdx.x; b < B; b += blockDim.x) acc -= kld[b] / (float)B;
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int sft = 128; sft >= 1; sft >>= 1) {
    if ((int)threadIdx.x < sft) red[threadIdx.x] += red[threadIdx.x + sft];
    __syncthreads();
  }
  if (threadIdx.x == 0) *obj = red[0];
}

static int fill_latent(LatentArgs& a, int M, int K, int B, int T, int Z, const float* const* bott, const float* const* noise,
                       const int* fam_post, int fam_prior, const float* pz_mu, const float* pz_s) {
  V_REQUIRE(M >= 1 && M <= MAXM, V_EUNSUPPORTED, "latent: 1..%d modalities supported (M=%d)", MAXM, M);
  V_REQUIRE(K >= 1 && B >= 0 && T >= 1 && Z >= 1, V_EBADSHAPE, "latent: bad shape");
  V_REQUIRE(bott && noise && fam_post, V_ENULL, "latent: null pointer table");
  a.M = M; a.K = K; a.B = B; a.T = T; a.Z = Z; a.fam_prior = fam_prior; a.pz_mu = pz_mu; a.pz_s = pz_s;
  for (int m = 0; m < MAXM; ++m) {
    a.bott[m] = m < M ? bott[m] : nullptr; a.noise[m] = m < M ? noise[m] : nullptr; a.fam_post[m] = m < M ? fam_post[m] : 0;
    a.mu[m] = a.s[m] = nullptr; a.dmu_ext[m] = a.ds_ext[m] = nullptr; a.dbott[m] = nullptr;
    if (m < M) {
      V_REQUIRE(bott[m] && noise[m], V_ENULL, "latent: null bottleneck/noise for modality %d", m);
      V_REQUIRE(fam_post[m] == 0 || fam_post[m] == 1, V_EUNSUPPORTED, "latent: family %d", fam_post[m]);
    }
  }
  return V_OK;
}

static inline int ew_grid2(long long total, int block) {
  long long g = (total + block - 1) / block;
  if (g > 148 * 16) g = 148 * 16;
  return g < 1 ? 1 : (int)g;
}

}  // namespace vaesne

using namespace vaesne;

extern "C" int vaesne_latent_fwd(int M, int K, int B, int T, int Z, const float* const* bott, const float* const* noise,
                                 const int* fam_post, int fam_prior, const float* pz_mu, const float* pz_s,
                                 float* z, float* const* mu, float* const* s, float* lat, float* pi, void* stream) {
  LatentArgs a{};
  int rc = fill_latent(a, M, K, B, T, Z, bott, noise, fam_post, fam_prior, pz_mu, pz_s); if (rc) return rc;
  V_REQUIRE(z && mu && s, V_ENULL, "latent_fwd: null outputs");
  V_REQUIRE(lat == nullptr || (pz_mu && pz_s), V_ENULL, "latent_fwd: latent terms need prior parameters");
  for (int m = 0; m < M; ++m) { a.mu[m] = mu[m]; a.s[m] = s[m]; V_REQUIRE(mu[m] && s[m], V_ENULL, "latent_fwd: null mu/s"); }
  a.z = z; a.lat = lat; a.pi = pi;
  if (B == 0) return V_OK;
  auto k = latent_fwd_kernel;
  VLAUNCH(k, dim3(ew_grid2((long long)M * K * B, 128)), dim3(128), 0, (cudaStream_t)stream, a);
  return check_launch("latent_fwd");
}

extern "C" int vaesne_latent_bwd(int M, int K, int B, int T, int Z, const float* const* bott, const float* const* noise,
                                 const int* fam_post, int fam_prior, const float* pz_mu, const float* pz_s,
                                 const float* dz, const float* dlat, const float* pi,
                                 const float* const* dmu_ext, const float* const* ds_ext, float kl_coef,
                                 float* const* dbott, void* stream) {
  LatentArgs a{};
  int rc = fill_latent(a, M, K, B, T, Z, bott, noise, fam_post, fam_prior, pz_mu, pz_s); if (rc) return rc;
  V_REQUIRE(dbott, V_ENULL, "latent_bwd: null outputs");
  V_REQUIRE(dlat == nullptr || (pi && pz_mu && pz_s), V_ENULL, "latent_bwd: dlat needs pi and prior parameters");
  V_REQUIRE(kl_coef == 0.f || (pz_mu && pz_s), V_ENULL, "latent_bwd: KL term needs prior parameters");
  for (int m = 0; m < M; ++m) {
    a.dbott[m] = dbott[m];
    a.dmu_ext[m] = dmu_ext ? dmu_ext[m] : nullptr; a.ds_ext[m] = ds_ext ? ds_ext[m] : nullptr;
  }
  a.dz = dz; a.dlat = dlat; a.pi = const_cast<float*>(pi); a.kl_coef = kl_coef;
  if (B == 0) return V_OK;
  auto k = latent_bwd_kernel;
  VLAUNCH(k, dim3(ew_grid2((long long)B * T * Z, 128)), dim3(128), 0, (cudaStream_t)stream, a);
  return check_launch("latent_bwd");
}

extern "C" int vaesne_kl_fwd(const float* mu, const float* s, int fam, const float* pz_mu, const float* pz_s, int B, int TZ, float* kld, void* stream) {
  V_REQUIRE(mu && s && pz_mu && pz_s && kld, V_ENULL, "kl_fwd: null argument");
  if (B == 0) return V_OK;
  auto k = kl_fwd_kernel;
  VLAUNCH(k, dim3(ew_grid2(B, 128)), dim3(128), 0, (cudaStream_t)stream, mu, s, fam, pz_mu, pz_s, B, TZ, kld);
  return check_launch("kl_fwd");
}

extern "C" int vaesne_loglik_fwd(const float* loc, const float* x, const unsigned char* mask, int R, int B, int L, int fam,
                                 float scale_masked, float scaling, float* lpx, int accumulate, void* stream) {
  V_REQUIRE(loc && x && lpx, V_ENULL, "loglik_fwd: null argument");
  V_REQUIRE(fam == 0 || fam == 1, V_EUNSUPPORTED, "loglik_fwd: family %d", fam);
  if ((long long)R * B == 0) return V_OK;
  auto k = loglik_fwd_kernel;
  long long g = ((long long)R * B + 7) / 8; if (g > 148 * 8) g = 148 * 8;
  VLAUNCH(k, dim3((int)g), dim3(256), 0, (cudaStream_t)stream, loc, x, mask, R, B, L, fam, scale_masked, scaling, lpx, accumulate);
  return check_launch("loglik_fwd");
}

extern "C" int vaesne_loglik_bwd(const float* loc, const float* x, const unsigned char* mask, int R, int B, int L, int fam,
                                 float scale_masked, float scaling, const float* coef, float gscale, const float* gptr, float* dloc, void* stream) {
  V_REQUIRE(loc && x && dloc, V_ENULL, "loglik_bwd: null argument");
  if ((long long)R * B * L == 0) return V_OK;
  auto k = loglik_bwd_kernel;
  VLAUNCH(k, dim3(ew_grid2((long long)R * B * L, 256)), dim3(256), 0, (cudaStream_t)stream, loc, x, mask, R, B, L, fam, scale_masked, scaling, coef, gscale, gptr, dloc);
  return check_launch("loglik_bwd");
}

extern "C" int vaesne_iwae_combine(const float* lat, const float* lpx, int R, int B, float* w, float* lw, float* obj, void* stream) {
  V_REQUIRE(lpx && w && obj, V_ENULL, "iwae_combine: null argument");
  V_REQUIRE(R >= 1 && B >= 1, V_EBADSHAPE, "iwae_combine: bad shape");
  auto k = iwae_kernel;
  VLAUNCH(k, dim3(1), dim3(256), 0, (cudaStream_t)stream, lat, lpx, R, B, w, lw, obj);
  return check_launch("iwae_combine");
}

extern "C" int vaesne_elbo_combine(const float* lpx, const float* kld, int K, int B, float* obj, void* stream) {
  V_REQUIRE(lpx && kld && obj, V_ENULL, "elbo_combine: null argument");
  V_REQUIRE(K >= 1 && B >= 1, V_EBADSHAPE, "elbo_combine: bad shape");
  auto k = elbo_kernel;
  VLAUNCH(k, dim3(1), dim3(256), 0, (cudaStream_t)stream, lpx, kld, K, B, obj);
  return check_launch("elbo_combine");
}

extern "C" int vaesne_kl_bwd(const float* mu, const float* s, int fam, const float* pz_mu, const float* pz_s, int B, int TZ,
                             float coef, const float* gptr, float* dmu, float* ds, void* stream) {
  V_REQUIRE(mu && s && pz_mu && pz_s && dmu && ds, V_ENULL, "kl_bwd: null argument");
  if (B == 0) return V_OK;
  auto k = kl_bwd_kernel;
  VLAUNCH(k, dim3(ew_grid2((long long)B * TZ, 128)), dim3(128), 0, (cudaStream_t)stream, mu, s, fam, pz_mu, pz_s, (long long)B * TZ, TZ, coef, gptr, dmu, ds);
  return check_launch("kl_bwd");
}

extern "C" int vaesne_scale(const float* src, long long n, float mult, const float* gptr, float* dst, void* stream) {
  V_REQUIRE(src && dst, V_ENULL, "scale: null argument");
  if (n == 0) return V_OK;
  auto k = scale_kernel;
  VLAUNCH(k, dim3(ew_grid2(n, 256)), dim3(256), 0, (cudaStream_t)stream, src, n, mult, gptr, dst);
  return check_launch("scale");
}
