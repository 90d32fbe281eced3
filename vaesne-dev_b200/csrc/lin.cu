// Token-wise linear layers with fused activation / dropout / residual / LayerNorm, fwd + bwd.
//
// Restates the arithmetic of nn.Linear, nn.LayerNorm(eps=1e-5), nn.GELU (erf), nn.ReLU and
// nn.Dropout as composed in the reference's TransformerBlock
// (/root/reference/package/VAESNe/util_layers.py:285-309) and its small MLPs (:9-34).
//
// Layout: one thread owns one token (row); a CTA owns a tile of 128 tokens.  Weights are
// staged once per CTA in shared memory and read as 16-byte broadcasts.  Weight gradients are
// reduced per CTA in shared memory (each output owned by one thread, no conflicts) and
// flushed with one atomicAdd per element per CTA.
#include "common.cuh"
#include "vaesne_b200.h"
#include "lin_args.cuh"

namespace vaesne {

constexpr int TT = 128;   // tokens per tile == threads per CTA

// one thread reads / writes C consecutive floats of its own row; 16-byte accesses when allowed
template <int C>
__device__ __forceinline__ void ld_row(float* v, const float* p, bool vec) {
  if (vec) {
#pragma unroll
    for (int j = 0; j < C / 4; ++j) {
      float4 t = reinterpret_cast<const float4*>(p)[j];
      v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < C; ++j) v[j] = p[j];
  }
}
template <int C>
__device__ __forceinline__ void st_row(float* p, const float* v, bool vec) {
  if (vec) {
#pragma unroll
    for (int j = 0; j < C / 4; ++j) reinterpret_cast<float4*>(p)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < C; ++j) p[j] = v[j];
  }
}
__host__ __device__ __forceinline__ bool vec_ok(const void* p, long long ld) {
  return p == nullptr || ((((uintptr_t)p) & 15) == 0 && (ld & 3) == 0);
}


template <int NC, bool LN>
__global__ void __launch_bounds__(TT) lin_fwd_kernel(LinFwd a) {
  VDYNSMEM(float, sm);
  const int tid = threadIdx.x;
  const int K = a.K, N = a.N;
  const int Np = ((N + NC - 1) / NC) * NC;
  float* sWt = sm;                 // [K][Np]
  float* sB = sWt + K * Np;        // [Np]
  float* sG = sB + Np;             // [32]
  float* sBe = sG + 32;            // [32]
  float* sX = sBe + 32;            // [TT][K+1]
  const int ldsx = K + 1;

  for (int i = tid; i < K * Np; i += TT) {
    int k = i / Np, n = i - k * Np;
    sWt[i] = n < N ? a.W[(long long)n * K + k] : 0.f;
  }
  for (int i = tid; i < Np; i += TT) sB[i] = (i < N && a.b) ? a.b[i] : 0.f;
  if (LN && tid < 32) { sG[tid] = a.gamma[tid]; sBe[tid] = a.beta[tid]; }
  DropCfg dc = make_drop(LN ? a.p_drop : 0.f, a.seed, a.stream_id);
  const bool vec = vec_ok(a.Y, a.ldy) && vec_ok(a.H, a.ldh) && vec_ok(a.R, a.ldr) && vec_ok(a.S, 32);

  const int ntiles = (a.T + TT - 1) / TT;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    const int t0 = tile * TT;
    const int rows = min(TT, a.T - t0);
    for (int i = tid; i < rows * K; i += TT) {
      int r = i / K, k = i - r * K;
      float v = a.X[(long long)(t0 + r) * a.ldx + k];
      if (a.Xadd) v += a.Xadd[(long long)(t0 + r) * a.ldxa + k];
      sX[r * ldsx + k] = v;
    }
    __syncthreads();
    if (tid < rows) {
      const long long t = t0 + tid;
      const float* xr = sX + tid * ldsx;
      for (int nc = 0; nc < Np; nc += NC) {
        float acc[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[j] = sB[nc + j];
        for (int k = 0; k < K; ++k) {
          const float xk = xr[k];
          const float4* w4 = reinterpret_cast<const float4*>(sWt + k * Np + nc);
#pragma unroll
          for (int j = 0; j < NC / 4; ++j) {
            float4 w = w4[j];
            acc[4 * j + 0] = fmaf(xk, w.x, acc[4 * j + 0]);
            acc[4 * j + 1] = fmaf(xk, w.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(xk, w.z, acc[4 * j + 2]);
            acc[4 * j + 3] = fmaf(xk, w.w, acc[4 * j + 3]);
          }
        }
        if (LN) {
          // y = LayerNorm(R + dropout(acc)); NC == N == 32 enforced by the host wrapper
          float s[NC];
          ld_row<NC>(s, a.R + t * a.ldr, vec);
          float mean = 0.f;
          const uint32_t rh = dc.on ? drop_row_hash(dc, (uint64_t)t) : 0u;
#pragma unroll
          for (int j = 0; j < NC; ++j) {
            float v = acc[j];
            if (dc.on) v *= drop_mult_row(dc, rh, j);
            s[j] += v;
            mean += s[j];
          }
          mean *= (1.f / NC);
          float var = 0.f;
#pragma unroll
          for (int j = 0; j < NC; ++j) { float d = s[j] - mean; var = fmaf(d, d, var); }
          const float rstd = 1.f / sqrtf(var * (1.f / NC) + a.eps);
          if (a.S) st_row<NC>(a.S + t * 32, s, vec);
#pragma unroll
          for (int j = 0; j < NC; ++j) s[j] = (s[j] - mean) * rstd * sG[j] + sBe[j];
          st_row<NC>(a.Y + t * a.ldy, s, vec);
        } else {
          if (vec && nc + NC <= N) {
            if (a.H) st_row<NC>(a.H + t * a.ldh + nc, acc, true);
#pragma unroll
            for (int j = 0; j < NC; ++j) {
              if (a.act == 1) acc[j] = fmaxf(acc[j], 0.f);
              else if (a.act == 2) acc[j] = gelu_erf(acc[j]);
            }
            st_row<NC>(a.Y + t * a.ldy + nc, acc, true);
          } else {
#pragma unroll
            for (int j = 0; j < NC; ++j) {
              const int n = nc + j;
              if (n < N) {
                float v = acc[j];
                if (a.H) a.H[t * a.ldh + n] = v;
                if (a.act == 1) v = fmaxf(v, 0.f);
                else if (a.act == 2) v = gelu_erf(v);
                a.Y[t * a.ldy + n] = v;
              }
            }
          }
        }
      }
    }
  }
}


template <bool LN>
__global__ void __launch_bounds__(TT) lin_bwd_kernel(LinBwd a) {
  VDYNSMEM(float, sm);
  const int tid = threadIdx.x;
  const int K = a.K, N = a.N;
  const int Kp = ((K + 7) / 8) * 8;
  const int ldz = N + 1, ldsx = Kp + 4;
  float* sW = sm;                       // [N][Kp]
  float* sAcc = sW + N * Kp;            // [N][Kp] (only when a.smem_acc)
  float* sDb = sAcc + (a.smem_acc ? N * Kp : 0);   // [N]
  float* sG = sDb + ((N + 3) / 4) * 4;  // [32] gamma
  float* sDg = sG + 32;                 // [32]
  float* sDbe = sDg + 32;               // [32]
  float* sX = sDbe + 32;                // [TT][Kp+4]
  float* sDZ = sX + TT * ldsx;          // [TT][N+1]
  const bool wgrad = a.dW != nullptr;

  for (int i = tid; i < N * Kp; i += TT) {
    int n = i / Kp, k = i - n * Kp;
    sW[i] = k < K ? a.W[(long long)n * K + k] : 0.f;
    if (a.smem_acc) sAcc[i] = 0.f;
  }
  for (int i = tid; i < N; i += TT) sDb[i] = 0.f;
  if (tid < 32) { sG[tid] = LN ? a.gamma[tid] : 0.f; sDg[tid] = 0.f; sDbe[tid] = 0.f; }
  DropCfg dc = make_drop(LN ? a.p_drop : 0.f, a.seed, a.stream_id);
  const bool vecl = vec_ok(a.S, 32) && vec_ok(a.dY, a.lddy) && vec_ok(a.dR, a.lddr);
  const bool vecx = vec_ok(a.dX, a.lddx);

  const int ntiles = (a.T + TT - 1) / TT;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    const int t0 = tile * TT;
    const int rows = min(TT, a.T - t0);
    // ---- stage X (zero padded to Kp) ----
    if (wgrad) {
      for (int i = tid; i < rows * Kp; i += TT) {
        int r = i / Kp, k = i - r * Kp;
        float v = 0.f;
        if (k < K) {
          v = a.X[(long long)(t0 + r) * a.ldx + k];
          if (a.Xadd) v += a.Xadd[(long long)(t0 + r) * a.ldxa + k];
        }
        sX[r * ldsx + k] = v;
      }
    }
    // ---- phase 1: dZ = gradient w.r.t. the linear output ----
    if (LN) {
      float dz[32], dyv[32];
      const long long t = t0 + tid;
      const bool active = tid < rows;
      if (active) {
        float s[32];
        ld_row<32>(s, a.S + t * 32, vecl);
        ld_row<32>(dyv, a.dY + t * a.lddy, vecl);
        float mean = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) mean += s[j];
        mean *= (1.f / 32);
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { s[j] -= mean; var = fmaf(s[j], s[j], var); }
        const float rstd = 1.f / sqrtf(var * (1.f / 32) + a.eps);
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          s[j] *= rstd;                    // xhat
          dz[j] = dyv[j] * sG[j];          // g
          m1 += dz[j]; m2 = fmaf(dz[j], s[j], m2);
        }
        m1 *= (1.f / 32); m2 *= (1.f / 32);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          sDZ[tid * ldz + j] = dyv[j] * s[j];               // for dgamma
          dz[j] = rstd * (dz[j] - m1 - s[j] * m2);          // dS
        }
        if (a.dR) {
          float* p = a.dR + t * a.lddr;
          if (a.dR_acc) {
            float o[32];
            ld_row<32>(o, p, vecl);
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] += dz[j];
            st_row<32>(p, o, vecl);
          } else {
            st_row<32>(p, dz, vecl);
          }
        }
        if (dc.on) {
          const uint32_t rh = drop_row_hash(dc, (uint64_t)t);
#pragma unroll
          for (int j = 0; j < 32; ++j) dz[j] *= drop_mult_row(dc, rh, j);
        }
      }
      __syncthreads();
      if (a.dgamma) {
        const int j = tid & 31, part = tid >> 5;
        float acc = 0.f;
        for (int r = part * 32; r < min(rows, part * 32 + 32); ++r) acc += sDZ[r * ldz + j];
        atomicAdd(&sDg[j], acc);
      }
      __syncthreads();
      if (active) {
#pragma unroll
        for (int j = 0; j < 32; ++j) sDZ[tid * ldz + j] = dyv[j];
      }
      __syncthreads();
      if (a.dbeta) {
        const int j = tid & 31, part = tid >> 5;
        float acc = 0.f;
        for (int r = part * 32; r < min(rows, part * 32 + 32); ++r) acc += sDZ[r * ldz + j];
        atomicAdd(&sDbe[j], acc);
      }
      __syncthreads();
      if (active) {
#pragma unroll
        for (int j = 0; j < 32; ++j) sDZ[tid * ldz + j] = dz[j];
      }
    } else {
      for (int i = tid; i < rows * N; i += TT) {
        int r = i / N, n = i - r * N;
        float v = a.dY[(long long)(t0 + r) * a.lddy + n];
        if (a.act == 1) v = a.A[(long long)(t0 + r) * a.lda + n] > 0.f ? v : 0.f;
        else if (a.act == 2) v *= gelu_erf_grad(a.A[(long long)(t0 + r) * a.lda + n]);
        sDZ[r * ldz + n] = v;
      }
    }
    __syncthreads();
    // ---- phase 2: dX = dZ . W ----
    if (a.dX && tid < rows) {
      const long long t = t0 + tid;
      const float* zr = sDZ + tid * ldz;
      for (int kc = 0; kc < Kp; kc += 32) {
        const int cw = min(32, Kp - kc);      // multiple of 8
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
        for (int n = 0; n < N; ++n) {
          const float dz = zr[n];
          const float4* w4 = reinterpret_cast<const float4*>(sW + n * Kp + kc);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (4 * j < cw) {
              float4 w = w4[j];
              acc[4 * j + 0] = fmaf(dz, w.x, acc[4 * j + 0]);
              acc[4 * j + 1] = fmaf(dz, w.y, acc[4 * j + 1]);
              acc[4 * j + 2] = fmaf(dz, w.z, acc[4 * j + 2]);
              acc[4 * j + 3] = fmaf(dz, w.w, acc[4 * j + 3]);
            }
          }
        }
        float* p = a.dX + t * a.lddx + kc;
        if (vecx && kc + 32 <= K) {
          if (a.dX_acc) {
            float o[32];
            ld_row<32>(o, p, true);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] += o[j];
          }
          st_row<32>(p, acc, true);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (kc + j < K) p[j] = a.dX_acc ? (p[j] + acc[j]) : acc[j];
          }
        }
      }
    }
    // ---- phase 3: dW += dZ^T X, db += sum dZ ----
    if (wgrad) {
      const int total = N * Kp;
      for (int base = 0; base < total; base += TT * 8) {
        const int idx = base + tid * 8;
        if (idx < total) {
          const int n = idx / Kp, k0 = idx - n * Kp;
          float acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = 0.f;
          float dbacc = 0.f;
          for (int r = 0; r < rows; ++r) {
            const float z = sDZ[r * ldz + n];
            const float4 x0 = *reinterpret_cast<const float4*>(sX + r * ldsx + k0);
            const float4 x1 = *reinterpret_cast<const float4*>(sX + r * ldsx + k0 + 4);
            acc[0] = fmaf(z, x0.x, acc[0]); acc[1] = fmaf(z, x0.y, acc[1]);
            acc[2] = fmaf(z, x0.z, acc[2]); acc[3] = fmaf(z, x0.w, acc[3]);
            acc[4] = fmaf(z, x1.x, acc[4]); acc[5] = fmaf(z, x1.y, acc[5]);
            acc[6] = fmaf(z, x1.z, acc[6]); acc[7] = fmaf(z, x1.w, acc[7]);
            dbacc += z;
          }
          if (a.smem_acc) {
#pragma unroll
            for (int j = 0; j < 8; ++j) sAcc[idx + j] += acc[j];
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (k0 + j < K) atomicAdd(&a.dW[(long long)n * K + k0 + j], acc[j]);
          }
          if (k0 == 0) sDb[n] += dbacc;
        }
      }
    }
  }
  __syncthreads();
  if (wgrad) {
    if (a.smem_acc) {
      for (int i = tid; i < N * Kp; i += TT) {
        int n = i / Kp, k = i - n * Kp;
        if (k < K) atomicAdd(&a.dW[(long long)n * K + k], sAcc[i]);
      }
    }
    if (a.db) for (int i = tid; i < N; i += TT) atomicAdd(&a.db[i], sDb[i]);
  }
  if (LN && tid < 32) {
    if (a.dgamma) atomicAdd(&a.dgamma[tid], sDg[tid]);
    if (a.dbeta) atomicAdd(&a.dbeta[tid], sDbe[tid]);
  }
}

static int grid_for(int T) {
  int ntiles = (T + TT - 1) / TT;
  int g = ntiles < 148 * 4 ? ntiles : 148 * 4;
  return g < 1 ? 1 : g;
}

// ------------------------------------------------------------------------------------------------
// output heads: Linear(32 -> 1) over every token (get_flux.fc2 / get_photo.fc2, util_layers.py:9-18).  One thread per token
// reads its 128-byte row with eight 16-byte loads (a warp covers 4 KB of consecutive rows), so these stream at memory
// speed instead of staging rows through shared memory for a single output column.
// ------------------------------------------------------------------------------------------------
constexpr int HT = 256;
__global__ void __launch_bounds__(HT) lin_head_fwd_kernel(LinFwd a) {
  __shared__ float sW[32];
  if (threadIdx.x < 32) sW[threadIdx.x] = a.W[threadIdx.x];
  __syncthreads();
  const float b = a.b ? a.b[0] : 0.f;
  for (long long t = (long long)blockIdx.x * HT + threadIdx.x; t < a.T; t += (long long)gridDim.x * HT) {
    float x[32];
    ld_row<32>(x, a.X + t * a.ldx, true);
    float acc0 = b, acc1 = 0.f;
#pragma unroll
    for (int k = 0; k < 32; k += 2) { acc0 = fmaf(x[k], sW[k], acc0); acc1 = fmaf(x[k + 1], sW[k + 1], acc1); }
    a.Y[t * a.ldy] = acc0 + acc1;
  }
}

__global__ void __launch_bounds__(HT) lin_head_bwd_kernel(LinBwd a) {
  __shared__ float sW[32];
  __shared__ float sAcc[HT / 32][33];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 32) sW[tid] = a.W[tid];
  __syncthreads();
  const bool wgrad = a.dW != nullptr;
  float dw[32], dbs = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) dw[k] = 0.f;
  for (long long t = (long long)blockIdx.x * HT + tid; t < a.T; t += (long long)gridDim.x * HT) {
    const float dy = a.dY[t * a.lddy];
    dbs += dy;
    if (wgrad) {
      float x[32];
      ld_row<32>(x, a.X + t * a.ldx, true);
#pragma unroll
      for (int k = 0; k < 32; ++k) dw[k] = fmaf(dy, x[k], dw[k]);
    }
    if (a.dX) {
      float g[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) g[k] = 0.f;
      if (a.dX_acc) ld_row<32>(g, a.dX + t * a.lddx, true);
#pragma unroll
      for (int k = 0; k < 32; ++k) g[k] = fmaf(dy, sW[k], g[k]);
      st_row<32>(a.dX + t * a.lddx, g, true);
    }
  }
  // column k of the per-thread sums -> lane k of every warp (butterfly transpose-reduction), then over the warps
  if (wgrad) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const float send = up ? dw[i] : dw[i + off];
        const float keep = up ? dw[i + off] : dw[i];
        dw[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    sAcc[warp][lane] = dw[0];
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) dbs += __shfl_xor_sync(0xffffffffu, dbs, off);
  if (lane == 0) sAcc[warp][32] = dbs;
  __syncthreads();
  if (tid < 33 && wgrad) {
    float x = 0.f;
#pragma unroll
    for (int w = 0; w < HT / 32; ++w) x += sAcc[w][tid];
    if (tid < 32) atomicAdd(&a.dW[tid], x);
    else if (a.db) atomicAdd(&a.db[0], x);
  }
}

static bool head_fwd_eligible(const LinFwd& a) {
  return a.N == 1 && a.K == 32 && a.act == 0 && !a.R && !a.H && !a.Xadd && vec_ok(a.X, a.ldx);
}
static bool head_bwd_eligible(const LinBwd& a) {
  return a.N == 1 && a.K == 32 && a.act == 0 && !a.S && !a.Xadd && (a.X == nullptr || vec_ok(a.X, a.ldx)) && vec_ok(a.dX, a.lddx);
}

}  // namespace vaesne

using namespace vaesne;

extern "C" int vaesne_lin_fwd(const float* X, long long ldx, const float* Xadd, long long ldxa,
                              int T, int K, int N, const float* W, const float* b, int act,
                              float* H, long long ldh,
                              const float* R, long long ldr, const float* gamma, const float* beta, float eps,
                              float* S, float p_drop, const uint64_t* seed, uint32_t stream_id,
                              float* Y, long long ldy, void* stream) {
  V_REQUIRE(T >= 0 && K >= 1 && K <= 128 && N >= 1 && N <= 128, V_EBADSHAPE, "lin_fwd: need 1<=K,N<=128 (K=%d N=%d)", K, N);
  V_REQUIRE(act >= 0 && act <= 2, V_EBADSHAPE, "lin_fwd: act %d", act);
  if (T == 0) return V_OK;                  // no tokens: torch hands out null pointers for empty tensors
  V_REQUIRE(X && W && Y, V_ENULL, "lin_fwd: null X/W/Y");
  const bool ln = R != nullptr;
  if (ln) {
    V_REQUIRE(N == 32 && gamma && beta, V_EUNSUPPORTED, "lin_fwd: LayerNorm epilogue needs N==32 and gamma/beta (N=%d)", N);
    V_REQUIRE(act == 0 && H == nullptr, V_EUNSUPPORTED, "lin_fwd: LayerNorm epilogue excludes activation");
  }
  LinFwd a{X, ldx, Xadd, ldxa, T, K, N, W, b, act, H, ldh, R, ldr, gamma, beta, eps, S, p_drop, seed, stream_id, Y, ldy};
#ifndef VAESNE_EMU
  if (lin_tc_fwd_eligible(a)) return lin_tc_fwd(a, (cudaStream_t)stream);
#endif
  if (head_fwd_eligible(a)) {
    const long long blocks = ((long long)T + HT - 1) / HT;
    VLAUNCH(lin_head_fwd_kernel, dim3((unsigned)(blocks < 148 * 16 ? blocks : 148 * 16)), dim3(HT), 0, (cudaStream_t)stream, a);
    return check_launch("lin_head_fwd");
  }
  const int NC = ln ? 32 : (N <= 4 ? 4 : (N <= 8 ? 8 : 32));
  const int Np = ((N + NC - 1) / NC) * NC;
  const size_t smem = sizeof(float) * ((size_t)K * Np + Np + 64 + (size_t)TT * (K + 1));
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(grid_for(T)), block(TT);
  if (ln) { auto k = lin_fwd_kernel<32, true>; VSET_SMEM(k, smem); VLAUNCH(k, grid, block, smem, st, a); }
  else if (NC == 4) { auto k = lin_fwd_kernel<4, false>; VSET_SMEM(k, smem); VLAUNCH(k, grid, block, smem, st, a); }
  else if (NC == 8) { auto k = lin_fwd_kernel<8, false>; VSET_SMEM(k, smem); VLAUNCH(k, grid, block, smem, st, a); }
  else { auto k = lin_fwd_kernel<32, false>; VSET_SMEM(k, smem); VLAUNCH(k, grid, block, smem, st, a); }
  return check_launch("lin_fwd");
}

extern "C" int vaesne_lin_bwd(const float* dY, long long lddy, int T, int K, int N,
                              const float* S, const float* gamma, float eps, float* dgamma, float* dbeta,
                              float* dR, long long lddr, int dR_acc,
                              float p_drop, const uint64_t* seed, uint32_t stream_id,
                              int act, const float* A, long long lda,
                              const float* X, long long ldx, const float* Xadd, long long ldxa,
                              const float* W, float* dW, float* db,
                              float* dX, long long lddx, int dX_acc, void* stream) {
  V_REQUIRE(T >= 0 && K >= 1 && K <= 128 && N >= 1 && N <= 128, V_EBADSHAPE, "lin_bwd: need 1<=K,N<=128 (K=%d N=%d)", K, N);
  if (T == 0) return V_OK;
  V_REQUIRE(dY && W, V_ENULL, "lin_bwd: null dY/W");
  V_REQUIRE(act >= 0 && act <= 2 && (act == 0 || A), V_EBADSHAPE, "lin_bwd: act %d needs the saved activation", act);
  V_REQUIRE(dW == nullptr || X != nullptr, V_ENULL, "lin_bwd: dW requested without X");
  const bool ln = S != nullptr;
  if (ln) V_REQUIRE(N == 32 && gamma && act == 0, V_EUNSUPPORTED, "lin_bwd: LayerNorm path needs N==32, gamma, act none");
  LinBwd a{dY, lddy, T, K, N, S, gamma, eps, dgamma, dbeta, dR, lddr, dR_acc, p_drop, seed, stream_id,
           act, A, lda, X, ldx, Xadd, ldxa, W, dW, db, dX, lddx, dX_acc, 1};
#ifndef VAESNE_EMU
  if (lin_tc_bwd_eligible(a)) return lin_tc_bwd(a, (cudaStream_t)stream);
#endif
  if (head_bwd_eligible(a)) {
    const long long blocks = ((long long)T + HT - 1) / HT;
    VLAUNCH(lin_head_bwd_kernel, dim3((unsigned)(blocks < 148 * 8 ? blocks : 148 * 8)), dim3(HT), 0, (cudaStream_t)stream, a);
    return check_launch("lin_head_bwd");
  }
  const int Kp = ((K + 7) / 8) * 8;
  a.smem_acc = (N * Kp <= 96 * 96) ? 1 : 0;
  const size_t smem = sizeof(float) * ((size_t)(1 + a.smem_acc) * N * Kp + ((N + 3) / 4) * 4 + 96 + (size_t)TT * (Kp + 4) + (size_t)TT * (N + 1));
  V_REQUIRE(smem <= 220 * 1024, V_EUNSUPPORTED, "lin_bwd: K=%d N=%d needs %zu B of shared memory", K, N, smem);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(grid_for(T)), block(TT);
  if (ln) { auto k = lin_bwd_kernel<true>; VSET_SMEM(k, smem); VLAUNCH(k, grid, block, smem, st, a); }
  else { auto k = lin_bwd_kernel<false>; VSET_SMEM(k, smem); VLAUNCH(k, grid, block, smem, st, a); }
  return check_launch("lin_bwd");
}
