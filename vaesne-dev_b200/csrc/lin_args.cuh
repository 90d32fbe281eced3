// Argument blocks shared by the general linear kernels (lin.cu) and the tcgen05 kernels (lin_tc.cu).
#pragma once
#include <stdint.h>

namespace vaesne {

struct LinFwd {
  const float* X; long long ldx; const float* Xadd; long long ldxa;
  int T, K, N;
  const float* W; const float* b;
  int act;
  float* H; long long ldh;
  const float* R; long long ldr;
  const float* gamma; const float* beta; float eps;
  float* S;
  float p_drop; const uint64_t* seed; uint32_t stream_id;
  float* Y; long long ldy;
};

struct LinBwd {
  const float* dY; long long lddy;
  int T, K, N;
  // LayerNorm + dropout part (S != nullptr enables it)
  const float* S; const float* gamma; float eps;
  float* dgamma; float* dbeta;
  float* dR; long long lddr; int dR_acc;
  float p_drop; const uint64_t* seed; uint32_t stream_id;
  // activation
  int act; const float* A; long long lda;
  // linear
  const float* X; long long ldx; const float* Xadd; long long ldxa;
  const float* W;
  float* dW; float* db;
  float* dX; long long lddx; int dX_acc;
  int smem_acc;     // 1: per-CTA dW accumulators live in shared memory; 0: flush every tile with atomics
};

#ifndef VAESNE_EMU
// tcgen05 path (lin_tc.cu): K == 32, N in {32, 64, 96}, 16-byte aligned rows.
bool lin_tc_fwd_eligible(const LinFwd& a);
bool lin_tc_bwd_eligible(const LinBwd& a);
int lin_tc_fwd(const LinFwd& a, cudaStream_t st);
int lin_tc_bwd(const LinBwd& a, cudaStream_t st);
#endif

}  // namespace vaesne
