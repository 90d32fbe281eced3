// Argument block shared by the general attention kernels (attn.cu) and the tcgen05 kernels (attn_tc.cu).
#pragma once
#include <stdint.h>

namespace vaesne {

struct AttnArgs {
  const float* q; long long ldq;
  const float* k; long long ldk;
  const float* v; long long ldv;
  int N, Lq, Lk;
  const unsigned char* mask; int mask_rows; int mask_len;
  float p_drop; const uint64_t* seed; uint32_t stream_id;
  float* O; long long ldo;
  float* LSE;                 // [N,H,Lq], natural log
  // backward only
  const float* dO; long long lddo;
  float* delta;               // [N,H,Lq] workspace (written by dq pass, read by dkv pass)
  float* dq; long long lddq;
  float* dk; long long lddk;
  float* dv; long long lddv;
  int flags;                  // bit 0: this call is one KEY BLOCK of a longer attention — a (row, head) whose keys are all
                              // masked inside the block contributes nothing (O = 0, LSE = -inf, zero gradients) instead of NaN
};
constexpr int kAttnPartial = 1;

#ifndef VAESNE_EMU
// tcgen05 path (attn_tc.cu): returns V_OK after enqueueing, or a negative code. `eligible` says whether the
// shape is served by the Blackwell-native kernels (long self-attention); fwd and bwd use the same predicate.
bool attn_tc_eligible(const AttnArgs& a);
bool attn_tc_has_bwd();
int attn_tc_fwd(const AttnArgs& a, cudaStream_t st);
int attn_tc_bwd(const AttnArgs& a, cudaStream_t st);
#endif
// few-key cross attention (attn_small.cu): Lk <= 8
bool attn_small_eligible(const AttnArgs& a);
int attn_small_fwd(const AttnArgs& a, cudaStream_t st);
int attn_small_bwd(const AttnArgs& a, cudaStream_t st);
// short sequences on both sides (attn_mid.cu): 8 < Lk <= 64, Lq <= 64 — one CTA per batch row
bool attn_mid_eligible(const AttnArgs& a);
int attn_mid_fwd(const AttnArgs& a, cudaStream_t st);
int attn_mid_bwd(const AttnArgs& a, cudaStream_t st);

}  // namespace vaesne
