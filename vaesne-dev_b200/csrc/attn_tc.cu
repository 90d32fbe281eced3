// Blackwell-native masked self-attention for the long sequences of the spectra stacks
// (982 decoder tokens / 983 encoder-context tokens, 4 heads x head_dim 8) — forward and backward.
//
// Same arithmetic as attn.cu (nn.MultiheadAttention core, util_layers.py:289,297): q*sqrt(1/8), QK^T,
// key-padding mask, softmax, dropout(P), PV — restructured for sm_100a.  The kernels share one skeleton:
//
//   kernel        TMEM lanes (rows)   staged once per CTA (columns)           per tile
//   fwd4 (default) 128 queries x 4 wg  K (tf32 hi+lo), V (fp16 hi|lo)          S=QK^T (64 keys) -> P=2^(S-m) -> O += P V
//   bwd  (default) 128 keys   x 2 wg   Q, dO (fp16 hi|lo), lse, delta, words   S^T, T^T -> P^T, dS^T -> dV += P^T dO, dK += dS^T Q,
//                                                                              dS^T also to smem -> dQ += dS K (per tile pair)
//   fwd           128 queries x 2 wg   as fwd4, 128-key tiles                  kept for comparison (VAESNE_TC_FWD2)
//   dq + dkv      queries / keys       tf32 operands                           the two-pass backward (VAESNE_TC_BWD_SPLIT)
//
//   * one CTA per (batch row, head); the column-side operands are staged ONCE in shared memory as tcgen05 operand
//     tiles.  Masked keys are compacted away while staging — the key-padding mask is never materialised and masked
//     keys cost nothing.
//   * every product runs on the tensor core with fp32 accumulation in TMEM.  Scores need fp32-level inputs (an absolute
//     error in S is a relative error in P): kind::tf32 truncates to 10 mantissa bits, so S uses hi + lo operand pairs —
//     tf32 (3 MMAs) in the forward, fp16 [hi | lo] with K = 16 (2 MMAs) in the fused backward; P and dS are 11-bit
//     operands (fp16, exact power-of-two range management), i.e. TF32-class second products.
//   * the row-side operands (Q or K/V rows) live in TMEM: each thread stores its own row, no smem staging.
//   * each warpgroup owns a 128-row tile whose thread r holds row r (32x32b TMEM loads: max / exp / sum need no
//     shuffles); P (and dS) go back to a separate TMEM region as fp16 pairs and are the A operand of the second
//     product (tcgen05.mma .ts form), so the next tile's first product runs under this tile's exponentials.
//   * one extra warp per warpgroup issues its MMAs from ONE ELECTED lane (elect.sync: without it ptxas wraps every
//     UTCMMA in a divergence loop, 103 instead of 24 clk per issue — tests/probe/tc_rates.cu) and tracks completion
//     with tcgen05.commit -> mbarrier.
//   * dropout: element (query i, key slot c) carries the 16-bit word a_i ^ b_c (per-query / per-slot hash halves) and is
//     dropped iff that word, read as an fp16 pattern, compares >= a threshold pattern — one LOP3 + one packed-half
//     compare per TWO elements in either orientation, so the row-major forward and the key-major backward regenerate
//     identical masks.  tests/attn_tc_ref.py restates it in numpy.
//
// At head_dim 8 the kernels are bound by MUFU.EX2 and instruction issue, not by tensor math (32 tensor FLOP per
// exponential); see DESIGN.md §4.
#include "common.cuh"
#include "vaesne_b200.h"
#include "attn_args.cuh"
#include "tc_common.cuh"
#include <stdlib.h>
#include <cuda_fp16.h>

namespace vaesne {
using namespace tc;

constexpr int TCQ = 128;            // rows per tile (= TMEM lanes)
constexpr int FK = 128;             // fwd: keys per tile
constexpr int BK = 64;              // bwd: columns per tile
constexpr int MAXL = 1024;          // staged column-side length
constexpr int NTHREADS = 320;       // 8 softmax warps (2 warpgroups) + 1 MMA-issuing warp per warpgroup
constexpr float kScale = 0.35355339059327373f;             // sqrt(1/8)
constexpr float kQScale = kScale * 1.4426950408889634f;    // ... * log2(e)
constexpr float kLazy = 8.f;        // rescale O only when the row max grows by more than 2^8
constexpr int RPT = (MAXL + NTHREADS - 1) / NTHREADS;   // staged rows per thread
constexpr int TILE_F = MAXL * 8;    // floats of one staged operand array (32 KB)

__device__ long long g_tc_prof[16];   // probe build (-DVAESNE_TC_PROFILE): per-phase clocks of CTA (0,0), warp 0 of the key pass
#ifdef VAESNE_TC_PROFILE
#define TPROF(slot, expr) do { long long _t0 = clock64(); expr; prof[slot] += clock64() - _t0; } while (0)
#define TPROF_ADD(slot, v) prof[slot] += (v)
#else
#define TPROF(slot, expr) do { expr; } while (0)
#define TPROF_ADD(slot, v) (void)0
#endif
__constant__ int g_tc_dbg = 0;      // timing experiments only (tests/probe): 1 = no second-product MMAs, 2 = no exponentials
// Dropout rule of the tcgen05 kernels: element (query i, key slot c) carries the 16-bit word r = a_i ^ b_c (a_i: low half of a
// per-query hash, b_c: low half of a per-slot hash) and is DROPPED iff r, read as an fp16 bit pattern, compares >= the pattern
// `thr` (NaN patterns compare false: kept).  Counting patterns makes the rate exact to 1/65536: a positive thr T drops the
// 0x7C01 - T patterns T..+inf; thr = 0x8000 | M drops every non-negative pattern, -0 and the M negative ones of smallest
// magnitude (31746 + M).  Why this shape: the xor of two PACKED 16-bit pairs is one LOP3, the compare one HSET2 / HSETP2 for two
// elements — 1.5 (forward, mask applied to the packed fp16 P) or 2 (backward) issue slots per element instead of 3 for a 32-bit
// multiply + compare + select — and it reads the same in the row-major forward (pairs of key slots) and the key-major
// backward (pairs of queries).  tests/attn_tc_ref.py restates it in numpy.
struct TcDrop { uint32_t s0, s1, stream, thr2; float scale; bool on; };      // thr2: the threshold pattern in both halves
__device__ __forceinline__ TcDrop make_tcdrop(float p, const uint64_t* seed, uint32_t stream) {
  TcDrop d; d.on = (p > 0.f) && seed != nullptr; d.s0 = d.s1 = 0; d.stream = stream; d.thr2 = 0x7C017C01u; d.scale = 1.f;
  if (d.on) {
    uint64_t s = *seed; d.s0 = (uint32_t)s; d.s1 = (uint32_t)(s >> 32);
    int n = (int)((double)p * 65536.0 + 0.5);                 // patterns to drop
    n = n < 1 ? 1 : (n > 63490 ? 63490 : n);
    uint32_t t;
    if (n <= 31744) t = 0x7C01u - (uint32_t)n;
    else { if (n < 31746) n = 31746; t = 0x8000u | (uint32_t)(n - 31746); }
    d.thr2 = t | (t << 16);
    d.scale = (float)(1.0 / (1.0 - (double)n * (1.0 / 65536.0)));
  }
  return d;
}
__device__ __forceinline__ uint32_t drop_row_word(const TcDrop& d, int nh, int Lq, int i) {      // a_i (low 16 bits used)
  return hash_ctr(d.s0, d.s1, d.stream, (uint64_t)nh * (uint64_t)Lq + (uint64_t)i) & 0xffffu;
}
__device__ __forceinline__ uint32_t drop_col_word(const TcDrop& d, int nh, int c) {               // b_c
  return hash_ctr(d.s1, d.s0, d.stream ^ 0x5bd1e995u, ((uint64_t)nh << 32) | (uint64_t)c) & 0xffffu;
}
// 0xffff in each half of the result whose r-half is KEPT (less-than-or-unordered against the threshold pattern)
__device__ __forceinline__ uint32_t drop_keep_mask2(uint32_t r2, uint32_t thr2) {
  return __hltu2_mask(*reinterpret_cast<const __half2*>(&r2), *reinterpret_cast<const __half2*>(&thr2));
}
// fp32 multipliers (mul or 0) of the two elements packed in r2
__device__ __forceinline__ void drop_mult2(uint32_t r2, uint32_t thr2, float mul, float& m0, float& m1) {
  asm("{\n\t.reg .pred p, q;\n\tsetp.ltu.f16x2 p|q, %2, %3;\n\tselp.f32 %0, %4, 0f00000000, p;\n\tselp.f32 %1, %4, 0f00000000, q;\n\t}\n"
      : "=f"(m0), "=f"(m1) : "r"(r2), "r"(thr2), "f"(mul));
}

// packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2): two lanes per issue slot — the exponentiation loops are bound by
// instruction issue next to the 16/clk MUFU, not by FP32 throughput
typedef uint64_t f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float max3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rn_tf32(float x) { uint32_t h; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x)); return __uint_as_float(h); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void ld8g(float* d, const float* p) {
  if (((uintptr_t)p & 15) == 0) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) d[c] = p[c];
  }
}
__device__ __forceinline__ void st8g(float* p, const float* d) {
  if (((uintptr_t)p & 15) == 0) {
    reinterpret_cast<float4*>(p)[0] = make_float4(d[0], d[1], d[2], d[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(d[4], d[5], d[6], d[7]);
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) p[c] = d[c];
  }
}
// layout L1: [rows x 8] K-major operand (contraction over the 8 features): two float4 per row
__device__ __forceinline__ void put_l1(float* dst, int row, const float* x) {
  float* p = dst + (row >> 3) * 64 + (row & 7) * 4;
  *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
  *reinterpret_cast<float4*>(p + 32) = make_float4(x[4], x[5], x[6], x[7]);
}
// layout L2: for every 8 consecutive rows one [8 features x 8 rows] K-major operand (contraction over rows)
__device__ __forceinline__ void put_l2(float* dst, int row, const float* x) {
  float* p = dst + (row >> 3) * 64 + ((row & 7) >> 2) * 32 + (row & 3);
#pragma unroll
  for (int d = 0; d < 8; ++d) p[d * 4] = x[d];
}
// layout L2h (second product, kind::f16): for every 16 consecutive rows one [8 features x 16 rows] K-major fp16 operand
// (two 8x8 core matrices 128 B apart); the lo parts live 16 KB (HALF_ARR halfs) after the hi parts and form the second
// 8-row group of the N=16 operand, so accumulator columns 8..15 collect the lo-part product for free.
constexpr int HALF_ARR = (MAXL / 16) * 128;     // halfs per hi (or lo) array = 16 KB
__device__ __forceinline__ void put_l2h(__half* dst, int row, const float* x) {
  __half* p = dst + (row >> 4) * 128 + ((row >> 3) & 1) * 64 + (row & 7);
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    const __half hi = __float2half_rn(x[d]);
    p[d * 8] = hi;
    p[HALF_ARR + d * 8] = __float2half_rn(x[d] - __half2float(hi));
  }
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// power of two that brings |x| into [1, 2)  (1 for zero / subnormal / non-finite x): exact fp16 range management
__device__ __forceinline__ float pow2_normaliser(float amax) {
  const uint32_t e = (__float_as_uint(amax) >> 23) & 255u;
  return (e == 0u || e >= 254u) ? 1.f : __uint_as_float((254u - e) << 23);
}
// range management of the fp16 operands: an exact power of two (clamped to 2^+-30 so that products and quotients of three
// normalisers stay finite and normal) that brings the largest magnitude of a row set into [1, 2); non-finite maxima leave the
// data alone (the results are then non-finite as in fp32 arithmetic).  Rows whose magnitudes lie below 2^-30 of ... 1 keep a
// proportionally smaller operand — a gradient row of 1e-38 (an importance weight that underflowed) rounds to zero.
__device__ __forceinline__ float pow2_normaliser_c(float amax) {
  uint32_t e = (__float_as_uint(amax) >> 23) & 255u;
  if (e == 0u || e >= 254u) return 1.f;
  e = e < 97u ? 97u : (e > 157u ? 157u : e);
  return __uint_as_float((254u - e) << 23);
}
__device__ __forceinline__ uint32_t absmax8_bits(const float* x, uint32_t m) {
#pragma unroll
  for (int c = 0; c < 8; ++c) m = max(m, __float_as_uint(fabsf(x[c])));     // non-negative floats order like their bit patterns; NaN/inf sort last
  return m;
}
__device__ __forceinline__ void split8(const float* x, float* hi, float* lo) {
#pragma unroll
  for (int c = 0; c < 8; ++c) split_tf32(x[c], hi[c], lo[c]);
}
__device__ __forceinline__ void rn8(const float* x, float* y) {
#pragma unroll
  for (int c = 0; c < 8; ++c) y[c] = rn_tf32(x[c]);
}
__device__ __forceinline__ void tmem_put8(uint32_t taddr, const float* x) {
  uint32_t u[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) u[c] = __float_as_uint(x[c]);
  tmem_st8(taddr, u);
}

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up: `narr` operand arrays of 32 KB, then the small tables
// ------------------------------------------------------------------------------------------------
struct TcSmem {
  float* arr[6];
  float* pad;            // [64] second 8-row group of the N=16 operands: ones row (fwd) or zeros (bwd)
  float* f0; float* f1;  // [MAXL] per-column floats (dkv: lse2, delta)
  uint32_t* w0;          // [MAXL] per-column dropout words
  uint16_t* idx;         // [MAXL] compacted slot -> key index
  uint32_t* ballot; uint32_t* pre; uint64_t* bars; uint32_t* tmem;
};
__host__ __device__ constexpr size_t tc_smem_bytes(int narr, bool cols) {
  return 128 + (size_t)narr * TILE_F * 4 + 256 + (cols ? 2 * MAXL * 4 : 0) + MAXL * 4 + MAXL * 2 + 32 * 4 + 36 * 4 + 16 * 8 + 16;
}
// (the dynamic shared window is declared __align__(1024); deriving every pointer from it by plain pointer
// arithmetic keeps the shared address space visible to ptxas: LDS/STS instead of generic LD/ST)
__device__ __forceinline__ TcSmem carve(unsigned char* raw, int narr, bool cols) {
  TcSmem s;
  float* f = reinterpret_cast<float*>(raw);
  for (int i = 0; i < 6; ++i) s.arr[i] = i < narr ? f + (size_t)i * TILE_F : nullptr;
  f += (size_t)narr * TILE_F;
  s.pad = f; f += 64;
  s.f0 = s.f1 = nullptr;
  if (cols) { s.f0 = f; f += MAXL; s.f1 = f; f += MAXL; }
  s.w0 = (uint32_t*)f; f += MAXL;
  s.idx = (uint16_t*)f; f += MAXL / 2;
  s.ballot = (uint32_t*)f; s.pre = s.ballot + 32;    // pre[0..31] exclusive prefix, pre[32] total
  s.bars = (uint64_t*)(s.pre + 36);
  s.tmem = (uint32_t*)(s.bars + 16);
  return s;
}

// key compaction: ballot of kept keys per 32-key chunk + exclusive prefix.  Returns the number of kept keys.
template <int NT = NTHREADS>
__device__ __forceinline__ int compact_keys(const AttnArgs& a, const TcSmem& s, int n, int tid, int warp, int lane) {
  const unsigned char* mrow = a.mask ? a.mask + (long long)(n % a.mask_rows) * a.mask_len : nullptr;
  for (int it = warp; it < 32; it += NT / 32) {
    const int j = it * 32 + lane;
    const bool keep = j < a.Lk && !(mrow && j < a.mask_len && mrow[j]);
    const uint32_t b = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s.ballot[it] = b;
  }
  __syncthreads();
  if (warp == 0) {
    const int v = __popc(s.ballot[lane]);
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
    s.pre[lane] = incl - v;
    if (lane == 31) s.pre[32] = incl;
  }
  __syncthreads();
  return (int)s.pre[32];
}
__device__ __forceinline__ int key_slot(const TcSmem& s, int j) {     // -1 if masked
  const uint32_t bw = s.ballot[j >> 5];
  if (!((bw >> (j & 31)) & 1u)) return -1;
  return (int)s.pre[j >> 5] + __popc(bw & ((1u << (j & 31)) - 1u));
}

// Stage the unmasked keys of (n, h) as tf32 hi + lo parts.  Khi/Klo: L1 of K.  V1hi/V1lo: L1 of V.
// V2hi/V2lo, K2hi/K2lo: L2 of V / K (the lo array sits one TILE_F after the hi array: it is the second 8-row
// group of the N=16 operand, so the accumulator's columns 8..15 collect the lo-part product for free).
// Two phases: load_key_rows() issues every thread's global loads (up to RPT rows) BEFORE the key compaction, so that
// their round trip overlaps it — the prologue is a chain of memory latencies with nothing else resident on the SM;
// stage_keys() then writes the rows to their compacted slots.
template <int NT> struct KeyRowsT { static constexpr int R = (MAXL + NT - 1) / NT; float kk[R][8], vv[R][8]; };
using KeyRows = KeyRowsT<NTHREADS>;
template <int NT>
__device__ __forceinline__ void load_key_rows(const AttnArgs& a, int n, int h, int tid, KeyRowsT<NT>& kr) {
#pragma unroll
  for (int u = 0; u < KeyRowsT<NT>::R; ++u) {
    const int j = tid + u * NT;
#pragma unroll
    for (int c = 0; c < 8; ++c) { kr.kk[u][c] = 0.f; kr.vv[u][c] = 0.f; }
    if (j < a.Lk) {
      ld8g(kr.kk[u], a.k + ((long long)n * a.Lk + j) * a.ldk + h * 8);
      ld8g(kr.vv[u], a.v + ((long long)n * a.Lk + j) * a.ldv + h * 8);
    }
  }
}
template <int NT>
__device__ __forceinline__ void stage_keys(const AttnArgs& a, const TcSmem& s, int n, int h, int tid, int LkC, int tile,
                                           float* Khi, float* Klo, float* V1hi, float* V1lo, __half* V2h, __half* K2h,
                                           const TcDrop& dc, const KeyRowsT<NT>& kr, float vscale = 1.f) {
  const int Lpad = ((LkC + tile - 1) / tile) * tile;
  float z[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) z[c] = 0.f;
  for (int c = LkC + tid; c < Lpad; c += NT) {
    put_l1(Khi, c, z); put_l1(Klo, c, z);
    if (V1hi) { put_l1(V1hi, c, z); put_l1(V1lo, c, z); }
    if (V2h) put_l2h(V2h, c, z);
    if (K2h) put_l2h(K2h, c, z);
  }
  const int nh = n * kH + h;
  if (dc.on) for (int c = tid; c < Lpad; c += NT) reinterpret_cast<uint16_t*>(s.w0)[c] = (uint16_t)drop_col_word(dc, nh, c);   // b_c, packed pairs
#pragma unroll
  for (int u = 0; u < KeyRowsT<NT>::R; ++u) {
    const int j = tid + u * NT;
    const int c = j < a.Lk ? key_slot(s, j) : -1;
    if (c < 0) continue;
    float hi[8], lo[8];
    split8(kr.kk[u], hi, lo);
    put_l1(Khi, c, hi); put_l1(Klo, c, lo);
    if (K2h) put_l2h(K2h, c, kr.kk[u]);
    split8(kr.vv[u], hi, lo);
    if (V1hi) { put_l1(V1hi, c, hi); put_l1(V1lo, c, lo); }
    if (V2h) {
      float vs[8];
#pragma unroll
      for (int c2 = 0; c2 < 8; ++c2) vs[c2] = kr.vv[u][c2] * vscale;
      put_l2h(V2h, c, vs);
    }
  }
}
// largest |v| over the rows a CTA holds in registers -> s.pre[33] (cleared before compact_keys); returns after a barrier
template <int NT>
__device__ __forceinline__ float v_normaliser(const TcSmem& s, const KeyRowsT<NT>& kr, int lane) {
  uint32_t m = 0u;
#pragma unroll
  for (int u = 0; u < KeyRowsT<NT>::R; ++u) m = absmax8_bits(kr.vv[u], m);
  m = __reduce_max_sync(0xffffffffu, m);
  if (lane == 0) atomicMax(&s.pre[33], m);
  __syncthreads();
  return pow2_normaliser_c(__uint_as_float(s.pre[33]));
}

// =================================================================================================
// pipeline skeleton shared by the three kernels
// =================================================================================================
// TMEM, per warpgroup w (256 columns at w*256):
//   IN  [0,128)  : score tiles written by the first product (fwd: S 128 ; bwd: S 64 | T 64)
//   OUT [128,192): what the warpgroup writes back as fp16 pairs (fwd: P ; bwd: P^T 32 | dS 32) = A operand of the second product
//   ACC [192,224): accumulators of the second product           X [224,256): this warpgroup's row operands
// mbarriers, per warpgroup: x_ready (128 arrivals: row operands stored), s_ready (commit: first product done),
//   in_free (128: the warpgroup holds the score tile in registers), p_ready (128: OUT stored),
//   out_free (commit: second product has consumed OUT), o_ready (commit: accumulators final).
// Because OUT is separate from IN, the issuer runs the first product of tile j+1 while the warpgroup is still
// exponentiating tile j, and the second product of tile j while it works on tile j+1: the warpgroups compute back to
// back and the exponential unit (MUFU, 16/clk/SM) is the only resource that saturates.  One issuer warp per warpgroup.
constexpr int C_IN = 0, C_OUT = 128, C_ACC = 192, C_X = 224, C_WG = 256;
constexpr int B_X = 0, B_S = 1, B_F = 2, B_P = 3, B_OF = 4, B_O = 5, B_PER_WG = 6;

__device__ __forceinline__ void init_pipeline(const TcSmem& s, int tid, int warp) {
  if (tid == 0) {
    for (int w = 0; w < 2; ++w) {
      uint64_t* b = s.bars + w * B_PER_WG;
      mbar_init(&b[B_X], 128); mbar_init(&b[B_S], 1); mbar_init(&b[B_F], 128); mbar_init(&b[B_P], 128); mbar_init(&b[B_OF], 1); mbar_init(&b[B_O], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc<512>(s.tmem);
}

// Issuer of warpgroup w: row tiles w, w+2, ... ; T column tiles each.
// issue_in(j): first product of column tile j into IN.   issue_acc(j): second product of tile j from OUT into ACC.
template <int STRIDE = 2, class FIn, class FAcc>
__device__ __forceinline__ void mma_issuer(uint64_t* b, int w, int nRT, int T, FIn issue_in, FAcc issue_acc) {
  if (T <= 0) return;
  uint32_t cF = 0, cP = 0;
  int it = 0;
  for (int rt = w; rt < nRT; rt += STRIDE, ++it) {
    mbar_wait(&b[B_X], it & 1);
    fence_after();
    if (elect_one()) { issue_in(0); commit(&b[B_S]); }
    __syncwarp();
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) {
        mbar_wait(&b[B_F], cF & 1); cF++;
        fence_after();
        if (elect_one()) { issue_in(j + 1); commit(&b[B_S]); }
        __syncwarp();
      }
      mbar_wait(&b[B_P], cP & 1); cP++;
      fence_after();
      if (elect_one()) { issue_acc(j); commit(j + 1 < T ? &b[B_OF] : &b[B_O]); }
      __syncwarp();
    }
  }
}

// warpgroup-side bookkeeping of barrier phases
struct WgPhase {
  uint32_t cs, cof;
  __device__ __forceinline__ void wait_s(uint64_t* b) { mbar_wait(&b[B_S], cs & 1); cs++; fence_after(); }
  __device__ __forceinline__ void wait_out_free(uint64_t* b) { mbar_wait(&b[B_OF], cof & 1); cof++; fence_after(); }
};
__device__ __forceinline__ void signal_in_free(uint64_t* b) { fence_before(); mbar_arrive(&b[B_F]); }

// =================================================================================================
// forward, four warpgroups (default): same arithmetic and pipeline as attn_tc_fwd_kernel with 64-key tiles, so that a
// warpgroup needs 128 TMEM columns (IN 64 | OUT 32 | ACC 16 | X 16) and FOUR query tiles are in flight per CTA.
// The exponentiation loop is a dependent chain per thread (subtract, MUFU, accumulate, convert) and with two warps per
// scheduler half the issue slots went to dependency stalls (ncu: 22 % wait + 26 % long scoreboard); four warps per
// scheduler hide them.  Tile MMAs get smaller (N=64), which the tensor pipe has room for in the forward (< 25 % busy).
// =================================================================================================
constexpr int F4_THREADS = 640;          // 16 softmax warps + 4 issuer warps (one per warpgroup)
constexpr int F2_THREADS = 384;          // the two-warpgroup variant: two CTAs share an SM
constexpr int F4_FK = 64;
constexpr int F4_IN = 0, F4_OUT = 64, F4_ACC = 96, F4_X = 112, F4_CW = 128;
constexpr size_t FWD4_SMEM = (size_t)3 * TILE_F * 4 + MAXL * 4 + 32 * 4 + 36 * 4 + 24 * 8 + 16;

// NWG = 4: one CTA per SM.  NWG = 2: two CTAs per SM (101 KB of shared memory and 256 TMEM columns each) — the same sixteen
// softmax warps per SM, but the staging prologue of one (row, head) runs under the key loop of the other instead of
// leaving the SM idle.
template <int NWG>
__global__ void __launch_bounds__(NWG * 128 + 128, NWG == 4 ? 1 : 2) attn_tc_fwdN_kernel(AttnArgs a) {
  // the issuers always form a FULL warpgroup (setmaxnreg is a warpgroup-wide instruction); with NWG = 2 its last two warps
  // only help staging
  constexpr int NT = NWG * 128 + 128, CW = NWG * 4;          // threads ; first issuer warp
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  TcSmem s;
  float* const fb = reinterpret_cast<float*>(tc_smem_raw);
  float* Khi = fb; float* Klo = fb + TILE_F; __half* V2h = reinterpret_cast<__half*>(fb + 2 * TILE_F);
  {
    float* f = fb + 3 * TILE_F;
    for (int i = 0; i < 6; ++i) s.arr[i] = nullptr;
    s.pad = nullptr; s.f0 = s.f1 = nullptr; s.idx = nullptr;
    s.w0 = (uint32_t*)f; f += MAXL;
    s.ballot = (uint32_t*)f; s.pre = s.ballot + 32;
    s.bars = (uint64_t*)(s.pre + 36);
    s.tmem = (uint32_t*)(s.bars + 24);
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x, n = blockIdx.y, nh = n * kH + h;
  const TcDrop dc = make_tcdrop(a.p_drop, a.seed, a.stream_id);

  KeyRowsT<NT> kr;
  load_key_rows<NT>(a, n, h, tid, kr);
  if (tid == 0) {
    for (int w = 0; w < NWG; ++w) {
      uint64_t* b = s.bars + w * B_PER_WG;
      mbar_init(&b[B_X], 128); mbar_init(&b[B_S], 1); mbar_init(&b[B_F], 128); mbar_init(&b[B_P], 128); mbar_init(&b[B_OF], 1); mbar_init(&b[B_O], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CW) tmem_alloc<NWG * 128>(s.tmem);
  if (tid == 0) s.pre[33] = 0u;
  const int LkC = compact_keys<NT>(a, s, n, tid, warp, lane);
  // V is an fp16 [hi | lo] operand: normalised per (row, head) by an exact power of two, undone in the epilogue
  const float vnorm = v_normaliser<NT>(s, kr, lane);
  stage_keys<NT>(a, s, n, h, tid, LkC, F4_FK, Khi, Klo, nullptr, nullptr, V2h, nullptr, dc, kr, vnorm);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = *s.tmem;
  const int T = (LkC + F4_FK - 1) / F4_FK;
  const int nQT = (a.Lq + TCQ - 1) / TCQ;

  if (warp >= CW) {
    if constexpr (NWG == 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");       // the issuer warpgroup hands its registers to the softmax warpgroups
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    const int w = warp - CW;
    if (w < NWG) {
    const uint32_t idQK = idesc_tf32(128, F4_FK), idPV = idesc_f16(128, 16);
    const uint32_t aKhi = smem_u32(Khi), aKlo = smem_u32(Klo), aV2 = smem_u32(V2h);
    const uint32_t tw = tb + (uint32_t)(w * F4_CW);
    auto issue_qk = [&](int j) {
      const uint32_t d = tw + F4_IN, q = tw + F4_X;
      const uint64_t dKhi = smem_desc(aKhi + j * (F4_FK * 32), 128, 256), dKlo = smem_desc(aKlo + j * (F4_FK * 32), 128, 256);
      mma_ts(d, q, dKhi, idQK, 0);
      mma_ts(d, q + 8, dKhi, idQK, 1);
      mma_ts(d, q, dKlo, idQK, 1);
    };
    auto issue_pv = [&](int j) {
      const int nsteps = (min(F4_FK, LkC - j * F4_FK) + 15) >> 4;
      for (int t = 0; t < nsteps; ++t) {
        const uint32_t v = aV2 + (uint32_t)(j * (F4_FK / 16) + t) * 256;
        mma_ts_f16(tw + F4_ACC, tw + F4_OUT + (uint32_t)t * 8, smem_desc(v, 128, HALF_ARR * 2), idPV, (j > 0 || t > 0) ? 1u : 0u);   // [Vhi | Vlo]
      }
    };
    mma_issuer<NWG>(s.bars + w * B_PER_WG, w, nQT, T, issue_qk, issue_pv);
    }
  } else {
    if constexpr (NWG == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int wg = warp >> 2, r = tid & 127;
    uint64_t* bars = s.bars + wg * B_PER_WG;
    const uint32_t tw = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(wg * F4_CW);
    const uint32_t tIN = tw + F4_IN, tOUT = tw + F4_OUT, tO = tw + F4_ACC, tQ = tw + F4_X;
    WgPhase ph = {0, 0};
    int it = 0;
    for (int qt = wg; qt < nQT; qt += NWG, ++it) {
      const int i = qt * TCQ + r;
      const bool valid = i < a.Lq;
      if (T == 0) {       // every key masked: softmax of an empty set (the reference yields NaN)
        if (valid) {
          float* op = a.O + ((long long)n * a.Lq + i) * a.ldo + h * 8;
          for (int c = 0; c < 8; ++c) op[c] = (a.flags & kAttnPartial) ? 0.f : __int_as_float(0x7fc00000);
          a.LSE[(long long)nh * a.Lq + i] = -INFINITY;
        }
        continue;
      }
      {   // Q row -> TMEM (scaled, hi/lo split)
        float q[8], hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) q[c] = 0.f;
        if (valid) {
          ld8g(q, a.q + ((long long)n * a.Lq + i) * a.ldq + h * 8);
#pragma unroll
          for (int c = 0; c < 8; ++c) q[c] *= kQScale;
        }
        split8(q, hi, lo);
        tmem_put8(tQ, hi); tmem_put8(tQ + 8, lo);
        tmem_wait_st();
        fence_before();
        mbar_arrive(&bars[B_X]);
      }
      const uint32_t rw2 = dc.on ? drop_row_word(dc, nh, a.Lq, valid ? i : 0) * 0x00010001u : 0u;      // a_i in both halves
      float m_used = -1e30f, lsum = 0.f;
      for (int j = 0; j < T; ++j) {
        ph.wait_s(bars);
        const int nvalid = min(F4_FK, LkC - j * F4_FK);
        uint32_t sr[64];
        tmem_ld32(tIN, sr); tmem_ld32(tIN + 32, sr + 32);
        tmem_wait_ld();
        if (j + 1 < T) signal_in_free(bars);          // QK^T of the next tile runs under this tile's softmax
        float mt = -1e30f;
        if (nvalid == F4_FK) {
          // four independent chains (a single chain of 32 dependent FMNMX3 is ~150 clk of exposed latency per tile)
          float m4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
          for (int c = 0; c < 64; c += 8) {
#pragma unroll
            for (int u = 0; u < 4; ++u) m4[u] = max3(m4[u], __uint_as_float(sr[c + 2 * u]), __uint_as_float(sr[c + 2 * u + 1]));
          }
          mt = fmaxf(max3(m4[0], m4[1], m4[2]), m4[3]);
        } else {
#pragma unroll
          for (int c = 0; c < 64; ++c) { if (c >= nvalid) sr[c] = 0xff800000u; mt = fmaxf(mt, __uint_as_float(sr[c])); }
        }
        const float m_new = fmaxf(m_used, mt);
        const bool resc = __any_sync(0xffffffffu, m_new > m_used + kLazy);    // warp-uniform: TMEM ld/st are warp-collective
        float alpha = 1.f;
        if (resc) { alpha = ex2(m_used - m_new); lsum *= alpha; m_used = m_new; }   // first tile: 2^(-1e30 - m) = 0
        uint32_t pk[32];
        float l0 = 0.f, l1 = 0.f;                     // two partial sums: shorter dependency chains
        if (!dc.on) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float p0 = ex2(__uint_as_float(sr[2 * c]) - m_used), p1 = ex2(__uint_as_float(sr[2 * c + 1]) - m_used);
            l0 += p0; l1 += p1;
            pk[c] = pack_h2(p0, p1);
          }
        } else {
          const uint4* bw = reinterpret_cast<const uint4*>(s.w0 + j * (F4_FK / 2));       // 16-bit words: 8 key slots per uint4
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const uint4 bq = bw[cc];
            const uint32_t bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = cc * 8 + e * 2;
              const float p0 = ex2(__uint_as_float(sr[c]) - m_used), p1 = ex2(__uint_as_float(sr[c + 1]) - m_used);
              l0 += p0; l1 += p1;
              pk[cc * 4 + e] = pack_h2(p0, p1) & drop_keep_mask2(rw2 ^ bb[e], dc.thr2);
            }
          }
        }
        lsum += l0 + l1;
        if (j > 0) {
          ph.wait_out_free(bars);                    // PV of tile j-1 has consumed OUT and updated O
          if (resc) {
            uint32_t o[16];
            tmem_ld16(tO, o); tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            tmem_st16(tO, o);
          }
        }
        tmem_st32(tOUT, pk);
        tmem_wait_st();
        fence_before();
        mbar_arrive(&bars[B_P]);
      }
      mbar_wait(&bars[B_O], it & 1);
      fence_after();
      uint32_t o[16];
      tmem_ld16(tO, o); tmem_wait_ld();
      if (valid) {
        const float inv = dc.scale / (lsum * vnorm);
        float out[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) out[c] = (__uint_as_float(o[c]) + __uint_as_float(o[8 + c])) * inv;
        st8g(a.O + ((long long)n * a.Lq + i) * a.ldo + h * 8, out);
        a.LSE[(long long)nh * a.Lq + i] = (m_used + log2f(lsum)) * kLn2;
      }
      fence_before();
    }
  }
  fence_before();
  __syncthreads();
  if (warp == CW) { fence_after(); tmem_dealloc<NWG * 128>(tb); }
}

// =================================================================================================
// backward, fused single pass (default): the key-major pass above extended by dQ, so S, P, dP and dS are computed ONCE.
// dQ = scale * sum_keys dS[query, key] K[key] contracts over the TMEM-lane index of dS^T, which no TMEM operand can do;
// the warpgroup therefore also writes dS^T (the fp16 pairs it stores to TMEM anyway) to shared memory as an MN-major
// A operand [queries x 128 keys] (eight conflict-free 16-byte stores per thread and tile), this key tile's K rows sit next to it
// as an MN-major B operand, and every second query tile 8 SS MMAs (M=128 = two query tiles, N=8, K=16) produce the pair's dQ
// contribution in 8 TMEM columns (row m -> lane m).  One CTA owns all of dq[n, :, h], so the contributions of its key tiles are summed
// with vector reductions into the zero-initialised output — no second pass, no extra exponentials.
// TMEM is full (IN 128 | OUT 64 | ACC 32 | X 32 per warpgroup), so the dV product takes dO as fp16 hi only and dQ takes K as
// fp16 hi only (their accumulators are 8 columns instead of 16).  every fp16 operand is range-managed by exact powers of two:
// dO, q and v per (row, head), k per key tile (undone by the exponent's fma and the output scales).
// =================================================================================================
constexpr int C_DQ = 216;                    // per warpgroup: ACC = dK hi|lo (192..207) | dV (208..215) | dQ tile (216..223)
constexpr int DS_BYTES = 128 * 128 * 2, KB_BYTES = 8 * 128 * 2;      // dS^T of a tile pair (128 queries x 128 keys fp16), K rows
constexpr size_t BWD_SMEM = (size_t)2 * TILE_F * 4 + 2 * DS_BYTES + 2 * KB_BYTES + 3 * MAXL * 4 + MAXL * 2 + 32 * 4 + 36 * 4 + 16 * 8 + 16 + 32 * 4;
static_assert(BWD_SMEM <= 232448, "fused backward does not fit the 227 KB shared-memory window");

__device__ __forceinline__ void red_add8(float* p, const float* v) {
  if (((uintptr_t)p & 15) == 0) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) atomicAdd(p + c, v[c]);
  }
}

// The first products (S^T = K Q^T, T^T = V dO^T) run as kind::f16 on fp16 hi/lo operands as well — K/V rows as
// [hi | lo] half pairs in TMEM, and the SAME fp16 arrays the second products use ([d][query], 8x8 core matrices) read as
// an MN-major B operand whose two 8-wide K chunks are the same memory (LBO = 0): [hi|lo] x [Qhi|Qhi] + [hi|0] x [Qlo|Qlo]
// = hi*hi + lo*hi + hi*lo in TWO MMAs per product instead of three (3xTF32 needed three and four more 32 KB arrays).
// The shared memory this frees holds dS^T of TWO query tiles, so the dQ product runs once per tile pair with M = 128.
__global__ void __launch_bounds__(NTHREADS, 1) attn_tc_bwd_kernel(AttnArgs a) {
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  TcSmem s;
  float* const f0 = reinterpret_cast<float*>(tc_smem_raw);
  __half* Q2h = reinterpret_cast<__half*>(f0);                       // Q  hi | lo  ([d][query] fp16, 16 KB each)
  __half* G2h = reinterpret_cast<__half*>(f0 + TILE_F);              // dO hi | lo
  unsigned char* const dsb = tc_smem_raw + (size_t)2 * TILE_F * 4;   // per warpgroup: dS^T of a tile pair (A of the dQ product)
  unsigned char* const kbb = dsb + 2 * DS_BYTES;                                      // per warpgroup: K rows (B of the dQ product)
  {
    float* f = reinterpret_cast<float*>(kbb + 2 * KB_BYTES);
    for (int i = 0; i < 6; ++i) s.arr[i] = nullptr;
    s.pad = nullptr;
    s.f0 = f; f += MAXL; s.f1 = f; f += MAXL;
    s.w0 = (uint32_t*)f; f += MAXL;
    s.idx = (uint16_t*)f; f += MAXL / 2;
    s.ballot = (uint32_t*)f; s.pre = s.ballot + 32;
    s.bars = (uint64_t*)(s.pre + 36);
    s.tmem = (uint32_t*)(s.bars + 16);
  }
  uint32_t* const nrm = s.tmem + 4;        // [2 warpgroups][8 key-tile iterations]: largest |k| of a key tile
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x, n = blockIdx.y, nh = n * kH + h;
  const TcDrop dc = make_tcdrop(a.p_drop, a.seed, a.stream_id);
#ifdef VAESNE_TC_PROFILE
  const long long tk0 = clock64();
#endif

  // query side: every thread owns up to RPT queries; their loads are issued first so that the round trip overlaps the key
  // compaction below (the prologue is a chain of global-memory latencies and nothing else runs on the SM meanwhile)
  // (first the largest |v| of the (row, head): V enters dP = dO V^T, whose scale the per-query delta must share, so its
  // normaliser has to be known before the query-side tables are written; K's normaliser is per key tile, see below)
  uint32_t vmb = 0u;
#pragma unroll
  for (int u = 0; u < RPT; ++u) {
    const int j = tid + u * NTHREADS;
    if (j < a.Lk) {
      float vv[8];
      ld8g(vv, a.v + ((long long)n * a.Lk + j) * a.ldv + h * 8);
      vmb = absmax8_bits(vv, vmb);
    }
  }
  float q[RPT][8], g[RPT][8], o[RPT][8], lse2[RPT];
#pragma unroll
  for (int u = 0; u < RPT; ++u) {
    const int i = tid + u * NTHREADS;
#pragma unroll
    for (int c = 0; c < 8; ++c) { q[u][c] = 0.f; g[u][c] = 0.f; o[u][c] = 0.f; }
    lse2[u] = INFINITY;
    if (i < a.Lq) {
      ld8g(q[u], a.q + ((long long)n * a.Lq + i) * a.ldq + h * 8);
      ld8g(g[u], a.dO + ((long long)n * a.Lq + i) * a.lddo + h * 8);
      ld8g(o[u], a.O + ((long long)n * a.Lq + i) * a.ldo + h * 8);
      lse2[u] = a.LSE[(long long)nh * a.Lq + i] * kLog2e;
    }
  }
  init_pipeline(s, tid, warp);
  if (tid == 0) { mbar_init(&s.bars[12], 1); mbar_init(&s.bars[13], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); s.pre[33] = 0u; s.pre[34] = 0u; s.pre[35] = 0u; }
  if (tid < 32) nrm[tid] = 0u;
  const int LkC = compact_keys(a, s, n, tid, warp, lane);
  const int nKT = (LkC + TCQ - 1) / TCQ;
  // Key tiles go to the two warpgroups in pairs; an odd last tile is SHARED: each warpgroup takes half of its query tiles
  // and the two partial dK/dV are summed by reductions into zero-initialised rows (an odd count would otherwise leave one
  // warpgroup idle for a whole tile: 4 vs 3 at the usual ~800-900 unmasked keys).
  const int nPair = nKT >> 1, shared_from = (nKT & 1) ? (nKT - 1) * TCQ : (1 << 30);
  for (int j = tid; j < a.Lk; j += NTHREADS) {          // slot -> key index; masked keys (and the shared tile's rows) start at zero
    const int c = key_slot(s, j);
    if (c >= 0) { s.idx[c] = (uint16_t)j; if (c < shared_from) continue; }
    float z[8];
#pragma unroll
    for (int c2 = 0; c2 < 8; ++c2) z[c2] = 0.f;
    st8g(a.dk + ((long long)n * a.Lk + j) * a.lddk + h * 8, z);
    st8g(a.dv + ((long long)n * a.Lk + j) * a.lddv + h * 8, z);
  }
  // query side (as in the key-major pass; delta = rowsum(dO * O) is computed here, dq starts at zero — NaN if every key is masked)
  const int NQ = (a.Lq + BK - 1) / BK;
  {
    float delta[RPT];
    float gm = 0.f;
    uint32_t qmb = 0u;
    const float init = (LkC > 0 || (a.flags & kAttnPartial)) ? 0.f : __int_as_float(0x7fc00000);
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      const int i = tid + u * NTHREADS;
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) { d = fmaf(g[u][c], o[u][c], d); o[u][c] = init; gm = fmaxf(gm, fabsf(g[u][c])); }
      delta[u] = d;
      qmb = absmax8_bits(q[u], qmb);
      if (i < a.Lq) st8g(a.dq + ((long long)n * a.Lq + i) * a.lddq + h * 8, o[u]);
    }
    gm = isfinite(gm) ? gm : 0.f;
    const uint32_t gmb = __reduce_max_sync(0xffffffffu, __float_as_uint(gm));
    qmb = __reduce_max_sync(0xffffffffu, qmb);
    vmb = __reduce_max_sync(0xffffffffu, vmb);
    if (lane == 0) { atomicMax(&s.pre[33], gmb); atomicMax(&s.pre[34], qmb); atomicMax(&s.pre[35], vmb); }       // cleared before the barriers of compact_keys
    __syncthreads();
    const float sc = pow2_normaliser_c(__uint_as_float(s.pre[33]));
    const float dsc = sc * pow2_normaliser_c(__uint_as_float(s.pre[35]));       // delta shares the scale of dP = (dO sc) (V v_norm)^T
    const float qs = kQScale * pow2_normaliser_c(__uint_as_float(s.pre[34]));
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      const int i = tid + u * NTHREADS;
      if (i >= NQ * BK) continue;
#pragma unroll
      for (int c = 0; c < 8; ++c) { q[u][c] *= qs; g[u][c] *= sc; }
      put_l2h(Q2h, i, q[u]); put_l2h(G2h, i, g[u]);
      s.f0[i] = -lse2[u]; s.f1[i] = -delta[u] * dsc;
      reinterpret_cast<uint16_t*>(s.w0)[i] = dc.on ? (uint16_t)drop_row_word(dc, nh, a.Lq, i < a.Lq ? i : 0) : (uint16_t)0;   // a_i, packed pairs
    }
  }
  const float cs_scale = pow2_normaliser_c(__uint_as_float(s.pre[33]));
  const float q_norm = pow2_normaliser_c(__uint_as_float(s.pre[34]));       // Q2h holds q * sqrt(1/8) * log2(e) * q_norm
  const float v_norm = pow2_normaliser_c(__uint_as_float(s.pre[35]));       // V rows enter TMEM as v * v_norm: dS carries cs_scale * v_norm
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = *s.tmem;
  const int nIter = nPair + (nKT & 1);
  // per warpgroup: IN = S^T (64) | T^T (64) ; OUT = P^T (32) | dS^T (32) fp16 pairs ; ACC = dK hi|lo (16) | dV (8) | dQ tile (8) ; X = Khi | Klo | Vhi | Vlo

  if (warp >= 8) {
    const int w = warp - 8;
    uint64_t* b = s.bars + w * B_PER_WG;
    uint64_t* bdq = s.bars + 12 + w;
    const uint32_t idK = idesc_f16(128, 16), idV = idesc_f16(128, 8), idQ = idesc_f16_mn(128, 8, true, true);
    const uint32_t aQ2 = smem_u32(Q2h), aG2 = smem_u32(G2h);
    const uint32_t aDS = smem_u32(dsb + w * DS_BYTES), aKB = smem_u32(kbb + w * KB_BYTES);
    const uint32_t tw = tb + (uint32_t)(w * C_WG);
    const uint32_t idF = idesc_f16_mn(128, BK, false, true);
    auto issue_st = [&](int j) {
      const uint32_t d = tw + C_IN, x = tw + C_X;
      // B = [N = 64 queries][K = 8 features, read twice], MN-major view of the [d][query] fp16 arrays
      const uint32_t off = (uint32_t)j * (BK / 8) * 128;
      mma_ts_f16(d, x, smem_desc(aQ2 + off, 0, 128), idF, 0);                          // [Khi | Klo] x [Qhi | Qhi]
      mma_ts_f16(d, x + 8, smem_desc(aQ2 + HALF_ARR * 2 + off, 0, 128), idF, 1);       // [Khi | 0  ] x [Qlo | Qlo]
      mma_ts_f16(d + 64, x + 16, smem_desc(aG2 + off, 0, 128), idF, 0);                // [Vhi | Vlo] x [Ghi | Ghi]
      mma_ts_f16(d + 64, x + 24, smem_desc(aG2 + HALF_ARR * 2 + off, 0, 128), idF, 1); // [Vhi | 0  ] x [Glo | Glo]
    };
    uint32_t cF = 0, cP = 0;
    int it = 0;
    for (; it < nIter; ++it) {
      const bool shared = it == nPair;
      const int kt = shared ? nKT - 1 : 2 * it + w;
      const int jb = (shared && w) ? NQ / 2 : 0, je = (shared && !w) ? NQ / 2 : NQ;
      const int ksteps = (min(TCQ, LkC - kt * TCQ) + 15) >> 4;
      mbar_wait(&b[B_X], it & 1);
      fence_after();
      if (elect_one()) { issue_st(jb); commit(&b[B_S]); }
      __syncwarp();
      for (int j = jb; j < je; ++j) {
        const bool last = j + 1 == je;
        if (!last) {
          mbar_wait(&b[B_F], cF & 1); cF++;
          fence_after();
          if (elect_one()) { issue_st(j + 1); commit(&b[B_S]); }
          __syncwarp();
        }
        mbar_wait(&b[B_P], cP & 1); cP++;
        fence_after();
        if (elect_one()) {
          const int nsteps = (min(BK, a.Lq - j * BK) + 15) >> 4;
          const uint32_t dK = tw + C_ACC, dV = dK + 16;
          for (int t = 0; t < nsteps; ++t) {
            const uint32_t off = (uint32_t)(j * (BK / 16) + t) * 256;
            const uint32_t acc = (j > jb || t > 0) ? 1u : 0u;
            mma_ts_f16(dV, tw + C_OUT + (uint32_t)t * 8, smem_desc(aG2 + off, 128, 256), idV, acc);                   // P^T dO
            mma_ts_f16(dK, tw + C_OUT + 32 + (uint32_t)t * 8, smem_desc(aQ2 + off, 128, HALF_ARR * 2), idK, acc);     // dS^T [Qhi | Qlo]
          }
          if (!last) commit(&b[B_OF]);
          if (((j - jb) & 1) || last) {                    // dS K -> dQ of the tile pair (M = 128: two 64-query tiles)
            for (int t = 0; t < ksteps; ++t)
              mma_ss_f16(tw + C_DQ, smem_desc(aDS + t * 256, 128, 2048), smem_desc(aKB + t * 256, 128, 2048), idQ, t > 0 ? 1u : 0u);
            commit(last ? &b[B_O] : bdq);
          }
        }
        __syncwarp();
      }
    }
  } else {
    const int wg = warp >> 2, r = tid & 127;
    uint64_t* bars = s.bars + wg * B_PER_WG;
    uint64_t* bdq = s.bars + 12 + wg;
    const uint32_t tw = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(wg * C_WG);
    const uint32_t tIN = tw + C_IN, tOUT = tw + C_OUT, tA = tw + C_ACC, tX = tw + C_X;
    unsigned char* const ds_row = dsb + wg * DS_BYTES + (r & 7) * 16 + (r >> 3) * 128;      // this key's 16-byte slot in each query group
    unsigned char* const kb_row = kbb + wg * KB_BYTES + (r & 7) * 16 + (r >> 3) * 128;
    const float ds_inv = 1.f / (cs_scale * v_norm);     // undoes the scale dS^T (and what is contracted with it) carries
    float dq_scale = kScale * ds_inv;            // times 1 / (this key tile's K normaliser), set per tile
    WgPhase ph = {0, 0};
    uint32_t cdq = 0;
    int it = 0;
#ifdef VAESNE_TC_PROFILE
    long long prof[16] = {0}; const long long tstart = clock64(); prof[6] = tstart - tk0;
#endif
    // dQ contribution of the tile pair starting at query tile jq: accumulator row m (query jq*64 + m) sits in lane m
    auto drain_dq = [&](int jq, int nrows) {
      uint32_t v[8];
      tmem_ld8(tw + C_DQ, v); tmem_wait_ld();
      const int i = jq * BK + r;
      if (r < nrows && i < a.Lq) {
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = __uint_as_float(v[c]) * dq_scale;
        red_add8(a.dq + ((long long)n * a.Lq + i) * a.lddq + h * 8, o);
      }
    };
    for (; it < nIter; ++it) {
      const bool shared = it == nPair;
      const int kt = shared ? nKT - 1 : 2 * it + wg;
      const int jb = (shared && wg) ? NQ / 2 : 0, je = (shared && !wg) ? NQ / 2 : NQ;
      const int cs = kt * TCQ + r;
      const bool valid = cs < LkC;
#ifdef VAESNE_TC_PROFILE
      const long long tset0 = clock64();
#endif
      const int jk = valid ? (int)s.idx[cs] : 0;
      float k[8], v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) { k[c] = 0.f; v[c] = 0.f; }
      if (valid) {
        ld8g(k, a.k + ((long long)n * a.Lk + jk) * a.ldk + h * 8);
        ld8g(v, a.v + ((long long)n * a.Lk + jk) * a.ldv + h * 8);
      }
      // K and V rows are fp16 operands ([hi | lo] in TMEM, hi in the dQ product), range-managed by exact powers of two: V per
      // (row, head) (above), K per key tile — the exponent (an fma instead of an add) and the dQ drain undo it for free
      float k_norm;
      {
        const uint32_t kmb = __reduce_max_sync(0xffffffffu, absmax8_bits(k, 0u));
        uint32_t* slot = nrm + wg * 8 + it;
        if (lane == 0) atomicMax(slot, kmb);
        asm volatile("bar.sync %0, 128;" :: "r"(wg + 1) : "memory");
        k_norm = pow2_normaliser_c(__uint_as_float(slot[0]));
#pragma unroll
        for (int c = 0; c < 8; ++c) { k[c] *= k_norm; v[c] *= v_norm; }
      }
      const float inv_qk = 1.f / (q_norm * k_norm);
      const f32x2 iqk2 = pk2(inv_qk, inv_qk);
      dq_scale = kScale * ds_inv / k_norm;
      *reinterpret_cast<uint4*>(kb_row) = make_uint4(pack_h2(k[0], k[1]), pack_h2(k[2], k[3]), pack_h2(k[4], k[5]), pack_h2(k[6], k[7]));
      {                // A operands: columns 0-3 = hi pairs, 4-7 = lo pairs (K index 0-7 hi, 8-15 lo); second operand: [hi | 0]
        uint32_t xa[8], xb[8];
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const float* src = pass ? v : k;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const __half2 h2 = __floats2half2_rn(src[2 * c], src[2 * c + 1]);
            const float2 hf = __half22float2(h2);
            xa[c] = *reinterpret_cast<const uint32_t*>(&h2);
            xa[4 + c] = pack_h2(src[2 * c] - hf.x, src[2 * c + 1] - hf.y);
            xb[c] = xa[c]; xb[4 + c] = 0u;
          }
          tmem_st8(tX + pass * 16, xa); tmem_st8(tX + pass * 16 + 8, xb);
        }
      }
      fence_async_smem();
      tmem_wait_st();
      fence_before();
      mbar_arrive(&bars[B_X]);
      TPROF_ADD(9, clock64() - tset0);
      const uint32_t cw2 = dc.on ? drop_col_word(dc, nh, cs) * 0x00010001u : 0u;      // b_c in both halves
      for (int j = jb; j < je; ++j) {
        TPROF(0, ph.wait_s(bars));
#ifdef VAESNE_TC_PROFILE
        const long long tc0 = clock64();
#endif
        uint32_t pk[32], dk2[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t sr[32], tr[32];
          tmem_ld32(tIN + half * 32, sr); tmem_ld32(tIN + 64 + half * 32, tr);
          tmem_wait_ld();
          if (half == 1 && j + 1 < je) signal_in_free(bars);
          const float4* l4 = reinterpret_cast<const float4*>(s.f0 + j * BK + half * 32);
          const float4* d4 = reinterpret_cast<const float4*>(s.f1 + j * BK + half * 32);
          const uint2* w4 = reinterpret_cast<const uint2*>(s.w0 + (j * BK + half * 32) / 2);      // 16-bit words: 4 queries per uint2
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const float4 lv = l4[cc], dv = d4[cc];
            float pp[4], ss[4];
            const int c = cc * 4;
            float x0, x1, x2, x3;
            upk2(fma2(pk2(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), iqk2, pk2(lv.x, lv.y)), x0, x1);
            upk2(fma2(pk2(__uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3])), iqk2, pk2(lv.z, lv.w)), x2, x3);
            const float p0 = ex2(x0), p1 = ex2(x1), p2 = ex2(x2), p3 = ex2(x3);
            const f32x2 pa = pk2(p0, p1), pb = pk2(p2, p3);
            const f32x2 ta = pk2(__uint_as_float(tr[c]), __uint_as_float(tr[c + 1])), tb2 = pk2(__uint_as_float(tr[c + 2]), __uint_as_float(tr[c + 3]));
            if (!dc.on) {
              pp[0] = p0; pp[1] = p1; pp[2] = p2; pp[3] = p3;
              upk2(mul2(pa, add2(ta, pk2(dv.x, dv.y))), ss[0], ss[1]);
              upk2(mul2(pb, add2(tb2, pk2(dv.z, dv.w))), ss[2], ss[3]);
            } else {
              const uint2 wv = w4[cc];
              float m0, m1, m2, m3;
              drop_mult2(wv.x ^ cw2, dc.thr2, dc.scale, m0, m1);
              drop_mult2(wv.y ^ cw2, dc.thr2, dc.scale, m2, m3);
              const f32x2 ma = pk2(m0, m1), mb = pk2(m2, m3);
              upk2(mul2(pa, ma), pp[0], pp[1]);
              upk2(mul2(pb, mb), pp[2], pp[3]);
              upk2(mul2(pa, fma2(ta, ma, pk2(dv.x, dv.y))), ss[0], ss[1]);
              upk2(mul2(pb, fma2(tb2, mb, pk2(dv.z, dv.w))), ss[2], ss[3]);
            }
            pk[half * 16 + cc * 2] = pack_h2(pp[0], pp[1]); pk[half * 16 + cc * 2 + 1] = pack_h2(pp[2], pp[3]);
            dk2[half * 16 + cc * 2] = pack_h2(ss[0], ss[1]); dk2[half * 16 + cc * 2 + 1] = pack_h2(ss[2], ss[3]);
          }
        }
        TPROF_ADD(1, clock64() - tc0);
        const int par = (j - jb) & 1;
        if (j > jb) TPROF(5, ph.wait_out_free(bars));
        if (par == 0 && j > jb) {
          TPROF(7, mbar_wait(bdq, cdq & 1)); cdq++;       // dQ product of the previous tile pair done: read its accumulator, reuse the dS^T buffer
          fence_after();
          TPROF(8, drain_dq(j - 2, 2 * BK));
        }
#ifdef VAESNE_TC_PROFILE
        const long long ts0 = clock64();
#endif
        tmem_st32(tOUT, pk); tmem_st32(tOUT + 32, dk2);
        // padded key rows (zero K, but P = 2^(-lse) may be huge) must contribute exact zeros to dQ
#pragma unroll
        for (int g8 = 0; g8 < 8; ++g8)
          *reinterpret_cast<uint4*>(ds_row + (par * 8 + g8) * 2048) = valid ? make_uint4(dk2[g8 * 4], dk2[g8 * 4 + 1], dk2[g8 * 4 + 2], dk2[g8 * 4 + 3]) : make_uint4(0u, 0u, 0u, 0u);
        fence_async_smem();
        tmem_wait_st();
        fence_before();
        mbar_arrive(&bars[B_P]);
        TPROF_ADD(2, clock64() - ts0);
      }
      TPROF(3, mbar_wait(&bars[B_O], it & 1));
#ifdef VAESNE_TC_PROFILE
      const long long ttail0 = clock64();
#endif
      fence_after();
      if ((je - jb) & 1) drain_dq(je - 1, BK); else drain_dq(je - 2, 2 * BK);     // an odd last tile sits alone in rows 0..63
      uint32_t o[24];
      tmem_ld16(tA, o); tmem_ld8(tA + 16, o + 16); tmem_wait_ld();
      if (valid) {
        float dk[8], dv[8];
        const float ck = kLn2 * ds_inv / q_norm;               // hi + lo parts; Q carried log2(e) and its normaliser
        const float cv = 1.f / cs_scale;                       // dV = P^T (dO cs_scale)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          dk[c] = (__uint_as_float(o[c]) + __uint_as_float(o[8 + c])) * ck;
          dv[c] = __uint_as_float(o[16 + c]) * cv;
        }
        float* pk_ = a.dk + ((long long)n * a.Lk + jk) * a.lddk + h * 8;
        float* pv_ = a.dv + ((long long)n * a.Lk + jk) * a.lddv + h * 8;
        if (shared) { red_add8(pk_, dk); red_add8(pv_, dv); }
        else { st8g(pk_, dk); st8g(pv_, dv); }
      }
      fence_before();
      TPROF_ADD(10, clock64() - ttail0);
    }
#ifdef VAESNE_TC_PROFILE
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) { prof[4] = clock64() - tstart; for (int q = 0; q < 11; ++q) g_tc_prof[q] = prof[q]; }
#endif
  }
  fence_before();
  __syncthreads();
  if (warp == 8) { fence_after(); tmem_dealloc<512>(tb); }
}

// =================================================================================================
// backward, ONE warpgroup per CTA and TWO CTAs per SM (default).  Same arithmetic, tiles, TMEM map and barrier protocol as the
// two-warpgroup kernel above, but a CTA = 4 compute warps + 1 issuer warp that walks ALL key tiles of its (row, head):
//   * two independent CTAs share an SM (113 KB of shared memory and 256 TMEM columns each), so the staging prologue and the
//     per-key-tile pipeline drain / refill of one run under the exponentiation loop of the other — with two warpgroups in ONE
//     CTA both sit in those phases together (measured in-kernel: 13 % + 12 % of the CTA's time with the SM idle);
//   * no key tile is shared between warpgroups: dK / dV are plain stores, only dQ is accumulated with reductions;
//   * 160 threads leave the register file for two such CTAs at 200 registers per thread.
// The prologue runs in two passes over the query rows (the normalisers must be known before the operands are written): pass 1
// only takes maxima, pass 2 re-reads the rows (L2 hits) and stages them.
// =================================================================================================
constexpr int B1_THREADS = 160;
constexpr int B1_RPT = (MAXL + B1_THREADS - 1) / B1_THREADS;      // 7 rows per thread and pass
// TMEM columns of this kernel: IN 0..127 (S^T | T^T), OUT 128..191 (P^T | dS^T fp16 pairs), dK 192..207, dV 208..223, dQ 224..239
// (each hi | lo: every second product takes its shared-memory operand as fp16 [hi | lo]), X 240..255 (K and V rows [hi | lo]).
// X needs no "[hi | 0]" copies: the low-order first-product MMA multiplies [Khi | Klo] by [Qlo | Qlo], whose extra Klo*Qlo
// term is the genuine fourth term of the exact product.
constexpr int E_DK = 192, E_DV = 208, E_DQ = 224, E_X = 240;
constexpr size_t BWD1_SMEM = (size_t)2 * TILE_F * 4 + DS_BYTES + 2 * KB_BYTES + 2 * MAXL * 4 + MAXL * 2 + MAXL * 2 + 32 * 4 + 36 * 4 + 8 * 8 + 16 + 8 * 4;
static_assert(2 * (BWD1_SMEM + 1024) <= 233472, "two one-warpgroup backward CTAs must fit one SM");

__global__ void __launch_bounds__(B1_THREADS, 2) attn_tc_bwd1_kernel(AttnArgs a) {
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  TcSmem s;
  float* const fb = reinterpret_cast<float*>(tc_smem_raw);
  __half* Q2h = reinterpret_cast<__half*>(fb);                       // Q  hi | lo  ([d][query] fp16, 16 KB each)
  __half* G2h = reinterpret_cast<__half*>(fb + TILE_F);              // dO hi | lo
  unsigned char* const dsb = tc_smem_raw + (size_t)2 * TILE_F * 4;   // dS^T of a tile pair (A of the dQ product)
  unsigned char* const kbb = dsb + DS_BYTES;                         // K rows, fp16 hi then lo (B of the dQ product)
  uint16_t* w16;                                                     // [MAXL] per-query dropout halves a_i
  {
    float* f = reinterpret_cast<float*>(kbb + 2 * KB_BYTES);
    for (int i = 0; i < 6; ++i) s.arr[i] = nullptr;
    s.pad = nullptr; s.w0 = nullptr;
    s.f0 = f; f += MAXL; s.f1 = f; f += MAXL;
    w16 = (uint16_t*)f; f += MAXL / 2;
    s.idx = (uint16_t*)f; f += MAXL / 2;
    s.ballot = (uint32_t*)f; s.pre = s.ballot + 32;
    s.bars = (uint64_t*)(s.pre + 36);
    s.tmem = (uint32_t*)(s.bars + 8);
  }
  uint32_t* const nrm = s.tmem + 4;        // [8 key tiles]: largest |k| of a key tile
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x, n = blockIdx.y, nh = n * kH + h;
  const TcDrop dc = make_tcdrop(a.p_drop, a.seed, a.stream_id);

  // ---- staging, ONE global round trip: the q / dO / O rows of the (row, head) travel global -> shared with cp.async into
  // regions that are free during the prologue (raw q in the dS^T buffer, raw dO where Q2h will be, raw O where G2h will be);
  // V rows (only their maximum is needed) and lse go through registers, all requested before anything is consumed.  The
  // operands are then converted shared -> shared in an order that never overwrites unread rows.  (A register-staged version
  // walked 7 rows per thread through 7 dependent round trips per pass: 0.54 of the kernel's 2.73 ms.)
  float* const rawQ = reinterpret_cast<float*>(dsb);       // [MAXL][8]
  float* const rawG = reinterpret_cast<float*>(Q2h);       // [MAXL][8]
  float* const rawO = reinterpret_cast<float*>(G2h);       // [MAXL][8]
  const int NQ = (a.Lq + BK - 1) / BK;
  auto stage_row = [&](float* dst, const float* src) {
    if (((uintptr_t)src & 15) == 0) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst + 4)), "l"(src + 4) : "memory");
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) dst[c] = src[c];
    }
  };
#pragma unroll
  for (int u = 0; u < B1_RPT; ++u) {
    const int i = tid + u * B1_THREADS;
    if (i < a.Lq) {
      stage_row(rawQ + i * 8, a.q + ((long long)n * a.Lq + i) * a.ldq + h * 8);
      stage_row(rawG + i * 8, a.dO + ((long long)n * a.Lq + i) * a.lddo + h * 8);
      stage_row(rawO + i * 8, a.O + ((long long)n * a.Lq + i) * a.ldo + h * 8);
    } else if (i < NQ * BK) {
#pragma unroll
      for (int c = 0; c < 8; ++c) { rawQ[i * 8 + c] = 0.f; rawG[i * 8 + c] = 0.f; rawO[i * 8 + c] = 0.f; }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  float lse2[B1_RPT];
  uint32_t gmb = 0u, qmb = 0u, vmb = 0u;
  {
    float vv[B1_RPT][8];
#pragma unroll
    for (int u = 0; u < B1_RPT; ++u) {
      const int i = tid + u * B1_THREADS;
#pragma unroll
      for (int c = 0; c < 8; ++c) vv[u][c] = 0.f;
      lse2[u] = INFINITY;
      if (i < a.Lk) ld8g(vv[u], a.v + ((long long)n * a.Lk + i) * a.ldv + h * 8);
      if (i < a.Lq) lse2[u] = a.LSE[(long long)nh * a.Lq + i];
    }
#pragma unroll
    for (int u = 0; u < B1_RPT; ++u) vmb = absmax8_bits(vv[u], vmb);
  }
  if (tid == 0) {
    uint64_t* b = s.bars;
    mbar_init(&b[B_X], 128); mbar_init(&b[B_S], 1); mbar_init(&b[B_F], 128); mbar_init(&b[B_P], 128); mbar_init(&b[B_OF], 1); mbar_init(&b[B_O], 1);
    mbar_init(&b[6], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s.pre[33] = 0u; s.pre[34] = 0u; s.pre[35] = 0u;
  }
  if (warp == 4) tmem_alloc<256>(s.tmem);
  if (tid < 8) nrm[tid] = 0u;
  const int LkC = compact_keys<B1_THREADS>(a, s, n, tid, warp, lane);
  const int nKT = (LkC + TCQ - 1) / TCQ;
  for (int j = tid; j < a.Lk; j += B1_THREADS) {          // slot -> key index; masked keys get zero gradients
    const int c = key_slot(s, j);
    if (c >= 0) { s.idx[c] = (uint16_t)j; continue; }
    float z[8];
#pragma unroll
    for (int c2 = 0; c2 < 8; ++c2) z[c2] = 0.f;
    st8g(a.dk + ((long long)n * a.Lk + j) * a.lddk + h * 8, z);
    st8g(a.dv + ((long long)n * a.Lk + j) * a.lddv + h * 8, z);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // maxima and delta = rowsum(dO * O) from the staged rows; per-query tables (delta is rescaled once the normalisers are known)
  {
    const float init = (LkC > 0 || (a.flags & kAttnPartial)) ? 0.f : __int_as_float(0x7fc00000);
#pragma unroll
    for (int u = 0; u < B1_RPT; ++u) {
      const int i = tid + u * B1_THREADS;
      if (i >= NQ * BK) continue;
      float q[8], g[8], o[8], d = 0.f;
      *reinterpret_cast<float4*>(q) = *reinterpret_cast<const float4*>(rawQ + i * 8); *reinterpret_cast<float4*>(q + 4) = *reinterpret_cast<const float4*>(rawQ + i * 8 + 4);
      *reinterpret_cast<float4*>(g) = *reinterpret_cast<const float4*>(rawG + i * 8); *reinterpret_cast<float4*>(g + 4) = *reinterpret_cast<const float4*>(rawG + i * 8 + 4);
      *reinterpret_cast<float4*>(o) = *reinterpret_cast<const float4*>(rawO + i * 8); *reinterpret_cast<float4*>(o + 4) = *reinterpret_cast<const float4*>(rawO + i * 8 + 4);
      qmb = absmax8_bits(q, qmb); gmb = absmax8_bits(g, gmb);
#pragma unroll
      for (int c = 0; c < 8; ++c) { d = fmaf(g[c], o[c], d); o[c] = init; }
      if (i < a.Lq) st8g(a.dq + ((long long)n * a.Lq + i) * a.lddq + h * 8, o);       // dq starts at zero (NaN if every key is masked)
      s.f0[i] = -lse2[u] * kLog2e; s.f1[i] = -d;
      w16[i] = dc.on ? (uint16_t)drop_row_word(dc, nh, a.Lq, i < a.Lq ? i : 0) : (uint16_t)0;
    }
  }
  gmb = __reduce_max_sync(0xffffffffu, gmb); qmb = __reduce_max_sync(0xffffffffu, qmb); vmb = __reduce_max_sync(0xffffffffu, vmb);
  if (lane == 0) { atomicMax(&s.pre[33], gmb); atomicMax(&s.pre[34], qmb); atomicMax(&s.pre[35], vmb); }      // cleared before the barriers of compact_keys
  __syncthreads();                                          // maxima final; every raw O row has been consumed
  const float cs_scale = pow2_normaliser_c(__uint_as_float(s.pre[33]));
  const float q_norm = pow2_normaliser_c(__uint_as_float(s.pre[34]));       // Q2h holds q * sqrt(1/8) * log2(e) * q_norm
  const float v_norm = pow2_normaliser_c(__uint_as_float(s.pre[35]));       // V rows enter TMEM as v * v_norm: dS carries cs_scale * v_norm
  {
    const float dsc = cs_scale * v_norm;
#pragma unroll
    for (int u = 0; u < B1_RPT; ++u) {                      // dO rows (in the Q2h region) -> G2h (where raw O was)
      const int i = tid + u * B1_THREADS;
      if (i >= NQ * BK) continue;
      float g[8];
      *reinterpret_cast<float4*>(g) = *reinterpret_cast<const float4*>(rawG + i * 8); *reinterpret_cast<float4*>(g + 4) = *reinterpret_cast<const float4*>(rawG + i * 8 + 4);
#pragma unroll
      for (int c = 0; c < 8; ++c) g[c] *= cs_scale;
      put_l2h(G2h, i, g);
      s.f1[i] *= dsc;
    }
    __syncthreads();                                        // every raw dO row has been consumed: Q2h may be written
    const float qs = kQScale * q_norm;
#pragma unroll
    for (int u = 0; u < B1_RPT; ++u) {                      // q rows (in the dS^T buffer) -> Q2h
      const int i = tid + u * B1_THREADS;
      if (i >= NQ * BK) continue;
      float q[8];
      *reinterpret_cast<float4*>(q) = *reinterpret_cast<const float4*>(rawQ + i * 8); *reinterpret_cast<float4*>(q + 4) = *reinterpret_cast<const float4*>(rawQ + i * 8 + 4);
#pragma unroll
      for (int c = 0; c < 8; ++c) q[c] *= qs;
      put_l2h(Q2h, i, q);
    }
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = *s.tmem;
  uint64_t* const bars = s.bars;
  uint64_t* const bdq = s.bars + 6;

  if (warp == 4) {
    const uint32_t idK = idesc_f16(128, 16), idQ = idesc_f16_mn(128, 16, true, true);
    const uint32_t aQ2 = smem_u32(Q2h), aG2 = smem_u32(G2h);
    const uint32_t aDS = smem_u32(dsb), aKB = smem_u32(kbb);
    const uint32_t tw = tb;
    const uint32_t idF = idesc_f16_mn(128, BK, false, true);
    auto issue_st = [&](int j) {
      const uint32_t d = tw + C_IN, x = tw + E_X;
      const uint32_t off = (uint32_t)j * (BK / 8) * 128;
      mma_ts_f16(d, x, smem_desc(aQ2 + off, 0, 128), idF, 0);                          // [Khi | Klo] x [Qhi | Qhi]
      mma_ts_f16(d, x, smem_desc(aQ2 + HALF_ARR * 2 + off, 0, 128), idF, 1);           // [Khi | Klo] x [Qlo | Qlo]
      mma_ts_f16(d + 64, x + 8, smem_desc(aG2 + off, 0, 128), idF, 0);                 // [Vhi | Vlo] x [Ghi | Ghi]
      mma_ts_f16(d + 64, x + 8, smem_desc(aG2 + HALF_ARR * 2 + off, 0, 128), idF, 1);  // [Vhi | Vlo] x [Glo | Glo]
    };
    uint32_t cF = 0, cP = 0;
    for (int kt = 0; kt < nKT; ++kt) {
      const int ksteps = (min(TCQ, LkC - kt * TCQ) + 15) >> 4;
      mbar_wait(&bars[B_X], kt & 1);
      fence_after();
      if (elect_one()) { issue_st(0); commit(&bars[B_S]); }
      __syncwarp();
      for (int j = 0; j < NQ; ++j) {
        const bool last = j + 1 == NQ;
        if (!last) {
          mbar_wait(&bars[B_F], cF & 1); cF++;
          fence_after();
          if (elect_one()) { issue_st(j + 1); commit(&bars[B_S]); }
          __syncwarp();
        }
        mbar_wait(&bars[B_P], cP & 1); cP++;
        fence_after();
        if (elect_one()) {
          const int nsteps = (min(BK, a.Lq - j * BK) + 15) >> 4;
          const uint32_t dK = tw + E_DK, dV = tw + E_DV;
          for (int t = 0; t < nsteps; ++t) {
            const uint32_t off = (uint32_t)(j * (BK / 16) + t) * 256;
            const uint32_t acc = (j > 0 || t > 0) ? 1u : 0u;
            mma_ts_f16(dV, tw + C_OUT + (uint32_t)t * 8, smem_desc(aG2 + off, 128, HALF_ARR * 2), idK, acc);          // P^T [dOhi | dOlo]
            mma_ts_f16(dK, tw + C_OUT + 32 + (uint32_t)t * 8, smem_desc(aQ2 + off, 128, HALF_ARR * 2), idK, acc);     // dS^T [Qhi | Qlo]
          }
          if (!last) commit(&bars[B_OF]);
          if ((j & 1) || last) {                    // dS K -> dQ of the tile pair (M = 128: two 64-query tiles)
            for (int t = 0; t < ksteps; ++t)
              mma_ss_f16(tw + E_DQ, smem_desc(aDS + t * 256, 128, 2048), smem_desc(aKB + t * 256, 128, 2048), idQ, t > 0 ? 1u : 0u);   // dS [Khi | Klo]
            commit(last ? &bars[B_O] : bdq);
          }
        }
        __syncwarp();
      }
    }
  } else {
    const int r = tid;
    const uint32_t tw = tb + ((uint32_t)(warp * 32) << 16);
    const uint32_t tIN = tw + C_IN, tOUT = tw + C_OUT, tX = tw + E_X;
    unsigned char* const ds_row = dsb + (r & 7) * 16 + (r >> 3) * 128;      // this key's 16-byte slot in each query group
    unsigned char* const kb_row = kbb + (r & 7) * 16 + (r >> 3) * 128;
    const float ds_inv = 1.f / (cs_scale * v_norm);     // undoes the scale dS^T (and what is contracted with it) carries
    float dq_scale = kScale * ds_inv;                   // times 1 / (this key tile's K normaliser), set per tile
    WgPhase ph = {0, 0};
    uint32_t cdq = 0;
    // dQ contribution of the tile pair starting at query tile jq: accumulator row m (query jq*64 + m) sits in lane m
    auto drain_dq = [&](int jq, int nrows) {
      uint32_t v[16];
      tmem_ld16(tw + E_DQ, v); tmem_wait_ld();
      const int i = jq * BK + r;
      if (r < nrows && i < a.Lq) {
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = (__uint_as_float(v[c]) + __uint_as_float(v[8 + c])) * dq_scale;
        red_add8(a.dq + ((long long)n * a.Lq + i) * a.lddq + h * 8, o);
      }
    };
    for (int kt = 0; kt < nKT; ++kt) {
      const int cs = kt * TCQ + r;
      const bool valid = cs < LkC;
      const int jk = valid ? (int)s.idx[cs] : 0;
      float k[8], v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) { k[c] = 0.f; v[c] = 0.f; }
      if (valid) {
        ld8g(k, a.k + ((long long)n * a.Lk + jk) * a.ldk + h * 8);
        ld8g(v, a.v + ((long long)n * a.Lk + jk) * a.ldv + h * 8);
      }
      // K and V rows are fp16 operands ([hi | lo] in TMEM, hi in the dQ product), range-managed by exact powers of two: V per
      // (row, head) (pass 1), K per key tile — the exponent (an fma instead of an add) and the dQ drain undo it for free
      float k_norm;
      {
        const uint32_t kmb = __reduce_max_sync(0xffffffffu, absmax8_bits(k, 0u));
        if (lane == 0) atomicMax(&nrm[kt], kmb);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        k_norm = pow2_normaliser_c(__uint_as_float(nrm[kt]));
#pragma unroll
        for (int c = 0; c < 8; ++c) { k[c] *= k_norm; v[c] *= v_norm; }
      }
      const float inv_qk = 1.f / (q_norm * k_norm);
      const f32x2 iqk2 = pk2(inv_qk, inv_qk);
      dq_scale = kScale * ds_inv / k_norm;
      {                // A operands: columns 0-3 = hi pairs, 4-7 = lo pairs (K index 0-7 hi, 8-15 lo); the K rows also go to
                       // shared memory as the [hi | lo] B operand of the dQ product
        uint32_t xa[8];
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const float* src = pass ? v : k;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const __half2 h2 = __floats2half2_rn(src[2 * c], src[2 * c + 1]);
            const float2 hf = __half22float2(h2);
            xa[c] = *reinterpret_cast<const uint32_t*>(&h2);
            xa[4 + c] = pack_h2(src[2 * c] - hf.x, src[2 * c + 1] - hf.y);
          }
          tmem_st8(tX + pass * 8, xa);
          if (pass == 0) {
            *reinterpret_cast<uint4*>(kb_row) = make_uint4(xa[0], xa[1], xa[2], xa[3]);
            *reinterpret_cast<uint4*>(kb_row + KB_BYTES) = make_uint4(xa[4], xa[5], xa[6], xa[7]);
          }
        }
      }
      fence_async_smem();
      tmem_wait_st();
      fence_before();
      mbar_arrive(&bars[B_X]);
      const uint32_t cw2 = dc.on ? drop_col_word(dc, nh, cs) * 0x00010001u : 0u;      // b_c in both halves
      for (int j = 0; j < NQ; ++j) {
        ph.wait_s(bars);
        uint32_t pk[32], dk2[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t sr[32], tr[32];
          tmem_ld32(tIN + half * 32, sr); tmem_ld32(tIN + 64 + half * 32, tr);
          tmem_wait_ld();
          if (half == 1 && j + 1 < NQ) signal_in_free(bars);
          const float4* l4 = reinterpret_cast<const float4*>(s.f0 + j * BK + half * 32);
          const float4* d4 = reinterpret_cast<const float4*>(s.f1 + j * BK + half * 32);
          const uint2* w4 = reinterpret_cast<const uint2*>(w16 + j * BK + half * 32);      // 16-bit words: 4 queries per uint2
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const float4 lv = l4[cc], dv = d4[cc];
            float pp[4], ss[4];
            const int c = cc * 4;
            float x0, x1, x2, x3;
            upk2(fma2(pk2(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), iqk2, pk2(lv.x, lv.y)), x0, x1);
            upk2(fma2(pk2(__uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3])), iqk2, pk2(lv.z, lv.w)), x2, x3);
            const float p0 = ex2(x0), p1 = ex2(x1), p2 = ex2(x2), p3 = ex2(x3);
            const f32x2 pa = pk2(p0, p1), pb = pk2(p2, p3);
            const f32x2 ta = pk2(__uint_as_float(tr[c]), __uint_as_float(tr[c + 1])), tb2 = pk2(__uint_as_float(tr[c + 2]), __uint_as_float(tr[c + 3]));
            if (!dc.on) {
              pp[0] = p0; pp[1] = p1; pp[2] = p2; pp[3] = p3;
              upk2(mul2(pa, add2(ta, pk2(dv.x, dv.y))), ss[0], ss[1]);
              upk2(mul2(pb, add2(tb2, pk2(dv.z, dv.w))), ss[2], ss[3]);
            } else {
              const uint2 wv = w4[cc];
              float m0, m1, m2, m3;
              drop_mult2(wv.x ^ cw2, dc.thr2, dc.scale, m0, m1);
              drop_mult2(wv.y ^ cw2, dc.thr2, dc.scale, m2, m3);
              const f32x2 ma = pk2(m0, m1), mb = pk2(m2, m3);
              upk2(mul2(pa, ma), pp[0], pp[1]);
              upk2(mul2(pb, mb), pp[2], pp[3]);
              upk2(mul2(pa, fma2(ta, ma, pk2(dv.x, dv.y))), ss[0], ss[1]);
              upk2(mul2(pb, fma2(tb2, mb, pk2(dv.z, dv.w))), ss[2], ss[3]);
            }
            pk[half * 16 + cc * 2] = pack_h2(pp[0], pp[1]); pk[half * 16 + cc * 2 + 1] = pack_h2(pp[2], pp[3]);
            dk2[half * 16 + cc * 2] = pack_h2(ss[0], ss[1]); dk2[half * 16 + cc * 2 + 1] = pack_h2(ss[2], ss[3]);
          }
        }
        const int par = j & 1;
        if (j > 0) ph.wait_out_free(bars);
        if (par == 0 && j > 0) {
          mbar_wait(bdq, cdq & 1); cdq++;       // dQ product of the previous tile pair done: read its accumulator, reuse the dS^T buffer
          fence_after();
          drain_dq(j - 2, 2 * BK);
        }
        tmem_st32(tOUT, pk); tmem_st32(tOUT + 32, dk2);
        // padded key rows (zero K, but P = 2^(-lse) may be huge) must contribute exact zeros to dQ
#pragma unroll
        for (int g8 = 0; g8 < 8; ++g8)
          *reinterpret_cast<uint4*>(ds_row + (par * 8 + g8) * 2048) = valid ? make_uint4(dk2[g8 * 4], dk2[g8 * 4 + 1], dk2[g8 * 4 + 2], dk2[g8 * 4 + 3]) : make_uint4(0u, 0u, 0u, 0u);
        fence_async_smem();
        tmem_wait_st();
        fence_before();
        mbar_arrive(&bars[B_P]);
      }
      mbar_wait(&bars[B_O], kt & 1);
      fence_after();
      if (NQ & 1) drain_dq(NQ - 1, BK); else drain_dq(NQ - 2, 2 * BK);     // an odd last tile sits alone in rows 0..63
      uint32_t o[32];
      tmem_ld32(tw + E_DK, o); tmem_wait_ld();                 // dK hi | lo | dV hi | lo
      if (valid) {
        float dk[8], dv[8];
        const float ck = kLn2 * ds_inv / q_norm;               // hi + lo parts; Q carried log2(e) and its normaliser
        const float cv = 1.f / cs_scale;                       // dV = P^T (dO cs_scale)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          dk[c] = (__uint_as_float(o[c]) + __uint_as_float(o[8 + c])) * ck;
          dv[c] = (__uint_as_float(o[16 + c]) + __uint_as_float(o[24 + c])) * cv;
        }
        st8g(a.dk + ((long long)n * a.Lk + jk) * a.lddk + h * 8, dk);
        st8g(a.dv + ((long long)n * a.Lk + jk) * a.lddv + h * 8, dv);
      }
      fence_before();
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 4) { fence_after(); tmem_dealloc<256>(tb); }
}

// ------------------------------------------------------------------------------------------------
static bool env_flag(const char* name) { const char* e = getenv(name); return e && e[0] && e[0] != '0'; }

bool attn_tc_eligible(const AttnArgs& a) {
  static const bool off = env_flag("VAESNE_NO_TC");
  if (off) return false;
  // from 96 tokens on either side: measured per score element (forward + backward, B200) the tensor-core kernels pass the
  // one-CTA-per-row CUDA-core kernels (attn_mid.cu) between 64 and 96 tokens — 10.7 vs 8.8 ps at 64, 6.2 vs 8.7 at 96, 4.0 vs
  // 9.1 at 128, 2.1 at 256 — so there is no cliff between the windows (tests/edge_cases.py run_window_timing)
  static const int lmin = [] { const char* e = getenv("VAESNE_TC_MIN"); const int v = e ? atoi(e) : 0; return v >= 16 ? v : 96; }();
  return a.Lq >= lmin && a.Lk >= lmin && a.Lk <= MAXL && a.Lq <= MAXL && a.N <= 65535;
}
bool attn_tc_has_bwd() { return true; }

template <typename K>
static int tc_configure(K k, size_t bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { set_error("%s: cannot reserve %zu B of shared memory: %s", what, bytes, cudaGetErrorString(e)); return V_ECUDA; }
  (void)cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);   // two CTAs per SM need the full window
  return V_OK;
}

// The per-device opt-in to the large dynamic shared-memory window is applied once per (kernel, device ordinal): a process may
// drive several GPUs from one thread.  (Keyed by the function pointer: both kernels have the same C++ type.)
static int tc_configure_dev(void (*k)(AttnArgs), size_t bytes, const char* what) {
  struct Entry { const void* fn; int dev; };
  static thread_local Entry done[32] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  for (const Entry& e : done) if (e.fn == (const void*)k && e.dev == dev) return V_OK;
  const int rc = tc_configure(k, bytes, what);
  if (rc == V_OK) for (Entry& e : done) if (!e.fn) { e.fn = (const void*)k; e.dev = dev; break; }
  return rc;
}

int attn_tc_fwd(const AttnArgs& a, cudaStream_t st) {
  static const bool one_cta = env_flag("VAESNE_TC_FWD4");      // the one-CTA-per-SM, four-warpgroup launch (comparison)
  if (one_cta) {
    const int cfg = tc_configure_dev(attn_tc_fwdN_kernel<4>, FWD4_SMEM, "attn_tc_fwd4");
    if (cfg) return cfg;
    attn_tc_fwdN_kernel<4><<<dim3(kH, a.N), dim3(F4_THREADS), FWD4_SMEM, st>>>(a);
    return check_launch("attn_tc_fwd4");
  }
  const int cfg = tc_configure_dev(attn_tc_fwdN_kernel<2>, FWD4_SMEM, "attn_tc_fwd2x2");
  if (cfg) return cfg;
  attn_tc_fwdN_kernel<2><<<dim3(kH, a.N), dim3(F2_THREADS), FWD4_SMEM, st>>>(a);
  return check_launch("attn_tc_fwd2x2");
}

int attn_tc_bwd(const AttnArgs& a, cudaStream_t st) {
  static const bool two_wg = env_flag("VAESNE_TC_BWD2");      // the two-warpgroup, one-CTA-per-SM launch (comparison)
  if (two_wg) {
    const int cfg = tc_configure_dev(attn_tc_bwd_kernel, BWD_SMEM, "attn_tc_bwd");
    if (cfg) return cfg;
    attn_tc_bwd_kernel<<<dim3(kH, a.N), dim3(NTHREADS), BWD_SMEM, st>>>(a);
    return check_launch("attn_tc_bwd");
  }
  const int cfg = tc_configure_dev(attn_tc_bwd1_kernel, BWD1_SMEM, "attn_tc_bwd1");
  if (cfg) return cfg;
  attn_tc_bwd1_kernel<<<dim3(kH, a.N), dim3(B1_THREADS), BWD1_SMEM, st>>>(a);
  return check_launch("attn_tc_bwd1");
}

}  // namespace vaesne
extern "C" int vaesne_debug_tc_prof(long long* out16) { return (int)cudaMemcpyFromSymbol(out16, vaesne::g_tc_prof, 16 * sizeof(long long)); }
extern "C" int vaesne_debug_tc(int flags) { return (int)cudaMemcpyToSymbol(vaesne::g_tc_dbg, &flags, sizeof(int)); }

