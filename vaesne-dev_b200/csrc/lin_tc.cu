// Blackwell-native token-wise linear layers of the transformer blocks (model_dim 32):
//   forward   Y = act(X W^T + b)            N in {32, 64, 96}   (in_proj, q / kv projections, ffn.0, MLP heads)
//             Y = LayerNorm(R + dropout(X W^T + b))   N = 32    (out_proj / ffn.2 + residual + LN, util_layers.py:292,303,307)
//   backward  dZ from (LayerNorm+dropout | activation) ; dX = dZ W ; dW += dZ^T X ; db, dgamma, dbeta
// Same arithmetic as lin.cu (its docstring cites the reference lines); restructured for sm_100a because these
// kernels move 384..640 B per token and must run at HBM speed:
//
//   * every global access is coalesced and moved in 128-token tiles: the forward and the pipelined backward use the TMA
//     unit (cp.async.bulk.tensor, 128-byte swizzle: a thread reads / writes ITS row with conflict-free 16-byte shared
//     accesses), the round-1 backward 16-byte cp.async into rows padded to 144 B.
//   * thread r owns token r of the tile = TMEM lane r: the LayerNorm / activation / dropout epilogues are
//     plain per-thread register code.
//   * the 32-wide contractions run on the tensor core: the row is split into tf32 hi + lo parts and stored
//     to TMEM as the A operand (tcgen05.st), W (hi + lo, K-major canonical layout) is staged once per CTA,
//     3 MMAs per 8-wide K step restore fp32-level products (kind::tf32 truncates its inputs), fp32
//     accumulation in TMEM, one elected lane issues, tcgen05.commit -> mbarrier.
//   * dW is a contraction over the 128 tokens of the tile, accumulated in TMEM across ALL tiles of the persistent CTA:
//     MN-major bf16 hi + lo operand tiles and kind::f16 MMAs in the pipelined backward (lin_tc_bwd2_kernel), transposed
//     tf32 operands in the round-1 backward (lin_tc_bwd_kernel); the parts are summed in shared memory and flushed with
//     one coalesced atomicAdd per 128-byte line per CTA.
//   * db / dgamma / dbeta: column sums read column-wise from the swizzled shared-memory tiles (pipelined backward) or
//     fp32 butterfly transpose-reductions (round-1 backward), shared-memory accumulators, one atomicAdd per element per CTA.
// Kernels: lin_tc_fwd_kernel (3 CTAs per SM), lin_tc_bwd2_kernel (one warp-specialised CTA per SM: TMA producer warp, MMA
// issuer warp, two row groups; the big calls of a training step), lin_tc_bwd_kernel (two CTAs per SM; small token counts,
// dX-only / dW-only calls, the Xadd heads).
#include "common.cuh"
#include "vaesne_b200.h"
#include "lin_args.cuh"
#include "tc_common.cuh"
#include <stdlib.h>
#include <stdio.h>
#include <cuda_bf16.h>
#include <cuda.h>          // CUtensorMap (the encoder itself is fetched through the runtime: no libcuda link)

namespace vaesne {
using namespace tc;

constexpr int LT = 128;              // tokens per tile == threads per CTA == TMEM lanes
constexpr int PITCH = 36;            // floats per padded shared-memory row (144 B)
constexpr int TILE = LT * PITCH;     // floats per staged tile (18 KB)
constexpr int TSW = LT * 32;         // floats per 128-byte-swizzled tile (16 KB, TMA destination)

__device__ __forceinline__ bool lelect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cp_async16(float* smem, const float* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// rows [t0, t0 + rows) x 32 floats of a row-strided matrix -> padded tile (coalesced: 8 lanes per 128-byte row)
__device__ __forceinline__ void load_tile(float* dst, const float* src, long long ld, long long t0, int rows, int tid) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int id = it * LT + tid, r = id >> 3, c = id & 7;
    if (r < rows) cp_async16(dst + r * PITCH + c * 4, src + (t0 + r) * ld + c * 4);
  }
}
template <bool ACC>
__device__ __forceinline__ void store_tile(float* dst, long long ld, const float* src, long long t0, int rows, int tid) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int id = it * LT + tid, r = id >> 3, c = id & 7;
    if (r < rows) {
      float4 v = *reinterpret_cast<const float4*>(src + r * PITCH + c * 4);
      float4* g = reinterpret_cast<float4*>(dst + (t0 + r) * ld + c * 4);
      if (ACC) { const float4 o = *g; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
      *g = v;
    }
  }
}
// TMA: one 128-token x 32-float tile (16 KB) per bulk tensor copy, 128-byte swizzle: the 16-byte chunk c of row r lands at
// chunk position c ^ (r & 7), which makes the row-per-thread 16-byte reads below conflict-free without padding.
__device__ __forceinline__ void tma_load_tile(float* dst, const CUtensorMap* tm, int row0, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(smem_u32(dst)), "l"(tm), "r"(0), "r"(row0), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lds_row_sw(float* v, const float* tile, int r) {
  const float* row = tile + r * 32;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(row + ((j ^ (r & 7)) << 2));
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
__device__ __forceinline__ void sts_row_sw(float* tile, int r, const float* v) {
  float* row = tile + r * 32;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(row + ((j ^ (r & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
// swizzled [128][32] staging tile -> global rows (coalesced: 8 lanes per 128-byte row)
__device__ __forceinline__ void store_tile_sw(float* dst, long long ld, const float* src, long long t0, int rows, int tid) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int id = it * LT + tid, r = id >> 3, c = id & 7;
    if (r < rows)
      *reinterpret_cast<float4*>(dst + (t0 + r) * ld + c * 4) = *reinterpret_cast<const float4*>(src + r * 32 + ((c ^ (r & 7)) << 2));
  }
}
__device__ __forceinline__ void lds_row(float* v, const float* row) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = reinterpret_cast<const float4*>(row)[j];
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
__device__ __forceinline__ void sts_row(float* row, const float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(row)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void tmem_put32(uint32_t taddr, const float* x) {
  uint32_t u[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) u[c] = __float_as_uint(x[c]);
  tmem_st32(taddr, u);
}
__device__ __forceinline__ float rna_tf32(float x) { uint32_t h; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x)); return __uint_as_float(h); }

// lane l <- sum over the warp of v[l]   (v is destroyed)
__device__ __forceinline__ float warp_colsum32(float* v, int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// K-major canonical (no swizzle) operand of a [rows x 32] matrix, contraction over the 32 columns:
// 8-row groups of 1 KB, inside a group 8 chunks (4 columns = 16 B) of 128 B.   LBO = 128 B, SBO = 1024 B.
__device__ __forceinline__ int wcanon(int row, int k) { return (row >> 3) * 256 + (k >> 2) * 32 + (row & 7) * 4 + (k & 3); }
// transposed operand of a 128-token tile, contraction over the tokens: element (row, token t); chunks of 4 tokens
// 144 B apart (conflict-free scalar stores by 32 consecutive tokens), 8-row groups 4608 B apart.
__device__ __forceinline__ int tcanon(int row, int t) { return (row >> 3) * 1152 + (t >> 2) * 36 + (row & 7) * 4 + (t & 3); }

// =================================================================================================
// forward
// =================================================================================================
template <int NCH>
__host__ __device__ constexpr size_t lin_tc_fwd_smem() { return 128 + sizeof(float) * (2 * NCH * 1024 + 96 + 64 + 4 * TSW) + 64; }

template <int NCH, bool LN>
__global__ void __launch_bounds__(LT, 3) lin_tc_fwd_kernel(LinFwd a, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmR) {
  constexpr int N = NCH * 32;
  constexpr int COLS = (64 + N) <= 128 ? 128 : 256;
  extern __shared__ __align__(1024) unsigned char lin_tc_raw[];      // plain pointer arithmetic from here on: keeps LDS/STS
  // two input sets {X tile, R | Xadd tile} (TMA destinations, 128-byte-swizzled [128][32] = 16 KB each); the set of the
  // tile being processed doubles as output staging (same swizzle) once its rows are in registers, while the TMA unit
  // fills the other set with the next tile
  float* IN = reinterpret_cast<float*>(lin_tc_raw);
  float* Whi = IN + 4 * TSW;
  float* Wlo = Whi + NCH * 1024;
  float* sB = Wlo + NCH * 1024;      // [96]
  float* sG = sB + 96;               // [32]
  float* sBe = sG + 32;              // [32]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sBe + 32);
  uint64_t* tbar = bar + 1;          // [2] TMA transaction barriers, one per input set
  uint32_t* tmem_s = reinterpret_cast<uint32_t*>(tbar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < N * 32; i += LT) {
    const int n = i >> 5, k = i & 31;
    float hi, lo;
    split_tf32(a.W[i], hi, lo);
    Whi[wcanon(n, k)] = hi; Wlo[wcanon(n, k)] = lo;
  }
  for (int i = tid; i < N; i += LT) sB[i] = a.b ? a.b[i] : 0.f;
  if (LN && tid < 32) { sG[tid] = a.gamma[tid]; sBe[tid] = a.beta[tid]; }
  if (tid == 0) { mbar_init(bar, 1); mbar_init(&tbar[0], 1); mbar_init(&tbar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc<COLS>(tmem_s);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = *tmem_s;
  const uint32_t tl = tb + ((uint32_t)(warp * 32) << 16);      // this thread's lane
  const bool second = LN || (a.Xadd != nullptr);
  auto tma_issue = [&](int tile, int set) {      // one thread; rows beyond T are zero-filled by the TMA unit
    mbar_expect_tx(&tbar[set], second ? 2u * LT * 128u : LT * 128u);
    tma_load_tile(IN + set * 2 * TSW, &tmX, tile * LT, &tbar[set]);
    if (second) tma_load_tile(IN + set * 2 * TSW + TSW, &tmR, tile * LT, &tbar[set]);
  };
  const uint32_t idesc = idesc_tf32(128, N);
  const uint32_t aWhi = smem_u32(Whi), aWlo = smem_u32(Wlo);
  const DropCfg dc = make_drop(LN ? a.p_drop : 0.f, a.seed, a.stream_id);
  uint32_t ph = 0;

  const int ntiles = (a.T + LT - 1) / LT;
  if (tid == 0 && (int)blockIdx.x < ntiles) tma_issue(blockIdx.x, 0);
  int kiter = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++kiter) {
    const int set = kiter & 1;
    float* bufX = IN + set * 2 * TSW; float* bufR = bufX + TSW;
    const long long t0 = (long long)tile * LT;
    const int rows = min(LT, a.T - (int)t0);
    const long long t = t0 + tid;
    // the other set was released by the barrier that ended the previous iteration: prefetch the next tile into it
    if (tid == 0 && tile + (int)gridDim.x < ntiles) tma_issue(tile + gridDim.x, set ^ 1);
    mbar_wait(&tbar[set], (uint32_t)((kiter >> 1) & 1));
    float res[LN ? 32 : 1];          // the residual row is read now, before the tiles turn into staging buffers
    if (LN) lds_row_sw(res, bufR, tid);
    {
      float x[32], hi[32], lo[32];
      lds_row_sw(x, bufX, tid);
      if (!LN && a.Xadd) {
        float xa[32];
        lds_row_sw(xa, bufR, tid);
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] += xa[j];
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) split_tf32(x[j], hi[j], lo[j]);
      tmem_put32(tl, hi); tmem_put32(tl + 32, lo);
      tmem_wait_st();
    }
    fence_before();
    __syncthreads();
    if (warp == 0) {
      fence_after();
      if (lelect_one()) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const uint64_t bhi = smem_desc(aWhi + s * 256, 128, 1024), blo = smem_desc(aWlo + s * 256, 128, 1024);
          mma_ts(tb + 64, tb + s * 8, bhi, idesc, s > 0 ? 1u : 0u);
          mma_ts(tb + 64, tb + 32 + s * 8, bhi, idesc, 1u);
          mma_ts(tb + 64, tb + s * 8, blo, idesc, 1u);
        }
        commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, ph); ph ^= 1u;
    fence_after();
    if (LN) {
      uint32_t d[32];
      tmem_ld32(tl + 64, d); tmem_wait_ld();
      float s[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) s[j] = res[LN ? j : 0];
      float mean = 0.f;
      const uint32_t rh = dc.on ? drop_row_hash(dc, (uint64_t)t) : 0u;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float v = __uint_as_float(d[j]) + sB[j];
        if (dc.on) v *= drop_mult_row(dc, rh, j);
        s[j] += v;
        mean += s[j];
      }
      mean *= (1.f / 32);
      float var = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) { const float dd = s[j] - mean; var = fmaf(dd, dd, var); }
      const float rstd = 1.f / sqrtf(var * (1.f / 32) + a.eps);
      if (a.S) sts_row_sw(bufX, tid, s);
#pragma unroll
      for (int j = 0; j < 32; ++j) s[j] = (s[j] - mean) * rstd * sG[j] + sBe[j];
      sts_row_sw(bufR, tid, s);
      __syncthreads();
      if (a.S) store_tile_sw(a.S, 32, bufX, t0, rows, tid);
      store_tile_sw(a.Y, a.ldy, bufR, t0, rows, tid);
    } else {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t d[32];
        tmem_ld32(tl + 64 + c * 32, d); tmem_wait_ld();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(d[j]) + sB[c * 32 + j];
        if (c > 0) __syncthreads();                 // previous chunk's coalesced stores have read the staging tiles
        if (a.H) sts_row_sw(bufR, tid, v);
        if (a.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (a.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
        sts_row_sw(bufX, tid, v);
        __syncthreads();
        if (a.H) store_tile_sw(a.H + c * 32, a.ldh, bufR, t0, rows, tid);
        store_tile_sw(a.Y + c * 32, a.ldy, bufX, t0, rows, tid);
      }
    }
    fence_async_smem();     // generic-proxy accesses of the staging tiles are ordered before the next TMA (async proxy) write
    fence_before();
    __syncthreads();        // staging tiles are free again; D has been read by every thread
    fence_after();
  }
  fence_before();
  __syncthreads();
  if (warp == 0) { fence_after(); tmem_dealloc<COLS>(tb); }
}

// =================================================================================================
// backward
// =================================================================================================
// 256 threads per CTA: TWO threads per token row (thread = (row r, column half hf)); warps 0-3 own columns 0..15, warps
// 4-7 columns 16..31 of TMEM lanes / rows 0..127.  Halving the per-thread work and doubling the warps is what hides
// the latency of the ~1000-instruction per-row arithmetic (LayerNorm backward, dropout, hi/lo splits, transposes).
// shared memory (tiles of 18 KB):  dZT_hi | B1 = dZT_lo | XT_hi | B2 = XT_lo | B0   then W^T hi/lo and the small tables.
//   B0: dY chunk (cp.async) -> dR staging.   B1: S / activation input (cp.async), then, once every thread has read
//   its row, the lo half of dZ^T.   B2: X tile (cp.async), then the lo half of X^T, then (after the MMAs) dX staging.
// dW for one 32-row chunk of W: ONE MMA per 8-token K step with A = [dZ^T_hi ; dZ^T_lo] stacked along M (rows 0-31 /
// 32-63) and B = [X^T_hi ; X^T_lo] stacked along N (64 columns): D[n][k] + D[n][32+k] + D[32+n][k] is the fp32-level
// product (hi*hi + hi*lo + lo*hi), accumulated in TMEM across all tiles of the CTA.
constexpr int BT = 256;              // threads of the backward CTA
template <int NCH>
__host__ __device__ constexpr size_t lin_tc_bwd_smem() { return 128 + sizeof(float) * (2 * NCH * 1024 + 5 * TILE + 32 * 3 + 64 + 128 * 8) + 32; }

__device__ __forceinline__ void load_tile256(float* dst, const float* src, long long ld, long long t0, int rows, int tid) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int id = it * BT + tid, r = id >> 3, c = id & 7;
    if (r < rows) cp_async16(dst + r * PITCH + c * 4, src + (t0 + r) * ld + c * 4);
  }
}
template <bool ACC>
__device__ __forceinline__ void store_tile256(float* dst, long long ld, const float* src, long long t0, int rows, int tid) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int id = it * BT + tid, r = id >> 3, c = id & 7;
    if (r < rows) {
      float4 v = *reinterpret_cast<const float4*>(src + r * PITCH + c * 4);
      float4* g = reinterpret_cast<float4*>(dst + (t0 + r) * ld + c * 4);
      if (ACC) { const float4 o = *g; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
      *g = v;
    }
  }
}
__device__ __forceinline__ void lds_half(float* v, const float* p) {          // 16 floats
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t = reinterpret_cast<const float4*>(p)[j];
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
__device__ __forceinline__ void sts_half(float* p, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(p)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void tmem_put16(uint32_t taddr, const float* x) {
  uint32_t u[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) u[c] = __float_as_uint(x[c]);
  tmem_st16(taddr, u);
}
// lane l (< 16) <- sum over the warp of v[l]   (v is destroyed; lanes >= 16 hold partial garbage)
__device__ __forceinline__ float warp_colsum16(float* v, int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <int NCH, bool LN>
__global__ void __launch_bounds__(BT, 2) lin_tc_bwd_kernel(LinBwd a) {
  constexpr int N = NCH * 32;
  constexpr int COLS = 256;            // A hi|lo 64 + dX 32 + dW 64*NCH  (NCH <= 2)
  extern __shared__ __align__(1024) unsigned char lin_tc_raw[];
  float* dZT = reinterpret_cast<float*>(lin_tc_raw);   // first: the M=128 MMA reads 16 row groups from here
  float* B1 = dZT + TILE;
  float* XT = B1 + TILE;
  float* B2 = XT + TILE;
  float* B0 = B2 + TILE;
  float* WThi = B0 + TILE;            // [32 rows k][N] canonical: (k>>3)*(N*8) + (n>>2)*32 + (k&7)*4 + (n&3)
  float* WTlo = WThi + NCH * 1024;
  float* sG = WTlo + NCH * 1024;      // [32]
  float* sDg = sG + 32; float* sDbe = sDg + 32;   // [32] each
  float* sDb = sDbe + 32;             // [64]
  float* red = sDb + 64;              // [128 rows][2 halves][4]: row-reduction exchange between the two threads of a row
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 128 * 8);
  uint32_t* tmem_s = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = tid & 127, hf = tid >> 7, c0 = hf * 16;       // row of the tile, column half, first column
  const bool wgrad = a.dW != nullptr;

  for (int i = tid; i < N * 32; i += BT) {
    const int n = i >> 5, k = i & 31;
    float hi, lo;
    split_tf32(a.W[i], hi, lo);
    const int o = (k >> 3) * (N * 8) + (n >> 2) * 32 + (k & 7) * 4 + (n & 3);
    WThi[o] = hi; WTlo[o] = lo;
  }
  if (tid < 32) { sG[tid] = LN ? a.gamma[tid] : 0.f; sDg[tid] = 0.f; sDbe[tid] = 0.f; }
  if (tid < 64) sDb[tid] = 0.f;
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc<COLS>(tmem_s);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = *tmem_s;
  const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t idX = idesc_tf32(128, 32), idW = idesc_tf32(128, 64);
  const uint32_t aWThi = smem_u32(WThi), aWTlo = smem_u32(WTlo), aZT = smem_u32(dZT), aXT = smem_u32(XT);
  const DropCfg dc = make_drop(LN ? a.p_drop : 0.f, a.seed, a.stream_id);
  float* my0 = B0 + r * PITCH + c0; float* my1 = B1 + r * PITCH + c0; float* my2 = B2 + r * PITCH + c0;
  uint32_t ph = 0;
  bool pending = false;          // an MMA batch has been committed and not yet waited for
  bool first_tile = true;
  float accg[LN ? 16 : 1], accb[LN ? 16 : 1], accd[NCH][16];     // dgamma / dbeta / db partials of this thread's rows
#pragma unroll
  for (int j = 0; j < (LN ? 16 : 1); ++j) { accg[j] = 0.f; accb[j] = 0.f; }
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int j = 0; j < 16; ++j) accd[c][j] = 0.f;

  const int ntiles = (a.T + LT - 1) / LT;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long t0 = (long long)tile * LT;
    const int rows = min(LT, a.T - (int)t0);
    const long long t = t0 + r;
    const bool active = r < rows;
    // the previous tile's MMAs read dZT (incl. B1) and XT (incl. B2): they must be done before the loads land
    if (pending) { mbar_wait(bar, ph); ph ^= 1u; pending = false; fence_after(); }
    __syncthreads();
    if (wgrad) load_tile256(B2, a.X, a.ldx, t0, rows, tid);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (c > 0) {               // chunk c-1's MMAs read B1 (= dZT_lo); B0 was consumed before its barrier
        if (pending) { mbar_wait(bar, ph); ph ^= 1u; pending = false; fence_after(); }
        __syncthreads();
      }
      load_tile256(B0, a.dY + c * 32, a.lddy, t0, rows, tid);
      if (LN) load_tile256(B1, a.S, 32, t0, rows, tid);
      else if (a.act != 0) load_tile256(B1, a.A + c * 32, a.lda, t0, rows, tid);
      cp_async_wait_all();
      __syncthreads();
      float dz[16];
      lds_half(dz, my0);
      if (LN) {
        float sv[16], g[16];
        lds_half(sv, my1);
        float ps = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) ps += sv[j];
        red[(r * 2 + hf) * 4] = ps;
        __syncthreads();
        const float mean = (red[(r * 2) * 4] + red[(r * 2 + 1) * 4]) * (1.f / 32);
        float pv = 0.f, pg = 0.f, pgd = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          sv[j] -= mean; pv = fmaf(sv[j], sv[j], pv);
          g[j] = dz[j] * sG[c0 + j];
          pg += g[j]; pgd = fmaf(g[j], sv[j], pgd);
        }
        red[(r * 2 + hf) * 4 + 1] = pv; red[(r * 2 + hf) * 4 + 2] = pg; red[(r * 2 + hf) * 4 + 3] = pgd;
        __syncthreads();
        const float var = (red[(r * 2) * 4 + 1] + red[(r * 2 + 1) * 4 + 1]) * (1.f / 32);
        const float rstd = 1.f / sqrtf(var + a.eps);
        const float m1 = (red[(r * 2) * 4 + 2] + red[(r * 2 + 1) * 4 + 2]) * (1.f / 32);
        const float m2 = (red[(r * 2) * 4 + 3] + red[(r * 2 + 1) * 4 + 3]) * (1.f / 32) * rstd;
        if (active) {            // dgamma / dbeta: accumulated per thread over all tiles of the CTA, reduced once at the end
#pragma unroll
          for (int j = 0; j < 16; ++j) { accg[j] = fmaf(dz[j] * rstd, sv[j], accg[j]); accb[j] += dz[j]; }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) dz[j] = rstd * (g[j] - m1 - sv[j] * rstd * m2);     // dS = dR
        if (a.dR) sts_half(my0, dz);          // own half row of B0: already consumed by this thread
        if (dc.on) {
          const uint32_t rh = drop_row_hash(dc, (uint64_t)t);
#pragma unroll
          for (int j = 0; j < 16; ++j) dz[j] *= drop_mult_row(dc, rh, c0 + j);
        }
      } else if (a.act != 0) {
        float av[16];
        lds_half(av, my1);
        if (a.act == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) dz[j] = av[j] > 0.f ? dz[j] : 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) dz[j] *= gelu_erf_grad(av[j]);
        }
      }
      if (!active) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dz[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) accd[c][j] += dz[j];
      if (c == 0 && wgrad) {
        // X half row -> X^T hi / lo.  XT_lo aliases the X tile: every thread must have read its row first.
        float x[16];
        lds_half(x, my2);
        if (a.Xadd && active) {               // rare (the MLP heads): the second addend comes straight from global memory
          const float4* xa = reinterpret_cast<const float4*>(a.Xadd + t * a.ldxa + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float4 v4 = xa[j]; x[4 * j] += v4.x; x[4 * j + 1] += v4.y; x[4 * j + 2] += v4.z; x[4 * j + 3] += v4.w; }
        }
        __syncthreads();                      // also: every thread has read its B1 row (dZT_lo aliases B1)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float hi, lo;
          split_tf32(active ? x[j] : 0.f, hi, lo);
          XT[tcanon(c0 + j, r)] = hi; B2[tcanon(c0 + j, r)] = lo;
        }
      } else {
        __syncthreads();                      // every thread has read its B1 row (dZT_lo aliases B1)
      }
      {
        float hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) split_tf32(dz[j], hi[j], lo[j]);
        if (a.dX) { tmem_put16(tl + c0, hi); tmem_put16(tl + 32 + c0, lo); }
        if (wgrad) {
#pragma unroll
          for (int j = 0; j < 16; ++j) { dZT[tcanon(c0 + j, r)] = hi[j]; B1[tcanon(c0 + j, r)] = lo[j]; }
        }
        tmem_wait_st();
      }
      fence_async_smem();
      fence_before();
      __syncthreads();
      if (warp == 0) {
        fence_after();
        if (lelect_one()) {
          if (a.dX) {
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2) {
              const uint32_t off = (uint32_t)(8 * c + 2 * s2) * 128;
              const uint64_t bhi = smem_desc(aWThi + off, 128, N * 32), blo = smem_desc(aWTlo + off, 128, N * 32);
              mma_ts(tb + 64, tb + s2 * 8, bhi, idX, (c > 0 || s2 > 0) ? 1u : 0u);
              mma_ts(tb + 64, tb + 32 + s2 * 8, bhi, idX, 1u);
              mma_ts(tb + 64, tb + s2 * 8, blo, idX, 1u);
            }
          }
          if (wgrad) {
#pragma unroll
            for (int s2 = 0; s2 < 16; ++s2)
              mma_ss(tb + 96 + c * 64, smem_desc(aZT + s2 * 288, 144, 4608), smem_desc(aXT + s2 * 288, 144, 4608), idW,
                     (!first_tile || s2 > 0) ? 1u : 0u);
          }
          commit(bar);
        }
        __syncwarp();
      }
      pending = true;
      if (LN && a.dR) {        // B0 holds dR rows (written before the barriers above)
        if (a.dR_acc) store_tile256<true>(a.dR, a.lddr, B0, t0, rows, tid);
        else store_tile256<false>(a.dR, a.lddr, B0, t0, rows, tid);
      }
    }
    first_tile = false;
    if (a.dX) {
      mbar_wait(bar, ph); ph ^= 1u; pending = false;
      fence_after();
      uint32_t d[16];
      tmem_ld16(tl + 64 + c0, d); tmem_wait_ld();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(d[j]);
      sts_half(my2, v);          // XT_lo (aliased) has been consumed: the MMAs are complete
      fence_before();
      __syncthreads();
      if (a.dX_acc) store_tile256<true>(a.dX, a.lddx, B2, t0, rows, tid);
      else store_tile256<false>(a.dX, a.lddx, B2, t0, rows, tid);
    }
  }
  if (pending) { mbar_wait(bar, ph); ph ^= 1u; pending = false; }
  fence_after();
  if (wgrad && !first_tile) {
    // D_dW region c (64 columns): lanes 0-31 hold [hi*hi | hi*lo] of row n = lane, lanes 32-63 hold [lo*hi | lo*lo].
    // The three parts are summed in shared memory first, so that the CTA sends one coalesced atomic per 128-byte line of
    // dW: every CTA of the grid flushes at about the same time, and one atomic per element per part serialises in L2
    // (measured: ~19 us per launch, against ~1 us for this form).
    float* part = dZT;                 // [N][33]; every operand tile is dead by now
    __syncthreads();
    if (warp < 2) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t d[32], e[32];
        tmem_ld32(tl + 96 + c * 64, d);
        tmem_ld32(tl + 96 + c * 64 + 32, e);
        tmem_wait_ld();
        if (warp == 0) {
#pragma unroll
          for (int k = 0; k < 32; ++k) part[(c * 32 + lane) * 33 + k] = __uint_as_float(d[k]) + __uint_as_float(e[k]);
        }
        __syncwarp();
        if (warp == 1) {
#pragma unroll
          for (int k = 0; k < 32; ++k) part[(N + c * 32 + lane) * 33 + k] = __uint_as_float(d[k]);
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < N * 32; i += BT) atomicAdd(&a.dW[i], part[(i >> 5) * 33 + (i & 31)] + part[(N + (i >> 5)) * 33 + (i & 31)]);
  }
  // column sums of this thread's rows -> shared accumulators (lane l < 16 holds column c0 + l)
  if (wgrad && a.db) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const float cs = warp_colsum16(accd[c], lane);
      if (lane < 16) atomicAdd(&sDb[c * 32 + c0 + lane], cs);
    }
  }
  if (LN) {
    const float cg = warp_colsum16(accg, lane), cb = warp_colsum16(accb, lane);
    if (lane < 16) { atomicAdd(&sDg[c0 + lane], cg); atomicAdd(&sDbe[c0 + lane], cb); }
  }
  __syncthreads();
  if (wgrad && a.db && tid < N) atomicAdd(&a.db[tid], sDb[tid]);
  if (LN && tid < 32) {
    if (a.dgamma) atomicAdd(&a.dgamma[tid], sDg[tid]);
    if (a.dbeta) atomicAdd(&a.dbeta[tid], sDbe[tid]);
  }
  fence_before();
  __syncthreads();
  if (warp == 0) { fence_after(); tmem_dealloc<COLS>(tb); }
}

// =================================================================================================
// backward, pipelined variant (K = 32, N = 32 * NCH): one persistent CTA per SM, warp-specialised
// =================================================================================================
// warps 0-3 = row group 0, warps 4-7 = row group 1 (one thread per token row, a whole 128-token tile per group, the two
// groups out of phase), warp 8 = TMA producer, warp 9 = MMA issuer.  Everything that crosses HBM moves through the TMA
// unit: a ring of input stages {dY chunk | S or activation input | X} filled ahead of the groups, outputs staged over
// the group's own, already consumed rows of its stage (dR over dY, dX of the previous tile over X) and drained by bulk
// tensor stores (bulk tensor reduce-adds when the destination accumulates).  dW comes from MN-major bf16 operands - each
// token writes 16-byte pieces of its own row, no transposition - split hi + lo (16 mantissa bits): A = [dZ_hi ; dZ_lo]
// stacked along M (64), B = [X_hi | X_lo] along N (64), one kind::f16 MMA per 16 tokens, all tiles of the CTA
// accumulating into one TMEM region per 32-row chunk of W.  dX keeps the fp32-level tf32 hi/lo product with the A
// operand in TMEM, accumulating over the chunks of a wide layer; the dX of a tile is collected one entry later, behind
// the next entry's row arithmetic.  Column sums (db, dgamma, dbeta) are warp transposes-by-shuffle, one register per lane.
//
// The work of a CTA is a sequence of entries (token tile, chunk): entry e belongs to group e & 1; a group walks the NCH
// chunks of its token tile in consecutive entries (the X operand tile is written once per token tile).  All three roles
// enumerate the same sequence and count its valid entries, which is what the ring stage and the barrier phases hang on.
constexpr int B2T = 320;
constexpr int B2_OP = 64 * LT * 2;                   // bytes per bf16 operand tile [64 mn][128 tokens]
template <int NCH>
__host__ __device__ constexpr size_t lin_tc_bwd2_smem() {     // ring: 3 stages of 3 tiles (NCH == 1), 4 of 2 tiles (wide)
  return sizeof(float) * ((NCH == 1 ? 9 : 8) * TSW + 2 * NCH * 1024 + 32 + NCH * 32 + 64) + 4 * (size_t)B2_OP + 12 * 8 + 16;
}

__device__ __forceinline__ void group_bar(int g) {
  if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
}
__device__ __forceinline__ void tma_load_tile_at(float* dst, const CUtensorMap* tm, int col0, int row0, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(smem_u32(dst)), "l"(tm), "r"(col0), "r"(row0), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* tm, int row0, const float* src, bool add) {
  if (add)
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                 :: "l"(tm), "r"(0), "r"(row0), "r"(smem_u32(src)) : "memory");
  else
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                 :: "l"(tm), "r"(0), "r"(row0), "r"(smem_u32(src)) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_drained() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_done() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// bf16 hi / lo of two neighbouring values, packed (low half = first value)
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h); lo = *reinterpret_cast<const uint32_t*>(&l);
}
// one token row (32 values) -> MN-major bf16 operand tile: rows 0-31 hi, rows 32-63 lo; element (mn, tok) at
// (mn & 7) * 2 + (mn >> 3) * 2048 + (tok & 7) * 16 + (tok >> 3) * 128 bytes   (LBO 128, SBO 2048)
__device__ __forceinline__ void sts_row_bf16mn(unsigned char* op, int r, const float* v) {
  unsigned char* base = op + (r & 7) * 16 + (r >> 3) * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_bf16x2(v[8 * q + 2 * e], v[8 * q + 2 * e + 1], h[e], l[e]);
    *reinterpret_cast<uint4*>(base + q * 2048) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(base + (4 + q) * 2048) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}
// lane l <- sum over rows [row0, row0 + 32) of column l of a 128-byte-swizzled [128][32] tile (conflict-free: the 32
// lanes read the 32 words of one row).  Replaces a 31-shuffle register transpose and needs no register array.
__device__ __forceinline__ float colsum_tile(const float* tile, int row0, int lane) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const int q = lane >> 2, w = lane & 3;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    a0 += tile[(row0 + i) * 32 + ((q ^ ((row0 + i) & 7)) << 2) + w];
    a1 += tile[(row0 + i + 1) * 32 + ((q ^ ((row0 + i + 1) & 7)) << 2) + w];
    a2 += tile[(row0 + i + 2) * 32 + ((q ^ ((row0 + i + 2) & 7)) << 2) + w];
    a3 += tile[(row0 + i + 3) * 32 + ((q ^ ((row0 + i + 3) & 7)) << 2) + w];
  }
  return (a0 + a1) + (a2 + a3);
}
__device__ __forceinline__ constexpr uint32_t idesc_bf16_mn(int M, int N) {
  return idesc_f16_mn(M, N, true, true) | (1u << 7) | (1u << 10);
}

template <bool LN, int NCH>
__global__ void __launch_bounds__(B2T, 1) lin_tc_bwd2_kernel(LinBwd a, const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmAUX,
                                                             const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDR,
                                                             const __grid_constant__ CUtensorMap tmDX) {
  static_assert(!LN || NCH == 1, "the LayerNorm variant is 32 wide");
  constexpr int N = NCH * 32;
  constexpr int COLS = NCH == 1 ? 256 : 512;   // per group: A hi|lo 64 + dX 32 (columns g*96 ..); dW of chunk c at 192 + 64 c
  constexpr int RING = (NCH == 1 ? 9 : 8) * TSW;
  extern __shared__ __align__(1024) unsigned char lin_tc_raw[];
  float* ST = reinterpret_cast<float*>(lin_tc_raw);          // input ring (plain pointer arithmetic keeps LDS / STS)
  unsigned char* OPA = reinterpret_cast<unsigned char*>(ST + RING);          // 2 x dZ operand
  unsigned char* OPB = OPA + 2 * B2_OP;                                      // 2 x X operand
  float* WThi = reinterpret_cast<float*>(OPB + 2 * B2_OP);   // [32 rows k][N] canonical: (k>>3)*(N*8) + (n>>2)*32 + (k&7)*4 + (n&3)
  float* WTlo = WThi + NCH * 1024;
  float* sG = WTlo + NCH * 1024;      // [32]
  float* sDb = sG + 32;               // [N]
  float* sDg = sDb + N; float* sDbe = sDg + 32;
  uint64_t* full = reinterpret_cast<uint64_t*>(sDbe + 32);   // [4] TMA bytes landed in stage s
  uint64_t* empty = full + 4;                                // [4] stage s may be refilled
  uint64_t* ready = empty + 4;                               // [2] operands of group g written
  uint64_t* done = ready + 2;                                // [2] MMAs of group g complete
  uint32_t* tmem_s = reinterpret_cast<uint32_t*>(done + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool has_aux = NCH == 1 && (LN || a.act != 0);
  const int nst = has_aux ? 3 : 4;                           // stages in the ring
  const int stage_floats = (has_aux ? 3 : 2) * TSW;          // {dY | S or activation input | X}  /  {dY | X}
  const int x_off = (has_aux ? 2 : 1) * TSW;

  for (int i = tid; i < N * 32; i += B2T) {
    const int n = i >> 5, k = i & 31;
    float hi, lo;
    split_tf32(a.W[i], hi, lo);
    const int o = (k >> 3) * (N * 8) + (n >> 2) * 32 + (k & 7) * 4 + (n & 3);
    WThi[o] = hi; WTlo[o] = lo;
  }
  if (tid < 32) { sG[tid] = LN ? a.gamma[tid] : 0.f; sDg[tid] = 0.f; sDbe[tid] = 0.f; }
  if (tid < N) sDb[tid] = 0.f;
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { mbar_init(&full[q], 1); mbar_init(&empty[q], 128); }
#pragma unroll
    for (int g = 0; g < 2; ++g) { mbar_init(&ready[g], 128); mbar_init(&done[g], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<COLS>(tmem_s);
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = *tmem_s;
  const int ntiles = (a.T + LT - 1) / LT;
  // entry e -> (group, chunk, token tile); -1 past the end of group 0's tiles (the sequence is over), -2 = this group has no such tile
  auto entry_tile = [&](int e, int& c) {
    const int g = e & 1, k = e >> 1;
    c = k % NCH;
    const int kk = k / NCH;
    const int t0 = blockIdx.x + (2 * kk) * gridDim.x;
    if (t0 >= ntiles) return -1;
    const int t = t0 + g * gridDim.x;
    return t < ntiles ? t : -2;
  };

  if (warp == 8) {
    // ---- TMA producer: the j-th valid entry goes to stage j % nst ----
    if (lane == 0) {
      int s = 0, u = 0;                  // stage, use count of that stage
      for (int e = 0;; ++e) {
        int c;
        const int tile = entry_tile(e, c);
        if (tile == -1) break;
        if (tile < 0) continue;
        if (u > 0) mbar_wait(&empty[s], (uint32_t)((u - 1) & 1));
        float* st = ST + s * stage_floats;
#ifdef B2_CHECK
        if (s < 3) reinterpret_cast<volatile int*>(tmem_s + 1)[s] = u;      // use index of stage s, published before the barrier is armed
        __threadfence_block();
#endif
        const bool with_x = c == 0;
        mbar_expect_tx(&full[s], ((has_aux ? 2u : 1u) + (with_x ? 1u : 0u)) * LT * 128u);
        tma_load_tile_at(st, &tmDY, c * 32, tile * LT, &full[s]);
        if (has_aux) tma_load_tile_at(st + TSW, &tmAUX, 0, tile * LT, &full[s]);
        if (with_x) tma_load_tile_at(st + x_off, &tmX, 0, tile * LT, &full[s]);
        if (++s == nst) { s = 0; ++u; }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ---- MMA issuer ----
    if (lane == 0) {
      const uint32_t idX = idesc_tf32(128, 32), idW = idesc_bf16_mn(64, 64);
      const uint32_t aWThi = smem_u32(WThi), aWTlo = smem_u32(WTlo);
      uint32_t started = 0u;             // chunks whose dW region holds a partial sum already
      int kg0 = 0, kg1 = 0;              // entries issued per group
      for (int e = 0;; ++e) {
        int c;
        const int tile = entry_tile(e, c);
        if (tile == -1) break;
        if (tile < 0) continue;
        const int g = e & 1;
        mbar_wait(&ready[g], (uint32_t)((g ? kg1 : kg0) & 1));
        if (g) ++kg1; else ++kg0;
        fence_after();
        const uint32_t tA = tb + g * 96, tD = tA + 64;
#pragma unroll
        for (int s2 = 0; s2 < 4; ++s2) {
          const uint32_t off = (uint32_t)(8 * c + 2 * s2) * 128;
          const uint64_t bhi = smem_desc(aWThi + off, 128, N * 32), blo = smem_desc(aWTlo + off, 128, N * 32);
          mma_ts(tD, tA + s2 * 8, bhi, idX, (c > 0 || s2 > 0) ? 1u : 0u);
          mma_ts(tD, tA + 32 + s2 * 8, bhi, idX, 1u);
          mma_ts(tD, tA + s2 * 8, blo, idX, 1u);
        }
        const uint32_t aA = smem_u32(OPA + g * B2_OP), aB = smem_u32(OPB + g * B2_OP);
        const bool first = !((started >> c) & 1u);
        started |= 1u << c;
#pragma unroll
        for (int t = 0; t < 8; ++t)
          mma_ss_f16(tb + 192 + c * 64, smem_desc(aA + t * 256, 128, 2048), smem_desc(aB + t * 256, 128, 2048), idW, (!first || t > 0) ? 1u : 0u);
        commit(&done[g]);
      }
    }
    __syncwarp();
  } else {
    // ---- row groups ----
    const int g = warp >> 2, r = tid & 127;
    const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16) + g * 96;
    const DropCfg dc = make_drop(LN ? a.p_drop : 0.f, a.seed, a.stream_id);
    unsigned char* opA = OPA + g * B2_OP;
    unsigned char* opB = OPB + g * B2_OP;
    const bool store_dr = LN && a.dR != nullptr;
    float acc_db[NCH], acc_dg = 0.f, acc_dbe = 0.f;        // lane l: column l, this warp's rows, all tiles
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc_db[c] = 0.f;
    int prev_row0 = -1;                                    // token tile whose dX is still in TMEM
    int k = 0;                                             // entries of this group so far
    // own entries only, state kept incrementally: the j-th valid entry of the CTA's sequence sits in stage j % nst.  Group
    // 1 follows group 0 entry by entry, except behind a last token tile that only group 0 has (its chunks are then consecutive).
    int c = 0, tile = blockIdx.x + g * gridDim.x;
    int s_cur = g, u_cur = 0;
    for (;;) {
      if (tile >= ntiles) break;
      float* st = ST + s_cur * stage_floats;
      const int row0 = tile * LT;
      const long long t = (long long)row0 + r;
      // With three stages the uses of a stage alternate between the two groups, and a group can get here while the other
      // group's load into this stage - the previous phase of full[s] - is still in flight: a parity wait for phase u is
      // then satisfied at once by the completed phase u - 2 and the rows of the tile before last are read (seen on the
      // GPU once slow accumulating stores starved the ring).  The stage's previous release, empty[s] phase u - 1, orders
      // it: that barrier can only be in phase u - 1 or u here (u - 2 was this group's own release, u needs this group's
      // arrivals), so its parity is unambiguous, and behind it full[s] is in phase u or u + 1.
      if (u_cur > 0) mbar_wait(&empty[s_cur], (uint32_t)((u_cur - 1) & 1));
      mbar_wait(&full[s_cur], (uint32_t)(u_cur & 1));
      const int wrow = (warp & 3) * 32;          // first row of this warp inside the tile
      float dz[32];
      lds_row_sw(dz, st, r);
      // Column sums (db, dgamma, dbeta) are taken from shared memory: a row lands in (or already is in) a tile slot of
      // the stage that this thread owns, and lane l adds up column l over the warp's 32 rows.
      if (LN) {
        float sv[32];
        lds_row_sw(sv, st + TSW, r);
        acc_dbe += colsum_tile(st, wrow, lane);                // dY tile as the TMA unit wrote it
#ifdef B2_CHECK
        if (t < a.T && sv[0] != (float)t && lane == 0)      // probe build: tests/probe/bwd2_acc_probe.py MARK=1 writes the row index into S[:, 0]
          printf("[b2check] cta %d warp %d tile %d (k %d, stage %d use %d, producer armed use %d): S[row][0] = %.0f, expected %lld (tile %d)\n", (int)blockIdx.x, warp, tile, k,
                 s_cur, u_cur, s_cur < 3 ? reinterpret_cast<volatile int*>(tmem_s + 1)[s_cur] : -1, sv[0], t, (int)(sv[0] / 128));
#endif
        float mean = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) mean += sv[j];
        mean *= (1.f / 32);
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { sv[j] -= mean; var = fmaf(sv[j], sv[j], var); }
        const float rstd = 1.f / sqrtf(var * (1.f / 32) + a.eps);
        {
          float tg[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) tg[j] = dz[j] * rstd * sv[j];
          sts_row_sw(st + TSW, r, tg);                         // own S row: consumed
        }
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { dz[j] *= sG[j]; m1 += dz[j]; m2 = fmaf(dz[j], sv[j], m2); }
        m1 *= (1.f / 32); m2 *= (1.f / 32) * rstd;
        __syncwarp();                                          // every lane: dY rows read (dbeta), dy * xhat rows written
        acc_dg += colsum_tile(st + TSW, wrow, lane);
#pragma unroll
        for (int j = 0; j < 32; ++j) dz[j] = rstd * (dz[j] - m1 - sv[j] * rstd * m2);      // dS = dR
        if (store_dr) sts_row_sw(st, r, dz);     // staged over this thread's own (already consumed) dY row of the stage
        if (dc.on) {
          const uint32_t rh = drop_row_hash(dc, (uint64_t)t);
#pragma unroll
          for (int j = 0; j < 32; ++j) dz[j] *= drop_mult_row(dc, rh, j);
        }
        __syncwarp();                                          // every lane: dgamma column reads done
        const float* dbsrc = st;                               // without dropout dZ = dR, which is staged already
        if (dc.on || !store_dr) { sts_row_sw(st + TSW, r, dz); dbsrc = st + TSW; }
        __syncwarp();
        acc_db[0] += colsum_tile(dbsrc, wrow, lane);
      } else if (NCH == 1 && a.act != 0) {
        float av[32];
        lds_row_sw(av, st + TSW, r);
        if (a.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) dz[j] = av[j] > 0.f ? dz[j] : 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) dz[j] *= gelu_erf_grad_fast(av[j]);
        }
        sts_row_sw(st + TSW, r, dz);                           // own row of the activation-input tile: consumed
        __syncwarp();
        acc_db[0] += colsum_tile(st + TSW, wrow, lane);
      } else {
        const float cs = colsum_tile(st, wrow, lane);          // dZ = dY: straight from the TMA-written tile
#pragma unroll
        for (int q = 0; q < NCH; ++q) if (q == c) acc_db[q] += cs;
      }
      // The group's previous entry must have left the tensor core before its operand tiles and TMEM A region are rewritten.
      // When that entry closed a token tile, its dX is collected here, behind this entry's row arithmetic, and staged
      // over this thread's own X row of the stage (c == 0 here: the stage carries an X tile).
      const bool collect = prev_row0 >= 0 && c == 0;
      {
        uint32_t d[32];
        if (k > 0) {
          mbar_wait(&done[g], (uint32_t)((k - 1) & 1));
          fence_after();
          if (collect) { tmem_ld32(tl + 64, d); tmem_wait_ld(); fence_before(); }
        }
        if (c == 0) {
          float x[32];
          lds_row_sw(x, st + x_off, r);
          sts_row_bf16mn(opB, r, x);
        }
        if (collect) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(d[j]);
          sts_row_sw(st + x_off, r, v);
        }
      }
      if (store_dr || collect) {
        fence_async_smem();
        group_bar(g);
        if (r == 0) {
          if (store_dr) tma_store_tile(&tmDR, row0, st, a.dR_acc != 0);
          if (collect) tma_store_tile(&tmDX, prev_row0, st + x_off, a.dX_acc != 0);
        }
      }
      if (r != 0) mbar_arrive(&empty[s_cur]);    // this thread is done with the stage (thread 0: once its stores have drained)
      {
        float hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) split_tf32(dz[j], hi[j], lo[j]);
        tmem_put32(tl, hi); tmem_put32(tl + 32, lo);
      }
      sts_row_bf16mn(opA, r, dz);
      tmem_wait_st();
      fence_async_smem();
      fence_before();
      if (r == 0) { tma_store_drained(); mbar_arrive(&empty[s_cur]); }
      mbar_arrive(&ready[g]);
      prev_row0 = row0;
      ++k;
      // next own entry
      const bool alone = g == 0 && tile + (int)gridDim.x >= ntiles;      // no group-1 entries in between
      if (++c == NCH) { c = 0; tile += 2 * gridDim.x; }
      s_cur += alone ? 1 : 2;
      if (s_cur >= nst) { s_cur -= nst; ++u_cur; }
    }
    if (prev_row0 >= 0) {
      // last token tile of the group: staged in the group's own dZ operand tile, idle once its MMAs are complete
      mbar_wait(&done[g], (uint32_t)((k - 1) & 1));
      fence_after();
      uint32_t d[32];
      tmem_ld32(tl + 64, d); tmem_wait_ld();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(d[j]);
      float* stg = reinterpret_cast<float*>(opA);
      sts_row_sw(stg, r, v);
      fence_async_smem();
      fence_before();
      group_bar(g);
      if (r == 0) tma_store_tile(&tmDX, prev_row0, stg, a.dX_acc != 0);
    }
    if (r == 0) tma_store_done();
#pragma unroll
    for (int c = 0; c < NCH; ++c) atomicAdd(&sDb[c * 32 + lane], acc_db[c]);
    if (LN) { atomicAdd(&sDg[lane], acc_dg); atomicAdd(&sDbe[lane], acc_dbe); }
  }
  fence_before();
  __syncthreads();
  fence_after();
  if ((int)blockIdx.x < ntiles) {
    // D_dW (M = 64): row m in TMEM lane (m & 15) + 32 * (m >> 4); rows 0-31 = dZ_hi x [X_hi | X_lo], rows 32-63 = dZ_lo x [X_hi | ..].
    // The parts are summed in shared memory (the input ring is idle now) and leave as one coalesced atomic per line.
    float* part = ST;                  // [NCH][64][33]: rows 0-31 hi parts, 32-63 lo parts
    if (warp < 4) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t d[32], e2[32];
        const uint32_t tq = tb + ((uint32_t)(warp * 32) << 16) + 192 + c * 64;
        tmem_ld32(tq, d); tmem_ld32(tq + 32, e2); tmem_wait_ld();
        if (lane < 16) {
          const int m = warp * 16 + lane;
#pragma unroll
          for (int k = 0; k < 32; ++k) part[(c * 64 + m) * 33 + k] = __uint_as_float(d[k]) + (warp < 2 ? __uint_as_float(e2[k]) : 0.f);
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < N * 32; i += B2T) {
      const int c = i >> 10, n = (i >> 5) & 31, k = i & 31;
      atomicAdd(&a.dW[i], part[(c * 64 + n) * 33 + k] + part[(c * 64 + 32 + n) * 33 + k]);
    }
    if (tid < N) atomicAdd(&a.db[tid], sDb[tid]);
    if (tid < 32) {
      if (LN && a.dgamma) atomicAdd(&a.dgamma[tid], sDg[tid]);
      if (LN && a.dbeta) atomicAdd(&a.dbeta[tid], sDbe[tid]);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) { fence_after(); tmem_dealloc<COLS>(tb); }
}

// ------------------------------------------------------------------------------------------------
static bool al16(const void* p, long long ld) { return p == nullptr || ((((uintptr_t)p) & 15) == 0 && (ld & 3) == 0); }
static bool tc_off() { static const bool off = [] { const char* e = getenv("VAESNE_NO_TC_LIN"); return e && e[0] && e[0] != '0'; }(); return off; }

bool lin_tc_fwd_eligible(const LinFwd& a) {
  if (tc_off() || a.K != 32 || !(a.N == 32 || a.N == 64 || a.N == 96)) return false;
  if (a.R && (a.N != 32 || a.Xadd || a.H || a.act != 0)) return false;
  return al16(a.X, a.ldx) && al16(a.Xadd, a.ldxa) && al16(a.H, a.ldh) && al16(a.R, a.ldr) && al16(a.S, 32) && al16(a.Y, a.ldy);
}
bool lin_tc_bwd_eligible(const LinBwd& a) {
  if (tc_off() || a.K != 32 || !(a.N == 32 || a.N == 64 || a.N == 96)) return false;
  if (a.S && (a.N != 32 || a.act != 0)) return false;
  if (a.N != 32 && a.act != 0) return false;
  if (!a.dX && !a.dW) return false;
  return al16(a.dY, a.lddy) && al16(a.S, 32) && al16(a.dR, a.lddr) && al16(a.A, a.lda) && al16(a.X, a.ldx) && al16(a.Xadd, a.ldxa) && al16(a.dX, a.lddx);
}

// Persistent grid: exactly as many CTAs as can be co-resident (a partial second wave would double the run time),
// never more than there are tiles.  The occupancy of each kernel instantiation is queried once.
struct TcKernelInfo { const void* fn; int ctas_per_sm; int sms; int dev; };
template <typename K>
static int tc_prepare(K k, size_t smem, const char* what, TcKernelInfo& out, int threads = LT) {
  // keyed by (kernel, device ordinal): cudaFuncSetAttribute opt-ins are per device, and one thread may drive several GPUs
  static thread_local TcKernelInfo cache[64] = {};
  int cur = 0;
  if (cudaGetDevice(&cur) != cudaSuccess) cur = 0;
  for (auto& e : cache) if (e.fn == (const void*)k && e.dev == cur) { out = e; return V_OK; }
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("%s: cannot reserve %zu B of shared memory: %s", what, smem, cudaGetErrorString(e)); return V_ECUDA; }
  // ask for the full shared-memory carve-out, otherwise the occupancy (and residency) is computed for a small default
  (void)cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  // Residency from the kernel's own resources.  (cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for
  // kernels that allocate tensor memory, whatever they allocate; the block scheduler itself only looks at
  // registers / shared memory / threads, and tcgen05.alloc waits for free columns.)
  int dev = cur, sms = 148, smem_sm = 233472, regs_sm = 65536;
  cudaFuncAttributes fa;
  e = cudaFuncGetAttributes(&fa, k);
  if (e != cudaSuccess) { set_error("%s: cudaFuncGetAttributes failed: %s", what, cudaGetErrorString(e)); return V_ECUDA; }
  {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
  }
  const int regs_cta = ((fa.numRegs + 7) / 8) * 8 * threads;
  const int by_regs = regs_sm / (regs_cta > 0 ? regs_cta : 1);
  const int by_smem = (int)((size_t)smem_sm / (smem + fa.sharedSizeBytes + 1024));
  int occ = by_regs < by_smem ? by_regs : by_smem;
  if (occ < 1) occ = 1;
  out = TcKernelInfo{(const void*)k, occ, sms, cur};
  if (getenv("VAESNE_DEBUG")) fprintf(stderr, "[vaesne] %s: %d CTAs/SM by occupancy, %d SMs, %zu B smem\n", what, occ, sms, smem);
  for (auto& c : cache) if (!c.fn) { c = out; break; }
  return V_OK;
}
template <typename K, typename A>
static int tc_launch(K k, size_t smem, int tmem_ctas, cudaStream_t st, const char* what, const A& args, int threads = LT) {
  TcKernelInfo ki;
  int rc = tc_prepare(k, smem, what, ki, threads); if (rc) return rc;
  const int per_sm = ki.ctas_per_sm < tmem_ctas ? ki.ctas_per_sm : tmem_ctas;      // TMEM: 512 columns per SM
  const int ntiles = (args.T + LT - 1) / LT;
  const int cap = ki.sms * per_sm;
  const int grid = ntiles < cap ? ntiles : cap;
  k<<<grid, threads, smem, st>>>(args);
  return check_launch(what);
}

// ---- TMA descriptors: [T, 32] fp32 view with row stride ld, 128-row boxes, 128-byte swizzle, zero fill out of bounds ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tma_encoder() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
static int make_tile_map(CUtensorMap* tm, const float* base, long long ld, int T, const char* what, int width = 32) {
  EncodeTiledFn enc = tma_encoder();
  if (!enc) { set_error("%s: cuTensorMapEncodeTiled is not available from this driver", what); return V_ECUDA; }
  const cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)T};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)LT}, estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("%s: cuTensorMapEncodeTiled failed (%d) for base %p ld %lld T %d", what, (int)r, (const void*)base, ld, T); return V_ECUDA; }
  return V_OK;
}

template <typename K>
static int lin_tc_fwd_launch(K k, size_t smem, int tmem_ctas, cudaStream_t st, const char* what, const LinFwd& a) {
  TcKernelInfo ki;
  int rc = tc_prepare(k, smem, what, ki); if (rc) return rc;
  CUtensorMap tmX, tmR;
  rc = make_tile_map(&tmX, a.X, a.ldx, a.T, what); if (rc) return rc;
  const float* second = a.R ? a.R : a.Xadd;
  rc = make_tile_map(&tmR, second ? second : a.X, second ? (a.R ? a.ldr : a.ldxa) : a.ldx, a.T, what); if (rc) return rc;
  const int per_sm = ki.ctas_per_sm < tmem_ctas ? ki.ctas_per_sm : tmem_ctas;      // TMEM: 512 columns per SM
  const int ntiles = (a.T + LT - 1) / LT;
  const int cap = ki.sms * per_sm;
  k<<<ntiles < cap ? ntiles : cap, LT, smem, st>>>(a, tmX, tmR);
  return check_launch(what);
}

int lin_tc_fwd(const LinFwd& a, cudaStream_t st) {
  const int g = a.N <= 64 ? 4 : 2;     // TMEM: 128 columns per CTA up to N = 64, else 256
  if (a.R) return lin_tc_fwd_launch(lin_tc_fwd_kernel<1, true>, lin_tc_fwd_smem<1>(), g, st, "lin_tc_fwd_ln", a);
  if (a.N == 32) return lin_tc_fwd_launch(lin_tc_fwd_kernel<1, false>, lin_tc_fwd_smem<1>(), g, st, "lin_tc_fwd", a);
  if (a.N == 64) return lin_tc_fwd_launch(lin_tc_fwd_kernel<2, false>, lin_tc_fwd_smem<2>(), g, st, "lin_tc_fwd", a);
  return lin_tc_fwd_launch(lin_tc_fwd_kernel<3, false>, lin_tc_fwd_smem<3>(), g, st, "lin_tc_fwd", a);
}
// pipelined variant: the big N = K = 32 calls of a training step (out_proj / ffn.2 with LayerNorm, q / ffn.0 plain)
static int bwd2_mode() {     // VAESNE_LIN_BWD2: 0 = never, 1 = when eligible (default), 2 = also for small T (tests)
  static const int m = [] { const char* e = getenv("VAESNE_LIN_BWD2"); return e && e[0] ? atoi(e) : 1; }();
  return m;
}
static bool lin_tc_bwd2_eligible(const LinBwd& a) {
  if (bwd2_mode() == 0 || a.K != 32 || !(a.N == 32 || a.N == 64 || a.N == 96)) return false;
  if (a.N != 32 && (a.S || a.act != 0)) return false;
  if (!a.dX || !a.dW || !a.db || a.Xadd) return false;
  if (a.S && a.lddr != 0 && a.dR && (a.lddr & 3)) return false;
  return bwd2_mode() >= 2 || a.T >= 148 * LT * 4;
}
template <typename K>
static int lin_tc_bwd2_launch(K k, size_t smem, cudaStream_t st, const char* what, const LinBwd& a) {
  TcKernelInfo ki;
  int rc = tc_prepare(k, smem, what, ki, B2T); if (rc) return rc;
  CUtensorMap tmDY, tmAUX, tmX, tmDR, tmDX;
  rc = make_tile_map(&tmDY, a.dY, a.lddy, a.T, what, a.N); if (rc) return rc;
  tmAUX = tmDY; tmDR = tmDY;
  if (a.S) { rc = make_tile_map(&tmAUX, a.S, 32, a.T, what); if (rc) return rc; }
  else if (a.act != 0) { rc = make_tile_map(&tmAUX, a.A, a.lda, a.T, what); if (rc) return rc; }
  rc = make_tile_map(&tmX, a.X, a.ldx, a.T, what); if (rc) return rc;
  if (a.S && a.dR) { rc = make_tile_map(&tmDR, a.dR, a.lddr, a.T, what); if (rc) return rc; }
  rc = make_tile_map(&tmDX, a.dX, a.lddx, a.T, what); if (rc) return rc;
  const int ntiles = (a.T + LT - 1) / LT;
  k<<<ntiles < ki.sms ? ntiles : ki.sms, B2T, smem, st>>>(a, tmDY, tmAUX, tmX, tmDR, tmDX);
  if (getenv("VAESNE_BWD2_SYNC")) {          // diagnosis: synchronise and print the arguments of every launch with its outcome
    const cudaError_t e = cudaStreamSynchronize(st);
    fprintf(stderr, "[bwd2] %s T=%d N=%d act=%d dY=%p/%lld S=%p A=%p/%lld X=%p/%lld dR=%p/%lld acc=%d dX=%p/%lld acc=%d dW=%p db=%p dg=%p p=%g -> %s\n", what, a.T, a.N,
            a.act, (const void*)a.dY, a.lddy, (const void*)a.S, (const void*)a.A, a.lda, (const void*)a.X, a.ldx, (void*)a.dR, a.lddr, a.dR_acc,
            (void*)a.dX, a.lddx, a.dX_acc, (void*)a.dW, (void*)a.db, (void*)a.dgamma, (double)a.p_drop, cudaGetErrorString(e));
  }
  return check_launch(what);
}
static int lin_tc_bwd_one(const LinBwd& a, cudaStream_t st) {
  if (lin_tc_bwd2_eligible(a))
    return a.S ? lin_tc_bwd2_launch(lin_tc_bwd2_kernel<true, 1>, lin_tc_bwd2_smem<1>(), st, "lin_tc_bwd2_ln", a)
               : lin_tc_bwd2_launch(lin_tc_bwd2_kernel<false, 1>, lin_tc_bwd2_smem<1>(), st, "lin_tc_bwd2", a);
  const int g = 2;                                   // TMEM: 256 columns per CTA
  if (a.S) return tc_launch(lin_tc_bwd_kernel<1, true>, lin_tc_bwd_smem<1>(), g, st, "lin_tc_bwd_ln", a, BT);
  if (a.N == 32) return tc_launch(lin_tc_bwd_kernel<1, false>, lin_tc_bwd_smem<1>(), g, st, "lin_tc_bwd", a, BT);
  return tc_launch(lin_tc_bwd_kernel<2, false>, lin_tc_bwd_smem<2>(), g, st, "lin_tc_bwd", a, BT);
}
int lin_tc_bwd(const LinBwd& a, cudaStream_t st) {
  if (a.N == 32) return lin_tc_bwd_one(a, st);
  if (lin_tc_bwd2_eligible(a))
    return a.N == 64 ? lin_tc_bwd2_launch(lin_tc_bwd2_kernel<false, 2>, lin_tc_bwd2_smem<2>(), st, "lin_tc_bwd2_w", a)
                     : lin_tc_bwd2_launch(lin_tc_bwd2_kernel<false, 3>, lin_tc_bwd2_smem<3>(), st, "lin_tc_bwd2_w", a);
  if (a.N <= 64) return lin_tc_bwd_one(a, st);
  // N = 96 (packed q|k|v projection): rows 0..63 of W, then rows 64..95 accumulating into dX (TMEM holds at most
  // two 64-column dW regions next to the dX accumulator and the A operand)
  LinBwd lo = a, hi = a;
  lo.N = 64;
  hi.N = 32; hi.dY = a.dY + 64; hi.W = a.W + 64 * 32;
  if (a.dW) hi.dW = a.dW + 64 * 32;
  if (a.db) hi.db = a.db + 64;
  if (a.dX) hi.dX_acc = 1;
  int rc = lin_tc_bwd_one(lo, st); if (rc) return rc;
  return lin_tc_bwd_one(hi, st);
}

}  // namespace vaesne

