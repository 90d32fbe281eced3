/* vaesne_b200.h — C ABI of the B200-native VAESNe hot path.
 *
 * The reference (YunyiShen/VAESNe-dev) is pure PyTorch and has no FFI; this header is the
 * lower surface defined in SURVEY.md §8(b).  Each entry point names the reference code whose
 * arithmetic it replaces (paths relative to /root/reference/package/VAESNe).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the comment says "host";
 *     kernels never allocate, free or synchronise; they enqueue on `stream` (a cudaStream_t
 *     passed as void*), so the whole step can be captured into a CUDA graph;
 *   - all float tensors are fp32, row-major; `ld*` are row strides in elements;
 *   - return value 0 = ok, negative = error (see VAESNE_E*), message in vaesne_last_error();
 *   - functions are re-entrant; the only global state is the thread-local error string;
 *   - gradient outputs named d<W>, d<b>, dgamma, dbeta, dtable ACCUMULATE (atomicAdd) into
 *     caller-zeroed buffers; dX / dR style outputs overwrite unless their *_acc flag is set;
 *   - dropout: `seed` points to one uint64 on the device (so a captured graph can advance it),
 *     `stream_id` separates the independent masks of one step; p_drop == 0 disables it.
 *   - model_dim 32 / 4 heads / head_dim 8 (what every reference script uses) is the supported
 *     geometry of the fused kernels; anything else returns VAESNE_EUNSUPPORTED.
 */
#ifndef VAESNE_B200_H
#define VAESNE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAESNE_B200_ABI_VERSION 1
#define VAESNE_OK 0
#define VAESNE_EBADSHAPE (-1)
#define VAESNE_EUNSUPPORTED (-2)
#define VAESNE_EALIGN (-3)
#define VAESNE_ECUDA (-4)
#define VAESNE_ENULL (-5)

const char* vaesne_last_error(void);
int vaesne_abi_version(void);
long long vaesne_launch_count(void);   /* kernels enqueued by this library since load (process-wide) */
int vaesne_is_emulated(void);   /* 1 only for the CPU test build under tests/emu */

/* ---- token-wise linear (+activation | +dropout+residual+LayerNorm) --------------------------
 * Y[T,N] = act((X [+ Xadd]) W^T + b)                      nn.Linear / MLPs  util_layers.py:9-34
 * Y      = LayerNorm(R + dropout(X W^T + b))  when R != NULL (N must be 32)
 *                                              TransformerBlock.forward util_layers.py:292,303,307
 * act: 0 none, 1 ReLU, 2 GELU(erf) (nn.GELU, util_layers.py:277).  H (optional) receives the
 * pre-activation, S (optional, [T,32]) the pre-LayerNorm sum; both are what the backward needs. */
int vaesne_lin_fwd(const float* X, long long ldx, const float* Xadd, long long ldxa,
                   int T, int K, int N, const float* W, const float* b, int act,
                   float* H, long long ldh,
                   const float* R, long long ldr, const float* gamma, const float* beta, float eps,
                   float* S, float p_drop, const uint64_t* seed, uint32_t stream_id,
                   float* Y, long long ldy, void* stream);

/* Backward of the above.  S != NULL selects the LayerNorm path (dgamma/dbeta accumulate, dR gets
 * the residual gradient).  A is the saved pre-activation (GELU) or output (ReLU).  dW/db accumulate. */
int vaesne_lin_bwd(const float* dY, long long lddy, int T, int K, int N,
                   const float* S, const float* gamma, float eps, float* dgamma, float* dbeta,
                   float* dR, long long lddr, int dR_acc,
                   float p_drop, const uint64_t* seed, uint32_t stream_id,
                   int act, const float* A, long long lda,
                   const float* X, long long ldx, const float* Xadd, long long ldxa,
                   const float* W, float* dW, float* db,
                   float* dX, long long lddx, int dX_acc, void* stream);

/* ---- masked multi-head attention, 4 heads x head_dim 8 -------------------------------------
 * nn.MultiheadAttention core (q*sqrt(1/8), QK^T, key-padding mask as -inf, softmax, dropout(P), PV)
 * util_layers.py:289,297,301.  q/k/v/O are [N, L, ld] with head h at columns h*8..h*8+7 of the
 * pointer passed.  mask is torch.bool storage [mask_rows, mask_len]; batch row n uses row
 * n % mask_rows (K-sample / source replication, PhotometricVAE.py:191-197), key j >= mask_len is
 * never masked (the appended phase token, SpectraLayers.py:129-131).  LSE is [N,4,Lq].
 * N = 0 returns VAESNE_OK without touching the (possibly null) pointers; N <= 65535 per call (the Python binding splits
 * larger batches, rows are independent).  A row whose keys are all masked yields NaN, as the reference does.
 * Four kernel families serve the call, chosen by shape only (forward and backward always agree): tcgen05 (96 <= Lq, Lk
 * <= 1024; TF32-class second products), few keys (Lk <= 8), one CTA per row (Lq, Lk <= 255 otherwise), general. */
int vaesne_attn_fwd(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv,
                    int N, int Lq, int Lk, const unsigned char* mask, int mask_rows, int mask_len,
                    float p_drop, const uint64_t* seed, uint32_t stream_id,
                    float* O, long long ldo, float* LSE, void* stream);
int vaesne_attn_bwd(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv,
                    int N, int Lq, int Lk, const unsigned char* mask, int mask_rows, int mask_len,
                    float p_drop, const uint64_t* seed, uint32_t stream_id,
                    const float* O, long long ldo, const float* LSE, const float* dO, long long lddo,
                    float* delta_ws /* [N,4,Lq] scratch */, float* dq, long long lddq, float* dk, long long lddk,
                    float* dv, long long lddv, void* stream);

/* Key-block form (sequences beyond the 1024 tokens the tcgen05 kernels stage at once).  `flags` bit 0 marks a call as ONE
 * KEY BLOCK of a longer attention: a (row, head) whose keys are all masked inside the block then yields O = 0, LSE = -inf and
 * zero gradients instead of NaN.  Forward: call per key block, then vaesne_attn_combine (O = sum_b exp(LSE_b - LSE) O_b,
 * LSE = logsumexp_b LSE_b; O_parts / LSE_parts are HOST arrays of nparts <= 8 device pointers to contiguous [N, Lb, 32] /
 * [N, 4, Lb] blocks; the result goes to rows q0 .. q0 + Lb of O / LSE).  Backward: call per key block with the COMBINED O and
 * LSE; dk / dv of the block are exact, the dq of the blocks add up.  Served by the tcgen05 kernels only (96 <= Lq, Lk <= 1024). */
int vaesne_attn_fwd_ex(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv,
                       int N, int Lq, int Lk, const unsigned char* mask, int mask_rows, int mask_len,
                       float p_drop, const uint64_t* seed, uint32_t stream_id,
                       float* O, long long ldo, float* LSE, int flags, void* stream);
int vaesne_attn_bwd_ex(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv,
                       int N, int Lq, int Lk, const unsigned char* mask, int mask_rows, int mask_len,
                       float p_drop, const uint64_t* seed, uint32_t stream_id,
                       const float* O, long long ldo, const float* LSE, const float* dO, long long lddo,
                       float* delta_ws, float* dq, long long lddq, float* dk, long long lddk,
                       float* dv, long long lddv, int flags, void* stream);
int vaesne_attn_combine(const float* const* O_parts /* host */, const float* const* LSE_parts /* host */, int nparts, long long N, int Lb,
                        int Lq, int q0, float* O, long long ldo, float* LSE, void* stream);

/* ---- embeddings and data movement -----------------------------------------------------------
 * sincos_feat: out[t, 0:nf] = sin(x[t]*div), out[t, nf:2nf] = cos(x[t]*div)   util_layers.py:125-129,142-146
 * gather/scatter_rows: nn.Embedding(num_bands, 32) forward / weight gradient   PhotometricLayers.py:61,129
 * expand_rows(_bwd): x.unsqueeze(0).expand(copies, ...) and its sum-backward   PhotometricVAE.py:191-197
 * copy3d: strided block copy used for torch.cat along tokens                   SpectraLayers.py:59,128 */
int vaesne_sincos_feat(const float* x, long long T, const float* div, int nf, float* out, long long ld, void* stream);
int vaesne_gather_rows(const long long* idx, long long T, const float* table, int nrows, float* out, long long ld, int accumulate, void* stream);
int vaesne_scatter_rows(const long long* idx, long long T, const float* dout, long long ld, float* dtable, int nrows, void* stream);
int vaesne_expand_rows(const float* src, long long row_elems, long long Bs, int copies, float* dst, void* stream);
int vaesne_expand_rows_bwd(const float* ddst, long long row_elems, long long Bs, int copies, float* dsrc, int accumulate, void* stream);
int vaesne_copy3d(const float* src, long long sgs, long long srs, float* dst, long long dgs, long long drs,
                  long long G, long long R, long long C, int accumulate, void* stream);

/* ---- posterior heads, sampling, mixture-of-experts latent terms ------------------------------
 * bott/noise/mu/s/dbott are HOST arrays of M device pointers.
 * fwd: mu = bott[:, :T], s = softplus(bott[:, T:])   PhotometricVAE.py:53-54 / SpectraVAE.py:48-49
 *      z[m,k,b] = rsample(mu_m, s_m; noise_m[k,b])   PhotometricVAE.py:162-163
 *      lat[r,k,b] = sum log p(z_r) - (logsumexp_m sum log q_m(z_r) - log M)   losses.py:53-54
 *      pi[r,k,b,m] = softmax_m of the expert log-densities (saved for the backward)
 * bwd: gradients of everything above wrt the bottleneck tokens, plus an optional closed-form
 *      KL(q||p) term with coefficient kl_coef (losses.py:21, torch/distributions/kl.py). */
int vaesne_latent_fwd(int M, int K, int B, int T, int Z, const float* const* bott, const float* const* noise,
                      const int* fam_post /* host */, int fam_prior, const float* pz_mu, const float* pz_s,
                      float* z, float* const* mu, float* const* s, float* lat, float* pi, void* stream);
int vaesne_latent_bwd(int M, int K, int B, int T, int Z, const float* const* bott, const float* const* noise,
                      const int* fam_post /* host */, int fam_prior, const float* pz_mu, const float* pz_s,
                      const float* dz, const float* dlat, const float* pi,
                      const float* const* dmu_ext, const float* const* ds_ext, float kl_coef,
                      float* const* dbott, void* stream);
int vaesne_kl_fwd(const float* mu, const float* s, int fam, const float* pz_mu, const float* pz_s, int B, int TZ, float* kld, void* stream);

/* ---- likelihood + objective reductions --------------------------------------------------------
 * loglik: lpx[r,b] (+)= scaling * sum_l log p(x[b,l] | loc[r,b,l], 1 or scale_masked)   losses.py:20,55-57
 * iwae_combine: lw = lat + lpx, obj = sum_b(logsumexp_r lw - log R), w = softmax_r lw    losses.py:60-62,93
 * elbo_combine: obj = mean_{k,b} lpx - mean_b kld                                        losses.py:24 */
int vaesne_loglik_fwd(const float* loc, const float* x, const unsigned char* mask, int R, int B, int L, int fam,
                      float scale_masked, float scaling, float* lpx, int accumulate, void* stream);
int vaesne_loglik_bwd(const float* loc, const float* x, const unsigned char* mask, int R, int B, int L, int fam,
                      float scale_masked, float scaling, const float* coef, float gscale, const float* gptr /* device scalar, nullable */,
                      float* dloc, void* stream);
/* gradient of coef * (*gptr) * KL(q||p) wrt (mu, s): closed forms of torch/distributions/kl.py:330-338,468-471 */
int vaesne_kl_bwd(const float* mu, const float* s, int fam, const float* pz_mu, const float* pz_s, int B, int TZ,
                  float coef, const float* gptr, float* dmu, float* ds, void* stream);
/* dst = mult * (*gptr) * src — applies the upstream scalar gradient without a host read-back */
int vaesne_scale(const float* src, long long n, float mult, const float* gptr, float* dst, void* stream);
int vaesne_iwae_combine(const float* lat, const float* lpx, int R, int B, float* w, float* lw, float* obj, void* stream);
int vaesne_elbo_combine(const float* lpx, const float* kld, int K, int B, float* obj, void* stream);

/* ---- optimiser -----------------------------------------------------------------------------------
 * torch.optim.AdamW over one flat buffer (cannon/test_photospectra.py:135); `step` is a device int
 * advanced by vaesne_step_advance so the update is CUDA-graph replayable; grad_scale multiplies g
 * first (1/world_size for mean-type objectives under data parallelism). */
int vaesne_adamw_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                      float eps, float weight_decay, const int* step, float grad_scale, void* stream);
int vaesne_step_advance(int* step, unsigned long long* seed, void* stream);
/* Dropout seed management: *cell = lcg(*cell), *out = *cell.  nn.Dropout / MHA dropout draw from torch's
 * global Philox stream (util_layers.py:265-271,283); here every forward call takes one fresh 64-bit seed. */
int vaesne_seed_next(unsigned long long* cell, unsigned long long* out, void* stream);

/* ---- contrastive objective ------------------------------------------------------------------------
 * negInfoNCE (losses.py:98-110): l2norm = F.normalize(z, dim=-1) (inv_norm [B] is saved for the backward);
 * ce_rows: loss[i] = logsumexp_j(inv_tau * a_i . b_j) - inv_tau * a_i . b_{i + label_off}  over rows A [n,P], columns Bm [m,P]
 * (the logits matrix is never materialised; P <= 64); bwd: dA / dB (+)= w * (*gptr) * d(sum_i loss[i]);
 * sum_scale: out = scale * (sum a [+ sum b]). */
int vaesne_l2norm_fwd(const float* x, int B, int P, float eps, float* y, float* inv_norm, void* stream);
int vaesne_l2norm_bwd(const float* y, const float* inv_norm, const float* dy, int B, int P, float* dx, int accumulate, void* stream);
int vaesne_ce_rows_fwd(const float* A, int n, const float* Bm, int m, int P, float inv_tau, int label_off, float* lse, float* loss, void* stream);
int vaesne_ce_rows_bwd(const float* A, int n, const float* Bm, int m, int P, float inv_tau, int label_off, const float* lse,
                       float w, const float* gptr /* device scalar, nullable */, float* dA, int dA_acc, float* dB, int dB_acc, void* stream);
int vaesne_sum_scale(const float* a, const float* b /* nullable */, int n, float scale, float* out, void* stream);

/* ---- device-side augmentation of a resident training set --------------------------------------------
 * cannon/test_photospectra.py:45-47,75-78, cannon/ZTF_photospect.py:46-66: output row r reads source row r % B (the 10x
 * repeat); x_out = x + sigma_elem * N(0,1) per element + sigma_row * N(0,1) per output row (the per-curve time shift);
 * mask_out = mask | (U < mask_p).  x / x_out or mask / mask_out may be NULL (only the other is produced).
 * Counter-based generator keyed by (*seed, stream_id, element): independent of the launch geometry. */
int vaesne_augment(const float* x, const unsigned char* mask, long long R, long long B, int L, float sigma_elem, float sigma_row,
                   float mask_p, const unsigned long long* seed, uint32_t stream_id, float* x_out, unsigned char* mask_out, void* stream);

/* ---- probe hooks (tests/probe only; not part of the operator surface) ---------------------------
 * vaesne_debug_tc(flags): timing experiments on the tcgen05 attention key pass (1 = skip the second-product MMAs,
 * 2 = skip the exponentials; results are then meaningless).  vaesne_debug_tc_prof(out16): per-phase clocks of
 * CTA (0,0) when the library was built with VAESNE_TC_PROFILE=1. */
int vaesne_debug_tc(int flags);
int vaesne_debug_tc_prof(long long* out16 /* host */);

#ifdef __cplusplus
}
#endif
#endif /* VAESNE_B200_H */
