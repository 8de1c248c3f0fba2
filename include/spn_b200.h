/*
 * spn_b200.h -- C ABI of libspn_b200.so: hand-written sm_100a CUDA kernels for the RAT-SPN / CSPN log-space
 * circuit evaluation (forward log-likelihood, backward, ancestral sampling, entropy building blocks).
 *
 * The reference (monoclecat/Conditional-RATSPN-for-RL) has no FFI: its "kernels" are eager ATen expressions inside
 * Python modules.  Each entry point below replaces one such group of expressions; the reference file:line it
 * replaces is cited per function.  The Python modules of this repo (rat_spn.py / cspn.py / layers.py /
 * distributions.py mirrors) bind these symbols with ctypes (see INTEGRATION.md for the stub a reference
 * maintainer would add).
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise
 *   - activations / parameters are fp32, indices int32, strides are in ELEMENTS (int64) and may be 0 (broadcast)
 *   - `stream` is a cudaStream_t passed as void*; work is stream ordered; nothing synchronises, nothing allocates
 *   - return value: 0 = ok, SPN_ERR_ARG = invalid argument, <= -1000 = -(1000 + cudaError_t)
 *   - row index b runs over prod(batch dims) * w with the conditional w innermost; the weight set used by row b is
 *     0 when Wsets == 1 (RatSpn: shared weights) and b % w when Wsets == w (CSPN: one set per conditional)
 *   - weight tensors are addressed as lw[ws*sw_w + s*sw_d + i*sw_i + o*sw_o + r*sw_r]; an inner Sum layer
 *     [W, d, ic, oc, R] and the root [W, 1, R*ic, C, 1] (channel = r*ic + i) are the same kernel with other strides
 */
#ifndef SPN_B200_H
#define SPN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPN_OK 0
#define SPN_ERR_ARG (-1)
#define SPN_ERR_UNSUPPORTED (-2)

/* Library / build identification.  Returns e.g. 100 for sm_100a and writes a version string (host pointer). */
int spn_version(char* buf, int buflen);

/* ---------------------------------------------------------------------------------------------------------------
 * Leaf layer: per-repetition feature permutation + Gaussian log-density + tanh change-of-variables correction +
 * NaN marginalisation + product over `card` consecutive features, in one pass.
 * Replaces rat_spn.py:155-170 (apply_permutation), distributions.py:199-246 (RatNormal.forward),
 * distributions.py:366-382 (pad + IndependentMultivariate.forward) and layers.py:276-312 (Product.forward).
 *   x     [B, F, 1|I, 1|R]  via strides (sx_i / sx_r = 0 when the dim is broadcast); pre-squash values
 *   perm  int32 [F, R] or NULL (x_needs_permutation=False)
 *   mu, sigma [Wsets, F, I, R] contiguous, already bounded (distributions.py:131-154)
 *   out   [B, d0, I, R] via strides, d0 = ceil(F / card)
 * ------------------------------------------------------------------------------------------------------------- */
int spn_leaf_fwd(const float* x, int64_t sx_b, int64_t sx_f, int64_t sx_i, int64_t sx_r, const int32_t* perm,
                 const float* mu, const float* sigma, int Wsets, int w, int B, int F, int I, int R, int card,
                 int d0, int tanh_corr, float* out, int64_t so_b, int64_t so_d, int64_t so_c, int64_t so_r,
                 void* stream);

/* Backward of spn_leaf_fwd w.r.t. the bounded parameters (autograd of the expressions above).
 *   g [B, d0, I, R] via strides; dmu, dsigma [Wsets, F, I, R] contiguous, fully overwritten. */
int spn_leaf_bwd_params(const float* g, int64_t sg_b, int64_t sg_d, int64_t sg_c, int64_t sg_r, const float* x,
                        int64_t sx_b, int64_t sx_f, int64_t sx_i, int64_t sx_r, const int32_t* perm,
                        const float* mu, const float* sigma, int Wsets, int w, int B, int F, int I, int R,
                        int card, int d0, float* dmu, float* dsigma, void* stream);

/* Backward of spn_leaf_fwd w.r.t. x (needed for reparameterised samples in the entropy estimator).
 *   dx has the shape/strides of x (broadcast dims are reduced); it is zeroed and accumulated into. */
int spn_leaf_bwd_x(const float* g, int64_t sg_b, int64_t sg_d, int64_t sg_c, int64_t sg_r, const float* x,
                   int64_t sx_b, int64_t sx_f, int64_t sx_i, int64_t sx_r, const int32_t* perm, const float* mu,
                   const float* sigma, int Wsets, int w, int B, int F, int I, int R, int card, int d0,
                   int tanh_corr, float* dx, int64_t dx_numel, void* stream);

/* Shared-parameter (RatSpn, one weight set) variants of spn_leaf_fwd / spn_leaf_bwd_params with layout-aware thread
 * mapping.  mu / sigma / dmu / dsigma are addressed as p[f*sp_f + i*sp_i + r*sp_r] (F*I*R elements); i_fast = 1 decodes
 * the thread index with the leaf channel fastest (pass it with the internal [scope][rep][row][channel] activation layout
 * and parameters stored [F][R][I]); i_fast = 0 is the reference order (r fastest, parameters [F][I][R]). */
int spn_leaf_shared_fwd(const float* x, int64_t sx_b, int64_t sx_f, int64_t sx_i, int64_t sx_r, const int32_t* perm,
                        const float* mu, const float* sigma, int64_t sp_f, int64_t sp_i, int64_t sp_r, int i_fast, int B,
                        int F, int I, int R, int card, int d0, int tanh_corr, float* out, int64_t so_b, int64_t so_d,
                        int64_t so_c, int64_t so_r, void* stream);
int spn_leaf_shared_bwd_params(const float* g, int64_t sg_b, int64_t sg_d, int64_t sg_c, int64_t sg_r, const float* x,
                               int64_t sx_b, int64_t sx_f, int64_t sx_i, int64_t sx_r, const int32_t* perm,
                               const float* mu, const float* sigma, int64_t sp_f, int64_t sp_i, int64_t sp_r, int i_fast,
                               int B, int F, int I, int R, int card, int d0, float* dmu, float* dsigma, void* stream);

/* Register-tiled variants for shared parameters and plain inputs x [B, F] (no channel / repetition dims): the fp32-pipe
 * bound path used by RatSpn.forward.  spn_leaf_transpose fills the caller-allocated workspace `xT`
 * (spn_leaf_workspace_floats(B, F) floats, 16-byte aligned) with the feature-major, zero-padded copy of x followed by
 * one NaN flag per 64 rows; both kernels gather whole row segments from it.  Parameter / output / gradient strides as
 * in spn_leaf_shared_*. */
int64_t spn_leaf_workspace_floats(int B, int F);
int spn_leaf_transpose(const float* x, int64_t sx_b, int64_t sx_f, int B, int F, float* xT, void* stream);
int spn_leaf_tiled_supported(int I, int R, int d0);
int spn_leaf_tiled_fwd(const float* xT, const int32_t* perm, const float* mu, const float* sigma, int64_t sp_f,
                       int64_t sp_i, int64_t sp_r, int B, int F, int I, int R, int card, int d0, int tanh_corr, float* out,
                       int64_t so_b, int64_t so_d, int64_t so_c, int64_t so_r, void* stream);
int spn_leaf_tiled_bwd_params(const float* xT, const float* g, int64_t sg_b, int64_t sg_d, int64_t sg_c, int64_t sg_r,
                              const int32_t* perm, const float* mu, const float* sigma, int64_t sp_f, int64_t sp_i,
                              int64_t sp_r, int B, int F, int I, int R, int card, int d0, float* dmu, float* dsigma,
                              void* stream);

/* Tensor-core variants (tcgen05 kind::tf32, leaf_tc.cu) of the two calls above for the internal layout: the leaf + product
 * over a scope as one GEMM per (scope, repetition) on operands split into two tf32 words each (x^2, x, valid against
 * -1/(2 sigma^2), mu/sigma^2, the constant term), fp32 accumulation; the parameter gradients as the transposed GEMM over
 * the rows.  out / g are contiguous [d0][R][B][I]; xT as written by spn_leaf_transpose; NaN inputs are marginalised; no
 * tanh correction (distributions.py:220 stays on the fp32 kernels).  `img`: caller-allocated scratch of
 * spn_leaf_tc_image_bytes bytes, 128-byte aligned (the parameter operand image, rebuilt by every forward call).
 * Supported: I a multiple of 8 up to 64, card <= 25, every scope non-empty.  Replaces distributions.py:199-246 +
 * layers.py:276-312 on the shared-parameter path (rat_spn.py:155-170 for the permutation). */
int spn_leaf_tc_supported(int I, int R, int card, int d0);
int64_t spn_leaf_tc_image_bytes(int I, int R, int card, int d0);
int spn_leaf_tc_fwd(const float* xT, const int32_t* perm, const float* mu, const float* sigma, int64_t sp_f, int64_t sp_i,
                    int64_t sp_r, int B, int F, int I, int R, int card, int d0, void* img, float* out, void* stream);
int spn_leaf_tc_bwd_params(const float* xT, const float* g, const int32_t* perm, const float* mu, const float* sigma,
                           int64_t sp_f, int64_t sp_i, int64_t sp_r, int B, int F, int I, int R, int card, int d0,
                           float* dmu, float* dsigma, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Sum layer, optionally fused with the CrossProduct layer below it (the c*c outer sum is never materialised).
 * Replaces layers.py:419-440 + :545-563 (CrossProduct.forward / split_shuffled_scopes), layers.py:110-128
 * (Sum.forward: broadcast add + logsumexp) and, with reduce_r = 1, rat_spn.py:230-236 (root merge over repetitions).
 *   xprod = 1: x [B, d_in, c, R]; scope s of the sum pairs input scopes (2s, 2s+1); scopes >= d_in are zero padding;
 *              in-channel i = l*c + r'.            xprod = 0: x [B, d_out, ic, R] is the sum input itself.
 *   out [B, d_out, oc, R]  (reduce_r = 1: [B, d_out, oc], logsumexp additionally runs over r; so_r ignored)
 * ------------------------------------------------------------------------------------------------------------- */
int spn_sum_fwd(int xprod, const float* x, int64_t sx_b, int64_t sx_d, int64_t sx_c, int64_t sx_r, int d_in, int c,
                const float* lw, int64_t sw_w, int64_t sw_d, int64_t sw_i, int64_t sw_o, int64_t sw_r, int Wsets,
                int w, int B, int d_out, int ic, int oc, int R, int reduce_r, float* out, int64_t so_b,
                int64_t so_d, int64_t so_o, int64_t so_r, void* stream);

/* Backward of spn_sum_fwd w.r.t. the log-weights: dlw[ws, s, i, o, r] = sum_b g * exp(x_i + lw - out)
 * (per conditional when Wsets == w).  dlw uses the strides of lw and is fully overwritten. */
int spn_sum_bwd_w(int xprod, const float* x, int64_t sx_b, int64_t sx_d, int64_t sx_c, int64_t sx_r, int d_in,
                  int c, const float* lw, int64_t sw_w, int64_t sw_d, int64_t sw_i, int64_t sw_o, int64_t sw_r,
                  int Wsets, int w, int B, int d_out, int ic, int oc, int R, int reduce_r, const float* out,
                  const float* g, int64_t so_b, int64_t so_d, int64_t so_o, int64_t so_r, float* dlw,
                  int64_t dlw_numel, void* stream);

/* Backward of spn_sum_fwd w.r.t. x.  dx [B, d_in, c, R] (xprod) or [B, d_out, ic, R], addressed through its own
 * strides (sdx_*); every element is written exactly once. */
int spn_sum_bwd_x(int xprod, const float* x, int64_t sx_b, int64_t sx_d, int64_t sx_c, int64_t sx_r, int d_in,
                  int c, const float* lw, int64_t sw_w, int64_t sw_d, int64_t sw_i, int64_t sw_o, int64_t sw_r,
                  int Wsets, int w, int B, int d_out, int ic, int oc, int R, int reduce_r, const float* out,
                  const float* g, int64_t so_b, int64_t so_d, int64_t so_o, int64_t so_r, float* dx, int64_t sdx_b,
                  int64_t sdx_d, int64_t sdx_c, int64_t sdx_r, void* stream);

/* Per-sample-weight (CSPN, Wsets == w) fused CrossProduct + Sum forward, HBM streaming: the contiguous weight slab
 * [ic, oc, R] of every (row, scope) is staged through shared memory by the TMA engine (cp.async.bulk) while one thread per
 * output accumulates.  Same semantics as spn_sum_fwd(xprod = 1, reduce_r = 0) for x [B, d_in, c, R] with c and R
 * contiguous (sx_b, sx_d in elements), lw [w, d_out, c*c, oc, R] contiguous, out [B, d_out, oc, R] contiguous. */
int spn_sum_ps_supported(int c, int oc, int R);
int spn_sum_ps_fwd(const float* x, int64_t sx_b, int64_t sx_d, int d_in, int c, const float* lw, int w, int B, int d_out,
                   int oc, int R, float* out, void* stream);

/* Fused backward of spn_sum_ps_fwd for ONE row per weight set (B == w, the CSPN log-likelihood training step): every
 * slab is read once and its gradient written once (TMA ring in, staged TMA bulk store out); dx of both children comes
 * out of the same pass.  dlw [w, d_out, c*c, oc, R] contiguous (fully overwritten); dx [B, d_in, c, R] with c, R
 * contiguous (sdx_b, sdx_d in elements; every scope < d_in is written).  Autograd of layers.py:110-128 + :419-440. */
int spn_sum_ps_bwd_supported(int c, int oc, int R);
int spn_sum_ps_bwd(const float* x, int64_t sx_b, int64_t sx_d, int d_in, int c, const float* lw, int w, int B, int d_out,
                   int oc, int R, const float* out, const float* g, float* dlw, float* dx, int64_t sdx_b,
                   int64_t sdx_d, void* stream);

/* The same pair for UNNORMALISED per-conditional weights (the raw head outputs of the gating network): the
 * F.log_softmax over the in-channels of cspn.py:274, :282 is fused into both directions.
 *   forward : out = lse_i(x_i + raw_i) - lse_i(raw_i); logz [B, d_out, oc, R] receives lse_i(raw_i)
 *   backward: draw[i] = g * (exp(x_i + raw_i - logz - out) - exp(raw_i - logz)); dx as above
 * One row per weight set (B == w). */
int spn_sum_ps_fwd_raw(const float* x, int64_t sx_b, int64_t sx_d, int d_in, int c, const float* raw, int w, int B,
                       int d_out, int oc, int R, float* out, float* logz, void* stream);
int spn_sum_ps_bwd_raw(const float* x, int64_t sx_b, int64_t sx_d, int d_in, int c, const float* raw, int w, int B,
                       int d_out, int oc, int R, const float* logz, const float* out, const float* g, float* draw,
                       float* dx, int64_t sdx_b, int64_t sdx_d, void* stream);

/* Stand-alone CrossProduct (materialising) and its backward: layers.py:419-440.  x [B, d_in, c, R] contiguous,
 * out [B, ceil_pow2(d_in)/2, c*c, R] contiguous; dx [B, d_in, c, R] contiguous. */
int spn_xprod_fwd(const float* x, int B, int d_in, int c, int R, float* out, void* stream);
int spn_xprod_bwd(const float* g, int B, int d_in, int c, int R, float* dx, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Top-down ancestral sampling, one Sum layer per call (the CrossProduct index un-ravelling above it is fused).
 * Replaces layers.py:130-237 (Sum.sample: weight gather, evidence re-weighting, Gumbel-max / MPE argmax) and
 * layers.py:442-511 (CrossProduct.sample).  Items are (q, s, rs): q in [0, Q) = (node, *n, w) flattened with w
 * innermost, s a scope of this layer, rs in [0, Rs) (Rs = 1 once a repetition is selected, else R).
 *   pa      int32 parent choice via strides (sp_q, sp_s, sp_r): unravel = 0 -> out-channel o = pa[q, s, rs];
 *           unravel = 1 -> pa holds the channel chosen over oc*oc pairs for scope s/2, o = pa / oc or pa % oc
 *   rep     int32 [Q] selected repetition or NULL (then r = rs)
 *   xc      NULL, or activations of the layer below from a forward pass on the evidence [Qe, xc_din, c, R] via
 *           strides; the posterior logit offset is xc[qe, 2s, l, r] + xc[qe, 2s+1, r', r], qe = q % Qe
 *   root    1: this is the root Sum over R*ic_top channels (d = 1, Rs = 1); the repetition is i / ic_top
 *   mpe     flags: bit 0 = MPE (argmax, no noise); bit 1 = kernel-variant selector (one repetition per lane group
 *           instead of four when every repetition is sampled with in-kernel noise; same indices, used for A/B tests)
 *   gumbel  [Q, d, ic, Rs] contiguous noise, or NULL with mpe = 1 (argmax) or mpe = 0 (in-kernel Philox, keyed by
 *           seed and the global element index q_offset-shifted so results do not depend on the GPU count);
 *           seed_dev != NULL: the seed is read from device memory instead (one uint64), so that a captured CUDA
 *           graph draws fresh noise on every replay when the caller refreshes that word inside the graph)
 *   out     int32 [Q, d, Rs] chosen in-channel (first maximal index)
 * ------------------------------------------------------------------------------------------------------------- */
int spn_sample_sum(const float* lw, int64_t sw_w, int64_t sw_d, int64_t sw_i, int64_t sw_o, int64_t sw_r, int Wsets,
                   int w, int Q, int d, int ic, int oc, int Rs, const int32_t* pa, int64_t sp_q, int64_t sp_s,
                   int64_t sp_r, int unravel, const int32_t* rep, const float* xc, int64_t sxc_b, int64_t sxc_d,
                   int64_t sxc_c, int64_t sxc_r, int xc_din, int c, int Qe, int root, int ic_top,
                   const float* gumbel, int mpe, uint64_t seed, uint64_t q_offset, const uint64_t* seed_dev,
                   int32_t* out, void* stream);

/* Leaf step of sampling fused with the post-processing: channel fan-out to features, (mu, sigma) gather, Gaussian
 * draw, inverse permutation of the selected repetition, evidence fill-in, clamp+tanh.
 * Replaces layers.py:343-362, distributions.py:248-302 + :384-393 and rat_spn.py:538-580.
 *   pa int32 via strides: channel chosen over I*I pairs for leaf scope s/2 (un-ravelled here)
 *   eps [Q, F, Rs] contiguous standard-normal noise indexed by *slot* (pre inverse permutation), NULL with mpe=1,
 *       or NULL with mpe=0 for in-kernel Philox + Box-Muller
 *   inv_perm int32 [F, R] or NULL (no inverse permutation); evidence via strides or NULL (NaN = sample it)
 *   out [Q, F, Rs] contiguous; ch_out int32 [Q, F, Rs] (leaf channel per slot) or NULL
 * ------------------------------------------------------------------------------------------------------------- */
int spn_sample_leaf(const float* mu, const float* sigma, int Wsets, int w, int Q, int F, int I, int R, int card,
                    int Rs, const int32_t* pa, int64_t sp_q, int64_t sp_s, int64_t sp_r, const int32_t* rep,
                    const float* eps, int mpe, uint64_t seed, uint64_t q_offset, const uint64_t* seed_dev,
                    const int32_t* inv_perm,
                    const float* evidence, int64_t se_q, int64_t se_f, int64_t se_r, int Qe, int squash,
                    float* out, int32_t* ch_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Backward of the differentiable ('onehot') sampling mode: straight-through Gumbel-softmax estimator, tau = 1.
 * Replaces autograd through layers.py:181-186, :206-234 (hard - soft.detach() + soft), :495-503 (CrossProduct one-hot
 * split), distributions.py:281-301 (one-hot selection of (mu, sigma)) and rat_spn.py:437-468 (root / repetition
 * one-hots).  The forward values are the ones of spn_sample_sum / spn_sample_leaf (same noise: the supplied tensors
 * or the same Philox seed); the backward runs bottom-up, one call per layer:
 *   spn_sample_leaf_bwd: gy [Q, F, Rs] = dL/d(out of spn_sample_leaf) (perm/evidence/squash as in that call; perm is
 *     the FORWARD permutation [F, R], y the forward output) -> dmu, dsigma [Wsets, F, I, R] (accumulated, atomics)
 *     and g0 [Q, dP, I, Rs] = dL/d(leaf-channel one-hot of every Product scope), dP = ceil(F / card).
 *   spn_sample_sum_bwd: gc [Q, gc_d, c, Rs] = gradient w.r.t. the one-hots of the gc_d child scopes (at the root it
 *     applies to the selected repetition only) -> dlw (same strides as lw, accumulated), dxc (same strides as xc,
 *     accumulated, NULL iff xc NULL), gparent [Q, d, oc, Rs] = gradient w.r.t. the parent one-hots (NULL at the top).
 *   D = 1 (the root sits directly on the leaf CrossProduct, no Sum layer clears the unselected repetitions): pass
 *     g0_allreps = gc_allreps = 1; g0 / gc are then [Q, dP, I, R], evaluated with every repetition's leaf parameters.
 * All other arguments as in the forward calls.  Accumulating outputs must be zero-initialised by the caller.
 * ------------------------------------------------------------------------------------------------------------- */
int spn_sample_sum_bwd(const float* lw, int64_t sw_w, int64_t sw_d, int64_t sw_i, int64_t sw_o, int64_t sw_r,
                       int Wsets, int w, int Q, int d, int ic, int oc, int Rs, const int32_t* pa, int64_t sp_q,
                       int64_t sp_s, int64_t sp_r, int unravel, const int32_t* rep, const float* xc, int64_t sxc_b,
                       int64_t sxc_d, int64_t sxc_c, int64_t sxc_r, int xc_din, int c, int Qe, int root, int ic_top,
                       const float* gumbel, uint64_t seed, uint64_t q_offset, const uint64_t* seed_dev, const float* gc,
                       int gc_d, int gc_allreps,
                       float* dlw, float* dxc, float* gparent, void* stream);
int spn_sample_leaf_bwd(const float* mu, const float* sigma, int Wsets, int w, int Q, int F, int I, int R, int card,
                        int Rs, int dP, const int32_t* ch, const int32_t* rep, const float* eps, int mpe, uint64_t seed,
                        uint64_t q_offset, const uint64_t* seed_dev, const int32_t* perm, const float* evidence, int64_t se_q,
                        int64_t se_f,
                        int64_t se_r, int Qe, int squash, const float* y, int g0_allreps, const float* gy, float* dmu,
                        float* dsigma, float* g0, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Tensor-core path (tcgen05 + TMEM) for layers whose weights are shared by MANY rows: fused CrossProduct + Sum as bf16
 * GEMMs with fp32 accumulation and fp32 contraction epilogues; replaces the same reference lines as
 * spn_sum_fwd(xprod = 1) (layers.py:90-128, :419-440) for
 *   - RatSpn models: nsets = 1, every row uses the one weight set;
 *   - CSPN layers with many rows per conditional (recursive entropy, rat_spn.py:625-750; cfg 5): nsets = number of
 *     conditionals, Bs rows per set.
 * Channel counts: c (per child) and oc multiples of 4 up to 64 (spn_sum_tc_supported); internally both are padded to
 * P = spn_sum_tc_padded_channels(c, oc) = max(c, oc) rounded up to a multiple of 8.
 * Activations use the internal layout [scope][repetition][row][channel] with the rows of one weight set contiguous:
 *   x [d_in, R, nsets*Bs, c] contiguous, out [d_out, R, nsets*Bs, oc] contiguous.
 * Weights lw [nsets, d_out, c*c, oc, R] (in-channel k = l*c + r'); nsd = nsets * d_out below.
 * spn_sum_tc_prepare turns them into two bf16 operand images (spn_sum_tc_image_bytes() bytes, caller-allocated) of exp(lw):
 * the k-major core-matrix image used by dX and its chunk-major permutation used by the forward; the kernels copy them
 * into shared memory with the TMA engine (resident per work item up to 128 KB, streamed in chunks above).
 * Weight normalisation (layers.py:85, F.log_softmax over the in-channels on every read) is fused: when `logz`
 * ([nsd, oc, R], caller-allocated) is non-NULL, `lw` is the UNNORMALISED parameter tensor, spn_sum_tc_prepare writes
 * logz = logsumexp_k lw and every consumer uses lw - logz; with logz == NULL, lw must already be normalised.
 * Outputs whose linear-domain sum underflows are recomputed exactly in fp32 from lw (logsumexp semantics are kept).
 * A protocol fault inside the kernels (timed-out mbarrier wait) traps the launch: the next CUDA call reports an error.
 * ------------------------------------------------------------------------------------------------------------- */
int spn_sum_tc_supported(int c, int oc, int R);
int spn_sum_tc_padded_channels(int c, int oc);
int64_t spn_sum_tc_image_bytes(int nsd, int c, int oc, int R);
int spn_sum_tc_prepare(const float* lw, int nsd, int c, int oc, int R, float* logz, void* wimg, void* stream);
/* Forward.  `fhand` (optional, spn_sum_tc_fhand_bytes() bytes): receives, per (set, scope, repetition, 128-row tile), the
 * packed bf16 row vectors exp(xL - mL), exp(xR - mR) and the shifts mL, mR; the backward kernels then load no activations
 * and exponentiate nothing. */
int64_t spn_sum_tc_fhand_bytes(int Bs, int nsd, int c, int oc, int R);
int spn_sum_tc_fwd(const float* x, const void* wimg, const float* lw, const float* logz, int Bs, int nsets, int d_in,
                   int d_out, int c, int oc, int R, float* out, void* fhand, void* stream);
/* Backward of spn_sum_tc_fwd (layouts as in the forward).
 *   spn_sum_tc_bwd_x: dx [d_in, R, nsets*Bs, c] fully overwritten: dE = G W^T as one small-K GEMM per row tile, contracted
 *                     with exp(xL - mL) / exp(xR - mR) straight out of tensor memory.  `fhand` (required): the forward's
 *                     hand-off (`x` itself is not read and may be NULL).  `hand` (required, spn_sum_tc_hand_bytes() bytes):
 *                     receives, per (set, scope, repetition, 128-row tile), G = g * exp(mL + mR - out) as a packed bf16 tile
 *                     (written by a G pre-pass, read back by TMA as this kernel's A operand and by spn_sum_tc_bwd_w as its
 *                     B operand).  `csum` (optional, [nsd, oc, R] floats): receives the column sums of g that the fused
 *                     softmax backward of spn_sum_tc_bwd_w needs (pass the tail of its workspace and csum_ready = 1 there).
 *   spn_sum_tc_bwd_w: needs `fhand` (from spn_sum_tc_fwd) and `hand` (from spn_sum_tc_bwd_x) of the same (x, out, g), same
 *                     stream; it recomputes nothing.
 *                     dlw [nsets, d_out, c*c, oc, R] = gradient w.r.t. the tensor passed as lw, fully overwritten: the
 *                     log-weights (logz == NULL) or the unnormalised parameters (logz from spn_sum_tc_prepare; the
 *                     softmax backward is fused); dwt_workspace: caller-provided scratch of
 *                     spn_sum_tc_dw_workspace_floats() floats. */
int64_t spn_sum_tc_hand_bytes(int Bs, int nsd, int c, int oc, int R);
int64_t spn_sum_tc_dw_workspace_floats(int nsd, int c, int oc, int R);
int spn_sum_tc_bwd_x(const float* x, const float* out, const float* g, const void* wimg, const void* fhand, int Bs, int nsets,
                     int d_in, int d_out, int c, int oc, int R, float* dx, void* hand, float* csum, void* stream);
int spn_sum_tc_bwd_w(const void* fhand, const void* hand, const float* out, const float* g, const float* lw, const float* logz,
                     int Bs, int nsets, int d_in, int d_out, int c, int oc, int R, float* dwt_workspace, int csum_ready,
                     float* dlw, void* stream);
/* ---------------------------------------------------------------------------------------------------------------
 * Root sum layer for SHARED weights on the internal layout: merges the repetitions into the channel axis and sums over
 * all of them, fused with the last CrossProduct.  Replaces rat_spn.py:230-236 + layers.py:110-128 for a RatSpn.
 *   x    [d_in (1 or 2), R, B, c] contiguous (output of the last spn_sum_tc_fwd)
 *   lw   root log-weights [1, 1, R*c*c, C, 1] contiguous (channel = r*c*c + l*c + r'), normalised over dim 2
 *   part caller workspace of R * B * C floats; out [B, C]
 * spn_root_shared_bwd: g = dL/dout [B, C]; dx [d_in, R, B, c] and / or dlw [R*c*c*C] (either may be NULL), overwritten.
 * ------------------------------------------------------------------------------------------------------------- */
int spn_root_shared_supported(int c, int C, int R);
int spn_root_shared_fwd(const float* x, int d_in, const float* lw, int B, int c, int C, int R, float* part, float* out,
                        void* stream);
int spn_root_shared_bwd(const float* x, int d_in, const float* lw, const float* out, const float* g, int B, int c, int C,
                        int R, float* dx, float* dlw, void* stream);

/* Debug aid (synchronises): number of CTAs that gave up on an mbarrier wait since the library was loaded. */
int spn_debug_tc_aborts(int* out_host);
/* Debug aid (synchronises): in builds with SPN_TC_PROFILE=1 warp 0 of CTA 0 of the spn_sum_tc_bwd_x kernel accumulates clock64
 * deltas per epilogue phase (see tests/_tc_time.py for the slot names); copies the 12 counters
 * to out_host (host pointer) and optionally resets them (all zero in the product build). */
int spn_debug_tc_prof(unsigned long long* out_host, int reset);

/* ---------------------------------------------------------------------------------------------------------------
 * Sum-level combination of the recursive entropy approximation (rat_spn.py:661-746 with weigh_tensors :591-623):
 *   H[n, o, r] = sum_k exp(lw[k,o,r]) * (-lw[k,o,r] + ce[k,r] + lwd[k,o,r] + own[k,r] - ll[k,o,r])
 * (weight entropy + weighted child entropies + weighted log-responsibilities; lwd = lw in value, the responsibility's
 * `log_weights` term that the reference detaches unless aux_with_grad).  lw, ll [n, ic, oc, R]; ce, own [n, ic, R];
 * H [n, oc, R], or [n, oc] with reduce_r = 1 (root: also summed over the repetitions).  All contiguous fp32.
 * Backward: gH like H -> dlw, dll like lw; dce, down like ce (any output pointer may be NULL).
 * ------------------------------------------------------------------------------------------------------------- */
int spn_entropy_combine_fwd(const float* lw, const float* ce, const float* own, const float* ll, int64_t n, int ic, int oc,
                            int R, int reduce_r, float* H, void* stream);
int spn_entropy_combine_bwd(const float* lw, const float* ce, const float* own, const float* ll, const float* gH, int64_t n,
                            int ic, int oc, int R, int reduce_r, int lwd_has_grad, float* dlw, float* dce, float* down,
                            float* dll, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Adam update over flat fp32 buffers (one launch for the whole model): torch.optim.Adam semantics without amsgrad /
 * weight decay, as used by mnist_experiments/mnist_gen_train.py:599, :669, :720.  step_size = lr / (1 - beta1^t), inv_sqrt_bc2 = 1 / sqrt(1 -
 * beta2^t); all pointers 16-byte aligned.
 * ------------------------------------------------------------------------------------------------------------- */
int spn_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float step_size, float beta1, float beta2,
                  float eps, float inv_sqrt_bc2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPN_B200_H */
