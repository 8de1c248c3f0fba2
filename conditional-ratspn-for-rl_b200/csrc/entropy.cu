// Sum-level combination step of the recursive entropy approximation (rat_spn.py:661-746, :591-623):
//     H[n, o, r] = sum_k W[k,o,r] * ( -lw[k,o,r] + ce[k,r] + lwd[k,o,r] + own[k,r] - ll[k,o,r] ),    W = exp(lw)
// = weight entropy + weighted child entropies + weighted log-responsibilities, where
//   lw  [n, ic, oc, R]  log-weights of the layer (n = conditionals x scopes),
//   ce  [n, ic, R]      entropies of the children,
//   own [n, ic, R]      log-likelihood of child k's own samples under child k (diagonal of the cross log-likelihoods),
//   ll  [n, ic, oc, R]  log-likelihood of child k's samples under every node o of this layer,
//   lwd = lw in value; it is the `log_weights` term of the responsibility, which the reference detaches unless
//   aux_with_grad (rat_spn.py:689-695), so that it only changes the weight gradient.
// reduce_r (root): the sum also runs over the repetitions (weights viewed per repetition, rat_spn.py:612-620) -> H [n, o].
// One fused kernel per direction replaces ~12 (forward) / ~25 (backward) element-wise and reduction launches per level.
#include "common.cuh"

__global__ void __launch_bounds__(256) entropy_combine_fwd_kernel(const float* __restrict__ lw, const float* __restrict__ ce,
                                                                  const float* __restrict__ own, const float* __restrict__ ll,
                                                                  int64_t n, int ic, int oc, int R, int reduce_r,
                                                                  float* __restrict__ H) {
    const int Rt = reduce_r ? 1 : R;
    const int64_t total = n * oc * Rt;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int r0 = (int)(tid % Rt);
    const int o = (int)((tid / Rt) % oc);
    const int64_t i = tid / ((int64_t)Rt * oc);
    const int r_lo = reduce_r ? 0 : r0, r_hi = reduce_r ? R : r0 + 1;
    float acc = 0.f;
    for (int k = 0; k < ic; ++k) {
        for (int r = r_lo; r < r_hi; ++r) {
            const int64_t e = ((i * ic + k) * oc + o) * R + r;
            const int64_t c = (i * ic + k) * R + r;
            const float w = lw[e];
            // (-lw + lwd) cancels in value; exp(-inf) * (finite - (-inf)) would be 0 * inf: a zero weight contributes 0
            const float t = ce[c] + own[c] - ll[e];
            acc += w == -CUDART_INF_F ? 0.f : __expf(w) * t;
        }
    }
    H[tid] = acc;
}

// thread per (i, k, r): loops over the outputs; writes dlw / dll for its (k, :, r) row and the two child gradients
__global__ void __launch_bounds__(256) entropy_combine_bwd_kernel(const float* __restrict__ lw, const float* __restrict__ ce,
                                                                  const float* __restrict__ own, const float* __restrict__ ll,
                                                                  const float* __restrict__ gH, int64_t n, int ic, int oc, int R,
                                                                  int reduce_r, int lwd_has_grad, float* __restrict__ dlw,
                                                                  float* __restrict__ dce, float* __restrict__ down,
                                                                  float* __restrict__ dll) {
    const int64_t total = n * ic * R;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int r = (int)(tid % R);
    const int k = (int)((tid / R) % ic);
    const int64_t i = tid / ((int64_t)R * ic);
    const float base = ce[tid] + own[tid];
    const float sub = lwd_has_grad ? 0.f : 1.f;
    float acc = 0.f;
    for (int o = 0; o < oc; ++o) {
        const int64_t e = ((i * ic + k) * oc + o) * R + r;
        const float g = reduce_r ? gH[i * oc + o] : gH[(i * oc + o) * R + r];
        const float w = lw[e];
        const float gw = w == -CUDART_INF_F ? 0.f : g * __expf(w);
        const float t = base - ll[e];
        if (dlw) dlw[e] = w == -CUDART_INF_F ? 0.f : gw * (t - sub);
        if (dll) dll[e] = -gw;
        acc += gw;
    }
    if (dce) dce[tid] = acc;
    if (down) down[tid] = acc;
}

extern "C" int spn_entropy_combine_fwd(const float* lw, const float* ce, const float* own, const float* ll, int64_t n, int ic,
                                       int oc, int R, int reduce_r, float* H, void* stream) {
    if (!lw || !ce || !own || !ll || !H || n < 0 || ic <= 0 || oc <= 0 || R <= 0) return SPN_ERR_ARG;
    if (n == 0) return SPN_OK;
    const int64_t total = n * oc * (reduce_r ? 1 : R);
    entropy_combine_fwd_kernel<<<spn_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(lw, ce, own, ll, n, ic, oc, R, reduce_r, H);
    return spn_launch_status();
}

extern "C" int spn_entropy_combine_bwd(const float* lw, const float* ce, const float* own, const float* ll, const float* gH,
                                       int64_t n, int ic, int oc, int R, int reduce_r, int lwd_has_grad, float* dlw,
                                       float* dce, float* down, float* dll, void* stream) {
    if (!lw || !ce || !own || !ll || !gH || n < 0 || ic <= 0 || oc <= 0 || R <= 0) return SPN_ERR_ARG;
    if (n == 0) return SPN_OK;
    const int64_t total = n * ic * R;
    entropy_combine_bwd_kernel<<<spn_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(lw, ce, own, ll, gH, n, ic, oc, R, reduce_r,
                                                                                         lwd_has_grad, dlw, dce, down, dll);
    return spn_launch_status();
}
