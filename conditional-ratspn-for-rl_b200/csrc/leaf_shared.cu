// Leaf kernels specialised for SHARED parameters (RatSpn: one weight set for the whole batch).
// Same arithmetic as leaf.cu, but the thread mapping follows the activation layout so that every global access of a
// warp is either contiguous or a broadcast:
//   * the (channel i, repetition r) pair is decoded from the fastest thread index in the order of the OUTPUT layout
//     (i fastest for the internal [scope][rep][row][channel] layout, r fastest for the reference layout), and the
//     parameters are passed with matching strides, so parameter loads and output / gradient accesses are coalesced;
//   * x[b, perm[f, r]] only depends on (b, r): with i fastest a warp reads at most two distinct addresses per load;
//   * forward: one thread owns TB rows (parameter loads + the SFU work are amortised);
//     backward: one thread owns FCH features of a scope (the upstream gradient g[b, s, i, r] is loaded once per row
//     for all of them) and walks a slice of the rows; slices are combined with one atomicAdd per feature.
// Reference semantics: rat_spn.py:155-170, distributions.py:199-246, :366-382, layers.py:276-312.
#include "common.cuh"

struct LeafSharedArgs {
    const float* x; int64_t sx_b, sx_f, sx_i, sx_r;
    const int32_t* perm;
    const float* mu; const float* sigma; int64_t sp_f, sp_i, sp_r;
    int B, F, I, R, card, d0, tanh_corr;
    float* out; const float* g; int64_t so_b, so_d, so_c, so_r;
    float* dmu; float* dsigma;
    int nsplit, nfch;
};

template <int TB, bool I_FAST>
__global__ void __launch_bounds__(256) leaf_shared_fwd_kernel(LeafSharedArgs a) {
    const int64_t IR = (int64_t)a.I * a.R;
    const int64_t nbt = (a.B + TB - 1) / TB;
    const int64_t total = nbt * a.d0 * IR;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int j = (int)(tid % IR);
    const int s = (int)((tid / IR) % a.d0);
    const int64_t b0 = (tid / (IR * a.d0)) * TB;
    const int i = I_FAST ? j % a.I : j / a.R;
    const int r = I_FAST ? j / a.I : j % a.R;

    float acc[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) acc[t] = 0.f;
    const int f_begin = s * a.card;
    const int f_end = min(a.F, f_begin + a.card);
    const float* xb = a.x + b0 * a.sx_b + (int64_t)i * a.sx_i + (int64_t)r * a.sx_r;
    const int nrow = (int)min((int64_t)TB, a.B - b0);
    for (int f = f_begin; f < f_end; ++f) {
        const int src = a.perm ? a.perm[(int64_t)f * a.R + r] : f;
        const int64_t p = (int64_t)f * a.sp_f + (int64_t)i * a.sp_i + (int64_t)r * a.sp_r;
        const float mu_v = a.mu[p];
        const float sg = a.sigma[p];
        const float inv2v = 1.f / (2.f * sg * sg);
        const float c_v = logf(sg) + SPN_LOG_SQRT_2PI;
        const float* xp = xb + (int64_t)src * a.sx_f;
        if (nrow == TB) {
            float xv[TB];
#pragma unroll
            for (int t = 0; t < TB; ++t) xv[t] = __ldg(xp + t * a.sx_b);
#pragma unroll
            for (int t = 0; t < TB; ++t) {
                const float d = xv[t] - mu_v;
                float lp = -(d * d) * inv2v - c_v;
                if (a.tanh_corr) lp -= spn_tanh_correction(xv[t]);
                acc[t] += isnan(xv[t]) ? 0.f : lp;
            }
        } else {
            for (int t = 0; t < nrow; ++t) {
                const float xv = __ldg(xp + t * a.sx_b);
                const float d = xv - mu_v;
                float lp = -(d * d) * inv2v - c_v;
                if (a.tanh_corr) lp -= spn_tanh_correction(xv);
                acc[t] += isnan(xv) ? 0.f : lp;
            }
        }
    }
    float* op = a.out + b0 * a.so_b + (int64_t)s * a.so_d + (int64_t)i * a.so_c + (int64_t)r * a.so_r;
#pragma unroll
    for (int t = 0; t < TB; ++t)
        if (t < nrow) op[t * a.so_b] = acc[t];
}

template <int FCH, bool I_FAST>
__global__ void __launch_bounds__(256) leaf_shared_bwd_kernel(LeafSharedArgs a) {
    const int64_t IR = (int64_t)a.I * a.R;
    const int64_t total = (int64_t)a.d0 * a.nfch * IR;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int j = (int)(tid % IR);
    const int fch = (int)((tid / IR) % a.nfch);
    const int s = (int)(tid / (IR * a.nfch));
    const int i = I_FAST ? j % a.I : j / a.R;
    const int r = I_FAST ? j / a.I : j % a.R;
    const int f0 = s * a.card + fch * FCH;
    const int f_end = min(a.F, min((s + 1) * a.card, f0 + FCH));
    if (f0 >= f_end) return;

    int64_t xoff[FCH];
    float mu_v[FCH], a0[FCH], a1[FCH], a2[FCH];
#pragma unroll
    for (int t = 0; t < FCH; ++t) {
        const int f = min(f0 + t, f_end - 1);
        const int src = a.perm ? a.perm[(int64_t)f * a.R + r] : f;
        xoff[t] = (int64_t)src * a.sx_f + (int64_t)i * a.sx_i + (int64_t)r * a.sx_r;
        mu_v[t] = a.mu[(int64_t)f * a.sp_f + (int64_t)i * a.sp_i + (int64_t)r * a.sp_r];
        a0[t] = a1[t] = a2[t] = 0.f;
    }
    const float* gp = a.g + (int64_t)s * a.so_d + (int64_t)i * a.so_c + (int64_t)r * a.so_r;
    const int64_t rows_per = (a.B + a.nsplit - 1) / a.nsplit;
    const int64_t b_begin = (int64_t)blockIdx.y * rows_per;
    const int64_t b_end = min((int64_t)a.B, b_begin + rows_per);
#pragma unroll 2
    for (int64_t b = b_begin; b < b_end; ++b) {
        const float gv = gp[b * a.so_b];
        const float* xr = a.x + b * a.sx_b;
#pragma unroll
        for (int t = 0; t < FCH; ++t) {
            const float xv = __ldg(xr + xoff[t]);
            const bool ok = !isnan(xv);
            const float d = ok ? xv - mu_v[t] : 0.f;
            const float t1 = gv * d;
            a0[t] += ok ? gv : 0.f;
            a1[t] += t1;
            a2[t] = fmaf(t1, d, a2[t]);
        }
    }
#pragma unroll
    for (int t = 0; t < FCH; ++t) {
        const int f = f0 + t;
        if (f < f_end) {
            const int64_t p = (int64_t)f * a.sp_f + (int64_t)i * a.sp_i + (int64_t)r * a.sp_r;
            const float inv_sg = 1.f / a.sigma[p];
            const float inv_var = inv_sg * inv_sg;
            const float dmu = a1[t] * inv_var;
            const float dsg = a2[t] * inv_var * inv_sg - a0[t] * inv_sg;
            if (a.nsplit == 1) {
                a.dmu[p] = dmu;
                a.dsigma[p] = dsg;
            } else {
                atomicAdd(a.dmu + p, dmu);
                atomicAdd(a.dsigma + p, dsg);
            }
        }
    }
}

extern "C" int spn_leaf_shared_fwd(const float* x, int64_t sx_b, int64_t sx_f, int64_t sx_i, int64_t sx_r,
                                   const int32_t* perm, const float* mu, const float* sigma, int64_t sp_f, int64_t sp_i,
                                   int64_t sp_r, int i_fast, int B, int F, int I, int R, int card, int d0, int tanh_corr,
                                   float* out, int64_t so_b, int64_t so_d, int64_t so_c, int64_t so_r, void* stream) {
    if (!x || !mu || !sigma || !out || B < 0 || F <= 0 || I <= 0 || R <= 0 || card <= 0 || d0 <= 0) return SPN_ERR_ARG;
    if (B == 0) return SPN_OK;
    LeafSharedArgs a{};
    a.x = x; a.sx_b = sx_b; a.sx_f = sx_f; a.sx_i = sx_i; a.sx_r = sx_r; a.perm = perm; a.mu = mu; a.sigma = sigma;
    a.sp_f = sp_f; a.sp_i = sp_i; a.sp_r = sp_r; a.B = B; a.F = F; a.I = I; a.R = R; a.card = card; a.d0 = d0;
    a.tanh_corr = tanh_corr; a.out = out; a.so_b = so_b; a.so_d = so_d; a.so_c = so_c; a.so_r = so_r;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t IR = (int64_t)I * R;
    if (B >= 64) {
        const int64_t total = ((int64_t)(B + 15) / 16) * d0 * IR;
        if (i_fast) leaf_shared_fwd_kernel<16, true><<<spn_blocks(total, 256), 256, 0, st>>>(a);
        else leaf_shared_fwd_kernel<16, false><<<spn_blocks(total, 256), 256, 0, st>>>(a);
    } else {
        const int64_t total = ((int64_t)(B + 3) / 4) * d0 * IR;
        if (i_fast) leaf_shared_fwd_kernel<4, true><<<spn_blocks(total, 256), 256, 0, st>>>(a);
        else leaf_shared_fwd_kernel<4, false><<<spn_blocks(total, 256), 256, 0, st>>>(a);
    }
    return spn_launch_status();
}

extern "C" int spn_leaf_shared_bwd_params(const float* g, int64_t sg_b, int64_t sg_d, int64_t sg_c, int64_t sg_r,
                                          const float* x, int64_t sx_b, int64_t sx_f, int64_t sx_i, int64_t sx_r,
                                          const int32_t* perm, const float* mu, const float* sigma, int64_t sp_f,
                                          int64_t sp_i, int64_t sp_r, int i_fast, int B, int F, int I, int R, int card,
                                          int d0, float* dmu, float* dsigma, void* stream) {
    if (!g || !x || !mu || !sigma || !dmu || !dsigma || B < 0 || F <= 0 || I <= 0 || R <= 0 || card <= 0) return SPN_ERR_ARG;
    constexpr int FCH = 8;
    LeafSharedArgs a{};
    a.x = x; a.sx_b = sx_b; a.sx_f = sx_f; a.sx_i = sx_i; a.sx_r = sx_r; a.perm = perm; a.mu = mu; a.sigma = sigma;
    a.sp_f = sp_f; a.sp_i = sp_i; a.sp_r = sp_r; a.B = B; a.F = F; a.I = I; a.R = R; a.card = card; a.d0 = d0;
    a.g = g; a.so_b = sg_b; a.so_d = sg_d; a.so_c = sg_c; a.so_r = sg_r; a.dmu = dmu; a.dsigma = dsigma;
    a.nfch = (card + FCH - 1) / FCH;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nparam = (int64_t)F * I * R;
    const int64_t total = (int64_t)d0 * a.nfch * I * R;
    int nsplit = (int)((300000 + total - 1) / total);
    if (nsplit > B / 64) nsplit = B / 64;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > 65535) nsplit = 65535;
    a.nsplit = nsplit;
    if (nsplit > 1) {
        cudaError_t e = cudaMemsetAsync(dmu, 0, nparam * sizeof(float), st);
        if (e != cudaSuccess) return spn_cuda_status(e);
        e = cudaMemsetAsync(dsigma, 0, nparam * sizeof(float), st);
        if (e != cudaSuccess) return spn_cuda_status(e);
    }
    if (B == 0) {
        if (nsplit == 1) {
            cudaMemsetAsync(dmu, 0, nparam * sizeof(float), st);
            cudaMemsetAsync(dsigma, 0, nparam * sizeof(float), st);
        }
        return SPN_OK;
    }
    dim3 grid(spn_blocks(total, 256), nsplit);
    if (i_fast) leaf_shared_bwd_kernel<FCH, true><<<grid, 256, 0, st>>>(a);
    else leaf_shared_bwd_kernel<FCH, false><<<grid, 256, 0, st>>>(a);
    return spn_launch_status();
}
