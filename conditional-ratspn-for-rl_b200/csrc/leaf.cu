// Leaf layer kernels: permutation gather + Gaussian log-density + tanh correction + NaN marginalisation +
// product over `card` consecutive features (forward), and the matching backward kernels.
// Reference semantics: rat_spn.py:155-170, distributions.py:199-246, :366-382, layers.py:276-312.
#include "common.cuh"

struct LeafArgs {
    const float* x; int64_t sx_b, sx_f, sx_i, sx_r;
    const int32_t* perm;
    const float* mu; const float* sigma;
    int Wsets, w, B, F, I, R, card, d0, tanh_corr;
    float* out; int64_t so_b, so_d, so_c, so_r;          // forward output / backward grad-in (same strides role)
    const float* g;
    float* dmu; float* dsigma; float* dx;
    int nsplit;
};

// One thread owns (row tile of TB rows, scope s, leaf channel i, repetition r); j = i*R + r is the fastest index so
// that parameter loads are coalesced.  TB > 1 amortises the parameter loads / SFU work over rows that share weights.
template <int TB>
__global__ void __launch_bounds__(256) leaf_fwd_kernel(LeafArgs a) {
    const int64_t IR = (int64_t)a.I * a.R;
    const int64_t nbt = (a.B + TB - 1) / TB;
    const int64_t total = nbt * a.d0 * IR;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int j = (int)(tid % IR);
    const int s = (int)((tid / IR) % a.d0);
    const int64_t b0 = (tid / (IR * a.d0)) * TB;
    const int i = j / a.R, r = j % a.R;

    float acc[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) acc[t] = 0.f;

    const int f_begin = s * a.card;
    const int f_end = min(a.F, f_begin + a.card);
#pragma unroll 4
    for (int f = f_begin; f < f_end; ++f) {
        const int src = a.perm ? a.perm[(int64_t)f * a.R + r] : f;
        const int64_t xoff = (int64_t)src * a.sx_f + (int64_t)i * a.sx_i + (int64_t)r * a.sx_r;
        float mu_v, inv2v, c_v;
        if (TB > 1 || a.Wsets == 1) {
            const int64_t p = ((int64_t)f * a.I + i) * a.R + r;
            mu_v = a.mu[p];
            const float sg = a.sigma[p];
            inv2v = 1.f / (2.f * sg * sg);
            c_v = logf(sg) + SPN_LOG_SQRT_2PI;
        }
#pragma unroll
        for (int t = 0; t < TB; ++t) {
            const int64_t b = b0 + t;
            if (b < a.B) {
                if (TB == 1 && a.Wsets != 1) {
                    const int64_t p = ((((int64_t)(b % a.w)) * a.F + f) * a.I + i) * a.R + r;
                    mu_v = a.mu[p];
                    const float sg = a.sigma[p];
                    inv2v = 1.f / (2.f * sg * sg);
                    c_v = logf(sg) + SPN_LOG_SQRT_2PI;
                }
                const float xv = a.x[b * a.sx_b + xoff];
                if (!isnan(xv)) {
                    const float d = xv - mu_v;
                    float lp = -(d * d) * inv2v - c_v;
                    if (a.tanh_corr) lp -= spn_tanh_correction(xv);
                    acc[t] += lp;
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) {
        const int64_t b = b0 + t;
        if (b < a.B) a.out[b * a.so_b + (int64_t)s * a.so_d + (int64_t)i * a.so_c + (int64_t)r * a.so_r] = acc[t];
    }
}

// One thread owns one parameter element (ws, f, i, r) and walks the rows that use it.  gridDim.y splits the rows
// (atomics only when nsplit > 1).
__global__ void __launch_bounds__(256) leaf_bwd_param_kernel(LeafArgs a) {
    const int64_t IR = (int64_t)a.I * a.R;
    const int64_t total = (int64_t)a.Wsets * a.F * IR;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int j = (int)(tid % IR);
    const int f = (int)((tid / IR) % a.F);
    const int ws = (int)(tid / (IR * a.F));
    const int i = j / a.R, r = j % a.R;
    const int s = f / a.card;
    const int src = a.perm ? a.perm[(int64_t)f * a.R + r] : f;
    const float mu_v = a.mu[tid];
    const float sg = a.sigma[tid];
    const float inv_sg = 1.f / sg;
    const float inv_var = inv_sg * inv_sg;
    const int64_t xoff = (int64_t)src * a.sx_f + (int64_t)i * a.sx_i + (int64_t)r * a.sx_r;
    const int64_t goff = (int64_t)s * a.so_d + (int64_t)i * a.so_c + (int64_t)r * a.so_r;
    const int64_t step = a.Wsets == 1 ? 1 : a.w;
    const int64_t first = a.Wsets == 1 ? 0 : ws;
    float dmu = 0.f, dsg = 0.f;
    for (int64_t b = first + (int64_t)blockIdx.y * step; b < a.B; b += step * a.nsplit) {
        const float xv = a.x[b * a.sx_b + xoff];
        if (!isnan(xv)) {
            const float gv = a.g[b * a.so_b + goff];
            const float d = xv - mu_v;
            dmu += gv * d * inv_var;
            dsg += gv * (d * d * inv_var * inv_sg - inv_sg);
        }
    }
    if (a.nsplit == 1) {
        a.dmu[tid] = dmu;
        a.dsigma[tid] = dsg;
    } else {
        atomicAdd(a.dmu + tid, dmu);
        atomicAdd(a.dsigma + tid, dsg);
    }
}

// One thread per (row, feature slot, i, r) evaluation; accumulates into dx (broadcast dims collapse via atomics).
__global__ void __launch_bounds__(256) leaf_bwd_x_kernel(LeafArgs a) {
    const int64_t IR = (int64_t)a.I * a.R;
    const int64_t total = (int64_t)a.B * a.F * IR;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int j = (int)(tid % IR);
    const int f = (int)((tid / IR) % a.F);
    const int64_t b = tid / (IR * a.F);
    const int i = j / a.R, r = j % a.R;
    const int s = f / a.card;
    const int src = a.perm ? a.perm[(int64_t)f * a.R + r] : f;
    const int64_t xo = b * a.sx_b + (int64_t)src * a.sx_f + (int64_t)i * a.sx_i + (int64_t)r * a.sx_r;
    const float xv = a.x[xo];
    if (isnan(xv)) return;
    const int ws = a.Wsets == 1 ? 0 : (int)(b % a.w);
    const int64_t p = (((int64_t)ws * a.F + f) * a.I + i) * a.R + r;
    const float sg = a.sigma[p];
    const float gv = a.g[b * a.so_b + (int64_t)s * a.so_d + (int64_t)i * a.so_c + (int64_t)r * a.so_r];
    float v = -(xv - a.mu[p]) / (sg * sg);
    if (a.tanh_corr) v += 2.f * tanhf(xv);
    atomicAdd(a.dx + xo, gv * v);
}

extern "C" int spn_leaf_fwd(const float* x, int64_t sx_b, int64_t sx_f, int64_t sx_i, int64_t sx_r,
                            const int32_t* perm, const float* mu, const float* sigma, int Wsets, int w, int B, int F,
                            int I, int R, int card, int d0, int tanh_corr, float* out, int64_t so_b, int64_t so_d,
                            int64_t so_c, int64_t so_r, void* stream) {
    if (!x || !mu || !sigma || !out || B < 0 || F <= 0 || I <= 0 || R <= 0 || card <= 0 || d0 <= 0) return SPN_ERR_ARG;
    if (Wsets != 1 && (Wsets != w || B % w != 0)) return SPN_ERR_ARG;
    if (B == 0) return SPN_OK;
    LeafArgs a{};
    a.x = x; a.sx_b = sx_b; a.sx_f = sx_f; a.sx_i = sx_i; a.sx_r = sx_r; a.perm = perm; a.mu = mu; a.sigma = sigma;
    a.Wsets = Wsets; a.w = w; a.B = B; a.F = F; a.I = I; a.R = R; a.card = card; a.d0 = d0; a.tanh_corr = tanh_corr;
    a.out = out; a.so_b = so_b; a.so_d = so_d; a.so_c = so_c; a.so_r = so_r;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t IR = (int64_t)I * R;
    if (Wsets == 1 && B >= 8) {
        const int64_t total = ((int64_t)(B + 7) / 8) * d0 * IR;
        leaf_fwd_kernel<8><<<spn_blocks(total, 256), 256, 0, st>>>(a);
    } else {
        const int64_t total = (int64_t)B * d0 * IR;
        leaf_fwd_kernel<1><<<spn_blocks(total, 256), 256, 0, st>>>(a);
    }
    return spn_launch_status();
}

extern "C" int spn_leaf_bwd_params(const float* g, int64_t sg_b, int64_t sg_d, int64_t sg_c, int64_t sg_r,
                                   const float* x, int64_t sx_b, int64_t sx_f, int64_t sx_i, int64_t sx_r,
                                   const int32_t* perm, const float* mu, const float* sigma, int Wsets, int w, int B,
                                   int F, int I, int R, int card, int d0, float* dmu, float* dsigma, void* stream) {
    if (!g || !x || !mu || !sigma || !dmu || !dsigma || B < 0 || F <= 0 || I <= 0 || R <= 0 || card <= 0) return SPN_ERR_ARG;
    if (Wsets != 1 && (Wsets != w || B % w != 0)) return SPN_ERR_ARG;
    LeafArgs a{};
    a.x = x; a.sx_b = sx_b; a.sx_f = sx_f; a.sx_i = sx_i; a.sx_r = sx_r; a.perm = perm; a.mu = mu; a.sigma = sigma;
    a.Wsets = Wsets; a.w = w; a.B = B; a.F = F; a.I = I; a.R = R; a.card = card; a.d0 = d0;
    a.g = g; a.so_b = sg_b; a.so_d = sg_d; a.so_c = sg_c; a.so_r = sg_r; a.dmu = dmu; a.dsigma = dsigma;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)Wsets * F * I * R;
    const int64_t rows = Wsets == 1 ? B : B / w;
    // split the row loop when there are few parameters but many rows, so the grid still fills 148 SMs
    int nsplit = 1;
    const int64_t blocks = (total + 255) / 256;
    if (blocks < 296 && rows > 64) {
        nsplit = (int)min((int64_t)64, min(rows / 32 + 1, (296 + blocks - 1) / blocks));
        if (nsplit < 1) nsplit = 1;
    }
    a.nsplit = nsplit;
    if (nsplit > 1) {
        cudaError_t e = cudaMemsetAsync(dmu, 0, total * sizeof(float), st);
        if (e != cudaSuccess) return spn_cuda_status(e);
        e = cudaMemsetAsync(dsigma, 0, total * sizeof(float), st);
        if (e != cudaSuccess) return spn_cuda_status(e);
    }
    dim3 grid(spn_blocks(total, 256), nsplit);
    leaf_bwd_param_kernel<<<grid, 256, 0, st>>>(a);
    return spn_launch_status();
}

extern "C" int spn_leaf_bwd_x(const float* g, int64_t sg_b, int64_t sg_d, int64_t sg_c, int64_t sg_r, const float* x,
                              int64_t sx_b, int64_t sx_f, int64_t sx_i, int64_t sx_r, const int32_t* perm,
                              const float* mu, const float* sigma, int Wsets, int w, int B, int F, int I, int R,
                              int card, int d0, int tanh_corr, float* dx, int64_t dx_numel, void* stream) {
    if (!g || !x || !mu || !sigma || !dx || B < 0 || F <= 0 || I <= 0 || R <= 0 || card <= 0) return SPN_ERR_ARG;
    if (Wsets != 1 && (Wsets != w || B % w != 0)) return SPN_ERR_ARG;
    LeafArgs a{};
    a.x = x; a.sx_b = sx_b; a.sx_f = sx_f; a.sx_i = sx_i; a.sx_r = sx_r; a.perm = perm; a.mu = mu; a.sigma = sigma;
    a.Wsets = Wsets; a.w = w; a.B = B; a.F = F; a.I = I; a.R = R; a.card = card; a.d0 = d0; a.tanh_corr = tanh_corr;
    a.g = g; a.so_b = sg_b; a.so_d = sg_d; a.so_c = sg_c; a.so_r = sg_r; a.dx = dx;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(dx, 0, dx_numel * sizeof(float), st);
    if (e != cudaSuccess) return spn_cuda_status(e);
    if (B == 0) return SPN_OK;
    const int64_t total = (int64_t)B * F * I * R;
    leaf_bwd_x_kernel<<<spn_blocks(total, 256), 256, 0, st>>>(a);
    return spn_launch_status();
}
