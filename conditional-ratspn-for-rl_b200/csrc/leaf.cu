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

// ---------------------------------------------------------------------------------------------------------------
// Per-conditional parameters (CSPN: Wsets == w), R % 4 == 0, contiguous
// [.., I, R] outputs: every parameter is read by exactly one row per n-sample, so both directions are HBM streams.  Each
// thread handles FOUR consecutive repetitions of one (row, scope / feature, channel): 16-byte loads of mu / sigma /
// permutation / gradient and 16-byte stores, four times the bytes in flight of the scalar kernels above.
__global__ void __launch_bounds__(256) leaf_ps_fwd_kernel(LeafArgs a) {
    const int IR4 = a.I * a.R / 4;
    const int64_t total = (int64_t)a.B * a.d0 * IR4;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int j = (int)(tid % IR4) * 4;
    const int s = (int)((tid / IR4) % a.d0);
    const int64_t b = tid / ((int64_t)IR4 * a.d0);
    const int i = j / a.R, r0 = j % a.R;
    const int64_t pbase = (((int64_t)(b % a.w)) * a.F * a.I + i) * a.R + r0;
    const float* xb = a.x + b * a.sx_b;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int f_begin = s * a.card;
    const int f_end = min(a.F, f_begin + a.card);
#pragma unroll 4
    for (int f = f_begin; f < f_end; ++f) {
        int src[4] = {f, f, f, f};
        if (a.perm) {
            const int4 p4 = __ldg(reinterpret_cast<const int4*>(a.perm + (int64_t)f * a.R + r0));
            src[0] = p4.x; src[1] = p4.y; src[2] = p4.z; src[3] = p4.w;
        }
        const int64_t p = pbase + (int64_t)f * a.I * a.R;
        const float4 m4 = __ldg(reinterpret_cast<const float4*>(a.mu + p));
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(a.sigma + p));
        const float mu_v[4] = {m4.x, m4.y, m4.z, m4.w}, sg_v[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float xv = xb[(int64_t)src[q] * a.sx_f + (int64_t)i * a.sx_i + (int64_t)(r0 + q) * a.sx_r];
            if (!isnan(xv)) {
                const float d = xv - mu_v[q];
                float lp = -(d * d) * (1.f / (2.f * sg_v[q] * sg_v[q])) - (logf(sg_v[q]) + SPN_LOG_SQRT_2PI);
                if (a.tanh_corr) lp -= spn_tanh_correction(xv);
                acc[q] += lp;
            }
        }
    }
    *reinterpret_cast<float4*>(a.out + b * a.so_b + (int64_t)s * a.so_d + (int64_t)i * a.so_c + r0) =
        make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// thread = (weight set, feature, channel, 4 repetitions); walks the n rows that use this weight set
__global__ void __launch_bounds__(256) leaf_ps_bwd_param_kernel(LeafArgs a) {
    const int IR4 = a.I * a.R / 4;
    const int64_t total = (int64_t)a.Wsets * a.F * IR4;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int j = (int)(tid % IR4) * 4;
    const int f = (int)((tid / IR4) % a.F);
    const int ws = (int)(tid / ((int64_t)IR4 * a.F));
    const int i = j / a.R, r0 = j % a.R;
    const int s = f / a.card;
    int src[4] = {f, f, f, f};
    if (a.perm) {
        const int4 p4 = __ldg(reinterpret_cast<const int4*>(a.perm + (int64_t)f * a.R + r0));
        src[0] = p4.x; src[1] = p4.y; src[2] = p4.z; src[3] = p4.w;
    }
    const int64_t p = (((int64_t)ws * a.F + f) * a.I + i) * a.R + r0;
    const float4 m4 = __ldg(reinterpret_cast<const float4*>(a.mu + p));
    const float4 s4 = __ldg(reinterpret_cast<const float4*>(a.sigma + p));
    const float mu_v[4] = {m4.x, m4.y, m4.z, m4.w}, sg_v[4] = {s4.x, s4.y, s4.z, s4.w};
    float inv_sg[4], inv_var[4], dmu[4] = {0.f, 0.f, 0.f, 0.f}, dsg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) { inv_sg[q] = 1.f / sg_v[q]; inv_var[q] = inv_sg[q] * inv_sg[q]; }
    const int64_t goff = (int64_t)s * a.so_d + (int64_t)i * a.so_c + r0;
    for (int64_t b = ws; b < a.B; b += a.w) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.g + b * a.so_b + goff));
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
        const float* xb = a.x + b * a.sx_b;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float xv = xb[(int64_t)src[q] * a.sx_f + (int64_t)i * a.sx_i + (int64_t)(r0 + q) * a.sx_r];
            if (!isnan(xv)) {
                const float d = xv - mu_v[q];
                dmu[q] += gv[q] * d * inv_var[q];
                dsg[q] += gv[q] * (d * d * inv_var[q] * inv_sg[q] - inv_sg[q]);
            }
        }
    }
    *reinterpret_cast<float4*>(a.dmu + p) = make_float4(dmu[0], dmu[1], dmu[2], dmu[3]);
    *reinterpret_cast<float4*>(a.dsigma + p) = make_float4(dsg[0], dsg[1], dsg[2], dsg[3]);
}

static bool leaf_ps_vec_ok(const LeafArgs& a) {
    return a.Wsets != 1 && a.Wsets == a.w && a.R % 4 == 0 && a.so_r == 1 &&
           a.so_c % 4 == 0 && a.so_d % 4 == 0 && a.so_b % 4 == 0;
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
    if (leaf_ps_vec_ok(a) && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(mu) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(sigma) & 15) == 0) {
        const int64_t total = (int64_t)B * d0 * (IR / 4);
        leaf_ps_fwd_kernel<<<spn_blocks(total, 256), 256, 0, st>>>(a);
    } else if (Wsets == 1 && B >= 8) {
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
    if (leaf_ps_vec_ok(a) && total / 4 >= 296 * 256 && rows <= 64 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(mu) & 15) == 0 && (reinterpret_cast<uintptr_t>(sigma) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(dmu) & 15) == 0 && (reinterpret_cast<uintptr_t>(dsigma) & 15) == 0) {
        leaf_ps_bwd_param_kernel<<<spn_blocks(total / 4, 256), 256, 0, st>>>(a);
        return spn_launch_status();
    }
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
