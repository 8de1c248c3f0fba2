// Backward of the differentiable ('onehot') ancestral sampling pass: straight-through Gumbel-softmax gradients
// (tau = 1) pushed bottom-up through the path that the index kernels of sample.cu selected top-down.
// Reference semantics: layers.py:181-186, :206-234 (hard - soft.detach() + soft), :495-503 (CrossProduct one-hot
// split), distributions.py:281-301 (leaf parameter selection by one-hot), rat_spn.py:437-468.
//
// For one Sum node with one-hot parent p (value e_o) and in-channel one-hot y = ST(softmax(z)), z_i = sum_o lw[i,o] p_o
// + xc_i * sum_o p_o + G_i, and upstream gradient g_i = dL/dy_i:
//     dz_i        = soft_i * (g_i - sum_j soft_j g_j)
//     dL/dlw[i,o] = dz_i                      (only the selected parent column o)
//     dL/dxc_i    = dz_i
//     dL/dp_o'    = sum_i dz_i * (lw[i,o'] + xc_i)           for every o'
// and a CrossProduct above hands g_i = gl[i / c] + gr[i % c] from its two child scopes.
#include "common.cuh"

struct SampleSumBwdArgs {
    const float* lw; int64_t sw_w, sw_d, sw_i, sw_o, sw_r;
    int Wsets, w, Q, d, ic, oc, Rs;
    const int32_t* pa; int64_t sp_q, sp_s, sp_r; int unravel;
    const int32_t* rep;
    const float* xc; int64_t sxc_b, sxc_d, sxc_c, sxc_r; int xc_din, c, Qe;
    int root, ic_top;
    const float* gumbel; uint64_t seed, q_offset; const uint64_t* seed_dev;
    int gc_allreps;              // root directly above the leaves (D = 1): gc is [Q, gc_d, c, R], every repetition
    const float* gc; int gc_d;   // upstream gradient w.r.t. the one-hots of the gc_d child scopes [Q, gc_d, c, Rs]
    float* dlw; float* dxc; float* gparent;  // gparent [Q, d, oc, Rs] or null
};

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int sh = G / 2; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
    return v;
}
template <int G>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int sh = G / 2; sh > 0; sh >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, sh));
    return v;
}

// One group of G lanes (G = 4..32, the power of two covering ic) per item (q, s, rs).  Dynamic shared memory: ic floats
// per group (dz), only for non-root layers.  Out-of-range groups run on the last item with their writes masked.
template <int G>
__global__ void __launch_bounds__(256) sample_sum_bwd_kernel(SampleSumBwdArgs a) {
    extern __shared__ float dz_all[];
    const int lane = threadIdx.x & (G - 1);
    const uint64_t seed_v = a.seed_dev ? *a.seed_dev : a.seed;  // device seed: fresh noise on CUDA-graph replays
    const int wib = threadIdx.x / G;
    const int64_t items = (int64_t)a.Q * a.d * a.Rs;
    int64_t item = (int64_t)blockIdx.x * (blockDim.x / G) + wib;
    const bool live = item < items;
    if (!live) item = items - 1;
    float* dz = dz_all + (size_t)wib * a.ic;
    const int rs = (int)(item % a.Rs);
    const int s = (int)((item / a.Rs) % a.d);
    const int64_t q = item / ((int64_t)a.Rs * a.d);
    const int ws = a.Wsets == 1 ? 0 : (int)(q % a.w);

    int o;
    if (a.unravel) {
        const int p = a.pa[q * a.sp_q + (int64_t)(s >> 1) * a.sp_s + (int64_t)rs * a.sp_r];
        o = (s & 1) ? p % a.oc : p / a.oc;
    } else {
        o = a.pa[q * a.sp_q + (int64_t)s * a.sp_s + (int64_t)rs * a.sp_r];
    }
    const int rsel = a.rep ? a.rep[q] : rs;
    const int64_t wbase = (int64_t)ws * a.sw_w + (int64_t)s * a.sw_d;
    const float* lwp = a.lw + wbase + (int64_t)o * a.sw_o;
    const int64_t qe = a.xc ? q % a.Qe : 0;
    const bool hl = a.xc && 2 * s < a.xc_din, hr = a.xc && 2 * s + 1 < a.xc_din;
    const int64_t xlo = qe * a.sxc_b + (int64_t)(2 * s) * a.sxc_d;
    const int64_t xro = qe * a.sxc_b + (int64_t)(2 * s + 1) * a.sxc_d;
    const uint64_t stream_id = ((uint64_t)a.d << 40) ^ ((uint64_t)a.ic << 8) ^ (uint64_t)a.Rs;
    const bool gl_ok = 2 * s < a.gc_d, gr_ok = 2 * s + 1 < a.gc_d;  // padded child scopes carry no gradient
    const float* gl = a.gc + ((q * a.gc_d + 2 * s) * (int64_t)a.c) * a.Rs + rs;
    const float* gr = a.gc + ((q * a.gc_d + 2 * s + 1) * (int64_t)a.c) * a.Rs + rs;

    auto logit = [&](int i, float& xoff) -> float {
        int r = rsel, ii = i;
        if (a.root) { r = i / a.ic_top; ii = i - r * a.ic_top; }
        float v = lwp[(int64_t)ii * a.sw_i + (int64_t)r * a.sw_r];
        xoff = 0.f;
        if (a.xc) {
            const int l = ii / a.c, rp = ii - l * a.c;
            if (hl) xoff += a.xc[xlo + (int64_t)l * a.sxc_c + (int64_t)r * a.sxc_r];
            if (hr) xoff += a.xc[xro + (int64_t)rp * a.sxc_c + (int64_t)r * a.sxc_r];
            v += xoff;
        }
        if (a.gumbel) v += a.gumbel[((q * a.d + s) * (int64_t)a.ic + i) * a.Rs + rs];
        else v += spn_gumbel(((((uint64_t)q + a.q_offset) * a.d + s) * (uint64_t)a.ic + i) * a.Rs + rs, stream_id, seed_v);
        return v;
    };
    auto upstream = [&](int i) -> float {
        if (a.root) {
            const int r = i / a.ic_top;
            i -= r * a.ic_top;
            if (a.gc_allreps) {  // no Sum layer below: the one-hots of unselected repetitions still reach the leaves
                const int Rall = a.ic / a.ic_top;
                const int l = i / a.c, rp = i - l * a.c;
                float g = gl_ok ? a.gc[((q * a.gc_d) * (int64_t)a.c + l) * Rall + r] : 0.f;
                if (gr_ok) g += a.gc[((q * a.gc_d + 1) * (int64_t)a.c + rp) * Rall + r];
                return g;
            }
            if (r != rsel) return 0.f;  // a Sum layer below clears unselected repetitions (layers.py:233-234)
        }
        const int l = i / a.c, rp = i - l * a.c;
        float g = gl_ok ? gl[(int64_t)l * a.Rs] : 0.f;
        if (gr_ok) g += gr[(int64_t)rp * a.Rs];
        return g;
    };

    // ic <= G: one element per lane, evaluate the logit (Philox + loads) and the upstream gradient once
    const bool single = a.ic <= G;
    float zc = -CUDART_INF_F, gcv = 0.f, xoc = 0.f;
    if (single && lane < a.ic) { zc = logit(lane, xoc); gcv = upstream(lane); }
    auto getz = [&](int i, float& xoff) -> float {
        if (single) { xoff = xoc; return zc; }
        return logit(i, xoff);
    };
    auto getg = [&](int i) -> float { return single ? gcv : upstream(i); };

    float xo;
    float m = -CUDART_INF_F;
    for (int i = lane; i < a.ic; i += G) m = fmaxf(m, getz(i, xo));
    m = group_max<G>(m);
    float Z = 0.f, SG = 0.f;
    for (int i = lane; i < a.ic; i += G) {
        const float e = __expf(getz(i, xo) - m);
        Z += e;
        SG += e * getg(i);
    }
    Z = group_sum<G>(Z);
    SG = group_sum<G>(SG);
    const float invZ = 1.f / Z, mean = SG * invZ;
    float T = 0.f;
    float* dlwp = a.dlw + wbase + (int64_t)o * a.sw_o;
    for (int i = lane; i < a.ic; i += G) {
        const float z = getz(i, xo);
        const float dl = __expf(z - m) * invZ * (getg(i) - mean);
        int r = rsel, ii = i;
        if (a.root) { r = i / a.ic_top; ii = i - r * a.ic_top; }
        if (live) atomicAdd(dlwp + (int64_t)ii * a.sw_i + (int64_t)r * a.sw_r, dl);
        T += dl * xo;
        if (!a.root) dz[i] = dl;
        else if (a.dxc && live) {
            const int l = ii / a.c, rp = ii - l * a.c;
            if (hl) atomicAdd(a.dxc + xlo + (int64_t)l * a.sxc_c + (int64_t)r * a.sxc_r, dl);
            if (hr) atomicAdd(a.dxc + xro + (int64_t)rp * a.sxc_c + (int64_t)r * a.sxc_r, dl);
        }
    }
    if (a.root) return;
    __syncwarp();
    if (a.dxc && live) {  // row / column sums of dz viewed as [c, c]
        for (int idx = lane; idx < 2 * a.c; idx += G) {
            float acc = 0.f;
            if (idx < a.c) {
                if (!hl) continue;
                for (int rp = 0; rp < a.c; ++rp) acc += dz[idx * a.c + rp];
                atomicAdd(a.dxc + xlo + (int64_t)idx * a.sxc_c + (int64_t)rsel * a.sxc_r, acc);
            } else {
                if (!hr) continue;
                const int rp = idx - a.c;
                for (int l = 0; l < a.c; ++l) acc += dz[l * a.c + rp];
                atomicAdd(a.dxc + xro + (int64_t)rp * a.sxc_c + (int64_t)rsel * a.sxc_r, acc);
            }
        }
    }
    if (a.gparent) {
        T = group_sum<G>(T);
        const float* lws = a.lw + wbase + (int64_t)rsel * a.sw_r;
        float* gp = a.gparent + ((q * a.d + s) * (int64_t)a.oc) * a.Rs + rs;
        for (int op = 0; op < a.oc; ++op) {
            float acc = 0.f;
            for (int i = lane; i < a.ic; i += G) acc += dz[i] * lws[(int64_t)i * a.sw_i + (int64_t)op * a.sw_o];
            acc = group_sum<G>(acc);
            if (lane == 0 && live) gp[(int64_t)op * a.Rs] = acc + T;
        }
    }
}

struct SampleLeafBwdArgs {
    const float* mu; const float* sigma;
    int Wsets, w, Q, F, I, R, card, Rs, dP;
    const int32_t* ch; const int32_t* rep;
    const float* eps; int mpe; uint64_t seed, q_offset; const uint64_t* seed_dev;
    const int32_t* perm;   // [F, R] slot -> output feature (null: identity / un-post-processed samples)
    const float* evidence; int64_t se_q, se_f, se_r; int Qe;
    int squash; const float* y;
    int g0_allreps;        // with rep != null: g0 is [Q, dP, I, R], evaluated with the leaf parameters of every repetition
    const float* gy;       // gradient w.r.t. the kernel's output [Q, F, Rs]
    float* dmu; float* dsigma; float* g0;   // g0 [Q, dP, I, Rs]: gradient w.r.t. the leaf-channel one-hots per scope
};

// One thread per (q, scope, channel, rs): sums the scope's features; the thread of the chosen channel also adds the
// (mu, sigma) gradients of that feature.
__global__ void __launch_bounds__(256) sample_leaf_bwd_kernel(SampleLeafBwdArgs a) {
    const int Rg = a.g0_allreps ? a.R : a.Rs;
    const int64_t total = (int64_t)a.Q * a.dP * a.I * Rg;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const uint64_t seed_v = a.seed_dev ? *a.seed_dev : a.seed;
    const int rg = (int)(tid % Rg);
    const int i = (int)((tid / Rg) % a.I);
    const int s = (int)((tid / ((int64_t)Rg * a.I)) % a.dP);
    const int64_t q = tid / ((int64_t)Rg * a.I * a.dP);
    const int ws = a.Wsets == 1 ? 0 : (int)(q % a.w);
    const int rs = a.g0_allreps ? 0 : rg;
    const int rsel = a.rep ? a.rep[q] : rs;   // repetition the sample was drawn from
    const int r = a.g0_allreps ? rg : rsel;   // repetition whose parameters this thread evaluates
    float acc = 0.f;
    const int f1 = min(a.F, (s + 1) * a.card);
    for (int f = s * a.card; f < f1; ++f) {
        const int fo = a.perm ? a.perm[(int64_t)f * a.R + rsel] : f;
        const int64_t oi = (q * a.F + fo) * (int64_t)a.Rs + rs;
        float g = a.gy[oi];
        if (a.evidence) {
            const float ev = a.evidence[(q % a.Qe) * a.se_q + (int64_t)fo * a.se_f + (int64_t)rs * a.se_r];
            if (!isnan(ev)) g = 0.f;
        }
        if (a.squash) {
            const float yv = a.y[oi];
            g = fabsf(yv) >= 0.99998771f ? 0.f : g * (1.f - yv * yv);  // clamp(-6, 6) then tanh
        }
        const int64_t pi = (((int64_t)ws * a.F + f) * a.I + i) * a.R + r;
        const int64_t ne = (q * a.F + f) * (int64_t)a.Rs + rs;
        float e = 0.f;
        if (!a.mpe) {
            if (a.eps) e = a.eps[ne];
            else e = spn_normal((((uint64_t)q + a.q_offset) * a.F + f) * (uint64_t)a.Rs + rs, 0x1eafULL, seed_v);
        }
        acc += g * (a.mu[pi] + a.sigma[pi] * e);
        if (a.ch[ne] == i && r == rsel && g != 0.f) {
            atomicAdd(a.dmu + pi, g);
            if (!a.mpe) atomicAdd(a.dsigma + pi, g * e);
        }
    }
    if (a.g0) a.g0[tid] = acc;
}

extern "C" int spn_sample_sum_bwd(const float* lw, int64_t sw_w, int64_t sw_d, int64_t sw_i, int64_t sw_o, int64_t sw_r,
                                  int Wsets, int w, int Q, int d, int ic, int oc, int Rs, const int32_t* pa,
                                  int64_t sp_q, int64_t sp_s, int64_t sp_r, int unravel, const int32_t* rep,
                                  const float* xc, int64_t sxc_b, int64_t sxc_d, int64_t sxc_c, int64_t sxc_r,
                                  int xc_din, int c, int Qe, int root, int ic_top, const float* gumbel, uint64_t seed,
                                  uint64_t q_offset, const uint64_t* seed_dev, const float* gc, int gc_d, int gc_allreps, float* dlw, float* dxc,
                                  float* gparent, void* stream) {
    if (!lw || !pa || !gc || !dlw || Q < 0 || d <= 0 || ic <= 0 || oc <= 0 || Rs <= 0 || c <= 0) return SPN_ERR_ARG;
    if (Wsets != 1 && (Wsets != w || Q % w != 0)) return SPN_ERR_ARG;
    if (xc && (Qe <= 0 || Q % Qe != 0 || !dxc)) return SPN_ERR_ARG;
    if (root && (d != 1 || Rs != 1 || ic_top <= 0 || ic % ic_top != 0 || c * c != ic_top || !rep)) return SPN_ERR_ARG;
    if ((!root && c * c != ic) || gc_d <= 0) return SPN_ERR_ARG;
    if (Q == 0) return SPN_OK;
    SampleSumBwdArgs a{};
    a.lw = lw; a.sw_w = sw_w; a.sw_d = sw_d; a.sw_i = sw_i; a.sw_o = sw_o; a.sw_r = sw_r;
    a.Wsets = Wsets; a.w = w; a.Q = Q; a.d = d; a.ic = ic; a.oc = oc; a.Rs = Rs;
    a.pa = pa; a.sp_q = sp_q; a.sp_s = sp_s; a.sp_r = sp_r; a.unravel = unravel; a.rep = rep;
    a.xc = xc; a.sxc_b = sxc_b; a.sxc_d = sxc_d; a.sxc_c = sxc_c; a.sxc_r = sxc_r; a.xc_din = xc_din; a.c = c; a.Qe = Qe;
    a.root = root; a.ic_top = ic_top; a.gumbel = gumbel; a.seed = seed; a.q_offset = q_offset; a.seed_dev = seed_dev;
    a.gc = gc; a.gc_d = gc_d; a.gc_allreps = (root && gc_allreps) ? 1 : 0; a.dlw = dlw; a.dxc = xc ? dxc : nullptr; a.gparent = root ? nullptr : gparent;
    int G = 4;
    while (G < 32 && G < ic) G <<= 1;
    int threads = 256;
    size_t smem = 0;
    if (!root) {
        while (threads > G && (size_t)(threads / G) * ic * sizeof(float) > 48 * 1024) threads >>= 1;
        smem = (size_t)(threads / G) * ic * sizeof(float);
        if (smem > 48 * 1024) return SPN_ERR_UNSUPPORTED;
    }
    const int64_t items = (int64_t)Q * d * Rs;
    const unsigned blocks = spn_blocks(items, threads / G);
    cudaStream_t st = (cudaStream_t)stream;
    switch (G) {
        case 4: sample_sum_bwd_kernel<4><<<blocks, threads, smem, st>>>(a); break;
        case 8: sample_sum_bwd_kernel<8><<<blocks, threads, smem, st>>>(a); break;
        case 16: sample_sum_bwd_kernel<16><<<blocks, threads, smem, st>>>(a); break;
        default: sample_sum_bwd_kernel<32><<<blocks, threads, smem, st>>>(a); break;
    }
    return spn_launch_status();
}

extern "C" int spn_sample_leaf_bwd(const float* mu, const float* sigma, int Wsets, int w, int Q, int F, int I, int R,
                                   int card, int Rs, int dP, const int32_t* ch, const int32_t* rep, const float* eps,
                                   int mpe, uint64_t seed, uint64_t q_offset, const uint64_t* seed_dev, const int32_t* perm,
                                   const float* evidence,
                                   int64_t se_q, int64_t se_f, int64_t se_r, int Qe, int squash, const float* y,
                                   int g0_allreps, const float* gy, float* dmu, float* dsigma, float* g0, void* stream) {
    if (!mu || !sigma || !ch || !gy || !dmu || !dsigma || Q < 0 || F <= 0 || I <= 0 || R <= 0 || card <= 0 || Rs <= 0)
        return SPN_ERR_ARG;
    if (dP <= 0 || (int64_t)dP * card < F) return SPN_ERR_ARG;
    if (Wsets != 1 && (Wsets != w || Q % w != 0)) return SPN_ERR_ARG;
    if (evidence && (Qe <= 0 || Q % Qe != 0)) return SPN_ERR_ARG;
    if (squash && !y) return SPN_ERR_ARG;
    if (g0_allreps && (!rep || Rs != 1)) return SPN_ERR_ARG;
    if (Q == 0) return SPN_OK;
    SampleLeafBwdArgs a{};
    a.mu = mu; a.sigma = sigma; a.Wsets = Wsets; a.w = w; a.Q = Q; a.F = F; a.I = I; a.R = R; a.card = card; a.Rs = Rs;
    a.dP = dP; a.ch = ch; a.rep = rep; a.eps = eps; a.mpe = mpe; a.seed = seed; a.q_offset = q_offset; a.seed_dev = seed_dev;
    a.perm = perm;
    a.evidence = evidence; a.se_q = se_q; a.se_f = se_f; a.se_r = se_r; a.Qe = Qe; a.squash = squash; a.y = y;
    a.g0_allreps = g0_allreps ? 1 : 0; a.gy = gy; a.dmu = dmu; a.dsigma = dsigma; a.g0 = g0;
    const int64_t total = (int64_t)Q * dP * I * (g0_allreps ? R : Rs);
    sample_leaf_bwd_kernel<<<spn_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(a);
    return spn_launch_status();
}
