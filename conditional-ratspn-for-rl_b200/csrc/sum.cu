// Sum layer kernels (exact log-sum-exp semantics), optionally fused with the CrossProduct layer below.
// These are the general kernels: any channel counts, shared (RatSpn) or per-conditional (CSPN) weights, inner layer
// or root.  For per-conditional weights every weight is used once, so the kernels are HBM-streaming by construction.
// Reference semantics: layers.py:110-128, :419-440, :545-563, rat_spn.py:230-236.
#include "common.cuh"

struct SumArgs {
    const float* x; int64_t sx_b, sx_d, sx_c, sx_r; int d_in, c;
    const float* lw; int64_t sw_w, sw_d, sw_i, sw_o, sw_r;
    int Wsets, w, B, d_out, ic, oc, R, reduce_r;
    float* out; const float* outc; const float* g; int64_t so_b, so_d, so_o, so_r;
    float* dlw; float* dx; int64_t sdx_b, sdx_d, sdx_c, sdx_r; int nsplit;
};

// value of in-channel i of sum scope s for row b, repetition r
template <bool XPROD>
__device__ __forceinline__ float sum_input(const SumArgs& a, int64_t b, int s, int i, int r) {
    if (XPROD) {
        const int l = i / a.c, rp = i - l * a.c;
        const int sl = 2 * s, sr = 2 * s + 1;
        float v = 0.f;
        if (sl < a.d_in) v += a.x[b * a.sx_b + (int64_t)sl * a.sx_d + (int64_t)l * a.sx_c + (int64_t)r * a.sx_r];
        if (sr < a.d_in) v += a.x[b * a.sx_b + (int64_t)sr * a.sx_d + (int64_t)rp * a.sx_c + (int64_t)r * a.sx_r];
        return v;
    } else {
        return a.x[b * a.sx_b + (int64_t)s * a.sx_d + (int64_t)i * a.sx_c + (int64_t)r * a.sx_r];
    }
}

// Forward.  One thread per output (b, s, o, r) (r fastest; reduce_r: per (b, s, o) looping over r as well).
// Fast path: shift by the maximum of the *inputs* (one exp per term, no exp of the weights needed because the
// exponent x_i - max + lw is <= 0); if that sum underflows (all mass on channels far below the max) the exact
// two-pass log-sum-exp over x_i + lw is recomputed for this output, which is what th.logsumexp does.
template <bool XPROD>
__global__ void __launch_bounds__(256) sum_fwd_kernel(SumArgs a) {
    const int Rt = a.reduce_r ? 1 : a.R;
    const int64_t total = (int64_t)a.B * a.d_out * a.oc * Rt;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int r0 = (int)(tid % Rt);
    const int o = (int)((tid / Rt) % a.oc);
    const int s = (int)((tid / ((int64_t)Rt * a.oc)) % a.d_out);
    const int64_t b = tid / ((int64_t)Rt * a.oc * a.d_out);
    const int ws = a.Wsets == 1 ? 0 : (int)(b % a.w);
    const float* lwp = a.lw + (int64_t)ws * a.sw_w + (int64_t)s * a.sw_d + (int64_t)o * a.sw_o;
    const int r_lo = a.reduce_r ? 0 : r0, r_hi = a.reduce_r ? a.R : r0 + 1;

    const float* xl = nullptr;
    const float* xr = nullptr;
    if (XPROD) {
        if (2 * s < a.d_in) xl = a.x + b * a.sx_b + (int64_t)(2 * s) * a.sx_d;
        if (2 * s + 1 < a.d_in) xr = a.x + b * a.sx_b + (int64_t)(2 * s + 1) * a.sx_d;
    }
    float mx = -CUDART_INF_F;
    for (int r = r_lo; r < r_hi; ++r) {
        if (XPROD) {
            float ml = xl ? -CUDART_INF_F : 0.f, mr = xr ? -CUDART_INF_F : 0.f;
            for (int l = 0; l < a.c; ++l) {
                if (xl) ml = fmaxf(ml, xl[(int64_t)l * a.sx_c + (int64_t)r * a.sx_r]);
                if (xr) mr = fmaxf(mr, xr[(int64_t)l * a.sx_c + (int64_t)r * a.sx_r]);
            }
            mx = fmaxf(mx, ml + mr);
        } else {
            for (int i = 0; i < a.ic; ++i) mx = fmaxf(mx, sum_input<false>(a, b, s, i, r));
        }
    }
    float result;
    if (mx == -CUDART_INF_F) {
        result = -CUDART_INF_F;
    } else {
        float acc = 0.f;
        for (int r = r_lo; r < r_hi; ++r) {
            const float* lwr = lwp + (int64_t)r * a.sw_r;
            if (XPROD) {
                for (int l = 0; l < a.c; ++l) {
                    const float lv = (xl ? xl[(int64_t)l * a.sx_c + (int64_t)r * a.sx_r] : 0.f) - mx;
                    const float* lwl = lwr + (int64_t)l * a.c * a.sw_i;
                    for (int rp = 0; rp < a.c; ++rp) {
                        const float rv = xr ? xr[(int64_t)rp * a.sx_c + (int64_t)r * a.sx_r] : 0.f;
                        acc += __expf(lv + rv + lwl[(int64_t)rp * a.sw_i]);
                    }
                }
            } else {
                for (int i = 0; i < a.ic; ++i)
                    acc += __expf(sum_input<false>(a, b, s, i, r) - mx + lwr[(int64_t)i * a.sw_i]);
            }
        }
        if (acc >= 1e-30f) {
            result = mx + logf(acc);
        } else {
            // exact rescue: two-pass log-sum-exp over t_i = x_i + lw_i
            float m2 = -CUDART_INF_F;
            for (int r = r_lo; r < r_hi; ++r)
                for (int i = 0; i < a.ic; ++i)
                    m2 = fmaxf(m2, sum_input<XPROD>(a, b, s, i, r) + lwp[(int64_t)r * a.sw_r + (int64_t)i * a.sw_i]);
            if (m2 == -CUDART_INF_F) {
                result = -CUDART_INF_F;
            } else {
                float acc2 = 0.f;
                for (int r = r_lo; r < r_hi; ++r)
                    for (int i = 0; i < a.ic; ++i)
                        acc2 += expf(sum_input<XPROD>(a, b, s, i, r) + lwp[(int64_t)r * a.sw_r + (int64_t)i * a.sw_i] - m2);
                result = m2 + logf(acc2);
            }
        }
    }
    a.out[b * a.so_b + (int64_t)s * a.so_d + (int64_t)o * a.so_o + (a.reduce_r ? 0 : (int64_t)r0 * a.so_r)] = result;
}

// Forward, one WARP per output: for layers with few outputs and long reductions (the root of a wide CSPN: one output per
// conditional over R * c^2 = 2048 terms) a thread per output leaves the GPU idle and serialises the exps.  Lanes stride
// over the flattened (r, i) terms; same two-stage semantics as sum_fwd_kernel (input-max shift, exact rescue on underflow).
template <bool XPROD>
__global__ void __launch_bounds__(256) sum_fwd_warp_kernel(SumArgs a) {
    const int Rt = a.reduce_r ? 1 : a.R;
    const int64_t total = (int64_t)a.B * a.d_out * a.oc * Rt;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= total) return;
    const int r0 = (int)(wid % Rt);
    const int o = (int)((wid / Rt) % a.oc);
    const int s = (int)((wid / ((int64_t)Rt * a.oc)) % a.d_out);
    const int64_t b = wid / ((int64_t)Rt * a.oc * a.d_out);
    const int ws = a.Wsets == 1 ? 0 : (int)(b % a.w);
    const float* lwp = a.lw + (int64_t)ws * a.sw_w + (int64_t)s * a.sw_d + (int64_t)o * a.sw_o;
    const int r_lo = a.reduce_r ? 0 : r0, nr = a.reduce_r ? a.R : 1;
    const int nterm = nr * a.ic;
    float mx = -CUDART_INF_F;
    for (int k = lane; k < nterm; k += 32) {
        const int rr = k / a.ic, i = k - rr * a.ic;
        mx = fmaxf(mx, sum_input<XPROD>(a, b, s, i, r_lo + rr));
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sh));
    float result = -CUDART_INF_F;
    if (mx != -CUDART_INF_F) {
        float acc = 0.f;
        for (int k = lane; k < nterm; k += 32) {
            const int rr = k / a.ic, i = k - rr * a.ic, r = r_lo + rr;
            acc += __expf(sum_input<XPROD>(a, b, s, i, r) - mx + lwp[(int64_t)r * a.sw_r + (int64_t)i * a.sw_i]);
        }
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
        if (acc >= 1e-30f) {
            result = mx + logf(acc);
        } else {
            float m2 = -CUDART_INF_F;
            for (int k = lane; k < nterm; k += 32) {
                const int rr = k / a.ic, i = k - rr * a.ic, r = r_lo + rr;
                m2 = fmaxf(m2, sum_input<XPROD>(a, b, s, i, r) + lwp[(int64_t)r * a.sw_r + (int64_t)i * a.sw_i]);
            }
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, sh));
            if (m2 != -CUDART_INF_F) {
                float acc2 = 0.f;
                for (int k = lane; k < nterm; k += 32) {
                    const int rr = k / a.ic, i = k - rr * a.ic, r = r_lo + rr;
                    acc2 += expf(sum_input<XPROD>(a, b, s, i, r) + lwp[(int64_t)r * a.sw_r + (int64_t)i * a.sw_i] - m2);
                }
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) acc2 += __shfl_xor_sync(0xffffffffu, acc2, sh);
                result = m2 + logf(acc2);
            }
        }
    }
    if (lane == 0)
        a.out[b * a.so_b + (int64_t)s * a.so_d + (int64_t)o * a.so_o + (a.reduce_r ? 0 : (int64_t)r0 * a.so_r)] = result;
}

// Backward w.r.t. the log-weights.  One thread per weight element, walking the rows that use it:
// dlw = sum_b g[b,s,o,(r)] * exp(x_i + lw - out[b,s,o,(r)]).  Thread order follows the weight memory order
// (inner layer: r fastest; root: o fastest) so out/g loads are coalesced.
template <bool XPROD>
__global__ void __launch_bounds__(256) sum_bwd_w_kernel(SumArgs a) {
    const int64_t per_set = (int64_t)a.d_out * a.ic * a.oc * a.R;
    const int64_t total = (int64_t)a.Wsets * per_set;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    int ws, s, i, o, r;
    {
        int64_t t = tid;
        if (!a.reduce_r) {
            r = (int)(t % a.R); t /= a.R;
            o = (int)(t % a.oc); t /= a.oc;
            i = (int)(t % a.ic); t /= a.ic;
        } else {
            o = (int)(t % a.oc); t /= a.oc;
            i = (int)(t % a.ic); t /= a.ic;
            r = (int)(t % a.R); t /= a.R;
        }
        s = (int)(t % a.d_out); t /= a.d_out;
        ws = (int)t;
    }
    const int64_t woff = (int64_t)ws * a.sw_w + (int64_t)s * a.sw_d + (int64_t)i * a.sw_i + (int64_t)o * a.sw_o +
                         (int64_t)r * a.sw_r;
    const float lwv = a.lw[woff];
    const int64_t ooff = (int64_t)s * a.so_d + (int64_t)o * a.so_o + (a.reduce_r ? 0 : (int64_t)r * a.so_r);
    const int64_t step = a.Wsets == 1 ? 1 : a.w;
    const int64_t first = a.Wsets == 1 ? 0 : ws;
    float acc = 0.f;
    for (int64_t b = first + (int64_t)blockIdx.y * step; b < a.B; b += step * a.nsplit) {
        const float ov = a.outc[b * a.so_b + ooff];
        if (ov == -CUDART_INF_F) continue;
        const float gv = a.g[b * a.so_b + ooff];
        acc += gv * __expf(sum_input<XPROD>(a, b, s, i, r) + lwv - ov);
    }
    if (a.nsplit == 1) a.dlw[woff] = acc;
    else atomicAdd(a.dlw + woff, acc);
}

// Backward w.r.t. x.  One thread per input element (b, scope, channel, r), r fastest; every element is written
// exactly once (no atomics): dx = sum over partner channel and o of g * exp(x_i + lw - out).
template <bool XPROD>
__global__ void __launch_bounds__(256) sum_bwd_x_kernel(SumArgs a) {
    const int nsc = XPROD ? a.d_in : a.d_out;
    const int nch = XPROD ? a.c : a.ic;
    const int64_t total = (int64_t)a.B * nsc * nch * a.R;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int r = (int)(tid % a.R);
    const int ch = (int)((tid / a.R) % nch);
    const int sc = (int)((tid / ((int64_t)a.R * nch)) % nsc);
    const int64_t b = tid / ((int64_t)a.R * nch * nsc);
    const int ws = a.Wsets == 1 ? 0 : (int)(b % a.w);
    const int s = XPROD ? sc >> 1 : sc;
    const float* lwp = a.lw + (int64_t)ws * a.sw_w + (int64_t)s * a.sw_d + (int64_t)r * a.sw_r;
    const float* op = a.outc + b * a.so_b + (int64_t)s * a.so_d + (a.reduce_r ? 0 : (int64_t)r * a.so_r);
    const float* gp = a.g + b * a.so_b + (int64_t)s * a.so_d + (a.reduce_r ? 0 : (int64_t)r * a.so_r);
    float acc = 0.f;
    if (XPROD) {
        const int side = sc & 1;
        const int psc = sc ^ 1;
        const float xv = a.x[b * a.sx_b + (int64_t)sc * a.sx_d + (int64_t)ch * a.sx_c + (int64_t)r * a.sx_r];
        const float* xp = psc < a.d_in ? a.x + b * a.sx_b + (int64_t)psc * a.sx_d + (int64_t)r * a.sx_r : nullptr;
        for (int pc = 0; pc < a.c; ++pc) {
            const float xi = xv + (xp ? xp[(int64_t)pc * a.sx_c] : 0.f);
            const int i = side == 0 ? ch * a.c + pc : pc * a.c + ch;
            const float* lwi = lwp + (int64_t)i * a.sw_i;
            for (int o = 0; o < a.oc; ++o) {
                const float ov = op[(int64_t)o * a.so_o];
                if (ov == -CUDART_INF_F) continue;
                acc += gp[(int64_t)o * a.so_o] * __expf(xi + lwi[(int64_t)o * a.sw_o] - ov);
            }
        }
    } else {
        const float xi = a.x[b * a.sx_b + (int64_t)sc * a.sx_d + (int64_t)ch * a.sx_c + (int64_t)r * a.sx_r];
        const float* lwi = lwp + (int64_t)ch * a.sw_i;
        for (int o = 0; o < a.oc; ++o) {
            const float ov = op[(int64_t)o * a.so_o];
            if (ov == -CUDART_INF_F) continue;
            acc += gp[(int64_t)o * a.so_o] * __expf(xi + lwi[(int64_t)o * a.sw_o] - ov);
        }
    }
    a.dx[b * a.sdx_b + (int64_t)sc * a.sdx_d + (int64_t)ch * a.sdx_c + (int64_t)r * a.sdx_r] = acc;
}

// Stand-alone CrossProduct (materialising), contiguous in/out.
__global__ void __launch_bounds__(256) xprod_fwd_kernel(const float* __restrict__ x, int B, int d_in, int c, int R,
                                                       int d_out, float* __restrict__ out) {
    const int64_t cc = (int64_t)c * c;
    const int64_t total = (int64_t)B * d_out * cc * R;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int r = (int)(tid % R);
    const int i = (int)((tid / R) % cc);
    const int s = (int)((tid / (R * cc)) % d_out);
    const int64_t b = tid / (R * cc * d_out);
    const int l = i / c, rp = i % c;
    float v = 0.f;
    if (2 * s < d_in) v += x[((b * d_in + 2 * s) * c + l) * R + r];
    if (2 * s + 1 < d_in) v += x[((b * d_in + 2 * s + 1) * c + rp) * R + r];
    out[tid] = v;
}

__global__ void __launch_bounds__(256) xprod_bwd_kernel(const float* __restrict__ g, int B, int d_in, int c, int R,
                                                       int d_out, float* __restrict__ dx) {
    const int64_t total = (int64_t)B * d_in * c * R;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int r = (int)(tid % R);
    const int ch = (int)((tid / R) % c);
    const int sc = (int)((tid / ((int64_t)R * c)) % d_in);
    const int64_t b = tid / ((int64_t)R * c * d_in);
    const int s = sc >> 1;
    const float* gp = g + ((b * d_out + s) * (int64_t)c * c) * R + r;
    float acc = 0.f;
    if ((sc & 1) == 0) {
        for (int pc = 0; pc < c; ++pc) acc += gp[(int64_t)(ch * c + pc) * R];
    } else {
        for (int pc = 0; pc < c; ++pc) acc += gp[(int64_t)(pc * c + ch) * R];
    }
    dx[tid] = acc;
}

static int fill_args(SumArgs& a, int xprod, const float* x, int64_t sx_b, int64_t sx_d, int64_t sx_c, int64_t sx_r,
                     int d_in, int c, const float* lw, int64_t sw_w, int64_t sw_d, int64_t sw_i, int64_t sw_o,
                     int64_t sw_r, int Wsets, int w, int B, int d_out, int ic, int oc, int R, int reduce_r,
                     int64_t so_b, int64_t so_d, int64_t so_o, int64_t so_r) {
    if (!x || !lw || B < 0 || d_out <= 0 || ic <= 0 || oc <= 0 || R <= 0) return SPN_ERR_ARG;
    if (Wsets != 1 && (Wsets != w || B % w != 0)) return SPN_ERR_ARG;
    if (xprod) {
        if (c <= 0 || d_in <= 0 || (int64_t)c * c != ic || d_in > 2 * d_out) return SPN_ERR_ARG;
    }
    a.x = x; a.sx_b = sx_b; a.sx_d = sx_d; a.sx_c = sx_c; a.sx_r = sx_r; a.d_in = d_in; a.c = c;
    a.lw = lw; a.sw_w = sw_w; a.sw_d = sw_d; a.sw_i = sw_i; a.sw_o = sw_o; a.sw_r = sw_r;
    a.Wsets = Wsets; a.w = w; a.B = B; a.d_out = d_out; a.ic = ic; a.oc = oc; a.R = R; a.reduce_r = reduce_r;
    a.so_b = so_b; a.so_d = so_d; a.so_o = so_o; a.so_r = so_r; a.nsplit = 1;
    return SPN_OK;
}

extern "C" int spn_sum_fwd(int xprod, const float* x, int64_t sx_b, int64_t sx_d, int64_t sx_c, int64_t sx_r, int d_in,
                           int c, const float* lw, int64_t sw_w, int64_t sw_d, int64_t sw_i, int64_t sw_o,
                           int64_t sw_r, int Wsets, int w, int B, int d_out, int ic, int oc, int R, int reduce_r,
                           float* out, int64_t so_b, int64_t so_d, int64_t so_o, int64_t so_r, void* stream) {
    SumArgs a{};
    int rc = fill_args(a, xprod, x, sx_b, sx_d, sx_c, sx_r, d_in, c, lw, sw_w, sw_d, sw_i, sw_o, sw_r, Wsets, w, B,
                       d_out, ic, oc, R, reduce_r, so_b, so_d, so_o, so_r);
    if (rc != SPN_OK || !out) return SPN_ERR_ARG;
    if (B == 0) return SPN_OK;
    a.out = out;
    const int64_t total = (int64_t)B * d_out * oc * (reduce_r ? 1 : R);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nterm = (int64_t)ic * (reduce_r ? R : 1);
    if (total < 148 * 256 && nterm >= 128) {  // few outputs, long reductions: one warp per output
        if (xprod) sum_fwd_warp_kernel<true><<<spn_blocks(total * 32, 256), 256, 0, st>>>(a);
        else sum_fwd_warp_kernel<false><<<spn_blocks(total * 32, 256), 256, 0, st>>>(a);
        return spn_launch_status();
    }
    if (xprod) sum_fwd_kernel<true><<<spn_blocks(total, 256), 256, 0, st>>>(a);
    else sum_fwd_kernel<false><<<spn_blocks(total, 256), 256, 0, st>>>(a);
    return spn_launch_status();
}

extern "C" int spn_sum_bwd_w(int xprod, const float* x, int64_t sx_b, int64_t sx_d, int64_t sx_c, int64_t sx_r,
                             int d_in, int c, const float* lw, int64_t sw_w, int64_t sw_d, int64_t sw_i, int64_t sw_o,
                             int64_t sw_r, int Wsets, int w, int B, int d_out, int ic, int oc, int R, int reduce_r,
                             const float* out, const float* g, int64_t so_b, int64_t so_d, int64_t so_o, int64_t so_r,
                             float* dlw, int64_t dlw_numel, void* stream) {
    SumArgs a{};
    int rc = fill_args(a, xprod, x, sx_b, sx_d, sx_c, sx_r, d_in, c, lw, sw_w, sw_d, sw_i, sw_o, sw_r, Wsets, w, B,
                       d_out, ic, oc, R, reduce_r, so_b, so_d, so_o, so_r);
    if (rc != SPN_OK || !out || !g || !dlw) return SPN_ERR_ARG;
    a.outc = out; a.g = g; a.dlw = dlw;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)Wsets * d_out * ic * oc * R;
    if (dlw_numel != total) return SPN_ERR_ARG;
    const int64_t rows = Wsets == 1 ? B : B / w;
    const int64_t blocks = (total + 255) / 256;
    int nsplit = 1;
    if (blocks < 296 && rows > 64) {
        nsplit = (int)min((int64_t)64, min(rows / 32 + 1, (296 + blocks - 1) / blocks));
        if (nsplit < 1) nsplit = 1;
    }
    a.nsplit = nsplit;
    if (nsplit > 1) {
        cudaError_t e = cudaMemsetAsync(dlw, 0, total * sizeof(float), st);
        if (e != cudaSuccess) return spn_cuda_status(e);
    }
    dim3 grid(spn_blocks(total, 256), nsplit);
    if (xprod) sum_bwd_w_kernel<true><<<grid, 256, 0, st>>>(a);
    else sum_bwd_w_kernel<false><<<grid, 256, 0, st>>>(a);
    return spn_launch_status();
}

extern "C" int spn_sum_bwd_x(int xprod, const float* x, int64_t sx_b, int64_t sx_d, int64_t sx_c, int64_t sx_r,
                             int d_in, int c, const float* lw, int64_t sw_w, int64_t sw_d, int64_t sw_i, int64_t sw_o,
                             int64_t sw_r, int Wsets, int w, int B, int d_out, int ic, int oc, int R, int reduce_r,
                             const float* out, const float* g, int64_t so_b, int64_t so_d, int64_t so_o, int64_t so_r,
                             float* dx, int64_t sdx_b, int64_t sdx_d, int64_t sdx_c, int64_t sdx_r, void* stream) {
    SumArgs a{};
    int rc = fill_args(a, xprod, x, sx_b, sx_d, sx_c, sx_r, d_in, c, lw, sw_w, sw_d, sw_i, sw_o, sw_r, Wsets, w, B,
                       d_out, ic, oc, R, reduce_r, so_b, so_d, so_o, so_r);
    if (rc != SPN_OK || !out || !g || !dx) return SPN_ERR_ARG;
    if (B == 0) return SPN_OK;
    a.outc = out; a.g = g; a.dx = dx;
    a.sdx_b = sdx_b; a.sdx_d = sdx_d; a.sdx_c = sdx_c; a.sdx_r = sdx_r;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)B * (xprod ? d_in * c : d_out * ic) * R;
    if (xprod) sum_bwd_x_kernel<true><<<spn_blocks(total, 256), 256, 0, st>>>(a);
    else sum_bwd_x_kernel<false><<<spn_blocks(total, 256), 256, 0, st>>>(a);
    return spn_launch_status();
}

extern "C" int spn_xprod_fwd(const float* x, int B, int d_in, int c, int R, float* out, void* stream) {
    if (!x || !out || B < 0 || d_in <= 0 || c <= 0 || R <= 0) return SPN_ERR_ARG;
    if (B == 0) return SPN_OK;
    int P = 1;
    while (P < d_in) P <<= 1;
    const int d_out = P > 1 ? P / 2 : 1;
    const int64_t total = (int64_t)B * d_out * c * c * R;
    xprod_fwd_kernel<<<spn_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(x, B, d_in, c, R, d_out, out);
    return spn_launch_status();
}

extern "C" int spn_xprod_bwd(const float* g, int B, int d_in, int c, int R, float* dx, void* stream) {
    if (!g || !dx || B < 0 || d_in <= 0 || c <= 0 || R <= 0) return SPN_ERR_ARG;
    if (B == 0) return SPN_OK;
    int P = 1;
    while (P < d_in) P <<= 1;
    const int d_out = P > 1 ? P / 2 : 1;
    const int64_t total = (int64_t)B * d_in * c * R;
    xprod_bwd_kernel<<<spn_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(g, B, d_in, c, R, d_out, dx);
    return spn_launch_status();
}
