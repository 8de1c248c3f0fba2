// Top-down ancestral sampling kernels (index mode): one launch per Sum layer with the CrossProduct un-ravelling
// fused, plus a leaf kernel fused with the sample post-processing.
// Reference semantics: layers.py:130-237, :442-511, distributions.py:248-302, :384-393, rat_spn.py:538-580
// (SURVEY A.2).  Categorical draws are argmax(logits + Gumbel) with first-index tie break; the noise is either
// supplied (bit-exact parity tests) or generated in-kernel with Philox keyed by the global element index.
#include <stdlib.h>
#include "common.cuh"

struct SampleSumArgs {
    const float* lw; int64_t sw_w, sw_d, sw_i, sw_o, sw_r;
    int Wsets, w, Q, d, ic, oc, Rs;
    const int32_t* pa; int64_t sp_q, sp_s, sp_r; int unravel;
    const int32_t* rep;
    const float* xc; int64_t sxc_b, sxc_d, sxc_c, sxc_r; int xc_din, c, Qe;
    int root, ic_top;
    const float* gumbel; int mpe; uint64_t seed, q_offset; const uint64_t* seed_dev;
    int32_t* out;
};

// One group of G lanes (G = 4..32, the power of two covering ic) per item (q, s, rs); lanes stride over the in-channels
// and keep (best value, first index).  Out-of-range groups run on the last item with the write masked.
template <int G>
__global__ void __launch_bounds__(256) sample_sum_kernel(SampleSumArgs a) {
    const int lane = threadIdx.x & (G - 1);
    const uint64_t seed_v = a.seed_dev ? *a.seed_dev : a.seed;  // device seed: fresh noise on CUDA-graph replays
    const int64_t items = (int64_t)a.Q * a.d * a.Rs;
    int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool live = item < items;
    if (!live) item = items - 1;
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
    const float* lwp = a.lw + (int64_t)ws * a.sw_w + (int64_t)s * a.sw_d + (int64_t)o * a.sw_o;
    const int64_t qe = a.xc ? q % a.Qe : 0;
    const float* xl = nullptr;
    const float* xr = nullptr;
    if (a.xc) {
        if (2 * s < a.xc_din) xl = a.xc + qe * a.sxc_b + (int64_t)(2 * s) * a.sxc_d;
        if (2 * s + 1 < a.xc_din) xr = a.xc + qe * a.sxc_b + (int64_t)(2 * s + 1) * a.sxc_d;
    }
    // virtual noise tensor [Q_global, d, ic, Rs]; stream id separates layers (d, ic unique per layer) and calls
    const uint64_t stream_id = ((uint64_t)a.d << 40) ^ ((uint64_t)a.ic << 8) ^ (uint64_t)a.Rs;

    float best = -CUDART_INF_F;
    int besti = 0x7fffffff;
    auto logit = [&](int i) -> float {
        int r = rsel, ii = i;
        if (a.root) { r = i / a.ic_top; ii = i - r * a.ic_top; }
        float v = lwp[(int64_t)ii * a.sw_i + (int64_t)r * a.sw_r];
        if (a.xc) {
            const int l = ii / a.c, rp = ii - l * a.c;
            float off = 0.f;
            if (xl) off += xl[(int64_t)l * a.sxc_c + (int64_t)r * a.sxc_r];
            if (xr) off += xr[(int64_t)rp * a.sxc_c + (int64_t)r * a.sxc_r];
            v += off;
        }
        return v;
    };
    if (!a.mpe && !a.gumbel && a.Rs == 1 && (a.ic & 3) == 0) {
        // in-kernel noise, one repetition per item: the virtual noise elements of consecutive in-channels are consecutive,
        // so one Philox block (four words) serves four in-channels -- the same values spn_gumbel() gives one at a time
        const uint64_t base4 = ((((uint64_t)q + a.q_offset) * a.d + s) * (uint64_t)a.ic) >> 2;
        for (int i4 = lane; i4 < (a.ic >> 2); i4 += G) {
            const uint4 rnd = spn_philox4x32(base4 + i4, stream_id, seed_v);
            const uint32_t bits[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = 4 * i4 + j;
                const float e = -__logf(spn_u01(bits[j]));
                const float v = logit(i) - __logf(fmaxf(e, 1e-30f));
                if (v > best || besti == 0x7fffffff) { best = v; besti = i; }
            }
        }
    } else {
        for (int i = lane; i < a.ic; i += G) {
            float v = logit(i);
            if (!a.mpe) {
                if (a.gumbel) {
                    v += a.gumbel[((q * a.d + s) * (int64_t)a.ic + i) * a.Rs + rs];
                } else {
                    const uint64_t elem = ((((uint64_t)q + a.q_offset) * a.d + s) * (uint64_t)a.ic + i) * a.Rs + rs;
                    v += spn_gumbel(elem, stream_id, seed_v);
                }
            }
            if (v > best || besti == 0x7fffffff) { best = v; besti = i; }
        }
    }
#pragma unroll
    for (int sh = G / 2; sh > 0; sh >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, sh);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, sh);
        if (oi != 0x7fffffff && (besti == 0x7fffffff || ov > best || (ov == best && oi < besti))) { best = ov; besti = oi; }
    }
    if (lane == 0 && live) a.out[item] = besti;
}

// Variant for layers sampled in every repetition (rep == null, Rs % 4 == 0, in-kernel noise, no MPE): one group of G lanes
// serves FOUR repetitions of a (sample, scope) -- the noise elements (q, s, i, rs .. rs + 3) are consecutive, so one Philox
// block yields all four (the same values spn_gumbel() gives one at a time), and the Philox rounds, which dominate the
// instruction count of the scalar kernel, are paid once per four draws.
template <int G>
__global__ void __launch_bounds__(256) sample_sum_r4_kernel(SampleSumArgs a) {
    const int lane = threadIdx.x & (G - 1);
    const uint64_t seed_v = a.seed_dev ? *a.seed_dev : a.seed;
    const int R4 = a.Rs >> 2;
    const int64_t items = (int64_t)a.Q * a.d * R4;
    int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool live = item < items;
    if (!live) item = items - 1;
    const int r0 = (int)(item % R4) * 4;
    const int s = (int)((item / R4) % a.d);
    const int64_t q = item / ((int64_t)R4 * a.d);
    const int ws = a.Wsets == 1 ? 0 : (int)(q % a.w);
    const float* lwp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int o;
        if (a.unravel) {
            const int p = a.pa[q * a.sp_q + (int64_t)(s >> 1) * a.sp_s + (int64_t)(r0 + j) * a.sp_r];
            o = (s & 1) ? p % a.oc : p / a.oc;
        } else {
            o = a.pa[q * a.sp_q + (int64_t)s * a.sp_s + (int64_t)(r0 + j) * a.sp_r];
        }
        lwp[j] = a.lw + (int64_t)ws * a.sw_w + (int64_t)s * a.sw_d + (int64_t)o * a.sw_o + (int64_t)(r0 + j) * a.sw_r;
    }
    const int64_t qe = a.xc ? q % a.Qe : 0;
    const float* xl = nullptr;
    const float* xr = nullptr;
    if (a.xc) {
        if (2 * s < a.xc_din) xl = a.xc + qe * a.sxc_b + (int64_t)(2 * s) * a.sxc_d;
        if (2 * s + 1 < a.xc_din) xr = a.xc + qe * a.sxc_b + (int64_t)(2 * s + 1) * a.sxc_d;
    }
    const uint64_t stream_id = ((uint64_t)a.d << 40) ^ ((uint64_t)a.ic << 8) ^ (uint64_t)a.Rs;
    const uint64_t base = ((((uint64_t)q + a.q_offset) * a.d + s) * (uint64_t)a.ic) * a.Rs + r0;   // multiple of 4
    float best[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    int besti[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    for (int i = lane; i < a.ic; i += G) {
        const uint4 rnd = spn_philox4x32((base + (uint64_t)i * a.Rs) >> 2, stream_id, seed_v);
        const uint32_t bits[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
        int l = 0, rp = 0;
        if (a.xc) { l = i / a.c; rp = i - l * a.c; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = lwp[j][(int64_t)i * a.sw_i];
            if (a.xc) {
                if (xl) v += xl[(int64_t)l * a.sxc_c + (int64_t)(r0 + j) * a.sxc_r];
                if (xr) v += xr[(int64_t)rp * a.sxc_c + (int64_t)(r0 + j) * a.sxc_r];
            }
            const float e = -__logf(spn_u01(bits[j]));
            v -= __logf(fmaxf(e, 1e-30f));
            if (v > best[j] || besti[j] == 0x7fffffff) { best[j] = v; besti[j] = i; }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int sh = G / 2; sh > 0; sh >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best[j], sh);
            const int oi = __shfl_xor_sync(0xffffffffu, besti[j], sh);
            if (oi != 0x7fffffff && (besti[j] == 0x7fffffff || ov > best[j] || (ov == best[j] && oi < besti[j]))) {
                best[j] = ov;
                besti[j] = oi;
            }
        }
    }
    if (lane == 0 && live) {
        int32_t* dst = a.out + (q * a.d + s) * (int64_t)a.Rs + r0;
        dst[0] = besti[0]; dst[1] = besti[1]; dst[2] = besti[2]; dst[3] = besti[3];
    }
}

struct SampleLeafArgs {
    const float* mu; const float* sigma;
    int Wsets, w, Q, F, I, R, card, Rs;
    const int32_t* pa; int64_t sp_q, sp_s, sp_r;
    const int32_t* rep;
    const float* eps; int mpe; uint64_t seed, q_offset; const uint64_t* seed_dev;
    const int32_t* inv_perm;
    const float* evidence; int64_t se_q, se_f, se_r; int Qe;
    int squash;
    float* out; int32_t* ch_out;
};

// One thread per output element (q, feature, rs).
__global__ void __launch_bounds__(256) sample_leaf_kernel(SampleLeafArgs a) {
    const int64_t total = (int64_t)a.Q * a.F * a.Rs;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const uint64_t seed_v = a.seed_dev ? *a.seed_dev : a.seed;
    const int rs = (int)(tid % a.Rs);
    const int fo = (int)((tid / a.Rs) % a.F);
    const int64_t q = tid / ((int64_t)a.Rs * a.F);
    const int ws = a.Wsets == 1 ? 0 : (int)(q % a.w);
    const int r = a.rep ? a.rep[q] : rs;
    const int slot = a.inv_perm ? a.inv_perm[(int64_t)fo * a.R + r] : fo;
    const int s = slot / a.card;
    const int p = a.pa[q * a.sp_q + (int64_t)(s >> 1) * a.sp_s + (int64_t)rs * a.sp_r];
    const int ch = (s & 1) ? p % a.I : p / a.I;
    const int64_t pi = (((int64_t)ws * a.F + slot) * a.I + ch) * a.R + r;
    float y = a.mu[pi];
    if (!a.mpe) {
        const int64_t ne = (q * a.F + slot) * (int64_t)a.Rs + rs;
        float e;
        if (a.eps) e = a.eps[ne];
        else e = spn_normal((((uint64_t)q + a.q_offset) * a.F + slot) * (uint64_t)a.Rs + rs, 0x1eafULL, seed_v);
        y += a.sigma[pi] * e;
    }
    if (a.evidence) {
        const float ev = a.evidence[(q % a.Qe) * a.se_q + (int64_t)fo * a.se_f + (int64_t)rs * a.se_r];
        if (!isnan(ev)) y = ev;
    }
    if (a.squash) y = tanhf(fminf(fmaxf(y, -6.f), 6.f));
    a.out[tid] = y;
    if (a.ch_out) a.ch_out[(q * a.F + slot) * (int64_t)a.Rs + rs] = ch;
}

extern "C" int spn_sample_sum(const float* lw, int64_t sw_w, int64_t sw_d, int64_t sw_i, int64_t sw_o, int64_t sw_r,
                              int Wsets, int w, int Q, int d, int ic, int oc, int Rs, const int32_t* pa, int64_t sp_q,
                              int64_t sp_s, int64_t sp_r, int unravel, const int32_t* rep, const float* xc,
                              int64_t sxc_b, int64_t sxc_d, int64_t sxc_c, int64_t sxc_r, int xc_din, int c, int Qe,
                              int root, int ic_top, const float* gumbel, int mpe, uint64_t seed, uint64_t q_offset,
                              const uint64_t* seed_dev, int32_t* out, void* stream) {
    if (!lw || !pa || !out || Q < 0 || d <= 0 || ic <= 0 || oc <= 0 || Rs <= 0) return SPN_ERR_ARG;
    if (Wsets != 1 && (Wsets != w || Q % w != 0)) return SPN_ERR_ARG;
    if (xc && (Qe <= 0 || c <= 0 || Q % Qe != 0)) return SPN_ERR_ARG;
    if (root && (d != 1 || Rs != 1 || ic_top <= 0 || ic % ic_top != 0)) return SPN_ERR_ARG;
    if (Q == 0) return SPN_OK;
    SampleSumArgs a{};
    a.lw = lw; a.sw_w = sw_w; a.sw_d = sw_d; a.sw_i = sw_i; a.sw_o = sw_o; a.sw_r = sw_r;
    a.Wsets = Wsets; a.w = w; a.Q = Q; a.d = d; a.ic = ic; a.oc = oc; a.Rs = Rs;
    a.pa = pa; a.sp_q = sp_q; a.sp_s = sp_s; a.sp_r = sp_r; a.unravel = unravel; a.rep = rep;
    a.xc = xc; a.sxc_b = sxc_b; a.sxc_d = sxc_d; a.sxc_c = sxc_c; a.sxc_r = sxc_r; a.xc_din = xc_din; a.c = c; a.Qe = Qe;
    const bool one_rep_per_group = (mpe & 2) != 0;   // bit 1 of `mpe`: A/B selector of the kernel variant (same indices)
    mpe &= 1;
    a.root = root; a.ic_top = ic_top; a.gumbel = gumbel; a.mpe = mpe; a.seed = seed; a.q_offset = q_offset; a.seed_dev = seed_dev; a.out = out;
    int G = 4;
    while (G < 32 && G < ic) G <<= 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (!root && !rep && !mpe && !gumbel && Rs % 4 == 0 && !one_rep_per_group) {  // every repetition sampled, in-kernel noise:
        // 4 repetitions per group
        const unsigned blocks4 = spn_blocks((int64_t)Q * d * (Rs / 4) * G, 256);
        switch (G) {
            case 4: sample_sum_r4_kernel<4><<<blocks4, 256, 0, st>>>(a); break;
            case 8: sample_sum_r4_kernel<8><<<blocks4, 256, 0, st>>>(a); break;
            case 16: sample_sum_r4_kernel<16><<<blocks4, 256, 0, st>>>(a); break;
            default: sample_sum_r4_kernel<32><<<blocks4, 256, 0, st>>>(a); break;
        }
        return spn_launch_status();
    }
    const unsigned blocks = spn_blocks((int64_t)Q * d * Rs * G, 256);
    switch (G) {
        case 4: sample_sum_kernel<4><<<blocks, 256, 0, st>>>(a); break;
        case 8: sample_sum_kernel<8><<<blocks, 256, 0, st>>>(a); break;
        case 16: sample_sum_kernel<16><<<blocks, 256, 0, st>>>(a); break;
        default: sample_sum_kernel<32><<<blocks, 256, 0, st>>>(a); break;
    }
    return spn_launch_status();
}

extern "C" int spn_sample_leaf(const float* mu, const float* sigma, int Wsets, int w, int Q, int F, int I, int R,
                               int card, int Rs, const int32_t* pa, int64_t sp_q, int64_t sp_s, int64_t sp_r,
                               const int32_t* rep, const float* eps, int mpe, uint64_t seed, uint64_t q_offset,
                               const uint64_t* seed_dev, const int32_t* inv_perm, const float* evidence, int64_t se_q, int64_t se_f,
                               int64_t se_r, int Qe, int squash, float* out, int32_t* ch_out, void* stream) {
    if (!mu || !sigma || !pa || !out || Q < 0 || F <= 0 || I <= 0 || R <= 0 || card <= 0 || Rs <= 0) return SPN_ERR_ARG;
    if (Wsets != 1 && (Wsets != w || Q % w != 0)) return SPN_ERR_ARG;
    if (evidence && (Qe <= 0 || Q % Qe != 0)) return SPN_ERR_ARG;
    if (Q == 0) return SPN_OK;
    SampleLeafArgs a{};
    a.mu = mu; a.sigma = sigma; a.Wsets = Wsets; a.w = w; a.Q = Q; a.F = F; a.I = I; a.R = R; a.card = card; a.Rs = Rs;
    a.pa = pa; a.sp_q = sp_q; a.sp_s = sp_s; a.sp_r = sp_r; a.rep = rep; a.eps = eps; a.mpe = mpe; a.seed = seed;
    a.seed_dev = seed_dev; a.q_offset = q_offset; a.inv_perm = inv_perm; a.evidence = evidence; a.se_q = se_q; a.se_f = se_f; a.se_r = se_r;
    a.Qe = Qe; a.squash = squash; a.out = out; a.ch_out = ch_out;
    const int64_t total = (int64_t)Q * F * Rs;
    sample_leaf_kernel<<<spn_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(a);
    return spn_launch_status();
}
