// Per-sample-weight (CSPN) CrossProduct+Sum forward, HBM-streaming with TMA staging.
//
// A CSPN gives every conditional its own weight set (cspn.py:254-301): Sum weights are [w, d, ic, oc, R] and every weight
// is used by exactly one row (per n-sample), so the layer is a batched contraction with arithmetic intensity ~0.5 flop/B:
// bound by HBM, not by the tensor cores.  For one (row b, scope s) the weights are ONE contiguous slab [ic][oc][R]
// (131 KB at ic = 256, oc = 16, R = 8), which the TMA engine (cp.async.bulk) streams through a 4-deep shared-memory ring
// while the threads -- one per output (o, r) -- accumulate exp(xL[l,r] + xR[r',r] + lw - max) over the in-channels.
// Reads are fully sequential, every byte is fetched once, and ~48-64 KB per block are in flight, which is what keeps the
// HBM pipes full (the general kernel in sum.cu issues dependent scalar loads and reaches ~30 % of the bandwidth).
// Reference semantics: layers.py:110-128 + :419-440; exact fp32, th.logsumexp semantics incl. -inf rows.
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int PS_NSTG = 4;          // ring depth
constexpr int PS_STAGE_BYTES = 16384;

struct PsArgs {
    const float* x; int64_t sx_b, sx_d;   // x [B][d_in][c][R] (c, R contiguous)
    const float* lw;                      // [w][d_out][ic][oc][R] contiguous
    float* out;                           // [B][d_out][oc][R] contiguous
    int B, w, d_in, d_out, c, oc, R;
};

// grid (d_out, B); block = oc * R threads (one per output of the (row, scope) pair)
__global__ void __launch_bounds__(1024) sum_ps_fwd_kernel(PsArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int ocr = a.oc * a.R, c = a.c, ic = c * c;
    const int rows_per_stage = PS_STAGE_BYTES / (ocr * 4);
    float* ring = reinterpret_cast<float*>(smem);                                 // [PS_NSTG][rows_per_stage][ocr]
    float* xl = ring + PS_NSTG * rows_per_stage * ocr;                            // [c][R]
    float* xr = xl + c * a.R;                                                     // [c][R]
    uint64_t* full = reinterpret_cast<uint64_t*>(xr + c * a.R);                   // [PS_NSTG]
    const int s = blockIdx.x;
    const int64_t b = blockIdx.y;
    const int t = threadIdx.x;
    const int r = t % a.R;
    const float* slab = a.lw + (((int64_t)(b % a.w) * a.d_out + s) * ic) * ocr;
    const int nstage = (ic + rows_per_stage - 1) / rows_per_stage;

    if (t == 0) {
        for (int i = 0; i < PS_NSTG; ++i) mbar_init(full + i, 1);
        mbar_fence_init();
    }
    // child activations of this (row, scope): [c][R] each, zero for the padded scope
    for (int e = t; e < c * a.R; e += blockDim.x) {
        xl[e] = 2 * s < a.d_in ? a.x[b * a.sx_b + (int64_t)(2 * s) * a.sx_d + e] : 0.f;
        xr[e] = 2 * s + 1 < a.d_in ? a.x[b * a.sx_b + (int64_t)(2 * s + 1) * a.sx_d + e] : 0.f;
    }
    __syncthreads();
    auto issue = [&](int k) {  // thread 0 only
        const int row0 = k * rows_per_stage;
        const int nrow = min(rows_per_stage, ic - row0);
        const uint32_t bytes = (uint32_t)nrow * ocr * 4;
        uint64_t* bar = full + k % PS_NSTG;
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(ring + (k % PS_NSTG) * rows_per_stage * ocr, slab + (int64_t)row0 * ocr, bytes, bar);
    };
    if (t == 0)
        for (int k = 0; k < min(PS_NSTG, nstage); ++k) issue(k);

    // shift: max over the children of this repetition (every term exp(xL + xR + lw - mx) is <= 1 because lw <= 0)
    float ml = -CUDART_INF_F, mr = -CUDART_INF_F;
    for (int l = 0; l < c; ++l) {
        ml = fmaxf(ml, xl[l * a.R + r]);
        mr = fmaxf(mr, xr[l * a.R + r]);
    }
    const float mx = ml + mr;
    const bool dead = mx == -CUDART_INF_F;
    constexpr float L2E = 1.4426950408889634f;
    const float ms = dead ? 0.f : mx * L2E;
    // the right-child terms of this repetition, pre-scaled for ex2: thread-private column of a [c][blockDim] array is not
    // needed -- xr is re-used in place (all threads of a repetition read the same value: broadcast)
    float acc = 0.f;
    int l = 0, rp = 0;
    float lvs = fmaf(xl[r], L2E, -ms);
    for (int k = 0; k < nstage; ++k) {
        const int row0 = k * rows_per_stage;
        const int nrow = min(rows_per_stage, ic - row0);
        // wait for the stage (spin; the copy is a few microseconds at most)
        while (!mbar_try_wait(full + k % PS_NSTG, (k / PS_NSTG) & 1)) {}
        const float* wst = ring + (k % PS_NSTG) * rows_per_stage * ocr + t;
#pragma unroll 4
        for (int j = 0; j < nrow; ++j) {
            const float arg = fmaf(wst[j * ocr] + xr[rp * a.R + r], L2E, lvs);
            acc += ex2_approx(arg);
            if (++rp == c) {
                rp = 0;
                ++l;
                lvs = fmaf(xl[min(l, c - 1) * a.R + r], L2E, -ms);
            }
        }
        __syncthreads();  // everyone is done with this stage: refill it
        if (t == 0 && k + PS_NSTG < nstage) issue(k + PS_NSTG);
    }
    float res;
    if (dead) {
        res = -CUDART_INF_F;
    } else if (acc >= 1e-30f) {
        res = mx + __logf(acc);
    } else {
        // exact two-pass log-sum-exp over t_i = x_i + lw_i (underflow: all mass far below the child maxima)
        float m2 = -CUDART_INF_F;
        for (int i = 0; i < ic; ++i) m2 = fmaxf(m2, xl[(i / c) * a.R + r] + xr[(i % c) * a.R + r] + slab[(int64_t)i * ocr + t]);
        res = m2;
        if (m2 != -CUDART_INF_F) {
            float acc2 = 0.f;
            for (int i = 0; i < ic; ++i) acc2 += expf(xl[(i / c) * a.R + r] + xr[(i % c) * a.R + r] + slab[(int64_t)i * ocr + t] - m2);
            res = m2 + logf(acc2);
        }
    }
    a.out[(b * a.d_out + s) * ocr + t] = res;
}

}  // namespace

extern "C" int spn_sum_ps_supported(int c, int oc, int R) {
    const int ocr = oc * R;
    return (ocr >= 32 && ocr <= 1024 && ocr % 4 == 0 && PS_STAGE_BYTES / (ocr * 4) >= 1 && c >= 1 && c <= 256) ? 1 : 0;
}

extern "C" int spn_sum_ps_fwd(const float* x, int64_t sx_b, int64_t sx_d, int d_in, int c, const float* lw, int w, int B,
                              int d_out, int oc, int R, float* out, void* stream) {
    if (!x || !lw || !out || B < 0 || w <= 0 || B % w != 0 || d_in <= 0 || d_in > 2 * d_out) return SPN_ERR_ARG;
    if (!spn_sum_ps_supported(c, oc, R)) return SPN_ERR_UNSUPPORTED;
    if (B == 0) return SPN_OK;
    if (B > 65535) return SPN_ERR_UNSUPPORTED;
    PsArgs a{};
    a.x = x; a.sx_b = sx_b; a.sx_d = sx_d; a.lw = lw; a.out = out;
    a.B = B; a.w = w; a.d_in = d_in; a.d_out = d_out; a.c = c; a.oc = oc; a.R = R;
    const int ocr = oc * R;
    const int rows = PS_STAGE_BYTES / (ocr * 4);
    const int smem = PS_NSTG * rows * ocr * 4 + 2 * c * R * 4 + PS_NSTG * 8 + 16;
    cudaError_t e = cudaFuncSetAttribute(sum_ps_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return spn_cuda_status(e);
    sum_ps_fwd_kernel<<<dim3(d_out, B), ocr, smem, (cudaStream_t)stream>>>(a);
    return spn_launch_status();
}
