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
    float* logz;                          // RAW: log-normaliser of the unnormalised weights, [B][d_out][oc][R] (output)
    int B, w, d_in, d_out, c, oc, R;
};

// grid (B, d_out); block = oc * R threads (one per output of the (row, scope) pair).
// RAW: `lw` holds the UNNORMALISED head outputs of the gating network; the log_softmax over the in-channels
// (cspn.py:274, :282) is fused: out = lse_i(x_i + raw_i) - lse_i(raw_i), both sums relative to one running maximum of the
// raw weights, and the normaliser is written out for the backward kernel.
template <bool RAW>
__global__ void __launch_bounds__(1024) sum_ps_fwd_kernel(PsArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int ocr = a.oc * a.R, c = a.c, ic = c * c;
    const int rows_per_stage = PS_STAGE_BYTES / (ocr * 4);
    float* ring = reinterpret_cast<float*>(smem);                                 // [PS_NSTG][rows_per_stage][ocr]
    float* xl = ring + PS_NSTG * rows_per_stage * ocr;                            // [c][R]
    float* xr = xl + c * a.R;                                                     // [c][R]
    uint64_t* full = reinterpret_cast<uint64_t*>(xr + c * a.R);                   // [PS_NSTG]
    const int s = blockIdx.y;
    const int64_t b = blockIdx.x;
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
    float acc = 0.f, accz = 0.f, M = -CUDART_INF_F;   // RAW: accz = sum exp2(v - M), acc = sum exp2(v - M + child terms)
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
            if (RAW) {
                const float v = wst[j * ocr] * L2E;
                if (v > M) {
                    const float sc = ex2_approx(M - v);   // exp2(-inf) = 0 on the first element
                    acc *= sc;
                    accz *= sc;
                    M = v;
                }
                const float e = ex2_approx(v - M);
                accz += e;
                acc += e * ex2_approx(fmaf(xr[rp * a.R + r], L2E, lvs));
            } else {
                const float arg = fmaf(wst[j * ocr] + xr[rp * a.R + r], L2E, lvs);
                acc += ex2_approx(arg);
            }
            if (++rp == c) {
                rp = 0;
                ++l;
                lvs = fmaf(xl[min(l, c - 1) * a.R + r], L2E, -ms);
            }
        }
        __syncthreads();  // everyone is done with this stage: refill it
        if (t == 0 && k + PS_NSTG < nstage) issue(k + PS_NSTG);
    }
    const float lz = RAW ? (M + __log2f(accz)) * 0.6931471805599453f : 0.f;   // lse_i(raw_i)
    if (RAW) a.logz[(b * a.d_out + s) * ocr + t] = lz;
    float res;
    if (dead) {
        res = -CUDART_INF_F;
    } else if (acc >= 1e-30f) {
        res = mx + __logf(acc) - (RAW ? __logf(accz) : 0.f);
    } else {
        // exact two-pass log-sum-exp over t_i = x_i + lw_i (underflow: all mass far below the child maxima)
        float m2 = -CUDART_INF_F;
        for (int i = 0; i < ic; ++i) m2 = fmaxf(m2, xl[(i / c) * a.R + r] + xr[(i % c) * a.R + r] + slab[(int64_t)i * ocr + t] - lz);
        res = m2;
        if (m2 != -CUDART_INF_F) {
            float acc2 = 0.f;
            for (int i = 0; i < ic; ++i) acc2 += expf(xl[(i / c) * a.R + r] + xr[(i % c) * a.R + r] + slab[(int64_t)i * ocr + t] - lz - m2);
            res = m2 + logf(acc2);
        }
    }
    a.out[(b * a.d_out + s) * ocr + t] = res;
}

// ---------------------------------------------------------------------------------------------------------------
// Forward for MANY rows per weight set (B = n * w with n >> 1: the recursive entropy scores K * n samples per
// conditional, rat_spn.py:661-676).  One block per (weight set, scope) walks its rows in tiles of PSR_RT: the slab
// streams through the TMA ring once per row tile (from L2 after the first), is turned into linear weights W = exp(lw)
// in place, and every thread (one per output (o, r)) keeps PSR_RT accumulators:
//     out[b,o,r] = mL + mR + log sum_l eL[b,l,r] * ( sum_r' eR[b,r',r] * W[(l,r'),o,r] )
// i.e. 2 exps per input instead of one exp per (input pair, output); the inner loop is one broadcast LDS + one FFMA.
constexpr int PSR_RT = 32;

struct PsRowsArgs {
    const float* x; int64_t sx_b, sx_d;
    const float* lw; float* out;
    int B, w, d_in, d_out, oc, R, rows_per_stage;
};

template <int C>
__global__ void __launch_bounds__(256) sum_ps_rows_fwd_kernel(PsRowsArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int ocr = a.oc * a.R, R = a.R;
    constexpr int ic = C * C;
    const int rps = a.rows_per_stage;  // multiple of C
    float* ring = reinterpret_cast<float*>(smem);                                  // [PS_NSTG][rps][ocr]
    float* eL = ring + PS_NSTG * rps * ocr;                                        // [PSR_RT][C][R]
    float* eR = eL + PSR_RT * C * R;                                               // [PSR_RT][C][R]
    float* sh = eR + PSR_RT * C * R;                                               // [PSR_RT][R]  mL + mR (or -inf: dead)
    uint64_t* full = reinterpret_cast<uint64_t*>(sh + PSR_RT * R);                 // [PS_NSTG]
    const int j = blockIdx.x, s = blockIdx.y;
    const int t = threadIdx.x, r = t % R;
    const float* slab = a.lw + (((int64_t)j * a.d_out + s) * ic) * ocr;
    const int nstage = ic / rps;
    const int nrows = a.B / a.w;       // rows of this weight set: b = j + k * w
    const bool has_l = 2 * s < a.d_in, has_r = 2 * s + 1 < a.d_in;
    if (t == 0) {
        for (int i = 0; i < PS_NSTG; ++i) mbar_init(full + i, 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int k) {  // thread 0 only; k = global stage counter over all row tiles
        const int ks = k % nstage;
        const uint32_t bytes = (uint32_t)rps * ocr * 4;
        uint64_t* bar = full + k % PS_NSTG;
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(ring + (k % PS_NSTG) * rps * ocr, slab + (int64_t)ks * rps * ocr, bytes, bar);
    };
    const int ntile = (nrows + PSR_RT - 1) / PSR_RT;
    const int total_stages = ntile * nstage;
    if (t == 0)
        for (int k = 0; k < min(PS_NSTG, total_stages); ++k) issue(k);
    int kg = 0;  // global stage counter
    for (int tile = 0; tile < ntile; ++tile) {
        const int row0 = tile * PSR_RT;
        const int nr = min(PSR_RT, nrows - row0);
        __syncthreads();  // previous tile's eL / eR / sh are no longer read
        // exp(x - max) of both children for the rows of this tile: thread e handles (row, r) pairs
        for (int e = t; e < PSR_RT * R; e += blockDim.x) {
            const int row = e / R, rr = e % R;
            float ml = -CUDART_INF_F, mr = -CUDART_INF_F;
            const int64_t b = (int64_t)j + (int64_t)(row0 + min(row, nr - 1)) * a.w;
            const float* xl = a.x + b * a.sx_b + (int64_t)(2 * s) * a.sx_d + rr;
            const float* xr = a.x + b * a.sx_b + (int64_t)(2 * s + 1) * a.sx_d + rr;
            for (int q = 0; q < C; ++q) {
                if (has_l) ml = fmaxf(ml, xl[q * R]);
                if (has_r) mr = fmaxf(mr, xr[q * R]);
            }
            if (!has_l) ml = 0.f;
            if (!has_r) mr = 0.f;
            const bool dead = ml == -CUDART_INF_F || mr == -CUDART_INF_F;
            for (int q = 0; q < C; ++q) {
                eL[(row * C + q) * R + rr] = dead ? 0.f : (has_l ? __expf(xl[q * R] - ml) : 1.f);
                eR[(row * C + q) * R + rr] = dead ? 0.f : (has_r ? __expf(xr[q * R] - mr) : 1.f);
            }
            sh[row * R + rr] = dead ? -CUDART_INF_F : ml + mr;
        }
        __syncthreads();
        float acc[PSR_RT];
#pragma unroll
        for (int q = 0; q < PSR_RT; ++q) acc[q] = 0.f;
        int l = 0;
        for (int ks = 0; ks < nstage; ++ks, ++kg) {
            while (!mbar_try_wait(full + kg % PS_NSTG, (kg / PS_NSTG) & 1)) {}
            float* wst = ring + (kg % PS_NSTG) * rps * ocr + t;
            for (int j0 = 0; j0 < rps; j0 += C, ++l) {
                float tl[PSR_RT];
#pragma unroll
                for (int q = 0; q < PSR_RT; ++q) tl[q] = 0.f;
#pragma unroll 4
                for (int rp = 0; rp < C; ++rp) {
                    const float wv = __expf(wst[(j0 + rp) * ocr]);
                    const float* er = eR + rp * R + r;
#pragma unroll
                    for (int q = 0; q < PSR_RT; ++q) tl[q] = fmaf(er[q * C * R], wv, tl[q]);
                }
                const float* el = eL + l * R + r;
#pragma unroll
                for (int q = 0; q < PSR_RT; ++q) acc[q] = fmaf(el[q * C * R], tl[q], acc[q]);
            }
            __syncthreads();  // everyone is done with this stage: refill it
            if (t == 0 && kg + PS_NSTG < total_stages) issue(kg + PS_NSTG);
        }
#pragma unroll
        for (int q = 0; q < PSR_RT; ++q) {
            if (q < nr) {
                const int64_t b = (int64_t)j + (int64_t)(row0 + q) * a.w;
                const float shift = sh[q * R + r];
                float res;
                if (shift == -CUDART_INF_F) {
                    res = -CUDART_INF_F;
                } else if (acc[q] >= 1e-30f) {
                    res = shift + __logf(acc[q]);
                } else {
                    // exact two-pass log-sum-exp (underflow: all mass far below the child maxima)
                    const float* xl = a.x + b * a.sx_b + (int64_t)(2 * s) * a.sx_d + r;
                    const float* xr = a.x + b * a.sx_b + (int64_t)(2 * s + 1) * a.sx_d + r;
                    float m2 = -CUDART_INF_F;
                    for (int i = 0; i < ic; ++i)
                        m2 = fmaxf(m2, (has_l ? xl[(i / C) * R] : 0.f) + (has_r ? xr[(i % C) * R] : 0.f) + slab[(int64_t)i * ocr + t]);
                    res = m2;
                    if (m2 != -CUDART_INF_F) {
                        float acc2 = 0.f;
                        for (int i = 0; i < ic; ++i)
                            acc2 += expf((has_l ? xl[(i / C) * R] : 0.f) + (has_r ? xr[(i % C) * R] : 0.f) + slab[(int64_t)i * ocr + t] - m2);
                        res = m2 + logf(acc2);
                    }
                }
                a.out[(b * a.d_out + s) * ocr + t] = res;
            }
        }
    }
}

template <int C>
static int launch_ps_rows(const PsRowsArgs& a, cudaStream_t st) {
    const int ocr = a.oc * a.R;
    const int smem = PS_NSTG * a.rows_per_stage * ocr * 4 + 2 * PSR_RT * C * a.R * 4 + PSR_RT * a.R * 4 + PS_NSTG * 8 + 16;
    cudaError_t e = cudaFuncSetAttribute(sum_ps_rows_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return spn_cuda_status(e);
    sum_ps_rows_fwd_kernel<C><<<dim3(a.w, a.d_out), ocr, smem, st>>>(a);
    return spn_launch_status();
}

// ---------------------------------------------------------------------------------------------------------------
// Fused backward for one row per weight set (B == w: the CSPN log-likelihood training step).  The slab of a
// (conditional, scope) is read ONCE (TMA ring, as in the forward) and its gradient written ONCE (staged in shared memory,
// TMA bulk store); the activation gradients of the two children fall out of the same pass:
//     dlw[i,o,r] = g[o,r] * exp(xL[l,r] + xR[r',r] + lw[i,o,r] - out[o,r]),   i = l*c + r'
//     dxL[l,r]   = sum_{r',o} dlw[(l,r'),o,r],      dxR[r',r] = sum_{l,o} dlw[(l,r'),o,r]
// One thread per (o, r).  Per-thread partial sums (one for the current l, C for the r') are reduced over o with two
// shuffles inside the warp and shared-memory atomics across warps.  Autograd of layers.py:110-128 + :419-440 for
// per-sample weights.
struct PsBwdArgs {
    const float* x; int64_t sx_b, sx_d;
    const float* logz;   // RAW: lse_i(raw_i) from the forward, [B][d_out][oc][R]
    const float* lw; const float* out; const float* g;
    float* dlw; float* dx; int64_t sdx_b, sdx_d;
    int w, d_in, d_out, oc, R, rows_per_stage;
};

constexpr int PSB_NIN = 3;   // slab stages in flight (loads)
constexpr int PSB_NOUT = 2;  // gradient stages (stores)

// RAW: `lw` holds unnormalised weights; d raw[i] = g * (p_i - w_i) with p_i = exp(x_i + raw_i - logz - out) (responsibility)
// and w_i = exp(raw_i - logz) (the softmax backward folded in: sum_i dlw_i = g).
template <int C, bool RAW>
__global__ void __launch_bounds__(1024) sum_ps_bwd_kernel(PsBwdArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int ocr = a.oc * a.R, R = a.R;
    constexpr int ic = C * C;
    const int rps = a.rows_per_stage;                    // multiple of C
    float* in_ring = reinterpret_cast<float*>(smem);                              // [PSB_NIN][rps][ocr]
    float* out_ring = in_ring + PSB_NIN * rps * ocr;                              // [PSB_NOUT][rps][ocr]
    float* xl = out_ring + PSB_NOUT * rps * ocr;                                  // [C][R]
    float* xr = xl + C * R;                                                       // [C][R]
    float* dxl = xr + C * R;                                                      // [C][R]
    float* dxr = dxl + C * R;                                                     // [C][R]
    uint64_t* full = reinterpret_cast<uint64_t*>(dxr + C * R);                    // [PSB_NIN]
    const int s = blockIdx.y;
    const int64_t b = blockIdx.x;
    const int t = threadIdx.x, lane = t & 31;
    const int r = t % R;
    const int64_t slab_off = ((b * a.d_out + s) * (int64_t)ic) * ocr;
    const float* slab = a.lw + slab_off;
    float* gslab = a.dlw + slab_off;
    const int nstage = ic / rps;

    if (t == 0) {
        for (int i = 0; i < PSB_NIN; ++i) mbar_init(full + i, 1);
        mbar_fence_init();
    }
    const bool has_l = 2 * s < a.d_in, has_r = 2 * s + 1 < a.d_in;
    for (int e = t; e < C * R; e += blockDim.x) {
        xl[e] = has_l ? a.x[b * a.sx_b + (int64_t)(2 * s) * a.sx_d + e] : 0.f;
        xr[e] = has_r ? a.x[b * a.sx_b + (int64_t)(2 * s + 1) * a.sx_d + e] : 0.f;
        dxl[e] = 0.f;
        dxr[e] = 0.f;
    }
    __syncthreads();
    auto issue = [&](int k) {  // thread 0 only
        const uint32_t bytes = (uint32_t)rps * ocr * 4;
        uint64_t* bar = full + k % PSB_NIN;
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(in_ring + (k % PSB_NIN) * rps * ocr, slab + (int64_t)k * rps * ocr, bytes, bar);
    };
    if (t == 0)
        for (int k = 0; k < min(PSB_NIN, nstage); ++k) issue(k);

    constexpr float L2E = 1.4426950408889634f;
    const float ov = a.out[(b * a.d_out + s) * ocr + t];
    float gv = a.g[(b * a.d_out + s) * ocr + t];
    const bool dead = ov == -CUDART_INF_F;   // impossible output: no gradient (th.logsumexp backward gives 0 * nan -> we give 0)
    if (dead) gv = 0.f;
    const float zs = RAW ? a.logz[(b * a.d_out + s) * ocr + t] * L2E : 0.f;
    const float os = (dead ? 0.f : ov * L2E) + zs;
    float xrs[C];                            // right-child terms of this repetition, pre-scaled
#pragma unroll
    for (int q = 0; q < C; ++q) xrs[q] = xr[q * R + r] * L2E;
    float acc_r[C];
#pragma unroll
    for (int q = 0; q < C; ++q) acc_r[q] = 0.f;
    // reduction over o of a per-thread value: lanes with the same r inside a warp are `R` apart
    const int same_r_in_warp = 32 / R;       // R divides 32 (checked on the host)
    auto reduce_o_and_add = [&](float v, float* dst) {
        for (int sh = R; sh < 32; sh <<= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
        if (lane < R) atomicAdd(dst, v);
        (void)same_r_in_warp;
    };
    int l = 0;
    for (int k = 0; k < nstage; ++k) {
        while (!mbar_try_wait(full + k % PSB_NIN, (k / PSB_NIN) & 1)) {}
        const float* wst = in_ring + (k % PSB_NIN) * rps * ocr + t;
        float* ost = out_ring + (k % PSB_NOUT) * rps * ocr + t;
        if (k >= PSB_NOUT) {
            // the store that used this gradient stage two iterations ago must have read it
            if (t == 0) bulk_wait_group_read<PSB_NOUT - 1>();
            __syncthreads();
        }
        for (int j0 = 0; j0 < rps; j0 += C, ++l) {
            const float lvs = fmaf(xl[l * R + r], L2E, -os);
            float acc_l = 0.f;
#pragma unroll
            for (int q = 0; q < C; ++q) {
                const float wv = wst[(j0 + q) * ocr];
                const float val = gv * ex2_approx(fmaf(wv, L2E, lvs + xrs[q]));
                ost[(j0 + q) * ocr] = RAW ? val - gv * ex2_approx(fmaf(wv, L2E, -zs)) : val;
                acc_l += val;
                acc_r[q] += val;
            }
            reduce_o_and_add(acc_l, dxl + l * R + r);
        }
        fence_proxy_async();   // the staged gradients are read by the TMA engine
        __syncthreads();       // everyone is done with both stages
        if (t == 0) {
            bulk_s2g(gslab + (int64_t)k * rps * ocr, out_ring + (k % PSB_NOUT) * rps * ocr, (uint32_t)rps * ocr * 4);
            bulk_commit_group();
            if (k + PSB_NIN < nstage) issue(k + PSB_NIN);
        }
    }
#pragma unroll
    for (int q = 0; q < C; ++q) reduce_o_and_add(acc_r[q], dxr + q * R + r);
    if (t == 0) bulk_wait_group<0>();
    __syncthreads();
    for (int e = t; e < C * R; e += blockDim.x) {
        if (has_l) a.dx[b * a.sdx_b + (int64_t)(2 * s) * a.sdx_d + e] = dxl[e];
        if (has_r) a.dx[b * a.sdx_b + (int64_t)(2 * s + 1) * a.sdx_d + e] = dxr[e];
    }
}

template <int C, bool RAW>
static int launch_ps_bwd2(const PsBwdArgs& a, int B, cudaStream_t st) {
    const int ocr = a.oc * a.R;
    const int smem = (PSB_NIN + PSB_NOUT) * a.rows_per_stage * ocr * 4 + 4 * C * a.R * 4 + PSB_NIN * 8 + 16;
    cudaError_t e = cudaFuncSetAttribute(sum_ps_bwd_kernel<C, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return spn_cuda_status(e);
    sum_ps_bwd_kernel<C, RAW><<<dim3(B, a.d_out), ocr, smem, st>>>(a);
    return spn_launch_status();
}
template <int C>
static int launch_ps_bwd(const PsBwdArgs& a, int B, cudaStream_t st) {
    return a.logz ? launch_ps_bwd2<C, true>(a, B, st) : launch_ps_bwd2<C, false>(a, B, st);
}

static int ps_bwd_rows_per_stage(int c, int ocr) {
    int m = PS_STAGE_BYTES / (c * ocr * 4);
    if (m < 1) m = 1;
    while (m > 1 && c % m != 0) --m;   // the stages must tile the c*c rows of a slab
    return c * m;
}

}  // namespace

extern "C" int spn_sum_ps_bwd_supported(int c, int oc, int R);

extern "C" int spn_sum_ps_supported(int c, int oc, int R) {
    const int ocr = oc * R;
    return (ocr >= 32 && ocr <= 1024 && ocr % 4 == 0 && PS_STAGE_BYTES / (ocr * 4) >= 1 && c >= 1 && c <= 256) ? 1 : 0;
}

static int ps_fwd_rowwise(const float* x, int64_t sx_b, int64_t sx_d, int d_in, int c, const float* lw, int w, int B,
                          int d_out, int oc, int R, float* out, float* logz, cudaStream_t st) {
    PsArgs a{};
    a.x = x; a.sx_b = sx_b; a.sx_d = sx_d; a.lw = lw; a.out = out; a.logz = logz;
    a.B = B; a.w = w; a.d_in = d_in; a.d_out = d_out; a.c = c; a.oc = oc; a.R = R;
    const int ocr = oc * R;
    const int rows = PS_STAGE_BYTES / (ocr * 4);
    const int smem = PS_NSTG * rows * ocr * 4 + 2 * c * R * 4 + PS_NSTG * 8 + 16;
    if (logz) {
        cudaError_t e = cudaFuncSetAttribute(sum_ps_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return spn_cuda_status(e);
        sum_ps_fwd_kernel<true><<<dim3(B, d_out), ocr, smem, st>>>(a);
    } else {
        cudaError_t e = cudaFuncSetAttribute(sum_ps_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return spn_cuda_status(e);
        sum_ps_fwd_kernel<false><<<dim3(B, d_out), ocr, smem, st>>>(a);
    }
    return spn_launch_status();
}

extern "C" int spn_sum_ps_fwd(const float* x, int64_t sx_b, int64_t sx_d, int d_in, int c, const float* lw, int w, int B,
                              int d_out, int oc, int R, float* out, void* stream) {
    if (!x || !lw || !out || B < 0 || w <= 0 || B % w != 0 || d_in <= 0 || d_in > 2 * d_out) return SPN_ERR_ARG;
    if (!spn_sum_ps_supported(c, oc, R)) return SPN_ERR_UNSUPPORTED;
    if (B == 0) return SPN_OK;
    if (d_out > 65535) return SPN_ERR_UNSUPPORTED;
    // many rows per weight set: block per (weight set, scope), linear-domain accumulation (see sum_ps_rows_fwd_kernel)
    if (B / w >= 16 && oc * R <= 256 && spn_sum_ps_bwd_supported(c, oc, R) &&
        (size_t)PS_NSTG * ps_bwd_rows_per_stage(c, oc * R) * oc * R * 4 + 2 * PSR_RT * c * R * 4 <= 200 * 1024) {
        PsRowsArgs ra{};
        ra.x = x; ra.sx_b = sx_b; ra.sx_d = sx_d; ra.lw = lw; ra.out = out;
        ra.B = B; ra.w = w; ra.d_in = d_in; ra.d_out = d_out; ra.oc = oc; ra.R = R;
        ra.rows_per_stage = ps_bwd_rows_per_stage(c, oc * R);
        cudaStream_t st = (cudaStream_t)stream;
        if (c == 8) return launch_ps_rows<8>(ra, st);
        if (c == 16) return launch_ps_rows<16>(ra, st);
        return launch_ps_rows<32>(ra, st);
    }
    return ps_fwd_rowwise(x, sx_b, sx_d, d_in, c, lw, w, B, d_out, oc, R, out, nullptr, (cudaStream_t)stream);
}

// Same with UNNORMALISED weights (the gating network's raw head outputs): F.log_softmax over the in-channels
// (cspn.py:274, :282) fused; logz [B, d_out, oc, R] receives lse_i(raw_i) for spn_sum_ps_bwd_raw.  One row per weight set.
extern "C" int spn_sum_ps_fwd_raw(const float* x, int64_t sx_b, int64_t sx_d, int d_in, int c, const float* raw, int w, int B,
                                  int d_out, int oc, int R, float* out, float* logz, void* stream) {
    if (!x || !raw || !out || !logz || B < 0 || w <= 0 || d_in <= 0 || d_in > 2 * d_out) return SPN_ERR_ARG;
    if (B != w) return SPN_ERR_UNSUPPORTED;
    if (!spn_sum_ps_supported(c, oc, R) || d_out > 65535) return SPN_ERR_UNSUPPORTED;
    if (B == 0) return SPN_OK;
    return ps_fwd_rowwise(x, sx_b, sx_d, d_in, c, raw, w, B, d_out, oc, R, out, logz, (cudaStream_t)stream);
}

extern "C" int spn_sum_ps_bwd_supported(int c, int oc, int R) {
    const int ocr = oc * R;
    if (!(c == 8 || c == 16 || c == 32)) return 0;
    if (ocr < 32 || ocr > 1024 || ocr % 32 != 0 || R > 32 || (32 % R) != 0) return 0;
    const int rps = ps_bwd_rows_per_stage(c, ocr);
    if ((size_t)(PSB_NIN + PSB_NOUT) * rps * ocr * 4 > 160 * 1024) return 0;
    return 1;
}

extern "C" int spn_sum_ps_bwd(const float* x, int64_t sx_b, int64_t sx_d, int d_in, int c, const float* lw, int w, int B,
                              int d_out, int oc, int R, const float* out, const float* g, float* dlw, float* dx,
                              int64_t sdx_b, int64_t sdx_d, void* stream) {
    if (!x || !lw || !out || !g || !dlw || !dx || B < 0 || w <= 0 || d_in <= 0 || d_in > 2 * d_out) return SPN_ERR_ARG;
    if (B != w) return SPN_ERR_UNSUPPORTED;   // one row per weight set (several rows would have to accumulate into dlw)
    if (!spn_sum_ps_bwd_supported(c, oc, R)) return SPN_ERR_UNSUPPORTED;
    if (B == 0) return SPN_OK;
    if (d_out > 65535) return SPN_ERR_UNSUPPORTED;
    PsBwdArgs a{};
    a.x = x; a.sx_b = sx_b; a.sx_d = sx_d; a.lw = lw; a.out = out; a.g = g; a.dlw = dlw; a.dx = dx;
    a.sdx_b = sdx_b; a.sdx_d = sdx_d; a.w = w; a.d_in = d_in; a.d_out = d_out; a.oc = oc; a.R = R;
    a.rows_per_stage = ps_bwd_rows_per_stage(c, oc * R);
    cudaStream_t st = (cudaStream_t)stream;
    if (c == 8) return launch_ps_bwd<8>(a, B, st);
    if (c == 16) return launch_ps_bwd<16>(a, B, st);
    return launch_ps_bwd<32>(a, B, st);
}

// Backward of spn_sum_ps_fwd_raw: draw [w, d_out, c*c, oc, R] is the gradient w.r.t. the UNNORMALISED weights (softmax
// backward fused), dx as in spn_sum_ps_bwd.
extern "C" int spn_sum_ps_bwd_raw(const float* x, int64_t sx_b, int64_t sx_d, int d_in, int c, const float* raw, int w, int B,
                                  int d_out, int oc, int R, const float* logz, const float* out, const float* g, float* draw,
                                  float* dx, int64_t sdx_b, int64_t sdx_d, void* stream) {
    if (!x || !raw || !logz || !out || !g || !draw || !dx || B < 0 || w <= 0 || d_in <= 0 || d_in > 2 * d_out) return SPN_ERR_ARG;
    if (B != w) return SPN_ERR_UNSUPPORTED;
    if (!spn_sum_ps_bwd_supported(c, oc, R) || d_out > 65535) return SPN_ERR_UNSUPPORTED;
    if (B == 0) return SPN_OK;
    PsBwdArgs a{};
    a.x = x; a.sx_b = sx_b; a.sx_d = sx_d; a.lw = raw; a.logz = logz; a.out = out; a.g = g; a.dlw = draw; a.dx = dx;
    a.sdx_b = sdx_b; a.sdx_d = sdx_d; a.w = w; a.d_in = d_in; a.d_out = d_out; a.oc = oc; a.R = R;
    a.rows_per_stage = ps_bwd_rows_per_stage(c, oc * R);
    cudaStream_t st = (cudaStream_t)stream;
    if (c == 8) return launch_ps_bwd<8>(a, B, st);
    if (c == 16) return launch_ps_bwd<16>(a, B, st);
    return launch_ps_bwd<32>(a, B, st);
}
