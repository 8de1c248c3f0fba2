// Shared-weight (RatSpn) CrossProduct+Sum forward on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// For one sum scope s and repetition r the layer is a dense contraction over K = c*c in-channels:
//     out[b, o] = mL[b] + mR[b] + log( sum_{l,r'} eL[b,l] * eR[b,r'] * W[l*c + r', o] ),
//     eL = exp(xL - mL), eR = exp(xR - mR), W = exp(log-weights)            (SURVEY 0 "numerics de-risk", A.1)
// i.e. a [128 x K] x [K x N] GEMM per batch tile whose A operand is an outer product that never exists in memory.
// Mapping to the hardware:
//   * B operand: W(s, r) as bf16 "core-matrix image" [K/8][N/8][8][8], copied once per work item into shared memory
//     by the TMA engine (cp.async.bulk) and consumed through an MN-major, no-swizzle UMMA descriptor.
//   * A operand: produced by CUDA cores directly into TENSOR MEMORY (tcgen05.st), one bf16x2 multiply per two
//     K-elements; it never touches shared memory (an smem-staged A would need 170 B/clk/SM of smem writes at N = 48,
//     more than the 128 B/clk an SM has).  K order is k = l*c + r', two l's per pipeline stage.
//   * accumulator: fp32 in TMEM, read back with tcgen05.ld for the epilogue (log, shift, store).
//   * roles: two producer/epilogue warpgroups (one batch tile each, ping-pong) + one MMA-issuing warp; mbarrier
//     pipelines full/empty per A stage and one "accumulator ready" barrier per warpgroup.
// Activations use the internal layout [scope][repetition][batch][channel] so that one thread reads / writes one
// contiguous row.  Outputs whose linear-domain sum underflows are recomputed exactly in fp32 (same rescue rule as
// the general kernel in sum.cu), so results keep th.logsumexp semantics.
#include <stdlib.h>

#include "tc_common.cuh"

using namespace tc;

__device__ int g_tc_aborts = 0;
// SPN_TC_DEBUG=9: clock64 deltas per phase, accumulated by one thread per warpgroup of CTA 0 (spn_debug_tc_prof)
__device__ unsigned long long g_tc_prof[16];
#define TC_PROF_T(var) long long var = prof ? clock64() : 0
#define TC_PROF_ADD(slot, t0, t1) do { if (prof) atomicAdd(&g_tc_prof[slot], (unsigned long long)((t1) - (t0))); } while (0)

struct TcFwdArgs {
    const float* x;        // [d_in][R][B][C]
    const uint8_t* wimg;   // [d_out][R][K/8][NPAD/8][8][8] bf16
    const float* lw;       // fp32 log-weights [1][d_out][K][OC][R] (exact rescue path)
    const float* logz;     // NULL, or [d_out][OC][R]: lw is unnormalised and the log-weights are lw - logz
    float* out;            // [d_out][R][B][OC]
    int B, d_in, d_out, R, OC;
    int ntiles, nsplit;    // batch tiles of 128 rows; splits of the tile range per (s, r)
    int debug;             // profiling aid: 1 = skip MMA issue, 2 = skip A generation (results invalid)
};

// exact fp32 log-sum-exp for one output (rare rescue path)
template <int C>
__device__ __noinline__ float tc_exact_lse(const float* xl, const float* xr, const float* lwp, int64_t sw_i) {
    float m = -CUDART_INF_F;
    for (int l = 0; l < C; ++l)
        for (int rp = 0; rp < C; ++rp) {
            const float t = (xl ? xl[l] : 0.f) + (xr ? xr[rp] : 0.f) + lwp[(int64_t)(l * C + rp) * sw_i];
            m = fmaxf(m, t);
        }
    if (m == -CUDART_INF_F) return m;
    float acc = 0.f;
    for (int l = 0; l < C; ++l)
        for (int rp = 0; rp < C; ++rp) {
            const float t = (xl ? xl[l] : 0.f) + (xr ? xr[rp] : 0.f) + lwp[(int64_t)(l * C + rp) * sw_i];
            acc += expf(t - m);
        }
    return m + logf(acc);
}

template <int C, int OC, int NPAD, int NST, int LPS>
__global__ void __launch_bounds__(320, 1) sum_tc_fwd_kernel(TcFwdArgs a) {
    constexpr int NG = NPAD / 8;
    constexpr int K = C * C;
    constexpr int W_BYTES = K * NPAD * 2;
    constexpr int STEPS = C / LPS;            // pipeline stages per tile: LPS l's (LPS*C k-values) each
    constexpr int MMAS_PER_STAGE = LPS * C / 16;
    constexpr int ACC_COLS = 64;              // accumulator slot per warpgroup (NPAD <= 64)
    constexpr int A_COLS = LPS * C / 2;       // LPS*C bf16 = LPS*C/2 32-bit columns per stage
    constexpr int A_BASE = 2 * ACC_COLS;
    static_assert(C % 8 == 0 && NPAD % 16 == 0 && NPAD <= 64 && OC <= NPAD && OC % 4 == 0 && LPS % 2 == 0 && C % LPS == 0,
                  "unsupported shape");
    static_assert(A_BASE + 2 * NST * A_COLS <= 512, "TMEM budget exceeded");
    constexpr uint32_t IDESC = instr_desc_bf16(128, NPAD, 0, 1);

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* wsm = smem;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + W_BYTES);  // [2][NST]
    uint64_t* bar_empty = bar_full + 2 * NST;                          // [2][NST]
    uint64_t* bar_acc = bar_empty + 2 * NST;                           // [2]
    uint64_t* bar_w = bar_acc + 2;                                     // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 1);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * NST; ++i) {
            mbar_init(bar_full + i, 4);  // one arrival per producer warp
            mbar_init(bar_empty + i, 1);
        }
        mbar_init(bar_acc + 0, 1);
        mbar_init(bar_acc + 1, 1);
        mbar_init(bar_w, 1);
        *abort_flag = 0;
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    // The whole TMEM (512 columns) of this SM is ours, so the allocation starts at lane 0 / column 0.  Using the
    // constant keeps every MMA operand warp-uniform for ptxas; anything else is reported through the abort flag.
    if (threadIdx.x == 0 && *tmem_slot != 0) *abort_flag = 1;
    constexpr uint32_t tmem_base = 0;
    __syncthreads();

    const int nitems = a.d_out * a.R * a.nsplit;
    const int tiles_per_split = (a.ntiles + a.nsplit - 1) / a.nsplit;
    uint32_t g_stage = 0;   // producer: stages filled so far by this warpgroup; MMA warp: per-wg counters below
    uint32_t g_tile = 0;    // producer: tiles finished by this warpgroup
    uint32_t mma_stage[2] = {0, 0};
    uint32_t items_done = 0;

    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++items_done) {
        const int split = item % a.nsplit;
        const int sr = item / a.nsplit;
        const int r = sr % a.R, s = sr / a.R;
        const int tile_begin = split * tiles_per_split;
        const int tile_end = min(a.ntiles, tile_begin + tiles_per_split);
        __syncthreads();  // every MMA of the previous item has completed (all epilogues waited on bar_acc)
        if (*abort_flag) break;  // uniform: the flag is only written before the barrier above
        if (warp >= 8) {
            // ============================ MMA issuers (warp 8 -> warpgroup 0, warp 9 -> warpgroup 1) ============================
            const int wg = warp - 8;
            if (wg == 0 && lane == 0) {  // warp 8 also streams the weight image into shared memory (TMA bulk copy)
                mbar_arrive_expect_tx(bar_w, W_BYTES);
                const uint8_t* src = a.wimg + (size_t)sr * W_BYTES;
                constexpr int CH = 16384;
                for (int off = 0; off < W_BYTES; off += CH)
                    bulk_g2s(wsm + off, src + off, min(CH, W_BYTES - off), bar_w);
            }
            __syncwarp();
            bool ok = mbar_wait(bar_w, items_done & 1, abort_flag);
            fence_after_sync();
            const uint32_t wsm_addr = smem_u32(wsm);
            const uint32_t acc = tmem_base + wg * ACC_COLS;
            for (int tile = tile_begin + wg; tile < tile_end && ok; tile += 2) {
                for (int step = 0; step < STEPS; ++step) {
                    const uint32_t st = mma_stage[0] % NST, par = (mma_stage[0] / NST) & 1;
                    ok = mbar_wait(bar_full + wg * NST + st, par, abort_flag);
                    if (!ok) break;
                    fence_after_sync();
                    const uint32_t abase = tmem_base + A_BASE + (wg * NST + st) * A_COLS;
                    const int kg0 = step * (LPS * C / 8);  // first 8-row K group of this stage
                    if (a.debug != 1) {
#pragma unroll
                        for (int kk = 0; kk < MMAS_PER_STAGE; ++kk) {
                            const uint64_t bdesc = smem_desc(wsm_addr + (uint32_t)(kg0 + 2 * kk) * (NG * 128), NG * 128, 128);
                            mma_ts_uniform(acc, abase + kk * 8, bdesc, IDESC, (step | kk) ? 1u : 0u);
                        }
                    }
                    mma_commit_uniform(bar_empty + wg * NST + st);
                    if (step == STEPS - 1) mma_commit_uniform(bar_acc + wg);
                    ++mma_stage[0];
                }
            }
        } else {
            // ===================================== producer / epilogue warpgroups =====================================
            const int wg = warp >> 2;
            const int row = threadIdx.x & 127;
            const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
            bool ok = true;
            const bool prof = a.debug == 9 && blockIdx.x == 0 && row == 0;
            for (int tile = tile_begin + wg; tile < tile_end && ok; tile += 2) {
                TC_PROF_T(t_start);
                const int64_t b = (int64_t)tile * 128 + row;
                const bool valid = b < a.B;
                const bool has_l = 2 * s < a.d_in, has_r = 2 * s + 1 < a.d_in;
                const float* xl = a.x + (((int64_t)(2 * s) * a.R + r) * a.B + (valid ? b : 0)) * C;
                const float* xr = a.x + (((int64_t)(2 * s + 1) * a.R + r) * a.B + (valid ? b : 0)) * C;
                uint32_t eL[C / 2], eR[C / 2];
                float mL = 0.f, mR = 0.f;
                {
                    float vl[C], vr[C];  // both rows are requested before either is consumed (one exposed latency)
#pragma unroll
                    for (int q = 0; q < C / 4; ++q) {
                        const float4 t = (valid && has_l) ? reinterpret_cast<const float4*>(xl)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                        vl[4 * q] = t.x; vl[4 * q + 1] = t.y; vl[4 * q + 2] = t.z; vl[4 * q + 3] = t.w;
                    }
#pragma unroll
                    for (int q = 0; q < C / 4; ++q) {
                        const float4 t = (valid && has_r) ? reinterpret_cast<const float4*>(xr)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                        vr[4 * q] = t.x; vr[4 * q + 1] = t.y; vr[4 * q + 2] = t.z; vr[4 * q + 3] = t.w;
                    }
                    // pull the rows of this warpgroup's next tile towards L2 while this tile is being processed
                    if (tile + 2 < tile_end && b + 256 < a.B) {
                        const char* nl = reinterpret_cast<const char*>(xl + (int64_t)256 * C);
                        const char* nr = reinterpret_cast<const char*>(xr + (int64_t)256 * C);
                        if (has_l) { prefetch_l2(nl); prefetch_l2(nl + C * 4 - 4); }
                        if (has_r) { prefetch_l2(nr); prefetch_l2(nr + C * 4 - 4); }
                    }
                    float m = vl[0], m2 = vr[0];
#pragma unroll
                    for (int q = 1; q < C; ++q) { m = fmaxf(m, vl[q]); m2 = fmaxf(m2, vr[q]); }
                    TC_PROF_T(t_loaded);
                    TC_PROF_ADD(0, t_start, t_loaded);
                    mL = (m == -CUDART_INF_F) ? 0.f : m;
                    mR = (m2 == -CUDART_INF_F) ? 0.f : m2;
#pragma unroll
                    for (int q = 0; q < C / 2; ++q) {
                        eL[q] = pack_bf16(__expf(vl[2 * q] - mL), __expf(vl[2 * q + 1] - mL));
                        eR[q] = pack_bf16(__expf(vr[2 * q] - mR), __expf(vr[2 * q + 1] - mR));
                    }
                }
                TC_PROF_T(t_exp);
                TC_PROF_ADD(1, t_start, t_exp);
#pragma unroll
                for (int step = 0; step < STEPS; ++step) {
                    const uint32_t st = g_stage % NST, par = (g_stage / NST) & 1;
                    TC_PROF_T(t_w0);
                    ok = ok && mbar_wait(bar_empty + wg * NST + st, par ^ 1, abort_flag);
                    if (!ok) break;
                    fence_after_sync();
                    TC_PROF_T(t_w1);
                    TC_PROF_ADD(2, t_w0, t_w1);
                    const uint32_t stage_addr = tmem_base + lane_base + A_BASE + (wg * NST + st) * A_COLS;
#pragma unroll
                    for (int h = 0; h < LPS / 2; ++h) {
                        // two l's per store batch: l = 2*j (low half of eL[j]) and 2*j + 1 (high half), j = step*LPS/2 + h
                        uint32_t p[C];
                        const uint32_t e = eL[step * (LPS / 2) + h];
                        const uint32_t l0 = bcast_lo(e), l1 = bcast_hi(e);
#pragma unroll
                        for (int q = 0; q < C / 2; ++q) {
                            p[q] = hmul2_bf16(l0, eR[q]);
                            p[C / 2 + q] = hmul2_bf16(l1, eR[q]);
                        }
                        if (a.debug != 2) tmem_store_cols<C>(stage_addr + h * C, p);
                    }
                    TC_PROF_T(t_g);
                    TC_PROF_ADD(3, t_w1, t_g);
                    tmem_wait_st();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + wg * NST + st);
                    TC_PROF_T(t_s);
                    TC_PROF_ADD(4, t_g, t_s);
                    ++g_stage;
                }
                if (!ok) break;
                TC_PROF_T(t_a0);
                ok = mbar_wait(bar_acc + wg, g_tile & 1, abort_flag);
                ++g_tile;
                if (!ok) break;
                fence_after_sync();
                TC_PROF_T(t_a1);
                TC_PROF_ADD(5, t_a0, t_a1);
                uint32_t acc[NPAD];
#pragma unroll
                for (int q = 0; q < NPAD / 16; ++q) tmem_ld16(tmem_base + lane_base + wg * ACC_COLS + q * 16, acc + q * 16);
                tmem_wait_ld();
                fence_before_sync();
                if (valid) {
                    float* op = a.out + (((int64_t)s * a.R + r) * a.B + b) * OC;
                    const float shift = mL + mR;
#pragma unroll
                    for (int o4 = 0; o4 < OC; o4 += 4) {
                        float res[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float sum = __uint_as_float(acc[o4 + u]);
                            if (sum >= 1e-30f || a.debug) {
                                res[u] = shift + __logf(sum);
                            } else {
                                const float* lwp = a.lw + ((int64_t)s * K * OC + (o4 + u)) * a.R + r;
                                res[u] = tc_exact_lse<C>(has_l ? xl : nullptr, has_r ? xr : nullptr, lwp, (int64_t)OC * a.R);
                                if (a.logz) res[u] -= a.logz[((int64_t)s * OC + (o4 + u)) * a.R + r];
                            }
                        }
                        *reinterpret_cast<float4*>(op + o4) = make_float4(res[0], res[1], res[2], res[3]);
                    }
                }
                TC_PROF_T(t_end);
                TC_PROF_ADD(6, t_a1, t_end);
                TC_PROF_ADD(7, t_start, t_end);
                if (prof) atomicAdd(&g_tc_prof[8], 1ull);
            }
        }
    }
    // teardown
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (threadIdx.x == 0 && *abort_flag) atomicAdd(&g_tc_aborts, 1);
    if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// Weight operand preparation: fp32 log-weights [1][d][K][OC][R]  ->  bf16 exp() core-matrix image
// [d][R][K/8][NPAD/8][8 (k)][8 (o)], zero padded for o >= OC.  One block per (scope, group of 8 k-values).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tc_wimage_kernel(const float* __restrict__ lw, const float* __restrict__ logz,
                                                       uint8_t* __restrict__ wimg, int d, int K, int OC, int R, int NPAD) {
    extern __shared__ float tile[];  // [8][OC][R]
    const int kg = blockIdx.x % (K / 8);
    const int s = blockIdx.x / (K / 8);
    const int n = 8 * OC * R;
    const float* src = lw + ((int64_t)s * K + (int64_t)kg * 8) * OC * R;
    const int ocr = OC * R;
    const float* zp = logz ? logz + (int64_t)s * ocr : nullptr;
    for (int i = threadIdx.x; i < n; i += blockDim.x) tile[i] = __expf(src[i] - (zp ? zp[i % ocr] : 0.f));
    __syncthreads();
    const int NG = NPAD / 8;
    const int per_r = NG * 64;  // bf16 elements per (r, kg)
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(wimg);
    for (int i = threadIdx.x; i < R * per_r; i += blockDim.x) {
        const int r = i / per_r, e = i % per_r;
        const int og = e / 64, kk = (e / 8) % 8, oo = e % 8;
        const int o = og * 8 + oo;
        const float v = o < OC ? tile[(kk * OC + o) * R + r] : 0.f;
        dst[(((int64_t)s * R + r) * (K / 8) + kg) * per_r + e] = __float2bfloat16(v);
    }
}

// Log-normaliser of the unnormalised weights: logz[s][o][r] = logsumexp_k param[s][k][o][r]  (F.log_softmax(dim=2) of
// layers.py:85 is then lw = param - logz, applied by the consumers).  Block = 64 consecutive (o, r) columns x 4 k-phases;
// every warp load is one 128-byte line; online max/sum so the parameters are read once.
__global__ void __launch_bounds__(256) tc_logz_kernel(const float* __restrict__ param, float* __restrict__ logz, int K,
                                                     int ocr, int tiles_per_s) {
    __shared__ float sm_m[4][64], sm_s[4][64];
    const int s = blockIdx.x / tiles_per_s;
    const int col = (blockIdx.x % tiles_per_s) * 64 + (threadIdx.x & 63);
    const int ph = threadIdx.x >> 6;
    float m = -CUDART_INF_F, acc = 0.f;
    if (col < ocr) {
        const float* p = param + (int64_t)s * K * ocr + col;
        int k = ph;
        for (; k + 12 < K; k += 16) {  // four independent loads in flight per thread
            const float v0 = p[(int64_t)k * ocr], v1 = p[(int64_t)(k + 4) * ocr], v2 = p[(int64_t)(k + 8) * ocr],
                        v3 = p[(int64_t)(k + 12) * ocr];
            const float mn = fmaxf(fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)), m);
            if (mn != -CUDART_INF_F) {
                acc = acc * __expf(m - mn) + __expf(v0 - mn) + __expf(v1 - mn) + __expf(v2 - mn) + __expf(v3 - mn);
                m = mn;
            }
        }
        for (; k < K; k += 4) {
            const float v = p[(int64_t)k * ocr];
            const float mn = fmaxf(m, v);
            if (mn != -CUDART_INF_F) {
                acc = acc * __expf(m - mn) + __expf(v - mn);
                m = mn;
            }
        }
    }
    sm_m[ph][threadIdx.x & 63] = m;
    sm_s[ph][threadIdx.x & 63] = acc;
    __syncthreads();
    if (ph == 0 && col < ocr) {
        float mm = fmaxf(fmaxf(sm_m[0][threadIdx.x], sm_m[1][threadIdx.x]), fmaxf(sm_m[2][threadIdx.x], sm_m[3][threadIdx.x]));
        float tot = 0.f;
        if (mm != -CUDART_INF_F) {
#pragma unroll
            for (int q = 0; q < 4; ++q) tot += sm_s[q][threadIdx.x] * __expf(sm_m[q][threadIdx.x] - mm);
        }
        logz[(int64_t)s * ocr + col] = mm == -CUDART_INF_F ? -CUDART_INF_F : mm + logf(tot);
    }
}

template <int C, int OC, int NPAD, int NST, int LPS>
static int launch_tc_fwd(const TcFwdArgs& a, cudaStream_t st) {
    constexpr int W_BYTES = C * C * NPAD * 2;
    constexpr int SMEM = W_BYTES + (4 * NST + 3) * 8 + 16;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(sum_tc_fwd_kernel<C, OC, NPAD, NST, LPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return spn_cuda_status(e);
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nitems = a.d_out * a.R * a.nsplit;
    const int grid = nitems < sms ? nitems : sms;
    sum_tc_fwd_kernel<C, OC, NPAD, NST, LPS><<<grid, 320, SMEM, st>>>(a);
    return spn_launch_status();
}

extern "C" int spn_sum_tc_supported(int c, int oc, int R) {
    (void)R;
    return (c == 40 && oc == 40) || (c == 32 && oc == 32) || (c == 16 && oc == 16) ? 1 : 0;
}

extern "C" int spn_sum_tc_prepare(const float* lw, int d_out, int c, int oc, int R, float* logz, void* wimg, void* stream) {
    if (!lw || !wimg || !spn_sum_tc_supported(c, oc, R)) return SPN_ERR_UNSUPPORTED;
    const int K = c * c, npad = (oc + 15) / 16 * 16;
    if (logz) {
        const int ocr = oc * R, tiles = (ocr + 63) / 64;
        tc_logz_kernel<<<d_out * tiles, 256, 0, (cudaStream_t)stream>>>(lw, logz, K, ocr, tiles);
        int rc = spn_launch_status();
        if (rc != SPN_OK) return rc;
    }
    const size_t smem = (size_t)8 * oc * R * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(tc_wimage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return spn_cuda_status(e);
    }
    tc_wimage_kernel<<<d_out * (K / 8), 256, smem, (cudaStream_t)stream>>>(lw, logz, (uint8_t*)wimg, d_out, K, oc, R, npad);
    return spn_launch_status();
}

extern "C" int spn_sum_tc_fwd(const float* x, const void* wimg, const float* lw, const float* logz, int B, int d_in,
                              int d_out, int c, int oc, int R, float* out, void* stream) {
    if (!x || !wimg || !lw || !out || B <= 0 || d_in <= 0 || d_in > 2 * d_out) return SPN_ERR_ARG;
    if (!spn_sum_tc_supported(c, oc, R)) return SPN_ERR_UNSUPPORTED;
    TcFwdArgs a{};
    a.x = x; a.wimg = (const uint8_t*)wimg; a.lw = lw; a.logz = logz; a.out = out;
    a.B = B; a.d_in = d_in; a.d_out = d_out; a.R = R; a.OC = oc;
    a.ntiles = (B + 127) / 128;
    int nsplit = (4 * 148 + d_out * R - 1) / (d_out * R);
    if (nsplit > a.ntiles / 4) nsplit = a.ntiles / 4;
    if (nsplit < 1) nsplit = 1;
    a.nsplit = nsplit;
    {
        const char* dbg = getenv("SPN_TC_DEBUG");
        a.debug = dbg ? atoi(dbg) : 0;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (c == 40) return launch_tc_fwd<40, 40, 48, 2, 4>(a, st);
    if (c == 32) return launch_tc_fwd<32, 32, 32, 2, 4>(a, st);
    if (c == 16) return launch_tc_fwd<16, 16, 16, 2, 4>(a, st);
    return SPN_ERR_UNSUPPORTED;
}

extern "C" int spn_debug_tc_prof(unsigned long long* out_host, int reset) {
    unsigned long long v[16] = {0};
    cudaError_t e = cudaMemcpyFromSymbol(v, g_tc_prof, sizeof(v));
    if (e != cudaSuccess) return spn_cuda_status(e);
    if (out_host) for (int i = 0; i < 16; ++i) out_host[i] = v[i];
    if (reset) {
        unsigned long long z[16] = {0};
        e = cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z));
        if (e != cudaSuccess) return spn_cuda_status(e);
    }
    return SPN_OK;
}

// debug: number of CTAs that hit the bounded-wait timeout since the library was loaded (synchronises)
int tc_bwd_aborts_host(int* out);

extern "C" int spn_debug_tc_aborts(int* out_host) {
    int v = 0, w = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&v, g_tc_aborts, sizeof(int));
    if (e != cudaSuccess) return spn_cuda_status(e);
    int rc = tc_bwd_aborts_host(&w);
    if (rc != SPN_OK) return rc;
    if (out_host) *out_host = v + w;
    return SPN_OK;
}
