// dW of the shared-weight CrossProduct+Sum layer on the tensor cores (tcgen05 + TMEM); see sum_tc2.cu for the forward
// and dX formulation.  With E[b,k] = eL[b,l] eR[b,r'] (k = l*c + r'), W = exp(log-weights) and
// G[b,o] = g[b,o] * exp(mL + mR - out[b,o])  (= g / linear-domain sum), SURVEY A.5 becomes two GEMMs:
//
//   dX:  dE[b,k] = sum_o G[b,o] W[k,o]            (M = 128 rows, N = k-chunk, K = o)      -> spn_sum_tc_bwd_x
//        dxL[b,l] = eL[b,l] * sum_r' eR[b,r'] dE[b,l,r'],  dxR[b,r'] = eR[b,r'] * sum_l eL[b,l] dE[b,l,r']
//   dW:  dWlin[k,o] = sum_b E[b,k] G[b,o]         (M = k-chunk of 128, N = o, K = rows)   -> spn_sum_tc_bwd_w
//        d(log w)[k,o] = W[k,o] * dWlin[k,o]      (applied by the layout-restoring epilogue kernel)
//
// dX reuses the forward's weight image as a K-major B operand (same bytes, other descriptor); its A operand (G) is
// written to TMEM once per tile and the 128 x N fp32 product is consumed straight out of TMEM by the CUDA cores.
// dW contracts over the batch, so both operands are generated per row into shared memory (MN-major core matrices:
// one 16-byte store per 8 elements, conflict free) and the accumulators for up to 7 k-chunks stay resident in TMEM
// while the CTA walks over every batch tile.
#include <stdlib.h>

#include "tc_common.cuh"

using namespace tc;

__device__ int g_tc_bwd_aborts = 0;

struct TcBwdArgs {
    const float* x;       // [d_in][R][B][C]
    const float* out;     // [d_out][R][B][OC]
    const float* g;       // [d_out][R][B][OC]
    const uint8_t* wimg;  // forward weight image (dX only)
    float* dx;            // [d_in][R][B][C]                (dX)
    float* dwt;           // [d_out][R][K][OC] linear-domain weight gradient (dW)
    int B, d_in, d_out, R;
    int ntiles, nsplit;
};

// Recompute the forward's per-row quantities: packed bf16 eL / eR, and G = g * exp(mL + mR - out) packed bf16.
template <int C, int OC, int NPAD>
__device__ __forceinline__ void tc_bwd_prologue(const TcBwdArgs& a, int s, int r, int64_t b, bool valid, uint32_t* eL,
                                                uint32_t* eR, uint32_t* Gp) {
    const bool has_l = 2 * s < a.d_in, has_r = 2 * s + 1 < a.d_in;
    const float* xl = a.x + (((int64_t)(2 * s) * a.R + r) * a.B + (valid ? b : 0)) * C;
    const float* xr = a.x + (((int64_t)(2 * s + 1) * a.R + r) * a.B + (valid ? b : 0)) * C;
    const float* op = a.out + (((int64_t)s * a.R + r) * a.B + (valid ? b : 0)) * OC;
    const float* gp = a.g + (((int64_t)s * a.R + r) * a.B + (valid ? b : 0)) * OC;
    float vl[C], vr[C];
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
    float m = vl[0], m2 = vr[0];
#pragma unroll
    for (int q = 1; q < C; ++q) { m = fmaxf(m, vl[q]); m2 = fmaxf(m2, vr[q]); }
    const float mL = (m == -CUDART_INF_F) ? 0.f : m;
    const float mR = (m2 == -CUDART_INF_F) ? 0.f : m2;
#pragma unroll
    for (int q = 0; q < C / 2; ++q) {
        eL[q] = pack_bf16(__expf(vl[2 * q] - mL), __expf(vl[2 * q + 1] - mL));
        eR[q] = pack_bf16(__expf(vr[2 * q] - mR), __expf(vr[2 * q + 1] - mR));
    }
    const float shift = mL + mR;
    float G[NPAD];
#pragma unroll
    for (int q = 0; q < OC / 4; ++q) {
        const float4 o4 = valid ? reinterpret_cast<const float4*>(op)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 g4 = valid ? reinterpret_cast<const float4*>(gp)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        // out = -inf (impossible row): no gradient.  exp argument <= ~0 whenever out is finite.
        // G = g / (linear-domain sum) = g * exp(shift - out) >= g.  out = -inf (impossible row): no gradient.  The
        // exponent is capped so that rows whose sum underflowed in the forward (rescue path) stay finite.
        G[4 * q] = o4.x == -CUDART_INF_F ? 0.f : g4.x * __expf(fminf(shift - o4.x, 80.f));
        G[4 * q + 1] = o4.y == -CUDART_INF_F ? 0.f : g4.y * __expf(fminf(shift - o4.y, 80.f));
        G[4 * q + 2] = o4.z == -CUDART_INF_F ? 0.f : g4.z * __expf(fminf(shift - o4.z, 80.f));
        G[4 * q + 3] = o4.w == -CUDART_INF_F ? 0.f : g4.w * __expf(fminf(shift - o4.w, 80.f));
    }
#pragma unroll
    for (int q = OC; q < NPAD; ++q) G[q] = 0.f;
#pragma unroll
    for (int q = 0; q < NPAD / 2; ++q) Gp[q] = pack_bf16(G[2 * q], G[2 * q + 1]);
}

// ================================================================================================================
// dW
// ================================================================================================================
// The k axis (c*c products) is split into 8x8 blocks (8 left channels x 8 right channels); a chunk is two blocks =
// 128 k-values = the M extent of one MMA.  chunk -> k mapping (also used by tc_dw_finish_kernel):
//   block = 2*chunk + m/64, l = 8*(block / (C/8)) + (m%64)/8, r' = 8*(block % (C/8)) + m%8, k = l*C + r'.
// One work item = (scope, repetition, group of <= GC chunks); it walks over all row tiles and keeps its accumulators
// in TMEM.  Warps 0-3 / 4-7 produce the operands of alternating tiles into shared memory, warp 8 issues every MMA.
template <int C, int OC, int NPAD, int GC>
__global__ void __launch_bounds__(288, 1) sum_tc_bwd_w_kernel(TcBwdArgs a, int ngroups) {
    constexpr int NG = NPAD / 8;
    constexpr int K = C * C;
    constexpr int NBLK = (C / 8) * (C / 8);
    constexpr int NCHUNK = (NBLK + 1) / 2;
    constexpr int A_BYTES = 128 * 128 * 2;     // [16 row-groups][16 m-groups][8][8] bf16
    constexpr int G_BYTES = 128 * NPAD * 2;    // [16 row-groups][NG][8][8] bf16
    constexpr int NSTA = 2;                    // A stages per warpgroup
    static_assert(GC * NPAD <= 512, "TMEM budget exceeded");
    constexpr uint32_t IDESC = instr_desc_bf16(128, NPAD, 1, 1);

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* abuf = smem;                                   // [2 wg][NSTA][A_BYTES]
    uint8_t* gbuf = smem + 2 * NSTA * A_BYTES;              // [2 wg][G_BYTES]
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(gbuf + 2 * G_BYTES);  // [2][NSTA]
    uint64_t* bar_empty = bar_full + 2 * NSTA;                              // [2][NSTA]
    uint64_t* bar_gfree = bar_empty + 2 * NSTA;                             // [2] G tile no longer read
    uint64_t* bar_done = bar_gfree + 2;                                     // [1] all MMAs of the item complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * NSTA; ++i) {
            mbar_init(bar_full + i, 4);
            mbar_init(bar_empty + i, 1);
        }
        mbar_init(bar_gfree + 0, 1);
        mbar_init(bar_gfree + 1, 1);
        mbar_init(bar_done, 1);
        *abort_flag = 0;
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (threadIdx.x == 0 && *tmem_slot != 0) *abort_flag = 1;
    constexpr uint32_t tmem_base = 0;
    __syncthreads();

    const int nitems = a.d_out * a.R * ngroups;
    uint32_t n_stage[2] = {0, 0};   // producer: own counter in [0]; MMA warp: one per warpgroup
    uint32_t n_gtile[2] = {0, 0};
    uint32_t items_done = 0;

    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++items_done) {
        const int grp = item % ngroups;
        const int sr = item / ngroups;
        const int r = sr % a.R, s = sr / a.R;
        const int chunk_begin = grp * GC;
        const int chunk_end = min(NCHUNK, chunk_begin + GC);
        __syncthreads();
        if (__shfl_sync(0xffffffffu, *abort_flag, 0)) break;
        if (warp == 8) {
            // ---------------------------------------- MMA issuer ----------------------------------------
            const uint32_t abuf_addr = smem_u32(abuf), gbuf_addr = smem_u32(gbuf);
            for (int t0 = 0; t0 < a.ntiles; t0 += 2) {
#pragma unroll
                for (int wg = 0; wg < 2; ++wg) {
                    const int tile = t0 + wg;
                    if (tile >= a.ntiles) continue;
                    for (int ch = chunk_begin; ch < chunk_end; ++ch) {
                        const uint32_t st = n_stage[wg] % NSTA, par = (n_stage[wg] / NSTA) & 1;
                        mbar_wait_sticky(bar_full + wg * NSTA + st, par, abort_flag);
                        fence_after_sync();
                        const uint32_t aaddr = abuf_addr + (wg * NSTA + st) * A_BYTES;
                        const uint32_t gaddr = gbuf_addr + wg * G_BYTES;
                        const uint32_t dcol = tmem_base + (ch - chunk_begin) * NPAD;
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            // A[M = k, K = row] MN-major: 8-element groups along M are 128 B apart, 8-row groups along K 2048 B
                            const uint64_t adesc = smem_desc(aaddr + (uint32_t)(2 * t) * 2048, 2048, 128);
                            // B[K = row, N = o] MN-major: 8-element groups along N 128 B apart, 8-row groups NG*128 B
                            const uint64_t bdesc = smem_desc(gaddr + (uint32_t)(2 * t) * (NG * 128), NG * 128, 128);
                            mma_ss_uniform(dcol, adesc, bdesc, IDESC, (tile | t) ? 1u : 0u);
                        }
                        mma_commit_uniform(bar_empty + wg * NSTA + st);
                        ++n_stage[wg];
                    }
                    mma_commit_uniform(bar_gfree + wg);
                }
            }
            mma_commit_uniform(bar_done);
        } else {
            // ---------------------------------------- operand producers ----------------------------------------
            const int wg = warp >> 2;
            const int row = threadIdx.x & 127;
            uint8_t* my_g = gbuf + wg * G_BYTES + ((row >> 3) * NG) * 128 + (row & 7) * 16;
            for (int tile = wg; tile < a.ntiles; tile += 2) {
                const int64_t b = (int64_t)tile * 128 + row;
                const bool valid = b < a.B;
                uint32_t eL[C / 2], eR[C / 2], Gp[NPAD / 2];
                tc_bwd_prologue<C, OC, NPAD>(a, s, r, b, valid, eL, eR, Gp);
                if (!valid) {
#pragma unroll
                    for (int q = 0; q < NPAD / 2; ++q) Gp[q] = 0u;  // rows past the batch contribute nothing
                }
                // G tile of the previous tile of this warpgroup must have been consumed
                mbar_wait_sticky(bar_gfree + wg, (n_gtile[0] & 1) ^ 1, abort_flag);
                ++n_gtile[0];
#pragma unroll
                for (int og = 0; og < NG; ++og)
                    *reinterpret_cast<uint4*>(my_g + og * 128) = make_uint4(Gp[4 * og], Gp[4 * og + 1], Gp[4 * og + 2], Gp[4 * og + 3]);
#pragma unroll 1
                for (int ch = chunk_begin; ch < chunk_end; ++ch) {
                    const uint32_t st = n_stage[0] % NSTA, par = (n_stage[0] / NSTA) & 1;
                    mbar_wait_sticky(bar_empty + wg * NSTA + st, par ^ 1, abort_flag);
                    uint8_t* my_a = abuf + (wg * NSTA + st) * A_BYTES + ((row >> 3) * 16) * 128 + (row & 7) * 16;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int blk = 2 * ch + half;
                        const int lb = blk / (C / 8), rb = blk % (C / 8);
                        // eR for r' in [8 rb, 8 rb + 8): 4 packed registers, selected without dynamic register indexing
                        uint32_t er[4], el4[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) { er[q] = 0u; el4[q] = 0u; }
#pragma unroll
                        for (int j = 0; j < C / 8; ++j) {
                            if (j == rb && blk < NBLK) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) er[q] = eR[4 * j + q];
                            }
                            if (j == lb && blk < NBLK) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) el4[q] = eL[4 * j + q];
                            }
                        }
#pragma unroll
                        for (int li = 0; li < 8; ++li) {
                            const uint32_t e = el4[li >> 1];
                            const uint32_t l2 = (li & 1) ? bcast_hi(e) : bcast_lo(e);
                            *reinterpret_cast<uint4*>(my_a + (half * 8 + li) * 128) =
                                make_uint4(hmul2_bf16(l2, er[0]), hmul2_bf16(l2, er[1]), hmul2_bf16(l2, er[2]), hmul2_bf16(l2, er[3]));
                        }
                    }
                    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + wg * NSTA + st);
                    ++n_stage[0];
                }
            }
            // ---------------------------------------- epilogue: TMEM -> dwt ----------------------------------------
            mbar_wait_sticky(bar_done, items_done & 1, abort_flag);
            {
                fence_after_sync();
                const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
                for (int ch = chunk_begin + wg; ch < chunk_end; ch += 2) {
                    uint32_t v[NPAD];
#pragma unroll
                    for (int q = 0; q < NPAD / 8; ++q) tmem_ld8(tmem_base + lane_base + (ch - chunk_begin) * NPAD + q * 8, v + q * 8);
                    tmem_wait_ld();
                    const int blk = 2 * ch + (row >> 6);
                    if (blk < NBLK) {
                        const int l = 8 * (blk / (C / 8)) + ((row & 63) >> 3);
                        const int rp = 8 * (blk % (C / 8)) + (row & 7);
                        float* dst = a.dwt + (((int64_t)s * a.R + r) * K + (l * C + rp)) * OC;
#pragma unroll
                        for (int q = 0; q < OC / 4; ++q)
                            reinterpret_cast<float4*>(dst)[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                            __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                    }
                }
                fence_before_sync();
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (threadIdx.x == 0 && *abort_flag) atomicAdd(&g_tc_bwd_aborts, 1);
    if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// csum[s][o][r] = sum over rows b with out > -inf of g[s][r][b][o].  For normalised weights lw = param - logz the
// softmax backward needs sum_k d(lw)[k,o] = sum_b g[b,o] * sum_k pi[b,k,o] = sum_b g[b,o]  (pi sums to one over k), so this
// column sum of the upstream gradient replaces a second pass over the weight gradient.
__global__ void __launch_bounds__(256) tc_gsum_kernel(const float* __restrict__ out, const float* __restrict__ g,
                                                     float* __restrict__ csum, int B, int OC, int R, int rows_per_block) {
    __shared__ float4 part[256];
    const int sr = blockIdx.x;  // s * R + r
    const int quads = OC / 4;                   // float4 per row (OC % 4 == 0 on the tensor-core path)
    const int rpi = 256 / quads;                // rows per iteration
    const int q = threadIdx.x % quads, slot = threadIdx.x / quads;
    const int64_t b0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t b1 = min((int64_t)B, b0 + rows_per_block);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot < rpi) {
        const float4* op = reinterpret_cast<const float4*>(out + (int64_t)sr * B * OC) + q;
        const float4* gp = reinterpret_cast<const float4*>(g + (int64_t)sr * B * OC) + q;
#pragma unroll 4
        for (int64_t b = b0 + slot; b < b1; b += rpi) {
            const float4 ov = __ldg(op + b * quads), gv = __ldg(gp + b * quads);
            acc.x += ov.x == -CUDART_INF_F ? 0.f : gv.x;
            acc.y += ov.y == -CUDART_INF_F ? 0.f : gv.y;
            acc.z += ov.z == -CUDART_INF_F ? 0.f : gv.z;
            acc.w += ov.w == -CUDART_INF_F ? 0.f : gv.w;
        }
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < quads) {
        float4 t = part[threadIdx.x];
        for (int j = 1; j < rpi; ++j) {
            const float4 u = part[j * quads + threadIdx.x];
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        const int s = sr / R, r = sr % R;
        float* dst = csum + ((int64_t)s * OC + 4 * threadIdx.x) * R + r;
        atomicAdd(dst, t.x);
        atomicAdd(dst + R, t.y);
        atomicAdd(dst + 2 * R, t.z);
        atomicAdd(dst + 3 * R, t.w);
    }
}

// Layout-restoring epilogue of dW: dwt[s][r][k][o] (linear-domain gradient) -> gradient w.r.t. the stored tensor, in the
// parameter layout [1][s][k][o][r]:
//   logz == NULL: lw holds log-weights;           dlw[k,o]    = exp(lw) * dwt
//   logz != NULL: lw holds unnormalised params p; dparam[k,o] = softmax(p)[k,o] * (dwt[k,o] - csum[o])   (layers.py:85)
// One block per (s, tile of KT k); tiles go through shared memory so both sides stay coalesced.
__global__ void __launch_bounds__(256) tc_dw_finish_kernel(const float* __restrict__ lw, const float* __restrict__ logz,
                                                          const float* __restrict__ csum, const float* __restrict__ dwt,
                                                          float* __restrict__ dlw, int d, int K, int OC, int R) {
    extern __shared__ float tile[];  // [R][KT*OC + 1]
    constexpr int KT = 4;
    const int kt = blockIdx.x % (K / KT);
    const int s = blockIdx.x / (K / KT);
    const int n = KT * OC;
    for (int i = threadIdx.x; i < R * n; i += blockDim.x) {
        const int r = i / n, e = i % n;
        tile[r * (n + 1) + e] = dwt[(((int64_t)s * R + r) * K + (int64_t)kt * KT) * OC + e];
    }
    __syncthreads();
    const int64_t base = ((int64_t)s * K + (int64_t)kt * KT) * OC * R;
    const int ocr = OC * R;
    for (int i = threadIdx.x; i < n * R; i += blockDim.x) {
        const int e = i / R, r = i % R;
        float w = lw[base + i], dv = tile[r * (n + 1) + e];
        if (logz) {
            const int zc = s * ocr + i % ocr;
            w -= logz[zc];
            dv -= csum[zc];
        }
        dlw[base + i] = __expf(w) * dv;
    }
}

extern "C" int spn_sum_tc_supported(int c, int oc, int R);

template <int C, int OC, int NPAD, int GC>
static int launch_bwd_w(const TcBwdArgs& a, cudaStream_t st) {
    constexpr int NCHUNK = ((C / 8) * (C / 8) + 1) / 2;
    constexpr int NGROUPS = (NCHUNK + GC - 1) / GC;
    constexpr int SMEM = 2 * 2 * 128 * 128 * 2 + 2 * 128 * NPAD * 2 + (4 * 2 + 3) * 8 + 16;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(sum_tc_bwd_w_kernel<C, OC, NPAD, GC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return spn_cuda_status(e);
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nitems = a.d_out * a.R * NGROUPS;
    sum_tc_bwd_w_kernel<C, OC, NPAD, GC><<<nitems < sms ? nitems : sms, 288, SMEM, st>>>(a, NGROUPS);
    return spn_launch_status();
}

extern "C" int spn_sum_tc_bwd_w(const float* x, const float* out, const float* g, const float* lw, const float* logz,
                                int B, int d_in, int d_out, int c, int oc, int R, float* dwt_workspace, float* dlw,
                                void* stream) {
    if (!x || !out || !g || !lw || !dwt_workspace || !dlw || B <= 0 || d_in <= 0 || d_in > 2 * d_out) return SPN_ERR_ARG;
    if (!spn_sum_tc_supported(c, oc, R)) return SPN_ERR_UNSUPPORTED;
    TcBwdArgs a{};
    a.x = x; a.out = out; a.g = g; a.dwt = dwt_workspace;
    a.B = B; a.d_in = d_in; a.d_out = d_out; a.R = R;
    a.ntiles = (B + 127) / 128;
    a.nsplit = 1;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (c == 40) rc = launch_bwd_w<40, 40, 48, 7>(a, st);
    else if (c == 32) rc = launch_bwd_w<32, 32, 32, 8>(a, st);
    else if (c == 16) rc = launch_bwd_w<16, 16, 16, 2>(a, st);
    else return SPN_ERR_UNSUPPORTED;
    if (rc != SPN_OK) return rc;
    const int K = c * c;
    float* csum = nullptr;
    if (logz) {
        // the column sums live behind the linear-domain gradient in the caller's workspace
        csum = dwt_workspace + (size_t)d_out * R * K * oc;
        cudaError_t e = cudaMemsetAsync(csum, 0, (size_t)d_out * oc * R * sizeof(float), st);
        if (e != cudaSuccess) return spn_cuda_status(e);
        const int rows_per_block = 512;
        dim3 grid(d_out * R, (B + rows_per_block - 1) / rows_per_block);
        tc_gsum_kernel<<<grid, 256, 0, st>>>(out, g, csum, B, oc, R, rows_per_block);
        rc = spn_launch_status();
        if (rc != SPN_OK) return rc;
    }
    const size_t smem = (size_t)R * (4 * oc + 1) * sizeof(float);
    tc_dw_finish_kernel<<<d_out * (K / 4), 256, smem, st>>>(lw, logz, csum, dwt_workspace, dlw, d_out, K, oc, R);
    return spn_launch_status();
}

int tc_bwd_aborts_host(int* out) {
    int v = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&v, g_tc_bwd_aborts, sizeof(int));
    if (e != cudaSuccess) return spn_cuda_status(e);
    *out = v;
    return SPN_OK;
}
