// dW of the CrossProduct+Sum layer with weights shared by many rows, on the tensor cores (tcgen05 + TMEM); see sum_tc3.cu
// for the forward and dX formulation.  With E[b,k] = eL[b,l] eR[b,r'] (k = l*c + r'), W = exp(log-weights) and
// G[b,o] = g[b,o] * exp(mL + mR - out[b,o])  (= g / linear-domain sum), SURVEY A.5 gives
//
//   dWlin[k,o] = sum_b E[b,k] G[b,o]         (M = k-chunk of 128, N = o, K = rows)   -> spn_sum_tc_bwd_w
//   d(log w)[k,o] = W[k,o] * dWlin[k,o]      (applied by the layout-restoring epilogue kernel, which also fuses the
//                                             softmax backward of the weight normalisation)
#include <stdlib.h>

#include <atomic>
#include "tc_common.cuh"

using namespace tc;

__device__ int g_tc_bwd_aborts = 0;

#ifndef SPN_TC_BWD_DEBUG_MODE
#define SPN_TC_BWD_DEBUG_MODE 0   // compile-time experiment switch (timing only): 1 no MMA issue, 3 generators idle
#endif
#define BWD_DEBUG(a) SPN_TC_BWD_DEBUG_MODE

struct TcBwdArgs {
    unsigned int* counter;   // work counter of this launch (dynamic item scheduling), zeroed by the host
    const uint8_t* fhand; // per (set, s, r, row tile): packed eL, eR ([pair][128 rows] words) + shifts, written by the forward kernel
    const uint8_t* hand;  // per (set, s, r, row tile): G tile (B-operand layout), written by the dX kernel
    float* dwt;           // [nsd][R][P*P][P] linear-domain weight gradient (padded channel counts)
    int B, d_in, d_out, R;   // B = rows per weight set, d_out = weight sets x scopes (the hand-off index does not separate them)
    int ntiles;
};

// Issues the 8 MMAs (K = 16 rows each) of one (tile, chunk): D[128 k x NPAD] (+)= A[128 k x 128 rows] * B[128 rows x NPAD],
// both operands MN-major in shared memory, from ONE PTX block (uniform-register issue path, see sum_tc2.cu).
__device__ __forceinline__ void dw_issue_chunk(uint32_t dcol, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate_first,
                                               uint32_t bstep) {
#define DW_LINE(AOFF, T)                                                                                              \
    "add.s64 da, %1, " #AOFF ";\n\tmad.wide.u32 db, %5, " #T ", %2;\n\t"                                              \
    "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, {%6, %6, %6, %6}, p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
    asm volatile(
        "{\n\t.reg .pred q, p;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        DW_LINE(0, 0) DW_LINE(256, 1) DW_LINE(512, 2) DW_LINE(768, 3) DW_LINE(1024, 4) DW_LINE(1280, 5) DW_LINE(1536, 6) DW_LINE(1792, 7)
        "}" ::"r"(dcol), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first), "r"(bstep), "r"(0u)
        : "memory");
#undef DW_LINE
}

// ================================================================================================================
// dW
// ================================================================================================================
// The k axis (c*c products) is split into 8x8 blocks (8 left channels x 8 right channels); a chunk is two blocks =
// 128 k-values = the M extent of one MMA.  chunk -> k mapping (also used by the epilogue):
//   block = 2*chunk + m/64, l = 8*(block / (C/8)) + (m%64)/8, r' = 8*(block % (C/8)) + m%8, k = l*C + r'.
// One work item = (scope, repetition, group of <= GC chunks); it walks over all row tiles and keeps its accumulators in
// TMEM.  Both operands contract over the batch, so the A operand E[row, k] = eL[row, l] * eR[row, r'] is generated per row
// into shared memory (MN-major core matrices: one 16-byte store per 8 elements); the kernel is bound by shared-memory
// bandwidth (A is written once and read once by the tensor core: ~0.8 MB per 128-row tile), not by the tensor pipe.
//
// The per-row inputs are NOT recomputed here: the forward kernel (sum_tc2.cu, mode FWD) hands over, per (scope, repetition,
// row tile), the packed bf16 vectors eL, eR ([pair][row] words), and the dX kernel (mode DXR) G = g * exp(shift - out) already
// in the core-matrix layout of this kernel's B operand.  Three TMA bulk copies per tile bring them into shared memory, so
// this kernel has no global-load instructions, no exp and no prologue at all (one row per thread with 16-byte loads of
// four fp32 arrays cost ~5000 L1 wavefronts per tile and was the limiter of the previous version).
// Thread organisation (10 warps):
//   warps 0-7   GENERATORS: two sets of 128 threads (one per row) that take alternate (tile, chunk) stages: bf16x2
//               multiplies of the row's eL / eR sub-vectors -> 16-byte stores into the 4-deep A ring; at the end of the
//               item they run the epilogue (TMEM -> linear-domain weight gradient).
//   warp 8      TMA: streams the hand-off tiles (double buffered).
//   warp 9      MMA issuer.
template <int C, int OC, int NPAD, int GC, int NSTA>
__global__ void __launch_bounds__(320, 1) sum_tc_bwd_w_kernel(TcBwdArgs a, int ngroups) {
    constexpr int K = C * C;
    constexpr int CB = C / 8;                  // 8-channel blocks per child = o-groups of the G tile
    constexpr int NBLK = CB * CB;
    constexpr int NCHUNK = (NBLK + 1) / 2;
    constexpr int NPAIR = C / 2;
    constexpr int A_BYTES = 128 * 128 * 2;     // [16 row-groups][16 m-groups][8][8] bf16
    constexpr int V_BYTES = NPAIR * 512;       // one hand-off vector tile: eL / eR as [pair][128 rows] words
    constexpr int KG = NPAD / 8;               // o-groups of the G tile, padded to the MMA's N (zero pad groups)
    constexpr int G_BYTES = 16 * KG * 128;     // G tile: [16 row groups][KG][8 rows][8 bf16]
    static_assert(GC * NPAD <= 512 && OC == C && NSTA >= 2 && NSTA <= 4, "unsupported shape");   // NSTA: A ring depth
    constexpr uint32_t IDESC = instr_desc_bf16(128, NPAD, 1, 1);

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* abuf = smem;                                                     // [NSTA][A_BYTES]
    uint8_t* gbuf = smem + NSTA * A_BYTES;                                    // [2][G_BYTES]
    uint32_t* Ls = reinterpret_cast<uint32_t*>(gbuf + 2 * G_BYTES);           // [2][NPAIR][128] packed eL
    uint32_t* Rs = Ls + 2 * NPAIR * 128;                                      // [2][NPAIR][128] packed eR
    uint64_t* a_full = reinterpret_cast<uint64_t*>(Rs + 2 * NPAIR * 128);     // [NSTA] generators -> MMA
    uint64_t* a_empty = a_full + NSTA;                                        // [NSTA] MMA commit -> generators
    uint64_t* vec_full = a_empty + NSTA;                                      // [2] TMA -> generators, MMA
    uint64_t* vec_empty = vec_full + 2;                                       // [2] generators -> TMA
    uint64_t* g_empty = vec_empty + 2;                                        // [2] MMA commit -> TMA
    uint64_t* bar_done = g_empty + 2;                                         // [1] all MMAs of the item complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTA; ++i) {
            mbar_init(a_full + i, 4);
            mbar_init(a_empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(vec_full + i, 1);
            mbar_init(vec_empty + i, 8);
            mbar_init(g_empty + i, 1);
        }
        mbar_init(bar_done, 1);
        *abort_flag = 0;
        mbar_fence_init();
    }
    if (warp == 9) tmem_alloc(tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (threadIdx.x == 0 && *tmem_slot != 0) *abort_flag = 1;
    constexpr uint32_t tmem_base = 0;
    __syncthreads();

    const int nitems = a.d_out * a.R * ngroups;
    uint32_t n_stage = 0, n_tile = 0, items_done = 0;   // per-role ring positions (persist across items)
    uint32_t m_stage = 0, m_tile = 0;                   // the MMA warp's own counters (kept warp-uniform)

    // dynamic item scheduling (see opg_kernel): late CTAs -- SMs held by a concurrent NCCL kernel -- find less work left
    __shared__ int s_next[2];
    for (;; ++items_done) {
        if (threadIdx.x == 0) s_next[items_done & 1] = (int)atomicAdd(a.counter, 1u);
        __syncthreads();  // the previous item (incl. its epilogue reads of TMEM) is finished by every role
        const int item = __shfl_sync(0xffffffffu, s_next[items_done & 1], 0);
        if (item >= nitems) break;
        const int grp = item % ngroups;
        const int sr = item / ngroups;
        const int r = sr % a.R, s = sr / a.R;
        const int chunk_begin = grp * GC;
        const int chunk_end = min(NCHUNK, chunk_begin + GC);
        const int nch = chunk_end - chunk_begin;
        if (__shfl_sync(0xffffffffu, *abort_flag, 0)) break;
        if (warp == 9) {
            // =============================================== MMA issuer ===============================================
            const uint32_t abuf_addr = smem_u32(abuf), gbuf_addr = smem_u32(gbuf);
            for (int tile = 0; tile < a.ntiles; ++tile, ++m_tile) {
                const uint32_t vb = m_tile & 1;
                mbar_wait_sticky(vec_full + vb, (m_tile >> 1) & 1, abort_flag);   // G operand tile of this row tile
                fence_after_sync();
                // B[K = row, N = o] MN-major: 8-element groups along N 128 B apart, 8-row groups KG*128 B
                const uint64_t bdesc = smem_desc(gbuf_addr + vb * G_BYTES, KG * 128, 128);
#pragma unroll 1
                for (int ch = 0; ch < nch; ++ch, ++m_stage) {
                    const uint32_t st = m_stage % NSTA, par = (m_stage / NSTA) & 1;
                    mbar_wait_sticky(a_full + st, par, abort_flag);
                    fence_after_sync();
                    // A[M = k, K = row] MN-major: 8-element groups along M are 128 B apart, 8-row groups along K 2048 B
                    const uint64_t adesc = smem_desc(abuf_addr + st * A_BYTES, 2048, 128);
                    if (BWD_DEBUG(a) != 1) dw_issue_chunk(tmem_base + ch * NPAD, adesc, bdesc, IDESC, tile ? 1u : 0u, (2 * KG * 128) >> 4);
                    mma_commit_uniform(a_empty + st);
                }
                mma_commit_uniform(g_empty + vb);
            }
            mma_commit_uniform(bar_done);
        } else if (warp == 8) {
            // =============================================== TMA ===============================================
            for (int tile = 0; tile < a.ntiles; ++tile, ++n_tile) {
                const uint32_t vb = n_tile & 1;
                // the buffers of tile n - 2 must have been consumed: row vectors by the generators, G tile by the MMAs
                mbar_wait_sticky(vec_empty + vb, ((n_tile >> 1) & 1) ^ 1, abort_flag);
                mbar_wait_sticky(g_empty + vb, ((n_tile >> 1) & 1) ^ 1, abort_flag);
                if (lane == 0) {
                    const uint8_t* src = a.fhand + ((size_t)sr * a.ntiles + tile) * (2 * V_BYTES + 1024);
                    mbar_arrive_expect_tx(vec_full + vb, 2 * V_BYTES + G_BYTES);
                    bulk_g2s(Ls + vb * NPAIR * 128, src, V_BYTES, vec_full + vb);
                    bulk_g2s(Rs + vb * NPAIR * 128, src + V_BYTES, V_BYTES, vec_full + vb);
                    bulk_g2s(gbuf + vb * G_BYTES, a.hand + ((size_t)sr * a.ntiles + tile) * G_BYTES, G_BYTES, vec_full + vb);
                }
                __syncwarp();
            }
        } else {
            // =============================================== generators ===============================================
            const uint32_t gsel = warp >> 2;          // this set produces the stages with n_stage % 2 == gsel
            const int row = (warp & 3) * 32 + lane;
            for (int tile = 0; tile < a.ntiles; ++tile, ++n_tile) {
                const uint32_t vb = n_tile & 1;
                mbar_wait_sticky(vec_full + vb, (n_tile >> 1) & 1, abort_flag);
                const uint32_t* my_l = Ls + vb * NPAIR * 128 + row;
                const uint32_t* my_r = Rs + vb * NPAIR * 128 + row;
                for (int i = (int)((gsel - n_stage) & 1); i < nch; i += 2) {
                    const int ch = chunk_begin + i;
                    const uint32_t ns = n_stage + i;
                    const uint32_t st = ns % NSTA, use = ns / NSTA;
                    // the eL / eR sub-vectors of the chunk's two 8x8 blocks (fetched before the slot wait)
                    uint32_t el4[2][4], er4[2][4];
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int blk = 2 * ch + half;
                        const bool on = blk < NBLK;
                        const int lb = on ? blk / CB : 0, rb = on ? blk % CB : 0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            el4[half][q] = on ? my_l[(4 * lb + q) * 128] : 0u;
                            er4[half][q] = on ? my_r[(4 * rb + q) * 128] : 0u;
                        }
                    }
                    mbar_wait_sticky(a_empty + st, (use & 1) ^ 1, abort_flag);
                    uint8_t* my_a = abuf + st * A_BYTES + ((row >> 3) * 16) * 128 + (row & 7) * 16;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if (BWD_DEBUG(a) == 3) continue;
#pragma unroll
                        for (int li = 0; li < 8; ++li) {
                            const uint32_t e = el4[half][li >> 1];
                            const uint32_t l2 = (li & 1) ? bcast_hi(e) : bcast_lo(e);
                            *reinterpret_cast<uint4*>(my_a + (half * 8 + li) * 128) =
                                make_uint4(hmul2_bf16(l2, er4[half][0]), hmul2_bf16(l2, er4[half][1]), hmul2_bf16(l2, er4[half][2]),
                                           hmul2_bf16(l2, er4[half][3]));
                        }
                    }
                    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_full + st);
                }
                n_stage += nch;
                fence_proxy_async();   // ordinary loads of eL / eR are ordered before the TMA engine's next write into the buffers
                __syncwarp();
                if (lane == 0) mbar_arrive(vec_empty + vb);  // eL / eR of this tile are no longer read by this warp
            }
            // ---------------------------------------- epilogue: TMEM -> dwt ----------------------------------------
            mbar_wait_sticky(bar_done, items_done & 1, abort_flag);
            fence_after_sync();
            const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
            for (int ch = chunk_begin + (int)gsel; ch < chunk_end; ch += 2) {
                uint32_t v[NPAD];
#pragma unroll
                for (int q = 0; q < NPAD / 8; ++q) tmem_ld8(tmem_base + lane_base + (ch - chunk_begin) * NPAD + q * 8, v + q * 8);
                tmem_wait_ld();
                const int blk = 2 * ch + (row >> 6);
                if (blk < NBLK) {
                    const int l = 8 * (blk / CB) + ((row & 63) >> 3);
                    const int rp = 8 * (blk % CB) + (row & 7);
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
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 9) tmem_dealloc(tmem_base, 512);
    if (threadIdx.x == 0) tc_abort_trap(abort_flag, &g_tc_bwd_aborts);
}

// csum[sg][o][r] = sum over the rows b of weight set `set` with out > -inf of g[s][r][b][o]  (sg = set * d_out + s).  For
// normalised weights lw = param - logz the softmax backward needs sum_k d(lw)[k,o] = sum_b g[b,o] * sum_k pi[b,k,o] =
// sum_b g[b,o]  (pi sums to one over k), so this column sum of the upstream gradient replaces a second pass over the weight
// gradient.  Activations: [d_out][R][Btot = nsets * Bs][OC], rows of one set contiguous.
__global__ void __launch_bounds__(256) tc_gsum_kernel(const float* __restrict__ out, const float* __restrict__ g,
                                                     float* __restrict__ csum, int Bs, int Btot, int d_out, int OC, int R,
                                                     int rows_per_block) {
    __shared__ float4 part[256];
    const int sgr = blockIdx.x;  // (set * d_out + s) * R + r
    const int r = sgr % R, sg = sgr / R;
    const int set = sg / d_out, s = sg % d_out;
    const int quads = OC / 4;                   // float4 per row (OC % 4 == 0 on the tensor-core path)
    const int rpi = 256 / quads;                // rows per iteration
    const int q = threadIdx.x % quads, slot = threadIdx.x / quads;
    const int64_t b0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t b1 = min((int64_t)Bs, b0 + rows_per_block);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot < rpi) {
        const int64_t base = (((int64_t)s * R + r) * Btot + (int64_t)set * Bs) * OC;
        const float4* op = reinterpret_cast<const float4*>(out + base) + q;
        const float4* gp = reinterpret_cast<const float4*>(g + base) + q;
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
        float* dst = csum + ((int64_t)sg * OC + 4 * threadIdx.x) * R + r;
        atomicAdd(dst, t.x);
        atomicAdd(dst + R, t.y);
        atomicAdd(dst + 2 * R, t.z);
        atomicAdd(dst + 3 * R, t.w);
    }
}

// Layout-restoring epilogue of dW: dwt[sg][r][k' = l*P + r'][P] (linear-domain gradient, padded channel counts) -> gradient
// w.r.t. the stored tensor, in the parameter layout [sg][k = l*c + r'][o][r]:
//   logz == NULL: lw holds log-weights;           dlw[k,o]    = exp(lw) * dwt
//   logz != NULL: lw holds unnormalised params p; dparam[k,o] = softmax(p)[k,o] * (dwt[k,o] - csum[o])   (layers.py:85)
// One block per (sg, tile of KT = 4 consecutive k, which share their left channel since c % 4 == 0); tiles go through shared
// memory so both sides stay coalesced.
__global__ void __launch_bounds__(256) tc_dw_finish_kernel(const float* __restrict__ lw, const float* __restrict__ logz,
                                                          const float* __restrict__ csum, const float* __restrict__ dwt,
                                                          float* __restrict__ dlw, int c, int OC, int R, int P) {
    extern __shared__ float tile[];  // [R][KT*OC + 1]
    constexpr int KT = 4;
    const int K = c * c;
    const int kt = blockIdx.x % (K / KT);
    const int s = blockIdx.x / (K / KT);
    const int k0 = kt * KT;
    const int64_t kp0 = (int64_t)(k0 / c) * P + k0 % c;   // padded index of the tile's first k
    const int n = KT * OC;
    for (int i = threadIdx.x; i < R * n; i += blockDim.x) {
        const int r = i / n, e = i % n;
        tile[r * (n + 1) + e] = dwt[(((int64_t)s * R + r) * P * P + kp0 + e / OC) * P + e % OC];
    }
    __syncthreads();
    const int64_t base = ((int64_t)s * K + k0) * OC * R;
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

// Same, vectorised for R % 4 == 0 (OC % 4 == 0 always holds here): float4 on both sides (dwt rows of OC floats; parameter
// rows with r innermost), index arithmetic once per four values.
__global__ void __launch_bounds__(256) tc_dw_finish_vec_kernel(const float* __restrict__ lw, const float* __restrict__ logz,
                                                              const float* __restrict__ csum, const float* __restrict__ dwt,
                                                              float* __restrict__ dlw, int c, int OC, int R, int P) {
    extern __shared__ float tile[];  // [R][KT*OC + 1]
    constexpr int KT = 4;
    const int K = c * c;
    const int kt = blockIdx.x % (K / KT);
    const int s = blockIdx.x / (K / KT);
    const int k0 = kt * KT;
    const int64_t kp0 = (int64_t)(k0 / c) * P + k0 % c;
    const int n = KT * OC, n4 = n / 4, oc4 = OC / 4;
    {
        const int rows_per_pass = blockDim.x / n4;
        const int rr = threadIdx.x / n4, e4 = threadIdx.x - rr * n4;
        if (rr < rows_per_pass) {
            const int kk = e4 / oc4, o4 = e4 - kk * oc4;
            for (int r = rr; r < R; r += rows_per_pass) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(dwt + (((int64_t)s * R + r) * P * P + kp0 + kk) * P) + o4);
                float* t = tile + r * (n + 1) + 4 * e4;
                t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
            }
        }
    }
    __syncthreads();
    const int64_t base = ((int64_t)s * K + k0) * OC * R;
    const int ocr = OC * R, R4 = R / 4;
    const float4* lw4 = reinterpret_cast<const float4*>(lw + base);
    float4* out4 = reinterpret_cast<float4*>(dlw + base);
    for (int i4 = threadIdx.x; i4 < n * R4; i4 += blockDim.x) {
        const int e = i4 / R4, r = 4 * (i4 - e * R4);
        float4 w = __ldg(lw4 + i4);
        const float* t = tile + r * (n + 1) + e;
        float4 dv = make_float4(t[0], t[n + 1], t[2 * (n + 1)], t[3 * (n + 1)]);
        if (logz) {
            const int zc = s * ocr + (e % OC) * R + r;
            const float4 z = __ldg(reinterpret_cast<const float4*>(logz + zc));
            const float4 cs = __ldg(reinterpret_cast<const float4*>(csum + zc));
            w.x -= z.x; w.y -= z.y; w.z -= z.z; w.w -= z.w;
            dv.x -= cs.x; dv.y -= cs.y; dv.z -= cs.z; dv.w -= cs.w;
        }
        out4[i4] = make_float4(__expf(w.x) * dv.x, __expf(w.y) * dv.y, __expf(w.z) * dv.z, __expf(w.w) * dv.w);
    }
}

extern "C" int spn_sum_tc_supported(int c, int oc, int R);
extern "C" int spn_sum_tc_padded_channels(int c, int oc);
unsigned int* tc_next_counter(cudaStream_t st);                                             // sum_tc3.cu
bool tc_configured_on_device(std::atomic<unsigned long long>& mask, int* dev_out);          // sum_tc3.cu

template <int C, int OC, int NPAD, int GC, int NSTA>
static int launch_bwd_w(const TcBwdArgs& a, cudaStream_t st) {
    constexpr int NCHUNK = ((C / 8) * (C / 8) + 1) / 2;
    constexpr int NGROUPS = (NCHUNK + GC - 1) / GC;
    constexpr int SMEM = NSTA * 128 * 128 * 2 + 2 * NPAD * 256 + 2 * 2 * (C / 2) * 512 + (2 * 4 + 7) * 8 + 16;
    static_assert(SMEM <= 232448, "shared memory budget");
    static std::atomic<unsigned long long> configured{0};
    int dev = 0, sms = 148;
    if (!tc_configured_on_device(configured, &dev)) {
        cudaError_t e = cudaFuncSetAttribute(sum_tc_bwd_w_kernel<C, OC, NPAD, GC, NSTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return spn_cuda_status(e);
        if (dev < 64) configured.fetch_or(1ull << dev);
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nitems = a.d_out * a.R * NGROUPS;
    TcBwdArgs b = a;
    b.counter = tc_next_counter(st);
    if (!b.counter) return SPN_ERR_ARG;
    sum_tc_bwd_w_kernel<C, OC, NPAD, GC, NSTA><<<nitems < sms ? nitems : sms, 320, SMEM, st>>>(b, NGROUPS);
    return spn_launch_status();
}

// dwt workspace (floats): padded linear-domain gradient + the column sums of the fused softmax backward
extern "C" int64_t spn_sum_tc_dw_workspace_floats(int nsd, int c, int oc, int R) {
    const int64_t P = spn_sum_tc_padded_channels(c, oc);
    return (int64_t)nsd * R * P * P * P + (int64_t)nsd * oc * R;
}

extern "C" int spn_sum_tc_bwd_w(const void* fhand, const void* hand, const float* out, const float* g, const float* lw, const float* logz,
                                int Bs, int nsets, int d_in, int d_out, int c, int oc, int R, float* dwt_workspace, int csum_ready,
                                float* dlw, void* stream) {
    if (!fhand || !hand || !out || !g || !lw || !dwt_workspace || !dlw || Bs <= 0 || nsets <= 0 || d_in <= 0 || d_in > 2 * d_out)
        return SPN_ERR_ARG;
    if (!spn_sum_tc_supported(c, oc, R)) return SPN_ERR_UNSUPPORTED;
    const int P = spn_sum_tc_padded_channels(c, oc);
    const int nsd = nsets * d_out;
    TcBwdArgs a{};
    a.fhand = (const uint8_t*)fhand; a.hand = (const uint8_t*)hand; a.dwt = dwt_workspace;
    a.B = Bs; a.d_in = d_in; a.d_out = nsd; a.R = R;
    a.ntiles = (Bs + 127) / 128;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    switch (P) {   // <P, P, N = P padded to 16, chunks per item (GC * N <= 512 TMEM columns), A ring depth>
        case 8: rc = launch_bwd_w<8, 8, 16, 1, 4>(a, st); break;
        case 16: rc = launch_bwd_w<16, 16, 16, 2, 4>(a, st); break;
        case 24: rc = launch_bwd_w<24, 24, 32, 5, 4>(a, st); break;
        case 32: rc = launch_bwd_w<32, 32, 32, 8, 4>(a, st); break;
        case 40: rc = launch_bwd_w<40, 40, 48, 7, 4>(a, st); break;
        case 48: rc = launch_bwd_w<48, 48, 48, 9, 4>(a, st); break;
        case 56: rc = launch_bwd_w<56, 56, 64, 7, 3>(a, st); break;
        case 64: rc = launch_bwd_w<64, 64, 64, 8, 3>(a, st); break;
        default: return SPN_ERR_UNSUPPORTED;
    }
    if (rc != SPN_OK) return rc;
    const int K = c * c;
    float* csum = nullptr;
    if (logz) {
        // the column sums live behind the linear-domain gradient in the caller's workspace
        csum = dwt_workspace + (size_t)nsd * R * P * P * P;
        if (!csum_ready) {   // not already produced by the G pre-pass of spn_sum_tc_bwd_x
            cudaError_t e = cudaMemsetAsync(csum, 0, (size_t)nsd * oc * R * sizeof(float), st);
            if (e != cudaSuccess) return spn_cuda_status(e);
            const int rows_per_block = 512;
            dim3 grid(nsd * R, (Bs + rows_per_block - 1) / rows_per_block);
            tc_gsum_kernel<<<grid, 256, 0, st>>>(out, g, csum, Bs, Bs * nsets, d_out, oc, R, rows_per_block);
            rc = spn_launch_status();
            if (rc != SPN_OK) return rc;
        }
    }
    const size_t smem = (size_t)R * (4 * oc + 1) * sizeof(float);
    if (smem > 200 * 1024) return SPN_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(tc_dw_finish_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_dw_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return spn_cuda_status(e);
    }
    if (R % 4 == 0)
        tc_dw_finish_vec_kernel<<<nsd * (K / 4), 256, smem, st>>>(lw, logz, csum, dwt_workspace, dlw, c, oc, R, P);
    else
        tc_dw_finish_kernel<<<nsd * (K / 4), 256, smem, st>>>(lw, logz, csum, dwt_workspace, dlw, c, oc, R, P);
    return spn_launch_status();
}

int tc_bwd_aborts_host(int* out) {
    int v = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&v, g_tc_bwd_aborts, sizeof(int));
    if (e != cudaSuccess) return spn_cuda_status(e);
    *out = v;
    return SPN_OK;
}
