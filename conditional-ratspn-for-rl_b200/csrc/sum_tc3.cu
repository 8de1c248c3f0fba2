// CrossProduct + Sum on the 5th-generation tensor cores (tcgen05 + TMEM): forward and dX, for weights shared by many
// rows (RatSpn: one weight set; CSPN: one set per conditional with many rows per set -- the entropy estimate).
//
// For one weight set, sum scope s and repetition r the layer contracts K = c*c in-channel pairs:
//   forward  out[b,o]  = mL + mR + log sum_l eL[b,l] * T[b,(l,o)],   T[b,(l,o)] = sum_r' eR[b,r'] W[(l,r'),o]
//   dX       dxL[b,l]  = eL[b,l]  * sum_r' eR[b,r'] * dE[b,(l,r')],  dE[b,(l,r')] = sum_o G[b,o] W[(l,r'),o]
//            dxR[b,r'] = eR[b,r'] * sum_l  eL[b,l]  * dE[b,(l,r')]
// with eL = exp(xL - mL), eR = exp(xR - mR), G = g * exp(mL + mR - out).  Both are ONE small-K GEMM per 128-row tile
// (M = 128 rows, K = c resp. oc, N = c*oc resp. c*c) whose result is consumed straight out of tensor memory by a
// contraction epilogue on the CUDA cores:
//   * A operand: forward: the packed bf16 row vectors eR, written to TENSOR MEMORY by the prologue warps (tcgen05.st, lane =
//     row); dX: the G tile the G pre-pass left in core-matrix layout, copied by the TMA engine into shared memory (SS MMA:
//     the dX kernel has no prologue at all).
//   * B operand: the bf16 weight image of (set, s, r) in shared memory, copied by the TMA engine (cp.async.bulk).  Forward
//     reads image F (MN-major: K = r', N = (l, o)), dX image A (K-major: K = o, N = (l, r')).  Images up to 128 KB stay
//     resident for all row tiles of a work item; wider ones stream chunk by chunk through a ring.
//   * accumulators: fp32 in TMEM, NL*P columns per chunk of NL left channels, NB chunk buffers in flight.  Inside a chunk the
//     N axis is ordered so that the columns ONE epilogue thread consumes are consecutive (one or two wide tcgen05.ld).
//   * epilogue: tcgen05.ld (measured 1.6 KB/clk/SM on B200, profiles/r02_pipe_bench.txt) + packed fp32 FMAs (FFMA2: 128
//     FMA/clk/SM): forward 1 FMA per accumulator element (c*oc per row), dX 2 (one per child) on ONE load of the element,
//     against M*N*K/8192 clk of tensor pipe -- both are bound by the fp32 pipe (1600 / 3200 clk per 128-row tile at
//     c = oc = 40), not by the tensor pipe (1200 clk).
// This replaces the round-1 formulation (outer-product operand eL (x) eR generated into TMEM, K = c*c, N = oc; dX as two
// such GEMMs), which was chosen on the assumption that reading a 128 x c*c accumulator back costs 64 B/clk.
//
// Channel counts are padded to P = multiple of 8 (template); c and oc (<= P, multiples of 4) are run-time values: padded
// channels carry e = 0 / zero image rows and are never stored.  Activations use the internal layout
// [scope][repetition][row][channel]; rows of one weight set are contiguous (set-major).
// Outputs whose linear-domain sum underflows are recomputed exactly in fp32, so results keep th.logsumexp semantics.
#include <stdlib.h>

#include <atomic>
#include "tc_common.cuh"

using namespace tc;

__device__ int g_tc3_aborts = 0;
// Per-phase clock hooks of the dX epilogue (build with SPN_TC_PROFILE=1; warp 0 of CTA 0 accumulates, see spn_debug_tc_prof)
__device__ unsigned long long g_tc3_prof[12];
#ifdef SPN_TC_PROFILE
#define TSP_DECL long long tsp_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long tsp_t = clock64()
#define TSP_LAP(slot) do { const long long tsp_n = clock64(); tsp_acc[slot] += tsp_n - tsp_t; tsp_t = tsp_n; } while (0)
#define TSP_FLUSH() do { if (blockIdx.x == 0 && threadIdx.x == 0) for (int q = 0; q < 12; ++q) atomicAdd(&g_tc3_prof[q], (unsigned long long)tsp_acc[q]); } while (0)
#else
#define TSP_DECL do { } while (0)
#define TSP_LAP(slot) do { } while (0)
#define TSP_FLUSH() do { } while (0)
#endif

enum { TS_FWD = 0, TS_DX = 1 };

// Experiment switches (timing only, results are garbage): compile-time, -DSPN_TC_DEBUG_MODE=<n> (build.py):
//   1 no MMA issue (commits only), 2 epilogue skips loads and FMAs, 3 epilogue loads but no FMAs, 4 FMAs on stale registers
//   (no loads), 5 forward prologue skips the exp computation, 6 = 2 + 5
#ifndef SPN_TC_DEBUG_MODE
#define SPN_TC_DEBUG_MODE 0
#endif
#define TS_DBG SPN_TC_DEBUG_MODE
#define TS_SKIP_EPI (SPN_TC_DEBUG_MODE == 2 || SPN_TC_DEBUG_MODE == 6)
#define TS_SKIP_PRO (SPN_TC_DEBUG_MODE == 5 || SPN_TC_DEBUG_MODE == 6)

struct TsArgs {
    unsigned int* counter;   // work counter of this launch (dynamic item scheduling), zeroed by the host
    const float* x;          // [d_in][R][Btot][c]
    const uint8_t* wimg;     // FWD: image F, DX: image A; one image of P*P*P*2 bytes per (set, scope, repetition)
    const float* lw;         // fp32 weights [nsets][d_out][c*c][oc][R] (exact rescue path, FWD)
    const float* logz;       // NULL or [nsets*d_out][oc][R]
    float* y;                // FWD: out [d_out][R][Btot][oc];  DX: dx [d_in][R][Btot][c]
    uint8_t* fhand_w;        // FWD, optional: hand-off written per (set, s, r, row tile): packed bf16 eL, eR + row shifts
    const uint8_t* fhand_r;  // DX: the forward's hand-off
    const uint8_t* hand;     // DX: G tiles (written by the G pre-pass) in core-matrix layout [row / 8][o / 8 (padded)][row % 8][8 o]
    int Bs, Btot, nsets, d_in, d_out, R, c, oc;
    int ntiles, nsplit;
};

// ---- compile-time geometry ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int ts_kp(int P) { return (P + 15) / 16 * 16; }
// left channels per chunk: even, a divisor of P, NL*P a multiple of 16, three chunk buffers next to the forward's A tile in
// the 512 TMEM columns
__host__ __device__ constexpr int ts_nl(int P) {
    int best = 2;
    for (int nl = 2; nl <= P; nl += 2)
        if (3 * nl * P + ts_kp(P) / 2 <= 512 && nl * P <= 256 && (nl * P) % 16 == 0 && P % nl == 0) best = nl;
    return best;
}
__host__ __device__ constexpr int ts_nb(int P) {
    int nb = (512 - ts_kp(P) / 2) / (ts_nl(P) * P);
    return nb > 4 ? 4 : nb;
}
__host__ __device__ constexpr bool ts_resident(int P) { return P * P * P * 2 <= 131072; }
__host__ __device__ constexpr int ts_nslot(int P) { return ts_resident(P) ? 1 : 4; }
__host__ __device__ constexpr int ts_wbytes(int P) { return ts_resident(P) ? P * P * P * 2 : ts_nslot(P) * ts_nl(P) * P * P * 2; }
constexpr int TS_ZERO = 4096;   // zeroed bytes behind the weights: target of padded K-halves / descriptor overruns
constexpr int TS_NXS = 3;       // forward: ring of x stages (one child's 128 rows each)
// Position of (left channel l, column v) on the N axis of a chunk.  The columns one epilogue thread consumes are consecutive:
//   forward (v = output o):        thread = quarter of the outputs (QC = P/4), all NL left channels of the chunk
//   dX      (v = right channel r'): thread = half of the right channels (HC = P/2) x half of the chunk's left channels
// Both orders are "block index, then left channel inside the chunk, then position inside the block".
__host__ __device__ constexpr int ts_ncol(int P, int blk_cols, int li, int v) {
    return (v / blk_cols) * (ts_nl(P) * blk_cols) + li * blk_cols + v % blk_cols;
}
template <int P, int MODE>
__host__ __device__ constexpr int ts_smem_bytes() {
    const int npair = P / 2;
    const int vec = MODE == TS_FWD ? (TS_NXS * 128 * P * 4 /* x ring */ + 2 * npair * 512 /* packed eL, two tiles */ + npair * 512 /* output staging */)
                                   : (2 * 16 * (ts_kp(P) / 8) * 128 /* G tiles (A operand), K padded */ + 2 * npair * 512 /* packed eL */ +
                                      2 * P * 128 * 4 /* aL shares of the two halves */);
    return ts_wbytes(P) + TS_ZERO + vec + 2 * 2 * 128 * 4 /* shifts */ + 32 * 8 /* barriers */ + 16;
}

// exact fp32 log-sum-exp for NOUT outputs of one row (rare rescue path)
__device__ __noinline__ void ts_exact_row(const TsArgs& a, int sg, int s, int r, int64_t brow, int o_begin, int nout, float* op) {
    const int c = a.c, oc = a.oc;
    const bool has_l = 2 * s < a.d_in, has_r = 2 * s + 1 < a.d_in;
    const float* xl = a.x + (((int64_t)(2 * s) * a.R + r) * a.Btot + brow) * c;
    const float* xr = a.x + (((int64_t)(2 * s + 1) * a.R + r) * a.Btot + brow) * c;
    for (int oo = 0; oo < nout; ++oo) {
        const int o = o_begin + oo;
        if (o >= oc) break;
        const float* lwp = a.lw + ((int64_t)sg * c * c * oc + o) * a.R + r;
        const int64_t sw_i = (int64_t)oc * a.R;
        float m = -CUDART_INF_F;
        for (int l = 0; l < c; ++l)
            for (int rp = 0; rp < c; ++rp)
                m = fmaxf(m, (has_l ? xl[l] : 0.f) + (has_r ? xr[rp] : 0.f) + lwp[(int64_t)(l * c + rp) * sw_i]);
        float res = m;
        if (m != -CUDART_INF_F) {
            float acc = 0.f;
            for (int l = 0; l < c; ++l)
                for (int rp = 0; rp < c; ++rp)
                    acc += expf((has_l ? xl[l] : 0.f) + (has_r ? xr[rp] : 0.f) + lwp[(int64_t)(l * c + rp) * sw_i] - m);
            res = m + logf(acc);
            if (a.logz) res -= a.logz[((int64_t)sg * oc + o) * a.R + r];
        }
        op[oo] = res;
    }
}

__device__ __forceinline__ void ts_named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// load N (even, <= 64) consecutive columns of this thread's TMEM lane with the widest instructions available
template <int N>
__device__ __forceinline__ void ts_tmem_load(uint32_t taddr, uint32_t* r) {
    static_assert(N % 2 == 0 && N <= 128, "unsupported column count");
    int done = 0;
    if constexpr (N >= 32) { tmem_ld32(taddr, r); done = 32; }
    if constexpr (N >= 64) { tmem_ld32(taddr + 32, r + 32); done = 64; }
    if constexpr (N >= 96) { tmem_ld32(taddr + 64, r + 64); done = 96; }
    if constexpr (N >= 128) { tmem_ld32(taddr + 96, r + 96); done = 128; }
    if constexpr ((N % 32) >= 16) { tmem_ld16(taddr + done, r + done); done += 16; }
    if constexpr ((N % 16) >= 8) { tmem_ld8(taddr + done, r + done); done += 8; }
    if constexpr ((N % 8) >= 4) { tmem_ld4(taddr + done, r + done); done += 4; }
    if constexpr ((N % 4) >= 2) { tmem_ld2(taddr + done, r + done); done += 2; }
}

// work item of the persistent kernels: (weight set, scope, repetition, row-tile range)
struct TsItem {
    int sr, sg, s, r, tile_begin, nt;
    bool has_l, has_r;
    int64_t row0;
};
__device__ __forceinline__ TsItem ts_item(const TsArgs& a, int item) {
    TsItem it;
    const int tiles_per_split = (a.ntiles + a.nsplit - 1) / a.nsplit;
    const int split = item % a.nsplit;
    it.sr = item / a.nsplit;                    // (set, scope, repetition) index: also the image / hand-off index
    it.r = it.sr % a.R;
    it.sg = it.sr / a.R;                        // set * d_out + s
    it.s = it.sg % a.d_out;
    it.tile_begin = split * tiles_per_split;
    it.nt = min(a.ntiles, it.tile_begin + tiles_per_split) - it.tile_begin;
    it.has_l = 2 * it.s < a.d_in;
    it.has_r = 2 * it.s + 1 < a.d_in;
    it.row0 = (int64_t)(it.sg / a.d_out) * a.Bs;   // first row of this weight set in the [Btot] row axis
    return it;
}

// ====================================================================================================================
// forward
// ====================================================================================================================
// Thread organisation (27 warps, <= 72 registers per thread):
//   warps 0-7   PROLOGUE: one thread per (row, child) -- warps 0-3 the left child, 4-7 the right child.  Row max, exp, pack to
//               bf16; the right child's packed vector goes to the TMEM A tile, the left child's to shared memory (epilogue);
//               both are handed to the backward kernels through `fhand`.  The x rows arrive through a 3-deep ring of TMA
//               stages (one child's 128 rows each), so the HBM latency of a stage is covered by two stages of compute.
//   warps 8-23  EPILOGUE: four threads per row (same lane, warps of one TMEM lane quadrant), each a quarter of the outputs:
//               acc[o] += eL[l] * T[l, o] over the NL left channels of a chunk -- one tcgen05.ld of NL*P/4 consecutive
//               columns and NL*P/8 FFMA2 per chunk.  tcgen05.ld has > 100 clk of latency and is not double buffered here:
//               four warps per scheduler hide it.
//   warp 24     MMA issuer.    warp 25   weight TMA (whole image per work item, or chunks through a ring).
//   warp 26     x TMA.
// TMEM: NB chunk buffers (accumulators) + ONE A tile; the prologue writes the next tile's A operand when the MMAs of the
// current one have completed, while the epilogue still has up to NB chunks to consume.
// All hand-offs are mbarriers (producer arrive = release, consumer try_wait = acquire).
template <int P>
__global__ void __launch_bounds__(864, 1) tsum_fwd_kernel(TsArgs a) {
    constexpr int CG = P / 8;                  // 8-channel groups
    constexpr int KP = ts_kp(P);               // K extent of the GEMM (padded to the MMA's K = 16)
    constexpr int KSTEPS = KP / 16;
    constexpr int ACOLS = KP / 2;              // A tile: packed bf16 pairs, one 32-bit column per two K values
    constexpr int NL = ts_nl(P);               // left channels per chunk
    constexpr int NCH = P / NL;                // chunks per row tile
    constexpr int NCOLS = NL * P;              // accumulator columns per chunk buffer
    constexpr int NB = ts_nb(P);               // chunk buffers
    constexpr int QC = P / 4;                  // outputs per epilogue thread
    constexpr int TC = NL * QC;                // accumulator columns per epilogue thread and chunk
    constexpr int NPAIR = P / 2;               // packed words per vector
    constexpr bool RESIDENT = ts_resident(P);
    constexpr int NSLOT = ts_nslot(P);
    constexpr int CHUNK_BYTES = NL * P * P * 2;
    constexpr int IMG_BYTES = P * P * P * 2;
    constexpr int W_BYTES = ts_wbytes(P);
    constexpr int A_COL0 = NB * NCOLS;
    constexpr int FH_BYTES = 2 * NPAIR * 512 + 1024;   // forward hand-off per (item, tile): eL, eR packed + shifts
    constexpr int XS_FLOATS = 128 * P;                 // one x stage
    static_assert(P % 8 == 0 && P >= 8 && P <= 64 && NL % 2 == 0 && P % NL == 0, "unsupported channel count");
    static_assert(NB >= 2 && A_COL0 + ACOLS <= 512 && ACOLS % 4 == 0 && QC % 2 == 0 && TC <= 64, "TMEM / register budget");
    constexpr int W_MMA = 24, W_WTMA = 25, W_XTMA = 26;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* wsm = smem;                                       // weights: resident image or ring of NSLOT chunks
    uint8_t* zero = smem + W_BYTES;                            // TS_ZERO bytes of zeros
    float* Xr = reinterpret_cast<float*>(zero + TS_ZERO);      // [TS_NXS][128][c] fp32 rows staged by TMA
    uint32_t* ELp = reinterpret_cast<uint32_t*>(Xr + TS_NXS * XS_FLOATS);   // [2 buf][NPAIR][128] packed eL
    float* Ms = reinterpret_cast<float*>(ELp + 2 * NPAIR * 128);            // [2 buf][2][128] row shifts mL, mR
    float* OST = Ms + 2 * 2 * 128;                                          // [64][oc] staging of the upper half of an output tile
    uint64_t* vec_full = reinterpret_cast<uint64_t*>(OST + NPAIR * 128);  // [2]  prologue -> epilogue
    uint64_t* vec_empty = vec_full + 2;                                  // [2]  epilogue -> prologue
    uint64_t* acc_full = vec_empty + 2;                                  // [NB] MMA commit -> epilogue
    uint64_t* acc_empty = acc_full + 4;                                  // [NB] epilogue -> MMA
    uint64_t* w_full = acc_empty + 4;                                    // [NSLOT] TMA -> MMA
    uint64_t* w_empty = w_full + 4;                                      // [NSLOT] MMA commit -> TMA
    uint64_t* x_full = w_empty + 4;                                      // [TS_NXS] TMA -> prologue
    uint64_t* x_empty = x_full + TS_NXS;                                 // [TS_NXS] prologue -> TMA
    uint64_t* a_full = x_empty + TS_NXS;                                 // [1]  prologue (A tile in TMEM) -> MMA
    uint64_t* a_free = a_full + 1;                                       // [1]  MMA commit (last chunk of a tile) -> prologue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_free + 1);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(vec_full + i, 8);
            mbar_init(vec_empty + i, 16);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(acc_full + i, 1);
            mbar_init(acc_empty + i, 16);
            mbar_init(w_full + i, 1);
            mbar_init(w_empty + i, 1);
        }
        for (int i = 0; i < TS_NXS; ++i) {
            mbar_init(x_full + i, 1);
            mbar_init(x_empty + i, 8);
        }
        mbar_init(a_full, 4);
        mbar_init(a_free, 1);
        *abort_flag = 0;
        mbar_fence_init();
    }
    // zero region (and, for streamed weights, the ring: descriptor overruns must only ever see finite values)
    for (int i = threadIdx.x; i < (TS_ZERO + (RESIDENT ? 0 : W_BYTES)) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(RESIDENT ? zero : wsm)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (threadIdx.x == 0 && *tmem_slot != 0) *abort_flag = 1;   // the whole TMEM of this SM is ours: base must be 0
    constexpr uint32_t tmem_base = 0;
    __syncthreads();

    const int c = a.c, oc = a.oc;
    const int nitems = a.nsets * a.d_out * a.R * a.nsplit;
    uint32_t n_tile = 0, n_chunk = 0, n_wchunk = 0, n_x = 0, items_done = 0;   // per-role ring positions (persist across items)

    __shared__ int s_next[2];
    for (;; ++items_done) {
        if (threadIdx.x == 0) s_next[items_done & 1] = (int)atomicAdd(a.counter, 1u);
        __syncthreads();  // the previous item is finished by every role (its last epilogue waited for its last MMA)
        const int item = __shfl_sync(0xffffffffu, s_next[items_done & 1], 0);
        if (item >= nitems) break;
        const TsItem it = ts_item(a, item);
        const int nt = it.nt;
        if (__shfl_sync(0xffffffffu, *abort_flag, 0)) break;
        const uint8_t* img = a.wimg + (size_t)it.sr * IMG_BYTES;

        if (warp == W_WTMA) {
            // =============================================== weight TMA ===============================================
            if (RESIDENT) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(w_full, IMG_BYTES);
                    constexpr int CH = 16384;
                    for (int off = 0; off < IMG_BYTES; off += CH) bulk_g2s(wsm + off, img + off, min(CH, IMG_BYTES - off), w_full);
                }
                __syncwarp();
            } else {
                for (int t = 0; t < nt; ++t)
                    for (int ch = 0; ch < NCH; ++ch, ++n_wchunk) {
                        const uint32_t slot = n_wchunk % NSLOT, use = n_wchunk / NSLOT;
                        mbar_wait_sticky(w_empty + slot, (use & 1) ^ 1, abort_flag);   // the MMAs that read the slot are complete
                        if (lane == 0) {
                            mbar_arrive_expect_tx(w_full + slot, CHUNK_BYTES);
                            bulk_g2s(wsm + slot * CHUNK_BYTES, img + (size_t)ch * CHUNK_BYTES, CHUNK_BYTES, w_full + slot);
                        }
                        __syncwarp();
                    }
            }
        } else if (warp == W_XTMA) {
            // =============================================== x TMA ===============================================
            for (int t = 0; t < nt; ++t, ++n_tile) {
                const int tile = it.tile_begin + t;
                const int rows = min(128, a.Bs - tile * 128);
                const uint32_t bytes = (uint32_t)rows * c * 4;
#pragma unroll
                for (int child = 0; child < 2; ++child) {
                    // a missing child (odd scope count) takes no stage: its prologue warps substitute zeros.  (A stage that is
                    // completed without data would let those warps run two phases ahead of the barrier's parity.)
                    if (!(child ? it.has_r : it.has_l)) continue;
                    const uint32_t j = n_x++, slot = j % TS_NXS, use = j / TS_NXS;
                    mbar_wait_sticky(x_empty + slot, (use & 1) ^ 1, abort_flag);   // the prologue has taken its rows out of the slot
                    if (lane == 0) {
                        mbar_arrive_expect_tx(x_full + slot, bytes);
                        bulk_g2s(Xr + slot * XS_FLOATS,
                                 a.x + (((int64_t)(2 * it.s + child) * a.R + it.r) * a.Btot + it.row0 + (int64_t)tile * 128) * c, bytes,
                                 x_full + slot);
                    }
                    __syncwarp();
                }
            }
        } else if (warp == W_MMA) {
            // =============================================== MMA issuer ===============================================
            const uint32_t wsm_addr = smem_u32(wsm), zero_addr = smem_u32(zero);
            if (RESIDENT) {
                mbar_wait_sticky(w_full, items_done & 1, abort_flag);
                fence_after_sync();
            }
            constexpr uint32_t idesc = instr_desc_bf16(128, NCOLS, 0, 1);
            for (int t = 0; t < nt; ++t, ++n_tile) {
                mbar_wait_sticky(a_full, n_tile & 1, abort_flag);   // A tile of this row tile is in TMEM
                fence_after_sync();
                const uint32_t a_col = tmem_base + A_COL0;
#pragma unroll 1
                for (int ch = 0; ch < NCH; ++ch, ++n_chunk) {
                    const uint32_t buf = n_chunk % NB, use = n_chunk / NB;
                    mbar_wait_sticky(acc_empty + buf, (use & 1) ^ 1, abort_flag);     // the epilogue has read this buffer
                    uint32_t base = wsm_addr + ch * CHUNK_BYTES;
                    uint32_t slot = 0;
                    if (!RESIDENT) {
                        slot = n_wchunk % NSLOT;
                        mbar_wait_sticky(w_full + slot, (n_wchunk / NSLOT) & 1, abort_flag);
                        base = wsm_addr + slot * CHUNK_BYTES;
                        ++n_wchunk;
                    }
                    fence_after_sync();
                    const uint32_t acc = tmem_base + buf * NCOLS;
#pragma unroll
                    for (int kk = 0; kk < KSTEPS; ++kk) {
                        // B[K = r', N] MN-major: N groups 128 B apart, K groups (8 r') NL*CG*128 B apart.  The second K half of
                        // the last step does not exist when P % 16 == 8: it is read from the zero region.
                        constexpr uint32_t kstride = (uint32_t)NL * CG * 128;
                        const uint32_t start = base + 2 * kk * kstride;
                        const uint32_t lbo = (KP > P && kk == KSTEPS - 1) ? zero_addr - start : kstride;
                        const uint64_t desc = smem_desc(start, lbo, 128);
                        if (TS_DBG != 1) mma_ts_uniform(acc, a_col + kk * 8, desc, idesc, kk ? 1u : 0u);
                    }
                    mma_commit_uniform(acc_full + buf);
                    if (!RESIDENT) mma_commit_uniform(w_empty + slot);
                }
                mma_commit_uniform(a_free);   // the A tile may be overwritten with the next tile's operand
            }
        } else if (warp < 8) {
            // =============================================== prologue ===============================================
            const int child = warp >> 2;                    // 0: left, 1: right
            const int row = (warp & 3) * 32 + lane;
            const uint32_t a_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + A_COL0;
            const bool has_mine = child ? it.has_r : it.has_l;
            for (int t = 0; t < nt; ++t, ++n_tile) {
                const int tile = it.tile_begin + t;
                const uint32_t vb = n_tile & 1;
                const bool valid = (int64_t)tile * 128 + row < a.Bs;
                // Stages are counted over the children that exist.  EVERY prologue warp passes EVERY stage barrier in order (and
                // releases it), also the other child's: a warp that skipped a phase of a slot could not tell the phases
                // two uses apart by their parity.
                const uint32_t jl = n_x, jr = n_x + (it.has_l ? 1u : 0u);
                n_x += (it.has_l ? 1u : 0u) + (it.has_r ? 1u : 0u);
                const uint32_t j = child ? jr : jl, slot = j % TS_NXS;
                if (child == 1 && it.has_l) {   // the left child's stage comes first
                    mbar_wait_sticky(x_full + jl % TS_NXS, (jl / TS_NXS) & 1, abort_flag);
                    if (lane == 0) mbar_arrive(x_empty + jl % TS_NXS);
                }
                if (has_mine) mbar_wait_sticky(x_full + slot, (j / TS_NXS) & 1, abort_flag);
                float ev[P];
                float m;
                {
                    const float4* rowp = reinterpret_cast<const float4*>(Xr + slot * XS_FLOATS + row * c);
                    float m4[P / 4];
#pragma unroll
                    for (int q = 0; q < P / 4; ++q) {
                        float4 t4 = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);   // padded channel
                        if (4 * q < c) t4 = (valid && has_mine) ? rowp[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                        ev[4 * q] = t4.x; ev[4 * q + 1] = t4.y; ev[4 * q + 2] = t4.z; ev[4 * q + 3] = t4.w;
                        m4[q] = fmaxf(fmaxf(t4.x, t4.y), fmaxf(t4.z, t4.w));
                    }
                    // the rows are in registers: the TMA engine may refill the slot
                    fence_proxy_async();   // ordinary loads of the slot are ordered before the TMA engine's next write into it
                    __syncwarp();
                    if (lane == 0 && has_mine) mbar_arrive(x_empty + slot);
#pragma unroll
                    for (int st = 1; st < P / 4; st *= 2)   // tree reduction: short dependency chain
#pragma unroll
                        for (int q = 0; q + st < P / 4; q += 2 * st) m4[q] = fmaxf(m4[q], m4[q + st]);
                    m = m4[0];
                    m = (m == -CUDART_INF_F) ? 0.f : m;
                }
                uint32_t pk[NPAIR];
                {
                    const float ms = m * 1.4426950408889634f;
#pragma unroll
                    for (int q = 0; q < NPAIR; ++q) {
                        const float e0 = TS_SKIP_PRO ? ev[2 * q] : ex2_approx(fmaf(ev[2 * q], 1.4426950408889634f, -ms));
                        const float e1 = TS_SKIP_PRO ? ev[2 * q + 1] : ex2_approx(fmaf(ev[2 * q + 1], 1.4426950408889634f, -ms));
                        pk[q] = pack_bf16(e0, e1);
                    }
                }
                if (a.fhand_w) {   // hand-off: coalesced (lanes = consecutive rows of one packed word index)
                    uint8_t* fh = a.fhand_w + ((size_t)it.sr * a.ntiles + tile) * FH_BYTES;
                    uint32_t* dw = reinterpret_cast<uint32_t*>(fh + child * NPAIR * 512) + row;
#pragma unroll
                    for (int q = 0; q < NPAIR; ++q) dw[q * 128] = pk[q];
                    reinterpret_cast<float*>(fh + 2 * NPAIR * 512)[child * 128 + row] = m;
                }
                // buffer set vb: the epilogue (eL, shifts) of tile n - 2 is done with it
                mbar_wait_sticky(vec_empty + vb, ((n_tile >> 1) & 1) ^ 1, abort_flag);
                Ms[vb * 256 + child * 128 + row] = m;
                if (child == 0) {
                    uint32_t* el = ELp + vb * NPAIR * 128 + row;
#pragma unroll
                    for (int q = 0; q < NPAIR; ++q) el[q * 128] = pk[q];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(vec_full + vb);
                if (child == 0 && it.has_r) {   // pass the right child's stage (landed long ago)
                    mbar_wait_sticky(x_full + jr % TS_NXS, (jr / TS_NXS) & 1, abort_flag);
                    if (lane == 0) mbar_arrive(x_empty + jr % TS_NXS);
                }
                if (child == 1) {
                    // the A tile: the MMAs of the previous tile must have completed
                    mbar_wait_sticky(a_free, (n_tile & 1) ^ 1, abort_flag);
                    fence_after_sync();
                    {
                        uint32_t at[ACOLS];
#pragma unroll
                        for (int q = 0; q < ACOLS; ++q) at[q] = q < NPAIR ? pk[q] : 0u;
                        tmem_store_cols<ACOLS>(a_addr, at);
                    }
                    tmem_wait_st();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_full);
                }
            }
        } else {
            // =============================================== epilogue ===============================================
            const int quarter = (warp - 8) >> 2;
            const int row = (warp & 3) * 32 + lane;
            const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
            const int col0 = quarter * QC;             // this thread's outputs: o in [col0, col0 + QC)
            for (int t = 0; t < nt; ++t, ++n_tile) {
                const int tile = it.tile_begin + t;
                const uint32_t vb = n_tile & 1;
                const int64_t b = (int64_t)tile * 128 + row;
                const bool valid = b < a.Bs;
                mbar_wait_sticky(vec_full + vb, (n_tile >> 1) & 1, abort_flag);   // eL / shifts of this tile are in smem
                uint64_t acc2[QC / 2];
#pragma unroll
                for (int j = 0; j < QC / 2; ++j) acc2[j] = 0ull;
                const uint32_t* elp = ELp + vb * NPAIR * 128 + row;
#pragma unroll 1
                for (int ch = 0; ch < NCH; ++ch, ++n_chunk) {
                    const uint32_t buf = n_chunk % NB, use = n_chunk / NB;
                    mbar_wait_sticky(acc_full + buf, use & 1, abort_flag);
                    fence_after_sync();
                    uint32_t d[TC];
                    if (!TS_SKIP_EPI && TS_DBG != 4) ts_tmem_load<TC>(tmem_base + lane_base + buf * NCOLS + quarter * TC, d);
                    uint32_t elw[NL / 2];
#pragma unroll
                    for (int u = 0; u < NL / 2; ++u) elw[u] = elp[(ch * (NL / 2) + u) * 128];
                    tmem_wait_ld();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + buf);   // the accumulator may be overwritten
                    if (!TS_SKIP_EPI && TS_DBG != 3) {
#pragma unroll
                        for (int li = 0; li < NL; ++li) {
                            const float el = (li & 1) ? bf16_hi(elw[li >> 1]) : bf16_lo(elw[li >> 1]);
                            const uint64_t e2 = f2pack(el, el);
#pragma unroll
                            for (int j = 0; j < QC / 2; ++j)
                                acc2[j] = fma2(e2, u2pack(d[li * QC + 2 * j], d[li * QC + 2 * j + 1]), acc2[j]);
                        }
                    }
                }
                const float* my_m = Ms + vb * 256 + row;
                const float shift = my_m[0] + my_m[128];
                // ---- the output tile [rows][oc] is staged in shared memory and written by the TMA engine (one thread per row
                // writing 8-byte pieces of a strided row straight to global memory costs 32 sectors per instruction: measured
                // 0.3 ms of a 0.6 ms kernel).  Rows 0..63 go to this tile's (no longer needed) packed-eL buffer, rows 64..127
                // to a buffer of their own.
                float2 ov[QC / 2];
                bool bad = false;
#pragma unroll
                for (int j = 0; j < QC / 2; ++j) {   // before the barrier: the MUFU work overlaps the wait for the slowest warp
                    float s0, s1;
                    f2unpack(acc2[j], s0, s1);
                    bad |= valid && col0 + 2 * j < oc && !(fminf(s0, s1) >= 1e-30f);
                    ov[j] = make_float2(fmaf(lg2_approx(s0), 0.6931471805599453f, shift), fmaf(lg2_approx(s1), 0.6931471805599453f, shift));
                }
                ts_named_bar(1, 512);   // every epilogue warp is done with eL of this tile
                float* st = (row < 64 ? reinterpret_cast<float*>(ELp + vb * NPAIR * 128) + row * oc : OST + (row - 64) * oc) + col0;
#pragma unroll
                for (int j = 0; j < QC / 2; ++j)
                    if (col0 + 2 * j < oc) *reinterpret_cast<float2*>(st + 2 * j) = ov[j];
                if (bad && !TS_DBG) ts_exact_row(a, it.sg, it.s, it.r, it.row0 + b, col0, QC, st);
                fence_proxy_async();    // staging writes -> visible to the TMA engine
                ts_named_bar(1, 512);
                if (warp == 8 && lane == 0 && TS_DBG != 14) {
                    const int rows = min(128, a.Bs - tile * 128);
                    float* op = a.y + (((int64_t)it.s * a.R + it.r) * a.Btot + it.row0 + (int64_t)tile * 128) * oc;
                    bulk_s2g(op, ELp + vb * NPAIR * 128, (uint32_t)min(rows, 64) * oc * 4);
                    if (rows > 64) bulk_s2g(op + 64 * oc, OST, (uint32_t)(rows - 64) * oc * 4);
                    bulk_commit_group();
                    bulk_wait_group_read<0>();   // both staging buffers have been read
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(vec_empty + vb);   // eL / shifts / staging of this buffer set are free
            }
            if (warp == 8 && lane == 0) bulk_wait_group<0>();
        }
    }
    // teardown
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == W_MMA) tmem_dealloc(tmem_base, 512);
    if (threadIdx.x == 0) tc_abort_trap(abort_flag, &g_tc3_aborts);
}

// ====================================================================================================================
// dX
// ====================================================================================================================
// Thread organisation (11 warps, <= 184 registers per thread):
//   warps 0-7   EPILOGUE: two threads per row (same lane, the two warps of one TMEM lane quadrant): h = lower / upper half of
//               the right channels, all left channels.  Every accumulator element dE[l, r'] is loaded ONCE and used twice:
//               aR[r'] += eL[l] * dE (kept in registers over the tile: complete when the tile ends) and
//               aL[l] += eR[r'] * dE (a dot product over the thread's right channels; the two halves combine their partial
//               sums through shared memory, plain stores -- shared-memory fp32 atomics are CAS loops).
//   warp 8      MMA issuer (SS: the G tile in shared memory is the A operand).
//   warp 9      weight TMA.    warp 10   TMA of the per-tile operands: G tile (double buffered), packed eL (double
//               buffered) and eR (single buffer: the epilogue threads take their values into registers at once).
template <int P>
__global__ void __launch_bounds__(352, 1) tsum_dx_kernel(TsArgs a) {
    constexpr int CG = P / 8;
    constexpr int KP = ts_kp(P);
    constexpr int KSTEPS = KP / 16;
    constexpr int KG = KP / 8;                 // K groups of the (padded) A operand tile
    constexpr int NL = ts_nl(P);
    constexpr int NCH = P / NL;
    constexpr int NCOLS = NL * P;
    constexpr int NB = ts_nb(P);
    constexpr int HC = P / 2;                  // right channels per epilogue thread
    constexpr int TC = NL * HC;                // accumulator columns per epilogue thread and chunk
    constexpr int NPAIR = P / 2;
    constexpr bool RESIDENT = ts_resident(P);
    constexpr int NSLOT = ts_nslot(P);
    constexpr int CHUNK_BYTES = NL * P * P * 2;
    constexpr int IMG_BYTES = P * P * P * 2;
    constexpr int W_BYTES = ts_wbytes(P);
    constexpr int FH_BYTES = 2 * NPAIR * 512 + 1024;
    constexpr int GS_BYTES = 16 * KG * 128;            // G tile: [16 row groups][KG][8 rows][8 bf16], pad K groups are zero
    static_assert(P % 8 == 0 && P >= 8 && P <= 64 && NL % 2 == 0 && P % NL == 0, "unsupported channel count");
    static_assert(NB >= 2 && NB * NCOLS <= 512 && HC % 4 == 0 && TC <= 128, "TMEM / register budget");
    constexpr int W_MMA = 8, W_WTMA = 9, W_VTMA = 10;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* wsm = smem;
    uint8_t* zero = smem + W_BYTES;
    uint8_t* Gs = zero + TS_ZERO;                                             // [2][GS_BYTES]
    uint32_t* Us = reinterpret_cast<uint32_t*>(Gs + 2 * GS_BYTES);            // [2][NPAIR][128] packed eL
    float* SL = reinterpret_cast<float*>(Us + 2 * NPAIR * 128);               // [2 halves][P][128] partial aL sums
    uint64_t* u_full = reinterpret_cast<uint64_t*>(SL + 2 * P * 128);         // [2] TMA -> epilogue
    uint64_t* u_empty = u_full + 2;                                           // [2] epilogue -> TMA
    uint64_t* g_full = u_empty + 2;                                           // [2] TMA -> MMA
    uint64_t* g_empty = g_full + 2;                                           // [2] MMA commit -> TMA
    uint64_t* acc_full = g_empty + 2;                                         // [NB]
    uint64_t* acc_empty = acc_full + 4;                                       // [NB]
    uint64_t* w_full = acc_empty + 4;                                         // [NSLOT]
    uint64_t* w_empty = w_full + 4;                                           // [NSLOT]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + 4);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(u_full + i, 1);
            mbar_init(u_empty + i, 8);
            mbar_init(g_full + i, 1);
            mbar_init(g_empty + i, 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(acc_full + i, 1);
            mbar_init(acc_empty + i, 8);
            mbar_init(w_full + i, 1);
            mbar_init(w_empty + i, 1);
        }
        *abort_flag = 0;
        mbar_fence_init();
    }
    // zero: the zero region and, for streamed weights, the ring (descriptor overruns must only ever see finite values)
    for (int i = threadIdx.x; i < TS_ZERO / 16; i += blockDim.x) reinterpret_cast<uint4*>(zero)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (!RESIDENT)
        for (int i = threadIdx.x; i < W_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(wsm)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (threadIdx.x == 0 && *tmem_slot != 0) *abort_flag = 1;
    constexpr uint32_t tmem_base = 0;
    __syncthreads();

    const int c = a.c;
    const int nitems = a.nsets * a.d_out * a.R * a.nsplit;
    uint32_t n_tile = 0, n_chunk = 0, n_wchunk = 0, items_done = 0;

    __shared__ int s_next[2];
    for (;; ++items_done) {
        if (threadIdx.x == 0) s_next[items_done & 1] = (int)atomicAdd(a.counter, 1u);
        __syncthreads();
        const int item = __shfl_sync(0xffffffffu, s_next[items_done & 1], 0);
        if (item >= nitems) break;
        const TsItem it = ts_item(a, item);
        const int nt = it.nt;
        if (__shfl_sync(0xffffffffu, *abort_flag, 0)) break;
        const uint8_t* img = a.wimg + (size_t)it.sr * IMG_BYTES;

        if (warp == W_WTMA) {
            // =============================================== weight TMA ===============================================
            if (RESIDENT) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(w_full, IMG_BYTES);
                    constexpr int CH = 16384;
                    for (int off = 0; off < IMG_BYTES; off += CH) bulk_g2s(wsm + off, img + off, min(CH, IMG_BYTES - off), w_full);
                }
                __syncwarp();
            } else {
                for (int t = 0; t < nt; ++t)
                    for (int ch = 0; ch < NCH; ++ch, ++n_wchunk) {
                        const uint32_t slot = n_wchunk % NSLOT, use = n_wchunk / NSLOT;
                        mbar_wait_sticky(w_empty + slot, (use & 1) ^ 1, abort_flag);
                        if (lane == 0) {
                            mbar_arrive_expect_tx(w_full + slot, CHUNK_BYTES);
                            bulk_g2s(wsm + slot * CHUNK_BYTES, img + (size_t)ch * CHUNK_BYTES, CHUNK_BYTES, w_full + slot);
                        }
                        __syncwarp();
                    }
            }
        } else if (warp == W_VTMA) {
            // =============================================== per-tile operand TMA ===============================================
            for (int t = 0; t < nt; ++t, ++n_tile) {
                const int tile = it.tile_begin + t;
                const uint32_t vb = n_tile & 1;
                const size_t ti = (size_t)it.sr * a.ntiles + tile;
                const uint8_t* fh = a.fhand_r + ti * FH_BYTES;
                // G tile (A operand): the MMAs of tile n - 2 are complete
                mbar_wait_sticky(g_empty + vb, ((n_tile >> 1) & 1) ^ 1, abort_flag);
                if (lane == 0) {
                    mbar_arrive_expect_tx(g_full + vb, GS_BYTES);
                    bulk_g2s(Gs + vb * GS_BYTES, a.hand + ti * GS_BYTES, GS_BYTES, g_full + vb);
                }
                __syncwarp();
                // packed eL (read all through the tile; the epilogue threads fetch their eR values themselves)
                mbar_wait_sticky(u_empty + vb, ((n_tile >> 1) & 1) ^ 1, abort_flag);
                if (lane == 0) {
                    mbar_arrive_expect_tx(u_full + vb, NPAIR * 512);
                    bulk_g2s(Us + vb * NPAIR * 128, fh, NPAIR * 512, u_full + vb);
                }
                __syncwarp();
            }
        } else if (warp == W_MMA) {
            // =============================================== MMA issuer ===============================================
            const uint32_t wsm_addr = smem_u32(wsm), gs_addr = smem_u32(Gs);
            if (RESIDENT) {
                mbar_wait_sticky(w_full, items_done & 1, abort_flag);
                fence_after_sync();
            }
            constexpr uint32_t idesc = instr_desc_bf16(128, NCOLS, 0, 0);
            for (int t = 0; t < nt; ++t, ++n_tile) {
                const uint32_t vb = n_tile & 1;
                mbar_wait_sticky(g_full + vb, (n_tile >> 1) & 1, abort_flag);   // G tile of this row tile is in shared memory
                fence_after_sync();
                const uint32_t ga = gs_addr + vb * GS_BYTES;
#pragma unroll 1
                for (int ch = 0; ch < NCH; ++ch, ++n_chunk) {
                    const uint32_t buf = n_chunk % NB, use = n_chunk / NB;
                    mbar_wait_sticky(acc_empty + buf, (use & 1) ^ 1, abort_flag);
                    uint32_t base = wsm_addr + ch * CHUNK_BYTES;
                    uint32_t slot = 0;
                    if (!RESIDENT) {
                        slot = n_wchunk % NSLOT;
                        mbar_wait_sticky(w_full + slot, (n_wchunk / NSLOT) & 1, abort_flag);
                        base = wsm_addr + slot * CHUNK_BYTES;
                        ++n_wchunk;
                    }
                    fence_after_sync();
                    const uint32_t acc = tmem_base + buf * NCOLS;
#pragma unroll
                    for (int kk = 0; kk < KSTEPS; ++kk) {
                        // A[M = row, K = o] K-major: K halves 128 B apart, 8-row groups KG*128 B apart (pad K group = zeros).
                        // B[K = o, N] K-major: N groups (8 consecutive n) CG*128 B apart, K halves 128 B apart; a padded K half
                        // (P % 16 == 8) reads the next block -- finite image values times the A tile's zeros.
                        const uint64_t adesc = smem_desc(ga + 2 * kk * 128, 128, KG * 128);
                        const uint64_t bdesc = smem_desc(base + 2 * kk * 128, 128, CG * 128);
                        if (TS_DBG != 1) mma_ss_uniform(acc, adesc, bdesc, idesc, kk ? 1u : 0u);
                    }
                    mma_commit_uniform(acc_full + buf);
                    if (!RESIDENT) mma_commit_uniform(w_empty + slot);
                }
                mma_commit_uniform(g_empty + vb);   // the G stage may be refilled
            }
        } else if (warp < 8) {
            // =============================================== epilogue ===============================================
            const int h = warp >> 2;
            const int row = (warp & 3) * 32 + lane;
            const uint32_t my_col = ((uint32_t)((warp & 3) * 32) << 16) + h * (NL * HC);
            TSP_DECL;
            // This half's eR values come straight from the forward's hand-off in global memory (coalesced: lanes = consecutive
            // rows of one packed word), one tile ahead: the loads of tile t + 1 are in flight during tile t.
            auto er_ptr = [&](int tile) {
                return reinterpret_cast<const uint32_t*>(a.fhand_r + ((size_t)it.sr * a.ntiles + tile) * FH_BYTES + NPAIR * 512) +
                       (h * (HC / 2)) * 128 + row;
            };
            uint32_t er_next[HC / 2];
            if (nt > 0) {
                const uint32_t* ep = er_ptr(it.tile_begin);
#pragma unroll
                for (int j = 0; j < HC / 2; ++j) er_next[j] = __ldg(ep + j * 128);
            }
            for (int t = 0; t < nt; ++t, ++n_tile) {
                const int tile = it.tile_begin + t;
                const uint32_t vb = n_tile & 1;
                uint64_t er2[HC / 2];
#pragma unroll
                for (int j = 0; j < HC / 2; ++j) er2[j] = f2pack(bf16_lo(er_next[j]), bf16_hi(er_next[j]));
                if (t + 1 < nt) {
                    const uint32_t* ep = er_ptr(tile + 1);
#pragma unroll
                    for (int j = 0; j < HC / 2; ++j) er_next[j] = __ldg(ep + j * 128);
                }
                TSP_LAP(0);   // er unpack / prefetch issue
                mbar_wait_sticky(u_full + vb, (n_tile >> 1) & 1, abort_flag);
                TSP_LAP(1);   // wait for eL
                const uint32_t* up = Us + vb * NPAIR * 128 + row;
                float* slp = SL + h * P * 128 + row;
                uint64_t ar2[HC / 2];
#pragma unroll
                for (int j = 0; j < HC / 2; ++j) ar2[j] = 0ull;
                // Chunk loop.  tcgen05.ld has ~400 clk of latency and tcgen05.wait::ld waits for ALL loads of the thread, so
                // the kernel is bound by the accumulator bytes in flight: every thread fetches its whole share of a chunk
                // (NL x P/2 columns) with one batch of loads, then issues the FFMA2s; the two warps of a scheduler alternate.
#pragma unroll 1
                for (int ch = 0; ch < NCH; ++ch, ++n_chunk) {
                    const uint32_t buf = n_chunk % NB;
                    mbar_wait_sticky(acc_full + buf, (n_chunk / NB) & 1, abort_flag);
                    fence_after_sync();
                    TSP_LAP(ch == 0 ? 8 : 2);   // wait for the accumulator (first chunk of the tile / others)
                    uint32_t d[TC];
                    if (!TS_SKIP_EPI && TS_DBG != 4) ts_tmem_load<TC>(tmem_base + buf * NCOLS + my_col, d);
                    const int l0 = ch * NL;
                    uint32_t elw[NL / 2];
#pragma unroll
                    for (int u = 0; u < NL / 2; ++u) elw[u] = up[((l0 >> 1) + u) * 128];
                    tmem_wait_ld();
                    TSP_LAP(3);   // tcgen05.ld issue + latency
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + buf);   // the accumulator may be overwritten
                    TSP_LAP(4);   // release
                    if (TS_SKIP_EPI || TS_DBG == 3) continue;
#pragma unroll
                    for (int u = 0; u < NL; ++u) {
                        const float el = (u & 1) ? bf16_hi(elw[u >> 1]) : bf16_lo(elw[u >> 1]);
                        const uint64_t e2 = f2pack(el, el);
                        uint64_t s0 = 0ull, s1 = 0ull;
#pragma unroll
                        for (int j = 0; j < HC / 2; ++j) {
                            const uint64_t d2 = u2pack(d[u * HC + 2 * j], d[u * HC + 2 * j + 1]);
                            ar2[j] = fma2(e2, d2, ar2[j]);
                            if (j & 1) s1 = fma2(er2[j], d2, s1); else s0 = fma2(er2[j], d2, s0);
                        }
                        float p0, p1, p2, p3;
                        f2unpack(s0, p0, p1);
                        f2unpack(s1, p2, p3);
                        slp[(l0 + u) * 128] = (p0 + p1) + (p2 + p3);   // this half's share of aL[l]
                    }
                    TSP_LAP(5);   // FFMA2s
                }
                // ---- tile results: dxR[r'] = eR[r'] * aR[r'], dxL[l] = eL[l] * aL[l] (this half's channels).  A thread owns 80-byte
                // pieces of strided rows (32 sectors per store instruction if written directly: measured 0.37 ms of a 0.8 ms
                // kernel), and a TMA bulk store or a single copy warp needs > 2000 clk for the 40 KB of a tile, so the two
                // [rows][c] tiles are staged in shared memory (in the share buffer) and copied out by all epilogue threads
                // with coalesced 16-byte accesses: shares complete -> everybody reads its sums -> everybody stages its results
                // -> everybody copies -> the buffer is free for the next tile's shares.
                ts_named_bar(1, 256);   // both halves' shares of aL are in shared memory
                TSP_LAP(10);   // first tile-end barrier
                float xl[HC];
#pragma unroll
                for (int j2 = 0; j2 < HC / 2; ++j2) {
                    const int l = h * HC + 2 * j2;
                    const float* sp = SL + l * 128 + row;
                    const uint32_t w = up[(l >> 1) * 128];
                    xl[2 * j2] = bf16_lo(w) * (sp[0] + sp[P * 128]);
                    xl[2 * j2 + 1] = bf16_hi(w) * (sp[128] + sp[P * 128 + 128]);
                }
                fence_proxy_async();    // ordinary loads of eL are ordered before the TMA engine's next write into the buffer
                ts_named_bar(1, 256);   // the shares (and eL) have been read
                if (lane == 0) mbar_arrive(u_empty + vb);
                {
                    float* ol = SL + row * c + h * HC;              // staging: [2 children][128 rows][c]
                    float* orr = ol + 128 * c;
#pragma unroll
                    for (int j4 = 0; j4 < HC / 4; ++j4) {
                        const uint64_t p0 = mul2(er2[2 * j4], ar2[2 * j4]), p1 = mul2(er2[2 * j4 + 1], ar2[2 * j4 + 1]);
                        float v0, v1, v2, v3;
                        f2unpack(p0, v0, v1);
                        f2unpack(p1, v2, v3);
                        if (h * HC + 4 * j4 < c) {
                            *reinterpret_cast<float4*>(orr + 4 * j4) = make_float4(v0, v1, v2, v3);
                            *reinterpret_cast<float4*>(ol + 4 * j4) = make_float4(xl[4 * j4], xl[4 * j4 + 1], xl[4 * j4 + 2], xl[4 * j4 + 3]);
                        }
                    }
                }
                ts_named_bar(1, 256);   // the results are staged
                if (TS_DBG != 14) {
                    const int rows = min(128, a.Bs - tile * 128);
                    const int n4 = rows * c / 4;                   // float4 per child tile
                    const int64_t r0 = it.row0 + (int64_t)tile * 128;
                    const int tid = threadIdx.x;                   // 0 .. 255
#pragma unroll
                    for (int child = 0; child < 2; ++child) {
                        if (!(child ? it.has_r : it.has_l)) continue;
                        const float4* src = reinterpret_cast<const float4*>(SL + child * 128 * c);
                        float4* dst = reinterpret_cast<float4*>(a.y + (((int64_t)(2 * it.s + child) * a.R + it.r) * a.Btot + r0) * c);
                        float4 v[P / 8];                           // 128 * P / 4 float4 over 256 threads
#pragma unroll
                        for (int k = 0; k < P / 8; ++k)
                            if (tid + k * 256 < n4) v[k] = src[tid + k * 256];
#pragma unroll
                        for (int k = 0; k < P / 8; ++k)
                            if (tid + k * 256 < n4) dst[tid + k * 256] = v[k];
                    }
                }
                ts_named_bar(1, 256);   // the staging area (= the share buffer) may be written again
                TSP_LAP(6);   // tile end
#ifdef SPN_TC_PROFILE
                tsp_acc[7] += 1;   // tiles
#endif
            }
            TSP_FLUSH();
        }
    }
    // teardown
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == W_MMA) tmem_dealloc(tmem_base, 512);
    if (threadIdx.x == 0) tc_abort_trap(abort_flag, &g_tc3_aborts);
}

// G pre-pass of the backward (one block per (set, scope, repetition, 128-row tile)): G = g * exp(mL + mR - out) = g / (linear
// -domain sum) as packed bf16 in the core-matrix layout [row / 8][o / 8][row % 8][8 o] that both the dX kernel (A operand,
// K-major) and the dW kernel (B operand, MN-major) read straight from shared memory after ONE bulk copy per tile; the o-groups
// are padded to KG = (multiple of 16) / 8 with zero groups (the MMAs' K resp. N extent).  out = -inf (impossible row): no
// gradient; the exponent is capped so that rows whose sum underflowed in the forward (rescue path) stay finite.  Also
// accumulates the column sums csum[sg][o][r] = sum_b g[b, o] over the rows with out > -inf, which the fused softmax backward
// of dW needs (the rows of a tile are contiguous, so the loads are coalesced 16-byte accesses; the shifts come from the
// forward's hand-off).
__global__ void __launch_bounds__(128) tc_gprep_kernel(const float* __restrict__ out, const float* __restrict__ g,
                                                      const uint8_t* __restrict__ fhand, uint8_t* __restrict__ hand,
                                                      float* __restrict__ csum, int Bs, int Btot, int d_out, int oc, int R,
                                                      int P, int ntiles) {
    extern __shared__ __align__(16) uint8_t gsm[];
    const int NPAIR = P / 2, KG = ts_kp(P) / 8;
    uint4* Gt = reinterpret_cast<uint4*>(gsm);                          // [16][KG][8] 16-byte rows
    float* sh = reinterpret_cast<float*>(gsm + 16 * KG * 128);          // [128] mL + mR (in log2 units)
    float* cs = sh + 128;                                               // [P] column sums of this tile
    const int tile = blockIdx.x % ntiles;
    const int sr = blockIdx.x / ntiles;
    const int r = sr % R, sg = sr / R;
    const int set = sg / d_out, s = sg % d_out;
    const int rows = min(128, Bs - tile * 128);
    const int FH_BYTES = 2 * NPAIR * 512 + 1024;
    const float* shp = reinterpret_cast<const float*>(fhand + ((size_t)sr * ntiles + tile) * FH_BYTES + 2 * NPAIR * 512);
    sh[threadIdx.x] = (shp[threadIdx.x] + shp[128 + threadIdx.x]) * 1.4426950408889634f;
    for (int i = threadIdx.x; i < P; i += 128) cs[i] = 0.f;
    for (int i = threadIdx.x; i < 16 * KG * 8; i += 128) Gt[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    const int64_t base = (((int64_t)s * R + r) * Btot + (int64_t)set * Bs + (int64_t)tile * 128) * oc;
    const float4* op = reinterpret_cast<const float4*>(out + base);
    const float4* gp = reinterpret_cast<const float4*>(g + base);
    const int q4 = oc / 4;                                              // float4 chunks per row
    uint2* Gt2 = reinterpret_cast<uint2*>(gsm);
    for (int i = threadIdx.x; i < rows * q4; i += 128) {
        const int row = i / q4, o = 4 * (i - row * q4);
        const float4 ov = __ldg(op + i), gv = __ldg(gp + i);
        const float sh2 = sh[row];
        const float o_[4] = {ov.x, ov.y, ov.z, ov.w}, g_[4] = {gv.x, gv.y, gv.z, gv.w};
        float G[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool live = o_[u] != -CUDART_INF_F;
            G[u] = live ? g_[u] * tc::ex2_approx(fminf(fmaf(o_[u], -1.4426950408889634f, sh2), 115.f)) : 0.f;
            if (csum && live) atomicAdd(cs + o + u, g_[u]);
        }
        // 8-byte half of the 16-byte row (row, o / 8)
        Gt2[(((row >> 3) * KG + (o >> 3)) * 8 + (row & 7)) * 2 + ((o >> 2) & 1)] = make_uint2(tc::pack_bf16(G[0], G[1]), tc::pack_bf16(G[2], G[3]));
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(hand + ((size_t)sr * ntiles + tile) * (size_t)(16 * KG * 128));
    for (int i = threadIdx.x; i < 16 * KG * 8; i += 128) dst[i] = Gt[i];
    if (csum)
        for (int o = threadIdx.x; o < oc; o += 128) atomicAdd(csum + ((int64_t)sg * oc + o) * R + r, cs[o]);
}

// ---------------------------------------------------------------------------------------------------------------
// Weight operand preparation
// ---------------------------------------------------------------------------------------------------------------
// Log-normaliser of the unnormalised weights: logz[s][o][r] = logsumexp_k param[s][k][o][r]  (F.log_softmax(dim=2) of
// layers.py:85 is then lw = param - logz, applied by the consumers).  Block = 64 consecutive (o, r) columns (16 threads
// x float4 = one 256-byte segment per row) x 16 k-phases, four rows in flight per thread (16 KB per block); online
// max / sum per column so the parameters are read once.  Requires ocr % 4 == 0.
__device__ __forceinline__ void logz_update(float& m, float& acc, float v0, float v1, float v2, float v3) {
    const float mn = fmaxf(fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)), m);
    if (mn != -CUDART_INF_F) {
        acc = acc * __expf(m - mn) + __expf(v0 - mn) + __expf(v1 - mn) + __expf(v2 - mn) + __expf(v3 - mn);
        m = mn;
    }
}

__global__ void __launch_bounds__(256) tc_logz_kernel(const float* __restrict__ param, float* __restrict__ logz, int K,
                                                     int ocr, int tiles_per_s) {
    __shared__ float sm_m[16][64], sm_s[16][64];
    const int s = blockIdx.x / tiles_per_s;
    const int cq = threadIdx.x & 15, ph = threadIdx.x >> 4;
    const int col = (blockIdx.x % tiles_per_s) * 64 + cq * 4;
    float m[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (col < ocr) {
        const float4* p = reinterpret_cast<const float4*>(param + (int64_t)s * K * ocr + col);
        const int64_t rs = ocr / 4;
        int k = ph;
        for (; k + 48 < K; k += 64) {
            const float4 a0 = __ldg(p + (int64_t)k * rs), a1 = __ldg(p + (int64_t)(k + 16) * rs),
                         a2 = __ldg(p + (int64_t)(k + 32) * rs), a3 = __ldg(p + (int64_t)(k + 48) * rs);
            logz_update(m[0], acc[0], a0.x, a1.x, a2.x, a3.x);
            logz_update(m[1], acc[1], a0.y, a1.y, a2.y, a3.y);
            logz_update(m[2], acc[2], a0.z, a1.z, a2.z, a3.z);
            logz_update(m[3], acc[3], a0.w, a1.w, a2.w, a3.w);
        }
        for (; k < K; k += 16) {
            const float4 a0 = __ldg(p + (int64_t)k * rs);
            const float ninf = -CUDART_INF_F;
            logz_update(m[0], acc[0], a0.x, ninf, ninf, ninf);
            logz_update(m[1], acc[1], a0.y, ninf, ninf, ninf);
            logz_update(m[2], acc[2], a0.z, ninf, ninf, ninf);
            logz_update(m[3], acc[3], a0.w, ninf, ninf, ninf);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sm_m[ph][cq * 4 + j] = m[j];
        sm_s[ph][cq * 4 + j] = acc[j];
    }
    __syncthreads();
    const int cc = threadIdx.x;
    const int ocol = (blockIdx.x % tiles_per_s) * 64 + cc;
    if (cc < 64 && ocol < ocr) {
        float mm = -CUDART_INF_F;
#pragma unroll
        for (int q = 0; q < 16; ++q) mm = fmaxf(mm, sm_m[q][cc]);
        float tot = 0.f;
        if (mm != -CUDART_INF_F) {
#pragma unroll
            for (int q = 0; q < 16; ++q) tot += sm_s[q][cc] * __expf(sm_m[q][cc] - mm);
        }
        logz[(int64_t)s * ocr + ocol] = mm == -CUDART_INF_F ? -CUDART_INF_F : mm + logf(tot);
    }
}

// 16-byte row (8 outputs o = 8 og .. 8 og + 7 of in-channel pair (l, r')) of image A inside one (set, scope, repetition):
// chunk-major; inside a chunk the K-major core-matrix grid [n / 8][og][n % 8] with n = the dX kernel's column order.
__host__ __device__ __forceinline__ int ts_image_a_row(int P, int NL, int l, int rp, int og) {
    const int CG = P / 8, HC = P / 2;
    const int n = (rp / HC) * (NL * HC) + (l % NL) * HC + rp % HC;
    return (l / NL) * (NL * P * CG) + ((n >> 3) * CG + og) * 8 + (n & 7);
}

// Image A: fp32 (log-)weights [S][K = c*c][oc][R] (S = weight sets x scopes) -> bf16 exp() core-matrix image
// [S][R][kg = l*CG + rg][og][8 (k = r' % 8)][8 (o)], CG = P / 8 groups per padded channel axis.  Fast variant for c and oc
// multiples of 8: one block per (scope, group of 8 consecutive k): the [8][oc][R] slab is read with coalesced float4 loads
// into a padded shared-memory tile [8][oc][R|1] (+ pad so that the k-stride is 9 mod 32 banks); then every thread assembles
// one 16-byte image row (8 outputs of one k), consecutive threads taking consecutive rows of one repetition so that both the
// shared-memory reads (conflict free) and the global stores (contiguous per repetition) are efficient.  Padding blocks
// (P > c or P > oc) are zero-filled by a memset before the launch.
__global__ void __launch_bounds__(256) tc_wimage_kernel(const float* __restrict__ lw, const float* __restrict__ logz,
                                                       uint8_t* __restrict__ wimg, int c, int OC, int R, int P, int NL, int RP,
                                                       int KS) {
    extern __shared__ float tile[];  // [8][KS], row (kk, o) at kk * KS + o * RP
    const int K = c * c, CG = P / 8, cg = c / 8;
    const int kgr = blockIdx.x % (K / 8);   // real k group: l * (c / 8) + rg
    const int s = blockIdx.x / (K / 8);
    const int l = kgr / cg, rg = kgr % cg;
    const int ocr = OC * R;
    const float4* src = reinterpret_cast<const float4*>(lw + ((int64_t)s * K + (int64_t)kgr * 8) * ocr);
    const float4* zp = logz ? reinterpret_cast<const float4*>(logz + (int64_t)s * ocr) : nullptr;
    const int n4 = 2 * ocr;  // 8 * ocr / 4
    for (int i4 = threadIdx.x; i4 < n4; i4 += blockDim.x) {
        const int i = i4 * 4;
        const int kk = i / ocr, rem = i - kk * ocr;
        const float4 v = __ldg(src + i4);
        const float4 z = zp ? __ldg(zp + rem / 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float e[4] = {__expf(v.x - z.x), __expf(v.y - z.y), __expf(v.z - z.z), __expf(v.w - z.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // (OC * R) % 4 == 0, so the four values share kk; R itself may be odd
            const int o = (rem + j) / R, r = rem + j - o * R;
            tile[kk * KS + o * RP + r] = e[j];
        }
    }
    __syncthreads();
    const int NOG = OC / 8;
    uint4* dst = reinterpret_cast<uint4*>(wimg);
    // work item = (r, og, kk) with kk fastest: chunk index inside (s, r) is (kg * CG + og) * 8 + kk
    for (int i = threadIdx.x; i < NOG * 8 * R; i += blockDim.x) {
        const int kk = i & 7, og = (i >> 3) % NOG, r = i / (8 * NOG);
        const float* t = tile + kk * KS + og * 8 * RP + r;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(t[0], t[RP]), p1 = __floats2bfloat162_rn(t[2 * RP], t[3 * RP]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(t[4 * RP], t[5 * RP]), p3 = __floats2bfloat162_rn(t[6 * RP], t[7 * RP]);
        uint4 v;
        v.x = *reinterpret_cast<uint32_t*>(&p0); v.y = *reinterpret_cast<uint32_t*>(&p1);
        v.z = *reinterpret_cast<uint32_t*>(&p2); v.w = *reinterpret_cast<uint32_t*>(&p3);
        dst[((int64_t)s * R + r) * ((int64_t)P * CG * CG * 8) + ts_image_a_row(P, NL, l, rg * 8 + kk, og)] = v;
    }
}

// Same for wide layers (the [8][oc][R] slab would need > 64 KB of shared memory: one block per SM, measured 0.56 TB/s at
// oc = R = 64): a block takes a chunk of RC = 16 repetitions of the slab (64-byte runs: whole sectors), blockIdx.y = chunk.
// Requires R % 4 == 0.
__global__ void __launch_bounds__(256) tc_wimage_rchunk_kernel(const float* __restrict__ lw, const float* __restrict__ logz,
                                                              uint8_t* __restrict__ wimg, int c, int OC, int R, int P, int NL) {
    constexpr int RC = 16, RP = RC + 1;
    extern __shared__ float tile[];  // [8][KS], row (kk, o) at kk * KS + o * RP
    const int KS = OC * RP + ((9 - (OC * RP) % 32) + 32) % 32;
    const int K = c * c, CG = P / 8, cg = c / 8;
    const int kgr = blockIdx.x % (K / 8);
    const int s = blockIdx.x / (K / 8);
    const int l = kgr / cg, rg = kgr % cg;
    const int r0 = blockIdx.y * RC, rc = min(RC, R - r0), q4 = rc / 4;
    const float* src = lw + ((int64_t)s * K + (int64_t)kgr * 8) * OC * R + r0;
    const float* zp = logz ? logz + (int64_t)s * OC * R + r0 : nullptr;
    for (int i = threadIdx.x; i < 8 * OC * q4; i += blockDim.x) {
        const int q = i % q4, ko = i / q4;          // ko = kk * OC + o
        const int o = ko % OC, kk = ko / OC;
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)ko * R) + q);
        const float4 z = zp ? __ldg(reinterpret_cast<const float4*>(zp + (int64_t)o * R) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        float* t = tile + kk * KS + o * RP + 4 * q;
        t[0] = __expf(v.x - z.x); t[1] = __expf(v.y - z.y); t[2] = __expf(v.z - z.z); t[3] = __expf(v.w - z.w);
    }
    __syncthreads();
    const int NOG = OC / 8;
    uint4* dst = reinterpret_cast<uint4*>(wimg);
    for (int i = threadIdx.x; i < NOG * 8 * rc; i += blockDim.x) {
        const int kk = i & 7, og = (i >> 3) % NOG, rr = i / (8 * NOG);
        const float* t = tile + kk * KS + og * 8 * RP + rr;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(t[0], t[RP]), p1 = __floats2bfloat162_rn(t[2 * RP], t[3 * RP]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(t[4 * RP], t[5 * RP]), p3 = __floats2bfloat162_rn(t[6 * RP], t[7 * RP]);
        uint4 v;
        v.x = *reinterpret_cast<uint32_t*>(&p0); v.y = *reinterpret_cast<uint32_t*>(&p1);
        v.z = *reinterpret_cast<uint32_t*>(&p2); v.w = *reinterpret_cast<uint32_t*>(&p3);
        dst[((int64_t)s * R + r0 + rr) * ((int64_t)P * CG * CG * 8) + ts_image_a_row(P, NL, l, rg * 8 + kk, og)] = v;
    }
}

// Image A for channel counts that are multiples of 4 only: one thread per 16-byte image row (strided reads; such layers are small).
__global__ void __launch_bounds__(256) tc_wimage_generic_kernel(const float* __restrict__ lw, const float* __restrict__ logz,
                                                               uint4* __restrict__ wimg, int S, int c, int OC, int R, int P, int NL) {
    const int CG = P / 8;
    const int64_t per_sr = (int64_t)P * CG * CG * 8;   // 16-byte rows per (s, r)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)S * R * per_sr) return;
    const int64_t srr = i / per_sr;
    const int e = (int)(i % per_sr);
    const int kk = e % 8, og = (e / 8) % CG, kg = e / (8 * CG);
    const int l = kg / CG, rp = (kg % CG) * 8 + kk;
    const int s = (int)(srr / R), r = (int)(srr % R);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int o = og * 8 + j;
        float val = 0.f;
        if (l < c && rp < c && o < OC) {
            const float w = lw[(((int64_t)s * c * c + (l * c + rp)) * OC + o) * R + r];
            val = __expf(w - (logz ? logz[((int64_t)s * OC + o) * R + r] : 0.f));
        }
        v[j] = val;
    }
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 q;
    q.x = *reinterpret_cast<uint32_t*>(&p0); q.y = *reinterpret_cast<uint32_t*>(&p1);
    q.z = *reinterpret_cast<uint32_t*>(&p2); q.w = *reinterpret_cast<uint32_t*>(&p3);
    wimg[srr * per_sr + ts_image_a_row(P, NL, l, rp, og)] = q;
}

// Image F (forward) from image A: per chunk the MN-major core-matrix grid [rg = r' / 8][n / 8][r' % 8][8 n] with n = the
// forward kernel's column order (quarter of the outputs, left channel inside the chunk, output inside the quarter), so
// one 16-byte row of F gathers 8 single elements of A.  One block per (set, scope, repetition, chunk): the A chunk goes
// through shared memory, both global sides are coalesced 16-byte accesses.
__global__ void __launch_bounds__(256) tc_wimage_f_kernel(const uint4* __restrict__ img, uint4* __restrict__ imgf, int P, int NL) {
    extern __shared__ uint4 achunk[];
    const int CG = P / 8, HC = P / 2, QC = P / 4, NCOLS = NL * P;
    const int rows = NCOLS * CG;                       // 16-byte rows per chunk
    const int64_t base = (int64_t)blockIdx.x * rows;   // chunks are contiguous: blockIdx.x = (s, r) * NCH + chunk
    for (int i = threadIdx.x; i < rows; i += blockDim.x) achunk[i] = img[base + i];
    __syncthreads();
    const unsigned short* a16 = reinterpret_cast<const unsigned short*>(achunk);
    for (int f = threadIdx.x; f < rows; f += blockDim.x) {
        const int kk = f & 7, ng = (f >> 3) % (NCOLS / 8), rg = f / (NCOLS);
        const int rp = rg * 8 + kk;
        unsigned short e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = ng * 8 + j;
            const int quarter = n / (NL * QC), rem = n % (NL * QC);
            const int li = rem / QC, o = quarter * QC + rem % QC;
            const int na = (rp / HC) * (NL * HC) + li * HC + rp % HC;
            e[j] = a16[(((na >> 3) * CG + (o >> 3)) * 8 + (na & 7)) * 8 + (o & 7)];
        }
        uint4 v;
        v.x = e[0] | ((uint32_t)e[1] << 16); v.y = e[2] | ((uint32_t)e[3] << 16);
        v.z = e[4] | ((uint32_t)e[5] << 16); v.w = e[6] | ((uint32_t)e[7] << 16);
        imgf[base + f] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// launch plumbing
// ---------------------------------------------------------------------------------------------------------------
// One zeroed work counter per launch.  Eager launches take slots round-robin from a ring (launches in flight never come
// near its size).  Launches recorded into a CUDA graph keep their slot for the life of the graph (the memset node re-zeroes
// it on every replay), so they draw from a separate region that the eager ring never wraps onto; one ring per device.
constexpr unsigned int TC_EAGER_SLOTS = 1024, TC_GRAPH_SLOTS = 3072;
__device__ unsigned int g_tc_counters[TC_EAGER_SLOTS + TC_GRAPH_SLOTS];
unsigned int* tc_next_counter(cudaStream_t st) {
    constexpr int MAXDEV = 64;
    static std::atomic<unsigned int*> base[MAXDEV];     // forward and backward launch from different host threads
    static std::atomic<unsigned int> next_eager[MAXDEV], next_graph[MAXDEV];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAXDEV) return nullptr;
    unsigned int* b = base[dev].load();
    if (!b) {
        if (cudaGetSymbolAddress((void**)&b, g_tc_counters) != cudaSuccess) return nullptr;
        base[dev].store(b);
    }
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) return nullptr;
    unsigned int* cnt;
    if (cs == cudaStreamCaptureStatusActive) {
        const unsigned int i = next_graph[dev].fetch_add(1);
        if (i >= TC_GRAPH_SLOTS) return nullptr;          // > 3072 tensor-core launches captured in this process
        cnt = b + TC_EAGER_SLOTS + i;
    } else {
        cnt = b + (next_eager[dev].fetch_add(1) % TC_EAGER_SLOTS);
    }
    if (cudaMemsetAsync(cnt, 0, sizeof(unsigned int), st) != cudaSuccess) return nullptr;
    return cnt;
}

// cudaFuncSetAttribute is per device: remember which devices a kernel has been configured on
bool tc_configured_on_device(std::atomic<unsigned long long>& mask, int* dev_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    *dev_out = dev;
    return dev < 64 && ((mask.load() >> dev) & 1ull);
}

template <int P, int MODE>
static int launch_tsum(const TsArgs& a, cudaStream_t st) {
    constexpr int SMEM = ts_smem_bytes<P, MODE>();
    static_assert(SMEM <= 232448, "shared memory budget");
    static std::atomic<unsigned long long> configured{0};
    auto kernel = MODE == TS_FWD ? tsum_fwd_kernel<P> : tsum_dx_kernel<P>;
    int dev = 0, sms = 148;
    if (!tc_configured_on_device(configured, &dev)) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return spn_cuda_status(e);
        if (dev < 64) configured.fetch_or(1ull << dev);
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nitems = a.nsets * a.d_out * a.R * a.nsplit;
    TsArgs b = a;
    b.counter = tc_next_counter(st);
    if (!b.counter) return SPN_ERR_ARG;
    kernel<<<nitems < sms ? nitems : sms, MODE == TS_FWD ? 864 : 352, SMEM, st>>>(b);
    return spn_launch_status();
}

static int ts_padded(int c, int oc) {
    const int m = c > oc ? c : oc;
    return (m + 7) / 8 * 8;
}

template <int MODE>
static int dispatch_tsum(int P, const TsArgs& a, cudaStream_t st) {
    switch (P) {
        case 8: return launch_tsum<8, MODE>(a, st);
        case 16: return launch_tsum<16, MODE>(a, st);
        case 24: return launch_tsum<24, MODE>(a, st);
        case 32: return launch_tsum<32, MODE>(a, st);
        case 40: return launch_tsum<40, MODE>(a, st);
        case 48: return launch_tsum<48, MODE>(a, st);
        case 56: return launch_tsum<56, MODE>(a, st);
        case 64: return launch_tsum<64, MODE>(a, st);
    }
    return SPN_ERR_UNSUPPORTED;
}

// row-tile split of a work item: enough items to balance 148 persistent CTAs, at least four tiles per item
static void ts_split(TsArgs& a) {
    a.ntiles = (a.Bs + 127) / 128;
    const int base = a.nsets * a.d_out * a.R;
    int nsplit = (8 * 148 + base - 1) / base;
    if (nsplit > a.ntiles / 4) nsplit = a.ntiles / 4;
    if (nsplit < 1) nsplit = 1;
    a.nsplit = nsplit;
}

extern "C" int spn_sum_tc_supported(int c, int oc, int R) {
    (void)R;
    return (c >= 4 && oc >= 4 && c % 4 == 0 && oc % 4 == 0 && c <= 64 && oc <= 64) ? 1 : 0;
}

extern "C" int spn_sum_tc_padded_channels(int c, int oc) { return ts_padded(c, oc); }

// images A and F for nsd = weight sets x scopes, back to back (+ slack for the descriptor overrun of the last block)
extern "C" int64_t spn_sum_tc_image_bytes(int nsd, int c, int oc, int R) {
    const int64_t P = ts_padded(c, oc);
    return 2 * (int64_t)nsd * R * P * P * P * 2 + 1024;
}

extern "C" int spn_sum_tc_prepare(const float* lw, int nsd, int c, int oc, int R, float* logz, void* wimg, void* stream) {
    if (!lw || !wimg || nsd <= 0 || !spn_sum_tc_supported(c, oc, R)) return SPN_ERR_UNSUPPORTED;
    const int K = c * c, P = ts_padded(c, oc);
    cudaStream_t st = (cudaStream_t)stream;
    if (logz) {
        const int ocr = oc * R, tiles = (ocr + 63) / 64;
        tc_logz_kernel<<<nsd * tiles, 256, 0, st>>>(lw, logz, K, ocr, tiles);
        int rc = spn_launch_status();
        if (rc != SPN_OK) return rc;
    }
    const int64_t img_bytes = (int64_t)nsd * R * P * P * P * 2;
    const int RP = R | 1;
    const int KS = oc * RP + ((9 - (oc * RP) % 32) + 32) % 32;
    const size_t smem = (size_t)8 * KS * sizeof(float);
    if (c % 8 == 0 && oc % 8 == 0 && R % 4 == 0 && smem > 64 * 1024) {
        if (P != c || P != oc) {
            cudaError_t e = cudaMemsetAsync(wimg, 0, (size_t)img_bytes, st);
            if (e != cudaSuccess) return spn_cuda_status(e);
        }
        const int RPc = 17, KSc = oc * RPc + ((9 - (oc * RPc) % 32) + 32) % 32;
        dim3 grid(nsd * (K / 8), (R + 15) / 16);
        tc_wimage_rchunk_kernel<<<grid, 256, (size_t)8 * KSc * sizeof(float), st>>>(lw, logz, (uint8_t*)wimg, c, oc, R, P, ts_nl(P));
    } else if (c % 8 == 0 && oc % 8 == 0 && smem <= 200 * 1024) {
        if (P != c || P != oc) {
            cudaError_t e = cudaMemsetAsync(wimg, 0, (size_t)img_bytes, st);
            if (e != cudaSuccess) return spn_cuda_status(e);
        }
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(tc_wimage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return spn_cuda_status(e);
        }
        tc_wimage_kernel<<<nsd * (K / 8), 256, smem, st>>>(lw, logz, (uint8_t*)wimg, c, oc, R, P, ts_nl(P), RP, KS);
    } else {
        tc_wimage_generic_kernel<<<spn_blocks(img_bytes / 16, 256), 256, 0, st>>>(lw, logz, (uint4*)wimg, nsd, c, oc, R, P, ts_nl(P));
    }
    int rc = spn_launch_status();
    if (rc != SPN_OK) return rc;
    const int NL = ts_nl(P);
    tc_wimage_f_kernel<<<nsd * R * (P / NL), 256, NL * P * P * 2, st>>>((const uint4*)wimg, (uint4*)((uint8_t*)wimg + img_bytes), P, NL);
    return spn_launch_status();
}

// hand-off buffers: forward -> backward (packed eL, eR, shifts) and dX -> dW (G tiles), per (set, scope, repetition, row tile)
extern "C" int64_t spn_sum_tc_fhand_bytes(int Bs, int nsd, int c, int oc, int R) {
    const int64_t P = ts_padded(c, oc);
    return (int64_t)nsd * R * ((Bs + 127) / 128) * (2 * (P / 2) * 512 + 1024);
}
extern "C" int64_t spn_sum_tc_hand_bytes(int Bs, int nsd, int c, int oc, int R) {
    const int64_t P = ts_padded(c, oc);
    return (int64_t)nsd * R * ((Bs + 127) / 128) * (16 * (ts_kp((int)P) / 8) * 128) + 1024;
}

extern "C" int spn_sum_tc_fwd(const float* x, const void* wimg, const float* lw, const float* logz, int Bs, int nsets, int d_in,
                              int d_out, int c, int oc, int R, float* out, void* fhand, void* stream) {
    if (!x || !wimg || !lw || !out || Bs <= 0 || nsets <= 0 || d_in <= 0 || d_in > 2 * d_out) return SPN_ERR_ARG;
    if (!spn_sum_tc_supported(c, oc, R)) return SPN_ERR_UNSUPPORTED;
    const int P = ts_padded(c, oc);
    TsArgs a{};
    a.x = x; a.lw = lw; a.logz = logz; a.y = out; a.fhand_w = (uint8_t*)fhand;
    a.wimg = (const uint8_t*)wimg + (size_t)nsets * d_out * R * P * P * P * 2;   // image F
    a.Bs = Bs; a.nsets = nsets; a.Btot = Bs * nsets; a.d_in = d_in; a.d_out = d_out; a.R = R; a.c = c; a.oc = oc;
    ts_split(a);
    return dispatch_tsum<TS_FWD>(P, a, (cudaStream_t)stream);
}

extern "C" int spn_sum_tc_bwd_x(const float* x, const float* out, const float* g, const void* wimg, const void* fhand, int Bs,
                                int nsets, int d_in, int d_out, int c, int oc, int R, float* dx, void* hand, float* csum,
                                void* stream) {
    (void)x;
    if (!out || !g || !wimg || !fhand || !hand || !dx || Bs <= 0 || nsets <= 0 || d_in <= 0 || d_in > 2 * d_out) return SPN_ERR_ARG;
    if (!spn_sum_tc_supported(c, oc, R)) return SPN_ERR_UNSUPPORTED;
    const int P = ts_padded(c, oc);
    cudaStream_t st = (cudaStream_t)stream;
    TsArgs a{};
    a.wimg = (const uint8_t*)wimg;   // image A
    a.y = dx; a.hand = (const uint8_t*)hand; a.fhand_r = (const uint8_t*)fhand;
    a.Bs = Bs; a.nsets = nsets; a.Btot = Bs * nsets; a.d_in = d_in; a.d_out = d_out; a.R = R; a.c = c; a.oc = oc;
    ts_split(a);
    // G pre-pass: G tiles (read by this call's kernel and by spn_sum_tc_bwd_w) and, optionally, the column sums of g
    if (csum) {
        cudaError_t e = cudaMemsetAsync(csum, 0, (size_t)nsets * d_out * oc * R * sizeof(float), st);
        if (e != cudaSuccess) return spn_cuda_status(e);
    }
    const size_t gsmem = (size_t)16 * (ts_kp(P) / 8) * 128 + 128 * 4 + P * 4;
    tc_gprep_kernel<<<nsets * d_out * R * a.ntiles, 128, gsmem, st>>>(out, g, (const uint8_t*)fhand, (uint8_t*)hand, csum, Bs,
                                                                  Bs * nsets, d_out, oc, R, P, a.ntiles);
    int rc = spn_launch_status();
    if (rc != SPN_OK) return rc;
    return dispatch_tsum<TS_DX>(P, a, st);
}

extern "C" int spn_debug_tc_prof(unsigned long long* out_host, int reset) {
    unsigned long long v[12] = {0};
    cudaError_t e = cudaMemcpyFromSymbol(v, g_tc3_prof, sizeof(v));
    if (e != cudaSuccess) return spn_cuda_status(e);
    if (out_host) for (int i = 0; i < 12; ++i) out_host[i] = v[i];
    if (reset) {
        unsigned long long z[12] = {0};
        e = cudaMemcpyToSymbol(g_tc3_prof, z, sizeof(z));
        if (e != cudaSuccess) return spn_cuda_status(e);
    }
    return SPN_OK;
}

// debug: number of CTAs that hit the bounded-wait timeout since the library was loaded (synchronises).  A non-zero count
// also trapped the launch (tc_abort_trap), so this normally reads 0 or fails with the sticky launch error.
int tc_bwd_aborts_host(int* out);

extern "C" int spn_debug_tc_aborts(int* out_host) {
    int v = 0, w = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&v, g_tc3_aborts, sizeof(int));
    if (e != cudaSuccess) return spn_cuda_status(e);
    int rc = tc_bwd_aborts_host(&w);
    if (rc != SPN_OK) return rc;
    if (out_host) *out_host = v + w;
    return SPN_OK;
}
