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
//   * A operand: the row vectors eR (forward) / G (dX) as packed bf16, written to TENSOR MEMORY by the prologue warps
//     (tcgen05.st, lane = row) -- 20 to 32 columns per tile, no shared memory involved.
//   * B operand: the bf16 weight image of (set, s, r) in shared memory, copied by the TMA engine (cp.async.bulk).  Forward
//     reads the chunked image F (MN-major: K = r', N = (l, o)), dX the k-major image A (K-major: K = o, N = (l, r')).
//     Images up to 128 KB stay resident for all row tiles of a work item; wider ones stream chunk by chunk through a ring.
//   * accumulators: fp32 in TMEM, NL*P columns per chunk of NL left channels, NB chunk buffers in flight.
//   * epilogue: tcgen05.ld (measured 1.6 KB/clk/SM on B200, profiles/r02_pipe_bench.txt) + packed fp32 FMAs (FFMA2,
//     two per lane and instruction): forward 0.5 FMA per accumulator element, dX 2 (one per child), against
//     M*N*K/8192 clk of tensor pipe -- the forward is bound by the tensor pipe, dX by the fp32 pipe.
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
__device__ unsigned long long g_tc3_prof[24];
// Per-phase clock hooks (build with SPN_TC_PROFILE=1): every lane accumulates into registers, CTA 0 flushes once at the end.
#ifdef SPN_TC_PROFILE
#define TSP_DECL long long tsp_acc[6] = {0, 0, 0, 0, 0, 0}; long long tsp_t = 0
#define TSP_START() tsp_t = clock64()
#define TSP_LAP(slot) do { const long long tsp_n = clock64(); tsp_acc[slot] += tsp_n - tsp_t; tsp_t = tsp_n; } while (0)
#define TSP_COUNT(slot) tsp_acc[slot] += 1
#define TSP_T0() const long long tsp_t0 = clock64()
#define TSP_ADD(slot) tsp_acc[slot] += clock64() - tsp_t0
#define TSP_FLUSH(base)                                                                                          \
    do {                                                                                                         \
        if (blockIdx.x == 0 && lane == 0)                                                                        \
            for (int q = 0; q < 6; ++q) atomicAdd(&g_tc3_prof[(base) + q], (unsigned long long)tsp_acc[q]);      \
    } while (0)
#else
#define TSP_DECL do { } while (0)
#define TSP_START() do { } while (0)
#define TSP_LAP(slot) do { } while (0)
#define TSP_COUNT(slot) do { } while (0)
#define TSP_T0() do { } while (0)
#define TSP_ADD(slot) do { } while (0)
#define TSP_FLUSH(base) do { } while (0)
#endif

enum { TS_FWD = 0, TS_DX = 1 };

// Experiment switches (timing only, results are garbage): exist only in builds with -DSPN_TC_DEBUGSW (SPN_TC_DEBUG env):
//   1 no MMA issue (commits only), 2 epilogue skips loads and FMAs, 3 epilogue loads but no FMAs, 4 FMAs on stale registers (no loads)
#ifdef SPN_TC_DEBUGSW
#define TS_DBG(a) ((a).debug)
#else
#define TS_DBG(a) 0
#endif

struct TsArgs {
    unsigned int* counter;   // work counter of this launch (dynamic item scheduling), zeroed by the host
    const float* x;          // [d_in][R][Btot][c]
    const uint8_t* wimg;     // FWD: image F, DX: image A; one image of P*P*P*2 bytes per (set, scope, repetition)
    const float* lw;         // fp32 weights [nsets][d_out][c*c][oc][R] (exact rescue path, FWD)
    const float* logz;       // NULL or [nsets*d_out][oc][R]
    const float* out;        // DX: forward output [d_out][R][Btot][oc]
    const float* g;          // DX: upstream gradient, same layout
    float* y;                // FWD: out;  DX: dx [d_in][R][Btot][c]
    uint8_t* fhand_w;        // FWD, optional: hand-off written per (set, s, r, row tile): packed bf16 eL, eR + row shifts
    const uint8_t* fhand_r;  // DX: the forward's hand-off
    uint8_t* hand;           // DX, optional: G tiles in the dW kernel's B-operand layout
    int Bs, Btot, nsets, d_in, d_out, R, c, oc;
    int ntiles, nsplit, debug;
};

// ---- compile-time geometry ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int ts_kp(int P) { return (P + 15) / 16 * 16; }
// left channels per chunk: even, NL*P a multiple of 16, three chunk buffers next to the (single) A tile in 512 columns
__host__ __device__ constexpr int ts_nl(int P) {
    int best = 2, best_div = 2;
    for (int nl = 2; nl <= P; nl += 2)
        if (3 * nl * P + ts_kp(P) / 2 <= 512 && nl * P <= 256 && (nl * P) % 16 == 0) {
            best = nl;
            if (P % nl == 0) best_div = nl;
        }
    return 2 * best_div >= best ? best_div : best;   // prefer equal chunks
}
__host__ __device__ constexpr int ts_nb(int P) {
    int nb = (512 - ts_kp(P) / 2) / (ts_nl(P) * P);
    return nb > 4 ? 4 : nb;
}
__host__ __device__ constexpr bool ts_resident(int P) { return P * P * P * 2 <= 131072; }
__host__ __device__ constexpr int ts_nslot(int P) { return ts_resident(P) ? 1 : 4; }
__host__ __device__ constexpr int ts_wbytes(int P) { return ts_resident(P) ? P * P * P * 2 : ts_nslot(P) * ts_nl(P) * P * P * 2; }
constexpr int TS_ZERO = 4096;   // zeroed bytes behind the weights: target of padded K-halves / descriptor overruns
template <int P, int MODE>
__host__ __device__ constexpr int ts_smem_bytes() {
    const int vec = MODE == TS_FWD ? (2 * 128 * P * 4 /* x staging */ + 2 * P * 128 * 4 /* eL fp32 */ + (P / 2) * 512 /* packed eR */)
                                   : (2 * 2 * (P / 2) * 512 /* packed eL, eR */ + 2 * P * 128 * 4 /* partial aL */ +
                                      (P / 2) * 512 /* G tile staging */);
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

// Thread organisation (22 warps, 80 registers per thread):
//   warps 0-3   PROLOGUE: one thread per row.  FWD: per child: row max, exp, pack; the right child's packed vector goes to the
//               TMEM A tile, the left child's fp32 vector to shared memory; both are handed to the backward kernels.
//               DX: the G tile (written by the G pre-pass, staged by TMA) goes to the TMEM A tile.
//   warps 4-19  EPILOGUE: four threads per row (same lane, warps of one TMEM lane quadrant): "h" 0 / 1 = lower / upper half
//               of the P columns of every left channel's block of the accumulator, times a role bit.  tcgen05.ld has more than
//               100 clk of latency and is not double buffered here: four warps per scheduler hide it.
//               FWD: role = lower / upper quarter of the half: acc[o] += eL[l] * T[l, o] for the thread's P/4 outputs.
//               DX:  role 0: aR[r'] += eL[l] * dE[l, r'];  role 1: aL[l] += eR[r'] * dE[l, r'] (partial over the half, combined
//               through smem) -- both roles read the same half of the accumulator.
//   warp 20     MMA issuer.      warp 21     weight TMA (whole image per work item, or chunks through a ring).
// The epilogue warps have scheduling priority (higher warp ids) over the prologue.
// TMEM: NB chunk buffers (accumulators) + ONE A tile; the prologue writes the next tile's A operand when the MMAs of the
// current one have completed, while the epilogue still has up to NB chunks to consume.
// All hand-offs are mbarriers (producer arrive = release, consumer try_wait = acquire).
template <int P, int MODE>
__global__ void __launch_bounds__(704, 1) tsum_kernel(TsArgs a) {
    constexpr int CG = P / 8;                  // 8-channel groups
    constexpr int KP = ts_kp(P);               // K extent of the GEMM (padded to the MMA's K = 16)
    constexpr int KSTEPS = KP / 16;
    constexpr int ACOLS = KP / 2;              // A tile: packed bf16 pairs, one 32-bit column per two K values
    constexpr int NL = ts_nl(P);               // left channels per chunk
    constexpr int NCH = (P + NL - 1) / NL;     // chunks per row tile
    constexpr int NCOLS = NL * P;              // accumulator columns per chunk buffer
    constexpr int NB = ts_nb(P);               // chunk buffers
    constexpr int HC = P / 2;                  // columns per epilogue half
    constexpr int QC = P / 4;                  // FWD: columns per epilogue thread
    constexpr int DL = HC <= 20 ? 2 : 1;       // DX: left channels per load step (register budget)
    constexpr int NPAIR = P / 2;               // packed words per vector
    constexpr bool RESIDENT = ts_resident(P);
    constexpr int NSLOT = ts_nslot(P);
    constexpr int CHUNK_BYTES = NL * P * P * 2;
    constexpr int IMG_BYTES = P * P * P * 2;
    constexpr int W_BYTES = ts_wbytes(P);
    constexpr int A_COL0 = NB * NCOLS;
    constexpr int FH_BYTES = 2 * NPAIR * 512 + 1024;   // forward hand-off per (item, tile): eL, eR packed + shifts
    constexpr int G_BYTES = NPAIR * 512;               // G tile per (item, tile): [16 row groups][CG][8 rows][8 bf16]
    static_assert(P % 8 == 0 && P >= 8 && P <= 64 && NL % 2 == 0 && P % NL % 2 == 0, "unsupported channel count");
    static_assert(NB >= 2 && A_COL0 + ACOLS <= 512 && ACOLS % 4 == 0 && QC % 2 == 0 && NL * QC <= 40, "TMEM / register budget");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* wsm = smem;                                       // weights: resident image or ring of NSLOT chunks
    uint8_t* zero = smem + W_BYTES;                            // TS_ZERO bytes of zeros
    uint8_t* vec = zero + TS_ZERO;
    // FWD: Xs [2 children][128][c] fp32 staged by TMA, EL [2 buf][P][128] fp32, As [NPAIR][128] packed eR (thread-private staging)
    float* Xs = reinterpret_cast<float*>(vec);
    float* EL = Xs + 2 * 128 * P;
    uint32_t* As = reinterpret_cast<uint32_t*>(EL + 2 * P * 128);
    // DX: Us / Vs [2 buf][NPAIR][128] packed bf16 eL / eR (TMA from the forward's hand-off), PL [2 halves][P][128] fp32,
    //     Gs: one G tile staged by TMA
    uint32_t* Us = reinterpret_cast<uint32_t*>(vec);
    uint32_t* Vs = Us + 2 * NPAIR * 128;
    float* PL = reinterpret_cast<float*>(Vs + 2 * NPAIR * 128);
    uint8_t* Gs = reinterpret_cast<uint8_t*>(PL + 2 * P * 128);
    float* Ms = reinterpret_cast<float*>(vec + (MODE == TS_FWD ? (2 * 128 * P * 4 + 2 * P * 128 * 4 + NPAIR * 512)
                                                               : (2 * 2 * NPAIR * 512 + 2 * P * 128 * 4 + G_BYTES)));   // [2 buf][2][128]
    uint64_t* vec_full = reinterpret_cast<uint64_t*>(Ms + 2 * 2 * 128);   // [2]  FWD: prologue -> epilogue;  DX: TMA (hand-off) -> epilogue
    uint64_t* vec_empty = vec_full + 2;                                  // [2]  epilogue -> prologue
    uint64_t* acc_full = vec_empty + 2;                                  // [NB] MMA commit -> epilogue
    uint64_t* acc_empty = acc_full + 4;                                  // [NB] epilogue -> MMA
    uint64_t* w_full = acc_empty + 4;                                    // [NSLOT] TMA -> MMA
    uint64_t* w_empty = w_full + 4;                                      // [NSLOT] MMA commit -> TMA
    uint64_t* x_full = w_empty + 4;                                      // [1]  TMA (FWD: x rows, DX: G tile) -> prologue
    uint64_t* a_full = x_full + 1;                                       // [1]  prologue (A tile in TMEM) -> MMA
    uint64_t* a_free = a_full + 1;                                       // [1]  MMA commit (last chunk of a tile) -> prologue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_free + 1);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(vec_full + i, MODE == TS_FWD ? 4 : 1);
            mbar_init(vec_empty + i, 16);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(acc_full + i, 1);
            mbar_init(acc_empty + i, 16);
            mbar_init(w_full + i, 1);
            mbar_init(w_empty + i, 1);
        }
        mbar_init(x_full, 1);
        mbar_init(a_full, 4);
        mbar_init(a_free, 1);
        *abort_flag = 0;
        mbar_fence_init();
    }
    // zero region (and, for streamed weights, the ring: descriptor overruns must only ever see finite values)
    for (int i = threadIdx.x; i < (TS_ZERO + (RESIDENT ? 0 : W_BYTES)) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(RESIDENT ? zero : wsm)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 20) tmem_alloc(tmem_slot, 512);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (threadIdx.x == 0 && *tmem_slot != 0) *abort_flag = 1;   // the whole TMEM of this SM is ours: base must be 0
    constexpr uint32_t tmem_base = 0;
    __syncthreads();

    const int c = a.c, oc = a.oc;
    const int nitems = a.nsets * a.d_out * a.R * a.nsplit;
    const int tiles_per_split = (a.ntiles + a.nsplit - 1) / a.nsplit;
    uint32_t n_tile = 0, n_chunk = 0, n_wchunk = 0, items_done = 0;   // per-role ring positions (persist across items)
    TSP_DECL;

    __shared__ int s_next[2];
    for (;; ++items_done) {
        if (threadIdx.x == 0) s_next[items_done & 1] = (int)atomicAdd(a.counter, 1u);
        __syncthreads();  // the previous item is finished by every role (its last epilogue waited for its last MMA)
        const int item = __shfl_sync(0xffffffffu, s_next[items_done & 1], 0);
        if (item >= nitems) break;
        const int split = item % a.nsplit;
        const int sr = item / a.nsplit;            // (set, scope, repetition) index: also the image / hand-off index
        const int r = sr % a.R, sg = sr / a.R;     // sg = set * d_out + s
        const int set = sg / a.d_out, s = sg % a.d_out;
        const int tile_begin = split * tiles_per_split;
        const int tile_end = min(a.ntiles, tile_begin + tiles_per_split);
        const int nt = tile_end - tile_begin;
        const bool has_l = 2 * s < a.d_in, has_r = 2 * s + 1 < a.d_in;
        const int64_t row0 = (int64_t)set * a.Bs;  // first row of this weight set in the [Btot] row axis
        if (__shfl_sync(0xffffffffu, *abort_flag, 0)) break;
        const uint8_t* img = a.wimg + (size_t)sr * IMG_BYTES;

        if (warp == 21) {
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
                            const int nlc = min(NL, P - ch * NL);
                            const uint32_t bytes = (uint32_t)nlc * P * P * 2;
                            mbar_arrive_expect_tx(w_full + slot, bytes);
                            bulk_g2s(wsm + slot * CHUNK_BYTES, img + (size_t)ch * CHUNK_BYTES, bytes, w_full + slot);
                        }
                        __syncwarp();
                    }
            }
        } else if (warp == 20) {
            // =============================================== MMA issuer ===============================================
            const uint32_t wsm_addr = smem_u32(wsm), zero_addr = smem_u32(zero);
            TSP_START();
            if (RESIDENT) {
                mbar_wait_sticky(w_full, items_done & 1, abort_flag);
                fence_after_sync();
            }
            TSP_LAP(0);   // m.wait_w
            for (int t = 0; t < nt; ++t, ++n_tile) {
                mbar_wait_sticky(a_full, n_tile & 1, abort_flag);   // A tile of this row tile is in TMEM
                fence_after_sync();
                TSP_LAP(1);   // m.wait_vec
                TSP_COUNT(5);
                const uint32_t a_col = tmem_base + A_COL0;
#pragma unroll 1
                for (int ch = 0; ch < NCH; ++ch, ++n_chunk) {
                    const uint32_t buf = n_chunk % NB, use = n_chunk / NB;
                    mbar_wait_sticky(acc_empty + buf, (use & 1) ^ 1, abort_flag);     // the epilogue has read this buffer
                    TSP_LAP(2);   // m.wait_acc_empty
                    uint32_t base = wsm_addr + ch * CHUNK_BYTES;
                    uint32_t slot = 0;
                    if (!RESIDENT) {
                        slot = n_wchunk % NSLOT;
                        mbar_wait_sticky(w_full + slot, (n_wchunk / NSLOT) & 1, abort_flag);
                        base = wsm_addr + slot * CHUNK_BYTES;
                        ++n_wchunk;
                    }
                    fence_after_sync();
                    TSP_LAP(3);   // m.wait_w_chunk
                    const int nlc = min(NL, P - ch * NL);
                    const uint32_t idesc = instr_desc_bf16(128, nlc * P, 0, MODE == TS_FWD ? 1 : 0);
                    const uint32_t acc = tmem_base + buf * NCOLS;
#pragma unroll
                    for (int kk = 0; kk < KSTEPS; ++kk) {
                        uint64_t desc;
                        if (MODE == TS_FWD) {
                            // B[K = r', N = (l, o)] MN-major: N groups 128 B apart, K groups (8 r') nlc*CG*128 B apart.  The
                            // second K half of the last step does not exist when P % 16 == 8: it is read from the zero region.
                            const uint32_t kstride = (uint32_t)nlc * CG * 128;
                            const uint32_t start = base + 2 * kk * kstride;
                            const uint32_t lbo = (KP > P && kk == KSTEPS - 1) ? zero_addr - start : kstride;
                            desc = smem_desc(start, lbo, 128);
                        } else {
                            // B[K = o, N = (l, r')] K-major: N groups (8 consecutive k) CG*128 B apart, K halves 128 B apart; a
                            // padded K half (P % 16 == 8) reads the next block -- finite image values times the A tile's zeros.
                            desc = smem_desc(base + 2 * kk * 128, 128, CG * 128);
                        }
                        if (TS_DBG(a) != 1) mma_ts_uniform(acc, a_col + kk * 8, desc, idesc, kk ? 1u : 0u);
                    }
                    mma_commit_uniform(acc_full + buf);
                    if (!RESIDENT) mma_commit_uniform(w_empty + slot);
                    TSP_LAP(4);   // m.issue
                }
                mma_commit_uniform(a_free);   // the A tile may be overwritten with the next tile's operand
            }
        } else if (warp < 4) {
            // =============================================== prologue ===============================================
            const int row = warp * 32 + lane;
            const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
            const uint32_t a_addr = tmem_base + lane_base + A_COL0;
            if (MODE == TS_FWD) {
                auto stage_x = [&](int tile) {   // whole warp 0 (one elected lane issues)
                    const int rows = min(128, a.Bs - tile * 128);
                    const uint32_t bytes = (uint32_t)rows * c * 4;
                    mbar_arrive_expect_tx_elect(x_full, (has_l ? bytes : 0u) + (has_r ? bytes : 0u));
                    if (has_l)
                        bulk_g2s_elect(Xs, a.x + (((int64_t)(2 * s) * a.R + r) * a.Btot + row0 + (int64_t)tile * 128) * c, bytes, x_full);
                    if (has_r)
                        bulk_g2s_elect(Xs + 128 * P, a.x + (((int64_t)(2 * s + 1) * a.R + r) * a.Btot + row0 + (int64_t)tile * 128) * c,
                                       bytes, x_full);
                };
                // max-shifted exp of one child's row: ev = fp32 vector, returns the shift
                auto child = [&](int half, bool valid, float* ev) -> float {
                    const bool has_mine = half ? has_r : has_l;
                    const float4* rowp = reinterpret_cast<const float4*>(Xs + half * 128 * P + row * c);
                    float m4[P / 4];
#pragma unroll
                    for (int q = 0; q < P / 4; ++q) {
                        float4 t4 = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);   // padded channel
                        if (4 * q < c) t4 = (valid && has_mine) ? rowp[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                        ev[4 * q] = t4.x; ev[4 * q + 1] = t4.y; ev[4 * q + 2] = t4.z; ev[4 * q + 3] = t4.w;
                        m4[q] = fmaxf(fmaxf(t4.x, t4.y), fmaxf(t4.z, t4.w));
                    }
#pragma unroll
                    for (int st = 1; st < P / 4; st *= 2)   // tree reduction: short dependency chain
#pragma unroll
                        for (int q = 0; q + st < P / 4; q += 2 * st) m4[q] = fmaxf(m4[q], m4[q + st]);
                    float m = m4[0];
                    m = (m == -CUDART_INF_F) ? 0.f : m;
                    const float ms = m * 1.4426950408889634f;
#pragma unroll
                    for (int q = 0; q < P; ++q) ev[q] = ex2_approx(fmaf(ev[q], 1.4426950408889634f, -ms));
                    return m;
                };
                if (warp == 0 && nt > 0) stage_x(tile_begin);
                for (int t = 0; t < nt; ++t, ++n_tile) {
                    const int tile = tile_begin + t;
                    const uint32_t vb = n_tile & 1;
                    const bool valid = (int64_t)tile * 128 + row < a.Bs;
                    uint8_t* fh = a.fhand_w ? a.fhand_w + ((size_t)sr * a.ntiles + tile) * FH_BYTES : nullptr;
                    TSP_START();
                    mbar_wait_sticky(x_full, n_tile & 1, abort_flag);
                    TSP_LAP(0);   // p.wait_x
                    float ev[P];
                    // right child: packed bf16 -> thread-private staging (goes to TMEM when the A tile is free) + hand-off
                    const float mR = child(1, valid, ev);
                    {
                        uint32_t* as = As + row;
                        uint32_t* dw = fh ? reinterpret_cast<uint32_t*>(fh + NPAIR * 512) + row : nullptr;
#pragma unroll
                        for (int q = 0; q < NPAIR; ++q) {
                            const uint32_t w = pack_bf16(ev[2 * q], ev[2 * q + 1]);
                            as[q * 128] = w;
                            if (fh) dw[q * 128] = w;   // coalesced: lanes = consecutive rows of one packed word index
                        }
                    }
                    // left child: fp32 vector for the epilogue + packed hand-off
                    const float mL = child(0, valid, ev);
                    TSP_LAP(1);   // p.compute
                    // every prologue thread has taken its rows out of the staging area: the TMA engine may refill it
                    ts_named_bar(1, 128);
                    if (warp == 0 && t + 1 < nt) stage_x(tile + 1);
                    TSP_LAP(2);   // p.bar
                    if (fh) {
                        uint32_t* dw = reinterpret_cast<uint32_t*>(fh) + row;
#pragma unroll
                        for (int q = 0; q < NPAIR; ++q) dw[q * 128] = pack_bf16(ev[2 * q], ev[2 * q + 1]);
                        float* ms = reinterpret_cast<float*>(fh + 2 * NPAIR * 512);
                        ms[row] = mL;
                        ms[128 + row] = mR;
                    }
                    // buffer set vb: the epilogue (eL, shifts) of tile n - 2 is done with it
                    mbar_wait_sticky(vec_empty + vb, ((n_tile >> 1) & 1) ^ 1, abort_flag);
                    TSP_LAP(3);   // p.wait_vec_empty
                    Ms[vb * 256 + row] = mL;
                    Ms[vb * 256 + 128 + row] = mR;
                    {
                        float* el = EL + vb * P * 128 + row;
#pragma unroll
                        for (int q = 0; q < P; ++q) el[q * 128] = ev[q];
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(vec_full + vb);
                    // the A tile: the MMAs of the previous tile must have completed
                    mbar_wait_sticky(a_free, (n_tile & 1) ^ 1, abort_flag);
                    fence_after_sync();
                    {
                        uint32_t at[ACOLS];
#pragma unroll
                        for (int q = 0; q < ACOLS; ++q) at[q] = q < NPAIR ? As[q * 128 + row] : 0u;
                        tmem_store_cols<ACOLS>(a_addr, at);
                    }
                    tmem_wait_st();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_full);
                    TSP_LAP(4);   // p.write
                    TSP_COUNT(5);
                }
            } else {
                auto stage_g = [&](int tile) {   // whole warp 0: the G tile written by the G pre-pass
                    mbar_arrive_expect_tx_elect(x_full, G_BYTES);
                    bulk_g2s_elect(Gs, a.hand + ((size_t)sr * a.ntiles + tile) * G_BYTES, G_BYTES, x_full);
                };
                if (warp == 0 && nt > 0) stage_g(tile_begin);
                for (int t = 0; t < nt; ++t, ++n_tile) {
                    const int tile = tile_begin + t;
                    const uint32_t vb = n_tile & 1;
                    TSP_START();
                    if (warp == 0) {   // eL, eR (packed bf16) and the row shifts of this tile, straight from the forward's hand-off
                        mbar_wait_sticky(vec_empty + vb, ((n_tile >> 1) & 1) ^ 1, abort_flag);
                        const uint8_t* src = a.fhand_r + ((size_t)sr * a.ntiles + tile) * FH_BYTES;
                        mbar_arrive_expect_tx_elect(vec_full + vb, FH_BYTES);
                        bulk_g2s_elect(Us + vb * NPAIR * 128, src, NPAIR * 512, vec_full + vb);
                        bulk_g2s_elect(Vs + vb * NPAIR * 128, src + NPAIR * 512, NPAIR * 512, vec_full + vb);
                        bulk_g2s_elect(Ms + vb * 256, src + 2 * NPAIR * 512, 1024, vec_full + vb);
                    }
                    TSP_LAP(3);   // p.wait_vec_empty
                    mbar_wait_sticky(x_full, n_tile & 1, abort_flag);
                    TSP_LAP(0);   // p.wait_x (G tile)
                    uint32_t at[ACOLS];
                    {
                        const uint4* gs = reinterpret_cast<const uint4*>(Gs) + (row >> 3) * (CG * 8) + (row & 7);
#pragma unroll
                        for (int og = 0; og < ACOLS / 4; ++og) {
                            const uint4 v4 = og < CG ? gs[og * 8] : make_uint4(0u, 0u, 0u, 0u);
                            at[4 * og] = v4.x; at[4 * og + 1] = v4.y; at[4 * og + 2] = v4.z; at[4 * og + 3] = v4.w;
                        }
                    }
                    ts_named_bar(1, 128);   // every row has been taken out of the staging area
                    if (warp == 0 && t + 1 < nt) stage_g(tile + 1);
                    TSP_LAP(1);   // p.compute
                    mbar_wait_sticky(a_free, (n_tile & 1) ^ 1, abort_flag);
                    fence_after_sync();
                    TSP_LAP(2);   // p.bar (wait for the A tile)
                    tmem_store_cols<ACOLS>(a_addr, at);
                    tmem_wait_st();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_full);
                    TSP_LAP(4);   // p.write
                    TSP_COUNT(5);
                }
            }
        } else {
            // =============================================== epilogue ===============================================
            const int ew = warp - 4;
            const int h = (ew >> 2) & 1, role = ew >> 3;
            const int row = (warp & 3) * 32 + lane;
            const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
            for (int t = 0; t < nt; ++t, ++n_tile) {
                const int tile = tile_begin + t;
                const uint32_t vb = n_tile & 1;
                const int64_t b = (int64_t)tile * 128 + row;
                const bool valid = b < a.Bs;
                TSP_START();
                mbar_wait_sticky(vec_full + vb, (n_tile >> 1) & 1, abort_flag);   // eL / eR / shifts of this tile are in smem
                TSP_LAP(0);   // e.wait_vec
                TSP_COUNT(5);
                if (MODE == TS_FWD) {
                    // this thread's outputs: o in [col0, col0 + QC)
                    const int col0 = h * HC + role * QC;
                    uint64_t acc2[QC / 2];
#pragma unroll
                    for (int j = 0; j < QC / 2; ++j) acc2[j] = 0ull;
#pragma unroll 1
                    for (int ch = 0; ch < NCH; ++ch, ++n_chunk) {
                        const uint32_t buf = n_chunk % NB, use = n_chunk / NB;
                        const int nlc = min(NL, P - ch * NL);
                        {
                            TSP_T0();
                            mbar_wait_sticky(acc_full + buf, use & 1, abort_flag);
                            fence_after_sync();
                            TSP_ADD(1);   // e.wait_acc (also contained in e.chunks)
                        }
                        const uint32_t cbase = tmem_base + lane_base + buf * NCOLS + col0;
                        uint32_t d[NL * QC];
                        float el[NL];
#pragma unroll
                        for (int li = 0; li < NL; ++li)
                            if (li < nlc && TS_DBG(a) != 2 && TS_DBG(a) != 4) tmem_load_cols<QC>(cbase + li * P, d + li * QC);
#pragma unroll
                        for (int li = 0; li < NL; ++li) el[li] = li < nlc ? EL[vb * P * 128 + (ch * NL + li) * 128 + row] : 0.f;
                        {
                            TSP_T0();
                            tmem_wait_ld();
                            TSP_ADD(4);   // e.wait_ld (also contained in e.chunks)
                        }
                        fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty + buf);   // the accumulator may be overwritten
                        if (TS_DBG(a) != 2 && TS_DBG(a) != 3) {
#pragma unroll
                            for (int li = 0; li < NL; ++li) {
                                if (li < nlc) {
                                    const uint64_t e2 = f2pack(el[li], el[li]);
#pragma unroll
                                    for (int j = 0; j < QC / 2; ++j)
                                        acc2[j] = fma2(e2, u2pack(d[li * QC + 2 * j], d[li * QC + 2 * j + 1]), acc2[j]);
                                }
                            }
                        }
                    }
                    TSP_LAP(2);   // e.chunks
                    const float* my_m = Ms + vb * 256 + row;
                    const float shift = my_m[0] + my_m[128];
                    float* op = a.y + (((int64_t)s * a.R + r) * a.Btot + row0 + (valid ? b : 0)) * oc + col0;
                    bool bad = false;
#pragma unroll
                    for (int j = 0; j < QC / 2; ++j) {
                        float s0, s1;
                        f2unpack(acc2[j], s0, s1);
                        const bool on = valid && col0 + 2 * j < oc;
                        bad |= on && !(fminf(s0, s1) >= 1e-30f);
                        if (on)
                            *reinterpret_cast<float2*>(op + 2 * j) = make_float2(fmaf(lg2_approx(s0), 0.6931471805599453f, shift),
                                                                                 fmaf(lg2_approx(s1), 0.6931471805599453f, shift));
                    }
                    if (bad && !TS_DBG(a)) ts_exact_row(a, sg, s, r, row0 + b, col0, QC, op);
                } else {
                    // both roles walk over this half's HC columns of every left channel
                    uint64_t acc2[HC / 2];   // role 0: aR pairs;  role 1: this half's eR pairs as fp32
#pragma unroll
                    for (int j = 0; j < HC / 2; ++j) {
                        if (role == 0) {
                            acc2[j] = 0ull;
                        } else {
                            const uint32_t w = Vs[vb * NPAIR * 128 + (h * (HC / 2) + j) * 128 + row];
                            acc2[j] = f2pack(bf16_lo(w), bf16_hi(w));
                        }
                    }
#pragma unroll 1
                    for (int ch = 0; ch < NCH; ++ch, ++n_chunk) {
                        const uint32_t buf = n_chunk % NB, use = n_chunk / NB;
                        const int nlc = min(NL, P - ch * NL);
                        {
                            TSP_T0();
                            mbar_wait_sticky(acc_full + buf, use & 1, abort_flag);
                            fence_after_sync();
                            TSP_ADD(1);
                        }
                        const uint32_t cbase = tmem_base + lane_base + buf * NCOLS + h * HC;
#pragma unroll 1
                        for (int li = 0; li < nlc; li += DL) {
                            const int l = ch * NL + li;
                            uint32_t d[DL * HC];
#pragma unroll
                            for (int u = 0; u < DL; ++u)
                                if (TS_DBG(a) != 2 && TS_DBG(a) != 4) tmem_load_cols<HC>(cbase + (li + u) * P, d + u * HC);
                            const uint32_t wl = Us[vb * NPAIR * 128 + (l >> 1) * 128 + row];   // eL[l & ~1], eL[l | 1]
                            {
                                TSP_T0();
                                tmem_wait_ld();
                                TSP_ADD(4);
                            }
                            if (li + DL >= nlc) {   // last load of this chunk has completed
                                fence_before_sync();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(acc_empty + buf);
                            }
                            if (TS_DBG(a) == 2 || TS_DBG(a) == 3) continue;
#pragma unroll
                            for (int u = 0; u < DL; ++u) {
                                if (role == 0) {
                                    const float eu = ((l + u) & 1) ? bf16_hi(wl) : bf16_lo(wl);
                                    const uint64_t e2 = f2pack(eu, eu);
#pragma unroll
                                    for (int j = 0; j < HC / 2; ++j) acc2[j] = fma2(e2, u2pack(d[u * HC + 2 * j], d[u * HC + 2 * j + 1]), acc2[j]);
                                } else {
                                    uint64_t s0 = 0ull, s1 = 0ull;
#pragma unroll
                                    for (int j = 0; j < HC / 2; ++j) {
                                        const uint64_t d2 = u2pack(d[u * HC + 2 * j], d[u * HC + 2 * j + 1]);
                                        if (j & 1) s1 = fma2(acc2[j], d2, s1); else s0 = fma2(acc2[j], d2, s0);
                                    }
                                    float p0, p1, p2, p3;
                                    f2unpack(s0, p0, p1);
                                    f2unpack(s1, p2, p3);
                                    PL[(h * P + l + u) * 128 + row] = (p0 + p1) + (p2 + p3);
                                }
                            }
                        }
                    }
                    TSP_LAP(2);   // e.chunks
                    if (role == 0) {
                        // dxR[r'] = eR[r'] * aR[r'] for this half's right channels
                        if (valid && has_r) {
                            float* dp = a.y + (((int64_t)(2 * s + 1) * a.R + r) * a.Btot + row0 + b) * c + h * HC;
#pragma unroll
                            for (int j4 = 0; j4 < HC / 4; ++j4) {
                                const uint32_t w0 = Vs[vb * NPAIR * 128 + (h * (HC / 2) + 2 * j4) * 128 + row];
                                const uint32_t w1 = Vs[vb * NPAIR * 128 + (h * (HC / 2) + 2 * j4 + 1) * 128 + row];
                                float v0, v1, v2, v3;
                                f2unpack(acc2[2 * j4], v0, v1);
                                f2unpack(acc2[2 * j4 + 1], v2, v3);
                                if (h * HC + 4 * j4 < c)
                                    *reinterpret_cast<float4*>(dp + 4 * j4) = make_float4(bf16_lo(w0) * v0, bf16_hi(w0) * v1, bf16_lo(w1) * v2, bf16_hi(w1) * v3);
                            }
                        }
                    } else {
                        ts_named_bar(2, 256);   // both halves' partial aL of this tile are in shared memory (role-1 warps only)
                        if (valid && has_l) {
                            float* dp = a.y + (((int64_t)(2 * s) * a.R + r) * a.Btot + row0 + b) * c + h * HC;
#pragma unroll
                            for (int j4 = 0; j4 < HC / 4; ++j4) {
                                float res[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const int l = h * HC + 4 * j4 + u;
                                    const uint32_t wl = Us[vb * NPAIR * 128 + (l >> 1) * 128 + row];
                                    const float el = (l & 1) ? bf16_hi(wl) : bf16_lo(wl);
                                    res[u] = el * (PL[l * 128 + row] + PL[(P + l) * 128 + row]);
                                }
                                if (h * HC + 4 * j4 < c) *reinterpret_cast<float4*>(dp + 4 * j4) = make_float4(res[0], res[1], res[2], res[3]);
                            }
                        }
                        ts_named_bar(2, 256);   // the partial sums have been read: the next tile may overwrite them
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(vec_empty + vb);   // eL / eR / shifts of this buffer set are no longer read
                TSP_LAP(3);   // e.finalize
            }
        }
    }
    // teardown
    if (warp == 4) TSP_FLUSH(0);
    if (warp == 0) TSP_FLUSH(6);
    if (warp == 20) TSP_FLUSH(12);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 20) tmem_dealloc(tmem_base, 512);
    if (threadIdx.x == 0) tc_abort_trap(abort_flag, &g_tc3_aborts);
}

// G pre-pass of the backward (one block per (set, scope, repetition, 128-row tile)): G = g * exp(mL + mR - out) = g / (linear
// -domain sum) as packed bf16 in the core-matrix layout [row / 8][o / 8][row % 8][8 o] that both the dX kernel (A operand,
// via TMA + tcgen05.st) and the dW kernel (B operand) consume; out = -inf (impossible row): no gradient; the exponent is
// capped so that rows whose sum underflowed in the forward (rescue path) stay finite.  Also accumulates the column sums
// csum[sg][o][r] = sum_b g[b, o] over the rows with out > -inf, which the fused softmax backward of dW needs (the rows of a
// tile are contiguous, so the loads are coalesced 16-byte accesses; the shifts come from the forward's hand-off).
__global__ void __launch_bounds__(128) tc_gprep_kernel(const float* __restrict__ out, const float* __restrict__ g,
                                                      const uint8_t* __restrict__ fhand, uint8_t* __restrict__ hand,
                                                      float* __restrict__ csum, int Bs, int Btot, int d_out, int oc, int R,
                                                      int P, int ntiles) {
    extern __shared__ __align__(16) uint8_t gsm[];
    const int NPAIR = P / 2, CG = P / 8;
    uint4* Gt = reinterpret_cast<uint4*>(gsm);                          // [16][CG][8] 16-byte rows
    float* sh = reinterpret_cast<float*>(gsm + NPAIR * 512);            // [128] mL + mR (in log2 units)
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
    for (int i = threadIdx.x; i < 16 * CG * 8; i += 128) Gt[i] = make_uint4(0u, 0u, 0u, 0u);
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
        Gt2[(((row >> 3) * CG + (o >> 3)) * 8 + (row & 7)) * 2 + ((o >> 2) & 1)] = make_uint2(tc::pack_bf16(G[0], G[1]), tc::pack_bf16(G[2], G[3]));
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(hand + ((size_t)sr * ntiles + tile) * (size_t)(NPAIR * 512));
    for (int i = threadIdx.x; i < 16 * CG * 8; i += 128) dst[i] = Gt[i];
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

// Image A: fp32 (log-)weights [S][K = c*c][oc][R] (S = weight sets x scopes) -> bf16 exp() core-matrix image
// [S][R][kg = l*CG + rg][og][8 (k = r' % 8)][8 (o)], CG = P / 8 groups per padded channel axis.  Fast variant for c and oc
// multiples of 8: one block per (scope, group of 8 consecutive k): the [8][oc][R] slab is read with coalesced float4 loads
// into a padded shared-memory tile [8][oc][R|1] (+ pad so that the k-stride is 9 mod 32 banks); then every thread assembles
// one 16-byte image row (8 outputs of one k), consecutive threads taking consecutive rows of one repetition so that both the
// shared-memory reads (conflict free) and the global stores (contiguous per repetition) are efficient.  Padding blocks
// (P > c or P > oc) are zero-filled by a memset before the launch.
__global__ void __launch_bounds__(256) tc_wimage_kernel(const float* __restrict__ lw, const float* __restrict__ logz,
                                                       uint8_t* __restrict__ wimg, int c, int OC, int R, int P, int RP,
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
    const int64_t kg = (int64_t)l * CG + rg;
    // work item = (r, og, kk) with kk fastest: chunk index inside (s, r) is (kg * CG + og) * 8 + kk
    for (int i = threadIdx.x; i < NOG * 8 * R; i += blockDim.x) {
        const int kk = i & 7, og = (i >> 3) % NOG, r = i / (8 * NOG);
        const float* t = tile + kk * KS + og * 8 * RP + r;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(t[0], t[RP]), p1 = __floats2bfloat162_rn(t[2 * RP], t[3 * RP]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(t[4 * RP], t[5 * RP]), p3 = __floats2bfloat162_rn(t[6 * RP], t[7 * RP]);
        uint4 v;
        v.x = *reinterpret_cast<uint32_t*>(&p0); v.y = *reinterpret_cast<uint32_t*>(&p1);
        v.z = *reinterpret_cast<uint32_t*>(&p2); v.w = *reinterpret_cast<uint32_t*>(&p3);
        dst[(((int64_t)s * R + r) * ((int64_t)P * CG) + kg) * (CG * 8) + og * 8 + kk] = v;
    }
}

// Image A for channel counts that are multiples of 4 only: one thread per 16-byte image row (strided reads; such layers are small).
__global__ void __launch_bounds__(256) tc_wimage_generic_kernel(const float* __restrict__ lw, const float* __restrict__ logz,
                                                               uint4* __restrict__ wimg, int S, int c, int OC, int R, int P) {
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
    wimg[i] = q;
}

// Image F (forward): the same 128-byte blocks, chunk-major: [chunk][rg][l in chunk][og][8][8] so that within a chunk of NL
// left channels the (l, o) pairs form ONE uniformly strided N axis of the GEMM.  One thread per 16-byte row of the destination.
__global__ void __launch_bounds__(256) tc_wimage_f_kernel(const uint4* __restrict__ img, uint4* __restrict__ imgf, int P, int NL,
                                                         int64_t nrows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const int CG = P / 8;
    const int per_sr = P * CG * CG * 8;      // 16-byte rows per (s, r)
    const int64_t srr = i / per_sr;
    const int e = (int)(i % per_sr);
    const int per_l = CG * CG * 8;           // rows per left channel
    const int ch = e / (NL * per_l);
    const int ec = e - ch * NL * per_l;
    const int nlc = min(NL, P - ch * NL);
    // inside the chunk: [rg][li][og][kk]
    const int kk = ec % 8, og = (ec / 8) % CG, li = (ec / (8 * CG)) % nlc, rg = ec / (8 * CG * nlc);
    const int l = ch * NL + li;
    imgf[i] = img[srr * per_sr + (((int64_t)l * CG + rg) * CG + og) * 8 + kk];
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
    int dev = 0, sms = 148;
    if (!tc_configured_on_device(configured, &dev)) {
        cudaError_t e = cudaFuncSetAttribute(tsum_kernel<P, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return spn_cuda_status(e);
        if (dev < 64) configured.fetch_or(1ull << dev);
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nitems = a.nsets * a.d_out * a.R * a.nsplit;
    TsArgs b = a;
    b.counter = tc_next_counter(st);
    if (!b.counter) return SPN_ERR_ARG;
    tsum_kernel<P, MODE><<<nitems < sms ? nitems : sms, 704, SMEM, st>>>(b);
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
#ifdef SPN_TC_DEBUGSW
    const char* dbg = getenv("SPN_TC_DEBUG");
    a.debug = dbg ? atoi(dbg) : 0;
#endif
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
    if (c % 8 == 0 && oc % 8 == 0 && smem <= 200 * 1024) {
        if (P != c || P != oc) {
            cudaError_t e = cudaMemsetAsync(wimg, 0, (size_t)img_bytes, st);
            if (e != cudaSuccess) return spn_cuda_status(e);
        }
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(tc_wimage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return spn_cuda_status(e);
        }
        tc_wimage_kernel<<<nsd * (K / 8), 256, smem, st>>>(lw, logz, (uint8_t*)wimg, c, oc, R, P, RP, KS);
    } else {
        tc_wimage_generic_kernel<<<spn_blocks(img_bytes / 16, 256), 256, 0, st>>>(lw, logz, (uint4*)wimg, nsd, c, oc, R, P);
    }
    int rc = spn_launch_status();
    if (rc != SPN_OK) return rc;
    const int64_t nrows = img_bytes / 16;
    tc_wimage_f_kernel<<<spn_blocks(nrows, 256), 256, 0, st>>>((const uint4*)wimg, (uint4*)((uint8_t*)wimg + img_bytes), P, ts_nl(P), nrows);
    return spn_launch_status();
}

// hand-off buffers: forward -> backward (packed eL, eR, shifts) and dX -> dW (G tiles), per (set, scope, repetition, row tile)
extern "C" int64_t spn_sum_tc_fhand_bytes(int Bs, int nsd, int c, int oc, int R) {
    const int64_t P = ts_padded(c, oc);
    return (int64_t)nsd * R * ((Bs + 127) / 128) * (2 * (P / 2) * 512 + 1024);
}
extern "C" int64_t spn_sum_tc_hand_bytes(int Bs, int nsd, int c, int oc, int R) {
    const int64_t P = ts_padded(c, oc);
    return (int64_t)nsd * R * ((Bs + 127) / 128) * (P / 2) * 512 + 1024;
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
    a.out = out; a.g = g; a.y = dx; a.hand = (uint8_t*)hand; a.fhand_r = (const uint8_t*)fhand;
    a.Bs = Bs; a.nsets = nsets; a.Btot = Bs * nsets; a.d_in = d_in; a.d_out = d_out; a.R = R; a.c = c; a.oc = oc;
    ts_split(a);
    // G pre-pass: G tiles (read by this call's kernel and by spn_sum_tc_bwd_w) and, optionally, the column sums of g
    if (csum) {
        cudaError_t e = cudaMemsetAsync(csum, 0, (size_t)nsets * d_out * oc * R * sizeof(float), st);
        if (e != cudaSuccess) return spn_cuda_status(e);
    }
    const size_t gsmem = (size_t)(P / 2) * 512 + 128 * 4 + P * 4;
    tc_gprep_kernel<<<nsets * d_out * R * a.ntiles, 128, gsmem, st>>>(out, g, (const uint8_t*)fhand, (uint8_t*)hand, csum, Bs,
                                                                  Bs * nsets, d_out, oc, R, P, a.ntiles);
    int rc = spn_launch_status();
    if (rc != SPN_OK) return rc;
    return dispatch_tsum<TS_DX>(P, a, st);
}

extern "C" int spn_debug_tc_prof(unsigned long long* out_host, int reset) {
    unsigned long long v[24] = {0};
    cudaError_t e = cudaMemcpyFromSymbol(v, g_tc3_prof, sizeof(v));
    if (e != cudaSuccess) return spn_cuda_status(e);
    if (out_host) for (int i = 0; i < 24; ++i) out_host[i] = v[i];
    if (reset) {
        unsigned long long z[24] = {0};
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
