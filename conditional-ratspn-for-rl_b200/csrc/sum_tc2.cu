// Shared-weight (RatSpn) CrossProduct+Sum on the 5th-generation tensor cores (tcgen05 + TMEM): forward and dX.
//
// For one sum scope s and repetition r the layer is a dense contraction over K = c*c in-channels:
//   forward   out[b,o]  = mL + mR + log sum_{l,r'} eL[b,l] eR[b,r'] W[(l,r'),o]          eL = exp(xL - mL), eR = exp(xR - mR)
//   dX right  dxR[b,r'] = eR[b,r'] * sum_{l,o}  eL[b,l] G[b,o] W[(l,r'),o]                G  = g * exp(mL + mR - out)
//   dX left   dxL[b,l]  = eL[b,l]  * sum_{r',o} eR[b,r'] G[b,o] W[(l,r'),o]
// All three are the SAME shape of GEMM: [128 rows x K = nu*nv] x [K x N <= 48], whose A operand is an outer product
// A[b, u*C + v] = U[b,u] * V[b,v] of two per-row vectors that never exists in memory:
//   mode FWD: U = eL, V = eR, N = o;   mode DXR: U = eL, V = G, N = r';   mode DXL: U = eR, V = G, N = l.
// Mapping to the hardware:
//   * A operand: produced by CUDA cores straight into TENSOR MEMORY (tcgen05.st), one bf16x2 multiply per two K-elements
//     (an smem-staged A would need 170 B/clk/SM of shared-memory writes, more than an SM has); stages of 2*C K-values,
//     NST-deep ring per warpgroup; the TMEM store of stage i is overlapped with the multiplies of stage i + 1.
//   * B operand: the bf16 weight image [K/8][C/8][8][8] of W(s, r) (pitch C, no padding) copied once per work item by the
//     TMA engine (cp.async.bulk).  FWD reads it as an MN-major operand (K = (l,r'), N = o).  DXR reads THE SAME bytes as
//     a K-major operand with N = r' and K = (l, o): every 8x8 core matrix of the image is one (l, 8 r', 8 o) block, so
//     only the descriptor strides change.  DXL needs the pair-transposed image W2[(r',l),o] (a 16-byte-chunk permutation
//     written by spn_sum_tc_prepare).  N = 48 > C = 40: the sixth N-group aliases the following block; the accumulator
//     columns it produces are never read.
//   * accumulator: fp32 in TMEM (48 columns), read back with tcgen05.ld; branch-free epilogue.
//   * roles: two producer/epilogue warpgroups (one row tile each, ping-pong) + one MMA-issuing warp per warpgroup.
// Activations use the internal layout [scope][repetition][batch][channel] (one thread <-> one contiguous row).
// Outputs whose linear-domain sum underflows are recomputed exactly in fp32, so results keep th.logsumexp semantics.
#include <stdlib.h>

#include <atomic>
#include "tc_common.cuh"

using namespace tc;

__device__ int g_tc2_aborts = 0;
__device__ unsigned long long g_tc2_prof[24];
// Per-phase clock hooks (build with SPN_TC_PROFILE=1).  Every lane reads the clock and accumulates into registers; the
// totals are flushed once at the end of the kernel, so the hooks add no branches to the pipeline loops (a divergent
// `if (lane == 0)` inside the MMA issue loop costs the uniform-register code path).
#ifdef SPN_TC_PROFILE
#define TC2_DECL long long tc2_acc[6] = {0, 0, 0, 0, 0, 0}
#define TC2_T(var) const long long var = clock64()
#define TC2_ADD(slot, t0, t1) tc2_acc[(slot) % 6] += (t1) - (t0)
#define TC2_COUNT(slot) tc2_acc[(slot) % 6] += 1
#define TC2_FLUSH(base)                                                                                          \
    do {                                                                                                         \
        if (blockIdx.x == 0 && lane == 0)                                                                        \
            for (int q = 0; q < 6; ++q) atomicAdd(&g_tc2_prof[(base) + q], (unsigned long long)tc2_acc[q]);      \
    } while (0)
#else
#define TC2_DECL do { } while (0)
#define TC2_T(var) do { } while (0)
#define TC2_ADD(slot, t0, t1) do { } while (0)
#define TC2_COUNT(slot) do { } while (0)
#define TC2_FLUSH(base) do { } while (0)
#endif
// The experiment switches (SPN_TC_DEBUG: skip the MMAs / the TMEM stores / the multiplies ..., SPN_TC_N: MMA width) exist
// only in builds with -DSPN_TC_DEBUGSW.  In the product build they are the constant 0: a run-time switch on them became a jump
// table (LDC + BRX) on the MMA warp's per-stage critical path, ~150 clk per stage while the tensor pipe drained.
#ifdef SPN_TC_DEBUGSW
#define OPG_DEBUG(a) ((a).debug)
#define OPG_DBGN(a) ((a).dbg_n)
#else
#define OPG_DEBUG(a) 0
#define OPG_DBGN(a) 0
#endif

enum { OPG_FWD = 0, OPG_DXR = 1, OPG_DXL = 2 };

struct OpgArgs {
    unsigned int* counter;   // work counter of this launch (dynamic item scheduling), zeroed by the host
    const float* x;        // [d_in][R][B][C]
    const uint8_t* wimg;   // [2][d_out][R][K*C] bf16: image, then pair-transposed image
    const float* lw;       // fp32 weights [1][d_out][K][OC][R] (exact rescue path, FWD)
    const float* logz;     // NULL or [d_out][OC][R]
    const float* out;      // [d_out][R][B][OC]  (DX: forward output)
    const float* g;        // [d_out][R][B][OC]  (DX: upstream gradient)
    float* y;              // FWD: out;  DX: dx [d_in][R][B][C]
    uint8_t* hand;         // optional hand-off written by this launch, per (s, r, row tile):
                           //   FWD: packed eL, eR ([pair][128 rows] words) + row shifts mL, mR  -> read by the dX / dW kernels
                           //   DXR: G tile in the dW kernel's B-operand layout                    -> read by the dW kernel
    const uint8_t* fhand;  // DX, optional: the forward's hand-off (then no x rows are loaded and nothing is exponentiated)
    const uint8_t* ghand;  // DXL, optional: the G tiles the DXR launch of the same call handed off (then out / g are not read)
    int B, d_in, d_out, R;
    int ntiles, nsplit, debug, dbg_n;
};

// exact fp32 log-sum-exp for one output (rare rescue path)
template <int C, int NOUT>
__device__ __noinline__ void opg_exact_row(const OpgArgs& a, int s, int r, int64_t b, int o_begin, float* op) {
    const bool has_l = 2 * s < a.d_in, has_r = 2 * s + 1 < a.d_in;
    const float* xl = a.x + (((int64_t)(2 * s) * a.R + r) * a.B + b) * C;
    const float* xr = a.x + (((int64_t)(2 * s + 1) * a.R + r) * a.B + b) * C;
    for (int oo = 0; oo < NOUT; ++oo) {
        const int o = o_begin + oo;
        const float* lwp = a.lw + ((int64_t)s * C * C * C + o) * a.R + r;
        const int64_t sw_i = (int64_t)C * a.R;
        float m = -CUDART_INF_F;
        for (int l = 0; l < C; ++l)
            for (int rp = 0; rp < C; ++rp)
                m = fmaxf(m, (has_l ? xl[l] : 0.f) + (has_r ? xr[rp] : 0.f) + lwp[(int64_t)(l * C + rp) * sw_i]);
        float res = m;
        if (m != -CUDART_INF_F) {
            float acc = 0.f;
            for (int l = 0; l < C; ++l)
                for (int rp = 0; rp < C; ++rp)
                    acc += expf((has_l ? xl[l] : 0.f) + (has_r ? xr[rp] : 0.f) + lwp[(int64_t)(l * C + rp) * sw_i] - m);
            res = m + logf(acc);
            if (a.logz) res -= a.logz[((int64_t)s * C + o) * a.R + r];
        }
        op[oo] = res;
    }
}

// ---- MMA issue: one PTX block per stage (one elect, descriptor arithmetic on block-local registers) --------------------
// Descriptor of MMA kk of a stage = base descriptor of the stage + a compile-time constant:
//   FWD  (B MN-major, K groups NVG*128 B apart): start advances by two K groups per MMA.
//   DX   (B K-major over the flattened (u, o-group) index g = g0 + 2kk, g0 a multiple of 4*NVG): the start block is
//        ((g / NVG) * NVG*NVG + g % NVG) and the leading-byte-offset field is the distance to the block of g + 1.
template <int C, int MODE>
__host__ __device__ constexpr long long opg_desc_delta(int kk) {
    constexpr int NVG = C / 8;
    if (MODE == OPG_FWD) return (long long)kk * 2 * NVG * 8;  // units of 16 bytes
    const int ga = 2 * kk, gb = ga + 1;
    const int ua = ga / NVG, oa = ga % NVG, ub = gb / NVG, ob = gb % NVG;
    const long long start = (long long)(ua * NVG * NVG + oa) * 8;
    const long long lbo = (long long)((ub - ua) * NVG * NVG + (ob - oa)) * 8;
    return start + (lbo << 16);
}

#define OPG_MMA_LINE(IMM, AOFF)                                                                                   \
    "add.s64 d, %1, " IMM ";\n\tadd.s32 ta, %0, " #AOFF ";\n\t"                                                     \
    "@q tcgen05.mma.cta_group::1.kind::f16 [%5], [ta], d, %2, {%3, %3, %3, %3}, p;\n\tsetp.ne.b32 p, 1, 0;\n\t"

// Issues the C/4 MMAs of one stage: D[acc] (+)= A[tmem abase + 8 kk] * B[dbase + delta(kk)].
template <int C, int MODE>
__device__ __forceinline__ void opg_issue_stage(uint32_t acc, uint32_t abase, uint64_t dbase, uint32_t idesc, uint32_t accumulate_first) {
    constexpr int MMAS = C / 4;
    static_assert(MMAS == 10 || MMAS == 8 || MMAS == 4, "unsupported channel count");
#define OPG_D(i) "n"(opg_desc_delta<C, MODE>(i))
    if constexpr (MMAS == 10) {
        asm volatile(
            "{\n\t.reg .pred q, p;\n\t.reg .b64 d;\n\t.reg .b32 ta;\n\t"
            "elect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            OPG_MMA_LINE("%6", 0) OPG_MMA_LINE("%7", 8) OPG_MMA_LINE("%8", 16) OPG_MMA_LINE("%9", 24) OPG_MMA_LINE("%10", 32)
            OPG_MMA_LINE("%11", 40) OPG_MMA_LINE("%12", 48) OPG_MMA_LINE("%13", 56) OPG_MMA_LINE("%14", 64) OPG_MMA_LINE("%15", 72)
            "}" ::"r"(abase), "l"(dbase), "r"(idesc), "r"(0u), "r"(accumulate_first), "r"(acc),
            OPG_D(0), OPG_D(1), OPG_D(2), OPG_D(3), OPG_D(4), OPG_D(5), OPG_D(6), OPG_D(7), OPG_D(8), OPG_D(9)
            : "memory");
    } else if constexpr (MMAS == 8) {
        asm volatile(
            "{\n\t.reg .pred q, p;\n\t.reg .b64 d;\n\t.reg .b32 ta;\n\t"
            "elect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            OPG_MMA_LINE("%6", 0) OPG_MMA_LINE("%7", 8) OPG_MMA_LINE("%8", 16) OPG_MMA_LINE("%9", 24) OPG_MMA_LINE("%10", 32)
            OPG_MMA_LINE("%11", 40) OPG_MMA_LINE("%12", 48) OPG_MMA_LINE("%13", 56)
            "}" ::"r"(abase), "l"(dbase), "r"(idesc), "r"(0u), "r"(accumulate_first), "r"(acc),
            OPG_D(0), OPG_D(1), OPG_D(2), OPG_D(3), OPG_D(4), OPG_D(5), OPG_D(6), OPG_D(7)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred q, p;\n\t.reg .b64 d;\n\t.reg .b32 ta;\n\t"
            "elect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            OPG_MMA_LINE("%6", 0) OPG_MMA_LINE("%7", 8) OPG_MMA_LINE("%8", 16) OPG_MMA_LINE("%9", 24)
            "}" ::"r"(abase), "l"(dbase), "r"(idesc), "r"(0u), "r"(accumulate_first), "r"(acc),
            OPG_D(0), OPG_D(1), OPG_D(2), OPG_D(3)
            : "memory");
    }
#undef OPG_D
}

template <int C>
__device__ __forceinline__ void load_row(const float* p, bool on, float* v) {
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
        const float4 t = on ? __ldg(reinterpret_cast<const float4*>(p) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}

// max-shifted exp of one row, packed to bf16 pairs; returns the shift (0 for an all -inf row)
template <int C>
__device__ __forceinline__ float exp_pack(const float* v, uint32_t* e) {
    float m = v[0];
#pragma unroll
    for (int q = 1; q < C; ++q) m = fmaxf(m, v[q]);
    m = (m == -CUDART_INF_F) ? 0.f : m;
    const float ms = m * 1.4426950408889634f;
#pragma unroll
    for (int q = 0; q < C / 2; ++q)
        e[q] = pack_bf16(ex2_approx(fmaf(v[2 * q], 1.4426950408889634f, -ms)), ex2_approx(fmaf(v[2 * q + 1], 1.4426950408889634f, -ms)));
    return m;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Thread organisation: 17 warps, specialised so that the resources of the pipeline run concurrently:
//   warps 0-7   GENERATORS: the fp16 pipe.  Warp w serves TMEM lane quadrant w % 4 (32 rows); the two warps of a quadrant
//               take alternate stages of the 4-deep A ring (bf16x2 multiplies -> tcgen05.st), so one multiplies while the
//               other waits for its TMEM stores / stage slot.
//   warps 8-15  LOADERS / EPILOGUE: LSU + MUFU.  Two threads per row (same lane, warps w and w + 4): "half" 0 / 1 owns the
//               left / right child row and half of the outputs.  They run the prologue of tile t + 1 (max, exp, pack ->
//               shared memory; the global loads were issued one tile earlier) and the epilogue of tile t - 1 (tcgen05.ld,
//               log / scale, row stores) while tile t is being generated.
//   warp 16     MMA issuer (also streams the weight image with the TMA engine).
// Row vectors and accumulators are double buffered, A stages quadruple buffered; all hand-offs are mbarriers (producer
// arrive = release, consumer try_wait = acquire).
template <int C, int MODE>
__global__ void __launch_bounds__(544, 1) opg_kernel(OpgArgs a) {
    constexpr int NPAD = (C + 15) / 16 * 16;
    constexpr int NVG = C / 8;                 // 8-element groups per vector
    constexpr int K = C * C;
    constexpr int W_BYTES = K * C * 2;         // image pitch = C columns
    constexpr int NPAIR = C / 2;               // packed bf16 registers per vector
    constexpr int STEPS = NPAIR / 2;           // stages per tile: two packed U registers (four u's, 4*C k-values) each
    constexpr int MMAS = C / 4;                // K = 16 instructions per stage
    constexpr int NST = 4;                     // A stage ring depth
    constexpr int ACC_COLS = 64;
    constexpr int A_COLS = 2 * C;              // 4*C bf16 = 2*C 32-bit columns per stage
    constexpr int A_BASE = 2 * ACC_COLS;
    constexpr int HC = C / 2;                  // outputs per loader half
    static_assert(C % 8 == 0 && C <= 48 && NPAIR % 2 == 0 && A_BASE + NST * A_COLS <= 512, "unsupported shape");
    constexpr uint32_t IDESC = instr_desc_bf16(128, NPAD, 0, MODE == OPG_FWD ? 1 : 0);

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* wsm = smem;                                                       // W_BYTES + 1024 (aliased N-group overrun)
    uint32_t* Us = reinterpret_cast<uint32_t*>(smem + W_BYTES + 1024);         // [2 buf][NPAIR][128]  U vector (packed bf16)
    uint32_t* Vs = Us + 2 * NPAIR * 128;                                       // [2 buf][NPAIR][128]  V vector
    uint32_t* Es = Vs + 2 * NPAIR * 128;                                       // DX:  [2 buf][NPAIR][128] epilogue scale vector
    float* Xs = reinterpret_cast<float*>(Es);                                  // FWD: [2 children][128][C] fp32 rows staged by TMA
    constexpr int ES_WORDS = MODE == OPG_FWD ? 2 * 128 * C : 2 * NPAIR * 128;
    float* Ms = reinterpret_cast<float*>(Es + ES_WORDS);                       // [2 buf][2][128]      row shifts mL, mR
    uint64_t* st_full = reinterpret_cast<uint64_t*>(Ms + 2 * 2 * 128);         // [NST]  generators -> MMA
    uint64_t* st_empty = st_full + NST;                                        // [NST]  MMA commit -> generators
    uint64_t* vec_full = st_empty + NST;                                       // [2]    loaders -> generators
    uint64_t* vec_empty = vec_full + 2;                                        // [2]    generators -> loaders
    uint64_t* acc_full = vec_empty + 2;                                        // [2]    MMA commit -> epilogue
    uint64_t* acc_empty = acc_full + 2;                                        // [2]    epilogue -> MMA
    uint64_t* bar_w = acc_empty + 2;                                           // [1]
    uint64_t* x_full = bar_w + 1;                                              // [1]    FWD: TMA (x rows) -> loaders
    uint64_t* h_full = x_full + 1;                                             // [2]    DX: TMA (forward hand-off) -> loaders
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_full + 2);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

    // warp index through a shuffle: the idiom that tells the compiler the value (and every branch on it) is warp-uniform
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) {
            mbar_init(st_full + i, 4);   // the four quadrant warps that produce this stage
            mbar_init(st_empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(vec_full + i, 8);
            mbar_init(vec_empty + i, 8);
            mbar_init(acc_full + i, 1);
            mbar_init(acc_empty + i, 8);
        }
        mbar_init(bar_w, 1);
        mbar_init(x_full, 1);
        mbar_init(h_full + 0, 1);
        mbar_init(h_full + 1, 1);
        *abort_flag = 0;
        mbar_fence_init();
    }
    if (warp == 16) tmem_alloc(tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    // The whole TMEM (512 columns) of this SM is ours, so the allocation starts at lane 0 / column 0.
    if (threadIdx.x == 0 && *tmem_slot != 0) *abort_flag = 1;
    constexpr uint32_t tmem_base = 0;
    __syncthreads();

    const int nitems = a.d_out * a.R * a.nsplit;
    const int tiles_per_split = (a.ntiles + a.nsplit - 1) / a.nsplit;
    uint32_t n_stage = 0, n_tile = 0, n_epi = 0, items_done = 0;   // per-role ring positions (persist across items)
    uint32_t m_stage = 0, m_tile = 0;                               // the MMA warp's own counters (kept warp-uniform)
    TC2_DECL;

    // Items are handed out dynamically (one atomic per item): when another kernel holds some SMs -- NCCL's all-reduce
    // of the upper layers' gradients overlaps this backward -- the CTAs that start late simply find less work left,
    // instead of a static 1/gridDim share that would become a second wave.
    __shared__ int s_next[2];
    for (;; ++items_done) {
        if (threadIdx.x == 0) s_next[items_done & 1] = (int)atomicAdd(a.counter, 1u);
        __syncthreads();  // the previous item is finished by every role (its last epilogue waited for its last MMA)
        const int item = __shfl_sync(0xffffffffu, s_next[items_done & 1], 0);
        if (item >= nitems) break;
        const int split = item % a.nsplit;
        const int sr = item / a.nsplit;
        const int r = sr % a.R, s = sr / a.R;
        const int tile_begin = split * tiles_per_split;
        const int tile_end = min(a.ntiles, tile_begin + tiles_per_split);
        const bool has_l = 2 * s < a.d_in, has_r = 2 * s + 1 < a.d_in;
        if (__shfl_sync(0xffffffffu, *abort_flag, 0)) break;
        if (warp == 16) {
            // =============================================== MMA issuer ===============================================
            if (lane == 0) {
                mbar_arrive_expect_tx(bar_w, W_BYTES);
                const uint8_t* src = a.wimg + ((size_t)(MODE == OPG_DXL ? 1 : 0) * a.d_out * a.R + sr) * W_BYTES;
                constexpr int CH = 16000;
                for (int off = 0; off < W_BYTES; off += CH) bulk_g2s(wsm + off, src + off, min(CH, W_BYTES - off), bar_w);
            }
            __syncwarp();
            TC2_T(t_mw0);
            mbar_wait_sticky(bar_w, items_done & 1, abort_flag);
            fence_after_sync();
            TC2_T(t_mw1);
            TC2_ADD(3, t_mw0, t_mw1);
            const uint32_t wsm_addr = smem_u32(wsm);
            // The tensor pipe queues only a few MMAs, so the issue of a stage returns when that stage is nearly finished; a
            // blocking wait (~90 clk) for the next stage in between would leave the pipe idle.  The next hand-off is therefore
            // PROBED before the current stage is issued (the probe's latency hides behind the issue); the blocking wait is
            // only the fall-back when the generators are not ahead.
            bool acc_ready = false, st_ready = false;
            for (int tile = tile_begin; tile < tile_end; ++tile, ++m_tile) {
                const uint32_t ab = m_tile & 1;
                TC2_T(t_m0);
                if (!acc_ready) mbar_wait_sticky(acc_empty + ab, ((m_tile >> 1) & 1) ^ 1, abort_flag);  // epilogue of tile n - 2 has read it
                fence_after_sync();
                TC2_T(t_m1);
                TC2_ADD(0, t_m0, t_m1);
                const uint32_t acc = tmem_base + ab * ACC_COLS;
#pragma unroll 1
                for (int step = 0; step < STEPS; ++step, ++m_stage) {
                    const uint32_t st = m_stage % NST, par = (m_stage / NST) & 1;
                    TC2_T(t_m2);
                    if (!st_ready) mbar_wait_sticky(st_full + st, par, abort_flag);
                    fence_after_sync();
                    TC2_T(t_m3);
                    TC2_ADD(1, t_m2, t_m3);
                    {   // probe the hand-offs needed right after this stage
                        const uint32_t nx = m_stage + 1;
                        st_ready = mbar_probe_uniform(st_full + nx % NST, (nx / NST) & 1);
                        if (step == STEPS - 1) acc_ready = mbar_probe_uniform(acc_empty + (ab ^ 1), (((m_tile + 1) >> 1) & 1) ^ 1);
                    }
                    const uint32_t abase = tmem_base + A_BASE + st * A_COLS;
                    if (OPG_DEBUG(a) != 1 && OPG_DEBUG(a) != 8) {
                        // base descriptor of the stage (see opg_desc_delta): FWD advances 4*NVG K groups of NVG*128 bytes per
                        // stage; DX advances four u's = 4*NVG*NVG blocks of 128 bytes
                        const uint32_t sbytes = (uint32_t)step * (4 * NVG * NVG * 128);
                        const uint64_t dbase = MODE == OPG_FWD ? smem_desc(wsm_addr + sbytes, NVG * 128, 128) : smem_desc(wsm_addr + sbytes, 0, NVG * 128);
                        opg_issue_stage<C, MODE>(acc, abase, dbase, OPG_DBGN(a) ? instr_desc_bf16(128, OPG_DBGN(a), 0, MODE == OPG_FWD ? 1 : 0) : IDESC,
                                                 step ? 1u : 0u);
                    }
                    mma_commit_uniform(st_empty + st);
                    if (step == STEPS - 1) mma_commit_uniform(acc_full + ab);
                    TC2_T(t_m4);
                    TC2_ADD(2, t_m3, t_m4);
                }
                TC2_COUNT(5);
            }
            // the probes must not leak into the next item: its barriers are waited for from scratch
            (void)acc_ready; (void)st_ready;
        } else if (warp < 8) {
            // =============================================== generators ===============================================
            const uint32_t gsel = warp >> 2;          // this warp produces the stages with n_stage % 2 == gsel
            const int row = (warp & 3) * 32 + lane;
            const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
            for (int tile = tile_begin; tile < tile_end; ++tile, ++n_tile) {
                TC2_T(t_start);
                const uint32_t vb = n_tile & 1;
                mbar_wait_sticky(vec_full + vb, (n_tile >> 1) & 1, abort_flag);
                const uint32_t* my_u = Us + vb * NPAIR * 128 + row;
                const uint32_t* my_v = Vs + vb * NPAIR * 128 + row;
                uint32_t V[NPAIR];
                if (MODE == OPG_FWD) {
#pragma unroll
                    for (int q = 0; q < NPAIR; ++q) V[q] = my_v[q * 128];
                } else {
                    // DX: V = G lives in the core-matrix layout of the dW kernel's B operand: [row/8][o/8][row%8][4 words]
                    const uint4* gv = reinterpret_cast<const uint4*>(Vs + vb * NPAIR * 128 + (row >> 3) * (NVG * 32) + (row & 7) * 4);
#pragma unroll
                    for (int og = 0; og < NVG; ++og) {
                        const uint4 t4 = gv[og * 8];
                        V[4 * og] = t4.x; V[4 * og + 1] = t4.y; V[4 * og + 2] = t4.z; V[4 * og + 3] = t4.w;
                    }
                }
                TC2_T(t_v);
                TC2_ADD(0, t_start, t_v);
                // our stages of this tile; the first half of a stage is multiplied before its slot is known to be free and
                // while the previous stage's TMEM stores are still in flight.  The packed U registers of the stage (and of the
                // next one) are fetched from shared memory BEFORE the TMEM stores are issued: both go through the MIO queue,
                // and a load queued behind 5 KB of tcgen05.st data would stall the multiplies that depend on it.
                int i = (int)((gsel - n_stage) & 1);
                uint32_t p[C];
                uint32_t e0 = 0, e1 = 0, e0n = 0;
                if (i < STEPS) {
                    e0 = my_u[(2 * i) * 128];
                    e1 = my_u[(2 * i + 1) * 128];
                    if ((OPG_DEBUG(a) != 3 && OPG_DEBUG(a) < 7 || OPG_DEBUG(a) == 9)) {
                        const uint32_t u0 = bcast_lo(e0), u1 = bcast_hi(e0);
#pragma unroll
                        for (int q = 0; q < NPAIR; ++q) { p[q] = hmul2_bf16(u0, V[q]); p[NPAIR + q] = hmul2_bf16(u1, V[q]); }
                    }
                }
                for (; i < STEPS; i += 2) {
                    const uint32_t ns = n_stage + i;
                    const uint32_t st = ns % NST, use = ns / NST;
                    const uint32_t cols = tmem_base + lane_base + A_BASE + st * A_COLS;
                    const bool more = i + 2 < STEPS;
                    if (more) e0n = my_u[(2 * (i + 2)) * 128];
                    TC2_T(t_w0);
                    mbar_wait_sticky(st_empty + st, (use & 1) ^ 1, abort_flag);
                    fence_after_sync();
                    TC2_T(t_w1);
                    TC2_ADD(1, t_w0, t_w1);
                    if (OPG_DEBUG(a) != 2 && (OPG_DEBUG(a) != 3 && OPG_DEBUG(a) < 7 || OPG_DEBUG(a) == 9)) tmem_store_cols<C>(cols, p);
                    if ((OPG_DEBUG(a) != 3 && OPG_DEBUG(a) < 7 || OPG_DEBUG(a) == 9)) {
                        const uint32_t u0 = bcast_lo(e1), u1 = bcast_hi(e1);
#pragma unroll
                        for (int q = 0; q < NPAIR; ++q) { p[q] = hmul2_bf16(u0, V[q]); p[NPAIR + q] = hmul2_bf16(u1, V[q]); }
                    }
                    if (more) e1 = my_u[(2 * (i + 2) + 1) * 128];
                    if (OPG_DEBUG(a) != 2 && (OPG_DEBUG(a) != 3 && OPG_DEBUG(a) < 7 || OPG_DEBUG(a) == 9)) tmem_store_cols<C>(cols + C, p);
                    if (more && (OPG_DEBUG(a) != 3 && OPG_DEBUG(a) < 7 || OPG_DEBUG(a) == 9)) {
                        const uint32_t u0 = bcast_lo(e0n), u1 = bcast_hi(e0n);
#pragma unroll
                        for (int q = 0; q < NPAIR; ++q) { p[q] = hmul2_bf16(u0, V[q]); p[NPAIR + q] = hmul2_bf16(u1, V[q]); }
                    }
                    TC2_T(t_g);
                    TC2_ADD(2, t_w1, t_g);
                    tmem_wait_st();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(st_full + st);
                    TC2_T(t_s);
                    TC2_ADD(3, t_g, t_s);
                    TC2_COUNT(4);
                }
                n_stage += STEPS;
                __syncwarp();
                if (lane == 0) mbar_arrive(vec_empty + vb);  // U / V of this tile are no longer read by this warp
                TC2_COUNT(5);
            }
        } else {
            // ========================================== loaders / epilogue ==========================================
            const int lw_ = warp - 8;
            const int half = lw_ >> 2;
            const int row = (lw_ & 3) * 32 + lane;
            const uint32_t lane_base = (uint32_t)((lw_ & 3) * 32) << 16;
            const bool has_mine = half ? has_r : has_l;
            const float* xbase = a.x + ((int64_t)(2 * s + half) * a.R + r) * a.B * C;
            const int nt = tile_end - tile_begin;
            // FWD: the 128 rows of a child are contiguous (20 KB at C = 40), so the TMA engine stages them one tile ahead and
            // every thread picks its row from shared memory: a per-thread 16-byte global load touches 32 cache lines per warp
            // instruction (32 L1 wavefronts), which made these loads the most expensive part of the prologue.
            // DX: the loader also needs out / g slices; everything of a tile is requested together at its start.
            constexpr bool STAGED = MODE == OPG_FWD;
            auto stage_x = [&](int tile) {  // whole warp 8 (one elected lane issues)
                const int rows = min(128, a.B - tile * 128);
                const uint32_t bytes = (uint32_t)rows * C * 4;
                mbar_arrive_expect_tx_elect(x_full, (has_l ? bytes : 0u) + (has_r ? bytes : 0u));
                if (has_l) bulk_g2s_elect(Xs, a.x + (((int64_t)(2 * s) * a.R + r) * a.B + (int64_t)tile * 128) * C, bytes, x_full);
                if (has_r) bulk_g2s_elect(Xs + 128 * C, a.x + (((int64_t)(2 * s + 1) * a.R + r) * a.B + (int64_t)tile * 128) * C, bytes, x_full);
            };
            if (STAGED && warp == 8 && nt > 0) stage_x(tile_begin);
            float v[C];
            for (int i = 0; i <= nt; ++i) {
                if (i < nt) {
                    // ---------------- prologue of tile i: half 0 owns the left child row, half 1 the right child row
                    TC2_T(t_p0);
                    const int tile = tile_begin + i;
                    const uint32_t vb = n_tile & 1;
                    const int64_t b = (int64_t)tile * 128 + row;
                    const bool valid = b < a.B;
                    const int64_t bb = valid ? b : 0;
                    uint32_t* my_u = Us + vb * NPAIR * 128 + row;
                    uint32_t* my_v = Vs + vb * NPAIR * 128 + row;
                    uint32_t* my_e = Es + vb * NPAIR * 128 + row;
                    float* my_m = Ms + vb * 2 * 128 + row;
                    constexpr int FH_BYTES = 2 * NPAIR * 512 + 1024;   // forward hand-off per tile: eL, eR, shifts
                    const bool from_fwd = MODE != OPG_FWD && a.fhand != nullptr;
                    const bool g_ready = MODE == OPG_DXL && from_fwd && a.ghand != nullptr;   // G arrives by TMA as well
                    if (from_fwd) {
                    } else if (STAGED) {
                        mbar_wait_sticky(x_full, n_tile & 1, abort_flag);
                        const float4* rowp = reinterpret_cast<const float4*>(Xs + (half * 128 + row) * C);
#pragma unroll
                        for (int q = 0; q < C / 4; ++q) {
                            const float4 t4 = (valid && has_mine) ? rowp[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                            v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
                        }
                    } else {
                        load_row<C>(xbase + bb * C, valid && has_mine, v);
                    }
                    uint32_t e[NPAIR];
                    float m = 0.f;
                    if (from_fwd) {
                    } else if ((OPG_DEBUG(a) != 4 && OPG_DEBUG(a) < 7 || OPG_DEBUG(a) == 9)) m = exp_pack<C>(v, e);
                    else {
#pragma unroll
                        for (int j = 0; j < NPAIR; ++j) e[j] = __float_as_uint(v[j]);
                    }
                    {
                        if (!STAGED && !from_fwd && i + 2 < nt && b + 256 < a.B && has_mine) {  // x rows are only read without the forward's hand-off
                            const char* nx = reinterpret_cast<const char*>(xbase + (b + 256) * C);
                            prefetch_l2(nx);
                            prefetch_l2(nx + C * 4 - 4);
                        }
                    }
                    mbar_wait_sticky(vec_empty + vb, ((n_tile >> 1) & 1) ^ 1, abort_flag);  // generators are done with tile n - 2
                    if (MODE != OPG_DXL && a.hand && warp == 8) bulk_wait_group_read<1>();  // ... and so is the TMA store (issuing lane)
                    // every loader thread has finished the epilogue that read this buffer (partner rows' shift / scale) -- and,
                    // FWD, has taken its row out of the staging area, which the TMA engine may now refill with the next tile
                    named_bar_sync(1, 256);
                    if (STAGED && warp == 8 && i + 1 < nt) stage_x(tile + 1);
                    if (from_fwd) {
                        // DXR: U = eL, scale = eR.  DXL: U = eR, scale = eL.  Straight from the forward's hand-off by TMA.
                        if (warp == 8) {
                            const uint8_t* src = a.fhand + ((size_t)sr * a.ntiles + tile) * FH_BYTES;
                            mbar_arrive_expect_tx_elect(h_full + vb, FH_BYTES + (g_ready ? NPAIR * 512 : 0));
                            if (g_ready)
                                bulk_g2s_elect(Vs + vb * NPAIR * 128, a.ghand + ((size_t)sr * a.ntiles + tile) * (NPAIR * 512), NPAIR * 512,
                                               h_full + vb);
                            bulk_g2s_elect(MODE == OPG_DXR ? Us + vb * NPAIR * 128 : Es + vb * NPAIR * 128, src, NPAIR * 512, h_full + vb);
                            bulk_g2s_elect(MODE == OPG_DXR ? Es + vb * NPAIR * 128 : Us + vb * NPAIR * 128, src + NPAIR * 512, NPAIR * 512, h_full + vb);
                            bulk_g2s_elect(Ms + vb * 2 * 128, src + 2 * NPAIR * 512, 1024, h_full + vb);
                        }
                    } else {
                        my_m[half * 128] = m;
                        // FWD: U = eL, V = eR.  DXR: U = eL, scale = eR.  DXL: U = eR, scale = eL.
                        uint32_t* dst = MODE == OPG_FWD ? (half ? my_v : my_u) : MODE == OPG_DXR ? (half ? my_e : my_u) : (half ? my_u : my_e);
#pragma unroll
                        for (int j = 0; j < NPAIR; ++j) dst[j * 128] = e[j];
                    }
                    if (g_ready) {
                        mbar_wait_sticky(h_full + vb, (n_tile >> 1) & 1, abort_flag);  // eL, eR, shifts and G have landed
                    } else if (MODE != OPG_FWD) {
                        // G = g * exp(shift - out) = g / (linear-domain sum); out = -inf (impossible row): no gradient.  The
                        // exponent is capped so that rows whose sum underflowed in the forward (rescue path) stay finite.
                        // Half 0 produces the o-groups [0, GH0), half 1 [GH0, NVG) (whole 16-byte chunks of 8 outputs), written
                        // in the core-matrix layout [row/8][o/8][row%8][8 bf16] that the dW kernel's B operand uses.
                        constexpr int GH0 = (NVG + 1) / 2;
                        constexpr int NGH = GH0;                       // chunks handled per half (half 1 may have fewer)
                        const int og0 = half ? GH0 : 0;
                        const int nog = half ? NVG - GH0 : GH0;
                        const float* op = a.out + (((int64_t)s * a.R + r) * a.B + bb) * C + og0 * 8;
                        const float* gp = a.g + (((int64_t)s * a.R + r) * a.B + bb) * C + og0 * 8;
                        float vo[NGH * 8], vg[NGH * 8];
#pragma unroll
                        for (int q = 0; q < NGH * 2; ++q) {
                            const bool on = valid && q < 2 * nog;
                            const float4 t4 = on ? __ldg(reinterpret_cast<const float4*>(op) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float4 u4 = on ? __ldg(reinterpret_cast<const float4*>(gp) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                            vo[4 * q] = t4.x; vo[4 * q + 1] = t4.y; vo[4 * q + 2] = t4.z; vo[4 * q + 3] = t4.w;
                            vg[4 * q] = u4.x; vg[4 * q + 1] = u4.y; vg[4 * q + 2] = u4.z; vg[4 * q + 3] = u4.w;
                        }
                        if (from_fwd) mbar_wait_sticky(h_full + vb, (n_tile >> 1) & 1, abort_flag);  // eL, eR and the shifts have landed
                        else named_bar_sync(2, 256);  // both shifts are in shared memory
                        const float sh2 = (my_m[0] + my_m[128]) * 1.4426950408889634f;
                        uint4* gdst = reinterpret_cast<uint4*>(Vs + vb * NPAIR * 128 + (row >> 3) * (NVG * 32) + (row & 7) * 4);
#pragma unroll
                        for (int og = 0; og < NGH; ++og) {
                            uint32_t w4[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float o0 = vo[8 * og + 2 * q], o1 = vo[8 * og + 2 * q + 1];
                                const float g0v = (!valid || o0 == -CUDART_INF_F) ? 0.f : vg[8 * og + 2 * q] * ex2_approx(fminf(fmaf(o0, -1.4426950408889634f, sh2), 115.f));
                                const float g1v = (!valid || o1 == -CUDART_INF_F) ? 0.f : vg[8 * og + 2 * q + 1] * ex2_approx(fminf(fmaf(o1, -1.4426950408889634f, sh2), 115.f));
                                w4[q] = pack_bf16(g0v, g1v);
                            }
                            if (og < nog) gdst[(og0 + og) * 8] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                        }
                    }
                    if (MODE != OPG_DXL && a.hand) fence_proxy_async();  // the row vectors are also read by the TMA engine below
                    __syncwarp();
                    if (lane == 0) mbar_arrive(vec_full + vb);
                    if (MODE != OPG_DXL && a.hand && warp == 8) {
                        // hand-off by TMA bulk stores once all eight loader warps have written the tiles (the barrier completes
                        // on the 8th arrival).  FWD: eL, eR, shifts for the backward kernels.  DXR: the G tile for the dW kernel.
                        mbar_wait_sticky(vec_full + vb, (n_tile >> 1) & 1, abort_flag);
                        if (MODE == OPG_FWD) {
                            uint8_t* dst = a.hand + ((size_t)sr * a.ntiles + tile) * FH_BYTES;
                            bulk_s2g_elect(dst, Us + vb * NPAIR * 128, NPAIR * 512);
                            bulk_s2g_elect(dst + NPAIR * 512, Vs + vb * NPAIR * 128, NPAIR * 512);
                            bulk_s2g_elect(dst + 2 * NPAIR * 512, Ms + vb * 2 * 128, 1024);
                        } else {
                            bulk_s2g_elect(a.hand + ((size_t)sr * a.ntiles + tile) * (NPAIR * 512), Vs + vb * NPAIR * 128, NPAIR * 512);
                        }
                        bulk_commit_group();
                    }
                    ++n_tile;
                    TC2_T(t_p1);
                    TC2_ADD(0, t_p0, t_p1);
                }
                if (i >= 1) {
                    // ---------------- epilogue of tile i - 1: this thread owns outputs [half*HC, half*HC + HC)
                    TC2_T(t_e0);
                    const int tile = tile_begin + i - 1;
                    const uint32_t ab = n_epi & 1;
                    const int64_t b = (int64_t)tile * 128 + row;
                    const bool valid = b < a.B;
                    const int64_t bb = valid ? b : 0;
                    mbar_wait_sticky(acc_full + ab, (n_epi >> 1) & 1, abort_flag);
                    fence_after_sync();
                    TC2_T(t_e1);
                    TC2_ADD(1, t_e0, t_e1);
                    constexpr int LD0 = (HC / 8) * 8;          // first column block loaded by half 1 (multiple of 8 <= HC)
                    constexpr int NLD = ((HC + 7) / 8) * 8 + (HC % 8 ? 8 : 0);   // columns loaded per thread (multiple of 8)
                    constexpr int OFF1 = HC - LD0;             // position of half 1's first output inside its loaded columns
                    uint32_t accv[NLD];
                    {
                        const uint32_t cbase = tmem_base + lane_base + ab * ACC_COLS + (half ? LD0 : 0);
#pragma unroll
                        for (int q = 0; q < NLD / 8; ++q) tmem_ld8(cbase + q * 8, accv + q * 8);
                    }
                    tmem_wait_ld();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + ab);  // the accumulator may be overwritten by tile n + 2
                    float av[HC];
#pragma unroll
                    for (int q = 0; q < HC; ++q) av[q] = __uint_as_float(half ? accv[q + OFF1] : accv[q]);
                    const float* my_m = Ms + ab * 2 * 128 + row;
                    if ((OPG_DEBUG(a) == 4 || OPG_DEBUG(a) == 7 || OPG_DEBUG(a) == 8)) {
                    } else if (MODE == OPG_FWD) {
                        const float shift = my_m[0] + my_m[128];
                        float* op = a.y + (((int64_t)s * a.R + r) * a.B + bb) * C + half * HC;
                        bool bad = false;
#pragma unroll
                        for (int o4 = 0; o4 < HC; o4 += 4) {
                            float res[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float sum = av[o4 + u];
                                bad |= !(sum >= 1e-30f);
                                res[u] = fmaf(__log2f(sum), 0.6931471805599453f, shift);
                            }
                            if (valid) *reinterpret_cast<float4*>(op + o4) = make_float4(res[0], res[1], res[2], res[3]);
                        }
                        if (bad && valid && !OPG_DEBUG(a)) opg_exact_row<C, HC>(a, s, r, b, half * HC, op);
                    } else {
                        const uint32_t* my_e = Es + ab * NPAIR * 128 + row;
                        const int sc = MODE == OPG_DXR ? 2 * s + 1 : 2 * s;
                        if (valid && sc < a.d_in) {
                            float* dp = a.y + (((int64_t)sc * a.R + r) * a.B + b) * C + half * HC;
#pragma unroll
                            for (int q = 0; q < HC / 4; ++q) {
                                const uint32_t e0 = my_e[(half * (HC / 2) + 2 * q) * 128], e1 = my_e[(half * (HC / 2) + 2 * q + 1) * 128];
                                *reinterpret_cast<float4*>(dp + 4 * q) =
                                    make_float4(bf16_lo(e0) * av[4 * q], bf16_hi(e0) * av[4 * q + 1], bf16_lo(e1) * av[4 * q + 2],
                                                bf16_hi(e1) * av[4 * q + 3]);
                            }
                        }
                    }
                    ++n_epi;
                    TC2_T(t_e2);
                    TC2_ADD(2, t_e1, t_e2);
                }
            }
        }
    }
    // teardown
    if (MODE != OPG_DXL && warp == 8) bulk_wait_group<0>();
    if (warp == 0) TC2_FLUSH(0);
    if (warp == 8) TC2_FLUSH(6);
    if (warp == 16) TC2_FLUSH(12);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 16) tmem_dealloc(tmem_base, 512);
    if (threadIdx.x == 0) tc_abort_trap(abort_flag, &g_tc2_aborts);
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
    const int c = threadIdx.x;
    const int ocol = (blockIdx.x % tiles_per_s) * 64 + c;
    if (c < 64 && ocol < ocr) {
        float mm = -CUDART_INF_F;
#pragma unroll
        for (int q = 0; q < 16; ++q) mm = fmaxf(mm, sm_m[q][c]);
        float tot = 0.f;
        if (mm != -CUDART_INF_F) {
#pragma unroll
            for (int q = 0; q < 16; ++q) tot += sm_s[q][c] * __expf(sm_m[q][c] - mm);
        }
        logz[(int64_t)s * ocr + ocol] = mm == -CUDART_INF_F ? -CUDART_INF_F : mm + logf(tot);
    }
}

// fp32 (log-)weights [1][d][K][OC][R]  ->  bf16 exp() core-matrix image [d][R][K/8][OC/8][8 (k)][8 (o)].
// One block per (scope, group of 8 k-values): the [8][OC][R] slab is read with coalesced float4 loads into a padded
// shared-memory tile [8][OC][R|1] (+ pad so that the k-stride is 9 mod 32 banks); then every thread assembles one
// 16-byte image row (8 outputs of one k), consecutive threads taking consecutive rows of one repetition so that both
// the shared-memory reads (conflict free) and the global stores (640 contiguous bytes per repetition) are efficient.
__global__ void __launch_bounds__(256) tc_wimage_kernel(const float* __restrict__ lw, const float* __restrict__ logz,
                                                       uint8_t* __restrict__ wimg, int d, int K, int OC, int R, int RP,
                                                       int KS) {
    extern __shared__ float tile[];  // [8][KS], row (kk, o) at kk * KS + o * RP
    const int kg = blockIdx.x % (K / 8);
    const int s = blockIdx.x / (K / 8);
    const int ocr = OC * R;
    const float4* src = reinterpret_cast<const float4*>(lw + ((int64_t)s * K + (int64_t)kg * 8) * ocr);
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
    // work item = (r, og, kk) with kk fastest: chunk index inside (s, r) is (kg * NOG + og) * 8 + kk
    for (int i = threadIdx.x; i < NOG * 8 * R; i += blockDim.x) {
        const int kk = i & 7, og = (i >> 3) % NOG, r = i / (8 * NOG);
        const float* t = tile + kk * KS + og * 8 * RP + r;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(t[0], t[RP]), p1 = __floats2bfloat162_rn(t[2 * RP], t[3 * RP]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(t[4 * RP], t[5 * RP]), p3 = __floats2bfloat162_rn(t[6 * RP], t[7 * RP]);
        uint4 v;
        v.x = *reinterpret_cast<uint32_t*>(&p0); v.y = *reinterpret_cast<uint32_t*>(&p1);
        v.z = *reinterpret_cast<uint32_t*>(&p2); v.w = *reinterpret_cast<uint32_t*>(&p3);
        dst[(((int64_t)s * R + r) * (K / 8) + kg) * (NOG * 8) + og * 8 + kk] = v;
    }
}

// Pair-transposed image: W2[(r',l),o] = W[(l,r'),o], a permutation of the image's 16-byte rows (8 o-values each).
// One thread per 16-byte chunk of the destination (coalesced writes).
__global__ void __launch_bounds__(256) tc_wimage_t_kernel(const uint4* __restrict__ img, uint4* __restrict__ img2, int C,
                                                         int64_t nchunks) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nchunks) return;
    const int NVG = C / 8;
    const int per_sr = C * C * NVG;       // chunks per (s, r): K rows x NVG o-groups
    const int64_t sr = i / per_sr;
    const int e = (int)(i % per_sr);
    // destination chunk (k2 = rp*C + l, og): block (k2/8, og), row k2%8
    const int row = e % 8, og = (e / 8) % NVG, k2g = e / (8 * NVG);
    const int k2 = k2g * 8 + row;
    const int rp = k2 / C, l = k2 % C;
    const int k = l * C + rp;
    const int srce = ((k / 8) * NVG + og) * 8 + (k % 8);
    img2[i] = img[sr * per_sr + srce];
}

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
    unsigned int* c;
    if (cs == cudaStreamCaptureStatusActive) {
        const unsigned int i = next_graph[dev].fetch_add(1);
        if (i >= TC_GRAPH_SLOTS) return nullptr;          // > 3072 tensor-core launches captured in this process
        c = b + TC_EAGER_SLOTS + i;
    } else {
        c = b + (next_eager[dev].fetch_add(1) % TC_EAGER_SLOTS);
    }
    if (cudaMemsetAsync(c, 0, sizeof(unsigned int), st) != cudaSuccess) return nullptr;
    return c;
}

// cudaFuncSetAttribute is per device: remember which devices a kernel has been configured on
bool tc_configured_on_device(std::atomic<unsigned long long>& mask, int* dev_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    *dev_out = dev;
    return dev < 64 && ((mask.load() >> dev) & 1ull);
}

template <int C, int MODE>
static int launch_opg(const OpgArgs& a, cudaStream_t st) {
    constexpr int W_BYTES = C * C * C * 2;
    constexpr int SMEM = W_BYTES + 1024 + 2 * 2 * (C / 2) * 128 * 4 + (MODE == OPG_FWD ? 2 * 128 * C * 4 : 2 * (C / 2) * 128 * 4) +
                         2 * 2 * 128 * 4 + (2 * 4 + 12) * 8 + 16;
    static std::atomic<unsigned long long> configured{0};
    int dev = 0, sms = 148;
    if (!tc_configured_on_device(configured, &dev)) {
        cudaError_t e = cudaFuncSetAttribute(opg_kernel<C, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return spn_cuda_status(e);
        if (dev < 64) configured.fetch_or(1ull << dev);
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nitems = a.d_out * a.R * a.nsplit;
    OpgArgs b = a;
    b.counter = tc_next_counter(st);
    if (!b.counter) return SPN_ERR_ARG;
    opg_kernel<C, MODE><<<nitems < sms ? nitems : sms, 544, SMEM, st>>>(b);
    return spn_launch_status();
}

template <int MODE>
static int dispatch_opg(int c, const OpgArgs& a, cudaStream_t st) {
    if (c == 40) return launch_opg<40, MODE>(a, st);
    if (c == 32) return launch_opg<32, MODE>(a, st);
    if (c == 16) return launch_opg<16, MODE>(a, st);
    return SPN_ERR_UNSUPPORTED;
}

// SPN_TC_DEBUG (timing experiments only; results are garbage): 1 no MMA issue, 2 no TMEM stores, 3 generators idle,
// 4 loaders skip exp / epilogue arithmetic, 5 loaders skip the row loads, 7 = 3 + 4 + 5, 8 = 7 + 1, 9 phase profile.
static void opg_split(OpgArgs& a, int B, int d_out, int R) {
    a.ntiles = (B + 127) / 128;
    int nsplit = (8 * 148 + d_out * R - 1) / (d_out * R);
    if (nsplit > a.ntiles / 4) nsplit = a.ntiles / 4;
    if (nsplit < 1) nsplit = 1;
#ifdef SPN_TC_DEBUGSW   // tuning experiments: only in builds with the experiment switches
    const char* ns = getenv("SPN_TC_NSPLIT");
    if (ns && atoi(ns) > 0 && atoi(ns) <= a.ntiles) nsplit = atoi(ns);
    const char* dbg = getenv("SPN_TC_DEBUG");
    a.debug = dbg ? atoi(dbg) : 0;
    const char* dn = getenv("SPN_TC_N");
    a.dbg_n = dn ? atoi(dn) : 0;
#endif
    a.nsplit = nsplit;
}

extern "C" int spn_sum_tc_supported(int c, int oc, int R) {
    (void)R;
    return (c == 40 && oc == 40) || (c == 32 && oc == 32) || (c == 16 && oc == 16) ? 1 : 0;
}

extern "C" int64_t spn_sum_tc_image_bytes(int d_out, int c, int oc, int R) {
    return 2 * (int64_t)d_out * R * c * c * oc * 2 + 1024;
}

extern "C" int spn_sum_tc_prepare(const float* lw, int d_out, int c, int oc, int R, float* logz, void* wimg, void* stream) {
    if (!lw || !wimg || !spn_sum_tc_supported(c, oc, R)) return SPN_ERR_UNSUPPORTED;
    const int K = c * c;
    cudaStream_t st = (cudaStream_t)stream;
    if (logz) {
        const int ocr = oc * R, tiles = (ocr + 63) / 64;
        tc_logz_kernel<<<d_out * tiles, 256, 0, st>>>(lw, logz, K, ocr, tiles);
        int rc = spn_launch_status();
        if (rc != SPN_OK) return rc;
    }
    const int RP = R | 1;
    const int KS = oc * RP + ((9 - (oc * RP) % 32) + 32) % 32;
    const size_t smem = (size_t)8 * KS * sizeof(float);
    if (smem > 200 * 1024) return SPN_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(tc_wimage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return spn_cuda_status(e);
    }
    tc_wimage_kernel<<<d_out * (K / 8), 256, smem, st>>>(lw, logz, (uint8_t*)wimg, d_out, K, oc, R, RP, KS);
    int rc = spn_launch_status();
    if (rc != SPN_OK) return rc;
    const int64_t img_bytes = (int64_t)d_out * R * K * oc * 2;
    const int64_t nchunks = img_bytes / 16;
    tc_wimage_t_kernel<<<spn_blocks(nchunks, 256), 256, 0, st>>>((const uint4*)wimg, (uint4*)((uint8_t*)wimg + img_bytes), c, nchunks);
    return spn_launch_status();
}

extern "C" int spn_sum_tc_fwd(const float* x, const void* wimg, const float* lw, const float* logz, int B, int d_in,
                              int d_out, int c, int oc, int R, float* out, void* fhand, void* stream) {
    if (!x || !wimg || !lw || !out || B <= 0 || d_in <= 0 || d_in > 2 * d_out) return SPN_ERR_ARG;
    if (!spn_sum_tc_supported(c, oc, R)) return SPN_ERR_UNSUPPORTED;
    OpgArgs a{};
    a.x = x; a.wimg = (const uint8_t*)wimg; a.lw = lw; a.logz = logz; a.y = out; a.hand = (uint8_t*)fhand;
    a.B = B; a.d_in = d_in; a.d_out = d_out; a.R = R;
    opg_split(a, B, d_out, R);
    return dispatch_opg<OPG_FWD>(c, a, (cudaStream_t)stream);
}

// hand-off buffers (see OpgArgs): forward -> backward (packed eL, eR, shifts) and dX -> dW (G tiles)
extern "C" int64_t spn_sum_tc_fhand_bytes(int B, int d_out, int c, int R) {
    return (int64_t)d_out * R * ((B + 127) / 128) * (2 * (c / 2) * 512 + 1024);
}
extern "C" int64_t spn_sum_tc_hand_bytes(int B, int d_out, int c, int R) {
    return (int64_t)d_out * R * ((B + 127) / 128) * (c / 2) * 512;
}

extern "C" int spn_sum_tc_bwd_x(const float* x, const float* out, const float* g, const void* wimg, const void* fhand, int B,
                                int d_in, int d_out, int c, int oc, int R, float* dx, void* hand, void* stream) {
    if (!x || !out || !g || !wimg || !dx || B <= 0 || d_in <= 0 || d_in > 2 * d_out) return SPN_ERR_ARG;
    if (!spn_sum_tc_supported(c, oc, R)) return SPN_ERR_UNSUPPORTED;
    OpgArgs a{};
    a.x = x; a.wimg = (const uint8_t*)wimg; a.out = out; a.g = g; a.y = dx; a.hand = (uint8_t*)hand;
    a.fhand = (const uint8_t*)fhand;
    a.B = B; a.d_in = d_in; a.d_out = d_out; a.R = R;
    opg_split(a, B, d_out, R);
    // right child first: it also hands the G tiles over (to the dW kernel, and to the left-child launch below, which then
    // neither reads out / g nor recomputes G)
    int rc = dispatch_opg<OPG_DXR>(c, a, (cudaStream_t)stream);
    if (rc != SPN_OK) return rc;
    if (a.hand && a.fhand) a.ghand = a.hand;
    return dispatch_opg<OPG_DXL>(c, a, (cudaStream_t)stream);
}

extern "C" int spn_debug_tc_prof(unsigned long long* out_host, int reset) {
    unsigned long long v[24] = {0};
    cudaError_t e = cudaMemcpyFromSymbol(v, g_tc2_prof, sizeof(v));
    if (e != cudaSuccess) return spn_cuda_status(e);
    if (out_host) for (int i = 0; i < 24; ++i) out_host[i] = v[i];
    if (reset) {
        unsigned long long z[24] = {0};
        e = cudaMemcpyToSymbol(g_tc2_prof, z, sizeof(z));
        if (e != cudaSuccess) return spn_cuda_status(e);
    }
    return SPN_OK;
}

// debug: number of CTAs that hit the bounded-wait timeout since the library was loaded (synchronises)
int tc_bwd_aborts_host(int* out);

extern "C" int spn_debug_tc_aborts(int* out_host) {
    int v = 0, w = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&v, g_tc2_aborts, sizeof(int));
    if (e != cudaSuccess) return spn_cuda_status(e);
    int rc = tc_bwd_aborts_host(&w);
    if (rc != SPN_OK) return rc;
    if (out_host) *out_host = v + w;
    return SPN_OK;
}
