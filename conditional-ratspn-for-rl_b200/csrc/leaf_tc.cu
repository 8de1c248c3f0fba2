// Leaf layer on the tensor cores (tcgen05, kind::tf32) for SHARED parameters and plain inputs, internal layout.
//
// The Gaussian leaf + product over a scope is a quadratic form in x:
//     out[b, i] = sum_f  a[f,i] x[b,f]^2 + bq[f,i] x[b,f] + c[f,i] valid[b,f]
//     a = -1 / (2 sigma^2),  bq = mu / sigma^2,  c = -mu^2 / (2 sigma^2) - log sigma - log sqrt(2 pi)
// i.e. one [rows x 3 card] x [3 card x I] GEMM per (scope, repetition); a NaN input (marginalised feature,
// distributions.py:58-61, :243) zeroes its three columns.  bf16 / single tf32 operands cannot hold the cancellation between
// the three terms, so every operand is split into two tf32 words (hi = upper 19 bits, lo = remainder) and the three
// significant cross products are summed in the fp32 accumulator: 8 K entries per feature = ONE K = 8 tf32 MMA step:
//     A (rows):   x2h  x2h  x2l  xh   xh   xl   v    v
//     B (params): ah   al   ah   bh   bl   bh   ch   cl
// Measured against fp64 the result is within 3x of the exact fp32 kernel's own error (profiles/README.md, round 2).
// The parameter gradients are the matching transposed GEMM (K = rows): per (scope, repetition) the accumulator
// D[5 card x 2 I] = [x2h x2l xh xl v]^T [gh | gl] yields sum g, sum g x and sum g x^2, from which
//     dmu = sum g (x - mu) / sigma^2,   dsigma = sum g (x - mu)^2 / sigma^3 - sum g / sigma.
// Reference semantics: rat_spn.py:155-170, distributions.py:199-246, :366-382, layers.py:276-312; the tanh
// change-of-variables correction is not handled here (callers keep the fp32 kernels of leaf_tiled.cu for it).
#include <atomic>

#include "tc_common.cuh"

using namespace tc;

unsigned int* tc_next_counter(cudaStream_t st);                                             // sum_tc3.cu
bool tc_configured_on_device(std::atomic<unsigned long long>& mask, int* dev_out);          // sum_tc3.cu

__device__ int g_leaf_tc_aborts = 0;

// per-phase clocks of one CTA (block 0, lane 0 of one warp per role); experiment builds only (-DSPN_LEAF_TC_PROF)
#ifdef SPN_LEAF_TC_PROF
#define LT_DBG(bit) ((a.dbg & (bit)) != 0)   // spn_debug_leaf_tc_mode: 1 no MMA, 2 no x reads, 4 no A operand, 8 no bulk store, 16 no B operand
__device__ long long g_leaf_tc_prof[64];
#define LTP_DECL long long ltp_t = clock64(), ltp_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define LTP_LAP(i) do { const long long ltp_n = clock64(); ltp_acc[i] += ltp_n - ltp_t; ltp_t = ltp_n; } while (0)
#define LTP_FLUSH(base) do { if (blockIdx.x == 0 && lane == 0) for (int ltp_i = 0; ltp_i < 8; ++ltp_i) atomicAdd((unsigned long long*)&g_leaf_tc_prof[(base) + ltp_i], (unsigned long long)ltp_acc[ltp_i]); } while (0)
#else
#define LT_DBG(bit) false
#define LTP_DECL
#define LTP_LAP(i)
#define LTP_FLUSH(base)
#endif

namespace {

constexpr int LT_NSTAGE = 4;             // forward: A-operand stages in tensor memory (8 features = 64 columns each)
constexpr int LT_XD = 6;                 // forward: depth of the x ring in tiles (prefetch distance LT_XD - 1)
constexpr int LT_MAXCARD = 25;           // backward: 5 operand rows per feature, M = 128
constexpr int LT_BROWS = 64;             // rows (K extent) per backward stage
constexpr int LT_BSTAGES = 2;

struct LeafTcArgs {
    const float* xT;        // [F][ldx] feature-major copy of x (spn_leaf_transpose)
    const int32_t* perm;    // [F][R] or null
    const uint8_t* img;     // forward: [d0 * R][card * NP * 32] parameter operand image
    float* out;             // forward: [d0][R][B][I]
    const float* g;         // backward: [d0][R][B][I]
    const float* mu; const float* sigma; int64_t sp_f, sp_i, sp_r;
    float* dmu; float* dsigma;
    unsigned int* counter;   // backward: zeroed work counter of this launch (dynamic item scheduling)
    int B, ldx, F, I, R, card, d0, nst, img_bytes, gring, dbg;
};

__device__ __forceinline__ uint32_t tf32_hi(float v) { return __float_as_uint(v) & 0xFFFFE000u; }
__device__ __forceinline__ void tf32_split(float v, uint32_t& hi, uint32_t& lo) {
    hi = tf32_hi(v);
    lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void lt_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Instruction descriptor of kind::tf32: c_format F32 [4,6) = 1, a_format / b_format TF32 [7,10) / [10,13) = 2, K-major
// operands, n_dim = N >> 3 [17,23), m_dim = M >> 4 [24,29).
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss_tf32_uniform(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                    uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

// --------------------------------------------------------------------------------------------------------------------
// Parameter operand image: per (scope, repetition) the B operand [N = NP channels][K = 8 * card] tf32, K-major core
// matrices (8 channels x 4 K entries = 128 bytes): element (n, k) at (k / 4) * NP * 16 + (n / 8) * 128 + (n % 8) * 16 + (k % 4) * 4.
// --------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) leaf_tc_image_kernel(LeafTcArgs a, int NP) {
    const int item = blockIdx.x, s = item / a.R, r = item % a.R;
    const int f_begin = s * a.card;
    uint8_t* img = const_cast<uint8_t*>(a.img) + (size_t)item * a.img_bytes;
    for (int e = threadIdx.x; e < a.card * NP; e += blockDim.x) {
        const int jj = e / NP, n = e % NP, f = f_begin + jj;
        uint32_t ah = 0, al = 0, bh = 0, bl = 0, ch = 0, cl = 0;
        if (n < a.I && f < a.F) {
            const int64_t p = (int64_t)f * a.sp_f + (int64_t)n * a.sp_i + (int64_t)r * a.sp_r;
            const float m = a.mu[p], sg = a.sigma[p];
            const float av = -1.f / (2.f * sg * sg);
            const float bv = -2.f * av * m;
            const float cv = av * m * m - logf(sg) - SPN_LOG_SQRT_2PI;
            tf32_split(av, ah, al);
            tf32_split(bv, bh, bl);
            tf32_split(cv, ch, cl);
        }
        uint8_t* dst = img + (size_t)(2 * jj) * NP * 16 + (n >> 3) * 128 + (n & 7) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(ah, al, ah, bh);
        *reinterpret_cast<uint4*>(dst + NP * 16) = make_uint4(bl, bh, ch, cl);
    }
}

// --------------------------------------------------------------------------------------------------------------------
// Forward.  Work item = (scope, repetition); 128-row tiles.  14 warps:
//   warps 0-3   epilogue: accumulator -> registers -> this warp's 32-row staging block in shared memory -> one TMA bulk store
//               per warp and tile (32 rows are a contiguous run of 32 * I floats of the [d0][R][B][I] output);
//   warps 4-11  operand producers, two threads per row (warp % 4 = TMEM lane quadrant, (warp - 4) / 4 = feature half).  The A
//               operand never touches shared memory: the 8 split words of a feature (x2h x2h x2l xh xh xl v v) are ONE
//               tcgen05.st of 8 columns into the row's TMEM lane, and the MMA reads A from tensor memory (the shared-memory
//               variant moved 200 KB per tile through shared memory and ran at that bandwidth);
//   warp 12     MMA issuer (one K = 8 step per feature, N = NP);
//   warp 13     TMA: the item's parameter image, and the x values of every tile (one 512-byte row segment of xT per feature)
//               LT_XD - 1 tiles ahead into a ring (with loads issued right before their use the kernel ran at the memory
//               latency under the 839 MB output stream, 2.3 ms; per-thread cp.async cost 60 clk per feature and warp).
// Tensor memory: columns [0, 128) two accumulators, [128, 128 + 64 LT_NSTAGE) the A ring (8 features x 8 columns a stage).
// --------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_ts_tf32_uniform(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                    uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

template <int NP>
__global__ void __launch_bounds__(448, 1) leaf_tc_fwd_kernel(LeafTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int img_stride = (a.img_bytes + 127) & ~127;
    uint8_t* Ws = smem;                                                        // [2][img_stride]
    float* Os = reinterpret_cast<float*>(Ws + 2 * img_stride);                 // [4 warps][2][32][I]
    float* Xs = Os + 2 * 128 * a.I;                                            // [LT_XD][card][128]
    uint64_t* s_full = reinterpret_cast<uint64_t*>(Xs + LT_XD * a.card * 128); // [LT_NSTAGE] producers -> MMA
    uint64_t* s_empty = s_full + LT_NSTAGE;                                    // [LT_NSTAGE] MMA commit -> producers
    uint64_t* acc_full = s_empty + LT_NSTAGE;                                  // [2]
    uint64_t* acc_empty = acc_full + 2;                                        // [2]
    uint64_t* w_full = acc_empty + 2;                                          // [2]
    uint64_t* w_empty = w_full + 2;                                            // [2]
    uint64_t* x_full = w_empty + 2;                                            // [LT_XD] TMA -> producers
    uint64_t* x_empty = x_full + LT_XD;                                        // [LT_XD] producers -> TMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_empty + LT_XD);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
    constexpr int W_MMA = 12, W_TMA = 13;
    constexpr uint32_t A_COL0 = 128;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < LT_NSTAGE; ++i) {
            mbar_init(s_full + i, 8);
            mbar_init(s_empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(acc_full + i, 1);
            mbar_init(acc_empty + i, 4);
            mbar_init(w_full + i, 1);
            mbar_init(w_empty + i, 1);
        }
        for (int i = 0; i < LT_XD; ++i) {
            mbar_init(x_full + i, 1);
            mbar_init(x_empty + i, 8);
        }
        *abort_flag = 0;
        mbar_fence_init();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int nitems = a.d0 * a.R;
    const int ntiles = (a.B + 127) / 128;
    const int nst = a.nst;
    uint32_t n_stage = 0, n_tile = 0, n_item = 0;

    if (warp >= 4 && warp < 12) {
        // ===================================================== operand producers =====================================================
        const int quad = warp & 3, half = (warp - 4) >> 2;
        const int row = quad * 32 + lane;
        const int xslot = a.card * 128;
        uint32_t n_x = 0;
        LTP_DECL;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int cnt = min(a.card, a.F - (item / a.R) * a.card);
            for (int t = 0; t < ntiles; ++t, ++n_x) {
                const bool inb = t * 128 + row < a.B;
                const uint32_t xsl = n_x % LT_XD;
                LTP_LAP(0);
                mbar_wait_sticky(x_full + xsl, (n_x / LT_XD) & 1, abort_flag);
                LTP_LAP(1);   // wait for the x tile
                const float* xs = Xs + xsl * xslot + row;
                for (int c = 0; c < nst; ++c, ++n_stage) {
                    const uint32_t slot = n_stage % LT_NSTAGE, use = n_stage / LT_NSTAGE;
                    const int j0 = 8 * c + 4 * half;
                    float xv[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) xv[e] = (j0 + e < cnt && !LT_DBG(2)) ? xs[(j0 + e) * 128] : 0.f;
                    if (c == nst - 1) {   // last read of this x tile: hand the ring slot back to the TMA warp
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(x_empty + xsl);
                    }
                    LTP_LAP(2);   // x from the ring
                    mbar_wait_sticky(s_empty + slot, (use & 1) ^ 1, abort_flag);
                    fence_after_sync();
                    LTP_LAP(3);   // wait for the stage
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + A_COL0 + slot * 64 + half * 32;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (j0 + e < cnt && !LT_DBG(4)) {
                            float x = xv[e];
                            const bool ok = inb && (x == x);
                            x = ok ? x : 0.f;
                            const uint32_t v = ok ? 0x3F800000u : 0u;
                            uint32_t w[8];
                            tf32_split(x * x, w[0], w[2]);
                            tf32_split(x, w[3], w[5]);
                            w[1] = w[0]; w[4] = w[3]; w[6] = v; w[7] = v;
                            tmem_st8(taddr + e * 8, w);
                        }
                    }
                    tmem_wait_st();
                    LTP_LAP(4);   // split + tcgen05.st
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_full + slot);
                    LTP_LAP(5);   // fence + arrive
                }
            }
        }
        if (warp == 4) LTP_FLUSH(0);
    } else if (warp == W_MMA) {
        // ===================================================== MMA issuer =====================================================
        constexpr uint32_t idesc = instr_desc_tf32(128, NP);
        const uint32_t ws_addr = smem_u32(Ws);
        LTP_DECL;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++n_item) {
            const int s = item / a.R;
            const int cnt = min(a.card, a.F - s * a.card);
            const uint32_t ib = n_item & 1;
            LTP_LAP(0);
            mbar_wait_sticky(w_full + ib, (n_item >> 1) & 1, abort_flag);
            LTP_LAP(1);   // wait for the image
            fence_after_sync();
            const uint32_t wbase = ws_addr + ib * img_stride;
            for (int t = 0; t < ntiles; ++t, ++n_tile) {
                const uint32_t ab = n_tile & 1;
                LTP_LAP(0);
                mbar_wait_sticky(acc_empty + ab, ((n_tile >> 1) & 1) ^ 1, abort_flag);
                LTP_LAP(2);   // wait for the accumulator buffer
                const uint32_t acc = tmem_base + ab * 64;
                for (int c = 0; c < nst; ++c, ++n_stage) {
                    const uint32_t slot = n_stage % LT_NSTAGE, use = n_stage / LT_NSTAGE;
                    const int nf = min(8, cnt - c * 8);
                    mbar_wait_sticky(s_full + slot, use & 1, abort_flag);
                    LTP_LAP(3);   // wait for the operand stage
                    fence_after_sync();
                    const uint32_t abase = tmem_base + A_COL0 + slot * 64;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        if (jj < nf) {
                            // A [M = row][K = 8] in tensor memory: lane = row, 8 columns
                            // B [N = channel][K = 8]: K halves NP * 16 B apart, 8-channel groups 128 B apart
                            const uint64_t bdesc = smem_desc(wbase + (c * 8 + jj) * (NP * 32), NP * 16, 128);
                            if (!LT_DBG(1)) mma_ts_tf32_uniform(acc, abase + jj * 8, bdesc, idesc, (c | jj) ? 1u : 0u);
                        }
                    }
                    mma_commit_uniform(s_empty + slot);
                    LTP_LAP(4);   // MMA issue + commit
                }
                mma_commit_uniform(acc_full + ab);
            }
            mma_commit_uniform(w_empty + ib);
        }
        LTP_FLUSH(8);
    } else if (warp == W_TMA) {
        // ===================================================== TMA: parameter images and x tiles =====================================================
        uint32_t n_x = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++n_item) {
            const uint32_t ib = n_item & 1;
            const int s = item / a.R, r = item % a.R;
            const int f_begin = s * a.card, cnt = min(a.card, a.F - f_begin);
            mbar_wait_sticky(w_empty + ib, ((n_item >> 1) & 1) ^ 1, abort_flag);
            if (lane == 0) {
                mbar_arrive_expect_tx(w_full + ib, a.img_bytes);
                const uint8_t* src = a.img + (size_t)item * a.img_bytes;
                constexpr int CH = 16384;
                for (int off = 0; off < a.img_bytes; off += CH)
                    bulk_g2s(Ws + ib * img_stride + off, src + off, min(CH, a.img_bytes - off), w_full + ib);
            }
            __syncwarp();
            int64_t mysrc = 0;   // lane j: offset of the xT row of the scope's j-th feature
            if (lane < cnt) mysrc = (int64_t)(a.perm ? a.perm[(int64_t)(f_begin + lane) * a.R + r] : f_begin + lane) * a.ldx;
            for (int t = 0; t < ntiles; ++t, ++n_x) {
                const uint32_t xsl = n_x % LT_XD;
                const int rows4 = min(128, a.ldx - t * 128);   // whole 16-byte groups (xT rows are zero padded to ldx)
                mbar_wait_sticky(x_empty + xsl, ((n_x / LT_XD) & 1) ^ 1, abort_flag);
                if (lane == 0) mbar_arrive_expect_tx(x_full + xsl, cnt * rows4 * 4);
                __syncwarp();
                // one bulk copy per feature, each issued by its own lane (a single-lane loop cost ~90 clk per copy)
                if (lane < cnt) bulk_g2s(Xs + xsl * (a.card * 128) + lane * 128, a.xT + mysrc + t * 128, rows4 * 4, x_full + xsl);
                __syncwarp();
            }
        }
    } else {
        // ===================================================== epilogue =====================================================
        const int I = a.I;
        LTP_DECL;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            for (int t = 0; t < ntiles; ++t, ++n_tile) {
                const uint32_t ab = n_tile & 1;
                LTP_LAP(0);
                mbar_wait_sticky(acc_full + ab, (n_tile >> 1) & 1, abort_flag);
                LTP_LAP(1);   // wait for the accumulator
                fence_after_sync();
                uint32_t d[NP];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + ab * 64;
#pragma unroll
                for (int q = 0; q < NP / 16; ++q) tmem_ld16(taddr + q * 16, d + q * 16);
                tmem_wait_ld();
                fence_before_sync();
                __syncwarp();
                LTP_LAP(2);   // tcgen05.ld
                if (lane == 0) {
                    mbar_arrive(acc_empty + ab);
                    bulk_wait_group_read<1>();   // this warp's staging block of tile n - 2 has been read by its bulk store
                }
                __syncwarp();
                LTP_LAP(3);   // release + wait for the staging block
                float* os = Os + (warp * 2 + ab) * 32 * I;
                float* o = os + lane * I;
#pragma unroll
                for (int i4 = 0; i4 < NP / 4; ++i4)
                    if (4 * i4 < I)
                        *reinterpret_cast<uint4*>(o + 4 * i4) = make_uint4(d[4 * i4], d[4 * i4 + 1], d[4 * i4 + 2], d[4 * i4 + 3]);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    const int r0 = t * 128 + warp * 32;
                    const int rows = min(32, a.B - r0);
                    if (rows > 0 && !LT_DBG(8))
                        bulk_s2g(a.out + ((int64_t)item * a.B + r0) * I, os, (uint32_t)(rows * I * 4));
                    bulk_commit_group();
                }
                LTP_LAP(4);   // staging + store issue
            }
        }
        if (lane == 0) bulk_wait_group<0>();
        if (warp == 0) LTP_FLUSH(16);
    }
    // teardown
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == W_MMA) tmem_dealloc(tmem_base, 512);
    if (threadIdx.x == 0) tc_abort_trap(abort_flag, &g_leaf_tc_aborts);
}

// --------------------------------------------------------------------------------------------------------------------
// Backward (parameter gradients).  Work item = (scope, repetition); the accumulator D[M = 5 card][N = 2 I] stays in tensor
// memory over all rows of the batch (K), in stages of 64 rows.  22 warps:
//   warps 0-3   epilogue (once per item): accumulator -> shared memory -> dmu / dsigma;
//   warps 4-19  operand producers.  A [M = 5 jj + term][K = row] (terms x2h x2l xh xl valid): thread = (feature jj, K group of
//               4 rows): one 16-byte load from xT (one stage ahead), five 16-byte stores; the 8 lanes of a quarter warp hold
//               8 consecutive features, whose operand rows 5 jj + term fall into 8 different 16-byte bank groups.
//               B [N = channel (hi) | I + channel (lo)][K = row]: thread = (channel, K group) from the TMA-landed g tile;
//   warp 20     MMA issuer (8 K = 8 steps per stage);   warp 21   TMA of the g tiles (64 * I floats as four padded 16-row
//               blocks, ring of up to 6 stages).
// --------------------------------------------------------------------------------------------------------------------
constexpr int LB_A_BYTES = 128 * LT_BROWS * 4;   // 32 KB: [16 row groups][16 K groups][8][4]
constexpr int LB_GRING = 6;   // at most; the launch picks the depth that fits (a.gring)
constexpr uint32_t LB_ITRING = 16;   // work-item ring of the backward
__host__ __device__ constexpr int lb_gslot_floats(int I) { return 4 * (16 * I + 8); }   // four 16-row blocks, 32 bytes of padding each
__host__ __device__ constexpr int lb_b_bytes(int I) { return 2 * I * LT_BROWS * 4; }

// long wait (a whole work item): sleep between probes instead of spinning next to the producer warps
__device__ __forceinline__ void lt_wait_sleepy(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
    for (int i = 0; i < (1 << 22); ++i) {
        if (mbar_try_wait(bar, parity) || *abort_flag) return;
        __nanosleep(256);
    }
    *abort_flag = 1;
}

__global__ void __launch_bounds__(704, 1) leaf_tc_bwd_kernel(LeafTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int I = a.I, NG = I / 8;
    const int b_bytes = lb_b_bytes(I);
    const int graw_bytes = lb_gslot_floats(I) * 4;
    const uint32_t gring = a.gring;
    uint8_t* As = smem;                                                   // [2][LB_A_BYTES]
    uint8_t* Bs = As + LT_BSTAGES * LB_A_BYTES;                           // [2][b_bytes]
    uint8_t* Gr = Bs + LT_BSTAGES * b_bytes;                              // [LB_GRING][graw_bytes]
    float* red = reinterpret_cast<float*>(Gr + gring * graw_bytes);       // [128][I + 1]
    uint64_t* s_full = reinterpret_cast<uint64_t*>(red + 128 * (I + 1));   // 512 (I + 1) bytes: 8-byte aligned
    uint64_t* s_empty = s_full + 2;
    uint64_t* g_full = s_empty + 2;             // [LB_GRING]
    uint64_t* g_empty = g_full + LB_GRING;      // [LB_GRING]
    uint64_t* acc_full = g_empty + LB_GRING;    // [2]
    uint64_t* acc_empty = acc_full + 2;         // [2]
    uint64_t* it_full = acc_empty + 2;          // [LB_ITRING] work-item ring: fetched by the TMA warp, read by every warp
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(it_full + LB_ITRING);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
    volatile int* s_item = abort_flag + 1;      // [LB_ITRING]

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    constexpr int W_MMA = 20, W_TMA = 21;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(s_full + i, 16);
            mbar_init(s_empty + i, 1);
            mbar_init(acc_full + i, 1);
            mbar_init(acc_empty + i, 4);
        }
        for (int i = 0; i < LB_ITRING; ++i) mbar_init(it_full + i, 1);
        for (int i = 0; i < LB_GRING; ++i) {
            mbar_init(g_full + i, 1);
            mbar_init(g_empty + i, 16);
        }
        *abort_flag = 0;
        mbar_fence_init();
    }
    // operand rows 125 .. 127 (and, for short scopes, the rows of absent features) are zero: the former are never written
    for (int i = threadIdx.x; i < LT_BSTAGES * LB_A_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(As)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == W_MMA) tmem_alloc(tmem_slot, 256);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int nitems = a.d0 * a.R;
    const int nstages = (a.B + LT_BROWS - 1) / LT_BROWS;
    uint32_t n_stage = 0, n_item = 0;
    // Work items come from an atomic counter (the gradient all-reduce of the layer above runs next to this kernel at N > 1 and
    // holds SMs: a static share would leave the late CTAs' items for a second wave).  The TMA warp fetches item k + 1 while it
    // works on item k and publishes it through a ring of LB_ITRING slots.  A slot is never overwritten before its last reader:
    // the MMA warp starts item e + 2 only after the epilogue released the accumulator of item e, the producers are at most two
    // stages ahead of the MMA warp and the TMA warp at most LB_GRING stages ahead of the producers -- with >= 1 stage per item
    // the fetcher is fewer than 12 items ahead of the slowest warp (a ring of 4 deadlocked at 4 stages per item).
    auto item_at = [&](uint32_t k) -> int {
        mbar_wait_sticky(it_full + (k % LB_ITRING), (k / LB_ITRING) & 1, abort_flag);
        return s_item[k % LB_ITRING];
    };

    if (warp >= 4 && warp < 20) {
        // ===================================================== operand producers =====================================================
        const int pw = warp - 4;
        const int l8 = lane & 7, l4 = lane >> 3;
        // A: this thread's feature and its K group of a stage
        const int jj = (pw & 3) * 8 + l8;
        const int qa = (pw >> 2) * 4 + l4;
        uint32_t aoff[5];
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const int m = 5 * jj + t;
            aoff[t] = (m >> 3) * 2048 + (m & 7) * 16 + qa * 128;
        }
        const bool arow = jj < LT_MAXCARD;
        for (uint32_t k = 0;; ++k) {
            const int item = item_at(k);
            if (item >= nitems) break;
            const int s = item / a.R, r = item % a.R;
            const int f_begin = s * a.card, cnt = min(a.card, a.F - f_begin);
            const bool have = jj < cnt;
            const float* xrow = a.xT;
            if (have) xrow += (int64_t)(a.perm ? a.perm[(int64_t)(f_begin + jj) * a.R + r] : f_begin + jj) * a.ldx;
            auto xload = [&](int st) {
                const int bq = st * LT_BROWS + 4 * qa;
                return (have && st < nstages && bq < a.ldx && !LT_DBG(2)) ? *reinterpret_cast<const float4*>(xrow + bq)
                                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            float4 xn = xload(0);
            LTP_DECL;
            for (int st = 0; st < nstages; ++st, ++n_stage) {
                const uint32_t slot = n_stage & 1, gslot = n_stage % gring;
                const int nrow = min(LT_BROWS, a.B - st * LT_BROWS);
                const float4 xc = xn;
                xn = xload(st + 1);   // one stage ahead
                LTP_LAP(0);
                mbar_wait_sticky(s_empty + slot, ((n_stage >> 1) & 1) ^ 1, abort_flag);
                LTP_LAP(1);   // wait for the operand stage
                if (arow && !LT_DBG(4)) {
                    uint8_t* ast = As + slot * LB_A_BYTES;
                    const float xs[4] = {xc.x, xc.y, xc.z, xc.w};
                    uint32_t w[5][4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const bool ok = have && xs[e] == xs[e];
                        const float x = ok ? xs[e] : 0.f;
                        tf32_split(x * x, w[0][e], w[1][e]);
                        tf32_split(x, w[2][e], w[3][e]);
                        w[4][e] = ok ? 0x3F800000u : 0u;
                    }
#pragma unroll
                    for (int t = 0; t < 5; ++t) *reinterpret_cast<uint4*>(ast + aoff[t]) = make_uint4(w[t][0], w[t][1], w[t][2], w[t][3]);
                }
                LTP_LAP(2);   // A operand
                // g tile -> B operand: thread = (channel, K group): 4 rows in, hi and lo words out (16-byte stores)
                mbar_wait_sticky(g_full + gslot, (n_stage / gring) & 1, abort_flag);
                LTP_LAP(3);   // wait for the g tile
                const float* gr = reinterpret_cast<const float*>(Gr + gslot * graw_bytes);
                uint8_t* bst = Bs + slot * b_bytes;
                for (int p = pw; p < NG * 4 && !LT_DBG(16); p += 16) {
                    // K group q = 4 l4 + (p & 3): the quarter warps read different 16-row blocks, which the padding puts 8 banks apart
                    const int ig = p >> 2, ql = p & 3, q = l4 * 4 + ql;
                    const float* gp = gr + l4 * (16 * I + 8) + (4 * ql) * I + ig * 8 + l8;
                    uint32_t gh[4], gl[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) tf32_split(4 * q + u < nrow ? gp[u * I] : 0.f, gh[u], gl[u]);
                    uint8_t* dst = bst + ig * 2048 + q * 128 + l8 * 16;
                    *reinterpret_cast<uint4*>(dst) = make_uint4(gh[0], gh[1], gh[2], gh[3]);
                    *reinterpret_cast<uint4*>(dst + NG * 2048) = make_uint4(gl[0], gl[1], gl[2], gl[3]);
                }
                LTP_LAP(4);   // B operand
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(g_empty + gslot);
                    mbar_arrive(s_full + slot);
                }
                LTP_LAP(5);   // fence + arrive
            }
            if (warp == 4) LTP_FLUSH(24);
        }
    } else if (warp == W_MMA) {
        // ===================================================== MMA issuer =====================================================
        const uint32_t idesc = instr_desc_tf32(128, 2 * I);
        const uint32_t as_addr = smem_u32(As), bs_addr = smem_u32(Bs);
        for (;; ++n_item) {
            if (item_at(n_item) >= nitems) break;
            const uint32_t ab = n_item & 1;
            mbar_wait_sticky(acc_empty + ab, ((n_item >> 1) & 1) ^ 1, abort_flag);
            const uint32_t acc = tmem_base + ab * 128;
            LTP_DECL;
            for (int st = 0; st < nstages; ++st, ++n_stage) {
                const uint32_t slot = n_stage & 1;
                LTP_LAP(0);
                mbar_wait_sticky(s_full + slot, (n_stage >> 1) & 1, abort_flag);
                LTP_LAP(1);   // wait for the operands
                fence_after_sync();
#pragma unroll
                for (int kk = 0; kk < LT_BROWS / 8; ++kk) {
                    // A [M][K = rows] and B [N][K = rows], K-major: K halves 128 B apart, 8-row groups 2048 B apart
                    const uint64_t adesc = smem_desc(as_addr + slot * LB_A_BYTES + 2 * kk * 128, 128, 2048);
                    const uint64_t bdesc = smem_desc(bs_addr + slot * b_bytes + 2 * kk * 128, 128, 2048);
                    if (!LT_DBG(1)) mma_ss_tf32_uniform(acc, adesc, bdesc, idesc, (st | kk) ? 1u : 0u);
                }
                mma_commit_uniform(s_empty + slot);
                LTP_LAP(2);   // MMA issue
            }
            mma_commit_uniform(acc_full + ab);
            LTP_FLUSH(32);
        }
    } else if (warp == W_TMA) {
        // ===================================================== g tile TMA =====================================================
        if (lane == 0) {
            s_item[0] = (int)blockIdx.x < nitems ? (int)blockIdx.x : nitems;
            mbar_arrive(it_full);
        }
        __syncwarp();
        for (uint32_t k = 0;; ++k) {
            const int item = item_at(k);
            if (lane == 0) {   // publish item k + 1 before the long work on item k
                s_item[(k + 1) % LB_ITRING] = item < nitems ? (int)min((unsigned int)nitems, gridDim.x + atomicAdd(a.counter, 1u)) : nitems;
                mbar_arrive(it_full + ((k + 1) % LB_ITRING));
            }
            __syncwarp();
            if (item >= nitems) break;
            const float* gsrc = a.g + (int64_t)item * a.B * I;
            for (int st = 0; st < nstages; ++st, ++n_stage) {
                const uint32_t gslot = n_stage % gring;
                const int b0 = st * LT_BROWS;
                const int nrow = min(LT_BROWS, a.B - b0);
                mbar_wait_sticky(g_empty + gslot, ((n_stage / gring) & 1) ^ 1, abort_flag);
                if (lane == 0) {
                    mbar_arrive_expect_tx(g_full + gslot, nrow * I * 4);
                    float* dst = reinterpret_cast<float*>(Gr + gslot * graw_bytes);
                    for (int blk = 0; blk < 4 && 16 * blk < nrow; ++blk)   // four 16-row blocks, padded apart (bank spread of the B build)
                        bulk_g2s(dst + blk * (16 * I + 8), gsrc + (int64_t)(b0 + 16 * blk) * I, min(16, nrow - 16 * blk) * I * 4,
                                 g_full + gslot);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================================================== epilogue =====================================================
        const int m = warp * 32 + lane;
        const int ld = I + 1;
        for (;; ++n_item) {
            const int item = item_at(n_item);
            if (item >= nitems) break;
            const int s = item / a.R, r = item % a.R;
            const int f_begin = s * a.card, cnt = min(a.card, a.F - f_begin);
            const uint32_t ab = n_item & 1;
            lt_wait_sleepy(acc_full + ab, (n_item >> 1) & 1, abort_flag);
            __syncwarp();
            fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + ab * 128;
            const bool with_lo = (m % 5) % 2 == 0;   // x2h, xh and valid rows also take the g-lo columns
            for (int i0 = 0; i0 < I; i0 += 8) {
                uint32_t dh[8], dl[8];
                tmem_ld8(taddr + i0, dh);
                tmem_ld8(taddr + I + i0, dl);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    red[m * ld + i0 + e] = __uint_as_float(dh[e]) + (with_lo ? __uint_as_float(dl[e]) : 0.f);
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + ab);
            lt_named_bar(1, 128);
            for (int e = threadIdx.x; e < cnt * I; e += 128) {
                const int fj = e / I, i = e % I;
                const float* rp = red + (5 * fj) * ld + i;
                const float sxx = rp[0] + rp[ld], sx = rp[2 * ld] + rp[3 * ld], s0 = rp[4 * ld];
                const int64_t p = (int64_t)(f_begin + fj) * a.sp_f + (int64_t)i * a.sp_i + (int64_t)r * a.sp_r;
                const float mu = a.mu[p], inv_sg = 1.f / a.sigma[p];
                const float inv_var = inv_sg * inv_sg;
                const float s1 = sx - mu * s0;                          // sum g (x - mu)
                const float s2 = sxx - 2.f * mu * sx + mu * mu * s0;    // sum g (x - mu)^2
                a.dmu[p] = s1 * inv_var;
                a.dsigma[p] = s2 * inv_var * inv_sg - s0 * inv_sg;
            }
            lt_named_bar(1, 128);
        }
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == W_MMA) tmem_dealloc(tmem_base, 256);
    if (threadIdx.x == 0) tc_abort_trap(abort_flag, &g_leaf_tc_aborts);
}

int lt_np(int I) { return (I + 15) / 16 * 16; }
int lt_fwd_smem(int I, int card) {
    const int img_stride = (card * lt_np(I) * 32 + 127) & ~127;
    return 2 * img_stride + 2 * 128 * I * 4 + LT_XD * card * 512 + 512;
}
int lt_bwd_smem(int I, int gring) {
    return LT_BSTAGES * (LB_A_BYTES + lb_b_bytes(I)) + gring * lb_gslot_floats(I) * 4 + 128 * (I + 1) * 4 + 8 + 512;
}
int lt_gring(int I) {   // deepest g ring that fits
    int g = LB_GRING;
    while (g > 2 && lt_bwd_smem(I, g) > 232448) --g;
    return g;
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" int spn_leaf_tc_supported(int I, int R, int card, int d0) {
    if (I < 8 || I > 64 || I % 8 != 0 || card < 1 || card > LT_MAXCARD || R < 1 || d0 < 1) return 0;
    if ((int64_t)R * d0 > (1 << 24)) return 0;
    return (lt_fwd_smem(I, card) <= 232448 && lt_bwd_smem(I, lt_gring(I)) <= 232448) ? 1 : 0;
}

extern "C" int64_t spn_leaf_tc_image_bytes(int I, int R, int card, int d0) {
    return (int64_t)d0 * R * card * lt_np(I) * 32 + 256;
}

#ifdef SPN_LEAF_TC_PROF
extern "C" int spn_debug_leaf_tc_prof(long long* out_host, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out_host, g_leaf_tc_prof, sizeof(long long) * 64);
    if (e == cudaSuccess && reset) {
        long long z[64] = {0};
        e = cudaMemcpyToSymbol(g_leaf_tc_prof, z, sizeof(z));
    }
    return spn_cuda_status(e);
}
#endif
static int g_leaf_tc_dbg = 0;
#ifdef SPN_LEAF_TC_PROF
extern "C" int spn_debug_leaf_tc_mode(int mode) { g_leaf_tc_dbg = mode; return 0; }
#endif

static void lt_fill(LeafTcArgs& a, const float* xT, const int32_t* perm, const float* mu, const float* sigma, int64_t sp_f,
                    int64_t sp_i, int64_t sp_r, int B, int F, int I, int R, int card, int d0) {
    a.xT = xT; a.perm = perm; a.ldx = (B + 3) / 4 * 4;
    a.mu = mu; a.sigma = sigma; a.sp_f = sp_f; a.sp_i = sp_i; a.sp_r = sp_r;
    a.B = B; a.F = F; a.I = I; a.R = R; a.card = card; a.d0 = d0;
    a.nst = (card + 7) / 8;
    a.img_bytes = card * lt_np(I) * 32;
    a.dbg = g_leaf_tc_dbg;
}

/* Forward: builds the parameter operand image in `img` (spn_leaf_tc_image_bytes, 128-byte aligned) and evaluates
 * out[d0][R][B][I] (contiguous) from the feature-major copy xT of spn_leaf_transpose. */
extern "C" int spn_leaf_tc_fwd(const float* xT, const int32_t* perm, const float* mu, const float* sigma, int64_t sp_f,
                               int64_t sp_i, int64_t sp_r, int B, int F, int I, int R, int card, int d0, void* img, float* out,
                               void* stream) {
    if (!xT || !mu || !sigma || !img || !out || B < 0 || F <= 0 || (int64_t)card * d0 < F || (int64_t)card * (d0 - 1) >= F)
        return SPN_ERR_ARG;   // every scope holds at least one feature
    if (!spn_leaf_tc_supported(I, R, card, d0) || !aligned16(xT) || !aligned16(out) || (reinterpret_cast<uintptr_t>(img) & 127) ||
        (int64_t)F * ((B + 3) / 4 * 4) >= (1ll << 31))   // 32-bit row offsets into xT
        return SPN_ERR_UNSUPPORTED;
    if (B == 0) return SPN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    LeafTcArgs a{};
    lt_fill(a, xT, perm, mu, sigma, sp_f, sp_i, sp_r, B, F, I, R, card, d0);
    a.img = static_cast<const uint8_t*>(img);
    a.out = out;
    const int NP = lt_np(I);
    leaf_tc_image_kernel<<<d0 * R, 256, 0, st>>>(a, NP);
    int rc = spn_launch_status();
    if (rc != SPN_OK) return rc;
    const int smem = lt_fwd_smem(I, card);
    auto kernel = NP == 16 ? leaf_tc_fwd_kernel<16> : NP == 32 ? leaf_tc_fwd_kernel<32> : NP == 48 ? leaf_tc_fwd_kernel<48>
                                                                                                     : leaf_tc_fwd_kernel<64>;
    static std::atomic<unsigned long long> configured[4];
    int dev = 0, sms = 148;
    if (!tc_configured_on_device(configured[NP / 16 - 1], &dev)) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return spn_cuda_status(e);
        if (dev < 64) configured[NP / 16 - 1].fetch_or(1ull << dev);
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nitems = d0 * R;
    kernel<<<nitems < sms ? nitems : sms, 448, smem, st>>>(a);
    return spn_launch_status();
}

/* Parameter gradients from g[d0][R][B][I] (contiguous): dmu / dsigma in the layout of mu / sigma (strides sp_*). */
extern "C" int spn_leaf_tc_bwd_params(const float* xT, const float* g, const int32_t* perm, const float* mu, const float* sigma,
                                      int64_t sp_f, int64_t sp_i, int64_t sp_r, int B, int F, int I, int R, int card, int d0,
                                      float* dmu, float* dsigma, void* stream) {
    if (!xT || !g || !mu || !sigma || !dmu || !dsigma || B <= 0 || F <= 0 || (int64_t)card * d0 < F ||
        (int64_t)card * (d0 - 1) >= F)
        return SPN_ERR_ARG;
    if (!spn_leaf_tc_supported(I, R, card, d0) || !aligned16(xT) || !aligned16(g)) return SPN_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    LeafTcArgs a{};
    lt_fill(a, xT, perm, mu, sigma, sp_f, sp_i, sp_r, B, F, I, R, card, d0);
    a.g = g; a.dmu = dmu; a.dsigma = dsigma;
    a.gring = lt_gring(I);
    static std::atomic<unsigned long long> configured{0};
    int dev = 0, sms = 148;
    if (!tc_configured_on_device(configured, &dev)) {
        cudaError_t e = cudaFuncSetAttribute(leaf_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return spn_cuda_status(e);
        if (dev < 64) configured.fetch_or(1ull << dev);
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nitems = d0 * R;
    a.counter = tc_next_counter(st);
    if (!a.counter) return SPN_ERR_ARG;
    leaf_tc_bwd_kernel<<<nitems < sms ? nitems : sms, 704, lt_bwd_smem(I, a.gring), st>>>(a);
    return spn_launch_status();
}
