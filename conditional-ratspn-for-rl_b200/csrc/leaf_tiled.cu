// Register-tiled leaf kernels for SHARED parameters (RatSpn) and plain inputs x [B, F] (no channel / repetition dims).
//
// The leaf layer evaluates N*F*I*R Gaussian log-densities (cfg 2: 5.1e9) from only N*F inputs and 2*F*I*R parameters,
// so it is bound by the fp32 pipes, not by HBM -- provided operands are re-used from shared memory / registers:
//   * x is first transposed to feature-major xT [F][ldx] (spn_leaf_transpose; ldx = B rounded up to 4, zero padded), so
//     that the per-repetition feature permutation x[b, perm[f, r]] becomes a gather of whole contiguous row segments
//     instead of one 32-byte sector per value; the same pass records which 64-row groups contain NaN inputs;
//   * forward: block = (scope s, repetition r) walking over 128-row tiles; the features of the scope are staged in
//     chunks of 32 as xs[feature][row] by cp.async (double buffered: the next tile streams in while the current one is
//     evaluated) next to mu, 1/(2 sigma^2) and log sigma per (feature, channel); each thread owns a 4-row x 4-channel
//     register tile: 3 shared-memory vector loads feed 16 evaluations (sub, mul, fma) per feature;
//   * backward (parameters): block = (scope s, repetition r, feature chunk); each thread owns a 4-feature x 4-channel
//     tile of (sum g*d, sum g*d*d) accumulators and walks the rows in double-buffered cp.async tiles of 64; row
//     sub-groups of the block are combined through shared memory, so every gradient element is written once (no
//     atomics unless the grid has to be split over rows to fill the machine).
// NaN inputs (marginalised features, distributions.py:58-61, :243) take a block-uniform slow path with a validity mask.
// Reference semantics: rat_spn.py:155-170, distributions.py:199-246, :366-382, layers.py:276-312.
#include "common.cuh"

namespace {

constexpr int FC = 32;      // features per staged chunk
constexpr int FROWS = 128;  // rows per forward tile
constexpr int NANROWS = 64; // granularity of the NaN flags written by the transpose pass
constexpr int BROWS = 128;  // rows per backward tile

struct LeafTiledArgs {
    const float* xT;  // [F][ldx]
    const int* nanflag;  // [ceil(B / 64)]
    const int32_t* perm;
    const float* mu; const float* sigma; int64_t sp_f, sp_i, sp_r;
    int B, ldx, F, I, R, card, d0, tanh_corr;
    float* out; const float* g; int64_t so_b, so_d, so_c, so_r;
    float* dmu; float* dsigma;
    int nsplit, nchunk, fchunk, vec_io, ngroups;
};

__host__ __device__ inline int leaf_ldx(int B) { return (B + 3) / 4 * 4; }

// Packed fp32 arithmetic (two lanes of a 64-bit register pair per instruction, sm_100): the scalar sub / mul / fma chain of
// the leaf kernels reads three distinct registers per instruction and is limited by register-file bandwidth (1.6 clk per
// instruction measured, profiles/r02_pipe_bench.txt); the packed forms issue 2 FMAs per 2 clk with half the issue slots.
__device__ __forceinline__ uint64_t lf_pack(float lo, float hi) {
    uint64_t v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
    return v;
}
__device__ __forceinline__ void lf_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t lf_add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t lf_mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t lf_fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__global__ void __launch_bounds__(256) leaf_transpose_kernel(const float* __restrict__ x, int64_t sx_b, int64_t sx_f, int B,
                                                            int ldx, int F, float* __restrict__ xT, int* __restrict__ nanflag) {
    __shared__ float tile[32][33];
    const int b0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    int has_nan = 0;
    for (int j = ty; j < 32; j += 8) {
        const int b = b0 + j, f = f0 + tx;
        const float v = (b < B && f < F) ? x[(int64_t)b * sx_b + (int64_t)f * sx_f] : 0.f;
        if (isnan(v)) has_nan = 1;
        tile[j][tx] = v;
    }
    has_nan = __syncthreads_or(has_nan);
    if (has_nan && threadIdx.x == 0) atomicOr(nanflag + b0 / NANROWS, 1);
    for (int j = ty; j < 32; j += 8) {
        const int f = f0 + j, b = b0 + tx;
        if (f < F && b < ldx) xT[(int64_t)f * ldx + b] = tile[tx][j];
    }
}

// grid (row-tile groups, R, d0); block 32 * ceil(I / 4) threads: lane = row group (4 rows), warp = channel group (4 ch).
template <int MAXT>
__global__ void __launch_bounds__(MAXT) leaf_tiled_fwd_kernel(LeafTiledArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int I4 = (int)(blockDim.x >> 5) * 4;
    float* xs = sm;                       // [2][FC][FROWS]
    float* pm = sm + 2 * FC * FROWS;      // [FC][I4] -mu
    float* pa = pm + FC * I4;             // [FC][I4] -1 / (2 sigma^2)
    float* pk = pa + FC * I4;             // [FC][I4] log sigma + log sqrt(2 pi)
    float* ks = pk + FC * I4;             // [I4]     sum of pk over the chunk
    const int s = blockIdx.z, r = blockIdx.y;
    const int rg = threadIdx.x & 31, cg = threadIdx.x >> 5;
    const int f_begin = s * a.card, f_end = min(a.F, f_begin + a.card);
    const int nchunks = (f_end - f_begin + FC - 1) / FC;
    const int ntiles = (a.B + FROWS - 1) / FROWS;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nstages = my_tiles * nchunks;
    const int pi = threadIdx.x % I4, pj0 = threadIdx.x / I4;  // parameter staging: blockDim = 8 * I4

    auto issue = [&](int k) {
        const int tile = blockIdx.x + (k / nchunks) * gridDim.x, ch = k % nchunks;
        const int fc0 = f_begin + ch * FC, cnt = min(FC, f_end - fc0);
        const int b0 = tile * FROWS;
        float* dst = xs + (k & 1) * FC * FROWS;
        for (int c = threadIdx.x; c < cnt * 32; c += blockDim.x) {
            const int j = c >> 5, q = c & 31;
            const int f = fc0 + j;
            const int src = a.perm ? a.perm[(int64_t)f * a.R + r] : f;
            const int b = b0 + 4 * q;
            const int bytes = max(0, min(16, (a.ldx - b) * 4));
            spn_cp_async16(dst + j * FROWS + 4 * q, a.xT + (int64_t)src * a.ldx + (bytes ? b : 0), bytes);
        }
        spn_cp_async_commit();
    };

    uint64_t acc2[4][2];   // [row u][channel pair]: (acc[u][2 vp], acc[u][2 vp + 1])
#pragma unroll
    for (int u = 0; u < 4; ++u) acc2[u][0] = acc2[u][1] = 0ull;

    if (nstages > 0) issue(0);
    for (int k = 0; k < nstages; ++k) {
        const int tile = blockIdx.x + (k / nchunks) * gridDim.x, ch = k % nchunks;
        const int fc0 = f_begin + ch * FC, cnt = min(FC, f_end - fc0);
        const int b0 = tile * FROWS;
        if (k + 1 < nstages) issue(k + 1);
        if (nchunks > 1 || k == 0) {
            for (int j = pj0; j < cnt; j += 8) {
                float m = 0.f, av = 0.f, kv = 0.f;
                if (pi < a.I) {
                    const int64_t p = (int64_t)(fc0 + j) * a.sp_f + (int64_t)pi * a.sp_i + (int64_t)r * a.sp_r;
                    m = a.mu[p];
                    const float sg = a.sigma[p];
                    av = 1.f / (2.f * sg * sg);
                    kv = logf(sg) + SPN_LOG_SQRT_2PI;
                }
                pm[j * I4 + pi] = -m; pa[j * I4 + pi] = -av; pk[j * I4 + pi] = kv;   // negated: d = x + (-mu), acc += (d * -a) * d
            }
            __syncthreads();
            if (threadIdx.x < I4) {
                float t = 0.f;
                for (int j = 0; j < cnt; ++j) t += pk[j * I4 + threadIdx.x];
                ks[threadIdx.x] = t;
            }
        }
        if (k + 1 < nstages) spn_cp_async_wait<1>(); else spn_cp_async_wait<0>();
        __syncthreads();
        const float* xb = xs + (k & 1) * FC * FROWS;
        int has_nan = a.nanflag[b0 / NANROWS];
        if (b0 + NANROWS < a.B) has_nan |= a.nanflag[b0 / NANROWS + 1];
        if (!has_nan && !a.tanh_corr) {
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const float4 x4 = *reinterpret_cast<const float4*>(xb + j * FROWS + 4 * rg);
                const float4 m4 = *reinterpret_cast<const float4*>(pm + j * I4 + 4 * cg);
                const float4 a4 = *reinterpret_cast<const float4*>(pa + j * I4 + 4 * cg);
                const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
                const uint64_t nm2[2] = {lf_pack(m4.x, m4.y), lf_pack(m4.z, m4.w)};
                const uint64_t na2[2] = {lf_pack(a4.x, a4.y), lf_pack(a4.z, a4.w)};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint64_t x2 = lf_pack(xv[u], xv[u]);
#pragma unroll
                    for (int vp = 0; vp < 2; ++vp) {
                        const uint64_t d2 = lf_add2(x2, nm2[vp]);                        // x - mu
                        acc2[u][vp] = lf_fma2(lf_mul2(d2, na2[vp]), d2, acc2[u][vp]);   // -= (x - mu)^2 / (2 sigma^2)
                    }
                }
            }
            const float4 k4 = *reinterpret_cast<const float4*>(ks + 4 * cg);
            const uint64_t nk2[2] = {lf_pack(-k4.x, -k4.y), lf_pack(-k4.z, -k4.w)};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int vp = 0; vp < 2; ++vp) acc2[u][vp] = lf_add2(acc2[u][vp], nk2[vp]);
        } else {
            // slow path: NaN -> contributes nothing; tanh change-of-variables correction per (row, feature)
            for (int j = 0; j < cnt; ++j) {
                const float4 x4 = *reinterpret_cast<const float4*>(xb + j * FROWS + 4 * rg);
                const float4 m4 = *reinterpret_cast<const float4*>(pm + j * I4 + 4 * cg);
                const float4 a4 = *reinterpret_cast<const float4*>(pa + j * I4 + 4 * cg);
                const float4 k4 = *reinterpret_cast<const float4*>(pk + j * I4 + 4 * cg);
                const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
                const float nmv[4] = {m4.x, m4.y, m4.z, m4.w};   // -mu
                const float nav[4] = {a4.x, a4.y, a4.z, a4.w};   // -1 / (2 sigma^2)
                const float kv[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (isnan(xv[u])) continue;
                    const float corr = a.tanh_corr ? spn_tanh_correction(xv[u]) : 0.f;
                    float t[4];
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float d = xv[u] + nmv[v];
                        t[v] = (d * d) * nav[v] - kv[v] - corr;
                    }
                    acc2[u][0] = lf_add2(acc2[u][0], lf_pack(t[0], t[1]));
                    acc2[u][1] = lf_add2(acc2[u][1], lf_pack(t[2], t[3]));
                }
            }
        }
        if (ch == nchunks - 1) {
            // every thread writes its 4 x 4 tile straight from registers
            float* obase = a.out + (int64_t)s * a.so_d + (int64_t)r * a.so_r;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t b = (int64_t)b0 + 4 * rg + u;
                float acc[4];
                lf_unpack(acc2[u][0], acc[0], acc[1]);
                lf_unpack(acc2[u][1], acc[2], acc[3]);
                if (b < a.B) {
                    if (a.vec_io) {
                        *reinterpret_cast<float4*>(obase + b * a.so_b + 4 * cg) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    } else {
#pragma unroll
                        for (int v = 0; v < 4; ++v)
                            if (4 * cg + v < a.I) obase[b * a.so_b + (int64_t)(4 * cg + v) * a.so_c] = acc[v];
                    }
                }
                acc2[u][0] = acc2[u][1] = 0ull;
            }
        }
        __syncthreads();  // stage buffers / parameters may be overwritten from here on
    }
}

// grid (R * d0, nchunk, nsplit); block = ngroups * nfg * nig threads, nfg = ceil(fchunk / 4), nig = ceil(I / 4).
__global__ void __launch_bounds__(256, 3) leaf_tiled_bwd_kernel(LeafTiledArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int nig = (a.I + 3) / 4, I4 = nig * 4;
    const int nfg = (a.fchunk + 3) / 4, F4 = nfg * 4;
    const int ntile = nfg * nig;
    const int ngroups = a.ngroups;  // blockDim is rounded up to whole warps: threads past ngroups * ntile only copy
    float* gs = sm;                        // [2][BROWS][I4]
    float* xs = gs + 2 * BROWS * I4;       // [2][BROWS][F4]
    float* crs = xs + 2 * BROWS * F4;      // [F4][I4] sum of g over the rows whose feature is NaN (slow path only)
    __shared__ int src_f[FC];
    const int s = blockIdx.x / a.R, r = blockIdx.x % a.R;
    const int f_begin = s * a.card + blockIdx.y * a.fchunk;
    const int f_end = min(min(a.F, (s + 1) * a.card), f_begin + a.fchunk);
    const int cnt = f_end - f_begin;
    if (cnt <= 0) return;
    const int tile = threadIdx.x % ntile, grp = threadIdx.x / ntile;
    const bool worker = grp < ngroups;
    const int ig = tile % nig, fg = tile / nig;

    if (threadIdx.x < F4) {
        const int f = f_begin + min((int)threadIdx.x, cnt - 1);
        src_f[threadIdx.x] = a.perm ? a.perm[(int64_t)f * a.R + r] : f;
    }
    for (int e = threadIdx.x; e < F4 * I4; e += blockDim.x) crs[e] = 0.f;
    float mv[4][4];  // mu of (feature u, channel v)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int f = f_begin + min(4 * fg + u, cnt - 1), i = min(4 * ig + v, a.I - 1);
            mv[u][v] = a.mu[(int64_t)f * a.sp_f + (int64_t)i * a.sp_i + (int64_t)r * a.sp_r];
        }
    // packed accumulators over channel pairs (see lf_fma2): a1 = sum g d, a2 = sum g d^2, a0 = sum g
    uint64_t nm2[4][2], a1p[4][2], a2p[4][2], a0p[2];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int vp = 0; vp < 2; ++vp) {
            nm2[u][vp] = lf_pack(-mv[u][2 * vp], -mv[u][2 * vp + 1]);
            a1p[u][vp] = 0ull;
            a2p[u][vp] = 0ull;
        }
    a0p[0] = a0p[1] = 0ull;
    const int64_t rows_per = (((int64_t)a.B + a.nsplit - 1) / a.nsplit + BROWS - 1) / BROWS * BROWS;
    const int64_t row_begin = (int64_t)blockIdx.z * rows_per;
    const int64_t row_end = min((int64_t)a.B, row_begin + rows_per);
    const int ntiles = row_end > row_begin ? (int)((row_end - row_begin + BROWS - 1) / BROWS) : 0;
    const float* gbase = a.g + (int64_t)s * a.so_d + (int64_t)r * a.so_r;
    __syncthreads();

    auto issue = [&](int t) {
        const int64_t b0 = row_begin + (int64_t)t * BROWS;
        const int nrow = (int)min((int64_t)BROWS, row_end - b0);
        float* gd = gs + (t & 1) * BROWS * I4;
        float* xd = xs + (t & 1) * BROWS * F4;
        if (a.vec_io) {  // the g tile is one contiguous run of nrow * I floats (I % 4 == 0)
            const float* src = gbase + b0 * a.so_b;
            const int nvalid = nrow * a.I / 4;
            for (int c = threadIdx.x; c < BROWS * I4 / 4; c += blockDim.x)
                spn_cp_async16(gd + 4 * c, src + (c < nvalid ? 4 * c : 0), c < nvalid ? 16 : 0);
        } else {
            for (int e = threadIdx.x; e < BROWS * I4; e += blockDim.x) {
                const int row = e / I4, i = e % I4;
                const bool ok = row < nrow && i < a.I;
                spn_cp_async4(gd + e, ok ? gbase + (b0 + row) * a.so_b + (int64_t)i * a.so_c : gbase, ok ? 4 : 0);
            }
        }
        // one warp per feature: the source row segment is contiguous, the destination is the transposed tile
        const int lane = threadIdx.x & 31, nwarp = (blockDim.x + 31) >> 5;
        for (int j = threadIdx.x >> 5; j < F4; j += nwarp) {
            const float* src = a.xT + (int64_t)src_f[j] * a.ldx + b0;
            const bool okj = j < cnt;
#pragma unroll
            for (int row = lane; row < BROWS; row += 32) {
                const bool ok = okj && row < nrow;
                spn_cp_async4(xd + row * F4 + j, ok ? src + row : a.xT, ok ? 4 : 0);
            }
        }
        spn_cp_async_commit();
    };

    if (ntiles > 0) issue(0);
    for (int t = 0; t < ntiles; ++t) {
        const int64_t b0 = row_begin + (int64_t)t * BROWS;
        const int nrow = (int)min((int64_t)BROWS, row_end - b0);
        if (t + 1 < ntiles) { issue(t + 1); spn_cp_async_wait<1>(); } else { spn_cp_async_wait<0>(); }
        __syncthreads();
        const float* gb = gs + (t & 1) * BROWS * I4;
        const float* xb = xs + (t & 1) * BROWS * F4;
        int has_nan = a.nanflag[b0 / NANROWS];
        if (b0 + NANROWS < a.B) has_nan |= a.nanflag[b0 / NANROWS + 1];
        if (!worker) {
        } else if (!has_nan) {
#pragma unroll 4
            for (int row = grp; row < nrow; row += ngroups) {
                const float4 g4 = *reinterpret_cast<const float4*>(gb + row * I4 + 4 * ig);
                const float4 x4 = *reinterpret_cast<const float4*>(xb + row * F4 + 4 * fg);
                const uint64_t g2[2] = {lf_pack(g4.x, g4.y), lf_pack(g4.z, g4.w)};
                const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
                a0p[0] = lf_add2(a0p[0], g2[0]);
                a0p[1] = lf_add2(a0p[1], g2[1]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint64_t x2 = lf_pack(xv[u], xv[u]);
#pragma unroll
                    for (int vp = 0; vp < 2; ++vp) {
                        const uint64_t d2 = lf_add2(x2, nm2[u][vp]);    // x - mu
                        const uint64_t t2 = lf_mul2(g2[vp], d2);
                        a1p[u][vp] = lf_add2(a1p[u][vp], t2);
                        a2p[u][vp] = lf_fma2(t2, d2, a2p[u][vp]);
                    }
                }
            }
        } else {
            for (int row = grp; row < nrow; row += ngroups) {
                const float4 g4 = *reinterpret_cast<const float4*>(gb + row * I4 + 4 * ig);
                const float4 x4 = *reinterpret_cast<const float4*>(xb + row * F4 + 4 * fg);
                const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
                const uint64_t g2[2] = {lf_pack(g4.x, g4.y), lf_pack(g4.z, g4.w)};
                const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
                a0p[0] = lf_add2(a0p[0], g2[0]);
                a0p[1] = lf_add2(a0p[1], g2[1]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (isnan(xv[u])) {
#pragma unroll
                        for (int v = 0; v < 4; ++v) atomicAdd(crs + (4 * fg + u) * I4 + 4 * ig + v, gv[v]);  // row does not count for feature u
                    } else {
                        const uint64_t x2 = lf_pack(xv[u], xv[u]);
#pragma unroll
                        for (int vp = 0; vp < 2; ++vp) {
                            const uint64_t d2 = lf_add2(x2, nm2[u][vp]);
                            const uint64_t t2 = lf_mul2(g2[vp], d2);
                            a1p[u][vp] = lf_add2(a1p[u][vp], t2);
                            a2p[u][vp] = lf_fma2(t2, d2, a2p[u][vp]);
                        }
                    }
                }
            }
        }
        __syncthreads();  // the buffer read here is refilled by the next iteration's issue
    }
    // combine the row groups through shared memory: red[grp][tile][36] (aliases the tile buffers, not crs)
    float* red = sm;
    if (worker) {
        float* my = red + ((int64_t)grp * ntile + tile) * 36;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int vp = 0; vp < 2; ++vp) {
                lf_unpack(a1p[u][vp], my[u * 4 + 2 * vp], my[u * 4 + 2 * vp + 1]);
                lf_unpack(a2p[u][vp], my[16 + u * 4 + 2 * vp], my[16 + u * 4 + 2 * vp + 1]);
            }
        lf_unpack(a0p[0], my[32], my[33]);
        lf_unpack(a0p[1], my[34], my[35]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ntile * 16; e += blockDim.x) {
        const int t = e / 16, uv = e % 16, u = uv / 4, v = uv % 4;
        const int fl = 4 * (t / nig) + u, i = 4 * (t % nig) + v;
        const int f = f_begin + fl;
        if (f >= f_end || i >= a.I) continue;
        float s1 = 0.f, s2 = 0.f, s0 = 0.f;
        for (int gq = 0; gq < ngroups; ++gq) {
            const float* src = red + ((int64_t)gq * ntile + t) * 36;
            s1 += src[uv]; s2 += src[16 + uv]; s0 += src[32 + v];
        }
        s0 -= crs[fl * I4 + i];
        const int64_t p = (int64_t)f * a.sp_f + (int64_t)i * a.sp_i + (int64_t)r * a.sp_r;
        const float inv_sg = 1.f / a.sigma[p];
        const float inv_var = inv_sg * inv_sg;
        const float dmu = s1 * inv_var;
        const float dsg = s2 * inv_var * inv_sg - s0 * inv_sg;
        if (a.nsplit == 1) {
            a.dmu[p] = dmu;
            a.dsigma[p] = dsg;
        } else {
            atomicAdd(a.dmu + p, dmu);
            atomicAdd(a.dsigma + p, dsg);
        }
    }
}

}  // namespace

extern "C" int64_t spn_leaf_workspace_floats(int B, int F) {
    return (int64_t)F * leaf_ldx(B) + (B + NANROWS - 1) / NANROWS + 4;
}

extern "C" int spn_leaf_transpose(const float* x, int64_t sx_b, int64_t sx_f, int B, int F, float* xT, void* stream) {
    if (!x || !xT || B < 0 || F <= 0) return SPN_ERR_ARG;
    if (B == 0) return SPN_OK;
    const int ldx = leaf_ldx(B);
    int* flags = reinterpret_cast<int*>(xT + (size_t)F * ldx);
    cudaError_t e = cudaMemsetAsync(flags, 0, ((B + NANROWS - 1) / NANROWS) * sizeof(int), (cudaStream_t)stream);
    if (e != cudaSuccess) return spn_cuda_status(e);
    leaf_transpose_kernel<<<dim3((ldx + 31) / 32, (F + 31) / 32), 256, 0, (cudaStream_t)stream>>>(x, sx_b, sx_f, B, ldx, F, xT, flags);
    return spn_launch_status();
}

extern "C" int spn_leaf_tiled_supported(int I, int R, int d0) { return (I >= 1 && I <= 128 && R <= 65535 && d0 <= 65535) ? 1 : 0; }

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int spn_leaf_tiled_fwd(const float* xT, const int32_t* perm, const float* mu, const float* sigma, int64_t sp_f,
                                  int64_t sp_i, int64_t sp_r, int B, int F, int I, int R, int card, int d0, int tanh_corr,
                                  float* out, int64_t so_b, int64_t so_d, int64_t so_c, int64_t so_r, void* stream) {
    if (!xT || !mu || !sigma || !out || B < 0 || F <= 0 || I <= 0 || R <= 0 || card <= 0 || d0 <= 0) return SPN_ERR_ARG;
    if (!spn_leaf_tiled_supported(I, R, d0) || !aligned16(xT)) return SPN_ERR_UNSUPPORTED;
    if (B == 0) return SPN_OK;
    LeafTiledArgs a{};
    a.xT = xT; a.ldx = leaf_ldx(B); a.nanflag = reinterpret_cast<const int*>(xT + (size_t)F * a.ldx);
    a.perm = perm; a.mu = mu; a.sigma = sigma; a.sp_f = sp_f; a.sp_i = sp_i; a.sp_r = sp_r;
    a.B = B; a.F = F; a.I = I; a.R = R; a.card = card; a.d0 = d0; a.tanh_corr = tanh_corr;
    a.out = out; a.so_b = so_b; a.so_d = so_d; a.so_c = so_c; a.so_r = so_r;
    a.vec_io = (so_c == 1 && I % 4 == 0 && so_b % 4 == 0 && so_d % 4 == 0 && so_r % 4 == 0 && aligned16(out)) ? 1 : 0;
    const int ncg = (I + 3) / 4, I4 = ncg * 4;
    const int smem = (2 * FC * FROWS + 3 * FC * I4 + I4) * (int)sizeof(float);
    // up to 16 channel groups (I <= 64): blocks of <= 512 threads may use 128 registers (deeper unrolling of the feature loop)
    auto kernel = 32 * ncg <= 512 ? leaf_tiled_fwd_kernel<512> : leaf_tiled_fwd_kernel<1024>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return spn_cuda_status(e);
    const int ntiles = (B + FROWS - 1) / FROWS;
    int gx = (16 * 148 + R * d0 - 1) / (R * d0);
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    kernel<<<dim3(gx, R, d0), 32 * ncg, smem, (cudaStream_t)stream>>>(a);
    return spn_launch_status();
}

extern "C" int spn_leaf_tiled_bwd_params(const float* xT, const float* g, int64_t sg_b, int64_t sg_d, int64_t sg_c,
                                         int64_t sg_r, const int32_t* perm, const float* mu, const float* sigma,
                                         int64_t sp_f, int64_t sp_i, int64_t sp_r, int B, int F, int I, int R, int card,
                                         int d0, float* dmu, float* dsigma, void* stream) {
    if (!xT || !g || !mu || !sigma || !dmu || !dsigma || B < 0 || F <= 0 || I <= 0 || R <= 0 || card <= 0 || d0 <= 0)
        return SPN_ERR_ARG;
    if (!spn_leaf_tiled_supported(I, R, d0)) return SPN_ERR_UNSUPPORTED;
    LeafTiledArgs a{};
    a.xT = xT; a.ldx = leaf_ldx(B); a.nanflag = reinterpret_cast<const int*>(xT + (size_t)F * a.ldx);
    a.perm = perm; a.mu = mu; a.sigma = sigma; a.sp_f = sp_f; a.sp_i = sp_i; a.sp_r = sp_r;
    a.B = B; a.F = F; a.I = I; a.R = R; a.card = card; a.d0 = d0;
    a.g = g; a.so_b = sg_b; a.so_d = sg_d; a.so_c = sg_c; a.so_r = sg_r; a.dmu = dmu; a.dsigma = dsigma;
    // contiguous g rows (internal layout): 16-byte cp.async of the whole tile
    a.vec_io = (sg_c == 1 && sg_b == I && I % 4 == 0 && sg_d % 4 == 0 && sg_r % 4 == 0 && aligned16(g)) ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int nig = (I + 3) / 4, I4 = nig * 4;
    // feature chunks per scope: at most FC features and at most 256 thread tiles per block
    int max_f = FC;
    while (max_f > 4 && ((max_f + 3) / 4) * nig > 256) max_f -= 4;
    a.nchunk = (card + max_f - 1) / max_f;
    a.fchunk = (card + a.nchunk - 1) / a.nchunk;
    const int nfg = (a.fchunk + 3) / 4, F4 = nfg * 4, ntile = nfg * nig;
    int ngroups = 256 / ntile;
    if (ngroups < 1) ngroups = 1;
    if (ngroups > 8) ngroups = 8;
    a.ngroups = ngroups;
    const int64_t blocks = (int64_t)R * d0 * a.nchunk;
    int nsplit = 1;
    if (blocks < 2 * 148) {
        nsplit = (int)((2 * 148 + blocks - 1) / blocks);
        const int max_split = (B + 4 * BROWS - 1) / (4 * BROWS);
        if (nsplit > max_split) nsplit = max_split;
        if (nsplit < 1) nsplit = 1;
    }
    a.nsplit = nsplit;
    const size_t nparam = (size_t)F * I * R;
    if (nsplit > 1 || B == 0) {
        cudaError_t e = cudaMemsetAsync(dmu, 0, nparam * sizeof(float), st);
        if (e != cudaSuccess) return spn_cuda_status(e);
        e = cudaMemsetAsync(dsigma, 0, nparam * sizeof(float), st);
        if (e != cudaSuccess) return spn_cuda_status(e);
    }
    if (B == 0) return SPN_OK;
    const int tile_f = 2 * BROWS * (I4 + F4);
    const int red_f = ngroups * ntile * 36;
    if (red_f > tile_f) return SPN_ERR_UNSUPPORTED;  // cannot happen: ngroups * ntile <= 256 and a tile holds >= 256 * 36 floats
    const int smem = (tile_f + F4 * I4) * (int)sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(leaf_tiled_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return spn_cuda_status(e);
    leaf_tiled_bwd_kernel<<<dim3(R * d0, a.nchunk, nsplit), (ngroups * ntile + 31) / 32 * 32, smem, st>>>(a);
    return spn_launch_status();
}
