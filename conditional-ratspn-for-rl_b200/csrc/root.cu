// Root sum layer for SHARED weights on the internal activation layout [scope][repetition][row][channel].
//
// The root merges the repetitions into the channel axis and sums over all of them (rat_spn.py:230-236):
//     out[b, k] = logsumexp_{r, l, r'} ( xL[b, l, r] + xR[b, r', r] + lw[r*c*c + l*c + r', k] )
// With eL = exp(xL - mL), eR = exp(xR - mR), W_r = exp(lw_r) this is one c x c bilinear form per (row, repetition),
//     t[b, r, k] = mL + mR + log( eL^T W_{r,k} eR ),      out[b, k] = logsumexp_r t[b, r, k],
// i.e. N = C outputs only (C = 1 for density estimation): GEMV-shaped, 2*B*R*c*c*C flops, far too narrow for the tensor
// cores but a good fit for CUDA cores with the weights of one repetition in shared memory and the row's eL / eR in a
// thread-private shared-memory column (conflict free) or registers.
// Backward with G[b, r, k] = g[b, k] * exp(mL + mR - out[b, k]):
//     dxL[l]  = eL[l]  * sum_k G_k * sum_r' W_k[l, r'] eR[r']        dxR[r'] = eR[r'] * sum_k G_k * sum_l eL[l] W_k[l, r']
//     dlw[r, l, r', k] = W_k[l, r'] * sum_b G[b, r, k] eL[b, l] eR[b, r']     (outer-product accumulation over the rows)
#include "common.cuh"

namespace {

constexpr int ROWS = 128;  // rows per block of the forward / dX kernels (one thread per row)

struct RootArgs {
    const float* x;    // [d_in][R][B][c]
    const float* lw;   // [R][c*c][C]   (= root log-weights [1, 1, R*c*c, C, 1])
    const float* out;  // [B][C]
    const float* g;    // [B][C]
    float* part;       // [R][B][C]
    float* dx;         // [d_in][R][B][c]
    float* dlw;        // [R][c*c][C]
    float* outw;       // [B][C]
    int B, d_in, c, C, R;
};

// Loads the row's left / right child activations, returns the shifts and leaves exp(x - m) in thread-private smem columns.
__device__ __forceinline__ void root_load_row(const RootArgs& a, int r, int64_t b, bool valid, float* eLs, float* eRs,
                                              float& mL, float& mR) {
    const int c = a.c;
    const bool has_l = a.d_in > 0, has_r = a.d_in > 1;
    const float* xl = a.x + ((int64_t)r * a.B + (valid ? b : 0)) * c;
    const float* xr = a.x + (((int64_t)a.R + r) * a.B + (valid ? b : 0)) * c;
    float m = -CUDART_INF_F, m2 = -CUDART_INF_F;
    for (int q = 0; q < c / 4; ++q) {
        const float4 u = (valid && has_l) ? __ldg(reinterpret_cast<const float4*>(xl) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 v = (valid && has_r) ? __ldg(reinterpret_cast<const float4*>(xr) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        eLs[(4 * q) * ROWS] = u.x; eLs[(4 * q + 1) * ROWS] = u.y; eLs[(4 * q + 2) * ROWS] = u.z; eLs[(4 * q + 3) * ROWS] = u.w;
        eRs[(4 * q) * ROWS] = v.x; eRs[(4 * q + 1) * ROWS] = v.y; eRs[(4 * q + 2) * ROWS] = v.z; eRs[(4 * q + 3) * ROWS] = v.w;
        m = fmaxf(m, fmaxf(fmaxf(u.x, u.y), fmaxf(u.z, u.w)));
        m2 = fmaxf(m2, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
    mL = (m == -CUDART_INF_F) ? 0.f : m;
    mR = (m2 == -CUDART_INF_F) ? 0.f : m2;
    for (int q = 0; q < c; ++q) {
        eLs[q * ROWS] = __expf(eLs[q * ROWS] - mL);
        eRs[q * ROWS] = __expf(eRs[q * ROWS] - mR);
    }
}

// v = eL^T W eR with W [c][c] in shared memory (row-major), eL / eR in thread-private columns (stride ROWS).
__device__ __forceinline__ float root_bilinear(const float* W, const float* eLs, const float* eRs, int c) {
    float v = 0.f;
    for (int rc = 0; rc < c; rc += 4) {
        const float e0 = eRs[rc * ROWS], e1 = eRs[(rc + 1) * ROWS], e2 = eRs[(rc + 2) * ROWS], e3 = eRs[(rc + 3) * ROWS];
        for (int l = 0; l < c; ++l) {
            const float4 w = *reinterpret_cast<const float4*>(W + l * c + rc);
            const float t = fmaf(w.x, e0, fmaf(w.y, e1, fmaf(w.z, e2, w.w * e3)));
            v = fmaf(eLs[l * ROWS], t, v);
        }
    }
    return v;
}

// grid (ceil(B / ROWS), R); dynamic smem: W [c*c] + eL [c][ROWS] + eR [c][ROWS]
__global__ void __launch_bounds__(ROWS) root_fwd_kernel(RootArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int c = a.c, cc = c * c;
    float* W = sm;
    float* eLs = sm + cc + threadIdx.x;
    float* eRs = eLs + c * ROWS;
    const int r = blockIdx.y;
    const int64_t b = (int64_t)blockIdx.x * ROWS + threadIdx.x;
    const bool valid = b < a.B;
    float mL, mR;
    root_load_row(a, r, b, valid, eLs, eRs, mL, mR);
    for (int k = 0; k < a.C; ++k) {
        __syncthreads();
        for (int i = threadIdx.x; i < cc; i += ROWS) W[i] = __expf(a.lw[((int64_t)r * cc + i) * a.C + k]);
        __syncthreads();
        const float v = root_bilinear(W, eLs, eRs, c);
        float res;
        if (v >= 1e-30f) {
            res = mL + mR + __logf(v);
        } else {
            // exact two-pass log-sum-exp (underflow / impossible rows), same rule as the general kernel in sum.cu
            const float* xl = a.d_in > 0 ? a.x + ((int64_t)r * a.B + (valid ? b : 0)) * c : nullptr;
            const float* xr = a.d_in > 1 ? a.x + (((int64_t)a.R + r) * a.B + (valid ? b : 0)) * c : nullptr;
            const float* lwp = a.lw + (int64_t)r * cc * a.C + k;
            float m = -CUDART_INF_F;
            for (int l = 0; l < c; ++l)
                for (int rp = 0; rp < c; ++rp)
                    m = fmaxf(m, (xl ? xl[l] : 0.f) + (xr ? xr[rp] : 0.f) + lwp[(int64_t)(l * c + rp) * a.C]);
            res = m;
            if (m != -CUDART_INF_F) {
                float acc = 0.f;
                for (int l = 0; l < c; ++l)
                    for (int rp = 0; rp < c; ++rp)
                        acc += expf((xl ? xl[l] : 0.f) + (xr ? xr[rp] : 0.f) + lwp[(int64_t)(l * c + rp) * a.C] - m);
                res = m + logf(acc);
            }
        }
        if (valid) a.part[((int64_t)r * a.B + b) * a.C + k] = res;
    }
}

// out[b][k] = logsumexp_r part[r][b][k]
__global__ void __launch_bounds__(256) root_merge_kernel(RootArgs a) {
    const int64_t n = (int64_t)a.B * a.C;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m = -CUDART_INF_F;
    for (int r = 0; r < a.R; ++r) m = fmaxf(m, a.part[(int64_t)r * n + i]);
    float res = m;
    if (m != -CUDART_INF_F) {
        float acc = 0.f;
        for (int r = 0; r < a.R; ++r) acc += expf(a.part[(int64_t)r * n + i] - m);
        res = m + logf(acc);
    }
    a.outw[i] = res;
}

// dX.  grid (ceil(B / ROWS), R); dynamic smem: W [c*c] + Wt [c*c] + eL, eR, dL, dR [c][ROWS] each
__global__ void __launch_bounds__(ROWS) root_bwd_x_kernel(RootArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int c = a.c, cc = c * c;
    float* W = sm;
    float* Wt = sm + cc;
    float* eLs = sm + 2 * cc + threadIdx.x;
    float* eRs = eLs + c * ROWS;
    float* dLs = eRs + c * ROWS;
    float* dRs = dLs + c * ROWS;
    const int r = blockIdx.y;
    const int64_t b = (int64_t)blockIdx.x * ROWS + threadIdx.x;
    const bool valid = b < a.B;
    float mL, mR;
    root_load_row(a, r, b, valid, eLs, eRs, mL, mR);
    for (int q = 0; q < c; ++q) { dLs[q * ROWS] = 0.f; dRs[q * ROWS] = 0.f; }
    for (int k = 0; k < a.C; ++k) {
        __syncthreads();
        for (int i = threadIdx.x; i < cc; i += ROWS) {
            const float w = __expf(a.lw[((int64_t)r * cc + i) * a.C + k]);
            W[i] = w;
            Wt[(i % c) * c + i / c] = w;
        }
        __syncthreads();
        float G = 0.f;
        if (valid) {
            const float ov = a.out[b * a.C + k];
            if (ov != -CUDART_INF_F) G = a.g[b * a.C + k] * __expf(fminf(mL + mR - ov, 80.f));
        }
        // dL[l] += G * sum_r' W[l][r'] eR[r']
        for (int l = 0; l < c; ++l) {
            float t = 0.f;
            for (int rc = 0; rc < c; rc += 4) {
                const float4 w = *reinterpret_cast<const float4*>(W + l * c + rc);
                t = fmaf(w.x, eRs[rc * ROWS], fmaf(w.y, eRs[(rc + 1) * ROWS], fmaf(w.z, eRs[(rc + 2) * ROWS], fmaf(w.w, eRs[(rc + 3) * ROWS], t))));
            }
            dLs[l * ROWS] = fmaf(G, t, dLs[l * ROWS]);
        }
        // dR[r'] += G * sum_l W[l][r'] eL[l]
        for (int rp = 0; rp < c; ++rp) {
            float t = 0.f;
            for (int lc = 0; lc < c; lc += 4) {
                const float4 w = *reinterpret_cast<const float4*>(Wt + rp * c + lc);
                t = fmaf(w.x, eLs[lc * ROWS], fmaf(w.y, eLs[(lc + 1) * ROWS], fmaf(w.z, eLs[(lc + 2) * ROWS], fmaf(w.w, eLs[(lc + 3) * ROWS], t))));
            }
            dRs[rp * ROWS] = fmaf(G, t, dRs[rp * ROWS]);
        }
    }
    if (valid) {
        if (a.d_in > 0) {
            float4* dl = reinterpret_cast<float4*>(a.dx + ((int64_t)r * a.B + b) * c);
            for (int q = 0; q < c / 4; ++q)
                dl[q] = make_float4(eLs[(4 * q) * ROWS] * dLs[(4 * q) * ROWS], eLs[(4 * q + 1) * ROWS] * dLs[(4 * q + 1) * ROWS],
                                    eLs[(4 * q + 2) * ROWS] * dLs[(4 * q + 2) * ROWS], eLs[(4 * q + 3) * ROWS] * dLs[(4 * q + 3) * ROWS]);
        }
        if (a.d_in > 1) {
            float4* dr = reinterpret_cast<float4*>(a.dx + (((int64_t)a.R + r) * a.B + b) * c);
            for (int q = 0; q < c / 4; ++q)
                dr[q] = make_float4(eRs[(4 * q) * ROWS] * dRs[(4 * q) * ROWS], eRs[(4 * q + 1) * ROWS] * dRs[(4 * q + 1) * ROWS],
                                    eRs[(4 * q + 2) * ROWS] * dRs[(4 * q + 2) * ROWS], eRs[(4 * q + 3) * ROWS] * dRs[(4 * q + 3) * ROWS]);
        }
    }
}

// dW.  grid (R, ceil(B / WROWS)); one thread per 4 x 4 block of (l, r'); A = G * eL and Bm = eR for WROWS rows in smem.
constexpr int WROWS = 128;

__global__ void __launch_bounds__(256) root_bwd_w_kernel(RootArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int c = a.c, cc = c * c, q4 = c / 4;
    float* A = sm;               // [WROWS][c]
    float* Bm = sm + WROWS * c;  // [WROWS][c]
    __shared__ float shift[WROWS], Gs[WROWS];
    const int r = blockIdx.x;
    const int64_t b0 = (int64_t)blockIdx.y * WROWS;
    const int nrow = (int)min((int64_t)WROWS, a.B - b0);
    const int nthr = blockDim.x;
    // every thread loads whole rows: exp(x - m) for both children
    for (int row = threadIdx.x; row < WROWS; row += nthr) {
        const bool valid = row < nrow;
        const float* xl = a.x + ((int64_t)r * a.B + b0 + (valid ? row : 0)) * c;
        const float* xr = a.x + (((int64_t)a.R + r) * a.B + b0 + (valid ? row : 0)) * c;
        float m = -CUDART_INF_F, m2 = -CUDART_INF_F;
        for (int q = 0; q < q4; ++q) {
            const float4 u = (valid && a.d_in > 0) ? __ldg(reinterpret_cast<const float4*>(xl) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 v = (valid && a.d_in > 1) ? __ldg(reinterpret_cast<const float4*>(xr) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            reinterpret_cast<float4*>(A + row * c)[q] = u;
            reinterpret_cast<float4*>(Bm + row * c)[q] = v;
            m = fmaxf(m, fmaxf(fmaxf(u.x, u.y), fmaxf(u.z, u.w)));
            m2 = fmaxf(m2, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
        }
        const float mL = (m == -CUDART_INF_F) ? 0.f : m, mR = (m2 == -CUDART_INF_F) ? 0.f : m2;
        for (int q = 0; q < c; ++q) {
            A[row * c + q] = valid ? __expf(A[row * c + q] - mL) : 0.f;
            Bm[row * c + q] = valid ? __expf(Bm[row * c + q] - mR) : 0.f;
        }
        shift[row] = mL + mR;
    }
    __syncthreads();
    const int t = threadIdx.x;
    const bool active = t < q4 * q4;
    const int l0 = 4 * (t / q4), r0 = 4 * (t % q4);
    for (int k = 0; k < a.C; ++k) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        __syncthreads();
        for (int row = threadIdx.x; row < nrow; row += nthr) {
            const float ov = a.out[(b0 + row) * a.C + k];
            Gs[row] = ov == -CUDART_INF_F ? 0.f : a.g[(b0 + row) * a.C + k] * __expf(fminf(shift[row] - ov, 80.f));
        }
        __syncthreads();
        if (active) {
            for (int row = 0; row < nrow; ++row) {
                const float G = Gs[row];
                const float4 u = *reinterpret_cast<const float4*>(A + row * c + l0);
                const float4 v = *reinterpret_cast<const float4*>(Bm + row * c + r0);
                const float ul[4] = {G * u.x, G * u.y, G * u.z, G * u.w};
                const float vr[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ul[i], vr[j], acc[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t wi = ((int64_t)r * cc + (l0 + i) * c + r0 + j) * a.C + k;
                    atomicAdd(a.dlw + wi, __expf(a.lw[wi]) * acc[i][j]);
                }
        }
    }
}

bool root_shape_ok(int c, int C, int R, int d_in) { return c >= 4 && c % 4 == 0 && c <= 64 && C >= 1 && R >= 1 && d_in >= 1 && d_in <= 2; }

}  // namespace

extern "C" int spn_root_shared_supported(int c, int C, int R) { return root_shape_ok(c, C, R, 2) ? 1 : 0; }

extern "C" int spn_root_shared_fwd(const float* x, int d_in, const float* lw, int B, int c, int C, int R, float* part,
                                   float* out, void* stream) {
    if (!x || !lw || !part || !out || B < 0) return SPN_ERR_ARG;
    if (!root_shape_ok(c, C, R, d_in)) return SPN_ERR_UNSUPPORTED;
    if (B == 0) return SPN_OK;
    RootArgs a{};
    a.x = x; a.lw = lw; a.part = part; a.outw = out; a.B = B; a.d_in = d_in; a.c = c; a.C = C; a.R = R;
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = (c * c + 2 * c * ROWS) * (int)sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(root_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return spn_cuda_status(e);
    root_fwd_kernel<<<dim3((B + ROWS - 1) / ROWS, R), ROWS, smem, st>>>(a);
    int rc = spn_launch_status();
    if (rc != SPN_OK) return rc;
    root_merge_kernel<<<spn_blocks((int64_t)B * C, 256), 256, 0, st>>>(a);
    return spn_launch_status();
}

extern "C" int spn_root_shared_bwd(const float* x, int d_in, const float* lw, const float* out, const float* g, int B,
                                   int c, int C, int R, float* dx, float* dlw, void* stream) {
    if (!x || !lw || !out || !g || B < 0) return SPN_ERR_ARG;
    if (!root_shape_ok(c, C, R, d_in)) return SPN_ERR_UNSUPPORTED;
    RootArgs a{};
    a.x = x; a.lw = lw; a.out = out; a.g = g; a.dx = dx; a.dlw = dlw; a.B = B; a.d_in = d_in; a.c = c; a.C = C; a.R = R;
    cudaStream_t st = (cudaStream_t)stream;
    if (dlw) {
        cudaError_t e = cudaMemsetAsync(dlw, 0, (size_t)R * c * c * C * sizeof(float), st);
        if (e != cudaSuccess) return spn_cuda_status(e);
    }
    if (B == 0) return SPN_OK;
    if (dx) {
        const int smem = (2 * c * c + 4 * c * ROWS) * (int)sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(root_bwd_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return spn_cuda_status(e);
        root_bwd_x_kernel<<<dim3((B + ROWS - 1) / ROWS, R), ROWS, smem, st>>>(a);
        int rc = spn_launch_status();
        if (rc != SPN_OK) return rc;
    }
    if (dlw) {
        const int smem = 2 * WROWS * c * (int)sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(root_bwd_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return spn_cuda_status(e);
        const int nthr = ((c / 4) * (c / 4) + 31) / 32 * 32;
        root_bwd_w_kernel<<<dim3(R, (B + WROWS - 1) / WROWS), nthr < 64 ? 64 : nthr, smem, st>>>(a);
        int rc = spn_launch_status();
        if (rc != SPN_OK) return rc;
    }
    return SPN_OK;
}
