// Adam update over flat buffers: ONE launch for the whole model (79.4 M parameters at cfg 2) instead of one multi-tensor
// pass per parameter group.  Pure HBM stream: reads p, g, m, v (16 B per parameter), writes p, m, v (12 B).
// Semantics of torch.optim.Adam (no amsgrad, no weight decay, maximize = False), which mnist_gen_train.py trains with:
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "common.cuh"

__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        float step_size, float b1, float b2, float eps, float inv_sqrt_bc2) {
    const int64_t n4 = n / 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    const float c1 = 1.f - b1, c2 = 1.f - b2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pv = p4[i], mv = m4[i], vv = v4[i];
        const float4 gv = __ldg(g4 + i);
#define SPN_ADAM1(c)                                                          \
        mv.c = b1 * mv.c + c1 * gv.c;                                         \
        vv.c = b2 * vv.c + c2 * gv.c * gv.c;                                  \
        pv.c -= step_size * mv.c / (sqrtf(vv.c) * inv_sqrt_bc2 + eps);
        SPN_ADAM1(x) SPN_ADAM1(y) SPN_ADAM1(z) SPN_ADAM1(w)
#undef SPN_ADAM1
        p4[i] = pv; m4[i] = mv; v4[i] = vv;
    }
    // tail (n % 4 elements)
    const int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float gv = g[i];
        const float mv = b1 * m[i] + c1 * gv;
        const float vv = b2 * v[i] + c2 * gv * gv;
        m[i] = mv; v[i] = vv;
        p[i] -= step_size * mv / (sqrtf(vv) * inv_sqrt_bc2 + eps);
    }
}

extern "C" int spn_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float step_size, float beta1,
                             float beta2, float eps, float inv_sqrt_bc2, void* stream) {
    if (!p || !g || !m || !v || n < 0) return SPN_ERR_ARG;
    if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) return SPN_ERR_ARG;
    if (n == 0) return SPN_OK;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    adam_step_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, step_size, beta1, beta2, eps, inv_sqrt_bc2);
    return spn_launch_status();
}
