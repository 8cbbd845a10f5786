// Step epilogue: the reference runs torch.optim.Adam over 61 small parameter tensors
// (exp_runner.py:115, 263).  Here parameters, gradients and both moments live in four flat fp32 buffers (the gradient
// buffer is the one the data-parallel all-reduce works on, rnb_b200/parallel.py), and one launch updates everything.
// HBM-bound byte work: 16 B read + 12 B written per parameter, float4 accesses, grid-stride.
//
// Arithmetic follows torch.optim.Adam's single-tensor path (amsgrad off, weight_decay 0, maximize off) operation by
// operation in fp32, with the step-dependent scalars prepared by the host in double:
//   m += (g - m) * (1 - beta1)                    (lerp_)
//   v  = v * beta2 + (1 - beta2) * g * g          (mul_ / addcmul_)
//   p -= step_size * m / (sqrt(v) / sqrt(bc2) + eps),   step_size = lr / bc1
// `grad_scale` multiplies g first (1/world after a summing all-reduce; 1 otherwise).
#include <cuda_runtime.h>
#include <stdint.h>

namespace rnb {

struct AdamScalars { float one_minus_b1, b2, one_minus_b2, step_size, sqrt_bc2, eps, grad_scale; };

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamScalars& s) {
    g *= s.grad_scale;
    m = m + (g - m) * s.one_minus_b1;
    v = v * s.b2 + s.one_minus_b2 * g * g;
    const float denom = sqrtf(v) / s.sqrt_bc2 + s.eps;
    p = p - s.step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        AdamScalars s) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = t0; i < n4; i += stride) {
        float4 P = reinterpret_cast<float4*>(p)[i];
        const float4 G = reinterpret_cast<const float4*>(g)[i];
        float4 M = reinterpret_cast<float4*>(m)[i];
        float4 V = reinterpret_cast<float4*>(v)[i];
        adam_elem(P.x, G.x, M.x, V.x, s);
        adam_elem(P.y, G.y, M.y, V.y, s);
        adam_elem(P.z, G.z, M.z, V.z, s);
        adam_elem(P.w, G.w, M.w, V.w, s);
        reinterpret_cast<float4*>(p)[i] = P;
        reinterpret_cast<float4*>(m)[i] = M;
        reinterpret_cast<float4*>(v)[i] = V;
    }
    for (int64_t i = (n4 << 2) + t0; i < n; i += stride) adam_elem(p[i], g[i], m[i], v[i], s);
}

cudaError_t launch_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                             double eps, int64_t step, double grad_scale, int sm_count, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (step < 1 || !p || !g || !m || !v) return cudaErrorInvalidValue;
    if ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) != 0) return cudaErrorMisalignedAddress;
    double b1t = 1.0, b2t = 1.0;                       // beta ** step like Python's float pow (exact enough: same doubles)
    b1t = pow(beta1, (double)step);
    b2t = pow(beta2, (double)step);
    const double bc1 = 1.0 - b1t, bc2 = 1.0 - b2t;
    AdamScalars s;
    s.one_minus_b1 = (float)(1.0 - beta1);
    s.b2 = (float)beta2;
    s.one_minus_b2 = (float)(1.0 - beta2);
    s.step_size = (float)(lr / bc1);
    s.sqrt_bc2 = (float)sqrt(bc2);
    s.eps = (float)eps;
    s.grad_scale = (float)grad_scale;
    const int64_t n4 = (n + 3) >> 2;
    int64_t blocks = (n4 + 255) / 256;
    const int64_t cap = (int64_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    adam_flat_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, s);
    return cudaGetLastError();
}

}  // namespace rnb
