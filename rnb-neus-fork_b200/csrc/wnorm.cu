// Weight-norm fold and its VJP for ALL layers of a network in one launch each.
// Reference: every SDFNetwork / RenderingNetwork layer is nn.utils.weight_norm(lin) (models/fields.py:72-74, 161-162):
// W = g * v / ||v||_row, recomputed by a hook in every forward; its backward is autograd's.  Through torch that is one
// small kernel per layer and direction (24 launches per train_rnb step for the 9 + 3 layers), which is what bounds the
// reference's own 512-ray batches once the MLP kernels are fast.  Parameter-sized row work: one warp per row.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rnb_b200.h"

namespace rnb {

struct WnArgs {
    int n_layers;
    int row_start[RNB_WN_MAX_LAYERS + 1];     // prefix sums of rows
    rnb_wn_layer_t L[RNB_WN_MAX_LAYERS];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ bool wn_locate(const WnArgs& a, int row, int& layer, int& r) {
    if (row >= a.row_start[a.n_layers]) return false;
    layer = 0;
    while (row >= a.row_start[layer + 1]) ++layer;
    r = row - a.row_start[layer];
    return true;
}

__global__ void __launch_bounds__(256) wn_fold_kernel(const __grid_constant__ WnArgs a) {
    const int lane = threadIdx.x & 31;
    for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); ; row += gridDim.x * 8) {
        int l, r;
        if (!wn_locate(a, row, l, r)) break;
        const rnb_wn_layer_t& L = a.L[l];
        const float* v = L.v + (size_t)r * L.cols;
        float ss = 0.f;
        for (int c = lane; c < L.cols; c += 32) ss = fmaf(v[c], v[c], ss);
        const float norm = sqrtf(warp_sum(ss));
        const float s = L.g[r] / norm;
        float* w = L.w + (size_t)r * L.cols;
        for (int c = lane; c < L.cols; c += 32) w[c] = v[c] * s;
        if (lane == 0) L.norm[r] = norm;
    }
}

// dg = <dW, v> / n ;  dv = g / n * (dW - v <dW, v> / n^2)
__global__ void __launch_bounds__(256) wn_vjp_kernel(const __grid_constant__ WnArgs a) {
    const int lane = threadIdx.x & 31;
    for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); ; row += gridDim.x * 8) {
        int l, r;
        if (!wn_locate(a, row, l, r)) break;
        const rnb_wn_layer_t& L = a.L[l];
        const float* v = L.v + (size_t)r * L.cols;
        const float* dw = L.w + (size_t)r * L.cols;       // here w = the incoming dW
        float dot = 0.f;
        for (int c = lane; c < L.cols; c += 32) dot = fmaf(dw[c], v[c], dot);
        dot = warp_sum(dot);
        const float n = L.norm[r], g = L.g[r];
        const float s = g / n, t = dot / (n * n);
        float* dv = L.dv + (size_t)r * L.cols;
        for (int c = lane; c < L.cols; c += 32) dv[c] = s * (dw[c] - v[c] * t);
        if (lane == 0) L.dg[r] = dot / n;
    }
}

static cudaError_t wn_launch(const rnb_wn_layer_t* layers, int n_layers, bool vjp, cudaStream_t st) {
    if (n_layers <= 0) return cudaSuccess;
    if (n_layers > RNB_WN_MAX_LAYERS || !layers) return cudaErrorInvalidValue;
    WnArgs a;
    a.n_layers = n_layers;
    a.row_start[0] = 0;
    for (int i = 0; i < n_layers; ++i) {
        const rnb_wn_layer_t& L = layers[i];
        if (L.rows <= 0 || L.cols <= 0 || !L.v || !L.g || !L.w || !L.norm || (vjp && (!L.dv || !L.dg))) return cudaErrorInvalidValue;
        a.L[i] = L;
        a.row_start[i + 1] = a.row_start[i] + L.rows;
    }
    const int total = a.row_start[n_layers];
    const int blocks = (total + 7) / 8;
    if (vjp) wn_vjp_kernel<<<blocks, 256, 0, st>>>(a);
    else wn_fold_kernel<<<blocks, 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_wn_fold(const rnb_wn_layer_t* layers, int n, cudaStream_t st) { return wn_launch(layers, n, false, st); }
cudaError_t launch_wn_vjp(const rnb_wn_layer_t* layers, int n, cudaStream_t st) { return wn_launch(layers, n, true, st); }

}  // namespace rnb
