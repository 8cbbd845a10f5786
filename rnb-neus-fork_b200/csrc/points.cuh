// Point sources and small shared device helpers of the chain kernels.
#pragma once
#include "chain.cuh"

namespace rnb {

// where the points of a launch come from
struct SdfPointSource {
    int64_t n_pts;
    const float* x;            // [n_pts,3]                          (explicit points)
    const float* rays_o;       // [B,3]  point = o + d*z             (ray samples)
    const float* rays_d;       // [B,3]
    const float* z;            // [B*n_per_ray]
    int n_per_ray;
    int grid_res;              // > 0: points of the R^3 lattice, x index offset by slab_x0
    int slab_x0;
    float bmin[3], bmax[3];
};

__device__ __forceinline__ void load_bias8(const float* b, float (&bb)[8]) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(b));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(b) + 1);
    bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w;
    bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
}

__device__ __forceinline__ void load_point(const SdfPointSource& src, int64_t p, float (&x)[3]) {
    if (src.grid_res > 0) {
        // u[ix,iy,iz] with ix = slab_x0 + p / R^2 : torch.linspace arithmetic of extract_fields
        // (reference models/renderer.py:12-14): start + step*i for the lower half, end - step*(R-1-i) above.
        const int R = src.grid_res;
        int64_t q = p < src.n_pts ? p : src.n_pts - 1;
        const int iz = (int)(q % R);
        q /= R;
        const int iy = (int)(q % R);
        const int ix = (int)(q / R) + src.slab_x0;
        const int idx[3] = {ix, iy, iz};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float lo = src.bmin[a], hi = src.bmax[a];
            const float step = __fdiv_rn(__fsub_rn(hi, lo), (float)(R - 1));
            x[a] = idx[a] < R / 2 ? __fadd_rn(lo, __fmul_rn(step, (float)idx[a]))
                                  : __fsub_rn(hi, __fmul_rn(step, (float)(R - 1 - idx[a])));
        }
    } else if (src.rays_o != nullptr) {
        // point = o + d * z  (reference models/renderer.py:863, 181); z may be a section mid-point
        const int64_t q = p < src.n_pts ? p : src.n_pts - 1;
        const int64_t ray = q / src.n_per_ray;
        const float z = __ldg(src.z + q);
#pragma unroll
        for (int a = 0; a < 3; ++a) x[a] = __ldg(src.rays_o + ray * 3 + a) + __ldg(src.rays_d + ray * 3 + a) * z;
    } else {
        const int64_t q = p < src.n_pts ? p : src.n_pts - 1;
#pragma unroll
        for (int a = 0; a < 3; ++a) x[a] = __ldg(src.x + q * 3 + a);
    }
}

// exponent of the scaled cotangent maximum: the largest input cotangent lands in [2^(RNB_COT_EXP-1), 2^RNB_COT_EXP)
#ifndef RNB_COT_EXP
#define RNB_COT_EXP 8
#endif
__device__ __forceinline__ float cot_scale_from_max(float m) {
    // power of two that maps the largest input cotangent to [128, 256): keeps every fp16 cotangent operand
    // far from overflow (x256 headroom) while the small ones stay in the normal range
    if (!(m > 0.f) || !isfinite(m)) return 1.f;
    int e;
    frexpf(m, &e);            // m = f * 2^e, f in [0.5, 1)
    return ldexpf(1.f, RNB_COT_EXP - e);
}

__device__ __forceinline__ uint32_t pack_h2_sat(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}


}  // namespace rnb
