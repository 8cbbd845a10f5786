// Positional encoding in registers (reference models/embedder.py:21-55):
//   e(x) = [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)],  column 3+6k+j = sin(2^k x_j),
//   column 6+6k+j = cos(2^k x_j).  One accurate sincosf per coordinate, then angle doubling (freqs are 2^k).
#pragma once
#include "common.cuh"

namespace rnb {

template <int L>
struct SinCos {
    float s[L][3];
    float c[L][3];
    __device__ __forceinline__ void compute(float x0, float x1, float x2) {
        const float x[3] = {x0, x1, x2};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float sj, cj;
            sincosf(x[j], &sj, &cj);
            s[0][j] = sj;
            c[0][j] = cj;
#pragma unroll
            for (int k = 1; k < L; ++k) {
                const float s2 = 2.f * sj * cj;
                const float c2 = 1.f - 2.f * sj * sj;
                sj = s2;
                cj = c2;
                s[k][j] = sj;
                c[k][j] = cj;
            }
        }
    }
};

// e[0..3+6L): the embedding itself
template <int L>
__device__ __forceinline__ void pe_embed(const float (&x)[3], const SinCos<L>& sc, float* e) {
#pragma unroll
    for (int j = 0; j < 3; ++j) e[j] = x[j];
#pragma unroll
    for (int k = 0; k < L; ++k)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            e[3 + 6 * k + j] = sc.s[k][j];
            e[6 + 6 * k + j] = sc.c[k][j];
        }
}

// one column of J_e * gbar (index i in [0, 3+6L))  -- compile-time i after unrolling
template <int L>
__device__ __forceinline__ float pe_jvp_col(int i, const SinCos<L>& sc, const float (&g)[3]) {
    if (i < 3) return g[i];
    const int k = (i - 3) / 6, r = (i - 3) % 6;
    const float f = (float)(1 << k);
    return r < 3 ? f * sc.c[k][r] * g[r] : -f * sc.s[k][r - 3] * g[r - 3];
}

// accumulate one column of J_e^T de into g
template <int L>
__device__ __forceinline__ void pe_vjp_col(int i, const SinCos<L>& sc, float de, float (&g)[3]) {
    if (i < 3) {
        g[i] += de;
        return;
    }
    const int k = (i - 3) / 6, r = (i - 3) % 6;
    const float f = (float)(1 << k);
    if (r < 3)
        g[r] += f * sc.c[k][r] * de;
    else
        g[r - 3] -= f * sc.s[k][r - 3] * de;
}

}  // namespace rnb
