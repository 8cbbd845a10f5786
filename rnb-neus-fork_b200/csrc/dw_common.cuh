// Pieces shared by the weight-gradient kernel (dw_gemm.cu) and the weight-gradient workers of the fused backward
// (sdf_chain.cu): the column sums the four otherwise idle warps take from the staged tiles.
#pragma once
#include "common.cuh"
#include "dw_params.h"

namespace rnb {

constexpr int DWC_STAGE_A = 32768;     // bytes of the A-side tile inside a stage (64 points x 256 columns fp16)
constexpr int DWC_WBUF_FLOATS = 3 * 64;   // per-warp scratch of the weighted column sums
constexpr int DWC_RIDER_BYTES = 16384;    // 64 points x 128 columns fp16

// Thread et (0..127) owns columns 2 et, 2 et + 1 (one 32-bit word of chunk et / 4) and walks the 64 point rows; the walk
// starts at row `chunk`, so the 8 chunks x 4 words a warp touches per step fall into 32 different banks.
// Producer lane, when it issues the stage of (pair pr, sub-tile sub): pull the 64 point weights of that sub-tile's weighted
// sums towards L2.  They are read exactly once (d_sdf: 4 bytes per point), i.e. always cold; without the hint all eight
// column-sum warps sat out a DRAM round trip per stage (the layer-8 CTAs ran 2.3x longer per stage than the others).
__device__ __forceinline__ void dw_prefetch_weights(const DwJob& job, int pr, int sub) {
#pragma unroll
    for (int s = 0; s < DW_MAX_CS; ++s) {
        if (s >= job.n_cs || job.cs[s].pair != pr || job.cs[s].n_w == 0) continue;
        const DwColsum& c = job.cs[s];
        if ((int64_t)sub * 64 >= c.n_valid) continue;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < c.n_w) {
                const float* w = c.w[k] + (int64_t)sub * 64;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(w));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(w + 32));
            }
        }
    }
}

struct DwColsumAcc {
    float v[DW_MAX_CS][3][2];
    float ws[DW_MAX_CS][3];        // sums of the weights (only thread et == 0 of a spec with wsum_partial keeps them)
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int s = 0; s < DW_MAX_CS; ++s)
#pragma unroll
            for (int k = 0; k < 3; ++k) v[s][k][0] = v[s][k][1] = ws[s][k] = 0.f;
    }
    // stage = base of the staged [A | B] tiles of 64-point sub-tile `sub`, holding pair `pr` of the job.
    // wbuf: this WARP's shared-memory scratch (DWC_WBUF_FLOATS floats): the 64 x n_w point weights of a weighted sum are
    // fetched once per stage with coalesced loads and then read back as conflict-free LDS (a global load per row and
    // weight measured 2.5 ms per launch: every one of them paid an L2 round trip).
    // Rows [row0, row0 + NR) of the sub-tile are this thread's share (two groups of four warps split the 64 rows).
    template <int NR>
    __device__ __forceinline__ void stage(const DwJob& job, int pr, int sub, const uint8_t* stage_base, int et, float* wbuf, int row0) {
        const int chunk = et >> 2, word = et & 3, lane = et & 31;
#pragma unroll
        for (int s = 0; s < DW_MAX_CS; ++s) {
            if (s >= job.n_cs || job.cs[s].pair != pr) continue;
            const DwColsum& c = job.cs[s];
            if (2 * et >= c.width) continue;
            const uint8_t* col = stage_base + c.tile_off + (size_t)chunk * 1024 + word * 4;
            if (c.n_w == 0) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll 16
                for (int r = 0; r < NR; ++r) {
                    const float2 f = unpack_h2(*reinterpret_cast<const uint32_t*>(col + (row0 + ((r + chunk) & (NR - 1))) * 16));
                    a0 += f.x;
                    a1 += f.y;
                }
                v[s][0][0] += a0;
                v[s][0][1] += a1;
            } else {
                const int64_t p0 = (int64_t)sub * 64;
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (k < c.n_w) {
#pragma unroll
                        for (int i = 0; i < NR; i += 32)
                            wbuf[k * 64 + row0 + i + lane] = p0 + row0 + i + lane < c.n_valid ? __ldg(c.w[k] + p0 + row0 + i + lane) : 0.f;
                    }
                }
                __syncwarp();
#pragma unroll 8
                for (int r = 0; r < NR; ++r) {
                    const int row = row0 + ((r + chunk) & (NR - 1));
                    const float2 f = unpack_h2(*reinterpret_cast<const uint32_t*>(col + row * 16));
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (k < c.n_w) {
                            const float wk = wbuf[k * 64 + row];
                            v[s][k][0] = fmaf(wk, f.x, v[s][k][0]);
                            v[s][k][1] = fmaf(wk, f.y, v[s][k][1]);
                            ws[s][k] += wk;
                        }
                    }
                }
            }
        }
    }
    // partial[k][split][256]
    __device__ __forceinline__ void store(const DwJob& job, int split, int et) const {
#pragma unroll
        for (int s = 0; s < DW_MAX_CS; ++s) {
            if (s >= job.n_cs) continue;
            const DwColsum& c = job.cs[s];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (k < (c.n_w == 0 ? 1 : c.n_w) && 2 * et < c.width) {
                    float* dst = c.partial[k] + (size_t)split * 256 + c.col0 + 2 * et;
                    dst[0] = v[s][k][0];
                    dst[1] = v[s][k][1];
                    if (c.wsum_partial && et == 0 && c.n_w > 0) c.wsum_partial[(size_t)split * 4 + k] = ws[s][k];
                }
            }
        }
    }
};

}  // namespace rnb
