// Parameter blocks of the weight-gradient kernels (dw_gemm.cu).
#pragma once
#include <stdint.h>

namespace rnb {

constexpr int DW_MAX_JOBS = 10;
constexpr int RED_MAX_JOBS = 40;
constexpr int DW_MAX_PAIRS = 3;
constexpr int DW_MAX_CS = 3;

// Column sums taken from a staged tile by the warps that are idle while the MMAs run: bias gradients (unweighted sum of an
// A-side cotangent tile) and the rank-1..3 weight gradients that used to need their own pass over a stream
// (SDF: dW_8[0,:] = sum_p uabar_7[p,:] + sum_p d_sdf[p] a_7[p,:];  albedo: dW_2[k,:] = sum_p dz2[k,p] h_1[p,:]).
struct DwColsum {
    int pair;                 // stage pair whose tile is summed
    int tile_off;             // byte offset of the summed tile inside the stage: 0 = A side, 32768 = B side, above = rider
    int width;                // columns of that tile (256, nw, or 128 for a rider half)
    int col0;                 // first output column (a rider half covers columns col0 .. col0 + 127 of the sum)
    int n_w;                  // 0: one unweighted sum -> partial[0];  1..3: that many weighted sums over the same tile
    const float* w[3];        // per-point weights (fp32, indexed by point); points >= n_valid weigh 0
    int64_t n_valid;
    float* partial[3];        // each [splits][256] fp32
    float* wsum_partial;      // optional [splits][4]: the sums of the weights themselves (db of a 1..3-row layer)
};
struct DwJob {
    const uint8_t* a[DW_MAX_PAIRS];      // A-side streams, 256 columns wide (M = 256 output rows of dW)
    const uint8_t* b[DW_MAX_PAIRS];      // B-side streams (unused for pairs >= mma_pairs)
    int b_chunks[DW_MAX_PAIRS];          // total 8-column chunks of each B stream (8, 32 or 40)
    // "rider": 128 columns (16 KB per sub-tile) of a 256-wide stream that is only column-summed, carried in the unused
    // part of a narrow job's B slot (nw = 64 leaves 24 KB) so that it costs no ring stage of its own; null = none
    const uint8_t* x[DW_MAX_PAIRS];
    int x_chunk0[DW_MAX_PAIRS];          // first 8-column chunk of the rider window (0 or 16)
    int b_chunk0;             // first chunk of the B column window
    int n_pairs;              // stages per 64-point sub-tile
    int mma_pairs;            // pairs [0, mma_pairs) are contracted; the rest only stage their A tile for column sums
    int nw;                   // window width N (multiple of 16, <= 256)
    float* partial;           // [splits][256][nw] fp32
    int n_cs;
    DwColsum cs[DW_MAX_CS];
};
struct DwParams {
    int n_jobs;
    int n_sub;                // number of 64-point sub-tiles
    DwJob jobs[DW_MAX_JOBS];
};

// fused backward (sdf_chain.cu, sdf_bwd_fused_kernel): K3a parameters + one weight-gradient job per layer + the hand-over
// queues.  job.partial / job.cs_partial are indexed by the worker's replica index instead of a split index.
constexpr int FUSED_MAX_WORKERS = 160;
struct ReduceJob {
    const float* partial;     // [splits][rows][nw]
    int splits, rows, nw;
    float* dst;               // row-major, pitch dst_pitch floats
    int dst_pitch, dst_row0, dst_col0;
    int dst_col_stride;       // distance between output columns (0 = 1): dst_pitch = 1 with a column stride writes transposed
    int out_rows, out_cols;
    const float* partial2;    // optional second source with its own scaling (same shape), null = none
    int splits2;
    float factor2;
    int use_cot_scale2;
    float factor;
    int use_cot_scale;        // divide by the power-of-two cotangent scale derived from *cot_absmax
    int fold_xlo;             // layer 0: add the x_lo columns 39..41 onto columns 0..2
    int accumulate;
};
struct ReduceParams {
    int n_jobs;
    const float* cot_absmax;
    ReduceJob jobs[RED_MAX_JOBS];
};

}  // namespace rnb
