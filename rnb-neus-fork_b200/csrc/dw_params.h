// Parameter blocks of the weight-gradient kernels (dw_gemm.cu).
#pragma once
#include <stdint.h>

namespace rnb {

constexpr int DW_MAX_JOBS = 16;
constexpr int RED_MAX_JOBS = 40;

struct DwJob {
    const uint8_t* a[2];      // A-side streams, 256 columns wide (M = 256 output rows of dW)
    const uint8_t* b[2];      // B-side streams
    int b_chunks[2];          // total 8-column chunks of each B stream (8, 32 or 40)
    int b_chunk0;             // first chunk of the B column window
    int n_pairs;
    int nw;                   // window width N (multiple of 16, <= 256)
    float* partial;           // [splits][256][nw] fp32
    int colsum_pair;          // pair whose A-side tile is also column-summed (bias gradient), -1 = none
    float* cs_partial;        // [splits][256] fp32
};
struct DwParams {
    int n_jobs;
    int n_sub;                // number of 64-point sub-tiles
    DwJob jobs[DW_MAX_JOBS];
};

struct ColsumJob {
    const uint8_t* stream;
    int chunks;               // stream width / 8
    int n_w;                  // 1..3 weighted sums taken in ONE pass over the stream
    const float* row_weight[3];   // optional per-point weights (fp32 [n_pts]); null = 1
    float* partial[3];            // each [splits][chunks*8]
};
struct ColsumParams {
    int n_jobs;
    int n_sub;
    int64_t n_pts;
    ColsumJob jobs[DW_MAX_JOBS * 2];
};

struct ReduceJob {
    const float* partial;     // [splits][rows][nw]
    int splits, rows, nw;
    float* dst;               // row-major, pitch dst_pitch floats
    int dst_pitch, dst_row0, dst_col0;
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
