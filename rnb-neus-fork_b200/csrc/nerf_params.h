// Parameter blocks and packed-weight layout of the NeRF++ background kernels (nerf.cu).
#pragma once
#include <stdint.h>
#include "chain.cuh"
#include "points.cuh"

namespace rnb {

constexpr int NERF_PE_COLS = 96;        // PE10 of the 4-D point: 84 columns, padded to a multiple of K = 32
constexpr int NERF_PEV_COLS = 32;       // PE4 of the view direction: 27 columns, padded

// packed fp16 operand images (byte offsets), "chunked" [rows x K] (common.cuh)
constexpr uint32_t NRFW_MAT = 256 * 256 * 2;
constexpr uint32_t NRFW_L0 = 0;                                   // pts_linears.0          [256 x 96]
constexpr uint32_t NRFW_L1 = NRFW_L0 + 256 * NERF_PE_COLS * 2;    // pts_linears.1..4       [256 x 256] x 4
constexpr uint32_t NRFW_L5H = NRFW_L1 + 4 * NRFW_MAT;             // pts_linears.5[:, 84:]  [256 x 256]  (h part of the skip)
constexpr uint32_t NRFW_L5E = NRFW_L5H + NRFW_MAT;                // pts_linears.5[:, :84]  [256 x 96]   (PE part)
constexpr uint32_t NRFW_L6 = NRFW_L5E + 256 * NERF_PE_COLS * 2;   // pts_linears.6, .7      [256 x 256] x 2
constexpr uint32_t NRFW_FEAT = NRFW_L6 + 2 * NRFW_MAT;            // feature_linear         [256 x 256]
constexpr uint32_t NRFW_VF = NRFW_FEAT + NRFW_MAT;                // views_linears.0[:, :256]   [128 x 256]
constexpr uint32_t NRFW_VE = NRFW_VF + 128 * 256 * 2;             // views_linears.0[:, 256:]   [128 x 32]
constexpr uint32_t NRFW_BYTES = NRFW_VE + 128 * NERF_PEV_COLS * 2;

// fp32 side table
constexpr int NRFX_B = 0;                     // pts_linears biases [8][256]
constexpr int NRFX_BFEAT = 8 * 256;           // feature_linear.bias [256]
constexpr int NRFX_BV = NRFX_BFEAT + 256;     // views_linears.0.bias [128]
constexpr int NRFX_WA = NRFX_BV + 128;        // alpha_linear.weight [256]
constexpr int NRFX_WRGB = NRFX_WA + 256;      // rgb_linear.weight [3][128]
constexpr int NRFX_BA = NRFX_WRGB + 384;      // alpha_linear.bias
constexpr int NRFX_BRGB = NRFX_BA + 1;        // rgb_linear.bias [3]
constexpr int NRFX_FLOATS = NRFX_BRGB + 3;

struct NerfFwdParams {
    int64_t n_pts;
    int n_tiles;
    // explicit inputs (module API: NeRF.forward(input_pts [n,4], input_views [n,3])) ...
    const float* pts4;
    const float* dirs;
    // ... or ray samples (render_core_outside, reference models/renderer.py:105-113): point = o + d * z,
    // pts4 = [p / r, 1 / r] with r = max(|p|, 1), view direction = d
    const float* rays_o;
    const float* rays_d;
    const float* z;
    int n_per_ray;
    const uint8_t* wblob;
    const float* aux;
    ChainTable tab;
    float* density;           // out [n]    raw alpha_linear output
    float* rgb;               // out [n,3]  raw rgb_linear output (sigmoid is applied by the caller / compositing kernel)
};

}  // namespace rnb
