// Parameter blocks and packed-weight layout of the albedo network kernels (albedo.cu).
#pragma once
#include <stdint.h>
#include "chain.cuh"
#include "points.cuh"

namespace rnb {

// packed fp16 operand images (byte offsets)
constexpr uint32_t ALBW_F0A = 0;                         // W0[:, features]   [256 out x 256]
constexpr uint32_t ALBW_F0B = ALBW_F0A + 131072;         // W0[:, PE block]   [256 out x 64]
constexpr uint32_t ALBW_F1 = ALBW_F0B + 32768;           // W1                [256 out x 256]
constexpr uint32_t ALBW_T1 = ALBW_F1 + 131072;           // W1^T              [256 in  x 256 out]
constexpr uint32_t ALBW_T0A = ALBW_T1 + 131072;          // W0[:, features]^T [256 feat x 256 out]
constexpr uint32_t ALBW_T0B = ALBW_T0A + 131072;         // W0[:, PE block]^T [64 x 256 out]
constexpr uint32_t ALBW_BYTES = ALBW_T0B + 32768;
// fp32 side table
constexpr int ALBX_B0 = 0, ALBX_B1 = 256, ALBX_W2 = 512, ALBX_B2 = 512 + 768, ALBX_FLOATS = ALBX_B2 + 4;

struct AlbedoFwdParams {
    SdfPointSource src;
    int n_tiles;
    const uint8_t* wblob;
    const float* aux;
    ChainTable tab;
    const float* normals;      // [n,3] = d sdf / d x
    const uint8_t* st_feat;    // fp16 stream [Npad x 256] from sdf_fwd_grad
    float* albedo;             // out [n,3]
    uint8_t* st_pe;            // out fp16 stream [Npad x 64]
    uint8_t* st_h0;            // out fp16 stream [Npad x 256]
    uint8_t* st_h1;            // out fp16 stream [Npad x 256]
};

struct AlbedoBwdParams {
    SdfPointSource src;
    int n_tiles;
    const uint8_t* wblob;
    const float* aux;
    ChainTable tab;
    const float* normals;
    const float* albedo;       // [n,3] forward output
    const float* d_albedo;     // [n,3] cotangent
    const float* cot_absmax;
    const uint8_t* st_h0;
    const uint8_t* st_h1;
    int64_t n_pad;
    uint8_t* st_dz2;           // out fp16 stream [Npad x 16] (scaled): columns 0..2 = dz2, the rest zero
    uint8_t* st_dz1;           // out fp16 stream (scaled)
    uint8_t* st_dz0;           // out fp16 stream (scaled)
    float* d_feat;             // out [n,256] fp32 (optional)
    uint8_t* st_dfeat16;       // out fp16 stream [Npad x 256] = d_feat * cot scale (optional)
    float* dfeat_max;          // out device scalar: max |stored d_feat| (atomicMax; caller zeroes)
    float* d_normal;           // out [n,3] fp32
};

}  // namespace rnb
