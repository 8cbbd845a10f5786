// Host/device parameter blocks of the SDF chain kernels (plain C structs; passed as __grid_constant__).
#pragma once
#include <stdint.h>
#include "chain.cuh"
#include "points.cuh"

namespace rnb {

// packed fp16 weight blob of the SDF network (pack.cu), byte offsets
constexpr uint32_t SDFW_F0 = 0;                                   // W_0   [256 x 64]
constexpr uint32_t SDFW_F1 = 32768;                               // W_1..W_7 [256 x 256] each (W_4 pre-scaled by 1/sqrt2)
constexpr uint32_t SDFW_MAT = 131072;
constexpr uint32_t SDFW_F8 = SDFW_F1 + 7 * SDFW_MAT;              // W_8[1:,:] (feature rows)
constexpr uint32_t SDFW_T7 = SDFW_F8 + SDFW_MAT;                  // W_7^T .. W_1^T  [in x out]
constexpr uint32_t SDFW_T0 = SDFW_T7 + 7 * SDFW_MAT;              // W_0^T [64 x 256]
constexpr uint32_t SDFW_T8 = SDFW_T0 + 32768;                     // W_8[1:,:]^T [in 256 x feat 256]
constexpr uint32_t SDFW_BYTES = SDFW_T8 + SDFW_MAT;
__host__ __device__ constexpr uint32_t sdfw_fwd(int l) { return l == 0 ? SDFW_F0 : SDFW_F1 + (uint32_t)(l - 1) * SDFW_MAT; }
__host__ __device__ constexpr uint32_t sdfw_tr(int l) { return l == 0 ? SDFW_T0 : SDFW_T7 + (uint32_t)(7 - l) * SDFW_MAT; }

// fp32 side table: biases [9][256] (row 8 = b_8[1:257]), W_8[0,:], b_8[0]
constexpr int AUX_W8ROW = 9 * 256;
constexpr int AUX_B8_0 = AUX_W8ROW + 256;
constexpr int AUX_FLOATS = AUX_B8_0 + 4;

struct SdfFwdParams {
    SdfPointSource src;
    int n_tiles;
    const uint8_t* wblob;
    const float* aux;
    ChainTable tab;
    float* out;                // [n_pts] = out_scale * sdf
    float out_scale;
};

struct SdfFwdGradParams {
    SdfPointSource src;
    int n_tiles;
    const uint8_t* wblob;
    const float* aux;
    ChainTable tab;
    float* out_sdf;            // [n_pts]
    float* out_grad;           // [n_pts,3]
    float* out_full;           // optional [n_pts,257] fp32 (module API); may be null
    uint8_t* st_feat;          // fp16 stream [Npad x 256]
    uint8_t* st_in0;           // fp16 stream [Npad x 64]   layer-0 input
    uint8_t* st_in;            // 8 fp16 streams, index l = in_{l+1} = a_l
    uint8_t* st_w;             // 8 fp16 streams, w_l = s_l * ua_l
    size_t stream_stride;      // bytes of one 256-wide stream = Npad * 512
    int keep_mask;             // bit l: a_l is stored with the default (write-back) policy instead of evict-first
};

struct SdfBwdParams {
    SdfPointSource src;
    int n_tiles;
    const uint8_t* wblob;
    const float* aux;
    ChainTable tab;
    const float* d_sdf;        // [n_pts]
    const float* d_grad;       // [n_pts,3]
    const float* d_feat;       // [n_pts,256] fp32 row-major (or null)
    const uint8_t* d_feat16;   // alternative: fp16 stream [Npad x 256] scaled by the power of two derived from ...
    const float* d_feat16_cot_absmax;   // ... this device scalar (the albedo backward's cotangent absmax)
    const float* cot_absmax;   // device scalar: max |cotangent| over the three inputs
    const uint8_t* st_in;      // 8 streams a_l (softplus' is recovered from them)
    const uint8_t* st_w;
    uint8_t* st_uin0;          // fp16 stream [Npad x 64]    uin_0 (scaled)
    uint8_t* st_uin;           // 8 streams, index l = uin_{l+1} = ua_bar_l (scaled)
    uint8_t* st_zbar;          // 8 streams, zbar_l (scaled)
    uint8_t* st_dfeat;         // 1 stream, d_feat (scaled)
    size_t stream_stride;
    int keep_streams;          // store uin / zbar with the default L2 policy instead of evict-first (fused launch: read back soon)
    int thread_prefetch;       // epilogue threads hint the next step's stream chunks into L2 (prefetch.global.L2)
};

}  // namespace rnb

#include "dw_params.h"
namespace rnb {
struct SdfBwdFusedParams {
    SdfBwdParams chain;
    DwJob jobs[9];                            // job l: dW_l (l = 8: the feature rows of W_8)
    int n_dw;                                 // blocks [0, n_dw) are weight-gradient workers
    uint8_t dw_layer[FUSED_MAX_WORKERS];      // layer of worker j
    uint8_t dw_replica[FUSED_MAX_WORKERS];    // its index among the workers of that layer (partial slot)
    int* q;                                   // [9][n_tiles] tile indices in completion order, -1 = not yet published
    int* q_tail;                              // [9] entries published
    int* q_head;                              // [9] tickets taken
    int* q_cnt;                               // [n_tiles][9] epilogue warps done with (tile, layer), zeroed before launch
    int stagger_ns;                           // spread of the chain CTAs' start times (0 = none)
    unsigned long long* dbg;                  // optional [2][FUSED_MAX_WORKERS]: block end / start times (globaltimer ns)
};
}  // namespace rnb
