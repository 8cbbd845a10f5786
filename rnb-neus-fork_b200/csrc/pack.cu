// Parameter-sized preparation kernels: fold the effective fp32 weights of the SDF network into the packed fp16
// UMMA operand images ("chunked" K-major, common.cuh) the chain kernels stream with cp.async.bulk, plus the fp32
// side table (biases, W_8[0,:]).  One launch per optimiser step; ~2 MB written.
#include "sdf_params.h"

namespace rnb {

struct SdfPackArgs {
    const float* W[9];     // effective weights, row-major [out_l, in_l]: 256x39, 256x256 x2, 217x256, 256x256 x4, 257x256
    const float* b[9];
    uint8_t* blob;
    float* aux;
};

__device__ __forceinline__ float sdf_w_fwd(const SdfPackArgs& a, int l, int n, int k) {
    // element (out n, in k) of the layer-l operand in kernel column layout
    if (l == 0) {
        if (k < 39) return a.W[0][n * 39 + k];
        if (k < 42) return a.W[0][n * 39 + (k - 39)];      // x_lo columns reuse the x weights
        return 0.f;
    }
    if (l == 3) return n < 217 ? a.W[3][n * 256 + k] : 0.f;
    if (l == 4) return a.W[4][n * 256 + k] * 0.70710678118654752f;   // cat[a_3, e] / sqrt2 (reference fields.py:94-96)
    if (l == 8) return a.W[8][(n + 1) * 256 + k];           // feature rows
    return a.W[l][n * 256 + k];
}

__global__ void sdf_pack_kernel(const __grid_constant__ SdfPackArgs a) {
    const uint32_t cid = blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte chunk row per thread
    if (cid < SDFW_BYTES / 16) {
        const uint32_t off = cid * 16;
        int l, rows;
        bool transposed;
        uint32_t base;
        if (off < SDFW_F1) { l = 0; rows = 256; transposed = false; base = SDFW_F0; }
        else if (off < SDFW_T7) { l = 1 + (off - SDFW_F1) / SDFW_MAT; rows = 256; transposed = false; base = SDFW_F1 + (l - 1) * SDFW_MAT; }
        else if (off < SDFW_T0) { const int i = (off - SDFW_T7) / SDFW_MAT; l = 7 - i; rows = 256; transposed = true; base = SDFW_T7 + i * SDFW_MAT; }
        else if (off < SDFW_T8) { l = 0; rows = 64; transposed = true; base = SDFW_T0; }
        else { l = 8; rows = 256; transposed = true; base = SDFW_T8; }
        const uint32_t local = (off - base) / 16;
        const int kc = local / rows, n = local % rows;                // chunk index along K, row along N
        __half h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = kc * 8 + j;
            // transposed images: row n = input index, k = output index
            const float v = transposed ? sdf_w_fwd(a, l, k, n) : sdf_w_fwd(a, l, n, k);
            h[j] = __float2half_rn(v);
        }
        *reinterpret_cast<uint4*>(a.blob + off) = *reinterpret_cast<uint4*>(h);
    }
    if (cid < AUX_FLOATS) {
        float v = 0.f;
        if (cid < 9 * 256) {
            const int l = cid / 256, c = cid % 256;
            if (l == 8) v = a.b[8][c + 1];
            else if (l == 3) v = c < 217 ? a.b[3][c] : 0.f;
            else v = a.b[l][c];
        } else if (cid < AUX_B8_0) v = a.W[8][cid - AUX_W8ROW];
        else if (cid == AUX_B8_0) v = a.b[8][0];
        a.aux[cid] = v;
    }
}

cudaError_t launch_sdf_pack(const float* const* W, const float* const* b, uint8_t* blob, float* aux, cudaStream_t st) {
    SdfPackArgs a;
    for (int l = 0; l < 9; ++l) { a.W[l] = W[l]; a.b[l] = b[l]; }
    a.blob = blob;
    a.aux = aux;
    const int threads = 256, total = SDFW_BYTES / 16;
    sdf_pack_kernel<<<(total + threads - 1) / threads, threads, 0, st>>>(a);
    return cudaGetLastError();
}

// fp32 [n, cols] row-major -> fp16 stream image [Npad/64][cols/8][64 rows][16 B]; rows >= n are zero-filled
__global__ void stream_from_rowmajor_kernel(const float* x, int64_t n, int cols, int64_t n_pad, uint8_t* out) {
    const int nch = cols / 8;
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // one 16-byte chunk row per thread
    if (id >= n_pad * nch) return;
    const int64_t p = id / nch;
    const int ch = (int)(id % nch);
    __half h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = __float2half_rn(p < n ? x[p * cols + ch * 8 + j] : 0.f);
    *reinterpret_cast<uint4*>(out + stream_off(p, ch, nch)) = *reinterpret_cast<uint4*>(h);
}

cudaError_t launch_stream_from_rowmajor(const float* x, int64_t n, int cols, int64_t n_pad, uint8_t* out, cudaStream_t st) {
    if (n_pad == 0) return cudaSuccess;
    const int64_t total = n_pad * (cols / 8);
    stream_from_rowmajor_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, n, cols, n_pad, out);
    return cudaGetLastError();
}

}  // namespace rnb
