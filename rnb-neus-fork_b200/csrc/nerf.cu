// NeRF++ background field, forward (reference models/fields.py:219-314 `NeRF`, use_viewdirs head, D=8, W=256, skips=[4],
// multires=10 on the 4-D inverted-sphere point, multires_view=4; called from render_core_outside,
// models/renderer.py:93-130).  Inference only: the reference reaches it through render() -> render_novel_image
// (exp_runner.py:541), whose result is detached; render_rnb* cannot run with n_outside > 0 upstream.
//
// Same CTA anatomy as the SDF chains (chain.cuh).  Twelve GEMM steps per 128-point tile:
//   0      pts_linears.0      K = 96  (PE10(pts4): 84 columns + padding)
//   1..4   pts_linears.1..4   K = 256
//   5, 6   pts_linears.5      K = 256 (h part) then K = 96 (PE part, accumulating into the same TMEM tile): the skip
//                             concatenation cat[PE, h] (fields.py:296-298) is never materialised
//   7, 8   pts_linears.6, .7
//   9      feature_linear     (no activation)
//   10,11  views_linears.0    K = 256 (feature part) then K = 32 (PE4(dir) part), N = 128
// alpha_linear (1 output) and rgb_linear (3 outputs) are dot products in the epilogue registers.
#include "chain.cuh"
#include "pe.cuh"
#include "nerf_params.h"

namespace rnb {

constexpr int NERF_A_COLS = 256;

__device__ __forceinline__ void nerf_load_inputs(const NerfFwdParams& P, int64_t p, float (&x4)[4], float (&dir)[3]) {
    const int64_t q = p < P.n_pts ? p : P.n_pts - 1;
    if (P.pts4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) x4[j] = __ldg(P.pts4 + q * 4 + j);
#pragma unroll
        for (int j = 0; j < 3; ++j) dir[j] = __ldg(P.dirs + q * 3 + j);
    } else {
        const int64_t ray = q / P.n_per_ray;
        const float z = __ldg(P.z + q);
        float pt[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            dir[j] = __ldg(P.rays_d + ray * 3 + j);
            pt[j] = __ldg(P.rays_o + ray * 3 + j) + dir[j] * z;
        }
        // dis_to_center = |p|.clip(1, 1e10);  pts4 = [p / dis, 1 / dis]   (reference models/renderer.py:108-109)
        const float r = fminf(fmaxf(sqrtf(pt[0] * pt[0] + pt[1] * pt[1] + pt[2] * pt[2]), 1.0f), 1e10f);
        x4[0] = pt[0] / r; x4[1] = pt[1] / r; x4[2] = pt[2] / r; x4[3] = 1.0f / r;
    }
}

// A[:, 0:96] <- PE10(pts4): column j = x_j, 4+8k+j = sin(2^k x_j), 8+8k+j = cos(2^k x_j)   (embedder.py:21-55, d=4)
__device__ __noinline__ void nerf_emit_pe_pts(const Epi& ep, const float (&x4)[4]) {
    float e[NERF_PE_COLS];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        e[j] = x4[j];
        float sj, cj;
        sincosf(x4[j], &sj, &cj);
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            e[4 + 8 * k + j] = sj;
            e[8 + 8 * k + j] = cj;
            const float s2 = 2.f * sj * cj, c2 = 1.f - 2.f * sj * sj;
            sj = s2;
            cj = c2;
        }
    }
#pragma unroll
    for (int i = 84; i < NERF_PE_COLS; ++i) e[i] = 0.f;
#pragma unroll
    for (int c = 0; c < NERF_PE_COLS / 8; ++c) {
        uint4 u;
        u.x = pack_h2(e[8 * c], e[8 * c + 1]); u.y = pack_h2(e[8 * c + 2], e[8 * c + 3]);
        u.z = pack_h2(e[8 * c + 4], e[8 * c + 5]); u.w = pack_h2(e[8 * c + 6], e[8 * c + 7]);
        ep.st_a(c, u);
    }
}

// A[:, 0:32] <- PE4(dir) (27 columns + padding)
__device__ __noinline__ void nerf_emit_pe_dir(const Epi& ep, const float (&dir)[3]) {
    float e[NERF_PEV_COLS];
    SinCos<4> sc;
    sc.compute(dir[0], dir[1], dir[2]);
    pe_embed<4>(dir, sc, e);
#pragma unroll
    for (int i = 27; i < NERF_PEV_COLS; ++i) e[i] = 0.f;
#pragma unroll
    for (int c = 0; c < NERF_PEV_COLS / 8; ++c) {
        uint4 u;
        u.x = pack_h2(e[8 * c], e[8 * c + 1]); u.y = pack_h2(e[8 * c + 2], e[8 * c + 3]);
        u.z = pack_h2(e[8 * c + 4], e[8 * c + 5]); u.w = pack_h2(e[8 * c + 6], e[8 * c + 7]);
        ep.st_a(c, u);
    }
}

// h = act(acc + b) over this thread's 128 columns -> A operand; optionally accumulates <h, wdot>
template <bool RELU, bool DOT>
__device__ __forceinline__ float nerf_layer(const Epi& ep, const float* bias, const float* wdot) {
    float acc = 0.f;
    ep.sweep_half_bias(bias, [&](int c0, const float (&z)[16]) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float ww[8], a[8];
            if (DOT) load_bias8(wdot + c0 + q * 8, ww);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a[j] = RELU ? fmaxf(z[q * 8 + j], 0.f) : z[q * 8 + j];
                if (DOT) acc = fmaf(a[j], ww[j], acc);
            }
            uint4 h;
            h.x = pack_h2(a[0], a[1]); h.y = pack_h2(a[2], a[3]); h.z = pack_h2(a[4], a[5]); h.w = pack_h2(a[6], a[7]);
            ep.st_a((c0 >> 3) + q, h);
        }
    });
    return acc;
}

__global__ void __launch_bounds__(CHAIN_THREADS, 2) nerf_fwd_kernel(const __grid_constant__ NerfFwdParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const ChainSmem s = chain_carve(smem, NERF_A_COLS);
    const uint32_t tmem = chain_setup(s);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (warp == 0) {
        if (lane == 0) chain_producer(s, P.tab, P.wblob, n_my);
    } else if (warp == 1) {
        chain_mma_warp(s, P.tab, tmem, n_my);
    } else {
        Epi ep;
        ep.init(s, tmem);
        const float* bias = P.aux + NRFX_B;
        for (int t = 0; t < n_my; ++t) {
            const int64_t p = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE_M + ep.row;
            const bool live = p < P.n_pts;
            float x4[4], dir[3];
            nerf_load_inputs(P, p, x4, dir);
            if (ep.half == 0) nerf_emit_pe_pts(ep, x4);
            ep.signal();
            // pts_linears.0 .. .4
#pragma unroll 1
            for (int l = 0; l < 5; ++l) {
                ep.wait_acc();
                nerf_layer<true, false>(ep, bias + l * 256, nullptr);
                ep.signal();
            }
            // pts_linears.5: the h part has been consumed -> stage the PE part, then the accumulated tile
            ep.wait_acc();
            if (ep.half == 0) nerf_emit_pe_pts(ep, x4);
            ep.signal();
            ep.wait_acc();
            nerf_layer<true, false>(ep, bias + 5 * 256, nullptr);
            ep.signal();
            // pts_linears.6
            ep.wait_acc();
            nerf_layer<true, false>(ep, bias + 6 * 256, nullptr);
            ep.signal();
            // pts_linears.7 and the density head  alpha = <h, w_alpha> + b_alpha
            ep.wait_acc();
            float dens = nerf_layer<true, true>(ep, bias + 7 * 256, P.aux + NRFX_WA);
            ep.signal();
            // feature_linear (no activation); A (= h_7) is dead: the two halves of a row combine the density
            ep.wait_acc();
            if (ep.half == 1) *ep.xchg() = dens;
            ep.sync_epi();
            if (ep.half == 0 && live) P.density[p] = dens + *ep.xchg() + __ldg(P.aux + NRFX_BA);
            nerf_layer<false, false>(ep, P.aux + NRFX_BFEAT, nullptr);
            ep.signal();
            // views_linears.0: feature part consumed -> stage PE4(dir)
            ep.wait_acc();
            if (ep.half == 0) nerf_emit_pe_dir(ep, dir);
            ep.signal();
            // views_linears.0 accumulated (N = 128: all columns belong to half 0); rgb = W_rgb relu(.) + b_rgb
            ep.wait_acc();
            if (ep.half == 0) {
                float o[3] = {0.f, 0.f, 0.f};
                const float* bv = P.aux + NRFX_BV;
                const float* wr = P.aux + NRFX_WRGB;
                ep.sweep_half([&](int c0, const uint32_t (&v)[16]) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        float bb[8], a[8], w[8];
                        load_bias8(bv + c0 + q * 8, bb);
#pragma unroll
                        for (int j = 0; j < 8; ++j) a[j] = fmaxf(__uint_as_float(v[q * 8 + j]) + bb[j], 0.f);
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            load_bias8(wr + k * 128 + c0 + q * 8, w);
#pragma unroll
                            for (int j = 0; j < 8; ++j) o[k] = fmaf(a[j], w[j], o[k]);
                        }
                    }
                });
                if (live) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) P.rgb[p * 3 + k] = o[k] + __ldg(P.aux + NRFX_BRGB + k);
                }
            }
        }
    }
    chain_teardown(s, tmem);
}

// ---------------------------------------------------------------------------------------- packing
struct NerfPackArgs {
    const float* W[8];        // pts_linears.{0..7}.weight: [256,84], [256,256] x4, [256,340], [256,256] x2
    const float* b[8];
    const float* Wf; const float* bf;     // feature_linear [256,256]
    const float* Wa; const float* ba;     // alpha_linear   [1,256]
    const float* Wv; const float* bv;     // views_linears.0 [128,283]
    const float* Wr; const float* br;     // rgb_linear     [3,128]
    uint8_t* blob; float* aux;
};

__global__ void nerf_pack_kernel(const __grid_constant__ NerfPackArgs a) {
    const uint32_t cid = blockIdx.x * blockDim.x + threadIdx.x;
    if (cid < NRFW_BYTES / 16) {
        const uint32_t off = cid * 16;
        uint32_t base;
        int rows, kind, l = 0;
        if (off < NRFW_L1) { base = NRFW_L0; rows = 256; kind = 0; }
        else if (off < NRFW_L5H) { l = 1 + (off - NRFW_L1) / NRFW_MAT; base = NRFW_L1 + (l - 1) * NRFW_MAT; rows = 256; kind = 1; }
        else if (off < NRFW_L5E) { base = NRFW_L5H; rows = 256; kind = 2; }
        else if (off < NRFW_L6) { base = NRFW_L5E; rows = 256; kind = 3; }
        else if (off < NRFW_FEAT) { l = 6 + (off - NRFW_L6) / NRFW_MAT; base = NRFW_L6 + (l - 6) * NRFW_MAT; rows = 256; kind = 1; }
        else if (off < NRFW_VF) { base = NRFW_FEAT; rows = 256; kind = 4; }
        else if (off < NRFW_VE) { base = NRFW_VF; rows = 128; kind = 5; }
        else { base = NRFW_VE; rows = 128; kind = 6; }
        const uint32_t local = (off - base) / 16;
        const int kc = local / rows, n = local % rows;
        __half h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = kc * 8 + j;
            float v;
            switch (kind) {
                case 0: v = k < 84 ? a.W[0][n * 84 + k] : 0.f; break;
                case 1: v = a.W[l][n * 256 + k]; break;
                case 2: v = a.W[5][n * 340 + 84 + k]; break;             // cat[PE, h]: h columns follow the 84 PE columns
                case 3: v = k < 84 ? a.W[5][n * 340 + k] : 0.f; break;
                case 4: v = a.Wf[n * 256 + k]; break;
                case 5: v = a.Wv[n * 283 + k]; break;                    // cat[feature, PE(dir)]
                default: v = k < 27 ? a.Wv[n * 283 + 256 + k] : 0.f; break;
            }
            h[j] = __float2half_rn(v);
        }
        *reinterpret_cast<uint4*>(a.blob + off) = *reinterpret_cast<uint4*>(h);
    }
    if (cid < NRFX_FLOATS) {
        float v;
        if (cid < NRFX_BFEAT) v = a.b[cid / 256][cid % 256];
        else if (cid < NRFX_BV) v = a.bf[cid - NRFX_BFEAT];
        else if (cid < NRFX_WA) v = a.bv[cid - NRFX_BV];
        else if (cid < NRFX_WRGB) v = a.Wa[cid - NRFX_WA];
        else if (cid < NRFX_BA) v = a.Wr[cid - NRFX_WRGB];
        else if (cid == NRFX_BA) v = a.ba[0];
        else v = a.br[cid - NRFX_BRGB];
        a.aux[cid] = v;
    }
}

cudaError_t launch_nerf_pack(const float* const* W, const float* const* b, const float* Wf, const float* bf, const float* Wa,
                             const float* ba, const float* Wv, const float* bv, const float* Wr, const float* br,
                             uint8_t* blob, float* aux, cudaStream_t st) {
    NerfPackArgs a;
    for (int l = 0; l < 8; ++l) { a.W[l] = W[l]; a.b[l] = b[l]; }
    a.Wf = Wf; a.bf = bf; a.Wa = Wa; a.ba = ba; a.Wv = Wv; a.bv = bv; a.Wr = Wr; a.br = br;
    a.blob = blob; a.aux = aux;
    const int total = NRFW_BYTES / 16;
    nerf_pack_kernel<<<(total + 255) / 256, 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_nerf_fwd(const NerfFwdParams& P, int sm_count, cudaStream_t st) {
    const int smem = chain_smem_bytes(NERF_A_COLS);
    {
        cudaError_t e = ensure_dynamic_smem((const void*)nerf_fwd_kernel, smem);
        if (e != cudaSuccess) return e;
    }
    if (P.n_tiles == 0) return cudaSuccess;
    const int g = 2 * sm_count;
    nerf_fwd_kernel<<<P.n_tiles < g ? P.n_tiles : g, CHAIN_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace rnb
