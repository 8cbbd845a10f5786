// Albedo network (reference RenderingNetwork, mode 'no_view_dir', models/fields.py:177-215) forward and backward:
//   in = [PE4(points) 27 | PE4(normals) 27 | features 256] -> 256 ReLU -> 256 ReLU -> 3 -> sigmoid.
// Same CTA anatomy as the SDF chains (chain.cuh).  The 310-wide first layer is split over two A operands that
// accumulate into one TMEM tile: the 256 feature columns (copied chunk-for-chunk from the fp16 feature stream the
// SDF kernel wrote) and a 64-wide block holding the two positional encodings.  The 3-wide output layer and its
// adjoint run in the epilogue registers (3 dot products per point).
#include "chain.cuh"
#include "pe.cuh"
#include "points.cuh"
#include "albedo_params.h"

namespace rnb {

constexpr int ALB_A_COLS = 256;

// 64-wide PE block: [PE4(pts) 27 | PE4(normals) 27 | 0 x 10], fp16
__device__ __forceinline__ void build_pe64(const float (&x)[3], const float (&nr)[3], uint32_t (&h)[32]) {
    float e[64];
    SinCos<4> sp, sn;
    sp.compute(x[0], x[1], x[2]);
    sn.compute(nr[0], nr[1], nr[2]);
    pe_embed<4>(x, sp, e);
    pe_embed<4>(nr, sn, e + 27);
#pragma unroll
    for (int i = 54; i < 64; ++i) e[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) h[i] = pack_h2(e[2 * i], e[2 * i + 1]);
}

// A[:, 0:64] <- [PE4(point) | PE4(normal)] (half-0 threads)
__device__ __noinline__ void emit_pe64(const Epi& ep, const float (&x)[3], const float (&nr)[3], uint8_t* stream, int64_t p) {
    uint32_t h[32];
    build_pe64(x, nr, h);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 u = make_uint4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
        ep.st_a(c, u);
        if (stream) st_stream(stream, p, c, 8, u);
    }
}

__global__ void __launch_bounds__(CHAIN_THREADS, 2) albedo_fwd_kernel(const __grid_constant__ AlbedoFwdParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const ChainSmem s = chain_carve(smem, ALB_A_COLS);
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
#ifdef RNB_TRACE
        ep.tab_trace = P.tab.trace;
#endif
        const float* b0 = P.aux + ALBX_B0;
        const float* b1 = P.aux + ALBX_B1;
        const float* w2 = P.aux + ALBX_W2;
        const int ch0 = ep.col0 >> 3;
        for (int t = 0; t < n_my; ++t) {
            const int64_t p = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE_M + ep.row;
            const bool live = p < P.src.n_pts;
            // A <- feature tile (already fp16, chunked)
#pragma unroll 8
            for (int k = 0; k < 16; ++k) ep.st_a(ch0 + k, ld_stream(P.st_feat, p, ch0 + k, 32));
            // the next tile's feature rows: hint them into L2 now, three GEMM steps before they are copied
            if (t + 1 < n_my) prefetch_stream_chunks(P.st_feat, p + (int64_t)gridDim.x * TILE_M, ch0, 16);
            ep.signal();
            // step 0a done (features consumed): A[:, 0:64] <- positional encodings
            ep.wait_acc();
            if (ep.half == 0) {
                float x[3], nr[3] = {0.f, 0.f, 0.f};
                load_point(P.src, p, x);
                if (live) { nr[0] = __ldg(P.normals + p * 3); nr[1] = __ldg(P.normals + p * 3 + 1); nr[2] = __ldg(P.normals + p * 3 + 2); }
                emit_pe64(ep, x, nr, P.st_pe, p);
            }
            ep.signal();
            // step 0b: h0 = relu(z0 + b0)
            ep.wait_acc();
            ep.sweep_half_bias(b0, [&](int c0, const float (&z)[16]) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float a[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = fmaxf(z[q * 8 + j], 0.f);
                    uint4 h;
                    h.x = pack_h2(a[0], a[1]); h.y = pack_h2(a[2], a[3]); h.z = pack_h2(a[4], a[5]); h.w = pack_h2(a[6], a[7]);
                    ep.st_a((c0 >> 3) + q, h);
                    if (P.st_h0) st_stream(P.st_h0, p, (c0 >> 3) + q, 32, h);
                }
            });
            ep.signal();
            // step 1: h1 = relu(z1 + b1); albedo = sigmoid(W2 h1 + b2)
            ep.wait_acc();
            // the three output dot products accumulate on column pairs (packed fp32), folded at the end
            float2 o2[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            ep.sweep_half([&](int c0, const uint32_t (&v)[16]) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float bb[8], a[8], w[8];
                    load_bias8(b1 + c0 + q * 8, bb);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 z = f2_add(make_float2(__uint_as_float(v[q * 8 + 2 * j]), __uint_as_float(v[q * 8 + 2 * j + 1])),
                                                make_float2(bb[2 * j], bb[2 * j + 1]));
                        a[2 * j] = fmaxf(z.x, 0.f);
                        a[2 * j + 1] = fmaxf(z.y, 0.f);
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        load_bias8(w2 + k * 256 + c0 + q * 8, w);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            o2[k] = f2_fma(make_float2(a[2 * j], a[2 * j + 1]), make_float2(w[2 * j], w[2 * j + 1]), o2[k]);
                    }
                    uint4 h;
                    h.x = pack_h2(a[0], a[1]); h.y = pack_h2(a[2], a[3]); h.z = pack_h2(a[4], a[5]); h.w = pack_h2(a[6], a[7]);
                    if (P.st_h1) st_stream(P.st_h1, p, (c0 >> 3) + q, 32, h);
                }
            });
            // the two column halves of a row combine their partial dot products through the dead A buffer
            const float o[3] = {o2[0].x + o2[0].y, o2[1].x + o2[1].y, o2[2].x + o2[2].y};
            if (ep.half == 1) {
                float* xc = ep.xchg();
                xc[0] = o[0]; xc[1] = o[1]; xc[2] = o[2];
            }
            ep.sync_epi();
            if (ep.half == 0 && live) {
                const float* xc = ep.xchg();
#pragma unroll
                for (int k = 0; k < 3; ++k) P.albedo[p * 3 + k] = 1.f / (1.f + expf(-(o[k] + xc[k] + __ldg(P.aux + ALBX_B2 + k))));
            }
            // all reads of the exchange slots precede the next tile's writes of chunk 8 by the same thread
        }
    }
    chain_teardown(s, tmem);
}

__global__ void __launch_bounds__(CHAIN_THREADS, 2) albedo_bwd_kernel(const __grid_constant__ AlbedoBwdParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const ChainSmem s = chain_carve(smem, ALB_A_COLS);
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
#ifdef RNB_TRACE
        ep.tab_trace = P.tab.trace;
#endif
        const float* w2 = P.aux + ALBX_W2;
        const float scale = cot_scale_from_max(__ldg(P.cot_absmax));
        const float inv_scale = 1.f / scale;
        const int ch0 = ep.col0 >> 3;
        const int c_last = ep.col0 + EPI_HALF_COLS - 16;
        for (int t = 0; t < n_my; ++t) {
            const int64_t p = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE_M + ep.row;
            const bool live = p < P.src.n_pts;
            // dz2 = d_albedo * albedo (1 - albedo)   (sigmoid'); its scaled fp16 image is a 16-wide stream, the B operand of
            // the dW_2 = dz2^T h_1 job of the weight-gradient GEMM
            float dz2[3] = {0.f, 0.f, 0.f};
            if (live) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float a = __ldg(P.albedo + p * 3 + k);
                    dz2[k] = __ldg(P.d_albedo + p * 3 + k) * a * (1.f - a);
                }
            }
            if (ep.half == 0) {
                const uint4 u0 = make_uint4(pack_h2_sat(dz2[0] * scale, dz2[1] * scale), pack_h2_sat(dz2[2] * scale, 0.f), 0u, 0u);
                st_stream(P.st_dz2, p, 0, 2, u0);
                st_stream(P.st_dz2, p, 1, 2, make_uint4(0u, 0u, 0u, 0u));
            }
            // dz1 = (W2^T dz2) * (h1 > 0)
#pragma unroll 4
            for (int k = 0; k < 16; ++k) {
                const int ch = ch0 + k;
                const uint4 hh = ld_stream(P.st_h1, p, ch, 32);
                const uint32_t ha[4] = {hh.x, hh.y, hh.z, hh.w};
                float w0[8], w1[8], w2r[8];
                load_bias8(w2 + ch * 8, w0);
                load_bias8(w2 + 256 + ch * 8, w1);
                load_bias8(w2 + 512 + ch * 8, w2r);
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 hv = unpack_h2(ha[j]);
                    float2 d = f2_mul(f2_splat(dz2[0]), make_float2(w0[2 * j], w0[2 * j + 1]));
                    d = f2_fma(f2_splat(dz2[1]), make_float2(w1[2 * j], w1[2 * j + 1]), d);
                    d = f2_fma(f2_splat(dz2[2]), make_float2(w2r[2 * j], w2r[2 * j + 1]), d);
                    d = f2_mul(d, f2_splat(scale));
                    if (!(hv.x > 0.f)) d.x = 0.f;
                    if (!(hv.y > 0.f)) d.y = 0.f;
                    o[j] = pack_h2_sat(d.x, d.y);
                }
                const uint4 u = make_uint4(o[0], o[1], o[2], o[3]);
                ep.st_a(ch, u);
                st_stream(P.st_dz1, p, ch, 32, u);
            }
            ep.signal();
            // step B1: dh0 = dz1 W1 ; dz0 = dh0 * (h0 > 0)
            {
                uint4 hh_n[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) hh_n[q] = ld_stream(P.st_h0, p, ch0 + q, 32);
                ep.wait_acc();
                ep.sweep_half([&](int c0, const uint32_t (&v)[16]) {
                    uint4 hh_c[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) hh_c[q] = hh_n[q];
                    if (c0 < c_last) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) hh_n[q] = ld_stream(P.st_h0, p, (c0 >> 3) + 2 + q, 32);
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int ch = (c0 >> 3) + q;
                        const uint4 hh = hh_c[q];
                        const uint32_t ha[4] = {hh.x, hh.y, hh.z, hh.w};
                        uint32_t o[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 hv = unpack_h2(ha[j]);
                            const float d0 = hv.x > 0.f ? __uint_as_float(v[q * 8 + 2 * j]) : 0.f;
                            const float d1 = hv.y > 0.f ? __uint_as_float(v[q * 8 + 2 * j + 1]) : 0.f;
                            o[j] = pack_h2_sat(d0, d1);
                        }
                        const uint4 u = make_uint4(o[0], o[1], o[2], o[3]);
                        ep.st_a(ch, u);
                        st_stream(P.st_dz0, p, ch, 32, u);
                    }
                });
            }
            ep.signal();
            // step B0a: d_feat = dz0 W0[:, feat]: fp16 stream in this kernel's cotangent scale (input of the SDF backward,
            // which rescales by a power of two) + its running max; optionally also fp32 [n,256] (module API / tests)
            ep.wait_acc();
            {
                float m = 0.f;
                ep.sweep_half([&](int c0, const uint32_t (&v)[16]) {
                    if (P.st_dfeat16) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            uint4 h;
                            h.x = pack_h2_sat(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
                            h.y = pack_h2_sat(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
                            h.z = pack_h2_sat(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
                            h.w = pack_h2_sat(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
                            st_stream(P.st_dfeat16, p, (c0 >> 3) + q, 32, h);
                        }
#pragma unroll
                        for (int k = 0; k < 16; ++k) m = fmaxf(m, fabsf(__uint_as_float(v[k])));
                    }
                    if (P.d_feat && live) {
                        float4* dst = reinterpret_cast<float4*>(P.d_feat + (size_t)p * 256 + c0);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dst[j] = make_float4(__uint_as_float(v[4 * j]) * inv_scale, __uint_as_float(v[4 * j + 1]) * inv_scale,
                                                 __uint_as_float(v[4 * j + 2]) * inv_scale, __uint_as_float(v[4 * j + 3]) * inv_scale);
                    }
                });
                if (P.st_dfeat16) {
                    if (!live) m = 0.f;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                    m = fminf(m, 65504.f);
                    if (lane == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(P.dfeat_max), __float_as_uint(m));
                }
            }
            ep.signal();       // A (dz0) is reused unchanged by the next GEMM
            // step B0b (N = 64): d_pe = dz0 W0[:, pe] ; d_normal = J_4(normal)^T d_pe[27:54]
            ep.wait_acc();
            if (ep.half == 0) {
                float nr[3] = {0.f, 0.f, 0.f};
                if (live) { nr[0] = __ldg(P.normals + p * 3); nr[1] = __ldg(P.normals + p * 3 + 1); nr[2] = __ldg(P.normals + p * 3 + 2); }
                SinCos<4> sn;
                sn.compute(nr[0], nr[1], nr[2]);
                float g[3] = {0.f, 0.f, 0.f};
                uint32_t v[16];
                ep.ld_acc16(16, v);
#pragma unroll
                for (int j = 11; j < 16; ++j) pe_vjp_col<4>(16 + j - 27, sn, __uint_as_float(v[j]), g);
                ep.ld_acc16(32, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) pe_vjp_col<4>(32 + j - 27, sn, __uint_as_float(v[j]), g);
                ep.ld_acc16(48, v);
#pragma unroll
                for (int j = 0; j < 6; ++j) pe_vjp_col<4>(48 + j - 27, sn, __uint_as_float(v[j]), g);
                if (live) {
                    P.d_normal[p * 3 + 0] = g[0] * inv_scale;
                    P.d_normal[p * 3 + 1] = g[1] * inv_scale;
                    P.d_normal[p * 3 + 2] = g[2] * inv_scale;
                }
            }
        }
    }
    chain_teardown(s, tmem);
}

// ---------------------------------------------------------------------------------------- packing
struct AlbedoPackArgs {
    const float* W0; const float* b0; const float* W1; const float* b1; const float* W2; const float* b2;
    uint8_t* blob; float* aux;
};

__device__ __forceinline__ float alb_w0(const AlbedoPackArgs& a, int n, int k_feat_or_pe, bool pe) {
    // reference column order of lin0: [PE(points) 0..26 | PE(normals) 27..53 | features 54..309]
    if (pe) return k_feat_or_pe < 54 ? a.W0[n * 310 + k_feat_or_pe] : 0.f;
    return a.W0[n * 310 + 54 + k_feat_or_pe];
}

__global__ void albedo_pack_kernel(const __grid_constant__ AlbedoPackArgs a) {
    const uint32_t cid = blockIdx.x * blockDim.x + threadIdx.x;
    if (cid < ALBW_BYTES / 16) {
        const uint32_t off = cid * 16;
        uint32_t base;
        int rows, kind;
        if (off < ALBW_F0B) { base = ALBW_F0A; rows = 256; kind = 0; }
        else if (off < ALBW_F1) { base = ALBW_F0B; rows = 256; kind = 1; }
        else if (off < ALBW_T1) { base = ALBW_F1; rows = 256; kind = 2; }
        else if (off < ALBW_T0A) { base = ALBW_T1; rows = 256; kind = 3; }
        else if (off < ALBW_T0B) { base = ALBW_T0A; rows = 256; kind = 4; }
        else { base = ALBW_T0B; rows = 64; kind = 5; }
        const uint32_t local = (off - base) / 16;
        const int kc = local / rows, n = local % rows;
        __half h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = kc * 8 + j;
            float v;
            switch (kind) {
                case 0: v = alb_w0(a, n, k, false); break;           // [out n][feat k]
                case 1: v = alb_w0(a, n, k, true); break;            // [out n][pe k]
                case 2: v = a.W1[n * 256 + k]; break;                // [out n][in k]
                case 3: v = a.W1[k * 256 + n]; break;                // [in n][out k]
                case 4: v = alb_w0(a, k, n, false); break;           // [feat n][out k]
                default: v = alb_w0(a, k, n, true); break;           // [pe n][out k]
            }
            h[j] = __float2half_rn(v);
        }
        *reinterpret_cast<uint4*>(a.blob + off) = *reinterpret_cast<uint4*>(h);
    }
    if (cid < ALBX_FLOATS) {
        float v = 0.f;
        if (cid < ALBX_B1) v = a.b0[cid];
        else if (cid < ALBX_W2) v = a.b1[cid - ALBX_B1];
        else if (cid < ALBX_B2) v = a.W2[cid - ALBX_W2];
        else if (cid < ALBX_B2 + 3) v = a.b2[cid - ALBX_B2];
        a.aux[cid] = v;
    }
}

cudaError_t launch_albedo_pack(const float* W0, const float* b0, const float* W1, const float* b1, const float* W2,
                               const float* b2, uint8_t* blob, float* aux, cudaStream_t st) {
    AlbedoPackArgs a{W0, b0, W1, b1, W2, b2, blob, aux};
    const int total = ALBW_BYTES / 16;
    albedo_pack_kernel<<<(total + 255) / 256, 256, 0, st>>>(a);
    return cudaGetLastError();
}

static inline int chain_grid2(int n_tiles, int sm_count) {
    const int g = 2 * sm_count;
    return n_tiles < g ? n_tiles : g;
}

cudaError_t launch_albedo_fwd(const AlbedoFwdParams& P, int sm_count, cudaStream_t st) {
    const int smem = chain_smem_bytes(ALB_A_COLS);
    {
        cudaError_t e = ensure_dynamic_smem((const void*)albedo_fwd_kernel, smem);
        if (e != cudaSuccess) return e;
    }
    if (P.n_tiles == 0) return cudaSuccess;
    albedo_fwd_kernel<<<chain_grid2(P.n_tiles, sm_count), CHAIN_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_albedo_bwd(const AlbedoBwdParams& P, int sm_count, cudaStream_t st) {
    const int smem = chain_smem_bytes(ALB_A_COLS);
    {
        cudaError_t e = ensure_dynamic_smem((const void*)albedo_bwd_kernel, smem);
        if (e != cudaSuccess) return e;
    }
    if (P.n_tiles == 0) return cudaSuccess;
    albedo_bwd_kernel<<<chain_grid2(P.n_tiles, sm_count), CHAIN_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace rnb
