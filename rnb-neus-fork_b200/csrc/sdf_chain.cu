// SDF MLP kernels (reference models/fields.py:82-127 and its autograd double-backward, exp_runner.py:261):
//   sdf_fwd_kernel<GRID>      K1/K8  PE + 8 hidden layers, sdf = <a_7, W_8[0,:]> + b_8[0]        (no_grad evals, grid)
//   sdf_fwd_grad_kernel       K2     forward + feature layer + analytic dx-chain, saves the streams K3 needs
//   sdf_bwd_data_kernel       K3a    adjoint of the dx-chain (phase A) + ordinary backward (phase B); writes the
//                                    cotangent streams the dW GEMM (dw_gemm.cu) contracts over the points
// All GEMMs are tcgen05.mma kind::f16 (fp16 operands, fp32 accumulate in TMEM); see chain.cuh for the CTA anatomy.
#include "chain.cuh"
#include "pe.cuh"
#include "sdf_params.h"

namespace rnb {

constexpr int SDF_A_COLS = 256;
constexpr int SKIP_COL = 217;      // layer 3 has 217 outputs; columns 217..255 of layer 4's input are the PE

// fp16 image of the 64-wide layer-0 input: [x_hi(3) | sin/cos(36) | x_lo(3) | 0...]; x = x_hi + x_lo keeps the
// linear term of the SDF at fp32-like accuracy although the operands are fp16.
__device__ __forceinline__ void build_in0(const float (&x)[3], const SinCos<6>& sc, uint32_t (&h)[32]) {
    float e[64];
    pe_embed<6>(x, sc, e);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float hi = __half2float(__float2half_rn(x[j]));
        e[39 + j] = x[j] - hi;
    }
#pragma unroll
    for (int i = 42; i < 64; ++i) e[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) h[i] = pack_h2(e[2 * i], e[2 * i + 1]);
}

// ======================================================================================= K1 / K8
// One hidden layer's epilogue over this thread's 128 columns: z = acc + b, a = softplus(z) -> next A operand.
// LAST also reduces this half's part of the sdf row.
__device__ __forceinline__ void fwd_layer_plain(const Epi& ep, const float* bias) {
    // (the software-pipelined bias sweep of K2 costs this kernel spills: plain loads here)
    ep.sweep_half([&](int c0, const uint32_t (&v)[16]) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float bb[8];
            load_bias8(bias + c0 + q * 8, bb);
            uint32_t hh[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 z = f2_add(make_float2(__uint_as_float(v[q * 8 + 2 * j]), __uint_as_float(v[q * 8 + 2 * j + 1])),
                                        make_float2(bb[2 * j], bb[2 * j + 1]));
                const float2 a = softplus100_x2(z);
                hh[j] = pack_h2(a.x, a.y);
            }
            ep.st_a((c0 >> 3) + q, make_uint4(hh[0], hh[1], hh[2], hh[3]));
        }
    });
}
// last hidden layer: this half's part of sdf = <softplus(z_7), W_8[0,:]>
__device__ __forceinline__ float fwd_layer_last(const Epi& ep, const float* bias, const float* w8row) {
    float acc = 0.f;
    ep.sweep_half([&](int c0, const uint32_t (&v)[16]) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float bb[8], ww[8];
            load_bias8(bias + c0 + q * 8, bb);
            load_bias8(w8row + c0 + q * 8, ww);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 z = f2_add(make_float2(__uint_as_float(v[q * 8 + 2 * j]), __uint_as_float(v[q * 8 + 2 * j + 1])),
                                        make_float2(bb[2 * j], bb[2 * j + 1]));
                const float2 a = softplus100_x2(z);
                acc = fmaf(a.x, ww[2 * j], acc);
                acc = fmaf(a.y, ww[2 * j + 1], acc);
            }
        }
    });
    return acc;
}

// overwrite columns 217..255 of the A operand (and optionally a global stream) with 39 values
template <bool WRITE_A = true, class F>
__device__ __forceinline__ void write_skip_cols(const Epi& ep, F&& val, uint8_t* stream, int64_t p) {
    // chunk 27 holds columns 216..223: keep column 216
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const __half h = __float2half_rn(val(i));
        if (WRITE_A) ep.st_a_half(SKIP_COL + i, h);
        if (stream) *reinterpret_cast<__half*>(stream + stream_off(p, 27, 32) + (1 + i) * 2) = h;
    }
#pragma unroll
    for (int c = 28; c < 32; ++c) {
        const int i0 = c * 8 - SKIP_COL;
        uint4 h;
        h.x = pack_h2(val(i0 + 0), val(i0 + 1)); h.y = pack_h2(val(i0 + 2), val(i0 + 3));
        h.z = pack_h2(val(i0 + 4), val(i0 + 5)); h.w = pack_h2(val(i0 + 6), val(i0 + 7));
        if (WRITE_A) ep.st_a(c, h);
        if (stream) st_stream(stream, p, c, 32, h);
    }
}

__device__ __forceinline__ void write_in0(const Epi& ep, const uint32_t (&h)[32], uint8_t* stream, int64_t p) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 u = make_uint4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
        ep.st_a(c, u);
        if (stream) st_stream(stream, p, c, 8, u);
    }
}

// layer-0 operand of a point (half-0 threads): PE + x_lo columns
__device__ __noinline__ void emit_in0(const Epi& ep, const float (&x)[3], uint8_t* stream, int64_t p) {
    SinCos<6> sc;
    sc.compute(x[0], x[1], x[2]);
    uint32_t h[32];
    build_in0(x, sc, h);
    write_in0(ep, h, stream, p);
}
// skip connection (half-1 threads): columns 217..255 of layer 4's input are the PE of the point
__device__ __noinline__ void emit_skip_pe(const Epi& ep, const float (&x)[3], uint8_t* stream, int64_t p) {
    SinCos<6> sc;
    sc.compute(x[0], x[1], x[2]);
    float e[39];
    pe_embed<6>(x, sc, e);
    write_skip_cols(ep, [&](int i) { return e[i]; }, stream, p);
}

__global__ void __launch_bounds__(CHAIN_THREADS, 2) sdf_fwd_kernel(const __grid_constant__ SdfFwdParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const ChainSmem s = chain_carve(smem, SDF_A_COLS);
    const uint32_t tmem = chain_setup(s);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (warp == 0) {
        if (lane == 0) chain_producer(s, P.tab, P.wblob, n_my);
    } else if (warp == 1) {
#ifdef RNB_MMA_V1
        if (lane == 0) chain_mma(s, P.tab, tmem, n_my);
#else
        chain_mma_warp(s, P.tab, tmem, n_my);
#endif
    } else {
        Epi ep;
        ep.init(s, tmem);
#ifdef RNB_TRACE
        ep.tab_trace = P.tab.trace;
#endif
        const float* bias = P.aux;
        const float* w8row = P.aux + AUX_W8ROW;
        const float b8 = __ldg(P.aux + AUX_B8_0);
        for (int t = 0; t < n_my; ++t) {
            const int64_t p = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE_M + ep.row;
            float x[3];
            load_point(P.src, p, x);
            if (ep.half == 0) emit_in0(ep, x, nullptr, p);
            ep.signal();
            float part = 0.f;
#pragma unroll 1
            for (int l = 0; l < 8; ++l) {
                ep.wait_acc();
                if (l < 7) {
                    fwd_layer_plain(ep, bias + l * 256);
                    if (l == 3 && ep.half == 1) emit_skip_pe(ep, x, nullptr, p);
                    ep.signal();
                } else {
                    part = fwd_layer_last(ep, bias + l * 256, w8row);
                }
            }
            // sdf = <a_7, W_8[0,:]> + b_8[0]: the two column halves of a row meet through the (now dead) A buffer
            if (ep.half == 1) *ep.xchg() = part;
            ep.sync_epi();
            if (ep.half == 0 && p < P.src.n_pts) P.out[p] = P.out_scale * (part + *ep.xchg() + b8);
        }
    }
    chain_teardown(s, tmem);
}

// ======================================================================================= K2
__global__ void __launch_bounds__(CHAIN_THREADS, 2) sdf_fwd_grad_kernel(const __grid_constant__ SdfFwdGradParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const ChainSmem s = chain_carve(smem, SDF_A_COLS);
    const uint32_t tmem = chain_setup(s);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (warp == 0) {
        if (lane == 0) chain_producer(s, P.tab, P.wblob, n_my);
    } else if (warp == 1) {
#ifdef RNB_MMA_V1
        if (lane == 0) chain_mma(s, P.tab, tmem, n_my);
#else
        chain_mma_warp(s, P.tab, tmem, n_my);
#endif
    } else {
        Epi ep;
        ep.init(s, tmem);
#ifdef RNB_TRACE
        ep.tab_trace = P.tab.trace;
#endif
        const float* bias = P.aux;
        const float* w8row = P.aux + AUX_W8ROW;
        const size_t SS = P.stream_stride;
        const int c_last = ep.col0 + EPI_HALF_COLS - 16;
        for (int t = 0; t < n_my; ++t) {
            const int64_t p = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE_M + ep.row;
            const bool live = p < P.src.n_pts;
            float x[3];
            load_point(P.src, p, x);
            if (ep.half == 0) emit_in0(ep, x, P.st_in0, p);
            ep.signal();
            // ---------------- forward, layers 0..7: a_l -> A operand and stream in_{l+1}
            float sdf = 0.f;
#pragma unroll 1
            for (int l = 0; l < 8; ++l) {
                ep.wait_acc();
                uint8_t* st_a_next = P.st_in + (size_t)l * SS;       // in_{l+1} = a_l
                const bool keep_l = (P.keep_mask >> l) & 1;
                const float* bl = bias + l * 256;
                ep.sweep_half_bias(bl, [&](int c0, const float (&z)[16]) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        float a[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 aa = softplus100_x2(make_float2(z[q * 8 + 2 * j], z[q * 8 + 2 * j + 1]));
                            a[2 * j] = aa.x;
                            a[2 * j + 1] = aa.y;
                        }
                        uint4 ha;
                        ha.x = pack_h2(a[0], a[1]); ha.y = pack_h2(a[2], a[3]); ha.z = pack_h2(a[4], a[5]); ha.w = pack_h2(a[6], a[7]);
                        const int ch = (c0 >> 3) + q;
                        ep.st_a(ch, ha);
                        if (keep_l) *reinterpret_cast<uint4*>(st_a_next + stream_off(p, ch, 32)) = ha;
                        else st_stream(st_a_next, p, ch, 32, ha);
                    }
                });
                if (l == 3 && ep.half == 1) emit_skip_pe(ep, x, st_a_next, p);
                ep.signal();
            }
            // ---------------- layer 8 features; then seed the dx-chain: w_7 = s_7 * W_8[0,:]
            ep.wait_acc();
            {
                const float* b8f = bias + 8 * 256;
                uint8_t* st_w7 = P.st_w + (size_t)7 * SS;
                ep.sweep_half([&](int c0, const uint32_t (&v)[16]) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        float bb[8], ww[8];
                        load_bias8(b8f + c0 + q * 8, bb);
                        load_bias8(w8row + c0 + q * 8, ww);
                        const int ch = (c0 >> 3) + q;
                        float f[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[q * 8 + j]) + bb[j];
                        uint4 hf;
                        hf.x = pack_h2(f[0], f[1]); hf.y = pack_h2(f[2], f[3]); hf.z = pack_h2(f[4], f[5]); hf.w = pack_h2(f[6], f[7]);
                        st_stream(P.st_feat, p, ch, 32, hf);
                        if (P.out_full && live) {
                            float* o = P.out_full + (size_t)p * 257 + 1 + c0 + q * 8;
#pragma unroll
                            for (int j = 0; j < 8; ++j) o[j] = f[j];
                        }
                        // a_7 is still in the A operand buffer (this GEMM just consumed it): s_7 = 1 - exp(-100 a_7)
                        const uint4 ha7 = ep.ld_a(ch);
                        const float2 s0 = unpack_h2(ha7.x), s1 = unpack_h2(ha7.y), s2 = unpack_h2(ha7.z), s3 = unpack_h2(ha7.w);
                        uint4 hw;
                        // sdf = <a_7, W_8[0,:]> + b_8[0]: this half's part, from the operand image of a_7 (fp16, exactly what
                        // the 256 feature rows of the same layer are computed from) -- a special case inside the layer-7
                        // sweep cost that step ~6000 cycles per tile (registers), here W_8[0,:] is loaded anyway
                        sdf = fmaf(s0.x, ww[0], sdf); sdf = fmaf(s0.y, ww[1], sdf); sdf = fmaf(s1.x, ww[2], sdf); sdf = fmaf(s1.y, ww[3], sdf);
                        sdf = fmaf(s2.x, ww[4], sdf); sdf = fmaf(s2.y, ww[5], sdf); sdf = fmaf(s3.x, ww[6], sdf); sdf = fmaf(s3.y, ww[7], sdf);
                        const float2 w0 = sigmul_x2(s0, make_float2(ww[0], ww[1])), w1 = sigmul_x2(s1, make_float2(ww[2], ww[3]));
                        const float2 w2 = sigmul_x2(s2, make_float2(ww[4], ww[5])), w3 = sigmul_x2(s3, make_float2(ww[6], ww[7]));
                        hw.x = pack_h2(w0.x, w0.y); hw.y = pack_h2(w1.x, w1.y); hw.z = pack_h2(w2.x, w2.y); hw.w = pack_h2(w3.x, w3.y);
                        ep.st_a(ch, hw);
                        if (P.st_w) st_stream(st_w7, p, ch, 32, hw);
                    }
                });
            }
            ep.signal();
            // ---------------- dx-chain: GEMM l (= 7..1) yields ua_{l-1}; w_{l-1} = s_{l-1} * ua_{l-1}
            float g[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
            for (int l = 7; l >= 1; --l) {
                const uint8_t* st_sp = P.st_in + (size_t)(l - 1) * SS;      // a_{l-1}; s_{l-1} = 1 - exp(-100 a)
                uint8_t* st_wp = P.st_w + (size_t)(l - 1) * SS;
                uint4 hs_n[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) hs_n[q] = ld_stream(st_sp, p, (ep.col0 >> 3) + q, 32);
                ep.wait_acc();
                if (l == 7) {
                    // A (= w_7) is dead: the two halves of the row combine their parts of <a_7, W_8[0,:]>
                    if (ep.half == 1) *ep.xchg() = sdf;
                    ep.sync_epi();
                    if (ep.half == 0) {
                        sdf += *ep.xchg() + __ldg(P.aux + AUX_B8_0);
                        if (live) {
                            P.out_sdf[p] = sdf;
                            if (P.out_full) P.out_full[(size_t)p * 257] = sdf;
                        }
                    }
                }
                ep.sweep_half([&](int c0, const uint32_t (&v)[16]) {
                    uint4 hs_c[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) hs_c[q] = hs_n[q];
                    if (c0 < c_last) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) hs_n[q] = ld_stream(st_sp, p, (c0 >> 3) + 2 + q, 32);
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int ch = (c0 >> 3) + q;
                        const uint4 hs = hs_c[q];
                        const float2 s0 = unpack_h2(hs.x), s1 = unpack_h2(hs.y), s2 = unpack_h2(hs.z), s3 = unpack_h2(hs.w);
                        float u[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) u[j] = __uint_as_float(v[q * 8 + j]);
                        uint4 hw;
                        const float2 w0 = sigmul_x2(s0, make_float2(u[0], u[1])), w1 = sigmul_x2(s1, make_float2(u[2], u[3]));
                        const float2 w2 = sigmul_x2(s2, make_float2(u[4], u[5])), w3 = sigmul_x2(s3, make_float2(u[6], u[7]));
                        hw.x = pack_h2(w0.x, w0.y); hw.y = pack_h2(w1.x, w1.y); hw.z = pack_h2(w2.x, w2.y); hw.w = pack_h2(w3.x, w3.y);
                        ep.st_a(ch, hw);
                        if (P.st_w) st_stream(st_wp, p, ch, 32, hw);
                    }
                });
                if (l == 4) {
                    if (ep.half == 1) {
                        // columns 217..255 of layer 4's input are the PE, not activations: w_3 = 0 there
                        write_skip_cols(ep, [](int) { return 0.f; }, P.st_w ? st_wp : nullptr, p);
                    } else {
                        // uin_4[217:] is d sdf / d e through the skip connection: fold it into the gradient now
                        SinCos<6> sc;
                        sc.compute(x[0], x[1], x[2]);
                        uint32_t v[16];
                        ep.ld_acc16(208, v);
#pragma unroll
                        for (int j = SKIP_COL - 208; j < 16; ++j) pe_vjp_col<6>(208 + j - SKIP_COL, sc, __uint_as_float(v[j]), g);
                        ep.ld_acc16(224, v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) pe_vjp_col<6>(224 + j - SKIP_COL, sc, __uint_as_float(v[j]), g);
                        ep.ld_acc16(240, v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) pe_vjp_col<6>(240 + j - SKIP_COL, sc, __uint_as_float(v[j]), g);
                    }
                }
                ep.signal();
            }
            // ---------------- last dx GEMM (N = 64): uin_0 = d sdf / d e
            ep.wait_acc();
            if (ep.half == 0) {
                SinCos<6> sc;
                sc.compute(x[0], x[1], x[2]);
                uint32_t v[16];
                ep.ld_acc16(0, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) pe_vjp_col<6>(j, sc, __uint_as_float(v[j]), g);
                ep.ld_acc16(16, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) pe_vjp_col<6>(16 + j, sc, __uint_as_float(v[j]), g);
                ep.ld_acc16(32, v);
#pragma unroll
                for (int j = 0; j < 7; ++j) pe_vjp_col<6>(32 + j, sc, __uint_as_float(v[j]), g);
                if (live) {
                    P.out_grad[p * 3 + 0] = g[0];
                    P.out_grad[p * 3 + 1] = g[1];
                    P.out_grad[p * 3 + 2] = g[2];
                }
            }
            // the next tile's prologue signals; nothing to do here
        }
    }
    chain_teardown(s, tmem);
}

// ======================================================================================= K3a
// uin_0 = J_e gbar in the 64-wide layer-0 column layout (x_lo columns carry no cotangent); half-0 threads
__device__ __noinline__ void emit_uin0(const Epi& ep, const float (&x)[3], const float (&gb)[3], uint8_t* stream, int64_t p) {
    SinCos<6> sc;
    sc.compute(x[0], x[1], x[2]);
    uint32_t h[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float e0 = 2 * i < 39 ? pe_jvp_col<6>(2 * i, sc, gb) : 0.f;
        const float e1 = 2 * i + 1 < 39 ? pe_jvp_col<6>(2 * i + 1, sc, gb) : 0.f;
        h[i] = pack_h2_sat(e0, e1);
    }
    write_in0(ep, h, stream, p);
}
// uin_4 = cat[ua_bar_3, ebar] (the 1/sqrt2 lives in the packed W_4); half-1 threads
__device__ __noinline__ void emit_skip_ebar(const Epi& ep, const float (&x)[3], const float (&gb)[3], uint8_t* stream, int64_t p) {
    SinCos<6> sc;
    sc.compute(x[0], x[1], x[2]);
    float e[39];
#pragma unroll
    for (int i = 0; i < 39; ++i) e[i] = pe_jvp_col<6>(i, sc, gb);
    write_skip_cols(ep, [&](int i) { return fminf(fmaxf(e[i], -65504.f), 65504.f); }, stream, p);
}

// zbar = s abar + z2,  z2 = softplus''(z) ua wbar = 100 (1 - s) w (uin / s)   with s = 1 - e, e = exp(-100 a);
// w = s ua and uin = s wbar are the fp16 streams.  Where s underflows (a -> 0) both streams are (sub)normal-tiny and
// z2 <= 100 |ua| |uin| is far below the fp16 resolution of any zbar that matters: the quotient is clamped to 0 there.
__device__ __forceinline__ float zbar_elem(float a, float w, float u, float abar) {
    const float e = ex2_approx(-144.26950408889634f * a);
    const float s = 1.f - e;
    const float r = s > 1e-6f ? rcp_approx(s) : 0.f;
    return fmaf(s, abar, 100.f * e * w * (u * r));
}
// two columns at once (packed fp32; same operations and rounding lane by lane)
__device__ __forceinline__ float2 zbar_x2(float2 a, float2 w, float2 u, float2 abar) {
    const float2 y = f2_mul(a, f2_splat(144.26950408889634f));
    float2 e;
    e.x = ex2_approx(-y.x);
    e.y = ex2_approx(-y.y);
    const float2 s = f2_fma(e, f2_splat(-1.f), f2_splat(1.f));
    float2 r;
    r.x = s.x > 1e-6f ? rcp_approx(s.x) : 0.f;
    r.y = s.y > 1e-6f ? rcp_approx(s.y) : 0.f;
    const float2 t = f2_mul(f2_mul(f2_mul(e, f2_splat(100.f)), w), f2_mul(u, r));
    return f2_fma(s, abar, t);
}

__global__ void __launch_bounds__(CHAIN_THREADS, 2) sdf_bwd_data_kernel(const __grid_constant__ SdfBwdParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const ChainSmem s = chain_carve(smem, SDF_A_COLS);
    const uint32_t tmem = chain_setup(s);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (warp == 0) {
        if (lane == 0) chain_producer(s, P.tab, P.wblob, n_my);
    } else if (warp == 1) {
#ifdef RNB_MMA_V1
        if (lane == 0) chain_mma(s, P.tab, tmem, n_my);
#else
        chain_mma_warp(s, P.tab, tmem, n_my);
#endif
    } else {
        Epi ep;
        ep.init(s, tmem);
#ifdef RNB_TRACE
        ep.tab_trace = P.tab.trace;
#endif
        const float* w8row = P.aux + AUX_W8ROW;
        const size_t SS = P.stream_stride;
        const float scale = cot_scale_from_max(__ldg(P.cot_absmax));
        const float dfeat_rescale = P.d_feat16 ? scale / cot_scale_from_max(__ldg(P.d_feat16_cot_absmax)) : 0.f;
        const int c_last = ep.col0 + EPI_HALF_COLS - 16;
        const int ch0 = ep.col0 >> 3;
        for (int t = 0; t < n_my; ++t) {
            const int64_t p = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE_M + ep.row;
            const bool live = p < P.src.n_pts;
            float x[3];
            load_point(P.src, p, x);
            float gb[3] = {0.f, 0.f, 0.f};
            float dsdf = 0.f;
            if (live) {
                gb[0] = __ldg(P.d_grad + p * 3 + 0) * scale;
                gb[1] = __ldg(P.d_grad + p * 3 + 1) * scale;
                gb[2] = __ldg(P.d_grad + p * 3 + 2) * scale;
                dsdf = __ldg(P.d_sdf + p) * scale;
            }
            if (ep.half == 0) emit_uin0(ep, x, gb, P.st_uin0, p);
            ep.signal();
            // ---------------- phase A, l = 0..7:  wbar = W_l uin_l ;  ua_bar_l = s_l * wbar  (-> uin_{l+1})
            // z2_l = softplus''(z_l) ua_l wbar_l is NOT formed here: phase B recovers it from the uin stream (which the
            // dW GEMM needs anyway) as z2 = 100 (1 - s) w (uin / s), saving one stream write + one stream read per layer.
#pragma unroll 1
            for (int l = 0; l < 8; ++l) {
                const uint8_t* st_s = P.st_in + (size_t)l * SS;      // a_l; s_l = 1 - exp(-100 a_l)
                uint8_t* st_un = P.st_uin + (size_t)l * SS;          // uin_{l+1} = ua_bar_l
                uint4 hs_n[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) hs_n[q] = ld_stream(st_s, p, ch0 + q, 32);
                if (P.thread_prefetch) prefetch_stream_chunks(st_s, p, ch0 + 2, P.thread_prefetch);
                ep.wait_acc();
                ep.sweep_half([&](int c0, const uint32_t (&v)[16]) {
                    uint4 hs_c[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) hs_c[q] = hs_n[q];
                    if (c0 < c_last) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) hs_n[q] = ld_stream(st_s, p, (c0 >> 3) + 2 + q, 32);
                    }
                    if (P.thread_prefetch && (threadIdx.x & 7) == 0) {
                        const int pc = (c0 >> 3) + 2 + P.thread_prefetch;        // chunks already covered: up to +1+pf
                        if (pc + 1 < ch0 + 16) {
                            prefetch_l2(st_s + stream_off(p, pc, 32));
                            prefetch_l2(st_s + stream_off(p, pc + 1, 32));
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int ch = (c0 >> 3) + q;
                        const uint4 hs = hs_c[q];
                        const uint32_t hsa[4] = {hs.x, hs.y, hs.z, hs.w};
                        uint32_t ub[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 av = unpack_h2(hsa[j]);
                            const float wb0 = __uint_as_float(v[q * 8 + 2 * j]), wb1 = __uint_as_float(v[q * 8 + 2 * j + 1]);
                            // s wbar = wbar - exp(-100 a) wbar
                            const float2 y = f2_mul(av, f2_splat(144.26950408889634f));
                            const float2 wb = make_float2(wb0, wb1);
                            const float2 uu2 = f2_fma(make_float2(ex2_approx(-y.x), ex2_approx(-y.y)), f2_mul(wb, f2_splat(-1.f)), wb);
                            ub[j] = pack_h2_sat(uu2.x, uu2.y);
                        }
                        const uint4 uu = make_uint4(ub[0], ub[1], ub[2], ub[3]);
                        st_stream(st_un, p, ch, 32, uu);
                        if (l < 7) ep.st_a(ch, uu);
                    }
                });
                if (l == 3 && ep.half == 1) emit_skip_ebar(ep, x, gb, st_un, p);
                if (l == 7) {
                    // A operand of phase B's first GEMM: d_feat (scaled, fp16); also streamed for dW_8
#pragma unroll 4
                    for (int k = 0; k < 16; ++k) {
                        const int ch = ch0 + k;
                        uint4 h = make_uint4(0, 0, 0, 0);
                        if (P.d_feat) {
                            float4 f0 = make_float4(0, 0, 0, 0), f1 = f0;
                            if (live) {
                                const float4* src = reinterpret_cast<const float4*>(P.d_feat + (size_t)p * 256 + ch * 8);
                                f0 = __ldg(src);
                                f1 = __ldg(src + 1);
                            }
                            h.x = pack_h2_sat(f0.x * scale, f0.y * scale); h.y = pack_h2_sat(f0.z * scale, f0.w * scale);
                            h.z = pack_h2_sat(f1.x * scale, f1.y * scale); h.w = pack_h2_sat(f1.z * scale, f1.w * scale);
                        } else if (P.d_feat16) {
                            // fp16 stream written by the albedo backward in ITS cotangent scale: rescale by the
                            // (power-of-two) ratio of the two scales
                            const uint4 u = ld_stream(P.d_feat16, p, ch, 32);
                            const float2 a0 = unpack_h2(u.x), a1 = unpack_h2(u.y), a2 = unpack_h2(u.z), a3 = unpack_h2(u.w);
                            h.x = pack_h2_sat(a0.x * dfeat_rescale, a0.y * dfeat_rescale);
                            h.y = pack_h2_sat(a1.x * dfeat_rescale, a1.y * dfeat_rescale);
                            h.z = pack_h2_sat(a2.x * dfeat_rescale, a2.y * dfeat_rescale);
                            h.w = pack_h2_sat(a3.x * dfeat_rescale, a3.y * dfeat_rescale);
                        }
                        ep.st_a(ch, h);
                        st_stream(P.st_dfeat, p, ch, 32, h);
                    }
                }
                ep.signal();
            }
            // ---------------- phase B: GEMM yields abar_l (l = 7..0);  zbar_l = s_l*abar_l + z2_l,
            //                  z2_l = softplus''(z_l) ua_l wbar_l = 100 (1 - s_l) w_l uin_{l+1} / s_l
#pragma unroll 1
            for (int l = 7; l >= 0; --l) {
                const uint8_t* st_s = P.st_in + (size_t)l * SS;      // a_l
                const uint8_t* st_w = P.st_w + (size_t)l * SS;       // w_l = s_l ua_l
                const uint8_t* st_u = P.st_uin + (size_t)l * SS;     // uin_{l+1} = s_l wbar_l
                uint8_t* st_zb = P.st_zbar + (size_t)l * SS;
                // one 16-byte chunk (8 columns) of each stream in flight ahead of the one being consumed
                uint4 hs_n = ld_stream(st_s, p, ch0, 32), hw_n = ld_stream(st_w, p, ch0, 32), hu_n = ld_stream(st_u, p, ch0, 32);
                if (P.thread_prefetch) {
                    prefetch_stream_chunks(st_s, p, ch0 + 1, P.thread_prefetch);
                    prefetch_stream_chunks(st_w, p, ch0 + 1, P.thread_prefetch);
                    prefetch_stream_chunks(st_u, p, ch0 + 1, P.thread_prefetch);
                }
                ep.wait_acc();
                ep.sweep_half([&](int c0, const uint32_t (&v)[16]) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int ch = (c0 >> 3) + q;
                        const uint4 hs = hs_n, hw = hw_n, hu = hu_n;
                        if (ch + 1 < ch0 + 16) {
                            hs_n = ld_stream(st_s, p, ch + 1, 32);
                            hw_n = ld_stream(st_w, p, ch + 1, 32);
                            hu_n = ld_stream(st_u, p, ch + 1, 32);
                        }
                        if (P.thread_prefetch && (threadIdx.x & 7) == 0 && ch + 1 + P.thread_prefetch < ch0 + 16) {
                            const size_t off = stream_off(p, ch + 1 + P.thread_prefetch, 32);
                            prefetch_l2(st_s + off);
                            prefetch_l2(st_w + off);
                            prefetch_l2(st_u + off);
                        }
                        const uint32_t hsa[4] = {hs.x, hs.y, hs.z, hs.w}, hwa[4] = {hw.x, hw.y, hw.z, hw.w},
                                       hua[4] = {hu.x, hu.y, hu.z, hu.w};
                        float ww[8];
                        if (l == 7) load_bias8(w8row + c0 + q * 8, ww);
                        uint32_t zb[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 av = unpack_h2(hsa[j]), wv = unpack_h2(hwa[j]), uv = unpack_h2(hua[j]);
                            float a0 = __uint_as_float(v[q * 8 + 2 * j]), a1 = __uint_as_float(v[q * 8 + 2 * j + 1]);
                            if (l == 7) {   // abar_7 = d_feat W_8[1:,:] + d_sdf W_8[0,:]
                                a0 = fmaf(dsdf, ww[2 * j], a0);
                                a1 = fmaf(dsdf, ww[2 * j + 1], a1);
                            }
                            const float2 zz = zbar_x2(av, wv, uv, make_float2(a0, a1));
                            zb[j] = pack_h2_sat(zz.x, zz.y);
                        }
                        const uint4 uz = make_uint4(zb[0], zb[1], zb[2], zb[3]);
                        st_stream(st_zb, p, ch, 32, uz);
                        if (l > 0) ep.st_a(ch, uz);
                    }
                });
                if (l == 3 && ep.half == 1) write_skip_cols(ep, [](int) { return 0.f; }, st_zb, p);    // zbar_3 = 0 on the PE columns
                if (l > 0) ep.signal();
            }
        }
    }
    chain_teardown(s, tmem);
}

// ======================================================================================= launchers
static inline int chain_grid(int n_tiles, int sm_count) {
    const int g = 2 * sm_count;
    return n_tiles < g ? n_tiles : g;
}

cudaError_t launch_sdf_fwd(const SdfFwdParams& P, int sm_count, cudaStream_t st) {
    const int smem = chain_smem_bytes(SDF_A_COLS);
    {
        cudaError_t e = ensure_dynamic_smem((const void*)sdf_fwd_kernel, smem);
        if (e != cudaSuccess) return e;
    }
    if (P.n_tiles == 0) return cudaSuccess;
    sdf_fwd_kernel<<<chain_grid(P.n_tiles, sm_count), CHAIN_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_sdf_fwd_grad(const SdfFwdGradParams& P, int sm_count, cudaStream_t st) {
    const int smem = chain_smem_bytes(SDF_A_COLS);
    {
        cudaError_t e = ensure_dynamic_smem((const void*)sdf_fwd_grad_kernel, smem);
        if (e != cudaSuccess) return e;
    }
    if (P.n_tiles == 0) return cudaSuccess;
    sdf_fwd_grad_kernel<<<chain_grid(P.n_tiles, sm_count), CHAIN_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_sdf_bwd_data(const SdfBwdParams& P, int sm_count, cudaStream_t st) {
    const int smem = chain_smem_bytes(SDF_A_COLS);
    {
        cudaError_t e = ensure_dynamic_smem((const void*)sdf_bwd_data_kernel, smem);
        if (e != cudaSuccess) return e;
    }
    if (P.n_tiles == 0) return cudaSuccess;
    sdf_bwd_data_kernel<<<chain_grid(P.n_tiles, sm_count), CHAIN_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace rnb
