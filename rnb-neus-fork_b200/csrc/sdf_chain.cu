// SDF MLP kernels (reference models/fields.py:82-127 and its autograd double-backward, exp_runner.py:261):
//   sdf_fwd_kernel<GRID>      K1/K8  PE + 8 hidden layers, sdf = <a_7, W_8[0,:]> + b_8[0]        (no_grad evals, grid)
//   sdf_fwd_grad_kernel       K2     forward + feature layer + analytic dx-chain, saves the streams K3 needs
//   sdf_bwd_data_kernel       K3a    adjoint of the dx-chain (phase A) + ordinary backward (phase B); writes the
//                                    cotangent streams the dW GEMM (dw_gemm.cu) contracts over the points
// All GEMMs are tcgen05.mma kind::f16 (fp16 operands, fp32 accumulate in TMEM); see chain.cuh for the CTA anatomy.
#include "chain.cuh"
#include "pe.cuh"
#include "sdf_params.h"
#include "dw_params.h"
#include "dw_common.cuh"

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

// epilogue of K1 for the tiles of one (virtual) CTA
__device__ __forceinline__ void sdf_fwd_epilogue(Epi& ep, const SdfFwdParams& P, const TileMap tm) {
    const float* bias = P.aux;
    const float* w8row = P.aux + AUX_W8ROW;
    const float b8 = __ldg(P.aux + AUX_B8_0);
    for (int t = 0; t < tm.n; ++t) {
        const int64_t p = (tm.first + (int64_t)t * tm.stride) * TILE_M + ep.row;
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

__global__ void __launch_bounds__(CHAIN_THREADS, 2) sdf_fwd_kernel(const __grid_constant__ SdfFwdParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const ChainSmem s = chain_carve(smem, SDF_A_COLS);
    const uint32_t tmem = chain_setup(s);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TileMap tm = tilemap_grid(P.n_tiles);
    if (warp == 0) {
        if (lane == 0) chain_producer(s, P.tab, P.wblob, tm);
    } else if (warp == 1) {
        chain_mma_warp(s, P.tab, tmem, tm.n);
    } else {
        Epi ep;
        ep.init(s, tmem);
#ifdef RNB_TRACE
        ep.tab_trace = P.tab.trace;
#endif
        sdf_fwd_epilogue(ep, P, tm);
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
        chain_mma_warp(s, P.tab, tmem, n_my);
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
            const size_t prow = stream_row(p);
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
                uint8_t* ran = st_a_next + prow;
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
                        st_stream_r(ran, ch, ha, keep_l);
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
                uint8_t* rfeat = P.st_feat + prow;
                uint8_t* rw7 = st_w7 + prow;
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
                        st_stream_r(rfeat, ch, hf);
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
                        if (P.st_w) st_stream_r(rw7, ch, hw);
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
                const uint8_t* rsp = st_sp + prow;
                uint8_t* rwp = st_wp + prow;
                uint4 hs_n[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) hs_n[q] = ld_stream_r(rsp, (ep.col0 >> 3) + q);
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
                        for (int q = 0; q < 2; ++q) hs_n[q] = ld_stream_r(rsp, (c0 >> 3) + 2 + q);
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
                        if (P.st_w) st_stream_r(rwp, ch, hw);
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

// Epilogue role of the backward chain (warps 2..9 of a (virtual) CTA) over the tiles of `tm`.  publish(ep, layer, tile) is
// called by all 256 epilogue threads once every stream the weight gradient of `layer` contracts is in memory.
struct NoPublish {
    __device__ __forceinline__ void operator()(const Epi&, int, int64_t) const {}
};
template <class Pub>
__device__ __forceinline__ void bwd_epilogue(Epi& ep, const SdfBwdParams& P, const TileMap tm, Pub&& publish) {
    const float* w8row = P.aux + AUX_W8ROW;
    const size_t SS = P.stream_stride;
    const float scale = cot_scale_from_max(__ldg(P.cot_absmax));
    const float dfeat_rescale = P.d_feat16 ? scale / cot_scale_from_max(__ldg(P.d_feat16_cot_absmax)) : 0.f;
    const int c_last = ep.col0 + EPI_HALF_COLS - 16;
    const int ch0 = ep.col0 >> 3;
    for (int t = 0; t < tm.n; ++t) {
        const int64_t tile = tm.first + (int64_t)t * tm.stride;
        const int64_t p = tile * TILE_M + ep.row;
        const size_t prow = stream_row(p);
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
            const uint8_t* rs = st_s + prow;
            uint8_t* run = st_un + prow;
            uint4 hs_n[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) hs_n[q] = ld_stream_r(rs, ch0 + q);
            if (P.thread_prefetch) prefetch_stream_chunks(st_s, p, ch0 + 2, P.thread_prefetch);
            ep.wait_acc();
            ep.sweep_half([&](int c0, const uint32_t (&v)[16]) {
                uint4 hs_c[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) hs_c[q] = hs_n[q];
                if (c0 < c_last) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) hs_n[q] = ld_stream_r(rs, (c0 >> 3) + 2 + q);
                }
                if (P.thread_prefetch && (threadIdx.x & 7) == 0) {
                    const int pc = (c0 >> 3) + 2 + P.thread_prefetch;        // chunks already covered: up to +1+pf
                    if (pc + 1 < ch0 + 16) {
                        prefetch_l2(rs + (size_t)pc * 1024);
                        prefetch_l2(rs + (size_t)(pc + 1) * 1024);
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
                    st_stream_r(run, ch, uu, P.keep_streams != 0);
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
            if (l == 7) publish(ep, 8, tile);      // d_feat and uin_8 are in memory: layer 8's weight gradient can be formed
        }
        // ---------------- phase B: GEMM yields abar_l (l = 7..0);  zbar_l = s_l*abar_l + z2_l,
        //                  z2_l = softplus''(z_l) ua_l wbar_l = 100 (1 - s_l) w_l uin_{l+1} / s_l
#pragma unroll 1
        for (int l = 7; l >= 0; --l) {
            const uint8_t* st_s = P.st_in + (size_t)l * SS;      // a_l
            const uint8_t* st_w = P.st_w + (size_t)l * SS;       // w_l = s_l ua_l
            const uint8_t* st_u = P.st_uin + (size_t)l * SS;     // uin_{l+1} = s_l wbar_l
            uint8_t* st_zb = P.st_zbar + (size_t)l * SS;
            const uint8_t* rs = st_s + prow;
            const uint8_t* rw = st_w + prow;
            const uint8_t* ru = st_u + prow;
            uint8_t* rzb = st_zb + prow;
            // one 16-byte chunk (8 columns) of each stream in flight ahead of the one being consumed
            uint4 hs_n = ld_stream_r(rs, ch0), hw_n = ld_stream_r(rw, ch0), hu_n = ld_stream_r(ru, ch0);
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
                        hs_n = ld_stream_r(rs, ch + 1);
                        hw_n = ld_stream_r(rw, ch + 1);
                        hu_n = ld_stream_r(ru, ch + 1);
                    }
                    if (P.thread_prefetch && (threadIdx.x & 7) == 0 && ch + 1 + P.thread_prefetch < ch0 + 16) {
                        const size_t off = (size_t)(ch + 1 + P.thread_prefetch) * 1024;
                        prefetch_l2(rs + off);
                        prefetch_l2(rw + off);
                        prefetch_l2(ru + off);
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
                    st_stream_r(rzb, ch, uz, P.keep_streams != 0);
                    if (l > 0) ep.st_a(ch, uz);
                }
            });
            if (l == 3 && ep.half == 1) write_skip_cols(ep, [](int) { return 0.f; }, st_zb, p);    // zbar_3 = 0 on the PE columns
            if (l > 0) ep.signal();
            publish(ep, l, tile);                  // zbar_l is in memory: both terms of dW_l can be formed
        }
    }
}

__global__ void __launch_bounds__(CHAIN_THREADS, 2) sdf_bwd_data_kernel(const __grid_constant__ SdfBwdParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const ChainSmem s = chain_carve(smem, SDF_A_COLS);
    const uint32_t tmem = chain_setup(s);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TileMap tm = tilemap_grid(P.n_tiles);
    if (warp == 0) {
        if (lane == 0) chain_producer(s, P.tab, P.wblob, tm);
    } else if (warp == 1) {
        chain_mma_warp(s, P.tab, tmem, tm.n);
    } else {
        Epi ep;
        ep.init(s, tmem);
#ifdef RNB_TRACE
        ep.tab_trace = P.tab.trace;
#endif
        bwd_epilogue(ep, P, tm, NoPublish());
    }
    chain_teardown(s, tmem);
}

// ======================================================================================= K3 fused
// Backward chain (K3a) and the weight-gradient contraction (K3b) in ONE launch, one 640-thread CTA per SM:
//   blocks [0, n_dw)     weight-gradient workers: one layer each, the 256 x N fp32 accumulator stays in TMEM (2 x 256
//                        columns) for the whole launch; they take point tiles from a per-layer queue in the order the
//                        chain completes them and pull the four operand tiles (w_l, uin_l, zbar_l, in_l) with cp.async.bulk
//                        while those lines are still in L2 -- the cotangent streams are written once and read back
//                        from L2, not from HBM (the separate dW kernel re-read 17.6 KB/point from DRAM).
//   blocks [n_dw, grid)  two "virtual" chain CTAs each (the anatomy of chain.cuh twice: own shared-memory region, own
//                        256 TMEM columns, own named barrier), running bwd_epilogue() and publishing (layer, tile) to the
//                        queues.  The chain never waits for a worker, so no placement of the blocks can deadlock.
// Memory ordering of a hand-over: stream stores (generic proxy) -> bar.sync of the 256 epilogue threads -> one thread:
// fence.acq_rel.gpu + st.release of the queue entry  ||  worker: ld.acquire of the entry -> fence.proxy.async ->
// cp.async.bulk reads (async proxy).
constexpr int FUSED_THREADS = 2 * CHAIN_THREADS;
constexpr int FUSED_VSMEM = (chain_smem_bytes(SDF_A_COLS) + 1023) / 1024 * 1024;     // one virtual CTA's region
constexpr int FDW_STAGES = 3;
constexpr int FDW_FIFO = 8;                                                        // scout -> producer tile FIFO
constexpr int FDW_STAGE_A = 32768, FDW_STAGE = 65536;                               // 64 points x 256 columns fp16, twice
constexpr int FUSED_SMEM = 2 * FUSED_VSMEM > FDW_STAGES * FDW_STAGE + 256 ? 2 * FUSED_VSMEM : FDW_STAGES * FDW_STAGE + 256;

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// No rendezvous of the epilogue threads: every warp fences its own stream stores and bumps the (tile, layer) counter;
// the warp that completes the count of eight pushes the tile (the "last block" pattern, fence cumulativity).
struct QueuePublish {
    int* q;
    int* tail;
    int* cnt;          // [n_tiles][9] warps that have finished (tile, layer)
    int n_tiles;
    __device__ __forceinline__ void operator()(const Epi&, int layer, int64_t tile) const {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
            __threadfence();
            if (atomicAdd(cnt + tile * 9 + layer, 1) == EPI_THREADS / 32 - 1) {
                __threadfence();
                const int pos = atomicAdd(tail + layer, 1);
                st_release_gpu(q + (size_t)layer * n_tiles + pos, (int)tile);
            }
        }
    }
};

__device__ __forceinline__ void fused_dw_worker(uint8_t* smem, const SdfBwdFusedParams& P, uint32_t tmem) {
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FDW_STAGES * FDW_STAGE);
    uint64_t* full = bars;
    uint64_t* empty = bars + FDW_STAGES;
    uint64_t* acc_full = bars + 2 * FDW_STAGES;
    volatile int* stage_stop = reinterpret_cast<volatile int*>(bars + 2 * FDW_STAGES + 1);    // [FDW_STAGES], zeroed at setup
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int layer = P.dw_layer[blockIdx.x], rep = P.dw_replica[blockIdx.x];
    const DwJob& job = P.jobs[layer];
    const int nw = job.nw;
    const uint32_t b_bytes = (uint32_t)(nw >> 3) * 1024u;
    const int n_tiles = P.chain.n_tiles;
    // warp 6 ("scout") takes tickets and waits for the queue entries; ready tile indices reach the copy-issuing lane
    // through a small shared-memory FIFO, so the L2 round trips of the ticket / the entry poll never sit between two copies
    volatile int* fifo = stage_stop + 4;              // [FDW_FIFO] tile indices, -1 = the queue is exhausted
    volatile int* fifo_tail = stage_stop + 4 + FDW_FIFO;      // entries written (scout)
    volatile int* fifo_head = fifo_tail + 1;                  // entries consumed (producer)
    volatile int* stage_sub = fifo_head + 1;                  // [FDW_STAGES] 64-point sub-tile staged in each slot
    if (warp == 6) {
        if (lane == 0) {
            int tail = 0;
            for (;;) {
                const int ticket = atomicAdd(P.q_head + layer, 1);
                int tile = -1;
                if (ticket < n_tiles) {
                    const int* entry = P.q + (size_t)layer * n_tiles + ticket;
                    const unsigned long long t0 = globaltimer_ns();
                    while ((tile = ld_acquire_gpu(entry)) < 0) {
                        __nanosleep(128);
                        if (globaltimer_ns() - t0 > 20000000000ull) __trap();      // 20 s: a lost hand-over must fail, not hang
                    }
                }
                while (tail - *fifo_head >= FDW_FIFO) __nanosleep(64);
                fifo[tail % FDW_FIFO] = tile;
                __threadfence_block();
                *fifo_tail = ++tail;
                if (tile < 0) break;
            }
        }
    } else if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            int head = 0;
            // the workers are the last readers of the cotangent streams: their lines may leave L2 first
            uint64_t drop;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(drop));
            for (;;) {
                while (*fifo_tail <= head) __nanosleep(32);
                __threadfence_block();
                const int tile = fifo[head % FDW_FIFO];
                *fifo_head = ++head;
                if (tile < 0) {
                    // queue exhausted: a stop sentinel through the ring
                    const uint32_t slot = it % FDW_STAGES, ph = (it / FDW_STAGES) & 1;
                    mbar_wait(&empty[slot], ph ^ 1);
                    stage_stop[slot] = 1;
                    mbar_arrive(&full[slot]);
                    break;
                }
                // the entry was acquired by the scout (same CTA); order this lane's async-proxy reads behind it
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                asm volatile("fence.proxy.async;" ::: "memory");
                for (int sub = 2 * tile; sub < 2 * tile + 2; ++sub)
                    for (int pr = 0; pr < job.n_pairs; ++pr, ++it) {
                        const uint32_t slot = it % FDW_STAGES, ph = (it / FDW_STAGES) & 1;
                        const bool mma = pr < job.mma_pairs;
                        dw_prefetch_weights(job, pr, sub);
                        mbar_wait(&empty[slot], ph ^ 1);
                        stage_sub[slot] = sub;
                        const uint8_t* rider = job.x[pr];
                        mbar_expect_tx(&full[slot], FDW_STAGE_A + (mma ? b_bytes : 0u) + (rider ? (uint32_t)DWC_RIDER_BYTES : 0u));
                        uint8_t* dst = smem + slot * FDW_STAGE;
                        bulk_g2s_hint(dst, job.a[pr] + (size_t)sub * FDW_STAGE_A, FDW_STAGE_A, &full[slot], drop);
                        if (mma)
                            bulk_g2s_hint(dst + FDW_STAGE_A, job.b[pr] + ((size_t)sub * job.b_chunks[pr] + job.b_chunk0) * 1024u,
                                          b_bytes, &full[slot], drop);
                        if (rider)
                            bulk_g2s_hint(dst + FDW_STAGE_A + b_bytes, rider + ((size_t)sub * 32 + job.x_chunk0[pr]) * 1024u,
                                          DWC_RIDER_BYTES, &full[slot], drop);
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(128, nw, FMT_F16, FMT_F16, MAJOR_MN, MAJOR_MN);
            uint32_t n_mma = 0;
            for (uint32_t it = 0;; ++it) {
                const uint32_t slot = it % FDW_STAGES, ph = (it / FDW_STAGES) & 1;
                mbar_wait(&full[slot], ph);
                if (stage_stop[slot]) break;
                if ((int)(it % (uint32_t)job.n_pairs) >= job.mma_pairs) {       // staged for the column sums only
                    mbar_arrive(&empty[slot]);
                    continue;
                }
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + slot * FDW_STAGE), sb = sa + FDW_STAGE_A;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t ad = umma_desc(sa + h * 16384 + ks * 256, 128, 1024);
                        const uint64_t bd = umma_desc(sb + ks * 256, 128, 1024);
                        umma_f16(tmem + h * 256, ad, bd, idesc, (n_mma | ks) != 0);
                    }
                ++n_mma;
                umma_commit(&empty[slot]);
            }
            umma_commit(acc_full);
        }
    } else if (warp != 6 && warp < 11) {
        // the job's column sums (bias gradients, the sdf row of W_8) from the staged tiles (dw_common.cuh): warps 2-5 take the
        // point rows 0-31 of every sub-tile, warps 7-10 rows 32-63 (warp 6 is the scout); afterwards each group dumps one
        // 128-row half of the accumulator
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int grp = warp > 6 ? 1 : 0;
        const int et = ((warp - (grp ? 7 : 2)) & 3) * 32 + lane;            // 0..127: owner of columns 2 et, 2 et + 1
        DwColsumAcc cs;
        cs.clear();
        uint32_t it = 0;
        // the scout does not tell these warps which tile a stage belongs to: the producer leaves the sub-tile index of
        // every stage next to the stop flags
        for (;; ++it) {
            const uint32_t slot = it % FDW_STAGES, ph = (it / FDW_STAGES) & 1;
            mbar_wait(&full[slot], ph);
            if (stage_stop[slot]) break;
            cs.stage<32>(job, (int)(it % (uint32_t)job.n_pairs), stage_sub[slot], smem + slot * FDW_STAGE, et,
                         reinterpret_cast<float*>(smem + FDW_STAGES * FDW_STAGE + 256) + (warp - 2) * DWC_WBUF_FLOATS, grp * 32);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
        float* out = job.partial + (size_t)rep * 256 * nw;
        if (it > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const int h = grp;
#pragma unroll 1
            for (int c0 = 0; c0 < nw; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + h * 256 + c0, v);
                tmem_ld_wait();
                float4* dst = reinterpret_cast<float4*>(out + (size_t)(h * 128 + row) * nw + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (c0 + 4 * j < nw)
                        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                             __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            }
        } else {
            for (int c = 0; c < nw; ++c) out[(size_t)(grp * 128 + row) * nw + c] = 0.f;
        }
        cs.store(job, 2 * rep + grp, et);
    }
}

__global__ void __launch_bounds__(FUSED_THREADS, 1) sdf_bwd_fused_kernel(const __grid_constant__ SdfBwdFusedParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_dw = (int)blockIdx.x < P.n_dw;
    const int v = warp >= CHAIN_THREADS / 32 ? 1 : 0;          // virtual chain CTA of this warp
    const int lw = warp - v * (CHAIN_THREADS / 32);            // warp index inside it
    const ChainSmem s = chain_carve(smem + v * FUSED_VSMEM, SDF_A_COLS);
    if (is_dw) {
        if (threadIdx.x == 0) {
            uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FDW_STAGES * FDW_STAGE);
            for (int i = 0; i < FDW_STAGES; ++i) { mbar_init(&bars[i], 1); mbar_init(&bars[FDW_STAGES + i], 1 + 8); }
            mbar_init(&bars[2 * FDW_STAGES], 1);
            int* stop = reinterpret_cast<int*>(bars + 2 * FDW_STAGES + 1);
            for (int i = 0; i < 4 + FDW_FIFO + 2 + FDW_STAGES; ++i) stop[i] = 0;   // stage_stop[4], fifo[], tail, head, stage_sub[]
            mbar_fence_init();
        }
    } else if (lw == 0 && lane == 0) {
        for (int i = 0; i < RING_STAGES; ++i) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], 1);
        }
        mbar_init(s.acc_full, 1);
        mbar_init(s.a_ready, EPI_THREADS);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (P.dbg && threadIdx.x == 0) P.dbg[FUSED_MAX_WORKERS + blockIdx.x] = globaltimer_ns();
    if (is_dw) {
        if (warp < 11) fused_dw_worker(smem, P, tmem);
    } else {
        const int n_v = 2 * ((int)gridDim.x - P.n_dw);
        TileMap tm;
        tm.first = ((int64_t)blockIdx.x - P.n_dw) * 2 + v;
        tm.stride = n_v;
        tm.n = tm.first < P.chain.n_tiles ? (P.chain.n_tiles - (int)tm.first + n_v - 1) / n_v : 0;
        const uint32_t tm_v = tmem + (uint32_t)(v * TMEM_COLS);
        if (lw == 0) {
            if (lane == 0) chain_producer(s, P.chain.tab, P.chain.wblob, tm);
        } else if (lw == 1) {
            chain_mma_warp(s, P.chain.tab, tm_v, tm.n);
        } else {
            Epi ep;
            ep.init(s, tm_v, lw, 1 + v);
            // Staggered start: virtual CTA g begins g / n_v of `stagger_ns` (about one tile time) late, so the hand-overs
            // reach every layer's queue as a steady trickle the workers can consume while the lines are still in L2, not
            // as one burst of n_v tiles per layer.  The CTAs with the highest delays own one tile less (n_tiles % n_v).
            if (P.stagger_ns > 0) {
                const unsigned long long until = globaltimer_ns() + (unsigned long long)P.stagger_ns * (unsigned long long)tm.first / (unsigned long long)n_v;
                while (globaltimer_ns() < until) __nanosleep(512);
            }
            QueuePublish pub{P.q, P.q_tail, P.q_cnt, P.chain.n_tiles};
            bwd_epilogue(ep, P.chain, tm, pub);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
    if (P.dbg && threadIdx.x == 0) P.dbg[blockIdx.x] = globaltimer_ns();
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
#if defined(RNB_TRACE) || defined(RNB_DBG_HOOKS)
    if (getenv("RNB_DBG_SOLO")) {      // experiment: one CTA per SM (shared memory request too large for two)
        const int big = 120 * 1024;
        cudaFuncSetAttribute((const void*)sdf_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        sdf_fwd_kernel<<<sm_count, CHAIN_THREADS, big, st>>>(P);
        return cudaGetLastError();
    }
#endif
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


cudaError_t launch_sdf_bwd_fused(const SdfBwdFusedParams& P, int sm_count, cudaStream_t st) {
    {
        cudaError_t e = ensure_dynamic_smem((const void*)sdf_bwd_fused_kernel, FUSED_SMEM);
        if (e != cudaSuccess) return e;
    }
    if (P.chain.n_tiles == 0) return cudaSuccess;
    int chain_blocks = sm_count - P.n_dw;
    const int need = (P.chain.n_tiles + 1) / 2;
    if (chain_blocks > need) chain_blocks = need;
    if (chain_blocks < 1) chain_blocks = 1;
    sdf_bwd_fused_kernel<<<P.n_dw + chain_blocks, FUSED_THREADS, FUSED_SMEM, st>>>(P);
    return cudaGetLastError();
}

}  // namespace rnb
