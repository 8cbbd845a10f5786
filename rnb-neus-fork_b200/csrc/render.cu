// Per-ray kernels of NeuSRenderer (reference models/renderer.py): one warp per ray, shuffle scans, no atomics.
//   coarse_z_kernel      uniform samples + jitter                    (renderer.py:829-845)
//   upsample_kernel      merge previous new samples (cat_z_vals :178-192) + up_sample :132-176 + sample_pdf :39-69
//   final_merge_kernel   last cat_z_vals + section mid-points        (renderer.py:479-484)
//   composite_fwd/bwd    render_core_mvps after the networks + RNb shading sum + adjoint (:503-540, :904-918, :1008-1017)
// HBM-bound, a few KB per ray; the per-ray state lives in registers / shared memory.
#include "common.cuh"
#include "render_params.h"
#include <cstdio>

namespace rnb {

constexpr int RAYS_PER_BLOCK = 4;          // 4 warps, one ray each
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }
// MUFU-only forms for the adjoint kernel (no IEEE division / sqrt subroutine calls): ex2.approx, rcp.approx, rsqrt.approx
// are accurate to ~2 ulp, four orders of magnitude inside the 2e-3 the adjoint is checked at
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrt_fast(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoidf_fast(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    return rcp_fast(1.f + e);
}

// torch.linspace(start, end, steps) float32 arithmetic (ATen: start + step*i below the midpoint, end - step*(steps-1-i) above)
__device__ __forceinline__ float linspace_at(float start, float end, int steps, int i) {
    const float step = __fdiv_rn(__fsub_rn(end, start), (float)(steps - 1));
    return i < steps / 2 ? __fadd_rn(start, __fmul_rn(step, (float)i)) : __fsub_rn(end, __fmul_rn(step, (float)(steps - 1 - i)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
// inclusive scans across the warp
__device__ __forceinline__ float warp_scan_add(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v *= t;
    }
    return v;
}

// ------------------------------------------------------------------------------------------------ A8
__global__ void coarse_z_kernel(const float* near, const float* far, const float* t_rand, float* z, int n_rays,
                                int n_samples) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays * n_samples) return;
    const int ray = i / n_samples, s = i % n_samples;
    const float lin = linspace_at(0.f, 1.f, n_samples, s);
    const float nr = near[ray], fr = far[ray];
    float v = __fadd_rn(nr, __fmul_rn(__fsub_rn(fr, nr), lin));
    if (t_rand) v = __fadd_rn(v, __fdiv_rn(__fmul_rn(t_rand[ray], 2.0f), (float)n_samples));
    z[i] = v;
}

// ------------------------------------------------------------------------------------------------ A9-A11
// searchsorted(cdf[0..n), u, right=True): first index with cdf[idx] > u  (n if none)
__device__ __forceinline__ int searchsorted_right(const float* cdf, int n, float u) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// inverse-CDF sample j of n_new from (bins[0..n), cdf[0..n))  -- sample_pdf det=True, renderer.py:48-67
__device__ __forceinline__ float sample_from_cdf(const float* bins, const float* cdf, int n, int n_new, int j, int* ind_out) {
    const float u = linspace_at(0.5f / n_new, 1.f - 0.5f / n_new, n_new, j);
    const int ind = searchsorted_right(cdf, n, u);
    if (ind_out) *ind_out = ind;
    const int below = max(ind - 1, 0), above = min(ind, n - 1);
    float denom = cdf[above] - cdf[below];
    if (denom < 1e-5f) denom = 1.f;
    const float t = (u - cdf[below]) / denom;
    return bins[below] + t * (bins[above] - bins[below]);
}

struct RaySmem {
    float z[MAX_RAY_SAMPLES];
    float sdf[MAX_RAY_SAMPLES];
    float cdf[MAX_RAY_SAMPLES + 4];
};

// merge sorted (z_old, sdf_old)[n_old] with sorted (z_new, sdf_new)[n_new] into smem; ties keep old first
__device__ __forceinline__ void merge_into(RaySmem& S, const float* z_old, const float* sdf_old, int n_old, const float* z_new,
                                           const float* sdf_new, int n_new, int lane) {
    for (int i = lane; i < n_old; i += 32) {
        const float v = z_old[i];
        int cnt = 0;
        for (int j = 0; j < n_new; ++j) cnt += z_new[j] < v;
        S.z[i + cnt] = v;
        if (sdf_old) S.sdf[i + cnt] = sdf_old[i];
    }
    for (int j = lane; j < n_new; j += 32) {
        const float v = z_new[j];
        int lo = 0, hi = n_old;                       // count of old <= v
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (z_old[mid] <= v) lo = mid + 1; else hi = mid;
        }
        S.z[j + lo] = v;
        if (sdf_new) S.sdf[j + lo] = sdf_new[j];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK) upsample_kernel(const __grid_constant__ UpsampleParams P) {
    __shared__ RaySmem smem[RAYS_PER_BLOCK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ray = blockIdx.x * RAYS_PER_BLOCK + warp;
    if (ray >= P.n_rays) return;
    RaySmem& S = smem[warp];
    const int n = P.n_old + P.n_merge;                       // samples after merging the pending ones
    if (P.n_merge > 0) {
        merge_into(S, P.z_old + (size_t)ray * P.n_old, P.sdf_old + (size_t)ray * P.n_old, P.n_old,
                   P.z_pending + (size_t)ray * P.n_merge, P.sdf_pending + (size_t)ray * P.n_merge, P.n_merge, lane);
    } else {
        for (int i = lane; i < n; i += 32) {
            S.z[i] = P.z_old[(size_t)ray * P.n_old + i];
            S.sdf[i] = P.sdf_old[(size_t)ray * P.n_old + i];
        }
        __syncwarp();
    }
    if (P.z_merged) {
        for (int i = lane; i < n; i += 32) {
            P.z_merged[(size_t)ray * n + i] = S.z[i];
            P.sdf_merged[(size_t)ray * n + i] = S.sdf[i];
        }
    }
    // ---- up_sample: section weights.  Lane owns sections 4*lane .. 4*lane+3 (n-1 <= 127 sections)
    const float ox = P.rays_o[ray * 3], oy = P.rays_o[ray * 3 + 1], oz = P.rays_o[ray * 3 + 2];
    const float dx = P.rays_d[ray * 3], dy = P.rays_d[ray * 3 + 1], dz = P.rays_d[ray * 3 + 2];
    const int m = n - 1;
    float alpha[4], w[4];
    float tprod = 1.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = 4 * lane + k;
        float a = 0.f;
        if (i < m) {
            const float z0 = S.z[i], z1 = S.z[i + 1], s0 = S.sdf[i], s1 = S.sdf[i + 1];
            const float px0 = ox + dx * z0, py0 = oy + dy * z0, pz0 = oz + dz * z0;
            const float px1 = ox + dx * z1, py1 = oy + dy * z1, pz1 = oz + dz * z1;
            const float r0 = sqrtf(px0 * px0 + py0 * py0 + pz0 * pz0), r1 = sqrtf(px1 * px1 + py1 * py1 + pz1 * pz1);
            const bool inside = (r0 < 1.f) || (r1 < 1.f);
            const float cosv = (s1 - s0) / (z1 - z0 + 1e-5f);
            float prev = 0.f;
            if (i > 0) prev = (s0 - S.sdf[i - 1]) / (z0 - S.z[i - 1] + 1e-5f);
            float c = fminf(prev, cosv);
            c = fminf(fmaxf(c, -1e3f), 0.f) * (inside ? 1.f : 0.f);
            const float dist = z1 - z0, mid = (s0 + s1) * 0.5f;
            const float pe = mid - c * dist * 0.5f, ne = mid + c * dist * 0.5f;
            const float pc = sigmoidf_acc(pe * P.inv_s), nc = sigmoidf_acc(ne * P.inv_s);
            a = (pc - nc + 1e-5f) / (pc + 1e-5f);
        }
        alpha[k] = a;
    }
    // exclusive cumprod of (1 - alpha + 1e-7)
    {
        float f[4], run = 1.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { f[k] = (4 * lane + k < m) ? (1.f - alpha[k] + 1e-7f) : 1.f; run *= f[k]; }
        const float incl = warp_scan_mul(run, lane);
        float excl = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) excl = 1.f;
        tprod = excl;
#pragma unroll
        for (int k = 0; k < 4; ++k) { w[k] = alpha[k] * tprod + 1e-5f; tprod *= f[k]; }
    }
    // pdf / cdf (sample_pdf)
    float lsum = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (4 * lane + k >= m) w[k] = 0.f; lsum += w[k]; }
    const float total = warp_sum(lsum);
    float run = 0.f, p[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { p[k] = w[k] / total; run += p[k]; }
    const float incl = warp_scan_add(run, lane);
    float base = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) { base = 0.f; S.cdf[0] = 0.f; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        base += p[k];
        if (4 * lane + k < m) S.cdf[4 * lane + k + 1] = base;
    }
    __syncwarp();
    if (lane < P.n_new) {
        int ind;
        const float zs = sample_from_cdf(S.z, S.cdf, n, P.n_new, lane, &ind);
        P.z_new[(size_t)ray * P.n_new + lane] = zs;
        if (P.inds) P.inds[(size_t)ray * P.n_new + lane] = ind;
    }
    if (P.cdf_out) for (int i = lane; i < n; i += 32) P.cdf_out[(size_t)ray * n + i] = S.cdf[i];
}

// test hook: searchsorted indices + samples from a caller-supplied CDF (bit-exact index parity)
__global__ void sample_pdf_from_cdf_kernel(const float* bins, const float* cdf, int n_rays, int n, int n_new, float* samples,
                                           int64_t* inds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays * n_new) return;
    const int ray = i / n_new, j = i % n_new;
    int ind;
    samples[i] = sample_from_cdf(bins + (size_t)ray * n, cdf + (size_t)ray * n, n, n_new, j, &ind);
    inds[i] = ind;
}

__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK) final_merge_kernel(const float* z_old, int n_old, const float* z_new,
                                                                          int n_new, int n_rays, float sample_dist,
                                                                          float* z_out, float* mid_out) {
    __shared__ RaySmem smem[RAYS_PER_BLOCK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ray = blockIdx.x * RAYS_PER_BLOCK + warp;
    if (ray >= n_rays) return;
    RaySmem& S = smem[warp];
    const int n = n_old + n_new;
    if (n_new > 0) {
        merge_into(S, z_old + (size_t)ray * n_old, nullptr, n_old, z_new + (size_t)ray * n_new, nullptr, n_new, lane);
    } else {
        for (int i = lane; i < n; i += 32) S.z[i] = z_old[(size_t)ray * n_old + i];
        __syncwarp();
    }
    for (int i = lane; i < n; i += 32) {
        const float z = S.z[i];
        const float dist = i + 1 < n ? S.z[i + 1] - z : sample_dist;
        z_out[(size_t)ray * n + i] = z;
        mid_out[(size_t)ray * n + i] = z + dist * 0.5f;
    }
}

// ------------------------------------------------------------------------------------------------ A12-A13
struct SampleFw {
    float dist, alpha, alpha_raw, pc, nc, T, w, tc, ic, gn, relax, inside;
    float rq, rgn;      // FAST only: 1 / (pc + 1e-5), 1 / |g|
};

// per-sample forward quantities of render_core_mvps (renderer.py:503-540); T and w are filled by the caller
// FAST (the adjoint kernel): the same quantities from MUFU approximations; the forward kernels keep the IEEE forms their
// outputs (weights, cdf, inside_sphere) are pinned with.
template <bool FAST = false>
__device__ __forceinline__ void sample_forward(float sdf, const float (&g)[3], const float (&d)[3], const float (&pt)[3],
                                               float dist, float inv_s, float r, SampleFw& f) {
    f.dist = dist;
    f.tc = d[0] * g[0] + d[1] * g[1] + d[2] * g[2];
    f.ic = -(fmaxf(-f.tc * 0.5f + 0.5f, 0.f) * (1.f - r) + fmaxf(-f.tc, 0.f) * r);
    const float en = sdf + f.ic * dist * 0.5f, ep = sdf - f.ic * dist * 0.5f;
    const float p2 = pt[0] * pt[0] + pt[1] * pt[1] + pt[2] * pt[2];
    const float g2 = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
    if (FAST) {
        f.pc = sigmoidf_fast(ep * inv_s);
        f.nc = sigmoidf_fast(en * inv_s);
        f.rq = rcp_fast(f.pc + 1e-5f);
        f.alpha_raw = (f.pc - f.nc + 1e-5f) * f.rq;
        f.relax = p2 < 1.44f ? 1.f : 0.f;
        f.inside = 0.f;
        f.rgn = g2 > 0.f ? rsqrt_fast(g2) : 0.f;
        f.gn = g2 * f.rgn;
    } else {
        f.pc = sigmoidf_acc(ep * inv_s);
        f.nc = sigmoidf_acc(en * inv_s);
        f.alpha_raw = (f.pc - f.nc + 1e-5f) / (f.pc + 1e-5f);
        const float pn = sqrtf(p2);
        f.inside = pn < 1.f ? 1.f : 0.f;
        f.relax = pn < 1.2f ? 1.f : 0.f;
        f.gn = sqrtf(g2);
    }
    f.alpha = fminf(fmaxf(f.alpha_raw, 0.f), 1.f);
}

#ifndef RNB_COMP_BLOCKS
#define RNB_COMP_BLOCKS 5
#endif
template <bool BWD>
__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK, RNB_COMP_BLOCKS) composite_kernel(const __grid_constant__ CompositeParams P) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ray = blockIdx.x * RAYS_PER_BLOCK + warp;
    if (ray >= P.n_rays) return;
    constexpr int NS = 128, PER = 4;                       // 128 fine samples, lane owns 4 consecutive ones
    const float inv_s = fminf(fmaxf(expf(__ldg(P.variance) * 10.f), 1e-6f), 1e6f);     // fields.py:323-325, renderer.py:503
    const float o[3] = {P.rays_o[ray * 3], P.rays_o[ray * 3 + 1], P.rays_o[ray * 3 + 2]};
    const float d[3] = {P.rays_d[ray * 3], P.rays_d[ray * 3 + 1], P.rays_d[ray * 3 + 2]};
    const size_t base = (size_t)ray * NS + lane * PER;
    const float4 z4 = *reinterpret_cast<const float4*>(P.z + base);
    const float4 s4 = *reinterpret_cast<const float4*>(P.sdf + base);
    const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
    const float sd[4] = {s4.x, s4.y, s4.z, s4.w};
    float g[4][3], al[4][3];
    {
        const float4* gp = reinterpret_cast<const float4*>(P.grad + base * 3);
        const float4 a = gp[0], b = gp[1], c = gp[2];
        const float t[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 3; ++j) g[k][j] = t[k * 3 + j];
    }
    if (P.albedo) {
        const float4* ap = reinterpret_cast<const float4*>(P.albedo + base * 3);
        const float4 a = ap[0], b = ap[1], c = ap[2];
        const float t[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 3; ++j) al[k][j] = t[k * 3 + j];
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) al[k][0] = al[k][1] = al[k][2] = 1.f;                 // no_albedo: renderer.py:905-906
    }
    float znext = __shfl_down_sync(FULL, zz[0], 1);
    SampleFw f[4];
    float run = 1.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float zn = k < 3 ? zz[k + 1] : znext;
        const float dist = (lane == 31 && k == 3) ? P.sample_dist : zn - zz[k];
        const float mid = zz[k] + dist * 0.5f;
        const float pt[3] = {o[0] + d[0] * mid, o[1] + d[1] * mid, o[2] + d[2] * mid};
        sample_forward<BWD>(sd[k], g[k], d, pt, dist, inv_s, P.cos_anneal_ratio, f[k]);
        run *= 1.f - f[k].alpha + 1e-7f;
    }
    {
        const float incl = warp_scan_mul(run, lane);
        float T = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) T = 1.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            f[k].T = T;
            f[k].w = f[k].alpha * T;
            T *= 1.f - f[k].alpha + 1e-7f;
        }
    }
    if (!BWD) {
        float wsum = 0.f, wmax = 0.f, en = 0.f, ed = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            wsum += f[k].w;
            wmax = fmaxf(wmax, f[k].w);
            en += f[k].relax * (f[k].gn - 1.f) * (f[k].gn - 1.f);
            ed += f[k].relax;
        }
        wsum = warp_sum(wsum); wmax = warp_max(wmax); en = warp_sum(en); ed = warp_sum(ed);
        for (int l = 0; l < P.n_lights; ++l) {
            const float* lp = P.lights + (size_t)l * P.light_stride_l + (size_t)ray * P.light_stride_ray;
            const float lx = lp[0], ly = lp[1], lz = lp[2];
            float c[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float sh = g[k][0] * lx + g[k][1] * ly + g[k][2] * lz;
                if (P.warmup == 1) sh = fmaxf(sh, 0.f);
                else if (P.warmup == 2) sh = 1.f;                      // plain colour compositing (render_core)
                const float ws = f[k].w * sh;
                c[0] += al[k][0] * ws; c[1] += al[k][1] * ws; c[2] += al[k][2] * ws;
            }
            c[0] = warp_sum(c[0]); c[1] = warp_sum(c[1]); c[2] = warp_sum(c[2]);
            if (lane == 0) {
                float* dst = P.color + ((size_t)l * P.n_rays + ray) * 3;
                dst[0] = c[0]; dst[1] = c[1]; dst[2] = c[2];
            }
        }
        *reinterpret_cast<float4*>(P.weights + base) = make_float4(f[0].w, f[1].w, f[2].w, f[3].w);
        *reinterpret_cast<float4*>(P.cdf + base) = make_float4(f[0].pc, f[1].pc, f[2].pc, f[3].pc);
        *reinterpret_cast<float4*>(P.inside + base) = make_float4(f[0].inside, f[1].inside, f[2].inside, f[3].inside);
        if (lane == 0) {
            P.weight_sum[ray] = wsum;
            P.weight_max[ray] = wmax;
            P.eik_part[ray * 2] = en;
            P.eik_part[ray * 2 + 1] = ed;
        }
    } else {
        // ---- adjoint (SURVEY 8a' K5)
        const float d_ws = P.d_weight_sum ? P.d_weight_sum[ray] : 0.f;
        const float eik_coef = __ldg(P.d_eik) * 2.f / (__ldg(P.eik_den) + 1e-5f);
        float dw[4] = {d_ws, d_ws, d_ws, d_ws};
        float dg[4][3], dal[4][3];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 3; ++j) dg[k][j] = dal[k][j] = 0.f;
        for (int l = 0; l < P.n_lights; ++l) {
            const float* lp = P.lights + (size_t)l * P.light_stride_l + (size_t)ray * P.light_stride_ray;
            const float lv[3] = {lp[0], lp[1], lp[2]};
            const float* dc = P.d_color + ((size_t)l * P.n_rays + ray) * 3;
            const float c0 = dc[0], c1 = dc[1], c2 = dc[2];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float sh = g[k][0] * lv[0] + g[k][1] * lv[1] + g[k][2] * lv[2];
                bool on = P.warmup != 1 || sh > 0.f;
                if (!on) sh = 0.f;
                if (P.warmup == 2) { sh = 1.f; on = false; }
                const float ca = c0 * al[k][0] + c1 * al[k][1] + c2 * al[k][2];
                dw[k] += ca * sh;
                const float ws = f[k].w * sh;
                dal[k][0] += c0 * ws; dal[k][1] += c1 * ws; dal[k][2] += c2 * ws;
                const float dsh = on ? ca * f[k].w : 0.f;
                dg[k][0] += dsh * lv[0]; dg[k][1] += dsh * lv[1]; dg[k][2] += dsh * lv[2];
            }
        }
        // d alpha_i = T_i dw_i - (sum_{j>i} w_j dw_j) / (1 - alpha_i + 1e-7)
        float ww[4], lsum = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { ww[k] = f[k].w * dw[k]; lsum += ww[k]; }
        const float incl = warp_scan_add(lsum, lane);
        const float total = __shfl_sync(FULL, incl, 31);
        float suffix = total - incl;                           // sum over lanes > this one
        float dsdf[4], dinv = 0.f;
#pragma unroll
        for (int k = 3; k >= 0; --k) {
            float da = f[k].T * dw[k] - suffix * rcp_fast(1.f - f[k].alpha + 1e-7f);
            suffix += ww[k];
            if (!(f[k].alpha_raw >= 0.f && f[k].alpha_raw <= 1.f)) da = 0.f;
            // alpha = (pc - nc + 1e-5) / q, q = pc + 1e-5:  d alpha / d pc = (1 - alpha) / q,  d alpha / d nc = -1 / q
            const float pc = f[k].pc, nc = f[k].nc;
            const float dnc = -da * f[k].rq;
            const float dpc = -dnc * (1.f - f[k].alpha_raw);
            const float dep = dpc * pc * (1.f - pc), den = dnc * nc * (1.f - nc);
            const float ep = sd[k] - f[k].ic * f[k].dist * 0.5f, en = sd[k] + f[k].ic * f[k].dist * 0.5f;
            dinv += dep * ep + den * en;
            dsdf[k] = (dep + den) * inv_s;
            const float dic = (den - dep) * inv_s * f[k].dist * 0.5f;
            const float dtc = dic * (0.5f * (1.f - P.cos_anneal_ratio) * ((-f[k].tc * 0.5f + 0.5f) > 0.f ? 1.f : 0.f) +
                                     P.cos_anneal_ratio * ((-f[k].tc) > 0.f ? 1.f : 0.f));
            const float ek = eik_coef * f[k].relax * (1.f - f[k].rgn) * (f[k].rgn > 0.f ? 1.f : 0.f);      // (|g| - 1) / |g|
#pragma unroll
            for (int j = 0; j < 3; ++j) dg[k][j] += dtc * d[j] + ek * g[k][j];
        }
        dinv = warp_sum(dinv);
        *reinterpret_cast<float4*>(P.d_sdf + base) = make_float4(dsdf[0], dsdf[1], dsdf[2], dsdf[3]);
        float4* gp = reinterpret_cast<float4*>(P.d_grad + base * 3);
        gp[0] = make_float4(dg[0][0], dg[0][1], dg[0][2], dg[1][0]);
        gp[1] = make_float4(dg[1][1], dg[1][2], dg[2][0], dg[2][1]);
        gp[2] = make_float4(dg[2][2], dg[3][0], dg[3][1], dg[3][2]);
        if (P.d_albedo) {
            float4* ap = reinterpret_cast<float4*>(P.d_albedo + base * 3);
            ap[0] = make_float4(dal[0][0], dal[0][1], dal[0][2], dal[1][0]);
            ap[1] = make_float4(dal[1][1], dal[1][2], dal[2][0], dal[2][1]);
            ap[2] = make_float4(dal[2][2], dal[3][0], dal[3][1], dal[3][2]);
        }
        // d variance = d inv_s * 10 * inv_s (zero when the clip is active)
        if (lane == 0) {
            const float e = expf(__ldg(P.variance) * 10.f);
            P.d_var_part[ray] = (e >= 1e-6f && e <= 1e6f) ? dinv * 10.f * inv_s : 0.f;
        }
    }
}

// ------------------------------------------------------------------------------------------------ launchers
cudaError_t launch_coarse_z(const float* near, const float* far, const float* t_rand, float* z, int n_rays, int n_samples,
                            cudaStream_t st) {
    const int total = n_rays * n_samples;
    if (total == 0) return cudaSuccess;
    coarse_z_kernel<<<(total + 255) / 256, 256, 0, st>>>(near, far, t_rand, z, n_rays, n_samples);
    return cudaGetLastError();
}
cudaError_t launch_upsample(const UpsampleParams& P, cudaStream_t st) {
    if (P.n_rays == 0) return cudaSuccess;
    upsample_kernel<<<(P.n_rays + RAYS_PER_BLOCK - 1) / RAYS_PER_BLOCK, 32 * RAYS_PER_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_sample_pdf_from_cdf(const float* bins, const float* cdf, int n_rays, int n, int n_new, float* samples,
                                       int64_t* inds, cudaStream_t st) {
    const int total = n_rays * n_new;
    if (total == 0) return cudaSuccess;
    sample_pdf_from_cdf_kernel<<<(total + 127) / 128, 128, 0, st>>>(bins, cdf, n_rays, n, n_new, samples, inds);
    return cudaGetLastError();
}
cudaError_t launch_final_merge(const float* z_old, int n_old, const float* z_new, int n_new, int n_rays, float sample_dist,
                               float* z_out, float* mid_out, cudaStream_t st) {
    if (n_rays == 0) return cudaSuccess;
    final_merge_kernel<<<(n_rays + RAYS_PER_BLOCK - 1) / RAYS_PER_BLOCK, 32 * RAYS_PER_BLOCK, 0, st>>>(
        z_old, n_old, z_new, n_new, n_rays, sample_dist, z_out, mid_out);
    return cudaGetLastError();
}
// ------------------------------------------------------------------------------------------------ A15
// render_core with the NeRF++ background (renderer.py:194-285 with background_alpha / background_sampled_color):
// forward only.  One warp per ray: the 128 SDF samples 4 per lane as in composite_kernel, then the n_outside
// appended background samples one per lane (two passes for n_outside > 32), the transmittance carried across.
__device__ __forceinline__ float bg_alpha_of(float density, float dist) {
    // alpha = 1 - exp(-softplus(density) * dist)   (renderer.py:118; F.softplus beta = 1, threshold = 20)
    const float sp = density > 20.f ? density : log1pf(expf(density));
    return 1.f - expf(-sp * dist);
}

__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK) composite_bg_kernel(const __grid_constant__ rnb_composite_bg_t P) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ray = blockIdx.x * RAYS_PER_BLOCK + warp;
    if (ray >= P.n_rays) return;
    constexpr int NS = 128;
    const int n_tot = NS + P.n_outside;
    const float inv_s = fminf(fmaxf(expf(__ldg(P.variance) * 10.f), 1e-6f), 1e6f);
    const float o[3] = {P.rays_o[ray * 3], P.rays_o[ray * 3 + 1], P.rays_o[ray * 3 + 2]};
    const float d[3] = {P.rays_d[ray * 3], P.rays_d[ray * 3 + 1], P.rays_d[ray * 3 + 2]};
    const float* zf = P.z_feed + (size_t)ray * n_tot;
    const float* bd = P.bg_density + (size_t)ray * n_tot;
    const float* brgb = P.bg_rgb + (size_t)ray * n_tot * 3;
    float col[3] = {0.f, 0.f, 0.f};
    float wsum = 0.f, wmax = 0.f, en = 0.f, ed = 0.f;
    float alpha[4], c[4][3], pcs[4], ins[4];
    float run = 1.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = lane * 4 + k;
        const size_t gi = (size_t)ray * NS + i;
        const float z = P.z[gi];
        const float dist = i + 1 < NS ? P.z[gi + 1] - z : P.sample_dist;
        const float mid = z + dist * 0.5f;
        const float pt[3] = {o[0] + d[0] * mid, o[1] + d[1] * mid, o[2] + d[2] * mid};
        const float g[3] = {P.grad[gi * 3], P.grad[gi * 3 + 1], P.grad[gi * 3 + 2]};
        SampleFw f;
        sample_forward(P.sdf[gi], g, d, pt, dist, inv_s, P.cos_anneal_ratio, f);
        const float a_bg = bg_alpha_of(bd[i], zf[i + 1] - zf[i]);        // i + 1 <= 128 < n_tot
        alpha[k] = f.alpha * f.inside + a_bg * (1.f - f.inside);
#pragma unroll
        for (int j = 0; j < 3; ++j)
            c[k][j] = P.color_in[gi * 3 + j] * f.inside + sigmoidf_acc(brgb[i * 3 + j]) * (1.f - f.inside);
        pcs[k] = f.pc;
        ins[k] = f.inside;
        en += f.relax * (f.gn - 1.f) * (f.gn - 1.f);
        ed += f.relax;
        run *= 1.f - alpha[k] + 1e-7f;
    }
    float carry;
    {
        const float incl = warp_scan_mul(run, lane);
        float T = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) T = 1.f;
        carry = __shfl_sync(FULL, incl, 31);
        float w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            w[k] = alpha[k] * T;
            T *= 1.f - alpha[k] + 1e-7f;
            wsum += w[k];
            wmax = fmaxf(wmax, w[k]);
            col[0] += w[k] * c[k][0]; col[1] += w[k] * c[k][1]; col[2] += w[k] * c[k][2];
        }
        float* wd = P.weights + (size_t)ray * n_tot + lane * 4;
        wd[0] = w[0]; wd[1] = w[1]; wd[2] = w[2]; wd[3] = w[3];
        const size_t base = (size_t)ray * NS + lane * 4;
        *reinterpret_cast<float4*>(P.cdf + base) = make_float4(pcs[0], pcs[1], pcs[2], pcs[3]);
        *reinterpret_cast<float4*>(P.inside + base) = make_float4(ins[0], ins[1], ins[2], ins[3]);
    }
    for (int j0 = 0; j0 < P.n_outside; j0 += 32) {
        const int i = NS + j0 + lane;
        const bool on = i < n_tot;
        float a = 0.f, cc[3] = {0.f, 0.f, 0.f};
        if (on) {
            const float dist = i + 1 < n_tot ? zf[i + 1] - zf[i] : P.sample_dist;
            a = bg_alpha_of(bd[i], dist);
#pragma unroll
            for (int j = 0; j < 3; ++j) cc[j] = sigmoidf_acc(brgb[i * 3 + j]);
        }
        const float fac = on ? 1.f - a + 1e-7f : 1.f;
        const float incl = warp_scan_mul(fac, lane);
        float T = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) T = 1.f;
        T *= carry;
        carry *= __shfl_sync(FULL, incl, 31);
        const float w = a * T;
        if (on) {
            P.weights[(size_t)ray * n_tot + i] = w;
            wsum += w;
            wmax = fmaxf(wmax, w);
            col[0] += w * cc[0]; col[1] += w * cc[1]; col[2] += w * cc[2];
        }
    }
    wsum = warp_sum(wsum); wmax = warp_max(wmax); en = warp_sum(en); ed = warp_sum(ed);
    col[0] = warp_sum(col[0]); col[1] = warp_sum(col[1]); col[2] = warp_sum(col[2]);
    if (lane == 0) {
        P.color[ray * 3] = col[0]; P.color[ray * 3 + 1] = col[1]; P.color[ray * 3 + 2] = col[2];
        P.weight_sum[ray] = wsum;
        P.weight_max[ray] = wmax;
        P.eik_part[ray * 2] = en;
        P.eik_part[ray * 2 + 1] = ed;
    }
}

cudaError_t launch_composite_bg(const rnb_composite_bg_t& P, cudaStream_t st) {
    if (P.n_rays == 0) return cudaSuccess;
    if (P.n_outside < 1 || P.n_outside > 64) return cudaErrorInvalidValue;
    const int grid = (P.n_rays + RAYS_PER_BLOCK - 1) / RAYS_PER_BLOCK;
    composite_bg_kernel<<<grid, 32 * RAYS_PER_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ ray batches
// Dataset.ps_gen_random_rays_at_view_on_all_lights + near_far_from_sphere + the per-pixel light gather of train_rnb
// (reference models/dataset.py:351-376, 448-458; exp_runner.py:214-220) in one launch over device-resident images:
// one thread per ray.  p = K^-1 [x, y, 1];  v = p / |p|;  d = R v;  o = t;  near/far = -(o.d)/(d.d) -/+ 1.
__global__ void ray_batch_kernel(const __grid_constant__ rnb_ray_batch_t P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_rays) return;
    const int x = (int)P.pixels_x[i], y = (int)P.pixels_y[i];
    const float px = (float)x, py = (float)y;
    float p[3], v[3], d[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        // torch.matmul of a [3,3] with a [3,1]: a plain 3-term dot product per row
        p[r] = P.intrinsics_inv[r * 4 + 0] * px + P.intrinsics_inv[r * 4 + 1] * py + P.intrinsics_inv[r * 4 + 2];
    }
    const float nrm = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
#pragma unroll
    for (int r = 0; r < 3; ++r) v[r] = p[r] / nrm;
    float od = 0.f, dd = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        d[r] = P.pose[r * 4 + 0] * v[0] + P.pose[r * 4 + 1] * v[1] + P.pose[r * 4 + 2] * v[2];
        const float o = P.pose[r * 4 + 3];
        P.rays_o[i * 3 + r] = o;
        P.rays_d[i * 3 + r] = d[r];
        od += o * d[r];
        dd += d[r] * d[r];
    }
    const float mid = 0.5f * (-(2.f * od)) / dd;
    P.near[i] = mid - 1.f;
    P.far[i] = mid + 1.f;
    const size_t pix = (size_t)y * P.W + x;
    if (P.mask_out) P.mask_out[i] = P.mask[pix * P.mask_channels];
    const size_t plane = (size_t)P.H * P.W * 3;
    for (int l = 0; l < P.n_lights; ++l) {
        const size_t src = (size_t)l * plane + pix * 3, dst = ((size_t)l * P.n_rays + i) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (P.rgb) P.rgb[dst + c] = P.images[src + c];
            if (P.rgb2) P.rgb2[dst + c] = P.images2[src + c];
            if (P.lights) P.lights[dst + c] = P.light_dirs[src + c];
        }
    }
}
cudaError_t launch_ray_batch(const rnb_ray_batch_t& P, cudaStream_t st) {
    if (P.n_rays == 0) return cudaSuccess;
    ray_batch_kernel<<<(P.n_rays + 127) / 128, 128, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_composite(const CompositeParams& P, bool bwd, cudaStream_t st) {
    if (P.n_rays == 0) return cudaSuccess;
    const int grid = (P.n_rays + RAYS_PER_BLOCK - 1) / RAYS_PER_BLOCK;
    if (bwd) composite_kernel<true><<<grid, 32 * RAYS_PER_BLOCK, 0, st>>>(P);
    else composite_kernel<false><<<grid, 32 * RAYS_PER_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

}  // namespace rnb
