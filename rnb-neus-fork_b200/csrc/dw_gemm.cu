// Weight-gradient GEMMs: dW[256 x N] = sum over points of A_stream^T * B_stream (K = points).
//   SDF layer l (reference autograd, SURVEY 8a' K3):  dW_l = w_l^T uin_bar_l + zbar_l^T in_l   (two stream pairs)
// Both operands are read straight from the activation streams as MN-major SWIZZLE_NONE UMMA operands: a 64-point
// stream sub-tile [chunk][64 rows][16 B] IS the canonical MN-major image (8 points x 8 columns = one 128-byte core
// matrix), so one cp.async.bulk per operand per stage feeds tcgen05.mma with no transpose anywhere.
// Split-K over the points; each CTA keeps a full 256 x N fp32 accumulator in TMEM (2 x 256 columns), dumps one
// partial, and a small kernel reduces the partials deterministically.  HBM-bound by design (128 FLOP/B).
// The four warps that are idle while the MMAs run take the job's column sums from the staged tiles: bias gradients and
// the rank-1..3 weight gradients (sdf row of W_8, the 3-row W_2 of the albedo net) that round 1 computed with a separate
// pass over the streams (colsum_kernel, 2 x 0.27 ms per step).
#include "common.cuh"
#include "points.cuh"     // cot_scale_from_max
#include "dw_params.h"
#include "dw_common.cuh"

namespace rnb {

constexpr int DW_THREADS = 320;       // producer, MMA issuer, 2 x 4 column-sum / accumulator-dump warps
constexpr int DW_STAGES = 3;
constexpr int DW_STAGE_A = 32768;     // 64 points x 256 columns fp16
constexpr int DW_STAGE_B = 32768;     // 64 points x <=256 columns fp16
constexpr int DW_STAGE = DW_STAGE_A + DW_STAGE_B;
constexpr int DW_SMEM = DW_STAGES * DW_STAGE + 128 + 8 * DWC_WBUF_FLOATS * 4;

__global__ void __launch_bounds__(DW_THREADS, 1) dw_gemm_kernel(const __grid_constant__ DwParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DW_STAGES * DW_STAGE);
    uint64_t* full = bars;
    uint64_t* empty = bars + DW_STAGES;
    uint64_t* acc_full = bars + 2 * DW_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DW_STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const DwJob& job = P.jobs[blockIdx.y];
    const int split = blockIdx.x;
    const int per = (P.n_sub + (int)gridDim.x - 1) / (int)gridDim.x;
    const int sub0 = split * per;
    const int sub1 = min(P.n_sub, sub0 + per);
    const int nw = job.nw;
    const uint32_t b_bytes = (uint32_t)(nw >> 3) * 1024u;

    if (threadIdx.x == 0) {
        // empty[]: the MMA commit (or the MMA lane's plain arrival for a column-sum-only stage) + one arrival per
        // epilogue warp (they read the staged tiles for the column sums)
        for (int i = 0; i < DW_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1 + 8); }
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int sub = sub0; sub < sub1; ++sub)
                for (int pr = 0; pr < job.n_pairs; ++pr, ++it) {
                    const uint32_t slot = it % DW_STAGES, ph = (it / DW_STAGES) & 1;
                    const bool mma = pr < job.mma_pairs;
                    const uint8_t* rider = job.x[pr];
                    dw_prefetch_weights(job, pr, sub);
                    mbar_wait(&empty[slot], ph ^ 1);
                    mbar_expect_tx(&full[slot], DW_STAGE_A + (mma ? b_bytes : 0u) + (rider ? (uint32_t)DWC_RIDER_BYTES : 0u));
                    uint8_t* dst = smem + slot * DW_STAGE;
                    bulk_g2s(dst, job.a[pr] + (size_t)sub * DW_STAGE_A, DW_STAGE_A, &full[slot]);
                    if (mma)
                        bulk_g2s(dst + DW_STAGE_A, job.b[pr] + ((size_t)sub * job.b_chunks[pr] + job.b_chunk0) * 1024u, b_bytes,
                                 &full[slot]);
                    if (rider)
                        bulk_g2s(dst + DW_STAGE_A + b_bytes, rider + ((size_t)sub * 32 + job.x_chunk0[pr]) * 1024u, DWC_RIDER_BYTES,
                                 &full[slot]);
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(128, nw, FMT_F16, FMT_F16, MAJOR_MN, MAJOR_MN);
            uint32_t it = 0, n_mma = 0;
            for (int sub = sub0; sub < sub1; ++sub)
                for (int pr = 0; pr < job.n_pairs; ++pr, ++it) {
                    const uint32_t slot = it % DW_STAGES, ph = (it / DW_STAGES) & 1;
                    mbar_wait(&full[slot], ph);
                    if (pr >= job.mma_pairs) {      // staged for the column sums only
                        mbar_arrive(&empty[slot]);
                        continue;
                    }
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + slot * DW_STAGE), sb = sa + DW_STAGE_A;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            // MN-major: LBO = next 8-point group (128 B), SBO = next 8-column chunk (64 rows x 16 B)
                            const uint64_t ad = umma_desc(sa + h * 16384 + ks * 256, 128, 1024);
                            const uint64_t bd = umma_desc(sb + ks * 256, 128, 1024);
                            umma_f16(tmem + h * 256, ad, bd, idesc, (n_mma | ks) != 0);
                        }
                    ++n_mma;
                    umma_commit(&empty[slot]);
                }
            umma_commit(acc_full);
        }
    } else {
        // While the MMAs run these eight warps are idle: they take the job's column sums straight from the staged shared
        // memory (dw_common.cuh; warps 2-5 the point rows 0-31 of every sub-tile, warps 6-9 rows 32-63), then dump the
        // accumulator (warps 2-5 the rows 0-127 of dW, warps 6-9 rows 128-255).
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int grp = (warp - 2) >> 2;                  // 0 / 1
        const int et = ((warp - 2) & 3) * 32 + lane;      // 0..127: owner of columns 2 et, 2 et + 1
        DwColsumAcc cs;
        cs.clear();
        {
            uint32_t it = 0;
            for (int sub = sub0; sub < sub1; ++sub)
                for (int pr = 0; pr < job.n_pairs; ++pr, ++it) {
                    const uint32_t slot = it % DW_STAGES, ph = (it / DW_STAGES) & 1;
                    // always observe the fill of this use before arriving on empty[]: keeps every arrival in its own phase
                    mbar_wait(&full[slot], ph);
                    cs.stage<32>(job, pr, sub, smem + slot * DW_STAGE, et,
                                 reinterpret_cast<float*>(smem + DW_STAGES * DW_STAGE + 128) + (warp - 2) * DWC_WBUF_FLOATS, grp * 32);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[slot]);
                }
        }
        float* out = job.partial + (size_t)split * 256 * nw;
        if (sub1 > sub0) {
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
                        if (c0 + 4 * j < nw)          // nw = 16: only the first 16 of the 32 loaded columns exist
                            dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
        } else {
            for (int c = 0; c < nw; ++c) out[(size_t)(grp * 128 + row) * nw + c] = 0.f;
        }
        cs.store(job, 2 * split + grp, et);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------ deterministic reduction of split-K partials
__global__ void __launch_bounds__(256) reduce_kernel(const __grid_constant__ ReduceParams P) {
    const ReduceJob& job = P.jobs[blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= job.out_rows * job.out_cols) return;
    const int m = idx / job.out_cols, n = idx % job.out_cols;
    float inv_sc = 1.f;
    {
        const float mx = __ldg(P.cot_absmax);
        if (mx > 0.f && isfinite(mx)) {
            int e;
            frexpf(mx, &e);
            inv_sc = ldexpf(1.f, e - RNB_COT_EXP);
        }
    }
    const float factor = job.use_cot_scale ? job.factor * inv_sc : job.factor;
    float acc = 0.f;
    const size_t stride = (size_t)job.rows * job.nw;
    const float* src = job.partial + (size_t)m * job.nw + n;
    for (int s = 0; s < job.splits; ++s) acc += src[(size_t)s * stride];
    if (job.fold_xlo && n < 3) {
        const float* src2 = job.partial + (size_t)m * job.nw + 39 + n;
        for (int s = 0; s < job.splits; ++s) acc += src2[(size_t)s * stride];
    }
    acc *= factor;
    if (job.partial2) {
        float acc2 = 0.f;
        const float* s2 = job.partial2 + (size_t)m * job.nw + n;
        for (int s = 0; s < job.splits2; ++s) acc2 += s2[(size_t)s * stride];
        acc += acc2 * (job.use_cot_scale2 ? job.factor2 * inv_sc : job.factor2);
    }
    float* dst = job.dst + (size_t)(job.dst_row0 + m) * job.dst_pitch + (size_t)(job.dst_col0 + n) * (job.dst_col_stride ? job.dst_col_stride : 1);
    *dst = job.accumulate ? *dst + acc : acc;
}

// ------------------------------------------------------------------ max |x| over up to three fp32 arrays
// `fold` (optional, device float[2]): a maximum that was taken in another kernel's scaled units -- fold[1] = max of
// values stored with the power-of-two scale derived from the cotangent absmax fold[0] -- joins the result unscaled.
__global__ void absmax_kernel(const float* a, int64_t na, const float* b, int64_t nb, const float* c, int64_t nc,
                              const float* fold, float* out) {
    float m = 0.f;
    if (fold && blockIdx.x == 0 && threadIdx.x == 0) m = fold[1] / cot_scale_from_max(fold[0]);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = i0; i < na; i += stride) m = fmaxf(m, fabsf(a[i]));
    for (int64_t i = i0; i < nb; i += stride) m = fmaxf(m, fabsf(b[i]));
    for (int64_t i = i0; i < nc; i += stride) m = fmaxf(m, fabsf(c[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
}

cudaError_t launch_dw_gemm(const DwParams& P, int splits, cudaStream_t st) {
    {
        cudaError_t e = ensure_dynamic_smem((const void*)dw_gemm_kernel, DW_SMEM);
        if (e != cudaSuccess) return e;
    }
    if (P.n_jobs == 0) return cudaSuccess;
    dw_gemm_kernel<<<dim3(splits, P.n_jobs), DW_THREADS, DW_SMEM, st>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_reduce(const ReduceParams& P, cudaStream_t st) {
    if (P.n_jobs == 0) return cudaSuccess;
    int mx = 0;
    for (int j = 0; j < P.n_jobs; ++j) mx = max(mx, P.jobs[j].out_rows * P.jobs[j].out_cols);
    reduce_kernel<<<dim3((mx + 255) / 256, P.n_jobs), 256, 0, st>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_absmax(const float* a, int64_t na, const float* b, int64_t nb, const float* c, int64_t nc,
                          const float* fold, float* out, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), st);
    if (e != cudaSuccess) return e;
    absmax_kernel<<<296, 256, 0, st>>>(a, na, b, nb, c, nc, fold, out);
    return cudaGetLastError();
}

}  // namespace rnb
