// Weight-gradient GEMMs: dW[256 x N] = sum over points of A_stream^T * B_stream (K = points).
//   SDF layer l (reference autograd, SURVEY 8a' K3):  dW_l = w_l^T uin_bar_l + zbar_l^T in_l   (two stream pairs)
// Both operands are read straight from the activation streams as MN-major SWIZZLE_NONE UMMA operands: a 64-point
// stream sub-tile [chunk][64 rows][16 B] IS the canonical MN-major image (8 points x 8 columns = one 128-byte core
// matrix), so one cp.async.bulk per operand per stage feeds tcgen05.mma with no transpose anywhere.
// Split-K over the points; each CTA keeps a full 256 x N fp32 accumulator in TMEM (2 x 256 columns), dumps one
// partial, and a small kernel reduces the partials deterministically.  HBM-bound by design (128 FLOP/B).
#include "common.cuh"
#include "points.cuh"     // cot_scale_from_max
#include "dw_params.h"

namespace rnb {

constexpr int DW_THREADS = 192;
constexpr int DW_STAGES = 3;
constexpr int DW_STAGE_A = 32768;     // 64 points x 256 columns fp16
constexpr int DW_STAGE_B = 32768;     // 64 points x <=256 columns fp16
constexpr int DW_STAGE = DW_STAGE_A + DW_STAGE_B;
constexpr int DW_SMEM = DW_STAGES * DW_STAGE + 128;

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
        // empty[]: the MMA commit + one arrival per epilogue warp (they read the A tile for the column sums)
        for (int i = 0; i < DW_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1 + 4); }
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
                    mbar_wait(&empty[slot], ph ^ 1);
                    mbar_expect_tx(&full[slot], DW_STAGE_A + b_bytes);
                    uint8_t* dst = smem + slot * DW_STAGE;
                    bulk_g2s(dst, job.a[pr] + (size_t)sub * DW_STAGE_A, DW_STAGE_A, &full[slot]);
                    bulk_g2s(dst + DW_STAGE_A, job.b[pr] + ((size_t)sub * job.b_chunks[pr] + job.b_chunk0) * 1024u, b_bytes,
                             &full[slot]);
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(128, nw, FMT_F16, FMT_F16, MAJOR_MN, MAJOR_MN);
            uint32_t it = 0;
            for (int sub = sub0; sub < sub1; ++sub)
                for (int pr = 0; pr < job.n_pairs; ++pr, ++it) {
                    const uint32_t slot = it % DW_STAGES, ph = (it / DW_STAGES) & 1;
                    mbar_wait(&full[slot], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + slot * DW_STAGE), sb = sa + DW_STAGE_A;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            // MN-major: LBO = next 8-point group (128 B), SBO = next 8-column chunk (64 rows x 16 B)
                            const uint64_t ad = umma_desc(sa + h * 16384 + ks * 256, 128, 1024);
                            const uint64_t bd = umma_desc(sb + ks * 256, 128, 1024);
                            umma_f16(tmem + h * 256, ad, bd, idesc, (it | ks) != 0);
                        }
                    umma_commit(&empty[slot]);
                }
            umma_commit(acc_full);
        }
    } else {
        // While the MMAs run these four warps are idle: they column-sum the flagged A-side tile straight from the
        // staged shared memory (db_l = sum_p zbar_l[p,:]), then dump the accumulator.
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int et = (warp - 2) * 32 + lane;            // 0..127
        const int cs_row = et & 63, cs_cg = et >> 6;      // thread owns point row cs_row and chunks cs_cg*16 .. +15
        float cs[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) cs[i] = 0.f;
        {
            uint32_t it = 0;
            for (int sub = sub0; sub < sub1; ++sub)
                for (int pr = 0; pr < job.n_pairs; ++pr, ++it) {
                    const uint32_t slot = it % DW_STAGES, ph = (it / DW_STAGES) & 1;
                    // always observe the fill of this use before arriving on empty[]: keeps every arrival in its own phase
                    mbar_wait(&full[slot], ph);
                    if (pr == job.colsum_pair) {
                        const uint8_t* a_tile = smem + slot * DW_STAGE;
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const uint4 u = *reinterpret_cast<const uint4*>(a_tile + ((size_t)(cs_cg * 16 + c) * 64 + cs_row) * 16);
                            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 f = unpack_h2(w[j]);
                                cs[c * 8 + 2 * j] += f.x;
                                cs[c * 8 + 2 * j + 1] += f.y;
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[slot]);
                }
        }
        float* out = job.partial + (size_t)split * 256 * nw;
        if (sub1 > sub0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
#pragma unroll 1
            for (int h = 0; h < 2; ++h)
#pragma unroll 1
                for (int c0 = 0; c0 < nw; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + h * 256 + c0, v);
                    tmem_ld_wait();
                    float4* dst = reinterpret_cast<float4*>(out + (size_t)(h * 128 + row) * nw + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                             __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
        } else {
            for (int h = 0; h < 2; ++h)
                for (int c = 0; c < nw; ++c) out[(size_t)(h * 128 + row) * nw + c] = 0.f;
        }
        if (job.colsum_pair >= 0) {
            // reduce over the 64 point rows: lanes of a warp hold 32 rows, warps (2,3) / (4,5) hold the two halves
#pragma unroll
            for (int i = 0; i < 128; ++i) {
                float v = cs[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                cs[i] = v;
            }
            // the ring is drained (acc_full passed): reuse its first bytes as scratch for the cross-warp add
            float* scr = reinterpret_cast<float*>(smem);
            if (lane == 0 && (et & 32) != 0) {             // second warp of each chunk group publishes
#pragma unroll
                for (int i = 0; i < 128; ++i) scr[cs_cg * 128 + i] = cs[i];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (lane == 0 && (et & 32) == 0) {
                float* dst = job.cs_partial + (size_t)split * 256 + cs_cg * 128;
#pragma unroll
                for (int i = 0; i < 128; ++i) dst[i] = cs[i] + scr[cs_cg * 128 + i];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------ column sums over a stream (bias gradients)
__global__ void __launch_bounds__(256) colsum_kernel(const __grid_constant__ ColsumParams P) {
    const ColsumJob& job = P.jobs[blockIdx.z];
    const int chunk = blockIdx.x;
    if (chunk >= job.chunks) return;
    const int split = blockIdx.y;
    const int per = (P.n_sub + (int)gridDim.y - 1) / (int)gridDim.y;
    const int sub0 = split * per, sub1 = min(P.n_sub, sub0 + per);
    const int r = threadIdx.x & 63, sl = threadIdx.x >> 6;
    float acc[3][8];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
#pragma unroll 4
    for (int sub = sub0 + sl; sub < sub1; sub += 4) {
        const uint4 u = *reinterpret_cast<const uint4*>(job.stream + (((size_t)sub * job.chunks + chunk) * 64 + r) * 16);
        const int64_t p = (int64_t)sub * 64 + r;
        float wgt[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            wgt[k] = k < job.n_w ? 1.f : 0.f;
            if (k < job.n_w && job.row_weight[k]) wgt[k] = p < P.n_pts ? __ldg(job.row_weight[k] + p) : 0.f;
        }
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_h2(w[j]);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                acc[k][2 * j] = fmaf(wgt[k], f.x, acc[k][2 * j]);
                acc[k][2 * j + 1] = fmaf(wgt[k], f.y, acc[k][2 * j + 1]);
            }
        }
    }
    __shared__ float red[3][8][8];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = acc[k][j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5][j] = v;
        }
    __syncthreads();
    if (threadIdx.x < 8 * job.n_w) {
        const int k = threadIdx.x >> 3, j = threadIdx.x & 7;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[k][w][j];
        job.partial[k][(size_t)split * job.chunks * 8 + chunk * 8 + j] = v;
    }
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
            inv_sc = ldexpf(1.f, e - 8);
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
    float* dst = job.dst + (size_t)(job.dst_row0 + m) * job.dst_pitch + job.dst_col0 + n;
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

// sum of an fp32 array (d b_8[0] = sum d_sdf; d b_2 of the albedo net): a fixed grid of partial sums in caller scratch
// (SUM_BLOCKS floats), then one block adds them in a fixed order -- deterministic, no global state.
constexpr int SUM_BLOCKS = 64;

__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    return v;       // valid in thread 0
}
__global__ void __launch_bounds__(256) sum_partial_kernel(const float* x, int64_t n, float* partial) {
    __shared__ float red[8];
    float v = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)SUM_BLOCKS * 256) v += x[i];
    v = block_sum_256(v, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
}
__global__ void __launch_bounds__(256) sum_final_kernel(const float* partial, float* out) {
    __shared__ float red[8];
    const float v = block_sum_256(threadIdx.x < SUM_BLOCKS ? partial[threadIdx.x] : 0.f, red);
    if (threadIdx.x == 0) *out = v;
}
cudaError_t launch_sum(const float* x, int64_t n, float* partial, float* out, cudaStream_t st) {
    sum_partial_kernel<<<SUM_BLOCKS, 256, 0, st>>>(x, n, partial);
    sum_final_kernel<<<1, 256, 0, st>>>(partial, out);
    return cudaGetLastError();
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
cudaError_t launch_colsum(const ColsumParams& P, int splits, cudaStream_t st) {
    if (P.n_jobs == 0) return cudaSuccess;
    colsum_kernel<<<dim3(40, splits, P.n_jobs), 256, 0, st>>>(P);
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
