// Stand-alone probe: how many bytes per second can the SMs pull out of L2 with cp.async.bulk (the weight ring of the
// chain kernels), as a function of slice size, slices in flight, CTAs per SM, working-set size -- and whether a
// cluster-of-2 multicast (each CTA issues half a slice, both receive all of it; "mcast2"), one issuing lane per ring slot
// ("lanes") or a slice fetched as 2 / 4 concurrent part copies ("split2/4") deliver more bytes per SM than one lane issuing
// whole slices ("unicast").
// The chain kernels' roofline rests on this number.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o l2bw_probe l2bw_probe.cu
// Run:   ./l2bw_probe            (prints one line per configuration)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../common.cuh"
using namespace rnb;

struct ProbeParams {
    const uint8_t* src;
    unsigned long long ws_bytes;   // working set (power of two)
    int slice;                     // bytes per ring slot
    int stages;                    // slots in flight
    int iters;                     // slices per CTA
    int stride_ctas;               // distance (in slices) between the start offsets of consecutive CTAs
};

__global__ void __launch_bounds__(128) bulk_unicast(const __grid_constant__ ProbeParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < P.stages; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long mask = P.ws_bytes - 1;
        unsigned long long off = ((unsigned long long)blockIdx.x * P.stride_ctas * P.slice) & mask;
        for (int it = 0; it < P.iters + P.stages; ++it) {
            const int slot = it % P.stages;
            if (it >= P.stages) mbar_wait(&full[slot], ((it / P.stages) - 1) & 1);
            if (it < P.iters) {
                mbar_expect_tx(&full[slot], P.slice);
                bulk_g2s(smem + (size_t)slot * P.slice, P.src + off, P.slice, &full[slot]);
                off = (off + P.slice) & mask;
            }
        }
    }
    __syncthreads();
}

// every slice is fetched by SPLIT lanes, each copying 1/SPLIT of it (same barrier): does a slice land sooner when its bytes
// come as several concurrent copies?
template <int SPLIT>
__global__ void __launch_bounds__(128) bulk_split(const __grid_constant__ ProbeParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < P.stages; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const unsigned long long mask = P.ws_bytes - 1;
        unsigned long long off = ((unsigned long long)blockIdx.x * P.stride_ctas * P.slice) & mask;
        const uint32_t part = P.slice / SPLIT;
        for (int it = 0; it < P.iters + P.stages; ++it) {
            const int slot = it % P.stages;
            if (it >= P.stages) mbar_wait(&full[slot], ((it / P.stages) - 1) & 1);
            if (it < P.iters) {
                if (lane == 0) mbar_expect_tx(&full[slot], P.slice);
                __syncwarp();
                if (lane < SPLIT) bulk_g2s(smem + (size_t)slot * P.slice + lane * part, P.src + off + lane * part, part, &full[slot]);
                off = (off + P.slice) & mask;
            }
        }
    }
    __syncthreads();
}

// one issuing lane PER SLOT: lane s refills slot s as soon as it has seen it full (do the per-copy issue costs overlap?)
__global__ void __launch_bounds__(128) bulk_lanes(const __grid_constant__ ProbeParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < P.stages; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if ((int)threadIdx.x < P.stages) {
        const int slot = threadIdx.x;
        const unsigned long long mask = P.ws_bytes - 1;
        unsigned long long off = ((unsigned long long)blockIdx.x * P.stride_ctas * P.slice + (unsigned long long)slot * P.slice) & mask;
        const int n = P.iters / P.stages;
        for (int it = 0; it <= n; ++it) {
            if (it > 0) mbar_wait(&full[slot], (it - 1) & 1);
            if (it < n) {
                mbar_expect_tx(&full[slot], P.slice);
                bulk_g2s(smem + (size_t)slot * P.slice, P.src + off, P.slice, &full[slot]);
                off = (off + (unsigned long long)P.stages * P.slice) & mask;
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void bulk_g2s_mc(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}

// cluster of 2: each CTA issues HALF of every slice, multicast to both; a slot is re-used when both CTAs have seen it full
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) bulk_multicast(const __grid_constant__ ProbeParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[16];
    __shared__ uint64_t empty[16];
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        for (int i = 0; i < P.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); }
        mbar_fence_init();
    }
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x == 0) {
        const unsigned long long mask = P.ws_bytes - 1;
        unsigned long long off = ((unsigned long long)(blockIdx.x >> 1) * P.stride_ctas * P.slice) & mask;
        const uint32_t half = P.slice / 2;
        const uint32_t peer_empty = mapa_u32(smem_u32(empty), rank ^ 1);
        for (int it = 0; it < P.iters + P.stages; ++it) {
            const int slot = it % P.stages;
            if (it >= P.stages) {
                // consume: this CTA has seen the slot full -> tell both producers
                mbar_wait(&full[slot], ((it / P.stages) - 1) & 1);
                mbar_arrive(&empty[slot]);
                mbar_arrive_cluster(peer_empty + slot * 8);
                if (it < P.iters) mbar_wait_cluster(&empty[slot], ((it / P.stages) - 1) & 1);
            }
            if (it < P.iters) {
                mbar_expect_tx(&full[slot], P.slice);
                bulk_g2s_mc(smem + (size_t)slot * P.slice + rank * half, P.src + off + rank * half, half, &full[slot], (uint16_t)3);
                off = (off + P.slice) & mask;
            }
        }
    }
    __syncthreads();
    cluster_sync_all();
}

// plain 16-byte loads, 4 in flight per thread (the epilogue's stream reads)
__global__ void __launch_bounds__(256) ldg_probe(const __grid_constant__ ProbeParams P, float* sink) {
    const unsigned long long mask = P.ws_bytes - 1;
    unsigned long long off = ((unsigned long long)blockIdx.x * P.stride_ctas * P.slice + threadIdx.x * 16) & mask;
    uint4 acc = make_uint4(0, 0, 0, 0);
    const int per_iter = 256 * 16 * 4;
    const long long total = (long long)P.iters * P.slice;
    for (long long done = 0; done < total; done += per_iter) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const uint4*>(P.src + ((off + (unsigned long long)k * 4096) & mask));
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc.x ^= v[k].x; acc.y ^= v[k].y; acc.z ^= v[k].z; acc.w ^= v[k].w; }
        off = (off + per_iter) & mask;
    }
    if (acc.x == 0x12345678u) sink[0] = 1.f;
}

int main() {
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const size_t buf_bytes = 256ull << 20;
    uint8_t* buf;
    cudaMalloc(&buf, buf_bytes);
    cudaMemset(buf, 1, buf_bytes);
    float* sink;
    cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("sm_count %d, max clock %d kHz\n", sm, clk_khz);
    printf("%-10s %8s %7s %6s %5s %10s %10s %12s\n", "mode", "ws_MB", "slice", "stages", "cta/sm", "ms", "TB/s", "B/clk/SM@max");
    auto report = [&](const char* mode, const ProbeParams& P, int ctas_per_sm, double bytes, float ms) {
        const double tbs = bytes / (ms * 1e-3) / 1e12;
        printf("%-10s %8.1f %7d %6d %5d %10.3f %10.3f %12.1f\n", mode, P.ws_bytes / 1048576.0, P.slice, P.stages, ctas_per_sm, ms,
               tbs, bytes / (ms * 1e-3) / (clk_khz * 1e3) / sm);
        fflush(stdout);
    };
    const unsigned long long ws_list[] = {2ull << 20, 32ull << 20, 256ull << 20};
    const int slice_list[] = {8192, 16384, 32768};
    const int stage_list[] = {3, 6};
    for (unsigned long long ws : ws_list)
        for (int slice : slice_list)
            for (int stages : stage_list)
                for (int cps = 1; cps <= 2; ++cps) {
                    if ((size_t)slice * stages > 100 * 1024) continue;
                    ProbeParams P{buf, ws, slice, stages, (int)((48ull << 20) / slice), 37};
                    const int smem = slice * stages;
                    cudaFuncSetAttribute((const void*)bulk_unicast, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    cudaFuncSetAttribute((const void*)bulk_multicast, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    const int grid = sm * cps;
                    for (int rep = 0; rep < 2; ++rep) {
                        cudaEventRecord(e0);
                        bulk_unicast<<<grid, 128, smem>>>(P);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                    }
                    float ms = 0;
                    cudaEventElapsedTime(&ms, e0, e1);
                    report("unicast", P, cps, (double)grid * P.iters * slice, ms);
                    cudaFuncSetAttribute((const void*)bulk_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    for (int rep = 0; rep < 2; ++rep) {
                        cudaEventRecord(e0);
                        bulk_lanes<<<grid, 128, smem>>>(P);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                    }
                    cudaEventElapsedTime(&ms, e0, e1);
                    report("lanes", P, cps, (double)grid * (P.iters / stages) * stages * slice, ms);
                    cudaFuncSetAttribute((const void*)bulk_split<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    cudaFuncSetAttribute((const void*)bulk_split<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    for (int rep = 0; rep < 2; ++rep) {
                        cudaEventRecord(e0);
                        bulk_split<2><<<grid, 128, smem>>>(P);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                    }
                    cudaEventElapsedTime(&ms, e0, e1);
                    report("split2", P, cps, (double)grid * P.iters * slice, ms);
                    for (int rep = 0; rep < 2; ++rep) {
                        cudaEventRecord(e0);
                        bulk_split<4><<<grid, 128, smem>>>(P);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                    }
                    cudaEventElapsedTime(&ms, e0, e1);
                    report("split4", P, cps, (double)grid * P.iters * slice, ms);
                    for (int rep = 0; rep < 2; ++rep) {
                        cudaEventRecord(e0);
                        bulk_multicast<<<grid, 128, smem>>>(P);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                    }
                    cudaEventElapsedTime(&ms, e0, e1);
                    // bytes DELIVERED into shared memory (each CTA receives whole slices, issues half of them)
                    report("mcast2", P, cps, (double)grid * P.iters * slice, ms);
                    if (cudaGetLastError() != cudaSuccess) { printf("CUDA error\n"); return 1; }
                }
    for (unsigned long long ws : ws_list)
        for (int cps = 2; cps <= 8; cps *= 2) {
            ProbeParams P{buf, ws, 16384, 1, (int)((48ull << 20) / 16384), 37};
            const int grid = sm * cps;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                ldg_probe<<<grid, 256>>>(P, sink);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            report("ldg128", P, cps, (double)grid * P.iters * 16384, ms);
        }
    cudaError_t e = cudaDeviceSynchronize();
    printf("final: %s\n", cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}
