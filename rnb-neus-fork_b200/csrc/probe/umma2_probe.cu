// Stand-alone hardware probe for the CTA-pair form of tcgen05.mma (cta_group::2) with the SWIZZLE_NONE "chunked"
// operand images of common.cuh:
//   * a cluster of two CTAs; each holds its own A tile [128 rows x K] (K-major) and HALF of B: rows
//     [N/2 * rank, N/2 * (rank + 1)) of the [N x K] weight image;
//   * the leader issues M = 256 MMAs; D rows 128*rank .. +127 land in CTA `rank`'s TMEM, all N columns;
//   * completion is multicast to both CTAs' mbarriers (tcgen05.commit ... multicast::cluster).
// It also exercises the cross-CTA "operand ready" handshake the chain kernels would need: the peer CTA arrives on a
// barrier of the leader through a mapa'd shared::cluster address.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma2_probe umma2_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../common.cuh"
using namespace rnb;

constexpr int KDIM = 64;       // 4 MMAs of K = 16
constexpr int NDIM = 256;

struct Probe2Params {
    const uint8_t* a_img;   // [2][128 x K] chunked images, one per CTA
    const uint8_t* b_img;   // [2][N/2 x K] chunked images, one per CTA
    float* d;               // [256 x N] row-major
    int alloc_both;         // 1: both CTAs execute tcgen05.alloc.cta_group::2, 0: leader only
};

// cluster / pair helpers (cluster_ctarank, cluster_sync_all, mapa_u32, mbar_arrive_cluster, tmem_alloc2, tmem_dealloc2,
// umma2_f16, umma2_commit_mc): ../common.cuh

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2_kernel(Probe2Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_ready, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const uint32_t rank = cluster_ctarank();
    uint8_t* sa = smem;                               // 128 x 64 fp16 = 16 KB
    uint8_t* sb = smem + 16384;                       // 128 x 64 fp16 = 16 KB (this CTA's half of B)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bar_ready, 2);                     // one arrival per CTA of the pair (leader's copy is the one used)
        mbar_init(&bar_mma, 1);
        mbar_fence_init();
    }
    // operands: plain loads -> generic-proxy stores, then the proxy fence (what the chain epilogues do)
    for (int i = threadIdx.x; i < 16384 / 16; i += 128) {
        reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(p.a_img + (size_t)rank * 16384)[i];
        reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(p.b_img + (size_t)rank * 16384)[i];
    }
    fence_proxy_async();
    __syncthreads();
    cluster_sync_all();                               // barriers initialised in both CTAs before any remote arrive
    if (warp == 0 && (p.alloc_both || rank == 0)) tmem_alloc2(&tmem_base_s, 256);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    if (!p.alloc_both && rank == 1) {
        // leader-only allocation: the pair shares the column range; read it from the leader's shared memory
        uint32_t v;
        asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(mapa_u32(smem_u32(&tmem_base_s), 0)));
        tmem_base_s = v;
        __syncthreads();
    }
    const uint32_t tmem = tmem_base_s;
    // "my operand is ready": both CTAs arrive on the LEADER's barrier
    if (threadIdx.x == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_ready), 0));
    if (rank == 0 && threadIdx.x == 0) {
        mbar_wait(&bar_ready, 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc(256, NDIM, FMT_F16, FMT_F16);
        for (int k = 0; k < KDIM / 16; ++k) {
            // chunked K-major images with 128 rows: LBO (next 8-column chunk) = 128 * 16 B, SBO = 128 B; K = 16 = two chunks
            const uint64_t ad = umma_desc(smem_u32(sa) + k * 2 * 2048, 2048, 128);
            const uint64_t bd = umma_desc(smem_u32(sb) + k * 2 * 2048, 2048, 128);
            umma2_f16(tmem, ad, bd, idesc, k > 0);
        }
        umma2_commit_mc(&bar_mma);
    }
    __syncwarp();
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < NDIM; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) p.d[(size_t)(rank * 128 + row) * NDIM + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0 && (p.alloc_both || rank == 0)) tmem_dealloc2(tmem, 256);
}

static uint16_t f2h(float f) { __half h = __float2half_rn(f); return *reinterpret_cast<uint16_t*>(&h); }
static float h2f(uint16_t u) { __half h = *reinterpret_cast<__half*>(&u); return __half2float(h); }

// chunked image of a [rows x cols] matrix: offset(r, c) = ((c >> 3) * rows + r) * 16 + (c & 7) * 2
static void to_chunked(const std::vector<float>& m, int rows, int cols, int row0, uint8_t* out) {
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const uint16_t h = f2h(m[(size_t)(row0 + r) * cols + c]);
            *reinterpret_cast<uint16_t*>(out + ((size_t)(c >> 3) * rows + r) * 16 + (c & 7) * 2) = h;
        }
}

int main(int argc, char** argv) {
    const int alloc_both = argc > 1 ? atoi(argv[1]) : 1;
    std::vector<float> A(256 * KDIM), B(NDIM * KDIM);
    srand(1);
    for (auto& v : A) v = h2f(f2h((rand() % 2001 - 1000) / 1000.f));
    for (auto& v : B) v = h2f(f2h((rand() % 2001 - 1000) / 1000.f));
    std::vector<uint8_t> a_img(2 * 16384), b_img(2 * 16384);
    for (int r = 0; r < 2; ++r) {
        to_chunked(A, 128, KDIM, 128 * r, a_img.data() + r * 16384);
        to_chunked(B, NDIM / 2, KDIM, (NDIM / 2) * r, b_img.data() + r * 16384);
    }
    uint8_t *da, *db;
    float* dd;
    cudaMalloc(&da, a_img.size());
    cudaMalloc(&db, b_img.size());
    cudaMalloc(&dd, 256 * NDIM * 4);
    cudaMemcpy(da, a_img.data(), a_img.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, b_img.data(), b_img.size(), cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xff, 256 * NDIM * 4);
    Probe2Params p{da, db, dd, alloc_both};
    cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    probe2_kernel<<<2, 128, 32768>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    printf("alloc_both=%d launch: %s\n", alloc_both, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> D(256 * NDIM);
    cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
    double max_err = 0;
    for (int m = 0; m < 256; ++m)
        for (int n = 0; n < NDIM; ++n) {
            double ref = 0;
            for (int k = 0; k < KDIM; ++k) ref += (double)A[m * KDIM + k] * B[n * KDIM + k];
            max_err = fmax(max_err, fabs(ref - D[m * NDIM + n]));
        }
    printf("cta_group::2 M=256 N=%d K=%d: max |err| = %.3e  -> %s\n", NDIM, KDIM, max_err, max_err < 1e-3 ? "PASS" : "FAIL");
    return max_err < 1e-3 ? 0 : 2;
}
