// Tile-major MLP chain machinery shared by the SDF / albedo / NeRF kernels.
//
// One CTA = 10 warps = 320 threads, two CTAs co-resident per SM (112 KB smem + 256 TMEM columns each), so one
// CTA's tcgen05.mma phase overlaps the other's epilogue without explicit ping-pong code.
//   warp 0   : weight producer  -- cp.async.bulk (UBLKCP) of pre-packed K=32 weight slices into a 3-slot ring
//   warp 1   : MMA issuer       -- lane 0 issues tcgen05.mma (M=128, N<=256, K=16) from smem A x smem W into TMEM
//   warps 2-9: epilogue         -- two threads per point row (TMEM lane), each owning one 128-column half of the
//                                  accumulator: tcgen05.ld in 16-column chunks with the next chunk's load in flight,
//                                  bias + activation (+ derivative terms) in registers, write the next layer's A
//                                  operand (fp16, "chunked" K-major image, see common.cuh) into smem and any streams
//                                  to HBM.  Eight epilogue warps per CTA = four per SM sub-partition with both CTAs
//                                  resident: the thread-level parallelism that hides the TMEM / L2 latencies.
// A "chain" is a table of GEMM steps executed strictly in order for each 128-point tile; step s+1's A operand
// is written by step s's epilogue.
#pragma once
#include "common.cuh"

namespace rnb {

constexpr int TILE_M = 128;
#ifndef RNB_SLICE_K
#define RNB_SLICE_K 32
#endif
constexpr int SLICE_K = RNB_SLICE_K;      // K extent of one ring stage: 32, or 16 (twice as many stages of half the size)
constexpr int RING_STAGES = 3 * 32 / SLICE_K;
constexpr int STAGE_BYTES = 512 * SLICE_K;   // one K slice of a 256-row operand (16 KB at K = 32)
constexpr int CHAIN_THREADS = 320;
constexpr int EPI_THREADS = 256;
constexpr int EPI_HALF_COLS = 128;     // accumulator columns owned by one epilogue thread
constexpr int MAX_STEPS = 24;
constexpr int TMEM_COLS = 256;

struct ChainStep {
    uint32_t w_off;   // byte offset of the packed [n x k] fp16 weight image in the blob
    uint16_t n;       // UMMA N (rows of the weight image), multiple of 16, <= 256
    uint16_t k;       // K, multiple of 32
    uint32_t accumulate;   // 1: add onto the accumulator left by the previous step (split-K over two A operands)
    // 256-wide fp16 streams this step's EPILOGUE reads (0 = none): the producer warp pulls the tile's 64 KB of each
    // into L2 one step ahead with cp.async.bulk.prefetch, so the per-row loads of the epilogue are L2 hits
    const uint8_t* pf[3];
};
struct ChainTable {
    int n_steps;
    int weights_evict_last;   // load the weight slices with an L2 evict-last policy (pays off only under heavy stream traffic)
    ChainStep steps[MAX_STEPS];
    unsigned long long* trace;   // RNB_TRACE builds only: CTA 0 records clock64() timestamps (see profiles/_trace_chain.py)
};
#ifdef RNB_TRACE
#define RNB_TR(role, idx, k, cond) do { if ((cond) && tab_trace && blockIdx.x == 0 && (idx) < 4096) tab_trace[((role) * 4096 + (idx)) * 4 + (k)] = clock64(); } while (0)
#else
#define RNB_TR(role, idx, k, cond) do { } while (0)
#endif

struct ChainSmem {
    uint8_t* sA;
    uint8_t* ring;
    uint64_t* full;     // [RING_STAGES]
    uint64_t* empty;    // [RING_STAGES]
    uint64_t* acc_full;
    uint64_t* a_ready;
    uint32_t* tmem_slot;
};

__host__ __device__ constexpr int chain_smem_bytes(int a_cols) {
    return a_cols * TILE_M * 2 + RING_STAGES * STAGE_BYTES + 128;
}

__device__ __forceinline__ ChainSmem chain_carve(uint8_t* smem, int a_cols) {
    ChainSmem s;
    s.sA = smem;
    s.ring = smem + a_cols * TILE_M * 2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s.ring + RING_STAGES * STAGE_BYTES);
    s.full = bars;
    s.empty = bars + RING_STAGES;
    s.acc_full = bars + 2 * RING_STAGES;
    s.a_ready = bars + 2 * RING_STAGES + 1;
    s.tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * RING_STAGES + 2);
    return s;
}

// Called by all threads at kernel start.  Returns the TMEM base address.
__device__ __forceinline__ uint32_t chain_setup(const ChainSmem& s) {
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING_STAGES; ++i) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], 1);
        }
        mbar_init(s.acc_full, 1);
        mbar_init(s.a_ready, EPI_THREADS);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(s.tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *s.tmem_slot;
}

__device__ __forceinline__ void chain_teardown(const ChainSmem& s, uint32_t tmem) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 1) tmem_dealloc(tmem, TMEM_COLS);
}

// Which tiles a (virtual) CTA owns: first, first + stride, ... (n of them).  Plain kernels: one CTA per blockIdx; the
// fused backward (sdf_chain.cu) runs two virtual CTAs per block and gives every one its own map.
struct TileMap {
    int64_t first, stride;
    int n;
};
__device__ __forceinline__ TileMap tilemap_grid(int n_tiles) {
    TileMap m;
    m.first = blockIdx.x;
    m.stride = gridDim.x;
    m.n = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    return m;
}

// warp 0, one lane
__device__ __forceinline__ void chain_prefetch_step(const ChainTable& tab, int st, int64_t tile) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint8_t* b = tab.steps[st].pf[k];
        if (b) {
            bulk_prefetch_l2(b + (size_t)tile * 65536, 32768);
            bulk_prefetch_l2(b + (size_t)tile * 65536 + 32768, 32768);
        }
    }
}

__device__ __forceinline__ void chain_producer(const ChainSmem& s, const ChainTable& tab, const uint8_t* wblob, const TileMap tm) {
    uint32_t it = 0;
    unsigned long long* tab_trace = tab.trace; (void)tab_trace;
    const uint64_t keep = l2_policy_evict_last();
    for (int t = 0; t < tm.n; ++t) {
        const int64_t tile = tm.first + (int64_t)t * tm.stride;
        for (int st = 0; st < tab.n_steps; ++st) {
            // The producer reaches step st's first weight slice while MMA(st-1) is still running, i.e. about one
            // epilogue + one MMA phase before epilogue(st) reads these streams: enough to cover DRAM latency, short
            // enough that the lines are still in L2 (a longer lead measurably doubled DRAM reads).
            chain_prefetch_step(tab, st, tile);
            const uint32_t bytes = (uint32_t)(2 * SLICE_K) * tab.steps[st].n;   // SLICE_K/8 chunks x n rows x 16 B
            const uint8_t* src = wblob + tab.steps[st].w_off;
            const int nsl = tab.steps[st].k / SLICE_K;
            for (int ks = 0; ks < nsl; ++ks, ++it) {
                const uint32_t slot = it % RING_STAGES, ph = (it / RING_STAGES) & 1;
                RNB_TR(0, it, 0, true);
                mbar_wait(&s.empty[slot], ph ^ 1);
                RNB_TR(0, it, 1, true);
#if defined(RNB_TRACE) || defined(RNB_DBG_HOOKS)
                if (tab.weights_evict_last & 2) { mbar_arrive(&s.full[slot]); continue; }   // experiment: no copies, stale slot contents
#endif
                mbar_expect_tx(&s.full[slot], bytes);
                if (tab.weights_evict_last & 1) bulk_g2s_hint(s.ring + slot * STAGE_BYTES, src + (size_t)ks * bytes, bytes, &s.full[slot], keep);
                else bulk_g2s(s.ring + slot * STAGE_BYTES, src + (size_t)ks * bytes, bytes, &s.full[slot]);
                RNB_TR(0, it, 2, true);
            }
        }
    }
}
__device__ __forceinline__ void chain_producer(const ChainSmem& s, const ChainTable& tab, const uint8_t* wblob, int n_my_tiles) {
    TileMap tm = tilemap_grid(0);
    tm.n = n_my_tiles;
    chain_producer(s, tab, wblob, tm);
}

// warp 1, ALL lanes (convergent): the waits are warp-wide, one elected lane issues.  Compared with running the whole
// loop inside `if (lane == 0)`, loop counters and descriptors stay in uniform registers and each tcgen05 instruction
// needs no per-instruction elect loop; ring slot / phase are carried incrementally (no division by RING_STAGES) and
// the operand descriptors are formed by adding to per-step bases (address field = bytes >> 4; smem addresses stay
// below 2^18, so the adds never carry out of the 14-bit field).  The issue loop, not the weight ring, paced the MMA
// phase before (profiles/r01_notes.md: clock64 trace of one CTA).
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void chain_mma_warp(const ChainSmem& s, const ChainTable& tab, uint32_t tmem, int n_my_tiles) {
    uint32_t slot = 0, ph = 0, sig = 0;
    const uint64_t a_desc0 = umma_desc(smem_u32(s.sA), TILE_M * 16, 128);
    const uint32_t ring_base = smem_u32(s.ring);
#ifdef RNB_TRACE
    unsigned long long* tab_trace = tab.trace;
    uint32_t tr_it = 0;
    const bool tr_lane = (threadIdx.x & 31) == 0;
#endif
    for (int t = 0; t < n_my_tiles; ++t) {
        for (int st = 0; st < tab.n_steps; ++st, ++sig) {
            const uint32_t n = tab.steps[st].n;
            const uint32_t idesc = umma_idesc(TILE_M, n, FMT_F16, FMT_F16);
            const int nsl = tab.steps[st].k / SLICE_K;
            const uint64_t b_desc0 = umma_desc(ring_base, n * 16, 128);
            const uint32_t b_slab = (2 * n * 16) >> 4;          // one K=16 slab of the stage, in descriptor units
            const uint32_t acc0 = tab.steps[st].accumulate;
            RNB_TR(2, sig, 0, tr_lane);
            mbar_wait(s.a_ready, sig & 1);
            RNB_TR(2, sig, 1, tr_lane);
            tc_fence_after();
            for (int ks = 0; ks < nsl; ++ks) {
                RNB_TR(1, tr_it, 0, tr_lane);
                mbar_wait(&s.full[slot], ph);
                RNB_TR(1, tr_it, 1, tr_lane);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < SLICE_K / 16; ++j) {
                        const uint64_t ad = a_desc0 + (uint64_t)((uint32_t)(ks * (SLICE_K / 16) + j) * ((2 * TILE_M * 16) >> 4));
                        const uint64_t bd = b_desc0 + (uint64_t)(slot * (STAGE_BYTES >> 4) + (uint32_t)j * b_slab);
                        umma_f16(tmem, ad, bd, idesc, ((ks | j) != 0) || acc0);
                    }
                    umma_commit(&s.empty[slot]);
                }
                __syncwarp();
                RNB_TR(1, tr_it, 2, tr_lane);
#ifdef RNB_TRACE
                ++tr_it;
#endif
                if (++slot == RING_STAGES) { slot = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(s.acc_full);
            __syncwarp();
            RNB_TR(2, sig, 2, tr_lane);
        }
    }
}

// ---- activations (fp32, MUFU based)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// softplus(beta=100): a = log(1+exp(100 z))/100  (reference models/fields.py:80).  The ATen threshold branch
// (100 z > 20 -> a = z) differs from this closed form by < 2.1e-11, far below fp32 resolution of a.
//   a = max(z, 0) + log1p(q) / 100,   q = exp(-100 |z|) in (0, 1]
// One MUFU (ex2) per element: log1p(q) = q * P4(q) with a degree-4 minimax polynomial on the FMA pipe (relative error
// 1.2e-4, i.e. < 8.1e-7 absolute on a -- below half an fp16 ulp of every activation the correction matters for).
// The XU pipe (16 lanes/clk/SM) is what bounds the forward layers; a second MUFU (lg2) would double that floor.
__device__ __forceinline__ float softplus100_corr(float q) {
    // coefficients of log1p(q)/q, pre-divided by 100
    float p = fmaf(q, 0.04106372e-2f, -0.15602615e-2f);
    p = fmaf(q, p, 0.30467027e-2f);
    p = fmaf(q, p, -0.49636758e-2f);
    p = fmaf(q, p, 0.99988786e-2f);
    return p;
}
__device__ __forceinline__ float softplus100(float z) {
    const float q = ex2_approx(-144.26950408889634f * fabsf(z));
    return fmaf(q, softplus100_corr(q), fmaxf(z, 0.f));
}
// ---- packed fp32 (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE fp32 lanes per instruction).  The epilogues are bound by
// issue slots, not by the FMA pipe, so doing the bias add, the exponent scaling and the log1p polynomial on pairs of
// columns removes ~7 of 19 instructions per pair with bit-identical results (same operations, same rounding).
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 f2_splat(float v) { return make_float2(v, v); }
// softplus100 of two columns; same arithmetic as softplus100() lane by lane
__device__ __forceinline__ float2 softplus100_x2(float2 z) {
    const float2 y = f2_mul(z, f2_splat(144.26950408889634f));
    float2 q;
    q.x = ex2_approx(-fabsf(y.x));
    q.y = ex2_approx(-fabsf(y.y));
    float2 p = f2_fma(q, f2_splat(0.04106372e-2f), f2_splat(-0.15602615e-2f));
    p = f2_fma(q, p, f2_splat(0.30467027e-2f));
    p = f2_fma(q, p, f2_splat(-0.49636758e-2f));
    p = f2_fma(q, p, f2_splat(0.99988786e-2f));
    return f2_fma(q, p, make_float2(fmaxf(z.x, 0.f), fmaxf(z.y, 0.f)));
}

// softplus'(z) recovered from a = softplus(z):  a = log(1 + e^{100 z}) / 100  =>  sigmoid(100 z) = 1 - e^{-100 a}.
// Lets the backward kernels read the activation stream they need anyway instead of a separate s stream.
__device__ __forceinline__ float sig_from_a(float a) { return 1.f - ex2_approx(-144.26950408889634f * a); }
// (1 - exp(-100 a)) * u on two columns: same operations and rounding as sig_from_a(a) * u lane by lane
__device__ __forceinline__ float2 sigmul_x2(float2 a, float2 u) {
    const float2 y = f2_mul(a, f2_splat(144.26950408889634f));
    float2 e;
    e.x = ex2_approx(-y.x);
    e.y = ex2_approx(-y.y);
    return f2_mul(f2_fma(e, f2_splat(-1.f), f2_splat(1.f)), u);
}


// per-thread epilogue context (warps 2..9)
struct Epi {
    uint8_t* sA;
    uint64_t* acc_full;
    uint64_t* a_ready;
    uint32_t tmem_row;   // TMEM address of this warp's lane quadrant, column 0
    uint32_t acc_cnt;
    int row;             // 0..127 = TMEM lane = point row inside the tile
    int half;            // 0: accumulator columns 0..127, 1: columns 128..255 (warp-uniform)
    int col0;            // = half * 128
    int bar_id;          // named barrier of this (virtual) CTA's 256 epilogue threads

    // local_warp: warp index inside the (virtual) CTA, 2..9 for epilogue warps; the TMEM lane quadrant follows the
    // hardware warp id (threadIdx.x >> 5) & 3 -- each group of four consecutive epilogue warps covers all four.
    __device__ __forceinline__ void init(const ChainSmem& s, uint32_t tmem, int local_warp = -1, int barrier_id = 1) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int quad = warp & 3;           // a warp may only touch TMEM lanes 32*(warp%4) .. +31
        if (local_warp < 0) local_warp = warp;
        sA = s.sA;
        acc_full = s.acc_full;
        a_ready = s.a_ready;
        tmem_row = tmem + ((uint32_t)(quad * 32) << 16);
        row = quad * 32 + lane;
        half = (local_warp - 2) >> 2;
        col0 = half * EPI_HALF_COLS;
        acc_cnt = 0;
        bar_id = barrier_id;
        tab_trace = nullptr;
    }
    unsigned long long* tab_trace;
    __device__ __forceinline__ void wait_acc() {
        RNB_TR(3, acc_cnt, 0, threadIdx.x == 64);
        mbar_wait(acc_full, acc_cnt & 1);
        RNB_TR(3, acc_cnt, 1, threadIdx.x == 64);
        ++acc_cnt;
        tc_fence_after();
    }
    // "A operand written, accumulator drained": lets the MMA warp start the next step
    __device__ __forceinline__ void signal() {
        tc_fence_before();
        fence_proxy_async();
        RNB_TR(3, acc_cnt, 2, threadIdx.x == 64);
        mbar_arrive(a_ready);
    }
    // rendezvous of the 256 epilogue threads (named barrier bar_id; the producer / MMA warps never join)
    __device__ __forceinline__ void sync_epi() const { asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory"); }
    __device__ __forceinline__ void st_a(int chunk, uint4 v) const {
        *reinterpret_cast<uint4*>(sA + ((size_t)chunk * TILE_M + row) * 16) = v;
    }
    __device__ __forceinline__ uint4 ld_a(int chunk) const {
        return *reinterpret_cast<const uint4*>(sA + ((size_t)chunk * TILE_M + row) * 16);
    }
    __device__ __forceinline__ void st_a_half(int col, __half h) const {
        *reinterpret_cast<__half*>(sA + ((size_t)(col >> 3) * TILE_M + row) * 16 + (col & 7) * 2) = h;
    }
    // Per-row exchange slot (4 floats) between the two threads of a row, inside the A buffer at this row's own
    // 16 bytes of chunk 8 (a chunk of half 0).  Protocol: A must be dead (its GEMM completed, acc_full seen);
    // the half-1 thread writes, sync_epi(), the half-0 thread reads -- and only afterwards overwrites chunk 8 in
    // its own sweep, so one barrier suffices.
    __device__ __forceinline__ float* xchg() const {
        return reinterpret_cast<float*>(sA + ((size_t)8 * TILE_M + row) * 16);
    }
    __device__ __forceinline__ void ld_acc16(int c0, uint32_t (&v)[16]) const {
        tmem_ld16(tmem_row + (uint32_t)c0, v);
        tmem_ld_wait16(v);
    }
    // Sweep NCH 16-column chunks starting at column c_begin: f(c, v) gets the first column of the chunk and its 16
    // fp32 accumulator values (as bits).  The TMEM load of chunk i+1 is in flight while chunk i is processed.
    template <int NCH = 8, class F>
    __device__ __forceinline__ void sweep(int c_begin, F&& f) const {
        static_assert(NCH % 2 == 0, "even chunk count");
        uint32_t va[16], vb[16];
        tmem_ld16(tmem_row + (uint32_t)c_begin, va);
#pragma unroll 1
        for (int i = 0; i < NCH / 2; ++i) {
            const int c = c_begin + i * 32;
            tmem_ld_wait16(va);
            tmem_ld16(tmem_row + (uint32_t)(c + 16), vb);
            f(c, va);
            tmem_ld_wait16(vb);
            if (i + 1 < NCH / 2) tmem_ld16(tmem_row + (uint32_t)(c + 32), va);
            f(c + 16, vb);
        }
    }
    template <class F>
    __device__ __forceinline__ void sweep_half(F&& f) const { sweep<8>(col0, f); }
    // sweep_half with a per-column fp32 bias added on the fly: f(c0, z) gets z[k] = acc[c0 + k] + bias[c0 + k].  The
    // bias loads are software-pipelined one chunk ahead of their use, so their L1 round trip overlaps the previous
    // chunk's arithmetic instead of stalling the first FADD of every chunk.
    template <class F>
    __device__ __forceinline__ void sweep_half_bias(const float* bias, F&& f) const {
        float4 nb[4];
        const float4* sp = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
        for (int k = 0; k < 4; ++k) nb[k] = __ldg(sp + k);
        sweep<8>(col0, [&](int c0, const uint32_t (&v)[16]) {
            float z[16];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 z0 = f2_add(make_float2(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1])), make_float2(nb[k].x, nb[k].y));
                const float2 z1 = f2_add(make_float2(__uint_as_float(v[4 * k + 2]), __uint_as_float(v[4 * k + 3])), make_float2(nb[k].z, nb[k].w));
                z[4 * k] = z0.x; z[4 * k + 1] = z0.y; z[4 * k + 2] = z1.x; z[4 * k + 3] = z1.y;
            }
            if (c0 + 16 < col0 + EPI_HALF_COLS) {
                const float4* np = reinterpret_cast<const float4*>(bias + c0 + 16);
#pragma unroll
                for (int k = 0; k < 4; ++k) nb[k] = __ldg(np + k);
            }
            f(c0, z);
        });
    }
};

// ---- global "stream" images: [n_pts/64 subtiles][C/8 chunks][64 rows][16 B]  (data layout)
__device__ __forceinline__ size_t stream_off(int64_t p, int chunk, int nchunks) {
    return ((size_t)((p >> 6) * nchunks + chunk) * 64 + (size_t)(p & 63)) * 16;
}
// 256-wide streams with the point-dependent part of the address hoisted: row = stream_row(p) once per tile, then
// address = base + row + chunk * 1024 (the chunk term folds into the load / store's immediate offset where the chunk index
// is a compile-time constant, and is one shift-add otherwise; the general form cost ~6 integer instructions per access)
__device__ __forceinline__ size_t stream_row(int64_t p) { return (size_t)(p >> 6) * 32768 + (size_t)(p & 63) * 16; }
__device__ __forceinline__ uint4 ld_stream_r(const uint8_t* base_row, int chunk) {
    return *reinterpret_cast<const uint4*>(base_row + (size_t)chunk * 1024);
}
__device__ __forceinline__ void st_stream_r(uint8_t* base_row, int chunk, uint4 v, bool keep = false) {
    uint4* q = reinterpret_cast<uint4*>(base_row + (size_t)chunk * 1024);
    if (keep) *q = v;
    else __stcs(q, v);
}
// streams are written once and read by a later kernel (or many steps later): evict-first (.cs) keeps them from
// pushing the prefetched operands of the next step out of L2
__device__ __forceinline__ void st_stream(uint8_t* base, int64_t p, int chunk, int nchunks, uint4 v) {
    __stcs(reinterpret_cast<uint4*>(base + stream_off(p, chunk, nchunks)), v);
}
__device__ __forceinline__ uint4 ld_stream(const uint8_t* base, int64_t p, int chunk, int nchunks) {
    return *reinterpret_cast<const uint4*>(base + stream_off(p, chunk, nchunks));
}

// Pull this thread's part of the NEXT step's stream tile towards L2 a whole step ahead (no registers held, no
// completion to wait for).  A warp's 32 rows x 16 B of one chunk are 512 contiguous bytes = four 128-byte lines:
// one lane in eight issues the hint.  The epilogue loads then hit L2 instead of waiting a DRAM round trip per chunk.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 128 consecutive floats (4 lines) towards L1 ahead of un-pipelined __ldg reads
__device__ __forceinline__ void prefetch_l1_row128(const float* p) {
#pragma unroll
    for (int k = 0; k < 4; ++k) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 32 * k));
}
__device__ __forceinline__ void prefetch_stream_chunks(const uint8_t* base, int64_t p, int chunk0, int n_chunks) {
    if ((threadIdx.x & 7) == 0) {
#pragma unroll 4
        for (int k = 0; k < n_chunks; ++k) prefetch_l2(base + stream_off(p, chunk0 + k, 32));
    }
}

}  // namespace rnb
