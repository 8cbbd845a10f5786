// extern "C" boundary of librnb_b200.so (declarations and contracts: include/rnb_b200.h).
#include "../../include/rnb_b200.h"
#include "sdf_params.h"
#include "dw_params.h"
#include "render_params.h"
#include "albedo_params.h"
#include "nerf_params.h"
#include <algorithm>
#include <vector>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include <mutex>

namespace rnb {
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
cudaError_t launch_sdf_pack(const float* const* W, const float* const* b, uint8_t* blob, float* aux, cudaStream_t st);
cudaError_t launch_sdf_fwd(const SdfFwdParams& P, int sm_count, cudaStream_t st);
cudaError_t launch_sdf_fwd_grad(const SdfFwdGradParams& P, int sm_count, cudaStream_t st);
cudaError_t launch_sdf_bwd_data(const SdfBwdParams& P, int sm_count, cudaStream_t st);
cudaError_t launch_sdf_bwd_fused(const SdfBwdFusedParams& P, int sm_count, cudaStream_t st);
cudaError_t launch_dw_gemm(const DwParams& P, int splits, cudaStream_t st);
cudaError_t launch_reduce(const ReduceParams& P, cudaStream_t st);
cudaError_t launch_absmax(const float* a, int64_t na, const float* b, int64_t nb, const float* c, int64_t nc,
                          const float* fold, float* out, cudaStream_t st);
cudaError_t launch_coarse_z(const float* near, const float* far, const float* t_rand, float* z, int n_rays, int n_samples,
                            cudaStream_t st);
cudaError_t launch_upsample(const UpsampleParams& P, cudaStream_t st);
cudaError_t launch_sample_pdf_from_cdf(const float* bins, const float* cdf, int n_rays, int n, int n_new, float* samples,
                                       int64_t* inds, cudaStream_t st);
cudaError_t launch_final_merge(const float* z_old, int n_old, const float* z_new, int n_new, int n_rays, float sample_dist,
                               float* z_out, float* mid_out, cudaStream_t st);
cudaError_t launch_composite(const CompositeParams& P, bool bwd, cudaStream_t st);
cudaError_t launch_albedo_pack(const float* W0, const float* b0, const float* W1, const float* b1, const float* W2,
                               const float* b2, uint8_t* blob, float* aux, cudaStream_t st);
cudaError_t launch_albedo_fwd(const AlbedoFwdParams& P, int sm_count, cudaStream_t st);
cudaError_t launch_albedo_bwd(const AlbedoBwdParams& P, int sm_count, cudaStream_t st);
cudaError_t launch_nerf_pack(const float* const* W, const float* const* b, const float* Wf, const float* bf, const float* Wa,
                             const float* ba, const float* Wv, const float* bv, const float* Wr, const float* br,
                             uint8_t* blob, float* aux, cudaStream_t st);
cudaError_t launch_nerf_fwd(const NerfFwdParams& P, int sm_count, cudaStream_t st);
cudaError_t launch_composite_bg(const rnb_composite_bg_t& P, cudaStream_t st);
cudaError_t launch_ray_batch(const rnb_ray_batch_t& P, cudaStream_t st);
cudaError_t launch_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                             double eps, int64_t step, double grad_scale, int sm_count, cudaStream_t st);
cudaError_t launch_wn_fold(const rnb_wn_layer_t* layers, int n, cudaStream_t st);
cudaError_t launch_wn_vjp(const rnb_wn_layer_t* layers, int n, cudaStream_t st);
cudaError_t launch_mc_count(const float* u, int nx, int ny, int nz, float thr, const int8_t* tri_count, int32_t* counts, cudaStream_t st);
cudaError_t launch_mc_emit(const float* u, int nx, int ny, int nz, float thr, const int8_t* tri_table, const int64_t* offsets,
                           int x_global0, float* verts, int64_t* keys, cudaStream_t st);
cudaError_t launch_stream_from_rowmajor(const float* x, int64_t n, int cols, int64_t n_pad, uint8_t* out, cudaStream_t st);

struct AlbedoBwdScratch {
    size_t absmax, sum_part, dz2, dz1, dz0, dw_part, dwcs_part, cs_part, total;
    int dw_splits;
};
static AlbedoBwdScratch albedo_bwd_scratch(int64_t n_pts) {
    AlbedoBwdScratch L;
    const size_t s256 = rnb_stream_bytes(n_pts, 256);
    const int64_t n_pad = rnb_padded_points(n_pts);
    const int n_sub = (int)(n_pad / 64);
    L.dw_splits = std::max(1, std::min(n_sub, 96));
    size_t o = 0;
    L.absmax = o; o += 256;
    L.sum_part = o; o += 3 * 256;
    L.dz2 = o; o += align_up((size_t)n_pad * 32, 256);          // fp16 stream [n_pad x 16]
    L.dz1 = o; o += s256;
    L.dz0 = o; o += s256;
    o = align_up(o, 256);
    L.dw_part = o; o += (size_t)L.dw_splits * 256 * (256 + 256 + 64 + 16) * 4;
    L.dwcs_part = o; o += (size_t)2 * L.dw_splits * 256 * 2 * 4;      // two column-sum partials per split (dw_gemm_kernel)
    L.cs_part = o; o += (size_t)2 * L.dw_splits * (3 * 256 + 4) * 4;  // W_2: three weighted sums + the weight sums
    L.total = align_up(o, 256);
    return L;
}

// scratch layout of rnb_sdf_bwd
struct SdfBwdScratch {
    size_t absmax, sum_part, uin0, uin, zbar, dfeat, dw_part, dwcs_part, cs_part, queue, total;
    int dw_splits;
};
static SdfBwdScratch sdf_bwd_scratch(int64_t n_pts) {
    SdfBwdScratch L;
    const size_t s256 = rnb_stream_bytes(n_pts, 256), s64 = rnb_stream_bytes(n_pts, 64);
    const int n_sub = (int)(rnb_padded_points(n_pts) / 64);
    L.dw_splits = std::max(1, std::min(n_sub, 48));
    size_t o = 0;
    L.absmax = o; o += 256;
    L.sum_part = o; o += 256;
    L.uin0 = o; o += s64;
    L.uin = o; o += 8 * s256;
    L.zbar = o; o += 8 * s256;
    L.dfeat = o; o += s256;
    o = align_up(o, 256);
    L.dw_part = o; o += (size_t)L.dw_splits * 256 * (64 + 8 * 256) * 4;
    // column-sum partials: two per split (the two warp groups of dw_gemm_kernel each sum half of the point rows)
    L.dwcs_part = o; o += (size_t)2 * L.dw_splits * 256 * 9 * 4;
    L.cs_part = o; o += (size_t)2 * L.dw_splits * (2 * 256 + 4) * 4;  // sdf row of W_8: two sums + the weight sum
    o = align_up(o, 256);
    // fused backward: counters (head[9], tail[9], padded to 64 ints) + the per-layer hand-over queues [9][n_tiles]
    L.queue = o; o += (64 + (size_t)18 * (rnb_padded_points(n_pts) / TILE_M)) * 4 + 2 * FUSED_MAX_WORKERS * 8;
    L.total = align_up(o, 256);
    return L;
}

// Workers of the fused backward per layer (layer 0 and 8 contract less: N = 64 / one stream pair).  RNB_DW_REPL overrides
// ("3,4,4,4,4,4,4,4,2"); the sum is the number of SMs taken from the chain.  Each count is clamped to dw_splits.
static void fused_replicas(int (&rep)[9], int max_rep) {
    static const int def[9] = {3, 4, 4, 4, 4, 4, 4, 4, 2};
    for (int l = 0; l < 9; ++l) rep[l] = def[l];
    if (const char* e = getenv("RNB_DW_REPL")) {
        int l = 0;
        const char* p = e;
        while (*p && l < 9) {
            rep[l++] = atoi(p);
            while (*p && *p != ',') ++p;
            if (*p == ',') ++p;
        }
    }
    for (int l = 0; l < 9; ++l) rep[l] = std::max(1, std::min(rep[l], std::min(max_rep, 16)));
}


// ---- optional per-kernel timing (cudaEvents on the launch stream) and a launch counter ------------------------
enum ProfTag { T_SDF_PACK, T_SDF_FWD, T_SDF_FWD_GRAD, T_SDF_BWD_DATA, T_SDF_BWD_FUSED, T_DW_GEMM, T_REDUCE, T_ABSMAX,
               T_COARSE_Z, T_UPSAMPLE, T_FINAL_MERGE, T_COMPOSITE_FWD, T_COMPOSITE_BWD, T_ALBEDO_PACK, T_ALBEDO_FWD,
               T_ALBEDO_BWD, T_SAMPLE_PDF, T_NERF_PACK, T_NERF_FWD, T_COMPOSITE_BG, T_RAY_BATCH, T_MC, T_ADAM, T_WNORM, T_COUNT };
static const char* const kProfNames[T_COUNT] = {
    "sdf_pack", "sdf_fwd", "sdf_fwd_grad", "sdf_bwd_data", "sdf_bwd_fused", "dw_gemm", "reduce", "absmax", "coarse_z",
    "upsample", "final_merge", "composite_fwd", "composite_bwd", "albedo_pack", "albedo_fwd", "albedo_bwd", "sample_pdf", "nerf_pack", "nerf_fwd", "composite_bg", "ray_batch", "marching_cubes", "adam", "weight_norm"};
struct ProfRec { int tag; cudaEvent_t a, b; };
// ctypes releases the GIL during a call and the backward runs on the autograd thread: the instrumentation state is shared
// between threads, so the counters are atomic and the event lists sit behind a mutex
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_ev_pool;
static std::atomic<long long> g_launches[T_COUNT];
static cudaEvent_t prof_event() {       // caller holds g_prof_mu
    cudaEvent_t e;
    if (!g_ev_pool.empty()) { e = g_ev_pool.back(); g_ev_pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}
template <class F>
static cudaError_t profiled(int tag, cudaStream_t st, F&& f) {
    g_launches[tag].fetch_add(1, std::memory_order_relaxed);
    if (!g_prof_on.load(std::memory_order_relaxed)) return f();
    ProfRec r;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        r = ProfRec{tag, prof_event(), prof_event()};
    }
    cudaEventRecord(r.a, st);
    cudaError_t e = f();
    cudaEventRecord(r.b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
    return e;
}

static int sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cache[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cache[dev] = n > 0 ? n : 148;
    }
    return cache[dev];
}

static SdfPointSource to_src(const rnb_points_t* p) {
    SdfPointSource s;
    s.n_pts = p->n_pts; s.x = p->x; s.rays_o = p->rays_o; s.rays_d = p->rays_d; s.z = p->z;
    s.n_per_ray = p->n_per_ray; s.grid_res = p->grid_res; s.slab_x0 = p->slab_x0;
    for (int a = 0; a < 3; ++a) { s.bmin[a] = p->bmin[a]; s.bmax[a] = p->bmax[a]; }
    return s;
}

static void add_step(ChainTable& t, uint32_t off, int n, int k, int accumulate = 0, const void* pf0 = nullptr,
                     const void* pf1 = nullptr, const void* pf2 = nullptr) {
    t.steps[t.n_steps].accumulate = (uint32_t)accumulate;
    static const int pf_mode = getenv("RNB_PF") ? atoi(getenv("RNB_PF")) : 7;   // bit0: first stream, bit1: second
    t.steps[t.n_steps].pf[0] = (pf_mode & 1) ? (const uint8_t*)pf0 : nullptr;
    t.steps[t.n_steps].pf[1] = (pf_mode & 2) ? (const uint8_t*)pf1 : nullptr;
    t.steps[t.n_steps].pf[2] = (pf_mode & 4) ? (const uint8_t*)pf2 : nullptr;
    t.steps[t.n_steps].w_off = off;
    t.steps[t.n_steps].n = (uint16_t)n;
    t.steps[t.n_steps].k = (uint16_t)k;
    ++t.n_steps;
}
static void table_forward(ChainTable& t) {           // layers 0..7
    add_step(t, sdfw_fwd(0), 256, 64);
    for (int l = 1; l < 8; ++l) add_step(t, sdfw_fwd(l), 256, 256);
}
static int n_tiles(int64_t n) { return (int)((n + TILE_M - 1) / TILE_M); }
#ifdef RNB_TRACE
static unsigned long long* g_trace = nullptr;   // device buffer [4 roles][4096][4] of clock64() stamps, CTA 0
#endif
}  // namespace rnb

using namespace rnb;

extern "C" {

#ifdef RNB_TRACE
__attribute__((visibility("default"))) void rnb_trace_set(void* p) { g_trace = (unsigned long long*)p; }
#endif
const char* rnb_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }
int rnb_version(void) { return 100; }
size_t rnb_sdf_wblob_bytes(void) { return SDFW_BYTES; }
size_t rnb_sdf_aux_floats(void) { return AUX_FLOATS; }
int64_t rnb_padded_points(int64_t n) { return (n + TILE_M - 1) / TILE_M * TILE_M; }
size_t rnb_stream_bytes(int64_t n, int cols) { return (size_t)rnb_padded_points(n) * (size_t)cols * 2; }

int rnb_sdf_pack(const float* const* W, const float* const* b, void* wblob, float* aux, void* stream) {
    return (int)profiled(T_SDF_PACK, (cudaStream_t)stream, [&] { return launch_sdf_pack(W, b, (uint8_t*)wblob, aux, (cudaStream_t)stream); });
}

int rnb_sdf_fwd(const rnb_points_t* pts, const void* wblob, const float* aux, float* out, float out_scale, void* stream) {
    SdfFwdParams P{};
    P.src = to_src(pts);
    P.n_tiles = n_tiles(pts->n_pts);
    P.wblob = (const uint8_t*)wblob;
    P.aux = aux;
    table_forward(P.tab);
#ifdef RNB_TRACE
    P.tab.trace = g_trace;
#endif
#if defined(RNB_TRACE) || defined(RNB_DBG_HOOKS)
    if (getenv("RNB_DBG_NOFILL")) P.tab.weights_evict_last |= 2;
#endif
    P.out = out;
    P.out_scale = out_scale;
    return (int)profiled(T_SDF_FWD, (cudaStream_t)stream, [&] { return launch_sdf_fwd(P, sm_count(), (cudaStream_t)stream); });
}

int rnb_sdf_fwd_grad(const rnb_points_t* pts, const void* wblob, const float* aux, float* out_sdf, float* out_grad,
                     float* out_full, void* st_feat, void* st_in0, void* st_in, void* st_w, void* stream) {
    SdfFwdGradParams P{};
    P.src = to_src(pts);
    P.n_tiles = n_tiles(pts->n_pts);
    P.wblob = (const uint8_t*)wblob;
    P.aux = aux;
    table_forward(P.tab);
#ifdef RNB_TRACE
    P.tab.trace = g_trace;
#endif
    // a_5 and a_6 are re-read by the dx-chain only 3-5 steps after they were written: stored evict-first they are dropped
    // from L2 right after the producer's prefetch found them still present, and the epilogue then waits for DRAM
    // (clock64 trace: those two steps 10.6 k / 9.0 k -> 7.3 k / 7.0 k cycles, the tile 228 k -> 215 k).  Older layers
    // keep the evict-first policy: keeping more (0x78, 0x7f) pushes the later steps' prefetched lines out instead.
    { static const int km = getenv("RNB_K2_KEEP") ? (int)strtol(getenv("RNB_K2_KEEP"), nullptr, 0) : 0x60; P.keep_mask = km; }
    add_step(P.tab, SDFW_F8, 256, 256);
    {
        const size_t SS = rnb_stream_bytes(pts->n_pts, 256);
        for (int l = 7; l >= 1; --l) add_step(P.tab, sdfw_tr(l), 256, 256, 0, (const uint8_t*)st_in + (size_t)(l - 1) * SS);
    }
    add_step(P.tab, sdfw_tr(0), 64, 256);
    P.out_sdf = out_sdf; P.out_grad = out_grad; P.out_full = out_full;
    P.st_feat = (uint8_t*)st_feat; P.st_in0 = (uint8_t*)st_in0; P.st_in = (uint8_t*)st_in;
    P.st_w = (uint8_t*)st_w;
    P.stream_stride = rnb_stream_bytes(pts->n_pts, 256);
    return (int)profiled(T_SDF_FWD_GRAD, (cudaStream_t)stream, [&] { return launch_sdf_fwd_grad(P, sm_count(), (cudaStream_t)stream); });
}


size_t rnb_sdf_bwd_scratch_bytes(int64_t n_pts) { return sdf_bwd_scratch(n_pts).total; }
// byte offset, inside the scratch buffer, of the fused backward's optional block time stamps (RNB_FUSED_DBG=1):
// uint64 [2][160] globaltimer ns, end times then start times, indexed by block
size_t rnb_sdf_bwd_debug_offset(int64_t n_pts) {
    const SdfBwdScratch L = sdf_bwd_scratch(n_pts);
    const size_t nt = (size_t)(rnb_padded_points(n_pts) / TILE_M);
    return (L.queue + (64 + 18 * nt) * 4 + 7) & ~(size_t)7;
}

int rnb_sdf_bwd(const rnb_points_t* pts, const void* wblob, const float* aux, const float* d_sdf, const float* d_grad,
                const float* d_feat, const void* d_feat16, const float* d_feat16_meta, const void* st_in0,
                const void* st_in, const void* st_w, void* scratch, float* const* dW, float* const* db, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = pts->n_pts;
    if (n <= 0) return 0;
    const SdfBwdScratch L = sdf_bwd_scratch(n);
    uint8_t* sc = (uint8_t*)scratch;
    float* absmax = (float*)(sc + L.absmax);
    const size_t SS = rnb_stream_bytes(n, 256);
    if (d_feat && d_feat16) return (int)cudaErrorInvalidValue;
    if (d_feat16 && !d_feat16_meta) return (int)cudaErrorInvalidValue;
    cudaError_t e = profiled(T_ABSMAX, st, [&] {
        return launch_absmax(d_sdf, n, d_grad, 3 * n, d_feat, d_feat ? 256 * n : 0, d_feat16 ? d_feat16_meta : nullptr, absmax, st); });
    if (e != cudaSuccess) return (int)e;
    // ---- K3a: cotangent streams
    SdfBwdParams P{};
    P.src = to_src(pts);
    P.n_tiles = n_tiles(n);
    P.wblob = (const uint8_t*)wblob;
    P.aux = aux;
    // L2 prefetch of the streams each epilogue reads.  A whole step of lead -- the producer-side bulk prefetch (RNB_BWD_PF)
    // or per-thread hints issued one step ahead -- SLOWS this kernel down (6.5 -> 6.7..7.1 ms at 1 M points): ~25 KB/point
    // turn the L2 over so fast that early fills are evicted before they are read.  Hints only RNB_BWD_TPF (= 3) chunks
    // ahead of the consuming load cover the DRAM latency with a few MB of L2 footprint: 6.5 -> 5.4 ms.
    {
        static const int bwd_pf = getenv("RNB_BWD_PF") ? atoi(getenv("RNB_BWD_PF")) : 0;
        static const int bwd_tpf = getenv("RNB_BWD_TPF") ? atoi(getenv("RNB_BWD_TPF")) : 3;
        P.thread_prefetch = bwd_tpf;
#ifdef RNB_TRACE
        P.tab.trace = g_trace;
#endif
        // 25 KB/point of streams flow through L2 in this kernel: pin the 2 MB of weights (5.47 -> 5.24 ms); the other
        // chain kernels measured 3-5 % slower with the hinted copy, so they keep the plain one
        P.tab.weights_evict_last = 1;
        const uint8_t* a = (const uint8_t*)st_in;
        const uint8_t* w = (const uint8_t*)st_w;
        const uint8_t* u = sc + L.uin;
        // phase A, step l: epilogue reads a_l
        add_step(P.tab, sdfw_fwd(0), 256, 64, 0, bwd_pf ? a : nullptr);
        for (int l = 1; l < 8; ++l) add_step(P.tab, sdfw_fwd(l), 256, 256, 0, bwd_pf ? a + (size_t)l * SS : nullptr);
        // phase B, GEMM yielding abar_l: epilogue reads a_l, w_l, uin_{l+1}
        add_step(P.tab, SDFW_T8, 256, 256, 0, bwd_pf ? a + (size_t)7 * SS : nullptr, bwd_pf ? w + (size_t)7 * SS : nullptr,
                 bwd_pf > 1 ? u + (size_t)7 * SS : nullptr);
        for (int l = 7; l >= 1; --l)
            add_step(P.tab, sdfw_tr(l), 256, 256, 0, bwd_pf ? a + (size_t)(l - 1) * SS : nullptr,
                     bwd_pf ? w + (size_t)(l - 1) * SS : nullptr, bwd_pf > 1 ? u + (size_t)(l - 1) * SS : nullptr);
    }
    P.d_sdf = d_sdf; P.d_grad = d_grad; P.d_feat = d_feat; P.cot_absmax = absmax;
    P.d_feat16 = (const uint8_t*)d_feat16; P.d_feat16_cot_absmax = d_feat16_meta;
    P.st_in = (const uint8_t*)st_in; P.st_w = (const uint8_t*)st_w;
    P.st_uin0 = sc + L.uin0; P.st_uin = sc + L.uin; P.st_zbar = sc + L.zbar; P.st_dfeat = sc + L.dfeat;
    P.stream_stride = SS;
    // RNB_BWD_FUSED=1: one launch for K3a + K3b (sdf_bwd_fused_kernel, the weight-gradient workers take the cotangent
    // streams over through L2).  Measured (profiles/r02_notes.md): parity-green but not faster -- with ~230 tiles in flight the
    // hand-over set exceeds what L2 retains, DRAM reads drop only 35.6 -> 32 GB and the chain loses the workers' SMs -- so the
    // two-kernel path stays the default.
    const bool fused = getenv("RNB_BWD_FUSED") && atoi(getenv("RNB_BWD_FUSED")) != 0;
    if (!fused) {
        e = profiled(T_SDF_BWD_DATA, st, [&] { return launch_sdf_bwd_data(P, sm_count(), st); });
        if (e != cudaSuccess) return (int)e;
    }
    // ---- K3b: dW_l = w_l^T uin_l + zbar_l^T in_l  (l = 0..7),  dW_8[1:] = dfeat^T in_8
    const int n_sub = (int)(rnb_padded_points(n) / 64);
    DwParams D{};
    D.n_sub = n_sub;
    ReduceParams R{};
    R.cot_absmax = absmax;
    float* part = (float*)(sc + L.dw_part);
    float* dwcs = (float*)(sc + L.dwcs_part);
    int rep[9];
    fused_replicas(rep, L.dw_splits);
    const uint8_t* in0 = (const uint8_t*)st_in0;
    const uint8_t* inl = (const uint8_t*)st_in;
    const uint8_t* sw = (const uint8_t*)st_w;
    float* cpart = (float*)(sc + L.cs_part);      // [2][2 dw_splits][256]: the two sums of the sdf row of W_8
    const int cs_mul = 2;                         // column-sum partials per split / worker: one per group of four warps
    const size_t cs_stride = (size_t)2 * L.dw_splits * 256;
    // layer 8 first: its CTAs take three column sums per tile and must not form the tail of the launch
    for (int li = 0; li < 9; ++li) {
        const int l = li == 0 ? 8 : li - 1;
        DwJob& j = D.jobs[fused ? l : D.n_jobs];
        ++D.n_jobs;
        j.nw = l == 0 ? 64 : 256;
        j.b_chunk0 = 0;
        const int splits = fused ? rep[l] : L.dw_splits;
        if (l < 8) {
            j.n_pairs = j.mma_pairs = 2;
            j.a[0] = sw + (size_t)l * SS;
            j.b[0] = l == 0 ? sc + L.uin0 : sc + L.uin + (size_t)(l - 1) * SS;
            j.a[1] = sc + L.zbar + (size_t)l * SS;
            j.b[1] = l == 0 ? in0 : inl + (size_t)(l - 1) * SS;
        } else {
            j.n_pairs = j.mma_pairs = 1;          // dW_8[1:,:] = dfeat^T a_7
            j.a[0] = sc + L.dfeat;
            j.b[0] = inl + (size_t)7 * SS;
        }
        for (int k = 0; k < DW_MAX_PAIRS; ++k) j.b_chunks[k] = j.nw / 8;
        j.partial = part;
        // bias gradient db_l = column sums of zbar_l (dfeat for l = 8), taken from the staged A tile inside the GEMM
        {
            DwColsum& c = j.cs[j.n_cs++];
            c.pair = l < 8 ? 1 : 0; c.tile_off = 0; c.width = 256; c.n_w = 0; c.partial[0] = dwcs;
            ReduceJob& rb = R.jobs[R.n_jobs++];
            rb.partial = dwcs; rb.splits = cs_mul * splits; rb.rows = 1; rb.nw = 256;
            rb.dst = db[l]; rb.dst_pitch = 0; rb.dst_row0 = 0; rb.dst_col0 = l == 8 ? 1 : 0;
            rb.out_rows = 1; rb.out_cols = l == 3 ? 217 : 256; rb.factor = 1.f; rb.use_cot_scale = 1;
            dwcs += cs_stride;
        }
        // dW_8[0,:] = sum_p uabar_7[p,:]  +  sum_p d_sdf[p] a_7[p,:]   (models/fields.py:104: sdf = row 0 of lin8)
        float* p0 = cpart;                  // [2 dw_splits][256]  first sum, taken by layer 0's CTAs
        float* p1 = cpart + cs_stride;      //                     second sum (+ the weight sum), taken by layer 8's CTAs
        if (l == 0) {
            // uabar_7 (= the uin_8 stream) is no GEMM operand anywhere: its two 128-column halves ride in the 24 KB that
            // layer 0's 64-wide B operands leave free in the two stages of every sub-tile
            for (int pr = 0; pr < 2; ++pr) {
                j.x[pr] = sc + L.uin + (size_t)7 * SS;
                j.x_chunk0[pr] = 16 * pr;
                DwColsum& c = j.cs[j.n_cs++];
                c.pair = pr; c.tile_off = 32768 + 8 * 1024; c.width = 128; c.col0 = 128 * pr; c.n_w = 0; c.partial[0] = p0;
            }
        }
        if (l == 8) {
            DwColsum& c1 = j.cs[j.n_cs++];
            c1.pair = 0; c1.tile_off = 32768; c1.width = 256; c1.n_w = 1; c1.w[0] = d_sdf; c1.n_valid = n; c1.partial[0] = p1;
            // db_8[0] = sum_p d_sdf[p]: the sum of the weights of the same pass
            c1.wsum_partial = cpart + 2 * cs_stride;
            ReduceJob& rs = R.jobs[R.n_jobs++];
            rs.partial = c1.wsum_partial; rs.splits = cs_mul * splits; rs.rows = 1; rs.nw = 4;
            rs.dst = db[8]; rs.dst_pitch = 0; rs.out_rows = 1; rs.out_cols = 1; rs.factor = 1.f; rs.use_cot_scale = 0;
            ReduceJob& r = R.jobs[R.n_jobs++];
            r.partial = p0; r.splits = cs_mul * (fused ? rep[0] : L.dw_splits); r.rows = 1; r.nw = 256;
            r.dst = dW[8]; r.dst_pitch = 0; r.out_rows = 1; r.out_cols = 256; r.factor = 1.f; r.use_cot_scale = 1;
            r.partial2 = p1; r.splits2 = cs_mul * splits; r.factor2 = 1.f; r.use_cot_scale2 = 0;
        }
        ReduceJob& r = R.jobs[R.n_jobs++];
        r.partial = part; r.splits = splits; r.rows = 256; r.nw = j.nw;
        r.dst = dW[l]; r.dst_pitch = l == 0 ? 39 : 256; r.dst_row0 = l == 8 ? 1 : 0; r.dst_col0 = 0;
        r.out_rows = l == 3 ? 217 : 256; r.out_cols = l == 0 ? 39 : 256;
        r.factor = l == 4 ? 0.70710678118654752f : 1.f;
        r.use_cot_scale = 1; r.fold_xlo = l == 0; r.accumulate = 0;
        part += (size_t)L.dw_splits * 256 * j.nw;
    }
    if (fused) {
        SdfBwdFusedParams F{};
        F.chain = P;
        F.chain.keep_streams = getenv("RNB_FUSED_KEEP") ? atoi(getenv("RNB_FUSED_KEEP")) : 1;
        int* qmem = (int*)(sc + L.queue);
        // [head 16 | tail 16 | pad 32 | cnt 9 n_tiles | q 9 n_tiles | dbg]
        F.q_head = qmem; F.q_tail = qmem + 16; F.q_cnt = qmem + 64; F.q = F.q_cnt + (size_t)9 * P.n_tiles;
        e = cudaMemsetAsync(qmem, 0, (64 + (size_t)9 * P.n_tiles) * sizeof(int), st);
        if (e != cudaSuccess) return (int)e;
        e = cudaMemsetAsync(F.q, 0xff, (size_t)9 * P.n_tiles * sizeof(int), st);
        if (e != cudaSuccess) return (int)e;
        F.stagger_ns = getenv("RNB_FUSED_STAGGER_US") ? 1000 * atoi(getenv("RNB_FUSED_STAGGER_US")) : 150000;
        if (P.n_tiles < 4 * sm_count()) F.stagger_ns = 0;           // a couple of tiles per CTA: nothing to de-phase
        if (getenv("RNB_FUSED_DBG")) F.dbg = (unsigned long long*)(((uintptr_t)(F.q + (size_t)9 * P.n_tiles) + 7) & ~(uintptr_t)7);
        for (int l = 0; l < 9; ++l) F.jobs[l] = D.jobs[l];      // (fused: D.jobs is indexed by layer)
        for (int l = 0; l < 9; ++l)
            for (int r = 0; r < rep[l]; ++r) {
                F.dw_layer[F.n_dw] = (uint8_t)l;
                F.dw_replica[F.n_dw] = (uint8_t)r;
                ++F.n_dw;
            }
        e = profiled(T_SDF_BWD_FUSED, st, [&] { return launch_sdf_bwd_fused(F, sm_count(), st); });
    } else {
        e = profiled(T_DW_GEMM, st, [&] { return launch_dw_gemm(D, L.dw_splits, st); });
    }
    if (e != cudaSuccess) return (int)e;
    return (int)profiled(T_REDUCE, st, [&] { return launch_reduce(R, st); });
}


int rnb_coarse_z(const float* near, const float* far, const float* t_rand, float* z, int n_rays, int n_samples, void* stream) {
    return (int)profiled(T_COARSE_Z, (cudaStream_t)stream, [&] { return launch_coarse_z(near, far, t_rand, z, n_rays, n_samples, (cudaStream_t)stream); });
}
int rnb_upsample_step(const rnb_upsample_t* p, void* stream) {
    if (p->n_old + p->n_merge > MAX_RAY_SAMPLES || p->n_new > 32) return (int)cudaErrorInvalidValue;
    return (int)profiled(T_UPSAMPLE, (cudaStream_t)stream, [&] { return launch_upsample(*p, (cudaStream_t)stream); });
}
int rnb_sample_pdf_from_cdf(const float* bins, const float* cdf, int n_rays, int n, int n_new, float* samples, int64_t* inds,
                            void* stream) {
    return (int)profiled(T_SAMPLE_PDF, (cudaStream_t)stream, [&] { return launch_sample_pdf_from_cdf(bins, cdf, n_rays, n, n_new, samples, inds, (cudaStream_t)stream); });
}
int rnb_final_merge(const float* z_old, int n_old, const float* z_new, int n_new, int n_rays, float sample_dist, float* z_out,
                    float* mid_out, void* stream) {
    if (n_old + n_new > MAX_RAY_SAMPLES) return (int)cudaErrorInvalidValue;
    return (int)profiled(T_FINAL_MERGE, (cudaStream_t)stream, [&] { return launch_final_merge(z_old, n_old, z_new, n_new, n_rays, sample_dist, z_out, mid_out, (cudaStream_t)stream); });
}
int rnb_composite_fwd(const rnb_composite_t* p, void* stream) { return (int)profiled(T_COMPOSITE_FWD, (cudaStream_t)stream, [&] { return launch_composite(*p, false, (cudaStream_t)stream); }); }
int rnb_composite_bwd(const rnb_composite_t* p, void* stream) { return (int)profiled(T_COMPOSITE_BWD, (cudaStream_t)stream, [&] { return launch_composite(*p, true, (cudaStream_t)stream); }); }


int rnb_mc_count(const float* u, int nx, int ny, int nz, float threshold, const int8_t* tri_count, int32_t* counts, void* stream) {
    return (int)profiled(T_MC, (cudaStream_t)stream, [&] { return launch_mc_count(u, nx, ny, nz, threshold, tri_count, counts, (cudaStream_t)stream); });
}
int rnb_mc_emit(const float* u, int nx, int ny, int nz, float threshold, const int8_t* tri_table, const int64_t* offsets,
                int x_global0, float* verts, int64_t* keys, void* stream) {
    return (int)profiled(T_MC, (cudaStream_t)stream, [&] {
        return launch_mc_emit(u, nx, ny, nz, threshold, tri_table, offsets, x_global0, verts, keys, (cudaStream_t)stream); });
}
int rnb_weight_norm_fold(const rnb_wn_layer_t* layers, int n_layers, void* stream) {
    return (int)profiled(T_WNORM, (cudaStream_t)stream, [&] { return launch_wn_fold(layers, n_layers, (cudaStream_t)stream); });
}
int rnb_weight_norm_vjp(const rnb_wn_layer_t* layers, int n_layers, void* stream) {
    return (int)profiled(T_WNORM, (cudaStream_t)stream, [&] { return launch_wn_vjp(layers, n_layers, (cudaStream_t)stream); });
}
int rnb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, int64_t step, double grad_scale, void* stream) {
    return (int)profiled(T_ADAM, (cudaStream_t)stream, [&] {
        return launch_adam_flat(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, grad_scale, sm_count(),
                                (cudaStream_t)stream); });
}
int rnb_ray_batch(const rnb_ray_batch_t* p, void* stream) {
    if (!p->intrinsics_inv || !p->pose || !p->pixels_x || !p->pixels_y || !p->rays_o || !p->rays_d || !p->near || !p->far)
        return (int)cudaErrorInvalidValue;
    if ((p->rgb && !p->images) || (p->rgb2 && !p->images2) || (p->lights && !p->light_dirs) || (p->mask_out && !p->mask))
        return (int)cudaErrorInvalidValue;
    return (int)profiled(T_RAY_BATCH, (cudaStream_t)stream, [&] { return launch_ray_batch(*p, (cudaStream_t)stream); });
}
int rnb_stream_from_rowmajor(const float* x, int64_t n, int cols, void* out, void* stream) {
    if (cols % 8 != 0 || n < 0) return (int)cudaErrorInvalidValue;
    return (int)launch_stream_from_rowmajor(x, n, cols, rnb_padded_points(n), (uint8_t*)out, (cudaStream_t)stream);
}
size_t rnb_nerf_wblob_bytes(void) { return NRFW_BYTES; }
size_t rnb_nerf_aux_floats(void) { return NRFX_FLOATS; }
int rnb_nerf_pack(const float* const* W, const float* const* b, const float* Wf, const float* bf, const float* Wa,
                  const float* ba, const float* Wv, const float* bv, const float* Wr, const float* br, void* wblob,
                  float* aux, void* stream) {
    return (int)profiled(T_NERF_PACK, (cudaStream_t)stream, [&] {
        return launch_nerf_pack(W, b, Wf, bf, Wa, ba, Wv, bv, Wr, br, (uint8_t*)wblob, aux, (cudaStream_t)stream); });
}
int rnb_nerf_fwd(const rnb_points_t* pts, const float* pts4, const float* dirs, const void* wblob, const float* aux,
                 float* density, float* rgb, void* stream) {
    NerfFwdParams P{};
    P.n_pts = pts->n_pts;
    P.n_tiles = n_tiles(pts->n_pts);
    P.pts4 = pts4; P.dirs = dirs;
    P.rays_o = pts->rays_o; P.rays_d = pts->rays_d; P.z = pts->z; P.n_per_ray = pts->n_per_ray;
    if (!pts4 && !(pts->rays_o && pts->rays_d && pts->z && pts->n_per_ray > 0)) return (int)cudaErrorInvalidValue;
    if (pts4 && !dirs) return (int)cudaErrorInvalidValue;
    P.wblob = (const uint8_t*)wblob;
    P.aux = aux;
    add_step(P.tab, NRFW_L0, 256, NERF_PE_COLS);
    for (int l = 1; l <= 4; ++l) add_step(P.tab, NRFW_L1 + (uint32_t)(l - 1) * NRFW_MAT, 256, 256);
    add_step(P.tab, NRFW_L5H, 256, 256);
    add_step(P.tab, NRFW_L5E, 256, NERF_PE_COLS, 1);
    add_step(P.tab, NRFW_L6, 256, 256);
    add_step(P.tab, NRFW_L6 + NRFW_MAT, 256, 256);
    add_step(P.tab, NRFW_FEAT, 256, 256);
    add_step(P.tab, NRFW_VF, 128, 256);
    add_step(P.tab, NRFW_VE, 128, NERF_PEV_COLS, 1);
    P.density = density; P.rgb = rgb;
    return (int)profiled(T_NERF_FWD, (cudaStream_t)stream, [&] { return launch_nerf_fwd(P, sm_count(), (cudaStream_t)stream); });
}
int rnb_composite_bg_fwd(const rnb_composite_bg_t* p, void* stream) {
    return (int)profiled(T_COMPOSITE_BG, (cudaStream_t)stream, [&] { return launch_composite_bg(*p, (cudaStream_t)stream); });
}

size_t rnb_albedo_wblob_bytes(void) { return ALBW_BYTES; }
size_t rnb_albedo_aux_floats(void) { return ALBX_FLOATS; }
int rnb_albedo_pack(const float* W0, const float* b0, const float* W1, const float* b1, const float* W2, const float* b2,
                    void* wblob, float* aux, void* stream) {
    return (int)profiled(T_ALBEDO_PACK, (cudaStream_t)stream, [&] { return launch_albedo_pack(W0, b0, W1, b1, W2, b2, (uint8_t*)wblob, aux, (cudaStream_t)stream); });
}

int rnb_albedo_fwd(const rnb_points_t* pts, const float* normals, const void* st_feat, const void* wblob, const float* aux,
                   float* albedo, void* st_pe, void* st_h0, void* st_h1, void* stream) {
    AlbedoFwdParams P{};
    P.src = to_src(pts);
    P.n_tiles = n_tiles(pts->n_pts);
    P.wblob = (const uint8_t*)wblob;
    P.aux = aux;
#ifdef RNB_TRACE
    P.tab.trace = g_trace;
#endif
    add_step(P.tab, ALBW_F0A, 256, 256);
    add_step(P.tab, ALBW_F0B, 256, 64, 1);
    add_step(P.tab, ALBW_F1, 256, 256);
    P.normals = normals; P.st_feat = (const uint8_t*)st_feat; P.albedo = albedo;
    P.st_pe = (uint8_t*)st_pe; P.st_h0 = (uint8_t*)st_h0; P.st_h1 = (uint8_t*)st_h1;
    return (int)profiled(T_ALBEDO_FWD, (cudaStream_t)stream, [&] { return launch_albedo_fwd(P, sm_count(), (cudaStream_t)stream); });
}

size_t rnb_albedo_bwd_scratch_bytes(int64_t n_pts) { return albedo_bwd_scratch(n_pts).total; }

int rnb_albedo_bwd(const rnb_points_t* pts, const float* normals, const float* albedo, const float* d_albedo,
                   const void* st_feat, const void* st_pe, const void* st_h0, const void* st_h1, const void* wblob,
                   const float* aux, void* scratch, float* d_normal, float* d_feat, void* d_feat16, float* d_feat16_meta,
                   float* dW0, float* db0, float* dW1, float* db1, float* dW2, float* db2, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = pts->n_pts;
    if (n <= 0) return 0;
    const AlbedoBwdScratch L = albedo_bwd_scratch(n);
    uint8_t* sc = (uint8_t*)scratch;
    // meta[0] = max |d_albedo| (the cotangent scale derives from it), meta[1] = max |stored fp16 d_feat|
    if (d_feat16 && !d_feat16_meta) return (int)cudaErrorInvalidValue;
    float* absmax = d_feat16_meta ? d_feat16_meta : (float*)(sc + L.absmax);
    const int64_t n_pad = rnb_padded_points(n);
    cudaError_t e = cudaMemsetAsync(absmax, 0, 2 * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    e = profiled(T_ABSMAX, st, [&] { return launch_absmax(d_albedo, 3 * n, nullptr, 0, nullptr, 0, nullptr, absmax, st); });
    if (e != cudaSuccess) return (int)e;
    AlbedoBwdParams P{};
    P.src = to_src(pts);
    P.n_tiles = n_tiles(n);
    P.wblob = (const uint8_t*)wblob;
    P.aux = aux;
#ifdef RNB_TRACE
    P.tab.trace = g_trace;
#endif
    add_step(P.tab, ALBW_T1, 256, 256);
    add_step(P.tab, ALBW_T0A, 256, 256);
    add_step(P.tab, ALBW_T0B, 64, 256);
    P.normals = normals; P.albedo = albedo; P.d_albedo = d_albedo; P.cot_absmax = absmax;
    P.st_h0 = (const uint8_t*)st_h0; P.st_h1 = (const uint8_t*)st_h1;
    P.n_pad = n_pad;
    P.st_dz2 = sc + L.dz2;
    P.st_dz1 = sc + L.dz1; P.st_dz0 = sc + L.dz0;
    P.d_feat = d_feat; P.d_normal = d_normal;
    P.st_dfeat16 = (uint8_t*)d_feat16; P.dfeat_max = absmax + 1;
    e = profiled(T_ALBEDO_BWD, st, [&] { return launch_albedo_bwd(P, sm_count(), st); });
    if (e != cudaSuccess) return (int)e;
    const int n_sub = (int)(n_pad / 64);
    DwParams D{};
    D.n_sub = n_sub;
    ReduceParams R{};
    R.cot_absmax = absmax;
    float* part = (float*)(sc + L.dw_part);
    float* dwcs = (float*)(sc + L.dwcs_part);
    auto add_dw = [&](const uint8_t* a, const uint8_t* b, int b_chunks, int nw, float* dst, int pitch, int col0, int out_cols,
                      float* db_dst) {
        DwJob& j = D.jobs[D.n_jobs++];
        j.n_pairs = j.mma_pairs = 1; j.a[0] = a; j.b[0] = b; j.b_chunk0 = 0; j.nw = nw; j.partial = part;
        for (int k = 0; k < DW_MAX_PAIRS; ++k) j.b_chunks[k] = b_chunks;
        if (db_dst) {
            DwColsum& c = j.cs[j.n_cs++];
            c.pair = 0; c.tile_off = 0; c.width = 256; c.n_w = 0; c.partial[0] = dwcs;
            ReduceJob& rb = R.jobs[R.n_jobs++];
            rb.partial = dwcs; rb.splits = 2 * L.dw_splits; rb.rows = 1; rb.nw = 256;
            rb.dst = db_dst; rb.dst_pitch = 0; rb.out_rows = 1; rb.out_cols = 256; rb.factor = 1.f; rb.use_cot_scale = 1;
            dwcs += (size_t)2 * L.dw_splits * 256;
        }
        ReduceJob& r = R.jobs[R.n_jobs++];
        r.partial = part; r.splits = L.dw_splits; r.rows = 256; r.nw = nw;
        r.dst = dst; r.dst_pitch = pitch; r.dst_row0 = 0; r.dst_col0 = col0; r.out_rows = 256; r.out_cols = out_cols;
        r.factor = 1.f; r.use_cot_scale = 1;
        part += (size_t)L.dw_splits * 256 * nw;
    };
    add_dw(sc + L.dz1, (const uint8_t*)st_h0, 32, 256, dW1, 256, 0, 256, db1);
    add_dw(sc + L.dz0, (const uint8_t*)st_feat, 32, 256, dW0, 310, 54, 256, db0);
    add_dw(sc + L.dz0, (const uint8_t*)st_pe, 8, 64, dW0, 310, 0, 54, nullptr);
    {
        // dW_2[k,:] = sum_p dz2[k,p] h_1[p,:] and db_2[k] = sum_p dz2[k,p], k = 0..2 (the 3-wide output layer,
        // models/fields.py:203-214): one more job of the same GEMM, D[256 x 16] = h_1^T dz2 with the 16-wide dz2 stream as
        // its B operand, written transposed into dW_2 [3,256]; db_2 = the column sums of the staged dz2 tile
        DwJob& j = D.jobs[D.n_jobs++];
        j.n_pairs = j.mma_pairs = 1; j.a[0] = (const uint8_t*)st_h1; j.b[0] = sc + L.dz2; j.b_chunk0 = 0; j.nw = 16; j.partial = part;
        for (int k = 0; k < DW_MAX_PAIRS; ++k) j.b_chunks[k] = 2;
        float* cpart = (float*)(sc + L.cs_part);
        DwColsum& c = j.cs[j.n_cs++];
        c.pair = 0; c.tile_off = 32768; c.width = 16; c.n_w = 0; c.partial[0] = cpart;
        ReduceJob& rb = R.jobs[R.n_jobs++];
        rb.partial = cpart; rb.splits = 2 * L.dw_splits; rb.rows = 1; rb.nw = 256;
        rb.dst = db2; rb.dst_pitch = 0; rb.out_rows = 1; rb.out_cols = 3; rb.factor = 1.f; rb.use_cot_scale = 1;
        ReduceJob& r = R.jobs[R.n_jobs++];
        r.partial = part; r.splits = L.dw_splits; r.rows = 256; r.nw = 16;
        r.dst = dW2; r.dst_pitch = 1; r.dst_row0 = 0; r.dst_col0 = 0; r.dst_col_stride = 256; r.out_rows = 256; r.out_cols = 3;
        r.factor = 1.f; r.use_cot_scale = 1;
        part += (size_t)L.dw_splits * 256 * 16;
    }
    e = profiled(T_DW_GEMM, st, [&] { return launch_dw_gemm(D, L.dw_splits, st); });
    if (e != cudaSuccess) return (int)e;
    return (int)profiled(T_REDUCE, st, [&] { return launch_reduce(R, st); });
}


void rnb_profile_enable(int on) { g_prof_on.store(on != 0); }
long long rnb_launch_count(void) {
    long long n = 0;
    for (int t = 0; t < T_COUNT; ++t) n += g_launches[t].load(std::memory_order_relaxed);
    return n;
}
int rnb_profile_collect(char* names, int name_stride, float* total_ms, int* counts, int max_tags) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    float ms[T_COUNT] = {0};
    int cnt[T_COUNT] = {0};
    for (const ProfRec& r : g_prof) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.tag] += t; ++cnt[r.tag]; }
        g_ev_pool.push_back(r.a);
        g_ev_pool.push_back(r.b);
    }
    g_prof.clear();
    int n = 0;
    for (int t = 0; t < T_COUNT && n < max_tags; ++t) {
        if (!cnt[t]) continue;
        std::strncpy(names + (size_t)n * name_stride, kProfNames[t], name_stride - 1);
        names[(size_t)n * name_stride + name_stride - 1] = 0;
        total_ms[n] = ms[t];
        counts[n] = cnt[t];
        ++n;
    }
    return n;
}

}  // extern "C"
