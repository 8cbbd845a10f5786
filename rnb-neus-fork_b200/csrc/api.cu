// extern "C" boundary of librnb_b200.so (declarations and contracts: include/rnb_b200.h).
#include "../../include/rnb_b200.h"
#include "sdf_params.h"

namespace rnb {
cudaError_t launch_sdf_pack(const float* const* W, const float* const* b, uint8_t* blob, float* aux, cudaStream_t st);
cudaError_t launch_sdf_fwd(const SdfFwdParams& P, int sm_count, cudaStream_t st);
cudaError_t launch_sdf_fwd_grad(const SdfFwdGradParams& P, int sm_count, cudaStream_t st);
cudaError_t launch_sdf_bwd_data(const SdfBwdParams& P, int sm_count, cudaStream_t st);

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static SdfPointSource to_src(const rnb_points_t* p) {
    SdfPointSource s;
    s.n_pts = p->n_pts; s.x = p->x; s.rays_o = p->rays_o; s.rays_d = p->rays_d; s.z = p->z;
    s.n_per_ray = p->n_per_ray; s.grid_res = p->grid_res; s.slab_x0 = p->slab_x0;
    for (int a = 0; a < 3; ++a) { s.bmin[a] = p->bmin[a]; s.bmax[a] = p->bmax[a]; }
    return s;
}

static void add_step(ChainTable& t, uint32_t off, int n, int k) {
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
}  // namespace rnb

using namespace rnb;

extern "C" {

const char* rnb_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }
int rnb_version(void) { return 100; }
size_t rnb_sdf_wblob_bytes(void) { return SDFW_BYTES; }
size_t rnb_sdf_aux_floats(void) { return AUX_FLOATS; }
int64_t rnb_padded_points(int64_t n) { return (n + TILE_M - 1) / TILE_M * TILE_M; }
size_t rnb_stream_bytes(int64_t n, int cols) { return (size_t)rnb_padded_points(n) * (size_t)cols * 2; }

int rnb_sdf_pack(const float* const* W, const float* const* b, void* wblob, float* aux, void* stream) {
    return (int)launch_sdf_pack(W, b, (uint8_t*)wblob, aux, (cudaStream_t)stream);
}

int rnb_sdf_fwd(const rnb_points_t* pts, const void* wblob, const float* aux, float* out, float out_scale, void* stream) {
    SdfFwdParams P{};
    P.src = to_src(pts);
    P.n_tiles = n_tiles(pts->n_pts);
    P.wblob = (const uint8_t*)wblob;
    P.aux = aux;
    table_forward(P.tab);
    P.out = out;
    P.out_scale = out_scale;
    return (int)launch_sdf_fwd(P, sm_count(), (cudaStream_t)stream);
}

int rnb_sdf_fwd_grad(const rnb_points_t* pts, const void* wblob, const float* aux, float* out_sdf, float* out_grad,
                     float* out_full, void* st_feat, void* st_in0, void* st_in, void* st_s, void* st_w, void* stream) {
    SdfFwdGradParams P{};
    P.src = to_src(pts);
    P.n_tiles = n_tiles(pts->n_pts);
    P.wblob = (const uint8_t*)wblob;
    P.aux = aux;
    table_forward(P.tab);
    add_step(P.tab, SDFW_F8, 256, 256);
    for (int l = 7; l >= 1; --l) add_step(P.tab, sdfw_tr(l), 256, 256);
    add_step(P.tab, sdfw_tr(0), 64, 256);
    P.out_sdf = out_sdf; P.out_grad = out_grad; P.out_full = out_full;
    P.st_feat = (uint8_t*)st_feat; P.st_in0 = (uint8_t*)st_in0; P.st_in = (uint8_t*)st_in;
    P.st_s = (uint8_t*)st_s; P.st_w = (uint8_t*)st_w;
    P.stream_stride = rnb_stream_bytes(pts->n_pts, 256);
    return (int)launch_sdf_fwd_grad(P, sm_count(), (cudaStream_t)stream);
}

}  // extern "C"
