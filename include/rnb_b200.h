/* rnb_b200.h -- C ABI of the B200-native RNb-NeuS hot path (librnb_b200.so).
 *
 * The reference (rti-team-imvia/RNb-NeuS-fork) has no FFI: its seam is the Python class API of models/
 * (SURVEY.md 8b).  Each entry point below replaces the ATen op sequence of one reference function and is what
 * a maintainer would bind (ctypes stub in INTEGRATION.md).  Conventions: plain pointers and sizes, all pointers
 * are DEVICE pointers unless stated, `stream` is a cudaStream_t passed as void*, the caller owns every buffer,
 * return value 0 = success, otherwise a cudaError_t code (rnb_error_string() describes it).  Kernels hold no
 * global state besides cached function attributes; they are re-entrant per stream.
 */
#ifndef RNB_B200_H
#define RNB_B200_H
#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define RNB_API __attribute__((visibility("default")))
#else
#define RNB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Where the points of an SDF launch come from (exactly one of the three sources is used):
 *   grid_res > 0 : lattice of extract_fields (reference models/renderer.py:10-25), x index offset by slab_x0
 *   rays_o != 0  : o + d*z ray samples (reference models/renderer.py:863, 181, 487)
 *   otherwise    : explicit points x[n_pts,3] (SDFNetwork.forward/sdf/gradient, reference models/fields.py:82-127) */
typedef struct {
    int64_t n_pts;
    const float* x;
    const float* rays_o;
    const float* rays_d;
    const float* z;
    int32_t n_per_ray;
    int32_t grid_res;
    int32_t slab_x0;
    float bmin[3];
    float bmax[3];
} rnb_points_t;

RNB_API const char* rnb_error_string(int code);
RNB_API int rnb_version(void);

/* sizes of caller-allocated buffers */
RNB_API size_t rnb_sdf_wblob_bytes(void);            /* packed fp16 weight images of the SDF net          */
RNB_API size_t rnb_sdf_aux_floats(void);             /* fp32 side table (biases, W_8[0,:], b_8[0])         */
RNB_API int64_t rnb_padded_points(int64_t n_pts);    /* n_pts rounded up to the 128-point tile             */
RNB_API size_t rnb_stream_bytes(int64_t n_pts, int cols);   /* one fp16 activation stream [Npad x cols]    */

/* weight-norm-folded fp32 weights W_l [out,in] / b_l of lin0..lin8 -> packed operands.
 * W, b: HOST arrays of 9 DEVICE pointers.  Replaces nothing in the reference (it re-folds per call,
 * models/fields.py:72-74); run once per optimiser step. */
RNB_API int rnb_sdf_pack(const float* const* W, const float* const* b, void* wblob, float* aux, void* stream);

/* SDFNetwork.sdf under no_grad (reference models/fields.py:106-108; call sites models/renderer.py:864, 186,
 * 1224): out[p] = out_scale * sdf(point p).  out_scale = -1 gives extract_fields' u = -sdf. */
RNB_API int rnb_sdf_fwd(const rnb_points_t* pts, const void* wblob, const float* aux, float* out, float out_scale, void* stream);

/* SDFNetwork.forward + SDFNetwork.gradient fused (reference models/fields.py:82-127; call site
 * models/renderer.py:492-498).  Writes sdf [n], grad [n,3], the feature stream (fp16 [Npad x 256]) and the
 * activation streams the backward needs: st_in0 [Npad x 64] (layer-0 input), st_in = 8 streams a_l [Npad x 256],
 * st_w = 8 streams w_l = softplus'(z_l) * (dx-chain cotangent).  out_full (optional, may be NULL): fp32 [n,257] like the reference.
 * Inference (no backward to follow, e.g. validate_image / render_novel_image): st_w and st_in0 may be NULL -- they are
 * only read by rnb_sdf_bwd -- which saves 4.1 of the 8.8 KB/point this kernel writes; st_in is still needed (the
 * kernel's own dx-chain reads it back) and st_feat feeds rnb_albedo_fwd. */
RNB_API int rnb_sdf_fwd_grad(const rnb_points_t* pts, const void* wblob, const float* aux, float* out_sdf, float* out_grad,
                     float* out_full, void* st_feat, void* st_in0, void* st_in, void* st_w, void* stream);

/* Double-backward of (sdf, features, gradient) w.r.t. the effective weights (what loss.backward() does to
 * SDFNetwork.forward + .gradient in the reference: exp_runner.py:261 through models/fields.py:82-127).
 * Cotangents: d_sdf [n], d_grad [n,3], and the feature cotangent either as d_feat [n,256] fp32 row-major or as
 * d_feat16 = the fp16 stream + d_feat16_meta = the device float[2] written by rnb_albedo_bwd (both NULL = zero).
 * st_*: the streams written by rnb_sdf_fwd_grad for the same points.  scratch: rnb_sdf_bwd_scratch_bytes(n).
 * dW, db: HOST arrays of 9 DEVICE pointers, dW[l] fp32 [out_l,in_l] row-major (256x39, 256x256, 256x256,
 * 217x256, 256x256 x4, 257x256), db[l] [out_l]; overwritten.
 * Launches: absmax, the backward chain (sdf_bwd_data), the weight-gradient GEMM (dw_gemm: all 9 layers, every bias /
 * sdf-row column sum taken from its staged tiles), one deterministic reduce.  Environment RNB_BWD_FUSED=1 runs the chain and
 * the weight-gradient contraction as ONE launch instead (sdf_bwd_fused; same results to summation order, measured not
 * faster: profiles/r02_notes.md); RNB_DW_REPL="3,4,4,4,4,4,4,4,2" sets its workers per layer. */
RNB_API size_t rnb_sdf_bwd_scratch_bytes(int64_t n_pts);
/* Diagnostics of the fused backward launch (env RNB_FUSED_DBG=1): byte offset inside `scratch` of uint64 [2][160]
 * globaltimer stamps (block end times, then block start times), indexed by block: blocks [0, n_workers) are the
 * weight-gradient workers, the rest the chain blocks. */
RNB_API size_t rnb_sdf_bwd_debug_offset(int64_t n_pts);
RNB_API int rnb_sdf_bwd(const rnb_points_t* pts, const void* wblob, const float* aux, const float* d_sdf, const float* d_grad,
                const float* d_feat, const void* d_feat16, const float* d_feat16_meta, const void* st_in0,
                const void* st_in, const void* st_w, void* scratch, float* const* dW, float* const* db, void* stream);


/* ---- per-ray kernels (reference models/renderer.py) ------------------------------------------------------- */

/* one hierarchical up-sampling step: merge the samples drawn by the previous step (cat_z_vals, renderer.py:178-192),
 * then up_sample (:132-176) + sample_pdf(det=True) (:39-69) -> n_new new depths per ray. */
typedef struct {
    int32_t n_rays;
    const float* rays_o;        /* [B,3] */
    const float* rays_d;        /* [B,3] */
    const float* z_old;         /* [B,n_old] sorted */
    const float* sdf_old;       /* [B,n_old] */
    int32_t n_old;
    const float* z_pending;     /* [B,n_merge] drawn by the previous step (merged first); may be NULL */
    const float* sdf_pending;   /* [B,n_merge] */
    int32_t n_merge;
    float* z_merged;            /* [B,n_old+n_merge] out; may be NULL */
    float* sdf_merged;
    float inv_s;                /* 64 * 2^i (renderer.py:872) */
    int32_t n_new;              /* <= 32 */
    float* z_new;               /* [B,n_new] out */
    int32_t* inds;              /* optional [B,n_new]: searchsorted indices */
    float* cdf_out;             /* optional [B,n_old+n_merge] */
} rnb_upsample_t;

/* render_core_mvps after the networks + RNb shading sum (renderer.py:503-540, 904-918, 1008-1017) and its adjoint.
 * 128 samples per ray.  Light l of ray b is read at lights + l*light_stride_l + b*light_stride_ray (floats). */
typedef struct {
    int32_t n_rays;
    const float* rays_o;
    const float* rays_d;
    const float* z;             /* [B,128] */
    const float* sdf;           /* [B*128] at the section mid-points */
    const float* grad;          /* [B*128,3] */
    const float* albedo;        /* [B*128,3] or NULL (no_albedo: ones) */
    const float* lights;
    int32_t n_lights;
    int64_t light_stride_l, light_stride_ray;
    const float* variance;      /* device scalar, SingleVarianceNetwork.variance */
    float cos_anneal_ratio;
    int32_t warmup;             /* 1 = relu on the shading (render_rnb_warmup) */
    float sample_dist;          /* 2 / n_samples */
    float* color;               /* out [L,B,3] */
    float* weights;             /* out [B,128] */
    float* cdf;                 /* out [B,128] */
    float* inside;              /* out [B,128] */
    float* weight_sum;          /* out [B] */
    float* weight_max;          /* out [B] */
    float* eik_part;            /* out [B,2] per-ray numerator / denominator of the eikonal term */
    const float* d_color;       /* bwd in [L,B,3] */
    const float* d_weight_sum;  /* bwd in [B] or NULL */
    const float* d_eik;         /* bwd in, device scalar */
    const float* eik_den;       /* bwd in, device scalar: sum of relax_inside_sphere */
    float* d_sdf;               /* bwd out [B*128] */
    float* d_grad;              /* bwd out [B*128,3] */
    float* d_albedo;            /* bwd out [B*128,3] or NULL */
    float* d_var_part;          /* bwd out [B] per-ray d loss / d variance */
} rnb_composite_t;

/* z = near + (far-near)*linspace(0,1,n) + t_rand*2/n   (renderer.py:829-845); t_rand may be NULL */
RNB_API int rnb_coarse_z(const float* near, const float* far, const float* t_rand, float* z, int n_rays, int n_samples, void* stream);
RNB_API int rnb_upsample_step(const rnb_upsample_t* p, void* stream);
/* test hook for "searchsorted indices bit-exact given the same CDF": sample_pdf's inverse-CDF on a caller CDF */
RNB_API int rnb_sample_pdf_from_cdf(const float* bins, const float* cdf, int n_rays, int n, int n_new, float* samples,
                            int64_t* inds, void* stream);
/* last cat_z_vals (z only) + section mid-points z + dist/2 (renderer.py:479-484) */
RNB_API int rnb_final_merge(const float* z_old, int n_old, const float* z_new, int n_new, int n_rays, float sample_dist,
                    float* z_out, float* mid_out, void* stream);
RNB_API int rnb_composite_fwd(const rnb_composite_t* p, void* stream);
RNB_API int rnb_composite_bwd(const rnb_composite_t* p, void* stream);

/* ---- ray batches on the device (reference models/dataset.py:351-376 ps_gen_random_rays_at_view_on_all_lights,
 *      :448-458 near_far_from_sphere, and the light-direction gather of exp_runner.py:214-220).  The reference indexes
 *      three CPU-resident [views,L,H,W,3] tensors with CPU pixel indices and copies five tensors to the GPU every step;
 *      here the images stay in HBM and one kernel gathers everything for the given pixels. --------------------------- */
typedef struct {
    int32_t n_rays, n_lights, H, W;
    const float* intrinsics_inv;  /* [4,4] row-major, the view's K^-1 (rows/cols 0..2 used) */
    const float* pose;            /* [4,4] row-major camera-to-world */
    const int64_t* pixels_x;      /* [B] */
    const int64_t* pixels_y;      /* [B] */
    const float* images;          /* [L,H,W,3] of the view, or NULL */
    const float* images2;         /* a second image set with the same layout (warm-up / regular), or NULL */
    const float* mask;            /* [H,W,mask_channels], or NULL */
    int32_t mask_channels;
    const float* light_dirs;      /* [L,H,W,3] per-pixel light directions of the view, or NULL */
    float* rays_o;                /* out [B,3] */
    float* rays_d;                /* out [B,3] */
    float* near;                  /* out [B] */
    float* far;                   /* out [B] */
    float* mask_out;              /* out [B] or NULL */
    float* rgb;                   /* out [L,B,3] or NULL */
    float* rgb2;                  /* out [L,B,3] or NULL */
    float* lights;                /* out [L,B,3] or NULL */
} rnb_ray_batch_t;
RNB_API int rnb_ray_batch(const rnb_ray_batch_t* p, void* stream);

/* ---- weight norm of all layers of a network at once (reference models/fields.py:72-74, 161-162: nn.utils.weight_norm on
 *      every layer; forward hook W = g v / |v|_row per layer and call, backward by autograd).  rnb_weight_norm_fold:
 *      w[rows,cols] = g[r] * v[r,:] / |v[r,:]|, norm[r] = |v[r,:]| for every layer in ONE launch.  rnb_weight_norm_vjp:
 *      w holds the incoming dW; dg[r] = <dW,v>/norm, dv = g/norm * (dW - v <dW,v>/norm^2), ONE launch. -------------- */
#define RNB_WN_MAX_LAYERS 16
typedef struct {
    int32_t rows, cols;
    const float* v;      /* weight_v [rows,cols] */
    const float* g;      /* weight_g [rows] */
    float* w;            /* fold: out W [rows,cols];  vjp: in dW [rows,cols] */
    float* norm;         /* fold: out [rows];  vjp: in */
    float* dv;           /* vjp: out [rows,cols] */
    float* dg;           /* vjp: out [rows] */
} rnb_wn_layer_t;
RNB_API int rnb_weight_norm_fold(const rnb_wn_layer_t* layers, int n_layers, void* stream);
RNB_API int rnb_weight_norm_vjp(const rnb_wn_layer_t* layers, int n_layers, void* stream);

/* ---- step epilogue (reference exp_runner.py:115 torch.optim.Adam over 61 parameter tensors, :263
 *      optimizer.step()).  One launch over flat fp32 buffers of n elements (16-byte aligned): param, grad, exp_avg and
 *      exp_avg_sq, torch.optim.Adam arithmetic with amsgrad off and weight_decay 0; `step` is the 1-based step count
 *      of this update, grad is multiplied by grad_scale first (1/world after a summing all-reduce). -------------------- */
RNB_API int rnb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, int64_t step, double grad_scale, void* stream);

/* ---- marching cubes on the device lattice (reference models/renderer.py:28-36 calls PyMCubes on the
 *      host after copying the whole lattice).  u: fp32 [nx,ny,nz] C-contiguous, inside = u > threshold.  Two passes:
 *      rnb_mc_count -> counts[(nx-1)(ny-1)(nz-1)] triangles per cell; the caller scans them into offsets;
 *      rnb_mc_emit  -> verts [n_tri*3,3] in lattice index coordinates (x shifted by x_global0 for slabs) and
 *      keys [n_tri*3] identifying the cube edge of every vertex (equal key <=> same vertex: weld with a unique).
 *      tri_count [256] / tri_table [256*16] (int8, device) come from rnb_b200/mc_tables.py. ----------------------- */
RNB_API int rnb_mc_count(const float* u, int nx, int ny, int nz, float threshold, const int8_t* tri_count, int32_t* counts, void* stream);
RNB_API int rnb_mc_emit(const float* u, int nx, int ny, int nz, float threshold, const int8_t* tri_table, const int64_t* offsets,
                int x_global0, float* verts, int64_t* keys, void* stream);

/* fp32 [n, cols] row-major -> fp16 stream image (cols % 8 == 0; out: rnb_stream_bytes(n, cols)).  Used by the
 * stand-alone RenderingNetwork.forward (exp_runner.py:613-615 validate_mesh_texture), whose feature vectors arrive
 * as an fp32 tensor instead of the stream rnb_sdf_fwd_grad writes. */
RNB_API int rnb_stream_from_rowmajor(const float* x, int64_t n, int cols, void* out, void* stream);

/* ---- NeRF++ background (reference models/fields.py:219-314 NeRF; models/renderer.py:93-130 render_core_outside,
 *      :255-260 foreground/background blend inside render_core) -- inference only, like its one reference caller
 *      (render() <- render_novel_image, exp_runner.py:541) ---------------------------------------------------- */
RNB_API size_t rnb_nerf_wblob_bytes(void);
RNB_API size_t rnb_nerf_aux_floats(void);
/* W, b: HOST arrays of 8 DEVICE pointers (pts_linears.{0..7}: [256,84], [256,256] x4, [256,340], [256,256] x2);
 * Wf [256,256] feature_linear, Wa [1,256] alpha_linear, Wv [128,283] views_linears.0, Wr [3,128] rgb_linear. */
RNB_API int rnb_nerf_pack(const float* const* W, const float* const* b, const float* Wf, const float* bf, const float* Wa,
                  const float* ba, const float* Wv, const float* bv, const float* Wr, const float* br, void* wblob,
                  float* aux, void* stream);
/* density [n] (raw alpha_linear output), rgb [n,3] (raw rgb_linear output) = NeRF(pts4, dirs).
 * Either explicit inputs pts4 [n,4] + dirs [n,3] (NeRF.forward), or ray samples: pts->rays_o/rays_d/z with
 * point = o + d*z, pts4 = [p/r, 1/r], r = max(|p|, 1), dirs = d (render_core_outside, renderer.py:105-113). */
RNB_API int rnb_nerf_fwd(const rnb_points_t* pts, const float* pts4, const float* dirs, const void* wblob, const float* aux,
                 float* density, float* rgb, void* stream);

/* render_core with a background model (renderer.py:194-285 with background_alpha / background_sampled_color):
 * alpha and colour of the 128 SDF samples are blended with the NeRF++ samples by inside_sphere, the n_outside
 * background samples are appended, then weights / colour are composited over 128 + n_outside samples. */
typedef struct {
    int32_t n_rays;
    const float* rays_o;
    const float* rays_d;
    const float* z;             /* [B,128] depths of the SDF pass */
    const float* sdf;           /* [B*128] at the section mid-points */
    const float* grad;          /* [B*128,3] */
    const float* color_in;      /* [B*128,3] colour-network output */
    const float* variance;      /* device scalar */
    float cos_anneal_ratio;
    float sample_dist;
    const float* z_feed;        /* [B,128+n_outside] sorted depths of the background pass */
    const float* bg_density;    /* [B*(128+n_outside)] raw NeRF density at the z_feed section mid-points */
    const float* bg_rgb;        /* [B*(128+n_outside),3] raw NeRF rgb */
    int32_t n_outside;          /* 1..64 */
    float* color;               /* out [B,3] */
    float* weights;             /* out [B,128+n_outside] */
    float* cdf;                 /* out [B,128] */
    float* inside;              /* out [B,128] */
    float* weight_sum;          /* out [B] */
    float* weight_max;          /* out [B] */
    float* eik_part;            /* out [B,2] */
} rnb_composite_bg_t;
RNB_API int rnb_composite_bg_fwd(const rnb_composite_bg_t* p, void* stream);

/* ---- albedo network (reference RenderingNetwork mode 'no_view_dir', models/fields.py:131-215) ------------- */
RNB_API size_t rnb_albedo_wblob_bytes(void);
RNB_API size_t rnb_albedo_aux_floats(void);
/* effective fp32 weights: W0 [256,310], W1 [256,256], W2 [3,256] and biases -> packed operands */
RNB_API int rnb_albedo_pack(const float* W0, const float* b0, const float* W1, const float* b1, const float* W2, const float* b2,
                    void* wblob, float* aux, void* stream);
/* albedo[n,3] = sigmoid(MLP([PE4(points), PE4(normals), features])); the view directions of the reference call
 * (models/renderer.py:501) are embedded and then unused in mode 'no_view_dir' (models/fields.py:179-192), so they
 * are not an argument.  st_feat: feature stream of rnb_sdf_fwd_grad; st_pe [Npad x 64], st_h0/st_h1 [Npad x 256]
 * are written for the backward. */
RNB_API int rnb_albedo_fwd(const rnb_points_t* pts, const float* normals, const void* st_feat, const void* wblob, const float* aux,
                   float* albedo, void* st_pe, void* st_h0, void* st_h1, void* stream);
RNB_API size_t rnb_albedo_bwd_scratch_bytes(int64_t n_pts);
/* VJP: d_albedo [n,3] -> d_normal [n,3], the feature cotangent and the effective-weight gradients
 * dW0 [256,310], db0 [256], dW1 [256,256], db1 [256], dW2 [3,256], db2 [3] (overwritten).
 * The feature cotangent comes out as d_feat [n,256] fp32 row-major (may be NULL) and/or as d_feat16, an fp16 stream
 * [Npad x 256] holding d_feat times this call's power-of-two cotangent scale, with d_feat16_meta (device float[2]:
 * [0] = max |d_albedo| the scale derives from, [1] = max |stored value|); rnb_sdf_bwd consumes the pair directly, which
 * saves the 1 KB/point fp32 round trip through HBM. */
RNB_API int rnb_albedo_bwd(const rnb_points_t* pts, const float* normals, const float* albedo, const float* d_albedo,
                   const void* st_feat, const void* st_pe, const void* st_h0, const void* st_h1, const void* wblob,
                   const float* aux, void* scratch, float* d_normal, float* d_feat, void* d_feat16, float* d_feat16_meta,
                   float* dW0, float* db0, float* dW1, float* db1, float* dW2, float* db2, void* stream);

/* ---- instrumentation -------------------------------------------------------------------------------------- */
/* total number of kernels this library has launched in this process */
RNB_API long long rnb_launch_count(void);
/* when enabled, every kernel launch is bracketed by cudaEvents on its stream; rnb_profile_collect synchronises the
 * device and returns, per kernel name, the summed device time (ms) and the launch count since the last collect.
 * names: max_tags rows of name_stride chars.  Returns the number of rows written. */
RNB_API void rnb_profile_enable(int on);
RNB_API int rnb_profile_collect(char* names, int name_stride, float* total_ms, int* counts, int max_tags);

#ifdef __cplusplus
}
#endif
#endif
