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
 * activation streams the backward needs.  out_full (optional, may be NULL): fp32 [n,257] like the reference. */
RNB_API int rnb_sdf_fwd_grad(const rnb_points_t* pts, const void* wblob, const float* aux, float* out_sdf, float* out_grad,
                     float* out_full, void* st_feat, void* st_in0, void* st_in, void* st_s, void* st_w, void* stream);

/* Double-backward of (sdf, features, gradient) w.r.t. the effective weights (what loss.backward() does to
 * SDFNetwork.forward + .gradient in the reference: exp_runner.py:261 through models/fields.py:82-127).
 * Cotangents: d_sdf [n], d_grad [n,3], d_feat [n,256] fp32 row-major (d_feat may be NULL = zero).
 * st_*: the streams written by rnb_sdf_fwd_grad for the same points.  scratch: rnb_sdf_bwd_scratch_bytes(n).
 * dW, db: HOST arrays of 9 DEVICE pointers, dW[l] fp32 [out_l,in_l] row-major (256x39, 256x256, 256x256,
 * 217x256, 256x256 x4, 257x256), db[l] [out_l]; overwritten. */
RNB_API size_t rnb_sdf_bwd_scratch_bytes(int64_t n_pts);
RNB_API int rnb_sdf_bwd(const rnb_points_t* pts, const void* wblob, const float* aux, const float* d_sdf, const float* d_grad,
                const float* d_feat, const void* st_in0, const void* st_in, const void* st_s, const void* st_w,
                void* scratch, float* const* dW, float* const* db, void* stream);

#ifdef __cplusplus
}
#endif
#endif
