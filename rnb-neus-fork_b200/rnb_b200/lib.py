"""ctypes binding of librnb_b200.so (C ABI in include/rnb_b200.h).

The library is the product: there is no Python/CPU fallback.  Importing this module without the built
shared object, or calling a kernel without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RNB_B200_LIB") or os.path.join(_HERE, "librnb_b200.so")   # override: A/B builds of the same ABI


class Points(C.Structure):
    """rnb_points_t"""
    _fields_ = [("n_pts", C.c_int64), ("x", C.c_void_p), ("rays_o", C.c_void_p), ("rays_d", C.c_void_p),
                ("z", C.c_void_p), ("n_per_ray", C.c_int32), ("grid_res", C.c_int32), ("slab_x0", C.c_int32),
                ("bmin", C.c_float * 3), ("bmax", C.c_float * 3)]


class Upsample(C.Structure):
    """rnb_upsample_t"""
    _fields_ = [("n_rays", C.c_int32), ("rays_o", C.c_void_p), ("rays_d", C.c_void_p), ("z_old", C.c_void_p),
                ("sdf_old", C.c_void_p), ("n_old", C.c_int32), ("z_pending", C.c_void_p), ("sdf_pending", C.c_void_p),
                ("n_merge", C.c_int32), ("z_merged", C.c_void_p), ("sdf_merged", C.c_void_p), ("inv_s", C.c_float),
                ("n_new", C.c_int32), ("z_new", C.c_void_p), ("inds", C.c_void_p), ("cdf_out", C.c_void_p)]


class Composite(C.Structure):
    """rnb_composite_t"""
    _fields_ = [("n_rays", C.c_int32), ("rays_o", C.c_void_p), ("rays_d", C.c_void_p), ("z", C.c_void_p),
                ("sdf", C.c_void_p), ("grad", C.c_void_p), ("albedo", C.c_void_p), ("lights", C.c_void_p),
                ("n_lights", C.c_int32), ("light_stride_l", C.c_int64), ("light_stride_ray", C.c_int64),
                ("variance", C.c_void_p), ("cos_anneal_ratio", C.c_float), ("warmup", C.c_int32),
                ("sample_dist", C.c_float), ("color", C.c_void_p), ("weights", C.c_void_p), ("cdf", C.c_void_p),
                ("inside", C.c_void_p), ("weight_sum", C.c_void_p), ("weight_max", C.c_void_p),
                ("eik_part", C.c_void_p), ("d_color", C.c_void_p), ("d_weight_sum", C.c_void_p), ("d_eik", C.c_void_p),
                ("eik_den", C.c_void_p), ("d_sdf", C.c_void_p), ("d_grad", C.c_void_p), ("d_albedo", C.c_void_p),
                ("d_var_part", C.c_void_p)]


class CompositeBg(C.Structure):
    """rnb_composite_bg_t"""
    _fields_ = [("n_rays", C.c_int32), ("rays_o", C.c_void_p), ("rays_d", C.c_void_p), ("z", C.c_void_p),
                ("sdf", C.c_void_p), ("grad", C.c_void_p), ("color_in", C.c_void_p), ("variance", C.c_void_p),
                ("cos_anneal_ratio", C.c_float), ("sample_dist", C.c_float), ("z_feed", C.c_void_p),
                ("bg_density", C.c_void_p), ("bg_rgb", C.c_void_p), ("n_outside", C.c_int32), ("color", C.c_void_p),
                ("weights", C.c_void_p), ("cdf", C.c_void_p), ("inside", C.c_void_p), ("weight_sum", C.c_void_p),
                ("weight_max", C.c_void_p), ("eik_part", C.c_void_p)]


class RayBatch(C.Structure):
    """rnb_ray_batch_t"""
    _fields_ = [("n_rays", C.c_int32), ("n_lights", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("intrinsics_inv", C.c_void_p), ("pose", C.c_void_p), ("pixels_x", C.c_void_p), ("pixels_y", C.c_void_p),
                ("images", C.c_void_p), ("images2", C.c_void_p), ("mask", C.c_void_p), ("mask_channels", C.c_int32),
                ("light_dirs", C.c_void_p), ("rays_o", C.c_void_p), ("rays_d", C.c_void_p), ("near", C.c_void_p),
                ("far", C.c_void_p), ("mask_out", C.c_void_p), ("rgb", C.c_void_p), ("rgb2", C.c_void_p),
                ("lights", C.c_void_p)]


class WnLayer(C.Structure):
    """rnb_wn_layer_t"""
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("v", C.c_void_p), ("g", C.c_void_p), ("w", C.c_void_p),
                ("norm", C.c_void_p), ("dv", C.c_void_p), ("dg", C.c_void_p)]


_lib = None

_VP = C.c_void_p
_SIGNATURES = {
    "rnb_error_string": (C.c_char_p, [C.c_int]),
    "rnb_version": (C.c_int, []),
    "rnb_sdf_wblob_bytes": (C.c_size_t, []),
    "rnb_sdf_aux_floats": (C.c_size_t, []),
    "rnb_padded_points": (C.c_int64, [C.c_int64]),
    "rnb_stream_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "rnb_sdf_pack": (C.c_int, [C.POINTER(_VP), C.POINTER(_VP), _VP, _VP, _VP]),
    "rnb_sdf_fwd": (C.c_int, [C.POINTER(Points), _VP, _VP, _VP, C.c_float, _VP]),
    "rnb_sdf_fwd_grad": (C.c_int, [C.POINTER(Points)] + [_VP] * 10),
    "rnb_sdf_bwd_scratch_bytes": (C.c_size_t, [C.c_int64]),
    "rnb_sdf_bwd_debug_offset": (C.c_size_t, [C.c_int64]),
    "rnb_albedo_wblob_bytes": (C.c_size_t, []),
    "rnb_albedo_aux_floats": (C.c_size_t, []),
    "rnb_albedo_pack": (C.c_int, [_VP] * 9),
    "rnb_albedo_fwd": (C.c_int, [C.POINTER(Points)] + [_VP] * 9),
    "rnb_albedo_bwd_scratch_bytes": (C.c_size_t, [C.c_int64]),
    "rnb_albedo_bwd": (C.c_int, [C.POINTER(Points)] + [_VP] * 21),
    "rnb_launch_count": (C.c_longlong, []),
    "rnb_profile_enable": (None, [C.c_int]),
    "rnb_profile_collect": (C.c_int, [C.c_char_p, C.c_int, _VP, _VP, C.c_int]),
    "rnb_coarse_z": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int, _VP]),
    "rnb_upsample_step": (C.c_int, [C.POINTER(Upsample), _VP]),
    "rnb_sample_pdf_from_cdf": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP]),
    "rnb_final_merge": (C.c_int, [_VP, C.c_int, _VP, C.c_int, C.c_int, C.c_float, _VP, _VP, _VP]),
    "rnb_composite_fwd": (C.c_int, [C.POINTER(Composite), _VP]),
    "rnb_composite_bwd": (C.c_int, [C.POINTER(Composite), _VP]),
    "rnb_nerf_wblob_bytes": (C.c_size_t, []),
    "rnb_nerf_aux_floats": (C.c_size_t, []),
    "rnb_nerf_pack": (C.c_int, [C.POINTER(_VP), C.POINTER(_VP)] + [_VP] * 11),
    "rnb_nerf_fwd": (C.c_int, [C.POINTER(Points)] + [_VP] * 7),
    "rnb_composite_bg_fwd": (C.c_int, [C.POINTER(CompositeBg), _VP]),
    "rnb_mc_count": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_float, _VP, _VP, _VP]),
    "rnb_mc_emit": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_float, _VP, _VP, C.c_int, _VP, _VP, _VP]),
    "rnb_weight_norm_fold": (C.c_int, [C.POINTER(WnLayer), C.c_int, _VP]),
    "rnb_weight_norm_vjp": (C.c_int, [C.POINTER(WnLayer), C.c_int, _VP]),
    "rnb_adam_step": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                C.c_int64, C.c_double, _VP]),
    "rnb_ray_batch": (C.c_int, [C.POINTER(RayBatch), _VP]),
    "rnb_stream_from_rowmajor": (C.c_int, [_VP, C.c_int64, C.c_int, _VP, _VP]),
    "rnb_sdf_bwd": (C.c_int, [C.POINTER(Points)] + [_VP] * 11 + [C.POINTER(_VP), C.POINTER(_VP), _VP]),
}


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(f"rnb_b200: {LIB_PATH} is missing -- build it with __graft_entry__.build() "
                               "(rnb-neus-fork_b200/csrc/build.sh); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def profile_enable(on: bool):
    load().rnb_profile_enable(1 if on else 0)


def profile_collect():
    """-> {kernel name: (total device ms, launches)} since the last collect (synchronises the device)."""
    lib = load()
    max_tags, stride = 32, 32
    names = C.create_string_buffer(max_tags * stride)
    ms = (C.c_float * max_tags)()
    cnt = (C.c_int * max_tags)()
    n = lib.rnb_profile_collect(names, stride, C.cast(ms, C.c_void_p), C.cast(cnt, C.c_void_p), max_tags)
    out = {}
    for i in range(n):
        out[names.raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode()] = (float(ms[i]), int(cnt[i]))
    return out


def launch_count() -> int:
    return int(load().rnb_launch_count())


def check(code: int, what: str):
    if code != 0:
        raise RuntimeError(f"rnb_b200.{what} failed: {load().rnb_error_string(code).decode()} ({code})")


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "rnb_b200 kernels take contiguous CUDA tensors"
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device.  torch.cuda.current_stream() builds a Python Stream
    object (~18 us; 18 calls per 512-ray step were 0.3 ms of a host-bound 1.8 ms iteration): the raw query is ~100x cheaper."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"rnb_b200.{what}: the hot path runs only on CUDA (sm_100a); got a {t.device} tensor. "
                           "There is no CPU fallback -- move the module and inputs to a B200.")
