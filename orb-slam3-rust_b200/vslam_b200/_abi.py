"""ctypes declarations of include/vslam_b200.h (the C ABI of libvslam_b200.so).

The library is the product: importing this module with the shared object missing raises immediately — there is
no CPU or PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VSLAM_B200_LIB") or os.path.join(_HERE, "libvslam_b200.so")  # env: A/B builds of the same ABI

ABI_VERSION = 1
OK, ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_CAPACITY, ERR_KEYPOINT_OVERFLOW, ERR_OUT_OF_MEMORY = range(6)
STAGE_PYRAMID, STAGE_BLURRED = 0, 1
SEARCH_LOCAL_MAP, SEARCH_MOTION_MODEL = 0, 1
MODE_REFERENCE, MODE_CLASSIC = 0, 1
GRID_COLS, GRID_ROWS = 64, 48

KP_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"),
     ("class_id", "<i4")]
)
DMATCH_DTYPE = np.dtype([("query_idx", "<i4"), ("train_idx", "<i4"), ("img_idx", "<i4"), ("distance", "<f4")])
assert KP_DTYPE.itemsize == 28 and DMATCH_DTYPE.itemsize == 16


class Camera(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("baseline", C.c_double)]


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("n_features", C.c_int32),
        ("n_levels", C.c_int32), ("scale_factor", C.c_float), ("edge_threshold", C.c_int32),
        ("fast_threshold", C.c_int32), ("camera", Camera), ("device", C.c_int32), ("max_batch", C.c_int32),
        ("keypoint_capacity", C.c_int32), ("mode", C.c_int32), ("classic_ini_fast", C.c_int32),
        ("classic_min_fast", C.c_int32), ("reserved", C.c_int32 * 5),
    ]


class FrameOut(C.Structure):
    _fields_ = [
        ("capacity", C.c_int32), ("kp_left", C.c_void_p), ("kp_right", C.c_void_p), ("desc_left", C.c_void_p),
        ("desc_right", C.c_void_p), ("matches", C.c_void_p), ("points_cam", C.c_void_p), ("points_valid", C.c_void_p),
        ("n_left", C.c_int32), ("n_right", C.c_int32), ("n_matches", C.c_int32), ("status", C.c_int32),
        ("timestamp_ns", C.c_uint64),
    ]


class SearchParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("radius", C.c_double), ("th_high", C.c_uint32), ("nn_ratio", C.c_float),
                ("img_w", C.c_double), ("img_h", C.c_double), ("bound_w", C.c_double), ("bound_h", C.c_double)]


class TriParams(C.Structure):
    _fields_ = [("F", C.c_double * 9), ("epipole", C.c_double * 2), ("max_dist", C.c_uint32), ("radius", C.c_float),
                ("image_width", C.c_uint32), ("image_height", C.c_uint32)]


# every symbol include/vslam_b200.h declares (tests/test_abi.py checks the .so exports all of them)
SYMBOLS = [
    "vslam_create", "vslam_destroy", "vslam_last_error", "vslam_keypoint_capacity", "vslam_process",
    "vslam_process_batch", "vslam_process_batch_device", "vslam_detect_and_compute", "vslam_match_features",
    "vslam_triangulate", "vslam_descriptor_distance", "vslam_grid_build", "vslam_grid_query",
    "vslam_search_by_projection", "vslam_bf_match_crosscheck", "vslam_level_info", "vslam_debug_level_image",
    "vslam_debug_fast_corners", "vslam_kernel_launches", "vslam_set_profiling", "vslam_last_stage_ms",
    "vslam_last_batch_ms", "vslam_fuse_search", "vslam_distinctive_descriptors", "vslam_vocab_upload",
    "vslam_bow_quantize", "vslam_search_for_triangulation", "vslam_search_for_triangulation_bow",
    "vslam_frame_retain", "vslam_frame_upload", "vslam_frame_release", "vslam_frame_num_features",
    "vslam_frame_search_by_projection", "vslam_frame_grid_query", "vslam_frame_bf_match_crosscheck",
    "vslam_last_search_ms", "vslam_classic_stereo_extra", "vslam_debug_classic_candidates",
]

_lib = None


def lib():
    """Loads libvslam_b200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). vslam_b200 has no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        vp, i32, u64, sz = C.c_void_p, C.c_int, C.c_uint64, C.c_size_t
        L.vslam_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.vslam_destroy.argtypes = [vp]
        L.vslam_destroy.restype = None
        L.vslam_last_error.argtypes = [vp]
        L.vslam_last_error.restype = C.c_char_p
        L.vslam_keypoint_capacity.argtypes = [vp]
        L.vslam_process.argtypes = [vp, vp, sz, vp, sz, u64, C.POINTER(FrameOut)]
        L.vslam_process_batch.argtypes = [vp, i32, vp, vp, sz, vp, C.POINTER(FrameOut)]
        L.vslam_process_batch_device.argtypes = [vp, i32, vp, vp, sz, sz, vp, C.POINTER(FrameOut)]
        L.vslam_detect_and_compute.argtypes = [vp, vp, sz, vp, vp, i32, C.POINTER(i32)]
        L.vslam_match_features.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, C.POINTER(i32)]
        L.vslam_triangulate.argtypes = [vp, vp, i32, vp, i32, vp, i32, vp, vp]
        L.vslam_descriptor_distance.argtypes = [vp, vp, vp, i32, vp]
        L.vslam_grid_build.argtypes = [vp, vp, i32, C.c_double, C.c_double, vp, vp]
        L.vslam_grid_query.argtypes = [vp, vp, i32, C.c_double, C.c_double, vp, i32, vp, vp, i32]
        L.vslam_search_by_projection.argtypes = [vp, vp, vp, i32, vp, vp, i32, C.POINTER(SearchParams), vp, vp]
        L.vslam_bf_match_crosscheck.argtypes = [vp, vp, i32, vp, i32, vp, C.POINTER(i32)]
        L.vslam_fuse_search.argtypes = [vp, vp, vp, i32, vp, vp, i32, C.c_uint32, vp, vp]
        L.vslam_distinctive_descriptors.argtypes = [vp, vp, vp, i32, vp]
        L.vslam_vocab_upload.argtypes = [vp, i32, vp, vp, vp, vp, vp]
        L.vslam_bow_quantize.argtypes = [vp, vp, i32, i32, vp, vp, vp]
        L.vslam_search_for_triangulation.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, i32, C.POINTER(TriParams), vp,
                                                     C.POINTER(i32)]
        L.vslam_search_for_triangulation_bow.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, i32, C.POINTER(TriParams), i32,
                                                         vp, vp, vp, vp, vp, C.POINTER(i32)]
        L.vslam_frame_retain.argtypes = [vp, i32, C.c_double, C.c_double, C.POINTER(i32)]
        L.vslam_frame_upload.argtypes = [vp, vp, vp, i32, C.c_double, C.c_double, C.POINTER(i32)]
        L.vslam_frame_release.argtypes = [vp, i32]
        L.vslam_frame_num_features.argtypes = [vp, i32]
        L.vslam_frame_search_by_projection.argtypes = [vp, i32, vp, vp, i32, C.POINTER(SearchParams), vp, vp]
        L.vslam_frame_grid_query.argtypes = [vp, i32, vp, i32, vp, vp, i32]
        L.vslam_frame_bf_match_crosscheck.argtypes = [vp, i32, i32, vp, C.POINTER(i32)]
        L.vslam_last_search_ms.argtypes = [vp]
        L.vslam_last_search_ms.restype = C.c_float
        L.vslam_classic_stereo_extra.argtypes = [vp, i32, vp, vp, i32]
        L.vslam_debug_classic_candidates.argtypes = [vp, i32, i32, vp, i32, C.POINTER(i32)]
        L.vslam_level_info.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_float), C.POINTER(i32)]
        L.vslam_debug_level_image.argtypes = [vp, i32, i32, i32, vp, sz]
        L.vslam_debug_fast_corners.argtypes = [vp, i32, i32, vp, i32, C.POINTER(i32)]
        L.vslam_kernel_launches.argtypes = [vp]
        L.vslam_kernel_launches.restype = u64
        L.vslam_set_profiling.argtypes = [vp, i32]
        L.vslam_last_stage_ms.argtypes = [vp, vp]
        L.vslam_last_batch_ms.argtypes = [vp]
        L.vslam_last_batch_ms.restype = C.c_float
        _lib = L
    return _lib
