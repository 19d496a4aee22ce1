"""ctypes view of the C ABI declared in include/rasterize_b200.h.

The shared library is built in-tree by rasterize_b200/build.py (nvcc, sm_100a).  Loading fails loudly when it
is missing: there is no CPU fallback in this package.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path as _P

PKG = _P(__file__).resolve().parent
LIB_PATH = PKG / "librasterize_b200.so"

OK, ERR_INVALID, ERR_CUDA, ERR_NAN, ERR_DEPTH, ERR_CAPACITY, ERR_WINDING = 0, -1, -2, -3, -4, -5, -6
JOB_MASK, JOB_COVERAGE, JOB_FILL, JOB_RENDER = 0, 1, 2, 3
OUT_LINCOLOR, OUT_RGBA8, OUT_COVERAGE = 0, 1, 2
BATCH_ORDERED, BATCH_INDEPENDENT = 0, 1
MAX_STOPS = 32


class RgpuError(RuntimeError):
    """Non-zero status from the C ABI (the Rust shim turns these into `panic!`)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"rasterize_b200 error {code}: {msg}")
        self.code = code


class CPath(C.Structure):
    _fields_ = [("points", C.POINTER(C.c_double)), ("kinds", C.POINTER(C.c_uint8)), ("subpath_offsets", C.POINTER(C.c_uint32)),
                ("closed", C.POINTER(C.c_uint8)), ("n_points", C.c_uint32), ("n_segments", C.c_uint32), ("n_subpaths", C.c_uint32)]


class CShape(C.Structure):
    _fields_ = [("start", C.c_size_t), ("width", C.c_size_t), ("height", C.c_size_t), ("row_stride", C.c_size_t),
                ("col_stride", C.c_size_t)]


class CPixel(C.Structure):
    _fields_ = [("x", C.c_size_t), ("y", C.c_size_t), ("alpha", C.c_double)]


class CPaint(C.Structure):
    _fields_ = [("kind", C.c_int32), ("units", C.c_int32), ("linear_colors", C.c_int32), ("spread", C.c_int32),
                ("tr", C.c_double * 6), ("p0", C.c_double * 2), ("p1", C.c_double * 2), ("r0", C.c_double), ("r1", C.c_double),
                ("solid", C.c_float * 4), ("n_stops", C.c_uint32), ("stop_pos", C.POINTER(C.c_double)),
                ("stop_colors", C.POINTER(C.c_float))]


class CJob(C.Structure):
    _fields_ = [("path", C.c_void_p), ("tr", C.c_double * 6), ("fill_rule", C.c_int32), ("mode", C.c_int32),
                ("paint", C.POINTER(CPaint)), ("path_bbox", C.POINTER(C.c_double)), ("canvas", C.c_void_p),
                ("origin", C.c_size_t), ("row_stride", C.c_size_t), ("width", C.c_uint32), ("height", C.c_uint32)]


class CStrokeStyle(C.Structure):
    _fields_ = [("width", C.c_double), ("miter_limit", C.c_double), ("line_join", C.c_int32), ("line_cap", C.c_int32)]


class CParseOptions(C.Structure):
    _fields_ = [("fit_width", C.c_uint32), ("fit_height", C.c_uint32), ("fit_align", C.c_int32)]


class CSceneFill(C.Structure):
    _fields_ = [("path", C.POINTER(CPath)), ("tr", C.c_double * 6), ("fill_rule", C.c_int32), ("paint", C.POINTER(CPaint)),
                ("path_bbox", C.POINTER(C.c_double)), ("x", C.c_uint32), ("y", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32)]


_lib = None

#: every symbol include/rasterize_b200.h declares (the ABI test walks this table)
SYMBOLS = [
    "rgpu_create", "rgpu_destroy", "rgpu_name", "rgpu_last_error", "rgpu_device_count", "rgpu_flatten", "rgpu_mask",
    "rgpu_mask_f32", "rgpu_mask_iter", "rgpu_coverage_f32", "rgpu_fill", "rgpu_path_upload", "rgpu_path_free",
    "rgpu_render_batch", "rgpu_batch_status", "rgpu_render_batch_sync", "rgpu_render_scene", "rgpu_render_scene_sync", "rgpu_render_scene_host", "rgpu_last_counts", "rgpu_last_transfer_bytes", "rgpu_set_profiling",
    "rgpu_last_stage_ms", "rgpu_to_rgba8_dev", "rgpu_layer_scale_by_mask_dev", "rgpu_layer_blend_over_dev", "rgpu_download_rgba8",
    "rgpu_fill_color_dev", "rgpu_stream", "rgpu_sync", "rgpu_device_alloc", "rgpu_device_free", "rgpu_device_zero",
    "rgpu_memcpy_h2d", "rgpu_memcpy_d2h", "rgpu_host_alloc", "rgpu_host_free",
    "rgpu_path_upload_batch", "rgpu_path_batch_get", "rgpu_path_batch_free", "rgpu_batch_create", "rgpu_batch_render", "rgpu_batch_free",
    "rgpu_fill_batch_host", "rgpu_mask_banded_host", "rgpu_multi_create", "rgpu_multi_destroy", "rgpu_multi_device_count",
    "rgpu_multi_last_error", "rgpu_multi_fill_batch_host", "rgpu_multi_mask_banded_host", "rgpu_set_winding_bits",
    "rgpu_path_stroke", "rgpu_dpath_stroke", "rgpu_dpath_info", "rgpu_dpath_download", "rgpu_parse_svg_batch", "rgpu_path_batch_info", "rgpu_path_batch_download",
]


def lib():
    """Load librasterize_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc required; rasterize_b200 has no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    vp, sz, dbl, i32, u32 = C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_uint32
    pd, pf = C.POINTER(C.c_double), C.POINTER(C.c_float)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("rgpu_create", i32, i32, dbl, C.POINTER(vp))
    sig("rgpu_destroy", None, vp)
    sig("rgpu_name", C.c_char_p)
    sig("rgpu_last_error", C.c_char_p, vp)
    sig("rgpu_device_count", i32)
    sig("rgpu_flatten", i32, vp, C.POINTER(CPath), pd, i32, pd, sz, C.POINTER(sz))
    sig("rgpu_mask", i32, vp, C.POINTER(CPath), pd, i32, vp, CShape)
    sig("rgpu_mask_f32", i32, vp, C.POINTER(CPath), pd, i32, vp, sz, sz)
    sig("rgpu_mask_iter", i32, vp, C.POINTER(CPath), pd, sz, sz, i32, C.POINTER(CPixel), sz, C.POINTER(sz))
    sig("rgpu_coverage_f32", i32, vp, C.POINTER(CPath), pd, i32, vp, sz, sz)
    sig("rgpu_fill", i32, vp, C.POINTER(CPath), pd, i32, C.POINTER(CPaint), pd, vp, CShape)
    sig("rgpu_path_upload", i32, vp, C.POINTER(CPath), C.POINTER(vp))
    sig("rgpu_path_free", None, vp, vp)
    sig("rgpu_path_stroke", i32, vp, C.POINTER(CPath), C.POINTER(CStrokeStyle), C.POINTER(vp))
    sig("rgpu_dpath_stroke", i32, vp, vp, C.POINTER(CStrokeStyle), C.POINTER(vp))
    pu32_ = C.POINTER(C.c_uint32)
    sig("rgpu_dpath_info", i32, vp, pu32_, pu32_, pu32_)
    sig("rgpu_dpath_download", i32, vp, vp, C.POINTER(C.c_double), C.POINTER(C.c_uint8), pu32_, C.POINTER(C.c_uint8))
    sig("rgpu_parse_svg_batch", i32, vp, C.c_char_p, pu32_, sz, C.POINTER(CParseOptions), C.POINTER(vp), vp)
    sig("rgpu_path_batch_info", i32, vp, C.POINTER(sz), pu32_, pu32_, pu32_)
    sig("rgpu_path_batch_download", i32, vp, vp, C.POINTER(C.c_double), C.POINTER(C.c_uint8), pu32_, C.POINTER(C.c_uint8), pu32_)
    sig("rgpu_render_batch", i32, vp, C.POINTER(CJob), sz, u32)
    sig("rgpu_batch_status", i32, vp)
    sig("rgpu_set_winding_bits", i32, vp, i32)
    sig("rgpu_render_batch_sync", i32, vp, C.POINTER(CJob), sz, u32)
    sig("rgpu_render_scene", i32, vp, C.POINTER(CJob), sz, vp, sz, sz, i32, pf, vp)
    sig("rgpu_render_scene_sync", i32, vp, C.POINTER(CJob), sz, vp, sz, sz, i32, pf, vp)
    sig("rgpu_render_scene_host", i32, vp, C.POINTER(CSceneFill), sz, sz, sz, pf, vp, vp)
    sig("rgpu_last_counts", i32, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64))
    sig("rgpu_last_transfer_bytes", i32, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64))
    sig("rgpu_set_profiling", i32, vp, i32)
    sig("rgpu_last_stage_ms", i32, vp, pf)
    sig("rgpu_to_rgba8_dev", i32, vp, vp, vp, sz)
    sig("rgpu_fill_color_dev", i32, vp, vp, sz, pf)
    sig("rgpu_layer_scale_by_mask_dev", i32, vp, vp, sz, sz, vp, sz, sz, sz, sz)
    sig("rgpu_layer_blend_over_dev", i32, vp, vp, sz, sz, vp, sz, sz, sz, sz, i32, C.c_float)
    sig("rgpu_download_rgba8", i32, vp, vp, sz, vp)
    sig("rgpu_stream", vp, vp)
    sig("rgpu_sync", i32, vp)
    sig("rgpu_device_alloc", i32, vp, sz, C.POINTER(vp))
    sig("rgpu_device_free", i32, vp, vp)
    sig("rgpu_device_zero", i32, vp, vp, sz)
    sig("rgpu_memcpy_h2d", i32, vp, vp, vp, sz)
    sig("rgpu_memcpy_d2h", i32, vp, vp, vp, sz)
    sig("rgpu_host_alloc", i32, vp, sz, C.POINTER(vp))
    sig("rgpu_host_free", i32, vp, vp)
    pu32 = C.POINTER(C.c_uint32)
    sig("rgpu_path_upload_batch", i32, vp, C.POINTER(CPath), pu32, sz, C.POINTER(vp))
    sig("rgpu_path_batch_get", vp, vp, sz)
    sig("rgpu_path_batch_free", None, vp, vp)
    sig("rgpu_batch_create", i32, vp, vp, sz, u32, C.POINTER(vp))
    sig("rgpu_batch_render", i32, vp, vp)
    sig("rgpu_batch_free", None, vp, vp)
    sig("rgpu_fill_batch_host", i32, vp, C.POINTER(CPath), pu32, sz, pd, i32, C.POINTER(CPaint), u32, u32, i32, vp)
    sig("rgpu_mask_banded_host", i32, vp, C.POINTER(CPath), pd, i32, vp, sz, sz, sz, u32, u32, u32)
    sig("rgpu_multi_create", i32, C.POINTER(C.c_int), i32, dbl, C.POINTER(vp))
    sig("rgpu_multi_destroy", None, vp)
    sig("rgpu_multi_device_count", i32, vp)
    sig("rgpu_multi_last_error", C.c_char_p, vp)
    sig("rgpu_multi_fill_batch_host", i32, vp, C.POINTER(CPath), pu32, sz, pd, i32, C.POINTER(CPaint), u32, u32, i32, vp)
    sig("rgpu_multi_mask_banded_host", i32, vp, C.POINTER(CPath), pd, i32, vp, sz, sz, sz, u32)
    _lib = L
    return L
