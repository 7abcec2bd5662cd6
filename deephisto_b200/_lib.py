"""ctypes binding of libdeephisto_b200.so (C-ABI declared in include/deephisto_b200.h).

There is no fallback: if the shared library is missing the import of any compute entry point
raises, and on a box without an sm_100 GPU `require_device()` raises.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("DEEPHISTO_B200_LIB", _HERE / "libdeephisto_b200.so"))

DH_F32, DH_BF16, DH_U8 = 0, 1, 2
DH_NHWC, DH_NCHW, DH_S2D16, DH_S2D48 = 0, 1, 2, 3
DH_FLIP_H, DH_FLIP_V = 1, 2
DH_SLOT_OK, DH_SLOT_MISS_LIMIT, DH_SLOT_EMPTY_RANGE = 0, 1, 2


class DeepHistoError(RuntimeError):
    pass


class RegionTables(C.Structure):
    """struct dh_region_tables (include/deephisto_b200.h)."""

    _fields_ = [
        ("edges", C.c_void_p),
        ("edge_off", C.c_void_p),
        ("reg_bbox", C.c_void_p),
        ("reg_area", C.c_void_p),
        ("reg_image", C.c_void_p),
        ("img_hw", C.c_void_p),
        ("tbl_cls_off", C.c_void_p),
        ("tbl_cls", C.c_void_p),
        ("cat_off", C.c_void_p),
        ("cat_region", C.c_void_p),
        ("cat_cdf", C.c_void_p),
        ("img_cdf", C.c_void_p),
        ("n_tables", C.c_int32),
        ("n_classes", C.c_int32),
        ("n_regions", C.c_int32),
        ("n_images", C.c_int32),
    ]


_i64, _i32, _u64, _vp, _f64 = C.c_int64, C.c_int, C.c_uint64, C.c_void_p, C.c_double

# name -> (restype, argtypes); must list every DH_API symbol of the header (tests check this)
SIGNATURES = {
    "dh_version": (C.c_int, []),
    "dh_last_error": (C.c_char_p, []),
    "dh_device_check": (C.c_int, []),
    "dh_synth_slide": (C.c_int, [_vp, _i64, _i64, _i64, _u64, _vp]),
    "dh_synth_slide_rows": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _i64, _u64, _vp]),
    "dh_dense_count": (_i64, [_i64, _i64, _i32, _i32, _i32, C.POINTER(_i64)]),
    "dh_dense_coords": (C.c_int, [_i64, _i64, _i32, _i32, _i32, _i64, _i64, _vp, _vp]),
    "dh_gather_normalize": (
        C.c_int,
        [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _i32, _vp, _i32, _i32, _i32, C.POINTER(C.c_float), C.POINTER(C.c_float), _vp, _vp],
    ),
    "dh_gather_normalize_multi": (
        C.c_int,
        [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _vp, _i32, _i32, _i32, C.POINTER(C.c_float), C.POINTER(C.c_float), _vp, _vp],
    ),
    "dh_gather_set_variant": (C.c_int, [_i32]),
    "dh_stitch_dense": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _i64, _vp]),
    "dh_stitch_dense_ex": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _i64, _vp]),
    "dh_stitch_scatter": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _i64, _i64, _i64, _vp]),
    "dh_stitch_finalize": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "dh_upload_rects": (C.c_int, [_vp, _i64, _i64, _vp, _i64, _vp, _vp]),
    "dh_host_device_pointer": (C.c_int, [_vp, C.POINTER(_u64)]),
    "dh_stitch_binned_scratch_bytes": (_i64, [_i64, _i32, _i32, _i32, _i64, _i64]),
    "dh_stitch_binned": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "dh_stitch_binned_set_tile_rows": (C.c_int, [_i32]),
    "dh_stitch_binned_set_variant": (C.c_int, [_i32]),
    "dh_stitch_dense_set_variant": (C.c_int, [_i32]),
    "dh_maxpool3x3s2_nhwc": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _vp]),
    "dh_maxpool3x3s2_d2s": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _vp]),
    "dh_colorize_overlay": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i32, _vp, _f64, _vp, _vp, _vp, _vp]),
    "dh_cover_scratch_words": (_i64, [_i64, _i64]),
    "dh_cover_init": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _vp]),
    "dh_cover_sample": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _u64, _u64, _vp, _vp, _vp, _i32, _vp]),
    "dh_cover_sample_group": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _u64, _u64, _i32, _vp, _vp, _vp, _vp]),
    "dh_cover_set_variant": (C.c_int, [_i32]),
    "dh_region_accept_dense": (C.c_int, [_vp, _i32, _i32, _i64, _i64, _i64, _i64, _i32, _i32, _f64, _vp, _vp, _vp]),
    "dh_compact_coords": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _i32, _vp, _vp, _vp]),
    "dh_region_sample": (
        C.c_int,
        [C.POINTER(RegionTables), _i64, _i32, _i32, _f64, _i32, _i32, _i32, _i64, _u64, _u64, _vp, _vp, _vp, _vp, _vp],
    ),
    "dh_rasterize_polygons": (C.c_int, [_vp, _vp, _vp, _i32, _f64, _vp, _i64, _i64, _vp]),
}

_lib = None
_device_ok = False


def load() -> C.CDLL:
    """Load the shared library (no GPU needed for loading)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise DeepHistoError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or deephisto_b200/csrc/build.sh. There is no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def require_device() -> C.CDLL:
    """Load the library and make sure the current CUDA device can run the sm_100a kernels."""
    global _device_ok
    lib = load()
    if not _device_ok:
        rc = lib.dh_device_check()
        if rc != 0:
            raise DeepHistoError(f"deephisto_b200 needs a B200 (sm_100) GPU: {last_error()}")
        _device_ok = True
    return lib


def last_error() -> str:
    return load().dh_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise DeepHistoError(f"{what} failed ({rc}): {last_error()}")
