"""ctypes binding of libgravomg_b200.so (the C ABI in include/gravomg_b200.h).

The shared library is built in-tree by ``gravo_mg_b200/csrc/Makefile`` (see
``__graft_entry__.build``). There is no Python or CPU fallback: if the library is missing the
import of this module raises, and every compute entry point fails loudly without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgravomg_b200.so")


class GmgParams(C.Structure):
    """struct gmg_params (include/gravomg_b200.h)."""

    _fields_ = [
        ("ratio", C.c_double),
        ("low_bound", C.c_int32),
        ("cycle_type", C.c_int32),
        ("tolerance", C.c_double),
        ("stopping_criteria", C.c_int32),
        ("pre_iters", C.c_int32),
        ("post_iters", C.c_int32),
        ("max_iter", C.c_int32),
        ("check_voronoi", C.c_int32),
        ("nested", C.c_int32),
        ("sampling_strategy", C.c_int32),
        ("weighting", C.c_int32),
        ("sig06", C.c_int32),
        ("verbose", C.c_int32),
        ("debug", C.c_int32),
        ("ablation", C.c_int32),
        ("ablation_num_points", C.c_int32),
        ("ablation_random", C.c_int32),
        ("smoother", C.c_int32),
        ("omega", C.c_double),
        ("dtype", C.c_int32),
        ("device", C.c_int32),
        ("build_hierarchy", C.c_int32),
        ("cheb_alpha", C.c_double),
    ]


_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_h = C.c_void_p

# name -> (restype, argtypes); exactly the symbols the header declares.
SIGNATURES = {
    "gmg_default_params": (C.c_int, [C.POINTER(GmgParams)]),
    "gmg_create": (C.c_int, [C.POINTER(GmgParams), C.c_int64, _f64p, _i32p, C.c_int32, _i32p, _i32p, _f64p, C.POINTER(_h)]),
    "gmg_destroy": (None, [_h]),
    "gmg_last_error": (C.c_char_p, [_h]),
    "gmg_set_option": (C.c_int, [_h, C.c_char_p, C.c_double]),
    "gmg_get_option": (C.c_int, [_h, C.c_char_p, _f64p]),
    "gmg_num_levels": (C.c_int, [_h, _i32p]),
    "gmg_prolongation_shape": (C.c_int, [_h, C.c_int32, _i64p, _i64p, _i64p]),
    "gmg_get_prolongation": (C.c_int, [_h, C.c_int32, _i32p, _i32p, _f64p]),
    "gmg_clear_prolongations": (C.c_int, [_h]),
    "gmg_set_prolongation": (C.c_int, [_h, C.c_int32, C.c_int64, C.c_int64, _i32p, _i32p, _f64p]),
    "gmg_get_samples": (C.c_int, [_h, C.c_int32, _i32p, _i64p]),
    "gmg_get_nearest_source": (C.c_int, [_h, C.c_int32, _i32p, _i64p]),
    "gmg_get_level_points": (C.c_int, [_h, C.c_int32, _f64p, _i64p]),
    "gmg_get_all_triangles": (C.c_int, [_h, C.c_int32, _i32p, _i64p]),
    "gmg_get_notrimap": (C.c_int, [_h, C.c_int32, _i32p, _i64p]),
    "gmg_solve": (C.c_int, [_h, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int32]),
    "gmg_stage_system": (C.c_int, [_h, C.c_int64, _i32p, _i32p, _f64p, _f64p, C.c_int32]),
    "gmg_solve_staged": (C.c_int, [_h]),
    "gmg_fetch_solution": (C.c_int, [_h, _f64p]),
    "gmg_direct_solve": (C.c_int, [_h, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int32]),
    "gmg_residual": (C.c_int, [_h, C.c_int64, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int32, C.c_int32, _f64p]),
    "gmg_dist_configure": (C.c_int, [_h, C.c_int32, C.c_int32, C.c_int64]),
    "gmg_dist_unique_id": (C.c_int, [C.c_void_p, C.c_int64, _i64p]),
    "gmg_dist_init": (C.c_int, [_h, C.c_void_p, C.c_int64]),
    "gmg_dist_layout": (C.c_int, [_h, C.c_int64, _i32p, _i32p]),
    "gmg_level_pattern": (C.c_int, [_h, C.c_int32, _i32p, _i32p, _i64p, _i64p]),
    "gmg_dist_windows": (C.c_int, [_h, C.c_int32, _i64p, _i64p, _i32p]),
    "gmg_dist_ranges": (C.c_int, [_h, C.c_int32, _i64p, _i32p]),
    "gmg_dist_halo": (C.c_int, [_h, C.c_int32, C.c_int32, C.c_int32, _i32p, _i64p, _i32p, _i64p]),
    "gmg_timing_keys": (C.c_int, [_h, C.c_int32, C.c_char_p, C.c_int64]),
    "gmg_get_timing": (C.c_int, [_h, C.c_int32, C.c_char_p, _f64p]),
    "gmg_get_convergence": (C.c_int, [_h, _f64p, _f64p, _i32p]),
    "gmg_level_info": (C.c_int, [_h, C.c_int32, _i64p, _i64p, _i64p]),
    "gmg_get_smoother_weights": (C.c_int, [_h, C.c_int32, _f64p, _f64p, _f64p]),
    "gmg_get_level_matrix": (C.c_int, [_h, C.c_int32, _i32p, _i32p, _f64p]),
    "gmg_level_op": (C.c_int, [_h, C.c_int32, C.c_int32, _f64p, _f64p, _f64p, C.c_int32]),
    "gmg_time_op": (C.c_int, [_h, C.c_int32, C.c_int32, C.c_int32, _f64p]),
    "gmg_kernel_profile": (C.c_int, [_h, C.c_int32, C.c_int32, _f64p, _i64p]),
    "gmg_reset_kernel_profile": (C.c_int, [_h]),
    "gmg_last_launch_count": (C.c_int, [_h, _i64p]),
    "gmg_get_trace": (C.c_int, [_h, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), _i64p]),
    "gmg_update_values_device": (C.c_int, [_h, C.c_void_p, C.c_void_p, C.c_int32]),
    "gmg_fetch_solution_device": (C.c_int, [_h, C.c_void_p]),
    "gmg_mesh_attach": (C.c_int, [_h, C.c_int64, _i32p]),
    "gmg_mesh_set_positions": (C.c_int, [_h, _f64p]),
    "gmg_mesh_stiffness": (C.c_int, [_h]),
    "gmg_mesh_mass": (C.c_int, [_h, C.c_int32]),
    "gmg_mesh_system": (C.c_int, [_h, C.c_double, C.c_double, _f64p, C.c_int32]),
    "gmg_mesh_flow": (C.c_int, [_h, C.c_double, C.c_int32, C.c_int32]),
    "gmg_mesh_get": (C.c_int, [_h, C.c_int32, _f64p]),
    "gmg_mesh_pattern": (C.c_int, [C.c_int64, C.c_int64, _i32p, _i32p, _i32p, _i64p]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C gravo_mg_b200/csrc`). gravo_mg_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()


def i32(a):
    return a.ctypes.data_as(_i32p)


def f64(a):
    return a.ctypes.data_as(_f64p)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def check(handle, status):
    if status != 0:
        msg = lib.gmg_last_error(handle)
        raise RuntimeError(msg.decode("utf-8", "replace") if msg else "gravomg_b200: unknown error")
