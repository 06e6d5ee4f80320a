"""ctypes binding of libhx_b200.so (include/hx_b200.h).

The product path never falls back to the CPU: if the shared library is missing or
a call fails, an exception is raised (``HxLibraryError`` / ``HxError``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhx_b200.so")

vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double


class HxLibraryError(RuntimeError):
    pass


class HxError(RuntimeError):
    pass


# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "hx_last_error": [],
    "hx_version": [],
    "hx_launch_count": [],
    "hx_launch_count_reset": [],
    "hx_launch_count_add": [i64],
    "hx_spmv_zz": [i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp],
    "hx_spmv_dz": [i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp],
    "hx_spmv_sell_zz": [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp],
    "hx_jacobi_sell": [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, f64, i32, vp],
    "hx_sell_slice_widths": [i32, vp, vp, i32, vp, vp],
    "hx_sell_fill": [i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp],
    "hx_sell_gather": [i64, vp, vp, vp, vp],
    "hx_spmv_cc": [i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp],
    "hx_spmv_sc": [i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp],
    "hx_spmv_sell_cc": [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp],
    "hx_spmv_sell_sc": [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp],
    "hx_sell_gather_s": [i64, vp, vp, vp, vp],
    "hx_jacobi_sell_c": [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, f64, i32, vp],
    "hx_jacobi_sweep_c": [i32, vp, vp, vp, vp, vp, vp, vp, f64, i32, vp],
    "hx_sell_gather_c": [i64, vp, vp, vp, vp],
    "hx_combine_abc": [i64, vp, vp, vp, vp, vp, vp, vp, vp],
    "hx_combine_zzz": [i64, vp, vp, vp, vp, vp, vp, vp, vp],
    "hx_lowrank_dots": [i32, vp, vp, vp, vp, vp, vp],
    "hx_lowrank_update": [i32, vp, vp, vp, vp, vp, vp, vp, vp],
    "hx_reduce_scratch_bytes": [i32],
    "hx_multi_dot": [i64, i32, vp, i64, vp, i32, vp, vp, vp],
    "hx_multi_axpy": [i64, i32, vp, i64, vp, vp, vp, vp, vp, vp],
    "hx_scale_copy": [i64, vp, vp, vp, vp, vp],
    "hx_axpby": [i64, vp, vp, vp, vp, vp],
    "hx_basis_rotate": [i64, i32, i32, vp, i64, vp, i32, vp, i64, vp],
    "hx_basis_rotate_dmma": [i64, i32, i32, vp, i64, vp, i32, vp, i64, vp],
    "hx_jacobi_sweep": [i32, vp, vp, vp, vp, vp, vp, vp, f64, i32, vp],
    "hx_extract_diag_inv": [i32, vp, vp, vp, vp, vp],
    "hx_ilu0_factor": [i32, vp, vp, vp, vp, i32, vp, vp, vp],
    "hx_ilu0_solve": [i32, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, vp, vp],
    "hx_diag_positions": [i32, vp, vp, vp, vp],
    "hx_dense_inverse": [i32, vp, vp, vp],
    "hx_dense_gemv": [i32, vp, vp, vp, vp],
    "hx_amg_tail": [vp, vp, vp],
    "hx_spgemm_symbolic": [i32, vp, vp, vp, vp, vp, vp, vp, i32, vp],
    "hx_spgemm_numeric": [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "hx_spgemm_numeric_small": [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "hx_dof_cell_count": [i64, i32, vp, i32, vp, vp],
    "hx_dof_cell_fill": [i64, i32, vp, i32, vp, vp, vp, vp],
    "hx_pattern_rows": [i32, i32, vp, vp, vp, vp, vp, vp, i32, vp],
    "hx_color_cells": [i64, i32, vp, vp, vp, vp, vp, i32, vp, vp],
    "hx_color_cells_h": [i64, i32, vp, i32, vp],
    "hx_assemble_AC": [i32, i64, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp],
    "hx_assemble_B": [i32, i64, vp, vp, vp, vp, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp],
    "hx_apply_dirichlet": [i32, vp, vp, vp, vp, i32, vp],
    "hx_facet_integrals": [i64, vp, vp, vp, vp, vp, vp],
    "hx_cell_volumes": [i64, vp, vp, vp, vp],
    "hx_flame_left": [i32, i64, vp, vp, vp, vp, f64, vp, i32, f64, vp, i32, i32, vp, vp, vp, vp],
    "hx_flame_right": [i32, i64, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp],
    "hx_locate_points": [i64, vp, vp, i32, vp, f64, vp, vp],
    "hx_point_dphidz": [i32, vp, vp, i32, vp, vp, vp, vp],
    "hx_shape_derivative": [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "hx_threshold": [i64, vp, f64, vp],
    "hx_peer_alloc": [i64, vp, vp],
    "hx_peer_open": [vp, vp],
    "hx_peer_close": [vp],
    "hx_peer_free": [vp],
    "hx_peer_halo_exchange": [vp, vp, i32, vp],
    "hx_peer_allreduce": [vp, vp, vp, i64, i32, vp],
}
_RESTYPES = {"hx_last_error": C.c_char_p, "hx_launch_count": i64, "hx_reduce_scratch_bytes": i64,
             "hx_launch_count_reset": None, "hx_launch_count_add": None}
_NO_CHECK = {"hx_last_error", "hx_version", "hx_launch_count", "hx_launch_count_reset", "hx_launch_count_add",
             "hx_reduce_scratch_bytes", "hx_color_cells_h"}

_lib = None


def load():
    """Load the shared library (once).  Raises HxLibraryError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HxLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(helmholtz_x_b200 has no CPU fallback)")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:
        raise HxLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, args in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise HxLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, i32)
    _lib = lib
    return lib


def call(name, *args):
    """Call an entry point and raise HxError on a negative return code."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if name not in _NO_CHECK and rc != 0:
        raise HxError(f"{name} failed ({rc}): {lib.hx_last_error().decode()}")
    return rc
