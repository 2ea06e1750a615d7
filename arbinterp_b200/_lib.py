"""ctypes binding of the C-ABI in ``include/arbinterp_b200.h``.

There is no CPU fallback: if the shared library is missing the import of anything that needs it
raises, loudly, with the build command.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libarbinterp_b200.so")

MODE_VECTOR, MODE_NORM, MODE_BOTH = 0, 1, 2

# every symbol include/arbinterp_b200.h declares (tests check the .so exports each one)
EXPORTED_SYMBOLS = (
    "arb_version", "arb_last_error", "arb_get_matrix",
    "arb_build_coeffs", "arb_build_coeffs_3d", "arb_build_coeffs_4d",
    "arb_query", "arb_query_host", "arb_query_grid", "arb_query_grid_host",
    "arb_build_nodes", "arb_query_nodes", "arb_query_nodes_host", "arb_query_gridil", "arb_query_gridil_host", "arb_query_routed", "arb_enable_peer_access", "arb_owner_keys", "arb_route_rows", "arb_query_inbox",
    "arb_push", "arb_push_steps", "arb_push_nodes", "arb_permute_rows", "arb_set_query_variant", "arb_set_build_variant",
)


class ArbGeom(ctypes.Structure):
    """``struct arb_geom`` (include/arbinterp_b200.h)."""
    _fields_ = [
        ("d", ctypes.c_int32),
        ("ncomp", ctypes.c_int32),
        ("ncell", ctypes.c_int64 * 4),
        ("slab_lo", ctypes.c_int64),
        ("slab_hi", ctypes.c_int64),
        ("int_min", ctypes.c_double * 4),
        ("int_max", ctypes.c_double * 4),
        ("h", ctypes.c_double * 4),
        ("flags", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


GEOM_FIXED_D4 = 1


class ArbError(RuntimeError):
    pass


_lib = None


def load():
    """Load (once) and return the ctypes handle; raise if the CUDA library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ArbError(
            f"{LIB_PATH} not found: the sm_100a CUDA library has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C arbinterp_b200/csrc`. "
            "arbinterp_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i64, i32, dbl_p = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p
    lib.arb_version.restype = ctypes.c_char_p
    lib.arb_version.argtypes = []
    lib.arb_last_error.restype = ctypes.c_char_p
    lib.arb_last_error.argtypes = []
    lib.arb_get_matrix.restype = i32
    lib.arb_get_matrix.argtypes = [i32, i32, i32, dbl_p]
    lib.arb_build_coeffs.restype = i32
    lib.arb_build_coeffs.argtypes = [i32, vp, i32, ctypes.POINTER(i64 * 4), vp, i32, vp]
    lib.arb_build_coeffs_3d.restype = i32
    lib.arb_build_coeffs_3d.argtypes = [vp, i32, i64, i64, i64, vp, vp]
    lib.arb_build_coeffs_4d.restype = i32
    lib.arb_build_coeffs_4d.argtypes = [vp, i32, i64, i64, i64, i64, vp, vp]
    lib.arb_query.restype = i32
    lib.arb_query.argtypes = [ctypes.POINTER(ArbGeom), vp, i32, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp]
    lib.arb_query_host.restype = i32
    lib.arb_query_host.argtypes = [ctypes.POINTER(ArbGeom), vp, i32, vp, i64, i64, vp, vp, vp, vp, i64]
    lib.arb_query_grid.restype = i32
    lib.arb_query_grid.argtypes = [ctypes.POINTER(ArbGeom), vp, i64, i32, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp]
    lib.arb_query_grid_host.restype = i32
    lib.arb_query_grid_host.argtypes = [ctypes.POINTER(ArbGeom), vp, i64, i32, vp, i64, i64, vp, vp, vp, vp, i64]
    lib.arb_build_nodes.restype = i32
    lib.arb_build_nodes.argtypes = [i32, vp, i32, ctypes.POINTER(i64 * 4), i64, vp, vp]
    lib.arb_query_nodes.restype = i32
    lib.arb_query_nodes.argtypes = lib.arb_query.argtypes
    lib.arb_query_nodes_host.restype = i32
    lib.arb_query_nodes_host.argtypes = lib.arb_query_host.argtypes
    lib.arb_query_gridil.restype = i32
    lib.arb_query_gridil.argtypes = lib.arb_query.argtypes
    lib.arb_query_gridil_host.restype = i32
    lib.arb_query_gridil_host.argtypes = lib.arb_query_host.argtypes
    lib.arb_query_routed.restype = i32
    lib.arb_query_routed.argtypes = [ctypes.POINTER(ArbGeom), vp, i32, vp, i64, i64, vp, vp, ctypes.POINTER(vp), i32, i64, vp]
    lib.arb_route_rows.restype = i32
    lib.arb_route_rows.argtypes = [ctypes.POINTER(ArbGeom), vp, i64, i64, vp, i32, i32, ctypes.POINTER(vp), ctypes.POINTER(vp),
                                   i64, vp, vp, vp, vp]
    lib.arb_query_inbox.restype = i32
    lib.arb_query_inbox.argtypes = [ctypes.POINTER(ArbGeom), vp, i32, vp, vp, i64, ctypes.POINTER(vp), i32, i64, vp]
    lib.arb_owner_keys.restype = i32
    lib.arb_owner_keys.argtypes = [ctypes.POINTER(ArbGeom), vp, i64, i64, vp, i32, vp, vp, vp]
    lib.arb_enable_peer_access.restype = i32
    lib.arb_enable_peer_access.argtypes = [i32]
    lib.arb_push.restype = i32
    lib.arb_push.argtypes = [ctypes.POINTER(ArbGeom), vp, i32, vp, vp, i64, ctypes.c_double, i64, ctypes.c_double,
                             ctypes.POINTER(ctypes.c_double * 3), vp, vp]
    lib.arb_push_nodes.restype = i32
    lib.arb_push_nodes.argtypes = lib.arb_push.argtypes
    lib.arb_push_steps.restype = i32
    lib.arb_push_steps.argtypes = [ctypes.POINTER(ArbGeom), vp, i32, vp, vp, vp, i64, ctypes.c_double, i64,
                                   ctypes.c_double, ctypes.POINTER(ctypes.c_double * 3), vp, vp]
    lib.arb_permute_rows.restype = i32
    lib.arb_permute_rows.argtypes = [vp, vp, vp, i64, i32, i32, vp]
    lib.arb_set_query_variant.restype = i32
    lib.arb_set_query_variant.argtypes = [i32]
    lib.arb_set_build_variant.restype = i32
    lib.arb_set_build_variant.argtypes = [i32]
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().arb_last_error().decode("utf-8", "replace")
        raise ArbError(f"{what} failed (code {rc}): {msg}")


def get_matrix(d: int, which: str, reference_quirk: bool = True):
    """inv(B) / D / A of makeAMatrix (A.py:107-175, 726-878) as a numpy array."""
    import numpy as np
    nm = 4 ** d
    out = np.empty((nm, nm), dtype=np.float64)
    code = {"invB": 0, "D": 1, "A": 2}[which]
    check(load().arb_get_matrix(d, code, int(reference_quirk), out.ctypes.data), "arb_get_matrix")
    return out


def permute_rows(src, order, scatter: bool = False):
    """Row gather ``src[order]`` / scatter ``out[order] = src`` of a contiguous float64 (N, w) CUDA tensor with the
    library's kernel (torch's row gather is ~30x slower for 24-64 byte rows); CPU tensors use torch."""
    import torch
    if src.dim() != 2 or src.dtype != torch.float64:
        raise TypeError(f"permute_rows moves float64 (N, w) rows, got {src.dtype} with shape {tuple(src.shape)}")
    if order.dtype != torch.int64 or order.shape != (src.shape[0],):
        raise TypeError("permute_rows needs an int64 order with one entry per row")
    if not src.is_cuda:
        if scatter:
            out = torch.empty_like(src)
            out[order] = src
            return out
        return src[order]
    src = src.contiguous()
    out = torch.empty_like(src)
    with torch.cuda.device(src.device):
        check(load().arb_permute_rows(out.data_ptr(), src.data_ptr(), order.contiguous().data_ptr(), src.shape[0],
                                      src.shape[1], int(scatter), torch.cuda.current_stream(src.device).cuda_stream),
              "arb_permute_rows")
    return out
