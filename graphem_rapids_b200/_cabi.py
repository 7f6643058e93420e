"""ctypes binding of libgraphem_b200.so (C ABI declared in include/graphem_b200.h).

The product path has NO fallback: if the shared library cannot be loaded the import of the
embedder fails loudly.  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64,
                    c_size_t, c_uint64, c_void_p)

from . import build as _build

_lib = None
ABI_VERSION = 2
STAGE_NAMES = ["sample", "spring_mid", "query_mid", "knn_bound", "knn_threshold", "knn_scan", "knn_select",
               "knn_fallback", "intersect", "update"]
_inited_devices = set()


class GemPlan(Structure):
    """Mirror of `struct gem_plan` (include/graphem_b200.h)."""
    _fields_ = [
        ("n", c_int64), ("e", c_int64), ("s", c_int64),
        ("d", c_int32), ("kp1", c_int32),
        ("k_attr", c_float), ("l_min", c_float), ("k_inter", c_float),
        ("seed", c_uint64),
        ("pos", c_void_p), ("edges", c_void_p), ("row_ptr", c_void_p), ("col", c_void_p),
        ("up_ptr", c_void_p), ("hubs", c_void_p), ("n_hubs", c_int64),
        ("force", c_void_p), ("mid", c_void_p),
        ("qmid", c_void_p), ("tau_hint", c_void_p), ("samp", c_void_p), ("knn_idx", c_void_p), ("knn_dist", c_void_p),
        ("iter_counter", c_void_p),
        ("knn_ws", c_void_p), ("knn_ws_bytes", c_size_t),
        ("stats_ws", c_void_p),
        ("external_sample", c_int32), ("mm_mode", c_int32),
    ]


# name -> (restype, argtypes); every symbol include/graphem_b200.h declares
SIGNATURES = {
    "gem_abi_version": (c_int, []),
    "gem_init": (c_int, []),
    "gem_error_string": (c_char_p, [c_int]),
    "gem_row_pitch": (c_int, [c_int]),
    "gem_mid_pitch": (c_int, [c_int]),
    "gem_spring_midpoints": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_float,
                                     c_void_p, c_void_p, c_void_p]),
    "gem_hub_degree": (c_int, []),
    "gem_spring_midpoints_csr": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64,
                                         c_int, c_float, c_float, c_void_p, c_void_p, c_int64, c_void_p]),
    "gem_spring_update_csr": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64,
                                      c_int, c_float, c_float, c_void_p, c_void_p, c_int64, c_void_p]),
    "gem_sample_edges": (c_int, [c_uint64, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p]),
    "gem_query_midpoints": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "gem_knn_workspace_bytes": (c_int, [c_int64, c_int, c_int64, c_int, POINTER(c_size_t)]),
    "gem_knn_midpoints": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gem_knn_midpoints_shard": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_int64, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gem_knn_fast_path": (c_int, [c_int64, c_int64, c_int, c_int64, c_int]),
    "gem_knn_prepare": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p,
                                c_void_p, c_void_p, c_size_t, c_void_p]),
    "gem_knn_query_prep": (c_int, [c_uint64, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                   c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "gem_knn_scan": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p,
                             c_void_p, c_size_t, c_void_p]),
    "gem_knn_linegraph_hint": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                       c_void_p, c_void_p]),
    "gem_knn_midpoints_exact": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p]),
    "gem_knn_debug_stats": (c_int, [c_int, c_int64, c_int, c_int64, c_int, POINTER(c_size_t), POINTER(c_size_t),
                                    POINTER(c_size_t), POINTER(c_int), POINTER(c_int)]),
    "gem_topk_merge": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "gem_topk_merge_strided": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int, c_void_p,
                                       c_void_p, c_void_p]),
    "gem_topk_merge_intersect": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_int, c_float, c_int64, c_int64, c_void_p,
                                         c_void_p, c_void_p]),
    "gem_intersection_forces": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64,
                                        c_int, c_float, c_void_p, c_void_p]),
    "gem_intersection_forces_range": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64,
                                              c_int, c_float, c_int64, c_int64, c_void_p, c_void_p]),
    "gem_update_workspace_bytes": (c_int, [c_int64, c_int, POINTER(c_size_t)]),
    "gem_update_positions": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p,
                                     c_int, c_void_p]),
    "gem_update_normalise_push": (c_int, [POINTER(c_void_p), c_int, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p,
                                          c_void_p, c_void_p]),
    "gem_remap_indices": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "gem_push_bytes": (c_int, [POINTER(c_void_p), c_int, c_size_t, c_void_p, c_size_t, c_void_p]),
    "gem_layout_step": (c_int, [POINTER(GemPlan), c_void_p]),
    "gem_profile_step": (c_int, [POINTER(GemPlan), c_void_p, POINTER(c_float)]),
    "gem_spmv_cols": (c_int, []),
    "gem_spmv_normalized_adjacency": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float,
                                              c_void_p, c_float, c_void_p]),
    "gem_graph_workspace_bytes": (c_int, [c_int64, POINTER(c_size_t)]),
    "gem_graph_count": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gem_graph_fill": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gem_pack_points": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "gem_check_line_intersections": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                             c_void_p, c_void_p]),
    "gem_fp32_peak_probe": (c_int, [POINTER(c_double), POINTER(c_double), c_void_p]),
}


def library_path() -> str:
    return _build.LIB


def load():
    """Load (building first if the .so is missing or stale and nvcc exists). Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if _build.needs_build():
        try:
            _build.build()
        except Exception as exc:  # no silent fallback: there is no other implementation
            if not os.path.exists(_build.LIB):
                raise ImportError(
                    "graphem_rapids_b200: libgraphem_b200.so is missing and could not be built "
                    f"({exc}). Run `python -c 'import __graft_entry__ as g; g.build()'`.") from exc
    lib = ctypes.CDLL(_build.LIB)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.gem_abi_version() != ABI_VERSION:
        raise ImportError("graphem_rapids_b200: ABI version mismatch between header and library")
    _lib = lib
    return lib


def error_string(code: int) -> str:
    return load().gem_error_string(int(code)).decode()


def check(code: int, what: str = "") -> None:
    """0 -> ok; GEM_E_KRANGE -> the RuntimeError torch.topk raises in the reference
    (embedder_pytorch.py:583); anything else -> RuntimeError with the library's message."""
    if code == 0:
        return
    msg = error_string(code)
    raise RuntimeError(f"graphem_b200 {what}: {msg} (code {code})" if what else f"graphem_b200: {msg} (code {code})")


def init_device(device_index: int) -> None:
    """gem_init() once per device (kernel attributes)."""
    import torch
    if device_index in _inited_devices:
        return
    with torch.cuda.device(device_index):
        check(load().gem_init(), "gem_init")
    _inited_devices.add(device_index)
