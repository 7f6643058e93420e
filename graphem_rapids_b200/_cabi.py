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
ABI_VERSION = 3
STAMP_NAMES = ["knn_prep", "spring_csr", "colsum", "knn_scan", "knn_select", "topk_merge_intersect", "normalise"]
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
        ("external_sample", c_int32), ("mm_mode", c_int32), ("coef_slot", c_int32),
    ]


class GemKnnPrepArgs(Structure):
    """Mirror of `struct gem_knn_prep_args`."""
    _fields_ = [
        ("d", c_int32), ("kp1", c_int32), ("s", c_int64), ("e", c_int64),
        ("pos", c_void_p), ("edges", c_void_p), ("e_total", c_int64), ("samp", c_void_p),
        ("draw", c_int32), ("bump", c_int32), ("seed", c_uint64), ("iter_counter", c_void_p),
        ("qmid_in", c_void_p), ("qmid_out", c_void_p),
        ("row_ptr", c_void_p), ("col", c_void_p), ("tau_hint_in", c_void_p), ("tau_hint_out", c_void_p),
        ("bound_mid", c_void_p), ("bound_edges", c_void_p), ("e_bound", c_int64), ("bound_samples", c_int64),
        ("coef_slot", c_int32),
        ("ws", c_void_p), ("ws_bytes", c_size_t),
    ]


class GemKnnPublish(Structure):
    """Mirror of `struct gem_knn_publish`."""
    _fields_ = [("remap", c_void_p), ("peer_base_host", POINTER(c_void_p)), ("world", c_int32),
                ("idx_offset_bytes", c_size_t), ("dist_offset_bytes", c_size_t)]


class GemMergePublish(Structure):
    """Mirror of `struct gem_merge_publish`."""
    _fields_ = [("peer_raw_host", POINTER(c_void_p)), ("peer_xchg_host", POINTER(c_void_p)), ("world", c_int32),
                ("rank", c_int32), ("stats_offset_bytes", c_size_t), ("touched", c_void_p), ("counters", c_void_p)]


# name -> (restype, argtypes); every symbol include/graphem_b200.h declares
SIGNATURES = {
    "gem_abi_version": (c_int, []),
    "gem_abi_struct_sizes": (c_int, [POINTER(c_size_t)]),
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
    "gem_spring_update_csr_push": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64,
                                           c_int, c_float, c_float, POINTER(c_void_p), c_int, c_int, c_void_p, c_void_p,
                                           c_int64, c_void_p, c_void_p]),
    "gem_debug_stamps": (c_int, [c_void_p]),
    "gem_debug_stamp_count": (c_int, []),
    "gem_debug_stamp_words": (c_int, []),
    "gem_coef_slots": (c_int, []),
    "gem_coef_slot_acquire": (c_int, [POINTER(c_int)]),
    "gem_coef_slot_release": (c_int, [c_int]),
    "gem_sample_edges": (c_int, [c_uint64, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p]),
    "gem_query_midpoints": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "gem_knn_workspace_bytes": (c_int, [c_int64, c_int, c_int64, c_int, POINTER(c_size_t)]),
    "gem_knn_midpoints": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "gem_knn_midpoints_shard": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_int64, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "gem_knn_fast_path": (c_int, [c_int64, c_int64, c_int, c_int64, c_int]),
    "gem_knn_prep": (c_int, [POINTER(GemKnnPrepArgs), c_void_p]),
    "gem_knn_scan": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p,
                             c_void_p, c_size_t, c_int, POINTER(GemKnnPublish), c_void_p]),
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
                                         c_void_p, POINTER(GemMergePublish), c_void_p]),
    "gem_intersection_forces": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64,
                                        c_int, c_float, c_void_p, c_void_p]),
    "gem_intersection_forces_range": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64,
                                              c_int, c_float, c_int64, c_int64, c_void_p, c_void_p]),
    "gem_update_workspace_bytes": (c_int, [c_int64, c_int, POINTER(c_size_t)]),
    "gem_update_positions": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p,
                                     c_int, c_void_p]),
    "gem_update_normalise_push": (c_int, [POINTER(c_void_p), c_int, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p,
                                          c_void_p, c_void_p]),
    "gem_update_normalise_all": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int, c_void_p]),
    "gem_remap_indices": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "gem_push_bytes": (c_int, [POINTER(c_void_p), c_int, c_size_t, c_void_p, c_size_t, c_void_p]),
    "gem_rows_scatter": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, POINTER(c_void_p), c_int, c_void_p]),
    "gem_rows_gather": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "gem_layout_step": (c_int, [POINTER(GemPlan), c_void_p]),
    "gem_profile_step": (c_int, [POINTER(GemPlan), c_void_p, POINTER(c_float)]),
    "gem_spmv_cols": (c_int, []),
    "gem_spmv_normalized_adjacency": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float,
                                              c_void_p, c_float, c_void_p]),
    "gem_spmv_normalized_adjacency_vec": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_void_p]),
    "gem_seed_select_max_k": (c_int, []),
    "gem_seed_select_workspace_bytes": (c_int, [c_int64, POINTER(c_size_t)]),
    "gem_seed_select": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
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
            # an older library exists, the sources are newer and the rebuild failed (e.g. no nvcc on this box):
            # say so instead of silently running stale code; the ABI version check below still applies
            import warnings
            warnings.warn(f"graphem_rapids_b200: sources are newer than {_build.LIB} and the rebuild failed ({exc}); "
                          "loading the existing library", RuntimeWarning)
    lib = ctypes.CDLL(_build.LIB)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.gem_abi_version() != ABI_VERSION:
        raise ImportError("graphem_rapids_b200: ABI version mismatch between header and library")
    sizes = (c_size_t * 4)()
    lib.gem_abi_struct_sizes(sizes)
    mine = [ctypes.sizeof(t) for t in (GemPlan, GemKnnPrepArgs, GemKnnPublish, GemMergePublish)]
    if list(sizes) != mine:
        raise ImportError(f"graphem_rapids_b200: struct layouts differ between _cabi.py {mine} and the library {list(sizes)}")
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
