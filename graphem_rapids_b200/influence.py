"""graphem_seed_selection -- graphem_rapids/influence.py:10-37 on the B200 embedder.
(The NDlib simulation helpers of that file are third-party driven and out of scope.)"""
import ctypes

import numpy as np


def seed_selection_reference(positions: np.ndarray, k: int) -> list:
    """The contract of the device kernel on host arrays: np.linalg.norm(positions, axis=1) (fp32, influence.py:31-32)
    ordered by (radius descending, vertex id ascending).  Identical to the reference's `np.argsort(-r)[:k].tolist()`
    whenever the k+1 largest radii are distinct (numpy's default argsort is not stable: the order of exactly equal
    radii is the one thing it leaves unspecified)."""
    r = np.linalg.norm(np.asarray(positions, dtype=np.float32), axis=1)
    return np.lexsort((np.arange(len(r)), -r.astype(np.float64)))[:k].tolist()


def device_seed_selection(embedder, k: int, return_radii: bool = False):
    """The k vertices with the largest radius from the embedder's DEVICE state (gem_seed_select: fused radius + 64-bit
    radix select; no (n, d) device->host copy)."""
    import torch
    from . import _cabi
    lib, dev = embedder._lib, embedder.device
    k = min(int(k), embedder.n)
    if k <= 0:
        return ([], []) if return_radii else []
    if k > lib.gem_seed_select_max_k():
        raise ValueError(f"device seed selection supports k <= {lib.gem_seed_select_max_k()}")
    nbytes = ctypes.c_size_t(0)
    _cabi.check(lib.gem_seed_select_workspace_bytes(k, ctypes.byref(nbytes)), "gem_seed_select_workspace_bytes")
    with torch.cuda.device(dev):
        ws = torch.empty((nbytes.value + 256,), device=dev, dtype=torch.uint8)
        out = torch.empty((k,), device=dev, dtype=torch.long)
        rad = torch.empty((k,), device=dev, dtype=torch.float32)
        pad = getattr(embedder, "_pad_index", None)
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(lib.gem_seed_select(ctypes.c_void_p(embedder._pos.data_ptr()), embedder.n, int(embedder.n_components),
                                        ctypes.c_void_p(pad.data_ptr()) if pad is not None else None, k,
                                        ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(rad.data_ptr()),
                                        ctypes.c_void_p(ws.data_ptr()), nbytes.value, st), "gem_seed_select")
        seeds = out.tolist()                                   # k ids leave the GPU, not n x d floats
        return (seeds, rad.tolist()) if return_radii else seeds


def graphem_seed_selection(embedder, k, num_iterations=20):
    """Run the layout, then return the k vertices with the largest radial distance
    (influence.py:28-37: `np.argsort(-radial_distances)[:k].tolist()`).

    SURVEY 8(f).3: with a B200 embedder the radial norm and the top-k selection run on the device in the library's
    own kernels (gem_seed_select), so a 10 M-vertex layout is never copied to the host just to pick k seeds; the
    result is the same python list[int] (descending radius; exactly equal radii by ascending id)."""
    if hasattr(embedder, "run_layout_device") and hasattr(embedder, "_pos") and hasattr(embedder, "_lib"):
        if num_iterations > 0:
            embedder.run_layout_device(num_iterations)
        if 0 < min(int(k), embedder.n) <= embedder._lib.gem_seed_select_max_k():
            return device_seed_selection(embedder, k)
        return seed_selection_reference(embedder.positions, k)
    embedder.run_layout(num_iterations=num_iterations)
    positions = np.array(embedder.positions)
    radial = np.linalg.norm(positions, axis=1)
    return np.argsort(-radial)[:k].tolist()
