"""Radial-correlation harness on the device (second half).

The reference measures the quality of a layout as the Spearman correlation between the vertices' distance from
the origin and a centrality measure (graphem_rapids/benchmark.py:166-243: scipy.stats.spearmanr on host arrays
after a full device->host copy of the positions, the centralities from networkx).  Here the radii, the average
ranks and the correlation are computed where the positions live, degree comes from the device edge list, and
PageRank is a power iteration with the library's pull SpMV (gem_spmv_normalized_adjacency, the operator of the
initial embedding): only scalars leave the GPU.  The rank / correlation functions are plain torch and
device-agnostic (unit-tested on CPU against scipy); the PageRank needs the CUDA library.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Mapping, Optional

import numpy as np
import torch

from . import _cabi


def rank_average(x: torch.Tensor) -> torch.Tensor:
    """1-based ranks of a 1-D tensor with ties given their average rank (scipy.stats.rankdata 'average'),
    float64, on x's device."""
    x = x.reshape(-1)
    n = x.numel()
    if n == 0:
        return torch.empty((0,), dtype=torch.float64, device=x.device)
    order = torch.argsort(x, stable=True)
    xs = x[order]
    new = torch.ones((n,), dtype=torch.bool, device=x.device)
    new[1:] = xs[1:] != xs[:-1]
    gid = torch.cumsum(new.to(torch.long), 0) - 1                 # tie group of every sorted element
    counts = torch.bincount(gid)
    ends = torch.cumsum(counts, 0)                                # last 1-based rank of the group
    avg = (2 * ends - counts + 1).to(torch.float64) * 0.5         # (first + last) / 2
    ranks = torch.empty((n,), dtype=torch.float64, device=x.device)
    ranks[order] = avg[gid]
    return ranks


def spearman(a: torch.Tensor, b: torch.Tensor) -> float:
    """Spearman rank correlation (Pearson correlation of the average ranks, like scipy.stats.spearmanr);
    nan when either input is constant."""
    a = torch.as_tensor(a).reshape(-1)
    b = torch.as_tensor(b).reshape(-1).to(a.device)
    if a.numel() != b.numel():
        raise ValueError("spearman: inputs must have the same length")
    ra, rb = rank_average(a), rank_average(b)
    ra = ra - ra.mean()
    rb = rb - rb.mean()
    den = torch.sqrt((ra * ra).sum() * (rb * rb).sum())
    return float((ra * rb).sum() / den) if float(den) > 0 else float("nan")


def radii(embedder) -> torch.Tensor:
    """(n,) distances of the vertices from the origin, on the embedder's device (no host copy)."""
    return torch.linalg.vector_norm(embedder._positions.to(torch.float32), dim=1)


def degrees(embedder) -> torch.Tensor:
    """(n,) vertex degrees from the device edge list (original vertex ids)."""
    return torch.bincount(embedder.edges.reshape(-1), minlength=embedder.n)


def _pagerank_power_iteration(deg: torch.Tensor, apply_m, alpha: float, tol: float, max_iter: int) -> torch.Tensor:
    """networkx.pagerank's iteration for an undirected graph, x <- alpha (x P + dangling mass / n) + (1 - alpha) / n
    until the L1 change is below n * tol, written with the symmetric operator M = D^-1/2 A D^-1/2:
        sum_u A[v,u] x[u] / deg(u) = sqrt(deg v) * (M (D^-1/2 x))[v].
    `apply_m(z)` returns M z for an (n,) fp32 tensor."""
    n = deg.numel()
    deg = deg.to(torch.float32)
    live = deg > 0
    dinv = torch.where(live, deg.clamp_min(1).rsqrt(), torch.zeros_like(deg))
    dsqrt = torch.sqrt(deg)
    x = torch.full((n,), 1.0 / n, device=deg.device, dtype=torch.float32)
    for _ in range(int(max_iter)):
        dangling = x[~live].sum()
        new = alpha * (apply_m(x * dinv) * dsqrt + dangling / n) + (1.0 - alpha) / n
        err = float((new - x).abs().sum())
        x = new
        if err < n * tol:
            break
    return x / x.sum()


def pagerank_device(embedder, alpha: float = 0.85, tol: float = 1e-6, max_iter: int = 100) -> torch.Tensor:
    """PageRank of the (undirected) graph by power iteration on the device; the operator is the library's pull
    SpMV over the symmetric CSR of the spring kernel (single right-hand-side form: one 4-byte gather per entry)."""
    if getattr(embedder, "_pad_index", None) is not None:
        raise NotImplementedError("pagerank_device runs on the single-GPU vertex numbering")
    lib, dev, n = embedder._lib, embedder.device, embedder.n
    deg = embedder._row_ptr[1:] - embedder._row_ptr[:-1]
    degf = deg.to(torch.float32)
    dinv = torch.where(degf > 0, degf.clamp_min(1).rsqrt(), torch.zeros_like(degf)).contiguous()

    def apply_m(v):
        v = v.contiguous()
        y = torch.empty_like(v)
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(lib.gem_spmv_normalized_adjacency_vec(
            ctypes.c_void_p(embedder._row_ptr.data_ptr()), ctypes.c_void_p(embedder._col.data_ptr()),
            ctypes.c_void_p(dinv.data_ptr()), ctypes.c_void_p(v.data_ptr()), ctypes.c_void_p(y.data_ptr()), n, 1.0, st),
            "gem_spmv_normalized_adjacency_vec")
        return y

    with torch.cuda.device(dev):
        return _pagerank_power_iteration(deg, apply_m, alpha, tol, max_iter)


def radial_correlations(embedder, measures: Optional[Mapping[str, object]] = None, pagerank: bool = True) -> Dict[str, float]:
    """{'degree': rho, 'pagerank': rho, **{name: rho}}: Spearman correlation of the layout radii with the vertex
    degree, the device PageRank and any further per-vertex measures given as arrays (moved to the device)."""
    r = radii(embedder)
    out = {"degree": spearman(r, degrees(embedder).to(torch.float32))}
    if pagerank:
        out["pagerank"] = spearman(r, pagerank_device(embedder))
    for name, values in (measures or {}).items():
        v = torch.as_tensor(np.asarray(values)).to(r.device)
        out[str(name)] = spearman(r, v)
    return out
