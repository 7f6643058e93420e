"""
CPU ORACLE for the GraphEm layout iteration -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (graphem_rapids_b200) never does.

This is a torch-CPU restatement of ONE hot path of sashakolpakov/graphem-rapids,
`GraphEmbedderPyTorch.update_positions` and what it calls.  All citations are
file:line inside the reference's graphem_rapids/backends/embedder_pytorch.py.

Parity status: PINNED.  The restatement is checked (tests/test_oracle_golden.py)
against golden vectors produced by importing the real reference in the build
container (tests/golden/make_golden.py, committed together with the vectors).

The arithmetic that lives in torch (third party, `torch>=2.0.0`, unpinned in the
reference; 2.11.0 here) is restated where it matters for neighbour selection:
  * torch.cdist, matmul mode (more than 25 rows on either side):
        x1_ = [-2x, |x|^2, 1], x2_ = [y, 1, |y|^2];  D = sqrt(max(0, x1_ @ x2_^T))
    which on CPU is bit-equal to a sequential fp32 FMA chain in K order
    (see knn_chain.c and `cdist_chain_sq`).
  * torch.cdist, direct mode (<= 25 rows on both sides):
        sqrt(fma(d2,d2, fma(d1,d1, d0*d0)))
  * torch.topk does NOT break ties by index; the specification used by the CUDA
    path (north star: "ties broken by index") orders by (distance, index) with a
    correctly rounded sqrt.  `knn_strict` implements that order; `knn_reference`
    is the literal cdist+topk call sequence of the reference.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------
# edge extraction  (embedder_pytorch.py:220-245)
# --------------------------------------------------------------------------
def extract_edges(adjacency) -> np.ndarray:
    """Upper-triangular (i<j) edge list in CSR nonzero() order (:235-240)."""
    rows, cols = adjacency.nonzero()
    mask = rows < cols
    return np.column_stack([rows[mask], cols[mask]])


# --------------------------------------------------------------------------
# (a) spring forces  (embedder_pytorch.py:595-636)
# --------------------------------------------------------------------------
def spring_forces(pos: torch.Tensor, edges: torch.Tensor, k_attr: float, L_min: float) -> torch.Tensor:
    p1 = pos[edges[:, 0]]                                    # :618
    p2 = pos[edges[:, 1]]                                    # :619
    diff = p2 - p1                                           # :622
    dist = torch.norm(diff, dim=1, keepdim=True) + 1e-6      # :623
    force_magnitude = -k_attr * (dist - L_min)               # :626
    edge_forces = force_magnitude * (diff / dist)            # :629
    forces = torch.zeros_like(pos)                           # :632
    forces.index_add_(0, edges[:, 0], edge_forces)           # :633
    forces.index_add_(0, edges[:, 1], -edge_forces)          # :634
    return forces


# --------------------------------------------------------------------------
# (b) midpoints (embedder_pytorch.py:785) and KNN (:381-424, :543-593)
# --------------------------------------------------------------------------
def midpoints(pos: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    return (pos[edges[:, 0]] + pos[edges[:, 1]]) / 2.0      # :785


def knn_reference(query: torch.Tensor, ref: torch.Tensor, kp1: int, chunk_size: int) -> torch.Tensor:
    """Literal `_compute_knn_torch` (:569-593): chunked cdist + topk, (S,kp1) int64.

    Tie order is whatever torch.topk returns.  Raises RuntimeError when kp1 > E,
    exactly like the reference (that is what its all-zero-adjacency test relies on).
    """
    out = []
    n_query = query.shape[0]
    for i in range(0, n_query, chunk_size):                  # :572
        chunk = query[i:min(i + chunk_size, n_query)]        # :573-574
        d = torch.cdist(chunk, ref, p=2)                     # :580
        _, idx = torch.topk(d, kp1, dim=1, largest=False)    # :583
        out.append(idx)
        del d
    return torch.cat(out, dim=0)                             # :593


def sq_norm_rows(x: torch.Tensor) -> torch.Tensor:
    """`x.pow(2).sum(-1)` as torch's _euclidean_dist computes it; for d<=3 this is the
    non-fused left-to-right sum ((x0^2+x1^2)+x2^2) [probed, SURVEY section 7]."""
    return x.pow(2).sum(-1)


def _fma32(a: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    """fp32 fused multiply-add emulated in fp64: the product of two fp32 numbers is
    exact in fp64 (48 bits); the fp64 sum is rounded once to 53 bits and once more to
    24.  Double rounding can only differ from a true fma on an exact 53-bit half-way
    pattern, which `knn_chain.c` (true fmaf) cross-checks in the tests."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def cdist_chain_sq(query: np.ndarray, ref: np.ndarray) -> np.ndarray:
    """Clamped squared distance of torch.cdist's matmul mode as an explicit FMA chain.

    acc = 0; for j<d: acc = fma(-2*q_j, y_j, acc); acc = fma(|q|^2, 1, acc);
    acc = fma(1, |y|^2, acc); return max(acc, 0)     (numpy, small sizes only)
    """
    q = np.ascontiguousarray(query, dtype=np.float32)
    y = np.ascontiguousarray(ref, dtype=np.float32)
    d = q.shape[1]
    qn = sq_norm_rows(torch.from_numpy(q)).numpy()
    yn = sq_norm_rows(torch.from_numpy(y)).numpy()
    acc = np.zeros((q.shape[0], y.shape[0]), dtype=np.float32)
    for j in range(d):
        acc = _fma32((np.float32(-2.0) * q[:, j])[:, None], y[None, :, j], acc)
    acc = _fma32(qn[:, None], np.float32(1.0), acc)
    acc = _fma32(np.float32(1.0), yn[None, :], acc)
    return np.maximum(acc, np.float32(0.0))


def cdist_direct_sq(query: np.ndarray, ref: np.ndarray) -> np.ndarray:
    """Squared distance of torch.cdist's direct mode (both sides <= 25 rows):
    acc = d0*d0; acc = fma(d_j, d_j, acc) for j = 1..d-1   [probed]."""
    q = np.ascontiguousarray(query, dtype=np.float32)
    y = np.ascontiguousarray(ref, dtype=np.float32)
    diff = q[:, None, :] - y[None, :, :]
    acc = diff[..., 0] * diff[..., 0]
    for j in range(1, q.shape[1]):
        acc = _fma32(diff[..., j], diff[..., j], acc)
    return acc


def uses_mm_mode(n_query: int, n_ref: int) -> bool:
    """torch.cdist default compute_mode: matmul form iff P > 25 or R > 25."""
    return n_query > 25 or n_ref > 25


_LIB = None


def _load_c():
    """Load oracle/_build/liboracle_knn.so (compiled by oracle/Makefile or build())."""
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle_knn.so")
        if not os.path.exists(path):
            build_c()
        lib = ctypes.CDLL(path)
        lib.oracle_knn_strict.restype = ctypes.c_int
        lib.oracle_knn_strict.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = lib
    return _LIB


def build_c() -> str:
    """gcc -O2 -ffp-contract=off -fopenmp knn_chain.c -> _build/liboracle_knn.so."""
    import subprocess
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "liboracle_knn.so")
    src = os.path.join(_HERE, "knn_chain.c")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp",
                               "-shared", "-fPIC", src, "-o", out, "-lm"])
    return out


def knn_strict(mid: torch.Tensor, samp: torch.Tensor, kp1: int, mm_mode: Optional[bool] = None):
    """KNN of mid[samp] among mid under the (distance, index) total order.

    distance = correctly-rounded sqrt of the clamped cdist value (matmul-mode FMA chain or
    direct-mode chain, chosen like torch.cdist does unless `mm_mode` is forced).
    Returns (idx (S,kp1) int64, dist (S,kp1) float32), rows sorted ascending.
    Raises RuntimeError when kp1 > E (the reference's topk does, :583).
    """
    mid_np = np.ascontiguousarray(mid.detach().cpu().numpy(), dtype=np.float32)
    samp_np = np.ascontiguousarray(samp.detach().cpu().numpy(), dtype=np.int64)
    E, d = mid_np.shape
    S = samp_np.shape[0]
    if kp1 > E:
        raise RuntimeError("selected index k out of range")
    if mm_mode is None:
        mm_mode = uses_mm_mode(S, E)
    out_idx = np.empty((S, kp1), dtype=np.int64)
    out_dist = np.empty((S, kp1), dtype=np.float32)
    lib = _load_c()
    rc = lib.oracle_knn_strict(mid_np.ctypes.data, E, d, samp_np.ctypes.data, S, kp1,
                               1 if mm_mode else 0, out_idx.ctypes.data, out_dist.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_knn_strict failed rc={rc}")
    return torch.from_numpy(out_idx), torch.from_numpy(out_dist)


# --------------------------------------------------------------------------
# (c) intersection forces (embedder_pytorch.py:638-736, :738-774)
# --------------------------------------------------------------------------
def check_line_intersections(p1, p2, q1, q2) -> torch.Tensor:
    def orientation(a, b, c):                                # :760-763
        return (b[..., 0] - a[..., 0]) * (c[..., 1] - a[..., 1]) - \
               (b[..., 1] - a[..., 1]) * (c[..., 0] - a[..., 0])
    o1 = orientation(p1, p2, q1)                             # :766
    o2 = orientation(p1, p2, q2)
    o3 = orientation(q1, q2, p1)
    o4 = orientation(q1, q2, p2)
    return (o1 * o2 < 0) & (o3 * o4 < 0)                     # :772


def intersection_pairs(pos, edges, knn, samp):
    """The surviving (edge_i, edge_j) pairs, in the reference's order (:666-715)."""
    _, n_neighbors = knn.shape
    cand_i = samp.unsqueeze(1).expand(-1, n_neighbors).flatten()    # :668
    cand_j = knn.flatten()                                          # :669
    valid = cand_i < cand_j                                         # :672
    vi, vj = cand_i[valid], cand_j[valid]
    ei, ej = edges[vi], edges[vj]
    share = ((ei[:, 0] == ej[:, 0]) | (ei[:, 0] == ej[:, 1]) |
             (ei[:, 1] == ej[:, 0]) | (ei[:, 1] == ej[:, 1]))       # :685-690
    keep = ~share
    vi, vj, ei, ej = vi[keep], vj[keep], ei[keep], ej[keep]
    if vi.numel() == 0:
        return vi, vj
    x = check_line_intersections(pos[ei[:, 0]], pos[ei[:, 1]], pos[ej[:, 0]], pos[ej[:, 1]])  # :708
    return vi[x], vj[x]


def intersection_forces(pos, edges, knn, samp, k_inter: float) -> torch.Tensor:
    forces = torch.zeros_like(pos)
    if knn.numel() == 0:
        return forces
    vi, vj = intersection_pairs(pos, edges, knn, samp)
    if vi.numel() == 0:                                      # early exits :674,:694,:710
        return forces
    ei, ej = edges[vi], edges[vj]
    p1, p2 = pos[ei[:, 0]], pos[ei[:, 1]]                    # :702-705
    q1, q2 = pos[ej[:, 0]], pos[ej[:, 1]]
    inter_mid = (p1 + p2 + q1 + q2) / 4.0                    # :722
    for vpos, verts in [(p1, ei[:, 0]), (p2, ei[:, 1]), (q1, ej[:, 0]), (q2, ej[:, 1])]:  # :727
        diff = vpos - inter_mid                              # :730
        dist = torch.norm(diff, dim=1, keepdim=True) + 1e-6  # :731
        repulsion = k_inter * diff / (dist ** 2)             # :732
        forces.index_add_(0, verts, repulsion)               # :734
    return forces


# --------------------------------------------------------------------------
# (d) update (embedder_pytorch.py:796-804)
# --------------------------------------------------------------------------
def update(pos, spring, inter) -> torch.Tensor:
    total = spring + inter                                   # :796
    new = pos + total                                        # :799
    new = new - torch.mean(new, dim=0, keepdim=True)         # :802
    std = torch.std(new, dim=0, keepdim=True) + 1e-6         # :803
    return new / std                                         # :804


# --------------------------------------------------------------------------
# one iteration with every intermediate (update_positions, :776-806)
# --------------------------------------------------------------------------
def layout_step(pos: torch.Tensor, edges: torch.Tensor, samp: torch.Tensor, *, n_neighbors: int,
                k_attr: float = 0.2, L_min: float = 1.0, k_inter: float = 0.5,
                strict: bool = True, chunk_size: int = 1 << 30) -> Dict[str, torch.Tensor]:
    """One `update_positions` from `pos` with the sampled edge ids `samp` given.

    strict=True  -> neighbours ordered by (distance, index)  (the CUDA path's contract)
    strict=False -> literal cdist+topk of the reference (tie order as torch returns it)
    MemoryManager / monitor_memory_usage (utils/memory_management.py:117-208) are
    omitted: they compute nothing.
    """
    F = spring_forces(pos, edges, k_attr, L_min)
    mid = midpoints(pos, edges)
    kp1 = n_neighbors + 1
    out = {"F_spring": F, "mid": mid}
    if strict:
        knn_full, knn_dist = knn_strict(mid, samp, kp1)
        out["knn_dist"] = knn_dist
    else:
        knn_full = knn_reference(mid[samp], mid, kp1, chunk_size)
    knn = knn_full[:, 1:]                                    # :421 drop column 0
    G = intersection_forces(pos, edges, knn, samp, k_inter)
    pi, pj = intersection_pairs(pos, edges, knn, samp) if knn.numel() else (samp[:0], samp[:0])
    new = update(pos, F, G)
    out.update({"knn_full": knn_full, "knn": knn, "F_inter": G, "pairs_i": pi, "pairs_j": pj,
                "new_pos": new})
    return out


def draw_sample(E: int, S: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """The reference's sampling (:404-413): randperm(E)[:S], or arange(E) when S >= E."""
    S = min(S, E)
    if S < E:
        return torch.randperm(E, generator=generator)[:S]
    return torch.arange(E)


def run_layout(pos, edges, num_iterations: int, *, sample_size: int, n_neighbors: int,
               k_attr=0.2, L_min=1.0, k_inter=0.5, generator=None, strict=False,
               chunk_size: int = 1 << 30, samples=None):
    """`run_layout` (:808-833) without tqdm/MemoryManager.  `samples[i]` overrides the draw."""
    E = edges.shape[0]
    for it in range(num_iterations):
        samp = samples[it] if samples is not None else draw_sample(E, sample_size, generator)
        pos = layout_step(pos, edges, samp, n_neighbors=n_neighbors, k_attr=k_attr, L_min=L_min,
                          k_inter=k_inter, strict=strict, chunk_size=chunk_size)["new_pos"]
    return pos
