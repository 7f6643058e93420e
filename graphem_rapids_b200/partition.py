"""Host-side graph layout for the CUDA path (numpy / scipy only; no torch, no CUDA -- unit-tested on CPU).

The device works on a *padded vertex numbering*: with G ranks, rank r stores the vertices it owns at
rows [r*slice, r*slice + count_r) of every per-vertex array, `slice` = the largest count.  Rows
r*slice + count_r .. (r+1)*slice - 1 are dummies (degree 0, position 0, never referenced).  Every
rank's block has the same size, so the position exchange of the multi-GPU iteration is one
equal-chunk store per peer and needs no unpacking.  Two ownership rules (`build_layout`):
'strided' (default for G > 1; vertex v -> rank v mod G, balanced for any vertex order) and
'contiguous' (cost-balanced ranges of original ids, monotonic numbering).  With G = 1 the mapping is
the identity and there are no dummies.

Edge ids are NOT renumbered in the public sense: the edge list stays in the reference's order
(embedder_pytorch.py:220-245, nonzero() order of the upper triangle), because edge ids are what the
sampler draws and what the KNN returns; `edges32` is indexed by original edge id.  What differs per
ownership is the order in which a rank's spring kernel PRODUCES the midpoints of the edges it owns
(rows in padded order, each row its owned edges in edge-list order).  Under contiguous ownership
that order is the edge-list order itself (rank r owns the contiguous edge range
[up_ptr[row_begin], up_ptr[row_end])); under strided ownership it is a permutation, recorded in
`edge_orig` (local-order number -> original edge id) and undone on the device by
`gem_remap_indices` before ties are broken.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


def edges_sorted_by_ij(edges: np.ndarray) -> bool:
    """True when the (E,2) list is strictly increasing in (i, j) with i < j (CSR nonzero() order
    of a canonical matrix) -- the precondition of the vertex-parallel spring kernel."""
    if len(edges) == 0:
        return True
    e = np.asarray(edges, dtype=np.int64)
    if not np.all(e[:, 0] < e[:, 1]):
        return False
    key = e[:, 0] * (int(e.max()) + 1) + e[:, 1]
    return bool(np.all(key[1:] > key[:-1]))


@dataclass
class GraphLayout:
    n: int                      # true vertex count
    n_edges: int
    world: int
    slice: int                  # rows per rank block
    n_pad: int                  # world * slice
    rank_count: np.ndarray      # (world,) valid rows of each rank block
    e_lo: np.ndarray            # (world,) range of the rank's edges in LOCAL-ORDER edge numbering (see edge_orig)
    e_hi: np.ndarray
    pad_of: np.ndarray          # (n,) int64: original id -> padded row
    edges32: np.ndarray         # (E,2) int32, padded ids, indexed by ORIGINAL edge id
    row_ptr: np.ndarray         # (n_pad+1,) int64  symmetric CSR over padded rows
    col: np.ndarray             # (2E,) int32 padded ids; per row ordered by ORIGINAL neighbour id
    up_ptr: np.ndarray          # (n_pad+1,) int64: #owned (first-endpoint) edges of the rows before this one
    edge_orig: Optional[np.ndarray] = None   # (E,) int64: local-order edge number -> original edge id; None = identity
    hubs: List[np.ndarray] = field(default_factory=list)   # per rank: padded ids with degree > hub_degree
    sorted_edges: bool = True
    ownership: str = "contiguous"
    v_lo: Optional[np.ndarray] = None        # contiguous ownership only: original id range per rank
    v_hi: Optional[np.ndarray] = None
    on_device: bool = False     # single-GPU arrays built by the library on the device (embedder._graph_arrays_device):
                                # pad_of / edges32 / row_ptr / col / up_ptr are then None here (identity numbering)

    def rank_rows(self, r: int):
        """[begin, end) of the rank's VALID rows in padded numbering."""
        b = r * self.slice
        return b, b + int(self.rank_count[r])

    def local_edge_ids(self, r: int) -> np.ndarray:
        """Original ids of the edges rank r owns, in the order its spring kernel writes their midpoints."""
        lo, hi = int(self.e_lo[r]), int(self.e_hi[r])
        return np.arange(lo, hi, dtype=np.int64) if self.edge_orig is None else self.edge_orig[lo:hi]

    def pad_positions(self, pos: np.ndarray, ld: int) -> np.ndarray:
        out = np.zeros((self.n_pad, ld), dtype=np.float32)
        out[self.pad_of if self.pad_of is not None else slice(None), : pos.shape[1]] = pos
        return out


def balanced_vertex_ranges(deg: np.ndarray, up: np.ndarray, world: int, w_entry: float = 1.0, w_edge: float = 4.5,
                           w_vertex: float = 3.0):
    """Contiguous vertex ranges with (approximately) equal cost
        cost(v) = w_entry*deg(v) + w_edge*up(v) + w_vertex
    (spring work ~ CSR entries, KNN work ~ owned candidate edges, update + position exchange ~
    vertices).  The weights are the measured single-B200 costs on the 1M-vertex BA graph in units of
    the per-entry spring cost: scan 45 us / M edges, spring 10 us / M entries, update 27 us / M rows.
    Every rank gets >= 1 vertex."""
    n = len(deg)
    if world > n:
        raise ValueError(f"cannot shard {n} vertices across {world} ranks")
    cost = w_entry * deg.astype(np.float64) + w_edge * up.astype(np.float64) + w_vertex
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    targets = cum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(cum, targets, side="left")
    bounds = np.concatenate([[0], cuts, [n]]).astype(np.int64)
    for r in range(1, world + 1):                       # strictly increasing, room for the ranks after r
        bounds[r] = max(bounds[r], bounds[r - 1] + 1)
    for r in range(world - 1, 0, -1):
        bounds[r] = min(bounds[r], bounds[r + 1] - 1)
    return bounds[:-1].copy(), bounds[1:].copy()


def build_layout(edges: np.ndarray, n: int, world: int = 1, hub_degree: int = 128,
                 ownership: str = "strided") -> GraphLayout:
    """ownership (world > 1):
      'strided'    vertex v belongs to rank v mod world (row (v mod world)*slice + v div world).  Rows, owned
                   edges and CSR entries are balanced for any vertex order -- in particular for
                   preferential-attachment graphs, whose hubs are the lowest ids and make contiguous ranges
                   3x unbalanced in rows at 8 ranks.  The padded numbering is then not monotonic: CSR rows stay
                   ordered by ORIGINAL neighbour id (the spring kernel relies on "the last up(v) entries of row v
                   are its owned edges, in edge-list order"), and a rank's edges are numbered in the order its rows
                   produce them (`edge_orig` maps that numbering back to original edge ids).
      'contiguous' cost-balanced contiguous ranges (monotonic numbering, edge_orig = identity)."""
    import scipy.sparse as sp
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    E = len(e)
    if world > n:
        raise ValueError(f"cannot shard {n} vertices across {world} ranks")
    is_sorted = edges_sorted_by_ij(e)
    deg = np.bincount(e.ravel(), minlength=n).astype(np.int64) if E else np.zeros(n, np.int64)
    up = np.bincount(e[:, 0], minlength=n).astype(np.int64) if E else np.zeros(n, np.int64)
    v_lo = v_hi = None
    if world == 1:
        ownership = "contiguous"
        v_lo, v_hi = np.array([0], np.int64), np.array([n], np.int64)
    elif ownership == "contiguous":
        v_lo, v_hi = balanced_vertex_ranges(deg, up, world)
    elif ownership != "strided":
        raise ValueError("ownership must be 'strided' or 'contiguous'")
    pad_of = np.empty(n, dtype=np.int64)
    if ownership == "contiguous":
        rank_count = (v_hi - v_lo).astype(np.int64)
        slice_rows = int(rank_count.max())
        for r in range(world):
            pad_of[v_lo[r]:v_hi[r]] = r * slice_rows + np.arange(v_hi[r] - v_lo[r])
    else:
        slice_rows = (n + world - 1) // world
        v = np.arange(n, dtype=np.int64)
        pad_of[:] = (v % world) * slice_rows + v // world
        rank_count = np.array([(n - r + world - 1) // world for r in range(world)], dtype=np.int64)
    n_pad = world * slice_rows
    if n_pad >= 2 ** 31:
        raise ValueError("padded vertex count must stay below 2^31 (int32 vertex ids on the device)")
    ep = pad_of[e] if E else e
    deg_pad = np.zeros(n_pad, np.int64)
    deg_pad[pad_of] = deg
    up_pad = np.zeros(n_pad, np.int64)
    up_pad[pad_of] = up
    row_ptr = np.concatenate([[0], np.cumsum(deg_pad)]).astype(np.int64)
    up_ptr = np.concatenate([[0], np.cumsum(up_pad)]).astype(np.int64)
    perm = np.argsort(pad_of, kind="stable")            # original vertex of the valid padded rows, in padded order
    monotonic = bool(np.all(perm == np.arange(n)))
    edge_orig = None
    if E:
        # symmetric CSR in ORIGINAL ids (scipy counting sort + per-row index sort), rows then permuted into padded
        # order and the column ids mapped: every row keeps its entries ordered by original neighbour id
        src = np.concatenate([e[:, 0], e[:, 1]])
        dst = np.concatenate([e[:, 1], e[:, 0]])
        sym = sp.csr_matrix((np.ones(2 * E, dtype=np.int8), (src, dst)), shape=(n, n))
        sym.sort_indices()
        assert sym.nnz == 2 * E, "duplicate edges in the edge list"
        if not monotonic:
            sym = sym[perm]
        col = pad_of[sym.indices].astype(np.int32)
        if not monotonic:
            # local-order edge numbering: rows in padded order, each contributing its owned edges in edge-list order
            up_cum = np.concatenate([[0], np.cumsum(up)])
            cnt = up[perm]
            starts = up_cum[perm]
            first = np.concatenate([[0], np.cumsum(cnt)])[:-1]
            edge_orig = np.repeat(starts - first, cnt) + np.arange(E, dtype=np.int64)
    else:
        col = np.zeros(0, np.int32)
    blk = np.arange(world, dtype=np.int64) * slice_rows
    e_lo = up_ptr[blk].astype(np.int64)
    e_hi = up_ptr[blk + rank_count].astype(np.int64)
    hubs = []
    for r in range(world):
        rows = np.arange(blk[r], blk[r] + rank_count[r])
        hubs.append(rows[deg_pad[rows] > hub_degree].astype(np.int32))
    return GraphLayout(n=n, n_edges=E, world=world, slice=slice_rows, n_pad=n_pad, rank_count=rank_count, e_lo=e_lo,
                       e_hi=e_hi, pad_of=pad_of, edges32=ep.astype(np.int32).reshape(-1, 2), row_ptr=row_ptr, col=col,
                       up_ptr=up_ptr, edge_orig=edge_orig, hubs=hubs, sorted_edges=is_sorted, ownership=ownership,
                       v_lo=v_lo, v_hi=v_hi)
