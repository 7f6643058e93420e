"""Shared helpers for the parity tests."""
import os

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def adjacency_from_edges(edges: np.ndarray, n: int):
    e = np.asarray(edges, dtype=np.int64)
    rows = np.concatenate([e[:, 0], e[:, 1]])
    cols = np.concatenate([e[:, 1], e[:, 0]])
    return sp.csr_matrix((np.ones(len(rows), dtype=np.int64), (rows, cols)), shape=(n, n))


def rel_inf(a, b) -> float:
    """||a-b||_inf / ||b||_inf  (per-array relative error)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    if den == 0:
        return float(np.abs(a).max())
    return float(np.abs(a - b).max() / den)


def golden_params(g):
    return dict(n_components=int(g["d"]), n_neighbors=int(g["n_neighbors"]), sample_size=int(g["sample_size"]),
                k_attr=float(g["k_attr"]), L_min=float(g["L_min"]), k_inter=float(g["k_inter"]))


def make_embedder(g, **extra):
    import graphem_rapids_b200 as gr
    adj = adjacency_from_edges(g["edges"], int(g["n"]))
    emb = gr.GraphEmbedderPyTorch(adj, device="cuda:0", verbose=False, seed=0, initial_positions=g["pos0"],
                                  **golden_params(g), **extra)
    assert np.array_equal(emb.edges.cpu().numpy(), g["edges"].astype(np.int64))
    return emb


def rows_match_modulo_ties(idx_a, dist_a, idx_b, dist_b, ulps=1):
    """Rows equal as sets, or differing only in elements within `ulps` of the boundary distance."""
    bad = []
    for r in range(idx_a.shape[0]):
        sa, sb = set(idx_a[r].tolist()), set(idx_b[r].tolist())
        if sa == sb:
            continue
        bound = max(dist_a[r].max(), dist_b[r].max())
        tol = ulps * np.spacing(np.float32(bound))
        da = dict(zip(idx_a[r].tolist(), dist_a[r].tolist()))
        db = dict(zip(idx_b[r].tolist(), dist_b[r].tolist()))
        for i in sa ^ sb:
            dd = da.get(i, db.get(i))
            if abs(dd - bound) > tol:
                bad.append((r, i, dd, bound))
    return bad


def spearman(a, b) -> float:
    from scipy.stats import spearmanr
    return float(spearmanr(a, b).correlation)
