"""CPU: host-side logic that needs no device -- generators, facade, seed selection, loud failure
without CUDA."""
import numpy as np
import pytest
import torch

import graphem_rapids_b200 as gr


def _check_adj(a, n):
    assert a.shape == (n, n)
    assert (a != a.T).nnz == 0 and a.diagonal().sum() == 0
    assert set(np.unique(a.data)) <= {1}


def test_generators_are_simple_symmetric_graphs():
    a = gr.erdos_renyi_graph(1000, 0.01, seed=0)
    _check_adj(a, 1000)
    assert abs(a.nnz // 2 - 4995) < 400
    a = gr.generate_ba(5000, 4, seed=0)
    _check_adj(a, 5000)
    deg = np.asarray(a.sum(1)).ravel()
    assert abs(a.nnz // 2 - 4 * (5000 - 4)) < 200 and deg.max() > 50 and deg.min() >= 1
    a = gr.generate_random_regular(2000, 8, seed=0)
    _check_adj(a, 2000)
    deg = np.asarray(a.sum(1)).ravel()
    assert deg.max() == 8 and deg.min() >= 6
    a, lab = gr.generate_sbm(100, 4, 0.1, 0.005, labels=True, seed=0)
    _check_adj(a, 400)
    assert lab.shape == (400,)
    r, c = a.nonzero()
    assert (lab[r] == lab[c]).mean() > 0.7
    # determinism in the seed
    assert (gr.generate_ba(500, 3, seed=5) != gr.generate_ba(500, 3, seed=5)).nnz == 0
    assert (gr.generate_ba(500, 3, seed=5) != gr.generate_ba(500, 3, seed=6)).nnz > 0


def test_edge_extraction_order_matches_reference_contract():
    """edges are the upper triangle in CSR nonzero() order: sorted by (i, j), i < j."""
    a = gr.generate_ba(300, 3, seed=1)
    rows, cols = a.nonzero()
    keep = rows < cols
    e = np.column_stack([rows[keep], cols[keep]])
    assert np.all(e[:, 0] < e[:, 1])
    key = e[:, 0] * 300 + e[:, 1]
    assert np.all(np.diff(key) > 0)


class _FakeEmbedder:
    def __init__(self, pos):
        self.positions = pos
        self.ran = None

    def run_layout(self, num_iterations=0):
        self.ran = num_iterations
        return self.positions


def test_seed_selection_semantics():
    pos = np.array([[0, 0], [3, 4], [1, 0], [0, -2], [6, 8]], dtype=np.float32)
    emb = _FakeEmbedder(pos)
    seeds = gr.graphem_seed_selection(emb, 3, num_iterations=7)
    assert emb.ran == 7 and seeds == [4, 1, 3] and all(isinstance(s, int) for s in seeds)


def test_seed_selection_reference_order_is_numpy_argsort_with_id_tie_break():
    from graphem_rapids_b200.influence import seed_selection_reference
    rng = np.random.default_rng(0)
    pos = rng.standard_normal((5000, 3)).astype(np.float32)
    r = np.linalg.norm(pos, axis=1)
    assert seed_selection_reference(pos, 40) == np.argsort(-r)[:40].tolist()          # distinct radii: numpy's list
    pos[[7, 3, 11]] = np.float32(9.0)                                                 # exact ties -> ascending id
    assert seed_selection_reference(pos, 4)[:3] == [3, 7, 11]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_cuda():
    adj = gr.generate_random_regular(50, 4, seed=0)
    with pytest.raises(RuntimeError, match="CUDA"):
        gr.GraphEmbedderPyTorch(adj, n_components=2, verbose=False)
    with pytest.raises(RuntimeError):
        gr.create_graphem(adj, n_components=2, backend="cpu")
    with pytest.raises(RuntimeError):
        gr.GraphEmbedderPyTorch(adj, n_components=2, device="cpu", verbose=False)


# ----------------------------------------------------------------------------- device graph build, specification
def _model_graph_count(indptr, indices, n):
    """numpy model of graph_count_kernel + the three scan launches (tile = 1024) of gem_graph_count."""
    NOT_SYM, NOT_CANON = 1, 2
    row_cnt = np.zeros(n, np.int64)
    up_cnt = np.zeros(n, np.int64)
    flags = 0
    for v in range(n):
        prev = -1
        for t in range(indptr[v], indptr[v + 1]):
            c = int(indices[t])
            if c <= prev:
                flags |= NOT_CANON
            prev = c
            if c < 0 or c >= n:
                flags |= NOT_CANON
                continue
            if c == v:
                continue
            row_cnt[v] += 1
            up_cnt[v] += c > v
            lo, hi = int(indptr[c]), int(indptr[c + 1])
            while lo < hi:
                m = (lo + hi) >> 1
                if indices[m] < v:
                    lo = m + 1
                else:
                    hi = m
            if lo >= indptr[c + 1] or indices[lo] != v:
                flags |= NOT_SYM

    def scan(a, tile=1024, threads=256):
        nt = (n + tile - 1) // tile
        sums = np.array([a[i * tile:(i + 1) * tile].sum() for i in range(nt)], np.int64)
        offs = np.zeros(nt, np.int64)
        carry = 0
        for b0 in range(0, nt, threads):                   # scan_tile_offsets_kernel: chunks of 256 with a carry
            chunk = sums[b0:b0 + threads]
            incl = np.cumsum(chunk)
            offs[b0:b0 + threads] = carry + incl - chunk
            carry += chunk.sum()
        out = np.empty(n + 1, np.int64)
        out[0] = 0
        for i in range(nt):
            out[1 + i * tile:1 + min(n, (i + 1) * tile)] = offs[i] + np.cumsum(a[i * tile:(i + 1) * tile])
        return out
    return scan(row_cnt), scan(up_cnt), flags


def _model_graph_fill(indptr, indices, n, row_ptr, up_ptr):
    col = np.full(int(row_ptr[n]), -1, np.int32)
    edges = np.full((int(up_ptr[n]), 2), -1, np.int32)
    for v in range(n):
        o, u = int(row_ptr[v]), int(up_ptr[v])
        for t in range(indptr[v], indptr[v + 1]):
            c = int(indices[t])
            if c == v:
                continue
            col[o] = c
            o += 1
            if c > v:
                edges[u] = (v, c)
                u += 1
    return col, edges


@pytest.mark.parametrize("name", ["ba", "er_loops", "rr"])
def test_device_graph_build_specification_equals_host_layout(name):
    """The two-pass construction the library runs on the device (gem_graph_count / gem_graph_fill), modelled in
    numpy, produces exactly partition.build_layout's arrays and the reference's edge list for a canonical
    symmetric adjacency -- with self loops, isolated vertices and non-unit data."""
    import scipy.sparse as sp
    from graphem_rapids_b200.partition import build_layout
    if name == "ba":
        a = gr.generate_ba(2500, 3, seed=2).tocsr()
    elif name == "rr":
        a = gr.generate_random_regular(1025, 4, seed=3).tocsr()
    else:
        a = gr.erdos_renyi_graph(1500, 0.002, seed=4).tolil()          # isolated vertices
        for v in (0, 7, 1499):
            a[v, v] = 3                                                # self loops: dropped by rows < cols
        a = a.tocsr().astype(np.float32)
        a.data *= 2.5
    a.sort_indices()
    assert a.has_canonical_format
    n = a.shape[0]
    indptr, indices = a.indptr.astype(np.int64), a.indices.astype(np.int32)
    row_ptr, up_ptr, flags = _model_graph_count(indptr, indices, n)
    assert flags == 0
    col, edges = _model_graph_fill(indptr, indices, n, row_ptr, up_ptr)
    r, c = a.nonzero()
    keep = r < c
    ref_edges = np.column_stack([r[keep], c[keep]])
    L = build_layout(ref_edges, n, 1)
    assert np.array_equal(edges, ref_edges) and np.array_equal(edges, L.edges32)
    assert np.array_equal(row_ptr, L.row_ptr) and np.array_equal(up_ptr, L.up_ptr) and np.array_equal(col, L.col)


def test_device_graph_build_specification_flags():
    import scipy.sparse as sp
    a = sp.triu(gr.generate_random_regular(200, 4, seed=5)).tocsr()   # upper triangle only: pattern not symmetric
    a.sort_indices()
    assert _model_graph_count(a.indptr.astype(np.int64), a.indices, 200)[2] == 1
    b = gr.generate_random_regular(200, 4, seed=5).tocsr()
    b.sort_indices()
    idx = b.indices.copy()
    idx[b.indptr[3]:b.indptr[4]] = idx[b.indptr[3]:b.indptr[4]][::-1]   # one row descending
    assert _model_graph_count(b.indptr.astype(np.int64), idx, 200)[2] & 2


# ----------------------------------------------------------------------------- correlation harness (torch part)
def test_rank_average_and_spearman_match_scipy():
    from scipy.stats import rankdata, spearmanr
    from graphem_rapids_b200.correlation import rank_average, spearman
    rng = np.random.default_rng(0)
    for n, levels in ((1, 1), (7, 3), (1000, 12), (5000, 100000)):
        a = rng.integers(0, levels, n).astype(np.float32)                  # heavy ties (degree-like)
        b = (a + rng.standard_normal(n) * 2).astype(np.float32)
        assert np.array_equal(rank_average(torch.from_numpy(a)).numpy(), rankdata(a, method="average"))
        if n > 1 and levels > 1:
            assert abs(spearman(torch.from_numpy(a), torch.from_numpy(b)) - spearmanr(a, b).correlation) < 1e-12
    assert np.isnan(spearman(torch.zeros(5), torch.arange(5.0)))
    assert rank_average(torch.empty(0)).numel() == 0


def test_pagerank_power_iteration_matches_networkx():
    """The iteration pagerank_device runs (with the normalised-adjacency operator supplied by scipy here, by the
    CUDA library on the device) reproduces networkx.pagerank, dangling (isolated) vertices included."""
    import networkx as nx
    import scipy.sparse as sp
    from graphem_rapids_b200.correlation import _pagerank_power_iteration
    a = gr.generate_ba(800, 3, seed=3).tolil()
    a.resize((803, 803))                                                    # three isolated vertices
    a = a.tocsr().astype(np.float64)
    deg = np.asarray(a.sum(1)).ravel()
    dinv = np.where(deg > 0, 1.0 / np.sqrt(np.maximum(deg, 1)), 0.0)
    m = sp.diags(dinv) @ a @ sp.diags(dinv)
    pr = _pagerank_power_iteration(torch.from_numpy(deg), lambda z: torch.from_numpy((m @ z.numpy().astype(np.float64))
                                                                                     .astype(np.float32)),
                                   0.85, 1e-8, 200).numpy()
    ref = nx.pagerank(nx.from_scipy_sparse_array(a), alpha=0.85, tol=1e-10, max_iter=500)
    ref = np.array([ref[i] for i in range(803)])
    assert abs(pr.sum() - 1) < 1e-5 and np.abs(pr - ref).max() / ref.max() < 1e-4
