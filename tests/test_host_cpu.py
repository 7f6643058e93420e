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


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_cuda():
    adj = gr.generate_random_regular(50, 4, seed=0)
    with pytest.raises(RuntimeError, match="CUDA"):
        gr.GraphEmbedderPyTorch(adj, n_components=2, verbose=False)
    with pytest.raises(RuntimeError):
        gr.create_graphem(adj, n_components=2, backend="cpu")
    with pytest.raises(RuntimeError):
        gr.GraphEmbedderPyTorch(adj, n_components=2, device="cpu", verbose=False)
