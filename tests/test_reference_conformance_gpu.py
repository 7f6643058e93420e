"""API conformance on the GPU: the scenarios the reference's own suite runs against GraphEmbedderPyTorch
(/root/reference/tests/test_pytorch_backend.py, test_embedder.py, test_integration.py -- SURVEY.md section 4; the
reference cannot travel to the GPU box, so the scenarios are restated here, not imported) executed against the B200
class: constructor contract, attributes, tiny / disconnected / empty graphs, repeated run_layout calls, dimensions,
reproducibility up to axis reflections, backend names.  Documented divergences are asserted as such: device='cpu' and
non-fp32 dtypes raise (there is no CPU or reduced-precision path), PyKeOps is reported absent."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

pytestmark = pytest.mark.gpu


def _gr():
    import graphem_rapids_b200 as gr
    return gr


TWO_TRIANGLES = np.array([[0, 1, 1, 0, 0, 0], [1, 0, 1, 0, 0, 0], [1, 1, 0, 0, 0, 0],
                          [0, 0, 0, 0, 1, 1], [0, 0, 0, 1, 0, 1], [0, 0, 0, 1, 1, 0]])
KW = dict(L_min=10.0, k_attr=0.5, k_inter=0.1, verbose=False)


def test_initialization_attributes_and_types():           # test_pytorch_backend.py:21-42, test_embedder.py:13-31
    gr = _gr()
    adj = gr.generate_random_regular(n=50, d=4, seed=42)
    emb = gr.GraphEmbedderPyTorch(adjacency=adj, n_components=3, n_neighbors=15, sample_size=256, batch_size=1024, **KW)
    assert emb.n == 50 and emb.n_components == 3 and emb.n_neighbors == 15 and emb.batch_size == 1024
    assert emb.L_min == 10.0 and emb.k_attr == 0.5 and emb.k_inter == 0.1 and emb.dtype == torch.float32
    assert isinstance(emb.device, torch.device) and emb.device.type == "cuda"
    assert emb.sample_size == min(256, emb.n_edges) and emb.memory_efficient is True
    assert sp.issparse(emb.adjacency) and emb.adjacency.shape == (50, 50)
    assert emb.edges.dtype == torch.long and emb.edges.shape == (emb.n_edges, 2) and emb.edges.is_cuda
    assert bool((emb.edges[:, 0] < emb.edges[:, 1]).all())
    assert isinstance(emb.positions, np.ndarray) and emb.positions.shape == (50, 3) and emb.positions.dtype == np.float32
    assert isinstance(emb._positions, torch.Tensor) and emb._positions.shape == (50, 3)
    assert emb._has_pykeops is False and "GraphEmbedderPyTorch(n_vertices=50" in repr(emb)


def test_cuda_device_argument_forms():                     # test_pytorch_backend.py:66-84
    gr = _gr()
    adj = gr.generate_random_regular(n=40, d=4, seed=1)
    for dev in ("cuda", "cuda:0", torch.device("cuda:0"), None):
        emb = gr.GraphEmbedderPyTorch(adj, n_components=2, device=dev, **KW)
        assert emb.device == torch.device("cuda:0") and emb._positions.is_cuda
        emb.close()
    with pytest.raises(RuntimeError):
        gr.GraphEmbedderPyTorch(adj, n_components=2, device="invalid_device", **KW)


def test_documented_divergences_raise_loudly():            # test_pytorch_backend.py:44-64, :151-182, :497-523
    gr = _gr()
    adj = gr.generate_random_regular(n=40, d=4, seed=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        gr.GraphEmbedderPyTorch(adj, n_components=2, device="cpu", **KW)
    for dt in (torch.float64, torch.float16):
        with pytest.raises(NotImplementedError, match="float32"):
            gr.GraphEmbedderPyTorch(adj, n_components=2, dtype=dt, **KW)
    emb = gr.GraphEmbedderPyTorch(adj, n_components=2, **KW)
    assert emb._check_pykeops_availability() is False
    with pytest.raises(ImportError):
        emb._compute_knn_pykeops(torch.zeros(4, 2), torch.zeros(8, 2), 2, 4)
    assert 1 <= emb._get_adaptive_chunk_size(100, 1000, "torch") <= 100


@pytest.mark.parametrize("d", [2, 3, 4])
def test_dimensions(d):                                    # test_pytorch_backend.py:86-104, test_integration.py:113-140
    gr = _gr()
    adj = gr.generate_random_regular(n=60, d=4, seed=42)
    emb = gr.GraphEmbedderPyTorch(adj, n_components=d, n_neighbors=10, sample_size=128, **KW)
    out = emb.run_layout(num_iterations=3)
    assert out.shape == (60, d) and np.all(np.isfinite(out)) and np.array_equal(out, emb.get_positions())


def test_disconnected_tiny_graph():                        # test_pytorch_backend.py:184-208, test_embedder.py:74-98
    gr = _gr()
    emb = gr.GraphEmbedderPyTorch(adjacency=TWO_TRIANGLES, n_components=2, n_neighbors=5, sample_size=6, **KW)
    assert emb.n_edges == 6
    emb.run_layout(num_iterations=2)
    assert emb.positions.shape == (6, 2) and np.all(np.isfinite(emb.positions))


def test_two_hexagons_and_k4():                            # test_integration.py:274-312, conftest.py:25-28
    gr = _gr()
    hexa = np.zeros((12, 12), dtype=int)
    for base in (0, 6):
        for i in range(6):
            a, b = base + i, base + (i + 1) % 6
            hexa[a, b] = hexa[b, a] = 1
    emb = gr.GraphEmbedderPyTorch(hexa, n_components=2, n_neighbors=5, sample_size=12, **KW)
    emb.run_layout(num_iterations=5)
    assert emb.positions.shape == (12, 2) and np.all(np.isfinite(emb.positions))
    k4 = np.ones((4, 4), dtype=int) - np.eye(4, dtype=int)
    emb = gr.GraphEmbedderPyTorch(k4, n_components=2, n_neighbors=3, sample_size=6, **KW)
    emb.run_layout(num_iterations=3)
    assert emb.positions.shape == (4, 2) and np.all(np.isfinite(emb.positions))


def test_layout_stability_over_repeated_calls():           # test_pytorch_backend.py:212-234, test_embedder.py:100-121
    gr = _gr()
    adj = gr.generate_random_regular(n=30, d=4, seed=42)
    emb = gr.GraphEmbedderPyTorch(adj, n_components=2, n_neighbors=15, sample_size=64, **KW)
    for _ in range(3):
        emb.run_layout(num_iterations=2)
        assert np.all(np.isfinite(emb.positions)) and np.max(np.abs(emb.positions)) < 1000


def test_sample_size_larger_than_edge_count():             # test_pytorch_backend.py:236-257
    gr = _gr()
    adj = gr.erdos_renyi_graph(n=200, p=0.02, seed=42)
    emb = gr.GraphEmbedderPyTorch(adj, n_components=2, n_neighbors=15, sample_size=512, **KW)
    assert emb.sample_size == emb.n_edges < 512 and emb.positions.shape == (200, 2)
    emb.run_layout(num_iterations=2)
    assert np.all(np.isfinite(emb.positions))


def test_parameter_validation():                           # test_pytorch_backend.py:259-289, test_integration.py:352-366,:386-
    gr = _gr()
    adj = gr.generate_random_regular(n=20, d=4, seed=42)
    with pytest.raises(ValueError):
        gr.GraphEmbedderPyTorch(adj, n_components=0, verbose=False)
    with pytest.raises(ValueError):
        gr.GraphEmbedderPyTorch(adj, n_components=2, k_attr=-0.5, verbose=False)
    with pytest.raises(ValueError):
        gr.GraphEmbedderPyTorch(np.zeros((3, 4)), n_components=2, verbose=False)
    with pytest.raises(ValueError):
        gr.GraphEmbedderPyTorch(np.zeros((0, 0)), n_components=2, verbose=False)
    emb = gr.GraphEmbedderPyTorch(adj, n_components=2, L_min=0.0, k_attr=0.0, k_inter=0.0, n_neighbors=1, sample_size=1,
                                  verbose=False)
    emb.run_layout(num_iterations=2)
    assert np.all(np.isfinite(emb.positions))


def test_empty_graph_raises_runtime_error_in_layout():     # test_integration.py:368-382
    gr = _gr()
    emb = gr.GraphEmbedderPyTorch(sp.csr_matrix((10, 10)), n_components=2, verbose=False)
    assert emb.n_edges == 0 and emb.positions.shape == (10, 2)
    with pytest.raises(RuntimeError):
        emb.run_layout(num_iterations=2)


def test_batch_size_is_accepted_and_irrelevant():          # test_pytorch_backend.py:291-325
    gr = _gr()
    adj = gr.generate_random_regular(n=100, d=6, seed=42)
    outs = []
    for bs in (None, 16, 4096):
        emb = gr.GraphEmbedderPyTorch(adj, n_components=2, n_neighbors=10, sample_size=128, batch_size=bs, seed=7,
                                      initial_positions=np.random.default_rng(0).standard_normal((100, 2)).astype(np.float32), **KW)
        outs.append(emb.run_layout(num_iterations=3))
    # (identical up to the order in which the atomics add the few intersection terms that hit one vertex)
    assert np.allclose(outs[0], outs[1], rtol=1e-5, atol=1e-5) and np.allclose(outs[0], outs[2], rtol=1e-5, atol=1e-5)


def test_reproducibility_up_to_axis_reflections():         # test_pytorch_backend.py:327-379, test_integration.py:215-249
    gr = _gr()
    adj = gr.generate_random_regular(n=50, d=4, seed=42)
    res = []
    for _ in range(2):
        torch.manual_seed(123)
        emb = gr.GraphEmbedderPyTorch(adj, n_components=2, n_neighbors=15, sample_size=256, **KW)
        res.append(emb.run_layout(num_iterations=3))
    ok = any(np.allclose(res[0], res[1] * np.array(sg), rtol=1e-6, atol=1e-6) for sg in ([1, 1], [-1, 1], [1, -1], [-1, -1]))
    assert ok
    res = []
    for _ in range(2):                                     # explicit seed argument (test_integration.py:215-249)
        emb = gr.GraphEmbedderPyTorch(adj, n_components=3, n_neighbors=10, sample_size=128, seed=42, **KW)
        res.append(emb.run_layout(num_iterations=5))
    assert min(np.mean(np.abs(res[0] - res[1] * np.array(sg))) for sg in
               ([1, 1, 1], [-1, 1, 1], [1, -1, 1], [1, 1, -1], [-1, -1, 1], [-1, 1, -1], [1, -1, -1], [-1, -1, -1])) < 1e-2


def test_knn_private_api_shape_and_range():                # test_pytorch_backend.py:465-472, :525-560
    gr = _gr()
    adj = gr.generate_random_regular(n=50, d=4, seed=42)
    emb = gr.GraphEmbedderPyTorch(adj, n_components=3, **KW)
    q = torch.randn(40, 3, device="cuda")
    ref = torch.randn(300, 3, device="cuda")
    for fn in (lambda: emb._compute_knn_chunked(q, ref, 5), lambda: emb._compute_knn_torch(q, ref, 5, 16)):
        idx = fn()
        assert idx.shape == (40, 5) and idx.dtype == torch.long and int(idx.min()) >= 0 and int(idx.max()) < 300
        want = torch.cdist(q, ref).topk(5, largest=False).indices
        assert (idx.sort(dim=1).values == want.sort(dim=1).values).float().mean() > 0.99     # the reference's neighbour sets


def test_create_graphem_backend_names():                   # graphem_rapids/__init__.py:78-136, test_integration.py:319-345
    gr = _gr()
    adj = gr.generate_random_regular(n=40, d=4, seed=1)
    for backend in (None, "auto", "pytorch", "cuvs"):
        emb = gr.create_graphem(adj, n_components=2, backend=backend, **KW)
        assert type(emb).__name__ == "GraphEmbedderPyTorch" and emb.run_layout(2).shape == (40, 2)
    with pytest.raises(ValueError):
        gr.create_graphem(adj, backend="no_such_backend")
    with pytest.raises(RuntimeError):
        gr.create_graphem(adj, backend="cpu")
