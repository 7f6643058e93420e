"""
GraphEmbedderPyTorch -- drop-in host class for the reference's
graphem_rapids/backends/embedder_pytorch.py::GraphEmbedderPyTorch, with the layout iteration
(`update_positions`, :776-806) executed by hand-written sm_100a kernels through the C ABI of
include/graphem_b200.h.  Same constructor signature, attributes, methods and exceptions.

Deliberate differences (DESIGN.md "Boundary"):
  * CUDA fp32 only.  device='cpu', a non-fp32 dtype or a missing extension raise -- there is no
    CPU / torch-op fallback on this path (the reference's backend dispatch, PyKeOps branch and
    MemoryManager are removed).
  * the S query edges are drawn on the device by a keyed bijection (gem_sample_edges) instead of
    torch.randperm(E)[:S]; `sampler='torch'` restores the reference's call and RNG stream.
  * extra keyword-only arguments: initial_positions, sampler, use_cuda_graph, init_method
    ('auto' | 'arpack' | 'device': the Laplacian initial embedding on the host like the reference, or by
    a device subspace iteration -- default for graphs of 20 000+ vertices), graph_build ('auto' | 'host' |
    'device': edge list + symmetric CSR extracted from the adjacency by numpy/scipy on the host, or by the
    library on the device -- default for canonical symmetric adjacencies of 20 000+ vertices).
"""
from __future__ import annotations

import ctypes
import logging
import weakref
from typing import Optional

import numpy as np
import scipy.sparse as sp
import torch

from . import _cabi
from .partition import GraphLayout, build_layout

logger = logging.getLogger(__name__)

_SUPPORTED_BACKENDS = (None, "auto", "pytorch", "cuda", "cuvs", "b200")


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class _nvtx_range:
    """NVTX range around the host-side phases (construction, graph build, initial embedding, run_layout, captures) when
    GEM_NVTX=1 -- the reference has no tracing hooks at all; the kernels themselves show up by
    name in Nsight Systems / Compute."""
    _on = None

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _nvtx_range._on is None:
            import os
            _nvtx_range._on = os.environ.get("GEM_NVTX", "0") == "1"
        if _nvtx_range._on:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if _nvtx_range._on:
            torch.cuda.nvtx.range_pop()
        return False


def _acquire_coef_slot(lib, device_index: int) -> int:
    slot = ctypes.c_int(-1)
    with torch.cuda.device(device_index):
        rc = lib.gem_coef_slot_acquire(ctypes.byref(slot))
        if rc != 0:                     # slots of unreachable embedders are released by their finalizers
            import gc
            gc.collect()
            rc = lib.gem_coef_slot_acquire(ctypes.byref(slot))
    if rc != 0:
        raise RuntimeError(f"graphem_rapids_b200: {_cabi.error_string(rc)} -- at most {lib.gem_coef_slots()} embedders "
                           f"can be alive per GPU; call close() on (or delete) the ones no longer needed")
    return int(slot.value)


def _release_coef_slot(lib, device_index: int, slot: int) -> None:
    try:
        with torch.cuda.device(device_index):
            lib.gem_coef_slot_release(int(slot))
    except Exception:  # interpreter shutdown  # pylint: disable=broad-exception-caught
        pass


class GraphEmbedderPyTorch:
    """Force-directed graph embedder; API of embedder_pytorch.py:27-180 (see module docstring)."""

    def __init__(self, adjacency, n_components=2, device=None, dtype=torch.float32, L_min=1.0,
                 k_attr=0.2, k_inter=0.5, n_neighbors=10, sample_size=256, batch_size=None,
                 memory_efficient=True, verbose=True, logger_instance=None, seed=None, *,
                 initial_positions=None, sampler="device", use_cuda_graph=True, init_method="auto",
                 graph_build="auto"):
        # seeding contract (embedder_pytorch.py:106-111)
        if seed is not None:
            np.random.seed(seed)
            torch.manual_seed(seed)
            if torch.cuda.is_available():
                torch.cuda.manual_seed(seed)
                torch.cuda.manual_seed_all(seed)

        # device (:114-117).  torch.device('invalid_device') raises RuntimeError like the reference.
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("graphem_rapids_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
            self.device = torch.device("cuda", torch.cuda.current_device())
        else:
            self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"graphem_rapids_b200 runs on CUDA only (got device={self.device}); "
                               "the reference's CPU backend is not part of this path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if dtype != torch.float32:
            raise NotImplementedError(f"graphem_rapids_b200 computes in float32 only (got {dtype})")

        if logger_instance is not None:
            self.logger = logger_instance
        else:
            self.logger = logger
            if verbose:
                logging.basicConfig(level=logging.INFO)

        adjacency = self._validate_adjacency(adjacency)
        self.adjacency = adjacency
        self.n = adjacency.shape[0]
        self.n_components = n_components
        self.dtype = dtype
        self.L_min = L_min
        self.k_attr = k_attr
        self.k_inter = k_inter
        self.n_neighbors = n_neighbors
        self.memory_efficient = memory_efficient
        self.batch_size = batch_size
        if n_components <= 0:                                   # :143-146
            raise ValueError(f"Number of components must be positive, got {n_components}")
        if k_attr < 0:
            raise ValueError(f"Attractive force constant k_attr must be non-negative, got {k_attr}")
        self.verbose = verbose
        if sampler not in ("device", "torch"):
            raise ValueError("sampler must be 'device' or 'torch'")
        self.sampler = sampler
        self.use_cuda_graph = bool(use_cuda_graph)
        if init_method not in ("auto", "arpack", "device"):
            raise ValueError("init_method must be 'auto', 'arpack' or 'device'")
        self.init_method = init_method
        if graph_build not in ("auto", "host", "device"):
            raise ValueError("graph_build must be 'auto', 'host' or 'device'")
        self.graph_build = graph_build

        _cabi.load()
        _cabi.init_device(self.device.index)
        self._lib = _cabi.load()
        self._ld = self._lib.gem_row_pitch(int(n_components))
        self._mld = self._lib.gem_mid_pitch(int(n_components))
        # constant-bank coefficient slot of this object's KNN scans (include/graphem_b200.h "Coefficient slots"):
        # owned for the object's lifetime, so embedders on different streams of one GPU never share filter state
        self._coef_slot = _acquire_coef_slot(self._lib, self.device.index)
        self._slot_finalizer = weakref.finalize(self, _release_coef_slot, self._lib, self.device.index, self._coef_slot)

        if self.n >= 2 ** 31:
            raise ValueError("graphem_rapids_b200 stores edge endpoints as int32: n must be < 2^31")
        # device-side graph arrays: padded vertex numbering (identity on one GPU), int32 edge endpoints,
        # symmetric CSR + per-vertex upper-edge offsets for the pull kernels.  Built by the library on the
        # device when the adjacency allows it, else by partition.py on the host.
        self._world, self._rank = self._world_and_rank()
        with _nvtx_range("gem.graph_build"):
            built = self._graph_arrays_device(adjacency)
        if built is not None:
            self._layout, self.edges, self._edges32, self._row_ptr, self._col, self._up_ptr, self._hubs = built
            self.n_edges = int(self._edges32.shape[0])
            self.sample_size = min(sample_size, self.n_edges)       # :156
            self._pad_index = None
        else:
            edges = self._extract_edges_from_adjacency(adjacency)
            self.n_edges = len(edges)
            self.sample_size = min(sample_size, self.n_edges)       # :156
            self.edges = torch.tensor(edges, device=self.device, dtype=torch.long).reshape(-1, 2)   # :159
            self._layout = build_layout(np.asarray(edges, dtype=np.int64).reshape(-1, 2), self.n, self._world,
                                        hub_degree=int(self._lib.gem_hub_degree()),
                                        ownership=getattr(self, "_ownership", "strided"))
            L = self._layout
            self._edges32 = torch.from_numpy(L.edges32).to(self.device).contiguous()
            self._row_ptr = torch.from_numpy(L.row_ptr).to(self.device)
            self._col = torch.from_numpy(L.col).to(self.device)
            self._up_ptr = torch.from_numpy(L.up_ptr).to(self.device)
            self._hubs = torch.from_numpy(L.hubs[self._rank]).to(self.device)
            identity = L.n_pad == self.n and bool(np.array_equal(L.pad_of, np.arange(self.n)))
            self._pad_index = None if identity else torch.from_numpy(L.pad_of).to(self.device)

        self._has_pykeops = False                               # the PyKeOps branch (:247-258) is removed
        if self.batch_size is None:
            self.batch_size = max(1, min(self.n, 1024))         # kept for API compatibility; unused
        self._sampler_seed = int(seed) if seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
        self._bufs = {}
        self._graph = None
        self._graph_unrolled = None
        self._graph_key = None
        self._torch_sample_ready = False
        self.last_sampled_indices = None
        self.last_knn_indices = None

        if self.verbose:
            self.logger.info("Initialized GraphEmbedderPyTorch (B200 kernels) on %s", self.device)
            self.logger.info("Graph: %d vertices, %d edges, %dD", self.n, self.n_edges, self.n_components)

        self._pos = torch.zeros((self._layout.n_pad, self._ld), device=self.device, dtype=torch.float32)
        if initial_positions is not None:
            self.positions = initial_positions
        else:
            with _nvtx_range("gem.initial_embedding"):
                self._positions = self._compute_laplacian_embedding()

    # ------------------------------------------------------------------ input handling
    def _validate_adjacency(self, adjacency):
        """embedder_pytorch.py:182-218: any array-like -> CSR, square, non-empty."""
        if sp.issparse(adjacency):
            adjacency = adjacency.tocsr()
        elif not isinstance(adjacency, np.ndarray):
            adjacency = np.asarray(adjacency)
        if adjacency.shape[0] != adjacency.shape[1]:
            raise ValueError(f"Adjacency matrix must be square, got shape {adjacency.shape}")
        if adjacency.shape[0] == 0:
            raise ValueError("Adjacency matrix cannot be empty")
        if not sp.issparse(adjacency):
            adjacency = sp.csr_matrix(adjacency)
        return adjacency

    def _extract_edges_from_adjacency(self, adjacency):
        """embedder_pytorch.py:220-245: nonzero() order, upper triangle i<j."""
        rows, cols = adjacency.nonzero()
        keep = rows < cols
        edges = np.column_stack([rows[keep], cols[keep]])
        if self.verbose and len(edges) == 0:
            self.logger.warning("No edges found in adjacency matrix")
        return edges

    _DEVICE_GRAPH_MIN_N = 20000

    def _graph_arrays_device(self, adjacency):
        """Edge list + symmetric CSR built on the device from the CSR adjacency (gem_graph_count /
        gem_graph_fill; replaces the host work of embedder_pytorch.py:220-245 and partition.build_layout).
        Returns None when the host path has to be used: several ranks, graph_build='host', a small graph under
        'auto', or an adjacency that is not canonical (sorted rows, no duplicates, no stored zeros) with a
        symmetric pattern -- the reference's nonzero() order is then not "entries above the diagonal in storage
        order", or the symmetric CSR is not the adjacency minus its diagonal.  graph_build='device' raises
        instead of falling back."""
        mode = self.graph_build
        forced = mode == "device"

        def give_up(why):
            if forced:
                raise ValueError(f"graph_build='device' is not applicable: {why}")
            return None

        if mode == "host":
            return None
        if self._world != 1:
            return give_up("the vertex partition of a multi-GPU run is built on the host")
        if not forced and self.n < self._DEVICE_GRAPH_MIN_N:
            return None
        nnz = int(adjacency.nnz)
        if nnz == 0:
            return give_up("no stored entries")
        if not adjacency.has_canonical_format:
            return give_up("CSR rows are not sorted / contain duplicates")
        if int(np.count_nonzero(adjacency.data)) != nnz:
            return give_up("stored zeros (nonzero() drops them)")
        lib, dev, n = self._lib, self.device, self.n
        with torch.cuda.device(dev):
            indptr = torch.from_numpy(np.ascontiguousarray(adjacency.indptr, dtype=np.int64)).to(dev)
            indices = torch.from_numpy(np.ascontiguousarray(adjacency.indices, dtype=np.int32)).to(dev)
            row_ptr = torch.empty((n + 1,), device=dev, dtype=torch.long)
            up_ptr = torch.empty((n + 1,), device=dev, dtype=torch.long)
            flags = torch.zeros((1,), device=dev, dtype=torch.int32)
            nbytes = ctypes.c_size_t(0)
            _cabi.check(lib.gem_graph_workspace_bytes(n, ctypes.byref(nbytes)), "gem_graph_workspace_bytes")
            ws = torch.empty((nbytes.value // 8 + 1,), device=dev, dtype=torch.long)
            st = self._stream()
            _cabi.check(lib.gem_graph_count(_ptr(indptr), _ptr(indices), n, _ptr(row_ptr), _ptr(up_ptr), _ptr(flags),
                                            _ptr(ws), nbytes.value, st), "gem_graph_count")
            two_e, n_edges, bad = (int(x) for x in torch.stack([row_ptr[n], up_ptr[n], flags[0].to(torch.long)]).tolist())
            if bad:
                return give_up("adjacency pattern is not symmetric" if bad & 1 else "CSR rows are not strictly ascending")
            if n_edges == 0:
                return give_up("no edges")
            if two_e != 2 * n_edges:
                raise RuntimeError(f"gem_graph_count: inconsistent offsets ({two_e} entries for {n_edges} edges)")
            col = torch.empty((two_e,), device=dev, dtype=torch.int32)
            edges32 = torch.empty((n_edges, 2), device=dev, dtype=torch.int32)
            _cabi.check(lib.gem_graph_fill(_ptr(indptr), _ptr(indices), n, _ptr(row_ptr), _ptr(up_ptr), _ptr(col),
                                           _ptr(edges32), st), "gem_graph_fill")
            deg = row_ptr[1:] - row_ptr[:-1]
            hubs = torch.nonzero(deg > int(lib.gem_hub_degree())).reshape(-1).to(torch.int32)
            edges64 = edges32.to(torch.long)                        # public attribute (:159)
        layout = GraphLayout(n=n, n_edges=n_edges, world=1, slice=n, n_pad=n, rank_count=np.array([n], np.int64),
                             e_lo=np.array([0], np.int64), e_hi=np.array([n_edges], np.int64), pad_of=None,
                             edges32=None, row_ptr=None, col=None, up_ptr=None, edge_orig=None, hubs=[],
                             sorted_edges=True, ownership="contiguous", v_lo=np.array([0], np.int64),
                             v_hi=np.array([n], np.int64), on_device=True)
        return layout, edges64, edges32, row_ptr, col, up_ptr, hubs

    def _check_pykeops_availability(self):
        return False

    def _get_adaptive_chunk_size(self, n_query, n_ref, backend):
        """API compatibility (:260-322): the fused KNN never materialises a distance matrix, so
        there is nothing to chunk; returns a positive size bounded by n_query."""
        return max(1, min(int(self.batch_size), int(n_query)))

    # ------------------------------------------------------------------ state
    def _world_and_rank(self):
        """(world size, rank) of the vertex partition: (1, 0) here; ShardedGraphEmbedder overrides."""
        return 1, 0

    @property
    def _positions(self):
        """(n, d) positions on the device (reference attribute `_positions`): a view of the padded
        buffer on one GPU with ld == d, a gather of the valid rows otherwise."""
        if self._pad_index is None and self._ld == self.n_components:
            return self._pos
        if self._pad_index is None:
            return self._pos[:, : self.n_components]
        return self._rows_to_public()

    @_positions.setter
    def _positions(self, value):
        value = torch.as_tensor(value)
        if tuple(value.shape) != (self.n, self.n_components):
            raise ValueError(f"positions must have shape {(self.n, self.n_components)}, got {tuple(value.shape)}")
        self._public_to_rows(value)

    def _io_stage(self, name="io_stage"):
        """(n, d) fp32 device staging buffer of the positions setter / getter (allocated once)."""
        st = getattr(self, "_io_bufs", None)
        if st is None:
            st = self._io_bufs = {}
        if name not in st:
            st[name] = torch.empty((self.n, self.n_components), device=self.device, dtype=torch.float32)
        return st[name]

    def _pinned_stage(self):
        st = getattr(self, "_io_bufs", None)
        if st is None:
            st = self._io_bufs = {}
        if "pinned" not in st:
            st["pinned"] = torch.empty((self.n, self.n_components), dtype=torch.float32, pin_memory=True)
        return st["pinned"]

    def _replica_ptrs(self):
        """Device pointers of every replica of the position buffer the setter must fill (one here)."""
        return (ctypes.c_void_p * 1)(self._pos.data_ptr()), 1

    def _public_to_rows(self, value: torch.Tensor, row0: int = 0):
        """(cnt, d) rows [row0, row0+cnt) of the public array (host or device tensor) -> padded (n_pad, ld) rows of
        the position buffer: one H2D copy (asynchronous when the source is pinned) + gem_rows_scatter."""
        d = self.n_components
        cnt = int(value.shape[0])
        with torch.cuda.device(self.device):
            if value.device.type == "cuda":
                src = value.to(device=self.device, dtype=torch.float32).contiguous()
            else:
                v = value.to(torch.float32).contiguous()
                if self._ld == d and self._pad_index is None and row0 == 0 and cnt == self.n:
                    self._pos.copy_(v, non_blocking=True)            # layout already matches: straight into place
                    return
                src = self._io_stage()[row0: row0 + cnt]
                src.copy_(v, non_blocking=True)
            ptrs, world = self._replica_ptrs()
            _cabi.check(self._lib.gem_rows_scatter(_ptr(src), row0, cnt, d, _ptr(self._pad_index), ptrs, world,
                                                   self._stream()), "gem_rows_scatter")

    def _rows_to_public(self, row0: int = 0, cnt: Optional[int] = None, out: Optional[torch.Tensor] = None):
        """Rows [row0, row0+cnt) of the public (n, d) array as a device tensor (gem_rows_gather)."""
        cnt = self.n - row0 if cnt is None else int(cnt)
        if out is None:
            out = torch.empty((cnt, self.n_components), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.gem_rows_gather(_ptr(self._pos), row0, cnt, self.n_components, _ptr(self._pad_index),
                                                  _ptr(out), self._stream()), "gem_rows_gather")
        return out

    @property
    def positions(self):
        """numpy copy (:324-327): a fresh host array, like the reference's `.cpu().numpy()`."""
        if self.n * self.n_components < 65536:
            return self._positions.detach().cpu().numpy()
        # large layouts: gather -> pinned staging (full-speed D2H) -> fresh pageable array.  The D2H runs in chunks and
        # the host fills the fresh array (page faults included) chunk by chunk behind it: the two copies overlap instead
        # of adding up (12 MB at C3: ~0.25 ms each)
        pinned = self._pinned_stage()
        d = self.n_components
        with torch.cuda.device(self.device):
            if self._ld == d and self._pad_index is None:
                src = self._pos
            else:
                src = self._rows_to_public(out=self._io_stage("d2h_stage"))
            nchunk = 4 if self.n >= 4 * 65536 else 1
            bounds = [(self.n * c) // nchunk for c in range(nchunk + 1)]
            events = []
            for c in range(nchunk):
                lo, hi = bounds[c], bounds[c + 1]
                pinned[lo:hi].copy_(src[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                events.append(ev)
            out = torch.empty((self.n, d), dtype=torch.float32)
            for c in range(nchunk):
                lo, hi = bounds[c], bounds[c + 1]
                events[c].synchronize()
                out[lo:hi].copy_(pinned[lo:hi])
        return out.numpy()

    @positions.setter
    def positions(self, value):
        """ndarray or tensor -> device (:329-335)."""
        if isinstance(value, np.ndarray):
            value = torch.from_numpy(np.ascontiguousarray(value, dtype=np.float32))
        self._positions = value

    def _compute_laplacian_embedding(self):
        """Initial embedding (embedder_pytorch.py:337-379): eigenvectors 2..d+1 of the normalised
        Laplacian; random*0.1 if the solver fails.  Small graphs use ARPACK on the host exactly like the
        reference; from `_DEVICE_INIT_MIN_N` vertices on (ARPACK takes ~1 min at 200 K vertices) a
        Chebyshev-filtered subspace iteration runs on the device.  One-off,
        outside the timed path."""
        self.logger.info("Computing Laplacian embedding")
        method = self.init_method
        if method == "auto":
            method = "device" if (self.n >= self._DEVICE_INIT_MIN_N and self.n_components + 1 <= 6
                                  and self.n_edges > 0) else "arpack"
        try:
            if method == "device":
                emb = self._laplacian_embedding_device()
                return emb.to(device=self.device, dtype=torch.float32)
            emb = self._laplacian_embedding_arpack()
        except Exception as exc:  # pylint: disable=broad-exception-caught
            self.logger.warning("Eigendecomposition failed: %s", exc)
            emb = np.random.randn(self.n, self.n_components) * 0.1
        return torch.tensor(emb, device=self.device, dtype=torch.float32)

    _DEVICE_INIT_MIN_N = 20000

    def _laplacian_embedding_arpack(self):
        import scipy.sparse.linalg as spla
        from scipy.sparse.csgraph import laplacian
        sym = sp.csr_matrix(self.adjacency + self.adjacency.transpose())
        sym.data = np.ones_like(sym.data)
        lap = laplacian(sym, normed=True)
        k = self.n_components + 1
        _, vecs = spla.eigsh(lap, k, which="SM")
        return vecs[:, 1:k]

    def _laplacian_embedding_device(self, tol=2e-4, max_outer=80, degree=12, return_info=False):
        """Eigenvectors of M = D^-1/2 A D^-1/2 with the largest eigenvalues (= smallest of L = I - M) by
        Chebyshev-filtered subspace iteration on a block of 8 vectors; the operator is the library's pull SpMV
        (gem_spmv_normalized_adjacency) over the symmetric CSR the spring kernel uses, the small dense algebra
        (thin QR, 8x8 Rayleigh-Ritz) is torch.  Returns the (n, d) tensor of Ritz vectors 2..d+1."""
        lib, dev = self._lib, self.device
        m = lib.gem_spmv_cols()
        k = int(self.n_components) + 1
        if k > m - 2:
            raise ValueError("device initial embedding supports n_components <= 5")
        rows = self._layout.n_pad
        deg = (self._row_ptr[1:] - self._row_ptr[:-1]).to(torch.float32)
        dinv = torch.where(deg > 0, deg.clamp_min(1).rsqrt(), torch.zeros_like(deg)).contiguous()
        live = (deg > 0).to(torch.float32).unsqueeze(1)
        st = self._stream()

        def spmv(x, alpha=1.0, beta=0.0, z=None, gamma=0.0):
            y = torch.empty_like(x)
            _cabi.check(lib.gem_spmv_normalized_adjacency(_ptr(self._row_ptr), _ptr(self._col), _ptr(dinv), _ptr(x), _ptr(y),
                                                          rows, float(alpha), float(beta), _ptr(z), float(gamma), st),
                        "gem_spmv_normalized_adjacency")
            return y

        gen = torch.Generator(device=dev)
        gen.manual_seed(int(self._sampler_seed) & 0x7FFFFFFF)
        with torch.cuda.device(dev):
            x = torch.randn((rows, m), device=dev, dtype=torch.float32, generator=gen) * live
            x, _ = torch.linalg.qr(x)
            lower, cut = -1.0, 0.0
            theta = None
            info = {"outer": 0, "spmv": 0, "residual": float("inf")}
            for it in range(max_outer):
                c, e = 0.5 * (lower + cut), 0.5 * (cut - lower)
                y0 = x
                y1 = spmv(x, 1.0 / e, 0.0, None, -c / e)
                for _ in range(2, degree + 1):
                    y0, y1 = y1, spmv(y1, 2.0 / e, -1.0, y0, -2.0 * c / e)
                info["spmv"] += degree
                x, _ = torch.linalg.qr(y1.contiguous())
                mx = spmv(x.contiguous())
                info["spmv"] += 1
                t = x.T @ mx
                theta, s = torch.linalg.eigh(0.5 * (t + t.T))           # ascending
                x = (x @ s).contiguous()
                mx = mx @ s
                res = (mx - x * theta.unsqueeze(0)).norm(dim=0)
                worst = float(res[m - k:].max())                         # the k wanted ones are the last (largest theta)
                info.update(outer=it + 1, residual=worst)
                cut = float(theta[0])                                    # damp everything below the block's smallest Ritz value
                cut = min(max(cut, -0.9), 0.999)
                if worst < tol:
                    break
            order = torch.argsort(theta, descending=True)
            vecs = x[:, order][:, 1:k]                                   # skip the trivial eigenvector (theta = 1)
            if self._pad_index is not None:
                vecs = vecs[self._pad_index]
            info["theta"] = theta[order][:k].tolist()
        if self.verbose:
            self.logger.info("device Laplacian embedding: %d outer iterations, %d SpMV, residual %.2e", info["outer"],
                             info["spmv"], info["residual"])
        return (vecs.contiguous(), info) if return_info else vecs.contiguous()

    # ------------------------------------------------------------------ buffers / plan
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _buffers(self):
        """Scratch owned by the object, (re)allocated when S, k or E change."""
        S = min(int(self.sample_size), self.n_edges)
        kp1 = int(self.n_neighbors) + 1
        key = (S, kp1, self.n_edges)
        if self._bufs.get("key") == key:
            return self._bufs
        dev, f32 = self.device, torch.float32
        E = max(self.n_edges, 1)
        ws_bytes = ctypes.c_size_t(0)
        _cabi.check(self._lib.gem_knn_workspace_bytes(E, int(self.n_components), max(S, 1), kp1,
                                                      ctypes.byref(ws_bytes)), "gem_knn_workspace_bytes")
        st_bytes = ctypes.c_size_t(0)
        _cabi.check(self._lib.gem_update_workspace_bytes(self.n, int(self.n_components), ctypes.byref(st_bytes)),
                    "gem_update_workspace_bytes")
        self._bufs = dict(
            key=key, S=S, kp1=kp1,
            force=torch.zeros((self.n, self._ld), device=dev, dtype=f32),
            mid=torch.zeros((E + 1, self._mld), device=dev, dtype=f32),
            qmid=torch.zeros((max(S, 1), self._mld), device=dev, dtype=f32),
            tau_hint=torch.zeros((max(S, 1),), device=dev, dtype=f32),
            samp=torch.zeros((max(S, 1),), device=dev, dtype=torch.long),
            samp_next=torch.zeros((max(S, 1),), device=dev, dtype=torch.long),
            knn_idx=torch.zeros((max(S, 1), kp1), device=dev, dtype=torch.long),
            knn_dist=torch.zeros((max(S, 1), kp1), device=dev, dtype=f32),
            iter=self._bufs.get("iter", torch.zeros((1,), device=dev, dtype=torch.long)),
            knn_ws=torch.zeros((ws_bytes.value + 256,), device=dev, dtype=torch.uint8),
            stats_ws=torch.zeros((st_bytes.value + 256,), device=dev, dtype=torch.uint8),
            knn_ws_bytes=ws_bytes.value,
        )
        self._graph = None
        self._graph_unrolled = None
        self._torch_sample_ready = False
        return self._bufs

    def _plan(self, external_sample: bool) -> _cabi.GemPlan:
        b = self._buffers()
        p = _cabi.GemPlan()
        p.n, p.e, p.s = self.n, self.n_edges, b["S"]
        p.d, p.kp1 = int(self.n_components), b["kp1"]
        p.k_attr, p.l_min, p.k_inter = float(self.k_attr), float(self.L_min), float(self.k_inter)
        p.seed = self._sampler_seed & (2 ** 64 - 1)
        p.pos = self._pos.data_ptr()
        p.edges = self._edges32.data_ptr()
        p.row_ptr = self._row_ptr.data_ptr()
        p.col = self._col.data_ptr() if self.n_edges > 0 else None
        if self._layout.sorted_edges:                       # precondition of the vertex-parallel spring kernel
            p.up_ptr = self._up_ptr.data_ptr()
            p.hubs = self._hubs.data_ptr() if self._hubs.numel() > 0 else None
            p.n_hubs = int(self._hubs.numel())
        p.tau_hint = b["tau_hint"].data_ptr()
        p.force = b["force"].data_ptr()
        p.mid = b["mid"].data_ptr()
        p.qmid = b["qmid"].data_ptr()
        p.samp = b["samp"].data_ptr()
        p.knn_idx = b["knn_idx"].data_ptr()
        p.knn_dist = b["knn_dist"].data_ptr()
        p.iter_counter = b["iter"].data_ptr()
        p.knn_ws = b["knn_ws"].data_ptr()
        p.knn_ws_bytes = b["knn_ws_bytes"]
        p.stats_ws = b["stats_ws"].data_ptr()
        p.external_sample = 1 if external_sample else 0
        p.mm_mode = -1
        p.coef_slot = self._coef_slot
        return p

    # ------------------------------------------------------------------ the hot path
    # torch.randperm offloads n < 30000 to the CPU generator + a synchronous copy (not capturable); graphs with the
    # torch sampler are only built above this size
    _TORCH_SAMPLER_GRAPH_MIN_E = 65536

    def _draw_torch_sample(self, out: torch.Tensor):
        """The reference's draw (embedder_pytorch.py:404-413): torch.randperm(E, device)[:S] from the seeded
        default generator (or arange(E) when S == E) -> out."""
        E, S = self.n_edges, out.numel()
        out.copy_(torch.randperm(E, device=self.device)[:S] if S < E else torch.arange(E, device=self.device))

    def update_positions(self, sampled_indices=None):
        """One layout iteration (embedder_pytorch.py:776-806) = one gem_layout_step call.

        `sampled_indices` (optional, extension) injects the S query edge ids -- this is how the
        parity tests drive this class and the oracle from the same sample."""
        if self.n_edges == 0:
            # the reference reaches torch.topk with k+1 > 0 candidates and raises (:583)
            raise RuntimeError("selected index k out of range")
        b = self._buffers()
        external = sampled_indices is not None or self.sampler == "torch"
        if sampled_indices is not None:
            s = torch.as_tensor(sampled_indices).to(device=self.device, dtype=torch.long).reshape(-1)
            if s.numel() != b["S"]:
                raise ValueError(f"sampled_indices must have {b['S']} entries, got {s.numel()}")
            b["samp"].copy_(s)
        elif self.sampler == "torch":
            with torch.cuda.device(self.device):
                if self._torch_sample_ready:             # drawn ahead by the last graph replay: next in the RNG stream
                    b["samp"].copy_(b["samp_next"])
                    self._torch_sample_ready = False
                else:
                    self._draw_torch_sample(b["samp"])                       # :408-413
        plan = self._plan(external)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.gem_layout_step(ctypes.byref(plan), self._stream()), "gem_layout_step")
        self.last_sampled_indices = b["samp"]
        self.last_knn_indices = b["knn_idx"][:, 1:]

    def profile_step(self):
        """One iteration with a CUDA event after every stage (gem_profile_step; synchronises).
        Returns {stage name: milliseconds}.  Used by bench.py for the per-kernel roofline."""
        plan = self._plan(False)
        ms = (ctypes.c_float * len(_cabi.STAGE_NAMES))()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.gem_profile_step(ctypes.byref(plan), self._stream(), ms), "gem_profile_step")
        return dict(zip(_cabi.STAGE_NAMES, [float(x) for x in ms]))

    def profile_kernels(self, iterations: int = 10, scan_ctas: bool = False):
        """Per-kernel timeline of the replayed iteration: {kernel: (begin_us, end_us)} relative to the first kernel's
        start, median over `iterations` replays, from %globaltimer stamps written by the kernels themselves
        (gem_debug_stamps)."""
        lib = self._lib
        nk, words = lib.gem_debug_stamp_count(), lib.gem_debug_stamp_words()
        with torch.cuda.device(self.device):
            self.run_layout_device(4)
            buf = torch.zeros((words,), device=self.device, dtype=torch.int64)
            reset = torch.zeros((words,), dtype=torch.int64, device=self.device)
            reset[0:2 * nk:2] = -1                                        # begin = ~0, end = 0
            torch.cuda.synchronize(self.device)
            _cabi.check(lib.gem_debug_stamps(ctypes.c_void_p(buf.data_ptr())), "gem_debug_stamps")
            rows, ctas = [], None
            try:
                for _ in range(int(iterations)):
                    buf.copy_(reset)
                    self.run_layout_device(1)
                    torch.cuda.synchronize(self.device)
                    raw = buf.cpu().numpy().astype(np.uint64)
                    v = raw[: 2 * nk].reshape(nk, 2)
                    ran = v[:, 1] > 0
                    t0 = v[ran, 0].min()
                    rows.append(np.where(ran[:, None], (v.astype(np.float64) - float(t0)) * 1e-3, np.nan))

            finally:
                _cabi.check(lib.gem_debug_stamps(None), "gem_debug_stamps")
        med = np.nanmedian(np.asarray(rows), axis=0)
        out = {name: (round(float(med[i, 0]), 1), round(float(med[i, 1]), 1)) for i, name in enumerate(_cabi.STAMP_NAMES)
               if not np.isnan(med[i, 0])}
        return out

    def fp32_peak_flops(self) -> float:
        """Measured FP32 FMA throughput of this device in flop/s (gem_fp32_peak_probe)."""
        out, out2 = ctypes.c_double(0.0), ctypes.c_double(0.0)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.gem_fp32_peak_probe(ctypes.byref(out), ctypes.byref(out2), self._stream()),
                        "gem_fp32_peak_probe")
        self.fp32_peak_flops_packed = float(out2.value)
        return float(out.value)

    def _graph_ok(self) -> bool:
        """Can run_layout replay a captured iteration?  Always with the device sampler; with sampler='torch' when
        torch.randperm itself is capturable (E large enough to stay on the device)."""
        if not self.use_cuda_graph or self.n_neighbors + 1 > self.n_edges:
            return False
        if self.sampler == "device":
            return True
        return self.n_edges >= self._TORCH_SAMPLER_GRAPH_MIN_E and not getattr(self, "_torch_graph_failed", False)

    _GRAPH_UNROLL = 8

    def _capture_iterations(self, count: int, torch_samp: bool):
        """One CUDA graph of `count` consecutive iterations (both streams of gem_layout_step each)."""
        b = self._buffers()
        plan = self._plan(torch_samp)
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize(self.device)
        it0 = b["iter"].clone()
        pos0 = self._pos.clone()
        next0 = b["samp_next"].clone()
        torch.cuda.synchronize(self.device)
        rc = 0
        with _nvtx_range("gem.capture_iteration"), torch.cuda.graph(graph):
            for _ in range(int(count)):
                if torch_samp:
                    cur = torch.cuda.current_stream(self.device)
                    b["samp"].copy_(b["samp_next"])
                    side = torch.cuda.Stream(device=self.device)
                    side.wait_stream(cur)
                    with torch.cuda.stream(side):
                        self._draw_torch_sample(b["samp_next"])
                rc = rc or self._lib.gem_layout_step(ctypes.byref(plan), self._stream())
                if torch_samp:
                    cur.wait_stream(side)
        _cabi.check(rc, "gem_layout_step (capture)")
        # capture does not execute, but keep state exactly as it was in any case
        b["iter"].copy_(it0)
        self._pos.copy_(pos0)
        b["samp_next"].copy_(next0)
        return graph

    def _run_graph(self, num_iterations: int) -> bool:
        """Replay one captured iteration `num_iterations` times.

        sampler='torch': the sample of iteration t+1 depends on nothing in iteration t, so the captured graph draws it
        (torch.randperm from the default generator, registered with the graph) on a side stream WHILE iteration t
        runs and hands it over at the start of the next replay; the very first sample is drawn eagerly.  The samples
        are the reference's RNG stream in order (`seed=s` reproduces embedder_pytorch.py:409); one sample is always
        drawn ahead."""
        b = self._buffers()
        torch_samp = self.sampler == "torch"
        key = (self._pos.data_ptr(), b["key"], float(self.k_attr), float(self.L_min), float(self.k_inter), self.sampler)
        if torch_samp and not self._torch_sample_ready:      # first use, or an eager step consumed the one drawn ahead
            self._draw_torch_sample(b["samp_next"])
            self._torch_sample_ready = True
        if self._graph is None or self._graph_key != key:
            self._graph, self._graph_key, self._graph_unrolled = self._capture_iterations(1, torch_samp), key, None
        todo = int(num_iterations)
        # Several iterations per graph: consecutive replays leave ~8-11 us of launch latency between the last kernel
        # of one graph and the first of the next (C3: 254 us of kernels in a 266 us step; C1: 46 in 54); inside one
        # graph a dependent kernel starts 2-4 us after its predecessor.  The device sampler keeps its iteration
        # counter on the device, so the unrolled graph is the same launches in the same order.
        U = self._GRAPH_UNROLL
        if not torch_samp and U > 1 and todo >= U:
            if self._graph_unrolled is None:
                self._graph_unrolled = self._capture_iterations(U, torch_samp)
            for _ in range(todo // U):
                self._graph_unrolled.replay()
            todo %= U
        for _ in range(todo):
            self._graph.replay()
        self.last_sampled_indices = b["samp"]
        self.last_knn_indices = b["knn_idx"][:, 1:]
        return True

    def run_layout(self, num_iterations=100):
        """embedder_pytorch.py:808-833; returns the positions as ndarray."""
        if self.verbose:
            self.logger.info("Running layout for %d iterations", num_iterations)
        if self.n_edges == 0 and num_iterations > 0:
            raise RuntimeError("selected index k out of range")
        with torch.cuda.device(self.device), _nvtx_range(f"gem.run_layout[{int(num_iterations)}]"):
            if self.n_neighbors + 1 > self.n_edges and num_iterations > 0:
                raise RuntimeError("selected index k out of range")
            if self._graph_ok() and (num_iterations > 1 or (num_iterations == 1 and self._graph is not None)):
                self._run_graph(int(num_iterations))          # a single iteration too, once its graph exists
            else:
                for _ in range(int(num_iterations)):
                    self.update_positions()
        if self.verbose:
            self.logger.info("Layout computation completed")
        return self.positions

    def get_positions(self):
        """embedder_pytorch.py:835-844."""
        return self.positions

    # host-buffer fast paths of the positions setter / getter (extensions)
    def load_positions(self, host_positions: torch.Tensor):
        """Asynchronous H2D of an (n, d) fp32 host tensor (pinned for full speed) into the device state
        (what `emb.positions = tensor` does, minus the shape/dtype conversions)."""
        d = self.n_components
        if tuple(host_positions.shape) != (self.n, d) or host_positions.dtype != torch.float32:
            raise ValueError(f"expected an fp32 tensor of shape {(self.n, d)}")
        self._public_to_rows(host_positions)

    def read_positions(self, out: torch.Tensor):
        """D2H of the current positions into an (n, d) fp32 host tensor (pinned for full speed), then stream sync."""
        d = self.n_components
        with torch.cuda.device(self.device):
            if self._ld == d and self._pad_index is None:
                out.copy_(self._pos, non_blocking=True)
            else:
                out.copy_(self._rows_to_public(out=self._io_stage("d2h_stage")), non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        return out

    def close(self):
        """Release the object's coefficient slot and captured graph (also done when it is garbage collected)."""
        self._graph = None
        self._graph_unrolled = None
        if getattr(self, "_slot_finalizer", None) is not None:
            self._slot_finalizer()

    def run_layout_device(self, num_iterations=100):
        """run_layout without the final device->host copy (positions stay in HBM).  Replays the captured
        CUDA graph of one iteration (capturing it on first use with more than one iteration)."""
        with torch.cuda.device(self.device):
            if self.n_edges == 0 or self.n_neighbors + 1 > self.n_edges:
                raise RuntimeError("selected index k out of range")
            if self._graph_ok() and (num_iterations > 1 or self._graph is not None):
                self._run_graph(int(num_iterations))
            else:
                for _ in range(int(num_iterations)):
                    self.update_positions()

    # ------------------------------------------------------------------ stage-level (private, unit-tested) API
    def _pad_rows(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(device=self.device, dtype=torch.float32)
        d = self.n_components
        if self._ld == d and x.is_contiguous():
            return x
        buf = torch.zeros((x.shape[0], self._ld), device=self.device, dtype=torch.float32)
        buf[:, :d] = x
        return buf

    def _edges_as_int32(self, edges: torch.Tensor) -> torch.Tensor:
        """int32 copy of an edge tensor in the CALLER's vertex numbering.  `self._edges32` may only stand in for
        `self.edges` when the device numbering is the identity: with several ranks it holds PADDED ids, while the
        private stage methods take positions indexed by original ids."""
        if edges is self.edges:
            if self._pad_index is None:
                return self._edges32
            if getattr(self, "_edges32_public", None) is None:
                self._edges32_public = self.edges.to(device=self.device, dtype=torch.int32).contiguous()
            return self._edges32_public
        return edges.to(device=self.device, dtype=torch.int32).contiguous()

    def _spring_stage(self, positions, edges, want_mid):
        """(force (n,d), mid (e,d) or None) by the stage kernel the iteration itself would use:
        the vertex-parallel CSR kernel for the object's own (sorted) edge list and d in {2,3},
        the edge-parallel kernel for any other edge tensor / dimension."""
        d = int(self.n_components)
        use_csr = (edges is self.edges and self._layout.sorted_edges and d in (2, 3) and self._pad_index is None
                   and positions.shape[0] == self.n and self.n_edges > 0)
        pos = self._pad_rows(positions)
        n_rows = pos.shape[0]
        e32 = self._edges_as_int32(edges)
        force = torch.empty_like(pos)
        mid = torch.zeros((e32.shape[0] + 1, self._mld), device=self.device, dtype=torch.float32) if want_mid else None
        with torch.cuda.device(self.device):
            if use_csr:
                _cabi.check(self._lib.gem_spring_midpoints_csr(
                    _ptr(pos), _ptr(self._row_ptr), _ptr(self._col), _ptr(self._up_ptr), 0, n_rows,
                    _ptr(self._hubs) if self._hubs.numel() else None, int(self._hubs.numel()), d,
                    float(self.k_attr), float(self.L_min), _ptr(force), _ptr(mid), 0, self._stream()),
                    "gem_spring_midpoints_csr")
            else:
                _cabi.check(self._lib.gem_spring_midpoints(_ptr(pos), _ptr(e32), n_rows, e32.shape[0], d,
                                                           float(self.k_attr), float(self.L_min), _ptr(force),
                                                           _ptr(mid), self._stream()), "gem_spring_midpoints")
        return force[:, :d], (mid[:-1, :d] if want_mid else None)

    def _compute_spring_forces(self, positions, edges):
        """embedder_pytorch.py:595-636 -> (n, d) tensor."""
        return self._spring_stage(positions, edges, False)[0]

    def _compute_midpoints(self, positions, edges):
        """The expression at embedder_pytorch.py:785 -> (e, d) tensor (extension, used by tests)."""
        return self._spring_stage(positions, edges, True)[1]

    def _knn_points(self, query, reference, k, exact=False, return_distances=False):
        query = query.to(device=self.device, dtype=torch.float32).contiguous()
        reference = reference.to(device=self.device, dtype=torch.float32).contiguous()
        nq, d = query.shape
        nr = reference.shape[0]
        if k > nr:
            raise RuntimeError("selected index k out of range")          # torch.topk (:583)
        mld = self._lib.gem_mid_pitch(int(d))
        rm = torch.zeros((nr + 1, mld), device=self.device, dtype=torch.float32)
        qm = torch.zeros((nq, mld), device=self.device, dtype=torch.float32)
        idx = torch.empty((nq, k), device=self.device, dtype=torch.long)
        dist = torch.empty((nq, k), device=self.device, dtype=torch.float32)
        st = self._stream()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.gem_pack_points(_ptr(reference), nr, int(d), _ptr(rm), st), "gem_pack_points")
            _cabi.check(self._lib.gem_pack_points(_ptr(query), nq, int(d), _ptr(qm), st), "gem_pack_points")
            if exact:
                _cabi.check(self._lib.gem_knn_midpoints_exact(_ptr(rm), nr, 0, int(d), _ptr(qm), nq, int(k), -1,
                                                              _ptr(idx), _ptr(dist), st), "gem_knn_midpoints_exact")
            else:
                nbytes = ctypes.c_size_t(0)
                _cabi.check(self._lib.gem_knn_workspace_bytes(nr, int(d), nq, int(k), ctypes.byref(nbytes)))
                ws = torch.zeros((nbytes.value + 256,), device=self.device, dtype=torch.uint8)
                _cabi.check(self._lib.gem_knn_midpoints(_ptr(rm), nr, 0, int(d), _ptr(qm), nq, int(k), -1, None, _ptr(idx),
                                                        _ptr(dist), _ptr(ws), nbytes.value, self._coef_slot, st),
                            "gem_knn_midpoints")
        return (idx, dist) if return_distances else idx

    def _compute_knn_chunked(self, query_points, reference_points, k):
        """embedder_pytorch.py:426-483 -> (n_query, k) int64, rows ascending by (distance, index)."""
        return self._knn_points(query_points, reference_points, k)

    def _compute_knn_torch(self, query_points, reference_points, k, chunk_size):
        """embedder_pytorch.py:543-593 (cdist + topk); chunk_size is irrelevant here."""
        return self._knn_points(query_points, reference_points, k)

    def _compute_knn_pykeops(self, query_points, reference_points, k, chunk_size):
        raise ImportError("PyKeOps not available")                       # :511-514

    def _locate_knn_midpoints(self, midpoints, k, sampled_indices=None):
        """embedder_pytorch.py:381-424 -> (knn (S,k) with column 0 dropped, sampled ids (S,))."""
        E = midpoints.shape[0]
        S = min(int(self.sample_size), E)
        if sampled_indices is not None:
            samp = torch.as_tensor(sampled_indices).to(device=self.device, dtype=torch.long)
        elif S < E:
            if self.sampler == "torch":
                samp = torch.randperm(E, device=self.device)[:S]
            else:
                b = self._buffers()
                samp = torch.empty((S,), device=self.device, dtype=torch.long)
                with torch.cuda.device(self.device):
                    _cabi.check(self._lib.gem_sample_edges(self._sampler_seed & (2 ** 64 - 1), _ptr(b["iter"]), 1, E, S,
                                                           _ptr(samp), self._stream()), "gem_sample_edges")
        else:
            samp = torch.arange(E, device=self.device)
        midpoints = midpoints.to(device=self.device, dtype=torch.float32)
        knn = self._knn_points(midpoints[samp], midpoints, k + 1)
        return knn[:, 1:], samp

    def _compute_intersection_forces(self, positions, edges, knn_indices, sampled_indices):
        """embedder_pytorch.py:638-736 -> (n, d) tensor."""
        if self.n_components < 2:
            raise IndexError("index 1 is out of bounds for dimension 1 with size 1")   # :762
        pos = self._pad_rows(positions)
        e32 = self._edges_as_int32(edges)
        knn = knn_indices.to(device=self.device, dtype=torch.long)
        samp = sampled_indices.to(device=self.device, dtype=torch.long).contiguous()
        S, k = knn.shape
        full = torch.zeros((S, k + 1), device=self.device, dtype=torch.long)
        full[:, 1:] = knn
        force = torch.zeros_like(pos)
        if S * k > 0:
            with torch.cuda.device(self.device):
                _cabi.check(self._lib.gem_intersection_forces(_ptr(pos), _ptr(e32), pos.shape[0], int(self.n_components),
                                                              _ptr(samp), _ptr(full), S, k + 1, float(self.k_inter),
                                                              _ptr(force), self._stream()), "gem_intersection_forces")
        return force[:, : self.n_components]

    def _check_line_intersections(self, p1, p2, q1, q2):
        """embedder_pytorch.py:738-774 -> bool (P,)."""
        ts = [t.to(device=self.device, dtype=torch.float32).contiguous() for t in (p1, p2, q1, q2)]
        P, d = ts[0].shape
        out = torch.zeros((P,), device=self.device, dtype=torch.uint8)
        if P > 0:
            with torch.cuda.device(self.device):
                _cabi.check(self._lib.gem_check_line_intersections(_ptr(ts[0]), _ptr(ts[1]), _ptr(ts[2]), _ptr(ts[3]),
                                                                   P, int(d), _ptr(out), self._stream()),
                            "gem_check_line_intersections")
        return out.bool()

    def _apply_update(self, positions, spring_forces, inter_forces):
        """Tail of update_positions (embedder_pytorch.py:796-804) -> (n, d) tensor (extension)."""
        pos = self._pad_rows(positions).clone()
        fs = self._pad_rows(spring_forces)
        fi = self._pad_rows(inter_forces) if inter_forces is not None else None
        nbytes = ctypes.c_size_t(0)
        _cabi.check(self._lib.gem_update_workspace_bytes(pos.shape[0], int(self.n_components), ctypes.byref(nbytes)))
        ws = torch.zeros((nbytes.value + 256,), device=self.device, dtype=torch.uint8)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.gem_update_positions(_ptr(pos), _ptr(fs), _ptr(fi), pos.shape[0], pos.shape[0],
                                                       int(self.n_components), _ptr(ws), 0, self._stream()),
                        "gem_update_positions")
        return pos[:, : self.n_components]

    # ------------------------------------------------------------------ misc
    def display_layout(self, edge_width=1, node_size=3, node_colors=None):
        """embedder_pytorch.py:846-969: plotly scatter of the layout (needs plotly)."""
        if self.n_components not in (2, 3):
            raise ValueError("Can only display 2D or 3D layouts")
        import plotly.graph_objects as go  # pylint: disable=import-outside-toplevel
        pos = self.get_positions()
        ed = self.edges.cpu().numpy()
        seg = np.full((len(ed), 3, self.n_components), np.nan)
        seg[:, 0] = pos[ed[:, 0]]
        seg[:, 1] = pos[ed[:, 1]]
        seg = seg.reshape(-1, self.n_components)
        marker = {"color": node_colors if node_colors is not None else "red", "colorscale": "Bluered",
                  "size": node_size, "showscale": node_colors is not None}
        if self.n_components == 2:
            traces = [go.Scatter(x=seg[:, 0], y=seg[:, 1], mode="lines", line={"color": "gray", "width": edge_width},
                                 hoverinfo="none"),
                      go.Scatter(x=pos[:, 0], y=pos[:, 1], mode="markers", marker=marker, hoverinfo="none")]
        else:
            traces = [go.Scatter3d(x=seg[:, 0], y=seg[:, 1], z=seg[:, 2], mode="lines",
                                   line={"color": "gray", "width": edge_width}, hoverinfo="none"),
                      go.Scatter3d(x=pos[:, 0], y=pos[:, 1], z=pos[:, 2], mode="markers", marker=marker,
                                   hoverinfo="none")]
        fig = go.Figure(data=traces)
        fig.update_layout(title=f"{self.n_components}D Graph Embedding (B200)", showlegend=False, width=800, height=800)
        fig.show()

    def __repr__(self):
        return (f"GraphEmbedderPyTorch(n_vertices={self.n}, n_components={self.n_components}, "
                f"device={self.device}, memory_efficient={self.memory_efficient})")


def create_graphem(adjacency, n_components=2, backend=None, **kwargs):
    """graphem_rapids/__init__.py:78-136.  Every GPU backend name routes to the B200 class; the
    reference's backend selection (utils/backend_selection.py) is not part of this path."""
    if backend == "cpu":
        raise RuntimeError("graphem_rapids_b200 has no CPU backend (backend='cpu' requested)")
    if backend not in _SUPPORTED_BACKENDS:
        raise ValueError(f"Unknown backend {backend!r}")
    return GraphEmbedderPyTorch(adjacency, n_components, **kwargs)
