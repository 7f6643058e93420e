#!/usr/bin/env python
"""
Generate the golden vectors under tests/golden/*.npz by running the REAL reference
(/root/reference, graphem_rapids v0.2.0, CPU, torch 2.11) in the build container.

The reference has no golden vectors of its own, so its outputs on fixed
inputs are recorded here.  Re-run:  python tests/golden/make_golden.py
(needs /root/reference; the GPU box does not have it, which is why the vectors are committed).

For every case, from FIXED positions `pos0` and the sample `samp` the reference itself draws
(captured by saving / restoring the torch RNG state around `torch.randperm`):
    F_spring   = emb._compute_spring_forces(pos0, edges)              embedder_pytorch.py:595-636
    mid        = (pos0[e0] + pos0[e1]) / 2.0                          :785
    knn_full   = emb._compute_knn_chunked(mid[samp], mid, k+1)        :426-483 (cdist + topk)
    knn_fdist  = torch.cdist(mid[samp], mid) gathered at knn_full     (the reference's distances)
    knn        = knn_full[:, 1:]                                      :421
    F_inter    = emb._compute_intersection_forces(pos0, edges, knn, samp)   :638-736
    new_pos    = emb.update_positions() from pos0 with the same RNG state   :776-806
and a short trajectory: `traj_samps` (T,S) drawn by the reference and `traj_pos` after T
iterations of `update_positions`.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    stubs = tempfile.mkdtemp(prefix="graphem_stubs_")
    for mod in ["plotly", "plotly/graph_objects", "plotly/express", "ndlib", "ndlib/models",
                "ndlib/models/ModelConfig", "ndlib/models/epidemics"]:
        if mod.count("/") == 0 or mod == "ndlib/models":
            os.makedirs(os.path.join(stubs, mod), exist_ok=True)
            open(os.path.join(stubs, mod, "__init__.py"), "w").close()
        else:
            open(os.path.join(stubs, mod + ".py"), "w").close()
    os.environ["GRAPHEM_RAPIDS_QUIET"] = "true"
    sys.path.insert(0, stubs)
    sys.path.insert(0, "/root/reference")
    import graphem_rapids  # noqa: F401
    from graphem_rapids.backends.embedder_pytorch import GraphEmbedderPyTorch
    import graphem_rapids.generators as gen
    return GraphEmbedderPyTorch, gen


def record(name, emb, torch, traj_steps=3, pre_steps=0):
    """Record one case from the embedder's current state (after `pre_steps` warm iterations)."""
    for _ in range(pre_steps):
        emb.update_positions()
    pos0 = emb._positions.clone()
    edges = emb.edges
    E = edges.shape[0]
    k = emb.n_neighbors

    # the sample the reference is about to draw
    state = torch.get_rng_state()
    S = min(emb.sample_size, E)
    samp = torch.randperm(E)[:S] if S < E else torch.arange(E)
    torch.set_rng_state(state)

    F_spring = emb._compute_spring_forces(pos0, edges)
    mid = (pos0[edges[:, 0]] + pos0[edges[:, 1]]) / 2.0
    knn_full = emb._compute_knn_chunked(mid[samp], mid, k + 1)
    dmat = torch.cdist(mid[samp], mid, p=2)
    knn_fdist = torch.gather(dmat, 1, knn_full)
    knn = knn_full[:, 1:]
    F_inter = emb._compute_intersection_forces(pos0, edges, knn, samp)

    torch.set_rng_state(state)
    emb.update_positions()
    new_pos = emb._positions.clone()

    # trajectory from pos0: T iterations, samples captured
    emb._positions = pos0.clone()
    torch.set_rng_state(state)
    traj_samps = []
    for _ in range(traj_steps):
        st = torch.get_rng_state()
        traj_samps.append((torch.randperm(E)[:S] if S < E else torch.arange(E)).numpy())
        torch.set_rng_state(st)
        emb.update_positions()
    traj_pos = emb._positions.clone()

    out = dict(
        edges=edges.numpy().astype(np.int32), n=np.int64(emb.n), d=np.int64(emb.n_components),
        n_neighbors=np.int64(k), sample_size=np.int64(emb.sample_size),
        k_attr=np.float64(emb.k_attr), L_min=np.float64(emb.L_min), k_inter=np.float64(emb.k_inter),
        pos0=pos0.numpy(), samp=samp.numpy().astype(np.int64),
        F_spring=F_spring.numpy(), mid=mid.numpy(), knn_full=knn_full.numpy().astype(np.int32),
        knn_fdist=knn_fdist.numpy(), F_inter=F_inter.numpy(), new_pos=new_pos.numpy(),
        traj_samps=np.stack(traj_samps).astype(np.int64), traj_pos=traj_pos.numpy(),
    )
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    n_pairs = int((F_inter.abs().sum(1) > 0).sum())
    print(f"{name}: N={emb.n} E={E} d={emb.n_components} S={S} k={k} "
          f"verts_with_inter_force={n_pairs} -> {os.path.getsize(path)/1024:.0f} KiB")


def main():
    import torch
    import scipy.sparse as sp
    Emb, gen = _import_reference()
    kw = dict(verbose=False)

    # C1: README quick start (BASELINE.json configs[0]); Laplacian init, then 5 warm iterations
    adj = gen.erdos_renyi_graph(n=1000, p=0.01, seed=0)
    record("er1000_d3_it0", Emb(adj, n_components=3, n_neighbors=10, seed=0, **kw), torch)
    record("er1000_d3_it5", Emb(adj, n_components=3, n_neighbors=10, seed=0, **kw), torch, pre_steps=5)

    # d=2, scaled-down C2 (random regular degree 8)
    adj = gen.generate_random_regular(n=2000, d=8, seed=0)
    record("rr2000_d2_it3", Emb(adj, n_components=2, n_neighbors=10, seed=0, **kw), torch, pre_steps=3)

    # scaled-down C3 (Barabasi-Albert m=4) from the reference's random-init fallback distribution
    adj = gen.generate_ba(n=3000, m=4, seed=0)
    emb = Emb(adj, n_components=3, n_neighbors=10, seed=0, **kw)
    emb.positions = (np.random.default_rng(0).standard_normal((3000, 3)) * 0.1).astype(np.float32)
    record("ba3000_d3_rand", emb, torch, pre_steps=2)

    # KNN-heavy (C4-like k=32) small SBM
    adj = gen.generate_sbm(n_per_block=150, num_blocks=4, p_in=0.05, p_out=0.005, seed=0)
    record("sbm600_d3_k32", Emb(adj, n_components=3, n_neighbors=32, sample_size=128, seed=0, **kw),
           torch, pre_steps=2)

    # RR n=50 d=4 seed=42 is the graph the reference's unit tests use throughout; n_components=4
    adj = gen.generate_random_regular(n=50, d=4, seed=42)
    record("rr50_d4", Emb(adj, n_components=4, n_neighbors=10, seed=42, **kw), torch, pre_steps=1)
    record("rr30_k15", Emb(gen.generate_random_regular(n=30, d=4, seed=42), n_components=2, L_min=10.0,
                           k_attr=0.5, k_inter=0.1, n_neighbors=15, sample_size=64, seed=1, **kw), torch)

    # tiny graphs of the reference's tests: cdist DIRECT mode (<= 25 rows both sides)
    two_tri = np.array([[0, 1, 1, 0, 0, 0], [1, 0, 1, 0, 0, 0], [1, 1, 0, 0, 0, 0],
                        [0, 0, 0, 0, 1, 1], [0, 0, 0, 1, 0, 1], [0, 0, 0, 1, 1, 0]])
    record("two_triangles", Emb(two_tri, n_components=2, L_min=10.0, k_attr=0.5, k_inter=0.1,
                                n_neighbors=5, sample_size=6, seed=0, **kw), torch)      # tests/test_pytorch_backend.py:187-205
    prism = np.array([[0, 1, 1, 0, 0, 1], [1, 0, 1, 1, 0, 0], [1, 1, 0, 0, 1, 0],
                      [0, 1, 0, 0, 1, 1], [0, 0, 1, 1, 0, 1], [1, 0, 0, 1, 1, 0]])
    record("prism6", Emb(prism, n_components=2, n_neighbors=3, seed=0, **kw), torch)       # tests/test_integration.py:181-196
    hexes = np.vstack([np.array([[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [5, 0]]),
                       np.array([[6, 7], [7, 8], [8, 9], [9, 10], [10, 11], [11, 6]])])
    adj = sp.csr_matrix((np.ones(len(hexes)), (hexes[:, 0], hexes[:, 1])), shape=(12, 12))
    adj = adj + adj.T
    record("two_hexagons", Emb(adj, n_components=2, seed=0, **kw), torch)                  # tests/test_integration.py:278-296
    k4 = np.ones((4, 4)) - np.eye(4)
    record("k4", Emb(k4, n_components=2, n_neighbors=3, seed=0, **kw), torch)              # tests/conftest.py:25-28


if __name__ == "__main__":
    main()
