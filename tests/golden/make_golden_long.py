#!/usr/bin/env python
"""
50-iteration golden runs of the REAL reference (north-star check: after 50 iterations the Spearman
correlation of radial distance with degree / betweenness centrality must be within 0.01 of the
reference's).  Records, per case: edges, the initial positions the reference started from, the 50
samples it drew (captured around torch.randperm), its final positions, degree and exact betweenness
centrality (networkx) and the two reference Spearman values.  Re-run (needs /root/reference):
    python tests/golden/make_golden_long.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _import_reference  # noqa: E402


def record(name, emb, torch, iters=50, exact_betweenness=True, store_pos0=True, ensemble=6):
    """`ensemble` replays of the same 50 samples by the oracle (bit-identical to the reference, see
    tests/test_oracle_golden.py) from initial positions perturbed by 1e-7 relative: the spread of the two
    Spearman values under a perturbation of the size of one fp32 rounding.  The iteration is chaotic (a 1e-7
    perturbation moves the final positions by O(1) at n ~ 1000), so this spread IS the reference's resolution."""
    import networkx as nx
    from scipy.stats import spearmanr
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import oracle
    pos0 = emb._positions.clone()
    E = emb.edges.shape[0]
    S = min(emb.sample_size, E)
    samples = []
    for _ in range(iters):
        st = torch.get_rng_state()
        samples.append((torch.randperm(E)[:S] if S < E else torch.arange(E)).numpy())
        torch.set_rng_state(st)
        emb.update_positions()
    final = emb._positions.clone().numpy()
    edges = emb.edges.numpy()
    g = nx.Graph()
    g.add_nodes_from(range(emb.n))
    g.add_edges_from(edges.tolist())
    deg = np.array([g.degree(i) for i in range(emb.n)], dtype=np.float64)
    if exact_betweenness:
        btw_d = nx.betweenness_centrality(g)
    else:                                   # pivot-sampled betweenness at >= 100 K vertices
        btw_d = nx.betweenness_centrality(g, k=64, seed=0)
    btw = np.array([btw_d[i] for i in range(emb.n)], dtype=np.float64)
    radius = np.linalg.norm(final, axis=1)
    rho_deg = float(spearmanr(radius, deg).correlation)
    rho_btw = float(spearmanr(radius, btw).correlation)
    ens_d, ens_b = [], []
    et = torch.from_numpy(edges.astype(np.int64))
    st = [torch.from_numpy(x.astype(np.int64)) for x in samples]
    for k in range(ensemble):
        rng = np.random.default_rng(100 + k)
        p = torch.from_numpy((pos0.numpy() * (1.0 + 1e-7 * rng.standard_normal(pos0.shape))).astype(np.float32))
        out = oracle.run_layout(p, et, iters, sample_size=emb.sample_size, n_neighbors=emb.n_neighbors, strict=True,
                                samples=st).numpy()
        rr = np.linalg.norm(out, axis=1)
        ens_d.append(float(spearmanr(rr, deg).correlation))
        ens_b.append(float(spearmanr(rr, btw).correlation))
    path = os.path.join(HERE, "long", name + ".npz")
    extra = dict(pos0=pos0.numpy(), final_pos=final, edges=edges.astype(np.int32)) if store_pos0 else {}
    np.savez_compressed(path, n=np.int64(emb.n), d=np.int64(emb.n_components),
                        n_neighbors=np.int64(emb.n_neighbors), sample_size=np.int64(emb.sample_size),
                        samples=np.stack(samples).astype(np.int32),
                        degree=deg.astype(np.float32), betweenness=btw.astype(np.float32), rho_degree=np.float64(rho_deg),
                        rho_betweenness=np.float64(rho_btw), ens_rho_degree=np.array(ens_d), ens_rho_betweenness=np.array(ens_b),
                        **extra)
    print(f"{name}: N={emb.n} E={E} rho(radius,degree)={rho_deg:+.4f} rho(radius,betweenness)={rho_btw:+.4f} "
          f"ensemble degree {np.mean(ens_d):+.4f}+-{np.std(ens_d):.4f} betweenness {np.mean(ens_b):+.4f}+-{np.std(ens_b):.4f} "
          f"-> {os.path.getsize(path)/1024:.0f} KiB")


def main():
    import torch
    Emb, gen = _import_reference()
    # C1 (BASELINE.json configs[0], README quick start): ER n=1000 p=0.01, d=3, k=10, 50 iterations, Laplacian init
    adj = gen.erdos_renyi_graph(n=1000, p=0.01, seed=0)
    record("c1_er1000_d3_50it", Emb(adj, n_components=3, n_neighbors=10, seed=0, verbose=False), torch)
    # preferential attachment (C3 scaled down), random init like the bench
    adj = gen.generate_ba(n=2000, m=4, seed=0)
    emb = Emb(adj, n_components=3, n_neighbors=10, seed=0, verbose=False)
    emb.positions = (np.random.default_rng(0).standard_normal((2000, 3)) * 0.1).astype(np.float32)
    record("ba2000_d3_50it", emb, torch)
    # BASELINE-scale statistics: BA n=100 000 m=4 (the generator and the initial positions are the
    # repo's deterministic ones, so the vectors hold only the samples and the reference's Spearman values)
    import graphem_rapids_b200.generators as mygen
    adj = mygen.generate_ba(100_000, 4, seed=0)
    emb = Emb(adj, n_components=3, n_neighbors=10, seed=0, verbose=False)
    emb.positions = (np.random.default_rng(0).standard_normal((100_000, 3)) * 0.1).astype(np.float32)
    record("ba100k_d3_50it", emb, torch, exact_betweenness=False, store_pos0=False, ensemble=3)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    main()
