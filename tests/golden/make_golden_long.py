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


def record(name, emb, torch, iters=50):
    import networkx as nx
    from scipy.stats import spearmanr
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
    btw_d = nx.betweenness_centrality(g)
    btw = np.array([btw_d[i] for i in range(emb.n)], dtype=np.float64)
    radius = np.linalg.norm(final, axis=1)
    rho_deg = float(spearmanr(radius, deg).correlation)
    rho_btw = float(spearmanr(radius, btw).correlation)
    path = os.path.join(HERE, "long", name + ".npz")
    np.savez_compressed(path, edges=edges.astype(np.int32), n=np.int64(emb.n), d=np.int64(emb.n_components),
                        n_neighbors=np.int64(emb.n_neighbors), sample_size=np.int64(emb.sample_size),
                        pos0=pos0.numpy(), samples=np.stack(samples).astype(np.int32), final_pos=final,
                        degree=deg, betweenness=btw, rho_degree=np.float64(rho_deg), rho_betweenness=np.float64(rho_btw))
    print(f"{name}: N={emb.n} E={E} rho(radius,degree)={rho_deg:+.4f} rho(radius,betweenness)={rho_btw:+.4f} "
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


if __name__ == "__main__":
    main()
