"""One-off robustness check at the largest configs (C4 / C5): two fused iterations, neighbour lists vs the
in-library exact kernel, positions vs the stage-by-stage path, timing.   usage: large_check.py c5"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench
import graphem_rapids_b200 as gr
from gem_testutil import rel_inf
wl = sys.argv[1] if len(sys.argv) > 1 else "c5"
w = bench.WORKLOADS[wl]
t0 = time.time(); adj = bench.make_graph(w); print(f"graph {time.time()-t0:.1f}s n={adj.shape[0]} nnz={adj.nnz}", flush=True)
n, d, k = adj.shape[0], w["d"], w["k"]
t0 = time.time()
emb = gr.GraphEmbedderPyTorch(adj, n_components=d, device="cuda:0", n_neighbors=k, sample_size=w["S"], verbose=False, seed=0,
                              initial_positions=bench.initial_positions(n, d))
print(f"embedder {time.time()-t0:.1f}s E={emb.n_edges} hubs={emb._hubs.numel()} mem={torch.cuda.memory_allocated()/2**30:.2f} GiB", flush=True)
for it in range(2):
    before = emb._positions.clone()
    mid = emb._compute_midpoints(before, emb.edges)
    F = emb._compute_spring_forces(before, emb.edges)
    emb.update_positions()
    samp = emb.last_sampled_indices.clone()
    knn_full = emb._bufs["knn_idx"].clone()
    ex_idx, ex_dist = emb._knn_points(mid[samp], mid, k + 1, exact=True, return_distances=True)
    G = emb._compute_intersection_forces(before, emb.edges, knn_full[:, 1:], samp)
    staged = emb._apply_update(before, F, G).cpu().numpy()
    got = emb.positions
    print(f"it{it}: knn==exact {torch.equal(knn_full, ex_idx)} dist==exact {torch.equal(emb._bufs['knn_dist'], ex_dist)} "
          f"pos vs staged {rel_inf(got, staged):.2e} finite {bool(np.all(np.isfinite(got)))}", flush=True)
    del mid, F, G, before
emb.run_layout_device(5); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); emb.run_layout_device(20); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"{wl}: {ms:.3f} ms/iteration back to back = {emb.n_edges / ms / 1e6:.2f}e9 edge-updates/s", flush=True)
print({k_: round(v, 4) for k_, v in emb.profile_step().items()})
