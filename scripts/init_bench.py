"""Time the device Laplacian initial embedding on a bench workload.  usage: init_bench.py c3"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, graphem_rapids_b200 as gr
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
w = bench.WORKLOADS[wl]
adj = bench.make_graph(w)
n, d = adj.shape[0], w["d"]
emb = gr.GraphEmbedderPyTorch(adj, n_components=d, device="cuda:0", n_neighbors=w["k"], verbose=False, seed=0,
                              initial_positions=np.zeros((n, d), np.float32))
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    vecs, info = emb._laplacian_embedding_device(return_info=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"{wl}: n={n} E={emb.n_edges} device init {dt*1e3:.1f} ms, outer {info['outer']}, SpMV {info['spmv']}, residual {info['residual']:.2e}, theta {[round(x,5) for x in info['theta']]}", flush=True)
