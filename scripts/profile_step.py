"""Run a few plain layout iterations of a bench workload (no diagnostics, no timing): the target of
`ncu -k regex:<kernel>` captures.   usage: profile_step.py [workload=c3] [iterations=6]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import graphem_rapids_b200 as gr
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
w = bench.WORKLOADS[wl]
adj = bench.make_graph(w)
emb = gr.GraphEmbedderPyTorch(adj, n_components=w["d"], device="cuda:0", n_neighbors=w["k"], sample_size=w["S"],
                              verbose=False, seed=0, initial_positions=bench.initial_positions(adj.shape[0], w["d"]))
for _ in range(iters):
    emb.update_positions()
torch.cuda.synchronize()
print("ok", wl, iters, float(emb._positions.abs().max()))
