"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck): a few iterations on small graphs that
exercise every kernel: fast path d=3 and d=2 (with hubs), generic d, tiny exact path, k=32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphem_rapids_b200 as gr
for kind, n, d, k in [("ba", 6000, 3, 10), ("rr", 5000, 2, 10), ("rr", 300, 5, 6), ("ba", 3000, 3, 32), ("rr", 40, 2, 5)]:
    adj = gr.generate_ba(n, 4, seed=1) if kind == "ba" else gr.generate_random_regular(n, 6, seed=1)
    pos0 = np.random.default_rng(0).standard_normal((n, d)).astype(np.float32)
    emb = gr.GraphEmbedderPyTorch(adj, n_components=d, device="cuda:0", n_neighbors=k, sample_size=64, verbose=False,
                                  seed=0, initial_positions=pos0, use_cuda_graph=False)
    for _ in range(2):
        emb.update_positions()
    torch.cuda.synchronize()
    p = emb.positions
    assert np.all(np.isfinite(p)), (kind, n, d)
    print("ok", kind, n, d, k, float(np.abs(p).max()), flush=True)
