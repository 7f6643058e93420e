"""Diagnostics of the KNN scan on a workload: survivors per query, thresholds, slow-path counters."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphem_rapids_b200 as gr
from graphem_rapids_b200 import _cabi
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
w = bench.WORKLOADS[wl]
adj = bench.make_graph(w)
emb = gr.GraphEmbedderPyTorch(adj, n_components=w["d"], device="cuda:0", n_neighbors=w["k"], sample_size=w["S"],
                              verbose=False, seed=0, initial_positions=bench.initial_positions(adj.shape[0], w["d"]))
lib = _cabi.load()
so, co, to = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
cap, g = ctypes.c_int(), ctypes.c_int()
lib.gem_knn_debug_stats(1, emb.n_edges, w["d"], min(w["S"], emb.n_edges), w["k"] + 1, ctypes.byref(so), ctypes.byref(co),
                        ctypes.byref(to), ctypes.byref(cap), ctypes.byref(g))
print("cap", cap.value, "g", g.value)
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 8):
    b = emb._buffers()
    ws = b["knn_ws"]
    ws[so.value:so.value + 64].zero_()
    ms = emb.profile_step()
    torch.cuda.synchronize()
    S = b["S"]
    stats = ws[so.value:so.value + 64].view(torch.int64).cpu().numpy()
    counts = ws[co.value:co.value + 4 * S].view(torch.int32).cpu().numpy()
    tau = ws[to.value:to.value + 4 * S].view(torch.float32).cpu().numpy()
    kd = b["knn_dist"][:, -1].cpu().numpy()
    pos = emb._positions
    r = pos.norm(dim=1)
    print(f"it{it}: scan {ms['knn_scan']:.3f} ms bound {ms['knn_bound']:.3f} thr {ms['knn_threshold']:.3f} sel {ms['knn_select']:.3f} "
          f"spring {ms['spring_mid']:.3f} upd {ms['update']:.3f} | rejected {stats[0]} accepted {stats[1]} inserts {stats[2]} warp-slow {stats[3]} | "
          f"counts mean {counts.mean():.0f} max {counts.max()} | tau/d11 ratio median {np.median(tau / np.maximum(kd, 1e-30)):.2f} max {np.max(tau / np.maximum(kd, 1e-30)):.3g} "
          f"| |pos| med {r.median():.3g} max {r.max():.3g}")
pk = emb.fp32_peak_flops()
print(f"fp32 peak: FFMA {pk/1e12:.1f} TFLOP/s, FFMA2 {emb.fp32_peak_flops_packed/1e12:.1f} TFLOP/s")
