#!/usr/bin/env python
"""Per-kernel times of ONE rank's share of the multi-GPU product flow, measured on a single GPU.

ncu must never wrap a multi-rank command, and CUDA events inside a real multi-rank step mix kernel time with rank
skew.  This script builds the R-rank layout of a workload, instantiates rank 0's engine only and "peer-maps" its
buffers to R-1 local dummy replicas (same store volume, no NVLink), then times every launch of the product flow in
isolation: each call is enqueued behind a long fill kernel, so host launch gaps do not enter the event interval.
usage: rank_share_profile.py [workload=c3] [world=8] [reps=10]      (ncu-friendly: one process, one GPU)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np          # noqa: E402
import torch                # noqa: E402

import bench                # noqa: E402
from graphem_rapids_b200 import _cabi                                   # noqa: E402
from graphem_rapids_b200.partition import build_layout                  # noqa: E402
from graphem_rapids_b200.sharded import CudaStages, ShardedLayoutEngine  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    w = bench.WORKLOADS[wl]
    adj = bench.make_graph(w)
    n, d, k, S = adj.shape[0], w["d"], w["k"], w["S"]
    rows, cols = adj.nonzero()
    keep = rows < cols
    e = np.column_stack([rows[keep], cols[keep]]).astype(np.int64)
    lib = _cabi.load()
    L = build_layout(e, n, R, hub_degree=lib.gem_hub_degree(), ownership="strided")
    dev = torch.device("cuda:0")
    st = CudaStages(L, 0, dev, n_components=d, k_attr=0.2, L_min=1.0, k_inter=0.5, seed=0)
    eng = ShardedLayoutEngine(L, 0, st, n_components=d, n_neighbors=k, sample_size=S)
    ld = st.ld
    raws = [torch.zeros((2, L.n_pad, ld), device=dev) for _ in range(R)]
    xb = CudaStages.exchange_bytes(R, eng._nb, ld)
    xchgs = [torch.zeros((xb,), device=dev, dtype=torch.uint8) for _ in range(R)]
    poss = [eng.pos] + [torch.zeros_like(eng.pos) for _ in range(R - 1)]
    st.attach_p2p([t.data_ptr() for t in poss], [t.data_ptr() for t in raws], [t.data_ptr() for t in xchgs], raws[0], xchgs[0],
                  eng._nb, S, k + 1, barrier=lambda ch: None)
    eng.set_positions(torch.from_numpy(bench.initial_positions(n, d)))
    # the absent ranks' partial lists: padding rows (+inf, -1), both parities
    ib, lb = st._ib, st._list_bytes
    for par in (0, 1):
        blk = xchgs[0][par * st._parity_bytes: par * st._parity_bytes + R * lb].view(R, lb)
        blk[1:, :ib].view(torch.int64).fill_(-1)
        blk[1:, ib: ib + S * (k + 1) * 4].view(torch.float32).fill_(float("inf"))
    big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    names = ["prep", "spring", "colsum", "scan_select", "merge_intersect", "normalise"]
    acc = {k_: [] for k_ in names}
    for it in range(reps + 2):
        st._marks = {}
        for _ in range(4):                        # ~4 x 170 us of GPU work: the launches below queue up behind it
            big.fill_(it & 255)
        # phase 1 in SERIES on one stream (side stream work moved to main by marking order): isolate each launch
        st.p2p_phase1(eng)
        torch.cuda.synchronize()
        big.fill_(1); big.fill_(1)
        st._mark("pre2")
        st.p2p_phase2(eng)
        torch.cuda.synchronize()
        big.fill_(2); big.fill_(2)
        st._mark("pre3")
        st.p2p_phase3(eng)
        eng.iteration += 1
        torch.cuda.synchronize()
        m = st._marks
        if it >= 2:
            acc["prep"].append(m["start"].elapsed_time(m["prep"]))
            acc["spring"].append(m["start"].elapsed_time(m["spring"]))
            acc["colsum"].append(m["spring"].elapsed_time(m["colsum"]))
            acc["scan_select"].append(max(m["spring"].elapsed_time(m["scan_select"]), 0.0))
            acc["merge_intersect"].append(m["pre2"].elapsed_time(m["merge_intersect"]))
            acc["normalise"].append(m["pre3"].elapsed_time(m["normalise"]))
    st._marks = None
    out = {k_: round(float(np.median(v)) * 1e3, 1) for k_, v in acc.items()}
    print(f"rank-0 share of {wl} at world={R} on one GPU (us; prep and spring run CONCURRENTLY, both measured from the step start; "
          f"colsum and scan_select from the end of spring): {out}")
    print(f"rows owned {eng.ve - eng.vb}, candidates owned {eng.e_hi - eng.e_lo}")


if __name__ == "__main__":
    main()
