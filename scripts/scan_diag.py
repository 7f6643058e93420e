#!/usr/bin/env python
"""Where does the FIXED cost of a knn_scan_kernel launch go?  (C1: 27 candidate blocks take 20 us.)

Uses the diagnostic build of the library (-DGEM_SCAN_DIAG=1 -> graphem_rapids_b200/libgraphem_b200_diag.so, built on
demand by graphem_rapids_b200.build.build_diag()): thread 0 of every scan CTA records %globaltimer at the kernel's phase boundaries; this script
replays the captured iteration of several workloads and prints, per phase, the mean / max over CTAs.
usage: scan_diag.py [lib] [workload ...]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np          # noqa: E402
import torch                # noqa: E402

import graphem_rapids_b200.build as _b      # noqa: E402

LIBP = sys.argv[1] if len(sys.argv) > 1 else _b.DIAG_LIB
if not os.path.exists(LIBP) or os.path.getmtime(LIBP) < os.path.getmtime(_b.SRC):
    assert os.path.abspath(LIBP) == os.path.abspath(_b.DIAG_LIB), f"{LIBP} is missing or older than the source"
    _b.build_diag()
_b.LIB = LIBP
_b.needs_build = lambda: False
import bench                                 # noqa: E402
import graphem_rapids_b200 as gr             # noqa: E402
from graphem_rapids_b200 import _cabi        # noqa: E402

PH = ["init(t0->sync)", "first TMA wait", "block0 main loop", "block0 slow path", "fetch + wait blk1", "block1 main loop",
      "block1 slow path", "rest of loop (warp 0)", "wait for other warps", "publish"]


def run(wl, lib):
    w = dict(bench.WORKLOADS[wl]) if wl in bench.WORKLOADS else None
    if w is None:                                   # "ba125k": one rank's share of C3 at 8 GPUs, as a whole problem
        w = dict(desc=wl, kind="ba", n=int(wl[2:]), m=4, d=3, k=10, S=256)
    adj = bench.make_graph(w)
    n, d = adj.shape[0], w["d"]
    dev = torch.device("cuda:0")
    emb = gr.GraphEmbedderPyTorch(adj, n_components=d, device=dev, n_neighbors=w["k"], sample_size=w["S"], verbose=False,
                                  seed=0, initial_positions=bench.initial_positions(n, d))
    emb.run_layout_device(6)
    torch.cuda.synchronize()
    stamps = emb.profile_kernels(5)
    ncta = 1024
    buf = torch.zeros((ncta * 12 + ncta * 128,), device=dev, dtype=torch.int64)
    lib.gem_debug_scan_diag.restype = ctypes.c_int
    lib.gem_debug_scan_diag.argtypes = [ctypes.c_void_p]
    _cabi.check(lib.gem_debug_scan_diag(ctypes.c_void_p(buf.data_ptr())))
    pbuf = torch.zeros((1200 * 12,), device=dev, dtype=torch.int64)
    lib.gem_debug_prep_diag.restype = ctypes.c_int
    lib.gem_debug_prep_diag.argtypes = [ctypes.c_void_p]
    _cabi.check(lib.gem_debug_prep_diag(ctypes.c_void_p(pbuf.data_ptr())))
    qbuf = torch.zeros((1024 * 8,), device=dev, dtype=torch.int64)
    lib.gem_debug_q_diag.restype = ctypes.c_int
    lib.gem_debug_q_diag.argtypes = [ctypes.c_void_p]
    _cabi.check(lib.gem_debug_q_diag(ctypes.c_void_p(qbuf.data_ptr())))
    sbuf = torch.zeros((1024 * 8,), device=dev, dtype=torch.int64)
    lib.gem_debug_sel_diag.restype = ctypes.c_int
    lib.gem_debug_sel_diag.argtypes = [ctypes.c_void_p]
    _cabi.check(lib.gem_debug_sel_diag(ctypes.c_void_p(sbuf.data_ptr())))
    rows = []
    for _ in range(5):
        sbuf.zero_()
        buf.zero_()
        pbuf.zero_()
        qbuf.zero_()
        pos_before = emb._pos.clone()
        emb.run_layout_device(1)
        torch.cuda.synchronize()
        raw = buf.cpu().numpy()
        v = raw[: ncta * 12].reshape(ncta, 12).astype(np.float64)
        live = v[:, 0] > 0
        v = v[live]
        rows.append(v)
        wd = raw[ncta * 12:].reshape(ncta, 16, 8)[live].astype(np.float64)      # per warp slow-path counters
    _cabi.check(lib.gem_debug_scan_diag(None))
    _cabi.check(lib.gem_debug_prep_diag(None))
    _cabi.check(lib.gem_debug_q_diag(None))
    _cabi.check(lib.gem_debug_sel_diag(None))
    sv = sbuf.cpu().numpy().reshape(1024, 8).astype(np.float64)
    sv = sv[sv[:, 0] > 0]
    if len(sv):
        s0 = sv[:, 0].min()
        f = lambda a: f"mean {a.mean() * 1e-3:.2f} p50 {np.percentile(a, 50) * 1e-3:.2f} max {a.max() * 1e-3:.2f}"
        print(f"== {wl}: select CTAs={len(sv)} stamps {stamps.get('knn_select')}; start spread {(sv[:, 0].max() - s0) * 1e-3:.1f} us; "
              f"kernel {(sv[:, 4].max() - s0) * 1e-3:.1f} us; keys per query: mean {sv[:, 5].mean():.0f} max {sv[:, 5].max():.0f}, after shrinking mean {sv[:, 6].mean():.0f} max {sv[:, 6].max():.0f}")
        print(f"   staging {f(sv[:, 1] - sv[:, 0])}; shrinking {f(sv[:, 2] - sv[:, 1])}; ranking {f(sv[:, 3] - sv[:, 2])}; intersection tail {f(sv[:, 4] - sv[:, 3])}; CTA life {f(sv[:, 4] - sv[:, 0])}")
        worst = np.argsort(-(sv[:, 4] - sv[:, 0]))[:4]
        for o in worst:
            print(f"   slowest CTA: life {(sv[o, 4] - sv[o, 0]) * 1e-3:.2f} us (start +{(sv[o, 0] - s0) * 1e-3:.1f}) n0 {int(sv[o, 5])} n {int(sv[o, 6])} staging {(sv[o, 1] - sv[o, 0]) * 1e-3:.2f} shrink {(sv[o, 2] - sv[o, 1]) * 1e-3:.2f} rank {(sv[o, 3] - sv[o, 2]) * 1e-3:.2f} tail {(sv[o, 4] - sv[o, 3]) * 1e-3:.2f}")
    qd = qbuf.cpu().numpy().reshape(1024, 8)[: w["S"]]
    tau = qd[:, 2].astype(np.uint32).view(np.float32)
    qn = qd[:, 4].astype(np.uint32).view(np.float32)
    samp = emb._buffers()["samp"].cpu().numpy()
    e32 = emb._edges32.cpu().numpy()
    P = pos_before.cpu().numpy()
    qd = qd.copy()
    qd[:, 0] += qd[:, 1]                  # exact evaluations = rejected at the bound + accepted
    print(f"== {wl}: per query: exact evaluations mean {qd[:, 0].mean():.0f} p50 {np.percentile(qd[:, 0], 50):.0f} max {qd[:, 0].max()}; "
          f"accepted mean {qd[:, 1].mean():.0f} max {qd[:, 1].max()}")
    deg = np.diff(emb._row_ptr.cpu().numpy())
    for q in np.argsort(-qd[:, 0])[:6]:
        msg = f"   query {q:3d}: passed {qd[q, 0]:6d} accepted {qd[q, 1]:6d} tau {tau[q]:.4e} |q| {np.sqrt(max(qn[q], 0)):.3f} edge id {samp[q]}"
        if e32 is not None and samp[q] < len(e32):
            u, v_ = e32[samp[q]]
            msg += f" = ({u},{v_})"
            if deg is not None:
                msg += f" degrees ({deg[u]},{deg[v_]})"
            msg += f" |p_u - p_v| {np.linalg.norm(P[u] - P[v_]):.3e}"
        print(msg)
    print(f"   position std per column {P[: emb.n].std(axis=0)}; tau percentiles p5 {np.percentile(tau, 5):.3e} p50 {np.percentile(tau, 50):.3e} p95 {np.percentile(tau, 95):.3e}")
    pv = pbuf.cpu().numpy().reshape(1200, 12).astype(np.float64)
    pv = pv[pv[:, 0] > 0]
    if len(pv):
        p0 = pv[:, 0].min()
        bound = pv[:, 1] > 0
        last = pv[:, 6] > 0
        print(f"== {wl}: prep CTAs={len(pv)} (bound {bound.sum()}, hint {(~bound).sum()}); stamps {stamps.get('knn_prep')}")
        print(f"   CTA start spread: bound {(pv[bound, 0].max() - p0) * 1e-3:.1f} us, hint {((pv[~bound, 0].max() - p0) * 1e-3) if (~bound).any() else 0:.1f} us")
        print(f"   bound CTA: query table mean {((pv[bound, 1] - pv[bound, 0]).mean()) * 1e-3:.2f} max {((pv[bound, 1] - pv[bound, 0]).max()) * 1e-3:.2f}; "
              f"bound pass mean {((pv[bound, 2] - pv[bound, 1]).mean()) * 1e-3:.2f} max {((pv[bound, 2] - pv[bound, 1]).max()) * 1e-3:.2f}; "
              f"last bound CTA done at {(pv[bound, 2].max() - p0) * 1e-3:.1f} us")
        if (~bound).any():
            print(f"   hint CTA: life mean {((pv[~bound, 2] - pv[~bound, 0]).mean()) * 1e-3:.2f} max {((pv[~bound, 2] - pv[~bound, 0]).max()) * 1e-3:.2f}; "
                  f"last hint CTA done at {(pv[~bound, 2].max() - p0) * 1e-3:.1f} us")
        print(f"   fence+ticket mean {((pv[:, 3] - pv[:, 2]).mean()) * 1e-3:.2f} max {((pv[:, 3] - pv[:, 2]).max()) * 1e-3:.2f}")
        if last.any():
            lv = pv[last][0]
            print(f"   last CTA ({'bound' if lv[1] > 0 else 'hint'}): ticket at {(lv[3] - p0) * 1e-3:.1f} us, thresholds+coefficients {(lv[4] - lv[3]) * 1e-3:.2f} us, "
                  f"clear+bump {(lv[5] - lv[4]) * 1e-3:.2f} us, end at {(lv[5] - p0) * 1e-3:.1f} us")
    v = rows[-1]
    t0 = v[:, 0].min()
    us = lambda a: (a * 1e-3)
    two = v[:, 11] >= 2          # CTAs whose warp 0 processed at least 2 blocks
    one = v[:, 11] >= 1
    print(f"== {wl}: E={emb.n_edges} scan CTAs={len(v)}  stamps {stamps.get('knn_scan')}  "
          f"blocks of warp 0: mean {v[:, 11].mean():.2f} max {v[:, 11].max():.0f}")
    print(f"   CTA start spread {us(v[:, 0].max() - t0):.1f} us; kernel (first start -> last end) {us(v[:, 7].max() - t0):.1f} us; "
          f"CTA life mean {us((v[:, 7] - v[:, 0]).mean()):.1f} max {us((v[:, 7] - v[:, 0]).max()):.1f}")

    def line(name, a):
        if len(a):
            print(f"   {name:24s} mean {us(a.mean()):7.2f}  max {us(a.max()):7.2f}  (n={len(a)})")
    line(PH[0], v[:, 1] - v[:, 0])
    line(PH[1], v[:, 2] - v[:, 1])
    line(PH[2], (v[:, 3] - v[:, 2])[one])
    line(PH[3], (v[:, 4] - v[:, 3])[one])
    line(PH[4], (v[:, 8] - v[:, 4])[one])
    line(PH[5], (v[:, 9] - v[:, 8])[two])
    line(PH[6], (v[:, 10] - v[:, 9])[two])
    last = np.where(two, v[:, 10], np.where(one, v[:, 4], v[:, 2]))
    line(PH[7], v[:, 5] - last)
    line(PH[8], v[:, 6] - v[:, 5])
    line(PH[9], v[:, 7] - v[:, 6])
    W = wd.reshape(-1, 8)
    ev, tm = W[:, 3], W[:, 4] * 1e-3
    pct = lambda a, q: float(np.percentile(a, q))
    print(f"   slow path per WARP: events mean {ev.mean():.1f} p50 {pct(ev, 50):.0f} p99 {pct(ev, 99):.0f} max {ev.max():.0f}; "
          f"time us mean {tm.mean():.2f} p50 {pct(tm, 50):.2f} p99 {pct(tm, 99):.2f} max {tm.max():.2f}; "
          f"us/event {tm.sum() / max(ev.sum(), 1):.3f}")
    print(f"   totals: events {ev.sum():.0f}, exact checks rejected {W[:, 0].sum():.0f} accepted {W[:, 1].sum():.0f} inserted {W[:, 2].sum():.0f}")
    cta_t = wd[:, :, 4].sum(axis=1) * 1e-3
    print(f"   slow-path time per CTA (sum over its warps, us): mean {cta_t.mean():.1f} p50 {pct(cta_t, 50):.1f} max {cta_t.max():.1f}")
    order = np.argsort(-W[:, 5])[:8]
    for o in order:
        r1, r2 = int(W[o, 1]), int(W[o, 2])
        print(f"   longest single call: {W[o, 5] * 1e-3:8.2f} us  block {int(W[o, 6]):6d}  events {int(W[o, 7]):4d}  (cta {o // 16}, warp {o % 16}; warp total {tm[o]:.1f} us, {int(ev[o])} events)"
              f"  resolve {W[o, 0] * 1e-3:.2f} us in {r2 & 0xffffffff} calls, {r2 >> 32} exact evaluations, {r1 >> 32} locked lanes, {r1 & 0xffffffff} lock spins")
    life = (v[:, 7] - v[:, 0]) * 1e-3
    print(f"   CTA life percentiles us: p5 {pct(life, 5):.1f} p50 {pct(life, 50):.1f} p95 {pct(life, 95):.1f} max {life.max():.1f}; "
          f"loop exit of warp 0 (since CTA start): p50 {pct((v[:, 5] - v[:, 0]) * 1e-3, 50):.1f} max {((v[:, 5] - v[:, 0]) * 1e-3).max():.1f}")
    emb.close()


def main():
    wls = sys.argv[2:] or ["c1", "c2", "ba125000", "c3"]
    lib = _cabi.load()
    _cabi.init_device(0)
    for wl in wls:
        run(wl, lib)


if __name__ == "__main__":
    main()
