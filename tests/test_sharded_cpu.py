"""CPU tests of the multi-GPU host logic: the vertex partition (partition.py) and the sharded
iteration's orchestration (sharded.ShardedLayoutEngine) under gloo with world_size 2 and 3, with
the compute stages bound to the oracle instead of the CUDA library (the engine is device-agnostic).
What is checked: every rank ends each iteration with the same replicated positions, and they are
the single-process oracle's (neighbour lists bit-exact, positions within 1e-5)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from graphem_rapids_b200 import generators as gen                      # noqa: E402
from graphem_rapids_b200.partition import (balanced_vertex_ranges, build_layout, edges_sorted_by_ij)  # noqa: E402
from oracle import oracle                                             # noqa: E402


def _edges(adj):
    return oracle.extract_edges(adj).astype(np.int64)


# ----------------------------------------------------------------------------- partition
@pytest.mark.parametrize("ownership", ["strided", "contiguous"])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("kind", ["ba", "rr", "er"])
def test_layout_invariants(world, kind, ownership):
    n = 3000
    adj = {"ba": lambda: gen.generate_ba(n, 4, seed=0), "rr": lambda: gen.generate_random_regular(n, 6, seed=0),
           "er": lambda: gen.erdos_renyi_graph(n, 8.0 / n, seed=0)}[kind]()
    e = _edges(adj)
    L = build_layout(e, n, world, hub_degree=32, ownership=ownership)
    assert L.sorted_edges and L.n_pad == world * L.slice and L.n_edges == len(e)
    assert int(L.rank_count.sum()) == n and np.all(L.rank_count >= 1)
    # every vertex has exactly one padded row, inside the valid part of its rank's block
    assert len(np.unique(L.pad_of)) == n
    owner = L.pad_of // L.slice
    assert np.all(L.pad_of - owner * L.slice < L.rank_count[owner])
    if world == 1 or ownership == "contiguous":
        assert np.all(np.diff(L.pad_of) > 0) and L.edge_orig is None      # monotonic numbering
        assert L.v_lo[0] == 0 and L.v_hi[-1] == n and np.all(L.v_lo[1:] == L.v_hi[:-1])
    else:
        assert np.array_equal(owner, np.arange(n) % world)                # v mod G
        assert L.rank_count.max() - L.rank_count.min() <= 1               # balanced rows for ANY vertex order
        own_edges = L.e_hi - L.e_lo
        assert own_edges.max() <= 1.35 * own_edges.mean() + 50            # and (statistically) balanced edges
    deg = np.diff(L.row_ptr)
    upc = np.diff(L.up_ptr)
    for r in range(world):
        b, t = L.rank_rows(r)
        assert np.array_equal(L.hubs[r], np.nonzero(deg[b:t] > 32)[0] + b)
        ids = L.local_edge_ids(r)                                         # the rank owns the edges whose first endpoint it owns
        assert np.all(L.pad_of[e[ids, 0]] // L.slice == r) and len(ids) == L.e_hi[r] - L.e_lo[r]
        assert np.all(np.diff(ids) > 0)                                   # local order == original order restricted
    assert L.e_lo[0] == 0 and L.e_hi[-1] == len(e) and np.all(L.e_lo[1:] == L.e_hi[:-1])
    all_ids = np.concatenate([L.local_edge_ids(r) for r in range(world)])
    assert np.array_equal(np.sort(all_ids), np.arange(len(e)))
    # edges32 = padded endpoints by ORIGINAL edge id; the last up(v) entries of row v are its owned edges, in
    # edge-list order: enumerating them over the padded rows gives the local-order edge numbering
    assert np.array_equal(L.edges32, L.pad_of[e].astype(np.int32))
    assert len(L.col) == 2 * len(e)
    src = np.repeat(np.arange(L.n_pad), deg)
    tpos = np.arange(len(L.col)) - np.repeat(L.row_ptr[:-1], deg)
    upper = tpos >= np.repeat(deg - upc, deg)
    assert np.array_equal(np.column_stack([src[upper], L.col[upper]]), L.edges32[all_ids])
    assert np.array_equal(upc, np.bincount(L.edges32[:, 0], minlength=L.n_pad))
    inv = np.full(L.n_pad, -1, np.int64)
    inv[L.pad_of] = np.arange(n)
    for v in np.random.default_rng(0).integers(0, L.n_pad, 50):
        row = L.col[L.row_ptr[v]:L.row_ptr[v + 1]]
        assert np.all(np.diff(inv[row]) > 0)                              # rows ordered by ORIGINAL neighbour id
    # positions round trip through the padding
    pos = np.random.default_rng(1).standard_normal((n, 3)).astype(np.float32)
    padded = L.pad_positions(pos, 4)
    assert np.array_equal(padded[L.pad_of, :3], pos) and padded.shape == (L.n_pad, 4)
    dummies = np.setdiff1d(np.arange(L.n_pad), L.pad_of)
    assert np.all(padded[dummies] == 0) and np.all(deg[dummies] == 0)


def test_balanced_ranges_cost_and_edge_cases():
    rng = np.random.default_rng(0)
    deg = rng.integers(1, 50, 10000)
    up = rng.integers(0, 25, 10000)
    lo, hi = balanced_vertex_ranges(deg, up, 8)
    cost = deg + 4.5 * up + 3.0
    per = np.array([cost[a:b].sum() for a, b in zip(lo, hi)])
    assert per.max() / per.mean() < 1.05
    # one hub holding most of the cost must not starve the other ranks of vertices
    deg = np.ones(10, np.int64); up = np.zeros(10, np.int64); up[0] = 1000
    lo, hi = balanced_vertex_ranges(deg, up, 4)
    assert np.all(hi > lo) and lo[0] == 0 and hi[-1] == 10
    with pytest.raises(ValueError):
        balanced_vertex_ranges(np.ones(2), np.ones(2), 3)
    assert edges_sorted_by_ij(np.array([[0, 1], [0, 2], [1, 2]]))
    assert not edges_sorted_by_ij(np.array([[0, 2], [0, 1]]))
    assert not edges_sorted_by_ij(np.array([[1, 0]]))


# ----------------------------------------------------------------------------- sharded engine under gloo
class OracleStages:
    """The engine's compute stages restated with the CPU oracle (row pitch = d, no padding lanes)."""

    def __init__(self, layout, d, k_attr=0.2, L_min=1.0, k_inter=0.5, seed=0):
        self.L, self.d, self.ld, self.mld = layout, d, d, d
        self.k_attr, self.L_min, self.k_inter, self.seed = k_attr, L_min, k_inter, seed
        self.device = torch.device("cpu")
        self.edges = torch.from_numpy(layout.edges32.astype(np.int64))
        self.row_ptr = torch.from_numpy(layout.row_ptr)
        self.up_ptr = torch.from_numpy(layout.up_ptr)
        self.edge_orig = None if layout.edge_orig is None else torch.from_numpy(layout.edge_orig)
        self.col = torch.from_numpy(layout.col.astype(np.int64))

    def alloc(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def sample(self, iteration, n_edges, samp):
        g = torch.Generator().manual_seed(self.seed * 1000 + iteration)
        samp.copy_(oracle.draw_sample(n_edges, samp.numel(), g))

    def spring(self, pos, vb, ve, force, mid, e_lo):
        # pull form over the CSR rows [vb, ve): F[v] = sum_w fm*((pos[w]-pos[v])/dist); the last up(v) entries of a
        # row are the edges v owns, in edge-list order
        r0, r1 = int(self.row_ptr[vb]), int(self.row_ptr[ve])
        deg = (self.row_ptr[vb + 1: ve + 1] - self.row_ptr[vb:ve])
        upc = (self.up_ptr[vb + 1: ve + 1] - self.up_ptr[vb:ve])
        src = torch.repeat_interleave(torch.arange(vb, ve), deg)
        dst = self.col[r0:r1]
        tpos = torch.arange(r1 - r0) - torch.repeat_interleave(self.row_ptr[vb:ve] - r0, deg)
        up = tpos >= torch.repeat_interleave(deg - upc, deg)
        diff = pos[dst] - pos[src]
        dd = torch.norm(diff, dim=1, keepdim=True) + 1e-6
        term = (-self.k_attr * (dd - self.L_min)) * (diff / dd)
        force.zero_()
        force.index_add_(0, src - vb, term)
        m = (pos[src[up]] + pos[dst[up]]) / 2.0
        mid[: m.shape[0]] = m
        assert m.shape[0] == mid.shape[0] - 1 and int(self.up_ptr[vb]) == e_lo

    def query_mid(self, pos, samp, qmid):
        e = self.edges[samp]
        qmid.copy_((pos[e[:, 0]] + pos[e[:, 1]]) / 2.0)

    def hint(self, pos, samp, kp1, tau_hint):
        tau_hint.fill_(float("inf"))

    def knn_local(self, mid, e_loc, e_total, e_lo, qmid, tau_hint, kp1, out_idx, out_dist):
        out_idx.fill_(-1)
        out_dist.fill_(float("inf"))
        if e_loc == 0:
            return
        assert oracle.uses_mm_mode(qmid.shape[0], e_total)
        d2 = oracle.cdist_chain_sq(qmid.numpy(), mid[:e_loc].numpy())
        dd = np.sqrt(d2)
        for q in range(qmid.shape[0]):
            order = np.lexsort((np.arange(e_loc), dd[q]))[:kp1]
            ids = torch.from_numpy(order + e_lo)                     # local-order numbers -> original edge ids
            out_idx[q, : len(order)] = ids if self.edge_orig is None else self.edge_orig[ids]
            out_dist[q, : len(order)] = torch.from_numpy(dd[q][order])

    def merge(self, g_idx, g_dist, out_idx, out_dist):
        parts, S, kp1 = g_idx.shape
        ai = g_idx.permute(1, 0, 2).reshape(S, -1).numpy()
        ad = g_dist.permute(1, 0, 2).reshape(S, -1).numpy()
        for q in range(S):
            order = np.lexsort((ai[q], ad[q]))[:kp1]
            out_idx[q] = torch.from_numpy(ai[q][order])
            out_dist[q] = torch.from_numpy(ad[q][order])

    def intersect(self, pos, samp, knn_idx, vb, ve, force):
        G = oracle.intersection_forces(pos, self.edges, knn_idx[:, 1:], samp, self.k_inter)
        force[: ve - vb] += G[vb:ve]

    def update_phase1(self, own, force, stats):
        new = own + force[: own.shape[0]]
        own.copy_(new)
        stats[: self.d] = new.double().sum(0)
        stats[self.ld: self.ld + self.d] = (new.double() ** 2).sum(0)

    def update_phase2(self, own, n_total, stats):
        nn = float(n_total)
        mean = stats[: self.d] / nn
        var = (stats[self.ld: self.ld + self.d] - nn * mean * mean) / (nn - 1.0)
        sd = var.clamp_min(0).sqrt().float() + 1e-6
        own.copy_((own - mean.float()) / sd)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, n, d, k, S, steps, out, ownership="strided"):
    from graphem_rapids_b200.sharded import ShardedLayoutEngine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        adj = gen.generate_ba(n, 3, seed=2) if kind == "ba" else gen.generate_random_regular(n, 6, seed=2)
        e = _edges(adj)
        L = build_layout(e, n, world, hub_degree=16, ownership=ownership)
        st = OracleStages(L, d, seed=5)
        eng = ShardedLayoutEngine(L, rank, st, n_components=d, n_neighbors=k, sample_size=S, inplace_allgather=False)
        pos0 = torch.from_numpy((np.random.default_rng(3).standard_normal((n, d)) * 0.5).astype(np.float32))
        eng.set_positions(pos0)
        ref = pos0.clone()
        edges = torch.from_numpy(e)
        ok = True
        worst = 0.0
        for it in range(steps):
            eng.step()
            samp = eng.samp.clone()
            o = oracle.layout_step(ref, edges, samp, n_neighbors=k, strict=True)
            ok &= bool(torch.equal(eng.knn_idx, o["knn_full"])) and bool(torch.equal(eng.knn_dist, o["knn_dist"]))
            got = eng.get_positions()
            err = float((got - o["new_pos"]).abs().max() / o["new_pos"].abs().max())
            worst = max(worst, err)
            ref = got.clone()                       # follow the sharded trajectory; compare step by step
        # replicated state identical on every rank, dummy rows untouched
        gathered = [torch.empty_like(eng.pos) for _ in range(world)]
        dist.all_gather(gathered, eng.pos)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        dummies = np.setdiff1d(np.arange(L.n_pad), L.pad_of)
        clean = bool((eng.pos[torch.from_numpy(dummies)] == 0).all()) if len(dummies) else True
        if rank == 0:
            out.put((ok, worst, same, clean))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind,n,d,k,ownership", [
    (2, "ba", 400, 3, 10, "strided"), (3, "rr", 400, 2, 5, "strided"), (2, "rr", 400, 3, 40, "contiguous"),
    (3, "rr", 60, 2, 45, "strided"),                  # shards hold fewer than k+1 edges
    (3, "ba", 401, 3, 10, "strided"), (2, "ba", 400, 3, 10, "contiguous")])
def test_sharded_engine_matches_single_process_oracle(world, kind, n, d, k, ownership):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    S, steps = 48, 3
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, n, d, k, S, steps, out, ownership))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    ok, worst, same, clean = out.get()
    assert ok, "merged neighbour lists differ from the single-process oracle"
    assert worst <= 1e-5, worst
    assert same and clean
