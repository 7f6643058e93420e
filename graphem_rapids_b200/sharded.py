"""Multi-GPU layout iteration: one process per GPU, vertex sharding.

Every rank holds the replicated positions and graph arrays and OWNS the vertices v with
v mod G == rank (partition.py).  Because the spring stage is vertex-parallel ("pull" over the CSR), a
rank produces the COMPLETE force of each vertex it owns -- there is no per-vertex force reduction
across ranks -- and the midpoints of the edges whose first endpoint it owns, which are its share
of the KNN candidates.

Product flow (CUDA stages + symmetric memory, `CudaStages.p2p_*`), per iteration and rank:

    side stream : KNN preparation on the WHOLE edge list (one launch: sample, query midpoints, bounds,
                  thresholds -- identical on every rank, no exchange)      | column sums of the own rows
    main stream : spring kernel: pos + F_spring of the own rows is stored into EVERY rank's raw buffer
                  (NVLink P2P stores: the position exchange starts here and runs underneath the scan)
                  -> scan of the own candidates -> select, which stores the partial lists into every
                  rank's exchange buffer                                                  == barrier A
                  -> merge + intersection forces into the own rows; the last CTA re-publishes the touched
                  rows and the rank's column sums to every rank                           == barrier B
                  -> every rank normalises ALL rows locally (raw -> pos)

Two device-side barriers per iteration (round 1: three, plus a position push behind the last one), no
NCCL collective, 8 kernels.  The raw buffers and the exchange area are double-buffered by iteration
parity (a fast rank may start iteration t+1 while a slow one still reads iteration t's).

Fallback flow (`ShardedLayoutEngine.step` without P2P stages: NCCL all-gather / all-reduce / all-gather)
is what the CPU tests drive under gloo with the stages bound to the oracle, and what runs when
symmetric memory is unavailable or a shard is too small for the fast KNN path.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi
from .embedder import GraphEmbedderPyTorch, _ptr
from .partition import GraphLayout


class ShardedLayoutEngine:
    """One `update_positions` (embedder_pytorch.py:776-806) across the ranks of `group`."""

    def __init__(self, layout: GraphLayout, rank: int, stages, *, n_components: int, n_neighbors: int,
                 sample_size: int, group=None, inplace_allgather: bool = True,
                 pos: Optional[torch.Tensor] = None):
        self.L, self.rank, self.st, self.group = layout, rank, stages, group
        self.world = layout.world
        self.d = int(n_components)
        self.kp1 = int(n_neighbors) + 1
        self.S = min(int(sample_size), layout.n_edges)
        self.inplace = inplace_allgather
        ld, mld = stages.ld, stages.mld
        self.vb, self.ve = layout.rank_rows(rank)                  # valid rows of this rank (padded numbering)
        self.e_lo, self.e_hi = int(layout.e_lo[rank]), int(layout.e_hi[rank])
        S, kp1 = max(self.S, 1), self.kp1
        a = stages.alloc
        self.pos = pos if pos is not None else a((layout.n_pad, ld), torch.float32)
        self.force = a((layout.slice, ld), torch.float32)
        self.mid = a((self.e_hi - self.e_lo + 1, mld), torch.float32)
        self.qmid = a((S, mld), torch.float32)
        self.tau_hint = a((S,), torch.float32)
        self.samp = a((S,), torch.int64)
        self.knn_idx = a((S, kp1), torch.int64)
        self.knn_dist = a((S, kp1), torch.float32)
        # packed partial list of one rank: [idx int64 S*kp1 | dist fp32 S*kp1 (+pad to 8 bytes)]
        self._ib = S * kp1 * 8
        self._nb = (self._ib + S * kp1 * 4 + 15) // 16 * 16           # 16-byte granules (P2P push kernel)
        self.part = a((self._nb,), torch.uint8)
        self.gathered = a((self.world, self._nb), torch.uint8)
        self.part_idx = self.part[: self._ib].view(torch.int64).view(S, kp1)
        self.part_dist = self.part[self._ib: self._ib + S * kp1 * 4].view(torch.float32).view(S, kp1)
        self.g_idx = self.gathered[:, : self._ib].view(torch.int64).view(self.world, S, kp1)
        self.g_dist = self.gathered[:, self._ib: self._ib + S * kp1 * 4].view(torch.float32).view(self.world, S, kp1)
        self.stats = a((2 * ld,), torch.float64)                   # column sums | sums of squares
        self.iteration = 0

    def bind_exchange(self, gathered: torch.Tensor):
        """Use `gathered` ((world, nb) uint8, e.g. a view of a symmetric-memory buffer that the peers
        store into) as the landing zone of the partial lists."""
        S, kp1 = max(self.S, 1), self.kp1
        assert gathered.shape == (self.world, self._nb) and gathered.dtype == torch.uint8
        self.gathered = gathered
        self.g_idx = gathered[:, : self._ib].view(torch.int64).view(self.world, S, kp1)
        self.g_dist = gathered[:, self._ib: self._ib + S * kp1 * 4].view(torch.float32).view(self.world, S, kp1)

    # replicated state in / out (original vertex numbering on the host side)
    def set_positions(self, pos_nd: torch.Tensor):
        self.pos.zero_()
        idx = torch.from_numpy(self.L.pad_of).to(self.pos.device)
        self.pos[idx, : self.d] = pos_nd.to(device=self.pos.device, dtype=torch.float32)

    def get_positions(self) -> torch.Tensor:
        idx = torch.from_numpy(self.L.pad_of).to(self.pos.device)
        return self.pos[idx][:, : self.d]

    # The iteration is three local phases separated by the three exchanges; `step` runs them with
    # torch.distributed, the single-GPU tests drive several engines ("virtual ranks") phase by phase.
    def phase_a(self, sampled_indices: Optional[torch.Tensor] = None):
        """sample -> spring + midpoints of the owned rows -> shard-local KNN (fills self.part)."""
        st, L = self.st, self.L
        if L.n_edges == 0 or self.kp1 > L.n_edges:
            raise RuntimeError("selected index k out of range")       # torch.topk in the reference (:583)
        if sampled_indices is not None:
            self.samp.copy_(sampled_indices.to(self.samp.device))
        if sampled_indices is None:
            st.sample(self.iteration, L.n_edges, self.samp)            # same ids on every rank
        self.iteration += 1
        # (a) complete spring forces of the owned vertices + midpoints of the owned edges
        st.spring(self.pos, self.vb, self.ve, self.force, self.mid, self.e_lo)
        # (b) KNN of the S query midpoints among the owned candidates
        st.query_mid(self.pos, self.samp, self.qmid)
        st.hint(self.pos, self.samp, self.kp1, self.tau_hint)
        st.knn_local(self.mid, self.e_hi - self.e_lo, L.n_edges, self.e_lo, self.qmid, self.tau_hint, self.kp1,
                     self.part_idx, self.part_dist)

    def phase_b(self):
        """merge the gathered partial lists -> intersection forces into the owned rows -> update pass 1."""
        st = self.st
        st.merge(self.g_idx, self.g_dist, self.knn_idx, self.knn_dist)
        # (c) every rank evaluates the <= S*k pairs, accumulates only into its own vertex rows
        st.intersect(self.pos, self.samp, self.knn_idx, self.vb, self.ve, self.force)
        # (d) update of the owned rows around a global reduction of the column sums
        st.update_phase1(self.pos[self.vb: self.ve], self.force, self.stats)

    def phase_c(self):
        """update pass 2 (normalise the owned rows with the reduced column sums)."""
        self.st.update_phase2(self.pos[self.vb: self.ve], self.L.n, self.stats)

    def own_block(self) -> torch.Tensor:
        """The rank's block of the position buffer (valid rows + dummies): its all-gather contribution."""
        return self.pos[self.rank * self.L.slice: (self.rank + 1) * self.L.slice]

    def step(self, sampled_indices: Optional[torch.Tensor] = None):
        st = self.st
        if getattr(st, "p2p_ready", False):
            # product flow: early position push + two barriers (module docstring)
            st.p2p_phase1(self, sampled_indices)
            st.barrier(0)
            st.p2p_phase2(self)
            st.barrier(1)
            st.p2p_phase3(self)
            self.iteration += 1
            return
        # fallback flow: three local phases separated by NCCL / gloo collectives
        self.phase_a(sampled_indices)
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered.view(-1), self.part, group=self.group)
        else:
            self.gathered.view(-1).copy_(self.part)
        self.phase_b()
        if self.world > 1:
            dist.all_reduce(self.stats, group=self.group)
        self.phase_c()
        if self.world > 1:
            block = self.own_block()
            dist.all_gather_into_tensor(self.pos.view(-1), (block if self.inplace else block.clone()).view(-1),
                                        group=self.group)


class CudaStages:
    """The compute stages of the engine on the CUDA C ABI (include/graphem_b200.h) for one rank of
    a GraphLayout.  `arrays` lets the owner share graph tensors it has already uploaded.

    Two sets of methods: the stage-by-stage ones (sample / spring / query_mid / hint / knn_local / merge /
    intersect / update_phase1 / update_phase2) serve the fallback flow of ShardedLayoutEngine (NCCL collectives
    between them) and the single-GPU "virtual rank" tests; p2p_phase1/2/3 + barrier are the product flow over
    peer-mapped buffers (attach_p2p)."""

    def __init__(self, layout: GraphLayout, rank: int, device, *, n_components: int, k_attr: float, L_min: float,
                 k_inter: float, seed: int, arrays: Optional[dict] = None, coef_slot: Optional[int] = None):
        self.L, self.rank = layout, rank
        self.device = torch.device(device)
        self.d = int(n_components)
        self.k_attr, self.L_min, self.k_inter = float(k_attr), float(L_min), float(k_inter)
        self.seed = int(seed) & (2 ** 64 - 1)
        self.lib = _cabi.load()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _cabi.init_device(dev_index)
        self.ld, self.mld = self.lib.gem_row_pitch(self.d), self.lib.gem_mid_pitch(self.d)
        if not layout.sorted_edges or self.d not in (2, 3):
            raise NotImplementedError("the multi-GPU path needs an (i,j)-sorted edge list and n_components in {2,3}")
        if coef_slot is None:                                         # stand-alone use (tests): own a slot
            from .embedder import _acquire_coef_slot, _release_coef_slot
            import weakref
            coef_slot = _acquire_coef_slot(self.lib, dev_index)
            self._slot_finalizer = weakref.finalize(self, _release_coef_slot, self.lib, dev_index, coef_slot)
        self.coef_slot = int(coef_slot)
        up = lambda x: torch.from_numpy(x).to(self.device)           # noqa: E731
        a = arrays or {}
        self.edges32 = a["edges32"] if "edges32" in a else up(layout.edges32).contiguous()
        self.row_ptr = a["row_ptr"] if "row_ptr" in a else up(layout.row_ptr)
        self.col = a["col"] if "col" in a else up(layout.col)
        self.up_ptr = a["up_ptr"] if "up_ptr" in a else up(layout.up_ptr)
        self.hubs = up(layout.hubs[rank])
        # the rank's own edges in the order its spring kernel writes their midpoints (= its KNN candidates);
        # l2g maps that numbering back to original edge ids (None: a contiguous slice of the edge list)
        own = layout.local_edge_ids(rank)
        self.l2g = None if layout.edge_orig is None else up(np.ascontiguousarray(own))
        self._iter = torch.zeros((1,), device=self.device, dtype=torch.int64)   # device-side iteration counter
        self._knn_ws = None
        self._stats_ws = None
        self._side = torch.cuda.Stream(device=self.device)
        self._fork = torch.cuda.Event()
        self._join = torch.cuda.Event()
        self._spring_done = torch.cuda.Event()
        self._stats_done = torch.cuda.Event()

    # ------------------------------------------------------------------ common helpers
    def alloc(self, shape, dtype):
        return torch.zeros(shape, device=self.device, dtype=dtype)

    def _s(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _side_ptr(self):
        return ctypes.c_void_p(self._side.cuda_stream)

    def _hub_args(self):
        return (_ptr(self.hubs) if self.hubs.numel() else None), int(self.hubs.numel())

    def _knn_workspace(self, e_loc, S, kp1):
        if self._knn_ws is None:
            nbytes = ctypes.c_size_t(0)
            _cabi.check(self.lib.gem_knn_workspace_bytes(max(e_loc, 1), self.d, S, kp1, ctypes.byref(nbytes)))
            self._knn_ws = torch.zeros((nbytes.value + 256,), device=self.device, dtype=torch.uint8)
            self._knn_ws_bytes = nbytes.value
        return self._knn_ws

    def _ws(self, n_rows):
        if self._stats_ws is None:
            nbytes = ctypes.c_size_t(0)
            _cabi.check(self.lib.gem_update_workspace_bytes(max(n_rows, 1), self.d, ctypes.byref(nbytes)))
            self._stats_ws = torch.zeros((nbytes.value + 256,), device=self.device, dtype=torch.uint8)
        return self._stats_ws

    # ------------------------------------------------------------------ optional phase marks (profile_phases)
    _marks = None

    def _mark(self, name, side=False):
        if self._marks is not None:
            if name not in self._marks:      # external events only exist for capture (timing nodes inside a CUDA graph)
                self._marks[name] = torch.cuda.Event(enable_timing=True, external=torch.cuda.is_current_stream_capturing())
            self._marks[name].record(self._side if side else torch.cuda.current_stream(self.device))

    # ------------------------------------------------------------------ product flow over peer-mapped buffers
    p2p_ready = False
    peer_ptrs = None            # kept for introspection: device pointers of every rank's position buffer

    def attach_p2p(self, pos_ptrs, raw_ptrs, xchg_ptrs, raw_local: torch.Tensor, xchg_local: torch.Tensor,
                   list_bytes: int, S: int, kp1: int, barrier, raw_multicast_ptr: int = 0):
        """pos_ptrs / raw_ptrs / xchg_ptrs: base device pointers of EVERY rank's position buffer (n_pad, ld), raw buffer
        (2, n_pad, ld) and exchange area (2 x [world x list_bytes | world x 2*ld doubles]) as seen from this process
        (torch.distributed._symmetric_memory buffer_ptrs, or plain pointers of sibling engines on one GPU);
        raw_local / xchg_local: this rank's own tensors; barrier(channel): cross-rank barrier on the current stream."""
        world = len(pos_ptrs)
        assert len(raw_ptrs) == world and len(xchg_ptrs) == world == self.L.world
        self.world = world
        self.peer_ptrs = (ctypes.c_void_p * world)(*[int(p) for p in pos_ptrs])
        self._raw_local, self._xchg_local = raw_local, xchg_local
        self._list_bytes = int(list_bytes)
        self._ib = S * kp1 * 8
        self._stats_bytes = world * 2 * self.ld * 8
        self._parity_bytes = (world * self._list_bytes + self._stats_bytes + 255) // 256 * 256
        raw_stride = self.L.n_pad * self.ld * 4
        self._raw_peers = [(ctypes.c_void_p * world)(*[int(p) + par * raw_stride for p in raw_ptrs]) for par in (0, 1)]
        # NVSwitch multicast mapping of the raw buffers (0: none -> world-1 unicast stores per row)
        self._raw_mc = [int(raw_multicast_ptr) + par * raw_stride for par in (0, 1)] if raw_multicast_ptr else None
        self._xchg_peers = (ctypes.c_void_p * world)(*[int(p) for p in xchg_ptrs])
        self._barrier = barrier
        self._touched = torch.zeros((max(4 * S * max(kp1 - 1, 1), 1),), device=self.device, dtype=torch.int32)
        self._counters = torch.zeros((2,), device=self.device, dtype=torch.int32)
        self._spring_work = torch.zeros((2,), device=self.device, dtype=torch.int32)
        self.p2p_ready = True

    @staticmethod
    def exchange_bytes(world: int, list_bytes: int, ld: int) -> int:
        return 2 * ((world * list_bytes + world * 2 * ld * 8 + 255) // 256 * 256)

    def barrier(self, channel):
        self._barrier(channel)
        self._mark("barrier_A" if channel == 0 else "barrier_B")

    def p2p_phase1(self, eng, sampled_indices=None):
        """KNN preparation (side stream) | spring kernel with the raw-row push -> column sums (side) -> scan + select
        publishing the partial lists."""
        L, lib = self.L, self.lib
        S, kp1 = max(eng.S, 1), eng.kp1
        if L.n_edges == 0 or kp1 > L.n_edges:
            raise RuntimeError("selected index k out of range")       # torch.topk in the reference (:583)
        par = eng.iteration & 1
        e_loc = eng.e_hi - eng.e_lo
        main = torch.cuda.current_stream(self.device)
        if sampled_indices is not None:
            eng.samp.copy_(sampled_indices.to(eng.samp.device))
        ws = self._knn_workspace(e_loc, S, kp1)
        self._mark("start")
        self._fork.record(main)
        self._side.wait_event(self._fork)
        a = _cabi.GemKnnPrepArgs()
        a.d, a.kp1, a.s, a.e = self.d, kp1, S, e_loc
        a.pos, a.edges, a.e_total = eng.pos.data_ptr(), self.edges32.data_ptr(), L.n_edges
        a.samp = eng.samp.data_ptr()
        a.draw = 0 if sampled_indices is not None else 1
        a.bump = a.draw
        a.seed, a.iter_counter = self.seed, self._iter.data_ptr()
        a.qmid_out = eng.qmid.data_ptr()
        a.row_ptr, a.col, a.tau_hint_out = self.row_ptr.data_ptr(), self.col.data_ptr(), eng.tau_hint.data_ptr()
        # bound with the WHOLE (replicated) edge list: the same global thresholds on every rank, no exchange
        # (the balanced sample size: measured on 2 and 8 B200 the preparation then ends with the NVLink-bound spring kernel)
        a.bound_edges, a.e_bound, a.bound_samples = self.edges32.data_ptr(), L.n_edges, 0
        a.coef_slot = self.coef_slot
        a.ws, a.ws_bytes = ws.data_ptr(), self._knn_ws_bytes
        _cabi.check(lib.gem_knn_prep(ctypes.byref(a), self._side_ptr()), "gem_knn_prep")
        self._mark("prep", side=True)
        self._join.record(self._side)
        hubs, n_hubs = self._hub_args()
        if eng.ve > eng.vb:
            _cabi.check(lib.gem_spring_update_csr_push(_ptr(eng.pos), _ptr(self.row_ptr), _ptr(self.col), _ptr(self.up_ptr),
                                                       eng.vb, eng.ve, hubs, n_hubs, self.d, self.k_attr, self.L_min,
                                                       self._raw_peers[par], self.world, self.rank,
                                                       ctypes.c_void_p(self._raw_mc[par]) if self._raw_mc else None,
                                                       _ptr(eng.mid), eng.e_lo, _ptr(self._spring_work), self._s()),
                        "gem_spring_update_csr_push")
        self._mark("spring")
        self._spring_done.record(main)
        # column sums of the own new rows: a read-only pass on the side stream, next to the scan
        self._side.wait_event(self._spring_done)
        sws = self._ws(eng.ve - eng.vb)
        own_raw = self._raw_local[par, eng.vb: eng.ve]
        if eng.ve > eng.vb:
            _cabi.check(lib.gem_update_positions(_ptr(own_raw), None, None, eng.ve - eng.vb, eng.ve - eng.vb, self.d,
                                                 _ptr(sws), 3, self._side_ptr()), "gem_update_positions(phase 3)")
        self._mark("colsum", side=True)
        self._stats_done.record(self._side)
        main.wait_event(self._join)
        pub = _cabi.GemKnnPublish()
        pub.remap = self.l2g.data_ptr() if self.l2g is not None else None
        pub.peer_base_host, pub.world = self._xchg_peers, self.world
        base = par * self._parity_bytes + self.rank * self._list_bytes
        pub.idx_offset_bytes, pub.dist_offset_bytes = base, base + self._ib
        off = eng.e_lo if self.l2g is None else 0
        _cabi.check(lib.gem_knn_scan(_ptr(eng.mid), e_loc, off, self.d, _ptr(eng.qmid), S, kp1, _ptr(eng.part_idx),
                                     _ptr(eng.part_dist), _ptr(ws), self._knn_ws_bytes, self.coef_slot, ctypes.byref(pub),
                                     self._s()), "gem_knn_scan")
        self._mark("scan_select")

    def _parity_views(self, eng, par):
        S, kp1 = max(eng.S, 1), eng.kp1
        blk = self._xchg_local[par * self._parity_bytes: (par + 1) * self._parity_bytes]
        lists = blk[: self.world * self._list_bytes].view(self.world, self._list_bytes)
        g_idx = lists[:, : self._ib].view(torch.int64).view(self.world, S, kp1)
        g_dist = lists[:, self._ib: self._ib + S * kp1 * 4].view(torch.float32).view(self.world, S, kp1)
        stats = blk[self.world * self._list_bytes: self.world * self._list_bytes + self._stats_bytes]
        return g_idx, g_dist, stats

    def p2p_phase2(self, eng):
        """merge the partial lists of all ranks + intersection forces into the own raw rows; the last CTA
        re-publishes the touched rows and the rank's column sums."""
        par = eng.iteration & 1
        S, kp1 = max(eng.S, 1), eng.kp1
        g_idx, g_dist, _ = self._parity_views(eng, par)
        torch.cuda.current_stream(self.device).wait_event(self._stats_done)     # the sums it corrects must be there
        pub = _cabi.GemMergePublish()
        pub.peer_raw_host, pub.peer_xchg_host = self._raw_peers[par], self._xchg_peers
        pub.world, pub.rank = self.world, self.rank
        pub.stats_offset_bytes = par * self._parity_bytes + self.world * self._list_bytes
        pub.touched, pub.counters = self._touched.data_ptr(), self._counters.data_ptr()
        own_raw = self._raw_local[par, eng.vb: eng.ve]
        sws = self._ws(eng.ve - eng.vb)
        _cabi.check(self.lib.gem_topk_merge_intersect(_ptr(g_dist), _ptr(g_idx), g_dist.stride(0), g_idx.stride(0),
                                                      self.world, S, kp1, _ptr(eng.knn_idx), _ptr(eng.knn_dist),
                                                      _ptr(eng.pos), _ptr(self.edges32), _ptr(eng.samp), self.d,
                                                      self.k_inter, eng.vb, eng.ve, _ptr(own_raw), _ptr(sws),
                                                      ctypes.byref(pub), self._s()), "gem_topk_merge_intersect")
        self._mark("merge_intersect")

    def p2p_phase3(self, eng):
        """every rank normalises ALL rows of its replica: raw -> pos."""
        par = eng.iteration & 1
        _, _, stats = self._parity_views(eng, par)
        _cabi.check(self.lib.gem_update_normalise_all(_ptr(eng.pos), _ptr(self._raw_local[par]), self.L.n_pad, self.L.n,
                                                      self.d, _ptr(stats), self.world, self._s()),
                    "gem_update_normalise_all")
        self._mark("normalise")

    # ------------------------------------------------------------------ stage-by-stage methods (fallback flow)
    def sample(self, iteration, n_edges, samp):
        # keyed bijection from (seed, device counter): every rank starts from 0 and therefore draws the same ids
        _cabi.check(self.lib.gem_sample_edges(self.seed, _ptr(self._iter), 1, n_edges, samp.numel(), _ptr(samp), self._s()),
                    "gem_sample_edges")

    def spring(self, pos, vb, ve, force, mid, e_lo):
        hubs, n_hubs = self._hub_args()
        _cabi.check(self.lib.gem_spring_midpoints_csr(
            _ptr(pos), _ptr(self.row_ptr), _ptr(self.col), _ptr(self.up_ptr), vb, ve, hubs, n_hubs, self.d, self.k_attr,
            self.L_min, _ptr(force), _ptr(mid), e_lo, self._s()), "gem_spring_midpoints_csr")

    def query_mid(self, pos, samp, qmid):
        _cabi.check(self.lib.gem_query_midpoints(_ptr(pos), _ptr(self.edges32), _ptr(samp), samp.numel(), self.d,
                                                 _ptr(qmid), self._s()), "gem_query_midpoints")

    def hint(self, pos, samp, kp1, tau_hint):
        _cabi.check(self.lib.gem_knn_linegraph_hint(_ptr(pos), _ptr(self.row_ptr), _ptr(self.col), _ptr(self.edges32),
                                                    _ptr(samp), samp.numel(), self.d, kp1, _ptr(tau_hint), self._s()),
                    "gem_knn_linegraph_hint")

    def knn_local(self, mid, e_loc, e_total, e_lo, qmid, tau_hint, kp1, out_idx, out_dist):
        S = qmid.shape[0]
        ws = self._knn_workspace(e_loc, S, kp1)
        mm = 1 if (S > 25 or e_total > 25) else 0                     # torch.cdist's rule on the WHOLE problem
        off = e_lo if self.l2g is None else 0                         # strided ownership: local numbers, mapped below
        _cabi.check(self.lib.gem_knn_midpoints_shard(_ptr(mid), e_loc, e_total, off, self.d, _ptr(qmid), S, kp1, mm,
                                                     _ptr(tau_hint), _ptr(out_idx), _ptr(out_dist), _ptr(ws),
                                                     self._knn_ws_bytes, self.coef_slot, self._s()),
                    "gem_knn_midpoints_shard")
        # local-order edge numbers -> original edge ids (ties in the merge are broken by ORIGINAL index; inside a rank
        # the local order is the original order restricted to its edges, so its own top-(k+1) is unaffected)
        if self.l2g is not None:
            _cabi.check(self.lib.gem_remap_indices(_ptr(out_idx), out_idx.numel(), _ptr(self.l2g), self._s()),
                        "gem_remap_indices")

    def merge(self, g_idx, g_dist, out_idx, out_dist):
        parts, S, kp1 = g_idx.shape
        _cabi.check(self.lib.gem_topk_merge_strided(_ptr(g_dist), _ptr(g_idx), g_dist.stride(0), g_idx.stride(0), parts,
                                                    S, kp1, _ptr(out_idx), _ptr(out_dist), self._s()),
                    "gem_topk_merge_strided")

    def intersect(self, pos, samp, knn_idx, vb, ve, force):
        S, kp1 = knn_idx.shape
        if kp1 > 1 and ve > vb:
            _cabi.check(self.lib.gem_intersection_forces_range(_ptr(pos), _ptr(self.edges32), pos.shape[0], self.d,
                                                               _ptr(samp), _ptr(knn_idx), S, kp1, self.k_inter,
                                                               vb, ve, _ptr(force), self._s()),
                        "gem_intersection_forces_range")

    def update_phase1(self, own, force, stats):
        ws = self._ws(own.shape[0])
        sums = ws[: stats.numel() * 8].view(torch.float64)
        if own.shape[0] > 0:
            _cabi.check(self.lib.gem_update_positions(_ptr(own), _ptr(force), None, own.shape[0], own.shape[0], self.d,
                                                      _ptr(ws), 1, self._s()), "gem_update_positions(phase 1)")
            stats.copy_(sums)
        else:
            stats.zero_()

    def update_phase2(self, own, n_total, stats):
        ws = self._ws(own.shape[0])
        ws[: stats.numel() * 8].view(torch.float64).copy_(stats)
        if own.shape[0] > 0:
            _cabi.check(self.lib.gem_update_positions(_ptr(own), None, None, own.shape[0], n_total, self.d, _ptr(ws),
                                                      2, self._s()), "gem_update_positions(phase 2)")


class ShardedGraphEmbedder(GraphEmbedderPyTorch):
    """GraphEmbedderPyTorch across the ranks of a torch.distributed process group (one process
    per GPU, backend nccl).  Same constructor; every rank passes the same adjacency / seed and ends
    every iteration with the same replicated positions.

    sampler='device' (default) draws the same ids on every rank from the shared seed and the device-side
    iteration counter; sampler='torch' draws torch.randperm(E)[:S] like the reference on every rank from the
    (identically seeded) default generator -- the ids are then compared across ranks in debug runs only."""

    def __init__(self, adjacency, n_components=2, *args, process_group=None, use_symmetric_memory=True,
                 use_multicast=False, ownership="strided", use_unrolled_graph=None, **kwargs):
        if not dist.is_initialized():
            raise RuntimeError("ShardedGraphEmbedder needs an initialised torch.distributed process group")
        self._group = process_group
        # several iterations per captured graph (GEM_SHARDED_UNROLL=0 switches it off)
        self.use_unrolled_graph = (os.environ.get("GEM_SHARDED_UNROLL", "1") != "0") if use_unrolled_graph is None \
            else bool(use_unrolled_graph)
        self._ownership = ownership            # 'strided' (v mod G: balanced for any vertex order) | 'contiguous'
        kwargs["graph_build"] = "host"        # the vertex partition (partition.build_layout) is host work
        super().__init__(adjacency, n_components, *args, **kwargs)
        grp = self._group if self._group is not None else dist.group.WORLD
        src0 = dist.get_global_rank(self._group, 0) if self._group is not None else 0
        # every rank must draw the SAME query edges: the sampler seed comes from rank 0 (with seed=None each process
        # would otherwise take it from its own torch RNG and merge partial lists of different queries)
        seed_t = torch.tensor([self._sampler_seed], device=self.device, dtype=torch.int64)
        dist.broadcast(seed_t, src=src0, group=self._group)
        self._sampler_seed = int(seed_t.item())
        stages = CudaStages(self._layout, self._rank, self.device, n_components=self.n_components, k_attr=self.k_attr,
                            L_min=self.L_min, k_inter=self.k_inter, seed=self._sampler_seed,
                            arrays=dict(edges32=self._edges32, row_ptr=self._row_ptr, col=self._col,
                                        up_ptr=self._up_ptr), coef_slot=self._coef_slot)
        self._engine = ShardedLayoutEngine(self._layout, self._rank, stages, n_components=self.n_components,
                                           n_neighbors=self.n_neighbors, sample_size=self.sample_size,
                                           group=self._group, pos=self._pos)
        eng = self._engine
        # product flow: position / raw / exchange buffers in symmetric memory, every rank can store into every replica
        self._symm = None
        S, kp1 = max(eng.S, 1), eng.kp1
        e_loc = eng.e_hi - eng.e_lo
        want = (self._world > 1 and use_symmetric_memory
                and bool(self._lib.gem_knn_fast_path(e_loc, self.n_edges, self.n_components, S, kp1)))
        ok = torch.tensor([1 if want else 0], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self._group)          # every rank must take the same flow
        if int(ok.item()) == 1:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                pos_buf = symm_mem.empty(tuple(self._pos.shape), dtype=torch.float32, device=self.device)
                pos_buf.copy_(self._pos)
                raw = symm_mem.empty((2,) + tuple(self._pos.shape), dtype=torch.float32, device=self.device)
                raw.zero_()
                xbytes = CudaStages.exchange_bytes(self._world, eng._nb, self._ld)
                xchg = symm_mem.empty((xbytes,), dtype=torch.uint8, device=self.device)
                xchg.zero_()
                h_pos = symm_mem.rendezvous(pos_buf, grp)
                h_raw = symm_mem.rendezvous(raw, grp)
                h_x = symm_mem.rendezvous(xchg, grp)
                self._symm = dict(pos=h_pos, raw=h_raw, xchg=h_x, raw_t=raw, xchg_t=xchg)
                self._pos = pos_buf
                eng.pos = pos_buf
                mc = 0
                # EXPERIMENTAL, off by default (GEM_MULTICAST=1 or use_multicast=True): the rows of the spring kernel go out
                # as NVSwitch multicast stores (multimem.st.relaxed.sys + fence.sys).  Measured on 2 B200: same iteration
                # time, but in the 2-rank test ONE replica ended with rows that differ -- the unicast re-publication of a
                # row that received intersection forces (merge kernel) can overtake the multicast copy of the same row on
                # another route.  A correct protocol keeps both on one route (or tags rows with the iteration number).
                if use_multicast or os.environ.get("GEM_MULTICAST", "0") == "1":
                    mc = int(getattr(h_raw, "multicast_ptr", 0) or 0)
                stages.attach_p2p(h_pos.buffer_ptrs, h_raw.buffer_ptrs, h_x.buffer_ptrs, raw, xchg, eng._nb, S, kp1,
                                  barrier=lambda ch: h_x.barrier(channel=int(ch)), raw_multicast_ptr=mc)
            except Exception as exc:  # pylint: disable=broad-exception-caught
                # never silent: the fallback is another design (NCCL collectives) with other performance
                self.logger.warning("symmetric memory unavailable (%s): falling back to the NCCL all-gather flow", exc)
                stages.p2p_ready = False
            ok = torch.tensor([1 if stages.p2p_ready else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self._group)
            if int(ok.item()) == 0:
                stages.p2p_ready = False
        # rank 0's initial positions are the truth (ARPACK start vectors are not reproducible across processes)
        dist.broadcast(self._pos, src=src0, group=self._group)
        self._engine.pos = self._pos
        self._sgraphs = {}
        self._warmed = False

    def _buffers(self):
        """Only the small per-object buffers the base-class helpers use here (the big ones live in the engine)."""
        if not self._bufs:
            S = min(int(self.sample_size), self.n_edges)
            dev = self.device
            self._bufs = dict(key=None, S=S, kp1=int(self.n_neighbors) + 1, iter=self._engine.st._iter,
                              samp=torch.zeros((max(S, 1),), device=dev, dtype=torch.long),
                              samp_next=torch.zeros((max(S, 1),), device=dev, dtype=torch.long))
        return self._bufs

    @property
    def exchange(self) -> str:
        """Which flow the iteration runs ('p2p': peer stores over symmetric memory + 2 barriers; 'nccl')."""
        return "p2p" if self._engine.st.p2p_ready else "nccl"

    @property
    def multicast(self) -> bool:
        """Are the position rows pushed with NVSwitch multicast stores (multimem.st)?"""
        return bool(self._engine.st.p2p_ready and self._engine.st._raw_mc)

    def _world_and_rank(self):
        return dist.get_world_size(self._group), dist.get_rank(self._group)

    def update_positions(self, sampled_indices=None):
        with torch.cuda.device(self.device):
            self._engine.pos = self._pos
            if sampled_indices is None and self.sampler == "torch":
                b = self._buffers()
                if self._torch_sample_ready:
                    b["samp"].copy_(b["samp_next"])
                    self._torch_sample_ready = False
                else:
                    self._draw_torch_sample(b["samp"])
                sampled_indices = b["samp"]
            self._engine.step(sampled_indices)
        self.last_sampled_indices = self._engine.samp
        self.last_knn_indices = self._engine.knn_idx[:, 1:]

    _GRAPH_UNROLL = 4          # even: an unrolled graph starts at parity 0 and ends there (8: same time on 2 GPUs, measured)

    def _capture(self, parity, count: int = 1):
        """Capture `count` whole sharded iterations starting at this parity -- kernels on both streams, peer stores and
        the two device barriers of each -- in a CUDA graph.  The raw / exchange buffers alternate with the iteration
        parity, so there is one single-iteration graph per parity, plus one of _GRAPH_UNROLL iterations from parity 0
        (key "unrolled": consecutive replays leave ~10-15 us between the last kernel of one graph and the first of the
        next, a dependent kernel inside a graph starts after 2-4 us)."""
        eng = self._engine
        torch_samp = self.sampler == "torch"
        b = self._buffers() if torch_samp else None
        it = eng.iteration
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            if torch_samp:
                cur = torch.cuda.current_stream(self.device)
                b["samp"].copy_(b["samp_next"])
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    self._draw_torch_sample(b["samp_next"])
                eng.step(b["samp"])
                cur.wait_stream(side)
            else:
                for _ in range(int(count)):
                    eng.step()
        eng.iteration = it                      # capture does not execute
        self._sgraphs[parity if count == 1 else "unrolled"] = graph

    def _replay(self, num_iterations: int):
        eng = self._engine
        eng.pos = self._pos
        torch_samp = self.sampler == "torch"
        if torch_samp and not self._torch_sample_ready:
            self._draw_torch_sample(self._buffers()["samp_next"])
            self._torch_sample_ready = True
        todo = int(num_iterations)
        if not self._warmed and todo > 0:
            # one eager iteration outside capture first (NCCL communicators, lazy allocations, symmetric-memory state)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self.update_positions()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self._warmed = True
            todo -= 1
            if torch_samp and not self._torch_sample_ready:
                self._draw_torch_sample(self._buffers()["samp_next"])
                self._torch_sample_ready = True
        U = self._GRAPH_UNROLL if (eng.st.p2p_ready and not torch_samp and self.use_unrolled_graph) else 1
        while todo > 0:
            par = eng.iteration & 1 if eng.st.p2p_ready else 0
            if U > 1 and par == 0 and todo >= U:
                if "unrolled" not in self._sgraphs:
                    self._capture(0, U)
                self._sgraphs["unrolled"].replay()
                eng.iteration += U
                todo -= U
                continue
            if par not in self._sgraphs:
                self._capture(par)
            self._sgraphs[par].replay()
            eng.iteration += 1
            todo -= 1
        self.last_sampled_indices = eng.samp
        self.last_knn_indices = eng.knn_idx[:, 1:]

    def _graph_ok(self) -> bool:
        if not self.use_cuda_graph:
            return False
        if self.sampler == "device":
            return True
        return self.n_edges >= self._TORCH_SAMPLER_GRAPH_MIN_E

    def close(self):
        """Release the captured CUDA graphs.  Call before torch.distributed.destroy_process_group(): tearing
        down an NCCL communicator whose collectives are still referenced by a live graph hangs."""
        self._sgraphs = {}
        import gc
        gc.collect()
        torch.cuda.synchronize(self.device)
        super().close()

    def run_layout_device(self, num_iterations=100):
        with torch.cuda.device(self.device):
            if self._graph_ok() and (num_iterations > 1 or self._sgraphs):
                self._replay(int(num_iterations))
            else:
                for _ in range(int(num_iterations)):
                    self.update_positions()

    def run_layout(self, num_iterations=100):
        self.run_layout_device(num_iterations)
        return self.positions

    def profile_step(self):
        raise NotImplementedError("per-stage events of the single-GPU step; use profile_phases() on several GPUs")

    PHASES = ["prep", "spring", "colsum", "scan_select", "barrier_A", "merge_intersect", "barrier_B", "normalise"]

    def profile_phases(self, iterations: int = 20):
        """Timeline of the captured product-flow iteration: microseconds from the start of the step to the END of each
        phase on this rank (median over `iterations` replays), measured with external CUDA events recorded INSIDE the
        captured graphs (same launches, same overlap as the timed path).  prep / colsum run on the side stream."""
        st = self._engine.st
        if not st.p2p_ready:
            raise RuntimeError("profile_phases needs the symmetric-memory flow")
        eng = self._engine
        with torch.cuda.device(self.device):
            if not self._warmed:
                self.run_layout_device(2)
            saved, self._sgraphs = self._sgraphs, {}
            st._marks = {}
            rows = []
            try:
                for it in range(int(iterations) + 2):
                    par = eng.iteration & 1
                    if par not in self._sgraphs:
                        self._capture(par)
                    self._sgraphs[par].replay()
                    eng.iteration += 1
                    torch.cuda.synchronize(self.device)
                    if it >= 2:
                        t0 = st._marks["start"]
                        rows.append([t0.elapsed_time(st._marks[k]) * 1e3 for k in self.PHASES])
            finally:
                st._marks = None
                self._sgraphs = saved
        med = np.median(np.asarray(rows), axis=0)
        return dict(zip(self.PHASES, [float(x) for x in med]))

    # ------------------------------------------------------------------ host I/O split across the ranks
    def chunk_rows(self):
        """[lo, hi) of the public rows this rank moves in load_positions_chunk / read_positions_chunk."""
        per = (self.n + self._world - 1) // self._world
        lo = min(self._rank * per, self.n)
        return lo, min(lo + per, self.n)

    def load_positions_chunk(self, host_rows: torch.Tensor):
        """Collective.  Every rank uploads ITS contiguous chunk of the public (n, d) array (rows chunk_rows(), fp32,
        pinned for full speed) over its own PCIe link and fans the rows out into every rank's replica over NVLink
        (gem_rows_scatter with peer pointers): the upload of N rows costs N/world rows per link."""
        lo, hi = self.chunk_rows()
        if tuple(host_rows.shape) != (hi - lo, self.n_components) or host_rows.dtype != torch.float32:
            raise ValueError(f"expected an fp32 tensor of shape {(hi - lo, self.n_components)}")
        st = self._engine.st
        if not st.p2p_ready:
            raise RuntimeError("load_positions_chunk needs the symmetric-memory flow")
        with torch.cuda.device(self.device):
            if hi > lo:
                src = self._io_stage()[lo:hi]
                src.copy_(host_rows, non_blocking=True)
                _cabi.check(self._lib.gem_rows_scatter(_ptr(src), lo, hi - lo, self.n_components, _ptr(self._pad_index),
                                                       st.peer_ptrs, self._world, self._stream()), "gem_rows_scatter")
            self._symm["pos"].barrier(channel=0)

    def read_positions_chunk(self, out_rows: torch.Tensor):
        """Every rank downloads its chunk_rows() of the (replicated) positions into `out_rows`, then stream sync."""
        lo, hi = self.chunk_rows()
        with torch.cuda.device(self.device):
            if hi > lo:
                stage = self._io_stage("d2h_stage")[lo:hi]
                self._rows_to_public(lo, hi - lo, out=stage)
                out_rows.copy_(stage, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        return out_rows
