"""Multi-GPU layout iteration: one process per GPU, vertex sharding.

Every rank holds the replicated positions and graph arrays and OWNS the vertices v with
v mod G == rank (partition.py).  Because the spring stage is vertex-parallel ("pull" over the CSR), a
rank produces the COMPLETE force of each vertex it owns -- there is no per-vertex force reduction
across ranks -- and the midpoints of the edges whose first endpoint it owns, which are its share
of the KNN candidates.  Per iteration the ranks exchange only

    1. the per-rank partial top-(k+1) lists        (S*(k+1)*12 bytes per rank)
    2. the 2*ld column sums of the update          (64 bytes per rank)
    3. the updated position rows                   (4*ld*N bytes in total)

With CUDA stages each exchange is a store phase of the producing kernel into every rank's buffer over
peer-mapped symmetric memory (NVLink P2P) plus one device-side barrier -- exchange 3 is fused into the
normalisation kernel; NCCL all-gather / all-reduce is the fallback (and what the CPU tests use under
gloo).  The orchestration (`ShardedLayoutEngine`) is device-agnostic: it calls a `stages` object for
the compute.  The product binds it to the CUDA C ABI (`CudaStages`); the CPU tests bind it to the
oracle to check the sharding logic.
"""
from __future__ import annotations

import ctypes
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
        begin = getattr(st, "begin_step", None)
        if begin is not None:
            begin()                                                    # CUDA stages: fork the side stream here
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
        p2p = self.world > 1 and getattr(st, "peer_ptrs", None) is not None
        self.phase_a(sampled_indices)
        if p2p:
            # CUDA stages with symmetric memory: each exchange is a P2P store phase from this rank into every
            # rank's buffer followed by one device-side cross-rank barrier -- no NCCL collective in the iteration
            st.push_lists(self.part, self.rank)
        elif self.world > 1:
            dist.all_gather_into_tensor(self.gathered.view(-1), self.part, group=self.group)
        else:
            self.gathered.view(-1).copy_(self.part)
        if p2p and getattr(st, "fused", False):
            # fused form (CUDA stages): the spring kernel has written pos+F of the owned rows into self.force and a
            # side-stream pass has taken its column sums; the merge kernel adds the intersection forces with a
            # correction of the sums; the sums go to the peers straight from the workspace
            st.merge_intersect(self.g_idx, self.g_dist, self.knn_idx, self.knn_dist, self.pos, self.samp, self.vb,
                               self.ve, self.force)
            st.push_stats(None, self.rank)
            st.normalise_and_push(self.pos, self.vb, self.ve, self.L.n, src=self.force)
            return
        self.phase_b()
        if p2p:
            # the barrier inside also tells every rank that all ranks have finished READING the old positions
            st.push_stats(self.stats, self.rank)
            st.normalise_and_push(self.pos, self.vb, self.ve, self.L.n)
            return
        if self.world > 1:
            dist.all_reduce(self.stats, group=self.group)
        self.phase_c()
        if self.world > 1:
            block = self.own_block()
            dist.all_gather_into_tensor(self.pos.view(-1), (block if self.inplace else block.clone()).view(-1),
                                        group=self.group)


class CudaStages:
    """The compute stages of the engine on the CUDA C ABI (include/graphem_b200.h) for one rank of
    a GraphLayout.  `arrays` lets the owner share graph tensors it has already uploaded."""

    def __init__(self, layout: GraphLayout, rank: int, device, *, n_components: int, k_attr: float, L_min: float,
                 k_inter: float, seed: int, arrays: Optional[dict] = None):
        self.L, self.rank = layout, rank
        self.device = torch.device(device)
        self.d = int(n_components)
        self.k_attr, self.L_min, self.k_inter = float(k_attr), float(L_min), float(k_inter)
        self.seed = int(seed) & (2 ** 64 - 1)
        self.lib = _cabi.load()
        _cabi.init_device(self.device.index if self.device.index is not None else torch.cuda.current_device())
        self.ld, self.mld = self.lib.gem_row_pitch(self.d), self.lib.gem_mid_pitch(self.d)
        if not layout.sorted_edges or self.d not in (2, 3):
            raise NotImplementedError("the multi-GPU path needs an (i,j)-sorted edge list and n_components in {2,3}")
        up = lambda x: torch.from_numpy(x).to(self.device)           # noqa: E731
        a = arrays or {}
        self.edges32 = a["edges32"] if "edges32" in a else up(layout.edges32).contiguous()
        self.row_ptr = a["row_ptr"] if "row_ptr" in a else up(layout.row_ptr)
        self.col = a["col"] if "col" in a else up(layout.col)
        self.up_ptr = a["up_ptr"] if "up_ptr" in a else up(layout.up_ptr)
        self.hubs = up(layout.hubs[rank])
        # the rank's own edges in the order its spring kernel writes their midpoints (= its KNN candidates)
        own = layout.local_edge_ids(rank)
        if layout.edge_orig is None:                                  # monotonic numbering: a slice of the edge list
            lo = int(layout.e_lo[rank])
            self.edges32_local = self.edges32[lo: lo + len(own)]
            self.l2g = None
        else:
            self.edges32_local = up(np.ascontiguousarray(layout.edges32[own]))
            self.l2g = up(np.ascontiguousarray(own))
        self._iter = torch.zeros((1,), device=self.device, dtype=torch.int64)   # device-side iteration counter
        self._knn_ws = None
        self._stats_ws = None
        self._draw = False
        self._bump = None
        # The KNN preparation (sample, query midpoints, line-graph hint, bound, thresholds) needs the
        # positions only: it runs on a side stream while the spring kernel runs on the current one.
        self._side = torch.cuda.Stream(device=self.device)
        self._fork = torch.cuda.Event()
        self._join = torch.cuda.Event()

    peer_ptrs = None            # set by attach_symmetric(): device pointers of every rank's position buffer
    _symm = None

    multicast = False

    def attach_symmetric(self, pos_handle, xchg_handle, xchg: torch.Tensor, list_bytes: int, use_multicast: bool = False,
                         fused: bool = True):
        """pos_handle / xchg_handle: torch.distributed._symmetric_memory rendezvous handles of the position
        buffer and of the small exchange buffer `xchg` = [world x list_bytes partial lists | world x 2*ld doubles]."""
        self._symm = pos_handle
        self._xsymm = xchg_handle
        world = pos_handle.world_size
        mc_pos = int(getattr(pos_handle, "multicast_ptr", 0) or 0) if use_multicast else 0
        mc_x = int(getattr(xchg_handle, "multicast_ptr", 0) or 0) if use_multicast else 0
        if mc_pos and mc_x:
            # EXPERIMENTAL, off by default.  NVSwitch multicast mapping of the same buffers: ONE store to this
            # address lands in every rank's replica (the switch replicates it), so a row leaves the GPU once
            # instead of world-1 times.  Measured on 2 B200: plain stores to the multicast address followed by the
            # symmetric-memory barrier give WRONG positions on the peers (the unicast signal overtakes the posted
            # multicast writes); it needs multimem.st + a system-scope fence/flag protocol inside the kernel.
            self.peer_ptrs = (ctypes.c_void_p * 1)(mc_pos)
            self.xchg_ptrs = (ctypes.c_void_p * 1)(mc_x)
            self.multicast = True
        else:
            self.peer_ptrs = (ctypes.c_void_p * world)(*[int(p) for p in pos_handle.buffer_ptrs])
            self.xchg_ptrs = (ctypes.c_void_p * world)(*[int(p) for p in xchg_handle.buffer_ptrs])
            self.multicast = False
        self._xchg = xchg
        self._list_bytes = list_bytes
        self._stats_off = world * list_bytes
        self._stats_bytes = 2 * self.ld * 8
        self._spring_done = torch.cuda.Event()
        self._stats_done = torch.cuda.Event()
        self._force_ref = None
        self.fused = bool(fused) and self.L.n_edges > 0

    def gathered_lists(self, world):
        return self._xchg[: world * self._list_bytes].view(world, self._list_bytes)

    def push_lists(self, part, rank):
        _cabi.check(self.lib.gem_push_bytes(self.xchg_ptrs, len(self.xchg_ptrs), rank * self._list_bytes, _ptr(part),
                                            part.numel(), self._s()), "gem_push_bytes(lists)")
        self._xsymm.barrier(channel=0)

    def push_stats(self, stats, rank):
        src = stats.view(torch.uint8) if stats is not None else self._ws(1)       # fused form: straight from the workspace
        _cabi.check(self.lib.gem_push_bytes(self.xchg_ptrs, len(self.xchg_ptrs), self._stats_off + rank * self._stats_bytes,
                                            _ptr(src), self._stats_bytes, self._s()), "gem_push_bytes(stats)")
        self._xsymm.barrier(channel=1)

    def normalise_and_push(self, pos, vb, ve, n_total, src=None):
        ws = self._ws(ve - vb)
        own = pos[vb:ve] if src is None else src
        rank_sums = self._xchg[self._stats_off:]
        _cabi.check(self.lib.gem_update_normalise_push(self.peer_ptrs, len(self.peer_ptrs), _ptr(own), vb, ve - vb, n_total,
                                                       self.d, _ptr(ws), _ptr(rank_sums), self._s()),
                    "gem_update_normalise_push")
        self._symm.barrier(channel=0)        # every rank's rows have landed everywhere before anyone reads them

    def begin_step(self):
        main = torch.cuda.current_stream(self.device)
        self._fork.record(main)
        self._side.wait_event(self._fork)

    def _side_ptr(self):
        return ctypes.c_void_p(self._side.cuda_stream)

    def alloc(self, shape, dtype):
        return torch.zeros(shape, device=self.device, dtype=dtype)

    def _s(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def sample(self, iteration, n_edges, samp):
        # drawn by the fused preparation launch in hint(); the device counter advances by itself (bumped by the
        # threshold kernel), so a captured CUDA graph replays correctly; every rank starts from 0 and therefore
        # draws the same ids
        self._draw = True

    fused = False               # set with attach_symmetric(): fused spring+update form of the iteration

    def spring(self, pos, vb, ve, force, mid, e_lo):
        self._pos_ref = pos
        if self.fused and self.peer_ptrs is not None and ve > vb:
            main = torch.cuda.current_stream(self.device)
            _cabi.check(self.lib.gem_spring_update_csr(
                _ptr(pos), _ptr(self.row_ptr), _ptr(self.col), _ptr(self.up_ptr), vb, ve,
                _ptr(self.hubs) if self.hubs.numel() else None, int(self.hubs.numel()), self.d, self.k_attr,
                self.L_min, _ptr(force), _ptr(mid), e_lo, self._s()), "gem_spring_update_csr")
            self._spring_done.record(main)
            self._n_own = ve - vb
            self._force_ref = force
            return
        _cabi.check(self.lib.gem_spring_midpoints_csr(
            _ptr(pos), _ptr(self.row_ptr), _ptr(self.col), _ptr(self.up_ptr), vb, ve,
            _ptr(self.hubs) if self.hubs.numel() else None, int(self.hubs.numel()), self.d, self.k_attr,
            self.L_min, _ptr(force), _ptr(mid), e_lo, self._s()), "gem_spring_midpoints_csr")

    def query_mid(self, pos, samp, qmid):
        self._qmid = qmid                                             # written by the fused launch in hint()

    def hint(self, pos, samp, kp1, tau_hint):
        # sample (when not injected) + query midpoints + line-graph bound: one launch on the side stream
        draw = 1 if self._draw else 0
        _cabi.check(self.lib.gem_knn_query_prep(self.seed, _ptr(self._iter), draw, _ptr(pos), _ptr(self.row_ptr),
                                                _ptr(self.col), _ptr(self.edges32), self.L.n_edges, _ptr(samp),
                                                samp.numel(), self.d, kp1, _ptr(self._qmid), _ptr(tau_hint),
                                                self._side_ptr()), "gem_knn_query_prep")
        self._bump = _ptr(self._iter) if draw else None
        self._draw = False

    def knn_local(self, mid, e_loc, e_total, e_lo, qmid, tau_hint, kp1, out_idx, out_dist):
        S = qmid.shape[0]
        if self._knn_ws is None:
            nbytes = ctypes.c_size_t(0)
            _cabi.check(self.lib.gem_knn_workspace_bytes(max(e_loc, 1), self.d, S, kp1, ctypes.byref(nbytes)))
            self._knn_ws = torch.zeros((nbytes.value + 256,), device=self.device, dtype=torch.uint8)
            self._knn_ws_bytes = nbytes.value
        main = torch.cuda.current_stream(self.device)
        if self.lib.gem_knn_fast_path(e_loc, e_total, self.d, S, kp1):
            # bound / thresholds from (pos, local edges) on the side stream, scan on the main one after the join
            e32 = self.edges32_local
            off = e_lo if self.l2g is None else 0                     # strided ownership: local numbers, mapped below
            _cabi.check(self.lib.gem_knn_prepare(None, _ptr(self._pos_ref), _ptr(e32), e_loc, self.d, _ptr(qmid), S, kp1,
                                                 _ptr(tau_hint), self._bump, _ptr(self._knn_ws), self._knn_ws_bytes,
                                                 self._side_ptr()), "gem_knn_prepare")
            self._bump = None
            self._join.record(self._side)
            main.wait_event(self._join)
            self._side_stats_pass()
            _cabi.check(self.lib.gem_knn_scan(_ptr(mid), e_loc, off, self.d, _ptr(qmid), S, kp1, _ptr(out_idx),
                                              _ptr(out_dist), _ptr(self._knn_ws), self._knn_ws_bytes, self._s()),
                        "gem_knn_scan")
            self._remap(out_idx)
            return
        self._join.record(self._side)
        main.wait_event(self._join)
        self._side_stats_pass()
        if self._bump is not None:                                    # no fast path here: bump the sample counter ourselves,
            self._iter.add_(1)                                        # after the join (the fused launch reads it on the side stream)
            self._bump = None
        mm = 1 if (S > 25 or e_total > 25) else 0                     # torch.cdist's rule on the WHOLE problem
        off = e_lo if self.l2g is None else 0
        _cabi.check(self.lib.gem_knn_midpoints_shard(_ptr(mid), e_loc, e_total, off, self.d, _ptr(qmid), S, kp1, mm,
                                                     _ptr(tau_hint), _ptr(out_idx), _ptr(out_dist), _ptr(self._knn_ws),
                                                     self._knn_ws_bytes, self._s()), "gem_knn_midpoints_shard")
        self._remap(out_idx)

    def _remap(self, out_idx):
        """local-order edge numbers -> original edge ids (ties in the merge are broken by ORIGINAL index; inside
        a rank the local order is the original order restricted to its edges, so its own top-(k+1) is unaffected)"""
        if self.l2g is not None:
            _cabi.check(self.lib.gem_remap_indices(_ptr(out_idx), out_idx.numel(), _ptr(self.l2g), self._s()),
                        "gem_remap_indices")

    def _side_stats_pass(self):
        """Fused form: column sums of the new positions (pos+F of the owned rows) on the side stream, next to the scan."""
        if not (self.fused and self.peer_ptrs is not None and getattr(self, "_force_ref", None) is not None):
            return
        ws = self._ws(self._n_own)
        self._side.wait_event(self._spring_done)
        _cabi.check(self.lib.gem_update_positions(_ptr(self._force_ref), None, None, self._n_own, self._n_own, self.d,
                                                  _ptr(ws), 3, self._side_ptr()), "gem_update_positions(phase 3)")
        self._stats_done.record(self._side)

    def merge_intersect(self, g_idx, g_dist, out_idx, out_dist, pos, samp, vb, ve, newpos):
        parts, S, kp1 = g_idx.shape
        ws = self._ws(ve - vb)
        torch.cuda.current_stream(self.device).wait_event(self._stats_done)     # the sums it corrects must be there
        _cabi.check(self.lib.gem_topk_merge_intersect(_ptr(g_dist), _ptr(g_idx), g_dist.stride(0), g_idx.stride(0), parts,
                                                      S, kp1, _ptr(out_idx), _ptr(out_dist), _ptr(pos), _ptr(self.edges32),
                                                      _ptr(samp), self.d, self.k_inter, vb, ve, _ptr(newpos), _ptr(ws),
                                                      self._s()), "gem_topk_merge_intersect")

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

    def _ws(self, n_rows):
        if self._stats_ws is None:
            nbytes = ctypes.c_size_t(0)
            _cabi.check(self.lib.gem_update_workspace_bytes(max(n_rows, 1), self.d, ctypes.byref(nbytes)))
            self._stats_ws = torch.zeros((nbytes.value + 256,), device=self.device, dtype=torch.uint8)
        return self._stats_ws

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
    every iteration with the same replicated positions."""

    def __init__(self, adjacency, n_components=2, *args, process_group=None, use_symmetric_memory=True,
                 use_multicast=False, fused_update=True, ownership="strided", **kwargs):
        if not dist.is_initialized():
            raise RuntimeError("ShardedGraphEmbedder needs an initialised torch.distributed process group")
        self._group = process_group
        self._ownership = ownership            # 'strided' (v mod G: balanced for any vertex order) | 'contiguous'
        kwargs["graph_build"] = "host"        # the vertex partition (partition.build_layout) is host work
        super().__init__(adjacency, n_components, *args, **kwargs)
        if self.sampler != "device":
            raise NotImplementedError("the multi-GPU path uses the device sampler (identical ids on every rank)")
        stages = CudaStages(self._layout, self._rank, self.device, n_components=self.n_components, k_attr=self.k_attr,
                            L_min=self.L_min, k_inter=self.k_inter, seed=self._sampler_seed,
                            arrays=dict(edges32=self._edges32, row_ptr=self._row_ptr, col=self._col,
                                        up_ptr=self._up_ptr))
        self._engine = ShardedLayoutEngine(self._layout, self._rank, stages, n_components=self.n_components,
                                           n_neighbors=self.n_neighbors, sample_size=self.sample_size,
                                           group=self._group, pos=self._pos)
        # the replicated position buffer lives in symmetric memory: every rank can store into every replica
        self._symm_handle = None
        if self._world > 1 and use_symmetric_memory:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                buf = symm_mem.empty(tuple(self._pos.shape), dtype=torch.float32, device=self.device)
                buf.copy_(self._pos)
                grp = self._group if self._group is not None else dist.group.WORLD
                self._symm_handle = symm_mem.rendezvous(buf, grp)
                eng = self._engine
                xbytes = self._world * eng._nb + self._world * 2 * self._ld * 8
                xchg = symm_mem.empty(((xbytes + 255) // 256 * 256,), dtype=torch.uint8, device=self.device)
                xchg.zero_()
                self._xchg_handle = symm_mem.rendezvous(xchg, grp)
                self._pos = buf
                eng.pos = buf
                stages.attach_symmetric(self._symm_handle, self._xchg_handle, xchg, eng._nb, use_multicast,
                                        fused=bool(fused_update) and self.n_neighbors + 1 <= 64)
                eng.bind_exchange(stages.gathered_lists(self._world))
            except Exception as exc:  # pylint: disable=broad-exception-caught
                if self.verbose:
                    self.logger.warning("symmetric memory unavailable (%s): falling back to the NCCL all-gather", exc)
            # the exchange method is part of the collective sequence: every rank must take the same one
            ok = torch.tensor([1 if stages.peer_ptrs is not None else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self._group)
            if int(ok.item()) == 0:
                stages.peer_ptrs = None
        # rank 0's initial positions are the truth (ARPACK start vectors are not reproducible across processes)
        dist.broadcast(self._pos, src=dist.get_global_rank(self._group, 0) if self._group is not None else 0,
                       group=self._group)
        self._engine.pos = self._pos
        self._sgraph = None
        self._sgraph_pos = None

    def _world_and_rank(self):
        return dist.get_world_size(self._group), dist.get_rank(self._group)

    def update_positions(self, sampled_indices=None):
        with torch.cuda.device(self.device):
            self._engine.pos = self._pos
            self._engine.step(sampled_indices)
        self.last_sampled_indices = self._engine.samp
        self.last_knn_indices = self._engine.knn_idx[:, 1:]

    def _replay(self, num_iterations: int):
        """Capture one whole sharded iteration -- kernels on both streams AND the three NCCL
        collectives -- in a CUDA graph and replay it: no Python / launch overhead per iteration."""
        if self._sgraph is None or self._sgraph_pos != self._pos.data_ptr():
            self._engine.pos = self._pos
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                    # warm-up outside capture (NCCL communicators, lazy buffers)
                self._engine.step()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._engine.step()
            self._sgraph, self._sgraph_pos = graph, self._pos.data_ptr()
            num_iterations -= 1                              # the warm-up step was a real iteration
        for _ in range(num_iterations):
            self._sgraph.replay()
        self.last_sampled_indices = self._engine.samp
        self.last_knn_indices = self._engine.knn_idx[:, 1:]

    def close(self):
        """Release the captured CUDA graph.  Call before torch.distributed.destroy_process_group(): tearing
        down an NCCL communicator whose collectives are still referenced by a live graph hangs."""
        self._sgraph = None
        import gc
        gc.collect()
        torch.cuda.synchronize(self.device)

    def run_layout_device(self, num_iterations=100):
        with torch.cuda.device(self.device):
            if self.use_cuda_graph and (num_iterations > 1 or self._sgraph is not None):
                self._replay(int(num_iterations))
            else:
                for _ in range(int(num_iterations)):
                    self.update_positions()

    def run_layout(self, num_iterations=100):
        self.run_layout_device(num_iterations)
        return self.positions

    def profile_step(self):
        raise NotImplementedError("per-stage profiling is a single-GPU tool")
