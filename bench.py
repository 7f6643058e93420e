#!/usr/bin/env python
"""
bench.py -- GraphEm layout-iteration benchmark (BASELINE.json metric:
"layout iters/sec & edge-updates/s, 1M-vertex BA graph, 1/2/4/8 B200 vs host CPU").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one update_positions (spring + midpoint KNN + intersection + update) on the
synthetic workload.  `value` = edge-updates/s = E * iterations / second over ALL ranks (the same
graph is edge-sharded across the ranks, so scaling is "strong").  Rank 0 prints ONE JSON line.

Timing: W >= 3 warm-up steps, CUDA events on the launching stream, barrier + synchronize on both sides, max over
ranks.  A workload whose iteration streams MORE than the 126 MB L2 (C3: 212 MB, C4, C5) is timed as EXACTLY K
iterations back to back in one event pair -- run_layout(K), the production path -- after one L2 flush; a smaller
one (C1, C2) as K single iterations, each bracketed by its own pair of events and preceded by an (untimed) L2
flush (a 512 MiB buffer write), durations summed.  `config.l2` says which; the large workloads also carry the
flushed single-iteration figure (`flushed_single_replays`).  `sustained` repeats the measurement as >= 2 s of
back-to-back replays (clocks recorded).  `e2e` times the drop-in API with host buffers: `emb.positions = x`
(pinned host tensor -> device), `emb.run_layout(1)`, which returns the result as a fresh ndarray; on
N > 1 GPUs every rank moves its 1/N of the rows (load_positions_chunk / read_positions_chunk).
N > 1 lines carry `parity` (sharded state vs a single-GPU embedder on rank 0) and `phase_us`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "layout iters/sec & edge-updates/s, 1M-vertex BA graph, 1/2/4/8 B200 vs host CPU"


def scan_traffic_from_profiles(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE knn_scan_kernel launch, from this round's `ncu --set full`
    capture of the same build (profiles/r02_scan_traffic.json, written by scripts/ncu_summary.py); None when no
    capture of this build/workload is committed -- a stale constant is worse than null."""
    path = os.path.join(ROOT, "profiles", "r02_scan_traffic.json")
    try:
        with open(path) as f:
            rec = json.load(f)
        return rec.get(workload)
    except Exception:
        return None

WORKLOADS = {
    # BASELINE.json configs[2] -- the headline
    "c3": dict(desc="barabasi_albert n=1000000 m=4 d=3 k=10 S=256", kind="ba", n=1_000_000, m=4, d=3, k=10, S=256),
    # configs[1]
    "c2": dict(desc="random_regular n=100000 deg=8 d=2 k=10 S=256", kind="rr", n=100_000, deg=8, d=2, k=10, S=256),
    # configs[3]
    "c4": dict(desc="sbm n=2000000 16 blocks d=3 k=32 S=256", kind="sbm", n=2_000_000, blocks=16, d=3, k=32, S=256),
    # configs[4]
    "c5": dict(desc="erdos_renyi n=10000000 avg_deg=10 d=3 k=10 S=256", kind="er", n=10_000_000, d=3, k=10, S=256),
    # configs[0] (README quick start; CPU-runnable)
    "c1": dict(desc="erdos_renyi n=1000 p=0.01 d=3 k=10 S=256", kind="er_small", n=1000, d=3, k=10, S=256),
    "tiny": dict(desc="barabasi_albert n=20000 m=4 d=3 k=10 S=256", kind="ba", n=20_000, m=4, d=3, k=10, S=256),
}


def make_graph(w):
    import graphem_rapids_b200.generators as gen
    if w["kind"] == "ba":
        return gen.generate_ba(w["n"], w["m"], seed=0)
    if w["kind"] == "rr":
        return gen.generate_random_regular(w["n"], w["deg"], seed=0)
    if w["kind"] == "sbm":
        npb = w["n"] // w["blocks"]
        return gen.generate_sbm(npb, w["blocks"], 6.4e-5, 1.07e-6, seed=0)
    if w["kind"] == "er":
        return gen.erdos_renyi_graph(w["n"], 10.0 / w["n"], seed=0)
    if w["kind"] == "er_small":
        return gen.erdos_renyi_graph(w["n"], 0.01, seed=0)
    raise ValueError(w["kind"])


def initial_positions(n, d):
    """randn*0.1 from default_rng(0): the reference's own fallback distribution
    (embedder_pytorch.py:369); the Laplacian init is outside the timed path."""
    return (np.random.default_rng(0).standard_normal((n, d)) * 0.1).astype(np.float32)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The region is short (K steps of ~0.3 ms), so
    the sampler is an NVML thread polling every ~2 ms (nvidia-smi -lms cannot start that fast); falls back
    to the nvidia-smi query of B200_PROFILING.md when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.samples = []           # (sm_mhz, reasons bitmask)
        self.sm_max = None
        self._stop = threading.Event()
        self.thread = None
        self.nvml = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.gpu_index < len(ids) and ids[self.gpu_index].isdigit():
                return int(ids[self.gpu_index])
        return self.gpu_index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                  str(self._physical_index())], capture_output=True, text=True, timeout=10).stdout
            parts = [x.strip() for x in out.strip().splitlines()[-1].split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = [nm for nm, val in zip(names, parts[5:9]) if val.lower().startswith("active")]
            return {"sm_mhz": float(parts[1]), "sm_max_mhz": float(parts[2]), "reasons": reasons, "samples": 1,
                    "source": "nvidia-smi query right after the timed region (pynvml unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}

    def stop(self):
        if self.nvml is None:
            return self._smi_once()
        self._stop.set()
        self.thread.join(timeout=1.0)
        n = self.nvml
        if not self.samples:
            return self._smi_once()
        def bit(new_name, old_name, default):
            return getattr(n, new_name, getattr(n, old_name, default))
        bits = {"hw_slowdown": bit("nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": bit("nvmlClocksEventReasonHwThermalSlowdown",
                                           "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": bit("nvmlClocksEventReasonSwThermalSlowdown",
                                           "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": bit("nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        seen = 0
        for _, rs in self.samples:
            seen |= rs
        reasons = sorted(nm for nm, bit in bits.items() if seen & bit)
        return {"sm_mhz": float(np.median([sm for sm, _ in self.samples])), "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(self.samples), "source": "NVML polled every ~2 ms during the timed region"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# graph cache: generate once per box (rank 0), every other rank loads the CSR arrays
# ----------------------------------------------------------------------------------------------
def load_graph(w, key, world, rank, barrier):
    import scipy.sparse as sp
    path = os.path.join(os.environ.get("GEM_GRAPH_CACHE", "/tmp"), f"gem_graph_{key}.npz")
    if world == 1:
        return make_graph(w)
    if rank == 0:
        adj = make_graph(w)
        tmp = path + ".tmp.npz"
        np.savez(tmp, indptr=adj.indptr, indices=adj.indices, n=np.int64(adj.shape[0]))
        os.replace(tmp, path)
    barrier()
    if rank != 0:
        with np.load(path) as z:
            n = int(z["n"])
            indices, indptr = z["indices"], z["indptr"]
        adj = sp.csr_matrix((np.ones(len(indices), dtype=np.int64), indices, indptr), shape=(n, n))
        adj.has_sorted_indices = True
    barrier()
    if rank == 0:
        try:
            os.remove(path)
        except OSError:
            pass
    return adj


L2_BYTES = 126e6          # B200: 126 MB of L2 (B200_PROFILING.md)


def iteration_bytes(w, n, E):
    """algorithmic bytes one iteration streams, 8E + 8dE + 28dN."""
    return 8.0 * E + 8.0 * w["d"] * E + 28.0 * w["d"] * n


def back_to_back(w, n, E):
    """Timing rule: between timed iterations either flush L2 or use inputs larger than L2.  A workload whose iteration
    streams more than the L2 holds is timed as K iterations back to back (the production run_layout(K)); a smaller one
    as K single iterations with an untimed L2 flush before each."""
    return iteration_bytes(w, n, E) > L2_BYTES


def workload_config(w, n, E):
    """The `config` object: identical in both arms (it names the workload, not the implementation)."""
    bi = iteration_bytes(w, n, E)
    l2 = (f"GPU arm: NOT flushed -- one iteration streams {bi / 1e6:.0f} MB (8E + 8dE + 28dN) > {L2_BYTES / 1e6:.0f} MB of L2; the K "
          "timed iterations run back to back as run_layout(K) does (one L2 flush before the region); the line also carries "
          "the flushed single-iteration figure (`flushed_single_replays`)"
          if back_to_back(w, n, E) else
          f"GPU arm: flushed before every timed step (512 MiB write, untimed): one iteration streams only {bi / 1e6:.1f} MB")
    return {"workload": w["desc"], "N": int(n), "E": int(E), "n_components": w["d"], "sample_size": int(min(w["S"], E)),
            "n_neighbors": w["k"], "init": "randn*0.1 default_rng(0)",
            "graph": "graphem_rapids_b200.generators (numpy, seed 0): same degree law as the reference's networkx generator, "
                     "not its random stream (E differs slightly from the reference's own graph of that name)",
            "l2": l2 + "; CPU arm: not applicable"}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port (torch CPU restatement of the reference) timed on
# the host cores.  The ONLY place besides tests/ and smoke() that executes oracle/.
# ----------------------------------------------------------------------------------------------
def cpu_reference_timing(adj, w, steps, warmup, chunk=32, with_gc=False, budget_s=150.0):
    """K timed iterations of the reference algorithm (oracle/oracle.py: the literal cdist + topk restatement) on all
    host cores.  with_gc: additionally the 4 gc.collect() per iteration that the reference's MemoryManager /
    cleanup_gpu_memory execute (utils/memory_management.py:117-128) -- the as-shipped-equivalent figure.
    When K full iterations would not fit `budget_s`, every step runs the KNN on the first S' of the S queries and
    its KNN time is scaled by S/S' (said in `sample`)."""
    import gc
    from oracle import oracle
    torch.set_num_threads(os.cpu_count() or 1)
    edges = torch.from_numpy(oracle.extract_edges(adj).astype(np.int64))
    E = edges.shape[0]
    S = min(w["S"], E)
    pos = torch.from_numpy(initial_positions(adj.shape[0], w["d"]))
    gen = torch.Generator().manual_seed(0)

    def one(pos, s_used):
        samp = oracle.draw_sample(E, S, gen)[:s_used]
        t0 = time.perf_counter()
        if with_gc:
            gc.collect()                       # spring stage exit (:616 via MemoryManager)
        out = oracle.layout_step(pos, edges, samp, n_neighbors=w["k"], strict=False, chunk_size=chunk)
        if with_gc:
            gc.collect(); gc.collect(); gc.collect()      # knn / intersection / update stage exits
        return out["new_pos"], time.perf_counter() - t0

    pos, t1 = one(pos, S)                      # first warm-up step doubles as the cost probe
    s_used = S
    if t1 * (steps + warmup) > budget_s and S > 8:
        s_used = max(8, int(S * budget_s / (t1 * (steps + warmup))))
    for _ in range(max(warmup - 1, 0)):
        pos, _ = one(pos, s_used)
    times = []
    for _ in range(steps):
        pos, dt = one(pos, s_used)
        times.append(dt)
    sec = float(np.mean(times))
    sample = (f"{steps} full iterations of the workload (E={E}) after {warmup} warm-up")
    if s_used < S:
        # the KNN dominates: scale the step to the full query count (conservative for the CPU: the non-KNN part is
        # scaled too)
        sec = sec * S / s_used
        sample = (f"{steps} iterations with the KNN on the first {s_used} of {S} queries, step time scaled by "
                  f"{S}/{s_used} (a full iteration takes {t1:.1f} s), after {warmup} warm-up")
    sample += (f"; literal cdist+topk restatement (oracle/oracle.py, strict=False), query chunk {chunk}; "
               + ("with the reference's 4 gc.collect() per iteration (MemoryManager, utils/memory_management.py:117-128)"
                  if with_gc else "gc.collect/MemoryManager of the reference omitted (the maths only)"))
    return dict(E=int(E), sec_per_step=sec, cores=torch.get_num_threads(), sample=sample)


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    adj = make_graph(w)
    steps, warm = max(1, args.steps), max(1, args.warmup)
    r = cpu_reference_timing(adj, w, steps, warm)
    val = r["E"] / r["sec_per_step"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "edge-updates/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "iters_per_s": 1.0 / r["sec_per_step"],
        "config": workload_config(w, adj.shape[0], r["E"]),
        "details": {"note": "reference algorithm on host CPU cores (oracle port)"},
        "cpu_baseline": {"value": val, "unit": "edge-updates/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": val, "unit": "edge-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def rel_inf(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def run_b200(args, w):
    import torch.distributed as dist
    import graphem_rapids_b200 as gr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        # a rank-divergent bug must not hold N GPUs for the default 10-minute watchdog
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    t_setup0 = time.perf_counter()
    adj = load_graph(w, f"{args.workload}_{os.environ.get('MASTER_PORT', '0')}", world, rank, barrier)
    t_graph = time.perf_counter() - t_setup0
    n, d = adj.shape[0], w["d"]
    pos0 = initial_positions(n, d)
    kw = dict(n_components=d, device=dev, n_neighbors=w["k"], sample_size=w["S"], verbose=False, seed=0,
              initial_positions=pos0, sampler=args.sampler)
    t_c0 = time.perf_counter()
    if world > 1:
        from graphem_rapids_b200.sharded import ShardedGraphEmbedder
        emb = ShardedGraphEmbedder(adj, **kw)
    else:
        emb = gr.GraphEmbedderPyTorch(adj, **kw)
    torch.cuda.synchronize(dev)
    t_construct = time.perf_counter() - t_c0
    E = emb.n_edges
    # one-off graph set-up (untimed by the metric; reported): the object was built above with CUDA already
    # initialised, so a second construction measures the set-up itself -- device build vs host build
    setup = None
    if world == 1 and not args.no_cpu_baseline and not args.profile_mode and E <= 12_000_000:
        setup = {}
        for mode in ("auto", "host"):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            tmp = gr.GraphEmbedderPyTorch(adj, graph_build=mode, **kw)
            torch.cuda.synchronize(dev)
            setup[f"{mode}_s"] = round(time.perf_counter() - t0, 4)
            setup[f"{mode}_on_device"] = bool(tmp._layout.on_device)
            tmp.close()
            del tmp

    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

    K, W = args.steps, max(args.warmup, 3)
    # warm-up: W iterations through the production path (run_layout's CUDA-graph replay; the graph of one
    # iteration -- both streams and, for N > 1, the peer stores and device barriers -- is captured here, untimed)
    emb.run_layout_device(W + (W & 1))      # even count: both buffer parities of the sharded flow are captured
    barrier()

    def one_step():
        emb.run_layout_device(1)               # one replay of the captured iteration

    # ---- device-resident timing: EXACTLY K steps, each bracketed by its own CUDA events on the launching
    # stream and preceded by an (untimed) L2 flush; barrier + synchronize on both sides of the region
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(K):
        flush_buf.fill_(i & 0xFF)
        starts[i].record()
        one_step()
        ends[i].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = float(np.sum(step_ms))
    flushed = None
    if back_to_back(w, n, E):
        # inputs larger than L2: EXACTLY K iterations back to back in one event pair (run_layout's production path:
        # graphs of several iterations + single ones), after one L2 flush; barrier + synchronize on both sides
        flushed = {"ms_per_step": total_ms / K, "min": float(np.min(step_ms)), "median": float(np.median(step_ms)),
                   "max": float(np.max(step_ms)),
                   "note": "K single-iteration graph replays, each with its own event pair behind an untimed 512 MiB L2 "
                           "flush: includes the launch latency of one graph per iteration"}
        emb.run_layout_device(emb._GRAPH_UNROLL + 2 + (emb._GRAPH_UNROLL & 1))      # untimed: the unrolled graph is captured
        barrier()
        flush_buf.fill_(7)
        a0, b0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.perf_counter()
        a0.record()
        emb.run_layout_device(K)
        b0.record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
        total_ms = float(a0.elapsed_time(b0))
        step_ms = [total_ms / K] * K
    clk = clocks.stop() if rank == 0 else None

    # ---- sustained: >= 2 s of back-to-back replays (no flush: every array of the step is larger than it can keep
    # in L2 across an iteration only at C4/C5; said in the line), clocks sampled over the whole region
    est_ms = total_ms / K
    if world > 1:                              # the iteration count must be the SAME on every rank (device barriers inside)
        t = torch.tensor([est_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        est_ms = float(t.item())
    sus_iters = int(min(max(2.2e3 / max(est_ms, 1e-3), K), 200000))
    if args.profile_mode:
        sus_iters = 2
    sus_iters += sus_iters & 1
    clocks2 = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks2.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    emb.run_layout_device(sus_iters)
    b.record()
    barrier()
    sus_ms = a.elapsed_time(b) / sus_iters
    clk2 = clocks2.stop() if rank == 0 else None

    # ---- end to end through the public API with HOST buffers
    sharded_io = world > 1 and emb.exchange == "p2p"
    Ke = max(3, min(K, 10))
    if sharded_io:
        lo, hi = emb.chunk_rows()
        cur = emb.positions
        host_in = torch.from_numpy(np.ascontiguousarray(cur[lo:hi])).pin_memory()
        host_out = torch.empty_like(host_in).pin_memory()
        h2d = d2h = int(n * d * 4)             # whole job: the ranks' chunks add up to the full array

        def e2e_step():
            emb.load_positions_chunk(host_in)  # H2D of the rank's rows + NVLink fan-out (collective)
            one_step()
            emb.read_positions_chunk(host_out) # D2H of the rank's rows + stream sync
        api = ("every rank: load_positions_chunk(pinned rows of its 1/N) -> run_layout_device(1) [replay of the captured "
               "iteration] -> read_positions_chunk(pinned)")
    else:
        host_in = torch.from_numpy(emb.positions).pin_memory()
        h2d = d2h = int(n * d * 4)
        state = {"x": host_in}

        def e2e_step():
            emb.positions = state["x"]         # the reference's setter (embedder_pytorch.py:329-335), pinned source
            out = emb.run_layout(1)            # the reference's driver (:808-833): returns a FRESH ndarray
            state["out"] = out
        api = ("emb.positions = pinned host tensor -> emb.run_layout(1) -> fresh ndarray (the reference's own setter / "
               "driver / return convention)")
    e2e_step()                                 # untimed: first-use allocations of the staging buffers
    barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_e0 = time.perf_counter()
    ea.record()
    for _ in range(Ke):
        e2e_step()
    eb.record()
    torch.cuda.synchronize(dev)
    e2e_wall_ms = (time.perf_counter() - t_e0) * 1e3 / Ke
    barrier()
    e2e_ms = max(ea.elapsed_time(eb) / Ke, e2e_wall_ms)     # host-side copies after the last event count too

    # ---- N > 1: parity of the sharded state against a single-GPU embedder on rank 0 (same seed, same device
    # counter => same samples), from the current positions, over T more iterations
    parity = None
    phase_us = None
    if world > 1:
        T = 3
        cur = emb.positions
        if rank == 0:
            ref_emb = gr.GraphEmbedderPyTorch(adj, **dict(kw, initial_positions=cur))
            if args.sampler == "device":
                ref_emb._buffers()["iter"].copy_(emb._engine.st._iter)
        emb.run_layout_device(T + (T & 1))
        torch.cuda.synchronize(dev)
        mine = emb._pos.clone()
        dist.broadcast(mine, src=0)
        same = torch.tensor([1 if torch.equal(mine, emb._pos) else 0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        if rank == 0:
            if args.sampler == "device":
                ref_emb.run_layout_device(T + (T & 1))
                torch.cuda.synchronize(dev)
                parity = {"iterations": T + (T & 1),
                          "vs": "single-GPU GraphEmbedderPyTorch on rank 0 from the same positions, seed and iteration counter",
                          "samples_equal": bool(torch.equal(ref_emb.last_sampled_indices, emb.last_sampled_indices)),
                          "knn_lists_equal": bool(torch.equal(ref_emb._bufs["knn_idx"], emb._engine.knn_idx)),
                          "pos_rel_inf": rel_inf(emb.positions, ref_emb.positions),
                          "replicas_bit_identical": bool(int(same.item()))}
                parity["ok"] = bool(parity["samples_equal"] and parity["knn_lists_equal"] and parity["replicas_bit_identical"]
                                    and parity["pos_rel_inf"] <= 1e-5 * parity["iterations"])
            else:
                parity = {"replicas_bit_identical": bool(int(same.item())), "ok": bool(int(same.item())),
                          "vs": "replica comparison only (sampler='torch' streams differ between two objects in one process)"}
            ref_emb.close()
        if emb.exchange == "p2p":
            try:
                ph = emb.profile_phases(20)
                vals = [ph[k] for k in emb.PHASES]
            except Exception as exc:  # pylint: disable=broad-exception-caught
                vals, ph = [-1.0] * len(emb.PHASES), {"error": repr(exc)}
            t = torch.tensor(vals, device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            phase_us = {"end_of_phase_us_since_step_start_max_over_ranks": dict(zip(emb.PHASES, [round(float(x), 1) for x in t.tolist()])),
                        "note": "external CUDA events inside the captured iteration; prep and colsum run on the side stream"}
            if "error" in ph:
                phase_us["error"] = ph["error"]
            try:
                kt = emb.profile_kernels(10, scan_ctas=True)
            except Exception as exc:  # pylint: disable=broad-exception-caught
                kt = {"error": repr(exc)}
            phase_us["kernel_begin_end_us_rank0"] = kt
            phase_us["kernel_note"] = ("%globaltimer stamps written by the kernels inside the replayed iteration of rank 0 "
                                       "(first CTA start, last CTA end), relative to the first kernel's start")

    # max over ranks
    if world > 1:
        t = torch.tensor([total_ms, sus_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, sus_ms, e2e_ms = [float(x) for x in t.tolist()]

    # ---- per-kernel roofline (rank 0, single GPU), measured live with CUDA events: gem_profile_step runs the
    # same launches in series on one stream with an event after every stage (in the timed steps above the
    # KNN preparation overlaps the spring kernel on a side stream, so stage times do not add up to the step)
    stage = None
    roof = None
    extra_roof = {}
    peaks, peak_src = measured_peaks()
    if rank == 0 and world == 1:
        runs = []
        for i in range(7):
            flush_buf.fill_(i)
            runs.append(emb.profile_step())
        stage = {k: float(np.median([r[k] for r in runs])) for k in runs[0]}
        fp32_peak = emb.fp32_peak_flops()
        S = min(w["S"], E)
        N = n
        flops = 2.0 * (d + 2) * S * E                        # 2(d+2) flop per query-candidate pair
        batched = S > 1024            # general path (KNN in batches of 1024 queries): no per-stage events inside the KNN
        scan_s = (total_ms / K if batched else stage["knn_scan"]) * 1e-3
        executed = 2.0 * d * S * E                           # what the filter executes: d FMA per pair
        traffic = scan_traffic_from_profiles(args.workload)
        roof = {"kernel": "knn_scan_kernel" if not batched else
                "knn_scan_kernel (batched full-KNN regime: the WHOLE iteration time is used as its duration -- a lower "
                "bound of its rate)", "bound": "fp32", "achieved": flops / scan_s / 1e12,
                "peak": fp32_peak / 1e12, "unit": "TFLOP/s", "frac": flops / scan_s / fp32_peak,
                "traffic": traffic,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full of THIS build "
                                  "(profiles/r02_scan_traffic.json) or null when not captured; algorithmic bytes per launch "
                                  f"= 16*E = {16.0 * E:.0f}",
                "peak_source": "gem_fp32_peak_probe: dependent-chain-free FFMA loop measured in this run "
                               "(MEASURED_PEAKS.json has no FP32 figure); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.5",
                "frac_of_nominal_peak": flops / scan_s / 74.5e12,
                "algorithmic_flops_per_launch": flops, "ms": scan_s * 1e3,
                "executed_fma_tflops": executed / scan_s / 1e12,
                "executed_frac_of_peak": executed / scan_s / fp32_peak,
                "note": "algorithmic = SURVEY 8(d): 2(d+2)*S*E flop (the 5-term cdist chain per pair). The kernel "
                        "executes d FMA + a min/compare per pair as a conservative filter and re-checks the rare "
                        "passes with the exact chain, so `achieved` counts work it does not have to do: "
                        "executed_fma_tflops is the hardware rate"}
        hbm = peaks["hbm_gbs"]
        ka_bytes = 8.0 * E + 4.0 * d * N + 4.0 * d * N + 4.0 * d * E
        kd_bytes = 20.0 * d * N
        extra_roof = {
            "spring_csr_kernel": {"bound": "hbm", "achieved": ka_bytes / (stage["spring_mid"] * 1e-3) / 1e9,
                                  "peak": hbm, "unit": "GB/s",
                                  "frac": ka_bytes / (stage["spring_mid"] * 1e-3) / 1e9 / hbm,
                                  "algorithmic_bytes": ka_bytes, "ms": stage["spring_mid"],
                                  "note": "bound in practice by the L1 wavefront rate of 2E scattered 16-byte position "
                                          "gathers (DESIGN.md section 4a), not by DRAM; the stage includes the column-sum pass"},
            "update (normalise)": {
                "bound": "hbm", "achieved": kd_bytes / (stage["update"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                "frac": kd_bytes / (stage["update"] * 1e-3) / 1e9 / hbm, "algorithmic_bytes": kd_bytes,
                "ms": stage["update"],
                "note": "the `update` stage is the normalisation pass only: the add of pass 1 is done by the spring kernel "
                        "(it writes pos+F) and the column-sum pass is timed inside the spring stage; frac uses the full "
                        "20dN algorithmic bytes of the reference's update and is an upper bound for this stage alone"},
            "spring+update combined": {
                "bound": "hbm", "algorithmic_bytes": ka_bytes + kd_bytes, "ms": stage["spring_mid"] + stage["update"],
                "achieved": (ka_bytes + kd_bytes) / ((stage["spring_mid"] + stage["update"]) * 1e-3) / 1e9, "peak": hbm,
                "unit": "GB/s", "frac": (ka_bytes + kd_bytes) / ((stage["spring_mid"] + stage["update"]) * 1e-3) / 1e9 / hbm},
            "peak_source": peak_src,
        }
        if batched:                    # the stage slots after the query midpoints are not aligned on the general path
            extra_roof = {"peak_source": peak_src}
            stage = {k: stage[k] for k in ("sample", "spring_mid", "query_mid")}
        # whole-iteration roofline: T_roof = B_iter/BW + F_iter/P
        b_iter = 8.0 * E + 8.0 * d * E + 28.0 * d * N
        t_roof = b_iter / (hbm * 1e9) + flops / fp32_peak
        extra_roof["iteration"] = {"t_roof_ms": t_roof * 1e3, "achieved_ms": total_ms / K,
                                   "frac": t_roof * 1e3 / (total_ms / K),
                                   "frac_sustained": t_roof * 1e3 / sus_ms}

    kernel_timeline = None
    if world == 1 and rank == 0:
        try:
            kernel_timeline = emb.profile_kernels(10, scan_ctas=True)
        except Exception as exc:  # pylint: disable=broad-exception-caught
            kernel_timeline = {"error": repr(exc)}
    exchange = emb.exchange if world > 1 else None
    emb_multicast = bool(emb.multicast) if world > 1 else False
    if world > 1:
        emb.close()            # captured graphs hold barrier / NCCL kernels: release them before the communicator
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on this box's host cores (bounded sample), N = 1 only
    cpu = None
    cpu_gc = None
    if not args.no_cpu_baseline and not args.profile_mode and world == 1:
        cb_steps = 2 if E > 1_000_000 else 3
        r = cpu_reference_timing(adj, w, cb_steps, 1, budget_s=25.0)
        cpu = {"value": r["E"] / r["sec_per_step"], "unit": "edge-updates/s", "cores": r["cores"], "kind": "port",
               "sample": r["sample"], "iters_per_s": 1.0 / r["sec_per_step"]}
        r2 = cpu_reference_timing(adj, w, cb_steps, 1, with_gc=True, budget_s=25.0)
        cpu_gc = {"value": r2["E"] / r2["sec_per_step"], "unit": "edge-updates/s", "cores": r2["cores"], "kind": "port",
                  "sample": r2["sample"], "iters_per_s": 1.0 / r2["sec_per_step"]}

    ms_per_step = total_ms / K
    value = E / (ms_per_step * 1e-3)
    # kernels of libgraphem_b200.so per iteration (memset / memcpy nodes and torch / barrier kernels not counted):
    # 1 GPU: knn_prep (sample + query midpoints + bounds + thresholds), spring_csr, column sums, knn_scan,
    # knn_select (+intersection), normalise; N > 1: knn_prep, spring_csr (push), column sums, knn_scan, knn_select
    # (publishing the lists), topk_merge_intersect (+patch/stat publication), normalise_all
    launches_per_step = 6 if world == 1 else (7 if exchange == "p2p" else 12)
    S_eff = min(w["S"], E)
    if world == 1 and S_eff > 1024:
        # general path: sample, spring, query midpoints, hint, per batch of 1024 queries (prep, one scan per 256 queries,
        # select), intersection, two update passes
        nb, rem = divmod(S_eff, 1024)
        launches_per_step = 4 + nb * (2 + 4) + ((2 + (rem + 255) // 256) if rem else 0) + 3
    line = {
        "metric": METRIC, "value": value, "unit": "edge-updates/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "iters_per_s": 1e3 / ms_per_step,
        "config": workload_config(w, n, E),
        "details": {"profile_mode": bool(args.profile_mode), "parallelism": "single GPU" if world == 1 else
                    (f"vertex-sharded x{world} (v mod N): spring kernel pushes pos+F rows to every replica over NVLink "
                     f"({'NVSwitch multicast stores, multimem.st' if emb_multicast else 'unicast peer stores'}), select "
                     f"publishes the partial lists, 2 device barriers, every rank normalises all rows; no NCCL collective"
                     if exchange == "p2p" else f"vertex-sharded x{world}, NCCL all-gather / all-reduce fallback flow"),
                    "step": "one iteration of run_layout's production path: replays of the captured CUDA graphs (one of "
                            "several iterations + one of a single iteration)",
                    "sampler": "device (keyed bijection, gem_knn_prep)" if args.sampler == "device" else
                               "torch.randperm(E)[:S] from the seeded default generator (the reference's RNG stream), drawn one "
                               "iteration ahead on a side stream inside the captured graph",
                    "graph_generation_s": round(t_graph, 2), "constructor_s": round(t_construct, 2)},
        "sustained": {"iterations": sus_iters, "ms_per_step": sus_ms, "seconds": sus_ms * sus_iters * 1e-3,
                      "value": E / (sus_ms * 1e-3), "l2": "not flushed (back-to-back replays)", "clocks": clk2},
        "wall_s_timed_region": t_wall,
        "e2e": {"value": E / (e2e_ms * 1e-3), "unit": "edge-updates/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": api},
        "gpu_launches": launches_per_step * K,
        "timed_steps_ms": {"min": float(np.min(step_ms)), "median": float(np.median(step_ms)), "max": float(np.max(step_ms))},
        "flushed_single_replays": flushed,
        "clocks": clk,
        "roofline": roof,
        "roofline_other": extra_roof,
        "stage_ms": stage,
        "cpu_baseline": cpu,
    }
    if cpu_gc is not None:
        line["cpu_baseline_as_shipped_equivalent"] = cpu_gc
    if parity is not None:
        line["parity"] = parity
    if phase_us is not None:
        line["phase_us"] = phase_us
    if kernel_timeline is not None:
        line["kernel_begin_end_us"] = kernel_timeline      # %globaltimer stamps inside the replayed iteration
    if setup is not None:
        line["graph_setup_s"] = setup          # constructor with given initial positions: device vs host graph build
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--sampler", default="device", choices=["device", "torch"],
                    help="query-edge sampler: the library's keyed bijection, or the reference's torch.randperm stream")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-mode", action="store_true",
                    help="short run for ncu: no sustained region, no CPU baseline, no set-up timing (the line says so)")
    ap.add_argument("--sample-size", type=int, default=None,
                    help="override the workload's sample size S (>= E: the full-KNN regime, SURVEY 8(f).4)")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.sample_size is not None:
        w["S"] = int(args.sample_size)
        w["desc"] = w["desc"].replace("S=256", f"S={args.sample_size}")
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_b200(args, w)


if __name__ == "__main__":
    main()
