"""torchrun --nproc-per-node 2 scripts/symm_probe.py : does torch symmetric memory work here (peer pointers, barrier, CUDA graph)?"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
def log(*a): print(f"[{rank}]", *a, flush=True)
try:
    log("backend", symm_mem.get_backend(dev) if hasattr(symm_mem, "get_backend") else None)
    t = symm_mem.empty((1024, 4), dtype=torch.float32, device=dev)
    t.fill_(float(rank + 1))
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    log("rendezvous ok; world", hdl.world_size, "rank", hdl.rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", hdl.has_multicast_support if hasattr(hdl, "has_multicast_support") else None)
    log("multicast_ptr", hex(int(hdl.multicast_ptr or 0)))
    torch.cuda.synchronize(); dist.barrier()
    # push my rows [rank*512, (rank+1)*512) into every peer's buffer
    for r in range(world):
        peer = hdl.get_buffer(r, (1024, 4), torch.float32)
        peer[rank * 512:(rank + 1) * 512].copy_(t[rank * 512:(rank + 1) * 512])
    hdl.barrier(channel=0)
    torch.cuda.synchronize()
    log("after push+barrier: rows", t[0, 0].item(), t[600, 0].item(), "(expect 1.0 2.0)")
    # the same inside a CUDA graph
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        t[rank * 512:(rank + 1) * 512] += 10.0
        for r in range(world):
            peer = hdl.get_buffer(r, (1024, 4), torch.float32)
            peer[rank * 512:(rank + 1) * 512].copy_(t[rank * 512:(rank + 1) * 512])
        hdl.barrier(channel=0)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize(); dist.barrier()
    with torch.cuda.graph(g):
        t[rank * 512:(rank + 1) * 512] += 10.0
        for r in range(world):
            peer = hdl.get_buffer(r, (1024, 4), torch.float32)
            peer[rank * 512:(rank + 1) * 512].copy_(t[rank * 512:(rank + 1) * 512])
        hdl.barrier(channel=0)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    log("after 1 eager + 3 graph replays: rows", t[0, 0].item(), t[600, 0].item(), "(expect 41.0 42.0)")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): g.replay()
    b.record(); torch.cuda.synchronize()
    log(f"graph replay (add + {world} pushes of 8 KB + barrier): {a.elapsed_time(b) / 50 * 1e3:.1f} us")
    del g
except Exception as e:
    import traceback; traceback.print_exc()
    log("FAILED", repr(e))
torch.cuda.synchronize()
dist.destroy_process_group()
log("done")
