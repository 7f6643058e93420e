"""torchrun --nproc-per-node G scripts/shard_profile.py [workload]: per-phase CUDA-event timing of the sharded step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from graphem_rapids_b200.sharded import ShardedGraphEmbedder
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
adj = bench.make_graph(w)
emb = ShardedGraphEmbedder(adj, n_components=w["d"], device=dev, n_neighbors=w["k"], sample_size=w["S"], verbose=False,
                           seed=0, initial_positions=bench.initial_positions(adj.shape[0], w["d"]), use_cuda_graph=False,
                           use_symmetric_memory=False)
g = emb._engine
L = g.L
print(f"[{rank}] rows {g.ve - g.vb} of slice {L.slice} edges {g.e_hi - g.e_lo} of {L.n_edges} hubs {emb._hubs.numel()}", flush=True)
for _ in range(5):
    emb.update_positions()
names = ["phase_a", "allgather_lists", "phase_b", "allreduce_stats", "phase_c", "allgather_pos"]
acc = np.zeros(len(names)); iters = 30
for it in range(iters):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    dist.barrier(); torch.cuda.synchronize()
    ev[0].record(); g.phase_a(); ev[1].record()
    dist.all_gather_into_tensor(g.gathered.view(-1), g.part); ev[2].record()
    g.phase_b(); ev[3].record()
    dist.all_reduce(g.stats); ev[4].record()
    g.phase_c(); ev[5].record()
    dist.all_gather_into_tensor(g.pos.view(-1), g.own_block().view(-1)); ev[6].record()
    torch.cuda.synchronize()
    acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))])
print(f"[{rank}] " + " ".join(f"{n} {1e3 * a / iters:.1f}us" for n, a in zip(names, acc)) + f" | total {1e3 * acc.sum() / iters:.1f}us", flush=True)
dist.destroy_process_group()
