"""torchrun --nproc-per-node 2 scripts/shard_debug.py : progress markers through the sharded path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import graphem_rapids_b200 as gr
from graphem_rapids_b200.sharded import ShardedGraphEmbedder
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
def mark(m):
    torch.cuda.synchronize(); print(f"[{rank}] {time.time():.1f} {m}", flush=True)
n = 200000
adj = gr.generate_ba(n, 4, seed=1)
pos0 = np.random.default_rng(2).standard_normal((n, 3)).astype(np.float32)
emb = ShardedGraphEmbedder(adj, n_components=3, device=dev, n_neighbors=10, sample_size=256, verbose=False, seed=4,
                           initial_positions=pos0, use_cuda_graph=("graph" in sys.argv))
mark("constructed")
for i in range(3):
    emb.update_positions(); mark(f"eager step {i}")
if "graph" in sys.argv:
    emb.run_layout_device(5); mark("graph replay done")
torch.cuda.synchronize(); t0 = time.time(); emb.run_layout_device(20); torch.cuda.synchronize(); mark(f"20 more steps {1e3*(time.time()-t0)/20:.3f} ms/step")
print(f"[{rank}] pos checksum {float(emb._pos.double().sum()):.6f}", flush=True)
emb.close(); mark('closed')
dist.destroy_process_group(); print(f'[{rank}] destroyed', flush=True)
