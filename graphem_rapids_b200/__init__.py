"""
graphem_rapids_b200 -- B200-native implementation of GraphEm's force-directed layout iteration
behind the reference's Python API (sashakolpakov/graphem-rapids, graphem_rapids/__init__.py).

    import graphem_rapids_b200 as gr
    emb = gr.create_graphem(adjacency, n_components=3)     # or gr.GraphEmbedderPyTorch(...)
    emb.run_layout(num_iterations=50); pos = emb.get_positions()
    seeds = gr.graphem_seed_selection(emb, k=10)

The compute path is libgraphem_b200.so (hand-written sm_100a CUDA, C ABI in
include/graphem_b200.h).  There is no CPU / torch-op fallback: without the library or a CUDA
device the embedder raises.
"""
from .embedder import GraphEmbedderPyTorch, create_graphem
from .influence import graphem_seed_selection
from .correlation import radial_correlations
from . import generators
from .generators import (erdos_renyi_graph, generate_ba, generate_random_regular, generate_sbm)

__version__ = "0.1.0"

__all__ = [
    "GraphEmbedderPyTorch", "create_graphem", "graphem_seed_selection", "radial_correlations", "generators",
    "erdos_renyi_graph", "generate_ba", "generate_random_regular", "generate_sbm",
]
