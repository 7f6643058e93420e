"""graphem_seed_selection -- graphem_rapids/influence.py:10-37 on the B200 embedder.
(The NDlib simulation helpers of that file are third-party driven and out of scope.)"""
import numpy as np


def graphem_seed_selection(embedder, k, num_iterations=20):
    """Run the layout, then return the k vertices with the largest radial distance
    (influence.py:28-37: `np.argsort(-radial_distances)[:k].tolist()`)."""
    embedder.run_layout(num_iterations=num_iterations)
    positions = np.array(embedder.positions)
    radial = np.linalg.norm(positions, axis=1)
    return np.argsort(-radial)[:k].tolist()
