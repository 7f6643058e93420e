"""graphem_seed_selection -- graphem_rapids/influence.py:10-37 on the B200 embedder.
(The NDlib simulation helpers of that file are third-party driven and out of scope.)"""
import numpy as np


def graphem_seed_selection(embedder, k, num_iterations=20):
    """Run the layout, then return the k vertices with the largest radial distance
    (influence.py:28-37: `np.argsort(-radial_distances)[:k].tolist()`).

    SURVEY 8(f).3: with a device embedder the radial norm and the top-k selection run on the device, so
    a 10 M-vertex layout is never copied to the host just to pick k seeds; the result is the same
    python list[int] (descending radius)."""
    import torch
    pos_dev = None
    if hasattr(embedder, "run_layout_device") and hasattr(embedder, "_positions"):
        embedder.run_layout_device(num_iterations) if num_iterations > 0 else None
        pos_dev = embedder._positions
    else:
        embedder.run_layout(num_iterations=num_iterations)
    if isinstance(pos_dev, torch.Tensor) and pos_dev.is_cuda:
        radial = torch.linalg.vector_norm(pos_dev, dim=1)
        k = min(int(k), radial.numel())
        return torch.topk(radial, k, largest=True, sorted=True).indices.tolist()
    positions = np.array(embedder.positions)
    radial = np.linalg.norm(positions, axis=1)
    return np.argsort(-radial)[:k].tolist()
