"""CPU: pin oracle/ (the restatement) against golden vectors recorded from the REAL reference
(tests/golden/make_golden.py).  Bit-exact wherever the reference is deterministic."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle


def _t(g):
    pos = torch.from_numpy(g["pos0"])
    edges = torch.from_numpy(g["edges"].astype(np.int64))
    samp = torch.from_numpy(g["samp"])
    par = dict(n_neighbors=int(g["n_neighbors"]), k_attr=float(g["k_attr"]), L_min=float(g["L_min"]),
               k_inter=float(g["k_inter"]))
    return pos, edges, samp, par


def test_spring_and_midpoints_bit_exact(golden):
    pos, edges, _, par = _t(golden)
    F = oracle.spring_forces(pos, edges, par["k_attr"], par["L_min"])
    assert np.array_equal(F.numpy(), golden["F_spring"])
    assert np.array_equal(oracle.midpoints(pos, edges).numpy(), golden["mid"])


def test_knn_literal_bit_exact(golden):
    """cdist + topk restated literally returns the very indices the reference returned."""
    _, _, samp, par = _t(golden)
    mid = torch.from_numpy(golden["mid"])
    knn = oracle.knn_reference(mid[samp], mid, par["n_neighbors"] + 1, 1 << 30)
    assert np.array_equal(knn.numpy(), golden["knn_full"].astype(np.int64))


def rows_match_modulo_ties(idx_a, dist_a, idx_b, dist_b, ulps=1):
    """Row sets are equal, or differ only in elements whose distance is within `ulps` of the
    row's boundary (k+1-th) distance -- torch.topk breaks ties arbitrarily and torch's CPU
    sqrt is not correctly rounded (1 ulp)."""
    bad = []
    for r in range(idx_a.shape[0]):
        sa, sb = set(idx_a[r].tolist()), set(idx_b[r].tolist())
        if sa == sb:
            continue
        bound = max(dist_a[r].max(), dist_b[r].max())
        tol = ulps * np.spacing(np.float32(bound))
        da = {i: d for i, d in zip(idx_a[r].tolist(), dist_a[r].tolist())}
        db = {i: d for i, d in zip(idx_b[r].tolist(), dist_b[r].tolist())}
        for i in sa ^ sb:
            dd = da.get(i, db.get(i))
            if abs(dd - bound) > tol:
                bad.append((r, i, dd, bound))
    return bad


def test_knn_strict_matches_reference_modulo_ties(golden):
    """The (distance, index) ordered KNN (the CUDA path's contract, C restatement with true
    fmaf) selects the reference's neighbour sets; differences only among boundary ties."""
    _, _, samp, par = _t(golden)
    mid = torch.from_numpy(golden["mid"])
    idx, dist = oracle.knn_strict(mid, samp, par["n_neighbors"] + 1)
    bad = rows_match_modulo_ties(idx.numpy(), dist.numpy(), golden["knn_full"].astype(np.int64),
                                 golden["knn_fdist"])
    assert not bad, bad[:5]
    # distances agree with the reference's cdist to 1 ulp (torch CPU sqrt is not correctly rounded)
    ref_sorted = np.sort(golden["knn_fdist"], axis=1)
    assert np.all(np.abs(dist.numpy() - ref_sorted) <= np.spacing(np.maximum(ref_sorted, np.float32(1e-30))))


def test_chain_numpy_equals_c_and_torch(golden):
    """cdist arithmetic: numpy FMA-chain emulation == torch.cdist**2 pattern == C fmaf chain."""
    _, _, samp, _ = _t(golden)
    mid = golden["mid"]
    q = mid[samp.numpy()[:32]]
    if oracle.uses_mm_mode(len(samp), mid.shape[0]):
        sq = oracle.cdist_chain_sq(q, mid)
    else:
        sq = oracle.cdist_direct_sq(q, mid)
    ref = torch.cdist(torch.from_numpy(mid[samp.numpy()]), torch.from_numpy(mid))[:32].numpy()
    mine = np.sqrt(sq)                     # numpy sqrt is correctly rounded, torch's is within 1 ulp
    assert np.all(np.abs(mine - ref) <= np.spacing(np.maximum(ref, np.float32(1e-30))))
    if oracle.uses_mm_mode(len(samp), mid.shape[0]):
        # matmul mode ends in torch's vectorised sqrt_: against it the chain is bit-equal
        assert np.array_equal(torch.from_numpy(sq).sqrt().numpy(), ref)
    else:
        # direct mode ends in std::sqrt (correctly rounded): bit-equal to numpy's sqrt
        assert np.array_equal(mine, ref)


def test_intersection_forces_bit_exact(golden):
    pos, edges, samp, par = _t(golden)
    knn = torch.from_numpy(golden["knn_full"].astype(np.int64))[:, 1:]
    G = oracle.intersection_forces(pos, edges, knn, samp, par["k_inter"])
    assert np.array_equal(G.numpy(), golden["F_inter"])


def test_full_step_and_trajectory_bit_exact(golden):
    pos, edges, samp, par = _t(golden)
    out = oracle.layout_step(pos, edges, samp, strict=False, **par)
    assert np.array_equal(out["new_pos"].numpy(), golden["new_pos"])
    samples = [torch.from_numpy(s) for s in golden["traj_samps"]]
    fin = oracle.run_layout(pos, edges, len(samples), sample_size=int(golden["sample_size"]),
                            samples=samples, strict=False, **par)
    assert np.array_equal(fin.numpy(), golden["traj_pos"])


def test_strict_step_close_to_reference(golden):
    """With the strict order the step differs from the reference only through tie choices."""
    pos, edges, samp, par = _t(golden)
    out = oracle.layout_step(pos, edges, samp, strict=True, **par)
    ref = golden["new_pos"]
    same_sets = all(set(a) == set(b) for a, b in zip(out["knn_full"].numpy().tolist(),
                                                     golden["knn_full"].astype(np.int64).tolist()))
    same_first = np.array_equal(out["knn_full"].numpy()[:, 0], golden["knn_full"][:, 0].astype(np.int64))
    if same_sets and same_first:
        err = np.abs(out["new_pos"].numpy() - ref).max() / np.abs(ref).max()
        assert err <= 1e-6


def test_kp1_larger_than_E_raises():
    mid = torch.randn(5, 2)
    with pytest.raises(RuntimeError):
        oracle.knn_strict(mid, torch.arange(5), 7)
    with pytest.raises(RuntimeError):
        oracle.knn_reference(mid, mid, 7, 100)


# ----------------------------------------------------------------------------- 50-iteration runs of the real reference
LONG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "long")
LONG_CASES = sorted(f[:-4] for f in os.listdir(LONG_DIR) if f.endswith(".npz"))


def long_case_inputs(z):
    """(edges int64 (E,2), pos0 (n,d)) of a long golden case; the 100 K case stores neither: graph and
    initial positions are the repo's deterministic ones (tests/golden/make_golden_long.py)."""
    if "pos0" in z.files:
        return z["edges"].astype(np.int64), z["pos0"]
    import graphem_rapids_b200.generators as gen
    n, d = int(z["n"]), int(z["d"])
    edges = oracle.extract_edges(gen.generate_ba(n, 4, seed=0)).astype(np.int64)
    pos0 = (np.random.default_rng(0).standard_normal((n, d)) * 0.1).astype(np.float32)
    return edges, pos0


@pytest.mark.parametrize("name", LONG_CASES)
def test_oracle_replays_reference_50_iterations(name):
    """tests/golden/make_golden_long.py: with the samples the reference drew, the oracle lands on the
    reference's final positions (and therefore on its Spearman(radius, degree/betweenness))."""
    from scipy.stats import spearmanr
    z = np.load(os.path.join(LONG_DIR, name + ".npz"))
    edges, pos0 = long_case_inputs(z)
    samples = [torch.from_numpy(s.astype(np.int64)) for s in z["samples"]]
    out = oracle.run_layout(torch.from_numpy(pos0), torch.from_numpy(edges), len(samples),
                            sample_size=int(z["sample_size"]), n_neighbors=int(z["n_neighbors"]), strict=True,
                            samples=samples).numpy()
    if "final_pos" in z.files:
        assert np.abs(out - z["final_pos"]).max() <= 1e-5 * np.abs(z["final_pos"]).max()
    r = np.linalg.norm(out, axis=1)
    # small cases replay bit for bit; at 100 K vertices torch.topk's tie order / threaded index_add_ make the
    # reference's own trajectory differ from the strict (distance, index) replay -- within the north star's 0.01
    tol = 1e-6 if "final_pos" in z.files else 0.01
    assert abs(spearmanr(r, z["degree"]).correlation - float(z["rho_degree"])) <= tol
    assert abs(spearmanr(r, z["betweenness"]).correlation - float(z["rho_betweenness"])) <= tol
