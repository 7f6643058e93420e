"""
Vectorised O(E) synthetic graph generators for the BASELINE.json configurations.

The reference's generators (graphem_rapids/generators.py) wrap networkx and return
`scipy.sparse.csr_matrix` int adjacency (symmetric, 0/1, no self loops; generators.py:13-15).
networkx is far too slow at 1M-10M vertices (erdos_renyi_graph is O(n^2)), so the bench and the
large parity tests use these numpy generators with documented seeds.  They return the same
kind of object; they do NOT reproduce networkx's random streams (parity is defined on an
identical adjacency, not on the generator).
"""
import numpy as np
import scipy.sparse as sp


def _sym_from_sorted_upper(lo, hi, n):
    """Symmetric 0/1 CSR adjacency from the distinct pairs lo < hi, already sorted by (lo, hi): the upper triangle IS
    a canonical CSR (indptr = offsets of lo, indices = hi); its transpose (scipy's O(E) counting sort in C) is the
    lower triangle with sorted rows, and the canonical sum of both is the adjacency.  No Python-level sort."""
    e = len(lo)
    idx_t = np.int32 if n < 2 ** 31 and 2 * e < 2 ** 31 else np.int64
    indptr = np.concatenate([[0], np.cumsum(np.bincount(lo, minlength=n))]).astype(idx_t)
    upper = sp.csr_matrix((np.ones(e, dtype=np.int64), hi.astype(idx_t), indptr), shape=(n, n))
    upper.has_sorted_indices = True
    upper.has_canonical_format = True
    lower = upper.T.tocsr()
    lower.has_sorted_indices = True
    lower.has_canonical_format = True
    adj = (upper + lower).tocsr()
    adj.sort_indices()
    return adj


def _to_adjacency(src, dst, n):
    """Undirected simple graph from endpoint arrays: drop loops, dedupe, symmetrise."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    keep = src != dst
    lo = np.minimum(src[keep], dst[keep])
    hi = np.maximum(src[keep], dst[keep])
    key = np.unique(lo * n + hi)                      # sorted by (lo, hi)
    return _sym_from_sorted_upper(key // n, key % n, n)


_ER_SORTED_MIN_EDGES = 5_000_000


def _er_sorted_pairs(n, m, rng):
    """m (minus the rare collisions) distinct unordered pairs of [0, n), SORTED by (i, j), without a sort: the order
    statistics of m uniforms are the normalised partial sums of m+1 exponentials; each value picks the pair with
    that rank in the row-major enumeration of the upper triangle."""
    total = n * (n - 1) // 2
    c = np.cumsum(rng.exponential(size=m + 1))
    r = np.floor(c[:-1] * (total / c[-1])).astype(np.int64)
    r = np.minimum(r, total - 1)
    r = r[np.concatenate([[True], r[1:] != r[:-1]])]              # collisions are consecutive
    # rank -> (i, j): row i starts at off(i) = i*(2n-i-1)/2
    t = 2.0 * n - 1.0
    i = np.floor((t - np.sqrt(np.maximum(t * t - 8.0 * r.astype(np.float64), 0.0))) / 2.0).astype(np.int64)
    i = np.clip(i, 0, n - 2)
    for _ in range(3):                                             # float64 rounding: fix by +-1
        off = i * (2 * n - i - 1) // 2
        i = np.where(off > r, i - 1, i)
        off_next = (i + 1) * (2 * n - i - 2) // 2
        i = np.where(off_next <= r, i + 1, i)
    off = i * (2 * n - i - 1) // 2
    j = i + 1 + (r - off)
    assert np.all((off <= r) & (j < n) & (j > i))
    return i, j


def erdos_renyi_graph(n, p, seed=0):
    """G(n, p): draw m ~ Binomial(n(n-1)/2, p) distinct unordered pairs (generators.py:32-49).  Small graphs
    (fewer than 5 M edges) keep round 1's rejection sampler (same graphs as before); large ones take the sort-free
    order-statistics sampler."""
    rng = np.random.default_rng(seed)
    total = n * (n - 1) // 2
    m = int(rng.binomial(total, p)) if total < 2 ** 62 else int(total * p)
    if m >= _ER_SORTED_MIN_EDGES:
        lo, hi = _er_sorted_pairs(n, m, rng)
        return _sym_from_sorted_upper(lo, hi, n)
    got = np.empty(0, dtype=np.int64)
    while len(got) < m:
        need = int((m - len(got)) * 1.1) + 16
        a = rng.integers(0, n, need, dtype=np.int64)
        b = rng.integers(0, n, need, dtype=np.int64)
        ok = a != b
        key = np.minimum(a[ok], b[ok]) * n + np.maximum(a[ok], b[ok])
        got = np.unique(np.concatenate([got, key]))
    if len(got) > m:
        got = rng.permutation(got)[:m]
    return _to_adjacency(got // n, got % n, n)


def generate_ba(n=300, m=3, seed=0):
    """Barabasi-Albert preferential attachment (generators.py:112-129), Batagelj-Brandes linear
    scheme vectorised by pointer jumping: edge t (source t//m + m) picks a uniformly random
    earlier endpoint slot; odd slots copy the target stored in the slot they point to.
    Multi-edges are merged, so E is slightly below m(n-m)."""
    rng = np.random.default_rng(seed)
    nv = n - m
    ne = nv * m
    # slot 2t = source of edge t, slot 2t+1 = its target.  Seed: the first vertex attaches to 0..m-1.
    src = np.repeat(np.arange(m, n, dtype=np.int64), m)
    tgt = np.empty(ne, dtype=np.int64)
    tgt[:m] = np.arange(m)
    # edge t >= m draws a slot among the 2*m*(t//m) slots of earlier vertices
    t = np.arange(m, ne, dtype=np.int64)
    hi = 2 * m * (t // m)
    slot = (rng.random(ne - m) * hi).astype(np.int64)
    ptr = np.full(ne, -1, dtype=np.int64)          # >=0: copy the target of edge ptr
    even = (slot % 2) == 0
    tgt[m:][even] = src[slot[even] // 2]
    ptr[m:][~even] = slot[~even] // 2
    unresolved = np.nonzero(ptr >= 0)[0]
    while len(unresolved):
        p = ptr[unresolved]
        nxt = ptr[p]                                # read before any update of this round
        done = nxt < 0
        tgt[unresolved[done]] = tgt[p[done]]
        ptr[unresolved[done]] = -1
        ptr[unresolved[~done]] = nxt[~done]         # jump
        unresolved = unresolved[~done]
    return _to_adjacency(src, tgt, n)


def generate_random_regular(n=100, d=3, seed=0):
    """Random d-regular-like graph (generators.py:235-252): union of d/2 random Hamiltonian cycles
    (plus a random perfect matching if d is odd); duplicate edges are merged, so a few vertices can
    have degree d-1 or d-2."""
    rng = np.random.default_rng(seed)
    src, dst = [], []
    for _ in range(d // 2):
        perm = rng.permutation(n)
        src.append(perm)
        dst.append(np.roll(perm, -1))
    if d % 2:
        perm = rng.permutation(n - (n % 2))
        src.append(perm[0::2])
        dst.append(perm[1::2])
    return _to_adjacency(np.concatenate(src), np.concatenate(dst), n)


def generate_sbm(n_per_block=75, num_blocks=4, p_in=0.15, p_out=0.01, labels=False, seed=0):
    """Stochastic block model (generators.py:67-109): Binomial edge counts per block pair, uniform
    endpoints inside the blocks."""
    rng = np.random.default_rng(seed)
    n = n_per_block * num_blocks
    src, dst = [], []
    for a in range(num_blocks):
        for b in range(a, num_blocks):
            if a == b:
                cnt = int(rng.binomial(n_per_block * (n_per_block - 1) // 2, p_in))
            else:
                cnt = int(rng.binomial(n_per_block * n_per_block, p_out))
            if cnt == 0:
                continue
            src.append(rng.integers(0, n_per_block, cnt, dtype=np.int64) + a * n_per_block)
            dst.append(rng.integers(0, n_per_block, cnt, dtype=np.int64) + b * n_per_block)
    adj = _to_adjacency(np.concatenate(src) if src else [], np.concatenate(dst) if dst else [], n)
    if labels:
        return adj, np.repeat(np.arange(num_blocks), n_per_block)
    return adj
