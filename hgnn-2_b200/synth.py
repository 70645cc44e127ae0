"""Seeded synthetic graph sets in the reference's instance format ``[X, A, t, W, WL, Pm, Pd]``
(functions/data_generator.py:85).  QM9 / rdkit are unavailable offline and the reference has no SBM
generator, so bench.py and the tests use these (distributions fixed in SURVEY.md 8d):

* ``sbm_instance``: 2 equal communities, p_in = a/N, p_out = b/N, A in {0,1}; features = [degree,
  4 x N(0,1)]; label = graph id parity (class 1 swaps a and b).
* ``qm9_shaped_instance``: molecule-shaped graphs (<= 29 atoms, heavy-atom tree + hydrogens, bond
  weights in {1, 1.5, 2, 3}), one-hot (H,C,N,O,F) features, 13 targets ~ N(1,1).
"""
import numpy as np
import torch

from .functions.operators import graph_operators
from .pack import SparseAdj


def _finish(X, A, t, J, sparse, dual=True):
    ops = graph_operators([X, A], J, dual, sparse=sparse)
    return [X, A, t] + list(ops)


def sbm_instance(graph_id, N=1000, a=7.0, b=3.0, J=1, sparse=True, n_feat=5):
    gen = torch.Generator().manual_seed(1000 + graph_id)
    label = graph_id % 2
    if label == 1:
        a, b = b, a
    half = N // 2
    comm = torch.arange(N) >= half
    if N <= 4096:
        same = comm.view(-1, 1) == comm.view(1, -1)
        prob = torch.where(same, torch.tensor(a / N), torch.tensor(b / N))
        up = (torch.rand(N, N, generator=gen) < prob).triu(1)
        iu, ju = up.nonzero(as_tuple=True)
        iu, ju = iu.numpy(), ju.numpy()
    else:   # sparse sampling: binomial edge counts per block, uniform pairs (duplicates dropped)
        rng = np.random.default_rng(1000 + graph_id)
        parts = []
        for (r0, r1, c0, c1, p, tri) in ((0, half, 0, half, a / N, True), (half, N, half, N, a / N, True),
                                         (0, half, half, N, b / N, False)):
            npairs = (r1 - r0) * (r1 - r0 - 1) // 2 if tri else (r1 - r0) * (c1 - c0)
            m = rng.binomial(npairs, p)
            i = rng.integers(r0, r1, size=m)
            j = rng.integers(c0, c1, size=m)
            lo, hi = np.minimum(i, j), np.maximum(i, j)
            keep = lo != hi
            parts.append(np.stack([lo[keep], hi[keep]], 1))
        e = np.unique(np.concatenate(parts, 0), axis=0)
        iu, ju = e[:, 0], e[:, 1]
    rows = np.concatenate([iu, ju])
    cols = np.concatenate([ju, iu])
    vals = np.ones(rows.shape[0], dtype=np.float32)
    deg = np.bincount(rows, minlength=N).astype(np.float32)
    X = torch.cat([torch.from_numpy(deg).view(-1, 1), torch.randn(N, n_feat - 1, generator=gen)], 1)
    A = SparseAdj(N, rows, cols, vals)
    if not sparse:
        A = A.to_dense()
    t = torch.tensor([label], dtype=torch.int64)
    return _finish(X, A, t, J, sparse)


def sbm_dataset(n_graphs, N=1000, a=7.0, b=3.0, J=1, sparse=True, first_id=0):
    return [sbm_instance(first_id + i, N, a, b, J, sparse) for i in range(n_graphs)]


def qm9_shaped_instance(graph_id, J=1, sparse=True, dual=True):
    rng = np.random.default_rng(2000 + graph_id)
    n = int(np.clip(np.rint(rng.normal(18, 3)), 3, 29))
    n_heavy = max(1, min(n, int(np.rint(n / 2))))
    A = np.zeros((n, n), dtype=np.float32)
    valence = np.zeros(n, dtype=np.int64)
    bond_w = np.array([1.0, 1.5, 2.0, 3.0], dtype=np.float32)
    for v in range(1, n_heavy):                      # random heavy-atom tree, valence <= 4
        cand = [u for u in range(v) if valence[u] < 4]
        u = int(rng.choice(cand)) if cand else int(rng.integers(0, v))
        w = bond_w[rng.choice(4, p=[0.8, 0.1, 0.08, 0.02])]
        A[u, v] = A[v, u] = w
        valence[u] += 1
        valence[v] += 1
    for _ in range(int(rng.integers(0, 3))):         # 0-2 ring closures
        if n_heavy >= 3:
            u, v = rng.choice(n_heavy, 2, replace=False)
            if A[u, v] == 0 and valence[u] < 4 and valence[v] < 4:
                A[u, v] = A[v, u] = 1.0
                valence[u] += 1
                valence[v] += 1
    for h in range(n_heavy, n):                      # hydrogens on heavy atoms with free valence
        cand = [u for u in range(n_heavy) if valence[u] < 4]
        u = int(rng.choice(cand)) if cand else int(rng.integers(0, n_heavy))
        A[u, h] = A[h, u] = 1.0
        valence[u] += 1
    X = np.zeros((n, 5), dtype=np.float32)
    X[np.arange(n_heavy), rng.choice([1, 2, 3, 4], size=n_heavy, p=[0.7, 0.1, 0.15, 0.05])] = 1.0
    X[n_heavy:, 0] = 1.0
    t = torch.from_numpy(rng.normal(1.0, 1.0, size=13).astype(np.float32))
    return _finish(torch.from_numpy(X), torch.from_numpy(A), t, J, sparse, dual)


def qm9_shaped_dataset(n_graphs, J=1, sparse=True, first_id=0):
    return [qm9_shaped_instance(first_id + i, J, sparse) for i in range(n_graphs)]
