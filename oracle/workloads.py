"""TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product path.

Dense twins of the seeded synthetic workloads of ``hgnn_b200.synth`` (binary SBM, QM9-shaped
molecular graphs) for the CPU baseline and the reference arm of bench.py: the same graphs, generated WITHOUT
importing the CUDA package (``bench.py --impl reference`` must not load libhgnn_b200.so).
tests/test_oracle_golden.py::test_workloads_match_synth pins them to ``synth`` bit for bit."""
import numpy as np
import torch


def sbm_dense(graph_id, N=1000, a=7.0, b=3.0, n_feat=5):
    """[X (N, n_feat), A (N, N) dense in {0,1}, t (1,) int64] - same draws as synth.sbm_instance (N <= 4096)."""
    if N > 4096:
        raise ValueError("the dense twin follows synth.sbm_instance's dense sampling path (N <= 4096)")
    gen = torch.Generator().manual_seed(1000 + graph_id)
    label = graph_id % 2
    if label == 1:
        a, b = b, a
    half = N // 2
    comm = torch.arange(N) >= half
    same = comm.view(-1, 1) == comm.view(1, -1)
    prob = torch.where(same, torch.tensor(a / N), torch.tensor(b / N))
    up = (torch.rand(N, N, generator=gen) < prob).triu(1)
    A = (up | up.t()).float()
    deg = A.sum(1)
    X = torch.cat([deg.view(-1, 1), torch.randn(N, n_feat - 1, generator=gen)], 1)
    return [X, A, torch.tensor([label], dtype=torch.int64)]


def qm9_shaped_dense(graph_id):
    """[X (n, 5) one-hot (H,C,N,O,F), A (n, n) bond weights, t (13,)] - same draws as synth.qm9_shaped_instance."""
    rng = np.random.default_rng(2000 + graph_id)
    n = int(np.clip(np.rint(rng.normal(18, 3)), 3, 29))
    n_heavy = max(1, min(n, int(np.rint(n / 2))))
    A = np.zeros((n, n), dtype=np.float32)
    valence = np.zeros(n, dtype=np.int64)
    bond_w = np.array([1.0, 1.5, 2.0, 3.0], dtype=np.float32)
    for v in range(1, n_heavy):
        cand = [u for u in range(v) if valence[u] < 4]
        u = int(rng.choice(cand)) if cand else int(rng.integers(0, v))
        w = bond_w[rng.choice(4, p=[0.8, 0.1, 0.08, 0.02])]
        A[u, v] = A[v, u] = w
        valence[u] += 1
        valence[v] += 1
    for _ in range(int(rng.integers(0, 3))):
        if n_heavy >= 3:
            u, v = rng.choice(n_heavy, 2, replace=False)
            if A[u, v] == 0 and valence[u] < 4 and valence[v] < 4:
                A[u, v] = A[v, u] = 1.0
                valence[u] += 1
                valence[v] += 1
    for h in range(n_heavy, n):
        cand = [u for u in range(n_heavy) if valence[u] < 4]
        u = int(rng.choice(cand)) if cand else int(rng.integers(0, n_heavy))
        A[u, h] = A[h, u] = 1.0
        valence[u] += 1
    X = np.zeros((n, 5), dtype=np.float32)
    X[np.arange(n_heavy), rng.choice([1, 2, 3, 4], size=n_heavy, p=[0.7, 0.1, 0.15, 0.05])] = 1.0
    X[n_heavy:, 0] = 1.0
    t = torch.from_numpy(rng.normal(1.0, 1.0, size=13).astype(np.float32))
    return [torch.from_numpy(X), torch.from_numpy(A), t]
