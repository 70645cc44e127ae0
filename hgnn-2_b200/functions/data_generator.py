"""Synthetic "three collinear points" classification set - mirror of the reference's
functions/data_generator.py (:19-87).  Same instance format ``[X, A, y, W, WL, Pm, Pd]`` (:85) and
the same sampling recipe; the operators come from this package's ``graph_operators``.
``sparse=True`` stores operator handles instead of dense tensors.  The SBM and QM9-shaped
generators used by the benchmark live in ``hgnn_b200.synth``.
"""
import os
import pickle
from random import shuffle

import torch

from .operators import graph_operators

save_path = os.environ.get("HGNN_DATA_PATH", "data/generated")


def three_collinear_points(n, Nmax, d, p, c, sparse=False):
    """n random graphs with 3..Nmax-1 nodes; with probability p three of the node feature vectors
    are collinear (label 1); adjacency entries are on with probability 1-c, symmetrised and
    clipped to 1, edge (0,1) always present (reference :45-87)."""
    data = []
    y = torch.rand(n) < p
    sizes = torch.randint(low=0, high=Nmax - 3, size=[n])
    for i in range(n):
        extra = int(sizes[i].item())
        total = extra + 3
        if y[i] == 1:
            direction = torch.randn(1, d)
            pts = [10 * torch.randn(1) * direction for _ in range(3)]
            pool = torch.cat([torch.randn(extra, d)] + pts, dim=0)
            order = list(range(total))
            shuffle(order)
            X = pool[torch.tensor(order)]
        else:
            X = torch.randn(total, d)
        A = (torch.rand(total, total) > c).float()
        A[0, 1] = 1.0
        A = torch.min(A + A.t(), torch.ones(total))
        W, WL, Pm, Pd = graph_operators([X, A], dual=True, sparse=sparse)
        data.append([X, A, torch.tensor([int(y[i].item())], dtype=torch.int64), W, WL, Pm, Pd])
    return data


def load_graph_sets(n=1000, Nmax=50, d=5, p=0.5, c=0.5):
    """80/10/10 split pickled under ``save_path`` (reference :19-42)."""
    n_train, n_valid = int(0.8 * n), int(0.1 * n)
    data = three_collinear_points(n, Nmax, d, p, c)
    os.makedirs(save_path, exist_ok=True)
    splits = {"cp_train_%d" % n_train: data[:n_train],
              "cp_valid_%d" % n_valid: data[n_train:n_train + n_valid],
              "cp_test_%d" % (n - n_valid - n_train): data[n_train + n_valid:]}
    for name, part in splits.items():
        with open(os.path.join(save_path, name + ".pickle"), "wb") as f:
            pickle.dump(part, f)
