"""``graph_operators`` - mirror of the reference's functions/operators.py:11-83.

The reference fills dense ``W (N,N,J+2)``, ``WL (M,M,J+2)``, ``Pm, Pd (N,M)`` with Python loops,
O(M^2) for the line graph (1 s at N=50, ~5 min at N=1000).  Here the non-zeros are
enumerated on the host in O(nnz) (sparse_ops.GraphOps - bit-exact, quirks included), the powers
``A^(2^j)`` are squared by the SpGEMM kernel and, when dense tensors are asked for, they are
scattered on the GPU (hgnn_csr_to_dense).
"""
import numpy as np
import torch

from ..pack import BatchPack, GraphHandle, OperatorHandle, SparseAdj
from ..sparse_ops import GraphOps


def graph_ops_of(A, dual=True):
    """Host sparse operators of an adjacency given as a dense tensor/array or a ``SparseAdj``."""
    if isinstance(A, SparseAdj):
        return GraphOps.from_coo(A.N, A.rows, A.cols, A.vals, dual=dual)
    if torch.is_tensor(A):
        A = A.detach().cpu().numpy()
    return GraphOps.from_dense(np.asarray(A, dtype=np.float32), dual=dual)


def graph_operators(graph, J=1, dual=False, sparse=False):
    """Operators of G = (V, A): I, D, A, ..., A^(2^(J-1)) and, with ``dual``, the line-graph twins
    plus Pm / Pd.  Same signature and return order as the reference (:11, :33, :83).

    sparse=False (default): dense CPU tensors, bit-identical to the reference.
    sparse=True: ``GraphHandle`` objects sharing one host ``GraphOps`` (nothing dense is built);
    ``prepare_batch`` consumes either form.
    """
    V, A = graph
    g = graph_ops_of(A, dual=dual)
    if int(V.shape[0]) != g.N:
        raise ValueError("V has %d rows but A is %d x %d" % (V.shape[0], g.N, g.N))
    if sparse:
        if not dual:
            return GraphHandle(g, "W", J)
        return tuple(GraphHandle(g, n, J) for n in ("W", "WL", "Pm", "Pd"))
    pack = BatchPack.from_graphs([g], J, dual=dual)
    W = pack.dense_W()[0].cpu()
    if not dual:
        return W
    return W, pack.dense_WL()[0].cpu(), pack.dense_P(False)[0].cpu(), pack.dense_P(True)[0].cpu()
