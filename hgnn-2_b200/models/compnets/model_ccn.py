"""``CCN_1D`` / ``CCN_2D`` - mirror of the reference's models/compnets/model_ccn.py (:18-64, :68-105).

Same constructor signatures, parameter names (``w1..wL``, ``fc``) and init scales; ``forward(X, adj)``
processes one graph like the reference (scripts/train_ccn.py:49).  ``forward_batch`` runs a whole
list of graphs through the same kernels in one launch per level (the reference has no batching;
SURVEY.md section 8e "CCN").
"""
import torch
import torch.nn as nn

from ... import ops
from ...functions.utils_ccn import CcnStructure, CompnetUtils, PackedFeatures


class _CCN(nn.Module):
    order = 2

    def _build(self, input_feats, n_outputs, hidden_size, layers, cudaflag):
        self.input_feats = input_feats
        self.n_outputs = n_outputs
        self.layers = layers
        self.utils = CompnetUtils(cudaflag)
        self.w1 = nn.Linear(input_feats * self.num_contractions, hidden_size)
        for i in range(layers - 1):
            self.add_module('w{}'.format(i + 2), nn.Linear(hidden_size * self.num_contractions, hidden_size))
        self.fc = nn.Linear(self.layers * hidden_size + input_feats, self.n_outputs)

    def _levels(self, F0, st):
        upd = ops.Ccn2Update if self.order == 2 else ops.Ccn1Update
        goff = st.goff2 if self.order == 2 else st.goff1
        feats, cur = [ops.SegmentSum.apply(F0, goff, st.n_graphs)], F0
        for i in range(self.layers):
            w = self._modules['w{}'.format(i + 1)]
            cur = upd.apply(cur, w.weight, w.bias, st)
            feats.append(ops.SegmentSum.apply(cur, goff, st.n_graphs))
        return torch.cat(feats, 1)       # (n_graphs, C_in + layers*hidden)

    def forward(self, X, adj):
        """One graph: X (n, input_feats), adj (n, n) with self-loops -> (n_outputs,)."""
        cur = self.utils.get_F0(X, adj) if self.order == 2 else self.utils.get_F0_1D(X, adj)
        return self.fc(self._levels(cur.packed, cur.st)[0])

    def forward_batch(self, Xs, adjs, structure=None):
        """A list of graphs in one pass -> (n_graphs, n_outputs)."""
        st = structure if structure is not None else CcnStructure.from_graphs(adjs)
        X = torch.cat([x.float() for x in Xs], 0).to(st.device)
        rows = st.row_vertex2 if self.order == 2 else st.row_vertex1
        return self.fc(self._levels(X.index_select(0, rows), st))


class CCN_1D(_CCN):
    order = 1

    def __init__(self, input_feats, n_outputs=1, hidden_size=2, layers=2, cudaflag=False):
        super(CCN_1D, self).__init__()
        self.hidden_size = hidden_size
        self.num_contractions = 2
        self._build(input_feats, n_outputs, hidden_size, layers, cudaflag)
        for l in [self._modules['w{}'.format(i + 1)] for i in range(layers)] + [self.fc]:   # :35-39
            l.weight.data.normal_(0, 0.1)
            l.bias.data.normal_(0, 0.1)


class CCN_2D(_CCN):
    order = 2

    def __init__(self, input_feats=2, n_outputs=1, hidden_size=2, layers=2, cudaflag=True):
        super(CCN_2D, self).__init__()
        self.hidden_size = 2          # the reference hard-codes this attribute (:73); unused
        self.num_contractions = 18
        self.cudaflag = cudaflag
        self._build(input_feats, n_outputs, hidden_size, layers, cudaflag)
        for i in range(layers):                                                          # :86-91
            self._modules['w{}'.format(i + 1)].weight.data.normal_(0, 0.1)
        self.fc.weight.data.normal_(0, 0.5)
