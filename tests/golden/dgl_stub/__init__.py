"""Minimal stand-in for the DGL 1.1.0 call surface used by the reference's models.py hot path.

Test tooling only: lets tests/golden/make_golden.py import the UNMODIFIED /root/reference/models.py
in a container without DGL/PyG/ogb/pyro/torch_scatter and run Mainmodel.forward on CPU to
record golden vectors.  The DGL semantics below are restated from the library's v1.1.0
behaviour; everything else executed for the golden vectors is the reference's
own code.
"""
import sys
import types

import torch
import torch.nn as nn


class StubGraph:
    """What the reference needs from a (batched) DGLGraph."""

    def __init__(self, seg_ptr, src, dst, num_nodes):
        self._seg_ptr = torch.as_tensor(seg_ptr, dtype=torch.int64)
        self._src = torch.as_tensor(src, dtype=torch.int64)
        self._dst = torch.as_tensor(dst, dtype=torch.int64)
        self._n = int(num_nodes)
        self.ndata = {}

    def batch_num_nodes(self):
        return self._seg_ptr[1:] - self._seg_ptr[:-1]

    def num_nodes(self):
        return self._n

    def edges(self):
        return self._src, self._dst

    def to(self, device):
        return self

    def adj(self):
        g = self

        class _Adj:
            def to_dense(self_inner):
                a = torch.zeros(g._n, g._n)
                a[g._src, g._dst] = 1.0
                return a
        return _Adj()


def sum_nodes(g, key):
    h = g.ndata[key]
    n = g.batch_num_nodes()
    seg = torch.repeat_interleave(torch.arange(n.numel()), n)
    return torch.zeros(n.numel(), h.shape[1], dtype=h.dtype).index_add(0, seg, h)


class GINConv(nn.Module):
    """dgl.nn.pytorch.conv.GINConv(apply_func, aggregator_type='sum', init_eps=0, learn_eps=False)."""

    def __init__(self, apply_func=None, aggregator_type="sum", init_eps=0, learn_eps=False, activation=None):
        super().__init__()
        assert aggregator_type == "sum" and not learn_eps and activation is None
        self.apply_func = apply_func
        self.register_buffer("eps", torch.FloatTensor([init_eps]))

    def forward(self, g, feat):
        neigh = torch.zeros_like(feat).index_add(0, g._dst, feat[g._src])
        rst = (1 + self.eps) * feat + neigh
        if self.apply_func is not None:
            rst = self.apply_func(rst)
        return rst


class SAGEConv(nn.Module):
    """dgl.nn.SAGEConv(in_feats, out_feats, 'mean'), restated from DGL 1.1.x: fc_neigh (no bias) on the mean of the
    in-neighbours (zero for isolated nodes), fc_self (with the bias), Xavier-uniform weights with the ReLU gain.  The linear
    map is applied before the aggregation when in_feats > out_feats (same function, the library's operation order)."""

    def __init__(self, in_feats, out_feats, aggregator_type, feat_drop=0.0, bias=True, norm=None, activation=None):
        super().__init__()
        assert aggregator_type == "mean" and feat_drop == 0.0 and norm is None and activation is None
        self._in, self._out = in_feats, out_feats
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=False)
        self.fc_self = nn.Linear(in_feats, out_feats, bias=bias)
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def forward(self, g, feat):
        deg = torch.zeros(g._n, dtype=feat.dtype).index_add(0, g._dst, torch.ones(g._dst.numel(), dtype=feat.dtype))
        lin_before_mp = self._in > self._out
        src = self.fc_neigh(feat) if lin_before_mp else feat
        neigh = torch.zeros(g._n, src.shape[1], dtype=feat.dtype).index_add(0, g._dst, src[g._src])
        neigh = neigh / deg.clamp(min=1)[:, None]
        if not lin_before_mp:
            neigh = self.fc_neigh(neigh)
        return self.fc_self(feat) + neigh


class GraphConv(nn.Module):
    """dgl.nn.pytorch.conv.GraphConv(in_feats, out_feats, norm='both', weight=True, bias=True,
    allow_zero_in_degree=True), restated from DGL 1.1.x: D_out^-1/2 on the source rows, sum over the in-edges,
    D_in^-1/2 on the result (degrees clamped to 1), weight [in, out] applied first when in_feats > out_feats."""

    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True, activation=None,
                 allow_zero_in_degree=False):
        super().__init__()
        assert norm == "both" and weight and bias and activation is None
        self._in, self._out = in_feats, out_feats
        self.weight = nn.Parameter(torch.Tensor(in_feats, out_feats))
        self.bias = nn.Parameter(torch.Tensor(out_feats))
        nn.init.xavier_uniform_(self.weight)
        nn.init.zeros_(self.bias)

    def forward(self, g, feat):
        one = torch.ones(g._src.numel(), dtype=feat.dtype)
        dout = torch.zeros(g._n, dtype=feat.dtype).index_add(0, g._src, one).clamp(min=1)
        din = torch.zeros(g._n, dtype=feat.dtype).index_add(0, g._dst, one).clamp(min=1)
        h = feat * torch.pow(dout, -0.5)[:, None]
        if self._in > self._out:
            h = torch.matmul(h, self.weight)
            rst = torch.zeros(g._n, h.shape[1], dtype=feat.dtype).index_add(0, g._dst, h[g._src])
        else:
            rst = torch.zeros(g._n, h.shape[1], dtype=feat.dtype).index_add(0, g._dst, h[g._src])
            rst = torch.matmul(rst, self.weight)
        return rst * torch.pow(din, -0.5)[:, None] + self.bias


class Set2Set(nn.Module):
    """dgl.nn.pytorch.glob.Set2Set(input_dim, n_iters, n_layers), restated from DGL 1.1.0: broadcast_nodes /
    softmax_nodes / sum_nodes are written out with segment ids."""

    def __init__(self, input_dim, n_iters, n_layers):
        super().__init__()
        self.input_dim, self.output_dim, self.n_iters, self.n_layers = input_dim, 2 * input_dim, n_iters, n_layers
        self.lstm = nn.LSTM(self.output_dim, self.input_dim, n_layers)
        self.lstm.reset_parameters()

    def forward(self, graph, feat):
        n = graph.batch_num_nodes()
        batch_size = n.numel()
        seg = torch.repeat_interleave(torch.arange(batch_size), n)
        h = (feat.new_zeros((self.n_layers, batch_size, self.input_dim)),
             feat.new_zeros((self.n_layers, batch_size, self.input_dim)))
        q_star = feat.new_zeros(batch_size, self.output_dim)
        for _ in range(self.n_iters):
            q, h = self.lstm(q_star.unsqueeze(0), h)
            q = q.view(batch_size, self.input_dim)
            e = (feat * q[seg]).sum(dim=-1, keepdim=True)                      # feat * broadcast_nodes(graph, q)
            alpha = torch.cat([torch.softmax(t, dim=0) for t in torch.split(e, n.tolist())])   # softmax_nodes
            r = feat * alpha
            readout = torch.zeros(batch_size, self.input_dim, dtype=feat.dtype).index_add(0, seg, r)   # sum_nodes
            q_star = torch.cat([q, readout], dim=-1)
        return q_star


class _Dummy:
    def __init__(self, *a, **k):
        raise NotImplementedError


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Register stand-ins for every third-party import at the top of reference models.py:4-35."""
    dummy_names = ["GCNConv", "global_mean_pool", "AsGraphPredDataset", "GraphDataLoader",
                   "collate_dgl", "DglGraphPropPredDataset", "Evaluator", "AtomEncoder", "GraphConv",
                   "scatter_mean", "scatter_add", "scatter_std", "SumPooling", "GINDataset"]
    d = {n: _Dummy for n in dummy_names}
    dgl = _mod("dgl", sum_nodes=sum_nodes, StubGraph=StubGraph)
    dgl.nn = _mod("dgl.nn", Set2Set=Set2Set, GraphConv=GraphConv, SAGEConv=SAGEConv, GINConv=GINConv)
    dgl.sparse = _mod("dgl.sparse")
    dgl.function = _mod("dgl.function")
    dgl.data = _mod("dgl.data", AsGraphPredDataset=_Dummy, GINDataset=_Dummy)
    dgl.dataloading = _mod("dgl.dataloading", GraphDataLoader=_Dummy)
    pt = _mod("dgl.nn.pytorch")
    pt.glob = _mod("dgl.nn.pytorch.glob", SumPooling=_Dummy)
    pt.conv = _mod("dgl.nn.pytorch.conv", GINConv=GINConv, GraphConv=GraphConv)
    dgl.nn.pytorch = pt
    tg = _mod("torch_geometric")
    tg.nn = _mod("torch_geometric.nn", GCNConv=_Dummy, SAGEConv=_Dummy, global_mean_pool=_Dummy)
    ogb = _mod("ogb")
    ogb.graphproppred = _mod("ogb.graphproppred", collate_dgl=_Dummy, DglGraphPropPredDataset=_Dummy,
                             Evaluator=_Dummy)
    _mod("ogb.graphproppred.mol_encoder", AtomEncoder=_Dummy)
    _mod("pyro")
    _mod("torch_scatter", scatter_mean=_Dummy, scatter_add=_Dummy, scatter_std=_Dummy)
    return dgl
