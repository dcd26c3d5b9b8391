"""Pure-torch CPU restatement of the S-CGIB pre-training hot path.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker and the timed CPU baseline.

Each function cites the reference lines it follows (paths relative to /root/reference).
DGL calls are replaced by index operations that follow DGL 1.1.0 semantics;
that part is restated, not executed ("parity unpinned" for DGL internals; the reference's own
Python is pinned by tests/golden, see oracle/__init__.py).

Two flavours of the same math:

* ``faithful``   - loop-for-loop: per-graph Python loops with growing ``torch.cat``
                   (models.py:631-660, 738-748) and the dense N x N reconstruction
                   (models.py:762-768).  Ground truth at small B and the timed "reference CPU path".
* ``vectorised`` - segment ops + the Gram identity for the reconstruction loss; proven equal
                   to ``faithful`` in tests/test_oracle.py and used at B = 4096/8192.

Noise enters as explicit tensors (``gate_u [N]``, ``feat_u [N,d]``); when omitted the faithful
path draws it exactly as the reference does on a CPU run (models.py:599, 650).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .graph_ref import RefEgoBatch, RefGraph


# --------------------------------------------------------------------------------------
# graph tensors
# --------------------------------------------------------------------------------------
@dataclass
class TGraph:
    """Torch view of a RefGraph / RefEgoBatch: what DGLGraph supplies to the reference."""
    seg_ptr: torch.Tensor   # int64 [S+1] segment (graph / ego-net) offsets over rows
    indptr: torch.Tensor    # int64 [V+1]
    indices: torch.Tensor   # int64 [E]
    src: torch.Tensor       # int64 [E]
    dst: torch.Tensor       # int64 [E]

    @property
    def num_nodes(self):
        return self.indptr.numel() - 1

    def batch_num_nodes(self):
        return self.seg_ptr[1:] - self.seg_ptr[:-1]

    def seg_ids(self):
        n = self.batch_num_nodes()
        return torch.repeat_interleave(torch.arange(n.numel(), device=n.device), n)

    def to(self, device):
        """The same graph on another device (bench.py times the restatement with torch on the GPU as well)."""
        return TGraph(*[t.to(device) for t in (self.seg_ptr, self.indptr, self.indices, self.src, self.dst)])


def tgraph_from_ref(g: RefGraph) -> TGraph:
    indptr = torch.from_numpy(g.indptr.astype(np.int64))
    indices = torch.from_numpy(g.indices.astype(np.int64))
    dst = torch.repeat_interleave(torch.arange(g.num_nodes), indptr[1:] - indptr[:-1])
    return TGraph(torch.from_numpy(g.graph_ptr.astype(np.int64)), indptr, indices, indices, dst)


def tgraph_from_ego(e: RefEgoBatch) -> TGraph:
    indptr = torch.from_numpy(e.sub_indptr.astype(np.int64))
    indices = torch.from_numpy(e.sub_indices.astype(np.int64))
    dst = torch.repeat_interleave(torch.arange(e.num_rows), indptr[1:] - indptr[:-1])
    return TGraph(torch.from_numpy(e.ego_ptr.astype(np.int64)), indptr, indices, indices, dst)


def sum_nodes(g: TGraph, h: torch.Tensor) -> torch.Tensor:
    """dgl.sum_nodes: segment sum by batch_num_nodes (models.py:716, 725, 733, 684)."""
    out = torch.zeros(g.seg_ptr.numel() - 1, h.shape[1], dtype=h.dtype, device=h.device)
    return out.index_add(0, g.seg_ids(), h)


# --------------------------------------------------------------------------------------
# modules (state_dict keys identical to the reference's)
# --------------------------------------------------------------------------------------
class MLP(nn.Module):
    """models.py:38-49."""

    def __init__(self, num_features, num_classes, dims=16):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(num_features, dims), nn.ReLU(), nn.Linear(dims, num_classes))

    def forward(self, x):
        return self.mlp(x)


class GINConvRef(nn.Module):
    """DGL 1.1.0 GINConv(apply_func, 'sum', init_eps=0, learn_eps=False):
    rst = (1 + eps) * h_dst + sum_{u->v} h_u ; rst = apply_func(rst).  eps is a buffer."""

    def __init__(self, apply_func):
        super().__init__()
        self.apply_func = apply_func
        self.register_buffer("eps", torch.FloatTensor([0.0]))

    def forward(self, g: TGraph, h):
        neigh = torch.zeros_like(h).index_add(0, g.dst, h[g.src])
        rst = (1 + self.eps.to(h.dtype)) * h + neigh
        return self.apply_func(rst)


class GIN(nn.Module):
    """models.py:52-72.  ``num_gin_layers`` = number of GINConv (4 in the published code:
    ``num_layers = 5; range(num_layers - 1)``)."""

    def __init__(self, input_dim, hidden_dim=64, num_gin_layers=4):
        super().__init__()
        self.ginlayers = nn.ModuleList()
        self.batch_norms = nn.ModuleList()
        for layer in range(num_gin_layers):
            mlp = MLP(input_dim if layer == 0 else hidden_dim, hidden_dim, hidden_dim)
            self.ginlayers.append(GINConvRef(mlp))
            self.batch_norms.append(nn.BatchNorm1d(hidden_dim))

    def forward(self, g, h):
        if getattr(self, "emulate_bf16", False):
            return self.forward_bf16_emulated(g, h)
        for i, layer in enumerate(self.ginlayers):
            h = layer(g, h)
            h = self.batch_norms[i](h)
            h = F.relu(h)
        return h

    def forward_bf16_emulated(self, g, h):
        """The same layers with the rounding points of the product's bf16 mode (csrc/gin_bf16.cu), everything else in the
        module's own precision (fp64 in the tests): the layer input t, the aggregated input a, the hidden activation r, the
        pre-BN output y and both weight matrices are rounded to bf16 (round-to-nearest-even); accumulation, biases, the
        BatchNorm statistics (of the ROUNDED y) and BN + ReLU are exact.  Training-mode statistics only.  The casts are
        straight-through for autograd.  Test infrastructure: separates "the kernel computes the bf16 algorithm correctly"
        from "how far the bf16 algorithm is from the reference"."""
        bf = lambda x: x + (x.detach().float().bfloat16().to(x.dtype) - x.detach())
        h = bf(h)
        act = lambda z: z
        for i, layer in enumerate(self.ginlayers):
            hin = act(h)
            a = bf(hin + torch.zeros_like(hin).index_add(0, g.dst, hin[g.src]))
            l1, l2 = layer.apply_func.mlp[0], layer.apply_func.mlp[2]
            r = bf(F.relu(a @ bf(l1.weight).t() + l1.bias))
            y = bf(r @ bf(l2.weight).t() + l2.bias)
            bn = self.batch_norms[i]
            mu, var = y.mean(0), y.var(0, unbiased=False)
            sc = bn.weight / torch.sqrt(var + bn.eps)
            sh = bn.bias - mu * sc
            act = (lambda sc, sh: (lambda z: F.relu(z * sc + sh)))(sc, sh)
            h = y
        return act(h)


class SAGEConvRef(nn.Module):
    """DGL 1.1.x SAGEConv(in, out, 'mean') as models.py:94-96 builds it: rst = fc_self(h) + fc_neigh(mean_{u->v} h_u)
    (mean = 0 for a node without in-edges; fc_self carries the bias, fc_neigh has none).  Written in the
    aggregate-then-project order; DGL projects first when in > out (the same function)."""

    def __init__(self, in_feats, out_feats):
        super().__init__()
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=False)
        self.fc_self = nn.Linear(in_feats, out_feats, bias=True)
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def forward(self, g: TGraph, h):
        deg = (g.indptr[1:] - g.indptr[:-1]).to(h.dtype).clamp(min=1)
        neigh = torch.zeros_like(h).index_add(0, g.dst, h[g.src]) / deg[:, None]
        return self.fc_self(h) + self.fc_neigh(neigh)


class GraphSAGE(nn.Module):
    """models.py:91-104: conv1, ReLU, conv2, ReLU, conv2 AGAIN (the forward never calls conv3: quirk kept), no final ReLU."""

    def __init__(self, in_feats, h_feats):
        super().__init__()
        self.conv1 = SAGEConvRef(in_feats, h_feats)
        self.conv2 = SAGEConvRef(h_feats, h_feats)
        self.conv3 = SAGEConvRef(h_feats, h_feats)

    def forward(self, g, h):
        h = F.relu(self.conv1(g, h))
        h = F.relu(self.conv2(g, h))
        return self.conv2(g, h)


class GraphConvRef(nn.Module):
    """DGL 1.1.x GraphConv(in, out, norm='both', allow_zero_in_degree=True) as models.py:78-80 builds it:
    rst = D_in^-1/2 A^T D_out^-1/2 h W + b, degrees clamped to 1, weight stored [in, out]."""

    def __init__(self, in_feats, out_feats):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        self.bias = nn.Parameter(torch.zeros(out_feats))
        nn.init.xavier_uniform_(self.weight)

    def forward(self, g: TGraph, h):
        # symmetric adjacency (bidirected molecular graphs and their induced ego-nets): in-degree = out-degree
        nrm = torch.pow((g.indptr[1:] - g.indptr[:-1]).to(h.dtype).clamp(min=1), -0.5)
        hs = h * nrm[:, None]
        agg = torch.zeros_like(hs).index_add(0, g.dst, hs[g.src]) * nrm[:, None]
        return agg @ self.weight + self.bias


class GCN(nn.Module):
    """models.py:75-88: widths in -> 2h -> 2h -> h, ReLU between, none after conv3."""

    def __init__(self, num_features, hidden_dim=64):
        super().__init__()
        self.conv1 = GraphConvRef(num_features, hidden_dim * 2)
        self.conv2 = GraphConvRef(hidden_dim * 2, hidden_dim * 2)
        self.conv3 = GraphConvRef(hidden_dim * 2, hidden_dim)

    def forward(self, g, h):
        h = F.relu(self.conv1(g, h))
        h = F.relu(self.conv2(g, h))
        return self.conv3(g, h)


def make_encoder(encoder, d_transfer, hidden_dim, num_gin_layers=4):
    """models.py:573-587."""
    if encoder == "GIN":
        return GIN(d_transfer, hidden_dim, num_gin_layers)
    if encoder == "GCN":
        return GCN(d_transfer, hidden_dim)
    if encoder == "GraphSAGE":
        return GraphSAGE(d_transfer, hidden_dim)
    raise SystemExit("Bug there is no pre-defined Encoders")


class _LSTMHolder(nn.Module):
    """Stand-in for dgl.nn.Set2Set(hidden, 2, 1): only the parameter container (``s2s.lstm.*``)
    so that state_dict keys / parameter order / init RNG consumption match models.py:565."""

    def __init__(self, input_dim, n_iters, n_layers):
        super().__init__()
        self.lstm = nn.LSTM(2 * input_dim, input_dim, n_layers)
        self.lstm.reset_parameters()     # DGL's Set2Set.__init__ re-initialises the LSTM (a second RNG draw)


class OracleMainmodel(nn.Module):
    """models.py:546-782 (Mainmodel), GIN encoder, readout 'sum', recons_type 'adj', useAtt 1.
    Construction order follows models.py:547-593 so a seeded default init matches."""

    def __init__(self, in_dim, hidden_dim=64, d_transfer=32, num_gin_layers=4, encoder="GIN"):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.fc1 = nn.Linear(hidden_dim, 1)
        self.in_dim = d_transfer
        self.transfer_d = nn.Linear(in_dim, d_transfer, bias=False)
        self.embedding_h = nn.Linear(d_transfer, hidden_dim, bias=False)
        self.attn_layer = nn.Linear(hidden_dim * 2, 1)
        self.reduce_d = nn.Linear(2 * hidden_dim, hidden_dim)
        self.s2s = _LSTMHolder(hidden_dim, 2, 1)
        self.reconstructX = nn.Sequential(nn.Linear(hidden_dim, d_transfer))
        self.MLP = nn.Sequential(nn.Linear(2 * hidden_dim, hidden_dim), nn.ReLU(),
                                 nn.Linear(hidden_dim, hidden_dim))
        self.Encoder1 = make_encoder(encoder, d_transfer, hidden_dim, num_gin_layers)     # models.py:573-587
        self.Encoder2 = make_encoder(encoder, d_transfer, hidden_dim, num_gin_layers)
        self.compressor = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.BatchNorm1d(hidden_dim),
                                        nn.ReLU(), nn.Linear(hidden_dim, 1))

    # ---------------------------------------------------------------- faithful pieces
    def compress(self, graph_features, gate_u=None):
        """models.py:595-604.  ``gate_u`` [n,1] replaces torch.rand(p.size())."""
        p = self.compressor(graph_features)
        bias = 0.0 + 0.0001
        u = torch.rand(p.size()) if gate_u is None else gate_u.reshape(p.size()).to(p.dtype)     # CPU draw, as models.py:599
        eps = (bias - (1 - bias)) * u + (1 - bias)
        gate_inputs = (torch.log(eps) - torch.log(1 - eps)).to(p.dtype).to(p.device)             # models.py:601
        gate_inputs = (gate_inputs + p) / 1.0
        gate_inputs = torch.sigmoid(gate_inputs).squeeze()
        return gate_inputs, p

    def compression(self, graph_features, nodes_list, gate_u=None, feat_u=None):
        """models.py:631-660, loop-for-loop (incl. the KL overwrite of models.py:659)."""
        epsilon = 0.0000001
        noisy_all = torch.tensor((), dtype=graph_features.dtype, device=graph_features.device)
        p_all = torch.tensor((), dtype=graph_features.dtype, device=graph_features.device)
        KL_all = torch.tensor((), dtype=graph_features.dtype, device=graph_features.device)
        split = torch.split(graph_features, tuple(nodes_list))
        off = 0
        for i in range(len(nodes_list)):
            features = split[i]
            n = features.shape[0]
            gu = None if gate_u is None else gate_u[off:off + n]
            lambda_pos, p = self.compress(features, gu)
            lambda_pos = lambda_pos.reshape(-1, 1)
            lambda_neg = 1 - lambda_pos
            static = features.clone().detach()
            std, mean = torch.std_mean(static, dim=0)
            noisy_mean = lambda_pos * features + lambda_neg * mean
            noisy_std = lambda_neg * std
            fu = torch.rand_like(noisy_mean) if feat_u is None else feat_u[off:off + n].to(noisy_mean.dtype).to(noisy_mean.device)
            noisy = noisy_mean + fu * noisy_std
            noisy_all = torch.cat((noisy_all, noisy), 0)
            p_all = torch.cat((p_all, p), 0)
            KL = 0.5 * ((noisy_std ** 2) / (std + epsilon) ** 2) + torch.sum(
                ((noisy_mean - mean) / (std + epsilon)) ** 2, dim=0)
            KL_all = torch.cat((KL, KL), 0)
            off += n
        return noisy_all, p_all, KL_all

    def sim(self, z1, z2):
        """models.py:606-609."""
        return torch.mm(F.normalize(z1), F.normalize(z2).t())

    def batched_semi_loss(self, z1, z2, batch_size):
        """models.py:611-629."""
        num_nodes = z1.size(0)
        num_batches = (num_nodes - 1) // batch_size + 1
        f = lambda x: torch.exp(x / 1)
        indices = torch.arange(0, num_nodes, device=z1.device)
        losses = []
        for i in range(num_batches):
            mask = indices[i * batch_size:(i + 1) * batch_size]
            refl = f(self.sim(z1[mask], z1))
            btw = f(self.sim(z1[mask], z2))
            losses.append(-torch.log(btw[:, i * batch_size:(i + 1) * batch_size].diag()
                                     / (refl.sum(1) + btw.sum(1)
                                        - refl[:, i * batch_size:(i + 1) * batch_size].diag())))
        return torch.cat(losses).mean()

    def loss_recon_adj(self, interaction_map, g: TGraph):
        """models.py:762-768: dense N x N over the whole batched graph."""
        row_num = interaction_map.shape[0]
        adj = torch.zeros(row_num, row_num, dtype=interaction_map.dtype, device=interaction_map.device)
        adj[g.src, g.dst] = 1.0
        recon = torch.mm(interaction_map, interaction_map.t())
        return torch.sum((recon - adj) ** 2) / row_num

    def extract_features(self, g: TGraph, batch_x, eg: TGraph, x_subs, gate_u=None, feat_u=None):
        """models.py:702-750 (readout 'sum', useAtt)."""
        nodes_list = [int(v) for v in g.batch_num_nodes()]
        graph_features = self.Encoder1(g, batch_x)
        subgraphs_features = self.Encoder2(eg, x_subs)
        graph_features_readout = sum_nodes(g, graph_features)
        noisy, p, KL_tensor = self.compression(graph_features, nodes_list, gate_u, feat_u)
        sub_readout = sum_nodes(eg, subgraphs_features)
        noisy_readout = sum_nodes(g, noisy)
        subgs_att = torch.tensor((), dtype=noisy.dtype, device=noisy.device)
        split = torch.split(sub_readout, tuple(nodes_list))
        for i in range(len(split)):
            cp = noisy_readout[i].repeat(nodes_list[i], 1)
            interaction = torch.cat((cp, split[i]), -1)
            att = F.softmax(self.attn_layer(interaction), dim=0)
            subgs_att = torch.cat((subgs_att, split[i] * att), 0)
        interaction_map = torch.cat((noisy, subgs_att), -1)
        self._dbg = dict(graph_features=graph_features, sub_readout=sub_readout, p=p)
        return interaction_map, KL_tensor, noisy, graph_features_readout

    def forward_faithful(self, g: TGraph, x_norm, eg: TGraph, x_subs_norm, gate_u=None, feat_u=None,
                         batch_size=None, recon_logm_steps=0):
        """models.py:662-700 / 1158-1195.  x_norm / x_subs_norm are the F.normalize'd raw features
        (exp_pretraining.py:312-314).  Returns dict with the three losses and the embeddings."""
        batch_x = self.transfer_d(x_norm)
        x_subs = self.transfer_d(x_subs_norm)
        imap, KL_tensor, noisy, g_readout = self.extract_features(g, batch_x, eg, x_subs, gate_u, feat_u)
        Z = self.MLP(imap)
        KL_loss = torch.mean(KL_tensor)
        noisy2 = sum_nodes(g, noisy)
        bs = batch_size if batch_size is not None else noisy2.shape[0]
        con = self.batched_semi_loss(noisy2, g_readout, bs)
        rec = loss_recon_logm(Z, g, recon_logm_steps) if recon_logm_steps else self.loss_recon_adj(Z, g)
        return dict(KL=KL_loss, contrastive=con, recon=rec, interaction_map=imap, Z=Z, noisy=noisy,
                    graph_readout=g_readout, core_readout=noisy2)

    # ---------------------------------------------------------------- vectorised (same math)
    def forward_vectorised(self, g: TGraph, x_norm, eg: TGraph, ego_nodes, gate_u, feat_u):
        """Segment-op restatement (no per-graph loops, Gram identity for the recon loss;
        SURVEY.md F5/F14 + Appendix A).  ``ego_nodes`` [Ns] int64 maps ego rows to parent nodes."""
        dt = x_norm.dtype
        t = self.transfer_d(x_norm)
        H = self.Encoder1(g, t)
        S = self.Encoder2(eg, t[ego_nodes])
        seg = g.seg_ids()
        nB = g.seg_ptr.numel() - 1
        n = g.batch_num_nodes().to(dt)
        R = sum_nodes(g, H)
        # compressor with per-graph BatchNorm (models.py:589-596 applied per split, :642)
        lin1, bn, _, lin2 = self.compressor
        q = lin1(H)
        dv = x_norm.device
        qm = torch.zeros(nB, q.shape[1], dtype=dt, device=dv).index_add(0, seg, q) / n[:, None]
        qc = q - qm[seg]
        qv = torch.zeros(nB, q.shape[1], dtype=dt, device=dv).index_add(0, seg, qc * qc) / n[:, None]
        qh = qc / torch.sqrt(qv[seg] + bn.eps)
        if self.training:
            # B sequential EMA updates (one per graph, graph order) in closed form: r <- 0.9 r + 0.1 stat_i
            with torch.no_grad():
                w = 0.1 * 0.9 ** torch.arange(nB - 1, -1, -1, dtype=dt, device=dv)
                bn.running_mean.mul_(0.9 ** nB).add_((w[:, None] * qm).sum(0))
                bn.running_var.mul_(0.9 ** nB).add_((w[:, None] * qv * (n / (n - 1))[:, None]).sum(0))
                bn.num_batches_tracked += nB
        p = lin2(F.relu(qh * bn.weight + bn.bias))
        eps = (0.0001 - (1 - 0.0001)) * gate_u.to(dt) + (1 - 0.0001)
        lam = torch.sigmoid((torch.log(eps) - torch.log(1 - eps)).to(dt)[:, None] + p)
        Hd = H.detach()
        mu = torch.zeros(nB, H.shape[1], dtype=dt, device=dv).index_add(0, seg, Hd) / n[:, None]
        var = torch.zeros(nB, H.shape[1], dtype=dt, device=dv).index_add(0, seg, (Hd - mu[seg]) ** 2) / (n[:, None] - 1)
        sd = torch.sqrt(var)
        m = lam * H + (1 - lam) * mu[seg]
        s = (1 - lam) * sd[seg]
        noisy = m + feat_u.to(dt) * s
        # KL: last graph only (models.py:657-659)
        last = seg == (nB - 1)
        e = 0.0000001
        KLt = 0.5 * (s[last] ** 2) / (sd[-1] + e) ** 2 + torch.sum(((m[last] - mu[-1]) / (sd[-1] + e)) ** 2, dim=0)
        KL = KLt.mean()
        # ego pooling, attention (softmax shift invariance: core half and bias cancel, F14)
        C = sum_nodes(eg, S)
        core = sum_nodes(g, noisy)
        logit = self.attn_layer(torch.cat((core[seg], C), -1)).squeeze(-1)
        mx = torch.full((nB,), -float("inf"), dtype=dt, device=dv).scatter_reduce(0, seg, logit, "amax")
        ex = torch.exp(logit - mx[seg])
        den = torch.zeros(nB, dtype=dt, device=dv).index_add(0, seg, ex)
        alpha = ex / den[seg]
        imap = torch.cat((noisy, C * alpha[:, None]), -1)
        Z = self.MLP(imap)
        # contrastive (models.py:606-629), one chunk
        z1 = F.normalize(core)
        z2 = F.normalize(R)
        refl = torch.exp(z1 @ z1.t())
        btw = torch.exp(z1 @ z2.t())
        con = (-torch.log(btw.diag() / (refl.sum(1) + btw.sum(1) - refl.diag()))).mean()
        # recon via Gram identity: ||ZZ^T - A||_F^2 = ||Z^T Z||_F^2 - 2 sum_E z_i.z_j + nnz(A)
        G = Z.t() @ Z
        edge = (Z[g.src] * Z[g.dst]).sum()
        rec = ((G * G).sum() - 2 * edge + g.src.numel()) / Z.shape[0]
        return dict(KL=KL, contrastive=con, recon=rec, interaction_map=imap, Z=Z, noisy=noisy,
                    graph_readout=R, core_readout=core, H=H, C=C, alpha=alpha, lam=lam.squeeze(-1), t=t)


def normalize_rows(x: torch.Tensor) -> torch.Tensor:
    """F.normalize(x) as used at exp_pretraining.py:312-314 (p=2, dim=1, eps=1e-12)."""
    return F.normalize(x)


def draw_noise_like_reference(nodes_list, d, seed):
    """Reproduce the CPU RNG stream of one reference forward on CPU: per graph, n gate draws
    (torch.rand(p.size()), models.py:599) then n*d feature draws (rand_like, models.py:650)."""
    gen_state = torch.get_rng_state()
    torch.manual_seed(seed)
    gate, feat = [], []
    for n in nodes_list:
        gate.append(torch.rand(n, 1))
        feat.append(torch.rand(n, d))
    torch.set_rng_state(gen_state)
    return torch.cat(gate).squeeze(1), torch.cat(feat)


def oracle_train_step(model: OracleMainmodel, opt, g, x_norm, eg, x_subs_norm, gate_u=None, feat_u=None):
    """exp_pretraining.py:307-324: zero_grad, forward, loss = KL + recon + contrastive, backward, Adam."""
    opt.zero_grad()
    out = model.forward_faithful(g, x_norm, eg, x_subs_norm, gate_u, feat_u)
    loss = out["KL"] + out["recon"] + out["contrastive"]
    loss.backward()
    opt.step()
    return float(loss.detach())


# --------------------------------------------------------------------------------------
# fine-tuning head: Set2Set readout + predict MLP
# --------------------------------------------------------------------------------------
class Set2SetRef(nn.Module):
    """dgl.nn.Set2Set(input_dim, n_iters, n_layers) restated from DGL 1.1.0 (dgl/nn/pytorch/glob.py; call sites
    models.py:365, 515): zero-initialised LSTM state and q*; per iteration ``q = LSTM(q*)``,
    ``e_v = <feat_v, q_graph(v)>``, ``alpha = softmax of e over the nodes of each graph``,
    ``readout_g = sum_v alpha_v feat_v``, ``q* = [q || readout]``.  DGL internals: restated, unpinned."""

    def __init__(self, input_dim, n_iters, n_layers):
        super().__init__()
        self.input_dim, self.output_dim, self.n_iters, self.n_layers = input_dim, 2 * input_dim, n_iters, n_layers
        self.lstm = nn.LSTM(self.output_dim, self.input_dim, n_layers)
        self.lstm.reset_parameters()

    def forward(self, g: TGraph, feat):
        nB = g.seg_ptr.numel() - 1
        seg = g.seg_ids()
        h = (feat.new_zeros((self.n_layers, nB, self.input_dim)), feat.new_zeros((self.n_layers, nB, self.input_dim)))
        q_star = feat.new_zeros(nB, self.output_dim)
        for _ in range(self.n_iters):
            q, h = self.lstm(q_star.unsqueeze(0), h)
            q = q.view(nB, self.input_dim)
            e = (feat * q[seg]).sum(dim=-1)
            mx = torch.full((nB,), -float("inf"), dtype=feat.dtype).scatter_reduce(0, seg, e, "amax")
            ex = torch.exp(e - mx[seg])
            alpha = ex / torch.zeros(nB, dtype=feat.dtype).index_add(0, seg, ex)[seg]
            readout = torch.zeros(nB, self.input_dim, dtype=feat.dtype).index_add(0, seg, feat * alpha[:, None])
            q_star = torch.cat([q, readout], dim=-1)
        return q_star


def finetune_trainable(name: str) -> bool:
    """The freeze rule of Mainmodel_finetuning.__init__ for the LOADED model's parameters (models.py:424-435): the inner
    loop's last iteration wins, so exactly the names containing "layers.2" stay trainable (with num_layers = 4 the list
    is ["layers.4", "layers.3", "layers.2"])."""
    return "layers.2" in name


class OracleFinetune(nn.Module):
    """models.py:358-520 (Mainmodel_finetuning), GIN encoder: own transfer_d / MLP / s2s / predict around the loaded
    pre-trained model's extract_features.  Modules the forward never touches (embedding_h, reduce_d, attn_layer,
    Encoder1/2, compressor of the OUTER module) are kept so that state-dict keys match."""

    def __init__(self, inner: OracleMainmodel, in_dim, hidden_dim=64, d_transfer=32, num_classes=10,
                 task="graph_classification", regression_dataset=False, num_gin_layers=4):
        super().__init__()
        self.s2s = Set2SetRef(hidden_dim, 2, 1)
        self.transfer_d = nn.Linear(in_dim, d_transfer, bias=False)
        self.embedding_h = nn.Linear(d_transfer, hidden_dim, bias=False)
        self.reduce_d = nn.Linear(2 * hidden_dim, hidden_dim)
        self.attn_layer = nn.Linear(2 * hidden_dim, 1)
        out_dim = 1 if task == "graph_regression" else num_classes
        self.predict = nn.Sequential(nn.Linear(2 * hidden_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, out_dim))
        self.MLP = nn.Sequential(nn.Linear(2 * hidden_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim))
        self.Encoder1 = GIN(d_transfer, hidden_dim, num_gin_layers)
        self.Encoder2 = GIN(d_transfer, hidden_dim, num_gin_layers)
        self.model = inner
        for n, p in self.model.named_parameters():
            p.requires_grad = finetune_trainable(n)
        self.compressor = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.BatchNorm1d(hidden_dim), nn.ReLU(),
                                        nn.Linear(hidden_dim, 1))
        self.no_sigmoid = regression_dataset      # models.py:516-517 (dataset in self.tasks)

    def forward(self, g: TGraph, x_norm, eg: TGraph, x_subs_norm, gate_u=None, feat_u=None):
        """models.py:501-520 -> dict(scores, Z, interaction_map, readout)."""
        batch_x = self.transfer_d(x_norm)
        x_subs = self.transfer_d(x_subs_norm)
        imap, _, _, _ = self.model.extract_features(g, batch_x, eg, x_subs, gate_u, feat_u)
        Z = self.MLP(imap)
        q_star = self.s2s(g, Z)
        s = self.predict(q_star)
        if not self.no_sigmoid:
            s = torch.sigmoid(s)
        return dict(scores=s, Z=Z, interaction_map=imap, readout=q_star)


class OracleDomainAdapt(nn.Module):
    """models.py:107-355 (Mainmodel_domainadapt), GIN encoder: the feature path of the loaded model, Set2Set readout of
    Z, ``r_transfer_d`` MLP to 2*in_dim, and a Set2Set readout of the RAW (normalised) features as the target;
    ``X_loss = sum((r_transfer_d(s2s(Z)) - s2s_rev(x))**2)`` (models.py:254-275).  Every parameter of the loaded model
    is trainable (models.py:176-178)."""

    def __init__(self, inner: OracleMainmodel, in_dim, hidden_dim=64, d_transfer=32, num_classes=10, num_gin_layers=4):
        super().__init__()
        self.s2s = Set2SetRef(hidden_dim, 2, 1)
        self.s2s_rev = Set2SetRef(in_dim, 2, 1)
        self.transfer_d = nn.Linear(in_dim, d_transfer, bias=False)
        self.embedding_h = nn.Linear(d_transfer, hidden_dim, bias=False)
        self.reduce_d = nn.Linear(2 * hidden_dim, hidden_dim)
        self.attn_layer = nn.Linear(2 * hidden_dim, 1)
        self.r_transfer_d = nn.Sequential(nn.Linear(2 * hidden_dim, hidden_dim), nn.ReLU(),
                                          nn.Linear(hidden_dim, in_dim * 2))
        self.predict = nn.Sequential(nn.Linear(2 * hidden_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, num_classes))
        self.MLP = nn.Sequential(nn.Linear(2 * hidden_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim))
        self.Encoder1 = GIN(d_transfer, hidden_dim, num_gin_layers)
        self.Encoder2 = GIN(d_transfer, hidden_dim, num_gin_layers)
        self.model = inner
        for p in self.model.parameters():
            p.requires_grad = True
        self.compressor = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.BatchNorm1d(hidden_dim), nn.ReLU(),
                                        nn.Linear(hidden_dim, 1))
        self.reconstructX = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, in_dim))

    def forward(self, g: TGraph, x_norm, eg: TGraph, x_subs_norm, gate_u=None, feat_u=None):
        batch_x = self.transfer_d(x_norm)
        x_subs = self.transfer_d(x_subs_norm)
        imap, _, _, _ = self.model.extract_features(g, batch_x, eg, x_subs, gate_u, feat_u)
        Z = self.MLP(imap)
        rec = self.r_transfer_d(self.s2s(g, Z))
        org = self.s2s_rev(g, x_norm)
        return dict(X_loss=torch.sum((rec - org) ** 2), rec=rec, org=org, Z=Z)


# --------------------------------------------------------------------------------------
# --recons_type logM
# --------------------------------------------------------------------------------------
def get_prob_tran_mat(Ak: np.ndarray) -> np.ndarray:
    """util.py:60-71 (GetProbTranMat): log(Ak / colsum) - log(1/n), negatives / -inf / nan -> 0."""
    n = Ak.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        col = np.repeat(np.sum(Ak, axis=0).reshape(1, -1), n, axis=0)
        m = np.log(np.divide(Ak, col)) - np.log(1.0 / n)
    m[m < 0] = 0
    m[np.isnan(m)] = 0
    return m


def get_logm(dense_adj: np.ndarray, kstep: int):
    """util.py:74-91 (getM_logM): Ak = Adj^i for i = 1..kstep (fp64), logM_i = GetProbTranMat(Ak)."""
    n = dense_adj.shape[0]
    Ak = np.identity(n)
    out = []
    for _ in range(kstep):
        Ak = Ak @ dense_adj.astype(np.float64)
        out.append(get_prob_tran_mat(Ak.copy()))
    return out


def loss_recon_logm(Z: torch.Tensor, g: TGraph, k: int) -> torch.Tensor:
    """models.py:770-782 (loss_recon): per graph h = Z_g Z_g^T, sum_i sum((h - logM_i)^2) / n^2, all divided by k.
    The logM matrices are what exp_pretraining.py:260-264 stores offline (float32 tensors)."""
    ptr = g.seg_ptr.tolist()
    loss = 0
    for b in range(len(ptr) - 1):
        v0, v1 = ptr[b], ptr[b + 1]
        n = v1 - v0
        A = np.zeros((n, n))
        sel = (g.dst >= v0) & (g.dst < v1)
        A[(g.src[sel] - v0).numpy(), (g.dst[sel] - v0).numpy()] = 1.0
        logm = [torch.from_numpy(np.asarray(m)).float().to(Z.dtype) for m in get_logm(A, k)]
        h = Z[v0:v1] @ Z[v0:v1].t()
        for i in range(k):
            loss = loss + torch.sum((h - logm[i]) ** 2) / (n * n)
    return loss / k
