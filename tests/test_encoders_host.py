"""Host logic of the --encoder GraphSAGE / GCN step (s-cgib_b200/encoders.py: layer sequencing, the hand-derived backward,
weight sharing of conv2, parameter routing) on CPU: every operator kernel is replaced by a torch stand-in (test infrastructure,
monkeypatched into scgib_b200.ops for this test only), and the composed forward / backward must reproduce the golden vectors
of the UNMODIFIED reference (tests/golden/enc_*.pt).  The kernels themselves are checked on the GPU (tests/test_gpu_encoders.py)."""
import glob
import os
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.graph_ref import RefEgoBatch, RefGraph
from oracle.scgib_oracle import OracleMainmodel, draw_noise_like_reference

ENC_GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "enc_*.pt")))

def seg_ids(ptr):
    n = (ptr[1:] - ptr[:-1]).long(); return torch.repeat_interleave(torch.arange(n.numel()), n)
def input_proj(x, Wt): return F.normalize(x) @ Wt.t()
def fn(mode, deg): 
    d = deg.float().clamp(min=1)
    return torch.ones_like(d) if mode == 0 else (1 / d if mode == 1 else d.pow(-0.5))
def graph_aggregate(h, indptr, indices, sn, dn, row_map=None, add=None):
    ip, ix = indptr.long(), indices.long(); V = ip.numel() - 1
    deg = ip[1:] - ip[:-1]; dst = torch.repeat_interleave(torch.arange(V), deg)
    src = h[row_map.long()] if row_map is not None else h
    out = torch.zeros(V, h.shape[1]).index_add(0, dst, src[ix] * fn(sn, deg)[ix][:, None]) * fn(dn, deg)[:, None]
    return out + add if add is not None else out
def segment_sum_w(h, ptr): return torch.zeros(ptr.numel() - 1, h.shape[1]).index_add(0, seg_ids(ptr), h)
def linear_fwd(X0, W0, O, w0_kxo=False, X1=None, W1=None, w1_kxo=False, bias=None, relu=False, map0=None, M0=None, M1=None, V=None):
    X0 = X0 * (M0 > 0) if M0 is not None else X0
    if map0 is not None: X0 = X0[map0.long()]
    y = X0 @ (W0 if w0_kxo else W0.t())
    assert y.shape[1] == O, (y.shape, O)
    if X1 is not None:
        X1 = X1 * (M1 > 0) if M1 is not None else X1
        y = y + X1 @ (W1 if w1_kxo else W1.t())
    if bias is not None: y = y + bias
    return torch.relu(y) if relu else y
def linear_bwd_w(G, X, dW, db=None, M=None, map=None, kxo=False, accumulate=False):
    Gm = G * (M > 0) if M is not None else G
    Xr = X[map.long()] if map is not None else X
    w = Gm.t() @ Xr
    if kxo: w = w.t()
    assert w.shape == dW.shape, (w.shape, dW.shape)
    with torch.no_grad():
        if accumulate: dW += w
        else: dW.copy_(w)
        if db is not None:
            if accumulate: db += Gm.sum(0)
            else: db.copy_(Gm.sum(0))
    return dW, db
def transfer_bwd(x, g0, g1, map1, normalize=True, csr0=None, csr1=None):
    xn = F.normalize(x) if normalize else x
    return g0.t() @ xn + g1.t() @ xn[map1.long()]
def segment_sum_bwd(g_out, ptr, rows): return g_out[seg_ids(ptr)]

_or = OracleMainmodel(9)

class CoreGate:
    def __init__(self, H, W, b, g, be, w2, b2):
        self.m = OracleMainmodel(9, H); c = self.m.compressor
        with torch.no_grad():
            c[0].weight.copy_(W); c[0].bias.copy_(b); c[1].weight.copy_(g); c[1].bias.copy_(be); c[3].weight.copy_(w2.reshape(1, -1)); c[3].bias.copy_(b2.reshape(1))
        self.m.train()
    def forward(self, Hf, graph_ptr, gate_u, feat_u):
        self.Hf = Hf.detach().clone().requires_grad_()
        nodes = (graph_ptr[1:] - graph_ptr[:-1]).tolist()
        noisy, _, KLt = self.m.compression(self.Hf, nodes, gate_u, feat_u)
        seg = seg_ids(graph_ptr); B = graph_ptr.numel() - 1
        self.readout = torch.zeros(B, Hf.shape[1]).index_add(0, seg, self.Hf)
        self.core = torch.zeros(B, Hf.shape[1]).index_add(0, seg, noisy)
        self.noisy, self.kl = noisy, KLt.mean()
        return noisy.detach(), None, self.readout.detach(), self.core.detach(), self.kl.detach().reshape(1)
    def update_running(self, rm, rv):          # the stand-in's own BatchNorm started from the same default buffers
        with torch.no_grad():
            rm.copy_(self.m.compressor[1].running_mean); rv.copy_(self.m.compressor[1].running_var)

    def backward(self, g_noisy, g_core, g_readout, kl_scale=1.0):
        c = self.m.compressor
        ps = [c[0].weight, c[0].bias, c[1].weight, c[1].bias, c[3].weight, c[3].bias]
        L = (self.noisy * g_noisy).sum() + (self.core * g_core).sum() + (self.readout * g_readout).sum() + kl_scale * self.kl
        gs = torch.autograd.grad(L, [self.Hf] + ps, allow_unused=True)
        gs = [torch.zeros_like(p) if g is None else g for g, p in zip(gs, [self.Hf] + ps)]
        return (gs[0], gs[1], gs[2], gs[3], gs[4], gs[5].reshape(-1), gs[6])
def attn(C, graph_ptr, w):
    seg = seg_ids(graph_ptr); logit = C @ w
    return torch.cat([F.softmax(logit[seg == b], 0) for b in range(graph_ptr.numel() - 1)])
def core_cand_attn_fwd(C, graph_ptr, w): a = attn(C, graph_ptr, w); return a, C * a[:, None]
def core_cand_attn_bwd(C, alpha, gT, graph_ptr, w):
    C2, w2 = C.detach().clone().requires_grad_(), w.detach().clone().requires_grad_()
    a = attn(C2, graph_ptr, w2); T = C2 * a[:, None]
    return torch.autograd.grad((T * gT).sum(), [C2, w2])
class HeadMLP:
    def __init__(self, H, W1, b1, W2, b2): self.p = [t.detach().clone().requires_grad_() for t in (W1, b1, W2, b2)]
    def forward(self, noisy, C, alpha):
        self.noisy = noisy.detach().clone().requires_grad_(); self.aC = (C * alpha[:, None]).detach().clone().requires_grad_()
        imap = torch.cat((self.noisy, self.aC), -1); W1, b1, W2, b2 = self.p
        self.Z = torch.relu(imap @ W1.t() + b1) @ W2.t() + b2
        return self.Z.detach(), imap.detach()
    def backward(self, gZ):
        g = torch.autograd.grad((self.Z * gZ).sum(), [self.noisy, self.aC] + self.p)
        return torch.stack([g[0], g[1]]), g[2], g[3], g[4], g[5]
def recon_adj(Z, indptr, indices, scale=1.0, want_grad=True):
    Z2 = Z.detach().clone().requires_grad_(); N = Z.shape[0]
    ip, ix = indptr.long(), indices.long(); dst = torch.repeat_interleave(torch.arange(N), ip[1:] - ip[:-1])
    adj = torch.zeros(N, N); adj[ix, dst] = 1.0
    loss = ((Z2 @ Z2.t() - adj) ** 2).sum() / N
    return loss.detach().reshape(1), torch.autograd.grad(loss * scale, Z2)[0]
def contrastive(core, readout, scale=1.0, want_grad=True):
    c, r = core.detach().clone().requires_grad_(), readout.detach().clone().requires_grad_()
    loss = _or.batched_semi_loss(c, r, c.shape[0])
    g = torch.autograd.grad(loss * scale, [c, r])
    return loss.detach().reshape(1), g[0], g[1]
def _eg(f):
    def w(*a, **k):
        with torch.enable_grad(): return f(*a, **k)
    return w
for _c in (CoreGate, HeadMLP):
    _c.forward = _eg(_c.forward); _c.backward = _eg(_c.backward)
core_cand_attn_bwd, recon_adj, contrastive = _eg(core_cand_attn_bwd), _eg(recon_adj), _eg(contrastive)

_NAMES = ("input_proj graph_aggregate segment_sum_w linear_fwd linear_bwd_w transfer_bwd segment_sum_bwd CoreGate "
          "core_cand_attn_fwd core_cand_attn_bwd HeadMLP recon_adj contrastive").split()


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("path", ENC_GOLD, ids=[os.path.basename(p) for p in ENC_GOLD])
def test_composed_step_host_logic_matches_reference_golden(path, monkeypatch):
    import models
    from scgib_b200 import models as M, ops
    from scgib_b200.graph import BatchedGraph, EgoBatch
    for n in _NAMES:
        monkeypatch.setattr(ops, n, globals()[n])
    monkeypatch.setattr(M._HotPathMixin, "_composed_inputs", lambda self, g, ego, dev: (g, ego))
    a = types.SimpleNamespace(recons_type="adj", useAtt=1, readout_f="sum", d_transfer=32, device="cpu", batch_size=128,
                              task="graph_classification", k_transition=1)
    fx = torch.load(path, weights_only=False)
    g, e = RefGraph(**fx["graph"]), RefEgoBatch(**fx["ego"])
    enc, k = fx["meta"]["encoder"], fx["meta"]["k"]
    m = models.Mainmodel(a, 9, hidden_dim=64, num_layers=4, num_heads=4, k_transition=k, encoder=enc)
    assert not m.load_state_dict(fx["state"], strict=False).unexpected_keys
    m.train()
    gate_u, feat_u = draw_noise_like_reference(g.batch_num_nodes().tolist(), 64, fx["meta"]["noise_seed"])
    monkeypatch.setattr(m, "_noise", lambda N, dev: (gate_u, feat_u))
    pg = BatchedGraph(g.graph_ptr, g.indptr, g.indices)
    pg.ndata["x"] = torch.from_numpy(g.x)
    t = lambda a_: torch.from_numpy(a_.astype(np.int32))
    ego = EgoBatch.__new__(EgoBatch)
    ego.parent, ego.k, ego.ego_seed = pg, k, None
    ego.ego_ptr, ego.ego_nodes, ego.sub_indptr, ego.sub_indices = t(e.ego_ptr), t(e.ego_nodes), t(e.sub_indptr), t(e.sub_indices)
    x = F.normalize(pg.ndata["x"].float())
    _, kl, con, rec = m.forward(pg, x, ego, None, None, 1, None, 2, "cpu", 16)
    (kl + rec + con).backward()
    ref = fx["out"]
    for n, gv in (("KL", kl), ("contrastive", con), ("recon", rec)):
        assert abs(float(gv.detach()) - float(ref[n])) <= 2e-6 * abs(float(ref[n])), n
    got = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    assert set(fx["grads"]) <= set(got)
    gmax = max(float(v.abs().max()) for v in fx["grads"].values())
    for n, gref in fx["grads"].items():
        if float(gref.abs().max()) <= 1e-5 * gmax:
            assert float(got[n].abs().max()) <= 1e-4 * gmax, n
            continue
        assert rel(got[n], gref) <= 5e-5, (n, rel(got[n], gref))
    assert int(m.compressor[1].num_batches_tracked) == g.num_graphs
    for n in ("compressor.1.running_mean", "compressor.1.running_var"):
        assert rel(m.state_dict()[n], fx["state_after"][n]) <= 1e-5, n
