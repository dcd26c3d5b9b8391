"""--encoder GraphSAGE / GCN (reference models.py:75-104) on the B200 path: the operator kernels of csrc/encoder_ops.cu against
torch, and models.Mainmodel(encoder=...) against the golden vectors of the UNMODIFIED reference (tests/golden/enc_*.pt, made by
make_golden.py on the DGL stand-in) and against the oracle restatement at a larger batch."""
import glob
import os
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.graph_ref import RefEgoBatch, RefGraph, ego_batch_ref, synth_batch
from oracle.scgib_oracle import (OracleMainmodel, draw_noise_like_reference, normalize_rows, tgraph_from_ego, tgraph_from_ref)
from tests.helpers import product_graph, rel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ENC_GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "enc_*.pt")))


def _args(**kw):
    a = types.SimpleNamespace(recons_type="adj", useAtt=1, readout_f="sum", d_transfer=32, device=DEV, batch_size=128,
                              task="graph_classification", k_transition=1)
    a.__dict__.update(kw)
    return a


def _csr(g):
    ip = torch.from_numpy(g.indptr.astype(np.int64))
    ix = torch.from_numpy(g.indices.astype(np.int64))
    dst = torch.repeat_interleave(torch.arange(g.num_nodes), ip[1:] - ip[:-1])
    return ip, ix, dst


@pytest.mark.parametrize("W", [32, 64, 128, 256])
@pytest.mark.parametrize("norms", [(0, 1), (1, 0), (2, 2), (0, 0)])
def test_graph_aggregate_op(W, norms):
    from scgib_b200 import ops
    g = synth_batch(81, 37)
    ip, ix, dst = _csr(g)
    V = g.num_nodes
    torch.manual_seed(W)
    rows = V + 11
    h = torch.randn(rows, W, dtype=torch.float64)
    row_map = torch.randint(0, rows, (V,))
    add = torch.randn(V, W, dtype=torch.float64)
    deg = (ip[1:] - ip[:-1]).double().clamp(min=1)
    f = lambda mode: torch.ones(V, dtype=torch.float64) if mode == 0 else (1.0 / deg if mode == 1 else deg.pow(-0.5))
    fs, fd = f(norms[0]), f(norms[1])
    src = h[row_map]
    ref = torch.zeros(V, W, dtype=torch.float64).index_add(0, dst, src[ix] * fs[ix][:, None]) * fd[:, None] + add
    got = ops.graph_aggregate(h.float().to(DEV), ip.int().to(DEV), ix.int().to(DEV), norms[0], norms[1],
                              row_map=row_map.int().to(DEV), add=add.float().to(DEV))
    assert rel(got, ref) <= 2e-6
    got2 = ops.graph_aggregate(h[:V].float().to(DEV), ip.int().to(DEV), ix.int().to(DEV), norms[0], norms[1])
    ref2 = torch.zeros(V, W, dtype=torch.float64).index_add(0, dst, h[:V][ix] * fs[ix][:, None]) * fd[:, None]
    assert rel(got2, ref2) <= 2e-6


@pytest.mark.parametrize("W", [32, 64, 128, 256])
def test_segment_sum_w_op(W):
    from scgib_b200 import ops
    ptr = torch.tensor([0, 3, 3, 50, 64, 65], dtype=torch.int32)
    h = torch.randn(65, W)
    seg = torch.repeat_interleave(torch.arange(5), (ptr[1:] - ptr[:-1]).long())
    ref = torch.zeros(5, W, dtype=torch.float64).index_add(0, seg, h.double())
    assert rel(ops.segment_sum_w(h.to(DEV), ptr.to(DEV)), ref) <= 2e-6


@pytest.mark.parametrize("V,K0,K1,O", [(1000, 32, 32, 64), (777, 64, 64, 64), (130, 128, 0, 32), (513, 256, 0, 256), (64, 32, 0, 128)])
@pytest.mark.parametrize("kxo", [False, True])
def test_linear_fwd_op(V, K0, K1, O, kxo):
    from scgib_b200 import ops
    torch.manual_seed(V + O)
    rows = V + 5
    X0 = torch.randn(rows, K0, dtype=torch.float64)
    map0 = torch.randint(0, rows, (V,))
    M0 = torch.randn(rows, K0, dtype=torch.float64)
    W0 = torch.randn(O, K0, dtype=torch.float64)
    bias = torch.randn(O, dtype=torch.float64)
    ref = (X0 * (M0 > 0))[map0] @ W0.t() + bias
    f = lambda t: t.float().to(DEV).contiguous()
    kw = {}
    if K1:
        X1, W1 = torch.randn(V, K1, dtype=torch.float64), torch.randn(O, K1, dtype=torch.float64)
        ref = ref + X1 @ W1.t()
        kw = dict(X1=f(X1), W1=f(W1.t()) if kxo else f(W1), w1_kxo=kxo)
    ref = torch.relu(ref)
    got = ops.linear_fwd(f(X0), f(W0.t()) if kxo else f(W0), O, w0_kxo=kxo, bias=f(bias), relu=True, map0=map0.int().to(DEV),
                         M0=f(M0), V=V, **kw)
    assert rel(got, ref) <= 5e-6


@pytest.mark.parametrize("V,O,K", [(1000, 64, 32), (5000, 64, 64), (40000, 128, 128), (300, 256, 32), (2000, 128, 256)])
@pytest.mark.parametrize("kxo", [False, True])
def test_linear_bwd_w_op(V, O, K, kxo):
    from scgib_b200 import ops
    torch.manual_seed(V + K)
    rows = V + 3
    G, M = torch.randn(V, O, dtype=torch.float64), torch.randn(V, O, dtype=torch.float64)
    X = torch.randn(rows, K, dtype=torch.float64)
    mp = torch.randint(0, rows, (V,))
    Gm = G * (M > 0)
    dW_ref, db_ref = Gm.t() @ X[mp], Gm.sum(0)
    f = lambda t: t.float().to(DEV).contiguous()
    dW = torch.empty(K, O, device=DEV) if kxo else torch.empty(O, K, device=DEV)
    db = torch.empty(O, device=DEV)
    ops.linear_bwd_w(f(G), f(X), dW, db, M=f(M), map=mp.int().to(DEV), kxo=kxo)
    first = dW.clone()
    assert rel(dW.t() if kxo else dW, dW_ref) <= 1e-5 and rel(db, db_ref) <= 1e-5
    ops.linear_bwd_w(f(G), f(X), dW, db, M=f(M), map=mp.int().to(DEV), kxo=kxo, accumulate=True)
    assert rel(dW.t() if kxo else dW, 2 * dW_ref) <= 1e-5 and rel(db, 2 * db_ref) <= 1e-5
    again = torch.empty_like(first)
    ops.linear_bwd_w(f(G), f(X), again, None, M=f(M), map=mp.int().to(DEV), kxo=kxo)
    assert torch.equal(first, again)          # fixed-order reduction: bit-identical reruns


def test_transfer_bwd_op():
    from scgib_b200 import ops
    g = synth_batch(83, 29)
    e = ego_batch_ref(g, 1)
    x = torch.from_numpy(g.x).double()
    xn = F.normalize(x)
    N, Ns = g.num_nodes, e.num_rows
    en = torch.from_numpy(e.ego_nodes.astype(np.int64))
    g0, g1 = torch.randn(N, 32, dtype=torch.float64), torch.randn(Ns, 32, dtype=torch.float64)
    ref = g0.t() @ xn + g1.t() @ xn[en]
    got = ops.transfer_bwd(x.float().to(DEV), g0.float().to(DEV), g1.float().to(DEV), en.int().to(DEV), normalize=True)
    assert rel(got, ref) <= 1e-5
    # with the CSRs: the layer-0 aggregation backward of the GIN path moved to the features
    ip, ix, dst = _csr(g)
    sip = torch.from_numpy(e.sub_indptr.astype(np.int64))
    six = torch.from_numpy(e.sub_indices.astype(np.int64))
    sdst = torch.repeat_interleave(torch.arange(Ns), sip[1:] - sip[:-1])
    xa0 = xn + torch.zeros_like(xn).index_add(0, dst, xn[ix])
    xe = xn[en]
    xa1 = xe + torch.zeros_like(xe).index_add(0, sdst, xe[six])
    ref2 = g0.t() @ xa0 + g1.t() @ xa1
    got2 = ops.transfer_bwd(x.float().to(DEV), g0.float().to(DEV), g1.float().to(DEV), en.int().to(DEV), normalize=True,
                            csr0=(ip.int().to(DEV), ix.int().to(DEV)), csr1=(sip.int().to(DEV), six.int().to(DEV)))
    assert rel(got2, ref2) <= 1e-5


def _run_dropin(monkeypatch, encoder, hidden, g, e, state, gate_u, feat_u, k):
    import models
    from scgib_b200.graph import khop_ego_batch
    m = models.Mainmodel(_args(k_transition=k), 9, hidden_dim=hidden, num_layers=4, num_heads=4, k_transition=k, encoder=encoder)
    missing, unexpected = m.load_state_dict(state, strict=False)
    assert not unexpected, unexpected
    m = m.to(DEV)
    pg = product_graph(g, DEV)
    ego = khop_ego_batch(pg, k)
    x = F.normalize(pg.ndata["x"].float())
    monkeypatch.setattr(m, "_noise", lambda N, dev: (gate_u.to(dev), feat_u.to(dev)))
    m.train()
    _, kl, con, rec = m.forward(pg, x, ego, None, None, 1, None, 2, DEV, max(16, g.num_graphs))
    (kl + rec + con).backward()
    torch.cuda.synchronize()
    return m, kl, con, rec


@pytest.mark.parametrize("path", ENC_GOLD, ids=[os.path.basename(p) for p in ENC_GOLD])
def test_mainmodel_sage_gcn_matches_reference_golden(path, monkeypatch):
    fx = torch.load(path, weights_only=False)
    g, e = RefGraph(**fx["graph"]), RefEgoBatch(**fx["ego"])
    enc, hidden, k = fx["meta"]["encoder"], int(fx["meta"].get("hidden", 64)), fx["meta"]["k"]
    gate_u, feat_u = draw_noise_like_reference(g.batch_num_nodes().tolist(), hidden, fx["meta"]["noise_seed"])
    m, kl, con, rec = _run_dropin(monkeypatch, enc, hidden, g, e, fx["state"], gate_u, feat_u, k)
    ref = fx["out"]
    for name, got in (("KL", kl), ("contrastive", con), ("recon", rec)):
        assert abs(float(got) - float(ref[name])) <= 1e-5 * abs(float(ref[name])), (name, float(got), float(ref[name]))
    last = m._composed_last
    for name in ("interaction_map", "Z", "noisy", "graph_readout"):
        assert rel(last[name], ref[name]) <= 1e-5, (name, rel(last[name], ref[name]))
    for n in ("compressor.1.running_mean", "compressor.1.running_var"):      # B sequential per-graph updates (models.py:642)
        assert rel(m.state_dict()[n].cpu(), fx["state_after"][n]) <= 1e-5, n
    assert int(m.compressor[1].num_batches_tracked) == int(fx["state_after"]["compressor.1.num_batches_tracked"])
    got = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    assert set(fx["grads"]) <= set(got)
    gmax = max(float(v.abs().max()) for v in fx["grads"].values())
    for n, gref in fx["grads"].items():
        if float(gref.abs().max()) <= 1e-5 * gmax:            # attn_layer.bias / compressor.0.bias: zero up to rounding
            assert float(got[n].abs().max()) <= 1e-4 * gmax, n
            continue
        assert rel(got[n].cpu(), gref) <= 2e-4, (n, rel(got[n].cpu(), gref))


@pytest.mark.parametrize("encoder,hidden,k", [("GraphSAGE", 64, 1), ("GCN", 64, 2), ("GraphSAGE", 64, 3), ("GraphSAGE", 128, 2), ("GCN", 128, 1)])
def test_mainmodel_sage_gcn_matches_oracle_batch(encoder, hidden, k, monkeypatch):
    g = synth_batch(91 + hidden + k, 200)
    e = ego_batch_ref(g, k)
    torch.manual_seed(7)
    ref = OracleMainmodel(9, hidden, 32, 4, encoder=encoder).double()
    gate_u, feat_u = torch.rand(g.num_nodes), torch.rand(g.num_nodes, hidden)
    state = {n: t.float() for n, t in ref.state_dict().items()}
    m, kl, con, rec = _run_dropin(monkeypatch, encoder, hidden, g, e, state, gate_u, feat_u, k)
    xr = normalize_rows(torch.from_numpy(g.x)).double()
    en = torch.from_numpy(e.ego_nodes.astype(np.int64))
    out = ref.forward_vectorised(tgraph_from_ref(g), xr, tgraph_from_ego(e), en, gate_u.double(), feat_u.double())
    (out["KL"] + out["recon"] + out["contrastive"]).backward()
    for name, got in (("KL", kl), ("contrastive", con), ("recon", rec)):
        assert abs(float(got) - float(out[name])) <= 1e-5 * abs(float(out[name])), (name, float(got), float(out[name]))
    last = m._composed_last
    for name in ("interaction_map", "Z", "noisy", "graph_readout"):
        assert rel(last[name], out[name]) <= 1e-5, (name, rel(last[name], out[name]))
    refg = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    got = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    gmax = max(float(v.abs().max()) for v in refg.values())
    errs = []
    for n, gref in refg.items():
        if float(gref.abs().max()) <= 1e-6 * gmax:
            continue
        errs.append(rel(got[n].cpu(), gref))
        assert errs[-1] <= 5e-4, (n, errs[-1])
    errs.sort()
    assert errs[len(errs) // 2] <= 5e-5
    # one optimiser step on the module's own parameters, and a second forward / backward (fresh saved state)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    before = m.Encoder1.conv1.weight.clone() if encoder == "GCN" else m.Encoder1.conv1.fc_self.weight.clone()
    opt.step()
    after = m.Encoder1.conv1.weight if encoder == "GCN" else m.Encoder1.conv1.fc_self.weight
    assert not torch.equal(before, after)


@pytest.mark.parametrize("encoder", ["GraphSAGE", "GCN"])
def test_exp_pretraining_cli_stages_with_encoder(encoder, tmp_path, monkeypatch):
    """python exp_pretraining.py --encoder GraphSAGE|GCN: the three stages (Mainmodel, then Mainmodel_continue around the
    pickled model, twice) through the reference's CLI surface; finite parameters that moved."""
    import exp_pretraining as ep
    monkeypatch.chdir(tmp_path)
    ep.args = ep.build_parser().parse_args(["--device", DEV, "--pt_epoches", "2", "--batch_size", "32", "--synthetic", "96",
                                            "--encoder", encoder, "--output_path", str(tmp_path) + "/outputs/"])
    ep.device = torch.device(DEV)
    ep.main()
    names = sorted(os.listdir(tmp_path / "outputs"))
    assert names == ["pre_training_PCQM4Mv2_%s_64_4_1.pt" % encoder, "pre_training_PCQM4Mv2_QM9_%s_64_4_1.pt" % encoder,
                     "pre_training_PCQM4Mv2_QM9_mol-PCBA_%s_64_4_1.pt" % encoder]
    last = torch.load(tmp_path / "outputs" / names[-1], weights_only=False)
    assert type(last).__name__ == "Mainmodel_continue" and type(last.model.Encoder1).__name__ == ("GCN" if encoder == "GCN" else "GraphSAGE")
    assert all(torch.isfinite(p).all() for p in last.parameters())


def test_extract_features_api_with_gcn():
    import models
    from scgib_b200.graph import khop_ego_batch
    g = synth_batch(52, 16)
    torch.manual_seed(52)
    m = models.Mainmodel(_args(), 9, 64, 4, 4, 1, "GCN").to(DEV)
    pg = product_graph(g, DEV)
    ego = khop_ego_batch(pg, 1)
    t = m.transfer_d(F.normalize(pg.ndata["x"].float())).detach()
    imap, kl_t, noisy, readout = m.extract_features(pg.batch_num_nodes(), pg, t, ego, None, DEV)
    assert imap.shape == (g.num_nodes, 128) and noisy.shape == (g.num_nodes, 64) and readout.shape == (16, 64)
    assert torch.equal(imap[:, :64], noisy) and torch.isfinite(kl_t).all()


def test_sage_gcn_unsupported_width_fails_loudly():
    import models
    with pytest.raises(NotImplementedError):
        models.Mainmodel(_args(), 9, 96, 4, 4, 1, "GraphSAGE")


def test_sage_gcn_eval_mode_fails_loudly():
    """model.eval() would need the running statistics inside the per-graph BatchNorm: not built for the composed path."""
    import models
    from scgib_b200.graph import khop_ego_batch
    g = synth_batch(5, 8)
    m = models.Mainmodel(_args(), 9, 64, 4, 4, 1, "GCN").to(DEV).eval()
    pg = product_graph(g, DEV)
    with pytest.raises(NotImplementedError):
        m.forward(pg, F.normalize(pg.ndata["x"].float()), khop_ego_batch(pg, 1), None, None, 1, None, 2, DEV, 16)


def test_loss_operators_hidden_128():
    """scgib_recon_adj_h_f32 / scgib_contrastive_h_f32 at hidden 128 (value + gradient) against the oracle's dense formulas in fp64."""
    from scgib_b200 import ops
    from oracle.scgib_oracle import tgraph_from_ref as _tg
    g = synth_batch(8, 120)
    tg = _tg(g)
    torch.manual_seed(9)
    m = OracleMainmodel(9, 128).double()
    Z = (torch.randn(g.num_nodes, 128, dtype=torch.float64) * 0.2).requires_grad_(True)
    rec = m.loss_recon_adj(Z, tg)
    rec.backward()
    pg = product_graph(g, DEV)
    loss, gZ = ops.recon_adj(Z.detach().float().to(DEV), pg.indptr, pg.indices, scale=0.5)
    assert abs(float(loss) - float(rec)) <= 1e-5 * abs(float(rec))
    assert rel(gZ.cpu(), 0.5 * Z.grad) <= 2e-5
    for B in (7, 300):
        core = (torch.randn(B, 128, dtype=torch.float64) * 2).requires_grad_(True)
        ro = (torch.randn(B, 128, dtype=torch.float64) * 3).requires_grad_(True)
        con = m.batched_semi_loss(core, ro, B)
        con.backward()
        loss, g1, g2 = ops.contrastive(core.detach().float().to(DEV), ro.detach().float().to(DEV), scale=2.0)
        assert abs(float(loss) - float(con)) <= 1e-5 * abs(float(con)), B
        assert rel(g1.cpu(), 2.0 * core.grad) <= 5e-5 and rel(g2.cpu(), 2.0 * ro.grad) <= 5e-5, B
