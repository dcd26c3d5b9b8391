"""Fine-tuning path on the GPU: the Set2Set + predict head kernels against the oracle's restatement,
and the drop-in ``models.Mainmodel_finetuning`` against golden vectors recorded from the unmodified reference class."""
import glob
import os
import types

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle.graph_ref import RefEgoBatch, RefGraph, ego_batch_ref, synth_batch
from oracle.scgib_oracle import (OracleFinetune, OracleMainmodel, Set2SetRef, TGraph, normalize_rows, tgraph_from_ego,
                                 tgraph_from_ref, draw_noise_like_reference)
from tests.helpers import product_graph, rel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FT_GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "finetune_*.pt")))


def _seg_graph(sizes):
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    z = torch.zeros(0, dtype=torch.int64)
    return TGraph(seg_ptr=torch.from_numpy(ptr), indptr=torch.zeros(int(ptr[-1]) + 1, dtype=torch.int64), indices=z, src=z, dst=z)


@pytest.mark.parametrize("H,C,T,sigmoid,B", [(64, 10, 2, True, 37), (64, 1, 2, False, 8), (64, 10, 3, True, 100),
                                             (64, 5, 1, True, 9), (128, 10, 2, True, 50)])
def test_finetune_head_matches_oracle(H, C, T, sigmoid, B):
    """Set2Set(H, T, 1) + predict + sigmoid: forward and every gradient against the fp64 restatement (1e-5)."""
    from scgib_b200.engine import FinetuneHead
    rng = np.random.default_rng(H + C + T + B)
    sizes = rng.integers(1, 40, size=B)
    sizes[0] = 1                                            # a single-node graph: softmax over one element
    tg = _seg_graph(sizes)
    N = int(sizes.sum())
    torch.manual_seed(B)
    s2s = Set2SetRef(H, T, 1).double()
    predict = nn.Sequential(nn.Linear(2 * H, H), nn.ReLU(), nn.Linear(H, C)).double()
    Z = torch.randn(N, H, dtype=torch.float64) * 0.7
    Zr = Z.clone().requires_grad_(True)
    s = predict(s2s(tg, Zr))
    if sigmoid:
        s = torch.sigmoid(s)
    g_s = torch.randn(B, C, dtype=torch.float64)
    (s * g_s).sum().backward()

    head = FinetuneHead(H, C, n_iters=T, sigmoid=sigmoid, device=DEV)
    sd = {"s2s." + n: p for n, p in s2s.state_dict().items()}
    sd.update({"predict." + n: p for n, p in predict.state_dict().items()})
    head.load_state_dict({n: t.float() for n, t in sd.items()})
    gp = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)).to(DEV)
    scores, readout = head.forward(Z.float().to(DEV), gp, want_readout=True)
    assert rel(scores.cpu(), s.detach()) <= 1e-5
    assert rel(readout.cpu(), s2s(tg, Z).detach()) <= 1e-5
    gZ = head.backward(g_s.float().to(DEV))
    assert rel(gZ.cpu(), Zr.grad) <= 2e-5
    ref = {"s2s." + n: p.grad for n, p in s2s.named_parameters()}
    ref.update({"predict." + n: p.grad for n, p in predict.named_parameters()})
    for n, got in head.views(grads=True).items():
        if T == 1 and "weight" in n and "lstm" in n:        # the only LSTM input is zero
            assert float(got.abs().max()) == 0.0 and float(ref[n].abs().max()) == 0.0
            continue
        assert rel(got.cpu(), ref[n]) <= 2e-5, (n, rel(got.cpu(), ref[n]))
    # determinism: fixed-order reductions
    s2 = head.forward(Z.float().to(DEV), gp)
    gZ2 = head.backward(g_s.float().to(DEV))
    assert torch.equal(s2, scores) and torch.equal(gZ2, gZ)


def _args(**kw):
    a = types.SimpleNamespace(recons_type="adj", useAtt=1, readout_f="sum", d_transfer=32, device=DEV, batch_size=128,
                              task="graph_classification", dataset="Peptides-func", k_transition=1)
    a.__dict__.update(kw)
    return a


def _dropin_from_state(tmp_path, state, k, num_classes, hidden=64):
    import models
    pre = models.Mainmodel(_args(), 9, hidden_dim=hidden, num_layers=4, num_heads=4, k_transition=k, encoder="GIN")
    ckpt = str(tmp_path / ("pre_training_synth_GIN_%d_4_%d.pt" % (hidden, k)))
    torch.save(pre, ckpt)
    m = models.Mainmodel_finetuning(_args(), 9, hidden_dim=hidden, num_layers=4, num_heads=4, k_transition=k,
                                    num_classes=num_classes, cp_filename=ckpt, encoder="GIN")
    missing, unexpected = m.load_state_dict(state, strict=False)
    assert not unexpected
    return m.to(DEV)


@pytest.mark.parametrize("path", FT_GOLD, ids=[os.path.basename(p) for p in FT_GOLD])
def test_mainmodel_finetuning_matches_reference_golden(path, tmp_path, monkeypatch):
    """models.Mainmodel_finetuning (CUDA path) on the inputs / weights / noise of the golden run of the UNMODIFIED
    reference class: scores and loss to 1e-5, the 21 gradients of one train_pep_func step, the freeze rule."""
    from scgib_b200.graph import khop_ego_batch
    fx = torch.load(path, weights_only=False)
    g, e = RefGraph(**fx["graph"]), RefEgoBatch(**fx["ego"])
    k, C, H = fx["meta"]["k"], fx["meta"]["num_classes"], int(fx["meta"].get("hidden", 64))
    m = _dropin_from_state(tmp_path, fx["state"], k, C, H)
    assert sorted(n for n, p in m.named_parameters() if p.requires_grad) == fx["trainable"]
    pg = product_graph(g, DEV)
    ego = khop_ego_batch(pg, k)
    assert np.array_equal(ego.ego_nodes.cpu().numpy(), e.ego_nodes)
    x = F.normalize(pg.ndata["x"].float())
    gate_u, feat_u = draw_noise_like_reference(g.batch_num_nodes().tolist(), H, fx["meta"]["noise_seed"])
    monkeypatch.setattr(m, "_noise", lambda N, dev: (gate_u.to(dev), feat_u.to(dev)))
    m.train()
    scores, z1, z2, z3 = m.forward(pg, x, ego, None, 1, None, 2, DEV, g.num_graphs)
    assert (z1, z2, z3) == (0, 0, 0)
    assert rel(scores.detach().cpu(), fx["out"]["scores"]) <= 1e-5
    loss = m.loss(scores, fx["targets"].to(DEV)) / 2
    assert abs(float(loss) - float(fx["out"]["loss"])) <= 1e-5 * abs(float(fx["out"]["loss"]))
    loss.backward()
    got = {n: p.grad.cpu() for n, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(fx["grads"])
    gmax = max(float(v.abs().max()) for v in fx["grads"].values())
    errs = []
    for n, gref in fx["grads"].items():
        if float(gref.abs().max()) <= 1e-6 * gmax:          # mathematically zero (bias in front of a BatchNorm)
            assert float(got[n].abs().max()) <= 1e-4 * gmax, n
            continue
        errs.append((rel(got[n], gref), n))
    errs.sort()
    assert errs[len(errs) // 2][0] <= 5e-5 and errs[-1][0] <= 5e-3, errs[-3:]
    # evaluate_network (train_pep_func.py:187-230): model.eval() -> every BatchNorm uses its running statistics
    m.eval()
    gu2, fu2 = draw_noise_like_reference(g.batch_num_nodes().tolist(), H, fx["meta"]["noise_seed"] + 1)
    monkeypatch.setattr(m, "_noise", lambda N, dev: (gu2.to(dev), fu2.to(dev)))
    with torch.no_grad():
        scores_eval, _, _, _ = m.forward(pg, x, ego, None, 1, None, 2, DEV, g.num_graphs)
    assert rel(scores_eval.cpu(), fx["out"]["scores_eval"]) <= 1e-5
    m.train()
    # BN running statistics of the loaded model are live buffers updated by the forward
    sd = m.state_dict()
    for n, t in fx["state_after"].items():
        if t.dtype.is_floating_point:
            assert rel(sd[n].cpu(), t) <= 1e-5, n
        else:
            assert torch.equal(sd[n].cpu(), t), n


def test_mainmodel_finetuning_b256_vs_fp64_oracle(tmp_path, monkeypatch):
    """A batch of 256 molecules, regression head (no sigmoid): CUDA path against the fp64 oracle."""
    from scgib_b200.graph import khop_ego_batch
    B, k = 256, 1
    g = synth_batch(77, B)
    e = ego_batch_ref(g, k)
    torch.manual_seed(77)
    inner = OracleMainmodel(9)
    ref = OracleFinetune(inner, 9, num_classes=1, task="graph_regression", regression_dataset=True)
    ref.train()
    import models
    pre = models.Mainmodel(_args(), 9, hidden_dim=64, num_layers=4, num_heads=4, k_transition=k, encoder="GIN")
    ckpt = str(tmp_path / "pre.pt")
    torch.save(pre, ckpt)
    m = models.Mainmodel_finetuning(_args(task="graph_regression", dataset="ZINC"), 9, 64, 4, 4, k, 1, ckpt, "GIN")
    missing, unexpected = m.load_state_dict(ref.state_dict(), strict=False)
    assert not unexpected
    m = m.to(DEV).train()
    ref = ref.double()
    pg = product_graph(g, DEV)
    ego = khop_ego_batch(pg, k)
    gate_u, feat_u = torch.rand(g.num_nodes), torch.rand(g.num_nodes, 64)
    monkeypatch.setattr(m, "_noise", lambda N, dev: (gate_u.to(dev), feat_u.to(dev)))
    scores, _, _, _ = m.forward(pg, F.normalize(pg.ndata["x"].float()), ego, None, 1, None, 2, DEV, B)
    targets = torch.randn(B, 1)
    m.lossMAE(scores, targets.to(DEV)).backward()
    xr = normalize_rows(torch.from_numpy(g.x)).double()
    en = torch.from_numpy(e.ego_nodes.astype(np.int64))
    out = ref(tgraph_from_ref(g), xr, tgraph_from_ego(e), xr[en], gate_u.double(), feat_u.double())
    F.l1_loss(out["scores"], targets.double()).backward()
    assert rel(scores.detach().cpu(), out["scores"].detach()) <= 1e-5
    refg = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    got = {n: p.grad.cpu() for n, p in m.named_parameters() if p.grad is not None}
    assert set(refg) == set(got)
    errs = sorted((rel(got[n], refg[n]), n) for n in refg if float(refg[n].abs().max()) > 1e-6)
    assert errs[len(errs) // 2][0] <= 5e-5 and errs[-1][0] <= 5e-3, errs[-3:]


DA_GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "domainadapt_*.pt")))


@pytest.mark.parametrize("h,T,B", [(9, 2, 23), (11, 3, 40), (32, 2, 9), (1, 1, 5)])
def test_padded_set2set_small_width(h, T, B):
    """Set2Set over narrow features (width h <= 32) on the zero-padded H = 32 kernels: output and LSTM gradients."""
    from scgib_b200.engine import PaddedSet2Set
    rng = np.random.default_rng(h + T + B)
    sizes = rng.integers(1, 30, size=B)
    tg = _seg_graph(sizes)
    N = int(sizes.sum())
    torch.manual_seed(h)
    ref = Set2SetRef(h, T, 1).double()
    x = torch.randn(N, h, dtype=torch.float64)
    out = ref(tg, x)
    g_out = torch.randn(B, 2 * h, dtype=torch.float64)
    (out * g_out).sum().backward()
    s2s = PaddedSet2Set(h, n_iters=T, device=DEV)
    lstm = ref.lstm
    s2s.set_params(lstm.weight_ih_l0.to(DEV), lstm.weight_hh_l0.to(DEV), lstm.bias_ih_l0.to(DEV), lstm.bias_hh_l0.to(DEV))
    gp = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)).to(DEV)
    got = s2s.forward(x.float().to(DEV), gp)
    assert rel(got.cpu(), out.detach()) <= 1e-5
    grads = s2s.backward(g_out.float().to(DEV))
    for gg, name in zip(grads, ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")):
        r = getattr(lstm, name).grad
        if float(r.abs().max()) == 0.0:
            assert float(gg.abs().max()) == 0.0
        else:
            assert rel(gg.cpu(), r) <= 2e-5, (name, rel(gg.cpu(), r))


@pytest.mark.parametrize("path", DA_GOLD, ids=[os.path.basename(p) for p in DA_GOLD])
def test_mainmodel_domainadapt_matches_reference_golden(path, tmp_path, monkeypatch):
    """models.Mainmodel_domainadapt (CUDA path) vs the golden run of the unmodified reference class: X_loss to 1e-5 and
    the 73 gradients of one train_epoch_domainadaptation step."""
    import models
    from scgib_b200.graph import khop_ego_batch
    fx = torch.load(path, weights_only=False)
    g, e = RefGraph(**fx["graph"]), RefEgoBatch(**fx["ego"])
    k = fx["meta"]["k"]
    pre = models.Mainmodel(_args(), 9, hidden_dim=64, num_layers=4, num_heads=4, k_transition=k, encoder="GIN")
    ckpt = str(tmp_path / "pre.pt")
    torch.save(pre, ckpt)
    m = models.Mainmodel_domainadapt(_args(), 9, hidden_dim=64, num_layers=4, num_heads=4, k_transition=k, num_classes=10,
                                     cp_filename=ckpt, encoder="GIN")
    missing, unexpected = m.load_state_dict(fx["state"], strict=False)
    assert not unexpected
    m = m.to(DEV).train()
    pg = product_graph(g, DEV)
    ego = khop_ego_batch(pg, k)
    x = F.normalize(pg.ndata["x"].float())
    gate_u, feat_u = draw_noise_like_reference(g.batch_num_nodes().tolist(), 64, fx["meta"]["noise_seed"])
    monkeypatch.setattr(m, "_noise", lambda N, dev: (gate_u.to(dev), feat_u.to(dev)))
    loss = m.forward(pg, x, ego, None, None, 1, None, 2, DEV, g.num_graphs)
    assert abs(float(loss) - float(fx["out"]["X_loss"])) <= 1e-5 * abs(float(fx["out"]["X_loss"]))
    loss.backward()
    got = {n: p.grad.cpu() for n, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(fx["grads"])
    gmax = max(float(v.abs().max()) for v in fx["grads"].values())
    errs = []
    for n, gref in fx["grads"].items():
        if float(gref.abs().max()) <= 1e-6 * gmax:
            assert float(got[n].abs().max()) <= 1e-4 * gmax, n
            continue
        gg = got[n]
        if n == "model.attn_layer.weight":
            gg, gref = gg[:, 64:], gref[:, 64:]
        errs.append((rel(gg, gref), n))
    errs.sort()
    assert errs[len(errs) // 2][0] <= 5e-5 and errs[-1][0] <= 5e-3, errs[-3:]


def test_finetune_forward_full_size_vs_fp64_oracle():
    """B = 4096 PCQM4Mv2-shape graphs (the bench size): features-only forward + Set2Set/predict head on the GPU against the
    fp64 vectorised oracle of the feature path followed by the fp64 Set2Set restatement; gradients of the head at scale."""
    from scgib_b200.engine import DeviceBatch, FinetuneHead
    from scgib_b200.graph import khop_ego_batch
    from tests.helpers import engine_from_oracle
    B = 4096
    g = synth_batch(123, B) if B <= 512 else __import__("oracle.graph_ref", fromlist=["synth_batch_fast"]).synth_batch_fast(123, B)
    e = ego_batch_ref(g, 1)
    torch.manual_seed(123)
    m = OracleMainmodel(9)
    s2s = Set2SetRef(64, 2, 1)
    predict = nn.Sequential(nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 10))
    gen = torch.Generator().manual_seed(5)
    gate_u, feat_u = torch.rand(g.num_nodes, generator=gen), torch.rand(g.num_nodes, 64, generator=gen)
    eng = engine_from_oracle(m, DEV)
    pg = product_graph(g, DEV)
    b = DeviceBatch(pg, khop_ego_batch(pg, 1), pg.ndata["x"], normalize_x=True)
    Z = eng.forward_features(b, gate_u.to(DEV), feat_u.to(DEV))
    head = FinetuneHead(64, 10, n_iters=2, sigmoid=True, device=DEV)
    sd = {"s2s." + n: p for n, p in s2s.state_dict().items()}
    sd.update({"predict." + n: p for n, p in predict.state_dict().items()})
    head.load_state_dict({n: t.float() for n, t in sd.items()})
    scores = head.forward(Z, pg.graph_ptr)
    # fp64 truth
    m64 = OracleMainmodel(9).double()
    m64.load_state_dict({n: (v.double() if v.dtype.is_floating_point else v) for n, v in m.state_dict().items()})
    xr = normalize_rows(torch.from_numpy(g.x).double())
    en = torch.from_numpy(e.ego_nodes.astype(np.int64))
    tg = tgraph_from_ref(g)
    with torch.no_grad():
        Zt = m64.forward_vectorised(tg, xr, tgraph_from_ego(e), en, gate_u.double(), feat_u.double())["Z"]
    assert rel(Z.cpu(), Zt) <= 2e-5
    s2s64, pred64 = s2s.double(), predict.double()
    Zr = Zt.clone().requires_grad_(True)
    st = torch.sigmoid(pred64(s2s64(tg, Zr)))
    assert rel(scores.cpu(), st.detach()) <= 2e-5
    g_s = torch.randn(B, 10, dtype=torch.float64, generator=torch.Generator().manual_seed(6)) / B
    (st * g_s).sum().backward()
    gZ = head.backward(g_s.float().to(DEV))
    assert rel(gZ.cpu(), Zr.grad) <= 1e-4
    ref = {"s2s." + n: p.grad for n, p in s2s64.named_parameters()}
    ref.update({"predict." + n: p.grad for n, p in pred64.named_parameters()})
    for n, got in head.views(grads=True).items():
        assert rel(got.cpu(), ref[n]) <= 1e-4, (n, rel(got.cpu(), ref[n]))
