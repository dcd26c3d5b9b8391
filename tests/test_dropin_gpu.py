"""Drop-in boundary on the GPU: the reference's class names / signatures (models.Mainmodel, Mainmodel_continue,
exp_pretraining.train_epoch_pre_training) run the CUDA path and agree with the oracle."""
import os
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.graph_ref import ego_batch_ref, synth_batch
from oracle.scgib_oracle import OracleMainmodel, normalize_rows, tgraph_from_ego, tgraph_from_ref
from tests.helpers import fp64_truth, product_graph, rel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _args(**kw):
    a = types.SimpleNamespace(recons_type="adj", useAtt=1, readout_f="sum", d_transfer=32, device=DEV, batch_size=128,
                              task="graph_classification", k_transition=1)
    a.__dict__.update(kw)
    return a


def test_mainmodel_forward_backward_matches_oracle(monkeypatch):
    import models
    from scgib_b200.graph import khop_ego_batch
    g = synth_batch(51, 48)
    e = ego_batch_ref(g, 1)
    torch.manual_seed(51)
    ref = OracleMainmodel(9)
    torch.manual_seed(51)
    m = models.Mainmodel(_args(), 9, hidden_dim=64, num_layers=4, num_heads=4, k_transition=1, encoder="GIN")
    # same construction order + seed => identical default initialisation as the reference-ordered oracle
    for (n1, p1), (n2, p2) in zip(ref.state_dict().items(), m.state_dict().items()):
        assert n1 == n2 and torch.equal(p1, p2), (n1, n2)
    m = m.to(DEV)
    pg = product_graph(g, DEV)
    ego = khop_ego_batch(pg, 1)
    x = F.normalize(pg.ndata["x"].float())
    gate_u, feat_u = torch.rand(g.num_nodes), torch.rand(g.num_nodes, 64)
    monkeypatch.setattr(m, "_noise", lambda N, dev: (gate_u.to(dev), feat_u.to(dev)))
    m.train()
    _, kl, con, rec = m.forward(pg, x, ego, None, None, 1, None, 2, DEV, 48)
    (kl + rec + con).backward()
    xr = normalize_rows(torch.from_numpy(g.x))
    en = torch.from_numpy(e.ego_nodes.astype(np.int64))
    out = ref.forward_faithful(tgraph_from_ref(g), xr, tgraph_from_ego(e), xr[en], gate_u, feat_u)
    (out["KL"] + out["recon"] + out["contrastive"]).backward()
    for name, got in (("KL", kl), ("contrastive", con), ("recon", rec)):
        assert abs(float(got) - float(out[name])) <= 1e-5 * abs(float(out[name])), name
    refg = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    got = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    assert set(refg) <= set(got)
    # gradient policy of tests/helpers.py: measured against the fp64 oracle, next to the fp32 oracle's own distance
    # from it (a ReLU unit at the kink flips its mask between two correct fp32 implementations)
    _, truth = fp64_truth(ref, g, e, gate_u, feat_u)
    errs = []
    for n in refg:
        if float(truth[n].abs().max()) <= 1e-3:
            continue
        e_got, e_ref = rel(got[n].cpu(), truth[n]), rel(refg[n], truth[n])
        assert e_got <= max(5e-3, 5.0 * e_ref), (n, e_got, e_ref)
        errs.append(e_got)
    errs.sort()
    assert errs[len(errs) // 2] <= 5e-5
    # BN running statistics are live module buffers
    assert rel(m.Encoder1.batch_norms[0].running_mean, ref.Encoder1.batch_norms[0].running_mean) <= 1e-5
    assert int(m.compressor[1].num_batches_tracked) == 48
    # an ordinary optimiser step moves the aliased flat parameters
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    before = m._bridge.engine.params.clone()
    opt.step()
    assert not torch.equal(before, m._bridge.engine.params)


def test_exp_pretraining_cli_stages(tmp_path, monkeypatch):
    """One epoch of each of the three stages through the reference's CLI surface; checkpoints use the reference's
    naming scheme and reload through Mainmodel_continue."""
    import exp_pretraining as ep
    monkeypatch.chdir(tmp_path)
    ep.args = ep.build_parser().parse_args(["--device", DEV, "--pt_epoches", "2", "--batch_size", "32", "--synthetic", "96",
                                            "--output_path", str(tmp_path) + "/outputs/"])
    ep.device = torch.device(DEV)
    ep.main()
    names = sorted(os.listdir(tmp_path / "outputs"))
    assert names == ["pre_training_PCQM4Mv2_GIN_64_4_1.pt", "pre_training_PCQM4Mv2_QM9_GIN_64_4_1.pt",
                     "pre_training_PCQM4Mv2_QM9_mol-PCBA_GIN_64_4_1.pt"]
    last = torch.load(tmp_path / "outputs" / names[-1], weights_only=False)
    assert type(last).__name__ == "Mainmodel_continue" and type(last.model).__name__ == "Mainmodel_continue"
    assert all(torch.isfinite(p).all() for p in last.parameters())


def test_unsupported_encoder_exits_like_reference():
    import models
    with pytest.raises(SystemExit):
        models.Mainmodel(_args(), 9, 64, 4, 4, 1, "Transformer")       # outside SURVEY section 8 (GraphSAGE / GCN: test_gpu_encoders.py)


def test_extract_features_api():
    import models
    from scgib_b200.graph import khop_ego_batch
    g = synth_batch(52, 16)
    torch.manual_seed(52)
    m = models.Mainmodel(_args(), 9, 64, 4, 4, 1, "GIN").to(DEV)
    pg = product_graph(g, DEV)
    ego = khop_ego_batch(pg, 1)
    t = m.transfer_d(F.normalize(pg.ndata["x"].float()))
    imap, kl_t, noisy, readout = m.extract_features(pg.batch_num_nodes(), pg, t, ego, None, DEV)
    assert imap.shape == (g.num_nodes, 128) and noisy.shape == (g.num_nodes, 64) and readout.shape == (16, 64)
    assert torch.equal(imap[:, :64], noisy) and torch.isfinite(kl_t).all()


def test_gradient_accumulation_across_backward_calls(monkeypatch):
    """.grad adopts views of the engine's flat gradient buffer (no copies); a second backward without zero_grad - the
    grad-accumulation loop of train_pep_func.py:155-161 - must still add to the first one's gradients."""
    import models
    from scgib_b200.graph import khop_ego_batch
    torch.manual_seed(5)
    m = models.Mainmodel(_args(), 9, 64, 4, 4, 1, "GIN").to(DEV).train()
    batches = []
    for seed in (61, 62):
        g = synth_batch(seed, 24)
        pg = product_graph(g, DEV)
        gen = torch.Generator().manual_seed(seed)
        batches.append((pg, khop_ego_batch(pg, 1), F.normalize(pg.ndata["x"].float()),
                        torch.rand(g.num_nodes, generator=gen), torch.rand(g.num_nodes, 64, generator=gen)))
    cur = {}
    monkeypatch.setattr(m, "_noise", lambda N, dev: (cur["gu"].to(dev), cur["fu"].to(dev)))

    def run(i):
        pg, ego, x, cur["gu"], cur["fu"] = batches[i]
        _, kl, con, rec = m.forward(pg, x, ego, None, None, 1, None, 2, DEV, 24)
        (kl + rec + con).backward()

    single = []
    for i in range(2):
        m.zero_grad()
        run(i)
        single.append({n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None})
    m.zero_grad()
    run(0)
    run(1)                                                    # no zero_grad in between
    for n, p in m.named_parameters():
        if p.grad is None:
            continue
        want = single[0][n] + single[1][n]
        assert torch.allclose(p.grad, want, rtol=1e-6, atol=1e-6 * float(want.abs().max()) + 1e-12), n


def test_flat_adam_matches_torch_adam(monkeypatch):
    """scgib_b200.optim.FlatAdam (one kernel over the flat buffer) against torch.optim.Adam(lr, weight_decay=5e-5) on the
    same module and the SAME gradients: one backward (two accumulated passes), then three optimiser steps without
    recomputing the gradients (recomputing would compare two chaotic trajectories: Adam turns the rounding-noise
    gradients of dead ReLU units into +-lr steps)."""
    import models
    from scgib_b200.graph import khop_ego_batch
    from scgib_b200.optim import FlatAdam
    g = synth_batch(71, 32)
    pg = product_graph(g, DEV)
    ego = khop_ego_batch(pg, 1)
    x = F.normalize(pg.ndata["x"].float())
    gen = torch.Generator().manual_seed(71)
    gu, fu = torch.rand(g.num_nodes, generator=gen), torch.rand(g.num_nodes, 64, generator=gen)
    finals = []
    for kind in ("torch", "flat"):
        torch.manual_seed(9)
        m = models.Mainmodel(_args(), 9, 64, 4, 4, 1, "GIN").to(DEV).train()
        monkeypatch.setattr(m, "_noise", lambda N, dev: (gu.to(dev), fu.to(dev)))
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=5e-5) if kind == "torch" else \
            FlatAdam(m, lr=1e-3, weight_decay=5e-5)
        opt.zero_grad()
        for _rep in range(2):                                   # accumulated gradient (private copies, see FlatAdam.step)
            _, kl, con, rec = m.forward(pg, x, ego, None, None, 1, None, 2, DEV, 32)
            (kl + rec + con).backward()
        for _step in range(3):
            opt.step()
        finals.append({n: p.detach().clone() for n, p in m.named_parameters()})
    for n in finals[0]:
        d = (finals[1][n] - finals[0][n]).abs()
        assert float(d.max()) <= 2e-6, (n, float(d.max()))       # 0.2 % of one step (lr = 1e-3), three steps taken


def test_mainmodel_recons_type_logm(monkeypatch):
    """--recons_type logM through the drop-in class: reconstruction_loss = loss_recon with k = k_transition step matrices
    computed on the GPU (the reference reads them from pts/*_M_khop_k.pt)."""
    import models
    from scgib_b200.graph import khop_ego_batch
    k = 2
    g = synth_batch(81, 40)
    e = ego_batch_ref(g, k)
    torch.manual_seed(81)
    ref = OracleMainmodel(9)
    torch.manual_seed(81)
    m = models.Mainmodel(_args(recons_type="logM", k_transition=k), 9, 64, 4, 4, k, "GIN").to(DEV).train()
    pg = product_graph(g, DEV)
    ego = khop_ego_batch(pg, k)
    x = F.normalize(pg.ndata["x"].float())
    gate_u, feat_u = torch.rand(g.num_nodes), torch.rand(g.num_nodes, 64)
    monkeypatch.setattr(m, "_noise", lambda N, dev: (gate_u.to(dev), feat_u.to(dev)))
    _, kl, con, rec = m.forward(pg, x, ego, None, None, 1, None, 2, DEV, 40)
    xr = normalize_rows(torch.from_numpy(g.x))
    en = torch.from_numpy(e.ego_nodes.astype(np.int64))
    out = ref.forward_faithful(tgraph_from_ref(g), xr, tgraph_from_ego(e), xr[en], gate_u, feat_u, recon_logm_steps=k)
    for name, got in (("KL", kl), ("contrastive", con), ("recon", rec)):
        assert abs(float(got) - float(out[name])) <= 1e-5 * abs(float(out[name])), name


def test_exp_pep_func_cli_pretrain_adapt_finetune(tmp_path, monkeypatch):
    """The fine-tuning CLI surface end to end (reference exp_pep_func_5.py): pre-train, domain-adapt, fine-tune with
    evaluation (model.eval()) on synthetic Peptides-shape molecules; checkpoints use the reference's naming scheme."""
    import exp_pep_func_5 as ep
    monkeypatch.chdir(tmp_path)
    out = str(tmp_path) + "/outputs/"
    common = ["--device", DEV, "--batch_size", "16", "--synthetic", "48", "--output_path", out, "--num_layers", "4"]
    ep.args = ep.build_parser().parse_args(common + ["--pretrained_mode", "1", "--pt_epoches", "2"])
    ep.device = torch.device(DEV)
    with pytest.raises(SystemExit):                               # the reference quits after the pre-training stage
        ep.main()
    assert os.listdir(tmp_path / "outputs") == ["Peptides-func_GIN_64_4_1.pt"]
    ep.args = ep.build_parser().parse_args(common + ["--pretrained_ds", "Peptides-func", "--domain_adapt", "1",
                                                     "--adapt_epoches", "2", "--ft_epoches", "3"])
    metric = ep.main()
    assert sorted(os.listdir(tmp_path / "outputs")) == ["Peptides-func_GIN_64_4_1.pt", "Peptides-func_GIN_64_4_1_Peptides-func.pt"]
    assert metric == metric and 0.0 < metric < 10.0             # finite BCE-with-logits metric of the test split
