"""Generate golden vectors by running the UNMODIFIED reference models.py (CPU) on a DGL stand-in.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/pretrain_k{K}_b{B}.pt.  The fixtures travel to the GPU box; this script and
/root/reference do not need to.
"""
import argparse
import importlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle.graph_ref import ego_batch_ref, synth_batch  # noqa: E402


def load_reference_models(ref_root="/root/reference"):
    import dgl_stub
    dgl_stub.install()
    sys.path.insert(0, ref_root)
    return importlib.import_module("models"), dgl_stub


def make(seed, B, k, ref_models, dgl_stub, out_path, recons_type="adj", hidden=64, encoder="GIN"):
    g = synth_batch(seed, B)
    e = ego_batch_ref(g, k)
    s, d = g.edges()
    bg = dgl_stub.StubGraph(g.graph_ptr, s, d, g.num_nodes)
    bg.ndata["x"] = torch.from_numpy(g.x)
    # ego batch: sub-graph s is the ego-net of global node s (exp_pretraining.py:269-274, 308-309)
    es_dst = np.repeat(np.arange(e.num_rows), np.diff(e.sub_indptr))
    eg = dgl_stub.StubGraph(e.ego_ptr, e.sub_indices, es_dst, e.num_rows)
    eg.ndata["x"] = torch.from_numpy(g.x[e.ego_nodes])

    args = types.SimpleNamespace(recons_type=recons_type, useAtt=1, readout_f="sum", d_transfer=32, device="cpu")
    torch.manual_seed(seed)
    model = ref_models.Mainmodel(args, 9, hidden_dim=hidden, num_layers=4, num_heads=4, k_transition=k, encoder=encoder)
    model.train()
    state0 = {n: t.detach().clone() for n, t in model.state_dict().items()}
    batch_logMs = None
    if recons_type == "logM":
        # exp_pretraining.py:260-264, 353-356: per graph, the reference's own util.getM_logM on the (stub) DGL graph
        import util as ref_util
        batch_logMs = []
        for b in range(g.num_graphs):
            v0, v1 = int(g.graph_ptr[b]), int(g.graph_ptr[b + 1])
            sel = (d >= v0) & (d < v1)
            one = dgl_stub.StubGraph(np.asarray([0, v1 - v0]), s[sel] - v0, d[sel] - v0, v1 - v0)
            _, logM = ref_util.getM_logM(one, kstep=k)
            batch_logMs.append(torch.from_numpy(np.array(logM)).float())

    # exp_pretraining.py:300-322
    batch_x = F.normalize(bg.ndata["x"].float())
    x_subs = F.normalize(eg.ndata["x"].float())
    noise_seed = 1000 + seed
    torch.manual_seed(noise_seed)
    _, KL, con, rec = model.forward(bg, batch_x, eg, batch_logMs, x_subs, 1, bg.edges(), 2, "cpu", B)
    # re-run the two pieces the forward does not return (deterministic given the same RNG state)
    torch.manual_seed(noise_seed)
    imap, KLt, noisy, readout = model.extract_features(bg.batch_num_nodes(), bg,
                                                       model.transfer_d(batch_x), eg,
                                                       model.transfer_d(x_subs), "cpu")
    Z = model.MLP(imap)
    # BN running stats were updated twice now; restore and do the graded forward/backward once more
    model.load_state_dict(state0)
    model.zero_grad()
    torch.manual_seed(noise_seed)
    _, KL2, con2, rec2 = model.forward(bg, batch_x, eg, batch_logMs, x_subs, 1, bg.edges(), 2, "cpu", B)
    assert torch.equal(KL, KL2) and torch.equal(con, con2) and torch.equal(rec, rec2)
    loss = KL2 + rec2 + con2
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    state1 = {n: t.detach().clone() for n, t in model.state_dict().items()
              if "running" in n or "num_batches" in n}
    used = set(grads) | {n for n in state0 if "running" in n or "num_batches" in n or n.endswith(".eps")}
    fx = dict(
        meta=dict(seed=seed, B=B, k=k, noise_seed=noise_seed, reference="models.py Mainmodel (unmodified) on dgl_stub",
                  torch=torch.__version__, recons_type=recons_type, hidden=hidden, encoder=encoder),
        graph=dict(graph_ptr=g.graph_ptr, indptr=g.indptr, indices=g.indices, x=g.x),
        ego=dict(ego_ptr=e.ego_ptr, ego_nodes=e.ego_nodes, sub_indptr=e.sub_indptr, sub_indices=e.sub_indices),
        state={n: t for n, t in state0.items() if n in used},
        state_after=state1,
        out=dict(KL=KL.detach(), contrastive=con.detach(), recon=rec.detach(), interaction_map=imap.detach(),
                 Z=Z.detach(), noisy=noisy.detach(), graph_readout=readout.detach(), KL_tensor=KLt.detach()),
        grads=grads,
    )
    torch.save(fx, out_path)
    print(out_path, os.path.getsize(out_path), "bytes", "KL %.6f con %.6f rec %.6f" % (KL, con, rec),
          "grads:", len(grads))


def make_finetune(seed, B, k, ref_models, dgl_stub, out_path, num_classes=10, hidden=64):
    """Mainmodel_finetuning (models.py:358-543) around a pickled pre-trained Mainmodel, one training step of
    train_pep_func.train_epoch_graph_classification (train_pep_func.py:137-157): BCE(sigmoid scores, targets) / 2."""
    import functools
    import tempfile
    g = synth_batch(seed, B)
    e = ego_batch_ref(g, k)
    s, d = g.edges()
    bg = dgl_stub.StubGraph(g.graph_ptr, s, d, g.num_nodes)
    bg.ndata["x"] = torch.from_numpy(g.x)
    es_dst = np.repeat(np.arange(e.num_rows), np.diff(e.sub_indptr))
    eg = dgl_stub.StubGraph(e.ego_ptr, e.sub_indices, es_dst, e.num_rows)
    eg.ndata["x"] = torch.from_numpy(g.x[e.ego_nodes])

    pre_args = types.SimpleNamespace(recons_type="adj", useAtt=1, readout_f="sum", d_transfer=32, device="cpu")
    torch.manual_seed(seed)
    pre = ref_models.Mainmodel(pre_args, 9, hidden_dim=hidden, num_layers=4, num_heads=4, k_transition=k, encoder="GIN")
    ckpt = os.path.join(tempfile.mkdtemp(), "pre_training_synth_GIN_%d_4_%d.pt" % (hidden, k))
    torch.save(pre, ckpt)                                       # exp_pretraining.py:107 saves the whole module
    args = types.SimpleNamespace(dataset="Peptides-func", readout_f="sum", d_transfer=32, batch_size=B, useAtt=1,
                                 device="cpu", task="graph_classification")
    real_load = torch.load
    torch.load = functools.partial(real_load, weights_only=False)   # torch >= 2.6 default; the reference targets 2.0.1
    try:
        torch.manual_seed(seed + 7)
        model = ref_models.Mainmodel_finetuning(args, 9, hidden_dim=hidden, num_layers=4, num_heads=4, k_transition=k,
                                                num_classes=num_classes, cp_filename=ckpt, encoder="GIN")
    finally:
        torch.load = real_load
    model.train()
    state0 = {n: t.detach().clone() for n, t in model.state_dict().items()}
    trainable = sorted(n for n, p in model.named_parameters() if p.requires_grad)
    targets = (torch.rand(B, num_classes, generator=torch.Generator().manual_seed(seed)) < 0.1).float()

    batch_x = F.normalize(bg.ndata["x"].float())
    x_subs = F.normalize(eg.ndata["x"].float())
    noise_seed = 1000 + seed
    torch.manual_seed(noise_seed)
    scores, _, _, _ = model.forward(bg, batch_x, eg, x_subs, 1, bg.edges(), 2, "cpu", B)
    loss = model.loss(scores, targets) / 2
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    state1 = {n: t.detach().clone() for n, t in model.state_dict().items() if "running" in n or "num_batches" in n}
    # evaluate_network (train_pep_func.py:187-230): model.eval() forward with the running statistics just updated
    model.eval()
    torch.manual_seed(noise_seed + 1)
    with torch.no_grad():
        scores_eval, _, _, _ = model.forward(bg, batch_x, eg, x_subs, 1, bg.edges(), 2, "cpu", B)
    model.train()
    # tensors the forward reads: the outer transfer_d / MLP / s2s / predict and the loaded model's encoders, compressor
    # and attention layer (models.py:508-518); everything else is dead weight in the state dict
    used = set(grads) | {n for n in state0 if n.startswith("model.") and
                         n.split(".")[1] in ("Encoder1", "Encoder2", "compressor", "attn_layer")}
    fx = dict(
        meta=dict(seed=seed, B=B, k=k, noise_seed=noise_seed, num_classes=num_classes, hidden=hidden,
                  reference="models.py Mainmodel_finetuning (unmodified) on dgl_stub", torch=torch.__version__),
        graph=dict(graph_ptr=g.graph_ptr, indptr=g.indptr, indices=g.indices, x=g.x),
        ego=dict(ego_ptr=e.ego_ptr, ego_nodes=e.ego_nodes, sub_indptr=e.sub_indptr, sub_indices=e.sub_indices),
        state={n: t for n, t in state0.items() if n in used}, state_after={n: t for n, t in state1.items() if n in used},
        trainable=trainable, targets=targets,
        out=dict(scores=scores.detach(), loss=loss.detach(), scores_eval=scores_eval.detach()),
        grads=grads,
    )
    torch.save(fx, out_path)
    print(out_path, os.path.getsize(out_path), "bytes", "loss %.6f" % float(loss), "grads:", len(grads),
          "trainable:", len(trainable))


def make_domainadapt(seed, B, k, ref_models, dgl_stub, out_path):
    """Mainmodel_domainadapt (models.py:107-355) around a pickled pre-trained Mainmodel: X_loss + backward, one step of
    train_pep_func.train_epoch_domainadaptation (train_pep_func.py:91-124)."""
    import functools
    import tempfile
    g = synth_batch(seed, B)
    e = ego_batch_ref(g, k)
    s, d = g.edges()
    bg = dgl_stub.StubGraph(g.graph_ptr, s, d, g.num_nodes)
    bg.ndata["x"] = torch.from_numpy(g.x)
    es_dst = np.repeat(np.arange(e.num_rows), np.diff(e.sub_indptr))
    eg = dgl_stub.StubGraph(e.ego_ptr, e.sub_indices, es_dst, e.num_rows)
    eg.ndata["x"] = torch.from_numpy(g.x[e.ego_nodes])
    pre_args = types.SimpleNamespace(recons_type="adj", useAtt=1, readout_f="sum", d_transfer=32, device="cpu")
    torch.manual_seed(seed)
    pre = ref_models.Mainmodel(pre_args, 9, hidden_dim=64, num_layers=4, num_heads=4, k_transition=k, encoder="GIN")
    ckpt = os.path.join(tempfile.mkdtemp(), "pre_training_synth_GIN_64_4_%d.pt" % k)
    torch.save(pre, ckpt)
    args = types.SimpleNamespace(dataset="Peptides-func", readout_f="sum", d_transfer=32, batch_size=B, useAtt=1,
                                 device="cpu", task="graph_classification")
    real_load = torch.load
    torch.load = functools.partial(real_load, weights_only=False)
    try:
        torch.manual_seed(seed + 7)
        model = ref_models.Mainmodel_domainadapt(args, 9, hidden_dim=64, num_layers=4, num_heads=4, k_transition=k,
                                                 num_classes=10, cp_filename=ckpt, encoder="GIN")
    finally:
        torch.load = real_load
    model.train()
    state0 = {n: t.detach().clone() for n, t in model.state_dict().items()}
    batch_x = F.normalize(bg.ndata["x"].float())
    x_subs = F.normalize(eg.ndata["x"].float())
    noise_seed = 1000 + seed
    torch.manual_seed(noise_seed)
    loss = model.forward(bg, batch_x, eg, None, x_subs, 1, bg.edges(), 2, "cpu", B)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    state1 = {n: t.detach().clone() for n, t in model.state_dict().items() if "running" in n or "num_batches" in n}
    used = set(grads) | {n for n in state0 if n.startswith("model.") and
                         n.split(".")[1] in ("Encoder1", "Encoder2", "compressor", "attn_layer")}
    fx = dict(
        meta=dict(seed=seed, B=B, k=k, noise_seed=noise_seed,
                  reference="models.py Mainmodel_domainadapt (unmodified) on dgl_stub", torch=torch.__version__),
        graph=dict(graph_ptr=g.graph_ptr, indptr=g.indptr, indices=g.indices, x=g.x),
        ego=dict(ego_ptr=e.ego_ptr, ego_nodes=e.ego_nodes, sub_indptr=e.sub_indptr, sub_indices=e.sub_indices),
        state={n: t for n, t in state0.items() if n in used}, state_after={n: t for n, t in state1.items() if n in used},
        out=dict(X_loss=loss.detach()), grads=grads,
    )
    torch.save(fx, out_path)
    print(out_path, os.path.getsize(out_path), "bytes", "X_loss %.6f" % float(loss), "grads:", len(grads))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    ref_models, dgl_stub = load_reference_models(a.ref)
    make(0, 6, 1, ref_models, dgl_stub, os.path.join(HERE, "pretrain_k1_b6.pt"))
    make(1, 4, 2, ref_models, dgl_stub, os.path.join(HERE, "pretrain_k2_b4.pt"))
    make(2, 3, 3, ref_models, dgl_stub, os.path.join(HERE, "pretrain_k3_b3.pt"))
    make(6, 5, 2, ref_models, dgl_stub, os.path.join(HERE, "logm_k2_b5.pt"), recons_type="logM")
    make(7, 4, 3, ref_models, dgl_stub, os.path.join(HERE, "logm_k3_b4.pt"), recons_type="logM")
    make_finetune(3, 5, 1, ref_models, dgl_stub, os.path.join(HERE, "finetune_k1_b5.pt"))
    make_finetune(4, 11, 2, ref_models, dgl_stub, os.path.join(HERE, "finetune_k2_b11.pt"))
    make_domainadapt(5, 7, 1, ref_models, dgl_stub, os.path.join(HERE, "domainadapt_k1_b7.pt"))
    # hidden_dim = 128 (--dims 128; BASELINE configs[4] GIN-5x128): the same unmodified reference classes at the wider width
    make(8, 6, 1, ref_models, dgl_stub, os.path.join(HERE, "pretrain_h128_k1_b6.pt"), hidden=128)
    make(9, 4, 2, ref_models, dgl_stub, os.path.join(HERE, "pretrain_h128_k2_b4.pt"), hidden=128)
    make_finetune(10, 5, 1, ref_models, dgl_stub, os.path.join(HERE, "finetune_h128_k1_b5.pt"), hidden=128)
    # --encoder GraphSAGE / GCN (models.py:75-104): the same unmodified Mainmodel on the stub's SAGEConv / GraphConv
    make(11, 6, 1, ref_models, dgl_stub, os.path.join(HERE, "enc_sage_k1_b6.pt"), encoder="GraphSAGE")
    make(12, 4, 2, ref_models, dgl_stub, os.path.join(HERE, "enc_sage_k2_b4.pt"), encoder="GraphSAGE")
    make(13, 6, 1, ref_models, dgl_stub, os.path.join(HERE, "enc_gcn_k1_b6.pt"), encoder="GCN")
    make(14, 4, 2, ref_models, dgl_stub, os.path.join(HERE, "enc_gcn_k2_b4.pt"), encoder="GCN")
