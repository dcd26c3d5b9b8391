"""Stage-by-stage diagnostic of the CUDA path against the fp64 vectorised oracle (run on the GPU box):
    python -m tests.gpu_diag [B] [k]   -> table of max-norm relative errors for every intermediate and gradient.
Test tooling (imports oracle/)."""
import sys

import numpy as np
import torch
import torch.nn.functional as F

from oracle.graph_ref import ego_batch_ref, synth_batch
from oracle.scgib_oracle import OracleMainmodel, draw_noise_like_reference, normalize_rows, tgraph_from_ego, tgraph_from_ref
from tests.helpers import engine_from_oracle, is_zero_grad_param, product_graph, rel


def main(B=64, k=1, seed=7, nseed=None):
    dev = "cuda:0"
    from scgib_b200.engine import DeviceBatch
    from scgib_b200.graph import khop_ego_batch
    g = synth_batch(seed, B)
    e = ego_batch_ref(g, k)
    torch.manual_seed(seed)
    m = OracleMainmodel(9)
    gate_u, feat_u = draw_noise_like_reference(g.batch_num_nodes().tolist(), 64, seed + 1 if nseed is None else nseed)
    m64 = OracleMainmodel(9).double()
    m64.load_state_dict({n: v.double() if v.dtype.is_floating_point else v for n, v in m.state_dict().items()})
    tg, te = tgraph_from_ref(g), tgraph_from_ego(e)
    x = normalize_rows(torch.from_numpy(g.x).double())
    en = torch.from_numpy(e.ego_nodes.astype(np.int64))

    # oracle with retained intermediates
    keep = {}
    t = m64.transfer_d(x)
    def enc(encoder, gr, h, tag):
        for i, layer in enumerate(encoder.ginlayers):
            y = layer(gr, h); y.retain_grad(); keep["y%s_%d" % (tag, i)] = y
            h = F.relu(encoder.batch_norms[i](y))
        return h
    a_in = {}
    def pre_hook(tag):
        def f(mod, inp):
            inp[0].retain_grad(); a_in[tag] = inp[0]
        return f
    for tag, encm in (("1", m64.Encoder1), ("2", m64.Encoder2)):
        for l, layer in enumerate(encm.ginlayers):
            layer.apply_func.register_forward_pre_hook(pre_hook("%s_%d" % (tag, l)))
    out = m64.forward_vectorised(tg, x, te, en, gate_u.double(), feat_u.double())
    a_saved = dict(a_in)
    for name in ("H", "C", "Z", "noisy", "core_readout", "graph_readout"):
        out[name].retain_grad()
    enc(m64.Encoder1, tg, t, "1"); enc(m64.Encoder2, te, t[en], "2")     # per-layer pre-BN outputs (same math)
    (out["KL"] + out["recon"] + out["contrastive"]).backward()
    ref_grads = {n: p.grad for n, p in m64.named_parameters() if p.grad is not None}

    eng = engine_from_oracle(m, dev)
    pg = product_graph(g, dev)
    ego = khop_ego_batch(pg, k)
    print("ego: Ns %d (ref %d)  Es %d (ref %d)  nodes_equal %s" % (
        ego.num_nodes(), e.num_rows, ego.num_edges(), e.num_edges,
        np.array_equal(ego.ego_nodes.cpu().numpy(), e.ego_nodes) and np.array_equal(ego.sub_indices.cpu().numpy(), e.sub_indices)))
    b = DeviceBatch(pg, ego, pg.ndata["x"])
    losses, emb = eng.forward(b, gate_u.to(dev), feat_u.to(dev), want=True)
    eng.backward()
    torch.cuda.synchronize()
    N, Ns = b.N, b.Ns
    rows = []
    def cmp(name, got, ref):
        rows.append((name, rel(got, ref), float(ref.abs().max())))
    cmp("t", eng.debug_buffer("t", (N, 32)), out["t"])
    for l in range(4):
        cmp("y1_%d" % l, eng.debug_buffer("y1_%d" % l, (N, 64)), keep["y1_%d" % l])
    for l in range(4):
        cmp("y2_%d" % l, eng.debug_buffer("y2_%d" % l, (Ns, 64)), keep["y2_%d" % l])
    cmp("H", eng.debug_buffer("H", (N, 64)), out["H"])
    cmp("C", eng.debug_buffer("C", (N, 64)), out["C"])
    cmp("lam", eng.debug_buffer("lam", (N,)), out["lam"])
    cmp("alpha", eng.debug_buffer("alpha", (N,)), out["alpha"])
    cmp("noisy", emb["noisy"], out["noisy"])
    cmp("core", eng.debug_buffer("core", (b.B, 64)), out["core_readout"])
    cmp("readout", emb["graph_readout"], out["graph_readout"])
    cmp("imap", emb["interaction_map"], out["interaction_map"])
    cmp("Z", emb["Z"], out["Z"])
    for i, n in enumerate(("KL", "contrastive", "recon")):
        rows.append(("loss_" + n, abs(float(losses[i]) - float(out[n])) / abs(float(out[n])), float(out[n])))
    cmp("g_core", eng.debug_buffer("g_core", (b.B, 64)), out["core_readout"].grad)
    cmp("g_readout", eng.debug_buffer("g_readout", (b.B, 64)), out["graph_readout"].grad)
    cmp("gZ", eng.debug_buffer("gZ", (N, 64)), out["Z"].grad)
    cmp("gH(total)", eng.debug_buffer("gH", (N, 64)), out["H"].grad)
    cmp("gC", eng.debug_buffer("gC", (N, 64)), out["C"].grad)
    cmp("g_a enc1 layer0", eng.debug_buffer("ga0_1", (N, 32)), a_saved["1_0"].grad)
    cmp("g_a enc2 layer0", eng.debug_buffer("ga0_2", (Ns, 32)), a_saved["2_0"].grad)
    for l in range(4):
        cmp("a1_%d (saved)" % l, eng.debug_buffer("a1_%d" % l, (N, 32 if l == 0 else 64)), a_saved["1_%d" % l])
    seg = tg.seg_ids()
    lastmask = (seg == seg.max())
    gh_got, gh_ref = eng.debug_buffer("gH", (N, 64)).cpu().double(), out["H"].grad
    rows.append(("gH rows of last graph", float((gh_got[lastmask] - gh_ref[lastmask]).abs().max() / gh_ref.abs().max()), float(gh_ref[lastmask].abs().max())))
    rows.append(("gH rows of other graphs", float((gh_got[~lastmask] - gh_ref[~lastmask]).abs().max() / gh_ref.abs().max()), float(gh_ref[~lastmask].abs().max())))
    gv = eng.grad_views()
    gmax = max(float(v.abs().max()) for v in ref_grads.values())
    for n, got in gv.items():
        ref = ref_grads[n].reshape(got.shape)
        if is_zero_grad_param(n):
            rows.append(("grad " + n + " (zero)", float(got.abs().max()) / gmax, 0.0))
        elif n == "attn_layer.weight":
            rows.append(("grad " + n, rel(got[:, 64:], ref[:, 64:]), float(ref.abs().max())))
        else:
            rows.append(("grad " + n, rel(got, ref), float(ref.abs().max())))
    print("%-58s %12s %12s" % ("tensor (B=%d k=%d)" % (B, k), "rel_err", "ref_max"))
    for n, err, mx in rows:
        flag = "" if err <= 1e-5 else ("  <-- " if err > 2e-4 else "  .")
        print("%-58s %12.3e %12.4e%s" % (n, err, mx, flag))


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    main(*a)
