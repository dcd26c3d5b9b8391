"""The op-level C-ABI entries of SURVEY 8(b) (core gate, core-candidate attention, head MLP, sum_nodes backward) against torch
autograd of the oracle's own functions (reference call sites models.py:595-604, 631-660, 738-748, 676), hidden 64 and 128."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.graph_ref import synth_batch
from oracle.scgib_oracle import OracleMainmodel, draw_noise_like_reference, sum_nodes, tgraph_from_ref
from tests.helpers import rel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("H", [64, 128])
def test_core_gate_op(H):
    from scgib_b200 import ops
    g = synth_batch(71, 40)
    tg = tgraph_from_ref(g)
    nodes = g.batch_num_nodes().tolist()
    torch.manual_seed(H)
    m = OracleMainmodel(9, H).double()
    Hf = torch.randn(g.num_nodes, H, dtype=torch.float64).requires_grad_()      # signed: GraphSAGE / GCN encoders end without a ReLU
    gate_u, feat_u = draw_noise_like_reference(nodes, H, 5)
    noisy, _, KL_tensor = m.compression(Hf, nodes, gate_u.double(), feat_u.double())
    readout, core, kl = sum_nodes(tg, Hf), sum_nodes(tg, noisy), KL_tensor.mean()
    gn, gc, gr = torch.randn_like(noisy), torch.randn_like(core), torch.randn_like(readout)
    ((noisy * gn).sum() + (core * gc).sum() + (readout * gr).sum() + 0.7 * kl).backward()
    c = m.compressor
    f = lambda t: t.detach().float().to(DEV)
    op = ops.CoreGate(H, f(c[0].weight), f(c[0].bias), f(c[1].weight), f(c[1].bias), f(c[3].weight), f(c[3].bias))
    gp = torch.from_numpy(g.graph_ptr.astype(np.int32)).to(DEV)
    out = op.forward(f(Hf), gp, gate_u.to(DEV), feat_u.to(DEV))
    for got, ref in zip(out, (noisy, None, readout, core, kl.reshape(1))):
        if ref is not None:
            assert rel(got, ref) <= 1e-5
    grads = op.backward(f(gn), f(gc), f(gr), 0.7)
    torch.cuda.synchronize()
    refs = (Hf.grad, c[0].weight.grad, c[0].bias.grad, c[1].weight.grad, c[1].bias.grad, c[3].weight.grad, c[3].bias.grad)
    gmax = max(float(r.abs().max()) for r in refs)
    for i, (got, ref) in enumerate(zip(grads, refs)):
        if float(ref.abs().max()) <= 1e-9 * gmax:              # compressor.0.bias sits in front of a BatchNorm: zero gradient
            assert float(got.abs().max()) <= 1e-4 * gmax
            continue
        assert rel(got.reshape(ref.shape), ref) <= 2e-4, (i, rel(got.reshape(ref.shape), ref))


@pytest.mark.parametrize("H", [64, 128])
def test_core_cand_attention_op(H):
    from scgib_b200 import ops
    g = synth_batch(72, 33)
    seg = tgraph_from_ref(g).seg_ids()
    torch.manual_seed(H)
    C = torch.randn(g.num_nodes, H, dtype=torch.float64).requires_grad_()
    w = torch.randn(H, dtype=torch.float64, requires_grad=True)
    logit = C @ w
    alpha = torch.cat([F.softmax(logit[seg == b], 0) for b in range(g.num_graphs)])
    T = C * alpha[:, None]
    gT = torch.randn_like(T)
    (T * gT).sum().backward()
    gp = torch.from_numpy(g.graph_ptr.astype(np.int32)).to(DEV)
    f = lambda t: t.detach().float().to(DEV)
    a, Tg = ops.core_cand_attn_fwd(f(C), gp, f(w))
    assert rel(a, alpha) <= 1e-5 and rel(Tg, T) <= 1e-5
    gC, dw = ops.core_cand_attn_bwd(f(C), a, f(gT), gp, f(w))
    torch.cuda.synchronize()
    assert rel(gC, C.grad) <= 2e-5 and rel(dw, w.grad) <= 2e-5


@pytest.mark.parametrize("H", [64, 128])
def test_head_mlp_op(H):
    from scgib_b200 import ops
    torch.manual_seed(H)
    N = 777
    mlp = torch.nn.Sequential(torch.nn.Linear(2 * H, H), torch.nn.ReLU(), torch.nn.Linear(H, H)).double()
    noisy = torch.randn(N, H, dtype=torch.float64).requires_grad_()
    C = torch.randn(N, H, dtype=torch.float64)
    alpha = torch.rand(N, dtype=torch.float64)
    aC = (C * alpha[:, None]).requires_grad_()
    imap = torch.cat((noisy, aC), -1)
    Z = mlp(imap)
    gZ = torch.randn_like(Z)
    (Z * gZ).sum().backward()
    f = lambda t: t.detach().float().to(DEV)
    op = ops.HeadMLP(H, f(mlp[0].weight), f(mlp[0].bias), f(mlp[2].weight), f(mlp[2].bias))
    Zg, ig = op.forward(f(noisy), f(C), f(alpha))
    assert rel(Zg, Z) <= 1e-5 and rel(ig, imap) <= 1e-6
    gI, dW1, db1, dW2, db2 = op.backward(f(gZ))
    torch.cuda.synchronize()
    assert rel(gI[0], noisy.grad) <= 2e-5 and rel(gI[1], aC.grad) <= 2e-5
    for got, ref in ((dW1, mlp[0].weight.grad), (db1, mlp[0].bias.grad), (dW2, mlp[2].weight.grad), (db2, mlp[2].bias.grad)):
        assert rel(got, ref) <= 2e-5


def test_segment_sum_bwd_op():
    from scgib_b200 import ops
    ptr = torch.tensor([0, 3, 3, 50, 64], dtype=torch.int32)
    g = torch.randn(4, 128)
    ref = torch.repeat_interleave(g, (ptr[1:] - ptr[:-1]).long(), 0)
    assert torch.equal(ops.segment_sum_bwd(g.to(DEV), ptr.to(DEV), 64).cpu(), ref)
