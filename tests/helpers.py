"""Shared helpers for the GPU parity tests: build product-side batches/engines from oracle-side objects."""
import numpy as np
import torch

from oracle.graph_ref import RefEgoBatch, RefGraph, ego_batch_ref
from oracle.scgib_oracle import OracleMainmodel, normalize_rows, tgraph_from_ego, tgraph_from_ref


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def is_zero_grad_param(n):
    """Gradients that are mathematically zero (bias in front of a BatchNorm; softmax shift)."""
    return n.endswith("apply_func.mlp.2.bias") or n in ("attn_layer.bias", "compressor.0.bias")


def product_graph(g: RefGraph, device):
    from scgib_b200.graph import BatchedGraph
    return BatchedGraph(torch.from_numpy(g.graph_ptr), torch.from_numpy(g.indptr), torch.from_numpy(g.indices),
                        torch.from_numpy(g.x)).to(device)


def product_ego_from_ref(pg, e: RefEgoBatch, k, device):
    from scgib_b200.graph import EgoBatch
    seed = np.repeat(np.arange(len(e.ego_ptr) - 1, dtype=np.int32), np.diff(e.ego_ptr))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.int32))).to(device)
    return EgoBatch(pg, k, t(e.ego_ptr), t(e.ego_nodes), t(seed), t(e.sub_indptr), t(e.sub_indices))


def engine_from_oracle(m: OracleMainmodel, device, gin_layers=4, in_dim=9):
    from scgib_b200.engine import PretrainEngine
    eng = PretrainEngine(in_dim, gin_layers=gin_layers, device=device)
    eng.load_state_dict({k: v.detach().float().to(device) for k, v in m.state_dict().items()}, strict=False)
    return eng


def oracle_grads(m, out):
    (out["KL"] + out["recon"] + out["contrastive"]).backward()
    return {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}


def compare_grads(eng, ref_grads, tol, report=None):
    """Compare the engine's flat gradient buffer with reference-named gradients."""
    gv = eng.grad_views()
    gmax = max(float(v.abs().max()) for v in ref_grads.values())
    worst = ("", 0.0)
    for n, got in gv.items():
        ref = ref_grads[n].to(got.device).reshape(got.shape)
        if is_zero_grad_param(n):
            err = float(got.abs().max()) / gmax
            bound = 1e-5
        elif n == "attn_layer.weight":
            assert float(got[:, :64].abs().max()) == 0.0
            err = rel(got[:, 64:], ref[:, 64:]); bound = tol
        else:
            err = rel(got, ref); bound = tol
        if report is not None:
            report.append((n, err, bound))
        if err / bound > worst[1]:
            worst = (n, err / bound)
        assert err <= bound, (n, err, bound)
    return worst
