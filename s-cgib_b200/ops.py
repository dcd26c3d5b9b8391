"""Thin wrappers over the individual-operator entry points of the C ABI (unit-test surface and the
building blocks of the drop-in modules).  Every function requires CUDA tensors; nothing falls back to torch."""
from __future__ import annotations

import torch

from . import _lib

HID, DTR = 64, 32


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _cuda(*ts):
    for t in ts:
        if t is not None and t.device.type != "cuda":
            raise RuntimeError("scgib_b200 ops need CUDA tensors (no CPU fallback)")


def input_proj(x: torch.Tensor, Wt: torch.Tensor) -> torch.Tensor:
    """transfer_d(F.normalize(x))  (exp_pretraining.py:312, models.py:668)."""
    _cuda(x, Wt)
    x, Wt = x.contiguous().float(), Wt.contiguous().float()
    t = torch.empty(x.shape[0], DTR, device=x.device)
    _lib.check(_lib.load().scgib_input_proj_fwd_f32(_lib.ptr(x), _lib.ptr(Wt), x.shape[0], x.shape[1], Wt.shape[0],
                                                    _lib.ptr(t), _stream(x)), "input_proj_fwd")
    return t


def gin_layer_fwd(h, indptr, indices, W1, b1, W2, b2, row_map=None, bn_in=None, running=None, save=False):
    """One GINConv + batch statistics (models.py:66-72).  Returns (y, bn{mean,rstd}[2,64], a, r)."""
    _cuda(h, indptr, W1)
    lib = _lib.load()
    V = indptr.numel() - 1
    kin = h.shape[1]
    y = torch.empty(V, HID, device=h.device)
    a = torch.empty(V, kin, device=h.device) if save else None
    r = torch.empty(V, HID, device=h.device) if save else None
    bn_out = torch.zeros(4, HID, device=h.device)
    ws = torch.empty(lib.scgib_gin_workspace_bytes(V), dtype=torch.uint8, device=h.device)
    _lib.check(lib.scgib_gin_layer_fwd_f32(_lib.ptr(h.contiguous()), kin, _lib.ptr(row_map), _lib.ptr(bn_in),
                                           _lib.ptr(indptr), _lib.ptr(indices), V, _lib.ptr(W1.contiguous()),
                                           _lib.ptr(b1), _lib.ptr(W2.contiguous()), _lib.ptr(b2), _lib.ptr(a),
                                           _lib.ptr(r), _lib.ptr(y), _lib.ptr(bn_out), _lib.ptr(running), _lib.ptr(ws),
                                           ws.numel(), _stream(h)), "gin_layer_fwd")
    return y, bn_out[:2], a, r


def segment_sum(h, seg_ptr, bn=None):
    """dgl.sum_nodes (models.py:716, 725, 733, 684); optional fused relu(BN(.)) on load."""
    _cuda(h, seg_ptr)
    S = seg_ptr.numel() - 1
    out = torch.empty(S, HID, device=h.device)
    _lib.check(_lib.load().scgib_segment_sum_f32(_lib.ptr(h.contiguous()), _lib.ptr(seg_ptr), S, _lib.ptr(bn),
                                                 _lib.ptr(out), _stream(h)), "segment_sum")
    return out


def gin_layer_bwd(g_next, y, r, a, bn, W1, W2, indptr=None, indices=None):
    """Backward of one GINConv + BatchNorm(train) + ReLU layer.  ``g_next``: gradient wrt the layer output (``indptr`` None)
    or wrt the next layer's aggregated input (gathered through the CSR).  ``bn`` = [4,64] {mean, rstd, gamma, beta}.
    Returns (g_a, dW1, db1, dW2, db2, dgamma, dbeta)."""
    _cuda(g_next, y, r, a, bn, W1, W2)
    lib = _lib.load()
    V, kin, dev = y.shape[0], a.shape[1], y.device
    g_a = torch.empty(V, kin, device=dev)
    dW1, db1 = torch.empty(HID, kin, device=dev), torch.empty(HID, device=dev)
    dW2, db2 = torch.empty(HID, HID, device=dev), torch.empty(HID, device=dev)
    dgamma, dbeta = torch.empty(HID, device=dev), torch.empty(HID, device=dev)
    ws = torch.empty(lib.scgib_gin_layer_bwd_workspace_bytes(V, kin) + 256, dtype=torch.uint8, device=dev)
    _lib.check(lib.scgib_gin_layer_bwd_f32(_lib.ptr(g_next.contiguous()), _lib.ptr(indptr), _lib.ptr(indices), V, kin,
                                           _lib.ptr(y.contiguous()), _lib.ptr(r.contiguous()), _lib.ptr(a.contiguous()),
                                           _lib.ptr(bn.contiguous()), _lib.ptr(W1.contiguous()), _lib.ptr(W2.contiguous()),
                                           _lib.ptr(g_a), _lib.ptr(dW1), _lib.ptr(db1), _lib.ptr(dW2), _lib.ptr(db2),
                                           _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.ptr(ws), ws.numel(), _stream(y)),
               "gin_layer_bwd")
    return g_a, dW1, db1, dW2, db2, dgamma, dbeta


def recon_adj(Z, indptr, indices, scale=1.0, want_grad=True):
    """loss_recon_adj (models.py:762-768) and scale * d loss / d Z; returns (loss[1], gZ or None)."""
    _cuda(Z, indptr)
    lib = _lib.load()
    Z = Z.contiguous().float()
    N, E = Z.shape[0], indices.numel()
    loss = torch.empty(1, device=Z.device)
    gZ = torch.empty_like(Z) if want_grad else None
    ws = torch.empty(lib.scgib_loss_workspace_bytes(1) + 256, dtype=torch.uint8, device=Z.device)
    _lib.check(lib.scgib_recon_adj_f32(_lib.ptr(Z), _lib.ptr(indptr), _lib.ptr(indices), N, E, float(scale), _lib.ptr(loss),
                                       _lib.ptr(gZ), _lib.ptr(ws), ws.numel(), _stream(Z)), "recon_adj")
    return loss, gZ


def contrastive(core, readout, scale=1.0, want_grad=True):
    """batched_semi_loss (models.py:606-629) of the core / graph readouts and its gradients; returns (loss[1], g_core, g_readout)."""
    _cuda(core, readout)
    lib = _lib.load()
    core, readout = core.contiguous().float(), readout.contiguous().float()
    B = core.shape[0]
    loss = torch.empty(1, device=core.device)
    g1 = torch.empty_like(core) if want_grad else None
    g2 = torch.empty_like(core) if want_grad else None
    ws = torch.empty(lib.scgib_loss_workspace_bytes(B) + 256, dtype=torch.uint8, device=core.device)
    _lib.check(lib.scgib_contrastive_f32(_lib.ptr(core), _lib.ptr(readout), B, float(scale), _lib.ptr(loss), _lib.ptr(g1),
                                         _lib.ptr(g2), _lib.ptr(ws), ws.numel(), _stream(core)), "contrastive")
    return loss, g1, g2
