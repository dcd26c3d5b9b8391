"""Probe of the tcgen05 primitives on the GPU box: python -m tests.gpu_umma_probe"""
import torch
from scgib_b200 import _lib

lib = _lib.load()
dev = "cuda:0"
torch.manual_seed(0)
for M in (128, 64):
    for mode in (0, 1, 2):
        A = torch.randn(M, 64, device=dev)
        B = torch.randn(M if mode % 10 == 2 else 64, 64, device=dev)
        out = torch.full((128, 64), float("nan"), device=dev)
        rc = lib.scgib_debug_umma(_lib.ptr(A), _lib.ptr(B), _lib.ptr(out), M, mode, None)
        torch.cuda.synchronize()
        Ad, Bd = A.double(), B.double()
        ref = (Ad @ Bd.t()) if mode % 10 == 0 else (Ad @ Bd) if mode % 10 == 1 else (Ad.t() @ Bd)
        R = ref.shape[0]
        o = out.double()
        # find, for every result row, the TMEM lane that holds it
        lanes = []
        for r in range(R):
            d = (o - ref[r][None, :]).abs().max(1).values
            d = torch.nan_to_num(d, nan=1e30)
            lanes.append(int(d.argmin()))
        got = o[lanes]
        err = float((got - ref).abs().max() / ref.abs().max())
        ident = lanes == list(range(R))
        print("nan %d nonzero %d out[0,:4]=%s ref[0,:4]=%s" % (int(torch.isnan(out).sum()), int((out != 0).sum()), out[0, :4].tolist(), ref[0, :4].tolist()))
        print("M=%d mode=%d rc=%d  rel_err=%.3e  identity_lane_map=%s  lanes[:20]=%s lanes[-4:]=%s" % (
            M, mode, rc, err, ident, lanes[:20], lanes[-4:]))
