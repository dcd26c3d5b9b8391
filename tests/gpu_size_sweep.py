"""Robustness sweep (experiments; run on the GPU box): odd and extreme batch sizes through the whole step; losses finite,
gradients finite, two runs bit-identical.  python tests/gpu_size_sweep.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scgib_b200.engine import PretrainEngine
from scgib_b200.synth import synth_batch

dev = torch.device("cuda:0")
for B, k, shape in [(1, 1, "pcqm"), (2, 1, "pcqm"), (3, 2, "pcqm"), (100, 1, "pcqm"), (1000, 3, "pcqm"), (4097, 1, "pcqm"),
                    (16384, 1, "pcqm"), (7, 1, "peptides"), (300, 2, "peptides"), (129, 4, "pcqm")]:
    eng = PretrainEngine(9, gin_layers=4, device=dev, seed=1)
    g = synth_batch(B, B, shape).to(dev)
    b = eng.make_batch(g, k)
    gen = torch.Generator(device=dev).manual_seed(3)
    gu, fu = torch.rand(b.N, device=dev, generator=gen), torch.rand(b.N, 64, device=dev, generator=gen)
    res = []
    for _ in range(2):
        losses = eng.forward(b, gu, fu).clone()
        grads = eng.backward().clone()
        res.append((losses, grads))
    torch.cuda.synchronize()
    ok = torch.isfinite(res[0][0]).all() and torch.isfinite(res[0][1]).all()
    same = torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    print("B %6d k %d %-8s N %7d Ns %8d  losses %s  finite %s  deterministic %s" %
          (B, k, shape, b.N, b.Ns, [round(v, 4) for v in res[0][0].tolist()], bool(ok), same), flush=True)
    assert ok and same
    del eng
    torch.cuda.empty_cache()
print("ok")
