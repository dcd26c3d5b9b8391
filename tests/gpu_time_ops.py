"""Time the stand-alone loss operators (experiments): python -m tests.gpu_time_ops"""
import torch
from scgib_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = 4096
core, readout = torch.randn(B, 64, device=dev), torch.randn(B, 64, device=dev)
for want in (False, True):
    for _ in range(5):
        ops.contrastive(core, readout, 1.0, want_grad=want)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ops.contrastive(core, readout, 1.0, want_grad=want)
    e1.record()
    torch.cuda.synchronize()
    print("contrastive B=%d want_grad=%s: %.1f us per call" % (B, want, e0.elapsed_time(e1) * 1000 / 50))
