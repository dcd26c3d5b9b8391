"""Per-tile role timeline of gin_fwd_tc4 (experiments): SCGIB_FWD4=1 SCGIB_DBG=1024 python -m tests.gpu_tc4_trace"""
import ctypes
import numpy as np
import torch
from scgib_b200 import _lib
from scgib_b200.engine import PretrainEngine
from scgib_b200.synth import synth_batch

lib = _lib.load()
dev = torch.device("cuda:0")
eng = PretrainEngine(9, gin_layers=4, device=dev, seed=0)
g = synth_batch(1, 4096).to(dev)
b = eng.make_batch(g, 1)
for _ in range(3):
    eng.forward(b)
torch.cuda.synchronize()
n = 160 * 16 * 16
buf = (ctypes.c_longlong * n)()
lib.scgib_debug_tc4_trace(ctypes.cast(buf, ctypes.c_void_p), n)
t = np.frombuffer(buf, dtype=np.int64).reshape(160, 16, 16).astype(np.float64)
names = ["c.start", "c.hfree", "c.full", "a.free", "a.full", "m.agg", "m.g1", "m.g2", "e.agg", "e.a", "e.d1", "e.r", "e.d2", "e.end"]
for cta in (0, 40, 100):
    t0 = t[cta, 0, 0]
    print("CTA", cta)
    for i in range(15):
        if t[cta, i, 0] == 0 and i > 0:
            break
        print("  tile %2d: " % i + " ".join("%s=%6.2f" % (names[e], (t[cta, i, e] - t0) / 1965.0) for e in range(14)))
