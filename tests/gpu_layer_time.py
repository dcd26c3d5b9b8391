"""Time one GIN layer forward (ego-batch size) in isolation: python -m tests.gpu_layer_time [mode]"""
import sys
import torch
from oracle.graph_ref import synth_batch
from scgib_b200 import _lib, ops
from scgib_b200.graph import BatchedGraph
from tests.helpers import product_graph

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 3
lib = _lib.load()
dev = "cuda:0"
g = synth_batch(5, 12000)           # ~180 k rows
pg = product_graph(g, dev)
V = pg.indptr.numel() - 1
torch.manual_seed(0)
for kin in (64, 32):
    h = torch.randn(V, kin, device=dev)
    W1 = torch.randn(64, kin, device=dev) * 0.1; b1 = torch.randn(64, device=dev) * 0.1
    W2 = torch.randn(64, 64, device=dev) * 0.1; b2 = torch.randn(64, device=dev) * 0.1
    lib.scgib_set_tensor_cores(mode)
    for _ in range(3):
        ops.gin_layer_fwd(h, pg.indptr, pg.indices, W1, b1, W2, b2, save=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gin_layer_fwd(h, pg.indptr, pg.indices, W1, b1, W2, b2, save=True)
    e1.record()
    torch.cuda.synchronize()
    print("mode %d kin %d V %d E %d: %.1f us per layer call (incl. transposes/memset)" % (mode, kin, V, pg.indices.numel(), e0.elapsed_time(e1) * 100))
