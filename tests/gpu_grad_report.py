"""Per-parameter gradient errors of the CUDA path against the fp64 oracle (experiments): python -m tests.gpu_grad_report [hidden] [dtype] [B] [k]"""
import sys

import numpy as np
import torch

from oracle.graph_ref import ego_batch_ref, synth_batch
from oracle.scgib_oracle import OracleMainmodel, draw_noise_like_reference
from tests.helpers import engine_from_oracle, fp64_truth, product_graph, rel

H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dtype = sys.argv[2] if len(sys.argv) > 2 else "fp32"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 100
k = int(sys.argv[4]) if len(sys.argv) > 4 else 1
DEV = "cuda:0"
from scgib_b200.engine import DeviceBatch
from scgib_b200.graph import khop_ego_batch
g = synth_batch(51, B)
e = ego_batch_ref(g, k)
torch.manual_seed(51)
m = OracleMainmodel(9, H)
gate_u, feat_u = draw_noise_like_reference(g.batch_num_nodes().tolist(), H, 151)
truth_out, truth_grads = fp64_truth(m, g, e, gate_u, feat_u)
eng = engine_from_oracle(m, DEV, dtype=dtype)
pg = product_graph(g, DEV)
b = DeviceBatch(pg, khop_ego_batch(pg, k), pg.ndata["x"], normalize_x=True)
losses, emb = eng.forward(b, gate_u.to(DEV), feat_u.to(DEV), want=True)
eng.backward()
torch.cuda.synchronize()
for n in ("interaction_map", "Z", "noisy", "graph_readout"):
    print("%-20s %.2e" % (n, rel(emb[n], truth_out[n])))
for i, n in enumerate(("KL", "contrastive", "recon")):
    print("%-20s %.2e" % (n, abs(float(losses[i]) - float(truth_out[n])) / abs(float(truth_out[n]))))
for n, got in eng.grad_views().items():
    t = truth_grads[n].reshape(got.shape)
    print("%-55s %.2e   |truth| %.2e" % (n, rel(got, t), float(t.abs().max())))
