"""Where the time of the drop-in training loop goes (experiments; run on the GPU box):
    python tests/gpu_dropin_profile.py
Phases of exp_pretraining.train_epoch_pre_training's loop body on models.Mainmodel, each closed by a synchronize."""
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import models
from scgib_b200.graph import khop_ego_batch
from scgib_b200.synth import synth_batch

dev = torch.device("cuda:0")
B, k = 4096, 1
host = [synth_batch(i, B).pin_memory() for i in range(3)]
ns = types.SimpleNamespace(recons_type="adj", useAtt=1, readout_f="sum", d_transfer=32, device=str(dev), batch_size=B, k_transition=k)
m = models.Mainmodel(ns, 9, 64, 4, 4, k, "GIN").to(dev).train()
from exp_pretraining import make_optimizer
opt = make_optimizer(m, 1e-4)
acc = {}


def mark(name, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t1 - t0)
    return t1


steps = 30
for it in range(steps + 5):
    if it == 5:
        acc.clear()
    torch.cuda.synchronize()
    t = time.perf_counter()
    bg = host[it % 3].to(dev, non_blocking=True)
    bx = bg.ndata["x"].float()
    t = mark("h2d", t)
    opt.zero_grad()
    ego = khop_ego_batch(bg, k)
    t = mark("ego", t)
    bx = F.normalize(bx)
    _, kl, con, rec = m.forward(bg, bx, ego, None, None, 1, None, 2, dev, B)
    t = mark("forward", t)
    loss = kl + rec + con
    loss.backward()
    t = mark("backward", t)
    opt.step()
    t = mark("adam", t)
    v = loss.detach().item()
    t = mark("item", t)
print({k_: round(v_ / steps * 1e3, 3) for k_, v_ in acc.items()}, "ms per step; total", round(sum(acc.values()) / steps * 1e3, 3))
# the same loop without the per-phase synchronisations
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(steps):
    bg = host[it % 3].to(dev, non_blocking=True)
    bx = bg.ndata["x"].float()
    opt.zero_grad()
    ego = khop_ego_batch(bg, k)
    bx = F.normalize(bx)
    _, kl, con, rec = m.forward(bg, bx, ego, None, None, 1, None, 2, dev, B)
    loss = kl + rec + con
    loss.backward()
    opt.step()
    v = loss.detach().item()
torch.cuda.synchronize()
print("unsynchronised loop: %.3f ms per step" % ((time.perf_counter() - t0) / steps * 1e3))
