"""Training sanity (experiments; run on the GPU box): 1500 steps at B = 4096 on a resident dataset of 64k synthetic molecules,
lr 1e-3 - the losses must stay finite and the total must go down."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scgib_b200.engine import PretrainEngine
from scgib_b200.graph import DeviceDataset, batch
from scgib_b200.synth import synth_batch

dev = torch.device("cuda:0")
eng = PretrainEngine(9, gin_layers=4, device=dev, seed=0)
ds = DeviceDataset.from_batched(batch([synth_batch(i, 4096) for i in range(16)]), dev)
gen = torch.Generator().manual_seed(0)
hist = []
handle = eng.prefetch_ids(ds, torch.randperm(len(ds), generator=gen)[:4096].to(torch.int32).pin_memory(), 1)
for step in range(1500):
    b = eng.wait_batch(handle)
    losses = eng.train_step(b, lr=1e-3)
    handle = eng.prefetch_ids(ds, torch.randperm(len(ds), generator=gen)[:4096].to(torch.int32).pin_memory(), 1)
    if step % 100 == 0 or step == 1499:
        l = losses.cpu().tolist()
        hist.append(l)
        print("step %4d  KL %.4f  contrastive %.4f  recon %.4f  total %.4f" % (step, *l), flush=True)
        assert all(v == v and abs(v) < 1e12 for v in l), "non-finite loss"
assert hist[-1][3] < 0.5 * hist[0][3], "the loss did not go down"
assert torch.isfinite(eng.params).all()
print("ok")
