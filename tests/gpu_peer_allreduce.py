"""2+ GPU check of the fused peer-memory all-reduce + Adam (run under torchrun on a multi-GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/gpu_peer_allreduce.py
Every rank: random gradients for several steps; the fused kernel must (a) match NCCL all-reduce + the Adam kernel to
fp32 rounding and (b) leave the replicas bit-identical."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scgib_b200 import _lib  # noqa: E402
from scgib_b200.dist import PeerAllreduce  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
n = 80680
gen = torch.Generator(device=dev).manual_seed(0)
p0 = torch.randn(n, device=dev, generator=gen)
pa, ma, va = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)      # fused path
pb, mb, vb = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)      # NCCL + Adam kernel
peer = PeerAllreduce(lib, n, dev)
gr = torch.Generator(device=dev).manual_seed(100 + rank)
st = torch.cuda.current_stream().cuda_stream
for step in range(1, 8):
    g = torch.randn(n, device=dev, generator=gr) * (1.0 + rank)
    peer.grads(step).copy_(g)
    peer.step(pa, ma, va, step, step, 1e-3, (0.9, 0.999), 1e-8, 5e-5, st)
    gs = g.clone()
    dist.all_reduce(gs)
    _lib.check(lib.scgib_adam_step_f32(_lib.ptr(pb), _lib.ptr(gs), _lib.ptr(mb), _lib.ptr(vb), n, step, 1e-3, 0.9, 0.999, 1e-8,
                                       5e-5, 1.0 / world, st))
torch.cuda.synchronize()
err = float((pa - pb).abs().max() / pb.abs().max())
gathered = [torch.empty_like(pa) for _ in range(world)]
dist.all_gather(gathered, pa)
identical = all(bool(torch.equal(gathered[0], t)) for t in gathered)
if rank == 0:
    print("fused vs NCCL+Adam: max rel diff %.3e; replicas bit-identical: %s" % (err, identical), flush=True)
assert err < 1e-5 and identical
# timing of the two exchanges
for name, fn in (("fused", lambda s: peer.step(pa, ma, va, s, s, 1e-3, (0.9, 0.999), 1e-8, 5e-5, st)),
                 ("nccl+adam", None)):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(100):
        s = 8 + k if name == "fused" else 0
        if name == "fused":
            fn(s)
        else:
            dist.all_reduce(gs)
            lib.scgib_adam_step_f32(_lib.ptr(pb), _lib.ptr(gs), _lib.ptr(mb), _lib.ptr(vb), n, 8 + k, 1e-3, 0.9, 0.999, 1e-8, 5e-5,
                                    1.0 / world, st)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print("%s: %.1f us per exchange+update" % (name, e0.elapsed_time(e1) * 10), flush=True)
    if name == "fused":
        peer.seq = 107
dist.barrier()
dist.destroy_process_group()
