"""Host logic of the data-parallel path on CPU: world_size 2, gloo.  Each rank computes the (oracle) gradients of its
own shard; after scgib_b200.dist.allreduce_grads_ + grad_scale every rank holds the mean of the per-shard gradients
and an Adam step keeps the replicas identical."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_grads(lo, hi):
    from oracle.graph_ref import batch_ref, ego_batch_ref, synth_molecule
    from oracle.scgib_oracle import OracleMainmodel, draw_noise_like_reference, normalize_rows, tgraph_from_ego, tgraph_from_ref
    from scgib_b200.engine import param_names
    rng = np.random.default_rng(5)
    mols = [synth_molecule(rng) for _ in range(12)]
    g = batch_ref(mols[lo:hi])
    e = ego_batch_ref(g, 1)
    torch.manual_seed(0)
    m = OracleMainmodel(9)
    x = normalize_rows(torch.from_numpy(g.x))
    en = torch.from_numpy(e.ego_nodes.astype(np.int64))
    gu, fu = draw_noise_like_reference(g.batch_num_nodes().tolist(), 64, 100 + lo)
    out = m.forward_vectorised(tgraph_from_ref(g), x, tgraph_from_ego(e), en, gu, fu)
    (out["KL"] + out["recon"] + out["contrastive"]).backward()
    named = dict(m.named_parameters())
    flat = torch.cat([(named[n].grad if named[n].grad is not None else torch.zeros_like(named[n])).reshape(-1)
                      for n in param_names(4)])
    params = torch.cat([named[n].detach().reshape(-1) for n in param_names(4)])
    return flat, params


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from scgib_b200 import dist as sdist
    r, lr, w = sdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    lo, hi = sdist.shard_range(12, rank, world)
    grads, params = _shard_grads(lo, hi)
    scale = sdist.allreduce_grads_(grads, world)
    p = torch.nn.Parameter(params.clone())
    p.grad = grads * scale
    torch.optim.Adam([p], lr=1e-4, weight_decay=5e-5).step()
    sdist.broadcast_params_(params, 0)
    q.put((rank, (grads * scale).numpy(), p.detach().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_batch():
    from scgib_b200.dist import shard_range
    for n, w in ((12, 2), (13, 4), (8192 * 8, 8), (5, 8)):
        r = [shard_range(n, i, w) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_allreduce_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    nthreads = torch.get_num_threads()
    torch.set_num_threads(1)              # same summation order as the workers (ReLU masks are rounding-sensitive)
    try:
        g0, _ = _shard_grads(0, 6)
        g1, _ = _shard_grads(6, 12)
    finally:
        torch.set_num_threads(nthreads)
    mean = ((g0 + g1) / 2).numpy()
    for rank, g, p in res:
        # workers run torch single-threaded, the check multi-threaded: fp32 summation order differs
        assert np.allclose(g, mean, rtol=1e-4, atol=1e-5 * np.abs(mean).max())
    assert np.array_equal(res[0][2], res[1][2])        # replicas stay bit-identical after the optimiser step


def _worker_adam(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from scgib_b200 import dist as sdist
    from scgib_b200.optim import AllReduceAdam
    sdist.init_from_env("gloo")
    torch.manual_seed(3)
    ps = [torch.nn.Parameter(torch.randn(5, 7)), torch.nn.Parameter(torch.randn(11)), torch.nn.Parameter(torch.randn(3))]
    gen = torch.Generator().manual_seed(10 + rank)
    ps[0].grad, ps[1].grad = torch.randn(5, 7, generator=gen), torch.randn(11, generator=gen)      # ps[2]: no gradient
    opt = AllReduceAdam(ps, lr=1e-2, weight_decay=5e-5)
    opt.step()
    q.put((rank, [p.detach().numpy() for p in ps], [ps[0].grad.numpy(), ps[1].grad.numpy()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_allreduce_adam_keeps_module_parameter_replicas_identical():
    """the optimiser of the modules without a flat engine buffer (--encoder GraphSAGE / GCN) under torchrun"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_adam, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    gens = [torch.Generator().manual_seed(10 + r) for r in range(world)]
    g0 = [torch.randn(5, 7, generator=g) for g in gens]
    g1 = [torch.randn(11, generator=g) for g in gens]
    mean0, mean1 = ((g0[0] + g0[1]) / 2).numpy(), ((g1[0] + g1[1]) / 2).numpy()
    for rank, params, grads in res:
        assert np.allclose(grads[0], mean0, rtol=1e-6, atol=1e-7) and np.allclose(grads[1], mean1, rtol=1e-6, atol=1e-7)
    for a, b in zip(res[0][1], res[1][1]):
        assert np.array_equal(a, b)
