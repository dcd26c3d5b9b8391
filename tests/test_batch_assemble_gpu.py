"""GPU-side batch assembly from a dataset resident in HBM: bit-exact with the oracle's ``dgl.batch``
restatement (oracle.graph_ref.batch_ref) for any id list (shuffled, repeated ids), and a training step fed by ids."""
import numpy as np
import pytest
import torch

from oracle.graph_ref import batch_ref, synth_molecule

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _dataset(M, seed, shape="pcqm"):
    rng = np.random.default_rng(seed)
    mols = [synth_molecule(rng, shape) for _ in range(M)]
    return mols, batch_ref(mols)


@pytest.mark.parametrize("M,B,shape", [(64, 64, "pcqm"), (200, 37, "pcqm"), (40, 1, "pcqm"), (24, 50, "peptides"),
                                       (3000, 2500, "pcqm")])
def test_assemble_matches_dgl_batch_semantics(M, B, shape):
    from scgib_b200.graph import DeviceDataset
    mols, big = _dataset(M, M + B, shape)
    ds = DeviceDataset(big.graph_ptr, big.indptr, big.indices, big.x, DEV)
    assert len(ds) == M
    rng = np.random.default_rng(B)
    ids = rng.integers(0, M, size=B) if B != M else rng.permutation(M)        # repeats allowed / a full shuffle
    ref = batch_ref([mols[i] for i in ids])
    g = ds.assemble(torch.from_numpy(ids.astype(np.int32)))
    assert np.array_equal(g.graph_ptr.cpu().numpy(), ref.graph_ptr)
    assert np.array_equal(g.indptr.cpu().numpy(), ref.indptr)
    assert np.array_equal(g.indices.cpu().numpy(), ref.indices)
    assert np.array_equal(g.ndata["x"].cpu().numpy(), ref.x)
    # reusable buffers: a second, different batch through the same dict
    out = {}
    for trial in range(2):
        ids2 = rng.integers(0, M, size=max(1, B // 2 + trial))
        ref2 = batch_ref([mols[i] for i in ids2])
        g2 = ds.assemble(torch.from_numpy(ids2.astype(np.int32)).pin_memory(), out=out)
        assert np.array_equal(g2.indices.cpu().numpy(), ref2.indices) and np.array_equal(g2.ndata["x"].cpu().numpy(), ref2.x)


def test_train_step_from_ids_equals_step_from_host_batch():
    """Same molecules, same weights, same noise: the step fed by ids (assembly on the GPU) and the step fed by the
    host-collated batch give bit-identical losses and gradients."""
    from scgib_b200.engine import PretrainEngine
    from scgib_b200.graph import BatchedGraph, DeviceDataset
    mols, big = _dataset(300, 9)
    ds = DeviceDataset(big.graph_ptr, big.indptr, big.indices, big.x, DEV)
    ids = np.random.default_rng(1).permutation(300)[:128]
    ref = batch_ref([mols[i] for i in ids])
    host = BatchedGraph(torch.from_numpy(ref.graph_ptr), torch.from_numpy(ref.indptr), torch.from_numpy(ref.indices),
                        torch.from_numpy(ref.x))
    res = []
    for mode in ("ids", "host"):
        eng = PretrainEngine(9, gin_layers=4, device=DEV, seed=3)
        if mode == "ids":
            b = eng.wait_batch(eng.prefetch_ids(ds, torch.from_numpy(ids.astype(np.int32)).pin_memory(), 1))
        else:
            b = eng.wait_batch(eng.prefetch_batch(host.pin_memory(), 1))
        gen = torch.Generator().manual_seed(5)
        gu, fu = torch.rand(b.N, generator=gen).to(DEV), torch.rand(b.N, 64, generator=gen).to(DEV)
        losses = eng.forward(b, gu, fu).clone()
        grads = eng.backward().clone()
        res.append((losses, grads))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


def test_batch_validate_names_the_violation():
    """scgib_batch_validate: well-formed batches pass; each malformed one is rejected with its own message."""
    from scgib_b200.graph import BatchedGraph
    mols, big = _dataset(12, 3)
    def make(gp=None, ip=None, idx=None):
        return BatchedGraph(torch.from_numpy(big.graph_ptr if gp is None else gp), torch.from_numpy(big.indptr if ip is None else ip),
                            torch.from_numpy(big.indices if idx is None else idx), torch.from_numpy(big.x)).to(DEV)
    make().validate()
    gp = big.graph_ptr.copy(); gp[1] = gp[0] + 1                      # first graph: one node
    with pytest.raises(ValueError, match="fewer than 2 nodes"):
        make(gp=gp).validate()
    idx = big.indices.copy(); idx[0] = big.num_nodes - 1              # first node points into the last graph
    with pytest.raises(ValueError, match="leaves its graph"):
        make(idx=idx).validate()
    idx = big.indices.copy()
    v = int(np.argmax(np.diff(big.indptr) >= 2)); e = int(big.indptr[v]); idx[e + 1] = idx[e]     # duplicate neighbour
    with pytest.raises(ValueError, match="strictly ascending"):
        make(idx=idx).validate()
    ip = big.indptr.copy(); ip[5] = ip[6] + 1
    with pytest.raises(ValueError, match="monotone|leaves|ascending"):
        make(ip=ip).validate()
