"""Host-side graph plumbing of the product (no GPU): ``graph`` (= dgl.graph + to_bidirected, util.py:277-325),
``batch`` (= dgl.batch) and the packed shard format against the oracle's restatements."""
import numpy as np
import torch

from oracle.graph_ref import batch_ref, csr_from_edges, synth_molecule, to_bidirected_ref


def _random_edges(rng, n, m):
    src = rng.integers(0, n, size=m)
    dst = rng.integers(0, n, size=m)
    keep = src != dst
    return src[keep], dst[keep]


def test_graph_matches_to_bidirected_semantics():
    from scgib_b200.graph import graph
    rng = np.random.default_rng(0)
    for n, m in ((5, 4), (20, 30), (50, 200), (3, 0)):
        src, dst = _random_edges(rng, n, m)                       # one direction only, with duplicates
        g = graph((src, dst), num_nodes=n)
        nn_, s, d = to_bidirected_ref(src, dst, n)
        indptr, indices = csr_from_edges(n, s, d)
        assert g.num_nodes() == n
        assert np.array_equal(g.indptr.numpy(), indptr) and np.array_equal(g.indices.numpy(), indices)
    g = graph((np.array([0, 1]), np.array([1, 6])))              # dgl.graph: num_nodes = max id + 1
    assert g.num_nodes() == 7


def test_batch_and_shard_roundtrip(tmp_path):
    from scgib_b200.graph import batch, graph, load_shard, pack_shard
    rng = np.random.default_rng(1)
    mols = [synth_molecule(rng) for _ in range(9)]
    triples = []
    for mo in mols:
        s, d = mo.edges()
        keep = s < d                                              # one direction per bond, as a PyG edge list may hold
        triples.append((np.stack([s[keep], d[keep]]), mo.x, rng.integers(0, 2, size=3)))
    ref = batch_ref(mols)
    shard = pack_shard(triples, str(tmp_path / "toy_csr.pt"))
    for key, want in (("graph_ptr", ref.graph_ptr), ("indptr", ref.indptr), ("indices", ref.indices), ("x", ref.x)):
        assert np.array_equal(shard[key].numpy(), want), key
    assert shard["y"].shape == (9, 3)
    big, y = load_shard(str(tmp_path / "toy_csr.pt"))
    assert np.array_equal(big.indices.numpy(), ref.indices) and torch.equal(y, shard["y"])
    # dgl.batch semantics on the product side: offsets in list order
    gs = [graph((t[0][0], t[0][1]), num_nodes=t[1].shape[0], x=torch.from_numpy(t[1])) for t in triples]
    b2 = batch(gs[::-1])
    ref2 = batch_ref(mols[::-1])
    assert np.array_equal(b2.indices.numpy(), ref2.indices) and np.array_equal(b2.graph_ptr.numpy(), ref2.graph_ptr)
    assert b2.batch_num_nodes().tolist() == ref2.batch_num_nodes().tolist()
