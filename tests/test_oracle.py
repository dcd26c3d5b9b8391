"""Oracle self-consistency: graph semantics, ego-net BFS properties, faithful == vectorised, fp64."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle.graph_ref import (batch_ref, csr_from_edges, ego_batch_ref, graph_from_bonds, khop_ball_ref,
                              path_graph, synth_batch, to_bidirected_ref)
from oracle.scgib_oracle import (OracleMainmodel, draw_noise_like_reference, normalize_rows, sum_nodes,
                                 tgraph_from_ego, tgraph_from_ref)


def test_to_bidirected_sorted_dedup():
    n, s, d = to_bidirected_ref([0, 2, 1, 0], [1, 1, 0, 1])
    assert n == 3
    assert list(zip(s.tolist(), d.tolist())) == [(0, 1), (1, 0), (1, 2), (2, 1)]


def test_graph_num_nodes_is_max_id_plus_one():
    n, s, d = to_bidirected_ref([0], [4])
    assert n == 5
    indptr, indices = csr_from_edges(n, s, d)
    assert indptr.tolist() == [0, 1, 1, 1, 1, 2] and indices.tolist() == [4, 0]


def test_batch_offsets():
    a, b = path_graph(3), path_graph(2, seed=1)
    g = batch_ref([a, b])
    assert g.graph_ptr.tolist() == [0, 3, 5]
    assert g.indptr.tolist() == [0, 1, 3, 4, 5, 6]
    assert g.indices.tolist() == [1, 0, 2, 1, 4, 3]
    assert g.batch_num_nodes().tolist() == [3, 2]
    A = g.dense_adj()
    assert A.sum() == 6 and A[0, 1] == 1 and A[3, 4] == 1 and A[2, 3] == 0 and np.array_equal(A, A.T)


def test_khop_path_and_star():
    p = path_graph(6)
    assert khop_ball_ref(p.indptr, p.indices, 0, 1).tolist() == [0, 1]
    assert khop_ball_ref(p.indptr, p.indices, 2, 1).tolist() == [1, 2, 3]
    assert khop_ball_ref(p.indptr, p.indices, 2, 2).tolist() == [0, 1, 2, 3, 4]
    star = graph_from_bonds(5, [(0, 1), (0, 2), (0, 3), (0, 4)])
    assert khop_ball_ref(star.indptr, star.indices, 3, 1).tolist() == [0, 3]
    assert khop_ball_ref(star.indptr, star.indices, 3, 2).tolist() == [0, 1, 2, 3, 4]
    iso = graph_from_bonds(4, [(0, 3)])  # nodes 1,2 isolated
    assert khop_ball_ref(iso.indptr, iso.indices, 1, 3).tolist() == [1]


def test_ego_batch_triangle():
    t = graph_from_bonds(3, [(0, 1), (1, 2), (0, 2)])
    e = ego_batch_ref(t, 1)
    assert e.ego_ptr.tolist() == [0, 3, 6, 9]
    assert e.ego_nodes.tolist() == [0, 1, 2] * 3
    assert e.sub_indptr.tolist() == list(range(0, 19, 2))
    assert e.sub_indices.tolist() == [1, 2, 0, 2, 0, 1, 4, 5, 3, 5, 3, 4, 7, 8, 6, 8, 6, 7]


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 10_000), st.integers(1, 3))
def test_ego_properties(seed, k):
    g = synth_batch(seed, 3)
    e = ego_batch_ref(g, k)
    N = g.num_nodes
    # brute-force BFS distance
    for v in range(0, N, 5):
        dist = {v: 0}
        fr = [v]
        for h in range(k):
            nx = []
            for u in fr:
                for w in g.indices[g.indptr[u]:g.indptr[u + 1]]:
                    if int(w) not in dist:
                        dist[int(w)] = h + 1
                        nx.append(int(w))
            fr = nx
        ball = e.ego_nodes[e.ego_ptr[v]:e.ego_ptr[v + 1]]
        assert ball.tolist() == sorted(dist)            # ascending, contains v, exactly the k-ball
        assert v in ball
    # induced edges symmetric and inside the same ego-net
    dst = np.repeat(np.arange(e.num_rows), np.diff(e.sub_indptr))
    src = e.sub_indices
    ego_of = np.repeat(np.arange(N), np.diff(e.ego_ptr))
    assert np.array_equal(ego_of[src], ego_of[dst])
    pairs = set(zip(src.tolist(), dst.tolist()))
    assert all((b, a) in pairs for a, b in pairs)
    # every induced edge is a parent edge
    for a, b in list(pairs)[:200]:
        pa, pb = e.ego_nodes[a], e.ego_nodes[b]
        assert pa in g.indices[g.indptr[pb]:g.indptr[pb + 1]]


def _setup(seed, B, k, dtype=torch.float32):
    g = synth_batch(seed, B)
    e = ego_batch_ref(g, k)
    torch.manual_seed(seed)
    m = OracleMainmodel(9).to(dtype)
    x = normalize_rows(torch.from_numpy(g.x)).to(dtype)
    en = torch.from_numpy(e.ego_nodes.astype(np.int64))
    gu, fu = draw_noise_like_reference(g.batch_num_nodes().tolist(), 64, seed + 7)
    return g, e, m, x, en, gu, fu


@pytest.mark.parametrize("k", [1, 2])
def test_faithful_equals_vectorised_fp64(k):
    g, e, m, x, en, gu, fu = _setup(3, 24, k, torch.float64)
    tg, te = tgraph_from_ref(g), tgraph_from_ego(e)
    a = m.forward_faithful(tg, x, te, x[en], gu, fu)
    ga = torch.autograd.grad(a["KL"] + a["recon"] + a["contrastive"], [p for p in m.parameters()], allow_unused=True)
    b = m.forward_vectorised(tg, x, te, en, gu, fu)
    for key in ("KL", "contrastive", "recon"):
        assert abs(float(a[key]) - float(b[key])) <= 1e-11 * abs(float(a[key])), key
    for key in ("interaction_map", "Z", "noisy", "graph_readout", "core_readout"):
        assert float((a[key] - b[key]).abs().max()) <= 1e-11 * float(a[key].abs().max()), key
    gb = torch.autograd.grad(b["KL"] + b["recon"] + b["contrastive"], [p for p in m.parameters()], allow_unused=True)
    for (n, _), u, v in zip(m.named_parameters(), ga, gb):
        if u is None:
            assert v is None or float(v.abs().max()) == 0, n
            continue
        assert float((u - v).abs().max()) <= 1e-9 * float(u.abs().max()) + 1e-10, n  # pre-BN biases: grad == 0 up to rounding


def test_attention_independent_of_core_half():
    """the core half of attn_layer and its bias cancel in the per-graph softmax."""
    g, e, m, x, en, gu, fu = _setup(5, 8, 1, torch.float64)
    tg, te = tgraph_from_ref(g), tgraph_from_ego(e)
    a = m.forward_vectorised(tg, x, te, en, gu, fu)["alpha"]
    with torch.no_grad():
        m.attn_layer.weight[:, :64].normal_()
        m.attn_layer.bias.fill_(3.0)
    b = m.forward_vectorised(tg, x, te, en, gu, fu)["alpha"]
    assert float((a - b).abs().max()) < 1e-12


def test_kl_is_last_graph_only():
    g, e, m, x, en, gu, fu = _setup(6, 5, 1, torch.float64)
    tg, te = tgraph_from_ref(g), tgraph_from_ego(e)
    base = m.forward_faithful(tg, x, te, x[en], gu, fu)["KL"]
    gu2 = gu.clone()
    n_last = int(g.batch_num_nodes()[-1])
    gu2[:-n_last] = 0.5            # perturb the gate noise of every graph but the last
    assert float(m.forward_faithful(tg, x, te, x[en], gu2, fu)["KL"]) == float(base)
    gu3 = gu.clone()
    gu3[-1] = 0.123
    assert float(m.forward_faithful(tg, x, te, x[en], gu3, fu)["KL"]) != float(base)


def test_sum_nodes():
    g = batch_ref([path_graph(3), path_graph(2, seed=1)])
    h = torch.arange(10, dtype=torch.float32).reshape(5, 2)
    assert sum_nodes(tgraph_from_ref(g), h).tolist() == [[6.0, 9.0], [14.0, 16.0]]
