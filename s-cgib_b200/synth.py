"""Synthetic molecule batches of the shapes BASELINE.json names (datasets are not available offline).

PCQM4Mv2-shape: n ~ U[8,22] atoms, random tree (attach within a window of 6, degree <= 4) plus 2 ring
closures that create no triangle => ~n+1 bonds; 9 integer-valued features cast to fp32.
Peptides-shape: n ~ U[120,180], chain-biased tree + 3 ring closures.
"""
from __future__ import annotations

import numpy as np
import torch

from .graph import BatchedGraph


def _bonds(rng, n, window, p_chain, n_rings, max_deg=4):
    deg = np.zeros(n, dtype=np.int64)
    adj = [set() for _ in range(n)]
    bonds = []

    def add(a, b):
        bonds.append((a, b)); adj[a].add(b); adj[b].add(a); deg[a] += 1; deg[b] += 1

    for v in range(1, n):
        if p_chain > 0 and deg[v - 1] < max_deg and rng.random() < p_chain:
            add(v - 1, v)
            continue
        cand = [u for u in range(max(0, v - window), v) if deg[u] < max_deg] or \
               [u for u in range(v) if deg[u] < max_deg]
        add(cand[int(rng.integers(len(cand)))], v)
    for _ in range(n_rings):
        for _try in range(32):
            a = int(rng.integers(n))
            b = int(rng.integers(max(0, a - window), min(n - 1, a + window) + 1))
            if a == b or b in adj[a] or deg[a] >= max_deg or deg[b] >= max_deg or (adj[a] & adj[b]):
                continue
            add(a, b)
            break
    return np.asarray(bonds, dtype=np.int64), deg


def synth_arrays(seed: int, num_graphs: int, shape: str = "pcqm"):
    """Returns (graph_ptr, indptr, indices, x) numpy arrays of one batch."""
    rng = np.random.default_rng(seed)
    gp, ip, idx, xs = [0], [np.zeros(1, dtype=np.int64)], [], []
    noff = eoff = 0
    for _ in range(num_graphs):
        if shape == "pcqm":
            n = int(rng.integers(8, 23)); b, deg = _bonds(rng, n, 6, 0.0, 2)
        elif shape == "peptides":
            n = int(rng.integers(120, 181)); b, deg = _bonds(rng, n, 6, 0.5, 3)
        else:
            raise ValueError(shape)
        s = np.concatenate([b[:, 0], b[:, 1]]); d = np.concatenate([b[:, 1], b[:, 0]])
        key = np.sort(d * n + s)
        dd, ss = key // n, key % n
        cnt = np.bincount(dd, minlength=n)
        ip.append(np.cumsum(cnt) + eoff); idx.append(ss + noff)
        xs.append(np.stack([rng.integers(1, 36, n), rng.integers(0, 4, n), deg, rng.integers(0, 11, n),
                            rng.integers(0, 5, n), rng.integers(0, 5, n), rng.integers(0, 6, n),
                            rng.integers(0, 2, n), rng.integers(0, 2, n)], 1).astype(np.float32))
        noff += n; eoff += len(key); gp.append(noff)
    return (np.asarray(gp, dtype=np.int32), np.concatenate(ip).astype(np.int32),
            np.concatenate(idx).astype(np.int32), np.concatenate(xs, 0))


def synth_batch(seed: int, num_graphs: int, shape: str = "pcqm") -> BatchedGraph:
    gp, ip, idx, x = synth_arrays(seed, num_graphs, shape)
    return BatchedGraph(torch.from_numpy(gp), torch.from_numpy(ip), torch.from_numpy(idx), torch.from_numpy(x))
