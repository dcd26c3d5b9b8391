"""CPU oracle for the S-CGIB pre-training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  The product path (``s-cgib_b200/``) never imports
this package and fails loudly when the CUDA library is missing.

Parity status: the reference (``/root/reference/models.py``) needs DGL 1.1.0,
torch_geometric, ogb, pyro and torch_scatter, none of which is installable
here, and it ships no tests or golden vectors.  The oracle is therefore a
restatement.  It is pinned in two ways (see ``tests/golden/README.md``):

* ``tests/golden/make_golden.py`` imports the *unmodified* reference
  ``models.py`` with a minimal stand-in for the DGL call surface
  (``tests/golden/dgl_stub``) and records inputs/outputs of
  ``Mainmodel.forward`` + ``backward``; ``tests/test_oracle_golden.py``
  checks the oracle against those vectors.  Everything that is reference
  Python (compress/compression/attention loop/losses) is pinned this way.
* DGL's own semantics (GINConv, batch, sum_nodes, adj, khop_in_subgraph,
  to_bidirected) are restated from the library's documented v1.1.0 behaviour
  in both the stub and the oracle: **that part of parity is unpinned**.
"""
