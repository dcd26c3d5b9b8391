"""Name kept for drop-in compatibility only: the reference's build_multigraph.py imports a package that does not
exist (``struc_sim``) and is unreachable from pre-training.  Batched-graph construction:
``scgib_b200.graph.batch`` and ``scgib_b200.graph.khop_ego_batch``."""
from scgib_b200.graph import batch, graph, khop_ego_batch  # noqa: F401
