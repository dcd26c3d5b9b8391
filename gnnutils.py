"""Name kept for drop-in compatibility only.  In the reference this file holds node-classification utilities of an
earlier project that the pre-training path never reaches; graph batching lives in
``scgib_b200.graph`` / ``molecules.py``."""
from scgib_b200.graph import BatchedGraph, EgoBatch, batch, graph, khop_ego_batch, sum_nodes  # noqa: F401
