"""scgib_b200 - B200-native (sm_100a) implementation of the S-CGIB per-batch pre-training hot path.

Layout
  csrc/      hand-written CUDA kernels + the C ABI (include/scgib.h) -> lib/libscgib.so (build.py)
  _lib.py    ctypes binding of the C ABI (fails loudly when the library is missing: no CPU fallback)
  graph.py   DGL-free batched CSR graphs, GPU k-hop ego-net extraction, synthetic molecule batches
  engine.py  PretrainEngine: flat parameters, workspace, forward / backward / Adam, data-parallel step
  models.py  drop-in nn.Modules with the reference's class names and forward signatures
"""
__version__ = "0.1.0"
