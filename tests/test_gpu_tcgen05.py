"""tcgen05 / tensor-memory path on the GPU: the 3xTF32 tile-GEMM primitives (all operand-major modes, M = 64 and
128, TMEM row mapping) and the tensor-core GIN forward kernel against fp64."""
import pytest
import torch

from oracle.graph_ref import synth_batch
from oracle.scgib_oracle import GINConvRef, MLP, tgraph_from_ref
from tests.helpers import product_graph, rel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("M", [64, 128])
def test_umma_3xtf32_tile_gemm(M, mode):
    """mode 0: A K-major, B K-major; mode 1: B MN-major (natural weights); mode 2: both MN-major (reduction over the
    tile rows).  fp32-level accuracy (3xTF32) and the documented TMEM lane of every accumulator row."""
    from scgib_b200 import _lib
    from tests import probe_lib
    lib = probe_lib.load()
    torch.manual_seed(M + mode)
    A = torch.randn(M, 64, device=DEV)
    B = torch.randn(M if mode == 2 else 64, 64, device=DEV)
    out = torch.full((128, 64), float("nan"), device=DEV)
    _lib.check(lib.scgib_debug_umma(_lib.ptr(A), _lib.ptr(B), _lib.ptr(out), M, mode, None))
    torch.cuda.synchronize()
    Ad, Bd = A.double(), B.double()
    ref = (Ad @ Bd.t()) if mode == 0 else (Ad @ Bd) if mode == 1 else (Ad.t() @ Bd)
    R = ref.shape[0]
    lanes = list(range(128)) if R == 128 else [(i // 16) * 32 + i % 16 for i in range(64)]
    assert rel(out[lanes], ref) <= 3e-6          # plain TF32 would be ~1e-3


@pytest.mark.parametrize("mode,B", [(0, 300), (1, 300), (1, 6000), (1, 20000)])
@pytest.mark.parametrize("kin", [32, 64])
def test_gin_layer_forward_tensor_cores(kin, mode, B):
    """mode 1: gin_tc3.cu (warp-specialised tcgen05 kernel), mode 0: the FFMA cross-check.  B = 6000 graphs give
    ~90 k rows = several 128-row tiles per CTA (both pipeline stages reused)."""
    from scgib_b200 import _lib, ops
    lib = _lib.load()
    g = synth_batch(2, B)
    tg = tgraph_from_ref(g)
    torch.manual_seed(kin)
    conv = GINConvRef(MLP(kin, 64, 64)).double()
    h = torch.randn(g.num_nodes, kin, dtype=torch.float64)
    y_ref = conv(tg, h)
    lin1, lin2 = conv.apply_func.mlp[0], conv.apply_func.mlp[2]
    pg = product_graph(g, DEV)
    f = lambda t: t.detach().float().to(DEV)
    lib.scgib_set_tensor_cores(mode)
    try:
        y, bn, a, r = ops.gin_layer_fwd(f(h), pg.indptr, pg.indices, f(lin1.weight), f(lin1.bias), f(lin2.weight),
                                        f(lin2.bias), save=True)
        torch.cuda.synchronize()
    finally:
        lib.scgib_set_tensor_cores(-1)
    assert rel(y, y_ref) <= 5e-6
    assert rel(bn[0], y_ref.mean(0)) <= 1e-5
    assert rel(bn[1], 1.0 / torch.sqrt(y_ref.var(0, unbiased=False) + 1e-5)) <= 1e-5
    assert rel(r, torch.relu(lin1(h + torch.zeros_like(h).index_add(0, tg.dst, h[tg.src])))) <= 5e-6
