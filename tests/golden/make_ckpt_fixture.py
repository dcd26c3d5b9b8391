"""Extract the tensors of the reference's shipped checkpoint (outputs/pre_training_v1_GIN_64_5_1.pt, a whole-module
pickle of models.Mainmodel_continue around models.Mainmodel with FIVE GINConv layers per encoder: SURVEY F4) into a
plain state dict that travels to the GPU box.  Run in the build container only (needs /root/reference):
    python tests/golden/make_ckpt_fixture.py
Writes tests/golden/shipped_ckpt_v1_GIN_64_5_1.pt = {"state": {name: tensor}, "meta": {...}}: the tensors the hot path
uses, under the reference Mainmodel's names - transfer_d / MLP of the pickled wrapper and the encoders, compressor and
attention layer of its loaded ``self.model`` (what Mainmodel_continue.forward runs and trains, models.py:1167).
The pickle references models.* and dgl.* classes; a stub Unpickler rebuilds them as bare nn.Module shells, so neither
DGL nor the reference code is executed."""
import os
import pickle
import sys

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/outputs/pre_training_v1_GIN_64_5_1.pt"


class _Shell(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in ("models", "dgl", "torch_geometric", "ogb", "pyro", "gnnutils", "build_multigraph"):
            return type(name, (_Shell,), {})
        return super().find_class(module, name)


class _PickleModule:
    Unpickler = _Unpickler
    load = staticmethod(lambda f, **k: _Unpickler(f, **k).load())
    __name__ = "stub_pickle"


def tensors(mod, prefix=""):
    out = {}
    for n, p in mod._parameters.items():
        if p is not None:
            out[prefix + n] = p.detach().clone()
    for n, b in mod._buffers.items():
        if b is not None:
            out[prefix + n] = b.detach().clone()
    for n, m in mod._modules.items():
        if m is not None and n != "model":
            out.update(tensors(m, prefix + n + "."))
    return out


def main():
    obj = torch.load(SRC, map_location="cpu", weights_only=False, pickle_module=_PickleModule)
    inner = obj._modules.get("model")
    outer_t = tensors(obj)
    inner_t = tensors(inner) if inner is not None else {}
    L = len([k for k in (inner_t or outer_t) if k.startswith("Encoder1.ginlayers.") and k.endswith("mlp.0.weight")])
    meta = dict(source=os.path.basename(SRC), cls=type(obj).__name__, inner_cls=type(inner).__name__ if inner is not None else None,
                gin_layers=L, n_outer=sum(v.numel() for v in outer_t.values()), n_inner=sum(v.numel() for v in inner_t.values()))
    print(meta)
    for k, v in list(inner_t.items())[:6]:
        print("inner", k, tuple(v.shape))
    # the tensors the hot path uses (models.py:1158-1195): transfer_d / MLP of the wrapper, everything else of its loaded model
    state = {k: v for k, v in outer_t.items() if k.startswith(("transfer_d.", "MLP."))}
    state.update({k: v for k, v in inner_t.items() if k.startswith(("Encoder1.", "Encoder2.", "compressor.", "attn_layer."))})
    meta["n_hotpath"] = sum(v.numel() for k, v in state.items() if v.dtype.is_floating_point and "running" not in k and not k.endswith("eps"))
    print(meta)
    torch.save(dict(state=state, meta=meta), os.path.join(HERE, "shipped_ckpt_v1_GIN_64_5_1.pt"))


if __name__ == "__main__":
    main()
