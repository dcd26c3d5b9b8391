"""Optimiser for the drop-in modules: ``torch.optim.Adam(lr, weight_decay)`` (reference exp_pretraining.py:86,112) as
ONE kernel over the engine's flat parameter / gradient / moment buffers (``scgib_adam_step_f32``) instead of ~10 foreach
launches over the 61 parameter tensors (0.67 ms -> 0.01 ms per step).  Same ``zero_grad()`` / ``step()`` surface, so the
reference's training loop is unchanged.  Under ``torch.distributed`` (torchrun, one process per GPU) ``step()`` first
sum-all-reduces the flat gradient buffer, so the same loop trains data-parallel."""
from __future__ import annotations

import torch


class FlatAdam:
    """Adam with L2-in-gradient weight decay over every parameter of ``model`` that lives in its engine's flat buffer
    (all parameters the pre-training path uses).  Parameters outside the flat buffer never receive gradients on this
    path; if one ever does, ``step`` raises instead of silently skipping it."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-5):
        if not hasattr(model, "_bridge"):
            raise TypeError("FlatAdam needs an scgib_b200 drop-in module (Mainmodel / Mainmodel_continue)")
        self.model = model
        self.param_groups = [dict(params=list(model.parameters()), lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)]

    def zero_grad(self, set_to_none: bool = True):
        cached = getattr(self.model._bridge, "_cached", None)
        params = cached[0] if cached is not None else self.param_groups[0]["params"]   # only these ever get gradients
        for p in params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        bridge = self.model._bridge
        eng = bridge.engine
        cached = getattr(bridge, "_cached", None)
        if eng is None or cached is None:
            raise RuntimeError("FlatAdam.step() before the first forward/backward")
        key = (eng.grads.data_ptr(), id(cached))
        if getattr(self, "_ptr_key", None) != key:         # pointer of every slot of the flat gradient buffer, once
            gv = eng.grad_views()
            self._views = [gv[name] for name in bridge.slot_names]
            self._ptrs = [v.data_ptr() for v in self._views]
            flat = {id(p) for p in cached[0]}
            self._others = [p for p in self.param_groups[0]["params"] if id(p) not in flat]
            self._ptr_key = key
        for p, ptr, v, name in zip(cached[0], self._ptrs, self._views, bridge.slot_names):
            g_ = p.grad
            if g_ is None:
                raise RuntimeError("parameter %s has no gradient: call step() after loss.backward()" % name)
            if g_.data_ptr() != ptr:                       # accumulated / clipped copy: put it back into the flat buffer
                v.copy_(g_.reshape(v.shape))
        for p in self._others:
            if p.grad is not None:
                raise RuntimeError("a parameter outside the engine's flat buffer received a gradient; use torch.optim.Adam")
        g = self.param_groups[0]
        world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size()
            if world > 1:          # data parallelism: ONE sum all-reduce of the flat gradient buffer, mean applied in the kernel
                torch.distributed.all_reduce(eng.grads, op=torch.distributed.ReduceOp.SUM)
        eng.adam_step(lr=g["lr"], betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"], grad_scale=1.0 / world)

    def state_dict(self):
        eng = self.model._bridge.engine
        return dict(step=0 if eng is None else eng.step_count, exp_avg=None if eng is None else eng.exp_avg.clone(),
                    exp_avg_sq=None if eng is None else eng.exp_avg_sq.clone(),
                    param_groups=[{k: v for k, v in self.param_groups[0].items() if k != "params"}])

    def load_state_dict(self, sd):
        eng = self.model._bridge.engine
        if eng is not None and sd.get("exp_avg") is not None:
            eng.step_count = int(sd["step"])
            eng.exp_avg.copy_(sd["exp_avg"])
            eng.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_groups[0].update(sd["param_groups"][0])


class AllReduceAdam(torch.optim.Adam):
    """``torch.optim.Adam`` over module parameters that averages the gradients over the data-parallel ranks first: ONE sum
    all-reduce of the flattened gradients per step (the modules without a flat engine buffer: --encoder GraphSAGE / GCN).
    Single process: plain Adam."""

    @torch.no_grad()
    def step(self, closure=None):
        dist = torch.distributed
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            grads = [p.grad for g in self.param_groups for p in g["params"] if p.grad is not None]
            if grads:
                flat = torch.cat([g.reshape(-1) for g in grads])
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                flat /= dist.get_world_size()
                off = 0
                for g in grads:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()
        return super().step(closure)
