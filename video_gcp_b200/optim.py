"""Optimiser step on the device (reference: the `self.optimizer.step()` of gcp/prediction/train.py:162).

The reference trains with RAdam or Adam (gcp_builder.py:174-186,259; blox/torch/radam.py; torch.optim.Adam) wrapped in
`ClipGradOptimizer` (blox/torch/training.py:146-161), which clips the global gradient norm before every step.
`ClippedOptimizer` keeps that surface -- `zero_grad()`, `step()`, `state_dict()` / `load_state_dict()` in
torch.optim's format (so `CheckpointHandler.load_weights(..., load_step_and_opt=True)` restores it), `param_groups[0]`
with lr / betas / eps / weight_decay -- and runs the update as one `gcpb200_optim_step` launch per parameter tensor
(+ one `gcpb200_sq_norm` launch per gradient when clipping; the norm never leaves the device).  There is no CPU path.

Contract with the model: the engine computes with its own packed copy of the weights (bf16 tiles, folded BatchNorm), so
after `step()` the owning model must repack before its next forward.  Pass `model=` (a TreeModel / SequentialModel):
`step()` then marks it dirty and the next `model.engine` access repacks from the updated parameters
(tests/test_gpu_parity_optim.py::test_step_then_forward_uses_new_weights).  Two differences from
`torch.nn.utils.clip_grad_norm_` + step: `p.grad` itself is left unscaled (the scale is folded into the update kernel),
and `gradient_clip=None` / 0 disables clipping.
"""
import ctypes as C

import torch

from . import _C

KINDS = {"adam": 0, "radam": 1}


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class ClippedOptimizer:
    def __init__(self, params, engine, optimizer_type="radam", lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 gradient_clip=None, model=None):
        if optimizer_type not in KINDS:
            raise ValueError("Optimizer '{}' not supported!".format(optimizer_type))      # gcp_builder.py:185
        self.engine = engine
        self.model = model
        self.kind = KINDS[optimizer_type]
        self.gradient_clip = gradient_clip
        self.params = [p for p in params]
        for p in self.params:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise _C.GcpB200Error("ClippedOptimizer needs contiguous fp32 CUDA parameters (there is no CPU path)")
        self.param_groups = [dict(params=self.params, lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)]
        self.state = {}
        self._sq = torch.zeros(1, device=engine.device, dtype=torch.float64)

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def _state(self, p):
        st = self.state.get(p)
        if st is None:
            st = dict(step=0, exp_avg=torch.zeros_like(p.data), exp_avg_sq=torch.zeros_like(p.data))
            self.state[p] = st
        return st

    def step(self, closure=None):
        loss = closure() if closure is not None else None
        eng, g0 = self.engine, self.param_groups[0]
        live = [p for p in self.params if p.grad is not None]
        with torch.cuda.device(eng.index):
            stream = C.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
            sq = None
            if self.gradient_clip is not None:
                self._sq.zero_()
                for p in live:
                    g = p.grad.contiguous()
                    eng._check(eng.lib.gcpb200_sq_norm(eng.h, _ptr(g), g.numel(), _ptr(self._sq), stream))
                sq = _ptr(self._sq)
            for p in live:
                st = self._state(p)
                st["step"] += 1
                g = p.grad.contiguous()
                eng._check(eng.lib.gcpb200_optim_step(
                    eng.h, self.kind, _ptr(p.data), _ptr(g), _ptr(st["exp_avg"]), _ptr(st["exp_avg_sq"]), p.numel(),
                    float(g0["lr"]), float(g0["betas"][0]), float(g0["betas"][1]), float(g0["eps"]),
                    float(g0["weight_decay"]), int(st["step"]), sq, float(self.gradient_clip or 0.0), stream))
        if self.model is not None and live:
            self.model._dirty = True        # the engine's packed weights are stale: repack on the next engine access
        return loss

    # ---- torch.optim.Optimizer checkpoint format: {'state': {index: {...}}, 'param_groups': [{..., 'params': [indices]}]}
    def state_dict(self):
        index = {p: i for i, p in enumerate(self.params)}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(self.params)))
        return dict(state={index[p]: dict(step=st["step"], exp_avg=st["exp_avg"], exp_avg_sq=st["exp_avg_sq"])
                           for p, st in self.state.items()}, param_groups=[group])

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("loaded state dict has a different number of parameter groups / parameters")
        for k, v in groups[0].items():
            if k != "params":
                self.param_groups[0][k] = v
        self.state = {}
        for i, st in sd["state"].items():
            p = self.params[int(i)]
            self.state[p] = dict(step=int(st["step"]),
                                 exp_avg=st["exp_avg"].to(p.device, torch.float32).contiguous().clone(),
                                 exp_avg_sq=st["exp_avg_sq"].to(p.device, torch.float32).contiguous().clone())


def get_clipped_optimizer(params, engine, optimizer_type="radam", gradient_clip=None, model=None, **kwargs):
    """blox/torch/training.py:146-161 with the optimiser named instead of passed as a class."""
    return ClippedOptimizer(params, engine, optimizer_type=optimizer_type, gradient_clip=gradient_clip, model=model, **kwargs)
