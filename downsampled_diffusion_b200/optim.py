"""Optimizer step of the training loop's caller.

The reference trains with `torch.optim.Adam(model.parameters(), lr)` (trainers/trainer.py:69) and, every step,
`clip_grad_norm_(model.parameters(), 1.) -> opt.step() -> opt.zero_grad() -> EMA.reset / EMA.update`
(trainers/trainer_ddpm.py:128-139, :243-254; update_ema :105-111): ~2000 small launches over ~380 tensors.
`Adam` here is a `torch.optim.Optimizer` with torch's state layout (`step`, `exp_avg`, `exp_avg_sq` per parameter, so
optimizer checkpoints written by the reference's Trainer.save_checkpoint load unchanged) whose `step()` is three launches
of libddb200: gradient-norm partials, the clip coefficient computed on the device, and one fused clip + Adam + EMA
(+ gradient reset) pass.  No host synchronisation: the norm `clip_grad_norm_` would return stays on the device in
`opt.grad_norm`.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib as L

CHUNK = 65536


class Adam(torch.optim.Optimizer):
    """Adam(lr, betas, eps) with optional fused gradient clipping (`max_grad_norm`) and EMA (`attach_ema`).

    Same update as torch.optim.Adam (no weight decay / amsgrad -- the reference uses neither).  All parameters must be
    contiguous fp32 CUDA tensors; anything else raises (no CPU fallback).

    Differences from `clip_grad_norm_` + `torch.optim.Adam` a caller can observe (none matters to the reference's loop):
      * the clip coefficient is applied inside the fused pass: `p.grad` itself is NOT scaled, so with `zero_grad=False` the
        gradients left behind are the unclipped ones;
      * the step count is kept per parameter GROUP: a parameter whose first gradient arrives in a later step gets the bias
        correction of the group's step count, not of its own first step (torch counts per parameter);
      * `opt.grad_norm` is a view of a device buffer that the next `step()` overwrites -- clone it to keep a value."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, max_grad_norm: Optional[float] = None):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        # the param-group keys of torch.optim.Adam, so that state_dicts move between the two classes in both directions
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                                      foreach=None, capturable=False, differentiable=False, fused=None,
                                      decoupled_weight_decay=False))
        self.max_grad_norm = max_grad_norm
        self.grad_norm: Optional[torch.Tensor] = None      # device scalar: total gradient norm of the last step (before clipping)
        self._ema = None
        self._plans = {}
        self._steps = {}                                    # group index -> number of steps taken

    # ---- EMA fusion -------------------------------------------------------------------------------
    def attach_ema(self, ema, model: torch.nn.Module) -> None:
        """Fuse `ema.update(model)` / `ema.reset(model)` into `step(ema=...)`.  `model.parameters()` must be the parameters
        this optimizer holds (the zip of trainers/ema.py:37 pairs them by order)."""
        self._ema = (ema, model)
        self._plans.clear()

    def _shadow_of(self):
        if self._ema is None:
            return {}
        ema, model = self._ema
        return {id(p): s for p, s in zip(model.parameters(), ema.ema_model.parameters())}

    # ---- launch plan per parameter group ---------------------------------------------------------------
    # torch keeps one CPU `step` tensor per parameter and increments each of them every step (~380 host operations); here
    # the count lives in one Python number per group and is written into the per-parameter tensors when a state_dict is taken.
    def _sync_steps(self) -> None:
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                st = self.state.get(p)
                if st:
                    st["step"] = torch.tensor(float(self._steps.get(gi, 0)), dtype=torch.float32)

    def state_dict(self):
        self._sync_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        self._plans.clear()                                 # exp_avg / exp_avg_sq were replaced
        self._steps = {}
        for gi, group in enumerate(self.param_groups):
            steps = [float(self.state[p]["step"]) for p in group["params"] if self.state.get(p)]
            if steps:
                if min(steps) != max(steps):
                    raise RuntimeError("parameters of one group carry different step counts")
                self._steps[gi] = int(steps[0])

    def _plan(self, gi: int, group, with_shadow: bool):
        """Pointer table + chunk list of the group.  Parameters, moments and shadows stay where they are between steps
        (load_state_dict / attach_ema / a replaced ema_model rebuild the plan); gradients re-created by autograd after
        zero_grad(set_to_none=True) may move, so their pointers -- and the parameters' -- are re-read every step and the
        table goes up again only when one changed, from pinned memory on the current stream (the host never waits)."""
        params = [p for p in group["params"] if p.grad is not None]
        ema_id = id(self._ema[0].ema_model) if (with_shadow and self._ema is not None) else 0
        key = (ema_id, tuple(p.data_ptr() for p in params), tuple(p.grad.data_ptr() for p in params))
        plan = self._plans.get((gi, with_shadow))
        if plan is not None and plan["key"] == key:
            return plan
        shadow = self._shadow_of() if with_shadow else {}
        rows, chunks = [], []
        for i, p in enumerate(params):
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            g, m, v = p.grad, st["exp_avg"], st["exp_avg_sq"]
            s = shadow.get(id(p))
            for t in (p, g, m, v) + ((s,) if s is not None else ()):
                if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                    raise RuntimeError("downsampled_diffusion_b200.Adam needs contiguous fp32 CUDA tensors (no CPU fallback)")
            rows += [p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), s.data_ptr() if s is not None else 0, p.numel()]
            chunks += [[i, c] for c in range((p.numel() + CHUNK - 1) // CHUNK)]
        dev = params[0].device
        table = torch.tensor(rows, dtype=torch.int64).pin_memory().to(dev, non_blocking=True)     # uint64 bit patterns
        numels = tuple(rows[5::6])
        if plan is not None and plan["numels"] == numels:
            plan.update(key=key, params=params, table=table)
        else:
            plan = dict(key=key, params=params, table=table, numels=numels,
                        chunks=torch.tensor(chunks, dtype=torch.int32, device=dev), n=len(chunks),
                        partial=torch.empty(len(chunks), dtype=torch.float32, device=dev),
                        norm=torch.zeros(2, dtype=torch.float32, device=dev))
        self._plans[(gi, with_shadow)] = plan
        return plan

    @torch.no_grad()
    def step(self, closure=None, ema: Optional[str] = None, zero_grad: bool = False):
        """One optimizer step.  ema: None (parameters only), 'update' (EMA.update fused, trainers/ema.py:36-44) or 'reset'
        (the shadow becomes a copy, what EMA.reset does during the first steps, trainer_ddpm.py:107-109) -- both need
        `attach_ema`.  zero_grad=True clears the gradients in the same pass (they stay allocated)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if ema not in (None, "update", "reset"):
            raise ValueError(f"ema must be None, 'update' or 'reset', got {ema!r}")
        if ema is not None and self._ema is None:
            raise RuntimeError("step(ema=...) needs attach_ema(ema, model) first")
        ema_mode = {None: 0, "update": 1, "reset": 2}[ema]
        if len(self.param_groups) > 1 and self.max_grad_norm:
            raise RuntimeError("fused gradient clipping spans one parameter group (the reference has one)")
        for gi, group in enumerate(self.param_groups):
            if not any(p.grad is not None for p in group["params"]):
                continue
            if group.get("weight_decay", 0) or group.get("amsgrad", False) or group.get("maximize", False):
                raise NotImplementedError("weight_decay / amsgrad / maximize are not part of the reference's optimizer")
            plan = self._plan(gi, group, ema_mode != 0)
            params = plan["params"]
            self._steps[gi] = self._steps.get(gi, 0) + 1
            step = float(self._steps[gi])
            beta1, beta2 = group["betas"]
            bc1 = 1.0 - beta1 ** step
            bc2 = 1.0 - beta2 ** step
            norm_ptr = None
            if self.max_grad_norm:
                L.call("dd_grad_norm", L.ptr(plan["table"]), L.ptr(plan["chunks"]), plan["n"], CHUNK, float(self.max_grad_norm),
                       L.ptr(plan["partial"]), L.ptr(plan["norm"]), L.stream())
                norm_ptr = L.ptr(plan["norm"])
                self.grad_norm = plan["norm"][0]
            decay = float(self._ema[0].decay) if self._ema is not None else 0.0
            L.call("dd_adam_ema_step", L.ptr(plan["table"]), L.ptr(plan["chunks"]), plan["n"], CHUNK, norm_ptr,
                   float(1.0 - beta1), float(beta2), float(1.0 - beta2), float(math.sqrt(bc2)), float(group["eps"]),
                   float(-(group["lr"] / bc1)), ema_mode, decay, float(1.0 - decay), 1 if zero_grad else 0, L.stream())
            # the kernel wrote through raw pointers: tell autograd / the packed-weight caches that the parameters changed
            torch.autograd.graph.increment_version(params)
        if ema_mode and self._ema is not None:
            # the shadow's weights changed under its packed caches: same bookkeeping as EMA.update, through the version counters
            em = self._ema[0].ema_model
            if getattr(self, "_shadow_params", (None, None))[0] is not em:
                self._shadow_params = (em, list(em.parameters()))
            torch.autograd.graph.increment_version(self._shadow_params[1])
        return loss


def _invalidate(model: torch.nn.Module) -> None:
    """(kept for callers that write a model's weights through raw pointers)"""
    for m in model.modules():
        if hasattr(m, "invalidate"):
            m.invalidate()
        if hasattr(m, "_programs"):
            for prog in m._programs.values():
                prog.weights_version = None
