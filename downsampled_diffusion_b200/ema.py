"""Exponential moving average of model parameters.

Drop-in for trainers/ema.py:7-61 of the reference (`EMA(model, decay)`, `update`, `reset`, `eval`,
`sample`, `reconstruct`, `state_dict`, `load_state_dict`, attribute `ema_model`).  `update` is ONE
multi-tensor kernel launch (12 B / parameter) instead of ~3 ATen launches per parameter tensor.
"""
from __future__ import annotations

from copy import deepcopy

import torch
from torch import nn

from . import _lib as L

CHUNK = 65536


class EMA:
    def __init__(self, model: nn.Module, decay: float = 0.99):
        self.decay = decay
        self.ema_model = deepcopy(model)
        for p in self.ema_model.parameters():
            p.detach_()
        self._plan = None

    def eval(self):
        self.ema_model.eval()

    def reset(self, model: nn.Module):
        self.ema_model = deepcopy(model)
        self._plan = None

    def _build_plan(self, model: nn.Module):
        sh = list(self.ema_model.parameters())
        pr = list(model.parameters())
        assert len(sh) == len(pr)
        table, chunks = [], []
        for i, (s, p) in enumerate(zip(sh, pr)):
            if not (s.is_cuda and p.is_cuda):
                raise RuntimeError("EMA.update runs on CUDA parameters only (no CPU fallback)")
            assert s.dtype == torch.float32 and p.dtype == torch.float32 and s.is_contiguous() and p.is_contiguous()
            assert s.numel() == p.numel()
            table += [s.data_ptr(), p.data_ptr(), s.numel()]
            chunks += [[i, c] for c in range((s.numel() + CHUNK - 1) // CHUNK)]
        dev = sh[0].device
        key = tuple(table)
        tab = torch.tensor(table, dtype=torch.int64, device=dev)      # uint64 bit patterns
        chk = torch.tensor(chunks, dtype=torch.int32, device=dev)
        self._plan = (key, tab, chk, len(chunks))

    def update(self, model: nn.Module):
        """p_ema <- p_ema*decay + (1-decay)*p for every parameter (buffers untouched): ema.py:36-44."""
        key = []
        for s, p in zip(self.ema_model.parameters(), model.parameters()):
            key += [s.data_ptr(), p.data_ptr(), s.numel()]
        if self._plan is None or self._plan[0] != tuple(key):
            self._build_plan(model)
        _, tab, chk, n = self._plan
        L.call("dd_ema_update", L.ptr(tab), L.ptr(chk), n, CHUNK, float(self.decay), float(1 - self.decay), L.stream())
        for m in self.ema_model.modules():          # packed bf16 weight caches are now stale
            if hasattr(m, "invalidate"):
                m.invalidate()
            if hasattr(m, "_programs"):
                for prog in m._programs.values():
                    prog.weights_version = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.ema_model(x)

    @torch.no_grad()
    def sample(self, n: int):
        return self.ema_model.sample(n)

    @torch.no_grad()
    def reconstruct(self, x: torch.Tensor, n: int):
        return self.ema_model.reconstruct(x, n)

    def load_state_dict(self, state_dict) -> None:
        self.ema_model.load_state_dict(state_dict)

    def state_dict(self):
        return self.ema_model.state_dict()
