"""Functional wrappers over the elementwise C-ABI entry points (torch tensors in, torch tensors out).

Each wrapper is differentiable where the reference's training step differentiates through it
(custom autograd.Function whose backward is again a libddb200 kernel); nothing here calls ATen math.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L


def _flat(x: torch.Tensor):
    B = x.shape[0]
    chw = x.numel() // B
    return B, chw


def _f32c(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("downsampled_diffusion_b200 runs on CUDA tensors only (no CPU fallback)")
    return x.contiguous().float()


def _i64c(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous().to(torch.int64)


def q_sample_raw(x, eps, t, sqrt_ac, sqrt_1mac) -> torch.Tensor:
    x, eps, t = _f32c(x), _f32c(eps), _i64c(t)
    B, chw = _flat(x)
    out = torch.empty_like(x)
    L.call("dd_q_sample", L.ptr(x), L.ptr(eps), L.ptr(t), L.ptr(sqrt_ac), L.ptr(sqrt_1mac), L.ptr(out), B, chw, L.stream())
    return out


def predict_x0_raw(x_t, eps, t, sqrt_recip, sqrt_recipm1, clip: bool) -> torch.Tensor:
    x_t, eps, t = _f32c(x_t), _f32c(eps), _i64c(t)
    B, chw = _flat(x_t)
    out = torch.empty_like(x_t)
    L.call("dd_predict_x0", L.ptr(x_t), L.ptr(eps), L.ptr(t), L.ptr(sqrt_recip), L.ptr(sqrt_recipm1), 1 if clip else 0,
           L.ptr(out), B, chw, L.stream())
    return out


def scale_rows_raw(g: torch.Tensor, t: torch.Tensor, table: torch.Tensor, neg: bool = False) -> torch.Tensor:
    """g * table[t_b] per sample, via q_sample's kernel (b-term zero): used by the backward formulas."""
    zeros_tab = torch.zeros_like(table)
    tab = -table if neg else table
    return q_sample_raw(g, g, t, tab, zeros_tab)


class _QSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps, t, sa, sb):
        ctx.save_for_backward(t, sa, sb)
        return q_sample_raw(x, eps, t, sa, sb)

    @staticmethod
    def backward(ctx, g):
        t, sa, sb = ctx.saved_tensors
        gx = scale_rows_raw(g, t, sa) if ctx.needs_input_grad[0] else None
        ge = scale_rows_raw(g, t, sb) if ctx.needs_input_grad[1] else None
        return gx, ge, None, None, None


def q_sample(x, eps, t, sa, sb) -> torch.Tensor:
    if torch.is_grad_enabled() and (x.requires_grad or eps.requires_grad):
        return _QSample.apply(x, eps, t, sa, sb)
    return q_sample_raw(x, eps, t, sa, sb)


class _PredictX0(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_t, eps, t, ra, rb, clip):
        out = predict_x0_raw(x_t, eps, t, ra, rb, clip)
        ctx.clip = clip
        ctx.save_for_backward(t, ra, rb, out)
        return out

    @staticmethod
    def backward(ctx, g):
        t, ra, rb, out = ctx.saved_tensors
        if ctx.clip:
            raise NotImplementedError("gradient through the clamped x0 prediction is not part of the reference's training path")
        gx = scale_rows_raw(g, t, ra) if ctx.needs_input_grad[0] else None
        ge = scale_rows_raw(g, t, rb, neg=True) if ctx.needs_input_grad[1] else None
        return gx, ge, None, None, None, None


def predict_x0(x_t, eps, t, ra, rb, clip: bool) -> torch.Tensor:
    if torch.is_grad_enabled() and (x_t.requires_grad or eps.requires_grad):
        return _PredictX0.apply(x_t, eps, t, ra, rb, clip)
    return predict_x0_raw(x_t, eps, t, ra, rb, clip)


def mse_rowsum_raw(a, b, mean: bool) -> torch.Tensor:
    a, b = _f32c(a), _f32c(b)
    B, chw = _flat(a)
    out = torch.empty(B, dtype=torch.float32, device=a.device)
    L.call("dd_mse_rowsum", L.ptr(a), L.ptr(b), L.ptr(out), B, chw, (1.0 / chw) if mean else 1.0, L.stream())
    return out


class _MseRows(torch.autograd.Function):
    """Per-sample sum (or mean) of squared differences; gradients flow to both operands."""

    @staticmethod
    def forward(ctx, a, b, mean):
        ctx.mean = mean
        ctx.save_for_backward(a, b)
        return mse_rowsum_raw(a, b, mean)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        a, b = _f32c(a), _f32c(b)
        B, chw = _flat(a)
        gb = torch.empty_like(b)
        L.call("dd_mse_rowsum_bwd", L.ptr(a), L.ptr(b), L.ptr(_f32c(g)), L.ptr(gb), B, chw,
               (1.0 / chw) if ctx.mean else 1.0, L.stream())
        ga = None
        if ctx.needs_input_grad[0]:
            ga = torch.empty_like(a)
            L.call("dd_mse_rowsum_bwd", L.ptr(b), L.ptr(a), L.ptr(_f32c(g)), L.ptr(ga), B, chw,
                   (1.0 / chw) if ctx.mean else 1.0, L.stream())
        return ga, (gb if ctx.needs_input_grad[1] else None), None


def mse_rows(a, b, mean: bool) -> torch.Tensor:
    """flatten_loss(l2_loss(a, b, 'none')): models/utils/losses.py:12-14 + utils/utils.py:27-40."""
    if torch.is_grad_enabled() and (a.requires_grad or b.requires_grad):
        return _MseRows.apply(a, b, mean)
    return mse_rowsum_raw(a, b, mean)


class SquaredError:
    """`l2_loss(target, output, reduction='none')` (models/utils/losses.py:12-14) kept as its operand pair: the reference
    only ever reduces it (`flatten_loss(...)`, `.mean()`), which the fused row-sum kernel does without materialising it."""

    def __init__(self, target: torch.Tensor, output: torch.Tensor):
        assert target.shape == output.shape
        self.target, self.output = target, output
        self.shape = target.shape

    def tensor(self) -> torch.Tensor:
        a, b = _f32c(self.target), _f32c(self.output)
        y = torch.empty_like(a)
        L.call("dd_ew", 5, L.ptr(a), L.ptr(b), L.ptr(y), a.numel(), 1.0, 0, L.stream())
        return y

    def sum(self) -> torch.Tensor:
        return mse_rows(self.target, self.output, False).sum()

    def mean(self) -> torch.Tensor:
        return mse_rows(self.target, self.output, False).sum() / float(self.target.numel())


def reduce_rows(x: torch.Tensor, mean: bool) -> torch.Tensor:
    """Sum (or mean) over all non-batch dimensions of a plain tensor (utils/utils.py:27-40); forward only."""
    if torch.is_grad_enabled() and x.requires_grad:
        raise NotImplementedError("flatten_loss of a plain tensor is forward-only; pass get_loss(target, output) for the differentiable form")
    x = _f32c(x)
    B, chw = _flat(x)
    out = torch.zeros(B, dtype=torch.float32, device=x.device)
    L.call("dd_colsum", L.ptr(x), L.ptr(out), chw, B, 1, chw, L.stream())      # (1, B, chw) read as NCHW: channel sums = row sums
    return out / float(chw) if mean else out


def posterior_step_raw(x_t, eps_hat, noise, coef, t_idx, t_stride: int, noise_step_stride: int, T: int,
                       noise_period: int, clip: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, chw = _flat(x_t)
    out = torch.empty_like(x_t) if out is None else out
    L.call("dd_posterior_step", L.ptr(x_t), L.ptr(eps_hat), L.ptr(noise), L.ptr(coef), L.ptr(t_idx), t_stride,
           noise_step_stride, T, noise_period, 1 if clip else 0, L.ptr(out), B, chw, L.stream())
    return out


# ---- evaluation-side chain and output formatting (8(f).2) --------------------------------------
def vlb_terms_raw(x, x_t, eps_hat, tab, t32, T: int) -> torch.Tensor:
    """DDPM.vlb_terms (ddpm.py:317-365) given the U-Net output, per-sample int32 t: (B,) bits/dim."""
    x, x_t, eps_hat = _f32c(x), _f32c(x_t), _f32c(eps_hat)
    B, chw = _flat(x)
    out = torch.empty(B, dtype=torch.float32, device=x.device)
    L.call("dd_vlb_terms", L.ptr(x), L.ptr(x_t), L.ptr(eps_hat), None, 0, 0, L.ptr(tab), L.ptr(t32), 1, T, L.ptr(out), None, 1, 0,
           B, chw, L.stream())
    return out


def prior_kl(x, sqrt_ac_last: float, log_1mac_last: float) -> torch.Tensor:
    """DDPM.calc_prior (ddpm.py:367-389): (B,) bits/dim."""
    x = _f32c(x)
    B, chw = _flat(x)
    out = torch.empty(B, dtype=torch.float32, device=x.device)
    L.call("dd_prior_kl", L.ptr(x), float(sqrt_ac_last), float(log_1mac_last), L.ptr(out), B, chw, L.stream())
    return out


def fix_samples_raw(samples: torch.Tensor) -> torch.Tensor:
    """Per-image min-max normalisation, x255, NCHW -> NHWC (device tensor); utils/eval_helpers.py:37-41."""
    x = _f32c(samples)
    B, C, H, W = x.shape
    out = torch.empty(B, H, W, C, dtype=torch.float32, device=x.device)
    L.call("dd_fix_samples", L.ptr(x), L.ptr(out), B, C, H, W, L.stream())
    return out
