"""Noise schedule of the diffusion chain: host-side float64 constants cast to fp32.

Mirrors models/diffusion/beta_schedule.py:5-33 and the buffer set of models/diffusion/ddpm.py:55-105
of the reference (same numpy float64 arithmetic, same cast), so the 12 persistent buffers are
bit-identical and a reference checkpoint loads unchanged.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

SCHEDULE_NAMES = ("linear", "cosine")


def make_beta_schedule(schedule: str, n_timesteps: int, linear_start: float = 1e-4, linear_end: float = 2e-2,
                       cosine_s: float = 8e-3) -> np.ndarray:
    if schedule == "linear":
        # range scaled so that T=1000 gives the DDPM paper's 1e-4 .. 2e-2
        k = 1000 / n_timesteps
        return np.linspace(k * linear_start, k * linear_end, n_timesteps, dtype=np.float64)
    if schedule == "cosine":
        grid = torch.arange(n_timesteps + 1, dtype=torch.float64) / n_timesteps + cosine_s
        abar = torch.cos(grid / (1 + cosine_s) * np.pi / 2).pow(2)
        abar = abar / abar[0]
        return np.clip((1 - abar[1:] / abar[:-1]).numpy(), a_min=0, a_max=0.999)
    raise ValueError(f"schedule '{schedule}' unknown.")


def diffusion_buffers(schedule: str, T: int) -> "OrderedDict[str, torch.Tensor]":
    """Registration-ordered buffers of DDPM (ddpm.py:78-105); 'vlb_weights' is the non-persistent one."""
    betas = make_beta_schedule(schedule, T)
    assert (betas > 0).all() and (betas <= 1).all(), "betas must be in (0, 1]"
    alphas = 1.0 - betas
    abar = np.cumprod(alphas, axis=0)
    abar_prev = np.append(1.0, abar[:-1])
    post_var = (1.0 - abar_prev) / (1.0 - abar) * betas
    c_x0 = np.sqrt(abar_prev) * betas / (1.0 - abar)
    c_xt = np.sqrt(alphas) * (1.0 - abar_prev) / (1.0 - abar)
    post_logvar = np.log(np.append(post_var[1], post_var[1:]))   # variance is 0 at t=0: reuse t=1

    def f32(a):
        return torch.tensor(a, dtype=torch.float32)

    out = OrderedDict()
    out["betas"] = f32(betas)
    out["alphas_cumprod"] = f32(abar)
    out["alphas_cumprod_prev"] = f32(abar_prev)
    out["sqrt_alphas_cumprod"] = f32(np.sqrt(abar))
    out["sqrt_one_minus_alphas_cumprod"] = f32(np.sqrt(1.0 - abar))
    out["log_one_minus_alphas_cumprod"] = f32(np.log(1.0 - abar))
    out["sqrt_recip_alphas_cumprod"] = f32(np.sqrt(1.0 / abar))
    out["sqrt_recipm1_alphas_cumprod"] = f32(np.sqrt(1.0 / abar - 1))
    out["posterior_variance"] = f32(post_var)
    out["posterior_log_variance_clipped"] = f32(post_logvar)
    out["posterior_mean_coef1"] = f32(c_x0)
    out["posterior_mean_coef2"] = f32(c_xt)
    w = out["betas"] ** 2 / (2 * out["posterior_variance"] * f32(alphas) * (1 - out["alphas_cumprod"]))
    w[0] = w[1]
    out["vlb_weights"] = w
    return out


def posterior_coef_table(buf) -> torch.Tensor:
    """(T, 5) fp32 rows consumed by dd_posterior_step:
    {sqrt_recip_ac, sqrt_recipm1_ac, post_mean_coef1, post_mean_coef2, exp(0.5*post_logvar_clipped)}.
    The last column is evaluated with torch on the host in fp32, as ddpm.py:227 does per element."""
    sig = (0.5 * buf["posterior_log_variance_clipped"].detach().float().cpu()).exp()
    cols = [buf["sqrt_recip_alphas_cumprod"], buf["sqrt_recipm1_alphas_cumprod"],
            buf["posterior_mean_coef1"], buf["posterior_mean_coef2"]]
    cols = [c.detach().float().cpu() for c in cols] + [sig]
    return torch.stack(cols, dim=1).contiguous()


def eval_coef_table(buf) -> torch.Tensor:
    """(T, 8) fp32 rows consumed by dd_q_sample_step / dd_vlb_terms: {sqrt_ac, sqrt_1mac, sqrt_recip_ac,
    sqrt_recipm1_ac, post_mean_coef1, post_mean_coef2, post_logvar_clipped, 0} -- the scalars DDPM.vlb_terms
    (ddpm.py:339-342) gathers with extract()."""
    cols = [buf[k].detach().float().cpu() for k in (
        "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
        "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped")]
    return torch.stack(cols + [torch.zeros_like(cols[0])], dim=1).contiguous()
