"""Gaussian diffusion wrapper: schedule buffers, q_sample, posterior, ancestral sampling, training loss.

Drop-in for models/diffusion/ddpm.py:22-457 of the reference: same constructor
`DDPM(config, latent_model, device, color_channels=3)`, method names, return types and buffers.
The arithmetic runs in libddb200 kernels; the T-step chain replays ONE captured CUDA graph per step
(U-Net + posterior update + step-counter tick) with no host synchronisation in between.
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import torch
from torch import nn

from . import _lib as L
from . import ops
from .engine import EngineCache
from .schedule import diffusion_buffers, eval_coef_table, posterior_coef_table

OBJECTIVE_NAMES = ("simple", "vlb", "hybrid")
NOISE_RING_BYTES = 4 << 30      # pre-drawn chain noise is held in a ring of at most this many bytes
HOST_NOISE_BLOCK = 50           # steps per host-to-device block of pre-drawn noise (copied on a side stream)


class SamplingPlan:
    """Captured sampling step for one (batch, latent shape, precision)."""

    def __init__(self, ddpm: "DDPM", shape: Sequence[int]):
        B, C, H, W = shape
        self.shape = tuple(shape)
        self.T = ddpm.timesteps
        dev = next(ddpm.latent_model.parameters()).device
        self.eng = ddpm.latent_model.engine(B, H, W)
        self.chw = C * H * W
        self.t_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.coef = posterior_coef_table({k: getattr(ddpm, k) for k in (
            "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
            "posterior_mean_coef2", "posterior_log_variance_clipped")}).to(dev)
        per_step = B * self.chw * 4
        self.period = max(1, min(self.T, NOISE_RING_BYTES // per_step))
        self.noise = torch.empty(self.period, B, C, H, W, dtype=torch.float32, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.copy_stream = torch.cuda.Stream(device=dev)          # host noise blocks are uploaded here (p_sample_loop)
        self.clip = ddpm.clip_denoised

    def step_eager(self) -> None:
        e = self.eng
        e.run()
        ops.posterior_step_raw(e.x_in, e.eps_out, self.noise, self.coef, self.t_dev, 0, self.shape[0] * self.chw,
                               self.T, self.period, self.clip, out=e.x_in)
        L.call("dd_tick", L.ptr(self.t_dev), 1, L.stream())

    def prepare(self) -> None:
        """(Re)build the time-bias table and the graph when the U-Net's weights changed."""
        e = self.eng
        e.bind_table(e.time_table(self.T), self.t_dev, 0)       # refreshes the packed weights, refills the table in place if stale
        if self.graph is None:
            # warm-up on a side stream (lazy CUDA initialisation must not happen inside capture)
            self.t_dev.fill_(self.T - 1)
            self.noise[0].zero_()
            e.x_in.zero_()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.step_eager()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.step_eager()
            self.graph = g
            self.launches_per_step = len(e.ops) + 3


EVAL_TABLE_KEYS = ("sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                   "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
                   "posterior_log_variance_clipped")


class EvalPlan(SamplingPlan):
    """Captured step of the evaluation chain (DDPM.test_losses_, ddpm.py:391-442): q_sample at the device-side step
    counter -> U-Net -> variational-bound term and squared-error row sums written to column T-1-t -> counter tick.
    The reference runs the U-Net twice per step on identical inputs (ddpm.py:199 and :418); here it runs once."""

    def __init__(self, ddpm: "DDPM", shape: Sequence[int]):
        super().__init__(ddpm, shape)
        dev = self.t_dev.device
        B = shape[0]
        self.tab = eval_coef_table({k: getattr(ddpm, k) for k in EVAL_TABLE_KEYS}).to(dev)
        self.x0 = torch.zeros(*shape, dtype=torch.float32, device=dev)
        self.vlb = torch.zeros(B, self.T, dtype=torch.float32, device=dev)
        self.sq = torch.zeros(B, self.T, dtype=torch.float32, device=dev)

    def step_eager(self) -> None:
        e = self.eng
        B, n = self.shape[0], self.shape[0] * self.chw
        L.call("dd_q_sample_step", L.ptr(self.x0), L.ptr(self.noise), n, self.period, L.ptr(self.tab), L.ptr(self.t_dev), self.T,
               L.ptr(e.x_in), B, self.chw, L.stream())
        e.run()
        L.call("dd_vlb_terms", L.ptr(self.x0), L.ptr(e.x_in), L.ptr(e.eps_out), L.ptr(self.noise), n, self.period, L.ptr(self.tab),
               L.ptr(self.t_dev), 0, self.T, L.ptr(self.vlb), L.ptr(self.sq), self.T, 1, B, self.chw, L.stream())
        L.call("dd_tick", L.ptr(self.t_dev), 1, L.stream())


class DDPM(nn.Module):
    def __init__(self, config: dict, latent_model: nn.Module, device: str, color_channels: int = 3):
        super().__init__()
        self.in_channels = color_channels
        self.latent_model = latent_model
        self.device = device
        self.image_size = config["image_size"]
        self.timesteps = config["T"]
        self.sample_shape = [self.in_channels, self.image_size, self.image_size]
        self.clip_denoised = True
        self.clip_range = (-1.0, 1.0)
        self.L = config["loss_type"]
        self.lambda_ = 0.0001
        assert self.L in OBJECTIVE_NAMES
        if config["loss_flat"] not in ("mean", "sum"):
            raise ValueError(f'Can only do mean or sum for flatten of loss, but {config["loss_flat"]} was desired..')
        self.loss_flat = config["loss_flat"]
        for name, buf in diffusion_buffers(config["beta_schedule"], self.timesteps).items():
            self.register_buffer(name, buf, persistent=(name != "vlb_weights"))
        assert not torch.isnan(self.vlb_weights).all()
        self._plans = EngineCache()
        self.use_graph = config.get("cuda_graph", True)

    # reference spelling of the loss helpers (ddpm.py:45-52): `flatten_loss(get_loss(target, output))`
    def get_loss(self, target: torch.Tensor, output: torch.Tensor) -> "ops.SquaredError":
        """partial(l2_loss, reduction='none') of ddpm.py:46: the per-element squared error, held as the operand pair so
        that `flatten_loss` / `.mean()` run the fused reduction kernel; `.tensor()` materialises it."""
        return ops.SquaredError(target, output)

    def flatten_loss(self, x) -> torch.Tensor:
        """reduce_mean / reduce_sum over all non-batch dimensions (ddpm.py:47-50, utils/utils.py:27-40)."""
        if isinstance(x, ops.SquaredError):
            return ops.mse_rows(x.target, x.output, self.loss_flat == "mean")
        return ops.reduce_rows(x, self.loss_flat == "mean")

    def mse_rows(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        """flatten_loss(get_loss(a, b)): per-sample sum/mean of squared error (ddpm.py:45-52, 279)."""
        return ops.mse_rows(a, b, self.loss_flat == "mean")

    # ---- forward process ---------------------------------------------------------------------
    def q_sample(self, x: torch.Tensor, t: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
        """x_t ~ q(x_t | x): ddpm.py:256-273."""
        assert x.shape == eps.shape
        return ops.q_sample(x, eps, t, self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod)

    def q_mean_variance(self, x, t):
        """ddpm.py:108-124 (evaluation helper; plain gather glue)."""
        shape = (x.shape[0],) + (1,) * (x.dim() - 1)
        mean = self.sqrt_alphas_cumprod.gather(-1, t).reshape(shape) * x
        var = (1.0 - self.alphas_cumprod).gather(-1, t).reshape(shape)
        logvar = self.log_one_minus_alphas_cumprod.gather(-1, t).reshape(shape)
        return mean, var, logvar

    # ---- reverse process -----------------------------------------------------------------------
    def predict_x_from_eps(self, x_t, t, eps, clip: bool = True):
        """ddpm.py:149-158."""
        assert x_t.shape == eps.shape
        return ops.predict_x0(x_t, eps, t, self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod, clip)

    def q_posterior(self, x, x_t, t):
        """ddpm.py:160-185: mean = c1*x0 + c2*x_t (q_sample's kernel), variance / clipped log-variance gathers."""
        assert x.shape == x_t.shape
        mean = ops.q_sample_raw(x, x_t, t, self.posterior_mean_coef1, self.posterior_mean_coef2)
        shape = (x.shape[0],) + (1,) * (x.dim() - 1)
        var = self.posterior_variance.gather(-1, t).reshape(shape)
        logvar = self.posterior_log_variance_clipped.gather(-1, t).reshape(shape)
        return mean, var, logvar

    def p_mean_variance(self, x_t, t):
        """ddpm.py:187-201."""
        eps_hat = self.latent_model(x_t, t)
        x_recon = self.predict_x_from_eps(x_t, t, eps_hat, clip=True)
        return self.q_posterior(x_recon, x_t, t)

    def _coef(self) -> torch.Tensor:
        c = getattr(self, "_coef_cache", None)
        if c is None or c.device != self.betas.device:
            c = posterior_coef_table({k: getattr(self, k) for k in (
                "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
                "posterior_mean_coef2", "posterior_log_variance_clipped")}).to(self.betas.device)
            self._coef_cache = c
        return c

    @torch.no_grad()
    def p_sample(self, x_t: torch.Tensor, t: torch.Tensor, repeat_noise: bool = False, noise: torch.Tensor = None):
        """One ancestral step with a per-sample `t`: ddpm.py:203-227 (U-Net, then one fused posterior kernel)."""
        x_t = x_t.contiguous().float()
        eps_hat = self.latent_model(x_t, t)
        if noise is None:
            if repeat_noise:
                noise = torch.randn((1, *x_t.shape[1:]), device=x_t.device).repeat(x_t.shape[0], *((1,) * (x_t.dim() - 1)))
            else:
                noise = torch.randn(x_t.shape, device=x_t.device)
        t32 = t.to(torch.int32).contiguous()
        return ops.posterior_step_raw(x_t, eps_hat.contiguous(), noise.contiguous().float(), self._coef(), t32, 1, 0,
                                      self.timesteps, 0, self.clip_denoised)

    def sampling_plan(self, shape) -> SamplingPlan:
        key = (tuple(shape), getattr(self.latent_model, "precision", None))
        plan = self._plans.get(key)
        if plan is None:
            plan = SamplingPlan(self, shape)
            self._plans[key] = plan
        return plan

    @torch.no_grad()
    def p_sample_loop(self, shape, every: int = 1, early_stop: int = None,
                      noise: Union[torch.Tensor, Sequence[torch.Tensor], None] = None) -> torch.Tensor:
        """ddpm.py:229-249.  `every` is accepted and ignored like the reference.

        noise: optional pre-drawn chain noise, (n_steps+1, B, C, H, W) tensor or sequence of tensors
        (device or pinned host): entry 0 is the start image, entry 1+k the k-th step's z.  When None the
        same draws are made with torch.randn on the device, in the reference's order (start image, then
        one draw per step including the masked one at t=0)."""
        shape = tuple(shape)
        T = self.timesteps
        t_end = 0 if early_stop is None else early_stop
        n_steps = T - t_end
        plan = self.sampling_plan(shape)
        plan.prepare()
        eng = plan.eng
        dev = eng.device

        def draw(k):      # k = 0: start image; k >= 1: noise of step k-1
            if noise is None:
                return torch.randn(shape, device=dev)
            return noise[k]

        eng.x_in.copy_(draw(0), non_blocking=True)
        plan.t_dev.fill_(T - 1)
        done = 0
        while done < n_steps:
            chunk = min(plan.period - (done % plan.period), n_steps - done)
            base = done % plan.period
            if noise is None:
                for j in range(chunk):
                    torch.randn(shape, out=plan.noise[base + j])      # same Philox draws as a fresh tensor, no copy launch
            elif isinstance(noise, torch.Tensor) and not noise.is_cuda and chunk > HOST_NOISE_BLOCK:
                # host noise: blocks of HOST_NOISE_BLOCK steps go up on a side stream while earlier blocks are consumed
                # (2.1 GB per C3 chain = ~40 ms of PCIe time that would otherwise sit in front of the first step)
                main = torch.cuda.current_stream(dev)
                side = plan.copy_stream
                side.wait_stream(main)                  # the ring slots being overwritten were consumed by earlier replays
                events = []
                for b0 in range(0, chunk, HOST_NOISE_BLOCK):
                    b1 = min(chunk, b0 + HOST_NOISE_BLOCK)
                    with torch.cuda.stream(side):
                        plan.noise[base + b0:base + b1].copy_(noise[1 + done + b0:1 + done + b1], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(side)
                    events.append((b0, b1, ev))
                for b0, b1, ev in events:
                    main.wait_event(ev)
                    for _ in range(b1 - b0):
                        if self.use_graph:
                            plan.graph.replay()
                        else:
                            plan.step_eager()
                L._Counter.n += chunk * plan.launches_per_step if self.use_graph else 0
                done += chunk
                continue
            elif isinstance(noise, torch.Tensor):
                plan.noise[base:base + chunk].copy_(noise[1 + done:1 + done + chunk], non_blocking=True)
            else:
                for j in range(chunk):
                    plan.noise[base + j].copy_(noise[1 + done + j], non_blocking=True)
            for _ in range(chunk):
                if self.use_graph:
                    plan.graph.replay()
                else:
                    plan.step_eager()
            L._Counter.n += chunk * plan.launches_per_step if self.use_graph else 0
            done += chunk
        return eng.x_in.clone()

    @torch.no_grad()
    def sample(self, batch_size: int = 16, every: int = 1, early_stop: int = None, noise=None) -> torch.Tensor:
        """ddpm.py:251-254."""
        return self.p_sample_loop((batch_size, *self.sample_shape), every, early_stop, noise=noise)

    @torch.no_grad()
    def reconstruct(self, x: torch.Tensor, n: int) -> torch.Tensor:
        """ddpm.py:126-147."""
        assert x.shape[0] >= n
        x = x[:n]
        t = torch.linspace(0, self.timesteps - 1, n, device=x.device, dtype=torch.long)
        eps = torch.randn_like(x)
        x_0 = self.q_sample(x, t, eps)
        eps_hat = self.latent_model(x_0, t)
        return self.predict_x_from_eps(x_0, t, eps_hat, clip=False)

    # ---- evaluation-side chain ----------------------------------------------------
    def _eval_tab(self) -> torch.Tensor:
        c = getattr(self, "_eval_tab_cache", None)
        if c is None or c.device != self.betas.device:
            c = eval_coef_table({k: getattr(self, k) for k in EVAL_TABLE_KEYS}).to(self.betas.device)
            self._eval_tab_cache = c
        return c

    @torch.no_grad()
    def vlb_terms(self, x: torch.Tensor, x_t: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """ddpm.py:317-365: L_t = KL(q(x_{t-1}|x_t,x) || p(x_{t-1}|x_t)) for t > 0, L_0 = -log p(x|x_1) (discretised
        Gaussian); (N,) in bits/dim.  U-Net, then one fused kernel (evaluation only: no gradient, like its one caller)."""
        assert x.shape == x_t.shape
        eps_hat = self.latent_model(x_t, t)
        return ops.vlb_terms_raw(x, x_t, eps_hat, self._eval_tab(), t.to(torch.int32).contiguous(), self.timesteps)

    @torch.no_grad()
    def calc_prior(self, x: torch.Tensor) -> torch.Tensor:
        """ddpm.py:367-389: the prior term L_T, (N,) in bits/dim."""
        return ops.prior_kl(x, float(self.sqrt_alphas_cumprod[-1]), float(self.log_one_minus_alphas_cumprod[-1]))

    def eval_plan(self, shape) -> EvalPlan:
        key = ("eval", tuple(shape), getattr(self.latent_model, "precision", None))
        plan = self._plans.get(key)
        if plan is None:
            plan = EvalPlan(self, shape)
            self._plans[key] = plan
        return plan

    @torch.no_grad()
    def test_losses_(self, x: torch.Tensor, noise: Union[torch.Tensor, Sequence[torch.Tensor], None] = None) -> dict:
        """ddpm.py:391-442: every term of the variational bound and of L_simple for t = T-1 .. 0, one graph replay per
        step.  noise: optional pre-drawn (T, N, C, H, W) eps, entry k belonging to t = T-1-k; when None the draws are
        made on the device in the reference's order (one randn_like(x) per step)."""
        x = x.contiguous().float()
        shape, T = tuple(x.shape), self.timesteps
        plan = self.eval_plan(shape)
        plan.prepare()
        dev = plan.eng.device
        plan.x0.copy_(x)
        plan.t_dev.fill_(T - 1)
        done = 0
        while done < T:
            base = done % plan.period
            chunk = min(plan.period - base, T - done)
            if noise is None:
                for j in range(chunk):
                    torch.randn(shape, out=plan.noise[base + j])      # same Philox draws as a fresh tensor, no copy launch
            elif isinstance(noise, torch.Tensor):
                plan.noise[base:base + chunk].copy_(noise[done:done + chunk], non_blocking=True)
            else:
                for j in range(chunk):
                    plan.noise[base + j].copy_(noise[done + j], non_blocking=True)
            for _ in range(chunk):
                if self.use_graph:
                    plan.graph.replay()
                else:
                    plan.step_eager()
            L._Counter.n += chunk * plan.launches_per_step if self.use_graph else 0
            done += chunk
        vlb_t = plan.vlb.clone()
        L_simple_t = plan.sq.sum(dim=0) / float(x.numel())           # get_loss(eps, eps_hat).mean() per step (ddpm.py:419)
        prior = self.calc_prior(x)
        return {"vlb_t": vlb_t, "prior": prior, "vlb": vlb_t.sum(dim=1) + prior, "L_simple_t": L_simple_t,
                "L_simple": L_simple_t.sum()}

    def test_losses(self, x: torch.Tensor, noise=None) -> dict:
        """ddpm.py:444-446."""
        return self.test_losses_(x, noise=noise)

    # ---- training objective --------------------------------------------------------------------
    def loss_ddpm(self, eps: torch.Tensor, eps_hat: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """ddpm.py:275-288."""
        loss = self.mse_rows(eps, eps_hat)
        if self.L == "simple":
            return loss.mean()
        if self.L == "vlb":
            return (self.vlb_weights[t] * loss).mean()
        return (loss + self.lambda_ * self.vlb_weights[t] * loss).mean()

    def losses(self, x: torch.Tensor, t: torch.Tensor, eps: torch.Tensor = None):
        """ddpm.py:290-315.  `eps` may be passed in for reproducible parity runs."""
        if eps is None:
            eps = torch.randn_like(x)
        x_t = self.q_sample(x, t, eps)
        eps_hat = self.latent_model(x_t, t)
        return self.loss_ddpm(eps, eps_hat, t)

    def p_losses(self, *args, **kwargs):
        """The name BASELINE.json's north_star uses for the training objective (= `losses`)."""
        return self.losses(*args, **kwargs)

    def t_sample(self, n: int) -> torch.Tensor:
        """ddpm.py:448-450."""
        return torch.randint(0, self.timesteps, (n,), device=self.device).long()

    def forward(self, x: torch.Tensor):
        """ddpm.py:452-457."""
        t = self.t_sample(x.shape[0])
        return self.losses(x, t)
