"""Parity at the BASELINE.json sizes (the shapes bench.py times), through size-independent properties plus oracle
checks on a subset of the batch: every sample of a batch is independent (GroupNorm per sample, LayerNorm per pixel,
attention per sample -- SURVEY.md 8(e)), so the CPU oracle only has to evaluate a few rows of a full-size batch.

  C3  dDDPM x3, latent 8x32x32, batch 64 (bf16 tensor-core sampling path)
  C4  dDDPM x3 256x256 training, batch 32 (TF32 tensor-core training path vs the fp32 validation mode)
  C5  full-resolution DDPM U-Net on 3x256x256 (bf16)
"""
import pytest
import torch

import downsampled_diffusion_b200 as dd
from oracle import ddpm_oracle as O
from tests import common as tc

pytestmark = pytest.mark.gpu


def test_c3_batch64_eps_subset_vs_oracle_and_batch_invariance(cuda):
    net = tc.build_model(dict(tc.C3, precision="bf16"), dd, "unet").to(cuda).eval()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    B = 64
    x = tc.randn(21, B, 8, 32, 32)
    t = (torch.arange(B) * 37 + 5) % 1000
    with torch.no_grad():
        eps = net(x.to(cuda), t.to(cuda)).cpu()
        rows = [0, 17, 42, 63]
        ref = O.unet_forward(sd, tc.C3, x[rows], t[rows])
        small = net(x[:4].to(cuda), t[:4].to(cuda)).cpu()           # a different batch size takes different tiles / splits
    assert torch.isfinite(eps).all()
    assert tc.rel_l2(eps[rows], ref) < 2e-2
    assert tc.rel_l2(eps[:4], small) < 2e-2


def test_c3_batch64_chain_tail_vs_oracle(cuda):
    cfg = dict(tc.C3, precision="bf16")
    m = tc.build_model(cfg, dd, "dddpm_ae", device="cuda").to(cuda).eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    buf = O.schedule_buffers("linear", cfg["T"])
    B, steps = 64, 6
    torch.manual_seed(5)
    noise = torch.stack([torch.randn(B, 8, 32, 32) for _ in range(steps + 1)])
    with torch.no_grad():               # the first `steps` ancestral steps (t = T-1 ... T-steps), as early_stop truncates the chain
        z = m.p_sample_loop((B, 8, 32, 32), early_stop=cfg["T"] - steps, noise=noise.to(cuda))
    rows = [0, 31, 63]
    zr = O.p_sample_loop(sd, cfg, buf, [n[rows] for n in noise], "latent_model.", t_end=cfg["T"] - steps)
    assert tc.max_abs(z.cpu()[rows], zr) < 5e-2
    with torch.no_grad():
        xi = m.rescaled_upsample(z[:2])
        xr = O.rescaled_upsample(sd, cfg, z[:2].cpu())
    assert xi.shape == (2, 3, 256, 256)
    assert tc.max_abs(xi, xr) < 2e-2


def test_c4_training_step_tf32_vs_fp32_mode(cuda):
    """Full-size training step: the TF32 tensor-core programs against the fp32 CUDA-core validation mode (which the
    golden vectors pin to the reference's autograd at small sizes) on the same weights, batch, t and eps."""
    B = 8                                       # per-sample independent; 8 x 3 x 256 x 256 keeps the fp32 mode quick
    x = tc.rand_pm1(31, B, 3, 256, 256).to(cuda)
    t = torch.tensor([0, 3, 50, 99, 100, 400, 700, 999], device=cuda)
    eps = tc.randn(32, B, 8, 32, 32).to(cuda)
    out = {}
    for prec in ("fp32", "bf16"):
        m = tc.build_model(dict(tc.C3, precision=prec), dd, "dddpm_ae", device="cuda").to(cuda).train()
        obj, d = m.losses(x, t, eps=eps)
        obj.backward()
        out[prec] = (float(obj), float(d["latent"]), float(d["recon"]), [None if p.grad is None else p.grad.double().norm().item() for p in m.parameters()],
                     m.upsample.conv[1].c2.weight.grad.detach().clone(), m.latent_model.mid_block1.block1.block[0].weight.grad.detach().clone())
        del m
    a, b = out["fp32"], out["bf16"]
    for i in range(3):
        assert abs(a[i] - b[i]) <= 2e-3 * abs(a[i]), (i, a[i], b[i])
    for ga, gb in zip(a[3], b[3]):
        if ga is None:
            assert gb is None
            continue
        assert abs(ga - gb) <= 3e-2 * max(ga, 1e-6) + 1e-5
    assert tc.rel_l2(b[4], a[4]) < 2e-2 and tc.rel_l2(b[5], a[5]) < 2e-2


def test_c5_fullres_unet_eps_vs_oracle(cuda):
    cfg = dict(tc.C3, image_size=256, unet_in=3, n_downsamples=0, precision="bf16")
    net = tc.build_model(cfg, dd, "unet").to(cuda).eval()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    x = tc.randn(41, 1, 3, 256, 256)
    t = torch.tensor([321])
    with torch.no_grad():
        eps = net(x.to(cuda), t.to(cuda)).cpu()
        ref = O.unet_forward(sd, cfg, x, t)
    assert tc.rel_l2(eps, ref) < 2e-2
