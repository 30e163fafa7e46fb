"""Evaluation-side chain (DDPM.vlb_terms / calc_prior / test_losses, DownsampleDDPM.test_losses) and the
sampler caller's output formatting (8(f).2: fix_samples) on the B200, against the golden vectors of the real reference
(oracle/make_golden_eval.py) and the CPU oracle.

Tolerances.  KL terms (t > 0) and the prior are smooth in the U-Net output: 2e-4 relative in the fp32 validation mode.
The t = 0 term is the discretised-Gaussian log-likelihood through a tanh CDF evaluated in fp32: where cdf_plus - cdf_min
cancels to a few ulps, tanhf of the device and of the host differ in the last bit and log() of the difference moves by
O(1) for that element (out of C*H*W averaged), so that term is compared at 1e-2 relative."""
import numpy as np
import pytest
import torch

import downsampled_diffusion_b200 as dd
from oracle import ddpm_oracle as O
from tests import common as tc

pytestmark = pytest.mark.gpu


def T_(a):
    return torch.from_numpy(np.asarray(a))


def eval_noise(seed, x, n):
    torch.manual_seed(seed)
    return torch.stack([torch.randn_like(x) for _ in range(n)])


def close(got, ref, rtol, msg=""):
    np.testing.assert_allclose(got.detach().cpu().numpy(), np.asarray(ref), rtol=rtol, err_msg=msg)


def test_vlb_terms_kernel_vs_oracle_and_golden(cuda, golden):
    m = tc.build_model(dict(tc.C1, precision="fp32"), dd, "ddpm", device="cuda").to(cuda).eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    buf = O.schedule_buffers("linear", 1000)
    x = tc.eval_images(71, 4, 1, 28, 28)
    eps = tc.randn(72, 4, 1, 28, 28)
    from downsampled_diffusion_b200 import ops
    for key, t in (("eval.c1.vlb_terms", torch.tensor([0, 1, 500, 999])), ("eval.c1.vlb_terms_t0", torch.zeros(4, dtype=torch.long))):
        x_t = O.q_sample(buf, x, t, eps)
        with torch.no_grad():
            eps_hat = O.unet_forward(sd, tc.C1, x_t, t, "latent_model.")
        ref = O.vlb_terms(buf, x, x_t, t, eps_hat)
        # the kernel alone, on the oracle's U-Net output
        got = ops.vlb_terms_raw(x.to(cuda), x_t.to(cuda), eps_hat.to(cuda), m._eval_tab(), t.to(cuda).int(), 1000).cpu()
        first = (t == 0)
        close(got[~first], ref[~first], 1e-5, key)
        close(got[first], ref[first], 1e-2, key)
        # the public method (U-Net on the device) against the reference's own output
        full = m.vlb_terms(x.to(cuda), x_t.to(cuda), t.to(cuda)).cpu()
        close(full[~first], golden[key][~first.numpy()], 2e-4, key)
        close(full[first], golden[key][first.numpy()], 1e-2, key)
    close(m.calc_prior(x.to(cuda)), golden["eval.c1.prior"], 2e-6)


@pytest.mark.parametrize("tag,cfg,kind,shape,seed,xseed,precision,rtol", [
    ("c1", tc.C1, "ddpm", (2, 1, 28, 28), 9, 73, "fp32", 2e-4),
    ("c1", tc.C1, "ddpm", (2, 1, 28, 28), 9, 73, "bf16", 6e-2),          # 28 -> 14 -> 7 maps on padded tensor-core tiles
    ("cs", tc.CS, "dddpm_ae", (2, 3, 32, 32), 10, 74, "fp32", 2e-4),
    ("cs", tc.CS, "dddpm_ae", (2, 3, 32, 32), 10, 74, "bf16", 6e-2)])
def test_evaluation_chain_vs_reference(cuda, golden, tag, cfg, kind, shape, seed, xseed, precision, rtol):
    cfg = dict(cfg, T=50, precision=precision)
    m = tc.build_model(cfg, dd, kind, device="cuda").to(cuda).eval()
    x = tc.eval_images(xseed, *shape)
    with torch.no_grad():
        z_shape = x if kind == "ddpm" else m.rescaled_downsample(x.to(cuda)).cpu()
    noise = eval_noise(seed, z_shape, 50)                      # the draws the reference made with randn_like on CPU
    got = m.test_losses(x.to(cuda), noise=noise.to(cuda))
    g = lambda k: golden[f"eval.{tag}.test_losses.{k}"]         # noqa: E731
    assert set(got.keys()) == {"vlb_t", "prior", "vlb", "L_simple_t", "L_simple"}
    assert got["vlb_t"].shape == (shape[0], 50) and got["L_simple_t"].shape == (50,)
    close(got["vlb_t"][:, :-1], g("vlb_t")[:, :-1], rtol, "KL terms")
    close(got["vlb_t"][:, -1], g("vlb_t")[:, -1], max(rtol, 1e-2), "L_0")
    close(got["prior"], g("prior"), max(rtol, 2e-6) if kind == "ddpm" else rtol, "prior")
    close(got["vlb"], g("vlb"), max(rtol, 2e-3), "vlb")
    close(got["L_simple_t"], g("L_simple_t"), rtol, "L_simple_t")
    close(got["L_simple"], g("L_simple"), rtol, "L_simple")
    # graph replay and the eager launch list agree: to the bit in fp32; the bf16 convolutions accumulate their GroupNorm
    # statistics with atomics, so there only to rounding
    m.use_graph = False
    again = m.test_losses(x.to(cuda), noise=noise.to(cuda))
    if precision == "fp32":
        assert all(torch.equal(got[k], again[k]) for k in got)
    else:
        assert all(torch.allclose(got[k], again[k], rtol=2e-2) for k in got)


def test_evaluation_chain_device_rng_and_host_noise(cuda):
    cfg = dict(tc.C1, T=20, precision="fp32")
    m = tc.build_model(cfg, dd, "ddpm", device="cuda").to(cuda).eval()
    x = tc.eval_images(5, 3, 1, 28, 28).to(cuda)
    torch.manual_seed(3)
    a = m.test_losses(x)                                      # draws made on the device, one randn per step
    torch.manual_seed(3)
    noise = [torch.randn(3, 1, 28, 28, device=cuda) for _ in range(20)]
    b = m.test_losses(x, noise=noise)                         # the same draws handed in as a sequence
    c = m.test_losses(x, noise=torch.stack(noise).cpu())      # ... and as one host tensor
    assert all(torch.equal(a[k], b[k]) and torch.equal(a[k], c[k]) for k in a)


def test_c3_evaluation_chain_full_size_properties(cuda):
    """BASELINE size (latent 64x8x32x32, T = 1000, bf16 U-Net): consistency of the chain's outputs, non-negative KL terms,
    closed-form prior, and the chained kernels against the per-step public method on a few steps."""
    cfg = dict(tc.C3, precision="bf16")
    m = tc.build_model(cfg, dd, "dddpm_ae", device="cuda").to(cuda).eval()
    B, T = 64, 1000
    z = tc.eval_images(6, B, 8, 32, 32).to(cuda)
    torch.manual_seed(11)
    ring = torch.randn(8, B, 8, 32, 32, device=cuda)

    class Cyc:                                               # 1000 steps of noise from a ring of 8 draws (262 MB instead of 2 GB)
        def __getitem__(self, k):
            return ring[k % 8]
    out = m.test_losses_(z, noise=Cyc())
    assert all(torch.isfinite(v).all() for v in out.values())
    assert (out["vlb_t"][:, :-1] >= 0).all()
    assert torch.allclose(out["vlb"], out["vlb_t"].sum(dim=1) + out["prior"]) and torch.allclose(out["L_simple"], out["L_simple_t"].sum())
    buf = O.schedule_buffers("linear", T)
    close(out["prior"].cpu(), O.calc_prior(buf, z.cpu(), T), 2e-6)
    for k in (0, 499, 998, 999):                             # step k <-> t = T-1-k
        t = torch.full((B,), T - 1 - k, device=cuda, dtype=torch.long)
        z_t = m.q_sample(z, t, ring[k % 8])
        direct = m.vlb_terms(z, z_t, t)
        close(out["vlb_t"][:, k], direct.cpu(), 2e-2 if k != 999 else 5e-2, f"step {k}")   # per-sample-t U-Net path vs time-table path (bf16)


def test_fix_samples_bit_exact(cuda, golden):
    s = tc.randn(75, 3, 3, 32, 32)
    out = dd.fix_samples(s.to(cuda))
    assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == (3, 32, 32, 3)
    assert np.array_equal(out, golden["fix_samples.out"])
    big = tc.randn(76, 4, 3, 256, 256)                        # C3 output size; odd element count per image in the second case
    assert np.array_equal(dd.fix_samples(big.to(cuda)), O.fix_samples(big))
    odd = tc.randn(77, 2, 1, 5, 7)
    assert np.array_equal(dd.fix_samples(odd.to(cuda)), O.fix_samples(odd))


def test_generate_samples_loop(cuda):
    """generate_model_samples.py:41-51: batches of `sample` -> `fix_samples`, with the host copies overlapped."""
    cfg = dict(tc.CS, T=20, precision="fp32")
    m = tc.build_model(cfg, dd, "dddpm_ae", device="cuda").to(cuda).eval()
    torch.manual_seed(4)
    xs, zs = dd.generate_samples(m, 5, 2)
    assert len(xs) == len(zs) == 3 and xs[0].shape == (2, 32, 32, 3) and zs[0].shape == (2, 8, 8, 8)
    torch.manual_seed(4)
    for k in range(3):
        x, z = m.sample(2)
        assert np.array_equal(xs[k], O.fix_samples(x.cpu())) and np.array_equal(zs[k], O.fix_samples(z.cpu()))
    assert all(a.min() == 0.0 and a.max() == 255.0 for a in xs + zs)
    plain = tc.build_model(dict(tc.C1, T=20, precision="fp32"), dd, "ddpm", device="cuda").to(cuda).eval()
    xs, zs = dd.generate_samples(plain, 3, 3)
    assert len(xs) == 1 and zs == [] and xs[0].shape == (3, 28, 28, 1)
