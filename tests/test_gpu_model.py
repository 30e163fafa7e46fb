"""Model-level parity on the B200 through the public (reference-shaped) Python API: U-Net epsilon
prediction, single ancestral step, full chains, resampling nets -- against the golden vectors the real
reference produced and against the CPU oracle on the same seeded weights / inputs / noise.

Tolerances (BASELINE.json north_star): eps_hat relative L2 <= 1e-4 in fp32 validation mode, <= 2e-2 in
bf16; full chain latent max-abs <= 5e-2 / image max-abs <= 2e-2 in bf16 and <= 1e-3 in fp32 mode."""
import numpy as np
import pytest
import torch

import downsampled_diffusion_b200 as dd
from oracle import ddpm_oracle as O
from tests import common as tc

pytestmark = pytest.mark.gpu

TS = (torch.tensor([999, 0]), torch.tensor([500, 37]))


def G(golden, key):
    return torch.from_numpy(np.asarray(golden[key]))


def chain_noise(seed, shape, n):
    torch.manual_seed(seed)
    return torch.stack([torch.randn(shape) for _ in range(n + 1)])


@pytest.mark.parametrize("tag,cfg,hw,seed", [("c3", tc.C3, 32, 11), ("c2", tc.C2, 16, 14), ("cs", tc.CS, 8, 13),
                                             ("c1", tc.C1, 28, 12)])
def test_unet_eps_fp32(cuda, golden, tag, cfg, hw, seed):
    net = tc.build_model(dict(cfg, precision="fp32"), dd, "unet").to(cuda).eval()
    x = tc.randn(seed, 2, cfg["unet_in"], hw, hw).to(cuda)
    for j, t in enumerate(TS):
        with torch.no_grad():
            eps = net(x, t.to(cuda))
        assert eps.shape == x.shape and eps.dtype == torch.float32
        assert tc.rel_l2(eps, G(golden, f"unet.{tag}.eps{j}")) < 1e-4


@pytest.mark.parametrize("tag,cfg,hw,seed", [("c3", tc.C3, 32, 11), ("c2", tc.C2, 16, 14), ("cs", tc.CS, 8, 13),
                                             ("c1", tc.C1, 28, 12)])       # c1: 28 -> 14 -> 7 maps on padded tensor-core tiles
def test_unet_eps_bf16(cuda, golden, tag, cfg, hw, seed):
    net = tc.build_model(dict(cfg, precision="bf16"), dd, "unet").to(cuda).eval()
    x = tc.randn(seed, 2, cfg["unet_in"], hw, hw).to(cuda)
    for j, t in enumerate(TS):
        with torch.no_grad():
            eps = net(x, t.to(cuda))
        err = tc.rel_l2(eps, G(golden, f"unet.{tag}.eps{j}"))
        print(f"bf16 eps rel-L2 {tag}/{j}: {err:.3e}")
        assert err < 2e-2


def test_bf16_rejects_untileable_shapes(cuda):
    net = tc.build_model(dict(tc.C1, unet_chan=48, precision="bf16"), dd, "unet").to(cuda).eval()
    with pytest.raises((ValueError, RuntimeError)):                   # 48-channel operands have no 128-byte K rows
        net(torch.zeros(2, 1, 28, 28, device=cuda), torch.zeros(2, dtype=torch.long, device=cuda))
    net = tc.build_model(dict(tc.C1, precision="bf16"), dd, "unet").to(cuda).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(2, 1, 28, 28), torch.zeros(2, dtype=torch.long))
    with pytest.raises(ValueError, match="divisible"):
        tc.build_model(dict(tc.C1, precision="fp32"), dd, "unet").to(cuda)(
            torch.zeros(2, 1, 30, 30, device=cuda), torch.zeros(2, dtype=torch.long, device=cuda))


def test_p_sample_and_chains_c1_fp32(cuda, golden):
    cfg = dict(tc.C1, T=50, precision="fp32")
    m = tc.build_model(cfg, dd, "ddpm", device="cuda").to(cuda).eval()
    x = tc.randn(21, 2, 1, 28, 28).to(cuda)
    torch.manual_seed(22)
    z = torch.randn(2, 1, 28, 28)
    out = m.p_sample(x, torch.tensor([30, 0], device=cuda), noise=z.to(cuda))
    assert tc.max_abs(out, G(golden, "p_sample.c1.out")) < 1e-4
    full = m.sample(2, noise=chain_noise(5, (2, 1, 28, 28), 50).to(cuda))
    assert tc.max_abs(full, G(golden, "chain.c1.x")) < 1e-3
    early = m.sample(2, early_stop=40, noise=chain_noise(6, (2, 1, 28, 28), 10).to(cuda))
    assert tc.max_abs(early, G(golden, "chain.c1.early")) < 1e-3
    # graph replay and eager launches are the same program
    m.use_graph = False
    eager = m.sample(2, noise=chain_noise(5, (2, 1, 28, 28), 50).to(cuda))
    assert torch.equal(eager, full)


def test_chain_c1_bf16(cuda, golden):
    """The 50-step C1 chains of the fp32 test above on the bf16 tensor-core path (ragged 28 / 14 / 7 maps): bar 5e-2."""
    m = tc.build_model(dict(tc.C1, T=50, precision="bf16"), dd, "ddpm", device="cuda").to(cuda).eval()
    full = m.sample(2, noise=chain_noise(5, (2, 1, 28, 28), 50).to(cuda))
    early = m.sample(2, early_stop=40, noise=chain_noise(6, (2, 1, 28, 28), 10).to(cuda))
    e1, e2 = tc.max_abs(full, G(golden, "chain.c1.x")), tc.max_abs(early, G(golden, "chain.c1.early"))
    print(f"bf16 c1 chains: full max-abs {e1:.3e}, early-stop max-abs {e2:.3e}")
    assert e1 < 5e-2 and e2 < 5e-2


@pytest.mark.parametrize("precision,tol_z,tol_x", [("fp32", 1e-3, 1e-3), ("bf16", 5e-2, 2e-2)])
def test_dddpm_chain_cs(cuda, golden, precision, tol_z, tol_x):
    cfg = dict(tc.CS, T=50, precision=precision)
    m = tc.build_model(cfg, dd, "dddpm_ae", device="cuda").to(cuda).eval()
    m.downsample.precision = m.upsample.precision = precision
    noise = chain_noise(5, (2, 8, 8, 8), 50)
    x, z = m.sample(2, noise=noise.pin_memory())          # host (pinned) noise: the e2e entry
    ez, ex = tc.max_abs(z, G(golden, "chain.cs.z")), tc.max_abs(x, G(golden, "chain.cs.x"))
    print(f"{precision} chain: latent max-abs {ez:.3e}, image max-abs {ex:.3e}")
    assert ez < tol_z and ex < tol_x
    assert x.shape == (2, 3, 32, 32) and z.shape == (2, 8, 8, 8)


def test_chain_device_rng_order(cuda):
    """Without pre-drawn noise the chain draws torch.randn on the device in the reference's order
    (start image, then one draw per step): reproducing those draws by hand gives the same sample."""
    cfg = dict(tc.CS, T=50, precision="fp32")
    m = tc.build_model(cfg, dd, "ddpm", device="cuda").to(cuda).eval()
    m.sample_shape = [8, 8, 8]
    torch.manual_seed(3)
    a = m.sample(2)
    torch.manual_seed(3)
    noise = torch.stack([torch.randn((2, 8, 8, 8), device=cuda) for _ in range(51)])
    b = m.sample(2, noise=noise)
    assert torch.equal(a, b)


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 3e-2)])
def test_resample_nets_c2(cuda, golden, precision, tol):
    m = tc.build_model(tc.C2, dd, "dddpm_ae", device="cuda").to(cuda).eval()
    m.downsample.precision = m.upsample.precision = precision
    x = tc.rand_pm1(31, 2, 3, 64, 64).to(cuda)
    with torch.no_grad():
        z = m.rescaled_downsample(x)
        xh = m.rescaled_upsample(G(golden, "resample.c2.z").to(cuda))
    assert tc.max_abs(z, G(golden, "resample.c2.z")) < tol
    assert tc.max_abs(xh, G(golden, "resample.c2.xhat")) < tol


def test_resample_convolutional_mode(cuda, golden):
    cfg = dict(tc.CS, d_mode="convolutional", u_mode="convolutional")
    m = tc.build_model(cfg, dd, "dddpm", device="cuda").to(cuda).eval()
    m.downsample.precision = m.upsample.precision = "fp32"
    x = tc.rand_pm1(32, 2, 3, 32, 32).to(cuda)
    with torch.no_grad():
        z = m.rescaled_downsample(x)
        xh = m.rescaled_upsample(z)
    assert tc.max_abs(z, G(golden, "resample.convolutional.z")) < 2e-5
    assert tc.max_abs(xh, G(golden, "resample.convolutional.xhat")) < 2e-5
    with pytest.raises(NotImplementedError):
        dd.get_upsampling(dict(cfg, u_mode="nope"), (3, 32, 32))


def test_state_dict_roundtrip_and_ema_proxy(cuda, golden):
    cfg = dict(tc.CS, T=50, precision="fp32")
    a = tc.build_model(cfg, dd, "dddpm_ae", device="cuda").to(cuda).eval()
    b = tc.build_model(cfg, dd, "dddpm_ae", seed=9, device="cuda").to(cuda).eval()
    x = tc.randn(13, 2, 8, 8, 8).to(cuda)
    t = torch.tensor([10, 3], device=cuda)
    with torch.no_grad():
        before = b.latent_model(x, t)
        b.load_state_dict(a.state_dict())                   # packed-weight caches must be refreshed
        assert torch.equal(b.latent_model(x, t), a.latent_model(x, t)) and not torch.equal(before, a.latent_model(x, t))
    ema = dd.EMA(a, decay=0.5)
    ema.eval()
    with torch.no_grad():
        for p in a.parameters():
            p.mul_(1.5)
    ema.update(a)
    assert set(ema.state_dict().keys()) == set(a.state_dict().keys())
    noise = chain_noise(5, (2, 8, 8, 8), 50).to(cuda)
    xs, zs = ema.ema_model.sample(2, noise=noise)
    assert torch.isfinite(xs).all() and xs.abs().max() <= 1.0
