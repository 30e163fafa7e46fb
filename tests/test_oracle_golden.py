"""Pins the CPU oracle (oracle/ddpm_oracle.py) against golden vectors produced by the real reference
(oracle/make_golden.py).  CPU only; the GPU parity tests then use the oracle as their checker."""
import numpy as np
import pytest
import torch

import downsampled_diffusion_b200 as ours
from oracle import ddpm_oracle as O
from tests import common as tc


def T(a):
    return torch.from_numpy(np.asarray(a))


def sd_of(cfg, kind, seed=0):
    return {k: v.clone() for k, v in tc.build_model(cfg, ours, kind, seed).state_dict().items()}


@pytest.mark.parametrize("sched", ["linear", "cosine"])
def test_schedule_buffers_bit_exact(golden, sched):
    buf = O.schedule_buffers(sched, 1000)
    for k, v in buf.items():
        assert torch.equal(v, T(golden[f"sched.{sched}.{k}"])), k
    # the package's own schedule (host-side product code) must match to the bit as well
    from downsampled_diffusion_b200.schedule import diffusion_buffers
    for k, v in diffusion_buffers(sched, 1000).items():
        assert torch.equal(v, T(golden[f"sched.{sched}.{k}"])), k


def test_schedule_known_values():
    buf = O.schedule_buffers("linear", 1000)
    assert abs(float(buf["betas"][0]) - 1e-4) < 1e-10 and abs(float(buf["betas"][-1]) - 0.02) < 1e-8
    assert buf["posterior_log_variance_clipped"][0] == buf["posterior_log_variance_clipped"][1]
    assert float(buf["posterior_mean_coef1"][0]) == 1.0 and float(buf["posterior_mean_coef2"][0]) == 0.0
    with pytest.raises(ValueError):
        O.beta_schedule("quadratic", 10)


@pytest.mark.parametrize("tag,cfg,hw,seed", [("c3", tc.C3, 32, 11), ("c1", tc.C1, 28, 12), ("cs", tc.CS, 8, 13),
                                             ("c2", tc.C2, 16, 14)])
def test_unet_eps(golden, tag, cfg, hw, seed):
    sd = sd_of(cfg, "unet")
    x = tc.randn(seed, 2, cfg["unet_in"], hw, hw)
    for j, t in enumerate((torch.tensor([999, 0]), torch.tensor([500, 37]))):
        with torch.no_grad():
            eps = O.unet_forward(sd, cfg, x, t)
        assert tc.rel_l2(eps, T(golden[f"unet.{tag}.eps{j}"])) < 2e-6


def chain_noise(seed, shape, n):
    torch.manual_seed(seed)
    return [torch.randn(shape) for _ in range(n + 1)]


def test_p_sample_and_chain_c1(golden):
    cfg = dict(tc.C1, T=50)
    sd = sd_of(cfg, "ddpm")
    buf = O.schedule_buffers("linear", 50)
    x = tc.randn(21, 2, 1, 28, 28)
    t = torch.tensor([30, 0])
    torch.manual_seed(22)
    z = torch.randn(x.shape)
    with torch.no_grad():
        out = O.posterior_step(buf, x, t, O.unet_forward(sd, cfg, x, t, "latent_model."), z)
        assert tc.max_abs(out, T(golden["p_sample.c1.out"])) < 1e-5
        full = O.p_sample_loop(sd, cfg, buf, chain_noise(5, (2, 1, 28, 28), 50))
        assert tc.max_abs(full, T(golden["chain.c1.x"])) < 2e-4
        early = O.p_sample_loop(sd, cfg, buf, chain_noise(6, (2, 1, 28, 28), 10), t_end=40)
        assert tc.max_abs(early, T(golden["chain.c1.early"])) < 2e-4


def test_dddpm_chain_cs(golden):
    cfg = dict(tc.CS, T=50)
    sd = sd_of(cfg, "dddpm_ae")
    buf = O.schedule_buffers("linear", 50)
    with torch.no_grad():
        x, z = O.dddpm_sample(sd, cfg, buf, chain_noise(5, (2, 8, 8, 8), 50))
    assert tc.max_abs(z, T(golden["chain.cs.z"])) < 2e-4
    assert tc.max_abs(x, T(golden["chain.cs.x"])) < 2e-4


def test_resample_nets(golden):
    sd = sd_of(tc.C2, "dddpm_ae")
    x = tc.rand_pm1(31, 2, 3, 64, 64)
    with torch.no_grad():
        z = O.rescaled_downsample(sd, tc.C2, x)
        xh = O.rescaled_upsample(sd, tc.C2, z)
    assert tc.max_abs(z, T(golden["resample.c2.z"])) < 1e-5
    assert tc.max_abs(xh, T(golden["resample.c2.xhat"])) < 1e-5


def test_diffusion_arithmetic_bit_exact(golden):
    buf = O.schedule_buffers("linear", 1000)
    x, e = tc.randn(41, 4, 1, 28, 28), tc.randn(42, 4, 1, 28, 28)
    t = torch.tensor([0, 1, 500, 999])
    assert torch.equal(O.q_sample(buf, x, t, e), T(golden["ddpm.q_sample"]))
    assert torch.equal(O.predict_x_from_eps(buf, x, t, e, True), T(golden["ddpm.predict_x0.clip"]))
    assert torch.equal(O.predict_x_from_eps(buf, x, t, e, False), T(golden["ddpm.predict_x0.noclip"]))
    mean, var, logvar = O.q_posterior(buf, e.clamp(-1, 1), x, t)
    assert torch.equal(mean, T(golden["ddpm.q_posterior.mean"]))
    assert torch.equal(var, T(golden["ddpm.q_posterior.var"]))
    assert torch.equal(logvar, T(golden["ddpm.q_posterior.logvar"]))


@pytest.mark.parametrize("kind", ["dddpm_ae", "dddpm"])
def test_training_objective(golden, kind):
    sd = sd_of(tc.CS, kind)
    buf = O.schedule_buffers("linear", 1000)
    x = tc.rand_pm1(51, 4, 3, 32, 32)
    t = torch.tensor([3, 50, 99, 700])
    torch.manual_seed(7)
    eps = torch.randn(4, 8, 8, 8)
    with torch.no_grad():
        obj, d = O.dddpm_losses(sd, tc.CS, buf, x, t, eps, autoencoder=(kind == "dddpm_ae"))
    assert abs(float(obj) - float(golden[f"loss.{kind}.obj"])) <= 2e-5 * abs(float(golden[f"loss.{kind}.obj"]))
    assert abs(float(d["latent"]) - float(golden[f"loss.{kind}.latent"])) <= 2e-5 * abs(float(golden[f"loss.{kind}.latent"]))
    assert abs(float(d["recon"]) - float(golden[f"loss.{kind}.recon"])) <= 2e-5 * abs(float(golden[f"loss.{kind}.recon"])) + 1e-7


@pytest.mark.parametrize("lt,lf", [("vlb", "sum"), ("hybrid", "mean"), ("simple", "mean")])
def test_ddpm_objective_variants(golden, lt, lf):
    cfg = dict(tc.C1, loss_type=lt, loss_flat=lf)
    sd = sd_of(cfg, "ddpm")
    buf = O.schedule_buffers("linear", 1000)
    x = tc.rand_pm1(52, 4, 1, 28, 28)
    t = torch.tensor([0, 10, 400, 999])
    torch.manual_seed(8)
    eps = torch.randn(x.shape)
    with torch.no_grad():
        obj = O.ddpm_losses(sd, cfg, buf, x, t, eps)
    ref = float(golden[f"loss.c1.{lt}.{lf}.obj"])
    assert abs(float(obj) - ref) <= 2e-5 * abs(ref)


def test_ema(golden):
    net = tc.build_model(tc.CS, ours, "unet")
    shadow = [p.detach().clone() for p in net.parameters()]
    names = [n for n, _ in net.named_parameters()]
    for k in range(3):
        g = torch.Generator().manual_seed(60 + k)
        with torch.no_grad():
            for p in net.parameters():
                p.add_(0.01 * torch.randn(p.shape, generator=g))
        shadow = O.ema_update(shadow, [p.detach() for p in net.parameters()], 0.995)
    got = dict(zip(names, shadow))
    for n in ("final_conv.1.weight", "downs.0.0.block1.block.0.bias", "mid_attn.fn.norm.g", "time_mlp.3.weight"):
        assert torch.equal(got[n], T(golden[f"ema.{n}"])), n


# ---- evaluation-side chain, output formatting (golden_v2 from oracle/make_golden_eval.py) ----------
def eval_noise(seed, x, n):
    torch.manual_seed(seed)
    return [torch.randn_like(x) for _ in range(n)]


def test_vlb_terms_and_prior(golden):
    sd = sd_of(tc.C1, "ddpm")
    buf = O.schedule_buffers("linear", 1000)
    x = tc.eval_images(71, 4, 1, 28, 28)
    eps = tc.randn(72, 4, 1, 28, 28)
    for key, t in (("eval.c1.vlb_terms", torch.tensor([0, 1, 500, 999])), ("eval.c1.vlb_terms_t0", torch.zeros(4, dtype=torch.long))):
        x_t = O.q_sample(buf, x, t, eps)
        with torch.no_grad():
            eps_hat = O.unet_forward(sd, tc.C1, x_t, t, "latent_model.")
        got = O.vlb_terms(buf, x, x_t, t, eps_hat)
        np.testing.assert_allclose(got.numpy(), golden[key], rtol=2e-5)
    assert torch.equal(O.calc_prior(buf, x, 1000), T(golden["eval.c1.prior"]))


@pytest.mark.parametrize("tag,cfg,kind,shape,seed,xseed", [("c1", tc.C1, "ddpm", (2, 1, 28, 28), 9, 73),
                                                          ("cs", tc.CS, "dddpm_ae", (2, 3, 32, 32), 10, 74)])
def test_evaluation_chain(golden, tag, cfg, kind, shape, seed, xseed):
    cfg = dict(cfg, T=50)
    sd = sd_of(cfg, kind)
    buf = O.schedule_buffers("linear", 50)
    x = tc.eval_images(xseed, *shape)
    with torch.no_grad():
        z = O.rescaled_downsample(sd, cfg, x) if kind != "ddpm" else x          # dddpm.py:146-148
        got = O.test_losses(sd, cfg, buf, z, eval_noise(seed, z, 50))
    for k, v in got.items():
        np.testing.assert_allclose(v.numpy(), golden[f"eval.{tag}.test_losses.{k}"], rtol=3e-5, err_msg=k)


def test_fix_samples(golden):
    out = O.fix_samples(tc.randn(75, 3, 3, 32, 32))
    assert out.shape == (3, 32, 32, 3) and np.array_equal(out, golden["fix_samples.out"])


def test_trainer_optimizer_step(golden):
    """clip_grad_norm_(1.0) + Adam(lr=2e-4) of the reference's trainer, two steps on synthetic gradients (golden_v2)."""
    net = tc.build_model(tc.CS, ours, "unet")
    names = [n for n, _ in net.named_parameters()]
    ps = [p.detach().clone() for p in net.parameters()]
    ms, vs = [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    for step in range(2):
        g = torch.Generator().manual_seed(80 + step)
        grads = [(0.05 if step == 0 else 0.0005) * torch.randn(p.shape, generator=g) for p in ps]
        total, grads = O.clip_grad_norm(grads, 1.0)
        np.testing.assert_allclose(float(total), float(golden[f"optim.norm{step}"]), rtol=1e-6)
        for i in range(len(ps)):
            ps[i], ms[i], vs[i] = O.adam_step(ps[i], grads[i], ms[i], vs[i], step + 1, 2e-4)
        for n in ("final_conv.1.weight", "downs.0.0.block1.block.0.bias", "mid_attn.fn.norm.g", "time_mlp.3.weight"):
            ref = T(golden[f"optim.step{step}.{n}"])
            assert (ps[names.index(n)] - ref).abs().max() <= 2.4e-7 * max(1.0, float(ref.abs().max())), n
    assert float(golden["optim.norm0"]) > 1.0 > float(golden["optim.norm1"])      # step 0 was clipped, step 1 was not


def test_deterministic_bicubic_resampler(golden):
    """'deterministic' mode (wrapper.py:22-24, 49-53): the bicubic restatement against the reference's F.interpolate."""
    cfg = dict(tc.CS, d_mode="deterministic", u_mode="deterministic", unet_in=3)
    x = tc.rand_pm1(85, 4, 3, 32, 32)
    z = O.rescaled_downsample({}, cfg, x)
    assert (z - T(golden["det.z"])).abs().max() < 2e-6
    assert (O.rescaled_upsample({}, cfg, z) - T(golden["det.xhat"])).abs().max() < 2e-6
    # the objective of the non-autoencoder dDDPM through the bicubic up-sampler, with autograd on the restatement
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd_of(cfg, "dddpm").items()}
    buf = O.schedule_buffers("linear", 1000)
    t = torch.tensor([3, 50, 99, 700])
    torch.manual_seed(12)
    eps = torch.randn(4, 3, 8, 8)
    obj, d = O.dddpm_losses(sd, cfg, buf, x, t, eps, autoencoder=False)
    obj.backward()
    np.testing.assert_allclose(float(obj), float(golden["det.loss.obj"]), rtol=2e-5)
    np.testing.assert_allclose(float(d["recon"]), float(golden["det.loss.recon"]), rtol=2e-5)
    assert tc.rel_l2(sd["latent_model.final_conv.1.weight"].grad, T(golden["det.loss.grad_final"])) < 1e-4


# ---- full T=1000 chains at the BASELINE sizes (golden_v3.npz, oracle/make_golden_chain.py) -----------------------
@pytest.mark.parametrize("tag,cfg,hw,rows", [("c2", tc.C2, 16, 2), ("c3", tc.C3, 32, 1)])
def test_full_chain_oracle_vs_reference(golden, tag, cfg, hw, rows):
    """The oracle's decomposed loop over all 1000 steps lands on what the unmodified reference's `sample()` returned
    for the same pre-drawn noise (first `rows` rows: every row's chain is independent; bounded so the CPU suite stays short)."""
    sd = sd_of(cfg, "dddpm_ae")
    buf = O.schedule_buffers("linear", cfg["T"])
    noise = tc.chain_noise(tag, cfg["T"], tc.CHAIN_ROWS, cfg["unet_in"], hw, hw)[:, :rows]
    with torch.no_grad():
        x, z = O.dddpm_sample(sd, cfg, buf, list(noise))
    zr, xr = T(golden[f"fullchain.{tag}.z"])[:rows], T(golden[f"fullchain.{tag}.x"])[:rows]
    if xr.shape[-1] != x.shape[-1]:
        x = x[:, :, ::2, ::2]
    assert tc.max_abs(z, zr) < 2e-4 and tc.max_abs(x, xr) < 2e-4


def test_c1_full_chain_oracle_vs_reference(golden):
    """BASELINE configs[0] (standard DDPM, 1x28x28, T = 1000, batch 16; golden_v4.npz, oracle/make_golden_c1.py): the oracle's
    loop on the first two rows against the unmodified reference's `sample()`."""
    cfg, rows = tc.C1, 2
    sd = sd_of(cfg, "ddpm")
    buf = O.schedule_buffers("linear", cfg["T"])
    noise = tc.chain_noise("c1", cfg["T"], tc.C1_CHAIN_BATCH, 1, 28, 28)[:, :rows]
    with torch.no_grad():
        x = O.p_sample_loop(sd, cfg, buf, list(noise))
    assert tc.max_abs(x, T(golden["fullchain.c1.x"])[:rows]) < 2e-4


def test_full_size_training_objective_oracle_vs_reference(golden):
    cfg = tc.C3
    model = tc.build_model(cfg, ours, "dddpm_ae")
    names = [n for n, _ in model.named_parameters()]
    sd = {k: v.detach().clone().requires_grad_(k in names) for k, v in model.state_dict().items()}
    buf = O.schedule_buffers("linear", cfg["T"])
    x, t, eps = tc.rand_pm1(31, 2, 3, 256, 256), torch.tensor([50, 700]), tc.randn(32, 2, 8, 32, 32)
    obj, d = O.dddpm_losses(sd, cfg, buf, x, t, eps, autoencoder=True)
    obj.backward()
    for key, val in (("obj", obj), ("latent", d["latent"]), ("recon", d["recon"])):
        ref = float(golden[f"fulltrain.c4.{key}"])
        assert abs(float(val) - ref) <= 1e-5 * abs(ref) + 1e-7, key
    norms = np.asarray(golden["fulltrain.c4.grad_norms"])
    for n, ref in zip(names, norms):
        g = sd[n].grad
        got = 0.0 if g is None else float(g.double().norm())
        assert abs(got - ref) <= 1e-4 * max(ref, 1e-6) + 1e-7, n
