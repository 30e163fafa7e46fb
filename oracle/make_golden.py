"""Generate tests/golden/golden_v1.npz from the REAL reference (simonamtoft/downsampled-diffusion).

TEST INFRASTRUCTURE.  Runs only in the authoring container, where /root/reference exists:
    python oracle/make_golden.py
The reference is imported unmodified (with `utils.evaluator` stubbed: it needs TensorFlow, see
SURVEY.md 8(c)), fed the seeded weights/inputs of tests/common.py, and its outputs are stored.
The GPU box has no /root/reference; tests there compare against these vectors and the oracle.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("DD_REFERENCE", "/root/reference")


def import_reference():
    sys.path.insert(0, REF)
    stub = types.ModuleType("utils.evaluator")
    stub.Evaluator = None
    sys.modules["utils.evaluator"] = stub
    import models as ref_models                     # noqa: E402  (the reference's top-level package)
    from trainers.ema import EMA as RefEMA          # noqa: E402
    ref_models.EMA = RefEMA
    return ref_models


def main():
    ref = import_reference()
    import downsampled_diffusion_b200 as ours
    from tests import common as tc

    torch.set_num_threads(8)
    out = {}

    def put(name, t):
        out[name] = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)

    def ref_model(cfg, kind, seed=0):
        """Reference model carrying exactly the weights tests/common.py generates for our package."""
        mine = tc.build_model(cfg, ours, kind, seed)
        theirs = tc.build_model(cfg, ref, kind, seed)
        sd_m, sd_t = mine.state_dict(), theirs.state_dict()
        assert list(sd_m.keys()) == list(sd_t.keys()), "state_dict keys differ from the reference"
        for k in sd_m:
            assert torch.equal(sd_m[k], sd_t[k]), f"init differs from the reference at {k}"
        theirs.eval()
        return theirs

    # ---- 1. schedules -----------------------------------------------------------------------
    for sched in ("linear", "cosine"):
        cfg = dict(tc.C1, beta_schedule=sched)
        m = ref.DDPM(cfg, torch.nn.Identity(), "cpu", 1)
        for k, v in m.state_dict().items():
            put(f"sched.{sched}.{k}", v)
        put(f"sched.{sched}.vlb_weights", m.vlb_weights)

    # ---- 2. U-Net epsilon prediction ------------------------------------------------------------
    for tag, cfg, hw, seed in (("c3", tc.C3, 32, 11), ("c1", tc.C1, 28, 12), ("cs", tc.CS, 8, 13), ("c2", tc.C2, 16, 14)):
        net = ref_model(cfg, "unet")
        x = tc.randn(seed, 2, cfg["unet_in"], hw, hw)
        for j, t in enumerate((torch.tensor([999, 0]), torch.tensor([500, 37]))):
            with torch.no_grad():
                put(f"unet.{tag}.eps{j}", net(x, t))
        print("unet", tag, "done")

    # ---- 3. single ancestral step with per-sample t + full short chains ---------------------------
    cfg = dict(tc.C1, T=50)
    m = ref_model(cfg, "ddpm")
    x = tc.randn(21, 2, 1, 28, 28)
    torch.manual_seed(22)
    put("p_sample.c1.out", m.p_sample(x, torch.tensor([30, 0])))
    torch.manual_seed(5)
    put("chain.c1.x", m.sample(2))
    torch.manual_seed(6)
    put("chain.c1.early", m.sample(2, early_stop=40))
    cfg = dict(tc.CS, T=50)
    m = ref_model(cfg, "dddpm_ae")
    torch.manual_seed(5)
    xs, zs = m.sample(2)
    put("chain.cs.x", xs)
    put("chain.cs.z", zs)
    print("chains done")

    # ---- 4. down / up-sampling nets ---------------------------------------------------------------
    m = ref_model(tc.C2, "dddpm_ae")
    x = tc.rand_pm1(31, 2, 3, 64, 64)
    with torch.no_grad():
        z = m.rescaled_downsample(x)
        put("resample.c2.z", z)
        put("resample.c2.xhat", m.rescaled_upsample(z))
    for mode in ("convolutional",):
        cfg = dict(tc.CS, d_mode=mode, u_mode=mode)
        m = ref_model(cfg, "dddpm")
        x = tc.rand_pm1(32, 2, 3, 32, 32)
        with torch.no_grad():
            z = m.rescaled_downsample(x)
            put(f"resample.{mode}.z", z)
            put(f"resample.{mode}.xhat", m.rescaled_upsample(z))

    # ---- 5. q_sample / predict_x0 / q_posterior ----------------------------------------------------
    m = ref_model(tc.C1, "ddpm")
    x, e = tc.randn(41, 4, 1, 28, 28), tc.randn(42, 4, 1, 28, 28)
    t = torch.tensor([0, 1, 500, 999])
    put("ddpm.q_sample", m.q_sample(x, t, e))
    put("ddpm.predict_x0.clip", m.predict_x_from_eps(x.clone(), t, e, clip=True))
    put("ddpm.predict_x0.noclip", m.predict_x_from_eps(x.clone(), t, e, clip=False))
    mean, var, logvar = m.q_posterior(e.clamp(-1, 1), x, t)
    put("ddpm.q_posterior.mean", mean)
    put("ddpm.q_posterior.var", var)
    put("ddpm.q_posterior.logvar", logvar)

    # ---- 6. training objectives + gradients ---------------------------------------------------------
    for kind in ("dddpm_ae", "dddpm"):
        m = ref_model(tc.CS, kind)
        m.train()
        x = tc.rand_pm1(51, 4, 3, 32, 32)
        t = torch.tensor([3, 50, 99, 700])
        torch.manual_seed(7)
        obj, d = m.losses(x, t)
        obj.backward()
        put(f"loss.{kind}.obj", obj)
        put(f"loss.{kind}.latent", d["latent"])
        put(f"loss.{kind}.recon", d["recon"])
        names, norms = [], []
        for n, p in m.named_parameters():
            names.append(n)
            norms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
        put(f"loss.{kind}.grad_norms", np.asarray(norms))
        for n in ("latent_model.final_conv.1.weight", "latent_model.downs.0.0.block1.block.1.weight",
                  "latent_model.mid_attn.fn.norm.g", "latent_model.time_mlp.1.bias", "upsample.conv.0.weight",
                  "downsample.conv.7.bias", "latent_model.ups.0.3.conv.bias", "latent_model.downs.0.3.conv.bias"):
            g = dict(m.named_parameters())[n].grad
            put(f"loss.{kind}.grad.{n}", torch.zeros(1) if g is None else g)
    for lt, lf in (("vlb", "sum"), ("hybrid", "mean"), ("simple", "mean")):
        cfg = dict(tc.C1, loss_type=lt, loss_flat=lf)
        m = ref_model(cfg, "ddpm")
        x = tc.rand_pm1(52, 4, 1, 28, 28)
        t = torch.tensor([0, 10, 400, 999])
        torch.manual_seed(8)
        obj = m.losses(x, t)
        obj.backward()
        put(f"loss.c1.{lt}.{lf}.obj", obj)
        put(f"loss.c1.{lt}.{lf}.grad_final", dict(m.named_parameters())["latent_model.final_conv.1.weight"].grad)
    print("losses done")

    # ---- 7. EMA ---------------------------------------------------------------------------------------
    net = ref_model(tc.CS, "unet")
    ema = ref.EMA(net, decay=0.995)
    for k in range(3):
        g = torch.Generator().manual_seed(60 + k)
        with torch.no_grad():
            for p in net.parameters():
                p.add_(0.01 * torch.randn(p.shape, generator=g))
        ema.update(net)
    sd = ema.state_dict()
    for n in ("final_conv.1.weight", "downs.0.0.block1.block.0.bias", "mid_attn.fn.norm.g", "time_mlp.3.weight"):
        put(f"ema.{n}", sd[n])

    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
