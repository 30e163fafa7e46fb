"""Training denoising step on the B200: objectives and gradients of DDPM.losses / DownsampleDDPM(.Autoencoder).losses
against the golden vectors the real reference produced with torch autograd on CPU (oracle/make_golden.py), plus
kernel-level backward checks against torch-CPU autograd.  Training programs are fp32: tolerance 1e-4 relative."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import downsampled_diffusion_b200 as dd
from oracle import ddpm_oracle as O
from tests import common as tc

pytestmark = pytest.mark.gpu


def G(golden, key):
    return torch.from_numpy(np.asarray(golden[key]))


@pytest.mark.parametrize("kind", ["dddpm_ae", "dddpm"])
def test_dddpm_losses_and_gradients(cuda, golden, kind):
    m = tc.build_model(dict(tc.CS, precision="fp32"), dd, kind, device="cuda").to(cuda)
    m.train()
    x = tc.rand_pm1(51, 4, 3, 32, 32).to(cuda)
    t = torch.tensor([3, 50, 99, 700], device=cuda)
    torch.manual_seed(7)
    eps = torch.randn(4, 8, 8, 8).to(cuda)           # the draw the reference made with randn_like(z) on CPU
    obj, d = m.losses(x, t, eps=eps)
    obj.backward()
    for key, val in (("obj", obj), ("latent", d["latent"]), ("recon", d["recon"])):
        ref = float(golden[f"loss.{kind}.{key}"])
        assert abs(float(val) - ref) <= 1e-4 * abs(ref) + 1e-7, key
    names = [n for n, _ in m.named_parameters()]
    norms = np.asarray(golden[f"loss.{kind}.grad_norms"])
    worst = 0.0
    for (n, p), ref in zip(m.named_parameters(), norms):
        got = 0.0 if p.grad is None else float(p.grad.double().norm())
        err = abs(got - ref) / max(ref, 1e-6)
        worst = max(worst, err)
        assert err < 2e-3, f"{n}: |grad| {got} vs reference {ref}"
    print(f"{kind}: worst relative gradient-norm error over {len(names)} parameters: {worst:.2e}")
    params = dict(m.named_parameters())
    for n in ("latent_model.final_conv.1.weight", "latent_model.downs.0.0.block1.block.1.weight",
              "latent_model.mid_attn.fn.norm.g", "latent_model.time_mlp.1.bias", "upsample.conv.0.weight",
              "downsample.conv.7.bias", "latent_model.ups.0.3.conv.bias", "latent_model.downs.0.3.conv.bias"):
        ref = G(golden, f"loss.{kind}.grad.{n}")
        g = params[n].grad
        if ref.numel() == 1 and g is None:
            continue
        assert tc.rel_l2(g, ref) < 2e-4, n


@pytest.mark.parametrize("lt,lf", [("vlb", "sum"), ("hybrid", "mean"), ("simple", "mean")])
def test_ddpm_objective_variants(cuda, golden, lt, lf):
    cfg = dict(tc.C1, loss_type=lt, loss_flat=lf, precision="fp32")
    m = tc.build_model(cfg, dd, "ddpm", device="cuda").to(cuda)
    m.train()
    x = tc.rand_pm1(52, 4, 1, 28, 28).to(cuda)
    t = torch.tensor([0, 10, 400, 999], device=cuda)
    torch.manual_seed(8)
    eps = torch.randn(x.shape).to(cuda)
    obj = m.p_losses(x, t, eps=eps)
    obj.backward()
    ref = float(golden[f"loss.c1.{lt}.{lf}.obj"])
    assert abs(float(obj) - ref) <= 1e-4 * abs(ref)
    g = dict(m.named_parameters())["latent_model.final_conv.1.weight"].grad
    assert tc.rel_l2(g, G(golden, f"loss.c1.{lt}.{lf}.grad_final")) < 2e-4


def test_forward_draws_t_and_eps_like_the_reference(cuda):
    """DDPM.forward = t_sample (randint) then losses (randn_like): ddpm.py:448-457."""
    cfg = dict(tc.CS, precision="fp32")
    m = tc.build_model(cfg, dd, "dddpm_ae", device="cuda").to(cuda)
    x = tc.rand_pm1(3, 2, 3, 32, 32).to(cuda)
    torch.manual_seed(11)
    obj, d = m(x)
    torch.manual_seed(11)
    t = torch.randint(0, 1000, (2,), device=cuda).long()
    with torch.no_grad():
        z = m.rescaled_downsample(x)
    eps = torch.randn_like(z)
    obj2, _ = m.losses(x, t, eps=eps)
    assert abs(float(obj) - float(obj2)) <= 1e-5 * abs(float(obj2))
    assert set(d.keys()) == {"latent", "recon"}


def test_dropout_training_mode_runs_and_masks(cuda):
    cfg = dict(tc.CS, precision="fp32", unet_dropout=0.5)
    m = tc.build_model(cfg, dd, "ddpm", device="cuda").to(cuda)
    m.train()
    x = tc.randn(1, 2, 8, 8, 8).to(cuda)
    t = torch.tensor([5, 500], device=cuda)
    a = m.latent_model(x, t)
    b = m.latent_model(x, t)
    assert torch.isfinite(a).all() and not torch.equal(a, b)          # a fresh mask per call
    with pytest.raises(RuntimeError, match="overwritten"):             # b's forward reused the program's activation buffers
        a.sum().backward()
    b.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.latent_model.parameters())
    m.eval()
    with torch.no_grad():
        assert torch.equal(m.latent_model(x, t), m.latent_model(x, t))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graph_replay_matches_eager_launch_lists(cuda, monkeypatch, precision):
    """The captured forward / backward graphs of the three networks give the gradients of the eager launch lists, on the
    capturing call and on later replays (with new inputs and repacked weights)."""
    def run(m, seed):
        x = tc.rand_pm1(seed, 4, 3, 32, 32).to(cuda)
        t = torch.tensor([3, 50, 99, 700], device=cuda)
        torch.manual_seed(seed)
        eps = torch.randn(4, 8, 8, 8).to(cuda)
        m.zero_grad()
        obj, _ = m.losses(x, t, eps=eps)
        obj.backward()
        return float(obj), [p.grad.detach().clone() for p in m.parameters()]

    def fresh():
        m = tc.build_model(dict(tc.CS, precision=precision), dd, "dddpm_ae", device="cuda").to(cuda)
        m.train()
        return m

    tol = 1e-5 if precision == "fp32" else 1e-4         # weight-gradient atomics reorder; TF32 tiles are deterministic otherwise
    monkeypatch.setenv("DD_TRAIN_GRAPH", "0")
    m = fresh()
    ref = [run(m, s) for s in (61, 62, 63)]
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.01)                                # bumps the parameter version: weights are repacked outside the graphs
    ref.append(run(m, 64))
    monkeypatch.setenv("DD_TRAIN_GRAPH", "1")
    m = fresh()
    got = [run(m, s) for s in (61, 62, 63)]
    progs = [p for net in (m.latent_model, m.downsample, m.upsample) for p in net._train_programs.values()]
    assert len(progs) == 3 and all(p.graphs is not None for p in progs)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.01)
    got.append(run(m, 64))
    for (lo, go), (lr, gr) in zip(got, ref):
        assert abs(lo - lr) <= tol * abs(lr)
        for a, b in zip(go, gr):
            assert tc.rel_l2(a, b) < tol


def test_training_step_with_ema_and_grad_accumulation(cuda):
    """Two micro-batches (gradient_accumulate_every=2, trainer_ddpm.py:35,219-229), Adam step, EMA update."""
    cfg = dict(tc.CS, precision="fp32")
    m = tc.build_model(cfg, dd, "dddpm_ae", device="cuda").to(cuda)
    m.train()
    ema = dd.EMA(m, decay=0.995)
    opt = torch.optim.Adam(m.parameters(), lr=2e-4)
    before = [p.detach().clone() for p in m.parameters()]
    losses = []
    for step in range(2):
        for k in range(2):
            x = tc.rand_pm1(70 + 2 * step + k, 4, 3, 32, 32).to(cuda)
            obj, _ = m(x)
            (obj / 2).backward()
            losses.append(float(obj))
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
        ema.update(m)
    assert all(np.isfinite(losses))
    assert any(not torch.equal(a, b) for a, b in zip(before, m.parameters()))
    with torch.no_grad():                                  # the updated weights are what the next forward uses
        x = tc.rand_pm1(99, 4, 3, 32, 32).to(cuda)
        t = torch.tensor([1, 2, 3, 4], device=cuda)
        eps = tc.randn(98, 4, 8, 8, 8).to(cuda)
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        ref, _ = O.dddpm_losses(sd, cfg, O.schedule_buffers("linear", 1000), x.cpu(), t.cpu(), eps.cpu(), autoencoder=True)
        m.eval()
        got, _ = m.losses(x, t, eps=eps)
    assert abs(float(got) - float(ref)) <= 2e-4 * abs(float(ref))


# ---- kernel-level backward checks -------------------------------------------------------------------
def test_backward_kernels_against_torch_autograd(cuda):
    from downsampled_diffusion_b200 import _lib as L
    torch.manual_seed(0)
    B, C, H, W, G_ = 2, 32, 6, 6, 8
    # GroupNorm + Mish
    x = torch.randn(B, C, H, W, requires_grad=True)
    gamma, beta = torch.randn(C, requires_grad=True), torch.randn(C, requires_grad=True)
    dy = torch.randn(B, C, H, W)
    F.mish(F.group_norm(x, G_, gamma, beta, 1e-5)).backward(dy)
    nh = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().to(cuda)
    xd, dyd = nh(x), nh(dy)
    st = torch.empty(B, G_, 2, device=cuda)
    L.call("dd_gn_stats", L.ptr(xd), L.DD_F32, B, H * W, C, G_, 1e-5, L.ptr(st), L.stream())
    s1, s2, s3 = (torch.empty(B, C, device=cuda) for _ in range(3))
    dx = torch.empty_like(xd)
    gd, bd = gamma.detach().to(cuda), beta.detach().to(cuda)
    L.call("dd_gn_mish_bwd", L.ptr(xd), L.ptr(dyd), L.ptr(st), L.ptr(gd), L.ptr(bd), B, H * W, C, G_, L.ptr(s1), L.ptr(s2),
           L.ptr(s3), L.ptr(dx), 0, L.stream())
    assert tc.rel_l2(dx.permute(0, 3, 1, 2), x.grad) < 1e-5
    assert tc.rel_l2(s1.sum(0), gamma.grad) < 1e-5 and tc.rel_l2(s2.sum(0), beta.grad) < 1e-5
    assert tc.rel_l2(s3.sum(0), dy.sum((0, 2, 3))) < 1e-5
    # channel LayerNorm (eps on the std)
    x = torch.randn(B, 64, H, W, requires_grad=True)
    g, b = torch.randn(1, 64, 1, 1, requires_grad=True), torch.randn(1, 64, 1, 1, requires_grad=True)
    dy = torch.randn(B, 64, H, W)
    O.channel_layernorm(x, g, b).backward(dy)
    xd, dyd = nh(x), nh(dy)
    dx, dg, db = torch.empty_like(xd), torch.zeros(64, device=cuda), torch.zeros(64, device=cuda)
    gdev = g.detach().reshape(-1).to(cuda)
    L.call("dd_layernorm_c_bwd", L.ptr(xd), L.ptr(dyd), L.ptr(gdev), 1e-5, B * H * W, 64, L.ptr(dx), 0, L.ptr(dg), L.ptr(db), L.stream())
    assert tc.rel_l2(dx.permute(0, 3, 1, 2), x.grad) < 1e-5
    assert tc.rel_l2(dg, g.grad.reshape(-1)) < 1e-5 and tc.rel_l2(db, b.grad.reshape(-1)) < 1e-5
    # LinearAttention core
    heads, dh, n = 4, 32, H * W
    qkv = torch.randn(B, 3 * heads * dh, H, W, requires_grad=True)
    dout = torch.randn(B, heads * dh, H, W)
    q, k, v = qkv.reshape(B, 3, heads, dh, n).unbind(1)
    ctx = torch.einsum("bhdn,bhen->bhde", k.softmax(dim=-1), v)
    torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(B, heads * dh, H, W).backward(dout)
    qd, dod = nh(qkv), nh(dout)
    need = int(L.lib().dd_linattn_ws_floats(B, n, heads))
    ws = torch.empty(need, device=cuda)
    out = torch.empty(B, H, W, heads * dh, device=cuda)
    L.call("dd_linattn_core", L.ptr(qd), L.ptr(out), L.DD_F32, B, n, heads, dh, L.ptr(ws), need, L.stream())
    saved = torch.empty(B * heads * 1088, device=cuda)
    L.call("dd_linattn_save", L.ptr(ws), B, n, heads, L.ptr(saved), L.stream())
    dctx, dqkv = torch.empty(B * heads * 1024, device=cuda), torch.empty_like(qd)
    L.call("dd_linattn_bwd", L.ptr(qd), L.ptr(dod), L.ptr(saved), L.ptr(dctx), L.ptr(dqkv), B, n, heads, dh, L.stream())
    assert tc.rel_l2(dqkv.permute(0, 3, 1, 2), qkv.grad) < 1e-5


@pytest.mark.parametrize("kind", ["3x3", "1x1", "down", "up"])
def test_conv_weight_and_input_gradients(cuda, kind):
    from downsampled_diffusion_b200.autograd import TrainProgram
    from downsampled_diffusion_b200.engine import Act
    torch.manual_seed(1)
    B, C1, C2, Cout, H, W = 2, 16, 8 if kind == "3x3" else 0, 24, 8, 8
    Cin = C1 + C2
    conv = {"3x3": lambda: torch.nn.Conv2d(Cin, Cout, 3, 1, 1), "1x1": lambda: torch.nn.Conv2d(Cin, Cout, 1),
            "down": lambda: torch.nn.Conv2d(Cin, Cout, 3, 2, 1), "up": lambda: torch.nn.ConvTranspose2d(Cin, Cout, 4, 2, 1)}[kind]()
    x = torch.randn(B, Cin, H, W, requires_grad=True)
    y = conv(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    holder = torch.nn.ModuleList([conv]).to(cuda)
    prog = TrainProgram(holder, B)
    nh = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().to(cuda)
    xa = Act(nh(x[:, :C1]), B, H, W, C1)
    x2a = Act(nh(x[:, C1:]), B, H, W, C2) if C2 else None
    ya = prog.t_conv(xa, holder[0], x2=x2a, kind=kind)

    prog.gwritten.add(id(ya.t))             # the test plays the role of the consumer that wrote dL/dy
    prog.build_backward()
    prog.refresh_weights()
    prog.run_ops()
    assert tc.rel_l2(ya.t.permute(0, 3, 1, 2), y) < 1e-5
    prog.grad(ya).copy_(nh(dy))
    prog.run_backward()
    pg = prog.param_grads()
    ref = {n: p.grad for n, p in conv.named_parameters()}
    assert tc.rel_l2(pg[id(holder[0].weight)], ref["weight"]) < 1e-5
    assert tc.rel_l2(pg[id(holder[0].bias)], ref["bias"]) < 1e-5
    assert tc.rel_l2(prog.grad(xa).permute(0, 3, 1, 2), x.grad[:, :C1]) < 1e-5
    if C2:
        assert tc.rel_l2(prog.grad(x2a).permute(0, 3, 1, 2), x.grad[:, C1:]) < 1e-5


# ---- TF32 tensor-core training form (dd_conv_tc32) ----------------------------------------------------

@pytest.mark.parametrize("kind,C1,C2,Cout,H,W,B,pre_mish", [
    ("3x3", 32, 0, 32, 16, 16, 3, True),        # ConvResBlock c2/c3 shape (convblocks.py:103-104), pre-activation
    ("3x3", 64, 64, 128, 8, 8, 5, False),       # two sources (skip concat), 2 images per tile, ragged batch
    ("1x1", 64, 0, 32, 32, 32, 2, True),        # ConvResBlock c1
    ("3x3", 128, 0, 256, 4, 4, 9, False),       # 8 images per tile, 2 n-tiles
    ("down", 64, 0, 96, 16, 16, 2, False),      # Downsample (blocks.py:44) as a 3x3 conv on the space-to-depth input
    ("up", 64, 0, 32, 8, 8, 3, False),          # Upsample (blocks.py:35) as a 3x3 conv to 4*Cout sub-pixel channels
])
def test_tf32_conv_forward_and_gradients(cuda, kind, C1, C2, Cout, H, W, B, pre_mish):
    """dd_conv_tc32 forward, input gradient (same kernel, flipped weights) and the fp32 weight gradient of a conv
    lowered by TrainProgram(tf32=True), against torch autograd on CPU.  TF32 operands: tolerance 2e-3 relative."""
    from downsampled_diffusion_b200.autograd import TrainProgram
    from downsampled_diffusion_b200.engine import Act
    torch.manual_seed(2)
    Cin = C1 + C2
    conv = {"3x3": lambda: torch.nn.Conv2d(Cin, Cout, 3, 1, 1), "1x1": lambda: torch.nn.Conv2d(Cin, Cout, 1),
            "down": lambda: torch.nn.Conv2d(Cin, Cout, 3, 2, 1), "up": lambda: torch.nn.ConvTranspose2d(Cin, Cout, 4, 2, 1)}[kind]()
    x = torch.randn(B, Cin, H, W, requires_grad=True)
    y = conv(F.mish(x) if pre_mish else x)
    dy = torch.randn_like(y)
    y.backward(dy)
    holder = torch.nn.ModuleList([conv]).to(cuda)
    prog = TrainProgram(holder, B, tf32=True)
    nh = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().to(cuda)
    xa = Act(nh(x[:, :C1]), B, H, W, C1)
    x2a = Act(nh(x[:, C1:]), B, H, W, C2) if C2 else None
    ya = prog.t_conv(xa, holder[0], x2=x2a, kind=kind, pre_mish=pre_mish)
    assert "dd_conv_tc32" in prog.op_names
    prog.gwritten.add(id(ya.t))
    prog.build_backward()
    prog.refresh_weights()
    prog.run_ops()
    assert tc.rel_l2(ya.t.permute(0, 3, 1, 2), y) < 2e-3
    prog.grad(ya).copy_(nh(dy))
    prog.run_backward()
    pg = prog.param_grads()
    assert tc.rel_l2(pg[id(holder[0].weight)], conv.weight.grad) < 2e-3
    assert tc.rel_l2(pg[id(holder[0].bias)], conv.bias.grad) < 1e-5
    assert tc.rel_l2(prog.grad(xa).permute(0, 3, 1, 2), x.grad[:, :C1]) < 2e-3
    if C2:
        assert tc.rel_l2(prog.grad(x2a).permute(0, 3, 1, 2), x.grad[:, C1:]) < 2e-3


def test_tf32_training_objective_and_gradients(cuda, golden):
    """The default (precision='bf16' sampling) model trains with TF32 tensor-core convolutions: objective within 1e-3,
    gradient norms within 2e-2 of the reference's fp32 CPU autograd (golden vectors)."""
    kind = "dddpm_ae"
    m = tc.build_model(dict(tc.CS, precision="bf16"), dd, kind, device="cuda").to(cuda)
    m.train()
    x = tc.rand_pm1(51, 4, 3, 32, 32).to(cuda)
    t = torch.tensor([3, 50, 99, 700], device=cuda)
    torch.manual_seed(7)
    eps = torch.randn(4, 8, 8, 8).to(cuda)
    obj, d = m.losses(x, t, eps=eps)
    obj.backward()
    ref = float(golden[f"loss.{kind}.obj"])
    assert abs(float(obj) - ref) <= 1e-3 * abs(ref)
    norms = np.asarray(golden[f"loss.{kind}.grad_norms"])
    for (n, p), r in zip(m.named_parameters(), norms):
        got = 0.0 if p.grad is None else float(p.grad.double().norm())
        assert abs(got - r) <= 2e-2 * max(r, 1e-6) + 1e-5, f"{n}: |grad| {got} vs reference {r}"
    g = m.latent_model.final_conv[1].weight.grad
    assert tc.rel_l2(g, G(golden, f"loss.{kind}.grad.latent_model.final_conv.1.weight")) < 1e-2


# ---- kernel-level checks of the training-path entry points added with the tensor-core path ---------------------

def test_conv_tc32_fused_epilogue(cuda):
    """dd_conv_tc32: y = (conv + bias) * mish'(z) + addend and y_mish = mish(y), on both TF32 kernels (halo form: 3x3 32->32
    on a 16x16 map; generic persistent form: 1x1 64->32 and a two-n-tile 3x3)."""
    from downsampled_diffusion_b200 import _lib as L
    for (ks, Cin, Cout, H, W, B) in ((3, 32, 32, 16, 16, 3), (1, 64, 32, 32, 32, 2), (3, 64, 256, 8, 8, 5)):
        x, w, b = tc.randn(1, B, Cin, H, W), tc.randn(2, Cout, Cin, ks, ks) * 0.1, tc.randn(3, Cout)
        z, add = tc.randn(4, B, Cout, H, W), tc.randn(5, B, Cout, H, W)
        ref = F.conv2d(x, w, b, padding=ks // 2)
        zz = z.clone().requires_grad_(True)
        F.mish(zz).sum().backward()                                  # mish'(z)
        ref = ref * zz.grad + add
        nh = lambda t: t.permute(0, 2, 3, 1).contiguous().to(cuda)
        xd, zd, ad = nh(x), nh(z), nh(add)
        wp = w.permute(0, 2, 3, 1).reshape(Cout, ks * ks * Cin).contiguous().to(cuda)
        bd = b.to(cuda)
        y = torch.empty(B, H, W, Cout, device=cuda)
        ym = torch.empty_like(y)
        L.call("dd_conv_tc32", L.TC_CONV3x3 if ks == 3 else L.TC_CONV1x1, L.ptr(xd), None, Cin, 0, L.ptr(wp), Cout, L.ptr(bd), L.ptr(ad),
               L.ptr(y), L.ptr(ym), L.ptr(zd), B, H, W, Cout, L.stream())
        assert tc.rel_l2(y.permute(0, 3, 1, 2), ref) < 2e-3, (ks, Cin, Cout)
        assert tc.rel_l2(ym.permute(0, 3, 1, 2), F.mish(y.permute(0, 3, 1, 2).cpu())) < 1e-5


def test_thin_conv_kernels(cuda):
    """3 -> 64 / 64 -> 3 layers of the resampling nets: forward, tanh, weight gradient (both layouts), bias gradient."""
    from downsampled_diffusion_b200 import _lib as L
    B, H, W, Cs, Cw = 3, 16, 32, 3, 64
    xn, w, b = tc.randn(1, B, Cs, H, W), tc.randn(2, Cw, Cs) * 0.3, tc.randn(3, Cw)
    y = torch.empty(B, H, W, Cw, device=cuda)
    xnd, wtd, bd = xn.to(cuda), w.t().contiguous().to(cuda), b.to(cuda)          # named: raw pointers do not keep temporaries alive
    L.call("dd_conv1x1_thin_in", L.ptr(xnd), L.ptr(wtd), L.ptr(bd), L.ptr(y), B, H * W, Cs, Cw, 0, L.stream())
    ref = F.conv2d(xn, w.view(Cw, Cs, 1, 1), b)
    assert tc.rel_l2(y.permute(0, 3, 1, 2), ref) < 1e-6
    L.call("dd_conv1x1_thin_in", L.ptr(xnd), L.ptr(wtd), None, L.ptr(y), B, H * W, Cs, Cw, 1, L.stream())
    assert tc.rel_l2(y.permute(0, 3, 1, 2), 2 * ref - b.view(1, -1, 1, 1)) < 1e-6            # accumulate, no bias
    xw, w2, b2 = tc.randn(4, B, Cw, H, W), tc.randn(5, Cs, Cw) * 0.2, tc.randn(6, Cs)
    xwd = xw.permute(0, 2, 3, 1).contiguous().to(cuda)
    w2d, b2d = w2.to(cuda), b2.to(cuda)
    for do_tanh in (0, 1):
        yo = torch.empty(B, Cs, H, W, device=cuda)
        L.call("dd_conv1x1_thin_out", L.ptr(xwd), L.ptr(w2d), L.ptr(b2d), L.ptr(yo), B, H * W, Cw, Cs, do_tanh, L.stream())
        r = F.conv2d(xw, w2.view(Cs, Cw, 1, 1), b2)
        assert tc.rel_l2(yo, torch.tanh(r) if do_tanh else r) < 1e-5
    g = tc.randn(7, B, Cs, H, W)                                            # narrow operand (NCHW), wide operand xw (NHWC)
    want = torch.einsum("bshw,bchw->sc", g.double(), xw.double()).float()
    gd = g.to(cuda)
    for narrow_major in (1, 0):
        dw = torch.zeros(Cs, Cw, device=cuda) if narrow_major else torch.zeros(Cw, Cs, device=cuda)
        db = torch.zeros(Cw, device=cuda)
        L.call("dd_conv1x1_thin_wgrad", L.ptr(gd), L.ptr(xwd), L.ptr(dw), narrow_major, L.ptr(db), B, H * W, Cs, Cw, L.stream())
        assert tc.rel_l2(dw if narrow_major else dw.t(), want) < 1e-5
        assert tc.rel_l2(db, xw.sum((0, 2, 3))) < 1e-5


def test_operand_copies_and_space_to_depth(cuda):
    from downsampled_diffusion_b200 import _lib as L
    B, H, W, C = 2, 8, 16, 32
    x = tc.randn(1, B, H, W, C)
    xd = x.to(cuda)
    Wp = 32
    y = torch.full((3, B, C, H + 2, Wp), float("nan"), device=cuda)
    cs = torch.zeros(C, device=cuda)
    L.call("dd_nhwc_to_chw_pad", L.ptr(xd), L.ptr(y), B, C, H, W, Wp, 1, 3, L.ptr(cs), L.stream())
    ref = torch.zeros(3, B, C, H + 2, Wp)
    xc = x.permute(0, 3, 1, 2)
    for s in range(3):
        for wcol in range(Wp):                  # a shifted copy may carry x[W-1] into the first padding column: it meets a zero of dY
            src = wcol + s - 1
            if 0 <= src < W:
                ref[s, :, :, 1:H + 1, wcol] = xc[:, :, :, src]
    assert torch.equal(y.cpu(), ref)
    assert tc.rel_l2(cs, x.sum((0, 1, 2))) < 1e-5
    full = tc.randn(2, B, 2 * H, 2 * W, C).to(cuda)
    packed = torch.empty(B, H, W, 4 * C, device=cuda)
    L.call("dd_s2d_f32", L.ptr(full), L.ptr(packed), B, H, W, C, 1, L.stream())
    want = torch.stack([full[:, py::2, px::2, :] for py in (0, 1) for px in (0, 1)], dim=3).reshape(B, H, W, 4 * C)
    assert torch.equal(packed, want)
    back = torch.empty_like(full)
    L.call("dd_s2d_f32", L.ptr(packed), L.ptr(back), B, H, W, C, 0, L.stream())
    assert torch.equal(back, full)


# ---- optimizer step of the trainer: clip_grad_norm_ + Adam + EMA in three launches -------------------
OPT_NAMES = ("final_conv.1.weight", "downs.0.0.block1.block.0.bias", "mid_attn.fn.norm.g", "time_mlp.3.weight")


def _set_grads(net, step):
    g = torch.Generator().manual_seed(80 + step)
    for p in net.parameters():          # the gradients oracle/make_golden_eval.py fed the reference's optimizer
        p.grad = ((0.05 if step == 0 else 0.0005) * torch.randn(p.shape, generator=g)).to(p.device)


def test_fused_adam_matches_torch_adam_of_the_reference_trainer(cuda, golden):
    net = tc.build_model(tc.CS, dd, "unet").to(cuda)
    twin = tc.build_model(tc.CS, dd, "unet").to(cuda)               # torch's own CUDA implementation, same gradients
    opt = dd.Adam(net.parameters(), lr=2e-4, max_grad_norm=1.0)
    topt = torch.optim.Adam(twin.parameters(), lr=2e-4)
    for step in range(2):
        _set_grads(net, step)
        _set_grads(twin, step)
        opt.step()
        tnorm = torch.nn.utils.clip_grad_norm_(twin.parameters(), 1.0)
        topt.step()
        ref_norm = float(golden[f"optim.norm{step}"])
        assert abs(float(opt.grad_norm) - ref_norm) <= 2e-6 * ref_norm and abs(float(tnorm) - ref_norm) <= 1e-5 * ref_norm
        sd, tsd = net.state_dict(), twin.state_dict()
        for n in OPT_NAMES:                                          # against the reference's CPU run: within 2 ulp of the weights
            ref = G(golden, f"optim.step{step}.{n}").to(cuda)
            assert (sd[n] - ref).abs().max() <= 2.4e-7 * max(1.0, float(ref.abs().max())), n
        for n in sd:
            assert (sd[n] - tsd[n]).abs().max() <= 2.4e-7 * max(1.0, float(tsd[n].abs().max())), n
    # optimizer state has torch's layout: a torch.optim.Adam continues from it and vice versa
    state = opt.state_dict()
    assert set(state["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(state["state"][0]["step"]) == 2.0
    topt2 = torch.optim.Adam(net.parameters(), lr=2e-4)
    topt2.load_state_dict(state)
    opt2 = dd.Adam(twin.parameters(), lr=2e-4, max_grad_norm=1.0)
    opt2.load_state_dict(topt.state_dict())
    _set_grads(net, 2)
    _set_grads(twin, 2)
    torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
    topt2.step()
    opt2.step()
    for a, b in zip(net.parameters(), twin.parameters()):
        assert (a - b).abs().max() <= 2.4e-7 * max(1.0, float(a.abs().max()))


def test_fused_adam_ema_and_weight_refresh(cuda):
    m = tc.build_model(dict(tc.CS, precision="fp32"), dd, "dddpm_ae", device="cuda").to(cuda).train()
    twin = tc.build_model(dict(tc.CS, precision="fp32"), dd, "dddpm_ae", device="cuda").to(cuda).train()
    ema, tema = dd.EMA(m, decay=0.9), dd.EMA(twin, decay=0.9)
    opt = dd.Adam(m.parameters(), lr=1e-3, max_grad_norm=1.0)
    opt.attach_ema(ema, m)
    topt = dd.Adam(twin.parameters(), lr=1e-3, max_grad_norm=1.0)
    x = tc.rand_pm1(90, 4, 3, 32, 32).to(cuda)
    t = torch.tensor([3, 50, 99, 700], device=cuda)
    eps = tc.randn(91, 4, 8, 8, 8).to(cuda)
    losses = []
    for step, mode in enumerate(("reset", "update", "update")):
        for mod, o in ((m, opt), (twin, topt)):
            obj, _ = mod.losses(x, t, eps=eps)
            obj.backward()
        losses.append(float(obj))
        opt.step(ema=mode, zero_grad=(step == 1))                   # fused: clip + Adam + EMA (+ gradient reset)
        topt.step()                                                 # unfused: the same optimizer, then the EMA calls of the trainer
        (tema.reset if mode == "reset" else tema.update)(twin)
        if step == 1:
            assert all(p.grad is not None and not p.grad.any() for p in m.parameters())
        else:
            opt.zero_grad()
        topt.zero_grad()
        # (weight-gradient atomics make the two models' gradients differ in the last bits, hence not torch.equal)
        for a, b in zip(m.parameters(), twin.parameters()):
            assert torch.allclose(a, b, rtol=0, atol=2e-6)
        for a, b in zip(ema.ema_model.parameters(), tema.ema_model.parameters()):
            assert torch.allclose(a, b, rtol=0, atol=2e-6)
    assert losses[0] != losses[1] != losses[2]                      # the training programs saw the updated weights (version bump)
    with torch.no_grad():                                           # and so does the shadow model's sampling engine
        a = ema.ema_model.latent_model(eps, t)
        b = tema.ema_model.latent_model(eps, t)
    assert tc.rel_l2(a, b) < 1e-4
    with pytest.raises(RuntimeError):
        dd.Adam(twin.parameters(), lr=1e-3).step(ema="update")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_tabulated_relayouts_match_the_packing_expressions(cuda, precision):
    """relayout.py: weights -> packed buffers and packed gradients -> parameter gradients through one gather launch each,
    bit for bit what the torch expressions they were tabulated from produce, before and after an optimizer step."""
    m = tc.build_model(dict(tc.CS, precision=precision), dd, "dddpm_ae", device="cuda").to(cuda).train()
    opt = torch.optim.SGD(m.parameters(), lr=0.01)
    x = tc.rand_pm1(61, 4, 3, 32, 32).to(cuda)
    t = torch.tensor([3, 50, 99, 700], device=cuda)
    eps = tc.randn(61, 4, 8, 8, 8).to(cuda)
    for step in range(2):
        opt.zero_grad()
        obj, _ = m.losses(x, t, eps=eps)
        obj.backward()
        for net in (m.latent_model, m.downsample, m.upsample):
            (prog,) = net._train_programs.values()
            assert prog.repack_plan is not None and prog.grad_plan is not None
            # gradients: the gather launch against the per-parameter expressions, on the same arena
            fast = prog.param_grads()
            plan, prog.grad_plan = prog.grad_plan, None
            slow = prog.param_grads()
            prog.grad_plan = plan
            assert fast.keys() == slow.keys() and all(torch.equal(fast[k], slow[k]) for k in fast)
        opt.step()
        for net in (m.latent_model, m.downsample, m.upsample):
            (prog,) = net._train_programs.values()
            # packed weights of the updated parameters: the gather launch against every packing expression
            prog.refresh_weights()
            got = [b.clone() for b in prog.packed_bufs]
            for pk in prog.packers:
                pk()
            assert all(torch.equal(a, b) for a, b in zip(got, prog.packed_bufs))
            n_fast = len(prog.packers) - len(prog.slow_packers)
            if step == 0:
                print(type(net).__name__, "packers tabulated:", n_fast, "of", len(prog.packers), "; gradients tabulated:",
                      len(prog.grad_fast), "of", len(prog.pg_specs))
            assert n_fast >= 0.9 * len(prog.packers) and len(prog.grad_fast) >= 0.8 * len(prog.pg_specs)


# ---- 'deterministic' resampler mode (convblocks.py:8-26, wrapper.py:22-24, 49-53): bicubic kernels -----------------------
def test_deterministic_bicubic_mode_forward_and_gradient(cuda, golden):
    cfg = dict(tc.CS, d_mode="deterministic", u_mode="deterministic", unet_in=3, precision="fp32")
    m = tc.build_model(cfg, dd, "dddpm", device="cuda").to(cuda).train()
    x = tc.rand_pm1(85, 4, 3, 32, 32).to(cuda)
    with torch.no_grad():
        z = m.rescaled_downsample(x)
        xhat = m.rescaled_upsample(z)
    assert (z.cpu() - G(golden, "det.z")).abs().max() < 5e-6 and (xhat.cpu() - G(golden, "det.xhat")).abs().max() < 5e-6
    t = torch.tensor([3, 50, 99, 700], device=cuda)
    torch.manual_seed(12)
    eps = torch.randn(4, 3, 8, 8).to(cuda)
    obj, d = m.losses(x, t, eps=eps)
    obj.backward()
    for key, val in (("obj", obj), ("latent", d["latent"]), ("recon", d["recon"])):
        ref = float(golden[f"det.loss.{key}"])
        assert abs(float(val.detach()) - ref) <= 1e-4 * abs(ref), key
    params = dict(m.named_parameters())
    assert tc.rel_l2(params["latent_model.final_conv.1.weight"].grad, G(golden, "det.loss.grad_final")) < 2e-4
    assert tc.rel_l2(params["latent_model.downs.0.0.block1.block.0.weight"].grad, G(golden, "det.loss.grad_init")) < 2e-4
    # the kernels alone: forward against the restatement, backward against its transpose, odd sizes, up and down
    from downsampled_diffusion_b200.downsampled import Interpolate, get_interpolate
    for (h, w), size in (((32, 32), (8, 8)), ((8, 8), (32, 32)), ((7, 13), (20, 9)), ((5, 5), (1, 1))):
        a = tc.randn(86, 2, 3, h, w)
        a_dev = a.to(cuda).requires_grad_(True)
        y = Interpolate(size)(a_dev, tanh=True)
        gy = tc.randn(87, *y.shape)
        y.backward(gy.to(cuda))
        a_ref = a.clone().requires_grad_(True)
        y_ref = torch.tanh(O.bicubic_resize(a_ref, size))
        y_ref.backward(gy)
        assert (y.detach().cpu() - y_ref.detach()).abs().max() < 1e-5, (h, w, size)
        assert (a_dev.grad.cpu() - a_ref.grad).abs().max() < 1e-5 * max(1.0, float(a_ref.grad.abs().max())), (h, w, size)
    with pytest.raises(NotImplementedError):
        get_interpolate((8, 8), mode="nearest")
    with pytest.raises(RuntimeError):
        Interpolate((8, 8))(torch.zeros(1, 3, 32, 32))
