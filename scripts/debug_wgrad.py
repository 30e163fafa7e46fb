"""Isolated check of dd_conv_wgrad_tc32 against torch (structured inputs to expose layout mistakes)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from downsampled_diffusion_b200 import _lib as L
import torch.nn.functional as F
dev = torch.device("cuda:0")
L.lib()
def run(kind, B, C, Cout, H, W, mode):
    taps = 9 if kind == 0 else 1
    g = torch.Generator().manual_seed(0)
    if mode == "ones":
        x = torch.ones(B, H, W, C); dy = torch.ones(B, H, W, Cout)
    elif mode == "chan":      # x depends on channel only, dy on channel only
        x = torch.arange(C).float().view(1, 1, 1, C).expand(B, H, W, C).contiguous() + 1
        dy = torch.arange(Cout).float().view(1, 1, 1, Cout).expand(B, H, W, Cout).contiguous() * 0.5 + 1
    else:
        x = torch.randn(B, H, W, C, generator=g); dy = torch.randn(B, H, W, Cout, generator=g)
    Wp = max(32, W)
    xs, dys = x.to(dev), dy.to(dev)
    ns = 3 if taps == 9 else 1
    xd = torch.full((ns, B, C, H + 2, Wp), float("nan"), device=dev); dyd = torch.full((B, Cout, H, Wp), float("nan"), device=dev)
    L.call("dd_nhwc_to_chw_pad", L.ptr(xs), L.ptr(xd), B, C, H, W, Wp, 1, ns, None, L.stream())
    L.call("dd_nhwc_to_chw_pad", L.ptr(dys), L.ptr(dyd), B, Cout, H, W, Wp, 0, 1, None, L.stream())
    dw = torch.zeros(taps, C, Cout, device=dev)
    L.call("dd_conv_wgrad_tc32", kind, L.ptr(xd), None, C, 0, L.ptr(dyd), L.ptr(dw), B, H, W, Wp, Cout, L.stream())
    torch.cuda.synchronize()
    xn = x.permute(0, 3, 1, 2).double().requires_grad_(False)
    w = torch.zeros(Cout, C, 3 if taps == 9 else 1, 3 if taps == 9 else 1, dtype=torch.double, requires_grad=True)
    y = F.conv2d(xn, w, padding=1 if taps == 9 else 0)
    y.backward(dy.permute(0, 3, 1, 2).double())
    ref = w.grad.permute(2, 3, 1, 0).reshape(taps, C, Cout).float()
    got = dw.cpu()
    err = float((got - ref).norm() / ref.norm())
    print(f"kind={kind} B={B} C={C} Cout={Cout} {H}x{W} {mode}: rel err {err:.3e}  |got| {float(got.norm()):.3e} |ref| {float(ref.norm()):.3e}")
    if err > 1e-2:
        nz = (got != 0).float().mean()
        print("  nonzero fraction", float(nz), "got[0] row0[:8]", got[0][0, :8].tolist(), "col0[:8]", got[0][:8, 0].tolist())
        print("  ref[0] row0[:8]", ref[0][0, :8].tolist(), "col0[:8]", ref[0][:8, 0].tolist())
if len(sys.argv) > 1:      # kind B C Cout H W mode
    a = sys.argv[1:]
    run(int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]), int(a[5]), a[6])
else:
    for mode in ("ones", "chan", "rand"):
        run(1, 1, 32, 32, 32, 32, mode)
        run(0, 1, 32, 32, 32, 32, mode)
    run(0, 3, 64, 128, 32, 32, "rand")
    run(0, 5, 128, 256, 32, 64, "rand")
