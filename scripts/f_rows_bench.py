"""Timing of the SURVEY.md 8(f) rows on one B200 (device-resident, CUDA events, after warm-up):
  f.1 fused optimizer step (clip + Adam + EMA) over the 22.67 M parameters of the C4 model, against the HBM roofline
      (algorithmic bytes: norm pass 4 B/param; update pass reads p, g, m, v, shadow and writes p, m, v, shadow = 36 B/param);
      next to torch's clip_grad_norm_ + Adam(foreach) + the package's EMA.update on the same tensors
  f.2 fix_samples on a C3 batch (64 x 3 x 256 x 256): read twice (min/max pass, normalise pass) + write = 12 B/element
  f.3 evaluation chain test_losses_ at the C3 latent shape (64 x 8 x 32 x 32, T = 1000): ms per step next to the sampling step
Usage: python scripts/f_rows_bench.py [--cpu]   (--cpu also times the oracle on a bounded sample)"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import downsampled_diffusion_b200 as dd
from tests import common as tc

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6453.4))


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {"hbm_peak_gbps": hbm}
# ---- f.1 ------------------------------------------------------------------------------------------
m = tc.build_model(dict(tc.C3, precision="bf16"), dd, "dddpm_ae", device="cuda:0").to(dev).train()
n_params = sum(p.numel() for p in m.parameters())
for p in m.parameters():
    p.grad = torch.randn_like(p) * 1e-3
ema = dd.EMA(m, decay=0.995)
opt = dd.Adam(m.parameters(), lr=2e-4, max_grad_norm=1.0)
opt.attach_ema(ema, m)
ms = timed(lambda: opt.step(ema="update"))
bytes_ = n_params * (4 + 36)
out["f1_fused_step_ms"] = ms
out["f1_fused_gbps"] = bytes_ / ms / 1e6
out["f1_frac_of_hbm_peak"] = out["f1_fused_gbps"] / hbm
topt = torch.optim.Adam(m.parameters(), lr=2e-4)


def unfused():
    torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    topt.step()
    ema.update(m)
out["f1_torch_clip_adam_plus_ema_ms"] = timed(unfused)
t0 = time.perf_counter()
for _ in range(20):
    opt.step(ema="update")
out["f1_fused_host_ms"] = (time.perf_counter() - t0) / 20 * 1e3
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    unfused()
out["f1_torch_host_ms"] = (time.perf_counter() - t0) / 20 * 1e3
torch.cuda.synchronize()
out["f1_params"] = n_params
# ---- f.2 ------------------------------------------------------------------------------------------
from downsampled_diffusion_b200 import ops
x = torch.randn(64, 3, 256, 256, device=dev)
ms = timed(lambda: ops.fix_samples_raw(x))
out["f2_fix_samples_ms"] = ms
out["f2_gbps"] = x.numel() * 12 / ms / 1e6
out["f2_torch_ms"] = timed(lambda: ((x - x.view(64, -1).min(1).values[:, None, None, None]) /
                                    (x.view(64, -1).max(1).values - x.view(64, -1).min(1).values)[:, None, None, None] * 255.).permute(0, 2, 3, 1).contiguous())
# ---- f.3 ------------------------------------------------------------------------------------------
m.eval()
z = tc.eval_images(6, 64, 8, 32, 32).to(dev)
ring = torch.randn(8, 64, 8, 32, 32, device=dev)


class Cyc:
    def __getitem__(self, k):
        return ring[k % 8]
m.test_losses_(z, noise=Cyc())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
res = m.test_losses_(z, noise=Cyc())
e1.record()
torch.cuda.synchronize()
out["f3_eval_chain_ms_per_step"] = e0.elapsed_time(e1) / 1000
out["f3_eval_chain_samples_per_s"] = 64 / (e0.elapsed_time(e1) / 1e3)
out["f3_vlb_bits_per_dim_mean"] = float(res["vlb"].mean())
if "--cpu" in sys.argv:
    from oracle import ddpm_oracle as O
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    buf = O.schedule_buffers("linear", 1000)
    zc = z[:4].cpu()
    cfg = dict(tc.C3, T=1000)
    t0 = time.perf_counter()
    steps = 4
    with torch.no_grad():
        for k in range(steps):
            t = torch.full((4,), 999 - k, dtype=torch.long)
            eps = torch.randn_like(zc)
            z_t = O.q_sample(buf, zc, t, eps)
            eh = O.unet_forward(sd, cfg, z_t, t, "latent_model.")
            O.vlb_terms(buf, zc, z_t, t, eh)
    dt = (time.perf_counter() - t0) / steps
    out["f3_cpu_oracle_samples_per_s"] = 4 / (dt * 1000)
    out["f3_cpu_sample"] = f"{steps} steps on 4 latents, {torch.get_num_threads()} threads, extrapolated x1000/{steps}"
print(json.dumps(out, indent=1))
