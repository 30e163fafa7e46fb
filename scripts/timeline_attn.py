"""In-kernel timeline of the dd_linattn_mix launches of the C3 step (instrumented build: -DDD_ATTN_TIMELINE=1)."""
import os, sys, torch, numpy as np
os.environ.setdefault("DD_NO_FORK", "1")          # per-op replays: keep every launch on one stream
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import downsampled_diffusion_b200 as dd
from downsampled_diffusion_b200 import _lib as L
from tests import common as tc
dev = torch.device("cuda:0")
m = tc.build_model(dict(tc.C3, precision="bf16"), dd, "dddpm_ae", device="cuda:0").to(dev).eval()
plan = m.sampling_plan((64, 8, 32, 32)); plan.prepare()
eng = plan.eng
idx = [i for i, n in enumerate(eng.op_names) if n == "dd_linattn_mix"]
buf = torch.zeros(4096 * 8, dtype=torch.int64, device=dev)
names = ["start", "dep_wait", "context", "merged", "normalised", "end"]
for which in (0, 1, 2, 3):          # 32x32, 16x16, 8x8, 4x4
    prev, op = eng.ops[idx[which] - 1], eng.ops[idx[which]]
    for _ in range(3): prev(); op()
    torch.cuda.synchronize()
    buf.zero_()
    L.lib().dd_debug_set_attn_timeline(buf.data_ptr())
    prev(); op()
    torch.cuda.synchronize()
    L.lib().dd_debug_set_attn_timeline(None)
    t = buf.cpu().numpy().reshape(-1, 8)
    t = t[t[:, 0] > 0]
    rel = t[:, :6] - t[:, [0]]
    print(f"attention #{which}: {len(t)} CTAs, span {int(t[:, 5].max() - t[:, 0].min())} clk; median offsets:",
          {n: int(np.median(rel[:, i])) for i, n in enumerate(names)}, "; start spread p50/max:", int(np.median(t[:, 0] - t[:, 0].min())), int((t[:, 0] - t[:, 0].min()).max()))
