"""Warm per-op GPU time of the C3 step program: each op is captured 10x into its own CUDA graph and replayed,
so CPU launch overhead is excluded and operands are L2-resident like in the real step."""
import os, sys, collections, torch
os.environ.setdefault("DD_NO_FORK", "1")          # per-op replays: keep every launch on one stream
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import downsampled_diffusion_b200 as dd
from tests import common as tc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
m = tc.build_model(dict(tc.C3, precision="bf16"), dd, "dddpm_ae", device="cuda:0").to(dev).eval()
plan = m.sampling_plan((B, 8, 32, 32)); plan.prepare()
eng = plan.eng
REP = 10
res = []
for i, op in enumerate(eng.ops):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        op()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP): op()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    res.append((eng.op_names[i], e0.elapsed_time(e1) / (5 * REP) * 1000))
agg = collections.defaultdict(lambda: [0, 0.0]); ci = 0
for i, (n, us) in enumerate(res):
    extra = ""
    if n == "dd_conv_tc":
        fl = eng.conv_tc_flops[ci]; ci += 1
        extra = f"  {fl / us / 1e6:7.1f} TFLOP/s"
    print(f"{i:3d} {n:22s} {us:8.2f} us{extra}")
    agg[n][0] += 1; agg[n][1] += us
tot = sum(v[1] for v in agg.values())
print("sum of warm per-op times: %.1f us" % tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:22s} n={v[0]:3d} {v[1]:8.1f} us {100 * v[1] / tot:5.1f}%")
