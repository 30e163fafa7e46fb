"""In-kernel timeline of single conv_tc launches (clock64 stamps per CTA)."""
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
idx = [i for i, n in enumerate(eng.op_names) if n == "dd_conv_tc"]
buf = torch.zeros(4096 * 32, dtype=torch.int64, device=dev)
names = ["start", "prologue", "pdl_wait", "first_data", "last_mma", "acc_ready", "epi_done"]
for which in [int(a) for a in sys.argv[1:]] or (2, 16, 23):          # 3x3@32 (halo), 3x3@8 (8x8 halo form), 3x3@4 (split-K)
    op = eng.ops[idx[which]]
    for _ in range(3): op()
    torch.cuda.synchronize()
    buf.zero_()
    L.lib().dd_debug_set_timeline(buf.data_ptr())
    op()
    torch.cuda.synchronize()
    L.lib().dd_debug_set_timeline(None)
    t = buf.cpu().numpy().reshape(-1, 32)
    t = t[t[:, 0] > 0]
    if (t[:, 11] > 0).any() and (t[:, 10] > 0).any() and (t[:, 13] > 0).any() and not (t[:, 1] > 0).any() and (t[:, 7] > 0).any() and (t[:, 12] >= 0).all() and (t[:, 10] > t[:, 13]).all():
        # persistent GEMM: 5 / 6 = item 0 accumulator ready / stored, 11 = item 1 parameters staged, 13 / 10 = item 1 ready / stored, 12 / 14 = last item
        d = lambda a, b: int(np.median((t[:, a] - t[:, b])[(t[:, a] > 0) & (t[:, b] > 0)])) if ((t[:, a] > 0) & (t[:, b] > 0)).any() else -1
        print(f"conv #{which}: persistent GEMM, {len(t)} CTAs")
        print("   pdl->first data %d, item0: MMAs issued %d after first data, acc ready->stored %d" % (d(3, 2), d(4, 3), d(6, 5)))
        print("   item1: item0 stored->parameters staged %d, ->acc ready %d, acc ready->stored %d" % (d(11, 6), d(13, 11), d(10, 13)))
        print("   first data -> last item stored %d" % d(14, 3))
        med = lambda c: int(np.median(t[:, c]))
        print("   whole CTA (%d items): MMA warp waited %d clk for operand stages, %d for a free accumulator; epilogue warp waited %d for accumulators" % (med(20), med(8), med(9), med(15)))
        continue
    if (t[:, 14] > 0).any() or (t[:, 13] > 0).any():          # persistent kernel (conv_tc_persist.cu): its own stamp set
        d = lambda a, b: int(np.median((t[:, a] - t[:, b])[(t[:, a] > 0) & (t[:, b] > 0)])) if ((t[:, a] > 0) & (t[:, b] > 0)).any() else -1
        print(f"conv #{which}: persistent, {len(t)} CTAs")
        print("   start->pdl %d, pdl->first data %d, item0 MMA issue span %d, item0 MMAs issued->acc ready %d" % (d(2, 0), d(3, 2), d(4, 3), d(5, 4)))
        print("   item0 epilogue: drain %d, stats+atomics %d, wait for the image %d, normalise+store %d" % (d(10, 5), d(11, 10), d(12, 11), d(6, 12)))
        print("   item0 end -> item1 acc ready %d; first data -> last item's MMAs issued %d; -> last item's epilogue done %d" % (d(13, 6), d(7, 3), d(14, 3)))
        if (t[:, 16] > 0).any():
            print("   statistics warp, item0, from the accumulator being ready: partial sums of all epilogue warps in %d, packet stored %d, all packets of the image seen %d, scale / shift staged + epilogue released %d" % (d(16, 5), d(17, 5), d(18, 5), d(19, 5)))
        med = lambda c: int(np.median(t[:, c]))
        print("   whole CTA: MMA warp waited %d clk for halos, %d for weight tiles, %d for a free accumulator; epilogue waited %d for accumulators" % (med(1), med(8), med(9), med(15)))
        continue
    t0 = t[:, 0].min()
    rel = t[:, :7] - t[:, [0]]
    print(f"conv #{which}: {len(t)} CTAs; kernel span {int(t[:, 6].max() - t0)} clk")
    print("   median per-CTA offsets from its own start:", {n: int(np.median(rel[:, i])) for i, n in enumerate(names)})
    print("   CTA start offsets (from first CTA): p50 %d p90 %d max %d" % tuple(np.percentile(t[:, 0] - t0, [50, 90, 100])))
    print("   epilogue split: drain+stage %d, stats %d, write-out %d" % (np.median(t[:, 10] - t[:, 5]), np.median(t[:, 11] - t[:, 10]), np.median(t[:, 6] - t[:, 11])))
    if (t[:, 12] > 0).any():      # fused GroupNorm epilogue: stamp 12 = statistics of the image known (after the cluster barrier)
        print("   fused GN: barrier + peer sums %d, normalise + write-out %d" % (np.median(t[:, 12] - t[:, 11]), np.median(t[:, 6] - t[:, 12])))
    print("   CTA end offsets (from first CTA start): p10 %d p50 %d p90 %d max %d" % tuple(np.percentile(t[:, 6] - t0, [10, 50, 90, 100])))
    d = rel[:, 4] - rel[:, 3]
    print("   mainloop (first_data -> last_mma): median %d clk, epilogue (acc_ready -> done): median %d clk" % (np.median(d), np.median(rel[:, 6] - rel[:, 5])))
