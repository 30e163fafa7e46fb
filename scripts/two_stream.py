"""Experiment: one chain of B samples vs two concurrent chains of B/2 on two streams (graph replays interleaved), C3 latent.
python scripts/two_stream.py [B] [steps]"""
import copy, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import downsampled_diffusion_b200 as dd
from tests import common as tc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda:0")
m = tc.build_model(dict(tc.C3, precision="bf16"), dd, "dddpm_ae", device="cuda:0").to(dev).eval()
m2 = copy.deepcopy(m)


def plan_of(model, b):
    p = model.sampling_plan((b, 8, 32, 32))
    p.prepare()
    p.t_dev.fill_(999)
    p.noise.normal_()
    return p


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


with torch.no_grad():
    full = plan_of(m, B)
    ha, hb = plan_of(m, B // 2), plan_of(m2, B // 2)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def one():
        for _ in range(N):
            full.graph.replay()

    def two():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        for _ in range(N):
            with torch.cuda.stream(s1):
                ha.graph.replay()
            with torch.cuda.stream(s2):
                hb.graph.replay()
        cur.wait_stream(s1); cur.wait_stream(s2)

    def half():
        for _ in range(N):
            ha.graph.replay()
    for p in (full, ha, hb):
        p.t_dev.fill_(999)
    one(); two()
    for name, fn in (("one chain of %d" % B, one), ("two concurrent chains of %d" % (B // 2), two), ("one chain of %d" % (B // 2), half)):
        for p in (full, ha, hb):
            p.t_dev.fill_(999)
        print(name, "ms per step: %.4f" % (timed(fn) / N))
