#!/usr/bin/env python
"""Headline benchmark: samples/sec for full T=1000 ancestral sampling of dDDPM x3 (BASELINE.json configs[2]).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

A "step" is ONE full chain over one batch: 1000 x (U-Net + posterior update) on the 8x32x32 latent of a
batch of 64 per GPU, then the up-sampling net to 3x256x256.  Weak scaling: every rank owns 64 samples, no
communication during the chain, one gather of the images at the end.  One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_STEPS = 1000
LATENT = (8, 32, 32)
IMAGE = (3, 256, 256)
UNET_GFLOP = 4.595          # per sample per step
UPNET_GFLOP = 8.746


def c3_config(precision: str) -> dict:
    from tests import common as tc
    return dict(tc.C3, unet_dropout=0.1, precision=precision)      # reference defaults (train.py:20-47)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:      # noqa: BLE001
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# -------------------------------------------------------------------------------------------------
def cpu_kind() -> str:
    from oracle import build_ref
    return "reference" if build_ref.available() else "port"


def cpu_reference_rate(n_steps: int, batch: int, threads: int):
    """The reference's own CPU path on the host cores: `n_steps` ancestral steps on a batch of `batch` latents + one
    up-net pass, extrapolated to T=1000.  With oracle/_ref present (oracle/build_ref.py) this is the UNMODIFIED reference
    (`DownsampleDDPMAutoencoder.p_sample` / `rescaled_upsample`, models/diffusion/ddpm.py:203-227, dddpm.py:103-112);
    otherwise the oracle port of the same algorithm."""
    from oracle import build_ref
    from tests import common as tc
    torch.set_num_threads(threads)
    cfg = c3_config("fp32")
    if build_ref.available():
        ref = build_ref.load()
        model = tc.build_model(cfg, ref, "dddpm_ae").eval()
        g = torch.Generator().manual_seed(0)
        img = torch.randn(batch, *LATENT, generator=g)
        with torch.no_grad():
            model.p_sample(img, torch.full((batch,), T_STEPS - 1, dtype=torch.long))      # warm-up step
            t0 = time.perf_counter()
            for i in range(n_steps):
                img = model.p_sample(img, torch.full((batch,), T_STEPS - 1 - i, dtype=torch.long))
            t1 = time.perf_counter()
            model.rescaled_upsample(img[:max(1, batch // 4)])
            t2 = time.perf_counter()
        step_s = (t1 - t0) / n_steps
        up_s = (t2 - t1) / max(1, batch // 4) * batch
        return batch / (step_s * T_STEPS + up_s), step_s, up_s
    from oracle import ddpm_oracle as O
    import downsampled_diffusion_b200 as dd
    model = tc.build_model(cfg, dd, "dddpm_ae")
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    buf = O.schedule_buffers("linear", T_STEPS)
    g = torch.Generator().manual_seed(0)
    noises = [torch.randn(batch, *LATENT, generator=g) for _ in range(n_steps + 1)]
    with torch.no_grad():
        O.p_sample_loop(sd, cfg, buf, noises[:2], t_end=T_STEPS - 1)      # warm-up step
        t0 = time.perf_counter()
        z = O.p_sample_loop(sd, cfg, buf, noises, t_end=T_STEPS - n_steps)
        t1 = time.perf_counter()
        O.rescaled_upsample(sd, cfg, z[:max(1, batch // 4)])
        t2 = time.perf_counter()
    step_s = (t1 - t0) / n_steps
    up_s = (t2 - t1) / max(1, batch // 4) * batch
    chain_s = step_s * T_STEPS + up_s
    return batch / chain_s, step_s, up_s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    for _ in range(max(1, args.steps)):
        rate, step_s, up_s = cpu_reference_rate(args.cpu_steps, args.cpu_batch, threads)
        vals.append(rate)
    rate = sum(vals) / len(vals)
    kind = cpu_kind()
    sample = (f"{args.cpu_steps} ancestral steps + up-net on a batch of {args.cpu_batch} latents per bench step, "
              f"extrapolated x{T_STEPS}/{args.cpu_steps}; " + ("the unmodified reference from oracle/_ref (torch CPU)" if kind == "reference"
                                                               else "oracle port of the reference (oracle/_ref absent)"))
    line = {"impl": "reference", "metric": "samples/sec, full T=1000 dDDPM x3 sampling", "value": rate, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * args.cpu_batch / rate,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3: dDDPM x3, latent 8x32x32 -> 3x256x256, T=1000 (CPU sample)", "cpu_batch": args.cpu_batch},
            "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
def conv_roofline(model, plan, pk):
    """Average tcgen05 conv launch: algorithmic FLOPs / CUDA-event duration, measured live by replaying
    the U-Net step's dd_conv_tc launches (and only those) back to back from a captured graph."""
    eng = plan.eng
    idx = [i for i, n in enumerate(eng.op_names) if n == "dd_conv_tc"]
    flops = eng.conv_tc_flops          # algorithmic 2*M*N*K per launch, same order as idx
    assert len(flops) == len(idx)
    from downsampled_diffusion_b200 import _lib as L

    def conv_ops():
        # the persistent convolutions exchange GroupNorm sums through a workspace that must start out zero (a real step begins
        # with the same memset): it is part of every replay here, so the waits for the other tiles of an image are real
        if eng.stats_arena is not None:
            L.call("dd_zero", eng.stats_arena.data_ptr(), eng.stats_arena.numel() * 4, L.stream())
        for i in idx_run:
            eng.ops[i]()
    # a res_conv on a side stream (per-GPU batches of 32 and less) is a fork op named like a convolution; its join must be replayed too
    idx_run = [i for i, n in enumerate(eng.op_names) if n in ("dd_conv_tc", "join")]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        conv_ops()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        conv_ops()
    for _ in range(3):
        g.replay()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if eng.stats_arena is not None:
        eng.stats_arena.zero_()          # leave the arena as a chain expects it
    tf = sum(flops) / (ms * 1e-3) / 1e12
    traffic, tsrc = None, None        # DRAM bytes per launch from the committed ncu capture of the same launches (profiles/)
    tpath = os.path.join(ROOT, "profiles", "conv_tc_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic, tsrc = tj.get("avg_dram_bytes_per_launch"), tj.get("source")
    return {"bound": "tensor", "achieved": tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": tf / pk["tf_sustained"],
            "traffic": traffic, "traffic_source": tsrc, "kernel": "conv_tc_kernel", "launches_per_unet_step": len(idx),
            "avg_launch_us": 1000.0 * ms / len(idx), "conv_ms_per_unet_step": ms,
            "peak_source": f"{pk['src']} sustained bf16 (kernel timed inside a long step)"}


def step_time_ms(plan, reps=50):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    plan.t_dev.fill_(T_STEPS - 1)
    for _ in range(5):
        plan.graph.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        plan.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def time_chains(chain, noise, n, barrier, final_gather, x_host=None):
    """CUDA-event time of `n` full chains (ms), barrier + synchronize on both sides."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(n):
        x, _ = chain(noise)
        if x_host is not None:
            x_host.copy_(x, non_blocking=True)
        final_gather(x)
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def run_ours(args):
    import torch.distributed as dist
    import downsampled_diffusion_b200 as dd
    from downsampled_diffusion_b200 import _lib
    from tests import common as tc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    pk = peaks()
    # BASELINE configs[2] as written: ONE batch of `--batch` (64) samples, sharded over the N GPUs of the box (strong scaling:
    # 64 / N per GPU, no communication during the chain, one gather of the images).  The saturating figure the survey also
    # asks for (64 samples PER GPU, weak) is measured after it and reported under "weak_64_per_gpu".
    G = args.batch
    assert G % world == 0, f"global batch {G} must divide over {world} GPUs"
    B = G // world
    cfg = c3_config(args.precision)
    model = tc.build_model(cfg, dd, "dddpm_ae", device=str(dev)).to(dev).eval()
    model.downsample.precision = model.upsample.precision = args.precision

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_case(Bc):
        shape = (Bc, *LATENT)
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        noise_dev = torch.randn(T_STEPS + 1, *shape, generator=gen, device=dev)        # 2.1 GB at 64 per GPU
        gathered = [torch.empty(Bc, *IMAGE, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None

        @torch.no_grad()
        def chain(noise):
            z = model.p_sample_loop(shape, noise=noise)
            x = model.rescaled_upsample(z)
            return x, z

        def final_gather(x):
            if world > 1:
                dist.gather(x, gathered, dst=0)
        return shape, noise_dev, chain, final_gather

    shape, noise_dev, chain, final_gather = make_case(B)
    noise_host = torch.empty(noise_dev.shape, dtype=torch.float32, pin_memory=True)
    noise_host.copy_(noise_dev)
    x_host = torch.empty(B, *IMAGE, dtype=torch.float32, pin_memory=True)
    torch.cuda.synchronize()

    for _ in range(args.warmup):
        x, _ = chain(noise_dev)
        final_gather(x)
    barrier()

    # ---- device-resident throughput ---------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    n0 = _lib.launches()
    ms = time_chains(chain, noise_dev, args.steps, barrier, final_gather)
    launches = _lib.launches() - n0
    sampler.stop_flag.set()
    sampler.join(timeout=3)

    # ---- end to end: host noise in, host images out -------------------------------------------------
    chain(noise_host)            # one warm-up through the host path
    ms_e2e = time_chains(chain, noise_host, args.steps, barrier, final_gather, x_host)

    # ---- the default call: model.sample(B) with the noise drawn on the device (what a drop-in user runs) ----
    @torch.no_grad()
    def default_call(_):
        x, z = model.sample(B)
        return x, z
    n_def = max(1, args.steps // 4)
    default_call(None)
    ms_def = time_chains(default_call, None, n_def, barrier, final_gather, x_host)

    # ---- weak figure: 64 samples per GPU (only differs from the above for N > 1) --------------------
    ms_weak, n_weak = None, max(2, args.steps // 4)
    if world > 1:
        del noise_dev, noise_host
        shape_w, noise_w, chain_w, gather_w = make_case(G)
        for _ in range(2):
            chain_w(noise_w)
        ms_weak = time_chains(chain_w, noise_w, n_weak, barrier, gather_w)
        del noise_w

    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_def, ms_weak], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_def, ms_weak = (float(v) for v in t)

    # the step replayed on its own and its convolution launches replayed on their own: measured here, in the state the chains left
    # (after the training sub-record the same replay read 4 % slower than the chain's own average, r02_ev4)
    step_ms = roof = None
    if rank == 0:
        plan = model.sampling_plan(shape)
        step_ms = step_time_ms(plan)
        roof = conv_roofline(model, plan, pk)

    train = None
    if not args.no_train:
        del chain
        torch.cuda.empty_cache()
        train = train_record(args, world, rank, dev, pk, sub=True)

    if rank == 0:
        total = G * args.steps
        value = total / (ms / 1000.0)
        e2e = total / (ms_e2e / 1000.0)
        unet_tf = UNET_GFLOP * B / step_ms            # GFLOP / ms = TFLOP/s
        line = {
            "metric": "samples/sec, full T=1000 dDDPM x3 sampling", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": "C3 (BASELINE configs[2]): dDDPM x3, T=1000 ancestral chain on latent 8x32x32 + up-net to 3x256x256, "
                                   "one batch of 64 sharded over the GPUs",
                       "batch_per_gpu": B, "global_batch": G, "weights": "random init seed 0, 22.67M params",
                       "parallelism": f"batch-sharded x{world}, no comm in chain, final gather",
                       "l2": "per-chain noise stream (33 MB per sample) exceeds L2 at every batch size; weights (45 MB bf16) L2-resident by design",
                       "cuda_graph": "one graph per ancestral step, replayed 1000x"},
            "unet_step_ms": step_ms, "unet_step_tflops": unet_tf, "unet_step_frac_of_sustained_peak": unet_tf / pk["tf_sustained"],
            "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": int((T_STEPS + 1) * B * 8192 * 4),
                    "d2h_bytes_per_step": int(x_host.numel() * 4)},
            "default_call": {"value": G * n_def / (ms_def / 1000.0), "unit": "samples/s", "chains": n_def,
                             "what": "model.sample(B) with noise=None: start image and per-step z drawn with torch.randn on the device, images read back"},
            "roofline": roof, "clocks": sampler.summary(),
        }
        if ms_weak is not None:
            line["weak_64_per_gpu"] = {"value": world * G * n_weak / (ms_weak / 1000.0), "unit": "samples/s", "batch_per_gpu": G,
                                       "global_batch": world * G, "chains": n_weak, "scaling": "weak"}
        if train is not None:
            line["train"] = train
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            rate, step_s, up_s = cpu_reference_rate(args.cpu_steps, args.cpu_batch, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": threads, "kind": cpu_kind(),
                                    "sample": f"{args.cpu_steps} ancestral steps ({step_s * 1e3:.1f} ms/step) + up-net on {args.cpu_batch} latents, extrapolated to T=1000"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------------------
# Training workload (BASELINE.json configs[3], C4): dDDPM x3 256x256, p_losses forward + backward, gradient
# all-reduce (N > 1), clip, Adam, EMA.  `python bench.py --workload train [--gpus N]`; the driver's default run
# is the sampling workload above.
TRAIN_GFLOP = 57.5          # per sample, forward + backward (torch flop counter on the reference)
# Algorithmic HBM bytes per sample of a layer-per-kernel training step ("training roofline"): the conv /
# norm / attention outputs of the three networks are 67.0 M fp32 elements per sample (down-net 26.2 M, up-net 36.7 M, U-Net
# 4.1 M; activations fused into their producers); each is written and read once in forward, read once in backward, and its
# gradient is written and read once: 5 passes x 4 B.
TRAIN_GB_PER_SAMPLE = 67.0e6 * 4 * 5 / 1e9


def cpu_train_rate(batch: int, threads: int):
    """Reference algorithm of one training step (dddpm.py:155-177 through the oracle port + torch autograd) on the host."""
    from oracle import ddpm_oracle as O
    from tests import common as tc
    import downsampled_diffusion_b200 as dd
    torch.set_num_threads(threads)
    cfg = dict(tc.C3, unet_dropout=0.0)
    model = tc.build_model(cfg, dd, "dddpm_ae")
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and k in dict(model.named_parameters())) for k, v in model.state_dict().items()}
    buf = O.schedule_buffers("linear", T_STEPS)
    x = tc.rand_pm1(5, batch, *IMAGE)
    t = torch.randint(0, T_STEPS, (batch,), generator=torch.Generator().manual_seed(6))
    eps = tc.randn(7, batch, *LATENT)
    t0 = time.perf_counter()
    obj, _ = O.dddpm_losses(sd, cfg, buf, x, t, eps, autoencoder=True)
    obj.backward()
    return batch / (time.perf_counter() - t0)


def train_record(args, world, rank, dev, pk, sub=False):
    """Times the C4 training step on every rank (process group already initialised for world > 1); returns the record on
    rank 0, None elsewhere.  sub=True: the `train` sub-record of the sampling line (bounded: --train-steps steps)."""
    import torch.distributed as dist
    import downsampled_diffusion_b200 as dd
    from downsampled_diffusion_b200 import _lib, parallel
    from tests import common as tc

    B = args.train_batch
    steps = args.train_steps if sub else args.steps
    warmup = min(args.warmup, 5) if sub else args.warmup
    cfg = dict(tc.C3, unet_dropout=0.1, precision="fp32" if args.train_precision == "fp32" else "bf16")
    model = tc.build_model(cfg, dd, "dddpm_ae", device=str(dev)).to(dev).train()
    model.downsample.precision = model.upsample.precision = cfg["precision"]
    ema = dd.EMA(model, decay=0.995)
    # the trainer's clip_grad_norm_(1.0) -> Adam.step -> EMA.update (trainer_ddpm.py:243-254) as the package's fused optimizer
    opt = dd.Adam(model.parameters(), lr=2e-4, max_grad_norm=1.0)
    opt.attach_ema(ema, model)
    params = [p for p in model.parameters()]
    x_host = torch.empty(B, *IMAGE, dtype=torch.float32, pin_memory=True)
    x_host.copy_(tc.rand_pm1(100 + rank, B, *IMAGE))
    x_dev = x_host.to(dev)
    reducer = parallel.GradReducer(model) if world > 1 else None

    def step(x):
        if reducer is not None:
            reducer.arm()
        obj, _ = model(x)
        obj.backward()
        if reducer is not None:
            reducer.finish()
        opt.step(ema="update")
        opt.zero_grad()
        return obj

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step(x_dev)
    barrier()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    n0 = _lib.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step(x_dev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launches() - n0
    sampler.stop_flag.set()
    sampler.join(timeout=3)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e2.record()
    loss = 0.0
    for _ in range(steps):
        loss = float(step(x_host.to(dev, non_blocking=True)))        # H2D of the batch, D2H of the loss, every step
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    del model, ema, opt
    if rank != 0:
        return None
    total = world * B * steps
    value, e2e = total / (ms / 1e3), total / (ms_e2e / 1e3)
    step_ms = ms / steps
    tf = TRAIN_GFLOP * B / step_ms           # per GPU, GFLOP/ms = TFLOP/s
    gbs = TRAIN_GB_PER_SAMPLE * B / (step_ms * 1e-3)
    line = {"metric": "samples/sec, dDDPM x3 256x256 training step (p_losses fwd+bwd, clip, Adam, EMA)", "value": value,
            "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.train_precision == "fp32" else "tf32",
            "data": "synthetic",
            "config": {"workload": "C4 (BASELINE configs[3]): dDDPM x3 256x256 training, x ~ U(-1,1), dropout 0.1",
                       "batch_per_gpu": B, "global_batch": world * B,
                       "parallelism": f"data-parallel x{world}, per-network NCCL gradient all-reduce (AVG) launched from the backward pass on a side stream",
                       "l2": "activations of one step (tens of GB) exceed L2"},
            "gpu_launches": launches, "last_loss": loss,
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": int(x_host.numel() * 4), "d2h_bytes_per_step": 4},
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": None,
                         "algorithmic_bytes_per_sample": TRAIN_GB_PER_SAMPLE * 1e9,
                         "kernel": "whole step (the resampling nets' 32/64-channel convolutions at 64^2..256^2 are HBM-bound, SURVEY.md 8(d)); "
                                   "bytes = every activation written once and read once per consumer, forward + backward, fp32",
                         "tensor_tflops": tf, "tensor_frac_of_sustained_bf16": tf / pk["tf_sustained"],
                         "peak_source": pk["src"] + " HBM copy bandwidth"},
            "clocks": sampler.summary()}
    return line


def run_train(args):
    import torch.distributed as dist
    from downsampled_diffusion_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            threads = os.cpu_count() or 1
            rate = cpu_train_rate(args.cpu_batch_train, threads)
            print(json.dumps({"impl": "reference", "metric": "samples/sec, dDDPM x3 256x256 training step", "value": rate, "unit": "samples/s",
                              "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "dtype": "f32",
                              "data": "synthetic", "config": {"workload": "C4 (CPU sample)", "cpu_batch": args.cpu_batch_train},
                              "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
                                               "sample": f"one forward+backward on {args.cpu_batch_train} images"},
                              "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    line = train_record(args, world, rank, dev, peaks())
    if rank == 0:
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            rate = cpu_train_rate(args.cpu_batch_train, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
                                    "sample": f"one forward+backward (oracle port + torch autograd) on {args.cpu_batch_train} images"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="samples per GPU")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="sample", choices=["sample", "train"])
    ap.add_argument("--train-batch", type=int, default=32, help="images per GPU per training step")
    ap.add_argument("--train-precision", default="tf32", choices=["tf32", "fp32"], help="tf32: tcgen05 kind::tf32 convolutions; fp32: CUDA-core validation mode")
    ap.add_argument("--cpu-batch-train", type=int, default=2)
    ap.add_argument("--train-steps", type=int, default=20, help="steps of the C4 training sub-record of the sampling line")
    ap.add_argument("--no-train", action="store_true", help="skip the C4 training sub-record")
    args = ap.parse_args()
    if args.workload == "train":
        run_train(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
