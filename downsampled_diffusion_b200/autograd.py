"""Training path: forward + hand-written backward programs behind torch.autograd.Function.

The reference trains through autograd over ATen ops (trainers/trainer_ddpm.py:225-229 calls
`model(x)` then `.backward()`).  Here the U-Net (unet.py:74-104) and the down/up-sampling nets
(convblocks.py:133-159) are lowered to an fp32 launch list whose backward is a second launch list of
libddb200 kernels (backward.cu + dd_conv_direct for input gradients); torch.autograd only sees one
Function per network.  Activations stay in the program's buffers between forward and backward.

Training programs are fp32 (CUDA-core convolutions): gradients match the reference to ~1e-5; the bf16
tensor-core backward is the next step.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _lib as L
from .engine import Act, EngineCache, Program, ensure_lazy, params_version
from .relayout import EXACT, GatherPlan, apply_codes, codes_from_probes

GN_EPS = 1e-5


class TrainProgram(Program):
    """fp32 forward launch list + its backward launch list, built together."""

    def __init__(self, module: torch.nn.Module, B: int, tf32: bool = False):
        super().__init__(module, B, "fp32")
        ensure_lazy()
        self.scratch_need: Dict[str, int] = {}
        self.scratch_buf: Dict[str, torch.Tensor] = {}
        self.tf32 = tf32                                  # 3x3 / 1x1 convolutions (forward + input gradient) on tcgen05 kind::tf32
        self.bops: List[Callable[[], None]] = []
        self.bbuilders: List[Callable[[], None]] = []
        self.gbuf: Dict[int, torch.Tensor] = {}          # id(act.t) -> gradient tensor (same shape, fp32)
        self.gwritten = set()
        self.pg_specs: List[Tuple[torch.nn.Parameter, int, Tuple[int, ...], Callable]] = []
        self.pg_total = 0
        self.pg_arena: Optional[torch.Tensor] = None
        self._emit_bwd = False
        # dropout: one seed per forward call, drawn on the device (torch's CUDA generator) and read by the kernels from
        # device memory, so the launch lists hold no per-call host values and can be replayed as CUDA graphs
        self.seed_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.n_dropout = 0
        self.graphs = None                                # (forward graph, launches, backward graph, launches)

    # ---- launch recording -------------------------------------------------------------------
    def add(self, name: str, *args) -> None:
        if not self._emit_bwd:
            return super().add(name, *args)
        fn = getattr(L.lib(), name)

        def op():
            L._Counter.n += 1
            rc = fn(*args, L.stream())
            if rc != 0:
                L.check(rc, name)
        self.bops.append(op)

    def add_host(self, fn: Callable[[], None]) -> None:
        """A tiny torch-side glue step (slice copy) in the current list."""
        if self._emit_bwd:
            self.bops.append(fn)
        else:
            self.ops.append(fn)
            self.op_names.append("host_glue")

    def scratch(self, key: str, numel: int) -> "_ScratchPtr":
        """fp32 scratch shared by every launch that asks for `key` (launches are stream-ordered); sized at the end of the build."""
        self.scratch_need[key] = max(self.scratch_need.get(key, 0), int(numel))
        return _ScratchPtr(self, key)

    def on_backward(self, builder: Callable[[], None]) -> None:
        self.bbuilders.append(builder)

    def build_backward(self) -> None:
        self._emit_bwd = True
        for b in reversed(self.bbuilders):
            b()
        self._emit_bwd = False
        self.pg_arena = torch.zeros(max(self.pg_total, 1), dtype=torch.float32, device=self.device)
        for key, n in self.scratch_need.items():
            self.scratch_buf[key] = torch.empty(n, dtype=torch.float32, device=self.device)
        self.repack_plan = self.grad_plan = None
        if not os.environ.get("DD_NO_RELAYOUT_PLAN"):
            with torch.no_grad():
                self._build_repack_plan()
                self._build_grad_plan()

    # ---- tabulated re-layouts (relayout.py): weights -> packed buffers, packed gradients -> parameter gradients --------
    def _build_repack_plan(self) -> None:
        params = list(self.module.parameters())
        if not params or not self.packers or max(p.numel() for p in params) >= EXACT or len(params) >= EXACT:
            return
        dev = self.device
        cand = [i for i, b in enumerate(self.packed_bufs) if b.dtype == torch.float32 and b.is_contiguous()]
        saved = [p.data for p in params]
        snaps = []
        try:
            for which in (0, 1):            # sources hold "tensor ordinal + 1", then "element index"
                for i, p in enumerate(params):
                    p.data = (torch.full(p.shape, float(i + 1), dtype=torch.float32, device=dev) if which == 0 else
                              torch.arange(p.numel(), dtype=torch.float32, device=dev).view(p.shape))
                for b in self.packed_bufs:
                    b.zero_()
                for pk in self.packers:
                    pk()
                snaps.append([self.packed_bufs[i].detach().clone() for i in cand])
        finally:
            for p, d in zip(params, saved):
                p.data = d
        for b in self.packed_bufs:
            b.zero_()
        for pk in self.packers:             # the real weights again: also the reference the tables are checked against
            pk()
        numels = torch.tensor([p.numel() for p in params], dtype=torch.int64, device=dev)
        offsets = torch.cumsum(numels, 0) - numels
        flat = torch.cat([p.detach().reshape(-1).float() for p in params])
        dsts, codes, fast = [], [], set()
        for j, i in enumerate(cand):
            c = codes_from_probes(snaps[0][j], snaps[1][j], numels)
            if c is not None and torch.equal(apply_codes(c, flat, offsets), self.packed_bufs[i].reshape(-1)):
                dsts.append(self.packed_bufs[i])
                codes.append(c)
                fast.add(i)
        if not dsts:
            return
        self.repack_plan = GatherPlan(dsts, codes)
        self.repack_params = params
        self.slow_packers = [pk for i, pk in enumerate(self.packers) if i not in fast]

    def refresh_weights(self) -> None:
        if getattr(self, "repack_plan", None) is None:
            return super().refresh_weights()
        v = params_version(self.module)
        if v != self.weights_version:
            self.repack_plan.run(tuple(p.data_ptr() for p in self.repack_params))
            for pk in self.slow_packers:
                pk()
            self.weights_version = v
            self.pack_epoch += 1

    def _build_grad_plan(self) -> None:
        total = self.pg_total
        if total == 0 or total + 1 >= EXACT * 4096:
            return
        dev = self.device
        count: Dict[int, int] = {}
        for param, _, _, _ in self.pg_specs:
            count[id(param)] = count.get(id(param), 0) + 1
        cand = [k for k, (param, _, _, _) in enumerate(self.pg_specs) if count[id(param)] == 1]
        pos1 = torch.arange(1, total + 1, dtype=torch.int64, device=dev)          # position + 1: zero stays "no source"
        probes = ((pos1 >> 12).float(), (pos1 & 4095).float())
        check = (pos1.float() * 0.6180339887).frac() - 0.5            # validation data (no RNG draw: callers seed around the build)

        def through(arena, k):
            param, off, shape, to_param = self.pg_specs[k]
            n = 1
            for s_ in shape:
                n *= s_
            return to_param(arena[off:off + n].view(shape)).reshape(-1)
        codes, self.grad_fast, fast, start = [], [], set(), 0
        for k in cand:
            param, off, shape, _ = self.pg_specs[k]
            hi, lo = through(probes[0], k), through(probes[1], k)
            if hi.numel() != param.numel():
                continue
            hi_i, lo_i = hi.round().long(), lo.round().long()
            pos = hi_i * 4096 + lo_i                                               # = source position + 1, 0 = constant zero
            got = torch.where(pos > 0, check[(pos - 1).clamp(min=0, max=total - 1)], torch.zeros((), device=dev))
            ok = (hi == hi_i) & (lo == lo_i) & (pos >= 0) & (pos <= total) & (got == through(check, k))
            if not bool(ok.all()):                                                  # one host synchronisation per parameter
                continue
            c = torch.where(pos > 0, (1 << 32) | (pos - 1), torch.zeros_like(pos))
            n = param.numel()
            pad = (-n) % 4                                                          # every gradient starts 16-byte aligned
            codes.append(torch.cat([c, torch.zeros(pad, dtype=torch.int64, device=dev)]) if pad else c)
            self.grad_fast.append((param, start, n))
            fast.add(k)
            start += n + pad
        if not codes:
            return
        self.grad_plan = GatherPlan([None], [torch.cat(codes)], first_dst_per_call=True)
        self.slow_specs = [sp for k, sp in enumerate(self.pg_specs) if k not in fast]

    # ---- gradient buffers ---------------------------------------------------------------------
    def grad(self, a: Act) -> torch.Tensor:
        g = self.gbuf.get(id(a.t))
        if g is None:
            g = self.empty(*a.t.shape, dtype=torch.float32)
            self.gbuf[id(a.t)] = g
        return g

    def has_grad(self, a: Act) -> bool:
        return id(a.t) in self.gwritten

    def gy(self, a: Act) -> torch.Tensor:
        if not self.has_grad(a):
            raise RuntimeError("backward program reads a gradient nobody wrote (internal error)")
        return self.grad(a)

    def acc(self, a: Act) -> int:
        """accumulate flag for the next writer of a's gradient (0 = first writer)."""
        flag = 1 if id(a.t) in self.gwritten else 0
        self.gwritten.add(id(a.t))
        return flag

    def add_into(self, a: Act, src: torch.Tensor) -> None:
        """grad(a) (+)= src"""
        self.add("dd_ew", 3, L.ptr(src), None, L.ptr(self.grad(a)), src.numel(), 1.0, self.acc(a))

    def pgrad(self, param: torch.nn.Parameter, shape: Tuple[int, ...], to_param: Callable[[torch.Tensor], torch.Tensor]):
        """Reserve a zero-initialised fp32 accumulation buffer (kernel layout) for `param`'s gradient."""
        n = 1
        for s in shape:
            n *= s
        off = self.pg_total
        self.pg_total += (n + 3) // 4 * 4
        self.pg_specs.append((param, off, tuple(shape), to_param))
        return _PgPtr(self, off)

    def param_grads(self) -> Dict[int, torch.Tensor]:
        out: Dict[int, torch.Tensor] = {}
        specs = self.pg_specs
        if getattr(self, "grad_plan", None) is not None:
            # one launch writes every tabulated gradient, in parameter layout, into one fresh buffer; the gradients are views
            flat = torch.empty(self.grad_plan.total, dtype=torch.float32, device=self.device)
            self.grad_plan.run((self.pg_arena.data_ptr(),), dst0=flat)
            self.last_flat = flat                   # the data-parallel reducer all-reduces this buffer in place (parallel.GradReducer)
            for param, o, n in self.grad_fast:
                out[id(param)] = flat[o:o + n].view(param.shape)
            specs = self.slow_specs
        for param, off, shape, to_param in specs:
            n = 1
            for s in shape:
                n *= s
            g = to_param(self.pg_arena[off:off + n].view(shape)).reshape(param.shape)
            if g.untyped_storage().data_ptr() == self.pg_arena.untyped_storage().data_ptr():
                g = g.clone()                       # still a view of the arena (no permutation happened): detach it
            out[id(param)] = out[id(param)] + g if id(param) in out else g
        return out

    def pre_backward(self) -> None:
        """device-side resets at the head of the backward list (subclasses add theirs)"""
        L.call("dd_zero", self.pg_arena.data_ptr(), self.pg_arena.numel() * 4, L.stream())

    def _backward_list(self) -> None:
        self.pre_backward()
        for op in self.bops:
            op()

    # Both launch lists are static (fixed buffers, weights repacked in place, the dropout seed in device memory), so after
    # one eager pass each -- lazy CUDA initialisation and cudaFuncSetAttribute calls must not happen inside a capture --
    # they are captured once and replayed: a training step issues two graph launches per network instead of ~1000 kernel
    # launches, which is what keeps the GPU fed when the host has to wait for a loss value every step.
    def _capture_graphs(self) -> None:
        self.run_ops()
        self._backward_list()          # gradient inputs hold whatever they hold: the results are overwritten by the real pass
        torch.cuda.synchronize()
        out = []
        for fn in (self.run_ops, self._backward_list):
            g = torch.cuda.CUDAGraph()
            n0 = L._Counter.n
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                fn()
            out += [g, L._Counter.n - n0]
        self.graphs = tuple(out)

    @staticmethod
    def _graphs_enabled() -> bool:
        return os.environ.get("DD_TRAIN_GRAPH", "1") != "0" and not os.environ.get("DD_DEBUG")

    def run_forward(self) -> None:
        if not self._graphs_enabled():
            return self.run_ops()
        if self.graphs is None:
            self._capture_graphs()
        self.graphs[0].replay()
        L._Counter.n += self.graphs[1]

    def run_backward(self) -> None:
        if not self._graphs_enabled() or self.graphs is None:
            return self._backward_list()
        self.graphs[2].replay()
        L._Counter.n += self.graphs[3]

    # ---- differentiable building blocks --------------------------------------------------------
    def t_conv(self, x: Optional[Act], conv: torch.nn.Module, *, x2: Act = None, kind: str = "3x3", residual: Act = None,
               pre_mish: bool = False, tanh: bool = False, out_nchw: torch.Tensor = None, in_nchw: torch.Tensor = None,
               in_shape: Tuple[int, int, int] = None, bias: bool = True, need_dx: bool = True, emit_mish: bool = False) -> Act:
        """conv / transposed conv with optional pre-Mish, residual add, tanh, NCHW input / output (fp32)."""
        if (self.tf32 and kind in ("down", "up") and x is not None and x2 is None and residual is None and not pre_mish and not tanh
                and out_nchw is None and in_nchw is None and self._strided_tc_ok(x, conv, kind)):
            return self.t_conv_strided(x, conv, kind)
        w = conv.weight
        transposed = kind == "up"
        if in_nchw is not None:
            C1, H, W = in_shape
            B = in_nchw.shape[0]
        else:
            C1, H, W, B = x.C, x.H, x.W, x.B
        C2 = x2.C if x2 is not None else 0
        Cin = C1 + C2
        Cout = w.shape[1] if transposed else w.shape[0]
        ks = w.shape[2]
        if transposed:
            stride, pad, mode, Ho, Wo = 2, 1, 1, 2 * H, 2 * W
            wd = self.packed((ks * ks, Cin, Cout), torch.float32,
                             lambda buf: buf.copy_(w.detach().permute(2, 3, 0, 1).reshape(ks * ks, Cin, Cout)))
        else:
            stride = 2 if kind == "down" else 1
            pad = 1 if ks == 3 else 0
            mode = 0
            Ho, Wo = (H + 2 * pad - ks) // stride + 1, (W + 2 * pad - ks) // stride + 1
            wd = self.packed((ks * ks, Cin, Cout), torch.float32,
                             lambda buf: buf.copy_(w.detach().permute(2, 3, 1, 0).reshape(ks * ks, Cin, Cout)))
        has_bias = bias and conv.bias is not None
        b_t = self.f32(conv.bias) if has_bias else None
        y = None if out_nchw is not None else self.act(Ho, Wo, Cout, B)
        flags = (L.CONV_PRE_MISH if pre_mish else 0) | (L.CONV_TANH if tanh else 0) | \
                (L.CONV_OUT_NCHW if out_nchw is not None else 0) | (L.CONV_IN_NCHW if in_nchw is not None else 0)
        src = L.ptr(in_nchw) if in_nchw is not None else L.ptr(x.t)
        pow2 = lambda v: v > 0 and (v & (v - 1)) == 0
        use_tc = (self.tf32 and kind in ("3x3", "1x1") and in_nchw is None and out_nchw is None and not tanh and C1 % 32 == 0
                  and C2 % 32 == 0 and Cout % 32 == 0 and pow2(H) and pow2(W) and H * W >= 16)
        kcode = L.TC_CONV3x3 if ks == 3 else L.TC_CONV1x1
        xm = None
        thin_w = lambda c: c % 4 == 0 and 4 <= c <= 128 and (c & (c - 1)) == 0
        plain1x1 = ks == 1 and not transposed and stride == 1 and x2 is None and residual is None and not pre_mish
        thin_in = plain1x1 and in_nchw is not None and out_nchw is None and not tanh and C1 <= 8 and thin_w(Cout)
        thin_out = plain1x1 and out_nchw is not None and in_nchw is None and Cout <= 8 and thin_w(C1) and Cout <= C1 // 4
        if thin_in:        # 3 -> 64 input layer: HBM streaming kernel (K = 3 leaves nothing for a tensor core)
            thin_wt = self.packed((C1, Cout), torch.float32, lambda buf: buf.copy_(w.detach().reshape(Cout, C1).t()))
            self.add("dd_conv1x1_thin_in", src, L.ptr(thin_wt), L.ptr(b_t) if b_t is not None else None, L.ptr(y.t), B, H * W, C1, Cout, 0)
        elif thin_out:     # 64 -> 3 (+tanh) output layer
            thin_wt = self.packed((Cout, C1), torch.float32, lambda buf: buf.copy_(w.detach().reshape(Cout, C1)))
            self.add("dd_conv1x1_thin_out", L.ptr(x.t), L.ptr(thin_wt), L.ptr(b_t) if b_t is not None else None, L.ptr(out_nchw), B, H * W,
                     C1, Cout, 1 if tanh else 0)
        elif use_tc:
            K = ks * ks * Cin
            wp = self.packed((Cout, K), torch.float32, lambda buf: buf.copy_(w.detach().permute(0, 2, 3, 1).reshape(Cout, K)))
            xin = x
            if pre_mish:       # the activated copy feeds the tensor-core forward (TMA -> UMMA has no place for a pre-activation)
                xm = getattr(x, "mish", None)          # written by the producer's epilogue when it was asked to (emit_mish)
                if xm is None:
                    xm = self.act(H, W, C1, B)
                    self.add("dd_ew", 0, L.ptr(x.t), None, L.ptr(xm.t), x.t.numel(), 1.0, 0)
                xin = xm
            ym = self.act(Ho, Wo, Cout, B) if emit_mish else None
            self.add("dd_conv_tc32", kcode, L.ptr(xin.t), L.ptr(x2.t) if x2 is not None else None, C1, C2, L.ptr(wp), Cout,
                     L.ptr(b_t) if b_t is not None else None, L.ptr(residual.t) if residual is not None else None, L.ptr(y.t),
                     L.ptr(ym.t) if ym is not None else None, None, B, H, W, Cout)
            if ym is not None:
                y.mish = ym
        else:
            self.add("dd_conv_direct", src, L.ptr(x2.t) if x2 is not None else None, C1, C2, L.DD_F32, L.ptr(wd),
                     L.ptr(b_t) if b_t is not None else None, L.ptr(residual.t) if residual is not None else None,
                     L.ptr(out_nchw) if out_nchw is not None else L.ptr(y.t), L.DD_F32, B, H, W, Cout, ks, stride, pad, mode, flags)

        # ---- backward -------------------------------------------------------------------------
        if transposed:
            to_w = lambda g: g.view(ks, ks, Cin, Cout).permute(2, 3, 0, 1)          # -> (Cin, Cout, k, k)
        else:
            to_w = lambda g: g.view(ks, ks, Cin, Cout).permute(3, 2, 0, 1)          # -> (Cout, Cin, k, k)
        dw = self.pgrad(w, (ks * ks, Cin, Cout), to_w)
        db = self.pgrad(conv.bias, (Cout,), lambda g: g) if has_bias else None
        # input-gradient weights per source: wT[tap'][co][ci]
        srcs = []
        if need_dx and in_nchw is None:
            srcs.append((x, 0, C1))
        if x2 is not None:
            srcs.append((x2, C1, C2))
        wts = []
        for (_, c_lo, c_n) in srcs:
            def fill(buf, c_lo=c_lo, c_n=c_n):
                wdt = w.detach()
                if transposed:                       # (Cin, Cout, k, k): dX = conv(gY, stride 2) with wT[tap][co][ci]
                    buf.copy_(wdt[c_lo:c_lo + c_n].permute(2, 3, 1, 0).reshape(ks * ks, Cout, c_n))
                elif kind == "down":                 # dX = transposed conv of gY with wT[tap][co][ci] = W[co][ci][ky][kx]
                    buf.copy_(wdt[:, c_lo:c_lo + c_n].permute(2, 3, 0, 1).reshape(ks * ks, Cout, c_n))
                else:                                # stride 1: correlation with the flipped kernel
                    buf.copy_(wdt[:, c_lo:c_lo + c_n].flip(2, 3).permute(2, 3, 0, 1).reshape(ks * ks, Cout, c_n))
            if use_tc:          # rows = input channels of this source, K = tap * Cout + co, flipped taps
                wts.append(self.packed((c_n, ks * ks * Cout), torch.float32, lambda buf, c_lo=c_lo, c_n=c_n: buf.copy_(
                    w.detach()[:, c_lo:c_lo + c_n].flip(2, 3).permute(1, 2, 3, 0).reshape(c_n, ks * ks * Cout))))
            else:
                wts.append(self.packed((ks * ks, Cout, c_n), torch.float32, fill))
        ypre = self.empty(*out_nchw.shape, dtype=torch.float32) if (tanh and out_nchw is not None) else None

        def backward():
            if out_nchw is not None:
                g = self.out_grad                     # NCHW fp32 gradient handed in by autograd
                if tanh:
                    self.add("dd_ew", 2, L.ptr(out_nchw), L.ptr(g), L.ptr(ypre), out_nchw.numel(), 1.0, 0)
                    g = ypre
                gflags = L.CONV_OUT_NCHW
            else:
                g = self.gy(y)
                gflags = 0
            M = B * Ho * Wo
            if thin_in:         # dW (1, Cin, Cout) and the bias gradient from one pass over dY
                self.add("dd_conv1x1_thin_wgrad", src, L.ptr(g), dw, 1, db, B, H * W, C1, Cout)
                return
            if thin_out:        # g is the (tanh-corrected) NCHW output gradient
                if db is not None:
                    self.add("dd_colsum", L.ptr(g), db, M, Cout, 1, Ho * Wo)
                self.add("dd_conv1x1_thin_wgrad", L.ptr(g), L.ptr(x.t), dw, 0, None, B, H * W, Cout, C1)
                if need_dx:
                    self.add("dd_conv1x1_thin_in", L.ptr(g), L.ptr(thin_wt), None, L.ptr(self.grad(x)), B, H * W, Cout, C1, self.acc(x))
                return
            if db is not None and not use_tc:         # tensor-core path: fused into the dY operand copy below
                self.add("dd_colsum", L.ptr(g), db, M, Cout, 1 if out_nchw is not None else 0, Ho * Wo)
            if use_tc:
                # tensor-core weight gradient: K-major TF32 operands = padded channel-major copies of the input and of dY
                xin = xm if xm is not None else x
                Wp = max(32, W)
                ns = 3 if ks == 3 else 1                  # column shifts are baked into copies (TMA origin alignment)
                xT, gT = self.scratch("wg_x", ns * B * C1 * (H + 2) * Wp), self.scratch("wg_g", B * Cout * H * Wp)
                self.add("dd_nhwc_to_chw_pad", L.ptr(xin.t), xT, B, C1, H, W, Wp, 1, ns, None)
                x2T = None
                if x2 is not None:
                    x2T = self.scratch("wg_x2", ns * B * C2 * (H + 2) * Wp)
                    self.add("dd_nhwc_to_chw_pad", L.ptr(x2.t), x2T, B, C2, H, W, Wp, 1, ns, None)
                self.add("dd_nhwc_to_chw_pad", L.ptr(g), gT, B, Cout, H, W, Wp, 0, 1, db)        # + bias gradient
                self.add("dd_conv_wgrad_tc32", kcode, xT, x2T, C1, C2, gT, dw, B, H, W, Wp, Cout)
            else:
                self.add("dd_conv_wgrad", src, L.ptr(x2.t) if x2 is not None else None, C1, C2, L.DD_F32, L.ptr(g), dw, B, H, W, Cout,
                         ks, stride, pad, mode, (flags & (L.CONV_PRE_MISH | L.CONV_IN_NCHW)) | gflags)
            if residual is not None:
                self.add_into(residual, g)
            for (a, c_lo, c_n), wt in zip(srcs, wts):
                if use_tc:      # epilogue fusions: * mish'(a) for a pre-activated conv, += the gradient's earlier writers
                    ga = self.grad(a)
                    self.add("dd_conv_tc32", kcode, L.ptr(g), None, Cout, 0, L.ptr(wt), c_n, None, L.ptr(ga) if self.acc(a) else None,
                             L.ptr(ga), None, L.ptr(a.t) if pre_mish else None, B, Ho, Wo, c_n)
                    continue
                direct = (not pre_mish) and not self.has_grad(a)
                dst = self.grad(a) if direct else self.empty(*a.t.shape, dtype=torch.float32)
                gin = L.CONV_IN_NCHW if out_nchw is not None else 0
                if transposed:        # forward convT(4,2,1): dX = conv(gY, k, stride 2, pad 1)
                    self.add("dd_conv_direct", L.ptr(g), None, Cout, 0, L.DD_F32, L.ptr(wt), None, None, L.ptr(dst), L.DD_F32,
                             B, Ho, Wo, c_n, ks, 2, 1, 0, gin)
                elif kind == "down":  # forward conv(3, stride 2, pad 1): dX = transposed conv of gY
                    self.add("dd_conv_direct", L.ptr(g), None, Cout, 0, L.DD_F32, L.ptr(wt), None, None, L.ptr(dst), L.DD_F32,
                             B, Ho, Wo, c_n, ks, 2, 1, 1, gin)
                else:
                    self.add("dd_conv_direct", L.ptr(g), None, Cout, 0, L.DD_F32, L.ptr(wt), None, None, L.ptr(dst), L.DD_F32,
                             B, Ho, Wo, c_n, ks, 1, pad, 0, gin)
                if direct:
                    self.acc(a)
                elif pre_mish:        # d mish(x) / dx
                    self.add("dd_ew", 1, L.ptr(a.t), L.ptr(dst), L.ptr(self.grad(a)), dst.numel(), 1.0, self.acc(a))
                else:
                    self.add_into(a, dst)
        self.on_backward(backward)
        return y

    # ---- stride-2 / transposed convolutions as dense 3x3 convolutions on the tensor cores ------------------
    @staticmethod
    def _strided_tc_ok(x: Act, conv, kind: str) -> bool:
        pow2 = lambda v: v > 0 and (v & (v - 1)) == 0
        w = conv.weight
        cin, cout = (w.shape[0], w.shape[1]) if kind == "up" else (w.shape[1], w.shape[0])
        h, wd = (x.H, x.W) if kind == "up" else (x.H // 2, x.W // 2)
        return cin % 32 == 0 and cout % 32 == 0 and pow2(x.H) and pow2(x.W) and h * wd >= 16

    def t_conv_strided(self, x: Act, conv, kind: str) -> Act:
        """Downsample (Conv2d 3x3 s2 p1, blocks.py:44) = a 3x3 s1 conv on the space-to-depth input (4*Cin channels);
        Upsample (ConvTranspose2d 4x4 s2 p1, blocks.py:35) = a 3x3 s1 conv to 4*Cout sub-pixel channels + depth-to-space.
        Weight blocks of (tap, plane) pairs that do not occur are zero; forward, input and weight gradient then run on
        dd_conv_tc32 / dd_conv_wgrad_tc32 like every other 3x3 convolution."""
        B, H, W, C = x.B, x.H, x.W, x.C
        if kind == "down":
            h, wd = H // 2, W // 2
            x4 = self.act(h, wd, 4 * C, B)
            self.add("dd_s2d_f32", L.ptr(x.t), L.ptr(x4.t), B, h, wd, C, 1)

            def back_in():          # runs after the conv's backward wrote grad(x4)
                if self.has_grad(x):
                    tmp = self.empty(*x.t.shape, dtype=torch.float32)
                    self.add("dd_s2d_f32", L.ptr(self.gy(x4)), L.ptr(tmp), B, h, wd, C, 0)
                    self.add_into(x, tmp)
                else:
                    self.add("dd_s2d_f32", L.ptr(self.gy(x4)), L.ptr(self.grad(x)), B, h, wd, C, 0)
                    self.acc(x)
            self.on_backward(back_in)
            return self.t_conv(x4, _DownAsConv3(conv), kind="3x3")
        y4 = self.t_conv(x, _UpAsConv3(conv), kind="3x3")
        cout = conv.weight.shape[1]
        y = self.act(2 * H, 2 * W, cout, B)
        self.add("dd_s2d_f32", L.ptr(y4.t), L.ptr(y.t), B, H, W, cout, 0)

        def back_out():             # runs before the conv's backward: packed gradient of the sub-pixel channels
            self.add("dd_s2d_f32", L.ptr(self.gy(y)), L.ptr(self.grad(y4)), B, H, W, cout, 1)
            self.acc(y4)
        self.on_backward(back_out)
        return y

    def t_gn_mish(self, x: Act, gn: torch.nn.GroupNorm, *, tb: torch.Tensor = None, tb_col: int = None, dtb: torch.Tensor = None,
                  residual: Act = None, dropout: float = 0.0) -> Act:
        B, HW, C, G = x.B, x.H * x.W, x.C, gn.num_groups
        st = self.empty(B, G, 2, dtype=torch.float32)
        gamma, beta = self.f32(gn.weight), self.f32(gn.bias)
        y = self.act(x.H, x.W, C, B)
        self.add("dd_gn_stats", L.ptr(x.t), L.DD_F32, B, HW, C, G, GN_EPS, L.ptr(st))
        J = tb.shape[1] if tb is not None else 0
        self.add("dd_gn_mish", L.ptr(x.t), L.ptr(y.t), L.DD_F32, B, HW, C, G, L.ptr(st), 0, GN_EPS, L.ptr(gamma), L.ptr(beta),
                 (tb.data_ptr() + 4 * tb_col) if tb is not None else None, J, None, 0,
                 L.ptr(residual.t) if residual is not None else None)
        out = y
        if dropout > 0:
            out = self.act(x.H, x.W, C, B)
            self.n_dropout += 1
            salt = self.n_dropout                      # one mask stream per dropout layer, one seed per forward call
            self.add("dd_dropout", L.ptr(y.t), L.ptr(out.t), y.t.numel(), salt, L.ptr(self.seed_dev), float(dropout))
        s1, s2, s3 = (self.empty(B, C, dtype=torch.float32) for _ in range(3))
        dgam = self.pgrad(gn.weight, (C,), lambda g: g)
        dbet = self.pgrad(gn.bias, (C,), lambda g: g)

        def backward():
            g = self.gy(out)
            if dropout > 0:
                gm = self.grad(y)
                self.add("dd_dropout", L.ptr(g), L.ptr(gm), g.numel(), salt, L.ptr(self.seed_dev), float(dropout))
                self.gwritten.add(id(y.t))
                g = gm
            if residual is not None:
                self.add_into(residual, g)
            self.add("dd_gn_mish_bwd", L.ptr(x.t), L.ptr(g), L.ptr(st), L.ptr(gamma), L.ptr(beta), B, HW, C, G, L.ptr(s1), L.ptr(s2),
                     L.ptr(s3), L.ptr(self.grad(x)), self.acc(x))
            self.add("dd_colsum", L.ptr(s1), dgam, B, C, 0, 1)
            self.add("dd_colsum", L.ptr(s2), dbet, B, C, 0, 1)
            if tb is not None:
                self.add_host(lambda: dtb[:, tb_col:tb_col + C].copy_(s3))
        self.on_backward(backward)
        return out

    def t_mish(self, x: Act) -> Act:
        y = self.act(x.H, x.W, x.C, x.B)
        n = x.t.numel()
        self.add("dd_ew", 0, L.ptr(x.t), None, L.ptr(y.t), n, 1.0, 0)

        def backward():
            self.add("dd_ew", 1, L.ptr(x.t), L.ptr(self.gy(y)), L.ptr(self.grad(x)), n, 1.0, self.acc(x))
        self.on_backward(backward)
        return y


class _ScratchPtr:
    def __init__(self, prog, key):
        self.prog, self.key = prog, key

    @property
    def _as_parameter_(self):
        return int(self.prog.scratch_buf[self.key].data_ptr())


class _PgPtr:
    """ctypes pointer into the parameter-gradient arena, resolved at launch time."""

    def __init__(self, prog, off):
        self.prog, self.off = prog, off

    @property
    def _as_parameter_(self):
        return int(self.prog.pg_arena.data_ptr() + 4 * self.off)


def _ensure_types():
    """ctypes resolves `_as_parameter_` at call time, so _PgPtr / _ScratchPtr need no special argtypes."""
    ensure_lazy()


# =====================================================================================================
class UnetTrainEngine(TrainProgram):
    """U-Net forward/backward for a fixed (B, H, W) (training: per-sample t, optional dropout)."""

    def __init__(self, unet, B: int, H: int, W: int, need_input_grad: bool, tf32: bool = False):
        super().__init__(unet, B, tf32)
        _ensure_types()
        self.H, self.W, self.need_input_grad = H, W, need_input_grad
        n_levels = len(unet.downs)
        if H % (1 << (n_levels - 1)) or W % (1 << (n_levels - 1)):
            raise ValueError(f"input {H}x{W} must be divisible by 2^{n_levels - 1}")
        dim, cin = unet.dim, unet.in_channels
        self.x_in = self.empty(B, cin, H, W, dtype=torch.float32)
        self.eps_out = self.empty(B, cin, H, W, dtype=torch.float32)
        self.out_grad = self.empty(B, cin, H, W, dtype=torch.float32)
        self.dx_in = self.empty(B, cin, H, W, dtype=torch.float32) if need_input_grad else None
        self.t_float = self.empty(B, dtype=torch.float32)
        half = dim // 2
        self.freq = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1))).to(self.device)
        self.keep.append(self.freq)
        # ---- time path: sincos -> Linear -> Mish -> Linear -> Mish -> [all block MLPs as one Linear] ----
        res_blocks = [m for m in unet.modules() if type(m).__name__ == "ResnetBlock"]
        self.tb_off, J = {}, 0
        for rb in res_blocks:
            self.tb_off[id(rb)] = J
            J += rb.mlp[1].out_features
        emb = self.act(1, 1, dim)
        self.add("dd_sincos_emb", L.ptr(self.t_float), L.ptr(self.freq), L.ptr(emb.t), B, dim)
        h1 = self.t_linear(emb, unet.time_mlp[1], need_dx=False)
        a1 = self.t_mish(h1)
        temb = self.t_linear(a1, unet.time_mlp[3])
        a2 = self.t_mish(temb)
        tb = self.act(1, 1, J)
        self.tb_rows = tb.t.view(B, J)
        self.dtb = self.grad(tb).view(B, J)
        col = 0
        outs = []
        for rb in res_blocks:          # one small GEMM per block keeps each mlp.1 parameter's gradient separate
            o = self.t_linear(a2, rb.mlp[1])
            outs.append((rb, o))
        # gather the per-block outputs into one (B, J) row table read by the GroupNorm kernels
        def gather():
            for rb, o in outs:
                c0 = self.tb_off[id(rb)]
                self.tb_rows[:, c0:c0 + o.C].copy_(o.t.view(B, o.C))
        self.add_host(gather)

        def scatter_back():
            for rb, o in outs:
                c0 = self.tb_off[id(rb)]
                self.grad(o).view(B, o.C).copy_(self.dtb[:, c0:c0 + o.C])
                self.gwritten.add(id(o.t))
        # executed in the backward pass after every GroupNorm wrote its dtb slice
        def scatter_builder():
            for rb, o in outs:
                self.gwritten.add(id(o.t))
            self.add_host(scatter_back)
        self._build_main(unet, scatter_builder)
        self.build_backward()
        self.refresh_weights()

    def t_linear(self, x: Act, lin: torch.nn.Linear, need_dx: bool = True) -> Act:
        """nn.Linear on (B, C) rows, as a 1x1 convolution over a 1x1 image."""
        shim = _LinearAsConv(lin)
        return self.t_conv(x, shim, kind="1x1", need_dx=need_dx)

    def _resnet(self, rb, x: Act, x2: Act = None, need_dx: bool = True) -> Act:
        col = self.tb_off[id(rb)]
        c1, g1 = rb.block1.block[0], rb.block1.block[1]
        c2, g2 = rb.block2.block[0], rb.block2.block[1]
        h = self.t_conv(x, c1, x2=x2, kind="3x3", need_dx=need_dx)
        if not isinstance(rb.res_conv, torch.nn.Identity):
            res = self.t_conv(x, rb.res_conv, x2=x2, kind="1x1", need_dx=need_dx)
        else:
            res = x
        p = rb.dropout.p if (self.module.training and rb.dropout.p > 0) else 0.0
        h = self.t_gn_mish(h, g1, tb=self.tb_rows, tb_col=col, dtb=self.dtb, dropout=p)
        h = self.t_conv(h, c2, kind="3x3")
        return self.t_gn_mish(h, g2, residual=res)

    def _attn(self, res_mod, x: Act) -> Act:
        pre, attn = res_mod.fn, res_mod.fn.fn
        B, n, C = x.B, x.H * x.W, x.C
        g, b = self.f32(pre.norm.g), self.f32(pre.norm.b)
        xn = self.act(x.H, x.W, C, B)
        self.add("dd_layernorm_c", L.ptr(x.t), L.ptr(xn.t), L.DD_F32, B * n, C, L.ptr(g), L.ptr(b), pre.norm.eps)
        dg = self.pgrad(pre.norm.g, (C,), lambda t: t.view(1, C, 1, 1))
        dbb = self.pgrad(pre.norm.b, (C,), lambda t: t.view(1, C, 1, 1))

        def ln_backward():
            self.add("dd_layernorm_c_bwd", L.ptr(x.t), L.ptr(self.gy(xn)), L.ptr(g), pre.norm.eps, B * n, C, L.ptr(self.grad(x)),
                     self.acc(x), dg, dbb)
        self.on_backward(ln_backward)                 # registration order = forward order (walked in reverse)
        qkv = self.t_conv(xn, attn.to_qkv, kind="1x1", bias=False)
        hid = attn.heads * attn.dim_head
        o = self.act(x.H, x.W, hid, B)
        need = int(L.lib().dd_linattn_ws_floats(B, n, attn.heads))
        ws = self.empty(need, dtype=torch.float32)
        saved = self.empty(B * attn.heads * 1088, dtype=torch.float32)
        dctx = self.empty(B * attn.heads * 1024, dtype=torch.float32)
        self.add("dd_linattn_core", L.ptr(qkv.t), L.ptr(o.t), L.DD_F32, B, n, attn.heads, attn.dim_head, L.ptr(ws), need)
        self.add("dd_linattn_save", L.ptr(ws), B, n, attn.heads, L.ptr(saved))

        def core_backward():
            self.add("dd_linattn_bwd", L.ptr(qkv.t), L.ptr(self.gy(o)), L.ptr(saved), L.ptr(dctx), L.ptr(self.grad(qkv)), B, n,
                     attn.heads, attn.dim_head)
            self.gwritten.add(id(qkv.t))
        self.on_backward(core_backward)
        y = self.t_conv(o, attn.to_out, kind="1x1", residual=x)
        return y

    def _build_main(self, unet, scatter_builder) -> None:
        B, H, W, cin = self.B, self.H, self.W, unet.in_channels
        x = self.act(H, W, cin)
        self.add("dd_nchw_to_nhwc", L.ptr(self.x_in), L.ptr(x.t), L.DD_F32, B, cin, H, W)
        x0 = x

        def input_backward():
            if self.need_input_grad:
                self.add("dd_nhwc_to_nchw", L.ptr(self.gy(x0)), L.DD_F32, L.ptr(self.dx_in), B, cin, H, W)
        # the time-path scatter must run after all GroupNorm backwards, i.e. be registered BEFORE them
        self.on_backward(scatter_builder)
        self.on_backward(input_backward)
        skips: List[Act] = []
        first = True
        for rb1, rb2, attn, down in unet.downs:
            x = self._resnet(rb1, x, need_dx=(self.need_input_grad or not first))
            first = False
            x = self._resnet(rb2, x)
            x = self._attn(attn, x)
            skips.append(x)
            if not isinstance(down, torch.nn.Identity):
                x = self.t_conv(x, down.conv, kind="down")
        x = self._resnet(unet.mid_block1, x)
        x = self._attn(unet.mid_attn, x)
        x = self._resnet(unet.mid_block2, x)
        for rb1, rb2, attn, up in unet.ups:
            x = self._resnet(rb1, x, x2=skips.pop())
            x = self._resnet(rb2, x)
            x = self._attn(attn, x)
            if not isinstance(up, torch.nn.Identity):
                x = self.t_conv(x, up.conv, kind="up")
        blk, last = unet.final_conv[0], unet.final_conv[1]
        h = self.t_conv(x, blk.block[0], kind="3x3")
        h = self.t_gn_mish(h, blk.block[1])
        self.t_conv(h, last, kind="1x1", out_nchw=self.eps_out)

    def forward(self, x: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
        self.refresh_weights()
        self.x_in.copy_(x)
        self.t_float.copy_(time.to(torch.float32))
        if self.n_dropout:
            self.seed_dev.random_()
        self.run_forward()
        return self.eps_out.clone()

    def pre_backward(self) -> None:
        super().pre_backward()
        self.dtb.zero_()

    def backward(self, grad_out: torch.Tensor):
        self.out_grad.copy_(grad_out)
        self.run_backward()
        return self.dx_in.clone() if self.need_input_grad else None


class _LinearAsConv:
    """Presents an nn.Linear as a 1x1 convolution module (weight (out, in, 1, 1)) to t_conv."""

    def __init__(self, lin: torch.nn.Linear):
        self.lin = lin
        self.bias = lin.bias

    @property
    def weight(self):
        return _W4(self.lin.weight)


class _W4:
    """(out, in) parameter viewed as (out, in, 1, 1); identity-hashable as the underlying parameter."""

    def __init__(self, p):
        self.p = p
        self.shape = (p.shape[0], p.shape[1], 1, 1)

    def detach(self):
        return self.p.detach().view(*self.shape)


class _Virt:
    """A parameter seen through a linear re-layout: `build` maps the real tensor to the virtual one (what the kernels
    consume), `to_real` maps a gradient of the virtual tensor back onto the real parameter."""

    def __init__(self, real, shape, build, to_real):
        self.real, self.shape, self._build, self.to_real = real, tuple(shape), build, to_real

    def detach(self):
        return self._build(self.real.detach())

    def numel(self):
        n = 1
        for s_ in self.shape:
            n *= s_
        return n


_KH_OF = {(0, 0): 1, (0, -1): 3, (1, 0): 2, (1, 1): 0}          # ConvTranspose2d(4,2,1): (sub-pixel phase, input offset) -> kernel index
_DOWN_OF = {0: (-1, 1), 1: (0, 0), 2: (0, 1)}                    # Conv2d(3,2,1): kernel index -> (packed offset, plane parity)


class _UpAsConv3:
    """ConvTranspose2d(Cin, Cout, 4, 2, 1) as Conv2d(Cin, 4*Cout, 3, 1, 1) on the input grid (+ depth-to-space):
    out[2a+py][2b+px][co] = sum_{dh,dw,ci} x[a+dh][b+dw][ci] * w[ci][co][py+1-2dh][px+1-2dw]."""

    def __init__(self, conv):
        w, cin, cout = conv.weight, conv.weight.shape[0], conv.weight.shape[1]

        def build(wr):
            v = wr.new_zeros(4, cout, cin, 3, 3)
            for (py, dh), kh in _KH_OF.items():
                for (px, dw), kw in _KH_OF.items():
                    v[py * 2 + px, :, :, dh + 1, dw + 1] = wr[:, :, kh, kw].t()
            return v.view(4 * cout, cin, 3, 3)

        def to_real(g):
            g = g.reshape(4, cout, cin, 3, 3)
            r = g.new_zeros(cin, cout, 4, 4)
            for (py, dh), kh in _KH_OF.items():
                for (px, dw), kw in _KH_OF.items():
                    r[:, :, kh, kw] = g[py * 2 + px, :, :, dh + 1, dw + 1].t()
            return r
        self.weight = _Virt(w, (4 * cout, cin, 3, 3), build, to_real)
        self.bias = None if conv.bias is None else _Virt(conv.bias, (4 * cout,), lambda b: b.repeat(4), lambda g: g.reshape(4, cout).sum(0))


class _DownAsConv3:
    """Conv2d(Cin, Cout, 3, 2, 1) as Conv2d(4*Cin, Cout, 3, 1, 1) on the space-to-depth input (channel = plane*Cin + ci)."""

    def __init__(self, conv):
        w, cout, cin = conv.weight, conv.weight.shape[0], conv.weight.shape[1]

        def build(wr):
            v = wr.new_zeros(cout, 4, cin, 3, 3)
            for kh, (dh, py) in _DOWN_OF.items():
                for kw, (dw, px) in _DOWN_OF.items():
                    v[:, py * 2 + px, :, dh + 1, dw + 1] = wr[:, :, kh, kw]
            return v.view(cout, 4 * cin, 3, 3)

        def to_real(g):
            g = g.reshape(cout, 4, cin, 3, 3)
            r = g.new_zeros(cout, cin, 3, 3)
            for kh, (dh, py) in _DOWN_OF.items():
                for kw, (dw, px) in _DOWN_OF.items():
                    r[:, :, kh, kw] = g[:, py * 2 + px, :, dh + 1, dw + 1]
            return r
        self.weight = _Virt(w, (cout, 4 * cin, 3, 3), build, to_real)
        self.bias = conv.bias


# make pgrad accept the wrappers: gradients are registered against the real parameter
_orig_pgrad = TrainProgram.pgrad


def _pgrad(self, param, shape, to_param):
    if isinstance(param, _W4):
        real = param.p
        return _orig_pgrad(self, real, shape, lambda g, f=to_param, r=real: f(g).reshape(r.shape))
    if isinstance(param, _Virt):
        return _orig_pgrad(self, param.real, shape, lambda g, f=to_param, v=param: v.to_real(f(g)))
    return _orig_pgrad(self, param, shape, to_param)


TrainProgram.pgrad = _pgrad


# =====================================================================================================
class ResampleTrainProgram(TrainProgram):
    """ConvResNet / SimpleDownConv / SimpleUpConv forward + backward (fp32), NCHW in / out."""

    def __init__(self, net, B: int, C: int, H: int, W: int, tanh: bool, need_input_grad: bool, tf32: bool = False):
        super().__init__(net, B, tf32)
        _ensure_types()
        from .downsampled import ConvResBlock
        self.need_input_grad = need_input_grad
        self.x_in = self.empty(B, C, H, W, dtype=torch.float32)
        layers = list(net.conv)
        x: Optional[Act] = None
        self.out = None
        self.dx_in = None
        for i, m in enumerate(layers):
            last = i == len(layers) - 1
            if isinstance(m, ConvResBlock):
                if m.drop.p > 0 and net.training:
                    raise RuntimeError("Dropout2d in the resampling nets is only supported with p=0 (reference default)")
                nxt_block = (i + 1 < len(layers) and isinstance(layers[i + 1], ConvResBlock) and not (m.upsample or m.downsample))
                h = self.t_conv(x, m.c1, kind="1x1", pre_mish=True, emit_mish=True)       # c1..c3 outputs are only read through Mish
                h = self.t_conv(h, m.c2, kind="3x3", pre_mish=True, emit_mish=True)
                h = self.t_conv(h, m.c3, kind="3x3", pre_mish=True, emit_mish=True)
                x = self.t_conv(h, m.c4, kind="1x1", pre_mish=True, residual=x if m.residual else None, emit_mish=nxt_block)
                if m.upsample:
                    x = self.t_resample(x, up=True)
                elif m.downsample:
                    x = self.t_resample(x, up=False)
                continue
            transposed = isinstance(m, torch.nn.ConvTranspose2d)
            kind = "up" if transposed else ("down" if m.stride[0] == 2 else ("3x3" if m.kernel_size[0] == 3 else "1x1"))
            first = x is None
            if last:
                Cout = m.out_channels
                Hi, Wi = (H, W) if first else (x.H, x.W)
                Ho, Wo = (2 * Hi, 2 * Wi) if transposed else ((Hi // 2, Wi // 2) if kind == "down" else (Hi, Wi))
                self.out = self.empty(B, Cout, Ho, Wo, dtype=torch.float32)
                self.out_grad = self.empty(B, Cout, Ho, Wo, dtype=torch.float32)
            if first:
                if need_input_grad:
                    # materialise the NHWC copy so the input gradient has a home
                    x = self.act(H, W, C)
                    self.add("dd_nchw_to_nhwc", L.ptr(self.x_in), L.ptr(x.t), L.DD_F32, B, C, H, W)
                    x0 = x
                    self.dx_in = self.empty(B, C, H, W, dtype=torch.float32)

                    def input_backward(x0=x0):
                        self.add("dd_nhwc_to_nchw", L.ptr(self.gy(x0)), L.DD_F32, L.ptr(self.dx_in), B, C, H, W)
                    self.on_backward(input_backward)
                    x = self.t_conv(x, m, kind=kind, out_nchw=self.out if last else None, tanh=tanh and last)
                else:
                    x = self.t_conv(None, m, kind=kind, in_nchw=self.x_in, in_shape=(C, H, W),
                                    out_nchw=self.out if last else None, tanh=tanh and last)
            else:
                x = self.t_conv(x, m, kind=kind, out_nchw=self.out if last else None, tanh=tanh and last)
        if self.out is None:
            raise ValueError("resampling net must end in a plain convolution")
        self.build_backward()
        self.refresh_weights()

    def t_resample(self, x: Act, up: bool) -> Act:
        B, H, W, C = x.B, x.H, x.W, x.C
        if up:      # nearest x2 (convblocks.py:127); backward = 2x2 block sum
            y = self.act(2 * H, 2 * W, C, B)
            self.add("dd_unpool2", L.ptr(x.t), L.ptr(y.t), B, H, W, C, 1.0)

            def backward():
                tmp_needed = self.has_grad(x)
                dst = self.empty(*x.t.shape, dtype=torch.float32) if tmp_needed else self.grad(x)
                self.add("dd_pool2_sum", L.ptr(self.gy(y)), L.ptr(dst), B, 2 * H, 2 * W, C, 1.0)
                if tmp_needed:
                    self.add_into(x, dst)
                else:
                    self.acc(x)
        else:       # avg_pool2d(2,2) (convblocks.py:129); backward = 0.25 * nearest x2
            y = self.act(H // 2, W // 2, C, B)
            self.add("dd_pool2_sum", L.ptr(x.t), L.ptr(y.t), B, H, W, C, 0.25)

            def backward():
                tmp_needed = self.has_grad(x)
                dst = self.empty(*x.t.shape, dtype=torch.float32) if tmp_needed else self.grad(x)
                self.add("dd_unpool2", L.ptr(self.gy(y)), L.ptr(dst), B, H // 2, W // 2, C, 0.25)
                if tmp_needed:
                    self.add_into(x, dst)
                else:
                    self.acc(x)
        self.on_backward(backward)
        return y

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.refresh_weights()
        self.x_in.copy_(x)
        self.run_forward()
        return self.out.clone()

    def backward(self, grad_out: torch.Tensor):
        self.out_grad.copy_(grad_out)
        self.run_backward()
        return self.dx_in.clone() if self.need_input_grad else None


# =====================================================================================================
# Called as hook(flat) right after a network's parameter gradients were written into their flat buffer, i.e. in the middle of
# loss.backward(): parallel.GradReducer uses it to start that network's all-reduce while the other networks still back-propagate.
_grads_ready_hook = None


def set_grads_ready_hook(fn) -> None:
    global _grads_ready_hook
    _grads_ready_hook = fn


class _NetFn(torch.autograd.Function):
    """One autograd node per network: forward/backward are the program's two launch lists."""

    @staticmethod
    def forward(ctx, prog, x, extra, *params):
        ctx.prog = prog
        ctx.n_params = len(params)
        ctx.params = params
        # the activations backward() needs live in the program's buffers, not on ctx: stamp the forward so that a
        # second forward through the same program before this one's backward is an error, not a silent wrong gradient
        prog.generation = getattr(prog, "generation", 0) + 1
        ctx.generation = prog.generation
        with torch.no_grad():
            return prog.forward(x, extra) if extra is not None else prog.forward(x)

    @staticmethod
    def backward(ctx, grad_out):
        prog = ctx.prog
        if ctx.generation != prog.generation:
            raise RuntimeError("backward() of a forward whose saved activations were overwritten: the same network ran forward again "
                               "(same batch shape) before this backward.  Run forward -> backward per micro-batch, or wrap the "
                               "extra forward in torch.no_grad().")
        with torch.no_grad():
            dx = prog.backward(grad_out.contiguous().float())
            prog.last_flat = None
            pg = prog.param_grads()
            if _grads_ready_hook is not None and prog.last_flat is not None:
                _grads_ready_hook(prog.last_flat)
        grads = tuple(pg.get(id(p)) if ctx.needs_input_grad[3 + i] else None for i, p in enumerate(ctx.params))
        return (None, dx if ctx.needs_input_grad[1] else None, None) + grads


def train_math(module) -> str:
    """'fp32' (CUDA-core convolutions, bit-faithful validation mode) when the module runs in precision='fp32', else
    'tf32': convolutions on the tensor cores with TF32 operands / fp32 accumulate -- the arithmetic torch's default
    `cudnn.allow_tf32 = True` gives the reference on a GPU."""
    return "fp32" if getattr(module, "precision", "bf16") == "fp32" else "tf32"


def _train_cache(module) -> EngineCache:
    if not hasattr(module, "_train_programs"):
        module._train_programs = EngineCache()
    return module._train_programs


def unet_apply(unet, eng_unused, x: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
    B, _, H, W = x.shape
    need_dx = bool(x.requires_grad)
    tf32 = train_math(unet) == "tf32"
    key = (B, H, W, need_dx, unet.training, tf32)
    cache = _train_cache(unet)
    prog = cache.get(key)
    if prog is None:
        prog = UnetTrainEngine(unet, B, H, W, need_dx, tf32)
        cache[key] = prog
    params = tuple(unet.parameters())
    return _NetFn.apply(prog, x.contiguous().float(), time, *params)


def resample_apply(net, x: torch.Tensor, tanh: bool) -> torch.Tensor:
    B, C, H, W = x.shape
    need_dx = bool(x.requires_grad)
    tf32 = train_math(net) == "tf32"
    key = (B, C, H, W, bool(tanh), need_dx, net.training, tf32)
    cache = _train_cache(net)
    prog = cache.get(key)
    if prog is None:
        prog = ResampleTrainProgram(net, B, C, H, W, bool(tanh), need_dx, tf32)
        cache[key] = prog
    params = tuple(net.parameters())
    return _NetFn.apply(prog, x.contiguous().float(), None, *params)
