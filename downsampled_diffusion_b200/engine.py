"""Step programs: the U-Net forward (unet.py:74-104 of the reference) and the down/up-sampling nets
(convblocks.py:133-159) lowered to a fixed list of libddb200 kernel launches.

A program is built once per (module, batch, resolution, precision): every activation buffer is a
torch tensor allocated up front (Python/torch owns all memory, the C side none), weights are repacked
into the kernels' layouts, and `run()` is a flat loop of C-ABI calls on the current CUDA stream with
fixed pointers -- so a whole ancestral step can be captured in one CUDA graph and replayed T times.

precision 'bf16': NHWC bf16 activations, 3x3/1x1/strided/transposed convs on tcgen05 (dd_conv_tc),
                  GroupNorm statistics accumulated by the conv epilogue.
precision 'fp32': NHWC fp32 activations, CUDA-core direct convs (dd_conv_direct) -- the validation mode.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _lib as L

GN_EPS = 1e-5


class EngineCache(dict):
    """Per-module cache of built programs; never copied along with the module (EMA deep-copies models)."""

    def __deepcopy__(self, memo):
        return EngineCache()


class Act:
    """An NHWC activation tensor of a program."""
    __slots__ = ("t", "B", "H", "W", "C", "mish", "ln", "c_real", "released", "pinned")

    def __init__(self, t: torch.Tensor, B: int, H: int, W: int, C: int):
        self.t, self.B, self.H, self.W, self.C = t, B, H, W, C
        self.released = False   # its buffer went back to the program's pool (Program.release)
        self.pinned = False     # still needed later in the program (a skip connection): release() leaves it alone
        self.c_real = C         # logical channels when the tensor is zero-padded to C (the U-Net input on the tensor-core path)
        self.mish = None        # training programs: Act holding mish(t) when the producing conv's epilogue wrote it
        self.ln = None          # (stats (B*H*W, parts, 2) fp32, parts): per-pixel channel sums the producing launch left for a PreNorm


def _is_pow2(v: int) -> bool:
    return v > 0 and (v & (v - 1)) == 0


def params_version(module: torch.nn.Module) -> int:
    """Changes whenever a parameter is written in place (optimizer step, load_state_dict: `_version`) or
    rebound to new storage (`p.data = ...`: data_ptr).  Raw-pointer writers (EMA.update) call invalidate()."""
    return hash(tuple((p._version, p.data_ptr()) for p in module.parameters()))


class Program:
    """Shared machinery: buffer allocation, weight repacking, launch list."""

    def __init__(self, module: torch.nn.Module, B: int, precision: str):
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        self.module = module
        self.B = B
        self.precision = precision
        self.device = next(module.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("programs are built for CUDA modules only (no CPU fallback)")
        self.adt = torch.bfloat16 if precision == "bf16" else torch.float32
        self.dcode = L.dtype_code(self.adt)
        self.ops: List[Callable[[], None]] = []
        self.op_names: List[str] = []
        self.conv_tc_flops: List[float] = []        # algorithmic 2*M*N*K of every dd_conv_tc launch, in order
        self.keep: List[torch.Tensor] = []          # packed weights etc. referenced by raw pointer
        self.packers: List[Callable[[], None]] = [] # re-run when the module's parameters change
        self.packed_bufs: List[torch.Tensor] = []   # packers[i] fills packed_bufs[i]
        self.weights_version = None
        self.pack_epoch = 0                         # bumped every time the packers re-run: consumers of derived data compare it
        self.n_gn = 0
        self.tc_flags = 0                           # extra dd_conv_tc flags for every conv of the program (tests: L.TC_PAIR)
        self.gn_slots: List[Tuple[int, int]] = []   # (B*G*2 offset, G) per GroupNorm of a bf16 program
        self.stats_arena: Optional[torch.Tensor] = None
        # Inference programs hand the buffers of dead activations to later layers of the same shape: the ~1.5 GB of per-op outputs
        # of a C3 step otherwise all end up written back to HBM (417 MB per step, profiles/README.md), and a 50 MB output burst
        # stalls on the write-back of somebody else's dirty lines.  A reused buffer is still L2-resident from its last life.
        self.pooling = precision == "bf16" and not os.environ.get("DD_NO_POOL")
        self._pool: Dict[tuple, List[torch.Tensor]] = {}

    # ---- helpers ---------------------------------------------------------------------------
    def empty(self, *shape, dtype=None) -> torch.Tensor:
        # launches hold raw pointers: every program buffer must stay alive as long as the program does
        t = torch.empty(*shape, dtype=dtype or self.adt, device=self.device)
        self.keep.append(t)
        return t

    def act(self, H: int, W: int, C: int, B: int = None) -> Act:
        B = B or self.B
        free = self._pool.get((B, H, W, C, self.adt)) if self.pooling else None
        return Act(free.pop() if free else self.empty(B, H, W, C), B, H, W, C)

    def release(self, a: Optional[Act]) -> None:
        """The activation is dead from here on in launch order: its buffer may be the output of any LATER launch (every launch
        writes only after the grid dependency on its predecessors has resolved, side-stream launches after their fork event)."""
        if not self.pooling or a is None or a.released or a.pinned or a.t.dtype != self.adt or a.t.dim() != 4:
            return
        a.released = True
        self._pool.setdefault((a.B, a.H, a.W, a.C, self.adt), []).append(a.t)

    def packed(self, shape, dtype, fill: Callable[[torch.Tensor], None]) -> torch.Tensor:
        """A derived weight buffer: allocated once, (re)filled from the fp32 master parameters."""
        buf = torch.zeros(*shape, dtype=dtype, device=self.device)
        self.keep.append(buf)

        def pack():
            with torch.no_grad():
                fill(buf)
        self.packers.append(pack)
        self.packed_bufs.append(buf)
        return buf

    def f32(self, param: torch.Tensor, pad_to: int = None) -> torch.Tensor:
        """Flat fp32 copy of a parameter (gamma/beta/bias), optionally zero padded."""
        n = param.numel()
        return self.packed((pad_to or n,), torch.float32, lambda b: b[:n].copy_(param.detach().reshape(-1)))

    def refresh_weights(self) -> None:
        v = params_version(self.module)
        if v != self.weights_version:
            for pk in self.packers:
                pk()
            self.weights_version = v
            self.pack_epoch += 1

    def add(self, name: str, *args) -> None:
        fn = getattr(L.lib(), name)
        self.op_names.append(name)

        def op():
            L._Counter.n += 1
            rc = fn(*args, L.stream())
            if rc != 0:
                L.check(rc, name)
        self.ops.append(op)

    def _fork(self, branch: List[Callable[[], None]], names: List[str]) -> Callable[[], None]:
        """Append ONE op that runs `branch` on the program's side stream after everything issued so far; returns the op that makes
        the main stream wait for it.  Under graph capture the pair becomes a fork / join of the captured graph."""
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        side = self._side
        ev_fork, ev_join = torch.cuda.Event(), torch.cuda.Event()

        def fork():
            main = torch.cuda.current_stream(self.device)
            ev_fork.record(main)
            side.wait_event(ev_fork)
            with torch.cuda.stream(side):
                for op in branch:
                    op()
                ev_join.record(side)

        def join():
            torch.cuda.current_stream(self.device).wait_event(ev_join)
        self.ops.append(fork)
        assert len(names) >= 1
        return join

    def run_ops(self) -> None:
        if os.environ.get("DD_DEBUG"):
            for i, op in enumerate(self.ops):        # synchronise after every launch to localise a fault
                op()
                try:
                    torch.cuda.synchronize()
                except Exception as e:               # noqa: BLE001
                    raise RuntimeError(f"op {i} ({self.op_names[i]}) faulted: {e}") from e
            return
        for op in self.ops:
            op()

    # ---- convolution lowering --------------------------------------------------------------
    FUSED = ("fused", 3)        # stats handle of a conv whose epilogue already applied GroupNorm + Mish (dd_conv_tc_gn)

    def conv(self, x: Act, conv: torch.nn.Module, *, x2: Act = None, kind: str = "3x3", gn: torch.nn.GroupNorm = None,
             residual: Act = None, pre_mish: bool = False, tanh: bool = False, out_nchw: torch.Tensor = None,
             bias: bool = True, fuse: dict = None) -> Tuple[Act, Optional[Tuple[torch.Tensor, int]]]:
        """Lower one nn.Conv2d / nn.ConvTranspose2d.  Returns (output Act, GroupNorm stats handle or None).

        fuse (with gn, bf16 path): dict(tb_col=column of the time-bias table or None, residual=Act added after the
             activation or None).  When the layer can take it the conv's epilogue applies GroupNorm + Mish itself
             (dd_conv_tc_gn) and the returned handle is Program.FUSED: the output IS the activated tensor.

        kind: '3x3' (s1 p1), '1x1', 'down' (3x3 s2 p1), 'up' (ConvTranspose2d 4,2,1).
        gn: when given the conv feeds that GroupNorm: statistics are produced (epilogue atomics in bf16,
            a dd_gn_stats launch in fp32) and returned as (stats tensor, mode)."""
        w = conv.weight
        Cin = x.C + (x2.C if x2 is not None else 0)
        Cout = w.shape[1] if kind == "up" else w.shape[0]
        B, H, W = x.B, x.H, x.W
        if kind == "down":
            Ho, Wo = H // 2, W // 2
        elif kind == "up":
            Ho, Wo = H * 2, W * 2
        else:
            Ho, Wo = H, W
        y = None if out_nchw is not None else self.act(Ho, Wo, Cout, B)
        b_t = self.f32(conv.bias) if (bias and conv.bias is not None) else None
        stats = None

        use_tc = self.precision == "bf16" and not pre_mish and not tanh
        if use_tc:
            # maps that are not powers of two (28 -> 14 -> 7) run the same tcgen05 kernel on tiles padded to the next power of two
            # (TMA zero-fills what hangs over the edge); the library keeps them on the plain epilogue: GroupNorm + Mish as dd_gn_mish
            if x.C % 64 or (x2 is not None and x2.C % 64):
                raise ValueError(f"bf16 tensor-core path needs channel counts that are multiples of 64, got {x.C}")
            Cout_p = Cout if Cout % 16 == 0 else (Cout + 15) // 16 * 16
            rows = Cout_p if Cout_p < 128 else (Cout_p + 127) // 128 * 128
            if rows != Cout_p:
                raise ValueError(f"unsupported Cout={Cout} for the tensor-core path")
            if b_t is not None and Cout_p != Cout:
                b_t = self.f32(conv.bias, pad_to=Cout_p)
            cr = x.c_real                   # < x.C: the input carries zero pad channels, the packed weights zero columns for them
            if cr != x.C and (x2 is not None or kind not in ("3x3", "1x1") or w.shape[1] != cr):
                raise ValueError("a zero-padded input feeds plain 3x3 / 1x1 convolutions only")
            if kind in ("3x3", "down"):
                K = 9 * Cin
                if cr != x.C:
                    wp = self.packed((rows, K), torch.bfloat16, lambda buf: buf.view(rows, 9, Cin)[:Cout, :, :cr].copy_(
                        w.detach().permute(0, 2, 3, 1).reshape(Cout, 9, cr)))
                else:
                    wp = self.packed((rows, K), torch.bfloat16,
                                     lambda buf: buf[:Cout].copy_(w.detach().permute(0, 2, 3, 1).reshape(Cout, K)))
                kcode = L.TC_CONV3x3 if kind == "3x3" else L.TC_DOWN
            elif kind == "1x1":
                K = Cin
                if cr != x.C:
                    wp = self.packed((rows, K), torch.bfloat16, lambda buf: buf[:Cout, :cr].copy_(w.detach().reshape(Cout, cr)))
                else:
                    wp = self.packed((rows, K), torch.bfloat16, lambda buf: buf[:Cout].copy_(w.detach().reshape(Cout, K)))
                kcode = L.TC_CONV1x1
            else:   # 'up': ConvTranspose2d weight (Cin, Cout, 4, 4) -> 4 sub-pixel phase matrices
                K = 4 * Cin

                def fill(buf, w=w, Cin=Cin, Cout=Cout, rows=rows):
                    wd = w.detach()
                    for ph in range(4):
                        py, px = ph >> 1, ph & 1
                        for t in range(4):
                            a, b = t >> 1, t & 1
                            ky, kx = 2 * a + 1 - py, 2 * b + 1 - px
                            buf[ph * rows: ph * rows + Cout, t * Cin:(t + 1) * Cin].copy_(wd[:, :, ky, kx].t())
                wp = self.packed((4 * rows, K), torch.bfloat16, fill)
                kcode = L.TC_UPT
            src, src2 = x.t, (x2.t if x2 is not None else None)
            gh, gw = (Ho, Wo) if kind == "down" else (H, W)
            G = 0
            st_ptr = None
            flags = self.tc_flags
            if kind == "down" and os.environ.get("DD_NO_STRIDED_TMA"):
                planes = self.empty(4, B, Ho, Wo, x.C)           # the four parity planes of the input (space-to-depth copy)
                self.add("dd_space_to_depth2", L.ptr(x.t), L.ptr(planes), B, H, W, x.C)
                src = planes
            elif kind == "down":
                flags |= L.TC_STRIDED_IN                          # the kernel reads every other pixel through a stride-2 tensor map
            if gn is not None:
                G = gn.num_groups
                # low-resolution layers: split-K partials, summed by the GroupNorm kernel (dd_gn_mish_sum)
                S = int(L.lib().dd_conv_tc_splits(kcode, B, gh, gw, Cin, Cout_p)) if (residual is None and Cout_p == Cout) else 1
                if S > 1 and gh * gw * Cout <= 16384 and S * B * gh * gw * Cout <= self.SPLITK_WS_FLOATS:
                    flags |= L.TC_SPLITK
                    stats = (("split", S, b_t), 2)
                elif fuse is not None and residual is None and Cout_p == Cout and kind != "up" and self._gn_fusable(kcode, B, gh, gw, Cout, G):
                    stats = self.FUSED
                else:
                    stats = (self._new_stats_slot(B, G), 1)
                    st_ptr = stats[0]
            taps = {"3x3": 9, "down": 9, "1x1": 1, "up": 4}[kind]
            self.conv_tc_flops.append(2.0 * B * Ho * Wo * Cout * taps * (Cin if x.c_real == x.C else x.c_real))     # algorithmic: no pad channels
            if stats is self.FUSED:
                self._add_conv_gn(kcode, src, 0, src2, x.C, x2.C if x2 is not None else 0, wp, b_t, y, B, gh, gw, Cout, flags, gn, fuse)
                return y, stats
            self.add("dd_conv_tc", kcode, L.ptr(src), 0, L.ptr(src2) if src2 is not None else None, x.C,
                     x2.C if x2 is not None else 0, L.ptr(wp), wp.shape[0], L.ptr(b_t) if b_t is not None else None,
                     L.ptr(residual.t) if residual is not None else None,
                     L.ptr(out_nchw) if out_nchw is not None else L.ptr(y.t), 1 if out_nchw is not None else 0,
                     Cout if out_nchw is not None else 0, st_ptr, G, B, gh, gw, Cout_p, flags, *self.splitk_args())
            if y is not None and Cout_p != Cout:
                raise ValueError("padded Cout is only supported with NCHW fp32 output")
        else:
            if kind == "up":
                wd = self.packed((16, Cin, Cout), torch.float32,
                                 lambda buf: buf.copy_(w.detach().permute(2, 3, 0, 1).reshape(16, Cin, Cout)))
                ks, stride, pad, mode = 4, 2, 1, 1
            else:
                ks = w.shape[2]
                wd = self.packed((ks * ks, Cin, Cout), torch.float32,
                                 lambda buf: buf.copy_(w.detach().permute(2, 3, 1, 0).reshape(ks * ks, Cin, Cout)))
                stride, pad, mode = (2 if kind == "down" else 1), (1 if ks == 3 else 0), 0
            flags = (L.CONV_PRE_MISH if pre_mish else 0) | (L.CONV_TANH if tanh else 0) | \
                    (L.CONV_OUT_NCHW if out_nchw is not None else 0)
            self.add("dd_conv_direct", L.ptr(x.t), L.ptr(x2.t) if x2 is not None else None, x.C,
                     x2.C if x2 is not None else 0, L.dtype_code(x.t.dtype), L.ptr(wd),
                     L.ptr(b_t) if b_t is not None else None, L.ptr(residual.t) if residual is not None else None,
                     L.ptr(out_nchw) if out_nchw is not None else L.ptr(y.t), self.dcode, B, H, W, Cout,
                     ks, stride, pad, mode, flags)
            if gn is not None:
                G = gn.num_groups
                st = self.empty(B, G, 2, dtype=torch.float32)
                self.add("dd_gn_stats", L.ptr(y.t), self.dcode, B, Ho * Wo, Cout, G, GN_EPS, L.ptr(st))
                stats = (st, 0)
        return y, stats

    GN_FUSE_MAX_CLUSTER = int(os.environ.get("DD_GN_FUSE_MAX_CLUSTER", "2"))

    def _gn_fusable(self, kcode, B, gh, gw, Cout, G) -> bool:
        """The conv can apply GroupNorm + Mish itself: as the persistent kernel (statistics through a zeroed workspace), or with the
        image's tiles as one thread-block cluster -- by default only up to 2 CTAs: clusters of 4 / 8 halve the number of resident
        CTAs (cudaOccupancyMaxActiveClusters: 33 / 15 clusters) and lose to conv + dd_gn_mish (profiles/README.md, round 2)."""
        lib = L.lib()
        if int(lib.dd_conv_tc_gn_ws_floats(kcode, B, gh, gw, Cout, G)) > 0:
            return True
        return 0 < int(lib.dd_conv_tc_gn_cluster(kcode, B, gh, gw, Cout, G)) <= self.GN_FUSE_MAX_CLUSTER

    def _add_conv_gn(self, kcode, src, pitch, src2, C1, C2, wp, b_t, y: Act, B, gh, gw, Cout, flags, gn, fuse: dict) -> None:
        gamma, beta = self.f32(gn.weight), self.f32(gn.bias)
        tb_col, res = fuse.get("tb_col"), fuse.get("residual")
        G = gn.num_groups
        ws_floats = int(L.lib().dd_conv_tc_gn_ws_floats(kcode, B, gh, gw, Cout, G))
        ws = None
        if ws_floats > 0:            # persistent kernel: flagged {sum, 1, sumsq, 1} packets per (image, channel tile, pixel tile, group), zeroed with the arena every run
            ws = _ArenaPtr(self, sum(s_ for s_, _ in self.gn_slots))
            self.gn_slots.append((ws_floats, G))
        ln_buf = None
        if fuse.get("want_ln"):
            parts = int(L.lib().dd_conv_tc_gn_ln_parts(kcode, B, gh, gw, Cout, G, 1 if ws is not None else 0))
            ln_buf = self.empty(B * gh * gw, parts, 2, dtype=torch.float32)
            y.ln = (ln_buf, parts)
        self.op_names.append("dd_conv_tc")          # counted with the convolutions (bench.py's roofline replays them)
        fn = L.lib().dd_conv_tc_gn
        args = (kcode, L.ptr(src), pitch, L.ptr(src2) if src2 is not None else None, C1, C2, L.ptr(wp), wp.shape[0],
                L.ptr(b_t) if b_t is not None else None, L.ptr(y.t), B, gh, gw, Cout, flags, gn.num_groups, GN_EPS,
                L.ptr(gamma), L.ptr(beta), _TbPtr(self, tb_col) if tb_col is not None else None,
                self.tb[0] if tb_col is not None else 0, _TrowPtr(self) if tb_col is not None else None,
                _TrowStride(self) if tb_col is not None else 0, L.ptr(res.t) if res is not None else None,
                L.ptr(ln_buf) if ln_buf is not None else None, ws)

        def op():
            L._Counter.n += 1
            rc = fn(*args, L.stream())
            if rc != 0:
                L.check(rc, "dd_conv_tc_gn")
        self.ops.append(op)

    SPLITK_WS_FLOATS = 148 * 128 * 128        # one full wave of 128x128 fp32 tiles (9.7 MB)
    SPLITK_COUNTERS = 1024

    def splitk_args(self):
        """(ws, ws_floats, counters, n_counters) shared by every split-K conv of the program: zero when idle."""
        if getattr(self, "_splitk", None) is None:
            ws = torch.zeros(self.SPLITK_WS_FLOATS, dtype=torch.float32, device=self.device)
            cnt = torch.zeros(self.SPLITK_COUNTERS, dtype=torch.int32, device=self.device)
            self.keep += [ws, cnt]
            self._splitk = (L.ptr(ws), self.SPLITK_WS_FLOATS, L.ptr(cnt), self.SPLITK_COUNTERS)
        return self._splitk

    def _new_stats_slot(self, B: int, G: int) -> int:
        """Reserve a (B,G,2) fp32 slot of the statistics arena; returns its device pointer lazily."""
        off = sum(s for s, _ in self.gn_slots)
        self.gn_slots.append((B * G * 2, G))
        return _ArenaPtr(self, off)

    def finalize_arena(self) -> None:
        total = sum(s for s, _ in self.gn_slots)
        if total:
            self.stats_arena = torch.zeros(total, dtype=torch.float32, device=self.device)

    def gn_mish(self, x: Act, stats, gn: torch.nn.GroupNorm, *, tb_col: int = None, tb=None, residual: Act = None,
                want_ln: bool = False) -> Act:
        if stats is self.FUSED:
            return x                                # the conv's epilogue already did it (and added tb / residual)
        y = self.act(x.H, x.W, x.C, x.B)
        st, mode = stats
        gamma, beta = self.f32(gn.weight), self.f32(gn.bias)
        if mode == 2:       # x was never written: the conv left S fp32 partials in the shared split-K workspace
            _, S, b_t = st
            ln_buf = None
            if want_ln and not os.environ.get("DD_NO_LN_FOLD"):
                parts = int(L.lib().dd_gn_mish_sum_parts(x.C, gn.num_groups))
                cvp = x.C // 4 // parts
                if 0 < cvp <= 32 and cvp & (cvp - 1) == 0:
                    ln_buf = self.empty(x.B * x.H * x.W, parts, 2, dtype=torch.float32)
                    y.ln = (ln_buf, parts)
            self.add("dd_gn_mish_sum", self.splitk_args()[0], S, L.ptr(b_t) if b_t is not None else None, L.ptr(y.t), x.B, x.H * x.W,
                     x.C, gn.num_groups, GN_EPS, L.ptr(gamma), L.ptr(beta),
                     _TbPtr(self, tb_col) if tb_col is not None else None, tb[0] if tb else 0,
                     _TrowPtr(self) if tb_col is not None else None, _TrowStride(self) if tb_col is not None else 0,
                     L.ptr(residual.t) if residual is not None else None, L.ptr(ln_buf) if ln_buf is not None else None)
            return y
        self.add("dd_gn_mish", L.ptr(x.t), L.ptr(y.t), self.dcode, x.B, x.H * x.W, x.C, gn.num_groups,
                 st if isinstance(st, _ArenaPtr) else L.ptr(st), mode, GN_EPS, L.ptr(gamma), L.ptr(beta),
                 _TbPtr(self, tb_col) if tb_col is not None else None, tb[0] if tb else 0,
                 _TrowPtr(self) if tb_col is not None else None, _TrowStride(self) if tb_col is not None else 0,
                 L.ptr(residual.t) if residual is not None else None)
        return y


class _Lazy:
    """ctypes argument resolved at launch time (pointers that change between runs)."""

    def __init__(self, prog):
        self.prog = prog


class _ArenaPtr(_Lazy):
    def __init__(self, prog, off):
        super().__init__(prog)
        self.off = off

    @property
    def _as_parameter_(self):
        return L.C.c_void_p(self.prog.stats_arena.data_ptr() + 4 * self.off)


class _TbPtr(_Lazy):
    def __init__(self, prog, col):
        super().__init__(prog)
        self.col = col

    @property
    def _as_parameter_(self):
        return L.C.c_void_p(self.prog.tb_rows.data_ptr() + 4 * self.col)


class _TrowPtr(_Lazy):
    @property
    def _as_parameter_(self):
        t = self.prog.trow
        return L.C.c_void_p(t.data_ptr() if t is not None else None)


class _TrowStride(_Lazy):
    @property
    def _as_parameter_(self):
        return L.C.c_int(self.prog.trow_stride)


# ctypes calls with argtypes set convert through from_param; teach c_void_p / c_int our lazies
def _install_lazy_support():
    lib = L.lib()
    for name, args in L.SIGNATURES.items():
        getattr(lib, name).argtypes = [_LazyVoidP if a is L._p else (_LazyInt if a is L._i else a) for a in args]


class _LazyVoidP(L.C.c_void_p):
    @classmethod
    def from_param(cls, v):
        if isinstance(v, _Lazy):
            return v._as_parameter_
        return L.C.c_void_p.from_param(v)


class _LazyInt(L.C.c_int):
    @classmethod
    def from_param(cls, v):
        if isinstance(v, _Lazy):
            return v._as_parameter_
        return L.C.c_int(v)


_lazy_installed = False


def ensure_lazy():
    global _lazy_installed
    if not _lazy_installed:
        _install_lazy_support()
        _lazy_installed = True


class UnetEngine(Program):
    """The U-Net forward as a launch list (see module docstring).  Inputs/outputs are NCHW fp32."""

    def __init__(self, unet, B: int, H: int, W: int, precision: str):
        super().__init__(unet, B, precision)
        ensure_lazy()
        self.H, self.W = H, W
        n_levels = len(unet.downs)
        if H % (1 << (n_levels - 1)) or W % (1 << (n_levels - 1)):
            raise ValueError(f"input {H}x{W} must be divisible by 2^{n_levels - 1} (unet_dims has {n_levels} levels)")
        dim, cin = unet.dim, unet.in_channels
        if dim % 8:
            raise ValueError("unet_chan must be a multiple of 8 (GroupNorm groups)")
        self.x_in = self.empty(B, cin, H, W, dtype=torch.float32)      # NCHW staging (graph-stable pointer)
        self.eps_out = self.empty(B, cin, H, W, dtype=torch.float32)
        # ---- time-embedding bias columns: one block of dim_out columns per ResnetBlock ----
        self.res_blocks = [m for m in unet.modules() if type(m).__name__ == "ResnetBlock"]
        self.tb_off = {}
        J = 0
        for rb in self.res_blocks:
            self.tb_off[id(rb)] = J
            J += rb.mlp[1].out_features
        self.J = J
        Wcat = self.packed((J, dim), torch.float32,
                           lambda b: b.copy_(torch.cat([rb.mlp[1].weight.detach() for rb in self.res_blocks], 0)))
        bcat = self.packed((J,), torch.float32,
                           lambda b: b.copy_(torch.cat([rb.mlp[1].bias.detach() for rb in self.res_blocks], 0)))
        self.tm = (self.f32(unet.time_mlp[1].weight), self.f32(unet.time_mlp[1].bias),
                   self.f32(unet.time_mlp[3].weight), self.f32(unet.time_mlp[3].bias), Wcat, bcat)
        half = dim // 2
        k = math.log(10000) / (half - 1)
        self.freq = torch.exp(torch.arange(half) * -k).to(self.device)          # blocks.py:25-26, evaluated on host
        self.tb_batch = self.empty(B, J, dtype=torch.float32)   # rows for a per-sample `time` argument
        self.t_float = self.empty(B, dtype=torch.float32)
        self.tb_rows = self.tb_batch                            # table the gn_mish launches read (rebindable)
        self.trow: Optional[torch.Tensor] = None                # int32 row index per sample (None: row = b)
        self.trow_stride = 0
        self.tb = (J,)
        self._tables: Dict[int, list] = {}                      # T -> [(T, J) table, pack_epoch it was filled at]
        self._build(unet)
        self.finalize_arena()
        self.refresh_weights()

    # ---- program construction --------------------------------------------------------------
    def _resnet(self, rb, x: Act, x2: Act = None, first: bool = False, want_ln: bool = False) -> Act:
        col = self.tb_off[id(rb)]
        c1, g1 = rb.block1.block[0], rb.block1.block[1]
        c2, g2 = rb.block2.block[0], rb.block2.block[1]
        has_res = not isinstance(rb.res_conv, torch.nn.Identity)
        join = None
        if first and self.precision == "bf16" and x.c_real == x.C:
            # x is the im2col'd input (B,H,W,kpad): the 3x3 conv and the 1x1 res_conv are K=kpad GEMMs
            h, st = self._conv_im2col(x, c1, gn=g1, center_only=False, fuse=dict(tb_col=col))
            res = self._conv_im2col(x, rb.res_conv, gn=None, center_only=True)[0] if has_res else None
            if not has_res:
                raise ValueError("first ResnetBlock without res_conv is not supported on the tensor-core path")
        else:
            if x.c_real != x.C and not has_res:
                raise ValueError("first ResnetBlock without res_conv is not supported on the tensor-core path")
            # all blocks or none (measured, pass ar): at 64 samples x 32x32 the second branch only gets in the persistent kernels' way
            # (+0.15 %; forking the low-resolution blocks alone +0.8 %), from 32 samples down it is worth 2 % of the step
            fork_rows = int(os.environ.get("DD_FORK_MAX_ROWS", "32768"))
            if has_res and self.precision == "bf16" and self.B * self.H * self.W <= fork_rows and not os.environ.get("DD_NO_FORK"):
                # The 1x1 res_conv reads only the block's input and is needed only by the second convolution's epilogue: it is issued
                # FIRST, on a side stream (a parallel branch of the captured graph), and runs next to the first convolution instead
                # of between the two.  (Without the res_conv launches at all a step is 29 us shorter: profiles/README.md, pass ar.)
                n0 = len(self.ops)
                res, _ = self.conv(x, rb.res_conv, x2=x2, kind="1x1")
                branch = self.ops[n0:]
                del self.ops[n0:]
                join = self._fork(branch, self.op_names[n0:])
                del self.op_names[n0 + 1:]
                h, st = self.conv(x, c1, x2=x2, kind="3x3", gn=g1, fuse=dict(tb_col=col))
            else:
                h, st = self.conv(x, c1, x2=x2, kind="3x3", gn=g1, fuse=dict(tb_col=col))
                if has_res:
                    res, _ = self.conv(x, rb.res_conv, x2=x2, kind="1x1")
                else:
                    assert x2 is None
                    res = x
        h1 = self.gn_mish(h, st, g1, tb_col=col, tb=self.tb)
        if h1 is not h:
            self.release(h)
        if join is not None:
            self.ops.append(join)
            self.op_names.append("join")
        want_ln = want_ln and self.precision == "bf16" and not os.environ.get("DD_NO_LN_FOLD")
        h2, st = self.conv(h1, c2, kind="3x3", gn=g2, fuse=dict(residual=res, want_ln=want_ln))
        out = self.gn_mish(h2, st, g2, residual=res, want_ln=want_ln)
        if out is not h2:
            self.release(h2)
        self.release(h1)
        if has_res:
            self.release(res)
        return out

    def _conv_im2col(self, x: Act, conv, gn, center_only: bool, fuse: dict = None):
        w = conv.weight
        Cout, Cin = w.shape[0], w.shape[1]
        kpad = x.C

        def fill(buf):
            wd = w.detach()
            if center_only:
                buf[:, 4 * Cin:5 * Cin].copy_(wd.reshape(Cout, Cin))
            else:
                buf[:, :9 * Cin].copy_(wd.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin))
        wp = self.packed((Cout, kpad), torch.bfloat16, fill)
        b_t = self.f32(conv.bias)
        y = self.act(x.H, x.W, Cout, x.B)
        stats, st_ptr, G = None, None, 0
        self.conv_tc_flops.append(2.0 * x.B * x.H * x.W * Cout * Cin * (1 if center_only else 9))
        if gn is not None:
            G = gn.num_groups
            if fuse is not None and self._gn_fusable(L.TC_CONV1x1, x.B, x.H, x.W, Cout, G):
                self._add_conv_gn(L.TC_CONV1x1, x.t, 0, None, kpad, 0, wp, b_t, y, x.B, x.H, x.W, Cout, 0, gn, fuse)
                return y, self.FUSED
            stats = (self._new_stats_slot(x.B, G), 1)
            st_ptr = stats[0]
        self.add("dd_conv_tc", L.TC_CONV1x1, L.ptr(x.t), 0, None, kpad, 0, L.ptr(wp), Cout, L.ptr(b_t), None, L.ptr(y.t),
                 0, 0, st_ptr, G, x.B, x.H, x.W, Cout, 0, *self.splitk_args())
        return y, stats

    def _attn(self, res_mod, x: Act) -> Act:
        pre = res_mod.fn            # PreNorm
        attn = pre.fn               # LinearAttention
        hid = attn.heads * attn.dim_head
        xn = None
        if x.ln is not None and self.precision == "bf16" and x.C % 64 == 0:
            qkv = self._qkv_ln_folded(pre, attn, x)
        else:
            g, b = self.f32(pre.norm.g), self.f32(pre.norm.b)
            xn = self.act(x.H, x.W, x.C, x.B)
            self.add("dd_layernorm_c", L.ptr(x.t), L.ptr(xn.t), self.dcode, x.B * x.H * x.W, x.C, L.ptr(g), L.ptr(b),
                     pre.norm.eps)
            qkv, _ = self.conv(xn, attn.to_qkv, kind="1x1", bias=False)
        if self.precision == "bf16":
            # fused output: per-sample matrices M_b = ctx_b . W_out^T, then ONE tensor-core GEMM q . M_b + bias + x
            C = x.C
            if C % 64:
                raise ValueError("bf16 tensor-core attention needs channel counts that are multiples of 64")
            need = int(L.lib().dd_linattn_mix_ws_floats(x.B, x.H * x.W, attn.heads))    # partials + arrival tickets
            ws = torch.zeros(need, dtype=torch.float32, device=self.device)
            self.keep.append(ws)
            wout = self.packed((C, hid), torch.bfloat16, lambda b: b.copy_(attn.to_out.weight.detach().reshape(C, hid)))
            mb = self.empty(x.B, C, hid, dtype=torch.bfloat16)
            self.add("dd_linattn_mix", L.ptr(qkv.t), self.dcode, x.B, x.H * x.W, attn.heads, attn.dim_head, L.ptr(ws), need,
                     L.ptr(wout), C, L.ptr(mb))
            b_t = self.f32(attn.to_out.bias)
            y = self.act(x.H, x.W, C, x.B)
            self.conv_tc_flops.append(2.0 * x.B * x.H * x.W * C * hid)
            self.add("dd_conv_tc", L.TC_CONV1x1, L.ptr(qkv.t), qkv.C, None, hid, 0, L.ptr(mb), C, L.ptr(b_t), L.ptr(x.t),
                     L.ptr(y.t), 0, 0, None, 0, x.B, x.H, x.W, C, L.TC_W_PER_SAMPLE, *self.splitk_args())
            self.release(qkv)
            self.release(xn)
            return y
        o = self.act(x.H, x.W, hid, x.B)
        need = int(L.lib().dd_linattn_ws_floats(x.B, x.H * x.W, attn.heads))
        ws = self.empty(need, dtype=torch.float32)
        self.add("dd_linattn_core", L.ptr(qkv.t), L.ptr(o.t), self.dcode, x.B, x.H * x.W, attn.heads, attn.dim_head,
                 L.ptr(ws), need)
        y, _ = self.conv(o, attn.to_out, kind="1x1", residual=x)
        return y

    def _qkv_ln_folded(self, pre, attn, x: Act) -> Act:
        """to_qkv(LayerNorm(x)) as ONE GEMM on x itself (blocks.py:57-69, 123): the gain is folded into the weights, mean and
        1 / (std + eps) of every pixel are applied by the epilogue from the channel sums the producing launch left in x.ln."""
        w, gvec, bvec = attn.to_qkv.weight, pre.norm.g, pre.norm.b
        Cout, C = w.shape[0], w.shape[1]
        wp = self.packed((Cout, C), torch.bfloat16, lambda buf: buf.copy_(w.detach().reshape(Cout, C) * gvec.detach().reshape(1, C)))
        # row sums of the ROUNDED weights, so that a constant input cancels exactly as it does in (x - mean)
        wsum = self.packed((Cout,), torch.float32,
                           lambda buf: buf.copy_((w.detach().reshape(Cout, C) * gvec.detach().reshape(1, C)).to(torch.bfloat16).float().sum(1)))
        cb = self.packed((Cout,), torch.float32, lambda buf: buf.copy_(w.detach().reshape(Cout, C) @ bvec.detach().reshape(C)))
        stats, parts = x.ln
        y = self.act(x.H, x.W, Cout, x.B)
        self.conv_tc_flops.append(2.0 * x.B * x.H * x.W * Cout * C)
        self.add("dd_conv_tc_ln", L.ptr(x.t), C, L.ptr(wp), Cout, L.ptr(cb), L.ptr(wsum), L.ptr(stats), parts, float(pre.norm.eps),
                 L.ptr(y.t), x.B, x.H, x.W, Cout)
        self.op_names[-1] = "dd_conv_tc"            # counted with the convolutions
        return y

    def _build(self, unet) -> None:
        B, H, W, cin = self.B, self.H, self.W, unet.in_channels
        if self.precision == "bf16" and cin <= 64 and not os.environ.get("DD_FIRST_IM2COL"):
            # the input as an ordinary activation, zero-padded to 64 channels: the first ResnetBlock is then lowered like every other
            # (3x3 convolution with the fused GroupNorm + Mish epilogue, 1x1 res_conv); K = 9 * 64 instead of the im2col GEMM's 128,
            # but no separate dd_gn_mish pass and 8 MB instead of 17 MB of input copy (profiles/README.md, round 2 pass ap)
            x = self.act(H, W, 64)
            x.c_real = cin
            self.add("dd_nchw_to_nhwc_pad", L.ptr(self.x_in), L.ptr(x.t), B, cin, H, W, 64)
        elif self.precision == "bf16":
            kpad = (9 * cin + 63) // 64 * 64
            x = self.act(H, W, kpad)
            self.add("dd_im2col3x3_nchw", L.ptr(self.x_in), L.ptr(x.t), B, cin, H, W, kpad)
        else:
            x = self.act(H, W, cin)
            self.add("dd_nchw_to_nhwc", L.ptr(self.x_in), L.ptr(x.t), self.dcode, B, cin, H, W)
        skips: List[Act] = []

        def step(fn, x, **kw):
            """x -> fn(x): the input is dead afterwards (unless it is pinned as a skip connection)"""
            y = fn(x, **kw)
            self.release(x)
            return y
        for i, (rb1, rb2, attn, down) in enumerate(unet.downs):
            x = step(lambda a, **kw: self._resnet(rb1, a, **kw), x, first=(i == 0))
            x = step(lambda a, **kw: self._resnet(rb2, a, **kw), x, want_ln=True)
            x = step(lambda a: self._attn(attn, a), x)
            x.pinned = True                                 # read again by the matching block of the up path
            skips.append(x)
            if not isinstance(down, torch.nn.Identity):
                x, _ = self.conv(x, down.conv, kind="down")
        x = step(lambda a, **kw: self._resnet(unet.mid_block1, a, **kw), x, want_ln=True)
        x = step(lambda a: self._attn(unet.mid_attn, a), x)
        x = step(lambda a: self._resnet(unet.mid_block2, a), x)
        for rb1, rb2, attn, up in unet.ups:
            skip = skips.pop()
            x = step(lambda a: self._resnet(rb1, a, x2=skip), x)        # concat-free: two K ranges (unet.py:97)
            skip.pinned = False
            self.release(skip)
            x = step(lambda a, **kw: self._resnet(rb2, a, **kw), x, want_ln=True)
            x = step(lambda a: self._attn(attn, a), x)
            if not isinstance(up, torch.nn.Identity):
                x = step(lambda a: self.conv(a, up.conv, kind="up")[0], x)
        blk, last = unet.final_conv[0], unet.final_conv[1]
        h, st = self.conv(x, blk.block[0], kind="3x3", gn=blk.block[1], fuse=dict())
        h2 = self.gn_mish(h, st, blk.block[1])
        if h2 is not h:
            self.release(h)
        self.release(x)
        self.conv(h2, last, kind="1x1", out_nchw=self.eps_out)
        self.release(h2)

    # ---- execution -------------------------------------------------------------------------
    def run(self) -> None:
        """x_in -> eps_out with the currently bound time-bias rows.  Pure launch list (graph-capturable)."""
        if self.stats_arena is not None:
            L.call("dd_zero", self.stats_arena.data_ptr(), self.stats_arena.numel() * 4, L.stream())
        self.run_ops()

    def time_bias_rows(self, t_float: torch.Tensor, out: torch.Tensor) -> None:
        w1, b1, w2, b2, wc, bc = self.tm
        L.call("dd_time_bias", L.ptr(t_float), t_float.numel(), self.module.dim, L.ptr(self.freq), L.ptr(w1), L.ptr(b1),
               L.ptr(w2), L.ptr(b2), L.ptr(wc), L.ptr(bc), self.J, L.ptr(out), L.stream())

    def time_table(self, T: int) -> torch.Tensor:
        """(T, J) fp32: every ResnetBlock's time bias for every step -- they depend on t only, so the
        sampling chain looks them up instead of running the 18 Linear layers per step.

        ONE buffer per T for the life of the engine, refilled in place whenever the packed weights were refreshed
        since it was last filled: captured graphs hold its address, so it must neither move nor go stale."""
        self.refresh_weights()
        ent = self._tables.get(T)
        if ent is None:
            ent = [self.empty(T, self.J, dtype=torch.float32), -1]
            self._tables[T] = ent
        if ent[1] != self.pack_epoch:
            self.time_bias_rows(torch.arange(T, dtype=torch.float32, device=self.device), ent[0])
            ent[1] = self.pack_epoch
        return ent[0]

    def bind_table(self, table: torch.Tensor, trow: torch.Tensor, stride: int) -> None:
        self.tb_rows, self.trow, self.trow_stride = table, trow, stride

    def bind_batch_rows(self) -> None:
        self.tb_rows, self.trow, self.trow_stride = self.tb_batch, None, 0

    def forward(self, x: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
        """Unet.forward(x, time) semantics: per-sample `time`, returns a fresh (B,C,H,W) fp32 tensor."""
        self.refresh_weights()
        self.x_in.copy_(x)
        self.t_float.copy_(time.to(torch.float32))
        self.bind_batch_rows()
        self.time_bias_rows(self.t_float, self.tb_batch)
        self.run()
        return self.eps_out.clone()
