"""Static execution plan: symbolic NDHWC tensors + a list of libpvb200 launches.

A ``TRef`` is a channels-last-3d activation living in a (possibly shared) buffer:
element (n,t,h,w,c) is at ``buf + ((n*T+t)*H+h)*W+w) * row_stride + ch_off + c``.  Because ops
resolve pointers only when they run, a tensor can be *retargeted* into a channel slice of a wider
buffer after it was produced - that is how ``torch.cat([slow, fuse], dim=1)``
(reference models/slowfast.py:728) disappears: both producers write straight into the concat
buffer.
"""
import ctypes as C
import os

import torch

from .. import _lib as L
from . import packing as PK

_TORCH_DT = {L.PV_F16: torch.float16, L.PV_F32: torch.float32}
_ESIZE = {L.PV_F16: 2, L.PV_F32: 4}


class Buf:
    def __init__(self, numel, dt):
        self.numel = int(numel)
        self.dt = dt
        self.tensor = None


class RawIn:
    """Plan input passed through untouched (see Plan.raw_input)."""

    def __init__(self, tensor):
        self.tensor = tensor


class TRef:
    def __init__(self, buf, N, T, H, W, C, Cp=None, ch_off=0, row_stride=None):
        self.buf = buf
        self.N, self.T, self.H, self.W = int(N), int(T), int(H), int(W)
        self.C = int(C)
        self.Cp = PK.pad8(C) if Cp is None else int(Cp)
        self.ch_off = int(ch_off)
        self.row_stride = self.Cp if row_stride is None else int(row_stride)
        self.lazy_src = None      # network input whose NCDHW->NDHWC conversion is emitted by its first consumer
        self.padw = None          # (w_pad, w_phys): rows physically zero-padded along W (window-mode stems)

    @property
    def npos(self):
        return self.T * self.H * self.W

    @property
    def dt(self):
        return self.buf.dt

    def ptr(self):
        return self.buf.tensor.data_ptr() + self.ch_off * _ESIZE[self.buf.dt]

    def retarget(self, buf, ch_off, row_stride):
        self.buf, self.ch_off, self.row_stride = buf, int(ch_off), int(row_stride)

    def shape5(self):
        return (self.N, self.C, self.T, self.H, self.W)

    def __repr__(self):
        return "TRef(N=%d,C=%d(%d),T=%d,H=%d,W=%d,off=%d,rs=%d)" % (
            self.N, self.C, self.Cp, self.T, self.H, self.W, self.ch_off, self.row_stride)


def _conv_out(i, k, s, p, d):
    return (i + 2 * p - d * (k - 1) - 1) // s + 1


class Plan:
    """Collects launches; ``finalize`` allocates, ``run`` replays (eagerly or as a CUDA graph)."""

    def __init__(self, device, dt=L.PV_F16, use_tcgen05=True):
        self.lib = L.load()
        self.device = torch.device(device)
        self.dt = dt
        self.use_tcgen05 = bool(use_tcgen05) and dt == L.PV_F16
        self.ops = []          # (name, closure(stream_ptr))
        self.meta = []         # per-op {name, kind, flops, bytes} (algorithmic figures for the roofline)
        self.bufs = []
        self.consts = []       # keep device parameter tensors alive
        self.zero_bufs = []    # f32 accumulators that must be cleared every run (SE sums)
        self.finalized = False
        self.graph = None
        self.stats = {"tcgen05": 0, "direct": 0, "depthwise": 0, "other": 0}
        # ---- lanes: independent branches of the network (SlowFast pathways) run on separate CUDA streams /
        #      graph branches.  Every op records the tensors it reads / writes; finalize() turns cross-lane
        #      read-after-write pairs into event waits (see _schedule).
        self.lane = 0          # lane of the ops being emitted (set by the lowering)
        self.op_lane = []
        self.op_io = []        # (reads, writes) as lists of TRef / Buf, or None = unknown (acts as a full barrier)
        self.sched = None
        self.multi_lane = os.environ.get("PVB200_LANES", "1") != "0"
        # f16 engine, MViT: the residual token stream (16 blocks x 2 adds) is kept in fp32 - branch outputs stay f16, the
        # add + LayerNorm is one kernel (pv_add_layernorm).  PVB200_TRUNK32=0 restores the all-f16 stream (A/B only).
        self.trunk32 = dt == L.PV_F16 and os.environ.get("PVB200_TRUNK32", "1") != "0"
        self._streams = {}
        self._events = {}

    # ---- memory --------------------------------------------------------------------------
    def new_buf(self, numel, dt=None):
        b = Buf(numel, self.dt if dt is None else dt)
        self.bufs.append(b)
        return b

    def new_tensor(self, N, T, H, W, C, Cp=None, dt=None):
        Cp = PK.pad8(C) if Cp is None else Cp
        b = self.new_buf(N * T * H * W * Cp, dt)
        return TRef(b, N, T, H, W, C, Cp)

    def const(self, t, dtype=None):
        t = t.detach().to(device=self.device, dtype=dtype if dtype is not None else t.dtype).contiguous()
        self.consts.append(t)
        return t

    def finalize(self):
        for b in self.bufs:
            if b.tensor is None:
                # zero-init so that pad lanes / never-written slices are finite
                b.tensor = torch.zeros(max(b.numel, 8), dtype=_TORCH_DT[b.dt], device=self.device)
        self._schedule()
        self.finalized = True

    def _schedule(self):
        """Cross-lane dependencies.  For op i on lane L: for every other lane M, the LAST op j < i on M that wrote
        a buffer op i reads (stream order covers the earlier ones).  Ops with unknown I/O are full barriers.  No
        buffer is reused inside a plan, so read-after-write is the only hazard (two lanes writing one concat
        buffer write disjoint channel slices)."""
        def bufs_of(items):
            out = []
            for t in items or ():
                b = t if isinstance(t, Buf) else getattr(t, "buf", None)
                if b is not None:
                    out.append(b)
            return out
        n = len(self.ops)
        lanes = sorted(set(self.op_lane)) if self.op_lane else [0]
        waits = [[] for _ in range(n)]      # op -> list of op indices (on other lanes) to wait for
        writers = {}                        # id(buf) -> {lane: last writer op}
        last_on_lane = {}
        last_barrier = None                 # last op with unknown I/O
        waited = {}                         # lane -> {other lane: latest op already waited for}
        for i in range(n):
            L = self.op_lane[i]
            io = self.op_io[i]
            need = {}
            if io is None:
                for M, j in last_on_lane.items():
                    if M != L:
                        need[M] = j
            else:
                for b in bufs_of(io[0]):
                    for M, j in writers.get(id(b), {}).items():
                        if M != L:
                            need[M] = max(need.get(M, -1), j)
                if last_barrier is not None and self.op_lane[last_barrier] != L:
                    M = self.op_lane[last_barrier]
                    need[M] = max(need.get(M, -1), last_barrier)
            # a lane never needs to wait twice for the same (or an earlier) op of another lane
            seen = waited.setdefault(L, {})
            keep = []
            for M, j in sorted(need.items()):
                if seen.get(M, -1) < j:
                    seen[M] = j
                    keep.append(j)
            waits[i] = sorted(keep)
            if io is None:
                last_barrier = i
            else:
                for b in bufs_of(io[1]):
                    writers.setdefault(id(b), {})[L] = i
            last_on_lane[L] = i
        signals = sorted(set(j for w in waits for j in w))
        self.sched = {"lanes": lanes, "waits": waits, "signals": set(signals), "last_on_lane": last_on_lane}

    def bytes_allocated(self):
        return sum(b.numel * _ESIZE[b.dt] for b in self.bufs)

    # ---- execution -----------------------------------------------------------------------
    def add(self, name, fn, kind="other", flops=0.0, nbytes=0.0, reads=None, writes=None):
        """reads / writes: the TRefs (or Bufs) the launch touches; leave both None for "unknown" (the op then
        orders against everything on the other lanes)."""
        self.ops.append((name, fn))
        self.meta.append({"name": name, "kind": kind, "flops": float(flops), "bytes": float(nbytes), "lane": self.lane})
        self.stats[kind] = self.stats.get(kind, 0) + 1
        self.op_lane.append(self.lane)
        self.op_io.append(None if reads is None and writes is None else (list(reads or ()), list(writes or ())))

    def profile(self, iters=3):
        """Per-launch device times (ms, mean over `iters`) measured with CUDA events on the
        launching stream (torch's current stream); eager replay, one event pair per launch."""
        assert self.finalized
        stream = torch.cuda.current_stream(self.device)
        sp = stream.cuda_stream
        n = len(self.ops)
        acc = [0.0] * n
        for it in range(iters + 1):
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
            for (e0, e1), (_, fn) in zip(evs, self.ops):      # single stream: lanes are ignored here
                e0.record(stream)
                fn(sp)
                e1.record(stream)
            stream.synchronize()
            if it == 0:
                continue   # warm-up pass
            for i, (e0, e1) in enumerate(evs):
                acc[i] += e0.elapsed_time(e1)
        return [a / iters for a in acc]

    def run(self, stream_ptr, single_stream=False):
        """Enqueue every launch.  With more than one lane the extra lanes run on side streams that fork from /
        join back into ``stream_ptr`` with events, so a CUDA-graph capture of this call records a graph with
        parallel branches (and an eager call overlaps them the same way)."""
        assert self.finalized
        lanes = self.sched["lanes"]
        if single_stream or not self.multi_lane or len(lanes) <= 1:
            for _, fn in self.ops:
                fn(stream_ptr)
            return
        main = torch.cuda.ExternalStream(stream_ptr, device=self.device)
        streams = {lanes[0]: main}
        for L in lanes[1:]:
            if L not in self._streams:
                self._streams[L] = torch.cuda.Stream(device=self.device)
            streams[L] = self._streams[L]
        ev = self._events

        def event(key):
            e = ev.get(key)
            if e is None:
                e = ev[key] = torch.cuda.Event()
            return e
        fork = event("fork")
        fork.record(main)
        for L in lanes[1:]:
            streams[L].wait_event(fork)
        waits, signals = self.sched["waits"], self.sched["signals"]
        for i, (_, fn) in enumerate(self.ops):
            st = streams[self.op_lane[i]]
            for j in waits[i]:
                st.wait_event(event(j))
            fn(st.cuda_stream)
            if i in signals:
                event(i).record(st)
        for L in lanes[1:]:
            e = event(("join", L))
            e.record(streams[L])
            main.wait_event(e)

    def num_launches(self):
        return len(self.ops)

    # =====================================================================================
    # op emitters
    # =====================================================================================
    def emit_input_ncdhw(self, static_in, C, c_pad):
        """static_in: torch tensor [N,C,T,H,W] (f32|f16) whose storage is fixed for the plan.
        The layout conversion is emitted lazily by the first consumer (a stem conv may ask for
        physically W-padded rows, see pv_igemm.cu window mode)."""
        N, Cc, T, H, W = static_in.shape
        assert Cc == C
        out = TRef(None, N, T, H, W, C, Cp=c_pad)
        out.lazy_src = static_in
        return out

    def materialize_input(self, x, w_pad=0, w_phys=0):
        if x.lazy_src is None:
            return x
        src = x.lazy_src
        x.lazy_src = None
        N, C, T, H, W = src.shape
        src_dt = L.PV_F32 if src.dtype == torch.float32 else L.PV_F16
        lib = self.lib
        if w_pad > 0:
            x.buf = self.new_buf(N * T * H * w_phys * x.Cp + 64 * x.Cp)
            x.padw = (w_pad, w_phys)

            def fn(stream):
                L.check(lib.pv_ncdhw_to_ndhwc_padw(src.data_ptr(), src_dt, x.ptr(), x.dt, N, C, T, H, W, x.Cp,
                                                  w_pad, w_phys, stream), "pv_ncdhw_to_ndhwc_padw")
            self.add("ncdhw_to_ndhwc_padw", fn, "other", 0.0, src.numel() * src.element_size() + N * T * H * w_phys * x.Cp * 2,
                     reads=(), writes=(x,))
        else:
            x.buf = self.new_buf(N * T * H * W * x.Cp)

            def fn(stream):
                L.check(lib.pv_ncdhw_to_ndhwc(src.data_ptr(), src_dt, x.ptr(), x.dt, N, C, T, H, W, x.Cp,
                                              x.row_stride, stream), "pv_ncdhw_to_ndhwc")
            self.add("ncdhw_to_ndhwc", fn, "other", 0.0, src.numel() * src.element_size() + N * T * H * W * x.Cp * 2,
                     reads=(), writes=(x,))
        return x

    def emit_conv(self, x, weight, conv_bias, bn, stride, padding, dilation, groups, act=L.ACT_NONE,
                  residual=None, name="conv", force_algo=None, se_sums=False):
        """Conv3d (+folded BN/bias) (+residual) (+activation).  weight: [Co, Ci/g, kt, kh, kw].
        se_sums (depthwise only): also accumulate the per-(sample, channel) sums of the output inside
        the conv kernel (Squeeze-Excitation statistics); the buffer is attached as ``y.se_sums``."""
        co, cig, kt, kh, kw = weight.shape
        ci = cig * groups
        if ci != x.C:
            # same error type the reference raises for a wrong channel count
            raise RuntimeError("conv %s expects %d input channels, got %d" % (name, ci, x.C))
        st, sh, sw = stride
        pt, ph, pw = padding
        dlt, dlh, dlw = dilation
        To, Ho, Wo = _conv_out(x.T, kt, st, pt, dlt), _conv_out(x.H, kh, sh, ph, dlh), _conv_out(x.W, kw, sw, pw, dlw)
        if min(To, Ho, Wo) <= 0:
            raise RuntimeError("conv %s: kernel larger than (padded) input" % name)
        co_pad = PK.pad8(co)
        if residual is not None:
            self.materialize_input(residual)
        # ---- narrow stems with a temporal extent: factor (kt,kh,kw) -> (1,kh,kw) with kt*Co channels
        #      (all temporal taps in one tensor-core pass) + a temporal tap sum, see pv_temporal_tap_sum
        if (x.lazy_src is not None and self.use_tcgen05 and force_algo in (None, L.ALGO_TCGEN05) and groups == 1
                and x.Cp == 4 and kt > 1 and residual is None and co_pad * kt <= 256 and dlw == 1
                and (sw * x.Cp * 2) % 16 == 0 and kw * x.Cp <= 64 and pw > 0 and sh <= 8):
            w2 = torch.zeros(kt * co_pad, cig, 1, kh, kw, dtype=weight.dtype)
            wsrc = weight.detach().cpu()
            for j in range(kt):
                w2[j * co_pad: j * co_pad + co] = wsrc[:, :, j:j + 1]
            yk = self.emit_conv(x, w2, None, None, (1, sh, sw), (0, ph, pw), (1, dlh, dlw), 1, L.ACT_NONE, None,
                                name + ".taps")
            y = self.new_tensor(x.N, To, Ho, Wo, co, Cp=co_pad)
            scale, bias = PK.fold_bn(conv_bias, bn, co, co_pad)
            scale_d, bias_d = self.const(scale), self.const(bias)
            lib = self.lib
            hw = Ho * Wo

            def fn_sum(stream):
                L.check(lib.pv_temporal_tap_sum(yk.ptr(), y.ptr(), self.dt, x.N, x.T, To, hw, co_pad, kt, st, pt, dlt,
                                                scale_d.data_ptr(), bias_d.data_ptr(), act, yk.row_stride,
                                                y.row_stride, stream), "pv_temporal_tap_sum(%s)" % name)
            self.add(name + ".tapsum", fn_sum, "other", 0.0, (x.N * x.T * hw * kt * co_pad + x.N * To * hw * co_pad) * 2,
                     reads=(yk,), writes=(y,))
            return y
        # ---- network input: pick the layout its first consumer wants
        window = False
        if x.lazy_src is not None:
            window = (self.use_tcgen05 and force_algo in (None, L.ALGO_TCGEN05) and groups == 1 and x.Cp == 4
                      and dlw == 1 and (sw * x.Cp * 2) % 16 == 0 and (kw + 1) * x.Cp <= 64 and kt * kh <= 64
                      and st * sh <= 8 and pw > 0)
            if window:
                wp = (pw + 3) // 4 * 4          # left pad rounded up: enables the 4-pixel conversion kernel
                lead = PK.window_lead(wp, pw, x.Cp)
                win = PK.window_elems(kw, x.Cp, lead)
                need = wp - pw + max(x.W + 2 * pw, (Wo - 1) * sw + (win + x.Cp - 1) // x.Cp)
                self.materialize_input(x, w_pad=wp, w_phys=(need + 3) // 4 * 4)
            else:
                self.materialize_input(x)
        elif x.padw is not None:
            raise RuntimeError("W-padded stem input can only feed one window-mode convolution")
        y = self.new_tensor(x.N, To, Ho, Wo, co, Cp=co_pad)
        scale, bias = PK.fold_bn(conv_bias, bn, co, co_pad)
        scale_d, bias_d = self.const(scale), self.const(bias)
        depthwise = groups != 1
        if depthwise and not (groups == ci == co):
            raise RuntimeError("conv %s: only groups==1 or depthwise (groups==Cin==Cout) is supported" % name)
        tdt = _TORCH_DT[self.dt]
        ci_pad = x.Cp
        ci_pad64 = ci_pad if ci_pad < 64 else PK.pad_to(ci_pad, 64)   # < 64: gather-fed kernel, un-padded taps

        d = L.Conv3dDesc()
        d.dtype = self.dt
        d.N, d.Ti, d.Hi, d.Wi, d.Ci = x.N, x.T, x.H, x.W, ci_pad
        d.To, d.Ho, d.Wo, d.Co = To, Ho, Wo, co_pad
        d.kt, d.kh, d.kw = kt, kh, kw
        d.st, d.sh, d.sw = st, sh, sw
        d.pt, d.ph, d.pw = pt, ph, pw
        d.dt, d.dh, d.dw = dlt, dlh, dlw
        d.groups = ci_pad if depthwise else 1
        d.act = act
        d.has_residual = 1 if residual is not None else 0
        d.x_row_stride, d.y_row_stride = x.row_stride, y.row_stride
        d.res_row_stride = residual.row_stride if residual is not None else 0
        d.ci_pad64 = ci_pad64
        if window:
            d.x_w_pad, d.x_w_phys = x.padw
            w_lead = PK.window_lead(x.padw[0], pw, x.Cp)
            d.ci_pad64 = PK.window_elems(kw, x.Cp, w_lead)

        stem_rows, zero_row = False, None
        if depthwise:
            algo, kind = L.ALGO_DIRECT, "depthwise"
            w_d = self.const(PK.pack_depthwise(weight, co_pad, tdt))
        else:
            want_tc = self.use_tcgen05 and bool(self.lib.pv_conv3d_tcgen05_supported(C.byref(d)))
            if force_algo is not None:
                want_tc = force_algo == L.ALGO_TCGEN05
            if window and not want_tc:
                raise RuntimeError("internal: window-mode stem rejected by the library: " + L.last_error())
            stem_rows = (window and want_tc and residual is None and not os.environ.get("PVB200_NO_STEMROWS")
                         and bool(self.lib.pv_conv3d_stem_rows_supported(C.byref(d))))
            if stem_rows:
                # zero-copy im2col over raw input rows (csrc/pv_stem.cu): its own weight layout and entry point
                algo, kind = L.ALGO_TCGEN05, "tcgen05"
                w_d = self.const(PK.pack_stem_rows(weight, x.Cp, co_pad, w_lead))
                zero_row = self.const(torch.zeros(4096, dtype=torch.float16))
            elif want_tc:
                algo, kind = L.ALGO_TCGEN05, "tcgen05"
                w_d = self.const(PK.pack_dense_window(weight, x.Cp, co_pad, w_lead) if window
                                 else PK.pack_dense_tcgen05(weight, ci_pad64, co_pad))
            else:
                algo, kind = L.ALGO_DIRECT, "direct"
                w_d = self.const(PK.pack_dense_direct(weight, ci_pad, co_pad, tdt))
        lib = self.lib
        sums = None
        if se_sums and depthwise and residual is None and act == L.ACT_NONE:
            sums = self.new_buf(x.N * co_pad, L.PV_F32)
            self.zero_bufs.append(sums)
            y.se_sums = sums

            def fn_zero(stream):
                L.check(lib.pv_zero_f32(sums.tensor.data_ptr(), x.N * co_pad, stream), "pv_zero_f32")
            self.add(name + ".se_zero", fn_zero, reads=(), writes=(sums,))

        def fn_dw(stream):
            d.x_row_stride, d.y_row_stride = x.row_stride, y.row_stride
            L.check(lib.pv_dwconv3d_fwd(C.byref(d), x.ptr(), w_d.data_ptr(), scale_d.data_ptr(), bias_d.data_ptr(),
                                        y.ptr(), sums.tensor.data_ptr() if sums is not None else None, stream),
                    "pv_dwconv3d_fwd(%s)" % name)

        def fn_stem(stream):
            d.x_row_stride, d.y_row_stride = x.row_stride, y.row_stride
            L.check(lib.pv_conv3d_stem_rows_fwd(C.byref(d), x.ptr(), w_d.data_ptr(), scale_d.data_ptr(), bias_d.data_ptr(),
                                                zero_row.data_ptr(), y.ptr(), stream), "pv_conv3d_stem_rows_fwd(%s)" % name)

        def fn(stream):
            d.x_row_stride, d.y_row_stride = x.row_stride, y.row_stride   # may have been retargeted
            d.res_row_stride = residual.row_stride if residual is not None else 0
            L.check(lib.pv_conv3d_fwd(C.byref(d), algo, x.ptr(), w_d.data_ptr(), scale_d.data_ptr(),
                                      bias_d.data_ptr(), residual.ptr() if residual is not None else None,
                                      y.ptr(), stream), "pv_conv3d_fwd(%s)" % name)
        esz = _ESIZE[self.dt]
        m_out = x.N * To * Ho * Wo
        flops = 2.0 * m_out * co * cig * kt * kh * kw
        nbytes = (x.N * x.npos * ci + m_out * co * (2 if residual is not None else 1)) * esz + weight.numel() * esz
        if stem_rows:
            self.stats["stem_rows"] = self.stats.get("stem_rows", 0) + 1
        self.add(name, fn_stem if stem_rows else (fn_dw if (depthwise and residual is None) else fn), kind, flops, nbytes,
                 reads=(x,) + ((residual,) if residual is not None else ()), writes=(y,) + ((sums,) if sums is not None else ()))
        return y

    def fused_bottleneck_desc(self, x, cin_pad, cmid, cout, kt, sb, has_shortcut, act):
        d = L.BottleneckDesc()
        d.N, d.T, d.H, d.W = x.N, x.T, x.H, x.W
        d.Cin, d.Cmid, d.Cout = cin_pad, PK.pad8(cmid), PK.pad8(cout)
        d.kt, d.sb, d.has_shortcut, d.act = kt, sb, 1 if has_shortcut else 0, act
        d.x_row_stride, d.y_row_stride = x.row_stride, PK.pad8(cout)
        return d

    def emit_bottleneck_fused(self, x, conv_a, bn_a, conv_b, bn_b, conv_c, bn_c, sc_conv, sc_bn, act, name):
        """ONE launch for conv_a -> conv_b -> conv_c (+ projection / identity shortcut) + activation
        (csrc/pv_fastblock.cu); the caller has checked eligibility with pv_bottleneck_fused_supported."""
        self.materialize_input(x)
        kt, sb = int(conv_a.kernel_size[0]), int(conv_b.stride[1])
        cmid, cout = conv_a.out_channels, conv_c.out_channels
        d = self.fused_bottleneck_desc(x, x.Cp, cmid, cout, kt, sb, sc_conv is not None, act)
        Ho, Wo = (x.H - 1) // sb + 1, (x.W - 1) // sb + 1
        y = self.new_tensor(x.N, x.T, Ho, Wo, cout, Cp=d.Cout)
        wa = self.const(PK.pack_rows_k16(conv_a.weight, x.Cp, d.Cmid))
        wb = self.const(PK.pack_rows_k16(conv_b.weight, d.Cmid, d.Cmid))
        wc = self.const(PK.pack_rows_k16(conv_c.weight, d.Cmid, d.Cout))
        ws = self.const(PK.pack_rows_k16(sc_conv.weight, x.Cp, d.Cout)) if sc_conv is not None else None
        sa, ba = (self.const(t) for t in PK.fold_bn(conv_a.bias, bn_a, cmid, d.Cmid))
        sbb, bbb = (self.const(t) for t in PK.fold_bn(conv_b.bias, bn_b, cmid, d.Cmid))
        scc, bcc = (self.const(t) for t in PK.fold_bn(conv_c.bias, bn_c, cout, d.Cout))
        ss, bs = (self.const(t) for t in PK.fold_bn(sc_conv.bias, sc_bn, cout, d.Cout)) if sc_conv is not None else (None, None)
        lib = self.lib

        def fn(stream):
            d.x_row_stride, d.y_row_stride = x.row_stride, y.row_stride
            L.check(lib.pv_bottleneck_fused_fwd(C.byref(d), x.ptr(), wa.data_ptr(), wb.data_ptr(), wc.data_ptr(),
                                                ws.data_ptr() if ws is not None else None, sa.data_ptr(), ba.data_ptr(),
                                                sbb.data_ptr(), bbb.data_ptr(), scc.data_ptr(), bcc.data_ptr(),
                                                ss.data_ptr() if ss is not None else None,
                                                bs.data_ptr() if bs is not None else None, y.ptr(), stream),
                    "pv_bottleneck_fused_fwd(%s)" % name)
        m_in, m_out = x.N * x.T * x.H * x.W, x.N * x.T * Ho * Wo
        cin = conv_a.in_channels
        flops = 2.0 * (m_in * cmid * cin * kt + m_out * cmid * cmid * 9 + m_out * cout * cmid +
                       (m_out * cout * cin if sc_conv is not None else 0))
        nbytes = (m_in * cin + m_out * cout) * 2 + sum(t.numel() for t in (wa, wb, wc)) * 2
        self.add(name, fn, "fused_block", flops, nbytes, reads=(x,), writes=(y,))
        return y

    def emit_pool(self, x, mode, kernel, stride, padding, name="pool"):
        self.materialize_input(x)
        kt, kh, kw = kernel
        st, sh, sw = stride
        pt, ph, pw = padding
        To, Ho, Wo = (x.T + 2 * pt - kt) // st + 1, (x.H + 2 * ph - kh) // sh + 1, (x.W + 2 * pw - kw) // sw + 1
        if min(To, Ho, Wo) <= 0:
            raise RuntimeError("pool %s: kernel %s larger than input (%d,%d,%d)" % (name, kernel, x.T, x.H, x.W))
        y = self.new_tensor(x.N, To, Ho, Wo, x.C, Cp=x.Cp)
        d = L.Pool3dDesc()
        d.dtype, d.mode = self.dt, mode
        d.N, d.Ti, d.Hi, d.Wi, d.C = x.N, x.T, x.H, x.W, x.Cp
        d.To, d.Ho, d.Wo = To, Ho, Wo
        d.kt, d.kh, d.kw, d.st, d.sh, d.sw, d.pt, d.ph, d.pw = kt, kh, kw, st, sh, sw, pt, ph, pw
        lib = self.lib

        def fn(stream):
            d.x_row_stride, d.y_row_stride = x.row_stride, y.row_stride
            L.check(lib.pv_pool3d_fwd(C.byref(d), x.ptr(), y.ptr(), stream), "pv_pool3d_fwd(%s)" % name)
        self.add(name, fn, reads=(x,), writes=(y,))
        return y

    def emit_se_scale_act(self, x, w1, b1, w2, b2, act, name="se"):
        """In-place y = act(x * sigmoid(W2 relu(W1 mean(x) + b1) + b2)) (fvcore SqueezeExcitation)."""
        Cr, Cc = w1.shape[0], w1.shape[1]
        assert Cc == x.C
        Cp = x.Cp
        w1p = torch.zeros(Cr, Cp, dtype=torch.float32)
        w1p[:, :Cc] = w1.detach().float().cpu().reshape(Cr, Cc)
        w2p = torch.zeros(Cp, Cr, dtype=torch.float32)
        w2p[:Cc] = w2.detach().float().cpu().reshape(Cc, Cr)
        b2p = torch.zeros(Cp, dtype=torch.float32)
        b2p[:Cc] = b2.detach().float().cpu()
        w1d, b1d, w2d, b2d = self.const(w1p), self.const(b1.detach().float().cpu()), self.const(w2p), self.const(b2p)
        fused = getattr(x, "se_sums", None)       # already produced by the depthwise conv kernel
        sums = fused if fused is not None else self.new_buf(x.N * Cp, L.PV_F32)
        gate = self.new_buf(x.N * Cp, L.PV_F32)
        if fused is None:
            self.zero_bufs.append(sums)
        lib = self.lib
        npos = x.npos

        def fn_sum(stream):
            L.check(lib.pv_zero_f32(sums.tensor.data_ptr(), x.N * Cp, stream), "pv_zero_f32")
            L.check(lib.pv_channel_sum(x.ptr(), x.dt, x.row_stride, x.N, npos, Cp, sums.tensor.data_ptr(), stream),
                    "pv_channel_sum(%s)" % name)

        def fn_gate(stream):
            L.check(lib.pv_se_gate(sums.tensor.data_ptr(), npos, x.N, Cp, Cr, w1d.data_ptr(), b1d.data_ptr(),
                                   w2d.data_ptr(), b2d.data_ptr(), Cp, gate.tensor.data_ptr(), stream),
                    "pv_se_gate(%s)" % name)

        def fn_apply(stream):
            L.check(lib.pv_scale_act(x.ptr(), x.ptr(), x.dt, x.row_stride, x.row_stride, x.N, npos, Cp,
                                     gate.tensor.data_ptr(), act, stream), "pv_scale_act(%s)" % name)
        if fused is None:
            self.add(name + ".sum", fn_sum)
        self.add(name + ".gate", fn_gate)
        self.add(name + ".apply", fn_apply)
        return x

    def raw_input(self, static_in):
        """A plan input that is used as it is (fp32 device tensor, no layout / dtype conversion): the [K, 5] bounding
        boxes of the detection heads.  The handle carries the static tensor; ops read its data pointer at run time."""
        return RawIn(static_in)

    def emit_roi_align(self, x, rois, output_size, spatial_scale, sampling_ratio, name="roi_align"):
        """torchvision RoIAlign (aligned=False) on a T == 1 feature map (models/head.py:462-471): [N,1,H,W,C] x
        [K,5] -> [K,1,R_h,R_w,C]."""
        if x.T != 1:
            raise RuntimeError("Temporal dimension should be 1. Consider modifying the pool layer.")   # head.py:464-467
        self.materialize_input(x)
        K = int(rois.tensor.shape[0])
        rh, rw = int(output_size[0]), int(output_size[1])
        y = self.new_tensor(K, 1, rh, rw, x.C, Cp=x.Cp)
        lib = self.lib
        N, H, W, Cp = x.N, x.H, x.W, x.Cp

        def fn(stream):
            L.check(lib.pv_roi_align_fwd(x.ptr(), x.dt, x.row_stride, N, H, W, Cp, rois.tensor.data_ptr(), K, rh, rw,
                                         float(spatial_scale), int(sampling_ratio), y.ptr(), y.row_stride, stream),
                    "pv_roi_align_fwd(%s)" % name)
        self.add(name, fn, reads=(x,), writes=(y,))
        return y

    def emit_act(self, x, act, name="act"):
        lib = self.lib

        def fn(stream):
            L.check(lib.pv_scale_act(x.ptr(), x.ptr(), x.dt, x.row_stride, x.row_stride, x.N, x.npos, x.Cp,
                                     None, act, stream), "pv_scale_act(%s)" % name)
        self.add(name, fn)
        return x

    def emit_head_reduce(self, x, softmax, name="head_reduce"):
        out = self.new_buf(x.N * x.C, L.PV_F32)
        lib = self.lib

        def fn(stream):
            L.check(lib.pv_head_reduce(x.ptr(), x.dt, x.row_stride, x.N, x.npos, x.C, 1 if softmax else 0,
                                       out.tensor.data_ptr(), stream), "pv_head_reduce(%s)" % name)
        self.add(name, fn, reads=(x,), writes=(out,))
        return out, (x.N, x.C)

    def emit_to_ncdhw(self, x, name="to_ncdhw"):
        self.materialize_input(x)
        out = self.new_buf(x.N * x.C * x.npos, L.PV_F32)
        lib = self.lib

        def fn(stream):
            L.check(lib.pv_ndhwc_to_ncdhw(x.ptr(), x.dt, x.row_stride, out.tensor.data_ptr(), x.N, x.C, x.T,
                                          x.H, x.W, stream), "pv_ndhwc_to_ncdhw(%s)" % name)
        self.add(name, fn, reads=(x,), writes=(out,))
        return out, x.shape5()

    def emit_input_tokens(self, static_in):
        """static_in: torch tensor [B, N, C] (or [B, C]) f32|f16, contiguous: token-major network input
        (the layout MViT blocks exchange).  One cast/copy launch into the plan's dtype."""
        shp = tuple(static_in.shape)
        B, N, Cc = (shp[0], 1, shp[1]) if len(shp) == 2 else shp
        if Cc % 8 and self.dt == L.PV_F16:
            raise RuntimeError("token width %d is not a multiple of 8 (16-byte rows are required in f16 mode)" % Cc)
        x = self.new_tensor(B, 1, 1, N, Cc, Cp=Cc)
        src_dt = L.PV_F32 if static_in.dtype == torch.float32 else L.PV_F16
        total = B * N * Cc
        lib = self.lib

        def fn(stream):
            # a [1, 1, 1, 1, total] "clip" with one channel: the layout conversion degenerates to a cast
            L.check(lib.pv_ncdhw_to_ndhwc(static_in.data_ptr(), src_dt, x.ptr(), x.dt, 1, 1, 1, 1, total, 1, 1, stream),
                    "pv_ncdhw_to_ndhwc(tokens)")
        self.add("tokens_in", fn, "other", 0.0, total * (static_in.element_size() + _ESIZE[self.dt]), reads=(), writes=(x,))
        return x

    def emit_to_tokens(self, x, name="to_tokens", squeeze=False):
        """Token TRef [B, N, C] -> f32 output buffer laid out (B, N, C) (or (B, C) with squeeze)."""
        self.materialize_input(x)
        lib = self.lib
        if x.row_stride != x.C or x.Cp != x.C:
            dense = self.new_tensor(x.N, 1, 1, x.npos, x.C, Cp=x.C, dt=x.dt)
            src = x

            def fn_c(stream):
                L.check(lib.pv_copy_rows(src.ptr(), dense.ptr(), src.dt, src.N * src.npos, src.C, src.row_stride,
                                         dense.row_stride, stream), "pv_copy_rows(%s)" % name)
            self.add(name + ".dense", fn_c, reads=(src,), writes=(dense,))
            x = dense
        total = x.N * x.npos * x.C
        out = self.new_buf(total, L.PV_F32)

        def fn(stream):
            L.check(lib.pv_ndhwc_to_ncdhw(x.ptr(), x.dt, 1, out.tensor.data_ptr(), 1, 1, 1, 1, total, stream),
                    "pv_ndhwc_to_ncdhw(%s)" % name)
        self.add(name, fn, reads=(x,), writes=(out,))
        return out, ((x.N, x.C) if squeeze and x.npos == 1 else (x.N, x.npos, x.C))

    def concat_channels(self, parts):
        """Fuse torch.cat(parts, dim=1): retarget every part into one wide buffer (no copy)."""
        p0 = parts[0]
        for p in parts:
            assert (p.N, p.T, p.H, p.W) == (p0.N, p0.T, p0.H, p0.W), "concat shape mismatch"
            assert p.C == p.Cp or p is parts[-1], "only the last concat part may carry channel padding"
        c_total = sum(p.C for p in parts)
        cp_total = PK.pad8(sum(p.Cp for p in parts[:-1]) + parts[-1].Cp)
        buf = self.new_buf(p0.N * p0.npos * cp_total)
        off = 0
        for p in parts:
            p.retarget(buf, off, cp_total)
            off += p.Cp
        return TRef(buf, p0.N, p0.T, p0.H, p0.W, c_total, Cp=cp_total, ch_off=0, row_stride=cp_total)


# =============================================================================================
# Token-major (MViT) emitters.  A token tensor [B, Ntok, C] is a TRef with T=H=1, W=Ntok.
# =============================================================================================
def _tok(plan, B, ntok, C, dt=None):
    return plan.new_tensor(B, 1, 1, ntok, C, Cp=C, dt=dt)


def emit_linear(plan, x, weight, bias, act=L.ACT_NONE, residual=None, name="linear"):
    """nn.Linear on tokens = 1x1x1 convolution (tcgen05 implicit GEMM in f16 mode)."""
    w = weight.reshape(weight.shape[0], weight.shape[1], 1, 1, 1)
    return plan.emit_conv(x, w, bias, None, (1, 1, 1), (0, 0, 0), (1, 1, 1), 1, act, residual, name)


def emit_layernorm(plan, x, ln, name="ln", rows_stride=None, rows=None):
    """LayerNorm over the channel dim of every token (or of `rows` rows spaced rows_stride apart)."""
    C = x.C
    assert tuple(ln.normalized_shape) == (C,), "LayerNorm width mismatch"
    g = plan.const(ln.weight.detach().float().cpu())
    b = plan.const(ln.bias.detach().float().cpu())
    eps = float(ln.eps)
    n_rows = x.N * x.npos if rows is None else rows
    xs = x.row_stride if rows_stride is None else rows_stride
    y = plan.new_tensor(x.N, 1, 1, x.npos if rows is None else 1, C, Cp=C)
    lib = plan.lib
    if x.dt != plan.dt:          # fp32 trunk of the f16 engine: f32 in, f16 out
        assert x.dt == L.PV_F32 and plan.dt == L.PV_F16

        def fn32(stream):
            L.check(lib.pv_add_layernorm(x.ptr(), x.dt, xs, None, 0, None, 0, y.ptr(), y.row_stride, n_rows, C,
                                         g.data_ptr(), b.data_ptr(), eps, stream), "pv_add_layernorm(%s)" % name)
        plan.add(name, fn32, "other", 0.0, n_rows * C * 6)
        return y

    def fn(stream):
        L.check(lib.pv_layernorm(x.ptr(), y.ptr(), x.dt, n_rows, 1, C, xs, y.row_stride, g.data_ptr(), b.data_ptr(),
                                 eps, stream), "pv_layernorm(%s)" % name)
    plan.add(name, fn, "other", 0.0, n_rows * C * 4)
    return y


def emit_add_layernorm(plan, a, br, ln, name="add_ln", want_sum=True):
    """fp32 residual trunk of the f16 engine (layers/attention.py:746-757): s = a + br with a f16|f32 and br the f16
    branch output; returns (s as an fp32 token tensor or None, LayerNorm(s) as f16 or None when ln is None)."""
    C = a.C
    assert br.C == C and br.N == a.N and br.npos == a.npos, "residual / branch shape mismatch"
    assert br.dt == L.PV_F16 and plan.dt == L.PV_F16
    assert want_sum or ln is not None
    n_rows = a.N * a.npos
    s = _tok(plan, a.N, a.npos, C, dt=L.PV_F32) if want_sum else None
    y = g = b = None
    eps = 0.0
    if ln is not None:
        assert tuple(ln.normalized_shape) == (C,), "LayerNorm width mismatch"
        g = plan.const(ln.weight.detach().float().cpu())
        b = plan.const(ln.bias.detach().float().cpu())
        eps = float(ln.eps)
        y = _tok(plan, a.N, a.npos, C)
    lib = plan.lib

    def fn(stream):
        L.check(lib.pv_add_layernorm(a.ptr(), a.dt, a.row_stride, br.ptr(), br.row_stride,
                                     s.ptr() if s is not None else None, s.row_stride if s is not None else 0,
                                     y.ptr() if y is not None else None, y.row_stride if y is not None else 0,
                                     n_rows, C, g.data_ptr() if g is not None else None,
                                     b.data_ptr() if b is not None else None, eps, stream), "pv_add_layernorm(%s)" % name)
    plan.add(name, fn, "other", 0.0, n_rows * C * (_ESIZE[a.dt] + 2 + (4 if want_sum else 0) + (2 if ln is not None else 0)))
    return s, y


def emit_pos_cls(plan, x, pos_table, has_cls, name="posenc", out_dt=None):
    """x: patch tokens as produced by the patch-embed conv [B, T', H', W', C] -> [B, cls+THW, C] (out_dt = PV_F32
    starts the fp32 residual trunk of the f16 engine)."""
    n_patch = x.npos
    C = x.C
    pos = plan.const(pos_table.float().contiguous())
    y = _tok(plan, x.N, n_patch + (1 if has_cls else 0), C, dt=out_dt)
    lib = plan.lib

    def fn(stream):
        L.check(lib.pv_add_pos_cls_to(x.ptr(), x.dt, y.ptr(), y.dt, x.N, n_patch, C, x.row_stride, pos.data_ptr(),
                                      1 if has_cls else 0, stream), "pv_add_pos_cls_to")
    plan.add(name, fn, "other", 0.0, x.N * n_patch * C * (_ESIZE[x.dt] + _ESIZE[y.dt]))
    return y


def _pool_geometry(pool):
    kind = type(pool).__name__
    if kind == "Conv3d":
        k, s, p, dl = [tuple(int(v) for v in t) for t in (pool.kernel_size, pool.stride, pool.padding, pool.dilation)]
        return kind, k, s, p, dl, int(pool.in_channels)
    if kind == "MaxPool3d":
        k, s, p = [tuple(int(v) for v in (t if isinstance(t, (tuple, list)) else (t,) * 3))
                   for t in (pool.kernel_size, pool.stride, pool.padding)]
        return kind, k, s, p, (1, 1, 1), 0
    raise NotImplementedError("pool module %s unsupported" % kind)


def pools_fusable(pool_a, pool_b, norm_a, norm_b):
    """True when two _AttentionPool branches (pool_k / pool_v) can run as ONE depthwise launch + ONE LayerNorm launch over
    adjacent channel slices: same conv geometry, LayerNorms of the same width / eps."""
    if type(pool_a).__name__ != "Conv3d" or type(pool_b).__name__ != "Conv3d":
        return False
    if _pool_geometry(pool_a) != _pool_geometry(pool_b):
        return False
    na, nb = type(norm_a).__name__, type(norm_b).__name__
    if na != "LayerNorm" or nb != "LayerNorm":
        return False
    return tuple(norm_a.normalized_shape) == tuple(norm_b.normalized_shape) and float(norm_a.eps) == float(norm_b.eps)


def emit_token_pool(plan, x, thw, pool, norm, heads, has_cls, name="pool"):
    """_AttentionPool (layers/attention.py:162-212) on a token tensor/slice x [B, cls+THW, dim]:
    depthwise Conv3d / MaxPool3d over the (T,H,W) grid of the patch tokens (cls row passes through),
    then the per-head LayerNorm over head_dim (cls row included; it is read straight from x by the LayerNorm
    launch).  ``pool`` / ``norm`` may be tuples (pool_k, pool_v) / (norm_k, norm_v): x then holds the branches as
    adjacent channel slices and both run in one depthwise + one LayerNorm launch (see pools_fusable).
    Returns (tokens, thw')."""
    import ctypes as C_
    pools = list(pool) if isinstance(pool, (tuple, list)) else [pool]
    norms = list(norm) if isinstance(norm, (tuple, list)) else [norm] * len(pools)
    nset = len(pools)
    T, H, W = thw
    dim_all = x.C
    assert dim_all % nset == 0
    dim = dim_all // nset
    cls = 1 if has_cls else 0
    assert x.npos == cls + T * H * W, "token count does not match thw"
    kind, k, s, p, dl, pool_ch = _pool_geometry(pools[0])
    for q in pools[1:]:
        assert _pool_geometry(q) == (kind, k, s, p, dl, pool_ch)
    if kind == "Conv3d":
        for q in pools:
            if q.groups != q.in_channels or q.in_channels != q.out_channels or q.bias is not None:
                raise NotImplementedError("%s: only depthwise, bias-free pooling convs are supported" % name)
        if dim % pool_ch:
            raise RuntimeError("%s: pool channels do not divide the token width" % name)
    To = (T + 2 * p[0] - dl[0] * (k[0] - 1) - 1) // s[0] + 1
    Ho = (H + 2 * p[1] - dl[1] * (k[1] - 1) - 1) // s[1] + 1
    Wo = (W + 2 * p[2] - dl[2] * (k[2] - 1) - 1) // s[2] + 1
    y = _tok(plan, x.N, cls + To * Ho * Wo, dim_all, dt=x.dt)
    lib = plan.lib
    esz = _ESIZE[x.dt]
    if kind == "Conv3d":
        reps = dim // pool_ch
        w_full = torch.cat([q.weight.detach().cpu().repeat(reps, 1, 1, 1, 1) for q in pools], 0)   # same filter for every head
        w_d = plan.const(PK.pack_depthwise(w_full, dim_all, _TORCH_DT[x.dt]))
        ones = plan.const(torch.ones(dim_all, dtype=torch.float32))
        zeros = plan.const(torch.zeros(dim_all, dtype=torch.float32))
        d = L.Conv3dDesc()
        d.dtype = x.dt
        d.N, d.Ti, d.Hi, d.Wi, d.Ci = x.N, T, H, W, dim_all
        d.To, d.Ho, d.Wo, d.Co = To, Ho, Wo, dim_all
        d.kt, d.kh, d.kw = k
        d.st, d.sh, d.sw = s
        d.pt, d.ph, d.pw = p
        d.dt, d.dh, d.dw = dl
        d.groups, d.act, d.has_residual = dim_all, L.ACT_NONE, 0

        def fn(stream):
            d.x_row_stride, d.y_row_stride = x.row_stride, y.row_stride
            d.x_batch_stride, d.y_batch_stride = x.npos * x.row_stride, y.npos * y.row_stride
            # depthwise entry point: lane-per-channel-pair stencil for 3x3x3 in f16, generic stencil otherwise; the
            # batch strides step over the cls row in front of every sample
            L.check(lib.pv_dwconv3d_fwd(C_.byref(d), x.ptr() + cls * x.row_stride * esz, w_d.data_ptr(),
                                        ones.data_ptr(), zeros.data_ptr(), y.ptr() + cls * y.row_stride * esz, None,
                                        stream), "pv_dwconv3d_fwd(%s)" % name)
        plan.add(name + ".dwconv", fn, "depthwise", 2.0 * x.N * To * Ho * Wo * dim_all * k[0] * k[1] * k[2],
                 (x.N * T * H * W + x.N * To * Ho * Wo) * dim_all * esz)
    else:
        d = L.Pool3dDesc()
        d.dtype, d.mode = x.dt, L.POOL_MAX
        d.N, d.Ti, d.Hi, d.Wi, d.C = x.N, T, H, W, dim_all
        d.To, d.Ho, d.Wo = To, Ho, Wo
        d.kt, d.kh, d.kw = k
        d.st, d.sh, d.sw = s
        d.pt, d.ph, d.pw = p

        def fn(stream):
            d.x_row_stride, d.y_row_stride = x.row_stride, y.row_stride
            d.x_batch_stride, d.y_batch_stride = x.npos * x.row_stride, y.npos * y.row_stride
            L.check(lib.pv_pool3d_fwd(C_.byref(d), x.ptr() + cls * x.row_stride * esz,
                                      y.ptr() + cls * y.row_stride * esz, stream), "pv_pool3d_fwd(%s)" % name)
        plan.add(name + ".maxpool", fn, "other", 0.0, (x.N * T * H * W + x.N * To * Ho * Wo) * dim_all * esz)
    have_norm = [n is not None and type(n).__name__ != "Identity" for n in norms]
    if any(have_norm) and not all(have_norm):
        raise NotImplementedError("%s: fused pooling branches need a norm on every branch" % name)
    if not all(have_norm):
        if cls:
            def fn_cls(stream):
                L.check(lib.pv_copy_rows(x.ptr(), y.ptr(), x.dt, x.N, dim_all, x.npos * x.row_stride, y.npos * y.row_stride,
                                         stream), "pv_copy_rows(%s)" % name)
            plan.add(name + ".cls", fn_cls)
        return y, (To, Ho, Wo)
    for n in norms:
        if type(n).__name__ != "LayerNorm":
            raise NotImplementedError("%s: pool norm %s unsupported" % (name, type(n).__name__))
    hd = int(norms[0].normalized_shape[0])
    eps = float(norms[0].eps)
    for n in norms[1:]:
        assert int(n.normalized_shape[0]) == hd and float(n.eps) == eps
    assert dim % hd == 0
    g = plan.const(torch.cat([n.weight.detach().float().cpu() for n in norms]))
    b = plan.const(torch.cat([n.bias.detach().float().cpu() for n in norms]))

    def fn_ln(stream):
        # in place on the pooled rows; the cls row of every sample (row % npos == 0) is read from x: no copy launch
        L.check(lib.pv_layernorm_sets(y.ptr(), y.ptr(), y.dt, y.N * y.npos, dim_all // hd, hd, y.row_stride, y.row_stride,
                                      g.data_ptr(), b.data_ptr(), dim // hd, x.ptr() if cls else None,
                                      x.npos * x.row_stride, y.npos, eps, stream), "pv_layernorm_sets(%s)" % name)
    plan.add(name + ".norm", fn_ln, "other", 0.0, 2 * y.N * y.npos * dim_all * esz)
    return y, (To, Ho, Wo)


def emit_attention(plan, q, k, v, heads, scale, residual_pool, name="attn"):
    import ctypes as C_
    B, Nq, Nk, dim = q.N, q.npos, k.npos, q.C
    assert k.C == dim and v.C == dim and v.npos == Nk and dim % heads == 0
    o = _tok(plan, B, Nq, dim)
    d = L.AttentionDesc()
    d.dtype, d.B, d.H, d.Nq, d.Nk, d.D = plan.dt, B, heads, Nq, Nk, dim // heads
    d.scale, d.add_q_residual = float(scale), 1 if residual_pool else 0
    lib = plan.lib

    def fn(stream):
        d.q_row_stride, d.k_row_stride, d.v_row_stride, d.o_row_stride = q.row_stride, k.row_stride, v.row_stride, o.row_stride
        d.q_batch_stride, d.k_batch_stride = Nq * q.row_stride, Nk * k.row_stride
        d.v_batch_stride, d.o_batch_stride = Nk * v.row_stride, Nq * o.row_stride
        L.check(lib.pv_attention_fwd(C_.byref(d), q.ptr(), k.ptr(), v.ptr(), o.ptr(), stream), "pv_attention_fwd(%s)" % name)
    plan.add(name, fn, "attention", 4.0 * B * heads * Nq * Nk * (dim // heads),
             (B * Nq * dim * 2 + 2 * B * Nk * dim) * _ESIZE[plan.dt])
    return o


def channel_slice(x, off, C):
    """View of C channels starting at `off` (no copy)."""
    t = TRef(x.buf, x.N, x.T, x.H, x.W, C, Cp=C, ch_off=x.ch_off + off, row_stride=x.row_stride)
    return t
