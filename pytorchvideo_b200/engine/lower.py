"""Lower a PyTorchVideo-style module tree into a Plan (structural walk, no tracing).

Dispatch is by class *name* and attribute names, exactly the names the reference uses
(models/net.py, resnet.py, stem.py, head.py, slowfast.py, x3d.py, layers/convolutions.py), so the
same lowering accepts this package's own parameter-container modules and - where the reference
is importable - the reference's modules themselves (state_dict-compatible drop-in).
"""
import torch
import torch.nn as nn

from .. import _lib as L
from .plan import Plan, TRef


def _t3(v):
    if isinstance(v, (tuple, list)):
        assert len(v) == 3
        return tuple(int(i) for i in v)
    return (int(v),) * 3


def _act_code(m):
    if m is None:
        return L.ACT_NONE
    n = type(m).__name__
    if n == "ReLU":
        return L.ACT_RELU
    if n == "Swish":
        return L.ACT_SWISH
    if n == "GELU":
        return L.ACT_GELU
    if n == "Sigmoid":
        return L.ACT_SIGMOID
    if n == "Identity":
        return L.ACT_NONE
    raise NotImplementedError("activation %s has no B200 kernel" % n)


def _is_bn(m):
    return isinstance(m, nn.modules.batchnorm._BatchNorm) or type(m).__name__.startswith("NaiveSyncBatchNorm")


class Lowering:
    def __init__(self, plan: Plan, extra=()):
        self.p = plan
        self.extra = tuple(extra)     # non-tensor forward arguments of the root module (thw_shape)
        self.aux_out = None           # host-side second result of the root module (pooled thw)

    # ---- leaf helpers --------------------------------------------------------------------
    def conv(self, x: TRef, conv: nn.Conv3d, bn=None, act=None, residual=None, name="conv", se_sums=False):
        if type(conv).__name__ == "Conv2plus1d":
            return self.conv2plus1d(x, conv, bn, act, residual, name)
        if not isinstance(conv, nn.Conv3d):
            raise NotImplementedError("%s: conv module %s unsupported" % (name, type(conv).__name__))
        if isinstance(conv.padding, str):
            raise NotImplementedError("string padding unsupported")
        if conv.padding_mode != "zeros":
            raise NotImplementedError("padding_mode %s unsupported" % conv.padding_mode)
        if bn is not None and not _is_bn(bn):
            raise NotImplementedError("%s: norm %s unsupported (BatchNorm only)" % (name, type(bn).__name__))
        return self.p.emit_conv(x, conv.weight, conv.bias, bn, _t3(conv.stride), _t3(conv.padding),
                                _t3(conv.dilation), conv.groups, _act_code(act), residual, name, se_sums=se_sums)

    def conv2plus1d(self, x, m, bn, act, residual, name):
        # layers/convolutions.py:232-237: conv_t -> norm -> activation -> conv_xy, or conv_xy first when
        # conv_xy_first is set (the flag only swaps the two convolutions; norm/activation stay in between)
        first, second = ("conv_xy", "conv_t") if getattr(m, "conv_xy_first", False) else ("conv_t", "conv_xy")
        h = self.conv(x, getattr(m, first), getattr(m, "norm", None), getattr(m, "activation", None), None,
                      name + "." + first)
        return self.conv(h, getattr(m, second), bn, act, residual, name + "." + second)

    def lower_Conv2plus1d(self, m, x, name):
        return self.conv2plus1d(x, m, None, None, None, name)

    def lower_Conv3d(self, m, x, name):
        return self.conv(x, m, None, None, None, name or "conv")

    def lower_ConvReduce3D(self, m, x, name):
        # layers/convolutions.py:77-85: parallel convolutions of one input, summed (stack().sum(0)) or
        # concatenated along channels.  sum: every conv after the first takes the running sum as the fused
        # residual of its epilogue (fp32 add on the accumulator); cat: producers write channel slices.
        outs = []
        acc = None
        for i, c in enumerate(m.convs):
            if m.reduction_method == "sum":
                acc = self.conv(x, c, None, None, acc, "%s.convs.%d" % (name, i))
            else:
                outs.append(self.conv(x, c, None, None, None, "%s.convs.%d" % (name, i)))
        if m.reduction_method == "sum":
            return acc
        for o in outs[:-1]:
            if o.C != o.Cp:
                raise NotImplementedError("ConvReduce3D(cat): out_channels must be a multiple of 8 for the fused concat")
        return self.p.concat_channels(outs) if len(outs) > 1 else outs[0]

    def lower_Swish(self, m, x, name):
        self.p.materialize_input(x)
        return self.p.emit_act(x, L.ACT_SWISH, name or "swish")

    def lower_ReLU(self, m, x, name):
        self.p.materialize_input(x)
        return self.p.emit_act(x, L.ACT_RELU, name or "relu")

    def lower_SqueezeExcitation(self, m, x, name):
        blk = m.block
        if type(blk[1]).__name__ != "ReLU" or type(blk[3]).__name__ != "Sigmoid":
            raise NotImplementedError("SqueezeExcitation variant unsupported")
        self.p.materialize_input(x)
        return self.p.emit_se_scale_act(x, blk[0].weight, blk[0].bias, blk[2].weight, blk[2].bias, L.ACT_NONE,
                                        name or "se")

    def pool(self, x, m, name="pool"):
        n = type(m).__name__
        if n == "MaxPool3d":
            if _t3(m.dilation) != (1, 1, 1) or m.ceil_mode:
                raise NotImplementedError("MaxPool3d dilation/ceil_mode unsupported")
            k = _t3(m.kernel_size)
            s = _t3(m.stride if m.stride is not None else m.kernel_size)
            return self.p.emit_pool(x, L.POOL_MAX, k, s, _t3(m.padding), name)
        if n == "AvgPool3d":
            if m.ceil_mode or not m.count_include_pad or m.divisor_override is not None:
                raise NotImplementedError("AvgPool3d options unsupported")
            k = _t3(m.kernel_size)
            s = _t3(m.stride if m.stride is not None else m.kernel_size)
            return self.p.emit_pool(x, L.POOL_AVG, k, s, _t3(m.padding), name)
        if n == "AdaptiveAvgPool3d":
            osz = _t3(m.output_size)
            if osz != (1, 1, 1):
                raise NotImplementedError("AdaptiveAvgPool3d output_size %s unsupported" % (osz,))
            k = (x.T, x.H, x.W)
            return self.p.emit_pool(x, L.POOL_AVG, k, k, (0, 0, 0), name)
        if n == "Identity":
            return x
        raise NotImplementedError("pool module %s unsupported" % n)

    # ---- blocks --------------------------------------------------------------------------
    def lower(self, m, x, name=""):
        n = type(m).__name__
        fn = getattr(self, "lower_" + n, None)
        if fn is None:
            raise NotImplementedError("no B200 lowering for module %s (%s)" % (n, name))
        return fn(m, x, name)

    def lower_Identity(self, m, x, name):
        return x

    def lower_Dropout(self, m, x, name):
        return x   # eval mode

    def lower_MaxPool3d(self, m, x, name):
        return self.pool(x, m, name)

    lower_AvgPool3d = lower_MaxPool3d
    lower_AdaptiveAvgPool3d = lower_MaxPool3d

    def lower_Net(self, m, x, name):
        # models/net.py:41-44
        for i, blk in enumerate(m.blocks):
            x = self.lower(blk, x, "%sblocks.%d" % (name + "." if name else "", i))
        return x

    def lower_Sequential(self, m, x, name):
        for i, blk in enumerate(m):
            x = self.lower(blk, x, "%s.%d" % (name, i))
        return x

    def lower_ResNetBasicStem(self, m, x, name):
        # models/stem.py:252-260: conv -> norm -> activation -> pool
        x = self.conv(x, m.conv, m.norm, m.activation, None, name + ".conv")
        if getattr(m, "pool", None) is not None:
            x = self.pool(x, m.pool, name + ".pool")
        return x

    def lower_ResStage(self, m, x, name):
        for i, blk in enumerate(m.res_blocks):
            x = self.lower(blk, x, "%s.res_blocks.%d" % (name, i))
        return x

    def _fusable_bottleneck(self, m, x):
        """ResBlock whose whole body runs as ONE launch of the fused narrow-pathway kernel (csrc/pv_fastblock.cu):
        plain Conv3d / BatchNorm / ReLU bottleneck with a (kt,1,1) conv_a, a dense (1,3,3) conv_b of stride
        (1,s,s) and an inner width of 8 / 16 / 32 channels (the SlowFast Fast pathway, res2-res4)."""
        import ctypes as C_
        import os
        p = self.p
        if not p.use_tcgen05 or os.environ.get("PVB200_NO_FUSED"):
            return False
        b = m.branch2
        if type(b).__name__ != "BottleneckBlock" or not isinstance(x, TRef):
            return False
        convs = (b.conv_a, b.conv_b, b.conv_c) + ((m.branch1_conv,) if m.branch1_conv is not None else ())
        for c in convs:
            if type(c) is not nn.Conv3d or c.groups != 1 or _t3(c.dilation) != (1, 1, 1) or c.padding_mode != "zeros" \
                    or isinstance(c.padding, str):
                return False
        for n_ in (b.norm_a, b.norm_b, b.norm_c):
            if n_ is None or not _is_bn(n_):
                return False
        if type(b.act_a).__name__ != "ReLU" or type(b.act_b).__name__ != "ReLU":
            return False
        if m.activation is not None and type(m.activation).__name__ not in ("ReLU", "Identity"):
            return False
        ka, kb, kc = _t3(b.conv_a.kernel_size), _t3(b.conv_b.kernel_size), _t3(b.conv_c.kernel_size)
        if ka[1:] != (1, 1) or ka[0] not in (1, 3) or _t3(b.conv_a.stride) != (1, 1, 1) or _t3(b.conv_a.padding) != (ka[0] // 2, 0, 0):
            return False
        sb = _t3(b.conv_b.stride)
        if kb != (1, 3, 3) or sb[0] != 1 or sb[1] != sb[2] or sb[1] not in (1, 2) or _t3(b.conv_b.padding) != (0, 1, 1):
            return False
        if kc != (1, 1, 1) or _t3(b.conv_c.stride) != (1, 1, 1) or _t3(b.conv_c.padding) != (0, 0, 0):
            return False
        if b.conv_a.in_channels != x.C or x.C != x.Cp:      # (a 3-channel network input is padded to 4: not this kernel)
            return False
        if m.branch1_conv is not None:
            c1 = m.branch1_conv
            if _t3(c1.kernel_size) != (1, 1, 1) or _t3(c1.stride) != (1, sb[1], sb[1]) or _t3(c1.padding) != (0, 0, 0):
                return False
            n1 = getattr(m, "branch1_norm", None)
            if n1 is not None and not _is_bn(n1):
                return False
        act = L.ACT_RELU if (m.activation is not None and type(m.activation).__name__ == "ReLU") else L.ACT_NONE
        d = p.fused_bottleneck_desc(x, x.Cp, b.conv_a.out_channels, b.conv_c.out_channels, ka[0], sb[1],
                                    m.branch1_conv is not None, act)
        return bool(p.lib.pv_bottleneck_fused_supported(C_.byref(d)))

    def lower_ResBlock(self, m, x, name):
        # models/resnet.py:1179-1189; branch_fusion is x + y for every builder in scope.
        if self._fusable_bottleneck(m, x):
            b = m.branch2
            act = L.ACT_RELU if (m.activation is not None and type(m.activation).__name__ == "ReLU") else L.ACT_NONE
            return self.p.emit_bottleneck_fused(x, b.conv_a, b.norm_a, b.conv_b, b.norm_b, b.conv_c, b.norm_c,
                                                m.branch1_conv, getattr(m, "branch1_norm", None), act, name + ".fused")
        if m.branch1_conv is not None:
            shortcut = self.conv(x, m.branch1_conv, getattr(m, "branch1_norm", None), None, None, name + ".branch1")
        else:
            shortcut = x
        return self.bottleneck(m.branch2, x, shortcut, m.activation, name + ".branch2")

    def lower_BottleneckBlock(self, m, x, name):
        return self.bottleneck(m, x, None, None, name)

    def bottleneck(self, m, x, shortcut, final_act, name):
        # models/resnet.py:1345-1365 with the block's residual add + activation fused into conv_c
        if type(m).__name__ != "BottleneckBlock":
            raise NotImplementedError("branch2 module %s unsupported" % type(m).__name__)
        h = self.conv(x, m.conv_a, m.norm_a, m.act_a, None, name + ".conv_a")
        norm_b, se = m.norm_b, None
        if isinstance(norm_b, nn.Sequential):      # X3D: Sequential(BN|Identity, SE|Identity), models/x3d.py:199-208
            assert len(norm_b) == 2
            se = norm_b[1] if type(norm_b[1]).__name__ == "SqueezeExcitation" else None
            if se is None and type(norm_b[1]).__name__ != "Identity":
                raise NotImplementedError("norm_b[1] %s unsupported" % type(norm_b[1]).__name__)
            norm_b = norm_b[0] if _is_bn(norm_b[0]) else None
        if se is None:
            h = self.conv(h, m.conv_b, norm_b, m.act_b, None, name + ".conv_b")
        else:
            h = self.conv(h, m.conv_b, norm_b, None, None, name + ".conv_b", se_sums=True)
            blk = se.block
            if type(blk[1]).__name__ != "ReLU" or type(blk[3]).__name__ != "Sigmoid":
                raise NotImplementedError("SqueezeExcitation variant unsupported")
            h = self.p.emit_se_scale_act(h, blk[0].weight, blk[0].bias, blk[2].weight, blk[2].bias,
                                         _act_code(m.act_b), name + ".se")
        return self.conv(h, m.conv_c, m.norm_c, final_act, shortcut, name + ".conv_c")

    def lower_MultiPathWayWithFuse(self, m, x, name):
        # models/net.py:107-122
        assert isinstance(x, list), "input for MultiPathWayWithFuse needs to be a list of tensors"
        out = list(x)
        # the pathways are independent until the fusion: each gets its own lane (CUDA stream / graph branch)
        for i, blk in enumerate(m.multipathway_blocks):
            if blk is not None:
                self.p.lane = i
                out[i] = self.lower(blk, x[i], "%s.multipathway_blocks.%d" % (name, i))
        self.p.lane = 0
        if m.multipathway_fusion is not None:
            out = self.lower(m.multipathway_fusion, out, name + ".multipathway_fusion")
        return out

    def lower_FuseFastToSlow(self, m, x, name):
        # models/slowfast.py:720-729; the concat is fused away (both producers write into one buffer)
        x_s, x_f = x[0], x[1]
        self.p.lane = 1        # the lateral conv reads the Fast tensor: keep it on the Fast lane, the Slow lane only
        fuse = self.conv(x_f, m.conv_fast_to_slow, m.norm, m.activation, None, name + ".conv_fast_to_slow")
        self.p.lane = 0        # waits for it where the next Slow stage reads the concat buffer
        return [self.p.concat_channels([x_s, fuse]), x_f]

    def lower_PoolConcatPathway(self, m, x, name):
        # models/slowfast.py:608-620
        outs = []
        for i, xi in enumerate(x):
            if xi is None:
                continue
            if m.pool is not None and m.pool[i] is not None:
                self.p.lane = i
                xi = self.pool(xi, m.pool[i], "%s.pool.%d" % (name, i))
                self.p.lane = 0
            outs.append(xi)
        cat = self.p.concat_channels(outs) if len(outs) > 1 else outs[0]
        return [cat] if getattr(m, "retain_list", False) else cat

    def lower_ProjectedPool(self, m, x, name):
        # models/x3d.py:791-806
        x = self.conv(x, m.pre_conv, m.pre_norm, m.pre_act, None, name + ".pre_conv")
        x = self.pool(x, m.pool, name + ".pool")
        return self.conv(x, m.post_conv, m.post_norm, m.post_act, None, name + ".post_conv")

    def lower_DetectionBBoxNetwork(self, m, x, name):
        # models/net.py:62-74: features = model(x); out = detection_head(features, bboxes); out.view(K, -1)
        feats = self.lower(m.model, x[0] if len(x) == 2 else list(x[:-1]), (name + "." if name else "") + "model")
        return self.lower_ResNetRoIHead(m.detection_head, [feats, x[-1]], (name + "." if name else "") + "detection_head")

    def lower_ResNetRoIHead(self, m, x, name):
        # models/head.py:441-482
        if len(x) != 2:
            raise RuntimeError("ResNetRoIHead.forward(x, bboxes) takes one feature tensor and the boxes")
        x, boxes = x
        if isinstance(x, list):
            raise RuntimeError("ResNetRoIHead expects ONE feature tensor (PoolConcatPathway(retain_list=False))")
        name = name or "head"
        if getattr(m, "pool", None) is not None:
            x = self.pool(x, m.pool, name + ".pool")
        roi = getattr(m, "roi_layer", None)
        if roi is not None:
            if type(roi).__name__ != "RoIAlign":
                raise NotImplementedError("roi layer %s unsupported (RoIAlign only)" % type(roi).__name__)
            if getattr(roi, "aligned", False):
                raise NotImplementedError("RoIAlign(aligned=True) unsupported")
            osz = roi.output_size
            osz = (osz, osz) if isinstance(osz, int) else tuple(osz)
            x = self.p.emit_roi_align(x, boxes, osz, roi.spatial_scale, roi.sampling_ratio, name + ".roi_layer")
            ps = getattr(m, "pool_spatial", None)
            if ps is not None:
                pn = type(ps).__name__
                if pn not in ("MaxPool2d", "AvgPool2d"):
                    raise NotImplementedError("pool_spatial %s unsupported" % pn)
                k2 = ps.kernel_size if isinstance(ps.kernel_size, (tuple, list)) else (ps.kernel_size,) * 2
                s2 = ps.stride if isinstance(ps.stride, (tuple, list)) else (ps.stride,) * 2
                p2 = ps.padding if isinstance(ps.padding, (tuple, list)) else (ps.padding,) * 2
                if pn == "MaxPool2d" and (ps.dilation not in (1, (1, 1)) or ps.ceil_mode):
                    raise NotImplementedError("MaxPool2d dilation/ceil_mode unsupported")
                if pn == "AvgPool2d" and (ps.ceil_mode or not ps.count_include_pad or ps.divisor_override is not None):
                    raise NotImplementedError("AvgPool2d options unsupported")
                x = self.p.emit_pool(x, L.POOL_MAX if pn == "MaxPool2d" else L.POOL_AVG, (1,) + tuple(k2), (1,) + tuple(s2),
                                     (0,) + tuple(p2), name + ".pool_spatial")
        proj = m.proj
        if not isinstance(proj, nn.Linear):
            raise NotImplementedError("head proj %s unsupported" % type(proj).__name__)
        act = getattr(m, "activation", None)
        an = None if act is None else type(act).__name__
        if an not in (None, "Sigmoid", "ReLU", "Identity"):
            raise NotImplementedError("RoI head activation %s unsupported" % an)
        w = proj.weight.reshape(proj.out_features, proj.in_features, 1, 1, 1)
        x = self.p.emit_conv(x, w, proj.bias, None, (1, 1, 1), (0, 0, 0), (1, 1, 1), 1,
                             L.ACT_RELU if an == "ReLU" else L.ACT_NONE, None, name + ".proj")
        if an == "Sigmoid":
            x = self.p.emit_act(x, L.ACT_SIGMOID, name + ".activation")
        if getattr(m, "output_pool", None) is not None:
            return self.p.emit_head_reduce(x, False, name + ".output_pool")       # AdaptiveAvgPool3d(1) + view(K, -1)
        return self.p.emit_to_ncdhw(x, name + ".to_ncdhw")

    def lower_ResNetBasicHead(self, m, x, name):
        # models/head.py:371-391
        pool = getattr(m, "pool", None)
        if pool is not None:
            x = self.lower(pool, x, name + ".pool") if type(pool).__name__ == "ProjectedPool" else self.pool(x, pool, name + ".pool")
        proj = m.proj
        if not isinstance(proj, nn.Linear):
            raise NotImplementedError("head proj %s unsupported" % type(proj).__name__)
        w = proj.weight.reshape(proj.out_features, proj.in_features, 1, 1, 1)
        x = self.p.emit_conv(x, w, proj.bias, None, (1, 1, 1), (0, 0, 0), (1, 1, 1), 1, L.ACT_NONE, None, name + ".proj")
        act = getattr(m, "activation", None)
        softmax = False
        if act is not None:
            an = type(act).__name__
            if an == "Softmax":
                if act.dim != 1:
                    raise NotImplementedError("head softmax dim != 1")
                softmax = True
            elif an == "Sigmoid":
                x = self.p.emit_act(x, L.ACT_SIGMOID, name + ".activation")
            else:
                raise NotImplementedError("head activation %s unsupported" % an)
        if getattr(m, "output_pool", None) is not None:
            return self.p.emit_head_reduce(x, softmax, name + ".output_pool")
        if softmax:
            raise NotImplementedError("head softmax without output_pool unsupported")
        return self.p.emit_to_ncdhw(x, name + ".to_ncdhw")


class CompiledModel:
    """A frozen (model, input shapes) plan: static input/output buffers + one CUDA graph.

    Inputs are 5-D clips (B, C, T, H, W) - or, for the MViT layer modules, token tensors (B, N, C) / (B, C).
    ``extra`` carries non-tensor forward arguments (the ``thw_shape`` of MultiScaleBlock / MultiScaleAttention);
    a lowering handler may leave a second, host-side result in ``Lowering.aux_out`` (the pooled thw)."""

    def __init__(self, model, example_inputs, dtype="f16", use_tcgen05=True, use_graph=True, extra=()):
        L.require_device()
        multi = isinstance(example_inputs, (list, tuple))
        ins = list(example_inputs) if multi else [example_inputs]
        raw = _raw_inputs(model, ins)
        for i, t in enumerate(ins):
            if i not in raw and t.dim() not in (2, 3, 5):
                raise RuntimeError("expected a 5-D (B, C, T, H, W) clip or a (B, N, C) token tensor, got %s" % (tuple(t.shape),))
        device = ins[0].device
        if device.type != "cuda":
            raise RuntimeError("pytorchvideo_b200 has no CPU path: inputs must be CUDA tensors")
        dt = {"f16": L.PV_F16, "f32": L.PV_F32}[dtype]
        self.multi = multi
        self.plan = Plan(device, dt, use_tcgen05)
        self.static_in = [torch.empty(t.shape, dtype=t.dtype if (t.dtype in (torch.float16, torch.float32) and i not in raw)
                                      else torch.float32, device=device) for i, t in enumerate(ins)]
        low = Lowering(self.plan, extra)
        xs = [self.plan.raw_input(s) if i in raw else _emit_input(self.plan, s) for i, s in enumerate(self.static_in)]
        out = low.lower(model, xs if multi else xs[0], "")
        out = _emit_output(self.plan, out, tokens=ins[0].dim() != 5)
        self.out_buf, self.out_shape = out
        self.aux = low.aux_out
        self.plan.finalize()
        self.graph = None
        self.use_graph = use_graph
        self.key = tuple((tuple(t.shape), t.dtype) for t in ins)

    def _capture(self):
        dev = self.plan.device
        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            self.plan.run(stream.cuda_stream)       # warm-up (also sets func attributes outside capture)
        torch.cuda.current_stream(dev).wait_stream(stream)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            self.plan.run(torch.cuda.current_stream(dev).cuda_stream)
        self.graph = g

    def pipeline(self, depth=2):
        """Double-buffered host-in / host-out serving loop (engine/pipeline.py)."""
        from .pipeline import ClipPipeline
        return ClipPipeline(self, depth)

    def output_view(self):
        return self.out_buf.tensor[: int(torch.tensor(self.out_shape).prod())].view(*self.out_shape)

    def check_inputs(self, inputs):
        """The plan is frozen for one set of input shapes: anything else is an error (Tensor.copy_ would
        silently broadcast a smaller batch into the static buffer)."""
        if self.multi != isinstance(inputs, (list, tuple)):
            raise RuntimeError("this plan was compiled for %s" % ("a list of pathway tensors" if self.multi else "a single tensor"))
        ins = list(inputs) if self.multi else [inputs]
        if len(ins) != len(self.static_in):
            raise RuntimeError("expected %d input tensors, got %d" % (len(self.static_in), len(ins)))
        for s, t in zip(self.static_in, ins):
            if not torch.is_tensor(t) or tuple(t.shape) != tuple(s.shape):
                raise RuntimeError("input shape %s differs from the compiled shape %s (compile a plan per shape)" % (
                    tuple(t.shape) if torch.is_tensor(t) else type(t).__name__, tuple(s.shape)))
            if not (t.is_floating_point() or t.dtype == torch.uint8):
                raise RuntimeError("unsupported input dtype %s" % t.dtype)
        return ins

    def __call__(self, inputs):
        ins = self.check_inputs(inputs)
        dev = self.plan.device
        with torch.cuda.device(dev):
            for s, t in zip(self.static_in, ins):
                if t.data_ptr() != s.data_ptr():
                    s.copy_(t, non_blocking=True)     # H2D or D2D staging into the plan's static input
            if self.use_graph:
                if self.graph is None:
                    self._capture()
                self.graph.replay()
            else:
                self.plan.run(torch.cuda.current_stream(dev).cuda_stream)
        return self.output_view()


def _raw_inputs(model, ins):
    """Indices of inputs that are not clips / tokens: the trailing [K, 5] box tensor of the detection modules
    (DetectionBBoxNetwork.forward(x, bboxes), ResNetRoIHead.forward(x, bboxes))."""
    if type(model).__name__ in ("DetectionBBoxNetwork", "ResNetRoIHead"):
        b = ins[-1]
        if len(ins) < 2 or b.dim() != 2 or b.shape[1] != 5 or b.shape[0] < 1:
            raise RuntimeError("bboxes must be a [K, 5] tensor (batch index, x1, y1, x2, y2) with K >= 1; "
                               "RoIAlignRotated ([K, 6]) is unsupported")
        return {len(ins) - 1}
    return set()


def _emit_input(plan, t):
    if t.dim() == 5:
        return plan.emit_input_ncdhw(t, t.shape[1], 4 if t.shape[1] <= 4 else (t.shape[1] + 7) // 8 * 8)
    return plan.emit_input_tokens(t)


def _emit_output(plan, out, tokens):
    if isinstance(out, TRef):
        if tokens or (out.T == 1 and out.H == 1 and getattr(out, "is_tokens", False)):
            return plan.emit_to_tokens(out, "output.to_tokens", squeeze=getattr(out, "squeeze", False))
        return plan.emit_to_ncdhw(out, "output.to_ncdhw")
    if isinstance(out, list):
        raise NotImplementedError("models returning a list are unsupported")
    return out


def compile_model(model, example_inputs, dtype="f16", use_tcgen05=True, use_graph=True, extra=()):
    return CompiledModel(model, example_inputs, dtype, use_tcgen05, use_graph, extra)


def lower_only(model, example_inputs, dtype="f16", use_tcgen05=True, extra=()):
    """Host-side dry run (works without a GPU): build the plan on the CPU and return
    (plan, output_shape).  Nothing can be executed; used by the CPU test-suite to check the
    lowering, shape inference, channel padding, concat fusion and algorithm selection."""
    multi = isinstance(example_inputs, (list, tuple))
    ins = list(example_inputs) if multi else [example_inputs]
    plan = Plan("cpu", {"f16": L.PV_F16, "f32": L.PV_F32}[dtype], use_tcgen05)
    raw = _raw_inputs(model, ins)
    xs = [plan.raw_input(torch.empty(t.shape, dtype=torch.float32)) if i in raw
          else _emit_input(plan, torch.empty(t.shape, dtype=torch.float32)) for i, t in enumerate(ins)]
    low = Lowering(plan, extra)
    out = low.lower(model, xs if multi else xs[0], "")
    out = _emit_output(plan, out, tokens=ins[0].dim() != 5)
    plan.aux = low.aux_out
    return plan, out[1]


# =============================================================================================
# MViT lowering (models/vision_transformers.py:172-182, layers/attention.py)
# =============================================================================================
from . import plan as PL  # noqa: E402


def _check_thw(x, thw, has_cls, name):
    T, H, W = (int(v) for v in thw)
    if x.npos != (1 if has_cls else 0) + T * H * W:
        raise RuntimeError("%s: %d tokens do not match thw_shape %s%s" % (name, x.npos, (T, H, W), " + cls" if has_cls else ""))
    return (T, H, W)


def _lower_mvit_attention(low, attn, xn, thw, name, residual=None):
    """MultiScaleAttention.forward (layers/attention.py:467-544) on normalised tokens ``xn``:
    returns (proj(attention) [+ residual fused into the GEMM epilogue], pooled q thw)."""
    p = low.p
    if attn.pool_first:
        raise NotImplementedError("pool_first=True is not used by any hub MViT and has no B200 lowering")
    has_cls = attn.has_cls_embed
    heads, dim_att = attn.num_heads, attn.dim_out
    thw = _check_thw(xn, thw, has_cls, name)
    # q/k/v projections as ONE GEMM over concatenated weights; q, k, v are channel slices of its output
    if attn.separate_qkv:
        w = torch.cat([attn.q.weight, attn.k.weight, attn.v.weight], 0)
        b = None if attn.q.bias is None else torch.cat([attn.q.bias, attn.k.bias, attn.v.bias], 0)
    else:
        w, b = attn.qkv.weight, attn.qkv.bias
    qkv = PL.emit_linear(p, xn, w, b, L.ACT_NONE, None, name + ".qkv")
    q, k, v = (PL.channel_slice(qkv, i * dim_att, dim_att) for i in range(3))
    thw_q = thw
    if getattr(attn, "pool_q", None) is not None:
        q, thw_q = PL.emit_token_pool(p, q, thw, attn.pool_q, getattr(attn, "norm_q", None), heads, has_cls, name + ".pool_q")
    pool_k, pool_v = getattr(attn, "pool_k", None), getattr(attn, "pool_v", None)
    norm_k, norm_v = getattr(attn, "norm_k", None), getattr(attn, "norm_v", None)
    if pool_k is not None and pool_v is not None and PL.pools_fusable(pool_k, pool_v, norm_k, norm_v):
        # k | v are adjacent channel slices of the QKV GEMM output: ONE depthwise launch + ONE LayerNorm launch for both
        kv, _ = PL.emit_token_pool(p, PL.channel_slice(qkv, dim_att, 2 * dim_att), thw, (pool_k, pool_v), (norm_k, norm_v),
                                   heads, has_cls, name + ".pool_kv")
        k, v = PL.channel_slice(kv, 0, dim_att), PL.channel_slice(kv, dim_att, dim_att)
    else:
        if pool_k is not None:
            k, _ = PL.emit_token_pool(p, k, thw, pool_k, norm_k, heads, has_cls, name + ".pool_k")
        if pool_v is not None:
            v, _ = PL.emit_token_pool(p, v, thw, pool_v, norm_v, heads, has_cls, name + ".pool_v")
    o = PL.emit_attention(p, q, k, v, heads, attn.scale, attn.residual_pool, name + ".core")
    x = PL.emit_linear(p, o, attn.proj.weight, attn.proj.bias, L.ACT_NONE, residual, name + ".proj")
    return x, thw_q


def _lower_mlp(low, mlp, xn, name, residual=None):
    # layers/attention.py:102-114: fc1 -> act -> fc2 (dropout = identity in eval)
    act = _act_code(mlp.act)
    if act == L.ACT_GELU and getattr(mlp.act, "approximate", "none") != "none":
        raise NotImplementedError("Mlp activation must be the exact (erf) GELU")
    h = PL.emit_linear(low.p, xn, mlp.fc1.weight, mlp.fc1.bias, act, None, name + ".fc1")
    return PL.emit_linear(low.p, h, mlp.fc2.weight, mlp.fc2.bias, L.ACT_NONE, residual, name + ".fc2")


def _lower_mvit_block(low, blk, x, thw, name, xn=None, next_ln=None, want_sum=True):
    """MultiScaleBlock.forward (layers/attention.py:729-757); DropPath is the identity in eval.
    Returns (x, thw', xn_next).  f16 engine with ``plan.trunk32``: the residual stream x is fp32 - the branch outputs
    (attention proj, fc2) stay f16 and each residual add is fused with the LayerNorm that follows it (norm2; ``next_ln`` =
    the next block's norm1 / the model's norm_embed, whose f16 output comes back as xn_next; ``xn`` = this block's
    already normalised input handed over by the previous block)."""
    p = low.p
    attn = blk.attn
    if getattr(blk, "norm1_is_batchnorm_1d", False) or getattr(blk, "norm2_is_batchnorm_1d", False):
        raise NotImplementedError("batchnorm MViT variant unsupported")
    has_cls = attn.has_cls_embed
    thw = _check_thw(x, thw, has_cls, name)
    trunk32 = p.trunk32
    if xn is None:
        xn = PL.emit_layernorm(p, x, blk.norm1, name + ".norm1")
    widen = blk.dim != blk.dim_out
    if blk.dim_mul_in_att and widen:
        x = PL.emit_linear(p, xn, blk.proj.weight, blk.proj.bias, L.ACT_NONE, None, name + ".proj")
    x_res = x
    if getattr(blk, "pool_skip", None) is not None:
        x_res, _ = PL.emit_token_pool(p, x, thw, blk.pool_skip, None, 1, has_cls, name + ".pool_skip")
    if trunk32:
        br, thw_q = _lower_mvit_attention(low, attn, xn, thw, name + ".attn", residual=None)
        x, xn2 = PL.emit_add_layernorm(p, x_res, br, blk.norm2, name + ".norm2")
    else:
        x, thw_q = _lower_mvit_attention(low, attn, xn, thw, name + ".attn", residual=x_res)
        xn2 = PL.emit_layernorm(p, x, blk.norm2, name + ".norm2")
    if (not blk.dim_mul_in_att) and widen:
        x = PL.emit_linear(p, xn2, blk.proj.weight, blk.proj.bias, L.ACT_NONE, None, name + ".proj")
    if trunk32:
        br2 = _lower_mlp(low, blk.mlp, xn2, name + ".mlp", residual=None)
        x, xn_next = PL.emit_add_layernorm(p, x, br2, next_ln, name + ".add", want_sum=want_sum or next_ln is None)
        return x, thw_q, xn_next
    x = _lower_mlp(low, blk.mlp, xn2, name + ".mlp", residual=x)
    return x, thw_q, None


def _pos_table(enc):
    """Rows of the additive table of SpatioTemporalClsPositionalEncoding.forward
    (layers/positional_encoding.py:112-136); row 0 also carries the cls token itself."""
    has_cls = bool(enc.cls_embed_on)
    with torch.no_grad():
        if enc.sep_pos_embed:
            pos = enc.pos_embed_spatial.detach().float().cpu().repeat(1, enc.num_temporal_patch, 1) + \
                torch.repeat_interleave(enc.pos_embed_temporal.detach().float().cpu(), enc.num_spatial_patch, dim=1)
            if has_cls:
                pos = torch.cat([enc.pos_embed_class.detach().float().cpu(), pos], 1)
        else:
            pos = enc.pos_embed.detach().float().cpu().clone()
        pos = pos[0].clone()
        if has_cls:
            pos[0] += enc.cls_token.detach().float().cpu()[0, 0]
    return pos, has_cls


def _lower_vit_head(low, head, x, name="head"):
    """VisionTransformerBasicHead.forward (models/head.py:521-535) on (already normalised) tokens."""
    p = low.p
    mode = head.sequence_pool.mode if head.sequence_pool is not None else None
    if mode == "cls":
        if x.npos > 1:
            x = TRef(x.buf, x.N, 1, 1, 1, x.C, Cp=x.C, ch_off=x.ch_off, row_stride=x.npos * x.row_stride)
    elif mode == "mean":
        x = p.emit_pool(x, L.POOL_AVG, (1, 1, x.W), (1, 1, x.W), (0, 0, 0), name + ".sequence_pool")
    else:
        raise NotImplementedError("sequence_pool=None MViT heads are unsupported")
    x = PL.emit_linear(p, x, head.proj.weight, head.proj.bias, L.ACT_NONE, None, name + ".proj")
    act = head.activation
    softmax = False
    if act is not None:
        if type(act).__name__ == "Softmax":
            softmax = True
        elif type(act).__name__ == "Sigmoid":
            x = p.emit_act(x, L.ACT_SIGMOID, name + ".activation")
        else:
            raise NotImplementedError("head activation unsupported")
    return p.emit_head_reduce(x, softmax, name + ".output")


def _lower_mvit(self, m, x, name):
    p = self.p
    pe = m.patch_embed
    if type(pe).__name__ != "PatchEmbed":
        raise NotImplementedError("MViT without a conv patch embedding is unsupported")
    x = self.conv(x, pe.patch_model, None, None, None, "patch_embed.patch_model")
    enc = m.cls_positional_encoding
    T, H, W = enc.patch_embed_shape()
    if (x.T, x.H, x.W) != (T, H, W):
        raise RuntimeError("input clip gives a %s patch grid but the model was built for %s" % ((x.T, x.H, x.W), (T, H, W)))
    pos, has_cls = _pos_table(enc)
    ne = m.norm_embed
    ne_is_ln = type(ne).__name__ == "LayerNorm"
    if p.trunk32 and not ne_is_ln:
        p.trunk32 = False            # the head's GEMM needs f16 tokens: without a final LayerNorm keep the f16 stream
    x = PL.emit_pos_cls(p, x, pos, has_cls, "cls_positional_encoding", out_dt=L.PV_F32 if p.trunk32 else None)
    thw = (T, H, W)
    xn = None
    nb = len(m.blocks)
    for i, blk in enumerate(m.blocks):
        last = i + 1 == nb
        nxt = (ne if last else m.blocks[i + 1].norm1) if p.trunk32 else None
        x, thw, xn = _lower_mvit_block(self, blk, x, thw, "blocks.%d" % i, xn=xn, next_ln=nxt, want_sum=not last)
    head = m.head
    if type(head).__name__ == "Identity":
        raise NotImplementedError("headless MViT output is unsupported")
    if type(head).__name__ != "VisionTransformerBasicHead":
        raise NotImplementedError("MViT head %s unsupported" % type(head).__name__)
    mode = head.sequence_pool.mode if head.sequence_pool is not None else None
    if p.trunk32:
        x = xn                 # norm_embed was fused into the last block's residual add
    elif ne_is_ln:
        if mode == "cls":      # only the cls row reaches the head: normalise just those B rows
            x = PL.emit_layernorm(p, x, ne, "norm_embed", rows_stride=x.npos * x.row_stride, rows=x.N)
        else:
            x = PL.emit_layernorm(p, x, ne, "norm_embed")
    return _lower_vit_head(self, head, x, "head")


# ---- layer-level entry points: the MViT building blocks as stand-alone modules (token tensors in / out) ----
def _tok_out(x, squeeze=False):
    x.is_tokens = True
    x.squeeze = squeeze
    return x


def _lower_block_module(self, m, x, name):
    if len(self.extra) != 1:
        raise RuntimeError("MultiScaleBlock.forward(x, thw_shape): thw_shape is required")
    y, thw, _ = _lower_mvit_block(self, m, x, self.extra[0], name or "block")
    self.aux_out = list(thw)
    return _tok_out(y)


def _lower_attention_module(self, m, x, name):
    if len(self.extra) != 1:
        raise RuntimeError("MultiScaleAttention.forward(x, thw_shape): thw_shape is required")
    y, thw = _lower_mvit_attention(self, m, x, self.extra[0], name or "attn")
    self.aux_out = list(thw)
    return _tok_out(y)


def _lower_mlp_module(self, m, x, name):
    return _tok_out(_lower_mlp(self, m, x, name or "mlp"), squeeze=True)


def _lower_posenc_module(self, m, x, name):
    pos, has_cls = _pos_table(m)
    T, H, W = m.patch_embed_shape()
    if x.npos != T * H * W:
        raise RuntimeError("expected %d patch tokens, got %d" % (T * H * W, x.npos))
    return _tok_out(PL.emit_pos_cls(self.p, x, pos, has_cls, name or "cls_positional_encoding"))


def _lower_patch_embed_module(self, m, x, name):
    # stem.py:289-292: conv then flatten(2).transpose(1, 2) - the NDHWC conv output already is (B, THW, C)
    y = self.conv(x, m.patch_model, None, None, None, (name + "." if name else "") + "patch_model")
    t = TRef(y.buf, y.N, 1, 1, y.npos, y.C, Cp=y.Cp, ch_off=y.ch_off, row_stride=y.row_stride)
    return _tok_out(t)


def _lower_vit_head_module(self, m, x, name):
    return _lower_vit_head(self, m, x, name or "head")


def _lower_sequence_pool_module(self, m, x, name):
    if m.mode == "cls":
        t = TRef(x.buf, x.N, 1, 1, 1, x.C, Cp=x.C, ch_off=x.ch_off, row_stride=x.npos * x.row_stride)
    else:
        t = self.p.emit_pool(x, L.POOL_AVG, (1, 1, x.W), (1, 1, x.W), (0, 0, 0), name or "sequence_pool")
    return _tok_out(t, squeeze=True)


Lowering.lower_MultiScaleBlock = _lower_block_module
Lowering.lower_MultiScaleAttention = _lower_attention_module
Lowering.lower_Mlp = _lower_mlp_module
Lowering.lower_SpatioTemporalClsPositionalEncoding = _lower_posenc_module
Lowering.lower_PatchEmbed = _lower_patch_embed_module
Lowering.lower_VisionTransformerBasicHead = _lower_vit_head_module
Lowering.lower_SequencePool = _lower_sequence_pool_module
Lowering.lower_MultiscaleVisionTransformers = _lower_mvit
