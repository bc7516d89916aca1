"""ORACLE (test infrastructure only - never imported by the product package).

A CPU fp32 restatement of the reference's eval-mode forward for the hot-path model families,
written as a structural interpreter over a module tree: it is handed ANY module whose class and
attribute names follow facebookresearch/pytorchvideo (the reference's own modules, or this
repo's parameter containers) and evaluates it with plain torch.nn.functional ops on the CPU.

Pinning (see oracle/gen_golden.py, tests/test_oracle_pinning.py): in the authoring container the
interpreter is run over the REAL reference models built from /root/reference and must reproduce
``reference_model(x)`` bit-for-bit; the reference outputs are also committed as golden vectors
under tests/golden/ so the pin travels to the GPU box where /root/reference does not exist.

Each handler cites the reference file:line it restates (paths relative to the reference root).
"""
import torch
import torch.nn.functional as F


def _bn(x, m):
    # nn.BatchNorm3d in eval mode: (x - running_mean) / sqrt(running_var + eps) * weight + bias
    return F.batch_norm(x, m.running_mean, m.running_var, m.weight, m.bias, False, 0.0, m.eps)


def _leaf(m, x):
    """torch.nn leaves (the L0 numerics the reference delegates to ATen)."""
    n = type(m).__name__
    if n == "Conv3d":
        return F.conv3d(x, m.weight, m.bias, m.stride, m.padding, m.dilation, m.groups)
    if n in ("BatchNorm3d", "BatchNorm1d"):
        return _bn(x, m)
    if n == "ReLU":
        return F.relu(x)
    if n == "Swish":                       # layers/swish.py:25-28  x * sigmoid(x)
        return x * torch.sigmoid(x)
    if n == "Sigmoid":
        return torch.sigmoid(x)
    if n == "Softmax":
        return F.softmax(x, dim=m.dim)
    if n == "GELU":
        return F.gelu(x)
    if n == "MaxPool3d":
        return F.max_pool3d(x, m.kernel_size, m.stride, m.padding, m.dilation, m.ceil_mode)
    if n == "AvgPool3d":
        return F.avg_pool3d(x, m.kernel_size, m.stride, m.padding, m.ceil_mode, m.count_include_pad)
    if n == "AdaptiveAvgPool3d":
        return F.adaptive_avg_pool3d(x, m.output_size)
    if n == "Linear":
        return F.linear(x, m.weight, m.bias)
    if n == "LayerNorm":
        return F.layer_norm(x, m.normalized_shape, m.weight, m.bias, m.eps)
    if n in ("Identity", "Dropout", "DropPath"):   # eval mode
        return x
    return None


def roi_align_ref(x, rois, output_size, spatial_scale, sampling_ratio):
    """torchvision.ops.roi_align(aligned=False) restated (torchvision/csrc/ops/cpu/roi_align_kernel.cpp +
    roi_align_common.h pre_calc_for_bilinear_interpolate); the reference's RoI head calls it through
    torchvision.ops.RoIAlign (models/head.py:209-227, 470).  fp32 scalar arithmetic in the C++ order (numpy float32),
    tensor arithmetic over the channel dim in torch.  x: [N, C, H, W]; rois: [K, 5]; returns [K, C, ph, pw]."""
    import numpy as np
    f = np.float32
    N, C, H, W = x.shape
    ph_n, pw_n = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
    K = rois.shape[0]
    out = torch.zeros((K, C, ph_n, pw_n), dtype=torch.float32)
    r = rois.detach().float().cpu().numpy().astype(np.float32)
    scale = f(spatial_scale)
    for k in range(K):
        n = int(r[k, 0])
        roi_start_w, roi_start_h = r[k, 1] * scale, r[k, 2] * scale
        roi_end_w, roi_end_h = r[k, 3] * scale, r[k, 4] * scale
        roi_w = max(roi_end_w - roi_start_w, f(1.0))
        roi_h = max(roi_end_h - roi_start_h, f(1.0))
        bin_h, bin_w = roi_h / f(ph_n), roi_w / f(pw_n)
        grid_h = sampling_ratio if sampling_ratio > 0 else int(np.ceil(roi_h / f(ph_n)))
        grid_w = sampling_ratio if sampling_ratio > 0 else int(np.ceil(roi_w / f(pw_n)))
        count = f(max(grid_h * grid_w, 1))
        feat = x[n]
        for ph in range(ph_n):
            for pw in range(pw_n):
                acc = torch.zeros(C, dtype=torch.float32)
                for iy in range(grid_h):
                    yy = roi_start_h + f(ph) * bin_h + f(iy + 0.5) * bin_h / f(grid_h)
                    for ix in range(grid_w):
                        xx = roi_start_w + f(pw) * bin_w + f(ix + 0.5) * bin_w / f(grid_w)
                        y_, x_ = yy, xx
                        if y_ < -1.0 or y_ > H or x_ < -1.0 or x_ > W:
                            continue                                   # all four weights are 0
                        y_ = max(y_, f(0.0))
                        x_ = max(x_, f(0.0))
                        y_low, x_low = int(y_), int(x_)
                        if y_low >= H - 1:
                            y_high = y_low = H - 1
                            y_ = f(y_low)
                        else:
                            y_high = y_low + 1
                        if x_low >= W - 1:
                            x_high = x_low = W - 1
                            x_ = f(x_low)
                        else:
                            x_high = x_low + 1
                        ly, lx = y_ - f(y_low), x_ - f(x_low)
                        hy, hx = f(1.0) - ly, f(1.0) - lx
                        w1, w2, w3, w4 = hy * hx, hy * lx, ly * hx, ly * lx
                        acc = acc + (float(w1) * feat[:, y_low, x_low] + float(w2) * feat[:, y_low, x_high]
                                     + float(w3) * feat[:, y_high, x_low] + float(w4) * feat[:, y_high, x_high])
                out[k, :, ph, pw] = acc / float(count)
    return out


class Oracle:
    def __call__(self, m, x):
        return self.run(m, x)

    def run(self, m, x):
        if m is None:
            return x
        n = type(m).__name__
        fn = getattr(self, "f_" + n, None)
        if fn is not None:
            return fn(m, x)
        y = _leaf(m, x)
        if y is None:
            raise NotImplementedError("oracle: no restatement for %s" % n)
        return y

    # ---- containers -----------------------------------------------------------------------
    def f_Net(self, m, x):                       # models/net.py:41-44
        for blk in m.blocks:
            x = self.run(blk, x)
        return x

    def f_Sequential(self, m, x):
        for blk in m:
            x = self.run(blk, x)
        return x

    def f_MultiPathWayWithFuse(self, m, x):      # models/net.py:107-122 (without the in-place aliasing)
        assert isinstance(x, list)
        out = list(x)
        for i, blk in enumerate(m.multipathway_blocks):
            if blk is not None:
                out[i] = self.run(blk, x[i])
        if m.multipathway_fusion is not None:
            out = self.run(m.multipathway_fusion, out)
        return out

    def f_FuseFastToSlow(self, m, x):            # models/slowfast.py:720-729
        x_s, x_f = x[0], x[1]
        fuse = self.run(m.conv_fast_to_slow, x_f)
        fuse = self.run(m.norm, fuse)
        fuse = self.run(m.activation, fuse)
        return [torch.cat([x_s, fuse], 1), x_f]

    def f_PoolConcatPathway(self, m, x):         # models/slowfast.py:608-620
        outs = []
        for i, xi in enumerate(x):
            if xi is None:
                continue
            if m.pool is not None and m.pool[i] is not None:
                xi = self.run(m.pool[i], xi)
            outs.append(xi)
        cat = torch.cat(outs, 1)
        return [cat] if m.retain_list else cat

    # ---- CNN blocks -----------------------------------------------------------------------
    def f_ResNetBasicStem(self, m, x):           # models/stem.py:252-260
        x = self.run(m.conv, x)
        x = self.run(m.norm, x)
        x = self.run(m.activation, x)
        return self.run(m.pool, x)

    def f_Conv2plus1d(self, m, x):               # layers/convolutions.py:232-237
        first, second = (m.conv_xy, m.conv_t) if m.conv_xy_first else (m.conv_t, m.conv_xy)
        x = self.run(first, x)
        x = self.run(m.norm, x) if m.norm else x
        x = self.run(m.activation, x) if m.activation else x
        return self.run(second, x)

    def f_ConvReduce3D(self, m, x):              # layers/convolutions.py:77-85
        outs = [self.run(c, x) for c in m.convs]
        if m.reduction_method == "sum":
            return torch.stack(outs, dim=0).sum(dim=0, keepdim=False)
        return torch.cat(outs, dim=1)

    def f_BottleneckBlock(self, m, x):           # models/resnet.py:1345-1365
        x = self.run(m.act_a, self.run(m.norm_a, self.run(m.conv_a, x)))
        x = self.run(m.act_b, self.run(m.norm_b, self.run(m.conv_b, x)))
        return self.run(m.norm_c, self.run(m.conv_c, x))

    def f_SqueezeExcitation(self, m, x):
        # fvcore.nn.squeeze_excitation (not vendored): x * block(mean_{T,H,W} x), block =
        # Conv3d(C,Cr,1,bias) -> ReLU -> Conv3d(Cr,C,1,bias) -> Sigmoid; structure pinned by
        # layers/accelerator/mobile_cpu/attention.py:62-104.
        g = x.mean(dim=[2, 3, 4], keepdim=True)
        for blk in m.block:
            g = self.run(blk, g)
        return x * g

    def f_ResBlock(self, m, x):                  # models/resnet.py:1179-1189
        y = self.run(m.branch2, x)
        if m.branch1_conv is None:
            out = x + y
        else:
            s = self.run(m.branch1_conv, x)
            if m.branch1_norm is not None:
                s = self.run(m.branch1_norm, s)
            out = s + y
        return self.run(m.activation, out)

    def f_ResStage(self, m, x):                  # models/resnet.py:1397-1400
        for blk in m.res_blocks:
            x = self.run(blk, x)
        return x

    def f_ProjectedPool(self, m, x):             # models/x3d.py:791-806
        x = self.run(m.pre_act, self.run(m.pre_norm, self.run(m.pre_conv, x)))
        x = self.run(m.pool, x)
        return self.run(m.post_act, self.run(m.post_norm, self.run(m.post_conv, x)))

    def f_ResNetBasicHead(self, m, x):           # models/head.py:371-391
        x = self.run(m.pool, x)
        x = self.run(m.dropout, x)
        x = x.permute((0, 2, 3, 4, 1))
        x = self.run(m.proj, x)
        x = x.permute((0, 4, 1, 2, 3))
        x = self.run(m.activation, x)            # applied BEFORE the global average
        if m.output_pool is not None:
            x = self.run(m.output_pool, x)
            x = x.view(x.shape[0], -1)
        return x

    def f_ResNetRoIHead(self, m, x, bboxes):     # models/head.py:441-482
        x = self.run(m.pool, x)
        if m.roi_layer is not None:
            if x.shape[-3] != 1:
                raise Exception("Temporal dimension should be 1. Consider modifying the pool layer.")
            x = torch.squeeze(x, -3)
            r = m.roi_layer                      # torchvision.ops.RoIAlign(output_size, spatial_scale, sampling_ratio)
            assert not getattr(r, "aligned", False)
            x = roi_align_ref(x, bboxes, r.output_size, r.spatial_scale, r.sampling_ratio)
            if m.pool_spatial is not None:
                ps = m.pool_spatial
                if type(ps).__name__ == "MaxPool2d":
                    x = F.max_pool2d(x, ps.kernel_size, ps.stride, ps.padding, ps.dilation, ps.ceil_mode)
                else:
                    x = F.avg_pool2d(x, ps.kernel_size, ps.stride, ps.padding, ps.ceil_mode, ps.count_include_pad)
            x = x.unsqueeze(-3)
        x = self.run(m.dropout, x)
        x = x.permute((0, 2, 3, 4, 1))
        x = self.run(m.proj, x)
        x = x.permute((0, 4, 1, 2, 3))
        x = self.run(m.activation, x)
        if m.output_pool is not None:
            x = self.run(m.output_pool, x)
            x = x.view(x.shape[0], -1)
        return x

    def f_DetectionBBoxNetwork(self, m, x, bboxes):   # models/net.py:62-74
        features = self.run(m.model, x)
        out = self.f_ResNetRoIHead(m.detection_head, features, bboxes)
        return out.view(out.shape[0], -1)

    # ---- MViT ----------------------------------------------------------------------------
    def f_PatchEmbed(self, m, x):                # models/stem.py:289-292
        x = self.run(m.patch_model, x)
        return x.flatten(2).transpose(1, 2)

    def f_SpatioTemporalClsPositionalEncoding(self, m, x):   # layers/positional_encoding.py:112-136
        B, N, C = x.shape
        if m.cls_embed_on:
            x = torch.cat((m.cls_token.expand(B, -1, -1), x), dim=1)
        if m.sep_pos_embed:
            pos = m.pos_embed_spatial.repeat(1, m.num_temporal_patch, 1) + torch.repeat_interleave(
                m.pos_embed_temporal, m.num_spatial_patch, dim=1)
            if m.cls_embed_on:
                pos = torch.cat([m.pos_embed_class, pos], 1)
            return x + pos
        return x + m.pos_embed

    def _attention_pool(self, x, pool, thw, has_cls, norm):  # layers/attention.py:162-212
        if pool is None:
            return x, thw
        ndim = x.ndim
        if ndim == 3:
            x = x.unsqueeze(1)
        if has_cls:
            cls_tok, x = x[:, :, :1, :], x[:, :, 1:, :]
        B, Nh, L, C = x.shape
        T, H, W = thw
        x = x.reshape(B * Nh, T, H, W, C).permute(0, 4, 1, 2, 3).contiguous()
        x = self.run(pool, x)
        thw = [x.shape[2], x.shape[3], x.shape[4]]
        x = x.reshape(B, Nh, C, thw[0] * thw[1] * thw[2]).transpose(2, 3)
        if has_cls:
            x = torch.cat((cls_tok, x), dim=2)
        if norm is not None:
            x = self.run(norm, x)
        if ndim == 3:
            x = x.squeeze(1)
        return x, thw

    def f_MultiScaleAttention(self, m, x, thw):  # layers/attention.py:501-544
        B, N, C = x.shape
        H = m.num_heads
        if m.separate_qkv:
            q = self.run(m.q, x).reshape(B, N, H, -1).permute(0, 2, 1, 3)
            k = self.run(m.k, x).reshape(B, N, H, -1).permute(0, 2, 1, 3)
            v = self.run(m.v, x).reshape(B, N, H, -1).permute(0, 2, 1, 3)
        else:
            qkv = self.run(m.qkv, x).reshape(B, N, 3, H, -1).permute(2, 0, 3, 1, 4)
            q, k, v = qkv[0], qkv[1], qkv[2]
        q, q_thw = self._attention_pool(q, m.pool_q, thw, m.has_cls_embed, getattr(m, "norm_q", None))
        k, _ = self._attention_pool(k, m.pool_k, thw, m.has_cls_embed, getattr(m, "norm_k", None))
        v, _ = self._attention_pool(v, m.pool_v, thw, m.has_cls_embed, getattr(m, "norm_v", None))
        attn = (q * m.scale) @ k.transpose(-2, -1)
        attn = attn.softmax(dim=-1)
        N = q.shape[2]
        if m.residual_pool:
            x = (attn @ v + q).transpose(1, 2).reshape(B, -1, m.dim_out)
        else:
            x = (attn @ v).transpose(1, 2).reshape(B, -1, m.dim_out)
        x = self.run(m.proj, x)
        return x, q_thw

    def f_Mlp(self, m, x):                       # layers/attention.py:102-114 (dropout = identity)
        x = self.run(m.fc2, self.run(m.act, self.run(m.fc1, x)))
        return x

    def f_MultiScaleBlock(self, m, x, thw):      # layers/attention.py:729-757
        x_norm = self.run(m.norm1, x)
        x_block, thw_new = self.f_MultiScaleAttention(m.attn, x_norm, thw)
        if m.dim_mul_in_att and m.dim != m.dim_out:
            x = self.run(m.proj, x_norm)
        x_res, _ = self._attention_pool(x, m.pool_skip, thw, m.has_cls_embed, None)
        x = x_res + x_block
        x_norm = self.run(m.norm2, x)
        x_mlp = self.f_Mlp(m.mlp, x_norm)
        if not m.dim_mul_in_att and m.dim != m.dim_out:
            x = self.run(m.proj, x_norm)
        return x + x_mlp, thw_new

    def f_MultiscaleVisionTransformers(self, m, x):   # models/vision_transformers.py:172-182
        if m.patch_embed is not None:
            x = self.run(m.patch_embed, x)
        x = self.run(m.cls_positional_encoding, x)
        if m.pos_drop is not None:
            x = self.run(m.pos_drop, x)
        thw = list(m.cls_positional_encoding.patch_embed_shape()) if hasattr(
            m.cls_positional_encoding, "patch_embed_shape") and callable(
            m.cls_positional_encoding.patch_embed_shape) else list(m.cls_positional_encoding.patch_embed_shape)
        for blk in m.blocks:
            x, thw = self.f_MultiScaleBlock(blk, x, thw)
        if m.norm_embed is not None:
            x = self.run(m.norm_embed, x)
        if m.head is not None:
            x = self.run(m.head, x)
        return x

    def f_SequencePool(self, m, x):              # models/head.py:29-36
        if m.mode == "cls":
            return x[:, 0]
        if m.mode == "mean":
            return x.mean(1)
        raise NotImplementedError

    def f_VisionTransformerBasicHead(self, m, x):    # models/head.py:521-535
        x = self.run(m.sequence_pool, x)
        x = self.run(m.dropout, x)
        x = self.run(m.proj, x)
        return self.run(m.activation, x)


def oracle_forward(model, x, *extra):
    """Eval-mode fp32 CPU forward of ``model`` (a module tree, read only) on ``x``.  ``extra``: the
    non-tensor forward arguments of the MViT layer modules (thw_shape); those return (tensor, thw)."""
    with torch.no_grad():
        if isinstance(x, (list, tuple)):
            x = [t.detach().float().cpu() for t in x]
        else:
            x = x.detach().float().cpu()
        if extra and type(model).__name__ in ("DetectionBBoxNetwork", "ResNetRoIHead"):
            return getattr(Oracle(), "f_" + type(model).__name__)(model, x, extra[0].detach().float().cpu())
        if extra:
            fn = getattr(Oracle(), "f_" + type(model).__name__)
            y, thw = fn(model, x, list(extra[0]))
            return y, list(thw)
        return Oracle().run(model, x)
