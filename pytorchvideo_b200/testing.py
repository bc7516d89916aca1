"""Deterministic synthetic weights / inputs shared by tests, bench.py and the golden generator.

Fresh builder init zeroes the last BatchNorm gamma of every residual branch
(models/weight_init.py), which makes every bottleneck a no-op in eval mode - a parity test on such
a model passes even with broken kernels (hard part 1).  ``randomize_model`` gives
every BatchNorm random affine + running statistics (as the reference's tests/test_fuse_bn.py:58-63
does) and re-draws conv / linear weights from a fixed seed, on the CPU generator, so the same
state is reproduced bit-for-bit here and on the GPU box.
"""
import torch
import torch.nn as nn


def f16_exact(t):
    """Round to the nearest f16-representable value (kept in fp32)."""
    return t.half().float()


@torch.no_grad()
def randomize_model(model, seed=1234, f16_weights=False):
    """f16_weights: draw the conv / linear / positional weights on the f16 grid (fp32 tensors whose
    values are exactly representable in f16).  The reference CPU forward and the f16 tensor-core path
    then multiply IDENTICAL weights - what is left between them is activation rounding and summation
    order, not an input difference (the weight rounding of arbitrary fp32 weights alone moves 8-25 %
    of the logits out of the rtol 1e-3 / atol 1e-4 band, tests/test_oracle_pinning.py)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)

    def u(t, lo, hi):
        t.copy_(torch.rand(t.shape, generator=g, dtype=torch.float32) * (hi - lo) + lo)

    for m in model.modules():
        if isinstance(m, (nn.Conv3d, nn.Conv2d)):
            fan_in = m.weight.shape[1] * m.weight[0, 0].numel()
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5)
            if m.bias is not None:
                u(m.bias, -0.1, 0.1)
        elif isinstance(m, nn.modules.batchnorm._BatchNorm):
            if getattr(m, "block_final_bn", False):
                u(m.weight, 0.2, 0.6)       # keep the residual stream well conditioned
            else:
                u(m.weight, 0.5, 1.5)
            u(m.bias, -0.5, 0.5)
            u(m.running_var, 0.5, 1.5)
            u(m.running_mean, -0.5, 0.5)
        elif isinstance(m, nn.Linear):
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (1.0 / m.weight.shape[1]) ** 0.5)
            if m.bias is not None:
                u(m.bias, -0.1, 0.1)
        elif isinstance(m, nn.LayerNorm):
            u(m.weight, 0.5, 1.5)
            u(m.bias, -0.2, 0.2)
    for name, p in model.named_parameters():
        if name.endswith(("cls_token", "pos_embed_spatial", "pos_embed_temporal", "pos_embed_class", "pos_embed")):
            p.copy_(torch.randn(p.shape, generator=g) * 0.2)
    if f16_weights:
        for m in model.modules():
            if isinstance(m, (nn.Conv3d, nn.Conv2d, nn.Linear)):
                m.weight.copy_(f16_exact(m.weight))
    return model


def synthetic_clip(batch, t, h, w, seed=42, channels=3, f16_values=False):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    x = torch.rand((batch, channels, t, h, w), generator=g, dtype=torch.float32)
    return f16_exact(x) if f16_values else x


def slowfast_inputs(clip, alpha=4):
    """[slow, fast] as the reference builds them (uniform_temporal_subsample_repeated with
    frame_ratios=(alpha, 1), tests/test_models_slowfast.py:142-144): slow = frames at
    linspace(0, T-1, T//alpha).long()."""
    T = clip.shape[2]
    idx = torch.clamp(torch.linspace(0, T - 1, T // alpha), 0, T - 1).long()
    return [torch.index_select(clip, 2, idx), clip]


def synthetic_u8_clip(t, h, w, seed=0, channels=3):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randint(0, 256, (channels, t, h, w), generator=g, dtype=torch.uint8)


def tensor_checksum(t):
    t = t.detach().double().cpu().reshape(-1)
    return [float(t.sum()), float(t.abs().sum()), float((t * t).sum())]


def state_checksum(model):
    s = 0.0
    for _, v in sorted(model.state_dict().items()):
        if v.is_floating_point():
            s += float(v.double().abs().sum())
    return s


MODEL_CASES = {
    # name: (hub builder name, kwargs, batch, T, H, W, is_slowfast)
    "x3d_xs": ("x3d_xs", {}, 2, 4, 160, 160, False),
    "x3d_m": ("x3d_m", {}, 1, 16, 224, 224, False),
    "slowfast_r50": ("slowfast_r50", {}, 1, 32, 224, 224, True),
    "slow_r50": ("slow_r50", {}, 1, 8, 224, 224, False),
    "csn_r101": ("csn_r101", {}, 1, 32, 224, 224, False),
    "r2plus1d_r50": ("r2plus1d_r50", {}, 1, 16, 224, 224, False),
    "i3d_r50": ("i3d_r50", {}, 1, 8, 224, 224, False),
    "mvit_base_16x4": ("mvit_base_16x4", {}, 1, 16, 224, 224, False),
    # same architecture on a 8x112x112 clip (785 tokens): cheap enough for the f32 CUDA-core parity mode
    "mvit_base_8x112": ("mvit_base_16x4", {"spatial_size": 112, "temporal_size": 8}, 2, 8, 112, 112, False),
    # further hub entries of the same families: goldens pin the oracle / module trees (CPU); not in the GPU lists yet
    "slowfast_r101": ("slowfast_r101", {}, 1, 32, 224, 224, True),
    "c2d_r50": ("c2d_r50", {}, 1, 8, 224, 224, False),
    "x3d_s": ("x3d_s", {}, 1, 13, 160, 160, False),
    "x3d_l": ("x3d_l", {}, 1, 16, 312, 312, False),
    "mvit_base_32x3": ("mvit_base_32x3", {}, 1, 32, 224, 224, False),
    # BASELINE.json configs at their REAL batch sizes, weights and clips drawn on the f16 grid (CASE_OPTS)
    "c1_x3d_xs": ("x3d_xs", {}, 1, 4, 160, 160, False),
    "c2_slowfast_r50_b8": ("slowfast_r50", {}, 8, 32, 224, 224, True),
    "c3_mvit_base_16x4_b8": ("mvit_base_16x4", {}, 8, 16, 224, 224, False),
    "c4_x3d_m_b32": ("x3d_m", {}, 32, 16, 224, 224, False),
    "slow_r50_f16w": ("slow_r50", {}, 2, 8, 224, 224, False),
    "mvit_base_8x112_f16w": ("mvit_base_16x4", {"spatial_size": 112, "temporal_size": 8}, 2, 8, 112, 112, False),
}
# per-case options of the golden generator / tests: f16_grid = weights AND input clip exactly representable in f16
CASE_OPTS = {
    # X3D-L (312^2, 55 blocks): with random BatchNorm statistics the residual stream grows to |x| ~ 1e3 by stage 5 and the
    # squeeze-excitation gates there saturate (pre-activations ~ +-140), so a gate near its zero crossing turns a 1e-4
    # relative change of a channel mean (the f16 rounding of the WEIGHTS) into a percent-level change of that channel
    # (tools/debug_block.py x3d_l 4 6: 3.5e-3 with the SE, 8.7e-4 without, same with the CUDA-core depthwise kernel).
    # On the f16 grid both sides multiply identical weights and the comparison is well conditioned again.
    "x3d_l": {"f16_grid": True},
    "c1_x3d_xs": {"f16_grid": True},
    "c2_slowfast_r50_b8": {"f16_grid": True},
    "c3_mvit_base_16x4_b8": {"f16_grid": True},
    "c4_x3d_m_b32": {"f16_grid": True},
    "slow_r50_f16w": {"f16_grid": True},
    "mvit_base_8x112_f16w": {"f16_grid": True},
}


def build_case(case, hub_module, weight_seed=1234, input_seed=42):
    """(model, inputs, is_slowfast) of a MODEL_CASES entry, built from ``hub_module`` (this package's hub)."""
    hub, kw, B, T, H, W, is_sf = MODEL_CASES[case]
    grid = CASE_OPTS.get(case, {}).get("f16_grid", False)
    model = randomize_model(getattr(hub_module, hub)(**kw), seed=weight_seed, f16_weights=grid).eval()
    clip = synthetic_clip(B, T, H, W, seed=input_seed, f16_values=grid)
    return model, (slowfast_inputs(clip) if is_sf else clip), is_sf



# ---- layer-level cases (tests/golden/layers.pt): name -> (builder, input shape, thw or None) --------------------
def _layer_builders():
    import torch.nn as nn
    from functools import partial
    from .layers.attention import Mlp, MultiScaleAttention, MultiScaleBlock
    from .layers.convolutions import ConvReduce3D, create_conv_2plus1d
    from .layers.positional_encoding import SpatioTemporalClsPositionalEncoding
    from .models.head import create_vit_basic_head
    from .models.stem import create_conv_patch_embed
    ln = partial(nn.LayerNorm, eps=1e-6)
    return {
        "conv_reduce_sum": (lambda: ConvReduce3D(in_channels=16, out_channels=32, kernel_size=((1, 1, 1), (3, 3, 3), (1, 3, 3)),
                                                 stride=((1, 1, 1), (1, 1, 1), None), padding=((0, 0, 0), (1, 1, 1), (0, 1, 1)),
                                                 bias=(False, True, None), reduction_method="sum"), (2, 16, 4, 12, 12), None),
        "conv_reduce_cat": (lambda: ConvReduce3D(in_channels=16, out_channels=24, kernel_size=((1, 1, 1), (3, 1, 1)),
                                                 padding=((0, 0, 0), (1, 0, 0)), reduction_method="cat"), (2, 16, 4, 12, 12), None),
        "conv2plus1d_xy_first": (lambda: create_conv_2plus1d(in_channels=16, out_channels=32, inner_channels=24, conv_xy_first=True,
                                                             stride=(1, 2, 2)), (2, 16, 4, 12, 12), None),
        "conv2plus1d": (lambda: create_conv_2plus1d(in_channels=16, out_channels=32, stride=(2, 1, 1)), (2, 16, 4, 12, 12), None),
        "mlp": (lambda: Mlp(in_features=96, hidden_features=384, out_features=192), (2, 50, 96), None),
        # MViT-B block 1 geometry at a small grid: Q pooled (1,2,2), K/V pooled (1,4,4), 2 heads of 96
        "attention_pool_qkv": (lambda: MultiScaleAttention(192, num_heads=2, qkv_bias=True, kernel_q=(3, 3, 3), kernel_kv=(3, 3, 3),
                                                           stride_q=(1, 2, 2), stride_kv=(1, 4, 4), norm_layer=ln,
                                                           residual_pool=False), (2, 1 + 4 * 8 * 8, 192), (4, 8, 8)),
        "attention_residual_pool_nocls": (lambda: MultiScaleAttention(64, num_heads=2, kernel_kv=(3, 3, 3), stride_kv=(1, 2, 2),
                                                                      has_cls_embed=False, norm_layer=ln, residual_pool=True),
                                          (2, 2 * 8 * 8, 64), (2, 8, 8)),
        "block_widen_pool": (lambda: MultiScaleBlock(96, 192, 1, qkv_bias=True, norm_layer=ln, attn_norm_layer=ln,
                                                     kernel_q=(3, 3, 3), kernel_kv=(3, 3, 3), stride_q=(1, 2, 2),
                                                     stride_kv=(1, 2, 2)), (2, 1 + 4 * 8 * 8, 96), (4, 8, 8)),
        "block_dim_mul_in_att": (lambda: MultiScaleBlock(64, 128, 2, qkv_bias=True, norm_layer=ln, attn_norm_layer=ln,
                                                         dim_mul_in_att=True, kernel_kv=(3, 3, 3), stride_kv=(1, 2, 2)),
                                 (2, 1 + 2 * 8 * 8, 64), (2, 8, 8)),
        "posenc": (lambda: SpatioTemporalClsPositionalEncoding(96, (4, 7, 7), sep_pos_embed=True, has_cls=True), (2, 4 * 7 * 7, 96), None),
        "patch_embed": (lambda: create_conv_patch_embed(in_channels=3, out_channels=96, conv_kernel_size=(3, 7, 7),
                                                        conv_stride=(2, 4, 4), conv_padding=(1, 3, 3)), (2, 3, 8, 56, 56), None),
        "vit_head": (lambda: create_vit_basic_head(in_features=192, out_features=40, seq_pool_type="cls"), (3, 17, 192), None),
    }


LAYER_CASES = tuple(sorted(["conv_reduce_sum", "conv_reduce_cat", "conv2plus1d_xy_first", "conv2plus1d", "mlp",
                            "attention_pool_qkv", "attention_residual_pool_nocls", "block_widen_pool",
                            "block_dim_mul_in_att", "posenc", "patch_embed", "vit_head"]))


def build_layer_case(name, seed=77):
    """(module, input, thw) with weights AND input on the f16 grid (same operands for reference and engine)."""
    make, shape, thw = _layer_builders()[name]
    torch.manual_seed(seed)
    m = randomize_model(make(), seed=seed, f16_weights=True).eval()
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + 1)
    x = f16_exact(torch.randn(shape, generator=g) if len(shape) == 3 else torch.rand(shape, generator=g))
    return m, x, thw


# ---- detection cases (tests/golden/detection.pt): trunk + RoIAlign head, weights / clips on the f16 grid ----------
DETECTION_CASES = {
    # name: (hub builder, kwargs, batch, T, H, W, is_slowfast, number of boxes)
    # head_activation=None: compare LOGITS (with random weights the default Sigmoid saturates to 0 / 1 and hides errors)
    "slow_r50_detection": ("slow_r50_detection", {"head_activation": None}, 2, 4, 128, 128, False, 6),
    "slowfast_r50_detection": ("slowfast_r50_detection", {"head_activation": None}, 2, 32, 128, 128, True, 5),
    "slow_r50_detection_sigmoid": ("slow_r50_detection", {}, 1, 4, 96, 96, False, 4),      # the hub default head
}


def synthetic_boxes(n_boxes, batch, H, W, seed=9):
    """[K, 5] fp32 (batch index, x1, y1, x2, y2) in input pixels: random boxes plus the shapes the RoIAlign boundary
    rules care about - one reaching outside the frame, one smaller than a feature cell, the whole frame."""
    g = torch.Generator().manual_seed(seed)
    b = torch.randint(0, batch, (n_boxes,), generator=g).float()
    x1 = torch.rand(n_boxes, generator=g) * W * 0.6
    y1 = torch.rand(n_boxes, generator=g) * H * 0.6
    x2 = x1 + 8 + torch.rand(n_boxes, generator=g) * W * 0.5
    y2 = y1 + 8 + torch.rand(n_boxes, generator=g) * H * 0.5
    boxes = torch.stack([b, x1, y1, x2, y2], 1)
    if n_boxes >= 3:
        boxes[0, 1:] = torch.tensor([-20.0, -12.0, W + 30.0, H * 0.5])      # sticks out left / top / right
        boxes[1, 1:] = torch.tensor([W * 0.5, H * 0.5, W * 0.5 + 3.0, H * 0.5 + 2.0])   # sub-cell box -> size clamp 1
        boxes[2, 1:] = torch.tensor([0.0, 0.0, float(W), float(H)])
    return boxes.contiguous()


def build_detection_case(case, hub_module, weight_seed=1234, input_seed=42):
    """(model, clip inputs, boxes, is_slowfast) of a DETECTION_CASES entry."""
    hub, kw, B, T, H, W, is_sf, K = DETECTION_CASES[case]
    model = randomize_model(getattr(hub_module, hub)(**kw), seed=weight_seed, f16_weights=True).eval()
    clip = synthetic_clip(B, T, H, W, seed=input_seed, f16_values=True)
    return model, (slowfast_inputs(clip) if is_sf else clip), synthetic_boxes(K, B, H, W), is_sf


def roi_align_case(seed=5):
    """Op-level RoIAlign case: feature map on the f16 grid + boxes (see synthetic_boxes), several (output size, scale,
    sampling ratio) settings.  Golden = torchvision.ops.roi_align (tests/golden/detection.pt["roi_align"])."""
    g = torch.Generator().manual_seed(seed)
    x = f16_exact(torch.randn(2, 24, 9, 11, generator=g))
    boxes = synthetic_boxes(7, 2, 144, 176, seed=seed + 1)
    settings = [((7, 7), 1.0 / 16.0, 0), ((7, 7), 1.0 / 16.0, 2), ((3, 5), 0.25, 0), ((1, 1), 1.0 / 16.0, 3)]
    return x, boxes, settings
