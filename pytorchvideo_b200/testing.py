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
