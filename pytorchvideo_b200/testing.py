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


@torch.no_grad()
def randomize_model(model, seed=1234):
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
    return model


def synthetic_clip(batch, t, h, w, seed=42, channels=3):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.rand((batch, channels, t, h, w), generator=g, dtype=torch.float32)


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
}
