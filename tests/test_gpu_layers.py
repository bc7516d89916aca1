"""GPU: the `pytorchvideo.layers` / stem / head surface at LAYER granularity.

Every case builds this package's module with the reference's constructor arguments, runs its ``forward`` on the
B200 engine and compares with the output of the REAL reference class (tests/golden/layers.pt, produced by
oracle/gen_golden.py with the weights copied by load_state_dict(strict=True)) and with the oracle re-run here.
Weights and inputs lie on the f16 grid, so reference and engine multiply identical operands.
Tolerance (f16 storage): |d| <= 2e-3*|ref| + 1e-3*max|ref|; f32 mode: north-star rtol 1e-3 / atol 1e-4*scale.
The shape checks restate the reference's own layer tests (tests/test_layers_attention.py:16-103,
tests/test_layers_convolutions.py) with channel widths the tensor-core path supports (multiples of 8,
head_dim in {32, 64, 96})."""
import os

import pytest
import torch

from oracle.interp import oracle_forward
from pytorchvideo_b200 import testing as TS

pytestmark = pytest.mark.gpu
GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layers.pt"), weights_only=False)


def _run(m, x, thw):
    m = m.cuda()
    out = m(x.cuda(), thw) if thw is not None else m(x.cuda())
    m.cpu()
    return out


@pytest.mark.parametrize("name", TS.LAYER_CASES)
def test_layer_forward_matches_reference_golden_f16(name):
    m, x, thw = TS.build_layer_case(name)
    g = GOLD[name]
    assert abs(TS.state_checksum(m) - g["state_checksum"]) <= 1e-6 * abs(g["state_checksum"])
    out = _run(m, x, thw)
    if thw is not None:
        out, thw_out = out
        assert list(thw_out) == g["thw_out"]
    out = out.float().cpu()
    ref = g["output"]
    assert out.shape == ref.shape
    scale = float(ref.abs().max())
    err = (out - ref).abs()
    inside = float((err <= 1e-3 * ref.abs() + 1e-4 * max(1.0, scale)).float().mean())
    print("PARITY layer %s f16: max|d|/max|ref| = %.3e, in-band %.3f" % (name, float(err.max()) / scale, inside))
    assert bool((err <= 2e-3 * ref.abs() + 1e-3 * scale).all()), float(err.max()) / scale
    orc = oracle_forward(m, x, thw) if thw is not None else oracle_forward(m, x)
    orc = orc[0] if thw is not None else orc
    assert float((orc - ref).abs().max()) <= 1e-4 * max(1.0, scale)


@pytest.mark.parametrize("name", ["conv_reduce_sum", "conv_reduce_cat", "conv2plus1d_xy_first", "mlp", "block_widen_pool",
                                  "attention_residual_pool_nocls", "posenc", "patch_embed", "vit_head"])
def test_layer_forward_f32_parity_mode(name):
    from pytorchvideo_b200 import config
    m, x, thw = TS.build_layer_case(name)
    ref = GOLD[name]["output"]
    config.set_precision("f32")
    try:
        out = _run(m, x, thw)
    finally:
        config.set_precision("f16")
    out = (out[0] if thw is not None else out).float().cpu()
    scale = max(1.0, float(ref.abs().max()))
    assert bool(((out - ref).abs() <= 1e-3 * ref.abs() + 1e-4 * scale).all()), float((out - ref).abs().max())


def test_attention_and_block_shapes_like_the_reference_tests():
    """tests/test_layers_attention.py:16-103 with tensor-core friendly widths: seq = 1 + 2*2*5 tokens."""
    from pytorchvideo_b200.layers import MultiScaleAttention, MultiScaleBlock, Mlp
    seq_len, c_dim, c_out, thw = 21, 64, 128, (2, 2, 5)
    x = torch.rand(8, seq_len, c_dim).cuda()
    out, shp = MultiScaleAttention(c_dim, num_heads=2).eval().cuda()(x, thw)
    assert out.shape == (8, seq_len, c_dim) and list(shp) == [2, 2, 5]
    out, shp = MultiScaleAttention(c_dim, dim_out=c_out, num_heads=2).eval().cuda()(x, thw)
    assert out.shape == (8, seq_len, c_out)
    out, shp = MultiScaleAttention(c_dim, num_heads=2, stride_q=(2, 2, 1)).eval().cuda()(x, thw)
    assert out.shape == (8, 6, c_dim) and list(shp) == [1, 1, 5]
    xn = torch.rand(8, 20, c_dim).cuda()
    out, shp = MultiScaleAttention(c_dim, num_heads=2, stride_q=(2, 2, 1), has_cls_embed=False).eval().cuda()(xn, thw)
    assert out.shape == (8, 5, c_dim) and list(shp) == [1, 1, 5]
    out, shp = MultiScaleBlock(c_dim, c_out, 2).eval().cuda()(x, thw)
    assert out.shape == (8, seq_len, c_out) and list(shp) == [2, 2, 5]
    out, shp = MultiScaleBlock(c_dim, c_out, 2, dim_mul_in_att=True).eval().cuda()(x, thw)
    assert out.shape == (8, seq_len, c_out)
    out, shp = MultiScaleBlock(c_dim, c_out, 2, stride_q=(2, 2, 1)).eval().cuda()(x, thw)
    assert out.shape == (8, (seq_len - 1) // 4 + 1, c_out) and list(shp) == [1, 1, 5]
    # Mlp on (B, C) like tests/test_layers_attention.py:105-130
    y = Mlp(in_features=64, hidden_features=32, out_features=24).eval().cuda()(torch.rand(8, 64).cuda())
    assert y.shape == (8, 24)
    # wrong token count for the given thw is an error, not a silent reshape
    with pytest.raises(RuntimeError):
        MultiScaleBlock(c_dim, c_out, 2).eval().cuda()(x, (2, 2, 4))


def test_swish_and_squeeze_excitation_stand_alone():
    from pytorchvideo_b200.layers.swish import Swish
    from pytorchvideo_b200.layers.squeeze_excitation import SqueezeExcitation
    x = TS.f16_exact(torch.randn(2, 16, 3, 6, 6))
    y = Swish().eval().cuda()(x.cuda()).float().cpu()
    ref = x * torch.sigmoid(x)
    assert torch.allclose(y, ref, rtol=2e-3, atol=2e-3)
    se = TS.randomize_model(SqueezeExcitation(16, 8), seed=3, f16_weights=True).eval()
    ref = oracle_forward(se, x)
    y = se.cuda()(x.cuda()).float().cpu()
    assert torch.allclose(y, ref, rtol=2e-3, atol=2e-3 * float(ref.abs().max()))
