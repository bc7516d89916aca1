"""GPU: fused clip transform vs the numpy oracle and the reference-generated goldens."""
import os

import numpy as np
import pytest
import torch

from oracle import transforms_ref as O
from pytorchvideo_b200 import testing as TS

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_small_chain_goldens_f32_and_f16():
    from pytorchvideo_b200.transforms import FusedClipTransform
    for c in torch.load(os.path.join(GOLD, "transforms.pt"), weights_only=False)["chain_small"]:
        clip = TS.synthetic_u8_clip(c["T"], c["H"], c["W"], seed=c["seed"]).cuda()
        for dt, rtol, atol in ((torch.float32, 1e-5, 2e-6), (torch.float16, 1e-3, 1e-4)):
            tr = FusedClipTransform(c["n"], c["mean"], c["std"], short_side=c["side"], crop=("center", c["crop"]),
                                    out_dtype=dt)
            out = tr(clip).float().cpu()
            assert out.shape == c["out"].shape
            assert torch.allclose(out, c["out"], rtol=rtol, atol=atol), float((out - c["out"]).abs().max())


def test_thwc_strided_input_equals_cthw():
    """Decoder hand-off layout (data/utils.py:26-31): logical CTHW that is physically THWC."""
    from pytorchvideo_b200.transforms import FusedClipTransform
    clip = TS.synthetic_u8_clip(12, 50, 70, seed=9)
    thwc = clip.permute(1, 2, 3, 0).contiguous().cuda()
    view = thwc.permute(3, 0, 1, 2)            # (C,T,H,W) view with THWC strides
    tr = FusedClipTransform(5, (0.4, 0.5, 0.6), (0.2, 0.25, 0.3), short_side=32, crop=("center", 24), out_dtype=torch.float32)
    a, b = tr(view).cpu(), tr(clip.cuda()).cpu()
    assert torch.equal(a, b)
    ref = O.val_chain(clip.numpy(), 5, (0.4, 0.5, 0.6), (0.2, 0.25, 0.3), 32, 24)
    np.testing.assert_allclose(a.numpy(), ref, rtol=1e-5, atol=2e-6)


def test_full_size_config5_against_reference_golden():
    from pytorchvideo_b200.transforms import FusedClipTransform
    g = torch.load(os.path.join(GOLD, "transforms.pt"), weights_only=False)["chain_full"]
    clip = TS.synthetic_u8_clip(g["T"], g["H"], g["W"], seed=g["seed"]).cuda()
    tr = FusedClipTransform(g["n"], g["mean"], g["std"], short_side=g["side"], crop=("center", g["crop"]),
                            out_dtype=torch.float32)
    out = tr(clip).cpu()
    assert out.shape == (3, 16, 224, 224)
    assert torch.allclose(out[:, ::5, ::7, ::9], g["sample"], rtol=1e-5, atol=2e-6)
    cs = TS.tensor_checksum(out)
    np.testing.assert_allclose(cs, g["checksum"], rtol=1e-5)
    out16 = FusedClipTransform(g["n"], g["mean"], g["std"], short_side=g["side"], crop=("center", g["crop"]),
                               out_dtype=torch.float16)(clip).float().cpu()
    assert torch.allclose(out16, out, rtol=1e-3, atol=1e-4)


def test_single_ops_and_properties():
    from pytorchvideo_b200 import transforms as T
    x = torch.rand(3, 9, 20, 10, generator=torch.Generator().manual_seed(1))
    xd = x.cuda()
    # subsample: bit-exact frame selection (integer index work)
    sub = T.UniformTemporalSubsample(4)(xd).cpu()
    assert torch.equal(sub, x[:, O.linspace_indices(9, 4)])
    # short side scale shape + values (tests/test_transforms.py:104-144)
    sss = T.ShortSideScale(5)(xd).cpu()
    assert sss.shape == (3, 9, 10, 5)
    np.testing.assert_allclose(sss.numpy(), O.short_side_scale(x.numpy(), 5), rtol=1e-5, atol=1e-6)
    # normalize by own stats -> mean 0 / std 1 (tests/test_transforms.py:324-332)
    m, s = x.mean(dim=(1, 2, 3)), x.std(dim=(1, 2, 3), unbiased=False)
    y = T.Normalize(m.tolist(), s.tolist())(xd).cpu()
    assert torch.allclose(y.mean(dim=(1, 2, 3)), torch.zeros(3), atol=1e-5)
    assert torch.allclose(y.std(dim=(1, 2, 3), unbiased=False), torch.ones(3), atol=1e-4)
    # uint8 path == float/255 path (tests/test_transforms.py:1100-1126)
    u8 = TS.synthetic_u8_clip(6, 16, 16, seed=2)
    a = T.ConvertUint8ToFloat()(u8.cuda()).cpu()
    assert torch.equal(a, u8.float() / 255.0)
    # identity resize is exact
    same = T.ShortSideScale(10)(xd).cpu()
    assert torch.equal(same, x)


def test_train_chain_matches_reference_goldens_under_fixed_seed():
    """create_video_transform(mode="train") = ONE kernel launch; random short side, crop offsets and the flip are
    drawn on the host from torch's global RNG in the reference's order, so under the reference's seed the fused
    output equals the REAL reference chain's (tests/golden/transforms.pt["train_small"])."""
    from pytorchvideo_b200.transforms import create_video_transform
    for c in torch.load(os.path.join(GOLD, "transforms.pt"), weights_only=False)["train_small"]:
        clip = TS.synthetic_u8_clip(c["T"], c["H"], c["W"], seed=c["seed"]).cuda()
        for dt, rtol, atol in ((torch.float32, 1e-5, 2e-6), (torch.float16, 1e-3, 1e-4)):
            tr = create_video_transform(mode="train", num_samples=c["n"], min_size=c["min_size"], max_size=c["max_size"],
                                        crop_size=c["crop"], out_dtype=dt)
            torch.manual_seed(c["rng_seed"])
            out = tr(clip).float().cpu()
            assert out.shape == c["out"].shape
            assert torch.allclose(out, c["out"], rtol=rtol, atol=atol), (c["draws"], float((out - c["out"]).abs().max()))


def test_random_crop_and_short_side_modules_follow_torchvision_under_seed():
    import torchvision.transforms as TV
    from pytorchvideo_b200 import transforms as T
    x = torch.rand(3, 5, 40, 60, generator=torch.Generator().manual_seed(5))
    torch.manual_seed(123)
    ref = TV.RandomCrop(24)(x)
    torch.manual_seed(123)
    got = T.RandomCropVideo(24)(x.cuda()).cpu()
    assert torch.equal(got, ref)
    torch.manual_seed(7)
    size = torch.randint(20, 31, (1,)).item()
    torch.manual_seed(7)
    got = T.RandomShortSideScale(20, 30)(x.cuda()).cpu()
    np.testing.assert_allclose(got.numpy(), O.short_side_scale(x.numpy(), size), rtol=1e-5, atol=1e-6)
    # dict-level factory output (video_key / remove_key) like transforms_factory.py:262-284
    tr = T.create_video_transform(mode="val", video_key="video", remove_key=["audio"], num_samples=4, min_size=32, crop_size=24,
                                  out_dtype=torch.float32)
    sample = {"video": TS.synthetic_u8_clip(8, 40, 60, seed=1).cuda(), "audio": 1, "label": 3}
    out = tr(sample)
    assert "audio" not in out and out["label"] == 3 and out["video"].shape == (3, 4, 24, 24)
    with pytest.raises(NotImplementedError):
        T.create_video_transform(mode="train", aug_type="randaug")
