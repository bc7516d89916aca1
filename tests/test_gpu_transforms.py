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


def test_batched_chain_equals_per_clip_chain_and_emits_slowfast_pathways():
    """pv_clip_transform_batch: one launch over a batch, taps computed in the kernel (no tables) - bit-equal to
    the table-driven single-clip kernel; the optional second output is the SlowFast slow pathway
    (pytorchvideo_trainer/datamodule/transforms.py:129-136), bit-equal to index_select on the fast output."""
    from pytorchvideo_b200.transforms import FusedClipTransform, SlowFastPackPathway
    from pytorchvideo_b200.transforms import functional as Fv
    clips = torch.stack([TS.synthetic_u8_clip(40, 90, 160, seed=30 + i) for i in range(3)]).cuda()      # (B, C, T, H, W)
    mean, std = (0.45, 0.45, 0.45), (0.225, 0.225, 0.225)
    for dt in (torch.float32, torch.float16):
        tr = FusedClipTransform(32, mean, std, short_side=48, crop=("center", 40), out_dtype=dt)
        batch = tr(clips)
        assert batch.shape == (3, 3, 32, 40, 40)
        idx, hw, win, _ = tr.plan(clips.shape[1:])
        for b in range(3):
            one = Fv.clip_transform(clips[b], frame_idx=idx, resize_hw=hw, window=win, mean=mean, std=std, div255=True,
                                    out_dtype=dt)                                       # table-driven kernel
            # same taps and weights bit for bit; the two kernels' blend expressions may contract into different FMAs
            assert torch.allclose(batch[b].float(), one.float(), rtol=0, atol=2e-6 if dt == torch.float32 else 2e-3)
        ref = O.val_chain(clips[1].cpu().numpy(), 32, mean, std, 48, 40)
        tol = dict(rtol=1e-5, atol=2e-6) if dt == torch.float32 else dict(rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(batch[1].float().cpu().numpy(), ref, **tol)
        # SlowFast packing from the same pass
        slow, fast = FusedClipTransform(32, mean, std, short_side=48, crop=("center", 40), out_dtype=dt, slowfast_alpha=4)(clips)
        assert torch.equal(fast, batch)
        sidx = torch.linspace(0, 31, 8).long()
        assert sidx.tolist() == [0, 4, 8, 13, 17, 22, 26, 31]
        assert torch.equal(slow, torch.index_select(batch, 2, sidx.cuda()))
    # stand-alone pack module on an already transformed clip / batch
    s2, f2 = SlowFastPackPathway(4)(batch)
    assert torch.equal(f2, batch) and torch.equal(s2, torch.index_select(batch, 2, sidx.cuda()))
    # THWC-interleaved decoder frames, batched
    thwc = clips.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
    assert torch.equal(FusedClipTransform(32, mean, std, short_side=48, crop=("center", 40), out_dtype=torch.float16)(thwc), batch)


def test_batched_train_chain_draws_per_clip():
    """A batch through the train chain draws short side / crop / flip per clip, in clip order, from the global RNG:
    equal to transforming the clips one after the other under the same seed."""
    from pytorchvideo_b200.transforms import create_video_transform
    clips = torch.stack([TS.synthetic_u8_clip(12, 60, 80, seed=50 + i) for i in range(4)]).cuda()
    tr = create_video_transform(mode="train", num_samples=6, min_size=32, max_size=48, crop_size=24, out_dtype=torch.float32)
    torch.manual_seed(99)
    batch = tr(clips)
    torch.manual_seed(99)
    singles = torch.stack([tr(clips[b]) for b in range(4)])
    assert torch.equal(batch, singles)


def test_uniform_temporal_subsample_nd_and_uint8():
    """functional.py:19-41 / 134-160 on 4-D clips, 5-D batches (temporal_dim=2, tests/test_models_slowfast.py:142-144)
    and uint8 frames: pure index selection, bit-exact, dtype preserved."""
    from pytorchvideo_b200.transforms import functional as Fv
    g = torch.Generator().manual_seed(3)
    x5 = torch.rand(2, 3, 32, 20, 24, generator=g)
    slow, fast = Fv.uniform_temporal_subsample_repeated(x5.cuda(), (4, 1), temporal_dim=2)
    idx = torch.clamp(torch.linspace(0, 31, 8), 0, 31).long()
    assert torch.equal(slow.cpu(), torch.index_select(x5, 2, idx)) and torch.equal(fast.cpu(), x5)
    u8 = TS.synthetic_u8_clip(20, 18, 22, seed=4)
    out = Fv.uniform_temporal_subsample(u8.cuda(), 10)
    assert out.dtype == torch.uint8 and torch.equal(out.cpu(), u8[:, O.linspace_indices(20, 10)])
    x8 = torch.rand(8, 9, 10, 12, generator=g)            # more than 4 leading entries: treated as 8 one-channel clips
    assert torch.equal(Fv.uniform_temporal_subsample(x8.cuda(), 4).cpu(), x8[:, O.linspace_indices(9, 4)])
    h16 = torch.rand(3, 7, 8, 8, generator=g).half()
    assert torch.equal(Fv.uniform_temporal_subsample(h16.cuda(), 3, temporal_dim=1).cpu(), h16[:, O.linspace_indices(7, 3)])
    with pytest.raises(NotImplementedError):
        Fv.uniform_temporal_subsample(x5.cuda(), 4, temporal_dim=1)


def test_multiview_views_and_ensemble():
    """f2: 3 spatial x K temporal views of one video from ONE transform launch == the reference pipeline view by view
    (clip slice -> subsample -> /255 -> normalize -> short side scale -> uniform_crop(i)), and the on-device
    reduction of the per-view predictions == the host-side accumulation of video_classification.py:290-311."""
    from pytorchvideo_b200.multiview import MultiViewEnsemble, clip_start_frames, view_reduce
    import pytorchvideo_b200.models.hub as PH
    video = TS.synthetic_u8_clip(50, 48, 72, seed=8)
    assert clip_start_frames(50, 16, 4) == [0, 11, 22, 34]              # 34 * i / 3 floored (clip_sampling.py:375-379)
    mv = MultiViewEnsemble(None, clip_frames=16, num_samples=4, clips_per_video=4, crops=3, side_size=32, crop_size=32,
                           out_dtype=torch.float32)
    views = mv.make_views(video.cuda()).cpu()
    assert views.shape == (12, 3, 4, 32, 32)
    mean, std = (0.45,) * 3, (0.225,) * 3
    v = 0
    for start in clip_start_frames(50, 16, 4):
        clip = video[:, start:start + 16].numpy()
        x = O.short_side_scale(O.normalize(O.div_255(O.uniform_temporal_subsample(clip, 4)), mean, std), 32)
        for s in range(3):
            y, xo, h, w = O.uniform_crop_window(x.shape[2], x.shape[3], 32, s)
            np.testing.assert_allclose(views[v].numpy(), x[:, :, y:y + h, xo:xo + w], rtol=1e-5, atol=2e-6)
            v += 1
    preds = torch.rand(2 * 12, 40, generator=torch.Generator().manual_seed(2))
    for mode, ref in (("sum", preds.view(2, 12, 40).sum(1)), ("mean", preds.view(2, 12, 40).sum(1) / 12),
                      ("max", preds.view(2, 12, 40).max(1).values)):
        assert torch.allclose(view_reduce(preds.cuda(), 12, mode).cpu(), ref, rtol=1e-6, atol=1e-6)
    # end to end on a small model: ensembled prediction == mean of the per-view model outputs
    model = TS.randomize_model(PH.x3d_xs(), seed=4).eval().cuda()
    vid = TS.synthetic_u8_clip(24, 170, 200, seed=9).cuda()
    mv = MultiViewEnsemble(model, clip_frames=8, num_samples=4, clips_per_video=3, crops=3, side_size=160, crop_size=160,
                           ensemble="mean")
    out = mv(vid).cpu()
    per_view = model(mv.make_views(vid)).float().cpu()
    # (two separate model runs: X3D's squeeze-excitation sums use fp32 atomics, so they agree to rounding only)
    assert out.shape == (400,) and torch.allclose(out, per_view.mean(0), rtol=2e-3, atol=2e-3 * float(per_view.abs().max()))
