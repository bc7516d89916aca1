"""nn.Module wrappers with the reference's names (transforms/transforms.py) plus the fused chain.

``FusedClipTransform`` is the product: UniformTemporalSubsample -> /255 -> Normalize ->
ShortSideScale -> crop as ONE kernel launch reading uint8 frames (CTHW or the decoder's THWC
view) and writing the f16/f32 network input.  The single-op modules run the same kernel with
identity settings so that a torchvision ``Compose`` of them still works (one launch per op).
"""
import torch
import torch.nn as nn

from . import functional as Fv


class ApplyTransformToKey:
    def __init__(self, key, transform):
        self._key = key
        self._transform = transform

    def __call__(self, x):
        x[self._key] = self._transform(x[self._key])
        return x


class UniformTemporalSubsample(nn.Module):
    def __init__(self, num_samples, temporal_dim=-3):
        super().__init__()
        self._num_samples = num_samples
        self._temporal_dim = temporal_dim

    def forward(self, x):
        return Fv.uniform_temporal_subsample(x, self._num_samples, self._temporal_dim)


class ShortSideScale(nn.Module):
    def __init__(self, size, interpolation="bilinear", backend="pytorch"):
        super().__init__()
        self._size, self._interpolation, self._backend = size, interpolation, backend

    def forward(self, x):
        return Fv.short_side_scale(x, self._size, self._interpolation, self._backend)


class RandomShortSideScale(nn.Module):
    def __init__(self, min_size, max_size, interpolation="bilinear", backend="pytorch"):
        super().__init__()
        self._min_size, self._max_size = min_size, max_size
        self._interpolation, self._backend = interpolation, backend

    def forward(self, x):
        size = torch.randint(self._min_size, self._max_size + 1, (1,)).item()
        return Fv.short_side_scale(x, size, self._interpolation, self._backend)


class Normalize(nn.Module):
    """(x - mean[c]) / std[c] on a CTHW clip (reference transforms.py:177-195)."""

    def __init__(self, mean, std, inplace=False):
        super().__init__()
        self.mean, self.std, self.inplace = list(mean), list(std), inplace

    def forward(self, x):
        if not x.is_floating_point():
            raise TypeError("Input tensor should be a float tensor. Got %s." % x.dtype)
        return Fv.clip_transform(x, mean=self.mean, std=self.std, out_dtype=x.dtype)


class Div255(nn.Module):
    def forward(self, x):
        return Fv.div_255(x)


class ConvertUint8ToFloat(nn.Module):
    def forward(self, x):
        assert x.dtype == torch.uint8, "image must have dtype torch.uint8"
        return Fv.clip_transform(x, div255=True)


class CenterCropVideo(nn.Module):
    """torchvision CenterCrop semantics on the last two dims (a view, no kernel)."""

    def __init__(self, size):
        super().__init__()
        self.size = size

    def forward(self, x):
        top, left, h, w = Fv.center_crop_window(x.shape[-2], x.shape[-1], self.size)
        return x[..., top:top + h, left:left + w]


class RandomCropVideo(nn.Module):
    """torchvision RandomCrop semantics (offsets from torch's global RNG on the host)."""

    def __init__(self, size):
        super().__init__()
        self.size = size

    def forward(self, x):
        top, left, h, w = Fv.random_crop_window(x.shape[-2], x.shape[-1], self.size)
        return x[..., top:top + h, left:left + w]


class UniformCropVideo(nn.Module):
    def __init__(self, size, video_key="video", aug_index_key="aug_index"):
        super().__init__()
        self._size, self._video_key, self._aug_index_key = size, video_key, aug_index_key

    def __call__(self, x):
        x[self._video_key] = Fv.uniform_crop(x[self._video_key], self._size, x[self._aug_index_key])
        return x


class FusedClipTransform(nn.Module):
    """One-kernel eval/train chain.  crop: None | ("center", size) | ("random", size) |
    ("uniform", size, spatial_idx).  Input: uint8 (or float) CUDA clip (C, T, H, W)."""

    def __init__(self, num_samples=None, mean=None, std=None, short_side=None, crop=None, div255=True,
                 out_dtype=torch.float16, random_short_side=None):
        super().__init__()
        self.num_samples, self.mean, self.std = num_samples, mean, std
        self.short_side, self.crop, self.div255, self.out_dtype = short_side, crop, div255, out_dtype
        self.random_short_side = random_short_side

    def plan(self, shape):
        """Host-side index/window selection for an input of ``shape`` (C, T, H, W)."""
        _, T, H, W = shape
        idx = None if self.num_samples is None else Fv.temporal_indices(T, self.num_samples)
        side = self.short_side
        if self.random_short_side is not None:
            lo, hi = self.random_short_side
            side = torch.randint(lo, hi + 1, (1,)).item()
        hw = None if side is None else Fv.short_side_size(H, W, side)
        nh, nw = (H, W) if hw is None else hw
        win = None
        if self.crop is not None:
            kind = self.crop[0]
            if kind == "center":
                win = Fv.center_crop_window(nh, nw, self.crop[1])
            elif kind == "random":
                win = Fv.random_crop_window(nh, nw, self.crop[1])
            elif kind == "uniform":
                win = Fv.uniform_crop_window(nh, nw, self.crop[1], self.crop[2])
            else:
                raise ValueError("unknown crop kind %r" % (kind,))
        return idx, hw, win

    def forward(self, x, out=None):
        idx, hw, win = self.plan(x.shape)
        return Fv.clip_transform(x, frame_idx=idx, resize_hw=hw, window=win, mean=self.mean, std=self.std,
                                 div255=self.div255, out_dtype=self.out_dtype, out=out)


def create_video_transform(mode="val", num_samples=8, video_mean=(0.45, 0.45, 0.45),
                           video_std=(0.225, 0.225, 0.225), min_size=256, max_size=320, crop_size=224,
                           convert_to_float=True, out_dtype=torch.float16):
    """Fused equivalent of the reference factory's default chains (transforms_factory.py:109-261):
    train = subsample, /255, normalize, RandomShortSideScale(min,max), RandomCrop
    val   = subsample, /255, normalize, ShortSideScale(min_size), CenterCrop."""
    assert mode in ("train", "val")
    if mode == "val":
        return FusedClipTransform(num_samples, video_mean, video_std, short_side=min_size,
                                  crop=("center", crop_size), div255=convert_to_float, out_dtype=out_dtype)
    return FusedClipTransform(num_samples, video_mean, video_std, crop=("random", crop_size),
                              div255=convert_to_float, out_dtype=out_dtype,
                              random_short_side=(min_size, max_size))
