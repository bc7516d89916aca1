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
                 out_dtype=torch.float16, random_short_side=None, hflip_prob=0.0, slowfast_alpha=None):
        super().__init__()
        self.num_samples, self.mean, self.std = num_samples, mean, std
        self.short_side, self.crop, self.div255, self.out_dtype = short_side, crop, div255, out_dtype
        self.random_short_side = random_short_side
        self.hflip_prob = float(hflip_prob)
        # emit [slow, fast] (SlowFastPackPathway, pytorchvideo_trainer datamodule/transforms.py:99-138) from the same pass
        self.slowfast_alpha = slowfast_alpha

    def plan(self, shape):
        """Host-side index/window/flip selection for an input of ``shape`` (C, T, H, W).  Random draws
        come from torch's global RNG in the order the reference's Compose makes them:
        RandomShortSideScale (randint), RandomCrop (randint i, randint j), RandomHorizontalFlip (rand)."""
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
        flip = False
        if self.hflip_prob > 0.0:        # torchvision RandomHorizontalFlip.forward: torch.rand(1) < p
            flip = bool(torch.rand(1) < self.hflip_prob)
        return idx, hw, win, flip

    def forward(self, x, out=None):
        """x: one clip (C, T, H, W) or a batch (B, C, T, H, W) - ONE launch either way (a batch in train mode
        draws short side / crop / flip per clip, in clip order).  Returns the clip(s), or [slow, fast] when
        ``slowfast_alpha`` is set."""
        if x.dim() == 5 and self._is_random():
            plans = [self.plan(x.shape[1:]) for _ in range(x.shape[0])]
            idx, hw, win, _ = plans[0]
            geom = [(p[1], p[2], p[3]) for p in plans]
            return Fv.clip_transform_batch(x, frame_idx=idx, resize_hw=hw, window=win, mean=self.mean, std=self.std,
                                           div255=self.div255, out_dtype=self.out_dtype, geom=geom,
                                           slow_alpha=self.slowfast_alpha, out=out)
        idx, hw, win, flip = self.plan(x.shape[-4:])
        return Fv.clip_transform_batch(x, frame_idx=idx, resize_hw=hw, window=win, mean=self.mean, std=self.std,
                                       div255=self.div255, out_dtype=self.out_dtype, hflip=flip,
                                       slow_alpha=self.slowfast_alpha, out=out)

    def _is_random(self):
        return self.random_short_side is not None or self.hflip_prob > 0.0 or (self.crop is not None and self.crop[0] == "random")


class SlowFastPackPathway(nn.Module):
    """frames (C, T, H, W) or (B, C, T, H, W) -> [slow, fast]; slow = frames at linspace(0, T-1, T//alpha).long()
    (pytorchvideo_trainer/datamodule/transforms.py:99-138).  One gather launch; inside a FusedClipTransform the
    same list comes out of the transform kernel itself (``slowfast_alpha``)."""

    def __init__(self, alpha=4):
        super().__init__()
        self.alpha = alpha

    def forward(self, frames):
        out = Fv.clip_transform_batch(frames, out_dtype=frames.dtype, slow_alpha=self.alpha)
        return [out[0], frames]


class RemoveKey:
    """transforms.py:34-47: drops ``key`` from a dict sample."""

    def __init__(self, key):
        self._key = key

    def __call__(self, x):
        if self._key in x:
            del x[self._key]
        return x


class _Chain:
    """Minimal torchvision ``Compose`` (dict-level steps around the fused clip transform)."""

    def __init__(self, steps):
        self.transforms = list(steps)

    def __call__(self, x):
        for t in self.transforms:
            x = t(x)
        return x


def create_video_transform(mode, video_key=None, remove_key=None, num_samples=8, convert_to_float=True,
                           video_mean=(0.45, 0.45, 0.45), video_std=(0.225, 0.225, 0.225), min_size=256,
                           max_size=320, crop_size=224, horizontal_flip_prob=0.5, aug_type="default",
                           aug_paras=None, random_resized_crop_paras=None, out_dtype=torch.float16):
    """Fused equivalent of the reference factory's default chains (transforms_factory.py:109-284), same
    signature and argument checks:
      train = subsample, /255, normalize, RandomShortSideScale(min,max), RandomCrop, RandomHorizontalFlip(p)
      val   = subsample, /255, normalize, ShortSideScale(min_size), CenterCrop
    as ONE kernel launch.  RandAugment / AugMix (aug_type) and RandomResizedCrop have no B200 kernel:
    asking for them raises NotImplementedError instead of silently changing the augmentation."""
    if mode not in ("train", "val"):
        raise NotImplementedError("mode must be 'train' or 'val'")
    if isinstance(crop_size, int):
        assert crop_size <= min_size, "crop_size must be less than or equal to min_size"
    elif isinstance(crop_size, tuple):
        assert max(crop_size) <= min_size, "the height and width in crop_size must be less than or equal to min_size"
    else:
        raise TypeError
    if video_key is None:
        assert remove_key is None, "remove_key should be None if video_key is None"
    if aug_type == "default":
        assert aug_paras is None, "aug_paras should be None for ``default`` aug_type"
    elif aug_type in ("randaug", "augmix"):
        if mode == "train":
            raise NotImplementedError("aug_type=%r has no B200 kernel (only the 'default' chain is fused)" % aug_type)
    else:
        raise NotImplementedError
    if random_resized_crop_paras is not None and mode == "train":
        raise NotImplementedError("RandomResizedCrop has no B200 kernel (use RandomShortSideScale + RandomCrop)")
    if mode == "val":
        tr = FusedClipTransform(num_samples, video_mean, video_std, short_side=min_size,
                                crop=("center", crop_size), div255=convert_to_float, out_dtype=out_dtype)
    else:
        tr = FusedClipTransform(num_samples, video_mean, video_std, crop=("random", crop_size),
                                div255=convert_to_float, out_dtype=out_dtype,
                                random_short_side=(min_size, max_size), hflip_prob=horizontal_flip_prob)
    if video_key is None:
        return tr
    return _Chain([ApplyTransformToKey(key=video_key, transform=tr)] +
                  ([] if remove_key is None else [RemoveKey(k) for k in remove_key]))
