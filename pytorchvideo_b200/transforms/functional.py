"""Clip transforms executed by ONE fused CUDA kernel (pv_clip_transform_fwd).

Host side only computes small integer/weight tables (temporal indices, bilinear taps, crop
window) exactly as the reference's host code does; every pixel is touched on the GPU once.
Reference: transforms/functional.py:19-41, 92-160, 302-347, 604-615; transforms/transforms.py.
"""
import ctypes as C
import math

import numpy as np
import torch

from .. import _lib as L

_DT = {torch.uint8: L.PV_U8, torch.float32: L.PV_F32, torch.float16: L.PV_F16}
_TABLE_CACHE = {}


def temporal_indices(t, num_samples):
    """clamp(torch.linspace(0, t-1, n), 0, t-1).long() - literally the reference's host
    expression (functional.py:39-40), so indices are bit-identical to the reference's."""
    assert num_samples > 0 and t > 0
    idx = torch.clamp(torch.linspace(0, t - 1, num_samples), 0, t - 1).long()
    return idx


def short_side_size(h, w, size):
    if w < h:
        return int(math.floor((float(h) / w) * size)), size
    return size, int(math.floor((float(w) / h) * size))


def bilinear_table(in_size, out_size):
    """(i0, i1, lambda1) of ATen's upsample_bilinear2d(align_corners=False) in fp32."""
    if in_size == out_size:
        i = np.arange(out_size, dtype=np.int32)
        return i, i.copy(), np.zeros(out_size, np.float32)
    scale = np.float32(in_size) / np.float32(out_size)
    d = np.arange(out_size, dtype=np.float32) + np.float32(0.5)
    src = (np.float64(scale) * np.float64(d) - 0.5).astype(np.float32)   # fma(scale, d, -0.5)
    src = np.maximum(src, np.float32(0))
    i0 = np.minimum(np.floor(src).astype(np.int64), in_size - 1)
    l1 = np.clip(src - i0.astype(np.float32), np.float32(0), np.float32(1)).astype(np.float32)
    i1 = i0 + (i0 < in_size - 1)
    return i0.astype(np.int32), i1.astype(np.int32), l1


def center_crop_window(h, w, size):
    th, tw = (size, size) if isinstance(size, int) else size
    if th > h or tw > w:
        raise ValueError("crop size %s larger than image (%d, %d)" % ((th, tw), h, w))
    return int(round((h - th) / 2.0)), int(round((w - tw) / 2.0)), th, tw


def uniform_crop_window(h, w, size, spatial_idx):
    assert spatial_idx in (0, 1, 2)
    y = int(math.ceil((h - size) / 2))
    x = int(math.ceil((w - size) / 2))
    if h > w:
        y = 0 if spatial_idx == 0 else (h - size if spatial_idx == 2 else y)
    else:
        x = 0 if spatial_idx == 0 else (w - size if spatial_idx == 2 else x)
    return y, x, size, size


def random_crop_window(h, w, size):
    """torchvision RandomCrop.get_params: offsets drawn on the host from torch's global RNG."""
    th, tw = (size, size) if isinstance(size, int) else size
    if h < th or w < tw:
        raise ValueError("Required crop size %s is larger than input image size %s" % ((th, tw), (h, w)))
    if w == tw and h == th:
        return 0, 0, h, w
    i = torch.randint(0, h - th + 1, size=(1,)).item()
    j = torch.randint(0, w - tw + 1, size=(1,)).item()
    return i, j, th, tw


def clip_transform(x, frame_idx=None, resize_hw=None, window=None, mean=None, std=None, div255=False,
                   out_dtype=torch.float32, out=None, hflip=False):
    """Run the fused kernel on a CUDA clip ``x`` of logical shape (C, T, H, W) (any strides).

    frame_idx : int tensor/sequence of frames to keep (None = all)
    resize_hw : (new_h, new_w) bilinear target (None = no resize)
    window    : (top, left, h, w) crop in the resized frame (None = full)
    mean/std  : per-channel normalisation (None = skip)
    hflip     : mirror the (cropped) output along W - torchvision hflip fused for free by reversing the
                column tap tables on the host
    """
    if not torch.is_tensor(x) or x.dim() != 4:
        raise RuntimeError("expected a (C, T, H, W) tensor")
    if x.device.type != "cuda":
        raise RuntimeError("pytorchvideo_b200 transforms run on the GPU only (no CPU path)")
    if x.dtype not in _DT:
        raise RuntimeError("unsupported clip dtype %s" % x.dtype)
    lib = L.load()
    Cc, T, H, W = x.shape
    idx = torch.arange(T) if frame_idx is None else torch.as_tensor(frame_idx).long().cpu()
    if idx.numel() == 0 or int(idx.min()) < 0 or int(idx.max()) >= T:
        raise RuntimeError("frame index out of range")
    nh, nw = (H, W) if resize_hw is None else resize_hw
    top, left, oh, ow = (0, 0, nh, nw) if window is None else window
    if top < 0 or left < 0 or top + oh > nh or left + ow > nw:
        raise RuntimeError("crop window outside the frame")
    y0, y1, ly = bilinear_table(H, nh)
    x0, x1, lx = bilinear_table(W, nw)
    y0, y1, ly = y0[top:top + oh], y1[top:top + oh], ly[top:top + oh]
    x0, x1, lx = x0[left:left + ow], x1[left:left + ow], lx[left:left + ow]
    if hflip:
        x0, x1, lx = x0[::-1], x1[::-1], lx[::-1]
    per_channel = mean is not None or std is not None
    n_t = int(idx.numel())
    if per_channel:
        if Cc > 4:
            raise RuntimeError("per-channel normalisation supports at most 4 channels")
        mean_l = [float(m) for m in (mean if mean is not None else [0.0] * Cc)]
        std_l = [float(s) for s in (std if std is not None else [1.0] * Cc)]
        if len(mean_l) == 1:
            mean_l = mean_l * Cc
        if len(std_l) == 1:
            std_l = std_l * Cc
        kC, kidx, ksc = Cc, idx, x.stride(0)
    else:
        # no per-channel state: fold channels into the frame list so any C works
        mean_l, std_l = [0.0], [1.0]
        kC = 1
        kidx = (torch.arange(Cc).view(-1, 1) * 0 + idx.view(1, -1)).reshape(-1)
        # address = c*stride_c + t*stride_t is not expressible with one stride unless we bake the
        # channel offset into the frame index; do it when stride_c is a multiple of stride_t,
        # otherwise run channel by channel.
        ksc = 0
        if Cc > 1:
            st_c, st_t = x.stride(0), x.stride(1)
            if st_t != 0 and st_c % st_t == 0:
                kidx = (torch.arange(Cc).view(-1, 1) * (st_c // st_t) + idx.view(1, -1)).reshape(-1)
            else:
                outs = [clip_transform(x[c:c + 1], frame_idx, resize_hw, window, None, None, div255, out_dtype)
                        for c in range(Cc)]
                return torch.cat(outs, 0)
    dev = x.device
    # device-side tables are tiny but cost several H2D copies: cache them per (geometry, device)
    ckey = (dev.index, H, W, nh, nw, top, left, oh, ow, bool(hflip), tuple(int(i) for i in kidx.tolist()))
    cached = _TABLE_CACHE.get(ckey)
    if cached is None:
        tabs = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (y0, y1, ly, x0, x1, lx)]
        idx_d = kidx.to(torch.int32).to(dev)
        if len(_TABLE_CACHE) > 256:
            _TABLE_CACHE.clear()
        _TABLE_CACHE[ckey] = (tabs, idx_d)
    else:
        tabs, idx_d = cached
    if out is None:
        out = torch.empty((Cc, n_t, oh, ow), dtype=out_dtype, device=dev)
    else:
        assert out.shape == (Cc, n_t, oh, ow) and out.is_contiguous() and out.dtype == out_dtype
    d = L.ClipTransformDesc()
    d.C, d.n_t, d.out_h, d.out_w = kC, int(idx_d.numel()), oh, ow
    d.sc, d.st, d.sh, d.sw = ksc, x.stride(1), x.stride(2), x.stride(3)
    for i in range(4):
        d.mean[i] = mean_l[i] if i < len(mean_l) else 0.0
        d.stdv[i] = std_l[i] if i < len(std_l) else 1.0
    d.src_dtype = _DT[x.dtype]
    d.dst_dtype = L.PV_F16 if out_dtype == torch.float16 else L.PV_F32
    d.div255 = 1 if div255 else 0
    stream = torch.cuda.current_stream(dev).cuda_stream
    L.check(lib.pv_clip_transform_fwd(C.byref(d), x.data_ptr(), idx_d.data_ptr(), tabs[0].data_ptr(),
                                      tabs[1].data_ptr(), tabs[2].data_ptr(), tabs[3].data_ptr(),
                                      tabs[4].data_ptr(), tabs[5].data_ptr(), out.data_ptr(), stream),
            "pv_clip_transform_fwd")
    # tables must outlive the asynchronous launch
    out._pv_keepalive = (tabs, idx_d)
    return out


# ---- reference-named functional API -----------------------------------------------------------
def uniform_temporal_subsample(x, num_samples, temporal_dim=-3):
    if x.dim() != 4 or temporal_dim not in (-3, 1):
        raise RuntimeError("uniform_temporal_subsample: (C, T, H, W) clips with temporal_dim=-3 only")
    idx = temporal_indices(x.shape[1], num_samples)
    odt = x.dtype if x.dtype in (torch.float32, torch.float16) else torch.float32
    if x.dtype == torch.uint8:
        raise RuntimeError("uint8 clips keep their dtype only inside the fused chain; use "
                           "FusedClipTransform / create_video_transform for uint8 input")
    return clip_transform(x, frame_idx=idx, out_dtype=odt)


def uniform_temporal_subsample_repeated(frames, frame_ratios, temporal_dim=-3):
    t = frames.shape[temporal_dim]
    return [uniform_temporal_subsample(frames, t // r, temporal_dim) for r in frame_ratios]


def short_side_scale(x, size, interpolation="bilinear", backend="pytorch"):
    assert len(x.shape) == 4
    assert x.dtype == torch.float32
    assert backend in ("pytorch", "opencv")
    if interpolation != "bilinear" or backend != "pytorch":
        raise NotImplementedError("only bilinear / pytorch semantics have a B200 kernel")
    return clip_transform(x, resize_hw=short_side_size(x.shape[2], x.shape[3], size))


def div_255(x):
    return clip_transform(x, div255=True)


def uniform_crop(images, size, spatial_idx):
    y, xo, h, w = uniform_crop_window(images.shape[2], images.shape[3], size, spatial_idx)
    return images[:, :, y:y + h, xo:xo + w]   # a view, exactly like the reference's slicing
