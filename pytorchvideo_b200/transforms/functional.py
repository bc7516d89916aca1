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


_IDX_CACHE = {}


def _dev_i32(values, dev):
    key = (dev.index, tuple(int(v) for v in values))
    t = _IDX_CACHE.get(key)
    if t is None:
        if len(_IDX_CACHE) > 512:
            _IDX_CACHE.clear()
        t = _IDX_CACHE[key] = torch.tensor(list(key[1]), dtype=torch.int32, device=dev)
    return t


def slow_pathway_indices(n_frames, alpha):
    """SlowFastPackPathway (pytorchvideo_trainer/datamodule/transforms.py:129-136):
    torch.linspace(0, T - 1, T // alpha).long() - positions inside the (already subsampled) fast clip."""
    return torch.linspace(0, n_frames - 1, n_frames // alpha).long()


def clip_transform_batch(x, frame_idx=None, resize_hw=None, window=None, mean=None, std=None, div255=False,
                         out_dtype=torch.float16, hflip=False, geom=None, slow_alpha=None, out=None, out_slow=None):
    """The fused chain on a BATCH of clips in one launch (pv_clip_transform_batch; taps computed in the kernel).

    x         : (B, C, T, H, W) CUDA tensor, C <= 4, any strides (CTHW-contiguous or the decoder's THWC-interleaved
                frames); a 4-D clip is a batch of one
    geom      : optional per-clip list of (resize_hw, window, hflip[, first_frame]) - the train chain with its own
                random short side / crop / flip per clip, or the spatial x temporal test-time views of one video
                (expand the video to a batch with clip stride 0); resize_hw / window then only give the output size
    slow_alpha: also emit the SlowFast slow pathway (frames linspace(0, n_t-1, n_t//alpha).long() of the kept
                frames) from the same pass; returns [slow, fast] like SlowFastPackPathway
    out_dtype : torch.float16 | torch.float32, or torch.uint8 for a pure frame selection / crop of uint8 clips
    """
    squeeze = False
    if torch.is_tensor(x) and x.dim() == 4:
        x, squeeze = x.unsqueeze(0), True
        out = out.unsqueeze(0) if out is not None and out.dim() == 4 else out
        out_slow = out_slow.unsqueeze(0) if out_slow is not None and out_slow.dim() == 4 else out_slow
    if not torch.is_tensor(x) or x.dim() != 5:
        raise RuntimeError("expected a (B, C, T, H, W) or (C, T, H, W) tensor")
    if x.device.type != "cuda":
        raise RuntimeError("pytorchvideo_b200 transforms run on the GPU only (no CPU path)")
    if x.dtype not in _DT:
        raise RuntimeError("unsupported clip dtype %s" % x.dtype)
    B, Cc, T, H, W = x.shape
    if Cc > 4:
        raise RuntimeError("at most 4 channels per clip (got %d)" % Cc)
    lib = L.load()
    dev = x.device
    idx = torch.arange(T) if frame_idx is None else torch.as_tensor(frame_idx).long().cpu()
    if idx.numel() == 0 or int(idx.min()) < 0 or int(idx.max()) >= T:
        raise RuntimeError("frame index out of range")
    n_t = int(idx.numel())
    nh, nw = (H, W) if resize_hw is None else (int(resize_hw[0]), int(resize_hw[1]))
    top, left, oh, ow = (0, 0, nh, nw) if window is None else (int(v) for v in window)
    geom_d = None
    if geom is not None:
        if len(geom) != B:
            raise RuntimeError("geom needs one entry per clip")
        flat = []
        for entry in geom:
            ghw, gwin, gflip = entry[:3]
            goff = int(entry[3]) if len(entry) > 3 else 0
            if goff + int(idx.min()) < 0 or goff + int(idx.max()) >= T:
                raise RuntimeError("view frame range outside the clip")
            gh, gw = (H, W) if ghw is None else ghw
            gt, gl, goh, gow = (0, 0, gh, gw) if gwin is None else gwin
            if (goh, gow) != (oh, ow) or gt < 0 or gl < 0 or gt + goh > gh or gl + gow > gw:
                raise RuntimeError("per-clip crop windows must have the common output size and lie inside the resized frame")
            flat += [int(gh), int(gw), int(gt), int(gl), 1 if gflip else 0, goff]
        geom_d = torch.tensor(flat, dtype=torch.int32, device=dev)
    elif top < 0 or left < 0 or top + oh > nh or left + ow > nw:
        raise RuntimeError("crop window outside the frame")
    d = L.ClipBatchDesc()
    d.C, d.n_clips, d.n_t = Cc, B, n_t
    d.in_h, d.in_w, d.new_h, d.new_w = H, W, nh, nw
    d.top, d.left, d.out_h, d.out_w, d.hflip = top, left, oh, ow, 1 if hflip else 0
    d.s_clip, d.sc, d.st, d.sh, d.sw = x.stride(0), x.stride(1), x.stride(2), x.stride(3), x.stride(4)
    normalize = mean is not None or std is not None
    mean_l = [float(m) for m in (mean if mean is not None else [0.0] * Cc)]
    std_l = [float(v) for v in (std if std is not None else [1.0] * Cc)]
    mean_l = mean_l * Cc if len(mean_l) == 1 else mean_l
    std_l = std_l * Cc if len(std_l) == 1 else std_l
    for i in range(4):
        d.mean[i] = mean_l[i] if i < len(mean_l) else 0.0
        d.stdv[i] = std_l[i] if i < len(std_l) else 1.0
    d.div255, d.normalize = 1 if div255 else 0, 1 if normalize else 0
    d.src_dtype = _DT[x.dtype]
    if out_dtype not in _DT:
        raise RuntimeError("unsupported output dtype %s" % out_dtype)
    d.dst_dtype = _DT[out_dtype]
    if out is None:
        out = torch.empty((B, Cc, n_t, oh, ow), dtype=out_dtype, device=dev)
    elif tuple(out.shape) != (B, Cc, n_t, oh, ow) or not out.is_contiguous() or out.dtype != out_dtype:
        raise RuntimeError("out must be a contiguous %s tensor of shape %s" % (out_dtype, (B, Cc, n_t, oh, ow)))
    d.d_clip = out.stride(0)
    idx_d = _dev_i32(idx.tolist(), dev)
    slow_d, n_slow = None, 0
    if slow_alpha is not None:
        sidx = slow_pathway_indices(n_t, int(slow_alpha)).tolist()
        n_slow = len(sidx)
        if n_slow < 1:
            raise RuntimeError("slow pathway would be empty (n_t=%d, alpha=%d)" % (n_t, slow_alpha))
        pos = [-1] * n_t
        for k, j in enumerate(sidx):
            pos[j] = k           # linspace().long() is strictly increasing for alpha >= 1: every slot is unique
        if len(set(sidx)) != n_slow:
            raise RuntimeError("slow pathway indices repeat (alpha < 1?)")
        slow_d = _dev_i32(pos, dev)
        if out_slow is None:
            out_slow = torch.empty((B, Cc, n_slow, oh, ow), dtype=out_dtype, device=dev)
        elif tuple(out_slow.shape) != (B, Cc, n_slow, oh, ow) or not out_slow.is_contiguous() or out_slow.dtype != out_dtype:
            raise RuntimeError("out_slow has the wrong shape / dtype")
        d.n_slow, d.d_slow_clip = n_slow, out_slow.stride(0)
    stream = torch.cuda.current_stream(dev).cuda_stream
    L.check(lib.pv_clip_transform_batch(C.byref(d), x.data_ptr(), idx_d.data_ptr(),
                                        slow_d.data_ptr() if slow_d is not None else None,
                                        geom_d.data_ptr() if geom_d is not None else None, out.data_ptr(),
                                        out_slow.data_ptr() if out_slow is not None else None, stream),
            "pv_clip_transform_batch")
    out._pv_keepalive = (idx_d, slow_d, geom_d)     # device tables must outlive the asynchronous launch
    if squeeze:
        out = out[0]
        out_slow = out_slow[0] if out_slow is not None else None
    return [out_slow, out] if slow_alpha is not None else out


# ---- reference-named functional API -----------------------------------------------------------
def _as_clip_batch(x, temporal_dim):
    """View an N-D tensor with its temporal dim at -3 as (clips, channels <= 4, T, H, W) without copying."""
    nd = x.dim()
    if nd < 3:
        raise RuntimeError("uniform_temporal_subsample needs at least (T, H, W)")
    td = temporal_dim % nd
    if td != nd - 3:
        raise NotImplementedError("only temporal_dim == -3 (…, T, H, W) has a B200 kernel; got dim %d of %d" % (temporal_dim, nd))
    lead = x.shape[:td]
    v = x.reshape((-1,) + tuple(x.shape[td:])) if nd != 4 else x      # (L, T, H, W); a view for contiguous leading dims
    Ld = v.shape[0]
    if Ld <= 4:
        return v.unsqueeze(0), lead                  # one clip of L channels
    return v.unsqueeze(1), lead                      # L clips of one channel


def uniform_temporal_subsample(x, num_samples, temporal_dim=-3):
    """functional.py:19-41 on any (…, T, H, W) tensor (4-D clips, 5-D batches as used with temporal_dim=2 by
    tests/test_models_slowfast.py:142-144): frame selection only, dtype preserved (uint8 stays uint8)."""
    if not torch.is_tensor(x) or x.device.type != "cuda":
        raise RuntimeError("pytorchvideo_b200 transforms run on the GPU only (no CPU path)")
    if x.dtype not in _DT:
        raise RuntimeError("unsupported clip dtype %s" % x.dtype)
    v, lead = _as_clip_batch(x, temporal_dim)
    idx = temporal_indices(v.shape[2], num_samples)
    out = clip_transform_batch(v, frame_idx=idx, out_dtype=x.dtype)
    return out.reshape(tuple(lead) + (int(idx.numel()),) + tuple(x.shape[-2:]))


def uniform_temporal_subsample_repeated(frames, frame_ratios, temporal_dim=-3):
    """functional.py:134-160: one subsampled tensor per ratio (SlowFast: frame_ratios=(alpha, 1))."""
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
