"""Heads (reference models/head.py)."""
import torch.nn as nn

from ..layers.utils import set_attributes
from ..module import B200Module


class ResNetBasicHead(B200Module):
    """pool -> dropout -> Linear projection -> activation -> global average (head.py:330-391).
    The activation (softmax over classes) is applied per position BEFORE the average."""

    def __init__(self, pool=None, dropout=None, proj=None, activation=None, output_pool=None):
        super().__init__()
        set_attributes(self, locals())
        assert self.proj is not None


def _head_activation(activation):
    if activation is None:
        return None
    if activation == nn.Softmax:
        return activation(dim=1)
    return activation()


def create_res_basic_head(*, in_features, out_features, pool=nn.AvgPool3d, output_size=(1, 1, 1),
                          pool_kernel_size=(1, 7, 7), pool_stride=(1, 1, 1), pool_padding=(0, 0, 0),
                          dropout_rate=0.5, activation=None, output_with_global_average=True):
    if pool is None:
        pool_model = None
    elif pool == nn.AdaptiveAvgPool3d:
        pool_model = pool(output_size)
    else:
        pool_model = pool(kernel_size=pool_kernel_size, stride=pool_stride, padding=pool_padding)
    return ResNetBasicHead(
        proj=nn.Linear(in_features, out_features),
        activation=_head_activation(activation),
        pool=pool_model,
        dropout=nn.Dropout(dropout_rate) if dropout_rate > 0 else None,
        output_pool=nn.AdaptiveAvgPool3d(1) if output_with_global_average else None,
    )


class RoIAlign(nn.Module):
    """Parameter-free stand-in for ``torchvision.ops.RoIAlign`` (the reference's default ``roi`` callable,
    head.py:209): same constructor, attributes and repr.  The lowering reads the attributes by name, so the
    torchvision module itself is accepted as well; the sampling runs in ``pv_roi_align_fwd``."""

    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=False):
        super().__init__()
        self.output_size = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.aligned = bool(aligned)

    def extra_repr(self):
        return "output_size=%s, spatial_scale=%s, sampling_ratio=%d, aligned=%s" % (
            self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned)

    def forward(self, x, boxes):
        raise RuntimeError("RoIAlign runs inside a compiled detection head (pytorchvideo_b200 has no eager ATen path)")


class ResNetRoIHead(B200Module):
    """pool -> RoIAlign(x, bboxes) -> 2-D spatial pool -> dropout -> Linear -> activation -> optional global
    average (reference head.py:394-482).  ``forward(x, bboxes)``: bboxes is a float tensor [K, 5] =
    (batch index, x1, y1, x2, y2)."""

    def __init__(self, pool=None, pool_spatial=None, roi_layer=None, dropout=None, proj=None, activation=None,
                 output_pool=None):
        super().__init__()
        set_attributes(self, locals())
        assert self.proj is not None

    def forward(self, x, bboxes):
        ins = (list(x) if isinstance(x, (list, tuple)) else [x]) + [bboxes]
        return self._pv_compiled(ins)(ins).clone()


def create_res_roi_pooling_head(*, in_features, out_features, resolution, spatial_scale, sampling_ratio=0,
                                roi=RoIAlign, pool=nn.AvgPool3d, output_size=(1, 1, 1), pool_kernel_size=(1, 7, 7),
                                pool_stride=(1, 1, 1), pool_padding=(0, 0, 0), pool_spatial=nn.MaxPool2d,
                                dropout_rate=0.5, activation=None, output_with_global_average=True):
    """Reference head.py:199-327 (same arguments and module attribute names)."""
    if pool is None:
        pool_model = None
    elif pool == nn.AdaptiveAvgPool3d:
        pool_model = pool(output_size)
    else:
        pool_model = pool(kernel_size=pool_kernel_size, stride=pool_stride, padding=pool_padding)
    return ResNetRoIHead(
        proj=nn.Linear(in_features, out_features),
        activation=_head_activation(activation),
        pool=pool_model,
        pool_spatial=pool_spatial(resolution, stride=1) if pool_spatial else None,
        roi_layer=roi(output_size=resolution, spatial_scale=spatial_scale, sampling_ratio=sampling_ratio),
        dropout=nn.Dropout(dropout_rate) if dropout_rate > 0 else None,
        output_pool=nn.AdaptiveAvgPool3d(1) if output_with_global_average else None,
    )


class SequencePool(B200Module):
    """'cls': first token, 'mean': token average (head.py:11-36)."""

    def __init__(self, mode):
        super().__init__()
        assert mode in ["cls", "mean"], "Unsupported mode for SequencePool."
        self.mode = mode


class VisionTransformerBasicHead(B200Module):
    """sequence_pool -> dropout -> Linear -> activation (head.py:485-535)."""

    def __init__(self, sequence_pool=None, dropout=None, proj=None, activation=None):
        super().__init__()
        set_attributes(self, locals())
        assert self.proj is not None


def create_vit_basic_head(*, in_features, out_features, seq_pool_type="cls", dropout_rate=0.5, activation=None):
    assert seq_pool_type in ["cls", "mean", "none"]
    if seq_pool_type in ("cls", "mean"):
        pool = SequencePool(seq_pool_type)
    else:
        pool = None
    return VisionTransformerBasicHead(sequence_pool=pool, dropout=nn.Dropout(dropout_rate) if dropout_rate > 0.0 else None,
                                      proj=nn.Linear(in_features, out_features), activation=_head_activation(activation))
