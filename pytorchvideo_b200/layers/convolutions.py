"""(2+1)D factored convolution containers (reference layers/convolutions.py:88-237)."""
import torch.nn as nn

from ..module import B200Module
from .utils import set_attributes


class Conv2plus1d(B200Module):
    """conv_t -> norm -> activation -> conv_xy, in that order (convolutions.py:232-237).

    R(2+1)D uses a dense temporal then a dense spatial convolution; the X3D stem passes the
    spatial conv as ``conv_t`` and a depthwise temporal conv as ``conv_xy`` (models/x3d.py:83-88).
    """

    def __init__(self, *, conv_t=None, norm=None, activation=None, conv_xy=None, conv_xy_first=False):
        super().__init__()
        set_attributes(self, locals())
        assert self.conv_t is not None and self.conv_xy is not None
        if conv_xy_first:
            raise NotImplementedError("conv_xy_first ordering is not used by any in-scope model")


class ConvReduce3D(nn.Module):
    """API placeholder (reference convolutions.py:11-85, acoustic stem only - out of scope)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("ConvReduce3D is only used by the acoustic stem (out of scope)")


def create_conv_2plus1d(*, in_channels, out_channels, inner_channels=None, conv_xy_first=False,
                        kernel_size=(3, 3, 3), stride=(2, 2, 2), padding=(1, 1, 1), bias=False,
                        dilation=(1, 1, 1), groups=1, norm=nn.BatchNorm3d, norm_eps=1e-5,
                        norm_momentum=0.1, activation=nn.ReLU):
    """Factor a k_t x k_h x k_w convolution into (k_t,1,1) then (1,k_h,k_w) with BN+act between."""
    mid = out_channels if inner_channels is None else inner_channels
    temporal = nn.Conv3d(in_channels, mid, kernel_size=(kernel_size[0], 1, 1), stride=(stride[0], 1, 1),
                         padding=(padding[0], 0, 0), bias=bias, groups=groups, dilation=(dilation[0], 1, 1))
    spatial = nn.Conv3d(mid, out_channels, kernel_size=(1, kernel_size[1], kernel_size[2]),
                        stride=(1, stride[1], stride[2]), padding=(0, padding[1], padding[2]), bias=bias,
                        groups=groups, dilation=(1, dilation[1], dilation[2]))
    return Conv2plus1d(
        conv_t=temporal,
        norm=None if norm is None else norm(num_features=mid, eps=norm_eps, momentum=norm_momentum),
        activation=None if activation is None else activation(),
        conv_xy=spatial,
        conv_xy_first=conv_xy_first,
    )
