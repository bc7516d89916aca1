"""(2+1)D factored convolution containers (reference layers/convolutions.py:88-237)."""
import torch.nn as nn

from ..module import B200Module
from .utils import set_attributes


class Conv2plus1d(B200Module):
    """conv_t -> norm -> activation -> conv_xy (convolutions.py:232-237); ``conv_xy_first`` swaps the two
    convolutions (norm / activation stay in between).

    R(2+1)D uses a dense temporal then a dense spatial convolution; the X3D stem passes the
    spatial conv as ``conv_t`` and a depthwise temporal conv as ``conv_xy`` (models/x3d.py:83-88).
    """

    def __init__(self, *, conv_t=None, norm=None, activation=None, conv_xy=None, conv_xy_first=False):
        super().__init__()
        set_attributes(self, locals())
        assert self.conv_t is not None and self.conv_xy is not None


class ConvReduce3D(B200Module):
    """Several Conv3d branches over ONE input, reduced by "sum" or "cat" (reference convolutions.py:11-85;
    used by the acoustic ResNet stem, stem.py:179-192).  Every per-branch option is a tuple indexed like
    ``kernel_size``; ``None`` (for the tuple or an entry) keeps nn.Conv3d's default.  On device the sum is
    fused: branch i+1 adds the running sum in its epilogue; "cat" writes channel slices of one buffer."""

    _PER_BRANCH = ("stride", "padding", "padding_mode", "dilation", "groups", "bias")

    def __init__(self, *, in_channels, out_channels, kernel_size, stride=None, padding=None, padding_mode=None,
                 dilation=None, groups=None, bias=None, reduction_method="sum"):
        super().__init__()
        assert reduction_method in ("sum", "cat")
        self.reduction_method = reduction_method
        given = dict(stride=stride, padding=padding, padding_mode=padding_mode, dilation=dilation, groups=groups,
                     bias=bias)
        branches = []
        for i, k in enumerate(kernel_size):
            opts = {name: given[name][i] for name in self._PER_BRANCH
                    if given[name] is not None and given[name][i] is not None}
            branches.append(nn.Conv3d(in_channels, out_channels, k, **opts))
        self.convs = nn.ModuleList(branches)


def create_conv_2plus1d(*, in_channels, out_channels, inner_channels=None, conv_xy_first=False,
                        kernel_size=(3, 3, 3), stride=(2, 2, 2), padding=(1, 1, 1), bias=False,
                        dilation=(1, 1, 1), groups=1, norm=nn.BatchNorm3d, norm_eps=1e-5,
                        norm_momentum=0.1, activation=nn.ReLU):
    """Factor a k_t x k_h x k_w convolution into (k_t,1,1) and (1,k_h,k_w) with BN+act between (reference
    convolutions.py:88-188).  Whichever convolution runs first maps in_channels -> inner_channels, the second
    inner_channels -> out_channels; ``conv_xy_first`` decides which one that is."""
    mid = out_channels if inner_channels is None else inner_channels
    assert groups == 1, "Support for groups is not implemented in R2+1 convolution layer"
    assert max(dilation) == 1 and min(dilation) == 1, "Support for dillaiton is not implemented in R2+1 convolution layer"
    t_io = (mid, out_channels) if conv_xy_first else (in_channels, mid)
    xy_io = (in_channels, mid) if conv_xy_first else (mid, out_channels)
    temporal = nn.Conv3d(t_io[0], t_io[1], kernel_size=(kernel_size[0], 1, 1), stride=(stride[0], 1, 1),
                         padding=(padding[0], 0, 0), bias=bias)
    spatial = nn.Conv3d(xy_io[0], xy_io[1], kernel_size=(1, kernel_size[1], kernel_size[2]),
                        stride=(1, stride[1], stride[2]), padding=(0, padding[1], padding[2]), bias=bias)
    return Conv2plus1d(
        conv_t=temporal,
        norm=None if norm is None else norm(num_features=mid, eps=norm_eps, momentum=norm_momentum),
        activation=None if activation is None else activation(),
        conv_xy=spatial,
        conv_xy_first=conv_xy_first,
    )
