"""X3D (reference models/x3d.py): depthwise-separable 3-D ResNet with SE + Swish."""
import math

import numpy as np
import torch
import torch.nn as nn

from ..layers.convolutions import Conv2plus1d
from ..layers.squeeze_excitation import SqueezeExcitation
from ..layers.swish import Swish
from ..layers.utils import round_repeats, round_width, set_attributes
from ..module import B200Module
from .head import ResNetBasicHead, _head_activation
from .net import Net
from .resnet import BottleneckBlock, ResBlock, ResStage
from .stem import ResNetBasicStem


def _bn(norm, c, eps, momentum):
    return None if norm is None else norm(num_features=c, eps=eps, momentum=momentum)


def create_x3d_stem(*, in_channels, out_channels, conv_kernel_size=(5, 3, 3), conv_stride=(1, 2, 2),
                    conv_padding=(2, 1, 1), norm=nn.BatchNorm3d, norm_eps=1e-5, norm_momentum=0.1,
                    activation=nn.ReLU):
    """Spatial 1xkxk conv, then a depthwise kx1x1 temporal conv, BN, ReLU (x3d.py:19-102).
    NB the spatial conv sits in the ``conv_t`` slot and the temporal one in ``conv_xy``."""
    spatial = nn.Conv3d(in_channels, out_channels, kernel_size=(1, conv_kernel_size[1], conv_kernel_size[2]),
                        stride=(1, conv_stride[1], conv_stride[2]), padding=(0, conv_padding[1], conv_padding[2]),
                        bias=False)
    temporal = nn.Conv3d(out_channels, out_channels, kernel_size=(conv_kernel_size[0], 1, 1),
                         stride=(conv_stride[0], 1, 1), padding=(conv_padding[0], 0, 0), bias=False,
                         groups=out_channels)
    return ResNetBasicStem(conv=Conv2plus1d(conv_t=spatial, norm=None, activation=None, conv_xy=temporal),
                           norm=_bn(norm, out_channels, norm_eps, norm_momentum),
                           activation=None if activation is None else activation(), pool=None)


def create_x3d_bottleneck_block(*, dim_in, dim_inner, dim_out, conv_kernel_size=(3, 3, 3), conv_stride=(1, 2, 2),
                                norm=nn.BatchNorm3d, norm_eps=1e-5, norm_momentum=0.1, se_ratio=0.0625,
                                activation=nn.ReLU, inner_act=Swish):
    """1x1x1 expand -> depthwise 3x3x3 (+BN, optional SE, Swish) -> 1x1x1 project (x3d.py:105-228)."""
    se = SqueezeExcitation(num_channels=dim_inner, num_channels_reduced=round_width(dim_inner, se_ratio),
                           is_3d=True) if se_ratio > 0.0 else nn.Identity()
    return BottleneckBlock(
        conv_a=nn.Conv3d(dim_in, dim_inner, kernel_size=(1, 1, 1), bias=False),
        norm_a=_bn(norm, dim_inner, norm_eps, norm_momentum),
        act_a=None if activation is None else activation(),
        conv_b=nn.Conv3d(dim_inner, dim_inner, kernel_size=conv_kernel_size, stride=conv_stride,
                         padding=[k // 2 for k in conv_kernel_size], bias=False, groups=dim_inner,
                         dilation=(1, 1, 1)),
        norm_b=nn.Sequential(nn.Identity() if norm is None else _bn(norm, dim_inner, norm_eps, norm_momentum), se),
        act_b=None if inner_act is None else inner_act(),
        conv_c=nn.Conv3d(dim_inner, dim_out, kernel_size=(1, 1, 1), bias=False),
        norm_c=_bn(norm, dim_out, norm_eps, norm_momentum),
    )


def create_x3d_res_block(*, dim_in, dim_inner, dim_out, bottleneck=create_x3d_bottleneck_block,
                         use_shortcut=True, conv_kernel_size=(3, 3, 3), conv_stride=(1, 2, 2),
                         norm=nn.BatchNorm3d, norm_eps=1e-5, norm_momentum=0.1, se_ratio=0.0625,
                         activation=nn.ReLU, inner_act=Swish):
    widen = dim_in != dim_out
    project = (widen or int(np.prod(conv_stride)) > 1) and use_shortcut
    return ResBlock(
        branch1_conv=nn.Conv3d(dim_in, dim_out, kernel_size=(1, 1, 1), stride=conv_stride, bias=False)
        if project else None,
        # the shortcut is only normalised when the width changes (x3d.py:293-310)
        branch1_norm=norm(num_features=dim_out) if (norm is not None and widen and use_shortcut) else None,
        branch2=bottleneck(dim_in=dim_in, dim_inner=dim_inner, dim_out=dim_out, conv_kernel_size=conv_kernel_size,
                           conv_stride=conv_stride, norm=norm, norm_eps=norm_eps, norm_momentum=norm_momentum,
                           se_ratio=se_ratio, activation=activation, inner_act=inner_act),
        activation=None if activation is None else activation(),
        branch_fusion=lambda x, y: x + y,
    )


def create_x3d_res_stage(*, depth, dim_in, dim_inner, dim_out, bottleneck=create_x3d_bottleneck_block,
                         conv_kernel_size=(3, 3, 3), conv_stride=(1, 2, 2), norm=nn.BatchNorm3d, norm_eps=1e-5,
                         norm_momentum=0.1, se_ratio=0.0625, activation=nn.ReLU, inner_act=Swish):
    blocks = [create_x3d_res_block(
        dim_in=dim_in if i == 0 else dim_out, dim_inner=dim_inner, dim_out=dim_out, bottleneck=bottleneck,
        conv_kernel_size=conv_kernel_size, conv_stride=conv_stride if i == 0 else (1, 1, 1), norm=norm,
        norm_eps=norm_eps, norm_momentum=norm_momentum,
        se_ratio=se_ratio if (i + 1) % 2 else 0.0,   # SE on every other block, starting with the first
        activation=activation, inner_act=inner_act) for i in range(depth)]
    return ResStage(res_blocks=nn.ModuleList(blocks))


class ProjectedPool(B200Module):
    """pre_conv/norm/act -> pool -> post_conv/norm/act (x3d.py:730-806)."""

    def __init__(self, *, pre_conv=None, pre_norm=None, pre_act=None, pool=None, post_conv=None,
                 post_norm=None, post_act=None):
        super().__init__()
        set_attributes(self, locals())
        assert self.pre_conv is not None and self.pool is not None and self.post_conv is not None


def create_x3d_head(*, dim_in, dim_inner, dim_out, num_classes, pool_act=nn.ReLU, pool_kernel_size=(13, 5, 5),
                    norm=nn.BatchNorm3d, norm_eps=1e-5, norm_momentum=0.1, bn_lin5_on=False, dropout_rate=0.5,
                    activation=nn.Softmax, output_with_global_average=True):
    if activation is not None and activation not in (nn.Softmax, nn.Sigmoid):
        raise NotImplementedError("{} is not supported as an activationfunction.".format(activation))
    pool = ProjectedPool(
        pre_conv=nn.Conv3d(dim_in, dim_inner, kernel_size=(1, 1, 1), bias=False),
        pre_norm=norm(num_features=dim_inner, eps=norm_eps, momentum=norm_momentum),
        pre_act=None if pool_act is None else pool_act(),
        pool=nn.AdaptiveAvgPool3d((1, 1, 1)) if pool_kernel_size is None else nn.AvgPool3d(pool_kernel_size, stride=1),
        post_conv=nn.Conv3d(dim_inner, dim_out, kernel_size=(1, 1, 1), bias=False),
        post_norm=norm(num_features=dim_out, eps=norm_eps, momentum=norm_momentum) if bn_lin5_on else None,
        post_act=None if pool_act is None else pool_act(),
    )
    return ResNetBasicHead(proj=nn.Linear(dim_out, num_classes, bias=True), activation=_head_activation(activation),
                           pool=pool, dropout=nn.Dropout(dropout_rate) if dropout_rate > 0 else None,
                           output_pool=nn.AdaptiveAvgPool3d(1) if output_with_global_average else None)


def create_x3d(*, input_channel=3, input_clip_length=13, input_crop_size=160, model_num_class=400,
               dropout_rate=0.5, width_factor=2.0, depth_factor=2.2, norm=nn.BatchNorm3d, norm_eps=1e-5,
               norm_momentum=0.1, activation=nn.ReLU, stem_dim_in=12, stem_conv_kernel_size=(5, 3, 3),
               stem_conv_stride=(1, 2, 2), stage_conv_kernel_size=((3, 3, 3),) * 4,
               stage_spatial_stride=(2, 2, 2, 2), stage_temporal_stride=(1, 1, 1, 1),
               bottleneck=create_x3d_bottleneck_block, bottleneck_factor=2.25, se_ratio=0.0625, inner_act=Swish,
               head_dim_out=2048, head_pool_act=nn.ReLU, head_bn_lin5_on=False, head_activation=None,
               head_output_with_global_average=True):
    """X3D builder (reference x3d.py:539-727).  XS: clip 4x160^2, M: 16x224^2 (hub/x3d.py)."""
    torch._C._log_api_usage_once("PYTORCHVIDEO.model.create_x3d")
    stem_out = round_width(stem_dim_in, width_factor)
    blocks = [create_x3d_stem(in_channels=input_channel, out_channels=stem_out,
                              conv_kernel_size=stem_conv_kernel_size, conv_stride=stem_conv_stride,
                              conv_padding=[k // 2 for k in stem_conv_kernel_size], norm=norm, norm_eps=norm_eps,
                              norm_momentum=norm_momentum, activation=activation)]
    base_depths = [1, 2, 5, 3]
    base_dims = [stem_dim_in]
    for _ in range(3):
        base_dims.append(round_width(base_dims[-1], 2.0, divisor=8))
    dim_in = stem_out
    for s in range(4):
        dim_out = round_width(base_dims[s], width_factor)
        dim_inner = int(bottleneck_factor * dim_out)
        blocks.append(create_x3d_res_stage(
            depth=round_repeats(base_depths[s], depth_factor), dim_in=dim_in, dim_inner=dim_inner, dim_out=dim_out,
            bottleneck=bottleneck, conv_kernel_size=stage_conv_kernel_size[s],
            conv_stride=(stage_temporal_stride[s], stage_spatial_stride[s], stage_spatial_stride[s]), norm=norm,
            norm_eps=norm_eps, norm_momentum=norm_momentum, se_ratio=se_ratio, activation=activation,
            inner_act=inner_act))
        dim_in = dim_out
    spatial_total = stem_conv_stride[1] * np.prod(stage_spatial_stride)
    temporal_total = stem_conv_stride[0] * np.prod(stage_temporal_stride)
    assert input_clip_length >= temporal_total, "Clip length doesn't match temporal stride!"
    assert input_crop_size >= spatial_total, "Crop size doesn't match spatial stride!"
    side = int(math.ceil(input_crop_size / spatial_total))
    blocks.append(create_x3d_head(
        dim_in=dim_out, dim_inner=dim_inner, dim_out=head_dim_out, num_classes=model_num_class,
        pool_act=head_pool_act, pool_kernel_size=(int(input_clip_length // temporal_total), side, side), norm=norm,
        norm_eps=norm_eps, norm_momentum=norm_momentum, bn_lin5_on=head_bn_lin5_on, dropout_rate=dropout_rate,
        activation=head_activation, output_with_global_average=head_output_with_global_average))
    return Net(blocks=nn.ModuleList(blocks))
