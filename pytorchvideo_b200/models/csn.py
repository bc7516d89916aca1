"""Channel-Separated Network (reference models/csn.py:12-191): bottleneck conv_b is a depthwise
3x3x3 convolution (groups = dim_inner // width_per_group)."""
import torch
import torch.nn as nn

from .head import create_res_basic_head
from .net import Net
from .resnet import _MODEL_STAGE_DEPTH, create_bottleneck_block, create_res_stage
from .stem import create_res_basic_stem


def create_csn(*, input_channel=3, model_depth=50, model_num_class=400, dropout_rate=0, norm=nn.BatchNorm3d,
               activation=nn.ReLU, stem_dim_out=64, stem_conv_kernel_size=(3, 7, 7), stem_conv_stride=(1, 2, 2),
               stem_pool=None, stem_pool_kernel_size=(1, 3, 3), stem_pool_stride=(1, 2, 2),
               stage_conv_a_kernel_size=(1, 1, 1), stage_conv_b_kernel_size=(3, 3, 3),
               stage_conv_b_width_per_group=1, stage_spatial_stride=(1, 2, 2, 2),
               stage_temporal_stride=(1, 2, 2, 2), bottleneck=create_bottleneck_block, bottleneck_ratio=4,
               head_pool=nn.AvgPool3d, head_pool_kernel_size=(1, 7, 7), head_output_size=(1, 1, 1),
               head_activation=None, head_output_with_global_average=True):
    torch._C._log_api_usage_once("PYTORCHVIDEO.model.create_csn")
    assert model_depth in _MODEL_STAGE_DEPTH, f"{model_depth} is not in {_MODEL_STAGE_DEPTH.keys()}"
    depths = _MODEL_STAGE_DEPTH[model_depth]
    blocks = [create_res_basic_stem(
        in_channels=input_channel, out_channels=stem_dim_out, conv_kernel_size=stem_conv_kernel_size,
        conv_stride=stem_conv_stride, conv_padding=[k // 2 for k in stem_conv_kernel_size], pool=stem_pool,
        pool_kernel_size=stem_pool_kernel_size, pool_stride=stem_pool_stride,
        pool_padding=[k // 2 for k in stem_pool_kernel_size], norm=norm, activation=activation)]
    width_in, width_out = stem_dim_out, stem_dim_out * 4
    for s, depth in enumerate(depths):
        inner = width_out // bottleneck_ratio
        blocks.append(create_res_stage(
            depth=depth, dim_in=width_in, dim_inner=inner, dim_out=width_out, bottleneck=bottleneck,
            conv_a_kernel_size=stage_conv_a_kernel_size, conv_a_stride=(1, 1, 1),
            conv_a_padding=[k // 2 for k in stage_conv_a_kernel_size],
            conv_b_kernel_size=stage_conv_b_kernel_size,
            conv_b_stride=(stage_temporal_stride[s], stage_spatial_stride[s], stage_spatial_stride[s]),
            conv_b_padding=[k // 2 for k in stage_conv_b_kernel_size],
            conv_b_num_groups=inner // stage_conv_b_width_per_group, conv_b_dilation=(1, 1, 1), norm=norm,
            activation=activation))
        width_in, width_out = width_out, width_out * 2
    blocks.append(create_res_basic_head(
        in_features=width_in, out_features=model_num_class, pool=head_pool, output_size=head_output_size,
        pool_kernel_size=head_pool_kernel_size, dropout_rate=dropout_rate, activation=head_activation,
        output_with_global_average=head_output_with_global_average))
    return Net(blocks=nn.ModuleList(blocks))
