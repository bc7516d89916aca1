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
