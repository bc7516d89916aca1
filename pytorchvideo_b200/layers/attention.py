"""MViT building blocks (reference layers/attention.py): parameter containers whose ``forward`` runs the
B200 engine - ``Mlp(x)``, ``MultiScaleAttention(x, thw_shape) -> (x, thw)``, ``MultiScaleBlock(x, thw_shape)
-> (x, thw)`` on CUDA token tensors (B, N, C), lowered by engine/lower.py (no ATen forward).

Attribute names / registration order follow the reference so state_dict keys match, including the
``_attention_pool_{q,k,v}`` wrappers that re-register the pool conv and norm of each branch (the
same Parameter appears under ``attn.pool_k.weight`` and ``attn._attention_pool_k.pool.weight``)."""
import numpy
import torch.nn as nn

from ..module import B200Module
from .drop_path import DropPath


class Mlp(B200Module):
    """fc1 -> act (exact-erf GELU) -> fc2 (attention.py:51-114)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, dropout_rate=0.0,
                 bias_on=True):
        super().__init__()
        self.dropout_rate = dropout_rate
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias_on)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias_on)
        self.dropout = nn.Dropout(dropout_rate) if dropout_rate > 0.0 else nn.Identity()


class _AttentionPool(nn.Module):
    """Pool (cls token excluded) then norm (cls token included) - attention.py:117-212."""

    def __init__(self, pool, has_cls_embed, norm):
        super().__init__()
        self.has_pool = pool is not None
        self.pool = pool if pool is not None else nn.Identity()
        self.has_cls_embed = has_cls_embed
        if norm is not None:
            self.norm_before_pool = isinstance(norm, (nn.BatchNorm3d, nn.Identity))
            self.has_norm = True
            self.norm = norm
        else:
            self.norm_before_pool = False
            self.has_norm = False
            self.norm = nn.Identity()


def _prod(v):
    p = 1
    for i in v:
        p *= i
    return p


class MultiScaleAttention(B200Module):
    """Pooled multi-head attention (attention.py:215-544)."""
    _version = 3

    def __init__(self, dim, dim_out=None, num_heads=8, qkv_bias=False, dropout_rate=0.0, kernel_q=(1, 1, 1),
                 kernel_kv=(1, 1, 1), stride_q=(1, 1, 1), stride_kv=(1, 1, 1), norm_layer=nn.LayerNorm,
                 has_cls_embed=True, pool_mode="conv", pool_first=False, residual_pool=True, depthwise_conv=True,
                 bias_on=True, separate_qkv=True):
        super().__init__()
        assert pool_mode in ["conv", "avg", "max"]
        self.pool_first = pool_first
        self.dropout_rate = dropout_rate
        self.num_heads = num_heads
        dim_out = dim if not dim_out else dim_out
        self.dim_out = dim_out
        head_dim = dim_out // num_heads
        self.scale = head_dim ** -0.5
        self.has_cls_embed = has_cls_embed
        self.residual_pool = residual_pool
        self.separate_qkv = separate_qkv
        pad_q = [int(q // 2) for q in kernel_q]
        pad_kv = [int(kv // 2) for kv in kernel_kv]
        self.q = self.k = self.v = self.qkv = nn.Identity()
        if pool_first or separate_qkv:
            self.q = nn.Linear(dim, dim_out, bias=qkv_bias)
            self.k = nn.Linear(dim, dim_out, bias=qkv_bias)
            self.v = nn.Linear(dim, dim_out, bias=qkv_bias)
        else:
            self.qkv = nn.Linear(dim, dim_out * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim_out, dim_out, bias=True if bias_on else False)
        self.proj_drop = nn.Dropout(dropout_rate) if dropout_rate > 0.0 else nn.Identity()
        if kernel_q is not None and _prod(kernel_q) == 1 and _prod(stride_q) == 1:
            kernel_q = None
        if kernel_kv is not None and _prod(kernel_kv) == 1 and _prod(stride_kv) == 1:
            kernel_kv = None
        if pool_mode in ("avg", "max"):
            op = nn.MaxPool3d if pool_mode == "max" else nn.AvgPool3d
            self.pool_q = op(kernel_q, stride_q, pad_q, ceil_mode=False) if kernel_q is not None else None
            self.pool_k = op(kernel_kv, stride_kv, pad_kv, ceil_mode=False) if kernel_kv is not None else None
            self.pool_v = op(kernel_kv, stride_kv, pad_kv, ceil_mode=False) if kernel_kv is not None else None
        else:
            dim_conv = (dim if pool_first else dim_out) // num_heads

            def conv(kernel, stride, pad):
                return nn.Conv3d(dim_conv, dim_conv, kernel, stride=stride, padding=pad,
                                 groups=dim_conv if depthwise_conv else 1, bias=False)
            self.pool_q = conv(kernel_q, stride_q, pad_q) if kernel_q is not None else None
            self.norm_q = norm_layer(dim_conv) if kernel_q is not None else None
            self.pool_k = conv(kernel_kv, stride_kv, pad_kv) if kernel_kv is not None else None
            self.norm_k = norm_layer(dim_conv) if kernel_kv is not None else None
            self.pool_v = conv(kernel_kv, stride_kv, pad_kv) if kernel_kv is not None else None
            self.norm_v = norm_layer(dim_conv) if kernel_kv is not None else None
        self._attention_pool_q = _AttentionPool(self.pool_q, has_cls_embed, getattr(self, "norm_q", None))
        self._attention_pool_k = _AttentionPool(self.pool_k, has_cls_embed, getattr(self, "norm_k", None))
        self._attention_pool_v = _AttentionPool(self.pool_v, has_cls_embed, getattr(self, "norm_v", None))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        # pre-v2 checkpoints only carry pool_/norm_ keys; mirror them onto the wrapper aliases
        version = local_metadata.get("version", None)
        if version is None or version < 2:
            for layer in ("pool", "norm"):
                for pat in ("q", "k", "v"):
                    for typ in ("weight", "bias"):
                        old = f"{prefix}{layer}_{pat}.{typ}"
                        if old in state_dict:
                            state_dict[f"{prefix}_attention_pool_{pat}.{layer}.{typ}"] = state_dict[old]
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)


class MultiScaleBlock(B200Module):
    """norm1 -> attention (+ pooled skip) -> norm2 -> Mlp (+ skip / dim-expanding proj), attention.py:578-757."""

    def __init__(self, dim, dim_out, num_heads, mlp_ratio=4.0, qkv_bias=False, dropout_rate=0.0, droppath_rate=0.0,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, attn_norm_layer=nn.LayerNorm, dim_mul_in_att=False,
                 kernel_q=(1, 1, 1), kernel_kv=(1, 1, 1), stride_q=(1, 1, 1), stride_kv=(1, 1, 1), pool_mode="conv",
                 has_cls_embed=True, pool_first=False, residual_pool=False, depthwise_conv=True, bias_on=True,
                 separate_qkv=True):
        super().__init__()
        self.dim = dim
        self.dim_out = dim_out
        self.norm1 = norm_layer(dim)
        self.dim_mul_in_att = dim_mul_in_att
        self.norm1_is_batchnorm_1d = isinstance(self.norm1, nn.BatchNorm1d)
        kernel_skip = [s + 1 if s > 1 else s for s in stride_q]
        stride_skip = stride_q
        padding_skip = [int(k // 2) for k in kernel_skip]
        att_dim = dim_out if dim_mul_in_att else dim
        self.attn = MultiScaleAttention(dim=dim, dim_out=att_dim, num_heads=num_heads, qkv_bias=qkv_bias,
                                        dropout_rate=dropout_rate, kernel_q=kernel_q, kernel_kv=kernel_kv,
                                        stride_q=stride_q, stride_kv=stride_kv, norm_layer=attn_norm_layer,
                                        has_cls_embed=has_cls_embed, pool_mode=pool_mode, pool_first=pool_first,
                                        residual_pool=residual_pool, bias_on=bias_on, depthwise_conv=depthwise_conv,
                                        separate_qkv=separate_qkv)
        self.drop_path = DropPath(droppath_rate) if droppath_rate > 0.0 else nn.Identity()
        self.norm2 = norm_layer(att_dim)
        self.norm2_is_batchnorm_1d = isinstance(self.norm2, nn.BatchNorm1d)
        self.has_cls_embed = has_cls_embed
        self.mlp = Mlp(in_features=att_dim, hidden_features=int(att_dim * mlp_ratio), out_features=dim_out,
                       act_layer=act_layer, dropout_rate=dropout_rate, bias_on=bias_on)
        self.proj = nn.Linear(dim, dim_out, bias=bias_on) if dim != dim_out else nn.Identity()
        self.pool_skip = (nn.MaxPool3d(kernel_skip, stride_skip, padding_skip, ceil_mode=False)
                          if len(stride_skip) > 0 and numpy.prod(stride_skip) > 1 else None)
        self._attention_pool = _AttentionPool(self.pool_skip, has_cls_embed=has_cls_embed, norm=None)
