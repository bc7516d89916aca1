import torch.nn as nn

from ..module import B200Module


class SqueezeExcitation(B200Module):
    """Parameter container for the fvcore SqueezeExcitation used by X3D
    (reference models/x3d.py:190-198; fvcore is not vendored - its structure is pinned by
    layers/accelerator/mobile_cpu/attention.py:62-104 and by hub checkpoint keys
    ``...norm_b.1.block.{0,2}.{weight,bias}``).
    gate = sigmoid(W2 relu(W1 mean_THW(x) + b1) + b2);  y = x * gate  (fused on device)."""

    def __init__(self, num_channels, num_channels_reduced=None, reduction_ratio=2.0, is_3d=True,
                 activation=None):
        super().__init__()
        if num_channels_reduced is None:
            num_channels_reduced = int(num_channels // reduction_ratio)
        if not is_3d:
            raise NotImplementedError("2-D SqueezeExcitation is outside the video hot path")
        self.is_3d = True
        self.block = nn.Sequential(
            nn.Conv3d(num_channels, num_channels_reduced, kernel_size=1, bias=True),
            nn.ReLU() if activation is None else activation,
            nn.Conv3d(num_channels_reduced, num_channels, kernel_size=1, bias=True),
            nn.Sigmoid(),
        )
