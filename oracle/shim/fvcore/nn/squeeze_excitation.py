"""fvcore.nn.squeeze_excitation.SqueezeExcitation restatement.

Structure pinned by the reference's own use of it:
/root/reference/pytorchvideo/layers/accelerator/mobile_cpu/attention.py:62-104 indexes
``se.block[0]`` (conv), ``[1]`` (activation), ``[2]`` (conv), ``[3]`` (sigmoid); hub X3D
checkpoints carry keys ``...norm_b.1.block.0.weight``.
"""
import torch.nn as nn


class SqueezeExcitation(nn.Module):
    def __init__(self, num_channels, num_channels_reduced=None, reduction_ratio=2.0,
                 is_3d=False, activation=None):
        super().__init__()
        if num_channels_reduced is None:
            num_channels_reduced = int(num_channels // reduction_ratio)
        conv = nn.Conv3d if is_3d else nn.Conv2d
        self.is_3d = is_3d
        self.block = nn.Sequential(
            conv(num_channels, num_channels_reduced, kernel_size=1, stride=1, bias=True),
            nn.ReLU() if activation is None else activation,
            conv(num_channels_reduced, num_channels, kernel_size=1, stride=1, bias=True),
            nn.Sigmoid(),
        )

    def forward(self, x):
        dims = [2, 3, 4] if self.is_3d else [2, 3]
        return x * self.block(x.mean(dim=dims, keepdim=True))
