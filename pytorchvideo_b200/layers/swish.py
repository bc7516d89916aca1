import torch.nn as nn


class Swish(nn.Module):
    """x * sigmoid(x) (reference layers/swish.py:7-28).  A marker module: the engine fuses it
    into the producing kernel's epilogue (PV_ACT_SWISH); it owns no parameters."""

    def forward(self, x):
        raise RuntimeError("Swish is fused by the B200 engine; call the enclosing block instead")
