"""Top-level containers (reference models/net.py:11-122)."""
import torch.nn as nn

from ..layers.utils import set_attributes
from ..module import B200Module
from .weight_init import init_net_weights


class Net(B200Module):
    """Sequential container of blocks; ``forward`` compiles the whole tree into one plan."""

    def __init__(self, *, blocks):
        super().__init__()
        assert blocks is not None
        self.blocks = blocks
        init_net_weights(self)

    def forward(self, x, *extra):
        # A Net whose blocks were replaced one by one by the accelerator protocol (transmute_model(model, "b200"))
        # runs them in sequence like the reference's Net.forward (net.py:41-44); each converted block owns its plan and
        # hands an NCDHW tensor to the next.  An untouched Net compiles into ONE plan (no layout round trips).
        if any(type(b).__name__ == "B200Block" for b in self.blocks):
            for b in self.blocks:
                x = b(x)
            return x
        return super().forward(x, *extra)


class MultiPathWayWithFuse(B200Module):
    """Per-pathway blocks followed by an optional cross-pathway fusion (net.py:66-122).

    NB the reference's ``inplace=True`` overwrites the caller's input list; the engine never
    mutates the inputs it is given."""

    def __init__(self, *, multipathway_blocks, multipathway_fusion, inplace=True):
        super().__init__()
        set_attributes(self, locals())
