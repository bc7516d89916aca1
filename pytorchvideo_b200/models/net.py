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


class DetectionBBoxNetwork(B200Module):
    """A trunk followed by a head that also takes bounding boxes (reference net.py:47-74):
    ``forward(x, bboxes)`` -> ``[K, num_classes]`` for bboxes [K, 5] = (batch index, x1, y1, x2, y2).
    Trunk, RoIAlign and head compile into ONE plan per (clip shapes, K)."""

    def __init__(self, model, detection_head):
        super().__init__()
        self.model = model
        self.detection_head = detection_head

    def forward(self, x, bboxes):
        ins = (list(x) if isinstance(x, (list, tuple)) else [x]) + [bboxes]
        out = self._pv_compiled(ins)(ins).clone()
        return out.view(out.shape[0], -1)


class MultiPathWayWithFuse(B200Module):
    """Per-pathway blocks followed by an optional cross-pathway fusion (net.py:66-122).

    NB the reference's ``inplace=True`` overwrites the caller's input list; the engine never
    mutates the inputs it is given."""

    def __init__(self, *, multipathway_blocks, multipathway_fusion, inplace=True):
        super().__init__()
        set_attributes(self, locals())
