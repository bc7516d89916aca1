""""b200" target: wraps a supported block (this package's or the reference's own module - the
lowering dispatches on class/attribute names) into an EfficientBlockBase whose ``convert`` builds
the static plan + CUDA graph for the recorded input shape."""
from copy import deepcopy

import torch
import torch.nn as nn

from .efficient_block_base import EfficientBlockBase
from .model_transmuter import EFFICIENT_BLOCK_TRANSMUTER_REGISTRY, transmute_model

_WHOLE_BLOCKS = ("Net", "ResStage", "ResBlock", "BottleneckBlock", "ResNetBasicStem", "ResNetBasicHead",
                 "MultiPathWayWithFuse", "ProjectedPool")


class B200Block(EfficientBlockBase):
    def __init__(self, block: nn.Module, dtype="f16"):
        super().__init__()
        self.block = block            # parameters stay owned by the original module (same state_dict keys + "block.")
        self.dtype = dtype
        self._compiled = None

    def convert(self, input_blob_size=None, **kwargs):
        assert self._compiled is None, "B200Block: already converted, cannot be converted again"
        from ..engine import compile_model
        if input_blob_size is None:
            raise ValueError("convert() needs the input shape(s)")
        shapes = input_blob_size if isinstance(input_blob_size[0], (tuple, list, torch.Size)) else [input_blob_size]
        ex = [torch.empty(tuple(s), dtype=torch.float32, device="cuda") for s in shapes]
        self.block.eval()
        self._compiled = compile_model(self.block, ex if len(ex) > 1 else ex[0], dtype=self.dtype)

    def forward(self, x):
        if self._compiled is None:    # unconverted: compile lazily for this shape
            shapes = [tuple(t.shape) for t in x] if isinstance(x, (list, tuple)) else tuple(x.shape)
            self.convert(shapes)
        return self._compiled(x).clone()


def transmute_b200(module: nn.Module):
    """Transmuter fn: returns a B200Block for whole blocks the engine can lower, else None."""
    if type(module).__name__ in _WHOLE_BLOCKS:
        return B200Block(module)
    return None


EFFICIENT_BLOCK_TRANSMUTER_REGISTRY.setdefault("b200", []).append(transmute_b200)


def convert_to_deployable_form(model: nn.Module, input_tensor, **kwargs) -> nn.Module:
    """Reference protocol (mobile_cpu/utils/model_conversion.py:87-125): deep-copy, eval, record each
    efficient block's input shape with one hooked forward, then call ``convert(input_blob_size)``."""
    model = deepcopy(model).eval()
    if type(model).__name__ in _WHOLE_BLOCKS:
        model = B200Block(model)
        shapes = [tuple(t.shape) for t in input_tensor] if isinstance(input_tensor, (list, tuple)) else tuple(input_tensor.shape)
        model.convert(shapes)
        return model
    transmute_model(model, "b200")
    shapes = {}
    hooks = []
    for m in model.modules():
        if isinstance(m, B200Block):
            def hook(mod, inp):
                x = inp[0]
                shapes[mod] = [tuple(t.shape) for t in x] if isinstance(x, (list, tuple)) else tuple(x.shape)
            hooks.append(m.register_forward_pre_hook(hook))
    with torch.no_grad():
        model(input_tensor)
    for h in hooks:
        h.remove()
    return model
