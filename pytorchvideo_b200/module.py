"""Base class of every block in this package: a parameter container whose ``forward`` runs the
B200 engine.  There is no ATen / CPU forward: calling a block with CPU tensors, in training mode
or without the CUDA library raises."""
import torch
import torch.nn as nn

from . import config


class B200Module(nn.Module):
    def _pv_fingerprint(self):
        v = 0
        for t in self.parameters():
            v += t._version + (t.data_ptr() & 0xFFFF)
        for t in self.buffers():
            v += t._version + (t.data_ptr() & 0xFFFF)
        return v

    def _pv_compiled(self, x):
        from .engine import compile_model
        ins = x if isinstance(x, (list, tuple)) else [x]
        for t in ins:
            if not torch.is_tensor(t):
                raise RuntimeError("expected tensor inputs")
            if t.device.type != "cuda":
                raise RuntimeError(
                    "pytorchvideo_b200 runs on B200 GPUs only (no CPU path); got a %s tensor" % t.device.type)
        if self.training:
            raise RuntimeError("pytorchvideo_b200 is an eval-mode forward engine: call model.eval() first")
        key = (tuple((tuple(t.shape), t.dtype, t.device.index) for t in ins), config.get_precision(),
               config.get_use_tcgen05(), config.get_use_graph(), self._pv_fingerprint())
        cache = self.__dict__.setdefault("_pv_cache", {})
        cm = cache.get(key)
        if cm is None:
            cache.clear()
            cm = compile_model(self, list(ins) if isinstance(x, (list, tuple)) else x, config.get_precision(),
                               config.get_use_tcgen05(), config.get_use_graph())
            cache[key] = cm
        return cm

    def forward(self, x):
        cm = self._pv_compiled(x)
        return cm(x).clone()
