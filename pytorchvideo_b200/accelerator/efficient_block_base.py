from abc import abstractmethod

import torch.nn as nn


class EfficientBlockBase(nn.Module):
    """Hardware-specific block: ``convert`` specialises it to static shapes, ``forward`` runs it
    (reference efficient_block_base.py:8-35)."""

    @abstractmethod
    def convert(self, *args, **kwargs):
        raise NotImplementedError

    @abstractmethod
    def forward(self, *args, **kwargs):
        raise NotImplementedError
