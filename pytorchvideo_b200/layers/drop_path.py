import torch.nn as nn


class DropPath(nn.Module):
    """Stochastic depth marker (reference layers/drop_path.py:7-48); identity in eval mode."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob
