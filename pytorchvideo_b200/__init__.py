"""pytorchvideo_b200 - a Blackwell (sm_100a) forward-path engine behind PyTorchVideo's API.

Host code is Python/PyTorch (module trees, state_dict keys and builder signatures mirror
facebookresearch/pytorchvideo); every op on the hot path is a hand-written CUDA kernel in
``libpvb200.so`` reached through the C ABI of ``include/pv_b200.h``.  There is no CPU path.
"""
__version__ = "0.1.0"

from . import config  # noqa: F401
from .config import set_precision  # noqa: F401
