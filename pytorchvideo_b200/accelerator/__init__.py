"""The reference's accelerator plug-in protocol, with a "b200" target.

Mirrors accelerator/deployment/common/model_transmuter.py:16-86 (registry + ``transmute_model``)
and accelerator/efficient_blocks/efficient_block_base.py:8-35 (``EfficientBlockBase`` with
``convert(input_blob_size)``).  The "b200" transmuter replaces WHOLE blocks (a ``Net``,
``ResStage``, ``ResBlock``, stem, head, ...) - never single convolutions - because the unit of
fusion on Blackwell is the block (BN fold + residual + activation + concat fusion).
"""
from .model_transmuter import EFFICIENT_BLOCK_TRANSMUTER_REGISTRY, transmute_model  # noqa: F401
from .efficient_block_base import EfficientBlockBase  # noqa: F401
from .b200 import B200Block, convert_to_deployable_form, transmute_b200  # noqa: F401
