"""Registry of per-target transmuter functions and the in-place tree walk
(reference model_transmuter.py:16-86)."""
import logging
from typing import Callable, Dict, List, Optional

import torch.nn as nn

EFFICIENT_BLOCK_TRANSMUTER_REGISTRY: Dict[str, List[Callable[[nn.Module], Optional[nn.Module]]]] = {}


def _first_replacement(module, transmuters):
    for fn in transmuters:
        new = fn(module)
        if new is not None:
            return new            # first hit wins
    return None


def transmute_model(model: nn.Module, target_device: str = "b200", prefix: str = "") -> None:
    """Replace, IN PLACE, every child some registered transmuter of ``target_device`` recognises;
    recursion stops at a replaced node."""
    assert target_device in EFFICIENT_BLOCK_TRANSMUTER_REGISTRY, f"{target_device} not registered!"
    transmuters = EFFICIENT_BLOCK_TRANSMUTER_REGISTRY[target_device]
    for name, child in model.named_children():
        path = f"{prefix}.{name}" if prefix else name
        new = _first_replacement(child, transmuters)
        if new is not None:
            logging.info("transmuting %s (%s) for %s", path, type(child).__name__, target_device)
            model._modules[name] = new
        else:
            transmute_model(child, target_device, path)
