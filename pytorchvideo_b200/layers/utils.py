"""Builder arithmetic shared by the model factories (reference layers/utils.py:7-49)."""
import math


def set_attributes(obj, params=None):
    """Copy constructor arguments onto ``obj`` (everything but ``self``)."""
    for key, value in (params or {}).items():
        if key != "self" and key != "__class__":
            setattr(obj, key, value)


def round_width(width, multiplier, min_width=8, divisor=8, ceil=False):
    """Scale a channel count and snap it to ``divisor`` (never dropping below 90%)."""
    if not multiplier:
        return width
    scaled = width * multiplier
    floor_w = min_width or divisor
    snapped = int(math.ceil(scaled / divisor)) * divisor if ceil else int(scaled + divisor / 2) // divisor * divisor
    snapped = max(floor_w, snapped)
    if snapped < 0.9 * scaled:
        snapped += divisor
    return int(snapped)


def round_repeats(repeats, multiplier):
    return repeats if not multiplier else int(math.ceil(multiplier * repeats))
