from ..module import B200Module


class Swish(B200Module):
    """x * sigmoid(x) (reference layers/swish.py:7-28).  Inside a block the engine fuses it into the
    producing kernel's epilogue (PV_ACT_SWISH); called on its own it is one elementwise launch
    (engine/lower.py lower_Swish).  It owns no parameters."""
