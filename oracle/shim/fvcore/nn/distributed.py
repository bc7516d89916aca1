"""Import-only stubs (reference layers/batch_norm.py:6, models/simclr.py:9)."""


def differentiable_all_reduce(x):
    raise NotImplementedError("training-only collective; out of scope")


def differentiable_all_gather(x):
    raise NotImplementedError("training-only collective; out of scope")
