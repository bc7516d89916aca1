"""Plan compiler: lowers a PyTorchVideo-style module tree to a static list of libpvb200 launches
over a preallocated NDHWC arena, replayed as one CUDA graph per (model, input shape)."""
from .plan import Plan, TRef  # noqa: F401
from .lower import compile_model, CompiledModel  # noqa: F401
