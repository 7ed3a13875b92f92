"""Drop-in ``nn.LayerNorm`` / ``nn.Linear`` subclasses (same parameters and state dicts) whose CUDA paths use this
library's kernels for the transformer's token-wise work."""
from .layernorm import LayerNorm
from .linear import Linear, column_sum

__all__ = ["LayerNorm", "Linear", "column_sum"]
