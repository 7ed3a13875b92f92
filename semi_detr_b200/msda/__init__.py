from .functions import MSDeformAttnFunction
from .modules import MSDeformAttn

__all__ = ["MSDeformAttnFunction", "MSDeformAttn"]
