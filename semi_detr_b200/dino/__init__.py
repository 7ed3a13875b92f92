from .backbone import ResNet
from .detector import DINODETR
from .dn_components import dn_post_process, prepare_for_cdn
from .head import DINODETRHead
from .positional_encoding import SinePositionalEncodingHW
from .transformer import DINOTransformer

__all__ = ["ResNet", "DINODETR", "DINODETRHead", "DINOTransformer", "SinePositionalEncodingHW", "prepare_for_cdn",
           "dn_post_process"]
