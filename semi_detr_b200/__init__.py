"""semi_detr_b200 -- B200-native (sm_100a) implementation of Semi-DETR's data-parallel hot path.

CUDA kernels live in ``csrc/`` behind the C ABI of ``include/semidetr_b200.h``
(``lib/libsemidetr_b200.so``, built by ``python -m semi_detr_b200.build``); the sub-packages mirror the
reference's operator / assigner / hook interfaces for that path.  There is no CPU fallback anywhere in this
package: the CPU restatement used for parity checks lives in the top-level ``oracle/`` and is never imported
from here.
"""
import sys

__version__ = "0.1.0"


def install_as_reference_extension():
    """Expose the MSDA kernels under the reference's extension-module name so that the reference's own
    ``functions/ms_deform_attn_func.py`` (which does ``import MultiScaleDeformableAttention as MSDA`` at :18)
    runs on them unchanged."""
    from .msda import MultiScaleDeformableAttention as ext
    sys.modules["MultiScaleDeformableAttention"] = ext
    return ext
