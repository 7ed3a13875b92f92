"""Tiny driver for ncu: a few launches of the default MSDA forward / backward kernels at the microbench shape."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semi_detr_b200 import _lib  # noqa: E402
from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA  # noqa: E402
from semi_detr_b200.synthetic import MICROBENCH_LEVELS, msda_inputs  # noqa: E402

fv = int(sys.argv[1]) if len(sys.argv) > 1 else 0
bv = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mode = sys.argv[3] if len(sys.argv) > 3 else "encoder"
_lib.lib().sdb_msda_set_variant(fv, bv)
x = msda_inputs(MICROBENCH_LEVELS, N=2, mode=mode, Lq=17821, seed=0)
a = (x["value"], x["shapes"], x["start"], x["loc"], x["attn"])
for _ in range(3):
    MSDA.ms_deform_attn_forward(*a, 64)
    MSDA.ms_deform_attn_backward(*a, x["gout"], 64)
torch.cuda.synchronize()
