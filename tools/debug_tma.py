import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA
from semi_detr_b200.synthetic import msda_inputs
levels = [(19, 27), (10, 14), (5, 7), (3, 4)] if len(sys.argv) < 2 else [(100, 134), (50, 67), (25, 34), (13, 17)]
x = msda_inputs(levels, N=2, mode="encoder", seed=7)
a = (x["value"], x["shapes"], x["start"], x["loc"], x["attn"])
MSDA.USE_TMA = True
got = MSDA.ms_deform_attn_forward(*a, 64)
torch.cuda.synchronize()
MSDA.USE_TMA = False
want = MSDA.ms_deform_attn_forward(*a, 64)
torch.cuda.synchronize()
print("max abs diff", (got - want).abs().max().item(), "ref max", want.abs().max().item())
