"""Device-resident constants keyed by their host description.

Everything the step derives from *host-known* facts (GT counts, image sizes, level shapes) -- segment offsets,
index maps, masks, positional encodings -- is built once per distinct key and reused, instead of being rebuilt and
copied host->device every step as the reference does (``torch.tensor(range(..)).cuda()`` in dn_components.py:56-58,
``bbox_pred.new_tensor([w, h, w, h])`` in dino_detr_head.py:702-707, per-call meshgrids in transformer.py:676-691).
Besides removing a dozen small pageable copies per step this makes the whole step capturable in a CUDA graph.
"""
from collections import OrderedDict

import torch

_CACHE = OrderedDict()
_MAX = 4096


def device_const(device, tag, key, build):
    """``build()`` -> numpy array / CPU tensor / device tensor; cached under (device, tag, key)."""
    k = (str(device), tag, key)
    t = _CACHE.get(k)
    if t is None:
        v = build()
        t = v if (torch.is_tensor(v) and v.device == torch.device(device)) else torch.as_tensor(v).to(device)
        _CACHE[k] = t
        if len(_CACHE) > _MAX:
            _CACHE.popitem(last=False)
    else:
        _CACHE.move_to_end(k)
    return t


def clear():
    _CACHE.clear()
