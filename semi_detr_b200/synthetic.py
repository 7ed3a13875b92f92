"""Seeded synthetic inputs shared by the tests and bench.py."""
import torch

MICROBENCH_LEVELS = [(100, 134), (50, 67), (25, 34), (13, 17)]      # sum HW = 17821 (800x1066 image)
COCO_4SCALE_LEVELS = [(100, 167), (50, 84), (25, 42), (13, 21)]     # sum HW = 22223 (800x1333 image)


def level_tensors(levels, device):
    shapes = torch.as_tensor(levels, dtype=torch.long, device=device)
    start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    return shapes, start


def encoder_reference_points(levels, device):
    """Pixel centres of every level, (S, 2) as (x, y) in [0,1] (transformer.py:676-691 with valid_ratio 1)."""
    pts = []
    for (h, w) in levels:
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=device) + 0.5,
                                torch.arange(w, dtype=torch.float32, device=device) + 0.5, indexing="ij")
        pts.append(torch.stack([xs.reshape(-1) / w, ys.reshape(-1) / h], -1))
    return torch.cat(pts)


def msda_inputs(levels, N, M=8, D=32, P=4, Lq=None, mode="encoder", seed=0, device="cuda", dtype=torch.float32):
    """value ~ N(0,1); loc: 'encoder' = own grid point + U(-4,4) px per level, 'uniform' = U(0,1),
    'wide' = U(-0.3,1.3) (exercises borders / out-of-range); attn = softmax(N(0,1)) over L*P; grad ~ N(0,1)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    L = len(levels)
    S = sum(h * w for h, w in levels)
    shapes, start = level_tensors(levels, device)
    value = torch.randn(N, S, M, D, generator=g).to(device=device, dtype=dtype)
    if mode == "encoder":
        Lq = S
        ref = encoder_reference_points(levels, "cpu")                       # (S, 2)
        wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float32)  # (L, 2)
        off = (torch.rand(N, Lq, M, L, P, 2, generator=g) * 8 - 4) / wh[None, None, None, :, None, :]
        loc = ref[None, :, None, None, None, :] + off
    else:
        Lq = Lq or 1100
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g)
        if mode == "wide":
            loc = loc * 1.6 - 0.3
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    gout = torch.randn(N, Lq, M * D, generator=g)
    return dict(value=value, shapes=shapes, start=start, loc=loc.to(device=device, dtype=dtype).contiguous(),
                attn=attn.to(device=device, dtype=dtype).contiguous(), gout=gout.to(device=device, dtype=dtype))


def msda_bytes(N, S, Lq, M=8, D=32, L=4, P=4, elt=4):
    """Algorithmic HBM bytes: fwd = value + loc + attn + out;
    bwd = grad_out + value + loc + attn (reads) + grad_value + grad_loc + grad_attn (writes)."""
    v, o = N * S * M * D, N * Lq * M * D
    loc, att = N * Lq * M * L * P * 2, N * Lq * M * L * P
    return elt * (v + loc + att + o), elt * (o + 2 * v + 2 * loc + 2 * att)
