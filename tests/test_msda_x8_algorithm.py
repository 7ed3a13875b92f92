"""Lane-level emulation in numpy of the 4-lanes x 8-channels gather the encoder backward uses (step C of
csrc/msda_backward_tile.cu; first written for the stand-alone x8 kernel, which measured 2x slower than the 8-lane one
because it split every reduction line into half-sector requests and was removed -- profiles/msda_bwd_variants_r2.txt),
against the oracle.  It pins the ALGORITHM -- which lane prepares which
level, what each broadcast carries, the two-stage reduce-scatter that leaves lane j with the corner sums of point j,
the closed forms of grad_x / grad_y / grad_attn from the four corner dot products, and (fused form) the chain through
the location arithmetic and the softmax -- so that a first hardware run only has the transcription left to get wrong.
Every array below is indexed [lane] exactly where the kernel holds a per-lane register."""
import numpy as np
import pytest
import torch

from oracle import msda_oracle as O


def _prep(x, y, a, H, W, start):
    """x8_prep / make_tap: pixel offset of (h0, w0), corner-validity bits, fractional offsets, weight."""
    h_im, w_im = y * H - 0.5, x * W - 0.5
    ok = (h_im > -1) and (w_im > -1) and (h_im < H) and (w_im < W)
    h0, w0 = (int(np.floor(h_im)), int(np.floor(w_im))) if ok else (0, 0)
    lh, lw = (h_im - np.floor(h_im), w_im - np.floor(w_im)) if ok else (0.0, 0.0)
    hlo, wlo, hhi, whi = h0 >= 0, w0 >= 0, h0 + 1 <= H - 1, w0 + 1 <= W - 1
    bits = [ok and hlo and wlo, ok and hlo and whi, ok and hhi and wlo, ok and hhi and whi]
    return dict(pix=start + h0 * W + w0, bits=bits, lh=lh, lw=lw, a=a)


def emulate_pair(value_nm, g, levels, starts, loc_or_off, attn_or_logits, ref=None, ref_dim=0):
    """One (query, head) pair on 4 lanes.  value_nm (S, 32) of this image and head; g (32,) grad_out; unfused:
    loc (L, P, 2), attn (L, P); fused: raw offsets (L, P, 2), logits (L*P,), ref (L, ref_dim).
    -> grad_value contribution (S, 32), grad wrt loc/offsets (L, P, 2), grad wrt attn/logits (L, P)."""
    L, P = len(levels), 4
    LP = L * P
    fused = ref is not None
    S = value_nm.shape[0]
    gv = np.zeros((S, 32))
    gl = np.zeros((L, P, 2))
    ga_out = np.zeros((L, P))
    lanes = range(4)
    for c0 in range(0, LP, 16):
        prep = [[None] * 4 for _ in lanes]
        sx, sy = [0.0] * 4, [0.0] * 4
        if fused:   # softmax over the pair's logits: group4_max / group4_sum see every lane's four values
            lg = [[attn_or_logits[4 * (c0 // 4 + j) + r] if c0 // 4 + j < L else -np.inf for r in range(4)] for j in lanes]
            mx = max(max(row) for row in lg)
            e = [[np.exp(v - mx) for v in row] for row in lg]
            inv = 1.0 / sum(sum(row) for row in e)
        for j in lanes:                                   # lane j prepares the 4 points of level c0/4 + j
            lv = c0 // 4 + j
            on = lv < L
            lvc = min(lv, L - 1)
            H, W = levels[lvc]
            for r in range(4):
                if not on:
                    x = y = a = 0.0
                elif fused:
                    ox, oy = loc_or_off[lv, r]
                    a = e[j][r] * inv
                    if ref_dim == 2:
                        sx[j], sy[j] = 1.0 / W, 1.0 / H
                        x, y = ref[lv, 0] + ox / W, ref[lv, 1] + oy / H
                    else:
                        sx[j], sy[j] = ref[lv, 2] * 0.5 / P, ref[lv, 3] * 0.5 / P
                        x, y = ref[lv, 0] + ox / P * ref[lv, 2] * 0.5, ref[lv, 1] + oy / P * ref[lv, 3] * 0.5
                else:
                    (x, y), a = loc_or_off[lv, r], attn_or_logits[lv, r]
                prep[j][r] = _prep(x, y, a, H, W, starts[lvc])
        nb = min(4, (LP - c0) // 4)
        sm_a = [[0.0] * 4 for _ in lanes]
        sm_g = [[0.0] * 4 for _ in lanes]
        for b in range(nb):                               # batch b = the points lane b prepared
            lvl = min(c0 // 4 + b, L - 1)
            H, W = levels[lvl]
            d = np.zeros((4, 4, 4))                       # [lane][point r][corner k]
            for r in range(4):
                p = prep[b][r]                            # __shfl_sync(..., b, 4): every lane reads lane b's registers
                hh, hw = 1.0 - p["lh"], 1.0 - p["lw"]
                w = [hh * hw, hh * p["lw"], p["lh"] * hw, p["lh"] * p["lw"]]
                pix = [p["pix"], p["pix"] + 1, p["pix"] + W, p["pix"] + W + 1]
                for j in lanes:
                    ch = slice(8 * j, 8 * j + 8)
                    for k in range(4):
                        if p["bits"][k]:
                            gv[pix[k], ch] += w[k] * p["a"] * g[ch]                 # red_add_f8_if
                            d[j, r, k] = np.dot(g[ch], value_nm[pix[k], ch])       # dot8
            # reduce-scatter: stage 1 exchanges with lane j^2, stage 2 with lane j^1
            ee = np.zeros((4, 2, 4))
            for j in lanes:
                hi2 = bool(j & 2)
                for r in range(2):
                    keep = d[j, r + 2] if hi2 else d[j, r]
                    # the partner's rule decides what arrives: a hi2 lane sends d[r], the others d[r + 2]
                    partner_hi2 = bool((j ^ 2) & 2)
                    send_partner = d[j ^ 2, r] if partner_hi2 else d[j ^ 2, r + 2]
                    ee[j, r] = keep + send_partner
            f = np.zeros((4, 4))
            for j in lanes:
                hi1 = bool(j & 1)
                keep = ee[j, 1] if hi1 else ee[j, 0]
                partner_hi1 = bool((j ^ 1) & 1)
                send_partner = ee[j ^ 1, 0] if partner_hi1 else ee[j ^ 1, 1]
                f[j] = keep + send_partner
            for j in lanes:                               # lane j finalises point j of the batch
                p = prep[b][j]
                hh, hw = 1.0 - p["lh"], 1.0 - p["lw"]
                gx = W * p["a"] * (hh * (f[j, 1] - f[j, 0]) + p["lh"] * (f[j, 3] - f[j, 2]))
                gy = H * p["a"] * (hw * (f[j, 2] - f[j, 0]) + p["lw"] * (f[j, 3] - f[j, 1]))
                ga = hh * (hw * f[j, 0] + p["lw"] * f[j, 1]) + p["lh"] * (hw * f[j, 2] + p["lw"] * f[j, 3])
                point = c0 + 4 * b + j
                if fused:
                    gx, gy = gx * sx[b], gy * sy[b]       # __shfl_sync(sx, b, 4)
                    sm_a[j][b], sm_g[j][b] = p["a"], ga
                else:
                    ga_out[point // P, point % P] = ga
                gl[point // P, point % P] = (gx, gy)
        if fused:
            dotp = sum(sm_a[j][b] * sm_g[j][b] for j in lanes for b in range(4))   # group4_sum
            for j in lanes:
                for b in range(nb):
                    point = c0 + 4 * b + j
                    ga_out[point // P, point % P] = sm_a[j][b] * (sm_g[j][b] - dotp)
    return gv, gl, ga_out


@pytest.mark.parametrize("levels", [[(7, 9), (4, 5), (2, 3), (1, 2)], [(9, 11), (5, 6), (3, 3), (2, 2), (1, 1)], [(6, 5)]],
                         ids=["4lvl", "5lvl", "1lvl"])
def test_unfused_lane_algorithm_equals_the_oracle(levels):
    rng = np.random.default_rng(len(levels))
    L, P, M, D, N, Lq = len(levels), 4, 2, 32, 1, 6
    S = sum(h * w for h, w in levels)
    starts = np.concatenate([[0], np.cumsum([h * w for h, w in levels])[:-1]]).astype(np.int64)
    value = rng.standard_normal((N, S, M, D))
    loc = rng.uniform(-0.2, 1.2, (N, Lq, M, L, P, 2))          # incl. borders and out-of-range samples
    attn = rng.uniform(0, 1, (N, Lq, M, L, P))
    gout = rng.standard_normal((N, Lq, M * D))
    gv, gl, ga = O.msda_backward(value, levels, starts, loc, attn, gout)
    got_v = np.zeros_like(value)
    for q in range(Lq):
        for m in range(M):
            v, l, a = emulate_pair(value[0, :, m], gout[0, q, m * D:(m + 1) * D], levels, starts, loc[0, q, m], attn[0, q, m])
            got_v[0, :, m] += v
            np.testing.assert_allclose(l, gl[0, q, m], rtol=1e-9, atol=1e-10)
            np.testing.assert_allclose(a, ga[0, q, m], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(got_v, gv, rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_fused_lane_algorithm_equals_autograd_through_the_module_math(ref_dim):
    """Fused form: gradients w.r.t. the raw offsets and logits equal autograd through the module's own arithmetic
    (ms_deform_attn.py:98-112: softmax over L*P, loc = ref + off / (W, H) or ref_xy + off / P * ref_wh * 0.5)."""
    levels = [(7, 9), (4, 5), (2, 3), (1, 2)]
    g = torch.Generator().manual_seed(ref_dim)
    L, P, M, D, Lq = 4, 4, 2, 32, 5
    S = sum(h * w for h, w in levels)
    starts = torch.tensor(np.concatenate([[0], np.cumsum([h * w for h, w in levels])[:-1]]))
    value = torch.randn(1, S, M, D, generator=g, dtype=torch.float64)
    off = (torch.randn(1, Lq, M, L, P, 2, generator=g, dtype=torch.float64) * 2).requires_grad_(True)
    logits = torch.randn(1, Lq, M, L * P, generator=g, dtype=torch.float64).requires_grad_(True)
    if ref_dim == 2:
        ref = torch.rand(1, Lq, L, 2, generator=g, dtype=torch.float64)
        wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float64)
        loc = ref[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
    else:
        ref = torch.cat([torch.rand(1, Lq, L, 2, generator=g, dtype=torch.float64),
                         torch.rand(1, Lq, L, 2, generator=g, dtype=torch.float64) * 0.4 + 0.05], -1)
        loc = ref[:, :, None, :, None, :2] + off / P * ref[:, :, None, :, None, 2:] * 0.5
    w = torch.softmax(logits, -1).view(1, Lq, M, L, P)
    out = O.msda_forward_torch(value.requires_grad_(True), levels, loc, w)
    gout = torch.randn(out.shape, generator=g, dtype=torch.float64)
    out.backward(gout)
    got_v = np.zeros((S, M, D))
    for q in range(Lq):
        for m in range(M):
            v, l, a = emulate_pair(value[0, :, m].detach().numpy(), gout[0, q, m * D:(m + 1) * D].numpy(), levels,
                                   starts.numpy(), off[0, q, m].detach().numpy(), logits[0, q, m].detach().numpy(),
                                   ref=ref[0, q].numpy(), ref_dim=ref_dim)
            got_v[:, m] += v
            np.testing.assert_allclose(l, off.grad[0, q, m].numpy(), rtol=1e-7, atol=1e-9)
            np.testing.assert_allclose(a.reshape(-1), logits.grad[0, q, m].numpy(), rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(got_v, value.grad[0].numpy(), rtol=1e-7, atol=1e-9)
