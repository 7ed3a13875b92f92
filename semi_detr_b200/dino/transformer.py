"""DINO deformable transformer -- host-side mirror of detr_od/models/utils/transformer.py:435-1406.

Same module tree and parameter names as the reference (``encoder.layers.N.self_attn.*``,
``decoder.layers.N.cross_attn.*``, ``decoder.ref_point_head``, ``level_embed``, ``tgt_embed``, ``enc_output`` ...)
so reference checkpoints load; every ``MSDeformAttn`` runs on the sm_100a kernels.  Only the configuration the
shipped configs use is implemented (two_stage_type='standard', deformable encoder+decoder, 'sa'->'ca'->'ffn',
learnable tgt, no query scale); anything else raises.

Host-side differences that do not change numerics:
 * encoder reference points / level geometry are built once per distinct (shapes, valid ratios) instead of
   per-level python meshgrids each call;
 * the decoder's sine embedding table (``dim_t``) is cached.
"""
import copy
import math
import os
from collections import OrderedDict

import torch
import torch.nn.functional as F
from torch import nn

from ..consts import device_const
from ..msda import MSDeformAttn
from ..registry import TRANSFORMER
from ..layers import LayerNorm, Linear, ffn
from ..layers.attention import attention_enabled, fused_self_attention
from ..layers.linear import linear


def inverse_sigmoid(x, eps=1e-5):
    """transformer.py:435-451"""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


class MLP(nn.Module):
    """transformer.py:453-465"""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = layer(x, relu=i < self.num_layers - 1)
        return x


_DIM_T = {}
_GEOMETRY = OrderedDict()     # (device, host geometry key, level shapes) -> mask-derived tensors (_mask_geometry)

# Decoder self-attention runs over ~1.1 k queries in fp32: for that size plain matmul-softmax-matmul beats the fused
# memory-efficient fp32 kernel torch's SDPA picks on sm_100 (measured in profiles/), so the layer writes it out
# (DINOTransformerDecoderLayer._self_attention).


def gen_sineembed_for_position(pos):
    """(nq, bs, 2|4) in [0,1] -> (nq, bs, 256|512); 128 dims per coordinate, temperature 10000, order y,x,w,h
    (transformer.py:467-493).

    The reference embeds each coordinate on its own (scale, divide, ``sin`` of the even slots, ``cos`` of the odd
    ones, stack, flatten: ~6 small kernels per coordinate, 25 per call, 6 calls per pass).  Here all coordinates go
    through ONE scale / divide / sin / cos / select: the same elementwise values (``sin`` and ``cos`` of the same
    quotient, picked by slot parity), 6 kernels per call."""
    n = pos.size(-1)
    if n not in (2, 4):
        raise ValueError("Unknown pos_tensor shape(-1):{}".format(n))
    dev = pos.device
    key = str(dev)
    if key not in _DIM_T:
        i = torch.arange(128, dtype=torch.float32, device=dev)
        yx = torch.arange(1, -1, -1, device=dev)                  # device-side construction only: stays capturable
        _DIM_T[key] = (10000 ** (2 * (i // 2) / 128), (torch.arange(128, device=dev) % 2) == 0,
                       {2: yx, 4: torch.cat([yx, torch.arange(2, 4, device=dev)])})
    dim_t, even, order = _DIM_T[key]
    p = pos.index_select(-1, order[n])                                  # y, x(, w, h)
    e = (p * (2 * math.pi))[..., None] / dim_t                          # (nq, bs, n, 128)
    return torch.where(even, e.sin(), e.cos()).flatten(2)


def encoder_proposals(memory_padding_mask, spatial_shapes_list):
    """The mask-only part of the two-stage proposals (transformer.py:525-575): grid centre / valid size with
    0.05*2^l boxes in logit space, +inf where the row is padded or a coordinate leaves (0.01, 0.99)
    -> (proposals (N, S, 4), bad (N, S, 1) bool).  Depends on the padding masks alone, i.e. on the batch geometry."""
    N = memory_padding_mask.shape[0]
    device = memory_padding_mask.device
    proposals = []
    cur = 0
    for lvl, (H, W) in enumerate(spatial_shapes_list):
        m = memory_padding_mask[:, cur:cur + H * W].view(N, H, W)
        valid_h = (~m[:, :, 0]).sum(1)
        valid_w = (~m[:, 0, :]).sum(1)
        def build_grid(H=H, W=W):
            gy, gx = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=device),
                                    torch.arange(W, dtype=torch.float32, device=device), indexing="ij")
            return torch.stack([gx, gy], -1)
        grid = device_const(device, "proposal_grid", (H, W), build_grid)
        scale = torch.stack([valid_w, valid_h], 1).view(N, 1, 1, 2)
        grid = (grid[None].expand(N, -1, -1, -1) + 0.5) / scale
        wh = torch.ones_like(grid) * 0.05 * (2.0 ** lvl)
        proposals.append(torch.cat((grid, wh), -1).view(N, -1, 4))
        cur += H * W
    prop = torch.cat(proposals, 1)
    valid = ((prop > 0.01) & (prop < 0.99)).all(-1, keepdim=True)
    prop = torch.log(prop / (1 - prop))
    bad = memory_padding_mask.unsqueeze(-1) | ~valid
    return prop.masked_fill(bad, float("inf")), bad


def gen_encoder_output_proposals(memory, memory_padding_mask, spatial_shapes_list):
    """Two-stage proposals (transformer.py:525-575): ``encoder_proposals`` + the memory rows of padded / invalid
    positions zeroed.  ``spatial_shapes_list`` is the host list [(H, W), ...] (no device sync)."""
    prop, bad = encoder_proposals(memory_padding_mask, spatial_shapes_list)
    return memory.masked_fill(bad, 0.0), prop


class DINOTransformerEncoderLayer(nn.Module):
    """MSDA self-attention + FFN, post-norm (transformer.py:578-642)."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        assert activation == "relu"
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = LayerNorm(d_model)
        self.linear1 = Linear(d_model, d_ffn)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = LayerNorm(d_model)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, key_padding_mask=None,
                query=None, next_pos=None):
        """``query``: ``src + pos`` if the caller already has it (the previous layer's second output); ``next_pos``: also
        return ``out + next_pos`` -- both additions then ride in the LayerNorm kernels (``LayerNorm.add_norm``)."""
        q = query if query is not None else (src if pos is None else src + pos)
        src2 = self.self_attn(q, reference_points, src, spatial_shapes, level_start_index, key_padding_mask)
        if ffn.fused_ok(src, self.norm1, self.linear1, self.linear2, self.norm2,
                        (self.dropout1, self.dropout2, self.dropout3)):
            # norm1(src + src2) -> FFN -> norm2 as one autograd node (layers/ffn.py)
            return ffn.post_attention_block(src, src2, self.norm1, self.linear1, self.linear2, self.norm2, next_pos)
        src = self.norm1.add_norm(src, self.dropout1(src2))
        src2 = self.linear2(self.dropout2(self.linear1(src, relu=True)))
        return self.norm2.add_norm(src, self.dropout3(src2), next_pos)


class DINOTransformerEncoder(nn.Module):
    """transformer.py:644-744"""

    def __init__(self, encoder_layer, num_layers, norm=None, d_model=256):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = norm
        self.d_model = d_model

    @staticmethod
    def get_reference_points(spatial_shapes_list, valid_ratios, device):
        """Pixel centres / (valid_ratio * size), then scaled by every level's valid ratio (transformer.py:676-691)."""
        refs = []
        for lvl, (H, W) in enumerate(spatial_shapes_list):
            def grid(H=H, W=W):
                ry, rx = torch.meshgrid(torch.linspace(0.5, H - 0.5, H, dtype=torch.float32, device=device),
                                        torch.linspace(0.5, W - 0.5, W, dtype=torch.float32, device=device),
                                        indexing="ij")
                return torch.stack((rx.reshape(-1), ry.reshape(-1)))
            g = device_const(device, "enc_ref_grid", (H, W), grid)
            ry = g[1][None] / (valid_ratios[:, None, lvl, 1] * H)
            rx = g[0][None] / (valid_ratios[:, None, lvl, 0] * W)
            refs.append(torch.stack((rx, ry), -1))
        ref = torch.cat(refs, 1)
        return ref[:, :, None] * valid_ratios[:, None]

    def forward(self, src, pos, spatial_shapes, level_start_index, valid_ratios, key_padding_mask,
                spatial_shapes_list, reference_points=None):
        out = src
        if reference_points is None:
            reference_points = self.get_reference_points(spatial_shapes_list, valid_ratios, src.device)
        query = None
        for i, layer in enumerate(self.layers):
            # every layer but the last also emits the next layer's query (out + pos) from its final LayerNorm pass
            nxt = pos if (pos is not None and i + 1 < len(self.layers)) else None
            res = layer(out, pos, reference_points, spatial_shapes, level_start_index, key_padding_mask, query=query,
                        next_pos=nxt)
            out, query = res if nxt is not None else (res, None)
        if self.norm is not None:
            out = self.norm(out)
        return out


class DINOTransformerDecoderLayer(nn.Module):
    """self-attention (nn.MultiheadAttention) -> MSDA cross-attention -> FFN, post-norm (transformer.py:746-873)."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4,
                 decoder_sa_type="sa", module_seq=("sa", "ca", "ffn")):
        super().__init__()
        assert activation == "relu" and decoder_sa_type == "sa"
        assert sorted(module_seq) == ["ca", "ffn", "sa"]
        self.module_seq = list(module_seq)
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = LayerNorm(d_model)
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = LayerNorm(d_model)
        self.linear1 = Linear(d_model, d_ffn)
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = LayerNorm(d_model)

    def _self_attention(self, qk, value, attn_mask, attn_mask_t=None):
        """``self.self_attn(qk, qk, value, attn_mask=attn_mask)[0]`` (nn.MultiheadAttention, transformer.py:795-803)
        written out on the module's own parameters: q and k share one projection GEMM, the (T, T) mask is added by
        the ``baddbmm`` that forms the scores instead of a separate pass over the (N*H, T, T) tensor, and the
        fully-masked-row guard of torch's math SDPA (three more passes) is dropped -- the denoising mask never
        masks a whole row (dn_components.py:97-113).  Scaling follows the reference's torch (q * d^-0.5 first)."""
        mha = self.self_attn
        T, N, C = qk.shape
        H = mha.num_heads
        d = C // H
        w, b = mha.in_proj_weight, mha.in_proj_bias
        if (attention_enabled(qk, d) and not (self.training and mha.dropout > 0)
                and (attn_mask is None or (attn_mask.dtype == torch.float32 and attn_mask.dim() == 2))):
            # scores, mask, softmax and the value product in one kernel each way (csrc/attention.cu): q and k are read
            # in place from the shared projection's output, no (N*H, T, T) tensor exists
            if attn_mask is not None and attn_mask_t is None:
                attn_mask_t = attn_mask.t().contiguous()
            out = fused_self_attention(linear(qk, w[:2 * C], b[:2 * C]), linear(value, w[2 * C:], b[2 * C:]),
                                       attn_mask, attn_mask_t, H)
            return linear(out, mha.out_proj.weight, mha.out_proj.bias)
        q, k = F.linear(qk, w[:2 * C], b[:2 * C]).split(C, -1)
        v = F.linear(value, w[2 * C:], b[2 * C:])
        q = (q * (float(d) ** -0.5)).reshape(T, N * H, d).transpose(0, 1)
        k = k.reshape(T, N * H, d).transpose(0, 1)
        v = v.reshape(T, N * H, d).transpose(0, 1)
        if attn_mask is None:
            scores = torch.bmm(q, k.transpose(1, 2))
        else:
            if attn_mask.dtype == torch.bool:
                attn_mask = torch.zeros(attn_mask.shape, dtype=q.dtype, device=q.device).masked_fill_(
                    attn_mask, float("-inf"))
            scores = torch.baddbmm(attn_mask, q, k.transpose(1, 2))
        probs = F.dropout(F.softmax(scores, -1), mha.dropout, self.training)
        out = torch.bmm(probs, v).transpose(0, 1).reshape(T, N, C)
        return F.linear(out, mha.out_proj.weight, mha.out_proj.bias)

    def forward(self, tgt, query_pos, reference_points, memory, memory_key_padding_mask, level_start_index,
                spatial_shapes, self_attn_mask=None, self_attn_mask_t=None):
        """tgt / query_pos (nq, bs, C); reference_points (nq, bs, L, 4); memory (bs, S, C) batch-first."""
        seq, i = self.module_seq, 0
        while i < len(seq):
            name = seq[i]
            if name == "sa":
                tgt2 = self._self_attention(tgt + query_pos, tgt, self_attn_mask, self_attn_mask_t)
                tgt = self.norm2.add_norm(tgt, self.dropout2(tgt2))
            elif name == "ca":
                tgt2 = self.cross_attn((tgt + query_pos).transpose(0, 1), reference_points.transpose(0, 1).contiguous(),
                                       memory, spatial_shapes, level_start_index,
                                       memory_key_padding_mask).transpose(0, 1)
                if (i + 1 < len(seq) and seq[i + 1] == "ffn" and
                        ffn.fused_ok(tgt, self.norm1, self.linear1, self.linear2, self.norm3,
                                     (self.dropout1, self.dropout3, self.dropout4))):
                    # norm1(tgt + tgt2) -> FFN -> norm3 as one autograd node (layers/ffn.py)
                    tgt = ffn.post_attention_block(tgt, tgt2, self.norm1, self.linear1, self.linear2, self.norm3)
                    i += 2
                    continue
                tgt = self.norm1.add_norm(tgt, self.dropout1(tgt2))
            else:
                tgt2 = self.linear2(self.dropout3(self.linear1(tgt, relu=True)))
                tgt = self.norm3.add_norm(tgt, self.dropout4(tgt2))
            i += 1
        return tgt


class DINOTransformerDecoder(nn.Module):
    """Iterative box refinement decoder (transformer.py:875-1045)."""

    def __init__(self, decoder_layer, num_layers, norm, d_model=256, query_dim=4, num_feature_levels=4):
        super().__init__()
        assert query_dim == 4
        self.layers = nn.ModuleList([copy.deepcopy(decoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = norm
        self.query_dim = query_dim
        self.num_feature_levels = num_feature_levels
        self.ref_point_head = MLP(query_dim // 2 * d_model, d_model, d_model, 2)
        self.d_model = d_model

    def forward(self, tgt, memory, tgt_mask, memory_key_padding_mask, refpoints_unsigmoid, level_start_index,
                spatial_shapes, valid_ratios, fc_reg):
        """tgt (nq, bs, C); memory (bs, S, C); refpoints_unsigmoid (nq, bs, 4)
        -> ([n_dec x (bs, nq, C)], [(n_dec+1) x (bs, nq, 4)])"""
        output = tgt
        intermediate = []
        reference_points = refpoints_unsigmoid.sigmoid()
        ref_points = [reference_points]
        vr4 = torch.cat([valid_ratios, valid_ratios], -1)[None]          # (1, bs, L, 4)
        if tgt_mask is not None and tgt_mask.dtype == torch.bool:
            # the additive form every layer's score product needs: built once per pass, not once per layer
            tgt_mask = torch.zeros(tgt_mask.shape, dtype=tgt.dtype, device=tgt.device).masked_fill_(tgt_mask,
                                                                                                  float("-inf"))
        # the key-major copy the attention backward reads (csrc/attention.cu), also once per pass
        tgt_mask_t = tgt_mask.t().contiguous() if (tgt_mask is not None and tgt_mask.dim() == 2) else None
        for lid, layer in enumerate(self.layers):
            ref_in = reference_points[:, :, None] * vr4                   # (nq, bs, L, 4)
            query_pos = self.ref_point_head(gen_sineembed_for_position(ref_in[:, :, 0, :]))
            output = layer(output, query_pos, ref_in, memory, memory_key_padding_mask, level_start_index,
                           spatial_shapes, self_attn_mask=tgt_mask, self_attn_mask_t=tgt_mask_t)
            if fc_reg is not None:
                new_ref = (fc_reg[lid](output) + inverse_sigmoid(reference_points)).sigmoid()
                reference_points = new_ref.detach()
                ref_points.append(new_ref)
            intermediate.append(self.norm(output))
        return [o.transpose(0, 1) for o in intermediate], [r.transpose(0, 1) for r in ref_points]


@TRANSFORMER.register_module()
class DINOTransformer(nn.Module):
    """transformer.py:1047-1406 with the defaults the configs rely on (``transformer=dict(type='DINOTransformer')``)."""

    def __init__(self, d_model=256, nhead=8, num_queries=900, num_encoder_layers=6, num_decoder_layers=6,
                 dim_feedforward=2048, dropout=0.0, activation="relu", normalize_before=False,
                 return_intermediate_dec=True, query_dim=4, num_feature_levels=4, enc_n_points=4, dec_n_points=4,
                 two_stage_type="standard", embed_init_tgt=True, decoder_sa_type="sa",
                 module_seq=("sa", "ca", "ffn"), **unsupported):
        super().__init__()
        for k, v in unsupported.items():
            if v not in (None, False, 0, True) and k not in ("modulate_hw_attn", "deformable_encoder",
                                                             "deformable_decoder", "learnable_tgt_init",
                                                             "rm_enc_query_scale", "rm_dec_query_scale"):
                raise NotImplementedError(f"DINOTransformer option {k}={v!r} is outside the shipped configs")
        assert two_stage_type == "standard" and query_dim == 4 and return_intermediate_dec and embed_init_tgt
        self.num_feature_levels = num_feature_levels
        self.num_encoder_layers = num_encoder_layers
        self.num_decoder_layers = num_decoder_layers
        self.num_queries = num_queries
        self.d_model = self.embed_dims = d_model
        self.nhead = nhead
        enc_layer = DINOTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels,
                                                nhead, enc_n_points)
        self.encoder = DINOTransformerEncoder(enc_layer, num_encoder_layers,
                                              LayerNorm(d_model) if normalize_before else None, d_model)
        dec_layer = DINOTransformerDecoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels,
                                                nhead, dec_n_points, decoder_sa_type, module_seq)
        self.decoder = DINOTransformerDecoder(dec_layer, num_decoder_layers, LayerNorm(d_model), d_model,
                                              query_dim, num_feature_levels)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model)) if num_feature_levels > 1 else None
        self.tgt_embed = nn.Embedding(num_queries, d_model)
        self.enc_output = Linear(d_model, d_model)
        self.enc_output_norm = LayerNorm(d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        if self.level_embed is not None:
            nn.init.normal_(self.level_embed)

    @staticmethod
    def get_valid_ratio(mask):
        _, H, W = mask.shape
        valid_h = (~mask[:, :, 0]).sum(1)
        valid_w = (~mask[:, 0, :]).sum(1)
        return torch.stack([valid_w.float() / W, valid_h.float() / H], -1)

    def _mask_geometry(self, masks, shapes_list, geometry_key):
        """Everything the pass derives from the padding masks alone: flattened mask, valid ratios, encoder reference
        points, two-stage proposals.  The masks are per-geometry constants (``dino/head.py`` builds them once per
        ``(batch_input_shape, img_shapes)``), so with that host key the derived tensors are built once as well (~110
        small launches per step otherwise); without a key they are computed as written."""
        def build():
            mask_flat = torch.cat([m.flatten(1) for m in masks], 1)
            valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)
            ref = DINOTransformerEncoder.get_reference_points(shapes_list, valid_ratios, mask_flat.device)
            prop, bad = encoder_proposals(mask_flat, shapes_list)
            return mask_flat, valid_ratios, ref, prop, bad
        if geometry_key is None:
            return build()
        k = (str(masks[0].device), geometry_key, tuple(shapes_list))
        hit = _GEOMETRY.get(k)
        if hit is None:
            hit = build()
            if not (masks[0].is_cuda and torch.cuda.is_current_stream_capturing()):   # never keep graph-pool memory
                _GEOMETRY[k] = hit
                if len(_GEOMETRY) > 64:
                    _GEOMETRY.popitem(last=False)
        else:
            _GEOMETRY.move_to_end(k)
        return hit

    def forward(self, srcs, masks, refpoint_embed, pos_embeds, tgt, attn_mask=None, fc_reg=None, fc_cls=None,
                fc_enc_reg=None, fc_enc_cls=None, geometry_key=None):
        """srcs / pos_embeds: L x (bs, C, H_l, W_l); masks: L x (bs, H_l, W_l) bool (True = padding);
        refpoint_embed (bs, n_dn, 4) / tgt (bs, n_dn, C): the denoising part or None
        -> hs [n_dec x (bs, nq, C)], references [(n_dec+1) x (bs, nq, 4)], hs_enc (1, bs, 900, C),
           ref_enc (1, bs, 900, 4), init_box_proposal (bs, 900, 4)"""
        src_l, pos_l, shapes_list = [], [], []
        for lvl, (src, mask, pos) in enumerate(zip(srcs, masks, pos_embeds)):
            bs, c, h, w = src.shape
            shapes_list.append((h, w))
            pos = pos.flatten(2).transpose(1, 2)
            if self.level_embed is not None:
                pos = pos + self.level_embed[lvl].view(1, 1, -1)
            src_l.append(src.flatten(2).transpose(1, 2))
            pos_l.append(pos)
        src_flat = torch.cat(src_l, 1)
        pos_flat = torch.cat(pos_l, 1)
        mask_flat, valid_ratios, enc_refs, output_proposals, bad_rows = self._mask_geometry(masks, shapes_list,
                                                                                            geometry_key)
        starts = [0]
        for h, w in shapes_list[:-1]:
            starts.append(starts[-1] + h * w)
        skey = tuple(shapes_list)
        spatial_shapes = device_const(src_flat.device, "spatial_shapes", skey,
                                      lambda: torch.as_tensor(shapes_list, dtype=torch.long))
        level_start_index = device_const(src_flat.device, "level_start", skey,
                                         lambda: torch.as_tensor(starts, dtype=torch.long))

        # A batch in which no image is padded (every img_shape equals the batch shape -- known on the host from the
        # metas) has an all-False padding mask: the MSDA value projections then skip the mask in both directions
        # (ms_deform_attn.py:96-97 would fill nothing).  Same numbers; SDB_SKIP_EMPTY_MASK=0 keeps the masked path.
        msda_mask = mask_flat
        if (geometry_key is not None and os.environ.get("SDB_SKIP_EMPTY_MASK", "1") != "0"
                and all(tuple(hw) == (geometry_key[0], geometry_key[1]) for hw in geometry_key[2])):
            msda_mask = None
        memory = self.encoder(src_flat, pos_flat, spatial_shapes, level_start_index, valid_ratios, msda_mask,
                              shapes_list, reference_points=enc_refs)

        # two-stage query selection (transformer.py:1314-1346; gen_encoder_output_proposals with its mask-only part
        # taken from the geometry cache)
        output_memory = memory.masked_fill(bad_rows, 0.0)
        output_memory = self.enc_output_norm(self.enc_output(output_memory))
        enc_cls = fc_enc_cls(output_memory)
        enc_coord = fc_enc_reg(output_memory) + output_proposals
        topk = torch.topk(enc_cls.max(-1)[0], self.num_queries, dim=1)[1]
        idx4 = topk.unsqueeze(-1).expand(-1, -1, 4)
        refpoint_undetach = torch.gather(enc_coord, 1, idx4)
        refpoint_sel = refpoint_undetach.detach()
        init_box_proposal = torch.gather(output_proposals, 1, idx4).sigmoid()
        tgt_undetach = torch.gather(output_memory, 1, topk.unsqueeze(-1).expand(-1, -1, self.d_model))
        bs = src_flat.shape[0]
        tgt_sel = self.tgt_embed.weight[:self.num_queries][None].expand(bs, -1, -1)
        if refpoint_embed is not None:
            refpoint_embed = torch.cat([refpoint_embed, refpoint_sel], dim=1)
            tgt = torch.cat([tgt, tgt_sel], dim=1)
        else:
            refpoint_embed, tgt = refpoint_sel, tgt_sel

        hs, references = self.decoder(tgt.transpose(0, 1), memory, attn_mask, msda_mask,
                                      refpoint_embed.transpose(0, 1), level_start_index, spatial_shapes,
                                      valid_ratios, fc_reg)
        hs_enc = tgt_undetach.unsqueeze(0)
        ref_enc = refpoint_undetach.sigmoid().unsqueeze(0)
        return hs, references, hs_enc, ref_enc, init_box_proposal
