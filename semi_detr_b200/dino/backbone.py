"""ResNet backbone with the mmdet constructor arguments the configs pass
(configs/dino_detr/dino_detr_r50_8x2_12e_coco.py:9-17: depth 50, out_indices (1,2,3), frozen_stages 1,
BN frozen + norm_eval, style 'pytorch').  Parameter names follow torchvision / mmdet (``conv1``, ``bn1``,
``layerN.M.convK`` ...) so ``torchvision://resnet50`` checkpoints load.  Dense convolutions go to cuDNN (library
plumbing): the backbone is not part of the graded hot path but is needed to run the train step end to end.
"""
import weakref

import torch
import torch.nn.functional as F
from torch import nn

from ..registry import BACKBONES


class _ConvBiasAct(torch.autograd.Function):
    """relu(conv(x, w) + b [+ z]) as ONE cuDNN fused convolution-bias-activation call (``cudnn_convolution_relu`` /
    ``cudnn_convolution_add_relu``: library plumbing, like the convolution itself) instead of a convolution and two
    or three elementwise passes over the activation; backward = ReLU mask + the library's convolution backward."""

    @staticmethod
    def forward(ctx, x, w, b, z, stride, padding, dilation, groups):
        if z is None:
            y = torch.cudnn_convolution_relu(x, w, b, stride, padding, dilation, groups)
        else:
            y = torch.cudnn_convolution_add_relu(x, w, z, 1.0, b, stride, padding, dilation, groups)
        ctx.save_for_backward(x, w, y)
        ctx.conf = (stride, padding, dilation, groups, z is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        stride, padding, dilation, groups, has_z = ctx.conf
        g = torch.ops.aten.threshold_backward(gy, y, 0.0)
        gx = gw = gb = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            gx, gw, _ = torch.ops.aten.convolution_backward(
                g, x, w, None, list(stride), list(padding), list(dilation), False, [0, 0], groups,
                [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        if ctx.needs_input_grad[2]:
            gb = g.sum((0, 2, 3))
        gz = g if (has_z and ctx.needs_input_grad[3]) else None
        return gx, gw, gb, gz, None, None, None, None


def fused_conv_enabled():
    import os
    return os.environ.get("SDB_FUSED_CONV", "1") != "0"


def enable_frozen_bn_fold_cache(module, enabled=True):
    """Opt the BatchNorm layers of ``module`` into caching their folded affine map (``folded_conv``).  Call it for a
    model whose frozen BN tensors are written only through torch ops on the tensors themselves (``load_state_dict``,
    ``copy_`` ... bump the version counters the cache is keyed on; writes through ``.data`` or raw pointers do not)
    -- the train-step engines do.  Do NOT enable it for an EMA teacher: its tensors are
    rewritten by ``sdb_ema_update_f32`` through raw pointers, which no version counter sees."""
    for m in module.modules():
        if isinstance(m, nn.BatchNorm2d):
            (_FOLD_ENABLED.add if enabled else _FOLD_ENABLED.discard)(m)
            _FOLD_CACHE.pop(m, None)
    return module


# Side tables, not module attributes: a deepcopy / pickle of the model (e.g. to make a teacher) neither inherits the
# opt-in nor carries cached tensors.
_FOLD_ENABLED = weakref.WeakSet()
_FOLD_CACHE = weakref.WeakKeyDictionary()


def fold_cache_enabled(bn):
    return bn in _FOLD_ENABLED


def _fold_key(*tensors):
    return tuple((t.data_ptr(), t._version) for t in tensors)


def folded_conv(conv, bn):
    """-> (w, t): conv(x, w) + t == BN_eval(conv(x)), with s = gamma / sqrt(var + eps), w = weight * s, t = beta - mean * s.

    With frozen BN parameters (``requires_grad=False``; every shipped config) ``s`` and ``t`` are the same numbers every
    step, and so is ``w`` for a frozen convolution (stem + ``frozen_stages``): recomputing them cost 6 tiny kernels per
    convolution per step (53 convolutions -> ~300 of the step's ~2 700 launches).  When the layer has been opted in
    (``enable_frozen_bn_fold_cache``) they are computed once and reused until one of the tensors is written (data
    pointer + version counter).  An existing entry is used during CUDA-graph capture (it is ordinary device memory,
    like the parameters); a new one is never created there (it would live in the graph's private pool).  Nothing that
    requires grad is cached."""
    cacheable = bn in _FOLD_ENABLED and not bn.weight.requires_grad and not bn.bias.requires_grad
    key = None
    if cacheable:
        key = _fold_key(bn.weight, bn.bias, bn.running_mean, bn.running_var)
        hit = _FOLD_CACHE.get(bn)
        if hit is not None and hit[0] == key:
            s, t, w, wkey = hit[1]
            if w is not None and wkey == _fold_key(conv.weight):
                return w, t
            return conv.weight * s.view(-1, 1, 1, 1), t
    s = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
    t = bn.bias - bn.running_mean * s
    w = conv.weight * s.view(-1, 1, 1, 1)
    if cacheable and not (s.is_cuda and torch.cuda.is_current_stream_capturing()):
        frozen_w = not conv.weight.requires_grad
        _FOLD_CACHE[bn] = (key, (s, t, w if frozen_w else None, _fold_key(conv.weight) if frozen_w else None))
    return w, t


def conv_bn_relu(conv, bn, x, identity=None):
    """relu(BN(conv(x)) [+ identity]) -- with frozen statistics on the device: one fused cuDNN call."""
    if bn.training or not x.is_cuda or not fused_conv_enabled() or torch.is_autocast_enabled():
        out = conv_bn(conv, bn, x)      # (under autocast the library casts per op; the fused call takes one dtype)
        return F.relu(out if identity is None else out + identity, inplace=True)
    w, t = folded_conv(conv, bn)
    args = (tuple(conv.stride), tuple(conv.padding), tuple(conv.dilation), conv.groups)
    if not (torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or
                                         (identity is not None and identity.requires_grad))):
        if identity is None:
            return torch.cudnn_convolution_relu(x, w, t, *args)
        return torch.cudnn_convolution_add_relu(x, w, identity, 1.0, t, *args)
    return _ConvBiasAct.apply(x, w, t, identity, *args)


def conv_bn(conv, bn, x):
    """conv -> BatchNorm.  With the statistics frozen (``norm_eval``; every shipped config) the normalisation is a
    per-channel affine map that folds exactly into the convolution: conv(x, w * s) + t with s = gamma / sqrt(var +
    eps), t = beta - mean * s.  That removes one elementwise pass over every backbone activation in forward and
    backward (53 BN launches per step); gradients still reach ``w`` through the scale.  In training-statistics mode
    the two modules run as written."""
    if bn.training or not x.is_cuda:
        return bn(conv(x))
    w, t = folded_conv(conv, bn)
    return F.conv2d(x, w, t, conv.stride, conv.padding, conv.dilation, conv.groups)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)   # style='pytorch'
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        identity = x if self.downsample is None else conv_bn(self.downsample[0], self.downsample[1], x)
        out = conv_bn_relu(self.conv1, self.bn1, x)
        out = conv_bn_relu(self.conv2, self.bn2, out)
        return conv_bn_relu(self.conv3, self.bn3, out, identity)


@BACKBONES.register_module()
class ResNet(nn.Module):
    arch = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}

    def __init__(self, depth=50, num_stages=4, out_indices=(1, 2, 3), frozen_stages=1,
                 norm_cfg=dict(type="BN", requires_grad=False), norm_eval=True, style="pytorch", init_cfg=None):
        super().__init__()
        assert depth in self.arch and style == "pytorch" and num_stages == 4
        self.out_indices, self.frozen_stages, self.norm_eval = tuple(out_indices), frozen_stages, norm_eval
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        inplanes = 64
        for i, blocks in enumerate(self.arch[depth]):
            planes, stride = 64 * 2 ** i, 1 if i == 0 else 2
            down = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False),
                                 nn.BatchNorm2d(planes * 4))
            layers = [Bottleneck(inplanes, planes, stride, down)]
            inplanes = planes * 4
            layers += [Bottleneck(inplanes, planes) for _ in range(1, blocks)]
            setattr(self, f"layer{i + 1}", nn.Sequential(*layers))
        if not norm_cfg.get("requires_grad", True):
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    for p in m.parameters():
                        p.requires_grad = False
        self._freeze_stages()

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            self.bn1.eval()
            for m in (self.conv1, self.bn1):
                for p in m.parameters():
                    p.requires_grad = False
        for i in range(1, self.frozen_stages + 1):
            m = getattr(self, f"layer{i}")
            m.eval()
            for p in m.parameters():
                p.requires_grad = False

    def train(self, mode=True):
        super().train(mode)
        self._freeze_stages()
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()
        return self

    def forward(self, x, cut_before_last_stage=None):
        """``cut_before_last_stage`` (a list, optional): layer4 then runs on a detached copy of layer3's output and the
        pair (layer3 output, detached copy) is appended -- the engine's segmented backward differentiates the two halves
        of the backbone separately so that layer4's gradients can be exchanged while layer2 / layer3 still compute."""
        x = self.maxpool(conv_bn_relu(self.conv1, self.bn1, x))
        outs = []
        for i in range(4):
            if i == 3 and cut_before_last_stage is not None and x.requires_grad:
                leaf = x.detach().requires_grad_(True)
                cut_before_last_stage.append((x, leaf))
                x = leaf
            x = getattr(self, f"layer{i + 1}")(x)
            if i in self.out_indices:
                outs.append(x)
        return tuple(outs)
