"""Import pieces of the REFERENCE (read-only /root/reference) in the authoring container.

Used only by tests/golden/make_golden.py to generate golden vectors; never at test/bench time
(the GPU box has no /root/reference).  mmcv / mmdet cannot be imported as packages here
(mmcv-full 1.3.16 is not installable offline), so the handful of reference *files* on the hot
path are loaded individually with tiny stand-ins for the mmcv symbols they import
(Registry / build_from_cfg / Hook / is_module_wrapper).  No reference source is copied: the
files are exec'd from where they lie.
"""
import importlib.util
import sys
import types

REF = "/root/reference"
MMDET = REF + "/thirdparty/mmdetection/mmdet"


class _Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self.module_dict[key]


def _build_from_cfg(cfg, registry, default_args=None):
    cfg = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            cfg.setdefault(k, v)
    return registry.get(cfg.pop("type"))(**cfg)


def _pkg(name):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(_pkg(parent), child, m)
    return sys.modules[name]


def _load(name, path):
    if name in sys.modules and getattr(sys.modules[name], "__file__", None) == path:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    parent, child = name.rsplit(".", 1)
    setattr(_pkg(parent), child, mod)
    spec.loader.exec_module(mod)
    return mod


def _install_mmcv_stub():
    mmcv = _pkg("mmcv")
    utils = _pkg("mmcv.utils")
    utils.Registry = _Registry
    utils.build_from_cfg = _build_from_cfg
    par = _pkg("mmcv.parallel")
    par.is_module_wrapper = lambda m: False
    _pkg("mmcv.runner")
    hooks = _pkg("mmcv.runner.hooks")

    class Hook:
        pass
    hooks.Hook = Hook
    hooks.HOOKS = _Registry("hook")
    return mmcv


def load_msda_python():
    """functions/ms_deform_attn_func.py with a stub for the CUDA extension it imports at :18."""
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    return _load("ref_ops.ms_deform_attn_func",
                 REF + "/detr_od/models/utils/ops/functions/ms_deform_attn_func.py")


def load_hungarian():
    """The real mmdet HungarianAssigner + match costs (hungarian_assigner.py, match_cost.py,
    iou2d_calculator.py, transforms.py) on stubbed registries."""
    _install_mmcv_stub()
    for p in ("mmdet", "mmdet.core", "mmdet.core.bbox", "mmdet.core.bbox.iou_calculators",
              "mmdet.core.bbox.match_costs", "mmdet.core.bbox.assigners", "mmdet.utils"):
        _pkg(p)
    _load("mmdet.utils.util_mixins", MMDET + "/utils/util_mixins.py")
    b = MMDET + "/core/bbox"
    _load("mmdet.core.bbox.iou_calculators.builder", b + "/iou_calculators/builder.py")
    iou = _load("mmdet.core.bbox.iou_calculators.iou2d_calculator", b + "/iou_calculators/iou2d_calculator.py")
    sys.modules["mmdet.core.bbox.iou_calculators"].bbox_overlaps = iou.bbox_overlaps
    _load("mmdet.core.bbox.transforms", b + "/transforms.py")
    mcb = _load("mmdet.core.bbox.match_costs.builder", b + "/match_costs/builder.py")
    mc = _load("mmdet.core.bbox.match_costs.match_cost", b + "/match_costs/match_cost.py")
    sys.modules["mmdet.core.bbox.match_costs"].build_match_cost = mcb.build_match_cost
    _load("mmdet.core.bbox.builder", b + "/builder.py")
    _load("mmdet.core.bbox.assigners.assign_result", b + "/assigners/assign_result.py")
    _load("mmdet.core.bbox.assigners.base_assigner", b + "/assigners/base_assigner.py")
    lg = types.ModuleType("mmdet.core.bbox.assigners.logger")
    lg.log_image_with_boxes = lambda *a, **k: None
    sys.modules["mmdet.core.bbox.assigners.logger"] = lg
    ha = _load("mmdet.core.bbox.assigners.hungarian_assigner", b + "/assigners/hungarian_assigner.py")
    return ha, mc, iou


def load_mean_teacher():
    """detr_ssod/utils/hooks/mean_teacher.py on a stubbed mmcv Hook."""
    _install_mmcv_stub()
    for p in ("detr_ssod_ref", "detr_ssod_ref.utils", "detr_ssod_ref.utils.hooks"):
        _pkg(p)
    lg = types.ModuleType("detr_ssod_ref.utils.logger")
    lg.log_every_n = lambda *a, **k: None
    sys.modules["detr_ssod_ref.utils.logger"] = lg
    sys.modules["detr_ssod_ref.utils"].logger = lg
    return _load("detr_ssod_ref.utils.hooks.mean_teacher", REF + "/detr_ssod/utils/hooks/mean_teacher.py")


def load_dino_transformer():
    """detr_od/models/utils/transformer.py (the DINO classes, :435-1406) and the reference's own ``MSDeformAttn`` module
    (ops/modules/ms_deform_attn.py) from where they lie.  The file imports a long list of mmcv / timm / matplotlib
    symbols for its Swin helpers (dead code for DINO): those get inert stand-ins.  The compiled extension is stubbed,
    and ``MSDeformAttnFunction.apply`` is routed to the reference's own pure-PyTorch fallback
    (``ms_deform_attn_core_pytorch``, ms_deform_attn_func.py:41-61) so the module runs on the CPU."""
    import torch.nn as nn
    _install_mmcv_stub()

    class _Inert:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return a[0] if a else None

    def _decorator_factory(*a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda f: f

    for name in ("matplotlib", "matplotlib.pyplot", "timm", "timm.models", "timm.models.layers"):
        _pkg(name)
    sys.modules["timm.models.layers"].DropPath = nn.Identity
    sys.modules["timm.models.layers"].trunc_normal_ = nn.init.trunc_normal_
    runner = _pkg("mmcv.runner")
    runner.auto_fp16 = _decorator_factory
    runner.force_fp32 = _decorator_factory
    bm = _pkg("mmcv.runner.base_module")
    bm.BaseModule, bm.ModuleList, bm.Sequential = nn.Module, nn.ModuleList, nn.Sequential
    cnn = _pkg("mmcv.cnn")
    for n in ("build_activation_layer", "build_conv_layer", "build_norm_layer", "xavier_init"):
        setattr(cnn, n, _Inert())
    _pkg("mmcv.cnn.bricks")
    reg = _pkg("mmcv.cnn.bricks.registry")
    reg.TRANSFORMER_LAYER, reg.TRANSFORMER_LAYER_SEQUENCE, reg.ATTENTION = (_Registry("tl"), _Registry("tls"),
                                                                          _Registry("attn"))
    tr = _pkg("mmcv.cnn.bricks.transformer")
    for n in ("BaseTransformerLayer", "TransformerLayerSequence", "FFN", "MultiScaleDeformableAttention"):
        setattr(tr, n, type(n, (nn.Module,), {}))
    tr.build_transformer_layer_sequence = _Inert()
    _pkg("mmcv.cnn.bricks.drop").build_dropout = _Inert()
    utils = sys.modules["mmcv.utils"]
    utils.to_2tuple = lambda x: (x, x)
    utils.deprecated_api_warning = _decorator_factory
    _pkg("mmcv.ops")
    _pkg("mmcv.ops.multi_scale_deform_attn").MultiScaleDeformableAttention = tr.MultiScaleDeformableAttention
    for p in ("mmdet", "mmdet.models", "mmdet.models.utils"):
        _pkg(p)
    _pkg("mmdet.models.utils.builder").TRANSFORMER = _Registry("transformer")

    base = REF + "/detr_od/models/utils"
    for p in ("detr_od_ref", "detr_od_ref.models", "detr_od_ref.models.utils", "detr_od_ref.models.utils.ops",
              "detr_od_ref.models.utils.ops.functions", "detr_od_ref.models.utils.ops.modules"):
        _pkg(p)
    att = _pkg("detr_od_ref.models.utils.attention")
    att.MultiheadAttention = type("MultiheadAttention", (nn.Module,), {})
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    fn = _load("detr_od_ref.models.utils.ops.functions.ms_deform_attn_func",
               base + "/ops/functions/ms_deform_attn_func.py")
    sys.modules["detr_od_ref.models.utils.ops.functions"].MSDeformAttnFunction = fn.MSDeformAttnFunction

    def cpu_apply(value, shapes, level_start_index, sampling_locations, attention_weights, im2col_step):
        return fn.ms_deform_attn_core_pytorch(value, shapes, sampling_locations, attention_weights)
    fn.MSDeformAttnFunction.apply = staticmethod(cpu_apply)
    mod = _load("detr_od_ref.models.utils.ops.modules.ms_deform_attn", base + "/ops/modules/ms_deform_attn.py")
    sys.modules["detr_od_ref.models.utils.ops.modules"].MSDeformAttn = mod.MSDeformAttn
    return _load("detr_od_ref.models.utils.transformer", base + "/transformer.py"), mod


def load_dino_head():
    """detr_od/models/dense_heads/dino_detr_head.py (``DINODETRHead.loss`` and everything under it: ``loss_single``,
    ``get_targets``, ``_get_target_single`` / ``_get_target_single_dn``) with the REAL mmdet pieces it computes with --
    HungarianAssigner + match costs, PseudoSampler, FocalLoss (python branch), L1Loss, GIoULoss, bbox transforms,
    ``multi_apply`` -- loaded from thirdparty/mmdetection, and inert stand-ins for the rest (layer builders, the
    AnchorFreeHead base, registries).  ``reduce_mean`` is the single-process identity, as in mmdet without
    torch.distributed."""
    import functools

    import torch
    import torch.nn as nn
    ha, mc, iou = load_hungarian()
    T, _ = load_dino_transformer()
    mmcv = sys.modules["mmcv"]
    mmcv.jit = lambda *a, **k: (lambda f: f)
    ops = _pkg("mmcv.ops")
    ops.sigmoid_focal_loss = None                      # CUDA op; the CPU tensors below take py_sigmoid_focal_loss
    cnn = sys.modules["mmcv.cnn"]
    cnn.Conv2d, cnn.Linear = nn.Conv2d, nn.Linear
    cnn.bias_init_with_prob = lambda p: float(-__import__("math").log((1 - p) / p))
    tr = sys.modules["mmcv.cnn.bricks.transformer"]
    tr.build_positional_encoding = lambda cfg: None

    # mmdet.core helpers the head imports by name
    b = MMDET + "/core/bbox"
    transforms = sys.modules["mmdet.core.bbox.transforms"]
    _load("mmdet.core.bbox.samplers_sampling_result", b + "/samplers/sampling_result.py")
    for p in ("mmdet.core.bbox.samplers",):
        _pkg(p)
    sys.modules["mmdet.core.bbox.samplers.sampling_result"] = sys.modules["mmdet.core.bbox.samplers_sampling_result"]
    _pkg("mmdet.core.bbox.samplers").sampling_result = sys.modules["mmdet.core.bbox.samplers_sampling_result"]
    _load("mmdet.core.bbox.samplers.base_sampler", b + "/samplers/base_sampler.py")
    ps = _load("mmdet.core.bbox.samplers.pseudo_sampler", b + "/samplers/pseudo_sampler.py")

    def multi_apply(func, *args, **kwargs):            # mmdet/core/utils/misc.py:11-29
        pfunc = functools.partial(func, **kwargs) if kwargs else func
        return tuple(map(list, zip(*map(pfunc, *args))))
    core = _pkg("mmdet.core")
    core.bbox_cxcywh_to_xyxy, core.bbox_xyxy_to_cxcywh = transforms.bbox_cxcywh_to_xyxy, transforms.bbox_xyxy_to_cxcywh
    core.build_assigner = lambda cfg: None
    core.build_sampler = lambda cfg, **k: None
    core.multi_apply = multi_apply
    core.reduce_mean = lambda t: t
    core.bbox_overlaps = iou.bbox_overlaps
    sys.modules["mmdet.models.utils"].build_transformer = lambda cfg: None
    builder = _pkg("mmdet.models.builder")
    builder.HEADS, builder.LOSSES = _Registry("heads"), _Registry("losses")
    builder.build_loss = lambda cfg: None
    _pkg("mmdet.models.utils.transformer").inverse_sigmoid = T.inverse_sigmoid
    _pkg("mmdet.models.dense_heads")
    afh = _pkg("mmdet.models.dense_heads.anchor_free_head")
    if not hasattr(afh, "AnchorFreeHead"):

        class _BaseDenseHead(nn.Module):              # what super(AnchorFreeHead, self).__init__(init_cfg) reaches
            def __init__(self, init_cfg=None):
                super().__init__()
        afh.AnchorFreeHead = type("AnchorFreeHead", (_BaseDenseHead,), {})
    _pkg("mmdet.models.losses")
    _load("mmdet.models.losses.utils", MMDET + "/models/losses/utils.py")
    fl = _load("mmdet.models.losses.focal_loss", MMDET + "/models/losses/focal_loss.py")
    il = _load("mmdet.models.losses.iou_loss", MMDET + "/models/losses/iou_loss.py")
    sl = _load("mmdet.models.losses.smooth_l1_loss", MMDET + "/models/losses/smooth_l1_loss.py")
    for p in ("detr_od_ref.models.dense_heads",):
        _pkg(p)
    dn = _load("detr_od_ref.models.dense_heads.dn_components", REF + "/detr_od/models/dense_heads/dn_components.py")
    head = _load("detr_od_ref.models.dense_heads.dino_detr_head", REF + "/detr_od/models/dense_heads/dino_detr_head.py")
    return dict(head=head, dn=dn, focal=fl, iou=il, l1=sl, assigner=ha, sampler=ps, torch=torch)


def load_positional_encoding():
    """detr_od/models/utils/positional_encoding.py (SinePositionalEncodingHW)."""
    import torch.nn as nn
    _install_mmcv_stub()
    _pkg("mmcv.cnn")
    _pkg("mmcv.cnn.bricks")
    tr = _pkg("mmcv.cnn.bricks.transformer")
    tr.POSITIONAL_ENCODING = _Registry("pe")
    runner = _pkg("mmcv.runner")

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
    runner.BaseModule = BaseModule
    for p in ("detr_od_ref", "detr_od_ref.models", "detr_od_ref.models.utils"):
        _pkg(p)
    return _load("detr_od_ref.models.utils.positional_encoding", REF + "/detr_od/models/utils/positional_encoding.py")


def load_bbox_utils():
    """detr_ssod/models/utils/bbox_utils.py (Transform2D.transform_bboxes and helpers)."""
    for p in ("mmdet", "mmdet.core", "mmdet.core.mask"):
        _pkg(p)
    _pkg("mmdet.core.mask.structures").BitmapMasks = type("BitmapMasks", (), {})
    for p in ("detr_ssod_ref", "detr_ssod_ref.models", "detr_ssod_ref.models.utils"):
        _pkg(p)
    return _load("detr_ssod_ref.models.utils.bbox_utils", REF + "/detr_ssod/models/utils/bbox_utils.py")


def load_o2m_assigner():
    """detr_od/core/bbox/assigners/o2m_assigner.py + o2m_assign_result.py on the real mmdet bbox pieces."""
    load_hungarian()
    ds = _pkg("detr_ssod")
    _pkg("detr_ssod.utils").log_every_n = lambda *a, **k: None
    sys.modules["detr_ssod.utils"].log_image_with_boxes = lambda *a, **k: None
    for p in ("detr_od_ref", "detr_od_ref.core", "detr_od_ref.core.bbox", "detr_od_ref.core.bbox.assigners"):
        _pkg(p)
    base = REF + "/detr_od/core/bbox/assigners"
    _load("detr_od_ref.core.bbox.assigners.o2m_assign_result", base + "/o2m_assign_result.py")
    return _load("detr_od_ref.core.bbox.assigners.o2m_assigner", base + "/o2m_assigner.py")


def load_dino_ssod_head():
    """detr_od/models/dense_heads/dino_detr_ssod_head.py (``DINODETRSSODHead.loss`` with both assignment phases) plus
    the reference's TaskAlignedFocalLoss and O2MAssigner, on the same stand-ins as ``load_dino_head``."""
    m = load_dino_head()
    o2m = load_o2m_assigner()
    _pkg("mmcv.runner").get_dist_info = lambda: (0, 1)
    sys.modules["mmdet.core"].multiclass_nms = None          # test-time only (get_bboxes), never called by loss()
    for p in ("detr_od_ref.models.losses",):
        _pkg(p)
    tal = _load("detr_od_ref.models.losses.task_aligned_focal_loss",
                REF + "/detr_od/models/losses/task_aligned_focal_loss.py")
    head = _load("detr_od_ref.models.dense_heads.dino_detr_ssod_head",
                 REF + "/detr_od/models/dense_heads/dino_detr_ssod_head.py")
    m.update(ssod_head=head, o2m=o2m, tal=tal)
    return m


def load_methods(path, class_name, names, namespace):
    """Compile individual methods of a reference class straight from its source file (for classes whose module pulls
    in the whole mmdet detector stack): the function bodies are the reference's, only the surrounding module is not
    executed.  -> {name: function}"""
    import ast
    tree = ast.parse(open(path).read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name)
    out = {}
    for node in cls.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            ns = dict(namespace)
            exec(compile(mod, path, "exec"), ns)
            out[node.name] = ns[node.name]
    return out


def load_dino_head_buildable():
    """``load_dino_head`` with the head module's builder names bound to the loaded reference / mmdet classes, so
    ``DINODETRHead(**cfg)`` itself can be constructed (its ``__init__`` / ``_init_layers`` / ``forward`` run as written):
    transformer -> the reference DINOTransformer, positional encoding -> the reference SinePositionalEncodingHW,
    losses / assigner / sampler -> the real mmdet classes."""
    import torch.nn as nn
    m = load_dino_head()
    T, _ = load_dino_transformer()
    pe = load_positional_encoding()
    head = m["head"]
    strip = lambda cfg: {k: v for k, v in cfg.items() if k != "type"}
    losses = {"FocalLoss": m["focal"].FocalLoss, "L1Loss": m["l1"].L1Loss, "GIoULoss": m["iou"].GIoULoss}
    head.build_transformer = lambda cfg: T.DINOTransformer(**strip(cfg))
    head.build_positional_encoding = lambda cfg: pe.SinePositionalEncodingHW(**strip(cfg))
    head.build_loss = lambda cfg: losses[cfg["type"]](**strip(cfg))
    head.build_assigner = lambda cfg: m["assigner"].HungarianAssigner(**strip(cfg))
    head.build_sampler = lambda cfg, context=None: m["sampler"].PseudoSampler()
    head.build_activation_layer = lambda cfg: nn.ReLU(inplace=True)

    return m


def load_mmdet_resnet():
    """thirdparty/mmdetection/mmdet/models/backbones/resnet.py + models/utils/res_layer.py (the backbone the configs
    name: ResNet-50, frozen_stages=1, BN frozen + norm_eval, style 'pytorch').  mmcv's layer builders are stood in by
    what they return for the shipped config: ``nn.Conv2d`` and ``(name, nn.BatchNorm2d)`` with the ``requires_grad``
    flag of ``norm_cfg`` applied (mmcv/cnn/bricks/norm.py)."""
    import torch.nn as nn
    _install_mmcv_stub()
    cnn = _pkg("mmcv.cnn")

    def build_conv_layer(cfg, *args, **kwargs):
        assert cfg is None or cfg.get("type", "Conv2d") in ("Conv2d", "Conv"), cfg
        return nn.Conv2d(*args, **kwargs)

    def build_norm_layer(cfg, num_features, postfix=""):
        assert cfg["type"] == "BN", cfg
        layer = nn.BatchNorm2d(num_features, eps=cfg.get("eps", 1e-5))
        for p in layer.parameters():
            p.requires_grad = cfg.get("requires_grad", True)
        return "bn" + str(postfix), layer
    cnn.build_conv_layer, cnn.build_norm_layer = build_conv_layer, build_norm_layer
    cnn.build_plugin_layer = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("plugins are not configured"))
    runner = _pkg("mmcv.runner")

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

    class Sequential(BaseModule, nn.Sequential):
        def __init__(self, *args, init_cfg=None):
            BaseModule.__init__(self, init_cfg)
            nn.Sequential.__init__(self, *args)
    runner.BaseModule, runner.Sequential = BaseModule, Sequential
    for p in ("mmdet", "mmdet.models", "mmdet.models.utils", "mmdet.models.backbones"):
        _pkg(p)
    _pkg("mmdet.models.builder").BACKBONES = _Registry("backbone")
    rl = _load("mmdet.models.utils.res_layer", MMDET + "/models/utils/res_layer.py")
    sys.modules["mmdet.models.utils"].ResLayer = rl.ResLayer
    return _load("mmdet.models.backbones.resnet", MMDET + "/models/backbones/resnet.py")


def load_dino_ssod_head_buildable():
    """``load_dino_ssod_head`` with the SSOD head module's builder names bound like ``load_dino_head_buildable`` does
    for the supervised head (plus O2MAssigner and TaskAlignedFocalLoss), so ``DINODETRSSODHead(**cfg)`` is constructed
    by its own ``__init__``."""
    import torch.nn as nn
    load_dino_head_buildable()
    m = load_dino_ssod_head()
    T, _ = load_dino_transformer()
    pe = load_positional_encoding()
    head = m["ssod_head"]
    strip = lambda cfg: {k: v for k, v in cfg.items() if k != "type"}
    losses = {"FocalLoss": m["focal"].FocalLoss, "L1Loss": m["l1"].L1Loss, "GIoULoss": m["iou"].GIoULoss,
              "TaskAlignedFocalLoss": m["tal"].TaskAlignedFocalLoss}
    assigners = {"HungarianAssigner": m["assigner"].HungarianAssigner, "O2MAssigner": m["o2m"].O2MAssigner}
    head.build_transformer = lambda cfg: T.DINOTransformer(**strip(cfg))
    head.build_positional_encoding = lambda cfg: pe.SinePositionalEncodingHW(**strip(cfg))
    head.build_loss = lambda cfg: losses[cfg["type"]](**strip(cfg))
    head.build_assigner = lambda cfg: assigners[cfg["type"]](**strip(cfg))
    head.build_sampler = lambda cfg, context=None: m["sampler"].PseudoSampler()
    head.build_activation_layer = lambda cfg: nn.ReLU(inplace=True)
    return m
