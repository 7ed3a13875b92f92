"""``Config.fromfile`` / ``patch_config`` on the reference's OWN config files (the registry
surface -- "configs run unchanged"): ``_base_`` chains, ``_delete_``, ``${...}`` and the ``semi_wrapper`` swap resolve
to the dicts this package's registry builds ``DINODETR`` and ``DinoDetrSSOD`` from.  /root/reference exists only in the
build container (where this CPU suite runs); the synthetic-input dicts the GPU box uses are checked against it here."""
import os

import pytest
import torch

from semi_detr_b200.config import Config, ConfigDict, build_detector, patch_config, resolve

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present on this box")


def test_merge_delete_and_resolve(tmp_path):
    (tmp_path / "base.py").write_text(
        "model = dict(type='A', head=dict(k=1, loss=dict(w=2.0)), keep=3)\nrunner = dict(type='Epoch', max_epochs=12)\n")
    (tmp_path / "mid.py").write_text(
        "_base_ = ['base.py']\nimport os\nmodel = dict(head=dict(loss=dict(_delete_=True, kind='l1')))\nfold = 2\n")
    (tmp_path / "top.py").write_text(
        "_base_ = 'mid.py'\nrunner = dict(_delete_=True, type='Iter', max_iters=5)\n"
        "wrapper = dict(type='W', model='${model}')\nwork_dir = 'w/${cfg_name}/${fold}'\n")
    cfg = Config.fromfile(str(tmp_path / "top.py"))
    assert cfg.model.head.to_dict() == dict(k=1, loss=dict(kind="l1")) and cfg.model.keep == 3
    assert cfg.runner.to_dict() == dict(type="Iter", max_iters=5)
    assert "os" not in cfg                                  # imported modules are not config keys
    d = cfg.to_dict()
    d["cfg_name"] = "top"
    r = resolve(d)
    assert r["wrapper"]["model"] == r["model"] and r["work_dir"] == "w/top/2"
    assert isinstance(cfg.model, ConfigDict)
    with pytest.raises(KeyError):                           # two bases defining the same key
        (tmp_path / "b2.py").write_text("model = dict(type='B')\n")
        (tmp_path / "dup.py").write_text("_base_ = ['base.py', 'b2.py']\n")
        Config.fromfile(str(tmp_path / "dup.py"))
    with pytest.raises(FileNotFoundError):
        Config.fromfile(str(tmp_path / "missing.py"))


@needs_ref
def test_supervised_config_builds_dinodetr_from_the_reference_file():
    from semi_detr_b200.synthetic import DINO_R50_4SCALE
    cfg = Config.fromfile(os.path.join(REF, "configs/dino_detr/dino_detr_r50_8x2_12e_coco.py"))
    model = cfg.model.to_dict()
    assert model["backbone"]["init_cfg"] == dict(type="Pretrained", checkpoint="torchvision://resnet50")
    stripped = cfg.model.to_dict()
    stripped["backbone"].pop("init_cfg")
    assert stripped == DINO_R50_4SCALE                      # what bench.py / the GPU tests build without the file
    assert cfg.optimizer.to_dict() == dict(type="AdamW", lr=1e-4, weight_decay=1e-4, paramwise_cfg=dict(
        custom_keys=dict(backbone=dict(lr_mult=0.1, decay_mult=1.0))))
    assert cfg.optimizer_config.grad_clip.max_norm == 0.1
    torch.manual_seed(0)
    det = build_detector(cfg.model)                         # init_cfg accepted and ignored (no network)
    assert type(det).__name__ == "DINODETR" and type(det.bbox_head).__name__ == "DINODETRHead"
    assert type(det.bbox_head.assigner).__name__ == "HungarianAssigner"
    assert sum(p.numel() for p in det.parameters()) == 46938088


@needs_ref
def test_ssod_config_builds_the_teacher_student_wrapper_from_the_reference_file():
    from semi_detr_b200.engine import train_step_options
    from semi_detr_b200.synthetic import ssod_model_cfg
    raw = Config.fromfile(os.path.join(REF, "configs/detr_ssod/detr_ssod_dino_detr_r50_coco_120k.py"))
    assert raw.semi_wrapper.model == "${model}"
    assert raw.runner.to_dict() == dict(type="IterBasedRunner", max_iters=120000)      # _delete_ dropped max_epochs
    cfg = patch_config(raw)
    assert "semi_wrapper" not in cfg and cfg.model.type == "DinoDetrSSOD"
    assert cfg.work_dir == "work_dirs/detr_ssod_dino_detr_r50_coco_120k/1/1"
    assert cfg.data.train.sup.ann_file.endswith("instances_train2017.1@1.json")
    assert cfg.data.samples_per_gpu == 5 and cfg.data.sampler.train.sample_ratio == [1, 4]
    stripped = cfg.model.to_dict()
    stripped["model"]["backbone"].pop("init_cfg")
    assert stripped == ssod_model_cfg()
    opts = train_step_options(cfg)
    assert opts == dict(lr=1e-4, weight_decay=1e-4, backbone_lr_mult=0.1, max_grad_norm=0.1, momentum=0.999,
                        warm_up=0)
    # a narrow copy of the same config (fewer queries / layers) keeps the CPU build cheap; the type strings, the
    # nesting and the train_cfg plumbing are the file's
    m = cfg.model.to_dict()
    m["model"]["bbox_head"].update(num_query=30, transformer=dict(type="DINOTransformer", num_encoder_layers=1,
                                                                  num_decoder_layers=2))
    torch.manual_seed(0)
    wrapper = build_detector(m)
    assert type(wrapper).__name__ == "DinoDetrSSOD"
    assert type(wrapper.student.bbox_head).__name__ == "DINODETRSSODHead"
    assert type(wrapper.student.bbox_head.assigner1).__name__ == "O2MAssigner"
    assert wrapper.train_cfg["unsup_weight"] == 4.0 and wrapper.train_cfg["pseudo_label_initial_score_thr"] == 0.4


@needs_ref
@pytest.mark.parametrize("rel", ["configs/dino_detr/dino_detr_r50_8x2_12e_voc.py",
                                 "configs/detr_ssod/detr_ssod_dino_detr_r50_coco_full_240k.py",
                                 "configs/detr_ssod/detr_ssod_dino_detr_r50_voc_80k.py"])
def test_every_shipped_config_loads(rel):
    cfg = patch_config(Config.fromfile(os.path.join(REF, rel)))
    model = cfg.model.model if cfg.model.type == "DinoDetrSSOD" else cfg.model
    assert model.type == "DINODETR" and model.bbox_head.transformer.type == "DINOTransformer"
