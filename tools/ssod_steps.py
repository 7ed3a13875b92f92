"""Per-step wall time of the teacher-student step (Hungarian phase, frozen state) with what decides its shapes:
python tools/ssod_steps.py [seed] -> one line per step: ms, pseudo boxes per image, NMS survivors."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semi_detr_b200 import dino, ssod  # noqa: E402,F401
from semi_detr_b200.engine import FusedSSODTrainStep  # noqa: E402
from semi_detr_b200.registry import DETECTORS  # noqa: E402
from semi_detr_b200.ssod import dino_detr_ssod as S  # noqa: E402
from semi_detr_b200.synthetic import ssod_batch, ssod_model_cfg  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = DETECTORS.build(ssod_model_cfg()).to(dev).train()
fused = FusedSSODTrainStep(model, momentum=0.999, warm_up=0, world_size=1, lr=0.0)
host = ssod_batch(1, 4, 800, 1333, seed=seed)
data = dict(img=host["img"].to(dev), img_metas=[dict(m) for m in host["img_metas"]],
            gt_bboxes=[x.to(dev) for x in host["gt_bboxes"]], gt_labels=[x.to(dev) for x in host["gt_labels"]])
rec = {}
orig = S.DinoDetrSSOD.unsup_loss


def spy(self, student_info, teacher_info, pseudo_bboxes, pseudo_labels, pseudo_scores):
    rec["pseudo"] = [int(b.shape[0]) for b in pseudo_bboxes]
    return orig(self, student_info, teacher_info, pseudo_bboxes, pseudo_labels, pseudo_scores)


S.DinoDetrSSOD.unsup_loss = spy
for it in range(14):
    fused.iter = 60000
    torch.cuda.synchronize()
    t = time.time()
    loss, _ = fused(dict(data, img_metas=[dict(m) for m in data["img_metas"]]))
    torch.cuda.synchronize()
    print(f"step {it}: {(time.time() - t) * 1e3:8.1f} ms  loss {float(loss):.4f}  pseudo boxes {rec.get('pseudo')}  "
          f"reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB", flush=True)
