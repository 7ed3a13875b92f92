from .dino_detr_ssod import DinoDetrSSOD, Projector
from .o2m_assigner import O2MAssigner
from .ssod_head import DINODETRSSODHead
from .task_aligned_focal_loss import TaskAlignedFocalLoss

__all__ = ["DinoDetrSSOD", "Projector", "O2MAssigner", "DINODETRSSODHead", "TaskAlignedFocalLoss"]
