from .mean_teacher import EmaPlan, MeanTeacher, StepRecord

__all__ = ["EmaPlan", "MeanTeacher", "StepRecord"]
