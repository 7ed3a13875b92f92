from .mean_teacher import EmaPlan, MeanTeacher

__all__ = ["EmaPlan", "MeanTeacher"]
