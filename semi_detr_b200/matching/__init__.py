from .hungarian_assigner import AssignResult, HungarianAssigner, MatchTargets
from .match_cost import BBoxL1Cost, FocalLossCost, IoUCost, build_match_cost

__all__ = ["AssignResult", "HungarianAssigner", "MatchTargets", "BBoxL1Cost", "FocalLossCost", "IoUCost",
           "build_match_cost"]
