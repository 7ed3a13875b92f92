"""Oracle: mean-teacher EMA on the CPU.  TEST INFRASTRUCTURE ONLY.

Restates detr_ssod/utils/hooks/mean_teacher.py:37-64: the momentum schedule
``min(momentum, 1 - (1 + warm_up) / (iter + 1 + warm_up))`` and the per-parameter update
``teacher.mul_(m).add_(student, alpha=1 - m)`` (parameters only, no buffers, no requires_grad
filter).  Pinning: closed-form; the reference has no test for it ("parity unpinned" by reference
tests) -- tests/test_ema.py pins the schedule values listed in SURVEY.md appendix A.8.
"""
import torch


def ema_momentum(cur_iter, momentum=0.999, warm_up=0):
    return min(momentum, 1 - (1 + warm_up) / (cur_iter + 1 + warm_up))


@torch.no_grad()
def ema_update(teacher_params, student_params, momentum):
    """In-place python loop, one mul_ + one add_ per tensor, like the reference hook."""
    for t, s in zip(teacher_params, student_params):
        t.mul_(momentum).add_(s, alpha=1 - momentum)
