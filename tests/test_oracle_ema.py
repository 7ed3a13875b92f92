"""EMA oracle against goldens from the real MeanTeacher hook (mean_teacher.py:37-64)."""
import numpy as np
import torch

from oracle.ema_oracle import ema_momentum, ema_update


def test_schedule_kats():
    # iter 0 -> copy, iter 1 -> 0.5, saturates at iter 999
    assert ema_momentum(0) == 0
    assert ema_momentum(1) == 0.5
    assert ema_momentum(998) < 0.999 and ema_momentum(999) == 0.999 and ema_momentum(10 ** 6) == 0.999


def test_matches_reference_hook(ema_golden):
    g = ema_golden
    n = len(g["student"])
    student = [torch.from_numpy(g["student"][str(i)].copy()) for i in range(n)]
    teacher = [torch.from_numpy(g["teacher0"][str(i)].copy()) for i in range(n)]
    for it, m in zip(g["iters"], g["momenta"]):
        assert ema_momentum(int(it)) == m
        ema_update(teacher, student, ema_momentum(int(it)))
        for i in range(n):
            assert np.array_equal(teacher[i].numpy(), g[f"teacher_after_{int(it)}"][str(i)])
