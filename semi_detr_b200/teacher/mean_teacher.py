"""``MeanTeacher`` hook -- host-side mirror of detr_ssod/utils/hooks/mean_teacher.py:7-64 (same constructor,
``before_run`` / ``before_train_iter`` / ``after_train_iter`` / ``momentum_update`` and momentum schedule) whose
update is ONE kernel launch over all parameters (``sdb_ema_update_f32``) instead of a Python loop of
``mul_`` + ``add_`` per tensor.
"""
import ctypes
from bisect import bisect_right

import torch

from .. import _lib
from ..registry import HOOKS

_CHUNK = 1 << 15      # floats per table entry: 128 KB of teacher per CTA visit


class EmaPlan:
    """Device-resident chunk table pairing teacher and student parameters (parameters only, in
    ``named_parameters()`` order, no ``requires_grad`` filter -- mean_teacher.py:61-64)."""

    def __init__(self, teacher_params, student_params):
        self.teacher = list(teacher_params)
        self.student = list(student_params)
        assert len(self.teacher) == len(self.student)
        self._key = None
        self._table = None
        self.num_chunks = 0
        self.num_params = 0

    def _pointers(self):
        return tuple(t.data_ptr() for t in self.teacher) + tuple(s.data_ptr() for s in self.student)

    def _build(self, key):
        entries = []
        total = 0
        for t, s in zip(self.teacher, self.student):
            if not (t.is_cuda and s.is_cuda):
                raise RuntimeError("MeanTeacher: parameters must live on a CUDA device (no CPU path)")
            if t.dtype != torch.float32 or s.dtype != torch.float32:
                raise RuntimeError("MeanTeacher: fp32 parameters expected")
            # the blend is elementwise over storage: any dense layout works as long as both sides share it
            # (channels-last conv weights included)
            if t.shape != s.shape or t.stride() != s.stride() or not _dense(t):
                raise RuntimeError("MeanTeacher: teacher/student parameters must be dense and share one layout")
            n = t.numel()
            total += n
            tp, sp = t.data_ptr(), s.data_ptr()
            for o in range(0, n, _CHUNK):
                entries.append((tp + 4 * o, sp + 4 * o, min(_CHUNK, n - o)))
        arr = (_lib.EmaChunk * max(len(entries), 1))()
        for i, (tp, sp, c) in enumerate(entries):
            arr[i].teacher, arr[i].student, arr[i].count = tp, sp, c
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self._table = raw.to(self.teacher[0].device) if entries else None
        self.num_chunks = len(entries)
        self.num_params = total
        self._key = key

    @torch.no_grad()
    def step(self, momentum):
        if not self.teacher:
            return
        key = self._pointers()
        if key != self._key:
            self._build(key)
        if self.num_chunks == 0:
            return
        dev = self.teacher[0].device
        with torch.cuda.device(dev):
            rc = _lib.lib().sdb_ema_update_f32(_lib.current_stream(dev), self._table.data_ptr(), self.num_chunks,
                                               float(momentum))
        _lib.check(rc, "ema_update")
        _lib.LAUNCHES["ema_update"] += 1


def _dense(t):
    """True when the tensor's elements occupy ``numel`` consecutive storage slots (any permutation of a contiguous
    layout)."""
    expect = 1
    for size, stride in sorted(((sz, st) for sz, st in zip(t.shape, t.stride()) if sz > 1), key=lambda p: p[1]):
        if stride != expect:
            return False
        expect *= size
    return True


def _unwrap(model):
    return model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model


@HOOKS.register_module()
class MeanTeacher:
    def __init__(self, momentum=0.999, interval=1, warm_up=100, decay_intervals=None, decay_factor=0.1):
        assert 0 <= momentum <= 1
        assert isinstance(interval, int) and interval > 0
        assert isinstance(decay_intervals, list) or decay_intervals is None
        self.momentum = momentum
        self.interval = interval
        self.warm_up = warm_up
        self.decay_intervals = decay_intervals
        self.decay_factor = decay_factor
        self._plan = None
        self._plan_model = None

    def before_run(self, runner):
        model = _unwrap(runner.model)
        assert hasattr(model, "teacher") and hasattr(model, "student")
        if runner.iter == 0:
            self.momentum_update(model, 0)          # teacher <- student

    def before_train_iter(self, runner):
        curr_step = runner.iter
        if curr_step % self.interval != 0:
            return
        model = _unwrap(runner.model)
        momentum = min(self.momentum, 1 - (1 + self.warm_up) / (curr_step + 1 + self.warm_up))
        runner.log_buffer.output["ema_momentum"] = momentum
        self.momentum_update(model, momentum)

    def after_train_iter(self, runner):
        if self.decay_intervals is None:
            return
        self.momentum = 1 - (1 - self.momentum) / self.decay_factor ** bisect_right(self.decay_intervals, runner.iter)

    def momentum_update(self, model, momentum):
        if self._plan is None or self._plan_model is not model:
            self._plan = EmaPlan([p.data for _, p in model.teacher.named_parameters()],
                                 [p.data for _, p in model.student.named_parameters()])
            self._plan_model = model
        self._plan.step(momentum)


@HOOKS.register_module()
class StepRecord:
    """detr_ssod/utils/hooks/step_record.py:7-27: publishes the runner's iteration on the model."""

    def __init__(self, normalize=True, name="curr_step"):
        self.normalize, self.name = normalize, name

    def before_train_iter(self, runner):
        model = _unwrap(runner.model)
        assert hasattr(model, self.name)
        it = getattr(runner, "iter", None)
        setattr(model, self.name, it / 10000 if self.normalize else it)
