"""Oracle: linear-sum-assignment via oracle/lsap.c (ctypes).  TEST INFRASTRUCTURE ONLY.

The reference calls ``scipy.optimize.linear_sum_assignment`` (mmdet
hungarian_assigner.py:136; detr_ssod/models/dino_detr_ssod.py:279) -- scipy is un-vendored
and unpinned in /root/reference; oracle/lsap.c restates its published algorithm and is pinned
against scipy 1.18.1 in tests/test_oracle_lsap.py and tests/golden/lsap_golden.npz.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile oracle/lsap.c -> oracle/_build/liboracle.so (gcc)."""
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_lsap_f32.restype = ctypes.c_int
        _LIB.oracle_lsap_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.c_void_p, ctypes.c_void_p]
    return _LIB


def linear_sum_assignment(cost):
    """Same contract as scipy's for a float32 cost matrix: (rows ascending, cols), int64."""
    cost = np.ascontiguousarray(np.asarray(cost, dtype=np.float32))
    assert cost.ndim == 2
    nr, nc = cost.shape
    k = min(nr, nc)
    rows = np.empty(k, dtype=np.int64)
    cols = np.empty(k, dtype=np.int64)
    rc = _lib().oracle_lsap_f32(cost.ctypes.data, nr, nc, rows.ctypes.data, cols.ctypes.data)
    if rc == 1:
        raise ValueError("matrix contains invalid numeric entries")
    if rc == 2:
        raise ValueError("cost matrix is infeasible")
    return rows, cols
