"""oracle/lsap.c against scipy (the reference's solver: hungarian_assigner.py:136) and the committed goldens."""
import numpy as np
import pytest

from oracle.lsap_oracle import linear_sum_assignment as lsa

scipy_opt = pytest.importorskip("scipy.optimize")


def test_golden(lsap_golden):
    for name, c in lsap_golden.items():
        if name == "scipy_version":
            continue
        r, cc = lsa(c["cost"])
        assert np.array_equal(r, c["rows"]) and np.array_equal(cc, c["cols"]), name


def test_tie_kats():
    # SURVEY.md appendix B (probed on scipy 1.18.1)
    r, c = lsa(np.zeros((4, 2), np.float32))
    assert r.tolist() == [0, 1] and c.tolist() == [0, 1]
    r, c = lsa(np.array([[1, 1], [1, 1], [0, 0]], np.float32))
    assert r.tolist() == [1, 2] and c.tolist() == [1, 0]


def test_random_against_scipy():
    rng = np.random.default_rng(1)
    for t in range(600):
        nr, nc = rng.integers(1, 50, 2)
        kind = t % 3
        if kind == 0:
            c = rng.standard_normal((nr, nc)).astype(np.float32)
        elif kind == 1:
            c = rng.integers(0, 3, (nr, nc)).astype(np.float32)
        else:
            c = np.round(rng.standard_normal((nr, nc)) * 2).astype(np.float32)
        a = scipy_opt.linear_sum_assignment(c)
        b = lsa(c)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_empty_and_invalid():
    r, c = lsa(np.zeros((5, 0), np.float32))
    assert r.size == 0 and c.size == 0
    with pytest.raises(ValueError, match="invalid numeric"):
        lsa(np.array([[np.nan, 1.0]], np.float32))
    with pytest.raises(ValueError, match="invalid numeric"):
        lsa(np.array([[-np.inf, 1.0]], np.float32))
    with pytest.raises(ValueError, match="infeasible"):
        lsa(np.full((2, 2), np.inf, np.float32))
