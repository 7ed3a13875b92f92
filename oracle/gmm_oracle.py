"""Oracle: cost threshold from a 1-D two-component Gaussian mixture -- CPU restatement of
``DinoDetrSSOD._fit_gmm`` (detr_ssod/models/dino_detr_ssod.py:832-890).  TEST INFRASTRUCTURE ONLY: the product runs
``sdb_gmm_threshold_f32`` (csrc/ssod.cu) and is checked against this file.

The reference fits ``sklearn.mixture.GaussianMixture(2, covariance_type='diag', reg_covar=1e-5)`` initialised
with means [min, max], weights [.5, .5], precisions [1, 1] on the matched Hungarian costs pooled over images and
ranks (a few dozen to a few hundred scalars), and returns the cost of the most likely sample of component 0
(falling back to component 1).  sklearn is an un-vendored, unpinned dependency of the reference
(thirdparty/mmdetection/requirements/optional.txt:5); this is its EM loop (tol 1e-3, max_iter 100, one init)
restated in float64 numpy for that 1-D case, pinned against sklearn 1.9 in tests/test_ssod_host.py and against the
reference method itself in tests/test_dino_reference_golden.py (tests/golden/ssod_gmm_golden.npz).
"""
import numpy as np


def _e_step(x, w, mu, var):
    # log N(x | mu_k, var_k) + log w_k, for k = 0, 1
    log_prob = -0.5 * (np.log(2 * np.pi) + np.log(var)[None, :] + (x[:, None] - mu[None, :]) ** 2 / var[None, :])
    weighted = log_prob + np.log(w)[None, :]
    m = weighted.max(1, keepdims=True)
    log_norm = (m + np.log(np.exp(weighted - m).sum(1, keepdims=True)))[:, 0]
    return log_norm, weighted - log_norm[:, None]


def fit_two_gaussians(x, tol=1e-3, max_iter=100, reg_covar=1e-5):
    """-> (weights, means, variances) after EM from the reference's fixed initialisation."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    w = np.array([0.5, 0.5])
    mu = np.array([x.min(), x.max()])
    var = np.array([1.0, 1.0])                      # precisions_init = 1
    lower = -np.inf
    for _ in range(max_iter):
        prev = lower
        log_norm, log_resp = _e_step(x, w, mu, var)
        lower = log_norm.mean()
        resp = np.exp(log_resp)
        nk = resp.sum(0) + 10 * np.finfo(np.float64).eps
        mu = (resp * x[:, None]).sum(0) / nk
        var = (resp * x[:, None] ** 2).sum(0) / nk - 2 * mu * (resp * x[:, None]).sum(0) / nk + mu ** 2 + reg_covar
        w = nk / nk.sum()
        if abs(lower - prev) < tol:
            break
    return w, mu, var


def fit_gmm_threshold(costs):
    """Matched costs (any 1-D array-like) -> python float threshold (dino_detr_ssod.py:832-890)."""
    x = np.asarray(costs, dtype=np.float64).reshape(-1)
    if x.size == 0:
        return 0.0
    x = np.sort(x)
    if x.size < 2:
        return float(x[0])
    w, mu, var = fit_two_gaussians(x)
    log_norm, log_resp = _e_step(x, w, mu, var)      # score_samples / predict
    assign = log_resp.argmax(1)
    for comp in (0, 1):
        sel = assign == comp
        if sel.any():
            return float(x[sel][log_norm[sel].argmax()])
    return float(x[0])
