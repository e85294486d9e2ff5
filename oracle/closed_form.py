"""Closed-form profile of the analytic Gaussian toy.

TEST INFRASTRUCTURE ONLY (see oracle/reference_port.py header).  Pinned against
the reference's own scipy oracle `AnalyticalKernel.get_scipy_profile`
(prospect/kernels/analytical_kernel.py:116-122) through tests/golden/toy*_loglkl.npz.

The reference builds the residual as [varying..., fixed] but contracts it with
inv(covmat) in the original x1..xd order (analytical_kernel.py:59-66),
so with S = sym(inv(C)) the blocks are taken at slots 0..n-1 / n, NOT at the
profiled parameter's own row.
"""
import numpy as np


def sym_precision(covmat):
    a = np.linalg.inv(np.asarray(covmat, dtype=float))
    return 0.5 * (a + a.T)


def profile_minimum(means, covmat, fixed_index, value):
    """Returns (loglkl*, x_varying*) of min over the varying parameters at x_fixed = value."""
    means = np.asarray(means, dtype=float)
    d = len(means)
    s = sym_precision(covmat)
    n = d - 1
    varying = [i for i in range(d) if i != fixed_index]
    s_aa, s_ab, s_bb = s[:n, :n], s[:n, n], s[n, n]
    f = value - means[fixed_index]
    w = np.linalg.solve(s_aa, s_ab)
    loglkl = 0.5 * (s_bb - s_ab @ w) * f * f
    x = means[varying] - w * f
    return loglkl, x


def profile_curve(means, covmat, fixed_index, values):
    return np.array([profile_minimum(means, covmat, fixed_index, v)[0] for v in values])


def stationary_covariance(covmat, temperature=1.0):
    """Free (nothing fixed) chain at temperature T samples N(mu, T * S^-1)."""
    return temperature * np.linalg.inv(sym_precision(covmat))
