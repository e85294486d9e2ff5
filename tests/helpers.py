"""Shared builders for the test-suite (golden toy problems, synthetic SPD targets)."""
import os

import numpy as np

from prospect_public_b200.problem import build_problem

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def mcmc_rows(d):
    m = golden(f'toy{d}_mcmc.npz')
    return np.concatenate([m[f'chain_{i}'] for i in range(4)], axis=0)


def toy_profile_problem(d, points=None):
    """x2-profile of the d-D toy with the reference's own per-value start points / bin covariances."""
    k, ini = golden(f'toy{d}_kernel.npz'), golden(f'toy{d}_init.npz')
    sel = np.arange(len(ini['values'])) if points is None else np.asarray(points)
    hp = build_problem(k['means'], k['covmat'], [-2.5] * d, [2.5] * d, 1, ini['values'][sel], ini['covmat'][sel])
    starts = ini['position'][sel][:, :-1]
    return hp, starts, ini['initial_loglkl'][sel], ini['values'][sel], k


def toy_free_problem(d, covmat=None):
    """Nothing fixed (global optimisation / mcmc).  Default proposal: the kernel's default covmat."""
    k = golden(f'toy{d}_kernel.npz')
    cov = 0.005 * np.diag(np.full(d, 5.0)) if covmat is None else covmat
    return build_problem(k['means'], k['covmat'], [-2.5] * d, [2.5] * d, None, None, cov), k


def spd_target(d, seed=0, lam_range=(1e-3, 1e-1)):
    """Synthetic correlated SPD Gaussian for d beyond the reference recipe:
    C = Q diag(lam) Q^T, lam log-uniform, Q from the QR of a seeded Gaussian matrix, mu ~ U(0,1)."""
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    lam = np.exp(rng.uniform(np.log(lam_range[0]), np.log(lam_range[1]), d))
    cov = (q * lam) @ q.T
    return rng.uniform(0.0, 1.0, d), 0.5 * (cov + cov.T)


def conditional_covariance(covmat, fixed_index):
    """Covariance of the varying block given the fixed one (closed form) -- a good proposal covmat."""
    s = np.linalg.inv(covmat)
    s = 0.5 * (s + s.T)
    n = covmat.shape[0] - 1
    return np.linalg.inv(s[:n, :n])


class ToyHostKernel:
    """A likelihood that only exists on the HOST, behind the reference's kernel plugin surface
    (prospect/kernels/base_kernel.py:9-138: `param`, `set_fixed_parameters`, `loglkl(position dict)`, `logprior`,
    defaults): a correlated Gaussian plus a non-Gaussian ripple, so nothing on the device could have evaluated it.
    Parameters x1..xd, ranges `box`; every call is counted."""

    def __init__(self, d, seed=3, box=(-2.5, 2.5), ignore_prior=False):
        mu, cov = spd_target(d, seed=seed, lam_range=(5e-3, 2e-1))
        self.mu, self.prec = mu, np.linalg.inv(cov)
        self.names = [f'x{i}' for i in range(1, d + 1)]
        self.param = {'varying': {n: {'range': list(box)} for n in self.names}, 'fixed': {}, 'derived': {}}
        self.ignore_prior = ignore_prior
        self.calls = 0

    def set_fixed_parameters(self, fixed_param_dict):            # base_kernel.py:46-51
        for name, value in fixed_param_dict.items():
            self.param['fixed'][name] = self.param['varying'].pop(name)
            self.param['fixed'][name]['fixed_value'] = value

    @property
    def varying_param_names(self):
        return list(self.param['varying'].keys())

    def value(self, full):
        r = np.asarray(full, dtype=np.float64) - self.mu
        return float(0.5 * r @ self.prec @ r + 0.05 * np.sum(np.sin(3.0 * np.asarray(full)) ** 2))

    def loglkl(self, prop):                                      # base_kernel.py:72-81
        for name, par in self.param['fixed'].items():
            prop[name] = [par['fixed_value']]
        self.calls += 1
        return self.value([prop[n][0] for n in self.names])

    def logprior(self, position):                                # base_kernel.py:83-102
        if self.ignore_prior:
            return 0.0
        for name, val in position.items():
            if name in self.param['varying']:
                lo, hi = self.param['varying'][name]['range']
                if (lo is not None and val[0] < lo) or (hi is not None and val[0] > hi):
                    return np.inf
        return 0.0

    def get_default_covmat(self):
        return 0.005 * np.diag([5.0] * len(self.varying_param_names))

    def get_default_initial_position(self):
        return {n: [0.3] for n in self.varying_param_names}

    def finalize(self):
        pass
