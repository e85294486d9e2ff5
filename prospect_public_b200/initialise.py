"""Start point + proposal covariance for every profile value (or for a global optimisation).

Mirrors InitialiseOptimiserTask (prospect/tasks/initialise_optimiser_task.py:27-200): the same
`start_from_mcmc` binning -- N = ceil(start_bin_fraction * rows), the N rows nearest to the profile value
in numpy's argsort order, multiplicity-weighted bin covariance with the profile row/column removed, best
point of the bin as the start -- plus the `start_from_position` + `start_from_covmat` and
`start_from_profile` variants.  Index bookkeeping is done with the very numpy calls the reference makes
(argsort / take / argmin / ceil) so it is bit-exact; the covariance reduction runs on the GPU
(pb200_weighted_moments).

The reference repeats this work for every (value, repetition) task although the result only depends on
the value; here it is done once per value and the MCMC is loaded once per job.
"""
import glob
import os
import re

import numpy as np

from .mcmc import Chain


def load_mcmc(mcmc_path, collapse=True):
    """Read an unpacked MCMC `<root>_<i>.txt` + `<root>.paramnames` (MontePython / cobaya / GetDist layout:
    multiplicity, -log L, parameters) without GetDist: every chain file in index order, no burn-in removal,
    which is what reproduces the reference's own fixtures.  -> Chain or list[Chain]."""
    files = []
    for f in glob.glob(mcmc_path + '_*.txt'):
        m = re.search(r'_(\d+)\.txt$', f)
        if m:
            files.append((int(m.group(1)), f))
    if not files:
        raise FileNotFoundError(f"No chain files '{mcmc_path}_<i>.txt' found.")
    names = []
    with open(mcmc_path + '.paramnames') as fh:
        for line in fh:
            if line.strip():
                names.append(line.split()[0])
    chains = []
    for _, f in sorted(files):
        data = np.atleast_2d(np.loadtxt(f))
        ch = Chain()
        ch.mults, ch.loglkls = data[:, 0], data[:, 1]
        for i, nm in enumerate(names):
            ch.positions[nm] = data[:, 2 + i]
        chains.append(ch)
    if not collapse:
        return chains
    total = Chain()
    total.mults = np.concatenate([c.mults for c in chains])
    total.loglkls = np.concatenate([c.loglkls for c in chains])
    for nm in names:
        total.positions[nm] = np.concatenate([c.positions[nm] for c in chains])
    return total


def get_points_in_bin(chain, param_val, param_name, n_points):
    """Indices of the `n_points` rows closest to `param_val` (initialise_optimiser_task.py:186-196)."""
    distance = np.abs(np.array(chain.positions[param_name]) - param_val)
    return np.argsort(distance)[:n_points]


def get_fraction_of_points_in_bin(chain, param_val, param_name, fraction):
    n_points = int(np.ceil(fraction * len(chain.mults)))
    return get_points_in_bin(chain, param_val, param_name, n_points)


def _weighted_cov_gpu(x, mults, device):
    import torch
    from . import engine
    xt = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=device)
    w = torch.as_tensor(np.asarray(mults).astype(np.int32), device=device)
    return engine.covariance_from_moments(*engine.weighted_moments(xt, x.shape[1], w)).cpu().numpy()


class StartingPoints:
    """Per profile point: start position (varying order), proposal covmat, initial_loglkl, bin indices."""

    def __init__(self, names_varying, positions, covmats, initial_loglkl, indices=None):
        self.names = names_varying
        self.positions = positions            # [P, n]
        self.covmats = covmats                # [P, n, n]
        self.initial_loglkl = initial_loglkl  # [P]
        self.indices = indices                # list of int arrays or None


def initialise_optimiser(config, kernel):
    """`kernel` must not have its profile parameter fixed yet.  Returns (StartingPoints, values [P])."""
    prof = config.profile
    is_profile = config.run.jobtype == 'profile'
    all_names = list(kernel.param['varying'].keys())
    if is_profile:
        if prof.parameter not in all_names:
            raise KeyError(f'The requested profile parameter {prof.parameter} is not recognized by the kernel. '
                           'Check that your kernel parameter file has the parameter enabled.')
        pidx = all_names.index(prof.parameter)
        values = np.asarray(prof.values, dtype=np.float64)
        names = [n for n in all_names if n != prof.parameter]
    else:
        pidx, values, names = None, np.zeros(0), all_names
    n_pts = max(1, len(values))

    def drop(cov):
        if pidx is None:
            return cov
        return np.delete(np.delete(cov, [pidx], 0), [pidx], 1)

    positions, covmats, indices = [], [], None
    if prof.start_from_position is not None and prof.start_from_covmat is not None:
        pos = np.asarray(kernel.read_initial_position(prof.start_from_position), dtype=np.float64)
        cov = np.asarray(kernel.read_covmat(prof.start_from_covmat), dtype=np.float64)
        if pos.shape[0] == len(all_names) and pidx is not None:
            pos = np.delete(pos, pidx)
        if cov.shape[0] == len(all_names):
            cov = drop(cov)
        positions = [pos] * n_pts
        covmats = [cov] * n_pts
    elif prof.start_from_profile is not None:
        from .io import load_profile
        chain = None
        if prof.start_from_mcmc is not None:
            chain = _reduced_chain(load_mcmc(prof.start_from_mcmc), kernel)
        start_profile = load_profile(prof.start_from_profile, direct_txt=True)
        for p in range(n_pts):
            if chain is not None:
                idx = (get_fraction_of_points_in_bin(chain, values[p], prof.parameter, prof.start_bin_fraction)
                       if is_profile else np.arange(len(chain.mults)))
                x = np.stack([np.take(chain.positions[nm], idx) for nm in all_names], axis=1)
                covmats.append(drop(_weighted_cov_gpu(x, np.take(chain.mults, idx), kernel.device)))
            else:
                covmats.append(drop(np.asarray(kernel.read_covmat(prof.start_from_covmat), dtype=np.float64)))
            if is_profile:
                closest = int(np.argmin(np.abs(start_profile[prof.parameter] - values[p])))
            else:
                closest = int(np.argmin(start_profile['-loglkl']))
            positions.append(np.array([start_profile[nm][closest] for nm in names]))
    elif prof.start_from_mcmc is not None:
        chain = _reduced_chain(load_mcmc(prof.start_from_mcmc), kernel)
        indices = []
        for p in range(n_pts):
            if is_profile:
                idx = get_fraction_of_points_in_bin(chain, values[p], prof.parameter, prof.start_bin_fraction)
            else:
                idx = np.arange(len(chain.mults))
            indices.append(idx)
            x = np.stack([np.take(chain.positions[nm], idx) for nm in all_names], axis=1)
            covmats.append(drop(_weighted_cov_gpu(x, np.take(chain.mults, idx), kernel.device)))
            best = int(np.argmin(np.take(chain.loglkls, idx)))
            start = x[best]
            positions.append(np.delete(start, pidx) if pidx is not None else start)
    else:
        raise ValueError('InitialiseOptimiserTask could not construct an initial position and/or covariance '
                         'matrix. Make sure to supply either "start_from_mcmc" or "start_from_position" and '
                         '"start_from_covmat" in the profile section of the input.')
    positions = np.array(positions, dtype=np.float64)
    covmats = np.array(covmats, dtype=np.float64)
    # initial_loglkl = kernel.loglkl(start) with the profile parameter fixed (initialise_optimiser_task.py:124)
    if is_profile:
        kernel.set_fixed_parameters({prof.parameter: float(values[0])})
        prob = kernel.device_problem(fixed_values=values, proposal_covmats=covmats)
        init_ll = kernel.loglkl_batch(positions, points=np.arange(n_pts), problem=prob)
    else:
        prob = kernel.device_problem(proposal_covmats=covmats)
        init_ll = kernel.loglkl_batch(positions, problem=prob)
    sp = StartingPoints(names, positions, covmats, init_ll, indices)
    sp.problem = prob
    return sp, values


def _reduced_chain(chain, kernel):
    """Keep only the kernel's parameters (initialise_optimiser_task.py:98-114)."""
    for nm in kernel.param['varying']:
        if nm not in chain.positions:
            raise KeyError(f"Parameter '{nm}' not present in the MCMC I'm starting from. Ending initialisation.")
    keep = set(kernel.param['varying']) | set(kernel.param['fixed'])
    for nm in [k for k in chain.positions if k not in keep]:
        del chain.positions[nm]
    return chain
