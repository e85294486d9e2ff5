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


def load_mcmc(mcmc_path, collapse=True, indices=None):
    """Read an unpacked MCMC `<root>_<i>.txt` + `<root>.paramnames` (MontePython / cobaya / GetDist layout:
    multiplicity, -log L, parameters) without GetDist: every chain file in index order, no burn-in removal,
    which is what reproduces the reference's own fixtures.  -> Chain or list[Chain].
    `indices`: only these chain indices (a rank of a sharded, resumed mcmc job reads its own chains)."""
    files = []
    wanted = None if indices is None else set(int(i) for i in indices)
    for f in glob.glob(mcmc_path + '_*.txt'):
        m = re.search(r'_(\d+)\.txt$', f)
        if m and (wanted is None or int(m.group(1)) in wanted):
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


# Above this many MCMC rows the nearest-N selection of every profile value runs on the GPU (one distance matrix, one
# partial sort for all values); below it numpy's argsort is exact, cheap (5 ms for the 2 477-row toy) and is what the
# bit-exact bookkeeping tests pin.
GPU_BINNING_MIN_ROWS = 200_000


def get_fraction_of_points_in_bin(chain, param_val, param_name, fraction):
    n_points = int(np.ceil(fraction * len(chain.mults)))
    return get_points_in_bin(chain, param_val, param_name, n_points)


def get_points_in_bins_gpu(column, values, n_points, device):
    """get_points_in_bin for ALL profile values at once on the device: indices [P, n_points] of the rows nearest to each
    value, nearest first.  Same result as np.argsort(|x - v|)[:n_points] whenever the distances involved are distinct
    (numpy's default argsort is not stable either, so the order inside a tie is not defined by the reference)
    (tasks/initialise_optimiser_task.py:186-196)."""
    import torch
    x = torch.as_tensor(np.asarray(column, dtype=np.float64), device=device)
    v = torch.as_tensor(np.asarray(values, dtype=np.float64), device=device)
    out = []
    for lo in range(0, len(v), 16):                                  # bounded scratch: 16 values x N rows at a time
        dist = (x[None, :] - v[lo:lo + 16, None]).abs()
        out.append(torch.topk(dist, n_points, dim=1, largest=False, sorted=True).indices)
    return torch.cat(out, dim=0).cpu().numpy()


def _weighted_cov_gpu(x, mults, device):
    return _weighted_covs_gpu([x], [mults], device)[0]


def _weighted_covs_gpu(xs, mults, device):
    """Multiplicity-weighted (bias=True) covariances of several row sets -- one per profile value -- with ONE
    upload, one moments launch per set and ONE read-back (the launches are queued back to back; a synchronising
    read per profile value was the bulk of the initialisation time for 32-256 values)."""
    import torch
    from . import engine
    if not xs:
        return []
    n = xs[0].shape[1]
    sizes = [len(x) for x in xs]
    xt = torch.as_tensor(np.ascontiguousarray(np.concatenate(xs, axis=0), dtype=np.float64), device=device)
    wt = torch.as_tensor(np.concatenate([np.asarray(m) for m in mults]).astype(np.int32), device=device)
    out = torch.empty((len(xs), n, n), dtype=torch.float64, device=device)
    lo = 0
    for i, size in enumerate(sizes):
        out[i] = engine.covariance_from_moments(*engine.weighted_moments(xt[lo:lo + size], n, wt[lo:lo + size]))
        lo += size
    return list(out.cpu().numpy())


class StartingPoints:
    """Per profile point: start position (varying order), proposal covmat, initial_loglkl, bin indices."""

    def __init__(self, names_varying, positions, covmats, initial_loglkl, indices=None):
        self.names = names_varying
        self.positions = positions            # [P, n]
        self.covmats = covmats                # [P, n, n]
        self.initial_loglkl = initial_loglkl  # [P]
        self.indices = indices                # list of int arrays or None


def initialise_optimiser(config, kernel, shard=None):
    """`kernel` must not have its profile parameter fixed yet.  Returns (StartingPoints, values [P]).

    `shard` = (rank, world) of a torch.distributed job: the binning / covariance work of the profile values is split
    over the ranks (contiguous blocks of values) and the per-value results are all-gathered, instead of every rank
    repeating all of it (256 values on 8 GPUs)."""
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
    rank, world = shard if shard is not None else (0, 1)
    mine = np.array_split(np.arange(n_pts), world)[rank] if world > 1 else np.arange(n_pts)

    def drop(cov):
        if pidx is None:
            return cov
        return np.delete(np.delete(cov, [pidx], 0), [pidx], 1)

    positions, covmats, indices = {}, {}, None
    if prof.start_from_position is not None and prof.start_from_covmat is not None:
        pos = np.asarray(kernel.read_initial_position(prof.start_from_position), dtype=np.float64)
        cov = np.asarray(kernel.read_covmat(prof.start_from_covmat), dtype=np.float64)
        if pos.shape[0] == len(all_names) and pidx is not None:
            pos = np.delete(pos, pidx)
        if cov.shape[0] == len(all_names):
            cov = drop(cov)
        mine, world = np.arange(n_pts), 1          # nothing to compute: every rank fills all values itself
        positions = {p: pos for p in mine}
        covmats = {p: cov for p in mine}
    elif prof.start_from_profile is not None:
        from .io import load_profile
        chain = None
        if prof.start_from_mcmc is not None:
            chain = _reduced_chain(load_mcmc(prof.start_from_mcmc), kernel)
        start_profile = load_profile(prof.start_from_profile, direct_txt=True)
        xs, ms = [], []
        for p in mine:
            if chain is not None:
                idx = (get_fraction_of_points_in_bin(chain, values[p], prof.parameter, prof.start_bin_fraction)
                       if is_profile else np.arange(len(chain.mults)))
                xs.append(np.stack([np.take(chain.positions[nm], idx) for nm in all_names], axis=1))
                ms.append(np.take(chain.mults, idx))
            else:
                covmats[p] = drop(np.asarray(kernel.read_covmat(prof.start_from_covmat), dtype=np.float64))
            if is_profile:
                closest = int(np.argmin(np.abs(start_profile[prof.parameter] - values[p])))
            else:
                closest = int(np.argmin(start_profile['-loglkl']))
            positions[p] = np.array([start_profile[nm][closest] for nm in names])
        for p, cov in zip(mine, _weighted_covs_gpu(xs, ms, _device_of(kernel))):
            covmats[p] = drop(cov)
    elif prof.start_from_mcmc is not None:
        chain = _reduced_chain(load_mcmc(prof.start_from_mcmc), kernel)
        indices, xs, ms = {}, [], []
        gpu_idx = None
        if is_profile and len(chain.mults) >= GPU_BINNING_MIN_ROWS and len(mine):
            n_bin = int(np.ceil(prof.start_bin_fraction * len(chain.mults)))
            gpu_idx = dict(zip(mine, get_points_in_bins_gpu(chain.positions[prof.parameter], values[mine], n_bin,
                                                            _device_of(kernel))))
        for p in mine:
            if gpu_idx is not None:
                idx = gpu_idx[p]
            elif is_profile:
                idx = get_fraction_of_points_in_bin(chain, values[p], prof.parameter, prof.start_bin_fraction)
            else:
                idx = np.arange(len(chain.mults))
            indices[p] = idx
            x = np.stack([np.take(chain.positions[nm], idx) for nm in all_names], axis=1)
            xs.append(x)
            ms.append(np.take(chain.mults, idx))
            start = x[int(np.argmin(np.take(chain.loglkls, idx)))]
            positions[p] = np.delete(start, pidx) if pidx is not None else start
        for p, cov in zip(mine, _weighted_covs_gpu(xs, ms, _device_of(kernel))):
            covmats[p] = drop(cov)
    else:
        raise ValueError('InitialiseOptimiserTask could not construct an initial position and/or covariance '
                         'matrix. Make sure to supply either "start_from_mcmc" or "start_from_position" and '
                         '"start_from_covmat" in the profile section of the input.')
    if world > 1:           # every rank gets every value's start point / covariance (ALL ranks take part, whatever
                            # their own share -- also a rank that happens to own every value, or none)
        import torch.distributed as dist
        parts = [None] * world
        dist.all_gather_object(parts, (positions, covmats, indices))
        for pos_r, cov_r, idx_r in parts:
            positions.update(pos_r)
            covmats.update(cov_r)
            if indices is not None and idx_r is not None:
                indices.update(idx_r)
    positions = np.array([positions[p] for p in range(n_pts)], dtype=np.float64)
    covmats = np.array([covmats[p] for p in range(n_pts)], dtype=np.float64)
    indices = None if indices is None else [indices[p] for p in range(n_pts)]
    # initial_loglkl = kernel.loglkl(start) with the profile parameter fixed (initialise_optimiser_task.py:124): one
    # batched likelihood launch over all values on a likelihood-only device image (no per-point proposal factors --
    # those are only built for the points a rank actually anneals)
    if getattr(kernel, 'host_evaluated', False):       # hosted.HostKernelAdapter: one host batch
        if is_profile:
            kernel.set_fixed_parameters({prof.parameter: float(values[0])})
        init_ll = kernel.loglkl_batch(positions, values if is_profile else None)
    elif is_profile:
        kernel.set_fixed_parameters({prof.parameter: float(values[0])})
        prob = kernel.device_problem(fixed_values=values, likelihood_only=True)
        init_ll = kernel.loglkl_batch(positions, points=np.arange(n_pts), problem=prob)
    else:
        prob = kernel.device_problem(likelihood_only=True)
        init_ll = kernel.loglkl_batch(positions, problem=prob)
    covmats.setflags(write=False)      # results, not scratch: lets kernels.array_key cache device problems by identity
    return StartingPoints(names, positions, covmats, init_ll, indices), values


def _device_of(kernel):
    """Device of the kernel's GPU work (a host-evaluated kernel may leave it open: the current CUDA device)."""
    dev = getattr(kernel, 'device', None)
    if dev is None:
        import torch
        dev = f'cuda:{torch.cuda.current_device()}'
    return dev


def _reduced_chain(chain, kernel):
    """Keep only the kernel's parameters (initialise_optimiser_task.py:98-114)."""
    for nm in kernel.param['varying']:
        if nm not in chain.positions:
            raise KeyError(f"Parameter '{nm}' not present in the MCMC I'm starting from. Ending initialisation.")
    keep = set(kernel.param['varying']) | set(kernel.param['fixed'])
    for nm in [k for k in chain.positions if k not in keep]:
        del chain.positions[nm]
    return chain
