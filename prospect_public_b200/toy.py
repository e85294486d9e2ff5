"""Synthetic toy inputs in PROSPECT's on-disk formats (kernel pickle, kernel-settings yaml, bootstrap MCMC
chain files) and ready-made yaml dicts for them.  Used by the examples, the benchmark and the tests; the
file formats are those of input/example_toy/ in the reference."""
import os
import pickle

import numpy as np
import yaml


def write_toy_inputs(folder, means, covmat, chains, root='test'):
    """-> dict(kernel_pkl, kernel_yaml, mcmc_root).  chains: list of [N_i, 2 + d] arrays (mult, loglkl, x)."""
    os.makedirs(os.path.join(folder, 'mcmc'), exist_ok=True)
    d = len(means)
    kpkl = os.path.join(folder, 'random_gauss.pkl')
    with open(kpkl, 'wb') as f:
        pickle.dump({'means': {f'x{i + 1}': float(means[i]) for i in range(d)}, 'covmat': np.asarray(covmat),
                     'dimension': d}, f)
    kyaml = os.path.join(folder, 'kernel_load_random_gauss.yaml')
    with open(kyaml, 'w') as f:
        yaml.dump({'function': 'random_gaussian', 'load': kpkl}, f)
    mroot = os.path.join(folder, 'mcmc', root)
    for i, c in enumerate(chains or []):
        np.savetxt(f'{mroot}_{i}.txt', c, delimiter=' ')
    with open(mroot + '.paramnames', 'w') as f:
        for i in range(1, d + 1):
            f.write(f'\nx{i} x{i}')
    with open(mroot + '.ranges', 'w') as f:
        for i in range(1, d + 1):
            f.write(f'\nx{i} -2.5 2.5')
    return {'kernel_pkl': kpkl, 'kernel_yaml': kyaml, 'mcmc_root': mroot}


def annealing_yaml(paths, out_dir, jobtype='profile', write=True, seed=0, **profile):
    """The shipped example_toy.yaml / example_global_optimisation_toy.yaml with overridable profile knobs."""
    prof = {
        'optimiser': 'simulated annealing', 'temperature_schedule': 'exponential',
        'temperature_range': [0.1, 0.001], 'max_iterations': 20, 'step_size_schedule': 'adaptive',
        'step_size_adaptive_interval': [0.19, 0.21], 'step_size_adaptive_multiplier': 0.3,
        'step_size_adaptive_initial': 0.1, 'steps_per_iteration': 250, 'repetitions': 1,
        'start_from_mcmc': paths['mcmc_root'], 'start_bin_fraction': 0.1, 'plot_profile': False,
        'plot_schedule': False,
    }
    if jobtype == 'profile':
        prof.update({'parameter': 'x2', 'values': 'np.linspace(-1.0, 2.0, 10)'})
    prof.update(profile)
    return {
        'io': {'jobname': 'example_toy', 'write': write, 'dir': out_dir, 'overwrite_dir': True,
               'snapshot_interval': 10.0},
        'run': {'jobtype': jobtype, 'mode': 'gpu', 'numpy_random_seed': seed},
        'kernel': {'type': 'analytical', 'conf': '', 'param': paths['kernel_yaml']},
        'profile': prof,
    }


def mcmc_yaml(paths, out_dir, n_chains=4, steps=2000, rm1=0.01, write=True, seed=0, **mcmc):
    sec = {'algorithm': 'MetropolisHastings', 'N_chains': n_chains, 'steps_per_iteration': steps,
           'convergence_Rm1': rm1, 'unpack_at_dump': True}
    sec.update(mcmc)
    return {
        'io': {'jobname': 'test', 'write': write, 'dir': out_dir, 'overwrite_dir': True},
        'run': {'jobtype': 'mcmc', 'mode': 'gpu', 'numpy_random_seed': seed},
        'kernel': {'type': 'analytical', 'conf': '', 'param': paths['kernel_yaml']},
        'mcmc': sec,
    }


def spd_target(d, seed=0, lam_range=(1e-3, 1e-1)):
    """Synthetic correlated SPD Gaussian for dimensions where the reference recipe is indefinite (d >= 64,
    SURVEY.md F3): C = Q diag(lam) Q^T with lam log-uniform and Q from the QR of a seeded Gaussian matrix,
    means ~ U(0, 1).  Returns (means [d], covmat [d, d])."""
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    lam = np.exp(rng.uniform(np.log(lam_range[0]), np.log(lam_range[1]), d))
    cov = (q * lam) @ q.T
    return rng.uniform(0.0, 1.0, d), 0.5 * (cov + cov.T)


def conditional_covariance(covmat, n_keep=None):
    """Covariance of the first n slots given the last one in the reference's residual ordering: inv(S[:n, :n]),
    S = sym(inv(C)).  The ideal proposal covariance for a profile point (closed form, no bootstrap MCMC)."""
    s = np.linalg.inv(covmat)
    s = 0.5 * (s + s.T)
    n = covmat.shape[0] - 1 if n_keep is None else n_keep
    return np.linalg.inv(s[:n, :n])
