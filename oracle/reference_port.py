"""CPU restatement of PROSPECT's annealing / Metropolis-Hastings hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `prospect_public_b200/` may import this
module: only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline legs of
`bench.py` use it, and only as the checker / reported baseline.

Parity status: PINNED.  With the same numpy legacy-RNG seed this port consumes
random numbers in exactly the order the reference does and reproduces, bit for
bit, the golden vectors frozen from the unmodified reference in tests/golden/
(made by tests/golden/make_golden.py): the shipped bootstrap chains
input/example_toy/mcmc/test_{0..3}.txt, the per-value binning indices /
covariances / start points of InitialiseOptimiserTask, every OptimiseTask
bestfit / acceptance rate / next step size of a 20-iteration serial profile run,
and the reference's own `initial_loglkl` column in test/debug_reanneal/profile/x2.txt
(tests/test_oracle_golden.py).

Every function cites the reference lines (relative to /root/reference) it restates.
State is held in flat numpy arrays instead of the reference's dict-of-lists.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


# --------------------------------------------------------------------------------------
# Target: prospect/kernels/analytical_kernel.py
# --------------------------------------------------------------------------------------
def random_gaussian_recipe(dimension: int, std_scale: float = 0.1, off_diag_factor: float = 0.05):
    """analytical_kernel.py:28-40 -- draws from the *global* legacy numpy RNG:
    d scalar rand() for the means, rand(d) for the diagonal, rand(d, d) for the rest."""
    means = np.array([np.random.rand() for _ in range(dimension)])
    diag = np.diag(np.random.rand(dimension))
    offdiag = np.random.rand(dimension, dimension)
    covmat = std_scale * (diag + offdiag * off_diag_factor)
    return means, covmat


@dataclass
class GaussianTarget:
    """The `random_gaussian` / `gaussian` analytic likelihood with its prior box.

    `fixed` is the index of the profiled parameter (None when nothing is fixed);
    varying parameters keep their x1..xd order with the fixed one removed
    (base_kernel.py:46-55)."""
    means: np.ndarray
    covmat: np.ndarray
    lo: np.ndarray = None
    hi: np.ndarray = None
    fixed: int | None = None
    fixed_value: float = 0.0

    def __post_init__(self):
        d = len(self.means)
        if self.lo is None:
            self.lo = np.full(d, -2.5)      # analytical_kernel.py:109-114
        if self.hi is None:
            self.hi = np.full(d, 2.5)

    @property
    def dimension(self):
        return len(self.means)

    @property
    def varying(self):
        return [i for i in range(self.dimension) if i != self.fixed]

    def fix(self, index, value):
        """base_kernel.py:46-51."""
        return GaussianTarget(self.means, self.covmat, self.lo, self.hi, index, float(value))

    def loglkl(self, x_varying):
        """-log L.  analytical_kernel.py:59-66: the residual is [varying..., fixed] but is
        contracted with inv(covmat) in the ORIGINAL x1..xd order; the inverse is
        recomputed on every call exactly as the reference does."""
        v = self.varying
        residual = list(np.asarray(x_varying, dtype=float) - self.means[v])
        if self.fixed is not None:
            residual.append(self.fixed_value - self.means[self.fixed])
        return 0.5 * np.dot(residual, np.matmul(np.linalg.inv(self.covmat), residual))

    def logprior(self, x_varying):
        """base_kernel.py:83-102: +inf outside the box (strict <, >), fixed parameter unchecked."""
        v = self.varying
        for val, lo, hi in zip(x_varying, self.lo[v], self.hi[v]):
            if lo is not None and val < lo:
                return np.inf
            if hi is not None and val > hi:
                return np.inf
        return 0.0

    def default_covmat(self):
        """analytical_kernel.py:98-103 (used as a covariance, not as stds)."""
        v = self.varying
        return 0.005 * np.diag(self.hi[v] - self.lo[v])

    def default_start(self):
        """analytical_kernel.py:88-92."""
        return np.zeros(len(self.varying))


# --------------------------------------------------------------------------------------
# Chain + sampler: prospect/mcmc.py
# --------------------------------------------------------------------------------------
@dataclass
class Chain:
    """mcmc.py:25-82 as three growing lists; positions are rows of the varying parameters."""
    mults: list = field(default_factory=list)
    loglkls: list = field(default_factory=list)
    positions: list = field(default_factory=list)

    def push(self, loglkl, x):
        """mcmc.py:69-73."""
        self.mults.append(1)
        self.loglkls.append(loglkl)
        self.positions.append(np.array(x, dtype=float))

    def data(self, fixed_column=None):
        """mcmc.py:58-63: rows (mult, loglkl, params...).  `fixed_column` = (value,) appends the
        fixed parameter as the last column, which is where the reference's dict order puts it."""
        pos = np.array(self.positions)
        cols = [np.array(self.mults, dtype=float)[:, None], np.array(self.loglkls)[:, None], pos]
        if fixed_column is not None:
            cols.append(np.full((len(self.mults), 1), fixed_column))
        return np.concatenate(cols, axis=1)


def weighted_covariance(positions, mults):
    """mcmc.py:93-97.  The reference hands np.cov the transpose of a parameter-major array
    (Chain.data, mcmc.py:58-63); the same memory layout is used here because BLAS summation
    order -- hence the last bits -- depends on it."""
    param_major = np.ascontiguousarray(np.asarray(positions, dtype=float).T)
    return np.cov(param_major.T, rowvar=False, bias=True, fweights=list(mults))


def mh_step(target: GaussianTarget, chain: Chain, covmat, temperature, step_size):
    """mcmc.py:163-190.  RNG order per step: multivariate_normal (n normals through numpy's
    SVD factor of covmat*step^2), then one uniform which is drawn even when the proposal is
    outside the prior box."""
    current = chain.positions[-1]
    prop = np.random.multivariate_normal(current, covmat * step_size ** 2)
    logprior_prop = target.logprior(prop)
    if logprior_prop == np.inf:
        acceptance = 0
    else:
        loglkl_prop = target.loglkl(prop)
        logpost_current = chain.loglkls[-1] + target.logprior(current)
        logpost_prop = loglkl_prop + logprior_prop
        with np.errstate(over='ignore'):
            acceptance = np.exp((logpost_current - logpost_prop) / temperature)
    rand = np.random.uniform(low=0, high=1)
    if acceptance > rand:
        chain.push(loglkl_prop, prop)
        return True
    chain.mults[-1] += 1
    return False


def run_steps(target, chain, covmat, temperature, step_size, n_steps):
    """mcmc.py:131-138 without the per-step timers."""
    accepted = 0
    for _ in range(n_steps):
        accepted += mh_step(target, chain, covmat, temperature, step_size)
    return accepted


# --------------------------------------------------------------------------------------
# Simulated annealing: prospect/optimiser.py + tasks/optimise_task.py
# --------------------------------------------------------------------------------------
@dataclass
class Schedule:
    """The `profile:` yaml knobs that steer one chain (profile.py:129-362)."""
    temperature_range: tuple = (0.1, 0.001)
    max_iterations: int = 20
    step_size_schedule: str = 'adaptive'            # 'adaptive' | 'exponential'
    step_size_range: tuple | None = None            # exponential only
    step_size_adaptive_interval: tuple = (0.19, 0.21)
    step_size_adaptive_multiplier: float | str = 0.3   # number or 'adaptive'
    step_size_adaptive_initial: float = 0.1
    steps_per_iteration: int = 250
    chi2_tolerance: float | None = None

    def temperature_change(self):
        """initialise_optimiser_task.py:133-141."""
        return (self.temperature_range[1] / self.temperature_range[0]) ** (1 / self.max_iterations)

    def initial_step_size(self):
        """profile.py:205-213: adaptive schedules start at step_size_adaptive_initial."""
        if self.step_size_schedule == 'adaptive':
            return self.step_size_adaptive_initial
        return self.step_size_range[0]

    def initial_step_size_change(self):
        """initialise_optimiser_task.py:143-156."""
        if self.step_size_schedule == 'exponential':
            return (self.step_size_range[1] / self.step_size_range[0]) ** (1 / self.max_iterations)
        if self.step_size_adaptive_multiplier == 'adaptive':
            return {}
        return None


def anneal_iteration(target, start, covmat, temperature, step_size, n_steps):
    """optimiser.py:40-85 + optimise_task.py:16-27: seed the chain with the start point
    (multiplicity 1), run the steps, pick the first minimum."""
    chain = Chain()
    chain.push(target.loglkl(start), start)
    run_steps(target, chain, covmat, temperature, step_size, n_steps)
    best = int(np.argmin(chain.loglkls))
    return {
        'loglkl': chain.loglkls[best],
        'acceptance_rate': len(chain.loglkls) / np.sum(chain.mults),
        'accepted_steps': len(chain.loglkls),
        'position': chain.positions[best].copy(),
        'last_position': chain.positions[-1].copy(),
        'last_loglkl': chain.loglkls[-1],
        'sum_mults': int(np.sum(chain.mults)),
        'chain': chain,
    }


def next_settings(schedule: Schedule, settings: dict, bestfit: dict):
    """optimiser.py:87-162.  `settings` holds temperature, temperature_change, step_size,
    step_size_change.  The reference's edge paths are kept: with the 'adaptive' multiplier an
    acceptance rate inside the interval on iteration 0/1 hits an unbound local (the task fails);
    equal consecutive acceptance rates divide numpy scalars by zero (+-inf / nan, a warning, no
    exception) and the min/max clamp then yields 1.0 or -0.9."""
    new_temp = settings['temperature'] * settings['temperature_change']
    ar = bestfit['acceptance_rate']
    lo, hi = schedule.step_size_adaptive_interval
    change = settings['step_size_change']
    if schedule.step_size_schedule == 'exponential':
        new_step = settings['step_size'] * change
    elif schedule.step_size_adaptive_multiplier == 'adaptive':
        if 'previous_multiplier' not in change:
            if ar > hi:
                m = 0.05
            elif ar < lo:
                m = -0.05
            else:
                raise UnboundLocalError('initial_multiplier')      # optimiser.py:99-103
            new_step = settings['step_size'] * (1 + m)
            change['previous_multiplier'] = m
            change['previous_acceptance_rate'] = ar
        elif 'current_multiplier' not in change:
            if ar > hi:
                m = 0.1
            elif ar < lo:
                m = -0.1
            else:
                raise UnboundLocalError('new_multiplier')          # optimiser.py:108-112
            new_step = settings['step_size'] * (1 + m)
            change['current_multiplier'] = m
            change['current_acceptance_rate'] = ar
        else:
            ar_desired = np.mean(schedule.step_size_adaptive_interval)
            ar_prev = change['previous_acceptance_rate']
            m_cur = change['current_multiplier']
            m_prev = change['previous_multiplier']
            m_new = (ar_desired - ar) / (ar - ar_prev) * (m_cur - m_prev) + m_cur   # inf/nan on equal ARs, as the reference
            m_new = max(-0.9, min(m_new, 1.0))
            new_step = settings['step_size'] * (1 + m_new)
            change = {'previous_acceptance_rate': ar, 'current_multiplier': m_new,
                      'previous_multiplier': m_cur}
    else:
        mult = schedule.step_size_adaptive_multiplier
        if ar > hi:
            new_step = settings['step_size'] * (1 + mult)
        elif ar < lo:
            new_step = settings['step_size'] * (1 - mult)
        else:
            new_step = settings['step_size']
    return {
        'current_best_loglkl': bestfit['loglkl'],
        'temperature': new_temp,
        'temperature_change': settings['temperature_change'],
        'step_size': new_step,
        'step_size_change': change,
    }


def converged(schedule: Schedule, current_best_loglkl, bestfit):
    """optimise_task.py:33-50: the chi2_tolerance stop (None/0 => never)."""
    tol = schedule.chi2_tolerance
    if tol is None or tol == 0.0:
        return False
    improvement = current_best_loglkl - bestfit['loglkl']
    if improvement < 0.0 or current_best_loglkl == np.inf or 2 * improvement > tol:
        return False
    return bestfit['accepted_steps'] > 1


# --------------------------------------------------------------------------------------
# Initialisation from a finished MCMC: tasks/initialise_optimiser_task.py
# --------------------------------------------------------------------------------------
def nearest_fraction_indices(column, value, fraction):
    """initialise_optimiser_task.py:186-200: N = ceil(fraction * rows); the N rows whose
    profile-parameter value is nearest to `value` in numpy's default argsort order."""
    n_points = int(np.ceil(fraction * len(column)))
    return np.argsort(np.abs(np.asarray(column) - value))[:n_points], n_points


def initialise_from_mcmc(target: GaussianTarget, mcmc_rows, profile_index, value, fraction):
    """initialise_optimiser_task.py:27-76,94-130 for `start_from_mcmc`.

    mcmc_rows: all chain files concatenated, columns (mult, loglkl, x1..xd).
    profile_index=None is global optimisation: every row, nothing removed."""
    mults, loglkls, params = mcmc_rows[:, 0], mcmc_rows[:, 1], mcmc_rows[:, 2:]
    if profile_index is None:
        idx = np.arange(len(mults))
        fixed_target = target
    else:
        idx, _ = nearest_fraction_indices(params[:, profile_index], value, fraction)
        fixed_target = target.fix(profile_index, value)
    sub_m, sub_l, sub_p = np.take(mults, idx), np.take(loglkls, idx), np.take(params, idx, axis=0)
    cov = weighted_covariance(sub_p, sub_m)
    best = int(np.argmin(sub_l))
    start = sub_p[best]
    if profile_index is not None:
        cov = np.delete(np.delete(cov, [profile_index], 0), [profile_index], 1)
        start = np.delete(start, profile_index)
    return {
        'indices': idx, 'covmat': cov, 'position': start,
        'initial_loglkl': fixed_target.loglkl(start), 'target': fixed_target,
    }


# --------------------------------------------------------------------------------------
# Whole jobs in the reference's serial task order (communication.py:50-68, base_task.py:57-63)
# --------------------------------------------------------------------------------------
def run_annealing_job(target, mcmc_rows, schedule: Schedule, values, repetitions, profile_index,
                      fraction=0.1, n_iterations=None, on_task=None):
    """Serial-mode order of RNG consumption: every chain's iteration k (rep-major, value-minor,
    initialise_profile_task.py:16-26) before any chain's iteration k+1.  Iterations run
    k = 0 .. n_iterations-1 (the reference itself never stops)."""
    n_iterations = schedule.max_iterations if n_iterations is None else n_iterations
    chains = []
    points = [None] if profile_index is None else list(values)
    inits = {v: initialise_from_mcmc(target, mcmc_rows, profile_index, v, fraction) for v in points}
    for rep in range(repetitions):
        for v in points:
            ini = inits[v]
            chains.append({
                'value': v, 'rep': rep, 'target': ini['target'], 'covmat': ini['covmat'],
                'start': ini['position'].copy(), 'done': False,
                'settings': {'current_best_loglkl': ini['initial_loglkl'],
                             'temperature': schedule.temperature_range[0],
                             'temperature_change': schedule.temperature_change(),
                             'step_size': schedule.initial_step_size(),
                             'step_size_change': schedule.initial_step_size_change()},
            })
    records = []
    for k in range(n_iterations):
        for c in chains:
            if c['done']:
                continue
            s = c['settings']
            bf = anneal_iteration(c['target'], c['start'], c['covmat'], s['temperature'], s['step_size'],
                                  schedule.steps_per_iteration)
            rec = {'fixed_param_val': c['value'], 'repetition_number': c['rep'], 'iteration_number': k,
                   'temperature': s['temperature'], 'step_size': s['step_size'],
                   'current_best_loglkl': s['current_best_loglkl'],
                   'bestfit_loglkl': bf['loglkl'], 'acceptance_rate': bf['acceptance_rate'],
                   'accepted_steps': bf['accepted_steps'], 'bestfit_position': bf['position'],
                   'last_position': bf['last_position'], 'last_loglkl': bf['last_loglkl'],
                   'sum_mults': bf['sum_mults']}
            records.append(rec)
            if on_task is not None:
                on_task(rec, bf)
            if converged(schedule, s['current_best_loglkl'], bf):
                c['done'] = True
                continue
            c['settings'] = next_settings(schedule, s, bf)
            c['start'] = bf['last_position']
    return {'inits': inits, 'records': records}


def run_mcmc_job(target, n_chains, steps_per_round, n_rounds, temperature=1.0, step_size=1.0,
                 covmat=None, start=None):
    """jobtype mcmc, tasks/mcmc_task.py:17-23,63-81: every round continues each chain in turn."""
    covmat = target.default_covmat() if covmat is None else covmat
    chains = []
    for _ in range(n_rounds):
        for c in range(n_chains):
            if len(chains) <= c:
                x0 = target.default_start() if start is None else np.asarray(start, dtype=float)
                ch = Chain()
                ch.push(target.loglkl(x0), x0)      # mcmc.py:117-125
                chains.append(ch)
            run_steps(target, chains[c], covmat, temperature, step_size, steps_per_round)
    return chains


def bare_loop_rate(target, covmat, start, temperature, step_size, n_steps):
    """Single-core chain-steps/s of run_steps (the figure BASELINE.md calls 'bare MH loop')."""
    import time
    chain = Chain()
    chain.push(target.loglkl(start), start)
    t0 = time.perf_counter()
    run_steps(target, chain, covmat, temperature, step_size, n_steps)
    return n_steps / (time.perf_counter() - t0)
