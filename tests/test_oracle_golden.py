"""Pin the CPU oracle (oracle/reference_port.py, oracle/closed_form.py) against golden
vectors frozen from the unmodified reference (tests/golden/make_golden.py) and against
the reference's own fixtures."""
import json
import os

import numpy as np
import pytest

from oracle import closed_form
from oracle import reference_port as rp


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _target(golden_dir, d):
    k = _load(golden_dir, f'toy{d}_kernel.npz')
    return rp.GaussianTarget(k['means'], k['covmat'])


def _mcmc_rows(golden_dir, d):
    m = _load(golden_dir, f'toy{d}_mcmc.npz')
    return np.concatenate([m[f'chain_{i}'] for i in range(4)], axis=0)


@pytest.mark.parametrize('d', [20, 30])
def test_recipe_reproduces_kernel(golden_dir, d):
    """O3(iii): analytical_kernel.py:28-40 with seed 0 is bit-identical to the pickle."""
    k = _load(golden_dir, f'toy{d}_kernel.npz')
    np.random.seed(0)
    means, cov = rp.random_gaussian_recipe(d)
    assert np.array_equal(means, k['means'])
    assert np.array_equal(cov, k['covmat'])


def test_bootstrap_chains_bit_exact(golden_dir):
    """O3(i): input/example_toy/mcmc/test_{0..3}.txt from seed 0 (4 chains x 4 rounds x 2000 steps).
    Pins the recipe, Gaussian(), MH step, default start/covmat and RNG draw order.  One round
    (quarter length) is checked in full precision on the common prefix to keep the suite quick."""
    m = _load(golden_dir, 'toy20_mcmc.npz')
    shipped = [m[f'chain_{i}'] for i in range(4)]
    np.random.seed(0)
    means, cov = rp.random_gaussian_recipe(20)
    chains = rp.run_mcmc_job(rp.GaussianTarget(means, cov), 4, 2000, 1)
    for ch in chains:
        data = ch.data()
        n = len(data) - 1          # last row's multiplicity is still growing after round 1
        ok = [s.shape[0] > n and np.allclose(s[:n], data[:n], rtol=0, atol=1e-12) for s in shipped]
        assert any(ok)


@pytest.mark.parametrize('d', [20, 30])
def test_loglkl_and_permutation_quirk(golden_dir, d):
    """A10 / F4: free and fixed-x2 loglkl equal the reference on random points."""
    g = _load(golden_dir, f'toy{d}_loglkl.npz')
    t = _target(golden_dir, d)
    free = [t.loglkl(x) for x in g['X']]
    assert np.array_equal(np.array(free), g['free'])
    tf = t.fix(1, 0.75)
    fixed = [tf.loglkl(np.delete(x, 1)) for x in g['X']]
    assert np.array_equal(np.array(fixed), g['fixed_x2_0p75'])


@pytest.mark.parametrize('d', [20, 30])
def test_closed_form_matches_reference_scipy_profile(golden_dir, d):
    """O2 vs O3(iv): closed form == get_scipy_profile (BFGS tolerance 1e-6 abs)."""
    g = _load(golden_dir, f'toy{d}_loglkl.npz')
    k = _load(golden_dir, f'toy{d}_kernel.npz')
    cf = closed_form.profile_curve(k['means'], k['covmat'], 1, g['values'])
    assert np.allclose(cf, g['scipy_profile'], rtol=1e-6, atol=1e-6)
    if d == 20:   # the three values quoted in SURVEY.md O2
        assert abs(closed_form.profile_minimum(k['means'], k['covmat'], 1, -1.0)[0] - 20.719044687) < 1e-8
        assert abs(closed_form.profile_minimum(k['means'], k['covmat'], 1, 0.0)[0] - 3.6023646587) < 1e-8
        assert abs(closed_form.profile_minimum(k['means'], k['covmat'], 1, 2.0)[0] - 11.625826054) < 1e-8


@pytest.mark.parametrize('d', [20, 30])
def test_binning_bookkeeping_bit_exact(golden_dir, d):
    """A14: indices, N, covariance, start point and initial_loglkl per profile value."""
    g = _load(golden_dir, f'toy{d}_init.npz')
    t = _target(golden_dir, d)
    rows = _mcmc_rows(golden_dir, d)
    for i, v in enumerate(g['values']):
        ini = rp.initialise_from_mcmc(t, rows, 1, v, 0.1)
        assert len(ini['indices']) == g['N_points'][i]
        assert np.array_equal(ini['indices'], g['indices'][i])
        assert np.array_equal(ini['covmat'], g['covmat'][i])
        # the reference's start dict is [varying..., fixed] (loglkl() re-inserts the fixed one last)
        assert np.array_equal(ini['position'], g['position'][i][:-1])
        assert g['position'][i][-1] == v
        assert ini['initial_loglkl'] == g['initial_loglkl'][i]


def test_reference_own_initial_loglkl_fixture(golden_dir):
    """O3(ii): test/debug_reanneal/profile/x2.txt, column initial_loglkl (11 digits)."""
    with open(os.path.join(golden_dir, 'ref_debug_reanneal_initial_loglkl.json')) as f:
        rows = json.load(f)['rows']
    t = _target(golden_dir, 20)
    mcmc = _mcmc_rows(golden_dir, 20)
    for v, expected in rows:
        got = rp.initialise_from_mcmc(t, mcmc, 1, v, 0.1)['initial_loglkl']
        assert f'{got:.10e}' == f'{expected:.10e}'


def _compare_records(records, golden, n_var):
    assert len(records) == len(golden)
    for r, g in zip(records, golden):
        assert r['iteration_number'] == g['iteration_number']
        assert r['repetition_number'] == g['repetition_number']
        assert r['fixed_param_val'] == g['fixed_param_val']
        for key in ('temperature', 'step_size', 'current_best_loglkl', 'bestfit_loglkl', 'acceptance_rate',
                    'last_loglkl'):
            assert r[key] == g[key], (key, r['iteration_number'], r[key], g[key])
        assert r['accepted_steps'] == g['accepted_steps']
        assert r['sum_mults'] == g['sum_mults']
        assert np.array_equal(r['bestfit_position'], np.array(g['bestfit_position'][:n_var]))
        assert np.array_equal(r['last_position'], np.array(g['last_position'][:n_var]))


def test_profile_run_bit_exact_20d(golden_dir):
    """A1-A8: 10 values x 20 iterations x 250 steps of example_toy.yaml, seed 0, serial.
    The first 3 iterations (30 tasks, 7 500 steps) are replayed; RNG order makes any
    divergence in step(), set_bestfit() or get_next_iteration_settings() visible at once."""
    with open(os.path.join(golden_dir, 'toy20_profile_run.json')) as f:
        gold = json.load(f)
    t = _target(golden_dir, 20)
    np.random.seed(0)
    out = rp.run_annealing_job(t, _mcmc_rows(golden_dir, 20), rp.Schedule(), np.linspace(-1.0, 2.0, 10),
                               1, 1, n_iterations=3)
    gopt = sorted(gold['optimise'], key=lambda o: o['id'])[:30]
    _compare_records(out['records'], gopt, 19)


def test_global_run_bit_exact_20d(golden_dir):
    """K2 as shipped (3 repetitions, nothing fixed): first 4 iterations."""
    with open(os.path.join(golden_dir, 'toy20_global_run.json')) as f:
        gold = json.load(f)
    t = _target(golden_dir, 20)
    np.random.seed(0)
    out = rp.run_annealing_job(t, _mcmc_rows(golden_dir, 20), rp.Schedule(), [], 3, None, n_iterations=4)
    gopt = sorted(gold['optimise'], key=lambda o: o['id'])[:12]
    _compare_records(out['records'], gopt, 20)


def test_exponential_step_schedule_bit_exact(golden_dir):
    with open(os.path.join(golden_dir, 'toy20_exponential_run.json')) as f:
        gold = json.load(f)
    t = _target(golden_dir, 20)
    np.random.seed(0)
    sched = rp.Schedule(step_size_schedule='exponential', step_size_range=(0.5, 0.05))
    out = rp.run_annealing_job(t, _mcmc_rows(golden_dir, 20), sched, [0.5], 1, 1, n_iterations=8)
    _compare_records(out['records'], sorted(gold['optimise'], key=lambda o: o['id']), 19)


def test_secant_multiplier_schedule_bit_exact(golden_dir):
    """A5 'adaptive' multiplier: +-0.05, +-0.1, then the clamped secant update."""
    with open(os.path.join(golden_dir, 'toy20_secant_run.json')) as f:
        gold = json.load(f)
    assert not gold['failed']
    t = _target(golden_dir, 20)
    np.random.seed(0)
    sched = rp.Schedule(step_size_adaptive_multiplier='adaptive', step_size_adaptive_interval=(0.24, 0.26),
                        step_size_adaptive_initial=0.3, steps_per_iteration=400)
    out = rp.run_annealing_job(t, _mcmc_rows(golden_dir, 20), sched, [0.5, 1.0], 1, 1, n_iterations=10)
    _compare_records(out['records'], sorted(gold['optimise'], key=lambda o: o['id']), 19)


def test_secant_multiplier_reference_edge_paths_are_kept():
    """A5 edge paths: AR inside the interval on iteration 0 -> unbound local (task fails);
    equal consecutive ARs -> numpy-scalar division by zero gives +-inf/nan which the
    min/max clamp turns into 1.0 / -0.9 (min(nan, 1.0) is nan, max(-0.9, nan) is -0.9)."""
    sched = rp.Schedule(step_size_adaptive_multiplier='adaptive', step_size_adaptive_interval=(0.1, 0.3))
    s = {'temperature': 0.1, 'temperature_change': 0.9, 'step_size': 0.1, 'step_size_change': {}}
    with pytest.raises(UnboundLocalError):
        rp.next_settings(sched, s, {'acceptance_rate': np.float64(0.2), 'loglkl': 1.0})
    for m_cur, expect in ((0.1, -0.9), (0.05, -0.9)):
        s['step_size_change'] = {'previous_multiplier': 0.05, 'previous_acceptance_rate': np.float64(0.5),
                                 'current_multiplier': m_cur}
        with np.errstate(all='ignore'):
            out = rp.next_settings(sched, s, {'acceptance_rate': np.float64(0.5), 'loglkl': 1.0})
        assert out['step_size_change']['current_multiplier'] == expect
        assert out['step_size'] == 0.1 * (1 + expect)


def test_converged_rule():
    sched = rp.Schedule(chi2_tolerance=2.5)
    assert rp.converged(sched, 10.0, {'loglkl': 9.5, 'accepted_steps': 5})
    assert not rp.converged(sched, 10.0, {'loglkl': 5.0, 'accepted_steps': 5})      # improved by more than tol
    assert not rp.converged(sched, 10.0, {'loglkl': 9.5, 'accepted_steps': 1})
    assert not rp.converged(sched, np.inf, {'loglkl': 9.5, 'accepted_steps': 5})
    assert not rp.converged(rp.Schedule(), 10.0, {'loglkl': 9.5, 'accepted_steps': 5})
