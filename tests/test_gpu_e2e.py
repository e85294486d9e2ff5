"""End-to-end GPU tests through the reference-facing host API (yaml -> run() -> PROSPECT output files).

Tolerances are the north-star's: profiled chi^2 within Delta chi^2 <= 1e-3 of the analytic profile and of the
reference CPU run on identical inputs (frozen in tests/golden/toy20_cold_reference_run.json); bestfit
parameters within 1e-4 relative on the cold schedule; MH marginals by KS at p > 1e-3 / d (Bonferroni).
"""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import closed_form  # noqa: E402
from prospect_public_b200 import capi  # noqa: E402
from prospect_public_b200.config import Configuration  # noqa: E402
from prospect_public_b200.io import load_profile  # noqa: E402
from prospect_public_b200.run import analyse, run  # noqa: E402
from prospect_public_b200.toy import annealing_yaml, mcmc_yaml, write_toy_inputs  # noqa: E402
from tests.helpers import GOLDEN, golden, spd_target  # noqa: E402


def _inputs(tmp_path, d):
    k, m = golden(f'toy{d}_kernel.npz'), golden(f'toy{d}_mcmc.npz')
    paths = write_toy_inputs(str(tmp_path / f'inp{d}'), k['means'], k['covmat'], [m[f'chain_{i}'] for i in range(4)])
    return paths, k


def test_example_toy_profile_writes_prospect_layout(tmp_path):
    """K1: input/example_toy/example_toy.yaml as shipped (10 values x 1 rep x 20 it x 250 steps)."""
    paths, k = _inputs(tmp_path, 20)
    out_dir = str(tmp_path / 'out')
    res = run(annealing_yaml(paths, out_dir))
    for fn in ('log.yaml', 'config.pkl', 'state.pkl', 'status.txt', 'analytical/random_gauss.pkl', 'profile/x2.txt',
               'profile/x2_status.txt', 'profile/x2_intervals.txt'):
        assert os.path.isfile(os.path.join(out_dir, fn)), fn
    assert res['chain_steps'] == 10 * 20 * 250
    prof = load_profile(out_dir)                      # the reference's reader contract
    assert np.allclose(prof['x2'], np.linspace(-1.0, 2.0, 10), atol=1e-10)
    cf = closed_form.profile_curve(k['means'], k['covmat'], 1, np.linspace(-1.0, 2.0, 10))
    excess = prof['-loglkl'] - cf
    # T_final = 1.26e-3: the reference itself ends 5e-3 .. 1.3e-2 above the minimum
    assert np.all(excess > -1e-9) and np.all(excess < 5e-2)
    ini = golden('toy20_init.npz')
    assert np.allclose(prof['initial_loglkl'], ini['initial_loglkl'], rtol=1e-10)
    # reported chi^2 is exact for the reported position
    from oracle import reference_port as rp
    t = rp.GaussianTarget(k['means'], k['covmat'])
    names = [f'x{i}' for i in range(1, 21) if i != 2]
    for row in (0, 5, 9):
        x = np.array([prof[nm][row] for nm in names])
        assert abs(t.fix(1, prof['x2'][row]).loglkl(x) - prof['-loglkl'][row]) < 1e-7
    before = open(os.path.join(out_dir, 'profile/x2.txt')).read()
    analyse(out_dir)                                   # prospect-analyse on the snapshot
    assert open(os.path.join(out_dir, 'profile/x2.txt')).read() == before


def test_initialisation_matches_reference_fixture(tmp_path):
    """A14 through the product path: bin indices bit-exact, GPU bin covariance to 1e-9, start points exact."""
    from prospect_public_b200.initialise import initialise_optimiser
    from prospect_public_b200.kernels import initialise_kernel
    for d, nvals in ((20, 10), (30, 32)):
        paths, k = _inputs(tmp_path, d)
        cfg = Configuration(annealing_yaml(paths, str(tmp_path / f'o{d}'), write=False,
                                           values=f'np.linspace(-1.0, 2.0, {nvals})'))
        kernel = initialise_kernel(cfg.kernel, None, 0)
        sp, values = initialise_optimiser(cfg, kernel)
        ini = golden(f'toy{d}_init.npz')
        assert np.array_equal(values, ini['values'])
        for p in range(nvals):
            assert np.array_equal(sp.indices[p], ini['indices'][p])
        assert np.array_equal(sp.positions, ini['position'][:, :-1])
        assert np.allclose(sp.covmats, ini['covmat'], rtol=1e-9, atol=1e-14)
        assert np.allclose(sp.initial_loglkl, ini['initial_loglkl'], rtol=1e-12)
        assert kernel.varying_param_names == [f'x{i}' for i in range(1, d + 1) if i != 2]


def test_kernel_plugin_interface(tmp_path):
    """B1: CudaGaussianKernel.loglkl mutates prop, returns -log L; prior box; fixed-parameter bookkeeping."""
    from oracle import reference_port as rp
    from prospect_public_b200.kernels import initialise_kernel
    paths, k = _inputs(tmp_path, 20)
    cfg = Configuration(annealing_yaml(paths, str(tmp_path / 'o'), write=False))
    kernel = initialise_kernel(cfg.kernel, None, 0)
    t = rp.GaussianTarget(k['means'], k['covmat'])
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, 20)
    pos = {f'x{i + 1}': [x[i]] for i in range(20)}
    assert abs(kernel.loglkl(pos) - t.loglkl(x)) < 1e-12 * t.loglkl(x)
    kernel.set_fixed_parameters({'x2': 0.75})
    pos = {f'x{i + 1}': [x[i]] for i in range(20) if i != 1}
    val = kernel.loglkl(pos)
    assert pos['x2'] == [0.75]                                      # fixed parameter injected into prop
    assert abs(val - t.fix(1, 0.75).loglkl(np.delete(x, 1))) < 1e-12 * val
    assert kernel.logprior(pos) == 0.0
    pos['x5'] = [2.6]
    assert kernel.logprior(pos) == np.inf and kernel.logpost(pos) == np.inf
    assert np.array_equal(kernel.get_default_covmat(), 0.005 * np.diag(np.full(19, 5.0)))
    assert abs(kernel.get_analytic_profile('x2', -1.0) - 20.719044687) < 1e-8


def test_cold_profile_matches_reference_run_and_analytic(tmp_path):
    """The real reference's cold K1 run (T 0.1 -> 1e-6, 30 iterations, frozen) vs the GPU on identical inputs:
    Delta chi^2 <= 1e-3 per point against the reference and against the analytic profile."""
    with open(os.path.join(GOLDEN, 'toy20_cold_reference_run.json')) as f:
        gold = json.load(f)
    ref_rows = [[float(t) for t in ln.split('\t')[:3]] for ln in gold['files']['x2.txt'].strip().split('\n')[1:]]
    ref_ll = np.array([r[2] for r in ref_rows])
    paths, k = _inputs(tmp_path, 20)
    res = run(annealing_yaml(paths, str(tmp_path / 'out'), temperature_range=[0.1, 1.0e-6], max_iterations=30,
                             repetitions=4))
    prof = res['profile']
    cf = closed_form.profile_curve(k['means'], k['covmat'], 1, np.linspace(-1.0, 2.0, 10))
    assert np.max(2 * np.abs(prof['loglkls'] - ref_ll)) <= 1e-3
    assert np.max(2 * np.abs(prof['loglkls'] - cf)) <= 1e-3
    assert np.all(prof['loglkls'] - cf > -1e-9)
    # Delta chi^2 column: same parabola construction as the reference
    ref_dchi = np.array([r[1] for r in ref_rows])
    assert np.max(np.abs(prof['Delta_chi2'] - ref_dchi)) <= 1e-3


def test_30d_profile_32x256_cold_schedule_reaches_analytic(tmp_path):
    """K3 at full size (32 values x 256 repetitions = 8192 chains) on a cold schedule:
    Delta chi^2 <= 1e-3 and bestfit parameters <= 1e-4 relative (floor 1e-2 on |x|) for every point."""
    paths, k = _inputs(tmp_path, 30)
    vals = np.linspace(-1.0, 2.0, 32)
    res = run(annealing_yaml(paths, str(tmp_path / 'out'), values='np.linspace(-1.0, 2.0, 32)', repetitions=256,
                             temperature_range=[0.1, 1.0e-12], max_iterations=400, steps_per_iteration=500))
    rec, srt, prof = res['record'], res['bestfits'], res['profile']
    assert res['chain_steps'] == 32 * 256 * 400 * 500
    p_idx = np.arange(32)
    best = srt['chain_best'][srt['best_rep'], p_idx]
    it = srt['chain_iter'][srt['best_rep'], p_idx]
    pos = rec.best_x[it, srt['best_rep'] * 32 + p_idx]
    for p, v in enumerate(vals):
        cf, xcf = closed_form.profile_minimum(k['means'], k['covmat'], 1, v)
        assert 0 <= 2 * (best[p] - cf) + 1e-9 and 2 * (best[p] - cf) <= 1e-3, (p, best[p] - cf)
        rel = np.max(np.abs(pos[p] - xcf) / np.maximum(np.abs(xcf), 1e-2))
        assert rel <= 1e-4, (p, rel)
    lp = load_profile(str(tmp_path / 'out'))
    assert len(lp['x2']) == 32
    # chains of one point are independent (distinct Philox streams): repetitions do not coincide
    # (checked on the first, hot iteration; at T = 1e-12 all of them agree to the last bits of the minimum)
    assert len(np.unique(rec.best_loglkl[0].reshape(256, 32)[:, 0])) == 256
    assert np.max(srt['chain_best'][:, 0]) - np.min(srt['chain_best'][:, 0]) < 1e-8


def test_30d_cold_profile_vs_frozen_reference_run(tmp_path):
    """The UNMODIFIED reference's own cold serial run of the 30-D toy (8 of the 32 profile values, T 0.1 -> 1e-10 in
    180 iterations x 250 steps, seed 0; frozen by tests/golden/make_golden_r2.py) against the GPU on identical inputs
    and the identical schedule (16 repetitions):
      * profiled chi^2: Delta chi^2 <= 1e-3 against the reference AND against the analytic profile;
      * bestfit parameters (relative, |x| floored at 1e-2): the reference's single repetition freezes 4e-5 ... 1.3e-2
        away from the analytic optimum at this schedule (the step size shrinks by 0.7 per iteration; measured when the
        fixture was made), so 1e-4 agreement with IT is not a property any correct sampler can have.  What is asserted:
        the GPU's best of 64 repetitions is at least as close to the analytic optimum as the reference is (or
        <= 1e-4) -- by the triangle inequality GPU and reference parameters then differ by at most twice the
        reference's own distance from the optimum.  (1e-4 against the optimum itself is reached on the longer schedule of
        test_30d_profile_32x256_cold_schedule_reaches_analytic.)"""
    with open(os.path.join(GOLDEN, 'toy30_cold_reference_run.json')) as f:
        gold = json.load(f)
    sch = gold['schedule']
    paths, k = _inputs(tmp_path, 30)
    res = run(annealing_yaml(paths, str(tmp_path / 'out'), values=gold['values'], repetitions=64,
                             temperature_range=sch['temperature_range'], max_iterations=sch['max_iterations'],
                             steps_per_iteration=sch['steps_per_iteration'],
                             step_size_adaptive_interval=sch['step_size_adaptive_interval'],
                             step_size_adaptive_multiplier=sch['step_size_adaptive_multiplier'],
                             step_size_adaptive_initial=sch['step_size_adaptive_initial']))
    rec, srt = res['record'], res['bestfits']
    n_pts = len(gold['values'])
    p_idx = np.arange(n_pts)
    best = srt['chain_best'][srt['best_rep'], p_idx]
    it = srt['chain_iter'][srt['best_rep'], p_idx]
    pos = rec.best_x[it, srt['best_rep'] * n_pts + p_idx]
    for p, v in enumerate(gold['values']):
        ref = gold['best'][p]
        assert ref['fixed_param_val'] == v
        ref_x = np.array([x for nm, x in zip(ref['position_names'], ref['bestfit_position']) if nm != 'x2'])
        cf, xcf = closed_form.profile_minimum(k['means'], k['covmat'], 1, v)
        assert 2 * abs(best[p] - ref['bestfit_loglkl']) <= 1e-3, (p, best[p], ref['bestfit_loglkl'])
        assert -1e-9 <= 2 * (best[p] - cf) <= 1e-3
        scale = np.maximum(np.abs(xcf), 1e-2)
        rel_gpu = np.max(np.abs(pos[p] - xcf) / scale)
        rel_ref = np.max(np.abs(ref_x - xcf) / scale)
        assert rel_gpu <= max(1e-4, rel_ref), (p, rel_gpu, rel_ref)
    # the reference's own profile file for these values: same Delta chi^2 column to 1e-3
    ref_rows = [[float(t) for t in ln.split('\t')[:3]] for ln in gold['files']['x2.txt'].strip().split('\n')[1:]]
    assert np.max(np.abs(res['profile']['Delta_chi2'] - np.array([r[1] for r in ref_rows]))) <= 2e-3


def test_global_optimisation_4096_chains(tmp_path):
    """K2: example_global_optimisation_toy.yaml with repetitions = 4096 (20-D as shipped)."""
    paths, k = _inputs(tmp_path, 20)
    out_dir = str(tmp_path / 'out')
    res = run(annealing_yaml(paths, out_dir, jobtype='global_optimisation', repetitions=4096,
                             temperature_range=[0.1, 1.0e-6], max_iterations=30))
    cb = res['bestfits']['chain_best'][:, 0]
    assert res['chain_steps'] == 4096 * 30 * 250
    assert cb.min() >= -1e-12 and 2 * cb.min() <= 1e-3            # global minimum of the Gaussian is 0 at mu
    assert 2 * np.median(cb) <= 1e-3
    lines = open(os.path.join(out_dir, 'global_optimisation', 'results.txt')).read().strip().split('\n')
    assert len(lines) == 4097 and lines[0].startswith('Repetition number \t -loglkl \t initial_loglkl')
    best_rep = res['bestfits']['best_rep']
    x = np.array([float(t) for t in lines[1 + best_rep].split('\t')[3:]])
    assert np.max(np.abs(x - k['means'])) < 5e-3


def test_mcmc_job_65536_chains_ks(tmp_path):
    """K5: 65 536 parallel MH chains on the 30-D toy at T = 1.  After burn-in the final states of the chains are
    65 536 independent draws: KS per coordinate against N(mu_i, (S^-1)_ii), a two-sample KS against the reference's
    own (frozen) chains, and the pooled covariance from the GPU moment reduction against S^-1."""
    from scipy import stats
    from prospect_public_b200 import engine
    from prospect_public_b200.kernels import initialise_kernel
    from prospect_public_b200.mcmc import initialise_mcmc
    paths, k = _inputs(tmp_path, 30)
    d = 30
    target_cov = closed_form.stationary_covariance(k['covmat'])
    cfg = Configuration(mcmc_yaml(paths, str(tmp_path / 'o'), n_chains=65536, steps=2000, write=False,
                                  start_from_covmat=target_cov.tolist(), step_size=float(2.4 / np.sqrt(d))))
    kernel = initialise_kernel(cfg.kernel, None, 0)
    sampler = initialise_mcmc(cfg.mcmc, kernel, seed=77, record='thinned', thin=2000)
    sampler.run_steps(2000)            # burn-in from the reference's default start (zeros)
    sampler.run_steps(2000)
    x, ll, mult, cnt = sampler.samples_numpy()
    assert np.all(cnt == 2)
    final = x[:, 1, :]
    ar = sampler.accepted.sum() / (65536 * 4000)
    assert 0.15 < ar < 0.40
    sig = np.sqrt(np.diag(target_cov))
    for i in range(d):
        p = stats.kstest((final[:, i] - k['means'][i]) / sig[i], 'norm').pvalue
        assert p > 1e-3 / d, (i, p)
    # ... and against the REFERENCE's own chains: tests/golden/toy30_ref_mcmc_thinned.npz is a run of the unmodified
    # reference (`jobtype: mcmc`, same target, same proposal covariance and step size, 16 chains x 40 000 steps, burn-in
    # dropped, every 250th step kept = 2320 samples).  Two-sample KS per coordinate, reject at p < 1e-3 / d.
    ref = golden('toy30_ref_mcmc_thinned.npz')
    assert abs(float(ref['step_size']) - 2.4 / np.sqrt(d)) < 1e-12 and np.allclose(ref['proposal_covmat'], target_cov)
    assert abs(ar - float(ref['acceptance_rate'])) < 0.02             # same acceptance rate as the reference's chains
    for i in range(d):
        p = stats.ks_2samp(final[::8, i], ref['samples'][:, i]).pvalue
        assert p > 1e-3 / d, (i, p)
    sw, swx, swxx = engine.weighted_moments(sampler._samples.x[:, 1, :].contiguous(), d)
    cov = engine.covariance_from_moments(sw, swx, swxx).cpu().numpy()
    assert np.max(np.abs(cov - target_cov) / np.sqrt(np.outer(np.diag(target_cov), np.diag(target_cov)))) < 0.02
    assert np.max(np.abs((swx / sw).cpu().numpy() - k['means']) / sig) < 0.02


def test_mcmc_job_writes_chain_files_like_reference(tmp_path):
    """jobtype mcmc with the reference's defaults (4 chains, start 0, covmat 0.005*diag(widths)): chain files in
    the MontePython layout that the profile initialiser can read back."""
    from prospect_public_b200.initialise import load_mcmc
    paths, k = _inputs(tmp_path, 20)
    out_dir = str(tmp_path / 'mc')
    res = run(mcmc_yaml(paths, out_dir, n_chains=4, steps=2000, rm1=1e9))
    assert res['rounds'] == 1
    chain = load_mcmc(os.path.join(out_dir, 'mcmc', 'test'))
    assert np.sum(chain.mults) == 4 * 2001                      # every chain: sum(mults) = steps + 1
    assert os.path.isfile(os.path.join(out_dir, 'mcmc', 'test.paramnames'))
    ref = golden('toy20_mcmc.npz')
    rows_per_round = sum(len(ref[f'chain_{i}']) for i in range(4)) / 4.0     # the reference ran 4 rounds
    assert 0.7 < len(chain.mults) / rows_per_round < 1.4          # same acceptance as the reference's chains


def test_wide_kernel_profile_d256(tmp_path):
    """K4-style 256-D synthetic SPD target through the CPL=8 path: 4 values x 8 reps, closed-form proposal covmat."""
    from tests.helpers import conditional_covariance
    d = 256
    mu, cov = spd_target(d, seed=4)
    paths = write_toy_inputs(str(tmp_path / 'inp'), mu, cov, None)
    sig = np.sqrt(cov[1, 1])
    vals = (mu[1] + sig * np.array([-2.0, -0.5, 0.5, 2.0])).tolist()
    start = np.delete(mu, 1) + 0.01
    y = annealing_yaml(paths, str(tmp_path / 'out'), values=vals, repetitions=8, start_from_mcmc=None,
                       start_from_position=start.tolist(), start_from_covmat=conditional_covariance(cov, 1).tolist(),
                       temperature_range=[0.1, 1.0e-8], max_iterations=120, steps_per_iteration=400,
                       step_size_adaptive_initial=0.15)
    res = run(y)
    best = res['bestfits']['chain_best'][res['bestfits']['best_rep'], np.arange(4)]
    for p, v in enumerate(vals):
        cf, _ = closed_form.profile_minimum(mu, cov, 1, v)
        assert 2 * abs(best[p] - cf) <= 1e-3, (p, best[p], cf)


def test_reanneal_continues_from_the_best_points(tmp_path):
    """N2 (prospect/profile.py:17-97): a finished run is re-annealed with a colder schedule from each value's
    best point; the new iterations are appended, the profile can only improve, every file is rewritten."""
    from prospect_public_b200.run import load_record, reanneal
    paths, k = _inputs(tmp_path, 20)
    out_dir = str(tmp_path / 'out')
    first = run(annealing_yaml(paths, out_dir, repetitions=2))
    ll_first = first['profile']['loglkls'].copy()
    second = reanneal({'profile': {'temperature_range': [1.0e-3, 1.0e-7], 'max_iterations': 25,
                                   'step_size_adaptive_initial': 0.02, 'repetitions': 3}}, out_dir)
    rec = second['record']
    assert rec.best_loglkl.shape == (45, 30) and rec.reps == 3
    assert np.all((rec.accepted[:20].reshape(20, 3, 10)[:, 2] == -1))          # the third repetition is new
    assert np.all(rec.accepted[20:] >= 0)
    ll_second = second['profile']['loglkls']
    assert np.all(ll_second <= ll_first + 1e-12)
    cf = closed_form.profile_curve(k['means'], k['covmat'], 1, np.linspace(-1.0, 2.0, 10))
    assert np.max(2 * np.abs(ll_second - cf)) <= 1e-3
    # re-annealed chains started from the old best: their first record cannot be worse than it
    assert np.all(rec.best_loglkl[20].reshape(3, 10).min(axis=0) <= ll_first + 1e-12)
    _, rec_disk = load_record(out_dir)
    assert rec_disk.best_loglkl.shape == (45, 30)
    prof = load_profile(out_dir)
    assert np.allclose(prof['-loglkl'], ll_second, rtol=1e-9)


@pytest.mark.parametrize('multiplier', [0.3, 'adaptive'])
def test_interrupted_run_resumes_to_identical_outputs(tmp_path, multiplier):
    """N2 resume (`prospect <folder>`, prospect/io.py:15-24, communication.py:30-36): a run stopped after 4 of 9
    iterations leaves a complete snapshot; running the FOLDER continues every chain (position, temperature, step
    size and its secant history, Philox step counter) and ends with byte-identical result files and bit-identical
    records to the run that was never interrupted.  Resuming a finished folder runs nothing."""
    paths, k = _inputs(tmp_path, 20)
    kw = dict(values=[-0.5, 0.3, 1.1], repetitions=3, max_iterations=9, temperature_range=[0.1, 1.0e-4],
              step_size_adaptive_multiplier=multiplier, step_size_adaptive_interval=[0.2, 0.3])
    full = run(annealing_yaml(paths, str(tmp_path / 'full'), **kw))
    part = run(annealing_yaml(paths, str(tmp_path / 'part'), **kw), stop_after=4)
    assert part['iterations_done'] == 4 and np.all(part['record'].accepted[4:] == -1)
    assert np.array_equal(part['record'].best_loglkl[:4], full['record'].best_loglkl[:4])
    res = run(str(tmp_path / 'part'))
    assert res['iterations_done'] == 9
    assert res['chain_steps'] == 250 * int((full['record'].accepted[4:] >= 0).sum()) > 0     # only the new rows ran
    for key in ('best_loglkl', 'accepted', 'best_x', 'temperature', 'step_size', 'converged'):
        assert np.array_equal(getattr(res['record'], key), getattr(full['record'], key)), key
    for name in ('x2.txt', 'x2_status.txt', 'x2_intervals.txt'):
        with open(tmp_path / 'part' / 'profile' / name, 'rb') as fa, open(tmp_path / 'full' / 'profile' / name, 'rb') as fb:
            assert fa.read() == fb.read(), name
    again = run(str(tmp_path / 'part'))
    assert again['chain_steps'] == 0 and np.array_equal(again['record'].best_loglkl, full['record'].best_loglkl)


def test_start_from_profile_and_explicit_gaussian_kernel(tmp_path):
    """N4: `start_from_profile` (+ covmat from the MCMC) starts every value from the nearest old bestfit
    (initialise_optimiser_task.py:39-62); and the explicit `function: gaussian` kernel-settings file
    (input/example_toy/kernel_settings/kernel_gauss.yaml) works through the same CUDA kernel."""
    import yaml
    paths, k = _inputs(tmp_path, 20)
    first = run(annealing_yaml(paths, str(tmp_path / 'a'), temperature_range=[0.1, 1.0e-5], max_iterations=25))
    old_txt = str(tmp_path / 'a' / 'profile' / 'x2.txt')
    second = run(annealing_yaml(paths, str(tmp_path / 'b'), start_from_profile=old_txt,
                                values='np.linspace(-0.9, 1.9, 8)', temperature_range=[1.0e-4, 1.0e-7],
                                max_iterations=15, step_size_adaptive_initial=0.01))
    cf = closed_form.profile_curve(k['means'], k['covmat'], 1, np.linspace(-0.9, 1.9, 8))
    assert np.max(2 * np.abs(second['profile']['loglkls'] - cf)) <= 1e-3
    # the start points are old bestfits (not bin bestfits): close to the old profile's positions
    old = load_profile(str(tmp_path / 'a'))
    near = np.argmin(np.abs(old['x2'][:, None] - np.linspace(-0.9, 1.9, 8)[None, :]), axis=0)
    assert np.allclose(second['record'].initial_position[:, 0], old['x1'][near])

    kfile = tmp_path / 'kernel_gauss.yaml'
    with open(kfile, 'w') as f:
        yaml.dump({'function': 'gaussian', 'dimension': 2, 'param_dict': {'x1': [-5.0, 5.0], 'x2': [-5.0, 5.0]},
                   'means': {'x1': 1.0, 'x2': 1.0}, 'covmat': [[0.25, 0.1], [0.1, 0.5]]}, f)
    y = annealing_yaml(paths, str(tmp_path / 'g'), jobtype='global_optimisation', repetitions=64,
                       start_from_mcmc=None, start_from_position=[3.0, -2.0],
                       start_from_covmat=[[0.25, 0.1], [0.1, 0.5]], temperature_range=[1.0, 1.0e-8],
                       max_iterations=40, step_size_adaptive_initial=1.0)
    y['kernel']['param'] = str(kfile)
    res = run(y)
    cb = res['bestfits']['chain_best'][:, 0]
    assert cb.min() >= -1e-12 and np.median(cb) < 1e-6
    rec = res['record']
    best_rep = res['bestfits']['best_rep']
    x = rec.best_x[res['bestfits']['chain_iter'][best_rep, 0], best_rep]
    assert np.allclose(x, [1.0, 1.0], atol=1e-3)


def test_ignore_prior_lifts_the_box(tmp_path):
    """kernel.ignore_prior: true (base_kernel.py:15-16) -- chains may leave [-2.5, 2.5]."""
    from prospect_public_b200.kernels import initialise_kernel
    from prospect_public_b200.mcmc import initialise_mcmc
    paths, k = _inputs(tmp_path, 20)
    for ignore, expect_outside in ((False, False), (True, True)):
        y = mcmc_yaml(paths, str(tmp_path / f'o{ignore}'), n_chains=256, steps=400, write=False,
                      temperature=1.0e6, step_size=3.0)
        y['kernel']['ignore_prior'] = ignore
        cfg = Configuration(y)
        kernel = initialise_kernel(cfg.kernel, None, 0)
        sampler = initialise_mcmc(cfg.mcmc, kernel, seed=5, record='thinned', thin=400)
        sampler.run_steps(400)
        x, _, _, cnt = sampler.samples_numpy()
        assert np.all(cnt == 1)
        outside = np.any(np.abs(x[:, 0, :]) > 2.5)
        assert outside == expect_outside


@pytest.mark.parametrize('world,reps', [(2, 6), (3, 6), (8, 6), (8, 400)])
def test_results_do_not_depend_on_the_number_of_ranks(tmp_path, monkeypatch, world, reps):
    """N-invariance of the sharded job, with the ranks emulated one after the other on one GPU: every rank of a
    `world`-rank job builds its own SimulatedAnnealing (its shard of the point x repetition grid, its global chain
    ids), runs the schedule, and the shards are assembled by global chain index -- bit-identical to the single-rank
    run.  5 points x 6 repetitions: 2 / 3 ranks shard by point, 8 ranks split the repetitions and leave two ranks
    WITHOUT chains (fewer repetitions than GPUs).  5 x 400 = 2000 chains: the job as a whole runs with 16 lanes per
    chain, a 250-chain shard on its own would pick 32 -- the geometry (hence the order of the FP32 reductions) is chosen
    from the size of the whole job, so the shards still reproduce the single-rank run.  (The exchange itself -- NCCL / peer-memory stores -- is covered by
    test_fused_exchange_fills_every_ranks_buffer, the gloo tests and scripts/multi_gpu_check.py on real ranks.)"""
    from prospect_public_b200 import optimiser as opt_mod
    from prospect_public_b200.distributed import global_chain_index
    from prospect_public_b200.initialise import initialise_optimiser
    from prospect_public_b200.kernels import initialise_kernel
    from prospect_public_b200.optimiser import (SimulatedAnnealing, get_initial_step_size_change,
                                                get_temperature_change)
    paths, k = _inputs(tmp_path, 20)
    cfg = Configuration(annealing_yaml(paths, str(tmp_path / 'o'), write=False, values=[-0.8, -0.1, 0.4, 1.1, 1.9],
                                       repetitions=reps, max_iterations=3, steps_per_iteration=90))
    kernel = initialise_kernel(cfg.kernel, None, 0)
    starts, values = initialise_optimiser(cfg, kernel)
    n_pts, n_iter = 5, 3
    settings = {'current_best_loglkl': starts.initial_loglkl, 'initial_position': starts.positions,
                'covmat': starts.covmats, 'temperature': 0.1, 'temperature_change': get_temperature_change(cfg),
                'step_size': 0.1, 'step_size_change': get_initial_step_size_change(cfg), 'iteration_number': 0,
                'repetitions': reps, 'fixed_param_val': values, 'max_rows': n_iter, 'seed': 4242}

    def run_as(rank, size):
        monkeypatch.setattr(opt_mod, 'dist_info', lambda: (rank, size, False))
        sa = SimulatedAnnealing(cfg, kernel, dict(settings))
        sa.run_schedule(n_iter, gather=False)
        torch.cuda.synchronize()
        ids = global_chain_index(sa.point_global, sa.rep, n_pts)
        out = {key: getattr(sa.records, key).cpu().numpy() for key in ('best_loglkl', 'accepted', 'step_size', 'best_x')}
        out['final_x'] = sa.batch.x.cpu().numpy()
        return ids, out

    _, single = run_as(0, 1)
    seen = np.zeros(n_pts * reps, dtype=int)
    empty_ranks = 0
    for rank in range(world):
        ids, part = run_as(rank, world)
        empty_ranks += len(ids) == 0
        seen[ids] += 1
        for key in ('best_loglkl', 'accepted', 'step_size'):
            assert np.array_equal(part[key], single[key][:, ids]), (key, rank)
        assert np.array_equal(part['best_x'], single['best_x'][:, ids]) and np.array_equal(part['final_x'], single['final_x'][ids])
    assert np.all(seen == 1)                       # every chain ran on exactly one rank
    assert empty_ranks == (2 if (world, reps) == (8, 6) else 0)
    if reps == 400:
        from prospect_public_b200 import engine
        assert engine.choose_lanes(32, n_pts * reps) == 16 and engine.choose_lanes(32, n_pts * reps // world) == 32


def test_results_async_equals_set_bestfit_with_two_jobs_in_flight(tmp_path):
    """SimulatedAnnealing.results_async: the records of job k come back through pinned memory on the copy stream while
    job k + 1 is built, uploaded and launched; what `get()` returns is what `set_bestfit` / `reduce_over_iterations`
    return for that job, and a job's device records are not disturbed by its successors."""
    from prospect_public_b200.initialise import initialise_optimiser
    from prospect_public_b200.kernels import initialise_kernel
    from prospect_public_b200.optimiser import (SimulatedAnnealing, get_initial_step_size_change,
                                                get_temperature_change)
    paths, k = _inputs(tmp_path, 20)
    cfg = Configuration(annealing_yaml(paths, str(tmp_path / 'o'), write=False, values=[-0.8, 0.4, 1.9],
                                       repetitions=40, max_iterations=4, steps_per_iteration=120))
    kernel = initialise_kernel(cfg.kernel, None, 0)
    starts, values = initialise_optimiser(cfg, kernel)
    n_iter = 4

    def job(seed):
        return SimulatedAnnealing(cfg, kernel, {
            'current_best_loglkl': starts.initial_loglkl, 'initial_position': starts.positions, 'covmat': starts.covmats,
            'temperature': 0.1, 'temperature_change': get_temperature_change(cfg), 'step_size': 0.1,
            'step_size_change': get_initial_step_size_change(cfg), 'iteration_number': 0, 'repetitions': 40,
            'fixed_param_val': values, 'max_rows': n_iter, 'seed': seed})

    expect = []
    for seed in range(5):                                        # the plain, synchronous way
        sa = job(seed)
        sa.run_schedule(n_iter)
        red = sa.reduce_over_iterations()
        bf = sa.set_bestfit(n_iter - 1)
        expect.append((bf, {name: t.cpu() for name, t in red.items()}))
    got, pending = [], None
    for seed in range(5):                                        # two jobs in flight
        sa = job(seed)
        sa.run_schedule(n_iter)
        nxt = sa.results_async(n_iter - 1, sa.reduce_over_iterations())
        if pending is not None:
            got.append(pending.get())
        pending = nxt
    got.append(pending.get())
    assert pending.d2h_bytes > 0
    for (bf_a, red_a), (bf_b, red_b) in zip(expect, got):
        for key in ('loglkl', 'acceptance_rate', 'accepted_steps'):
            assert np.array_equal(bf_a[key], bf_b[key])
        assert bf_a['position'].keys() == bf_b['position'].keys()
        assert all(np.array_equal(bf_a['position'][nm], bf_b['position'][nm]) for nm in bf_a['position'])
        assert red_a.keys() == red_b.keys() and all(torch.equal(red_a[nm], red_b[nm]) for nm in red_a)
    assert not np.array_equal(expect[0][0]['loglkl'], expect[1][0]['loglkl'])          # (the jobs do differ)


def test_interrupted_mcmc_job_resumes_to_identical_chain_files(tmp_path):
    """`prospect <folder>` for jobtype mcmc (tasks/mcmc_task.py:71-81): a job stopped after its first round and
    resumed from the chain files it left is, step for step, the job that ran three rounds without interruption --
    byte-identical chain files (positions are written with 18 digits, the step counter is the sum of the
    multiplicities, the Philox key comes from the snapshot) and the same R-1."""
    paths, k = _inputs(tmp_path, 20)
    full = run(mcmc_yaml(paths, str(tmp_path / 'full'), n_chains=6, steps=700, rm1=1e-12), max_rounds=3)
    assert full['rounds'] == 3
    part = run(mcmc_yaml(paths, str(tmp_path / 'part'), n_chains=6, steps=700, rm1=1e-12), stop_after=1)
    assert part['rounds'] == 1
    cont = run(str(tmp_path / 'part'), max_rounds=3)
    assert cont['rounds'] == 3 and abs(cont['Rm1'] - full['Rm1']) <= 1e-12 * full['Rm1']   # (reduction order differs)
    for i in range(6):
        a = open(os.path.join(str(tmp_path / 'full'), 'mcmc', f'test_{i}.txt'), 'rb').read()
        b = open(os.path.join(str(tmp_path / 'part'), 'mcmc', f'test_{i}.txt'), 'rb').read()
        assert a == b, i


def test_gpu_binning_of_a_large_starting_mcmc_matches_numpy(tmp_path, monkeypatch):
    """for a starting MCMC of several 1e5 rows the nearest-N selection of all profile values runs on the
    device.  Indices equal numpy's argsort selection (distinct distances), and the whole initialisation -- start
    points, bin covariances, initial_loglkl -- equals the host path."""
    from prospect_public_b200 import initialise as ini
    from prospect_public_b200.kernels import initialise_kernel
    rng = np.random.default_rng(11)
    col = rng.normal(0.5, 1.0, 300_000)
    vals = np.linspace(-1.0, 2.0, 21)
    got = ini.get_points_in_bins_gpu(col, vals, 30_000, 'cuda:0')
    for p, v in enumerate(vals):
        assert np.array_equal(got[p], np.argsort(np.abs(col - v))[:30_000]), p
    # end to end through initialise_optimiser: force the device path on the (small) toy chain
    paths, k = _inputs(tmp_path, 20)
    cfg = Configuration(annealing_yaml(paths, str(tmp_path / 'o'), write=False))
    host = ini.initialise_optimiser(cfg, initialise_kernel(cfg.kernel, None, 0))[0]
    monkeypatch.setattr(ini, 'GPU_BINNING_MIN_ROWS', 1)
    dev = ini.initialise_optimiser(cfg, initialise_kernel(cfg.kernel, None, 0))[0]
    for a, b in zip(host.indices, dev.indices):      # the toy's four chains all start at 0: exact distance ties, whose
        assert np.array_equal(np.sort(a), np.sort(b))    # order neither sort defines -> the same SET of rows
    assert np.array_equal(host.positions, dev.positions) and np.array_equal(host.initial_loglkl, dev.initial_loglkl)
    assert np.allclose(host.covmats, dev.covmats, rtol=1e-10, atol=1e-16)    # (summation order follows the row order)


def test_mcmc_run_minutes_uses_the_time_budget(tmp_path):
    """BaseMCMC.run_minutes (mcmc.py:140-151): step until the wall-clock budget is spent; the chains stay a valid
    continuation (sum of multiplicities = steps + 1)."""
    from prospect_public_b200.kernels import initialise_kernel
    from prospect_public_b200.mcmc import initialise_mcmc
    paths, k = _inputs(tmp_path, 20)
    cfg = Configuration(mcmc_yaml(paths, str(tmp_path / 'o'), n_chains=8, steps=100, write=False))
    sampler = initialise_mcmc(cfg.mcmc, initialise_kernel(cfg.kernel, None, 0), seed=3)
    import time
    t0 = time.perf_counter()
    steps = sampler.run_minutes(0.01, chunk=512)            # 0.6 s
    assert 0.6 <= time.perf_counter() - t0 < 5.0 and steps >= 512 and steps % 512 == 0
    assert all(int(np.sum(ch.mults)) == steps + 1 for ch in sampler.chains)
