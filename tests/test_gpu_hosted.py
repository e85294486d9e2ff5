"""GPU tests of the batched proposal / accept machinery for HOST-evaluated likelihoods (SURVEY.md section 8f, N4:
prospect_public_b200/hosted.py over pb200_hosted_propose / pb200_hosted_accept).

The checker is oracle/hosted_model.py: the pinned reference port's MH step with the two numpy draws replaced by the
GPU's own Philox variates (pb200_rng_variates), and the port's next_settings / converged called unchanged.  Integer
results (accept counts, status, iteration bookkeeping) and the schedule (temperatures, step sizes: same IEEE operations
in the same order) must be identical; positions and likelihood values agree to FP64 rounding (the device accumulates
F z with fused multiply-adds, numpy does not): 1e-12 absolute, written below.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import closed_form, hosted_model as hm, ref_harness as rh, reference_port as rp  # noqa: E402
from prospect_public_b200 import capi, engine  # noqa: E402
from prospect_public_b200.config import Configuration  # noqa: E402
from prospect_public_b200.hosted import HostedSimulatedAnnealing, HostKernelAdapter  # noqa: E402
from prospect_public_b200.optimiser import (get_initial_step_size_change, get_temperature_change,  # noqa: E402
                                            initialise_optimiser)
from prospect_public_b200.toy import annealing_yaml, write_toy_inputs  # noqa: E402
from tests.helpers import ToyHostKernel, golden  # noqa: E402

DEV = 'cuda:0'
SEED = 0xB200_0000_1234


def _config(tmp_path, **profile):
    k, m = golden('toy20_kernel.npz'), golden('toy20_mcmc.npz')
    paths = write_toy_inputs(str(tmp_path / 'inp'), k['means'], k['covmat'], [m[f'chain_{i}'] for i in range(4)])
    return Configuration(annealing_yaml(paths, str(tmp_path / 'o'), write=False, **profile)), paths, k


@pytest.mark.parametrize('mode', ['adaptive', 'exponential', 'tolerance', 'secant'])
def test_hosted_annealing_matches_scalar_replay(tmp_path, mode):
    d, reps, n_iter, steps = 7, 3, 5, 40
    values = [-0.3, 0.2, 0.9]
    prof = dict(values=values, repetitions=reps, max_iterations=n_iter, steps_per_iteration=steps,
                temperature_range=[0.5, 0.05], step_size_adaptive_interval=[0.3, 0.5])
    step0 = 0.9
    if mode == 'exponential':
        prof.update(step_size_schedule='exponential', step_size_range=[0.9, 0.3])
    elif mode == 'tolerance':
        prof.update(chi2_tolerance=1.5)
    elif mode == 'secant':
        prof.update(step_size_adaptive_multiplier='adaptive', step_size_adaptive_interval=[0.02, 0.03])
    cfg, _, _ = _config(tmp_path, **prof)
    host = ToyHostKernel(d, box=(-0.1, 1.1))
    kernel = HostKernelAdapter(host, profile_parameter='x2')
    assert kernel.host_evaluated and kernel.varying_param_names == [f'x{i}' for i in range(1, d + 1) if i != 2]
    n, n_pts = d - 1, len(values)
    rng = np.random.default_rng(5)
    varying = [i for i in range(d) if i != 1]
    starts = np.clip(host.mu[varying] + 0.1 * rng.standard_normal((n_pts, n)), 0.0, 1.0)
    covs = np.empty((n_pts, n, n))
    for p in range(n_pts):
        a = rng.standard_normal((n, n))
        covs[p] = 0.04 * (a @ a.T / n + 0.5 * np.eye(n))
    init_ll = kernel.loglkl_batch(starts, values)
    assert host.calls == n_pts
    settings = {'current_best_loglkl': init_ll, 'initial_position': starts, 'covmat': covs, 'temperature': 0.5,
                'temperature_change': get_temperature_change(cfg), 'step_size': step0,
                'step_size_change': get_initial_step_size_change(cfg), 'iteration_number': 0, 'repetitions': reps,
                'fixed_param_val': values, 'max_rows': n_iter, 'seed': SEED}
    sa = initialise_optimiser(cfg, kernel, settings)
    assert isinstance(sa, HostedSimulatedAnnealing) and sa.n_chains == n_pts * reps
    launches0 = engine.LAUNCHES['count']
    sa.run_schedule(n_iter)
    torch.cuda.synchronize()
    launches = engine.LAUNCHES['count'] - launches0
    rec = sa.global_records()
    status = sa.global_status()
    final_x = sa.batch.positions()

    sched = rp.Schedule(temperature_range=(0.5, 0.05), max_iterations=n_iter,
                        step_size_schedule='exponential' if mode == 'exponential' else 'adaptive',
                        step_size_range=(0.9, 0.3) if mode == 'exponential' else None,
                        step_size_adaptive_interval=(0.02, 0.03) if mode == 'secant' else (0.3, 0.5),
                        step_size_adaptive_multiplier='adaptive' if mode == 'secant' else 0.3,
                        steps_per_iteration=steps, chi2_tolerance=1.5 if mode == 'tolerance' else None)
    lo, hi = np.full(n, -0.1), np.full(n, 1.1)
    evaluations, outside_total, stopped = n_pts + n_pts * reps, 0, 0     # start points: initialiser + optimiser
    for rep in range(reps):
        for p in range(n_pts):
            c = rep * n_pts + p                                          # initialise_profile_task.py:16-26
            z, e = engine.rng_variates(DEV, SEED, c, 0, n_iter * steps, n)
            z, e = z.cpu().numpy(), e.cpu().numpy()

            def loglkl(x, p=p):
                full = np.insert(x, 1, values[p])
                return host.value(full)

            try:
                rows, final = hm.anneal_chain(loglkl, starts[p], covs[p], lo, hi, sched, 0.5, step0, init_ll[p], z,
                                              e, n_iter)
                failed = False
            except UnboundLocalError:                 # the reference's secant edge path: the task fails
                failed = True
                rows = None
            if failed:
                assert status[c] == capi.FAILED
                continue
            for it, row in enumerate(rows):
                bf = row['bestfit']
                assert rec['accepted'][it, c] == row['accepted'], (c, it)
                assert rec['temperature'][it, c] == row['temperature'] and rec['step_size'][it, c] == row['step_size']
                assert abs(rec['best_loglkl'][it, c] - bf['loglkl']) <= 1e-12 * max(1.0, abs(bf['loglkl'])), (c, it)
                assert np.max(np.abs(rec['best_x'][it, c] - bf['position'])) <= 1e-12, (c, it)
                evaluations += steps - row['outside']
                outside_total += row['outside']
            for it in range(len(rows), n_iter):       # iterations a converged chain never ran
                assert rec['accepted'][it, c] == -1 and rec['best_loglkl'][it, c] == np.inf
            assert status[c] == (capi.CONVERGED if final['status'] == 'converged' else capi.ACTIVE)
            stopped += final['status'] == 'converged'
            assert np.max(np.abs(final_x[c] - final['x'])) <= 1e-12
    assert outside_total > 0                          # the box was exercised
    if mode == 'tolerance':
        assert 0 < stopped
    if mode != 'secant':
        # the likelihood of an out-of-box proposal and of a stopped chain is never asked for (mcmc.py:166-167)
        assert kernel.evaluations == evaluations and host.calls == evaluations
    # two launches per step + one schedule update per iteration, nothing else -- and nothing at all once every chain
    # has stopped
    assert launches % (2 * steps + 1) == 0 and launches <= n_iter * (2 * steps + 1)
    if mode in ('adaptive', 'exponential'):
        assert launches == n_iter * (2 * steps + 1)
    bf = sa.set_bestfit()
    assert set(bf['position']) == set(kernel.varying_param_names)


def test_hosted_vectorised_likelihood_and_worker_threads_agree(tmp_path):
    """The three ways the host can serve a batch -- one kernel object, several objects on worker threads, one
    vectorised callable -- give the same records."""
    d, reps, n_iter, steps = 5, 4, 2, 25
    values = [0.1, 0.6]
    cfg, _, _ = _config(tmp_path, values=values, repetitions=reps, max_iterations=n_iter, steps_per_iteration=steps)
    starts = np.full((2, d - 1), 0.4)
    covs = np.tile(0.02 * np.eye(d - 1), (2, 1, 1))

    def run_with(kernel):
        settings = {'current_best_loglkl': np.full(2, np.inf), 'initial_position': starts, 'covmat': covs,
                    'temperature': 0.3, 'temperature_change': 0.5, 'step_size': 0.7,
                    'step_size_change': get_initial_step_size_change(cfg), 'iteration_number': 0,
                    'repetitions': reps, 'fixed_param_val': values, 'max_rows': n_iter, 'seed': 99}
        sa = initialise_optimiser(cfg, kernel, settings)
        sa.run_schedule(n_iter)
        torch.cuda.synchronize()
        return sa.global_records()

    one = run_with(HostKernelAdapter(ToyHostKernel(d), profile_parameter='x2'))
    pool = run_with(HostKernelAdapter([ToyHostKernel(d) for _ in range(3)], profile_parameter='x2'))
    proto = ToyHostKernel(d)

    def batch(x, fixed):
        full = np.insert(x, 1, fixed, axis=1)
        r = full - proto.mu
        return 0.5 * np.einsum('ij,jk,ik->i', r, proto.prec, r) + 0.05 * np.sum(np.sin(3.0 * full) ** 2, axis=1)

    vec = run_with(HostKernelAdapter(ToyHostKernel(d), profile_parameter='x2', batch_loglkl=batch))
    for key in ('accepted', 'step_size', 'temperature'):
        assert np.array_equal(one[key], pool[key]) and np.array_equal(one[key], vec[key])
    assert np.array_equal(one['best_loglkl'], pool['best_loglkl']) and np.array_equal(one['best_x'], pool['best_x'])
    assert np.allclose(one['best_loglkl'], vec['best_loglkl'], rtol=1e-12)


def test_hosted_global_optimisation_without_box(tmp_path):
    """No profiled parameter, ignore_prior: every proposal is evaluated; a likelihood that fails (NaN / inf,
    base_kernel.py:78-81) never accepts."""
    d, n_iter, steps = 4, 3, 30
    cfg, _, _ = _config(tmp_path, repetitions=6, max_iterations=n_iter, steps_per_iteration=steps)
    host = ToyHostKernel(d, ignore_prior=True)
    calls = {'n': 0}

    def flaky(x, fixed):
        calls['n'] += len(x)
        r = x - host.mu
        out = 0.5 * np.einsum('ij,jk,ik->i', r, host.prec, r)
        out[x[:, 0] > host.mu[0] + 0.05] = np.nan               # "computation exception" region
        return out

    kernel = HostKernelAdapter(host, ignore_prior=True, batch_loglkl=flaky)
    settings = {'current_best_loglkl': np.inf, 'initial_position': host.mu[None, :] - 0.2,
                'covmat': 0.01 * np.eye(d), 'temperature': 0.2, 'temperature_change': 0.6, 'step_size': 1.0,
                'step_size_change': get_initial_step_size_change(cfg), 'iteration_number': 0, 'repetitions': 6,
                'max_rows': n_iter, 'seed': 7}
    sa = initialise_optimiser(cfg, kernel, settings)
    sa.run_schedule(n_iter)
    torch.cuda.synchronize()
    assert calls['n'] == 6 + 6 * n_iter * steps                 # no box: nothing is skipped
    x = sa.batch.positions()
    assert np.all(x[:, 0] <= host.mu[0] + 0.05)                 # the failing region was never entered
    rec = sa.global_records()
    assert np.all(np.isfinite(rec['best_loglkl'])) and np.all(rec['accepted'] >= 0)


@pytest.mark.skipif(not rh.available(), reason='oracle/_ref is absent (python oracle/make_ref.py builds it where '
                                               '/root/reference exists)')
def test_reference_analytical_kernel_behind_the_hosted_optimiser(tmp_path):
    """The UNMODIFIED reference AnalyticalKernel (CPU, prospect/kernels/analytical_kernel.py) as the host likelihood
    of a profile job on the shipped schedule: the profile reaches the closed form as the reference itself and the
    fused GPU path do (T_final = 1.26e-3: 5e-3 .. 1.3e-2 above the minimum)."""
    rh.activate()
    from prospect.kernels.analytical_kernel import AnalyticalKernel
    ini = golden('toy20_init.npz')                 # the reference's own start points / bin covariances per value
    sel = [1, 5, 8]
    values = [float(v) for v in ini['values'][sel]]
    cfg, paths, k = _config(tmp_path, values=values)
    out = str(tmp_path / 'ref_kernel')
    os.makedirs(os.path.join(out, 'analytical'))
    with rh.quiet():
        ref_kernel = AnalyticalKernel(cfg.kernel, 0, output_folder=out)
    kernel = HostKernelAdapter(ref_kernel, profile_parameter='x2')
    starts = ini['position'][sel][:, :-1]
    init = kernel.loglkl_batch(starts, values)
    assert np.allclose(init, ini['initial_loglkl'][sel], rtol=1e-12)       # the reference's own initial_loglkl column
    settings = {'current_best_loglkl': init, 'initial_position': starts, 'covmat': ini['covmat'][sel],
                'temperature': 0.1, 'temperature_change': get_temperature_change(cfg), 'step_size': 0.1,
                'step_size_change': get_initial_step_size_change(cfg), 'iteration_number': 0, 'repetitions': 1,
                'fixed_param_val': values, 'max_rows': 20, 'seed': 31}
    sa = initialise_optimiser(cfg, kernel, settings)
    sa.run_schedule(20)
    torch.cuda.synchronize()
    best = sa.global_records()['best_loglkl'].min(axis=0)
    cf = closed_form.profile_curve(k['means'], k['covmat'], 1, np.asarray(values))
    excess = best - cf
    assert np.all(excess > -1e-9) and np.all(excess < 5e-2), excess
    assert ref_kernel.param['fixed']['x2']['fixed_value'] in values          # the adapter re-values the fixed parameter


@pytest.mark.skipif(not rh.available(), reason='oracle/_ref is absent (python oracle/make_ref.py builds it where '
                                               '/root/reference exists)')
def test_whole_profile_job_with_a_host_kernel_through_run(tmp_path):
    """yaml -> run(kernel=...) -> PROSPECT's output tree with the UNMODIFIED reference AnalyticalKernel as a host
    likelihood: start points / bin covariances from the MCMC (GPU moments), initial_loglkl from one host batch, hosted
    annealing, the same analysis and writers; and an interrupted + resumed job (`prospect <folder>`) ends with the files
    and records of the uninterrupted one."""
    from prospect.kernels.analytical_kernel import AnalyticalKernel
    from prospect_public_b200.io import load_profile
    from prospect_public_b200.run import run
    rh.activate()
    k, m = golden('toy20_kernel.npz'), golden('toy20_mcmc.npz')
    paths = write_toy_inputs(str(tmp_path / 'inp'), k['means'], k['covmat'], [m[f'chain_{i}'] for i in range(4)])
    built = []

    def factory(config_kernel, output_folder, rank, device):
        with rh.quiet():
            ref = AnalyticalKernel(config_kernel, rank, output_folder=output_folder)
        built.append(ref)
        return HostKernelAdapter(ref, device=device)

    kw = dict(values=[-0.4, 0.5, 1.3], repetitions=2, max_iterations=6, steps_per_iteration=80)
    full = run(annealing_yaml(paths, str(tmp_path / 'full'), **kw), kernel=factory)
    assert isinstance(full['optimiser'], HostedSimulatedAnnealing) and full['chain_steps'] == 3 * 2 * 6 * 80
    for fn in ('log.yaml', 'config.pkl', 'state.pkl', 'profile/x2.txt', 'profile/x2_status.txt',
               'profile/x2_intervals.txt', 'analytical/random_gauss.pkl'):
        assert os.path.isfile(os.path.join(str(tmp_path / 'full'), fn)), fn
    prof = load_profile(str(tmp_path / 'full'))
    assert np.allclose(prof['x2'], kw['values'], atol=1e-12)
    cf = closed_form.profile_curve(k['means'], k['covmat'], 1, np.asarray(kw['values']))
    assert np.all(prof['-loglkl'] - cf > -1e-9) and np.all(prof['-loglkl'] < prof['initial_loglkl'])
    # the reported chi^2 IS the host kernel's value at the reported position
    names = [f'x{i}' for i in range(1, 21) if i != 2]
    ref = built[0]
    for row in range(3):
        ref.param['fixed']['x2']['fixed_value'] = float(prof['x2'][row])
        assert ref.loglkl({nm: [float(prof[nm][row])] for nm in names}) == pytest.approx(prof['-loglkl'][row], rel=1e-9)
    part = run(annealing_yaml(paths, str(tmp_path / 'part'), **kw), stop_after=2, kernel=factory)
    assert part['iterations_done'] == 2
    res = run(str(tmp_path / 'part'), kernel=factory)
    assert res['iterations_done'] == 6
    for key in ('best_loglkl', 'accepted', 'best_x', 'temperature', 'step_size'):
        assert np.array_equal(getattr(res['record'], key), getattr(full['record'], key)), key
    for name in ('x2.txt', 'x2_status.txt', 'x2_intervals.txt'):
        with open(tmp_path / 'part' / 'profile' / name, 'rb') as fa, open(tmp_path / 'full' / 'profile' / name, 'rb') as fb:
            assert fa.read() == fb.read(), name


def test_hosted_mh_sampler_matches_scalar_replay():
    """jobtype-mcmc sampler for a host kernel (initialise_mcmc -> HostedMetropolisHastings): every chain is, row for row
    (multiplicity, -log L, position), the scalar replay of MetropolisHastings.step with the GPU's variates; a sampler
    continued from the chains of another picks up their Philox streams and open last rows."""
    from types import SimpleNamespace
    from prospect_public_b200.mcmc import initialise_mcmc
    d, b = 5, 6
    host = ToyHostKernel(d, box=(-0.2, 1.2))
    kernel = HostKernelAdapter(host)
    rng = np.random.default_rng(3)
    a = rng.standard_normal((d, d))
    cov = 0.05 * (a @ a.T / d + 0.5 * np.eye(d))
    cfg = SimpleNamespace(algorithm='MetropolisHastings', N_chains=b, temperature=0.7, step_size=0.9,
                          start_from_covmat=cov, start_from_position=None, steps_per_iteration=37)
    mh = initialise_mcmc(cfg, kernel, seed=SEED)
    mh.run_steps(37)
    mh.run_steps(20)
    second = initialise_mcmc(cfg, kernel, seed=SEED, chains=mh.chains, steps_done=57)
    second.run_steps(30)
    chains = second.chains
    factor = np.linalg.cholesky(cov)
    lo, hi = np.full(d, -0.2), np.full(d, 1.2)
    outside = 0
    for c in range(b):
        z, e = engine.rng_variates(DEV, SEED, c, 0, 87, d)
        z, e = z.cpu().numpy(), e.cpu().numpy()
        x = np.full(d, 0.3)
        ll = host.value(x)
        rows = [[1, ll, x]]
        for t in range(87):
            acc, prop, llp, out = hm.step(host.value, x, ll, factor, lo, hi, 0.7, 0.9, z[t], e[t])
            outside += out
            if acc:
                x, ll = prop, llp
                rows.append([1, llp, prop])
            else:
                rows[-1][0] += 1
        ch = chains[c]
        assert list(ch.mults) == [r[0] for r in rows], c
        assert np.allclose(ch.loglkls, [r[1] for r in rows], rtol=1e-12, atol=0)
        got = np.stack([ch.positions[nm] for nm in kernel.varying_param_names], axis=1)
        assert np.max(np.abs(got - np.array([r[2] for r in rows]))) <= 1e-12
        assert sum(ch.mults) == 88                                  # start point + one count per step
    assert outside > 0
    assert int(mh.accepted.sum() + second.accepted.sum()) == sum(len(ch.mults) - 1 for ch in chains)


@pytest.mark.skipif(not rh.available(), reason='oracle/_ref is absent (python oracle/make_ref.py builds it where '
                                               '/root/reference exists)')
def test_mcmc_job_with_a_host_kernel_through_run_and_resume(tmp_path):
    """jobtype mcmc through run(kernel=...) with the reference AnalyticalKernel on the host: chain files in the reference's
    layout, R-1 from the device reductions, and a job stopped after its first round and resumed from its chain files
    (tasks/mcmc_task.py:71-81) writes byte-identical chain files to the job that was never interrupted."""
    from prospect.kernels.analytical_kernel import AnalyticalKernel
    from prospect_public_b200.hosted import HostedMetropolisHastings
    from prospect_public_b200.initialise import load_mcmc
    from prospect_public_b200.run import run
    from prospect_public_b200.toy import mcmc_yaml
    rh.activate()
    k = golden('toy20_kernel.npz')
    paths = write_toy_inputs(str(tmp_path / 'inp'), k['means'], k['covmat'], None)

    def factory(config_kernel, output_folder, rank, device):
        with rh.quiet():
            return HostKernelAdapter(AnalyticalKernel(config_kernel, rank, output_folder=output_folder), device=device)

    kw = dict(n_chains=5, steps=120, rm1=1e-12)
    full = run(mcmc_yaml(paths, str(tmp_path / 'full'), **kw), max_rounds=3, kernel=factory)
    assert full['rounds'] == 3 and isinstance(full['sampler'], HostedMetropolisHastings)
    assert np.isfinite(full['Rm1']) and full['chain_steps'] == 3 * 120 * 5
    chains = load_mcmc(os.path.join(str(tmp_path / 'full'), 'mcmc', 'test'), collapse=False)
    assert len(chains) == 5 and all(int(round(np.sum(ch.mults))) == 361 for ch in chains)
    part = run(mcmc_yaml(paths, str(tmp_path / 'part'), **kw), stop_after=1, kernel=factory)
    assert part['rounds'] == 1
    cont = run(str(tmp_path / 'part'), max_rounds=3, kernel=factory)
    assert cont['rounds'] == 3 and abs(cont['Rm1'] - full['Rm1']) <= 1e-12 * full['Rm1']
    for i in range(5):
        a = open(os.path.join(str(tmp_path / 'full'), 'mcmc', f'test_{i}.txt'), 'rb').read()
        b = open(os.path.join(str(tmp_path / 'part'), 'mcmc', f'test_{i}.txt'), 'rb').read()
        assert a == b, i
