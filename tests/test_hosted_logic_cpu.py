"""CPU run of the HOST logic of prospect_public_b200/hosted.py (record layout, per-iteration bookkeeping, early stops,
evaluation masks, chain recording and continuation) -- the logic tests/test_gpu_hosted.py checks on a B200.

There is no CPU implementation of the product's kernels, so this module supplies TEST DOUBLES for the five device
entry points the hosted path calls (pb200_hosted_propose, pb200_hosted_accept, pb200_schedule_update,
pb200_rng_variates and the device allocations), written in numpy on top of oracle/device_model (Philox variates,
schedule_next).  They exist only inside the `fake_device` fixture below; nothing under prospect_public_b200/ can reach
them, and the GPU tests remain the parity tests proper.  What is verified here is everything between those calls.
"""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import device_model as dm  # noqa: E402
from prospect_public_b200 import capi, engine, hosted, problem  # noqa: E402


@pytest.fixture
def fake_device(monkeypatch):
    def on_cpu(fn):
        def wrapped(*args, **kwargs):
            kwargs.pop('pin_memory', None)
            if 'device' in kwargs and str(kwargs['device']).startswith('cuda'):
                kwargs['device'] = 'cpu'
            return fn(*args, **kwargs)
        return wrapped

    for name in ('empty', 'zeros', 'full', 'as_tensor'):
        monkeypatch.setattr(torch, name, on_cpu(getattr(torch, name)))

    class Stream:
        cuda_stream = 0

        def synchronize(self):
            pass

    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: Stream())
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, 'current_device', lambda: 0)

    def hosted_problem_init(self, n, covmats, lo, hi, device):
        self.device, self.n, self.npad = 'cuda:0', int(n), problem.pad_of(n)
        covmats = np.asarray(covmats, dtype=np.float64).reshape(-1, n, n)
        self.n_points = covmats.shape[0]
        self.factors = [problem.proposal_factor(c) for c in covmats]
        self.lo, self.hi, self.c = np.asarray(lo), np.asarray(hi), None

    def upload_packed(arrays, device):
        out = {k: torch.as_tensor(np.ascontiguousarray(v).copy()) for k, v in arrays.items()}
        out['_storage'] = None
        return out

    monkeypatch.setattr(hosted.HostedProblem, '__init__', hosted_problem_init)
    monkeypatch.setattr(problem, 'upload_packed', upload_packed)
    monkeypatch.setattr(capi, 'ptr', lambda t: None)
    monkeypatch.setattr(capi, 'Chains', lambda *a: None)
    monkeypatch.setattr(capi, 'Records', lambda *a: None)
    monkeypatch.setattr(engine, 'choose_lanes', lambda npad, b: 32)

    def propose(hp, ch, seed, out_prop, out_outside, stream=None):
        engine._count()
        for c in range(ch.n_chains):
            live = int(ch.status[c]) == capi.ACTIVE
            z, _ = dm.variates(seed, int(ch.chain_id[c]) & 0xFFFFFFFF, int(ch.steps_done[c]), 1, hp.n)
            zz = z[0].astype(np.float64) if live else np.zeros(hp.n)
            prop = ch.x[c, :hp.n].numpy() + float(ch.step_size[c]) * (hp.factors[int(ch.point[c])] @ zz)
            out_prop[c, :hp.n] = torch.as_tensor(prop)
            out_outside[c] = int(live and bool(np.any(prop < hp.lo) or np.any(prop > hp.hi)))

    def accept(hp, ch, seed, prop, prop_ll, outside, flag=None, nacc=None, best_ll=None, best_x=None, stream=None):
        engine._count()
        for c in range(ch.n_chains):
            if int(ch.status[c]) != capi.ACTIVE:
                if flag is not None:
                    flag[c] = 0
                continue
            t = int(ch.steps_done[c])
            _, e = dm.variates(seed, int(ch.chain_id[c]) & 0xFFFFFFFF, t, 1, hp.n)
            ll, llp = float(ch.loglkl[c]), float(prop_ll[c])
            acc = int(outside[c]) == 0 and (llp - ll < float(ch.temperature[c]) * float(e[0]))
            if flag is not None:
                flag[c] = int(acc)
            if acc:
                ch.x[c] = prop[c]
                ch.loglkl[c] = llp
                if nacc is not None:
                    nacc[c] += 1
                if best_ll is not None and llp < float(best_ll[c]):
                    best_ll[c] = llp
                    best_x[c] = prop[c]
            ch.steps_done[c] = t + 1

    def schedule_update(ch, sched, accepted, best, stream=None):
        engine._count()
        ms = dm.DmSchedule(sched.steps_per_iteration, sched.step_mode, sched.temperature_change, sched.step_size_change,
                           sched.adaptive_lo, sched.adaptive_hi, sched.adaptive_multiplier, sched.chi2_tolerance)
        for c in range(ch.n_chains):
            if int(ch.status[c]) != capi.ACTIVE:
                continue
            mc = dm.DmChain(None, 0.0, float(ch.temperature[c]), float(ch.step_size[c]), float(ch.current_best[c]),
                            float(ch.secant_prev_mult[c]), float(ch.secant_cur_mult[c]), float(ch.secant_prev_ar[c]), 0,
                            int(ch.iteration[c]), 0, int(ch.secant_stage[c]), 0, 0)
            dm.schedule_next(ms, mc, int(accepted[c]), float(best[c]))
            ch.temperature[c], ch.step_size[c], ch.current_best[c] = mc.temperature, mc.step_size, mc.current_best
            ch.secant_prev_mult[c], ch.secant_cur_mult[c] = mc.prev_mult, mc.cur_mult
            ch.secant_prev_ar[c], ch.secant_stage[c] = mc.prev_ar, mc.stage
            ch.status[c], ch.iteration[c] = mc.status, mc.iteration

    def rng_variates(device, seed, chain_id, step0, n_steps, n, stream=None):
        z, e = dm.variates(seed, chain_id, step0, n_steps, n)
        return torch.as_tensor(z), torch.as_tensor(e)

    monkeypatch.setattr(engine, 'hosted_propose', propose)
    monkeypatch.setattr(engine, 'hosted_accept', accept)
    monkeypatch.setattr(engine, 'schedule_update', schedule_update)
    monkeypatch.setattr(engine, 'rng_variates', rng_variates)


def _gpu_module():
    import tests.test_gpu_hosted as g
    return g


@pytest.mark.parametrize('mode', ['adaptive', 'exponential', 'tolerance', 'secant'])
def test_annealing_bookkeeping(fake_device, tmp_path, mode):
    _gpu_module().test_hosted_annealing_matches_scalar_replay(tmp_path, mode)


def test_three_ways_to_serve_a_batch(fake_device, tmp_path):
    _gpu_module().test_hosted_vectorised_likelihood_and_worker_threads_agree(tmp_path)


def test_no_box_and_failing_likelihoods(fake_device, tmp_path):
    _gpu_module().test_hosted_global_optimisation_without_box(tmp_path)


def test_mh_sampler_recording_and_continuation(fake_device):
    _gpu_module().test_hosted_mh_sampler_matches_scalar_replay()


def test_reference_kernel_behind_the_optimiser(fake_device, tmp_path):
    g = _gpu_module()
    if not g.rh.available():
        pytest.skip('the reference is not available here')
    g.test_reference_analytical_kernel_behind_the_hosted_optimiser(tmp_path)


@pytest.mark.parametrize('mode', ['adaptive', 'exponential', 'secant'])
def test_hosted_optimiser_reproduces_the_unmodified_reference_chain_by_chain(fake_device, tmp_path, monkeypatch, mode):
    """The strongest statement the CPU suite can make about the hosted path: with the reference's own AnalyticalKernel
    as the host likelihood, every chain of HostedSimulatedAnnealing (its host logic + the device semantics the test
    doubles restate) is the UNMODIFIED reference's SimulatedAnnealing -- optimise / set_bestfit /
    get_next_iteration_settings (prospect/optimiser.py:40-162) over MetropolisHastings.step (mcmc.py:163-190) -- fed
    the draws the chain's variates stand for: accepted steps, bestfit value and position, next temperature and step
    size, last position, all EXACTLY equal, iteration by iteration."""
    from types import SimpleNamespace
    g = _gpu_module()
    if not g.rh.available():
        pytest.skip('the reference is not available here')
    g.rh.activate()
    from prospect.kernels.analytical_kernel import AnalyticalKernel
    from prospect.optimiser import SimulatedAnnealing as RefSimulatedAnnealing
    import os
    ini = g.golden('toy20_init.npz')
    sel, reps, n_iter, steps, seed = [2, 7], 2, 5, 40, 2024
    values = [float(v) for v in ini['values'][sel]]
    prof = dict(values=values, repetitions=reps, max_iterations=n_iter, steps_per_iteration=steps)
    if mode == 'exponential':
        prof.update(step_size_schedule='exponential', step_size_range=[0.5, 0.05])
    elif mode == 'secant':
        prof.update(step_size_adaptive_multiplier='adaptive', step_size_adaptive_interval=[0.02, 0.03])
    cfg, paths, k = g._config(tmp_path, **prof)
    step0 = 0.5 if mode == 'exponential' else 0.1

    def new_ref_kernel(tag):
        out = str(tmp_path / tag)
        os.makedirs(os.path.join(out, 'analytical'))
        with g.rh.quiet():
            return AnalyticalKernel(cfg.kernel, 0, output_folder=out)

    kernel = g.HostKernelAdapter(new_ref_kernel('ours'), profile_parameter='x2')
    names = kernel.varying_param_names
    starts, covs = ini['position'][sel][:, :-1], ini['covmat'][sel]
    init_ll = kernel.loglkl_batch(starts, values)
    settings = {'current_best_loglkl': init_ll, 'initial_position': starts, 'covmat': covs, 'temperature': 0.1,
                'temperature_change': g.get_temperature_change(cfg), 'step_size': step0,
                'step_size_change': g.get_initial_step_size_change(cfg), 'iteration_number': 0, 'repetitions': reps,
                'fixed_param_val': values, 'max_rows': n_iter, 'seed': seed}
    sa = g.initialise_optimiser(cfg, kernel, settings)
    sa.run_schedule(n_iter)
    rec, status, final_x = sa.global_records(), sa.global_status(), sa.batch.positions()
    final_t, final_s = sa.batch.temperature.numpy(), sa.batch.step_size.numpy()

    ref_cfg = SimpleNamespace(run=SimpleNamespace(jobtype='profile'),
                              profile=SimpleNamespace(parameter='x2', steps_per_iteration=steps, minutes_per_iteration=None,
                                                      temperature_schedule='exponential',
                                                      step_size_schedule=cfg.profile.step_size_schedule,
                                                      step_size_adaptive_multiplier=cfg.profile.step_size_adaptive_multiplier,
                                                      step_size_adaptive_interval=cfg.profile.step_size_adaptive_interval))
    ref_kernel = new_ref_kernel('ref')
    failed_chains = 0
    for rep in range(reps):
        for p in range(len(values)):
            c = rep * len(values) + p
            z, e = dm.variates(seed, c, 0, n_iter * steps, len(names))
            factor = np.linalg.cholesky(0.5 * (covs[p] + covs[p].T))       # the symmetrised covariance, lower factor
            if 'x2' in ref_kernel.param['fixed']:
                ref_kernel.param['fixed']['x2']['fixed_value'] = values[p]
            else:
                ref_kernel.set_fixed_parameters({'x2': values[p]})
            state = {'t': 0, 'step': step0}

            def normal(mean, cov, state=state, factor=factor, z=z):
                return mean + state['step'] * (factor @ z[state['t']].astype(np.float64))

            def uniform(low=0, high=1, state=state, e=e):
                u = np.exp(-float(e[state['t']]))
                state['t'] += 1
                return u

            monkeypatch.setattr(np.random, 'multivariate_normal', normal)
            monkeypatch.setattr(np.random, 'uniform', uniform)
            change = g.get_initial_step_size_change(cfg)
            ref_settings = {'current_best_loglkl': init_ll[p], 'initial_position': {nm: [v] for nm, v in zip(names, starts[p])},
                            'covmat': covs[p], 'temperature': 0.1, 'temperature_change': g.get_temperature_change(cfg),
                            'step_size': step0, 'step_size_change': change, 'iteration_number': 0,
                            'repetition_number': rep, 'fixed_param_val': values[p]}
            for it in range(n_iter):
                state['step'] = ref_settings['step_size']
                with g.rh.quiet():
                    ref = RefSimulatedAnnealing(ref_cfg, ref_kernel, ref_settings)
                    ref.optimise()
                    ref.set_bestfit()
                bf = ref.bestfit
                assert rec['accepted'][it, c] + 1 == bf['accepted_steps'], (c, it)
                assert rec['best_loglkl'][it, c] == bf['loglkl'], (c, it)
                assert np.array_equal(rec['best_x'][it, c], [bf['position'][nm][0] for nm in names]), (c, it)
                assert rec['temperature'][it, c] == ref_settings['temperature'], (c, it)
                assert rec['step_size'][it, c] == ref_settings['step_size'], (c, it)
                try:
                    nxt = ref.get_next_iteration_settings()
                except UnboundLocalError:               # the reference's own failure path of the 'adaptive' multiplier
                    assert status[c] == capi.FAILED
                    failed_chains += 1
                    break
                nxt['iteration_number'] = it + 1
                nxt['repetition_number'] = rep
                ref_settings = nxt
            else:
                assert status[c] == capi.ACTIVE
                assert np.array_equal(final_x[c], [ref_settings['initial_position'][nm][0] for nm in names])
                assert final_t[c] == ref_settings['temperature'] and final_s[c] == ref_settings['step_size']
    assert failed_chains < reps * len(values)
