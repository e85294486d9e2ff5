#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (imported from /root/reference) in this container.

Run:  python tests/golden/make_golden.py            (writes tests/golden/*.npz|*.json|*.txt)

The reference cannot travel to the GPU box, so its outputs are frozen here.
Import shims (tests/golden/_shims) stand in for getdist / matplotlib which are
not installed; nothing from the reference source is copied.

Harness obligations:
  * the reference never stops annealing -> OptimiseTask.emit_tasks is wrapped to
    stop after `max_iterations` iterations (k = 0 .. max_iterations-1);
  * Scheduler.finalize is broken -> the analysis task is run by hand;
  * everything runs in `mode: serial` with `numpy_random_seed` set.
"""
import contextlib
import copy
import io
import json
import os
import pickle
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
sys.path.insert(0, os.path.join(HERE, '_shims'))
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import yaml  # noqa: E402

import getdist  # the shim  # noqa: E402
from prospect.input import Configuration  # noqa: E402
from prospect.io import prepare_run  # noqa: E402
from prospect.communication import Scheduler, SerialContext  # noqa: E402
from prospect.tasks.optimise_task import OptimiseTask  # noqa: E402
from prospect.tasks.base_task import BaseTask  # noqa: E402
from prospect.kernels.initialisation import initialise_kernel  # noqa: E402
from prospect.tasks import initialise_optimiser_task as iot  # noqa: E402

QUIET = True


@contextlib.contextmanager
def quiet():
    if QUIET:
        with contextlib.redirect_stdout(io.StringIO()):
            yield
    else:
        yield


@contextlib.contextmanager
def truncated_annealing(max_iter):
    """Stop every chain after `max_iter` OptimiseTasks."""
    orig = OptimiseTask.emit_tasks

    def emit(self):
        if self.optimise_settings['iteration_number'] + 1 >= max_iter:
            return []
        return orig(self)
    OptimiseTask.emit_tasks = emit
    try:
        yield
    finally:
        OptimiseTask.emit_tasks = orig


def run_scheduler(config_yaml, max_iter=None):
    """Configuration -> prepare_run -> Scheduler.delegate(SerialContext)."""
    with quiet():
        config = Configuration(copy.deepcopy(config_yaml))
        prepare_run(config)
        sched = Scheduler(config, False)
        ctx = truncated_annealing(max_iter) if max_iter else contextlib.nullcontext()
        with ctx:
            sched.delegate(SerialContext())
    assert len(sched.tasks.failed) == 0, [t.error for t in sched.tasks.failed.values()]
    return config, sched


def base_yaml(jobtype, outdir, kernel_param, seed=0):
    return {
        'io': {'jobname': 'test', 'write': True, 'dir': outdir, 'overwrite_dir': True,
               'snapshot_interval': 1.0e9},
        'run': {'jobtype': jobtype, 'mode': 'serial', 'numpy_random_seed': seed},
        'kernel': {'type': 'analytical', 'conf': '', 'param': kernel_param},
    }


def gen_mcmc(dimension, work):
    """Bootstrap MCMC the way the shipped input/example_toy/mcmc/test_* were made
    (SURVEY.md F8a): seed 0, random_gaussian recipe, 4 chains x 4 rounds x 2000 steps."""
    kparam = os.path.join(work, f'kernel_random_gauss_{dimension}.yaml')
    with open(kparam, 'w') as f:
        yaml.dump({'function': 'random_gaussian', 'dimension': dimension,
                   'std_scale': 0.1, 'off_diag_factor': 0.05}, f)
    outdir = os.path.join(work, f'mcmc{dimension}')
    cfg = base_yaml('mcmc', outdir, kparam)
    cfg['mcmc'] = {'algorithm': 'MetropolisHastings', 'N_chains': 4, 'steps_per_iteration': 2000,
                   'convergence_Rm1': 0.01, 'unpack_at_dump': True}
    getdist.GELMAN_RUBIN_SCRIPT[:] = [1.0, 1.0, 1.0, 0.0]   # 4 rounds
    config, sched = run_scheduler(cfg)
    with open(os.path.join(outdir, 'analytical', 'random_gauss.pkl'), 'rb') as f:
        kd = pickle.load(f)
    names = [f'x{i}' for i in range(1, dimension + 1)]
    means = np.array([kd['means'][n] for n in names])
    chains = [np.loadtxt(os.path.join(outdir, 'mcmc', f'test_{i}.txt')) for i in range(4)]
    return {'means': means, 'covmat': np.array(kd['covmat']), 'dimension': dimension}, chains, outdir


def write_kernel_pickle(kernel, path):
    d = kernel['dimension']
    with open(path, 'wb') as f:
        pickle.dump({'means': {f'x{i+1}': float(kernel['means'][i]) for i in range(d)},
                     'covmat': kernel['covmat'], 'dimension': d}, f)


def write_chain_files(chains, d, root):
    os.makedirs(os.path.dirname(root), exist_ok=True)
    for i, c in enumerate(chains):
        np.savetxt(f'{root}_{i}.txt', c, delimiter=' ')
    with open(root + '.paramnames', 'w') as f:
        for i in range(1, d + 1):
            f.write(f'\nx{i} x{i}')
    with open(root + '.ranges', 'w') as f:
        for i in range(1, d + 1):
            f.write(f'\nx{i} -2.5 2.5')


def profile_yaml(work, tag, kernel_pkl, mcmc_root, jobtype='profile', **profile_kw):
    kparam = os.path.join(work, f'kernel_load_{tag}.yaml')
    with open(kparam, 'w') as f:
        yaml.dump({'function': 'random_gaussian', 'load': kernel_pkl}, f)
    cfg = base_yaml(jobtype, os.path.join(work, f'out_{tag}'), kparam)
    prof = {
        'optimiser': 'simulated annealing', 'temperature_schedule': 'exponential',
        'temperature_range': [0.1, 0.001], 'max_iterations': 20,
        'step_size_schedule': 'adaptive', 'step_size_adaptive_interval': [0.19, 0.21],
        'step_size_adaptive_multiplier': 0.3, 'step_size_adaptive_initial': 0.1,
        'steps_per_iteration': 250, 'repetitions': 1, 'start_from_mcmc': mcmc_root,
        'start_bin_fraction': 0.1, 'plot_profile': False, 'detailed_plot': False,
        'plot_Delta_chi2': False, 'plot_schedule': False, 'confidence_levels': [0.68, 0.95],
    }
    if jobtype == 'profile':
        prof['parameter'] = 'x2'
        prof['values'] = 'np.linspace(-1.0, 2.0, 10)'
    prof.update(profile_kw)
    cfg['profile'] = prof
    return cfg


def collect_tasks(config, sched):
    """Flatten done tasks into plain arrays (what set_bestfit / get_next_iteration_settings produced)."""
    init, opt = [], []
    for tid, t in sorted(sched.tasks.done.items()):
        if t.type == 'InitialiseOptimiserTask':
            init.append({
                'id': tid, 'fixed_param_val': getattr(t, 'fixed_param_val', None),
                'repetition_number': t.repetition_number,
                'initial_loglkl': float(t.initial_loglkl),
                'initial_names': list(t.initial_bestfit.keys()),
                'initial_position': [float(v[0]) for v in t.initial_bestfit.values()],
                'initial_covmat': np.asarray(t.initial_covmat).tolist(),
            })
        elif t.type == 'OptimiseTask':
            s = t.optimise_settings
            bf = t.optimiser.bestfit
            ch = t.optimiser.mcmc.chain
            opt.append({
                'id': tid, 'fixed_param_val': s.get('fixed_param_val'),
                'repetition_number': s['repetition_number'], 'iteration_number': s['iteration_number'],
                'temperature': float(s['temperature']), 'step_size': float(s['step_size']),
                'current_best_loglkl': float(s['current_best_loglkl']),
                'bestfit_loglkl': float(bf['loglkl']), 'acceptance_rate': float(bf['acceptance_rate']),
                'accepted_steps': int(bf['accepted_steps']),
                'position_names': list(bf['position'].keys()),
                'bestfit_position': [float(v[0]) for v in bf['position'].values()],
                'last_position': [float(v[0]) for v in ch.last_position.values()],
                'last_loglkl': float(ch.loglkls[-1]), 'sum_mults': int(np.sum(ch.mults)),
            })
    return {'initialise': init, 'optimise': opt}


def run_analysis(config, sched, subdir, files):
    with quiet():
        task = sched.get_analysis_task()
        task.run([t for t in sched.tasks.done.values() if t.id in task.required_task_ids])
    out = {}
    for fn in files:
        with open(os.path.join(config.io.dir, subdir, fn)) as f:
            out[fn] = f.read()
    return out


def binning_fixture(cfg_yaml, values):
    """A14 bookkeeping: indices picked by get_points_in_bin + what run() derives from them."""
    with quiet():
        config = Configuration(copy.deepcopy(cfg_yaml))
        prepare_run(config)
    rec = {'values': [], 'N_points': [], 'indices': [], 'covmat': [], 'position': [], 'initial_loglkl': [],
           'profile_param_idx': []}
    captured = {}
    orig = iot.get_points_in_bin

    def spy(chain, param_val, param_name, N_points):
        distance_points = np.abs(np.array(chain.positions[param_name]) - param_val)
        captured['idx'] = np.argsort(distance_points)[:N_points]
        captured['N'] = N_points
        return orig(chain, param_val, param_name, N_points)
    iot.get_points_in_bin = spy
    try:
        for v in values:
            BaseTask.idx_count = 1
            with quiet():
                t = iot.InitialiseOptimiserTask(config, v, 0)
                t.run(None)
            rec['values'].append(float(v))
            rec['N_points'].append(int(captured['N']))
            rec['indices'].append(np.asarray(captured['idx'], dtype=np.int64))
            rec['covmat'].append(np.asarray(t.initial_covmat))
            rec['position'].append(np.array([x[0] for x in t.initial_bestfit.values()]))
            rec['initial_loglkl'].append(float(t.initial_loglkl))
            rec['profile_param_idx'].append(int(t.profile_param_idx))
    finally:
        iot.get_points_in_bin = orig
    return {k: np.array(v) for k, v in rec.items()}


def scipy_profile(cfg_yaml, values):
    with quiet():
        config = Configuration(copy.deepcopy(cfg_yaml))
        prepare_run(config)
        kernel = initialise_kernel(config.kernel, config.io.dir, 0)
        kernel.set_fixed_parameters({'x2': 0.0})
        return np.array([kernel.get_scipy_profile('x2', float(v)) for v in values])


def loglkl_samples(cfg_yaml, n=16, seed=123):
    """Random positions -> reference loglkl, free and with x2 fixed (pins F4)."""
    rng = np.random.default_rng(seed)
    with quiet():
        config = Configuration(copy.deepcopy(cfg_yaml))
        prepare_run(config)
        k_free = initialise_kernel(config.kernel, config.io.dir, 0)
        k_fix = initialise_kernel(config.kernel, config.io.dir, 0)
        k_fix.set_fixed_parameters({'x2': 0.75})
    d = k_free.dimension
    X = rng.uniform(-1.0, 1.5, size=(n, d))
    free = [k_free.loglkl({f'x{i+1}': [X[r, i]] for i in range(d)}) for r in range(n)]
    names_fix = k_fix.varying_param_names
    fix = [k_fix.loglkl({nm: [X[r, int(nm[1:]) - 1]] for nm in names_fix}) for r in range(n)]
    return {'X': X, 'free': np.array(free), 'fixed_x2_0p75': np.array(fix)}


def main():
    work = tempfile.mkdtemp(prefix='golden_')
    os.chdir(REF)  # example yamls use paths relative to the reference root
    out = {}

    # ---- (iii) shipped kernel pickle + recipe reproduction (F3) --------------------
    with open(os.path.join(REF, 'input/example_toy/kernel_settings/random_gauss.pkl'), 'rb') as f:
        shipped = pickle.load(f)
    for d in (20, 30):
        kernel, chains, _ = gen_mcmc(d, work)
        if d == 20:
            names = [f'x{i}' for i in range(1, 21)]
            assert np.array_equal(kernel['covmat'], shipped['covmat'])
            assert np.array_equal(kernel['means'], np.array([shipped['means'][n] for n in names]))
            # (i) shipped chains == regenerated chains (multiset of chains; F8a)
            ship = [np.loadtxt(os.path.join(REF, f'input/example_toy/mcmc/test_{i}.txt')) for i in range(4)]
            for c in chains:
                assert any(c.shape == s.shape and np.allclose(c, s, rtol=0, atol=1e-10) for s in ship)
            chains = ship  # freeze the shipped order (what example_toy.yaml reads)
            print('20-D: recipe == shipped pickle; regenerated chains == shipped chains')
        np.savez(os.path.join(HERE, f'toy{d}_kernel.npz'), means=kernel['means'], covmat=kernel['covmat'])
        np.savez(os.path.join(HERE, f'toy{d}_mcmc.npz'), **{f'chain_{i}': c for i, c in enumerate(chains)})
        out[d] = (kernel, chains)

    summary = {}
    for d in (20, 30):
        kernel, chains = out[d]
        kpkl = os.path.join(work, f'random_gauss_{d}.pkl')
        write_kernel_pickle(kernel, kpkl)
        root = os.path.join(work, f'mcmc_in_{d}', 'test')
        write_chain_files(chains, d, root)
        nvals = 10 if d == 20 else 32
        values = np.linspace(-1.0, 2.0, nvals)

        # ---- A14 bookkeeping ----------------------------------------------------
        cfg = profile_yaml(work, f'bin{d}', kpkl, root, values=f'np.linspace(-1.0, 2.0, {nvals})')
        np.savez(os.path.join(HERE, f'toy{d}_init.npz'), **binning_fixture(cfg, values))
        # ---- O2 / (iv): scipy profile + loglkl samples -----------------------------
        sp = scipy_profile(cfg, values)
        ls = loglkl_samples(cfg)
        np.savez(os.path.join(HERE, f'toy{d}_loglkl.npz'), values=values, scipy_profile=sp, **ls)

        # ---- K1-style serial annealing run, seed 0 ---------------------------------
        cfg = profile_yaml(work, f'prof{d}', kpkl, root, values=f'np.linspace(-1.0, 2.0, {nvals})')
        if d == 30:
            cfg['profile']['values'] = 'np.linspace(-1.0, 2.0, 4)'   # keep the fixture small
        config, sched = run_scheduler(cfg, max_iter=20)
        rec = collect_tasks(config, sched)
        rec['files'] = run_analysis(config, sched, 'profile', ['x2.txt', 'x2_status.txt', 'x2_intervals.txt'])
        with open(os.path.join(HERE, f'toy{d}_profile_run.json'), 'w') as f:
            json.dump(rec, f)
        summary[f'toy{d}_profile_tasks'] = len(rec['optimise'])

    # ---- K2 as shipped (20-D global optimisation, 3 reps) ------------------------------
    kernel, chains = out[20]
    kpkl = os.path.join(work, 'random_gauss_20.pkl')
    root = os.path.join(work, 'mcmc_in_20', 'test')
    cfg = profile_yaml(work, 'glob20', kpkl, root, jobtype='global_optimisation', repetitions=3)
    config, sched = run_scheduler(cfg, max_iter=20)
    rec = collect_tasks(config, sched)
    rec['files'] = run_analysis(config, sched, 'global_optimisation', ['results.txt'])
    with open(os.path.join(HERE, 'toy20_global_run.json'), 'w') as f:
        json.dump(rec, f)

    # ---- schedule variants: exponential step size; 'adaptive' multiplier (A5) ----------
    cfg = profile_yaml(work, 'exp20', kpkl, root, step_size_schedule='exponential',
                       step_size_range=[0.5, 0.05], values='[0.5]')
    del cfg['profile']['step_size_adaptive_multiplier'], cfg['profile']['step_size_adaptive_initial']
    config, sched = run_scheduler(cfg, max_iter=8)
    with open(os.path.join(HERE, 'toy20_exponential_run.json'), 'w') as f:
        json.dump(collect_tasks(config, sched), f)

    cfg = profile_yaml(work, 'sec20', kpkl, root, step_size_adaptive_multiplier='adaptive',
                       step_size_adaptive_interval=[0.24, 0.26], step_size_adaptive_initial=0.3,
                       values='[0.5, 1.0]', steps_per_iteration=400)
    try:
        config, sched = run_scheduler(cfg, max_iter=10)
        rec = collect_tasks(config, sched)
        rec['failed'] = False
    except AssertionError as e:  # the reference's unbound-local / zero-division paths
        rec = {'failed': True, 'error': str(e)[-400:]}
    with open(os.path.join(HERE, 'toy20_secant_run.json'), 'w') as f:
        json.dump(rec, f)

    # ---- the reference's own fixture: test/debug_reanneal/profile/x2.txt initial_loglkl (F8b)
    rows = []
    with open(os.path.join(REF, 'test/debug_reanneal/profile/x2.txt')) as f:
        header = f.readline()
        for line in f:
            parts = [p.strip() for p in line.split('\t')]
            if 'no optimisations' in line:
                continue
            rows.append([float(parts[0]), float(parts[4])])
    with open(os.path.join(HERE, 'ref_debug_reanneal_initial_loglkl.json'), 'w') as f:
        json.dump({'source': 'test/debug_reanneal/profile/x2.txt (columns x2, initial_loglkl)', 'rows': rows}, f)

    # ---- input schema: what Configuration makes of the shipped example yamls (C10/C11 defaults) ----
    cfgs = {}
    for name, path in (('example_toy', 'input/example_toy/example_toy.yaml'),
                       ('example_global', 'input/example_toy/example_global_optimisation_toy.yaml')):
        with open(os.path.join(REF, path)) as f:
            y = yaml.full_load(f)
        y['io']['dir'] = os.path.join(work, 'cfg_' + name)
        with quiet():
            c = Configuration(y)
        cfgs[name] = {sec: {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in d_.items()}
                      for sec, d_ in c.config_dict.items()}
        cfgs[name]['io']['dir'] = '<dir>'
    y = base_yaml('mcmc', os.path.join(work, 'cfg_mcmc'), 'input/example_toy/kernel_settings/kernel_random_gauss.yaml')
    y['mcmc'] = {'algorithm': 'MetropolisHastings', 'N_chains': 4, 'steps_per_iteration': 2000,
                 'convergence_Rm1': 0.01, 'unpack_at_dump': True}
    with quiet():
        c = Configuration(y)
    cfgs['mcmc'] = {sec: dict(d_) for sec, d_ in c.config_dict.items()}
    cfgs['mcmc']['io']['dir'] = '<dir>'
    with open(os.path.join(HERE, 'config_defaults.json'), 'w') as f:
        json.dump(cfgs, f, indent=1)

    # ---- cold K1 run of the real reference: the "reference CPU run on identical inputs" for the GPU ----
    kernel, chains = out[20]
    cfg = profile_yaml(work, 'cold20', kpkl, root, temperature_range=[0.1, 1.0e-6], max_iterations=30)
    config, sched = run_scheduler(cfg, max_iter=30)
    rec = collect_tasks(config, sched)
    keep = {'files': run_analysis(config, sched, 'profile', ['x2.txt']),
            'final': [o for o in rec['optimise'] if o['iteration_number'] == 29]}
    with open(os.path.join(HERE, 'toy20_cold_reference_run.json'), 'w') as f:
        json.dump(keep, f)

    print(json.dumps(summary))
    shutil.rmtree(work, ignore_errors=True)


if __name__ == '__main__':
    main()
