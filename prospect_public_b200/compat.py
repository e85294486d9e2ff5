"""Snapshot files that PROSPECT itself can read.

`prospect-analyse <folder>`, `prospect.run.load`, `prospect.profile.load_profile` and `reanneal` unpickle
`<folder>/config.pkl` (a `prospect.input.Configuration`) and `<folder>/state.pkl` (a
`prospect.communication.TasksState` whose `done` dict holds `InitialiseOptimiserTask` / `OptimiseTask` objects,
prospect/communication.py:18-26,159-166; prospect/io.py:29-47).  This module writes exactly those pickles from a
dense `analysis.RunRecord` WITHOUT needing PROSPECT installed: pickle stores classes by module + qualified
name and instances as a `__dict__`, so light stand-in classes registered under PROSPECT's module names
serialise to byte streams that resolve to PROSPECT's real classes when PROSPECT loads them.

Attributes emitted are the ones the reference reads (tasks/analyse_profile_task.py:24-121,
tasks/analyse_global_optimisation_task.py:20-75, profile.py:17-97, communication.py:30-36, io.py:38-47).
By default one OptimiseTask per chain and per iteration that ran is written; `compact=True` keeps only each
chain's best and last iteration (what the analysis needs) so that 8192-chain runs stay small.
"""
import os
import pickle
import sys
import types
from collections import defaultdict
from datetime import datetime

import numpy as np

_STANDINS = {
    'prospect.input': ['Configuration'],
    'prospect.io': ['Arguments'],
    'prospect.run': ['Arguments'],
    'prospect.kernels.base_kernel': ['Arguments'],
    'prospect.profile': ['Arguments'],
    'prospect.mcmc': ['Arguments', 'Chain'],
    'prospect.communication': ['TasksState'],
    'prospect.optimiser': ['SimulatedAnnealing'],
    'prospect.tasks.optimise_task': ['OptimiseTask'],
    'prospect.tasks.initialise_optimiser_task': ['InitialiseOptimiserTask'],
}


class _Registered:
    """Context manager: make PROSPECT's module names importable (stand-ins unless the real package is there)."""

    def __enter__(self):
        self.added = []
        self.cls = {}
        try:
            import prospect.input  # noqa: F401  (real package present: use its classes)
            real = True
        except Exception:
            real = False
        for mod_name, names in _STANDINS.items():
            if real:
                try:
                    mod = __import__(mod_name, fromlist=names)
                    for n in names:
                        self.cls[(mod_name, n)] = getattr(mod, n)
                    continue
                except Exception:
                    pass
            parts = mod_name.split('.')
            for i in range(1, len(parts) + 1):
                name = '.'.join(parts[:i])
                if name not in sys.modules:
                    sys.modules[name] = types.ModuleType(name)
                    self.added.append(name)
            mod = sys.modules[mod_name]
            for n in names:
                if not hasattr(mod, n):
                    setattr(mod, n, type(n, (), {'__module__': mod_name, '__qualname__': n}))
                self.cls[(mod_name, n)] = getattr(mod, n)
        return self

    def new(self, mod_name, cls_name, **attrs):
        cls = self.cls[(mod_name, cls_name)]
        obj = cls.__new__(cls)
        obj.__dict__.update(attrs)
        return obj

    def __exit__(self, *exc):
        for name in reversed(self.added):
            sys.modules.pop(name, None)
        return False


def _plain(v):
    return v.tolist() if isinstance(v, np.ndarray) else v


def _reference_config(reg, config):
    """prospect.input.Configuration with per-section Arguments objects and config_dict."""
    cd = {sec: {k: v for k, v in entries.items()} for sec, entries in config.config_dict.items()}
    if cd['run'].get('mode') == 'gpu':
        cd['run']['mode'] = 'serial'                      # a value prospect.run.Arguments accepts
    if cd['kernel'].get('type') == 'cuda_gaussian':
        cd['kernel']['type'] = 'analytical'
    section_module = {'io': 'prospect.io', 'run': 'prospect.run', 'kernel': 'prospect.kernels.base_kernel',
                      'profile': 'prospect.profile', 'mcmc': 'prospect.mcmc'}
    attrs = {'config_dict': cd}
    for sec, entries in cd.items():
        attrs[sec] = reg.new(section_module[sec], 'Arguments', **entries)
    return reg.new('prospect.input', 'Configuration', **attrs)


def export_reference_snapshot(config, rec, folder, compact=False):
    """Write <folder>/config.pkl and <folder>/state.pkl in PROSPECT's own pickle layout.  Returns the number of
    task objects written."""
    is_profile = config.run.jobtype == 'profile'
    n_pts, reps = rec.n_points, rec.reps
    names = list(rec.names)
    now = datetime.now()
    with _Registered() as reg:
        ref_config = _reference_config(reg, config)
        done = {}
        tid = 1

        def base(task_type_attrs):
            nonlocal tid
            d = dict(config=ref_config, id=tid, required_task_ids=[], success=True, error=None, time_emit=now,
                     time_launch=now, time_finish=now, emitted_by=None)
            d.update(task_type_attrs)
            tid += 1
            return d

        for rep in range(reps):
            for p in range(n_pts):
                bestfit = {nm: [float(rec.initial_position[p][i])] for i, nm in enumerate(names)}
                attrs = dict(repetition_number=rep, initial_bestfit=bestfit,
                             initial_loglkl=float(rec.initial_loglkl[p]),
                             initial_covmat=np.asarray(rec.initial_covmat[p]))
                if is_profile:
                    bestfit[config.profile.parameter] = [float(rec.values[p])]
                    attrs['fixed_param_val'] = rec.values[p]
                    attrs['profile_param_idx'] = int(getattr(rec, 'profile_param_idx', 0))
                t = reg.new('prospect.tasks.initialise_optimiser_task', 'InitialiseOptimiserTask', **base(attrs))
                done[t.id] = t
        cb, ci = rec.per_chain_best()
        ran = rec.accepted >= 0
        steps = config.profile.steps_per_iteration
        for rep in range(reps):
            for p in range(n_pts):
                c = rep * n_pts + p
                its = np.nonzero(ran[:, c])[0]
                if compact and len(its):
                    its = sorted({int(ci[rep, p]), int(its[-1])})
                for k in its:
                    pos = {nm: [float(rec.best_x[k, c, i])] for i, nm in enumerate(names)}
                    settings = {
                        'current_best_loglkl': float(rec.best_loglkl[k - 1, c]) if k > 0 and ran[k - 1, c]
                        else float(rec.initial_loglkl[p]),
                        'initial_position': pos, 'covmat': np.asarray(rec.initial_covmat[p]),
                        'temperature': float(rec.temperature[k, c]) if rec.temperature is not None else None,
                        'temperature_change': None,
                        'step_size': float(rec.step_size[k, c]) if rec.step_size is not None else None,
                        'step_size_change': None, 'iteration_number': int(k), 'repetition_number': rep,
                    }
                    if is_profile:
                        pos[config.profile.parameter] = [float(rec.values[p])]
                        settings['fixed_param_val'] = rec.values[p]
                    acc = int(rec.accepted[k, c])
                    bf = {'loglkl': float(rec.best_loglkl[k, c]), 'acceptance_rate': (1 + acc) / (1.0 + steps),
                          'accepted_steps': 1 + acc, 'iteration_number': int(k), 'position': pos}
                    opt = reg.new('prospect.optimiser', 'SimulatedAnnealing', config=ref_config, settings=settings,
                                  initial_position=pos, covmat=settings['covmat'], bestfit=bf)
                    last = k == its[-1]
                    t = reg.new('prospect.tasks.optimise_task', 'OptimiseTask',
                                **base(dict(optimise_settings=settings, optimiser=opt,
                                            converged=bool(rec.converged[c]) and bool(last))))
                    done[t.id] = t
        state = reg.new('prospect.communication', 'TasksState', queued=[], unready={}, ongoing={}, done=done,
                        failed={}, dependencies=defaultdict(list))
        for name, obj in (('config', ref_config), ('state', state)):
            tmp = os.path.join(folder, f'{name}_backup.pkl')
            with open(tmp, 'wb') as f:
                pickle.dump(obj, f)
            os.replace(tmp, os.path.join(folder, f'{name}.pkl'))
    return len(done)
