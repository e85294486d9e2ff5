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
    task objects written.  `compact` (large runs): per chain only its best and its last iteration, and one
    InitialiseOptimiserTask per profile value (the reference's analysis reads the first one it finds for a value,
    analyse_profile_task.py:113-119) -- a 8192-chain run is then ~16 000 task objects instead of ~170 000.

    Everything numeric is converted to Python scalars in bulk (ndarray.tolist) and the attribute dicts are built
    directly: this function is pure host work on the critical path of "time to result", so per-object overhead
    matters (65 536 chains on 8 GPUs are half a million dict entries)."""
    is_profile = config.run.jobtype == 'profile'
    n_pts, reps = rec.n_points, rec.reps
    names = list(rec.names)
    now = datetime.now()
    param = config.profile.parameter if is_profile else None
    values = np.asarray(rec.values, dtype=np.float64)
    init_pos = np.asarray(rec.initial_position, dtype=np.float64).tolist()
    init_ll = np.asarray(rec.initial_loglkl, dtype=np.float64).tolist()
    init_cov = [np.asarray(c) for c in rec.initial_covmat]
    best_ll = np.asarray(rec.best_loglkl)
    ran = rec.accepted >= 0
    cb, ci = rec.per_chain_best()
    ci = ci.reshape(-1)
    k_rows, b = ran.shape
    last = np.where(ran.any(axis=0), k_rows - 1 - np.argmax(ran[::-1], axis=0), -1)
    if compact:                                   # (iteration, chain) pairs to emit, chain-major
        ks = np.stack([ci, last], axis=1)
        keep = (ks >= 0) & np.stack([np.ones(b, dtype=bool), last != ci], axis=1)
        ks.sort(axis=1)
        keep = (ks >= 0)
        keep[:, 1] &= ks[:, 1] != ks[:, 0]
        chain_of = np.repeat(np.arange(b), 2)[keep.reshape(-1)]
        iter_of = ks.reshape(-1)[keep.reshape(-1)]
    else:
        iter_of, chain_of = np.nonzero(ran)
        order = np.lexsort((iter_of, chain_of))
        iter_of, chain_of = iter_of[order], chain_of[order]
    # per emitted task, in bulk
    t_best = best_ll[iter_of, chain_of].tolist()
    prev_ok = (iter_of > 0) & ran[np.maximum(iter_of - 1, 0), chain_of]
    t_prev = np.where(prev_ok, best_ll[np.maximum(iter_of - 1, 0), chain_of],
                      np.asarray(rec.initial_loglkl)[chain_of % n_pts]).tolist()
    t_acc = np.asarray(rec.accepted)[iter_of, chain_of].tolist()
    t_temp = rec.temperature[iter_of, chain_of].tolist() if rec.temperature is not None else [None] * len(iter_of)
    t_step = rec.step_size[iter_of, chain_of].tolist() if rec.step_size is not None else [None] * len(iter_of)
    t_pos = np.asarray(rec.best_x)[iter_of, chain_of].tolist()
    t_conv = (np.asarray(rec.converged, dtype=bool)[chain_of] & (iter_of == last[chain_of])).tolist()
    iter_l, chain_l = iter_of.tolist(), chain_of.tolist()
    steps = config.profile.steps_per_iteration
    with _Registered() as reg:
        ref_config = _reference_config(reg, config)
        init_cls = reg.cls[('prospect.tasks.initialise_optimiser_task', 'InitialiseOptimiserTask')]
        opt_cls = reg.cls[('prospect.tasks.optimise_task', 'OptimiseTask')]
        sa_cls = reg.cls[('prospect.optimiser', 'SimulatedAnnealing')]
        done = {}
        tid = 1
        ppi = int(getattr(rec, 'profile_param_idx', 0))
        for rep in range(1 if compact else reps):
            for p in range(n_pts):
                bestfit = {nm: [v] for nm, v in zip(names, init_pos[p])}
                d = {'config': ref_config, 'id': tid, 'required_task_ids': [], 'success': True, 'error': None,
                     'time_emit': now, 'time_launch': now, 'time_finish': now, 'emitted_by': None,
                     'repetition_number': rep, 'initial_bestfit': bestfit, 'initial_loglkl': init_ll[p],
                     'initial_covmat': init_cov[p]}
                if is_profile:
                    bestfit[param] = [float(values[p])]
                    d['fixed_param_val'] = values[p]
                    d['profile_param_idx'] = ppi
                t = init_cls.__new__(init_cls)
                t.__dict__.update(d)
                done[tid] = t
                tid += 1
        for j in range(len(iter_l)):
            k, c = iter_l[j], chain_l[j]
            rep, p = divmod(c, n_pts)
            pos = {nm: [v] for nm, v in zip(names, t_pos[j])}
            settings = {'current_best_loglkl': t_prev[j], 'initial_position': pos, 'covmat': init_cov[p],
                        'temperature': t_temp[j], 'temperature_change': None, 'step_size': t_step[j],
                        'step_size_change': None, 'iteration_number': k, 'repetition_number': rep}
            if is_profile:
                pos[param] = [float(values[p])]
                settings['fixed_param_val'] = values[p]
            acc = t_acc[j]
            opt = sa_cls.__new__(sa_cls)
            opt.__dict__.update({'config': ref_config, 'settings': settings, 'initial_position': pos,
                                 'covmat': init_cov[p],
                                 'bestfit': {'loglkl': t_best[j], 'acceptance_rate': (1 + acc) / (1.0 + steps),
                                             'accepted_steps': 1 + acc, 'iteration_number': k, 'position': pos}})
            t = opt_cls.__new__(opt_cls)
            t.__dict__.update({'config': ref_config, 'id': tid, 'required_task_ids': [], 'success': True,
                               'error': None, 'time_emit': now, 'time_launch': now, 'time_finish': now,
                               'emitted_by': None, 'optimise_settings': settings, 'optimiser': opt,
                               'converged': t_conv[j]})
            done[tid] = t
            tid += 1
        state = reg.new('prospect.communication', 'TasksState', queued=[], unready={}, ongoing={}, done=done,
                        failed={}, dependencies=defaultdict(list))
        for name, obj in (('config', ref_config), ('state', state)):
            tmp = os.path.join(folder, f'{name}_backup.pkl')
            with open(tmp, 'wb') as f:
                pickle.dump(obj, f, protocol=pickle.HIGHEST_PROTOCOL)
            os.replace(tmp, os.path.join(folder, f'{name}.pkl'))
    return len(done)
