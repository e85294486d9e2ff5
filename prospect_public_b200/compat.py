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


# ----------------------------------------------------------------------------------------------------------------------
# Bulk writer: the same state.pkl, written as a pickle byte stream assembled with numpy instead of through half a
# million Python objects.  All OptimiseTask entries of a run have the same shape, so ONE task is laid out as a byte
# template (protocol-2 opcodes, fixed-width encodings: BININT for integers, BINFLOAT for floats, LONG_BINGET /
# LONG_BINPUT for memo references), the template is tiled once per task and the numeric fields are scattered into
# their columns.  Shared objects (classes, the Configuration, key strings, per-point covariance matrices and fixed
# values) are pickled once into a discarded prefix (MARK ... POP_MARK) under memo indices >= 2**20 and referenced from
# the tasks.  pickle.load builds exactly the object graph export_reference_snapshot(compact=True) writes
# (tests/test_host_logic.py compares the two; the unmodified reference loads and analyses it).
# ----------------------------------------------------------------------------------------------------------------------
_G = 1 << 20          # first memo index of the shared objects


def _body(obj):
    """Protocol-2 pickle of `obj` without the PROTO header and the STOP: leaves the object on the unpickler's stack.
    (Its internal memo indices are small and only referenced from inside the same body.)"""
    return pickle.dumps(obj, protocol=2)[2:-1]


class _Emitter:
    """Byte template with named fixed-width placeholders."""

    def __init__(self, memo):
        self.parts, self.fields, self.size, self.memo = [], [], 0, memo

    def raw(self, b):
        self.parts.append(b)
        self.size += len(b)

    def get(self, key):                          # LONG_BINGET of a shared object
        self.raw(b'j' + int(self.memo[key]).to_bytes(4, 'little'))

    def field(self, name, opcode, width):        # opcode + `width` bytes filled per task
        self.raw(opcode)
        self.fields.append((name, self.size, width))
        self.raw(b'\x00' * width)

    def kv(self, key):
        self.get(('key', key))


def export_reference_snapshot_bulk(config, rec, folder):
    """state.pkl / config.pkl of a large run (the `compact` selection of export_reference_snapshot: per chain its best
    and its last iteration, one InitialiseOptimiserTask per profile value), assembled as bytes.  Returns the number of
    task objects written."""
    import struct
    is_profile = config.run.jobtype == 'profile'
    n_pts, names = rec.n_points, list(rec.names)
    param = config.profile.parameter if is_profile else None
    values = np.asarray(rec.values, dtype=np.float64)
    best_ll, ran = np.asarray(rec.best_loglkl), rec.accepted >= 0
    _, ci = rec.per_chain_best()
    ci = ci.reshape(-1)
    k_rows, b = ran.shape
    last = np.where(ran.any(axis=0), k_rows - 1 - np.argmax(ran[::-1], axis=0), -1)
    ks = np.stack([ci, last], axis=1)
    ks.sort(axis=1)
    keep = ks >= 0
    keep[:, 1] &= ks[:, 1] != ks[:, 0]
    chain_of = np.repeat(np.arange(b), 2)[keep.reshape(-1)]
    iter_of = ks.reshape(-1)[keep.reshape(-1)]
    n_tasks = len(iter_of)
    point_of, rep_of = chain_of % n_pts, chain_of // n_pts
    steps = config.profile.steps_per_iteration
    now = datetime.now()
    with _Registered() as reg:
        ref_config = _reference_config(reg, config)
        # ---- shared objects -> memo
        keys = ['config', 'id', 'required_task_ids', 'success', 'error', 'time_emit', 'time_launch', 'time_finish',
                'emitted_by', 'optimise_settings', 'optimiser', 'converged', 'settings', 'initial_position', 'covmat',
                'bestfit', 'loglkl', 'acceptance_rate', 'accepted_steps', 'iteration_number', 'position',
                'current_best_loglkl', 'temperature', 'temperature_change', 'step_size', 'step_size_change',
                'repetition_number', 'fixed_param_val', 'initial_bestfit', 'initial_loglkl', 'initial_covmat',
                'profile_param_idx'] + names + ([param] if is_profile else [])
        memo, prefix, nxt = {}, [b'('], _G
        def share(key, body):
            nonlocal nxt
            memo[key] = nxt
            prefix.append(body + b'r' + nxt.to_bytes(4, 'little'))
            nxt += 1
        for cls_key in (('prospect.tasks.optimise_task', 'OptimiseTask'), ('prospect.optimiser', 'SimulatedAnnealing'),
                        ('prospect.tasks.initialise_optimiser_task', 'InitialiseOptimiserTask'),
                        ('prospect.communication', 'TasksState')):
            share(('cls', cls_key[1]), b'c' + cls_key[0].encode() + b'\n' + cls_key[1].encode() + b'\n')
        share('config', _body(ref_config))
        share('now', _body(now))
        for key in dict.fromkeys(keys):
            share(('key', key), _body(key))
        for p in range(n_pts):
            share(('cov', p), _body(np.asarray(rec.initial_covmat[p])))
            if is_profile:
                share(('val', p), _body(values[p]))
        prefix.append(b'1')                                   # POP_MARK: the stack is empty again, the memo is not

        def common(e, tid_field=True):                        # the BaseTask attributes
            e.kv('config'); e.get('config')
            e.kv('id'); e.field('tid', b'J', 4) if tid_field else None
            e.kv('required_task_ids'); e.raw(b']')
            e.kv('success'); e.raw(b'\x88')
            e.kv('error'); e.raw(b'N')
            for key in ('time_emit', 'time_launch', 'time_finish'):
                e.kv(key); e.get('now')
            e.kv('emitted_by'); e.raw(b'N')

        # ---- InitialiseOptimiserTask: one per value, emitted directly (few objects)
        init_pos = np.asarray(rec.initial_position, dtype=np.float64)
        init_ll = np.asarray(rec.initial_loglkl, dtype=np.float64)
        ppi = int(getattr(rec, 'profile_param_idx', 0))
        chunks = []
        tid = 1
        for p in range(n_pts):
            e = _Emitter(memo)
            e.raw(b'J' + struct.pack('<i', tid))              # key of the `done` dict
            e.get(('cls', 'InitialiseOptimiserTask')); e.raw(b')\x81}(')
            e.kv('config'); e.get('config')
            e.kv('id'); e.raw(b'J' + struct.pack('<i', tid))
            e.kv('required_task_ids'); e.raw(b']')
            e.kv('success'); e.raw(b'\x88')
            e.kv('error'); e.raw(b'N')
            for key in ('time_emit', 'time_launch', 'time_finish'):
                e.kv(key); e.get('now')
            e.kv('emitted_by'); e.raw(b'N')
            e.kv('repetition_number'); e.raw(b'J' + struct.pack('<i', 0))
            e.kv('initial_bestfit'); e.raw(b'}(')
            for i, nm in enumerate(names):
                e.kv(nm); e.raw(b']G' + struct.pack('>d', init_pos[p][i]) + b'a')
            if is_profile:
                e.kv(param); e.raw(b']G' + struct.pack('>d', float(values[p])) + b'a')
            e.raw(b'u')
            e.kv('initial_loglkl'); e.raw(b'G' + struct.pack('>d', init_ll[p]))
            e.kv('initial_covmat'); e.get(('cov', p))
            if is_profile:
                e.kv('fixed_param_val'); e.get(('val', p))
                e.kv('profile_param_idx'); e.raw(b'J' + struct.pack('<i', ppi))
            e.raw(b'ub')
            chunks.append(b''.join(e.parts))
            tid += 1
        first_opt_tid = tid

        # ---- OptimiseTask template
        e = _Emitter(memo)
        e.field('tid', b'J', 4)                               # key of the `done` dict
        e.get(('cls', 'OptimiseTask')); e.raw(b')\x81}(')
        e.kv('config'); e.get('config')
        e.kv('id'); e.field('tid2', b'J', 4)
        e.kv('required_task_ids'); e.raw(b']')
        e.kv('success'); e.raw(b'\x88')
        e.kv('error'); e.raw(b'N')
        for key in ('time_emit', 'time_launch', 'time_finish'):
            e.kv(key); e.get('now')
        e.kv('emitted_by'); e.raw(b'N')
        e.kv('optimise_settings'); e.raw(b'}(')
        e.kv('current_best_loglkl'); e.field('prev', b'G', 8)
        e.kv('initial_position'); e.raw(b'}(')
        for i, nm in enumerate(names):
            e.kv(nm); e.raw(b']'); e.field(('x', i), b'G', 8); e.raw(b'a')
        if is_profile:
            e.kv(param); e.raw(b']'); e.field('value', b'G', 8); e.raw(b'a')
        e.raw(b'u'); e.field('memo_pos', b'r', 4)             # ... and remember the position dict
        e.kv('covmat'); e.field('cov', b'j', 4)
        e.kv('temperature'); e.field('temp', b'G', 8)
        e.kv('temperature_change'); e.raw(b'N')
        e.kv('step_size'); e.field('step', b'G', 8)
        e.kv('step_size_change'); e.raw(b'N')
        e.kv('iteration_number'); e.field('iter', b'J', 4)
        e.kv('repetition_number'); e.field('rep', b'J', 4)
        if is_profile:
            e.kv('fixed_param_val'); e.field('val', b'j', 4)
        e.raw(b'u'); e.field('memo_set', b'r', 4)
        e.kv('optimiser'); e.get(('cls', 'SimulatedAnnealing')); e.raw(b')\x81}(')
        e.kv('config'); e.get('config')
        e.kv('settings'); e.field('get_set', b'j', 4)
        e.kv('initial_position'); e.field('get_pos', b'j', 4)
        e.kv('covmat'); e.field('cov2', b'j', 4)
        e.kv('bestfit'); e.raw(b'}(')
        e.kv('loglkl'); e.field('best', b'G', 8)
        e.kv('acceptance_rate'); e.field('ar', b'G', 8)
        e.kv('accepted_steps'); e.field('acc1', b'J', 4)
        e.kv('iteration_number'); e.field('iter2', b'J', 4)
        e.kv('position'); e.field('get_pos2', b'j', 4)
        e.raw(b'uub')                                         # bestfit dict, optimiser __dict__, BUILD
        e.kv('converged'); e.field('conv', b'', 1)
        e.raw(b'ub')
        template = np.frombuffer(b''.join(e.parts), dtype=np.uint8)
        out = np.tile(template, (n_tasks, 1))
        off = {name: (o, w) for name, o, w in e.fields}

        def put(name, arr, kind):
            o, w = off[name]
            if kind == 'f':
                out[:, o:o + w] = np.ascontiguousarray(arr, dtype='>f8').view(np.uint8).reshape(n_tasks, 8)
            else:
                out[:, o:o + w] = np.ascontiguousarray(arr, dtype='<i4').view(np.uint8).reshape(n_tasks, 4)

        tids = first_opt_tid + np.arange(n_tasks)
        acc = np.asarray(rec.accepted)[iter_of, chain_of].astype(np.int64)
        prev_ok = (iter_of > 0) & ran[np.maximum(iter_of - 1, 0), chain_of]
        put('tid', tids, 'i'); put('tid2', tids, 'i')
        put('prev', np.where(prev_ok, best_ll[np.maximum(iter_of - 1, 0), chain_of], init_ll[point_of]), 'f')
        xs = np.asarray(rec.best_x)[iter_of, chain_of]
        for i in range(len(names)):
            put(('x', i), xs[:, i], 'f')
        if is_profile:
            put('value', values[point_of], 'f')
            put('val', np.array([memo[('val', p)] for p in range(n_pts)])[point_of], 'i')
        cov_idx = np.array([memo[('cov', p)] for p in range(n_pts)])[point_of]
        put('cov', cov_idx, 'i'); put('cov2', cov_idx, 'i')
        put('temp', rec.temperature[iter_of, chain_of], 'f')
        put('step', rec.step_size[iter_of, chain_of], 'f')
        put('iter', iter_of, 'i'); put('iter2', iter_of, 'i'); put('rep', rep_of, 'i')
        per_task = nxt + 2 * np.arange(n_tasks)               # two memo slots per task: position dict, settings dict
        put('memo_pos', per_task, 'i'); put('get_pos', per_task, 'i'); put('get_pos2', per_task, 'i')
        put('memo_set', per_task + 1, 'i'); put('get_set', per_task + 1, 'i')
        put('best', best_ll[iter_of, chain_of], 'f')
        put('ar', (1 + acc) / (1.0 + steps), 'f')
        put('acc1', 1 + acc, 'i')
        conv = np.asarray(rec.converged, dtype=bool)[chain_of] & (iter_of == last[chain_of])
        o, _ = off['conv']
        out[:, o] = np.where(conv, 0x88, 0x89)

        stream = [b'\x80\x02', b''.join(prefix)]
        stream.append(b'j' + memo[('cls', 'TasksState')].to_bytes(4, 'little') + b')\x81}(')
        for key in ('queued',):
            stream.append(_body(key) + b']')
        for key in ('unready', 'ongoing'):
            stream.append(_body(key) + b'}')
        stream.append(_body('done') + b'}(' + b''.join(chunks))
        stream.append(out.tobytes())
        stream.append(b'u')
        stream.append(_body('failed') + b'}')
        stream.append(_body('dependencies') + _body(defaultdict(list)))
        stream.append(b'ub.')
        tmp = os.path.join(folder, 'state_backup.pkl')
        with open(tmp, 'wb') as f:
            for piece in stream:
                f.write(piece)
        os.replace(tmp, os.path.join(folder, 'state.pkl'))
        tmp = os.path.join(folder, 'config_backup.pkl')
        with open(tmp, 'wb') as f:
            pickle.dump(ref_config, f, protocol=pickle.HIGHEST_PROTOCOL)
        os.replace(tmp, os.path.join(folder, 'config.pkl'))
    return n_pts + n_tasks
