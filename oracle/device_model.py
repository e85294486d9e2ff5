"""ctypes wrapper of oracle/device_model.c -- TEST INFRASTRUCTURE ONLY.

The C file restates, one chain at a time, what the CUDA kernels of libpb200 compute, in the same
floating-point order, so tests can compare the GPU bit for bit (given the GPU's own variates through
pb200_rng_variates) or statistically (own Philox + float64 Box-Muller).  It takes the same padded
arrays as `prospect_public_b200.problem.HostProblem` purely as a data layout; no code is shared.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'libdevice_model.so')

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)


class DmProblem(C.Structure):
    _fields_ = [('d', C.c_int32), ('n', C.c_int32), ('npad', C.c_int32), ('n_points', C.c_int32),
                ('saa', _dp), ('sab', _dp), ('mu', _dp), ('lo', _dp), ('hi', _dp), ('f', _dp), ('c0', _dp),
                ('lt_t', _fp), ('lam', _fp), ('lt_norm', _fp), ('g_t', _dp), ('h', _dp)]


class DmSchedule(C.Structure):
    _fields_ = [('steps_per_iteration', C.c_int32), ('step_mode', C.c_int32), ('temperature_change', C.c_double),
                ('step_size_change', C.c_double), ('adaptive_lo', C.c_double), ('adaptive_hi', C.c_double),
                ('adaptive_multiplier', C.c_double), ('chi2_tolerance', C.c_double)]


class DmChain(C.Structure):
    _fields_ = [('x', _dp), ('loglkl', C.c_double), ('temperature', C.c_double), ('step_size', C.c_double),
                ('current_best', C.c_double), ('prev_mult', C.c_double), ('cur_mult', C.c_double),
                ('prev_ar', C.c_double), ('point', C.c_int32), ('iteration', C.c_int32), ('status', C.c_int32),
                ('stage', C.c_int32), ('chain_id', C.c_uint32), ('steps_done', C.c_uint32)]


def build(force=False):
    src = os.path.join(HERE, 'device_model.c')
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(['make', '-C', HERE, '-B'], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        h = C.CDLL(build())
        h.dm_exact_eval.restype = C.c_double
        h.dm_exact_eval.argtypes = [C.POINTER(DmProblem), C.c_int, _dp, _dp, C.c_int]
        h.dm_mh_chain.restype = C.c_int32
        h.dm_variates.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, _fp, _fp]
        _lib = h
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    return a.ctypes.data_as(_fp)


class Model:
    """Holds contiguous copies of the problem arrays and exposes the C entry points."""

    def __init__(self, hp, lanes=32):
        """lanes: lanes per chain of the device geometry being mirrored (32, 16 or 8)."""
        self.hp = hp
        self.lanes = int(lanes)
        self.np_ = hp.npad
        self.arr = {k: np.ascontiguousarray(getattr(hp, k), dtype=np.float64)
                    for k in ('saa', 'sab', 'mu', 'lo', 'hi', 'f', 'c0', 'g_t', 'h')}
        self.arr32 = {k: np.ascontiguousarray(getattr(hp, k), dtype=np.float32) for k in ('lt_t', 'lam', 'lt_norm')}
        a, b = self.arr, self.arr32
        self.c = DmProblem(hp.d, hp.n, hp.npad, hp.n_points, _d(a['saa']), _d(a['sab']), _d(a['mu']), _d(a['lo']),
                           _d(a['hi']), _d(a['f']), _d(a['c0']), _f(b['lt_t']), _f(b['lam']), _f(b['lt_norm']), _d(a['g_t']),
                           _d(a['h']))

    def exact_eval(self, x, point=0, with_grad=False):
        xp = np.zeros(self.np_)
        xp[:len(x)] = x
        gt = np.zeros(self.np_) if with_grad else None
        ll = lib().dm_exact_eval(C.byref(self.c), int(point), _d(xp), _d(gt) if with_grad else None, self.lanes)
        return (ll, gt) if with_grad else ll

    def new_chain(self, x0, point=0, chain_id=0, temperature=1.0, step_size=1.0, current_best=np.inf, steps_done=0):
        xp = np.zeros(self.np_)
        xp[:self.hp.n] = np.asarray(x0, dtype=np.float64)[:self.hp.n]
        ch = DmChain(_d(xp), 0.0, temperature, step_size, current_best, 0.0, 0.0, 0.0, int(point), 0, 0, 0,
                     int(chain_id), int(steps_done))
        ch._x = xp       # keep alive
        return ch

    def anneal_chain(self, chain, sched: DmSchedule, n_iter, seed, z=None, e=None):
        k, np_ = n_iter, self.np_
        out = {'best_loglkl': np.full(k, np.inf), 'last_loglkl': np.full(k, np.inf), 'temperature': np.zeros(k),
               'step_size': np.zeros(k), 'accepted': np.full(k, -1, dtype=np.int32),
               'best_x': np.zeros((k, np_)), 'last_x': np.zeros((k, np_))}
        if z is not None:
            z = np.ascontiguousarray(z, dtype=np.float32)
            e = np.ascontiguousarray(e, dtype=np.float32)
        lib().dm_anneal_chain(C.byref(self.c), C.byref(sched), C.byref(chain), C.c_int32(n_iter), C.c_uint64(seed),
                              C.c_int(self.lanes),
                              _f(z) if z is not None else None, _f(e) if e is not None else None,
                              _d(out['best_loglkl']), _d(out['last_loglkl']), _d(out['temperature']),
                              _d(out['step_size']), out['accepted'].ctypes.data_as(_ip), _d(out['best_x']),
                              _d(out['last_x']))
        return out

    def mh_chain(self, chain, n_steps, seed, z=None, e=None, mode=0, thin=1, capacity=0, samples=None):
        if z is not None:
            z = np.ascontiguousarray(z, dtype=np.float32)
            e = np.ascontiguousarray(e, dtype=np.float32)
        if mode and samples is None:
            samples = {'x': np.zeros((capacity, self.np_)), 'loglkl': np.zeros(capacity),
                       'mult': np.zeros(capacity, dtype=np.int32), 'count': np.zeros(1, dtype=np.int32)}
        nacc = lib().dm_mh_chain(
            C.byref(self.c), C.byref(chain), C.c_int32(n_steps), C.c_uint64(seed), C.c_int(self.lanes),
            _f(z) if z is not None else None, _f(e) if e is not None else None, C.c_int(mode), C.c_int(thin),
            C.c_int(capacity), _d(samples['x']) if mode else None, _d(samples['loglkl']) if mode else None,
            samples['mult'].ctypes.data_as(_ip) if mode else None,
            samples['count'].ctypes.data_as(_ip) if mode else None)
        return nacc, samples


def philox(counter, key):
    ctr = (C.c_uint32 * 4)(*counter)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    lib().dm_philox4x32_10(ctr, k, out)
    return list(out)


def variates(seed, chain_id, step0, n_steps, n):
    z = np.zeros((n_steps, n), dtype=np.float32)
    e = np.zeros(n_steps, dtype=np.float32)
    lib().dm_variates(C.c_uint64(seed), C.c_uint32(chain_id), C.c_uint32(step0), C.c_int32(n_steps), C.c_int32(n),
                      _f(z), _f(e))
    return z, e


def schedule_next(sched: DmSchedule, chain: DmChain, accepted, best_loglkl):
    lib().dm_schedule_next(C.byref(sched), C.byref(chain), C.c_int32(accepted), C.c_double(best_loglkl))
