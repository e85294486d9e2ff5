"""Device-resident chain batches and thin wrappers over the libpb200 entry points.

torch is used for device memory and streams only; every numerical step runs in the hand-written
sm_100a kernels behind the C ABI (include/pb200.h).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import capi
from .problem import DeviceProblem


LAUNCHES = {'count': 0}     # kernels of libpb200 launched through this module (bench.py reports it)


def _count(n=1):
    LAUNCHES['count'] += n


def _stream_ptr(stream=None):
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)


class ChainBatch:
    """State of B chains on one GPU (struct pb200_chains).  Chains are enumerated
    c = repetition * n_points + point (tasks/initialise_profile_task.py:16-26)."""

    _f64 = ('loglkl', 'temperature', 'step_size', 'current_best', 'secant_prev_mult', 'secant_cur_mult',
            'secant_prev_ar')
    _i32 = ('point', 'iteration', 'status', 'secant_stage')

    def __init__(self, problem: DeviceProblem, x0, point, chain_id=None, temperature=1.0, step_size=1.0,
                 current_best=math.inf, steps_done=0):
        dev = problem.device
        self.problem = problem
        x0 = np.asarray(x0, dtype=np.float64)
        b = x0.shape[0]
        self.n_chains = b
        xp = np.zeros((b, problem.npad))
        xp[:, :problem.n] = x0[:, :problem.n]
        self.x = torch.as_tensor(xp, device=dev)
        full = lambda v, dt: torch.as_tensor(np.broadcast_to(np.asarray(v, dtype=dt), (b,)).copy(), device=dev)
        self.loglkl = torch.zeros(b, dtype=torch.float64, device=dev)
        self.temperature = full(temperature, np.float64)
        self.step_size = full(step_size, np.float64)
        self.current_best = full(current_best, np.float64)
        self.point = full(point, np.int32)
        ids = np.arange(b, dtype=np.uint32) if chain_id is None else np.asarray(chain_id, dtype=np.uint32)
        self.chain_id = torch.as_tensor(ids.view(np.int32).copy(), device=dev)
        self.steps_done = full(steps_done, np.int32)
        self.iteration = torch.zeros(b, dtype=torch.int32, device=dev)
        self.status = torch.zeros(b, dtype=torch.int32, device=dev)
        self.secant_stage = torch.zeros(b, dtype=torch.int32, device=dev)
        self.secant_prev_mult = torch.zeros(b, dtype=torch.float64, device=dev)
        self.secant_cur_mult = torch.zeros(b, dtype=torch.float64, device=dev)
        self.secant_prev_ar = torch.zeros(b, dtype=torch.float64, device=dev)
        self.c = capi.Chains(b, *(capi.ptr(getattr(self, k)) for k in (
            'x', 'loglkl', 'temperature', 'step_size', 'current_best', 'point', 'chain_id', 'steps_done',
            'iteration', 'status', 'secant_stage', 'secant_prev_mult', 'secant_cur_mult', 'secant_prev_ar')))

    def positions(self):
        return self.x[:, :self.problem.n].cpu().numpy()


class Records:
    """Per-(iteration, chain) bestfit records (struct pb200_records)."""

    def __init__(self, problem: DeviceProblem, n_chains, max_iterations, keep_positions=True):
        dev = problem.device
        k, b = max_iterations, n_chains
        self.max_iterations = k
        self.best_loglkl = torch.full((k, b), math.inf, dtype=torch.float64, device=dev)
        self.last_loglkl = torch.full((k, b), math.inf, dtype=torch.float64, device=dev)
        self.temperature = torch.zeros((k, b), dtype=torch.float64, device=dev)
        self.step_size = torch.zeros((k, b), dtype=torch.float64, device=dev)
        self.accepted = torch.full((k, b), -1, dtype=torch.int32, device=dev)
        self.best_x = torch.zeros((k, b, problem.npad), dtype=torch.float64, device=dev) if keep_positions else None
        self.last_x = torch.zeros((k, b, problem.npad), dtype=torch.float64, device=dev) if keep_positions else None
        self.c = capi.Records(k, *(capi.ptr(getattr(self, n)) for n in (
            'best_loglkl', 'last_loglkl', 'temperature', 'step_size', 'accepted', 'best_x', 'last_x')))


class Samples:
    """Chain recording buffers of mh_run (struct pb200_samples)."""

    def __init__(self, problem: DeviceProblem, n_chains, capacity, mode=capi.RECORD_ACCEPTED, thin=1):
        dev = problem.device
        self.mode, self.thin, self.capacity = mode, thin, capacity
        self.x = torch.zeros((n_chains, capacity, problem.npad), dtype=torch.float64, device=dev)
        self.loglkl = torch.zeros((n_chains, capacity), dtype=torch.float64, device=dev)
        self.mult = torch.zeros((n_chains, capacity), dtype=torch.int32, device=dev)
        self.count = torch.zeros(n_chains, dtype=torch.int32, device=dev)
        self.c = capi.Samples(mode, thin, capacity, capi.ptr(self.x), capi.ptr(self.loglkl), capi.ptr(self.mult),
                              capi.ptr(self.count))


def make_schedule(steps_per_iteration, temperature_change, step_mode=capi.STEP_ADAPTIVE_FIXED,
                  step_size_change=1.0, adaptive_interval=(0.05, 0.15), adaptive_multiplier=0.0,
                  chi2_tolerance=None):
    tol = 0.0 if chi2_tolerance is None else float(chi2_tolerance)
    return capi.Schedule(int(steps_per_iteration), int(step_mode), float(temperature_change),
                         float(step_size_change), float(adaptive_interval[0]), float(adaptive_interval[1]),
                         float(adaptive_multiplier), tol)


def loglkl_batch(problem: DeviceProblem, x, point=None, stream=None):
    """x: [M, npad] float64 device tensor -> (loglkl [M], outside [M] int32)."""
    m = x.shape[0]
    out = torch.empty(m, dtype=torch.float64, device=problem.device)
    outside = torch.empty(m, dtype=torch.int32, device=problem.device)
    _count()
    capi.check(capi.lib().pb200_loglkl_batch(C.byref(problem.c), capi.ptr(x), capi.ptr(point), m, capi.ptr(out),
                                             capi.ptr(outside), _stream_ptr(stream)))
    return out, outside


def anneal_run(problem, chains: ChainBatch, schedule, records: Records, n_iter, seed, stream=None):
    _count()
    capi.check(capi.lib().pb200_anneal_run(C.byref(problem.c), C.byref(chains.c), C.byref(schedule),
                                           C.byref(records.c), int(n_iter), C.c_uint64(seed), _stream_ptr(stream)))


def mh_run(problem, chains: ChainBatch, n_steps, seed, samples: Samples = None, stream=None):
    accepted = torch.zeros(chains.n_chains, dtype=torch.int32, device=problem.device)
    _count()
    capi.check(capi.lib().pb200_mh_run(C.byref(problem.c), C.byref(chains.c), int(n_steps), C.c_uint64(seed),
                                       C.byref(samples.c) if samples is not None else None, capi.ptr(accepted),
                                       _stream_ptr(stream)))
    return accepted


def schedule_update(chains: ChainBatch, schedule, accepted, best_loglkl, stream=None):
    _count()
    capi.check(capi.lib().pb200_schedule_update(C.byref(chains.c), C.byref(schedule), capi.ptr(accepted),
                                                capi.ptr(best_loglkl), _stream_ptr(stream)))


def profile_reduce(records: Records, n_iter, n_points, reps, stream=None):
    """-> dict(chain_best [B], chain_iter [B], point_best [P], point_rep [P], point_avg [P]) device tensors."""
    dev = records.best_loglkl.device
    b = n_points * reps
    out = {
        'chain_best': torch.empty(b, dtype=torch.float64, device=dev),
        'chain_iter': torch.empty(b, dtype=torch.int32, device=dev),
        'point_best': torch.empty(n_points, dtype=torch.float64, device=dev),
        'point_rep': torch.empty(n_points, dtype=torch.int32, device=dev),
        'point_avg': torch.empty(n_points, dtype=torch.float64, device=dev),
    }
    _count(2)
    capi.check(capi.lib().pb200_profile_reduce(
        capi.ptr(records.best_loglkl), capi.ptr(records.accepted), int(n_iter), b, int(n_points), int(reps),
        capi.ptr(out['chain_best']), capi.ptr(out['chain_iter']), capi.ptr(out['point_best']),
        capi.ptr(out['point_rep']), capi.ptr(out['point_avg']), _stream_ptr(stream)))
    return out


def weighted_moments(x, n, weights=None, stream=None):
    """x: [N, ld] float64 device tensor (first n columns used), weights [N] int32 or None.
    -> (sum_w, sum_wx [n], sum_wxx [n, n]) device tensors."""
    x2 = x.reshape(-1, x.shape[-1])
    out = torch.empty(1 + n + n * n, dtype=torch.float64, device=x.device)
    _count()
    capi.check(capi.lib().pb200_weighted_moments(capi.ptr(x2), capi.ptr(weights), x2.shape[0], x2.shape[1], int(n),
                                                 capi.ptr(out), _stream_ptr(stream)))
    return out[0], out[1:1 + n], out[1 + n:].reshape(n, n)


def covariance_from_moments(sum_w, sum_wx, sum_wxx):
    """np.cov(..., bias=True, fweights=w) from the raw moments (prospect/mcmc.py:93-97)."""
    mean = sum_wx / sum_w
    return sum_wxx / sum_w - torch.outer(mean, mean)


def rng_variates(device, seed, chain_id, step0, n_steps, n, stream=None):
    z = torch.empty((n_steps, n), dtype=torch.float32, device=device)
    e = torch.empty(n_steps, dtype=torch.float32, device=device)
    _count()
    capi.check(capi.lib().pb200_rng_variates(C.c_uint64(seed), int(chain_id), int(step0), int(n_steps), int(n),
                                             capi.ptr(z), capi.ptr(e), _stream_ptr(stream)))
    return z, e
