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

    STATE_KEYS = ('x', 'loglkl', 'temperature', 'step_size', 'current_best', 'secant_prev_mult', 'secant_cur_mult',
                  'secant_prev_ar', 'steps_done', 'iteration', 'status', 'secant_stage')

    def __init__(self, problem: DeviceProblem, x0, point, chain_id=None, temperature=1.0, step_size=1.0,
                 current_best=math.inf, steps_done=0, state=None):
        from .problem import upload_packed
        dev = problem.device
        self.problem = problem
        x0 = np.asarray(x0, dtype=np.float64)
        b = x0.shape[0]
        self.n_chains = b
        xp = np.zeros((b, problem.npad))
        xp[:, :problem.n] = x0[:, :problem.n]
        full = lambda v, dt: np.broadcast_to(np.asarray(v, dtype=dt), (b,)).copy()
        ids = np.arange(b, dtype=np.uint32) if chain_id is None else np.asarray(chain_id, dtype=np.uint32)
        host = {
            'x': xp, 'loglkl': np.zeros(b), 'temperature': full(temperature, np.float64),
            'step_size': full(step_size, np.float64), 'current_best': full(current_best, np.float64),
            'secant_prev_mult': np.zeros(b), 'secant_cur_mult': np.zeros(b), 'secant_prev_ar': np.zeros(b),
            'point': full(point, np.int32), 'chain_id': ids.view(np.int32).copy(),
            'steps_done': full(steps_done, np.int32), 'iteration': np.zeros(b, dtype=np.int32),
            'status': np.zeros(b, dtype=np.int32), 'secant_stage': np.zeros(b, dtype=np.int32),
        }
        if state is not None:                 # resume: every per-chain quantity of an earlier run (host_state())
            for key in self.STATE_KEYS:
                if key == 'x':
                    host['x'][:, :problem.n] = np.asarray(state['x'], dtype=np.float64)[:, :problem.n]
                else:
                    host[key] = np.asarray(state[key], dtype=host[key].dtype).reshape(b).copy()
        t = upload_packed(host, dev)          # one pinned staging buffer, one H2D copy
        self._storage = t.pop('_storage')
        for name, tensor in t.items():
            setattr(self, name, tensor)
        self.c = capi.Chains(b, *(capi.ptr(getattr(self, k)) for k in (
            'x', 'loglkl', 'temperature', 'step_size', 'current_best', 'point', 'chain_id', 'steps_done',
            'iteration', 'status', 'secant_stage', 'secant_prev_mult', 'secant_cur_mult', 'secant_prev_ar')))

    def host_state(self):
        """Everything needed to continue these chains later (ChainBatch(..., state=...)): host arrays per chain."""
        out = {key: getattr(self, key).cpu().numpy() for key in self.STATE_KEYS}
        out['x'] = out['x'][:, :self.problem.n]
        return out

    def positions(self):
        return self.x[:, :self.problem.n].cpu().numpy()


class Records:
    """Per-(iteration, chain) bestfit records (struct pb200_records).

    One contiguous FP64 block per iteration -- [best_loglkl | last_loglkl | temperature | step_size | accepted
    (int32) | best_x] -- so a rank can all-gather `buf[k]` at an iteration boundary without packing; the last
    positions (only needed locally) live in a second buffer with the same iteration stride.
    `n_alloc` >= n_chains pads the block to a rank-independent size."""

    def __init__(self, problem: DeviceProblem, n_chains, max_iterations, keep_positions=True, n_alloc=None):
        dev = problem.device
        k, b, npad = max_iterations, int(n_alloc or n_chains), problem.npad
        self.max_iterations, self.n_chains, self.n_alloc, self.npad = k, n_chains, b, npad
        acc_d = (b + 1) // 2
        self.offsets = {'best_loglkl': 0, 'last_loglkl': b, 'temperature': 2 * b, 'step_size': 3 * b,
                        'accepted': 4 * b, 'best_x': 4 * b + acc_d}
        self.stride = max(4 * b + acc_d + (b * npad if keep_positions else 0), b * npad if keep_positions else 0)
        self.scalar_len = 4 * b + acc_d          # the per-iteration prefix ranks all-gather at every boundary
        self.buf = torch.zeros((k, self.stride), dtype=torch.float64, device=dev)
        self.buf_last = torch.zeros((k, self.stride), dtype=torch.float64, device=dev) if keep_positions else None
        o = self.offsets
        self.best_loglkl = self.buf[:, o['best_loglkl']:o['best_loglkl'] + n_chains]
        self.last_loglkl = self.buf[:, o['last_loglkl']:o['last_loglkl'] + n_chains]
        self.temperature = self.buf[:, o['temperature']:o['temperature'] + n_chains]
        self.step_size = self.buf[:, o['step_size']:o['step_size'] + n_chains]
        self.accepted = self.buf[:, o['accepted']:o['accepted'] + acc_d].view(torch.int32)[:, :n_chains]
        self.best_x = self.last_x = None
        if keep_positions:
            self.best_x = self.buf[:, o['best_x']:o['best_x'] + b * npad].unflatten(1, (b, npad))[:, :n_chains]
            self.last_x = self.buf_last[:, :b * npad].unflatten(1, (b, npad))[:, :n_chains]
        self.best_loglkl.fill_(math.inf)
        self.last_loglkl.fill_(math.inf)
        self.accepted.fill_(-1)
        base = self.buf.data_ptr()
        addr = lambda name: C.c_void_p(base + 8 * o[name])
        self.c = capi.Records(k, self.stride, addr('best_loglkl'), addr('last_loglkl'), addr('temperature'),
                              addr('step_size'), addr('accepted'), addr('best_x') if keep_positions else None,
                              C.c_void_p(self.buf_last.data_ptr()) if keep_positions else None)

    @staticmethod
    def decode(block, n_chains, n_alloc, npad, keep_positions=True):
        """Split gathered blocks [..., stride or scalar_len] (numpy) back into named arrays of the first n_chains
        chains (positions only if the block carries them)."""
        b, acc_d = n_alloc, (n_alloc + 1) // 2
        out = {'best_loglkl': block[..., 0:n_chains], 'last_loglkl': block[..., b:b + n_chains],
               'temperature': block[..., 2 * b:2 * b + n_chains], 'step_size': block[..., 3 * b:3 * b + n_chains]}
        acc = np.ascontiguousarray(block[..., 4 * b:4 * b + acc_d]).view(np.int32)
        out['accepted'] = acc[..., :n_chains]
        if keep_positions and block.shape[-1] >= 4 * b + acc_d + b * npad:
            o = 4 * b + acc_d
            bx = block[..., o:o + b * npad]
            out['best_x'] = bx.reshape(bx.shape[:-1] + (b, npad))[..., :n_chains, :]
        return out


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


def choose_lanes(npad, n_chains):
    """Lanes per chain the library picks for this batch (see pb200_choose_lanes)."""
    return int(capi.lib().pb200_choose_lanes(int(npad), int(n_chains)))


def uses_producer_warps(lanes, n_chains):
    """Whether anneal_run serves this batch with the producer-warp kernel (pb200_uses_producer_warps)."""
    return bool(capi.lib().pb200_uses_producer_warps(int(lanes), int(n_chains)))


def anneal_run(problem, chains: ChainBatch, schedule, records: Records, n_iter, seed, stream=None, lanes=0):
    _count()
    capi.check(capi.lib().pb200_anneal_run(C.byref(problem.c), C.byref(chains.c), C.byref(schedule),
                                           C.byref(records.c), int(n_iter), C.c_uint64(seed), int(lanes),
                                           _stream_ptr(stream)))


def anneal_run_signalled(problem, chains: ChainBatch, schedule, records: Records, n_iter, seed, progress, stream=None,
                         lanes=0):
    """anneal_run whose kernel counts finished iterations in `progress` (zeroed int32 device tensor [n_iter]).
    Returns the value every counter reaches (see pb200_anneal_run_signalled)."""
    target = C.c_int32(0)
    _count()
    capi.check(capi.lib().pb200_anneal_run_signalled(C.byref(problem.c), C.byref(chains.c), C.byref(schedule),
                                                     C.byref(records.c), int(n_iter), C.c_uint64(seed), int(lanes),
                                                     capi.ptr(progress), C.byref(target), _stream_ptr(stream)))
    return target.value


def make_peers(rank, n_alloc, rows, rank_stride, buffer_ptrs, progress_ptrs):
    """struct pb200_peers from the device addresses of every rank's exchange buffer / counters (peer memory)."""
    n = len(buffer_ptrs)
    if n > capi.MAX_PEERS:
        raise capi.Pb200Error(f'at most {capi.MAX_PEERS} ranks can share exchange buffers')
    peers = capi.Peers()
    peers.n_peers, peers.rank, peers.n_alloc, peers.rows = n, int(rank), int(n_alloc), int(rows)
    peers.row_stride, peers.rank_stride = int(n * rank_stride), int(rank_stride)
    for q in range(n):
        peers.buffer[q] = int(buffer_ptrs[q])
        peers.progress[q] = int(progress_ptrs[q])
    return peers


def anneal_run_shared(problem, chains: ChainBatch, schedule, records: Records, n_iter, seed, peers, progress=None,
                      stream=None, lanes=0):
    """anneal_run whose kernel also stores every iteration's per-chain scalars into all ranks' exchange buffers and
    counts arrivals there (pb200_anneal_run_shared).  Returns this rank's number of chain warps."""
    target = C.c_int32(0)
    _count()
    capi.check(capi.lib().pb200_anneal_run_shared(C.byref(problem.c), C.byref(chains.c), C.byref(schedule),
                                                  C.byref(records.c), int(n_iter), C.c_uint64(seed), int(lanes),
                                                  C.byref(peers), capi.ptr(progress), C.byref(target),
                                                  _stream_ptr(stream)))
    return target.value


def stream_wait_geq(stream, counter, index, value):
    """Make `stream` wait until counter[index] >= value (a device-side wait; the host does not block)."""
    capi.check(capi.lib().pb200_stream_wait_geq(_stream_ptr(stream), C.c_void_p(counter.data_ptr() + 4 * int(index)),
                                                int(value)))


def exchange_arrivals(npad, n_chains, lanes=0):
    """Arrivals a rank with `n_chains` chains adds to the counters of the signalled / shared launches."""
    n = int(capi.lib().pb200_exchange_arrivals(int(npad), int(n_chains), int(lanes)))
    if n < 0:
        capi.check(1)
    return n


def launch_status(stream=None):
    """Wait for the stream and raise Pb200Error if a launch recorded a device-side protocol fault (a producer-warp
    ring or a TMA / tensor-core pipeline stage that timed out)."""
    capi.check(capi.lib().pb200_launch_status(_stream_ptr(stream)))


def mh_run(problem, chains: ChainBatch, n_steps, seed, samples: Samples = None, stream=None, lanes=0,
           exact_box=False):
    """`exact_box`: test the prior box at every likelihood-accepted step (PB200_EXACT_BOX) instead of deferring it
    behind the whitened safe radius -- same decisions, faster for wide steps."""
    accepted = torch.zeros(chains.n_chains, dtype=torch.int32, device=problem.device)
    smp = samples.c if samples is not None else None
    if exact_box:
        smp = capi.Samples(capi.EXACT_BOX, 1, 0, None, None, None, None) if smp is None else capi.Samples(
            smp.mode | capi.EXACT_BOX, smp.thin, smp.capacity, smp.x, smp.loglkl, smp.mult, smp.count)
    _count()
    capi.check(capi.lib().pb200_mh_run(C.byref(problem.c), C.byref(chains.c), int(n_steps), C.c_uint64(seed),
                                       C.byref(smp) if smp is not None else None, capi.ptr(accepted),
                                       int(lanes), _stream_ptr(stream)))
    return accepted


def hosted_propose(hosted, chains: ChainBatch, seed, out_prop, out_outside, stream=None):
    """Proposals x' = x + step * F z of every active chain + their prior-box test (pb200_hosted_propose)."""
    _count()
    capi.check(capi.lib().pb200_hosted_propose(C.byref(hosted.c), C.byref(chains.c), C.c_uint64(seed),
                                               capi.ptr(out_prop), capi.ptr(out_outside), _stream_ptr(stream)))


def hosted_accept(hosted, chains: ChainBatch, seed, prop, prop_loglkl, outside, accepted_flag=None, n_accepted=None,
                  best_loglkl=None, best_x=None, stream=None):
    """Accept / reject with host-evaluated likelihood values and advance the chains (pb200_hosted_accept)."""
    _count()
    capi.check(capi.lib().pb200_hosted_accept(C.byref(hosted.c), C.byref(chains.c), C.c_uint64(seed), capi.ptr(prop),
                                              capi.ptr(prop_loglkl), capi.ptr(outside), capi.ptr(accepted_flag),
                                              capi.ptr(n_accepted), capi.ptr(best_loglkl), capi.ptr(best_x),
                                              _stream_ptr(stream)))


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
        capi.ptr(records.best_loglkl), capi.ptr(records.accepted), int(records.stride), 2 * int(records.stride),
        int(n_iter), b, int(n_points), int(reps),
        capi.ptr(out['chain_best']), capi.ptr(out['chain_iter']), capi.ptr(out['point_best']),
        capi.ptr(out['point_rep']), capi.ptr(out['point_avg']), _stream_ptr(stream)))
    return out


def chain_minimum(records: Records, n_iter, stream=None):
    """Per chain: first minimum over the iterations that ran -> dict(chain_best [B], chain_iter [B]) device tensors
    (the first pass of pb200_profile_reduce, with every chain as its own 'point')."""
    dev = records.best_loglkl.device
    b = records.n_chains
    out = {'chain_best': torch.empty(b, dtype=torch.float64, device=dev),
           'chain_iter': torch.empty(b, dtype=torch.int32, device=dev)}
    scratch = {'pb': torch.empty(b, dtype=torch.float64, device=dev), 'pr': torch.empty(b, dtype=torch.int32, device=dev),
               'pa': torch.empty(b, dtype=torch.float64, device=dev)}
    _count(2)
    capi.check(capi.lib().pb200_profile_reduce(
        capi.ptr(records.best_loglkl), capi.ptr(records.accepted), int(records.stride), 2 * int(records.stride),
        int(n_iter), b, b, 1, capi.ptr(out['chain_best']), capi.ptr(out['chain_iter']), capi.ptr(scratch['pb']),
        capi.ptr(scratch['pr']), capi.ptr(scratch['pa']), _stream_ptr(stream)))
    return out


def weighted_moments(x, n, weights=None, stream=None):
    """x: [N, ld] float64 device tensor (first n columns used), weights [N] int32 or None.
    -> (sum_w, sum_wx [n], sum_wxx [n, n]) device tensors."""
    x2 = x.reshape(-1, x.shape[-1])
    out = torch.empty(1 + n + n * n, dtype=torch.float64, device=x.device)
    _count(1 if x2.shape[0] <= 32 else 2)          # partial sums per CTA + their ordered reduction (no atomics)
    capi.check(capi.lib().pb200_weighted_moments(capi.ptr(x2), capi.ptr(weights), x2.shape[0], x2.shape[1], int(n),
                                                 capi.ptr(out), _stream_ptr(stream)))
    return out[0], out[1:1 + n], out[1 + n:].reshape(n, n)


def chain_moments(samples, n, stream=None):
    """Per-chain weighted first moments of recorded chains: (sum_w [B], sum_wx [B, n], sum_wx / sqrt(sum_w) [B, n])."""
    b = samples.x.shape[0]
    dev = samples.x.device
    sw = torch.empty(b, dtype=torch.float64, device=dev)
    swx = torch.empty((b, n), dtype=torch.float64, device=dev)
    scaled = torch.empty((b, n), dtype=torch.float64, device=dev)
    _count()
    capi.check(capi.lib().pb200_chain_moments(capi.ptr(samples.x), capi.ptr(samples.mult), capi.ptr(samples.count), b,
                                              samples.capacity, samples.x.shape[-1], int(n), capi.ptr(sw),
                                              capi.ptr(swx), capi.ptr(scaled), _stream_ptr(stream)))
    return sw, swx, scaled


def chain_statistics(samples, n, stream=None):
    """Flat FP64 device vector [sum_w, sum_wx (n), sum_wxx (n*n), sum_c swx_c swx_c^T / sw_c (n*n), n_chains] of the
    recorded chains of this rank: plain sums, so ranks all-reduce it (run.convergence_from_statistics)."""
    b = samples.x.shape[0]
    out = torch.zeros(2 + n + 2 * n * n, dtype=torch.float64, device=samples.x.device)
    if b == 0:
        return out
    sw, swx, scaled = chain_moments(samples, n, stream)
    tot_w, tot_wx, tot_wxx = weighted_moments(samples.x, n, samples.mult.reshape(-1), stream)
    _, _, syy = weighted_moments(scaled, n, None, stream)
    out[0] = tot_w
    out[1:1 + n] = tot_wx
    out[1 + n:1 + n + n * n] = tot_wxx.reshape(-1)
    out[1 + n + n * n:1 + n + 2 * n * n] = syy.reshape(-1)
    out[-1] = (sw > 0).sum()             # chains with at least one recorded row
    return out


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
