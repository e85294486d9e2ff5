#!/usr/bin/env python
"""bench.py -- chain-steps/s and time-to-profile of the annealing / Metropolis-Hastings hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...] [--no-subrecords]

Headline (`value`, BASELINE.json configs[1]): the reference's example_global_optimisation_toy.yaml on the 30-D toy
(reference recipe, seed 0) with `repetitions: 4096` PER GPU -- 4096 chains x 20 iterations x 250 steps = 2.048e7
chain-steps per job, exponential temperature 0.1 -> 1e-3, adaptive step size (interval [0.19, 0.21], multiplier 0.3,
initial 0.1), start point / proposal covariance from the reference's own bootstrap MCMC (tests/golden/toy30_mcmc.npz).
One "step" = JOBS_PER_STEP such jobs back to back, each on its own device-resident chain batch (so the timed region of
the default run is > 1 s and the clock sampler sees load).

One JSON line is printed by rank 0:
  value        whole-job chain-steps/s with chain state resident in HBM (CUDA events, max over ranks)
  e2e          the same job through the host API (SimulatedAnnealing built from HOST arrays: H2D of the starts and the
               problem matrices, launches, D2H of the records into pinned memory), wall clock; independent jobs are
               double-buffered (SimulatedAnnealing.results_async), every job's copies are inside the timed region
  roofline     fused annealing kernel vs the streaming-formulation HBM bound (272 B per
               chain-step at n_pad = 32), + the measured instruction-pipe bound -- DESIGN.md explains both
  cpu_baseline the UNMODIFIED reference (oracle/_ref: its own Scheduler in `mode: threaded`, one process per host core)
               on a bounded sample of the same workload
  k3_strong    BASELINE configs[2]: 30-D profile, 32 values x 256 repetitions = 8192 chains FIXED, sharded by profile
               point over the N ranks (strong scaling)
  k4           BASELINE configs[3]: 256-D synthetic SPD Gaussian profile, 8 points x 128 repetitions per GPU (wide path:
               steps kernel + tcgen05 position GEMM + FP64 anchor GEMM per iteration)
  k5           BASELINE configs[4]: jobtype mcmc, 65 536 chains in total at the reference-optimal wide step
               2.4 / sqrt(d), without recording and recording every 100th step
  time_to_profile   yaml -> run() -> profile/x2.txt on disk for the 30-D profile job (32 values per GPU x 256 reps)
`--impl reference` runs the reference arm alone: the unmodified reference through its own scheduler on all host cores.
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC, UNIT = 'chain-steps/s', 'chain-steps/s'
D = 30
JOBS_PER_STEP = 32


def workload_profile_kw(name, n_gpus, reps_per_gpu=4096):
    base = dict(temperature_range=[0.1, 0.001], max_iterations=20, step_size_schedule='adaptive',
                step_size_adaptive_interval=[0.19, 0.21], step_size_adaptive_multiplier=0.3,
                step_size_adaptive_initial=0.1, steps_per_iteration=250)
    if name == 'global30':
        return 'global_optimisation', dict(base, repetitions=reps_per_gpu * n_gpus), 1, reps_per_gpu * n_gpus
    if name == 'profile30':            # weak: 32 profile values per GPU
        p = 32 * n_gpus
        return 'profile', dict(base, repetitions=256, values=f'np.linspace(-1.0, 2.0, {p})'), p, 256
    if name == 'profile30_strong':     # BASELINE configs[2]: 32 x 256 fixed
        return 'profile', dict(base, repetitions=256, values='np.linspace(-1.0, 2.0, 32)'), 32, 256
    if name == 'profile256':           # BASELINE configs[3]: 64 points x 128 repetitions on 8 GPUs = 8 points per GPU
        p = 8 * n_gpus
        return 'profile', dict(base, repetitions=128, values=None, step_size_adaptive_initial=0.15), p, 128
    raise SystemExit(f'unknown workload {name}')


def spd_inputs(folder, n_points):
    """256-D synthetic SPD target (not from the reference recipe, which is indefinite at this size): start near the
    mean, proposal = closed-form conditional covariance, profile values within +-2 sigma of the profiled parameter."""
    from prospect_public_b200.toy import conditional_covariance, spd_target, write_toy_inputs
    mu, cov = spd_target(256, seed=4)
    paths = write_toy_inputs(folder, mu, cov, None)
    sig = float(np.sqrt(cov[1, 1]))
    extra = {'values': (mu[1] + sig * np.linspace(-2.0, 2.0, n_points)).tolist(), 'start_from_mcmc': None,
             'start_from_position': (np.delete(mu, 1) + 0.01).tolist(),
             'start_from_covmat': conditional_covariance(cov).tolist()}
    return paths, {'means': mu, 'covmat': cov}, None, extra


def toy_inputs(folder):
    from prospect_public_b200.toy import write_toy_inputs
    k = np.load(os.path.join(ROOT, 'tests', 'golden', f'toy{D}_kernel.npz'))
    m = np.load(os.path.join(ROOT, 'tests', 'golden', f'toy{D}_mcmc.npz'))
    chains = [m[f'chain_{i}'] for i in range(4)]
    return write_toy_inputs(folder, k['means'], k['covmat'], chains), k, np.concatenate(chains, axis=0)


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of this rank's GPU during the timed region (NVML)."""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap',
               0x80: 'hw_power_brake_slowdown'}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag:
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(
                    pynvml, 'nvmlDeviceGetCurrentClocksEventReasons') else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.01)
        except Exception as exc:  # NVML missing: report it rather than fail the bench
            self.reasons.add(f'nvml_unavailable:{type(exc).__name__}')

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'samples': len(s),
                'reasons': sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------------------------
# the CPU side: the UNMODIFIED reference through its own scheduler (oracle/_ref), or -- only if that copy is missing
# -- the pinned port (oracle/reference_port.py)
# ------------------------------------------------------------------------------------------------------------------
def reference_job_yaml(workload, tmp, tag, cores, kw, chains_per_core=1):
    """The yaml of a bounded sample of `workload` for the reference: `chains_per_core` chains per host core (global30) or
    the 32 profile values x 1 repetition (profile30*), ALL 20 iterations."""
    from prospect_public_b200.toy import annealing_yaml
    paths, _, _ = toy_inputs(os.path.join(tmp, 'in_' + tag))
    prof = {k: v for k, v in kw.items() if k not in ('repetitions', 'values')}
    prof.update(plot_profile=False, plot_schedule=False)
    if workload == 'global30':
        jobtype, prof['repetitions'] = 'global_optimisation', cores * chains_per_core
        sample = f'{cores * chains_per_core} repetitions ({chains_per_core} per host core) x 20 iterations x 250 steps'
    else:
        jobtype, prof['repetitions'], prof['values'] = 'profile', 1, 'np.linspace(-1.0, 2.0, 32)'
        sample = '32 profile values x 1 repetition x 20 iterations x 250 steps'
    y = annealing_yaml(paths, os.path.join(tmp, 'ref_' + tag), jobtype=jobtype, write=True, **prof)
    y['io']['snapshot_interval'] = 1.0e9
    return y, sample


def reference_rate(workload, tmp, tag, kw, cores, analyse=False, chains_per_core=1):
    """(chain-steps/s, seconds, kind, sample, extra) of one bounded reference job on `cores` processes."""
    from oracle import ref_harness as rh
    if rh.available():
        y, sample = reference_job_yaml(workload, tmp, tag, cores, kw, chains_per_core)
        out = rh.run_reference(y, 20, mode='threaded' if cores > 1 else 'serial', num_processes=cores, analyse=analyse)
        extra = {'seconds_to_profile': out.get('seconds_to_profile'), 'tasks': out['n_optimise_tasks']}
        return out['chain_steps'] / out['seconds'], out['seconds'], 'reference', sample + \
            " through the unmodified reference's Scheduler (mode: threaded, one process per core)", extra
    # fallback: the pinned numpy port, one chain per core
    import multiprocessing as mp
    k = np.load(os.path.join(ROOT, 'tests', 'golden', f'toy{D}_kernel.npz'))
    m = np.load(os.path.join(ROOT, 'tests', 'golden', f'toy{D}_mcmc.npz'))
    rows = np.concatenate([m[f'chain_{i}'] for i in range(4)], axis=0)
    jobtype = 'global_optimisation' if workload == 'global30' else 'profile'
    t0 = time.perf_counter()
    with mp.get_context('spawn').Pool(cores) as pool:
        done = sum(pool.map(_port_worker, [(k['means'], k['covmat'], rows, jobtype, 1000 + c, 20, 250)
                                           for c in range(cores)]))
    dt = time.perf_counter() - t0
    return done / dt, dt, 'port', f'{cores} chains x 20 iterations x 250 steps through oracle/reference_port.py ' \
                                  '(oracle/_ref is missing)', {}


def _port_worker(args):
    k_means, k_cov, rows, jobtype, seed, n_iter, steps = args
    from oracle import reference_port as rp
    sched = rp.Schedule(steps_per_iteration=steps, max_iterations=20)
    np.random.seed(seed)
    target = rp.GaussianTarget(k_means, k_cov)
    if jobtype == 'global_optimisation':
        rp.run_annealing_job(target, rows, sched, [], 1, None, n_iterations=n_iter)
    else:
        rp.run_annealing_job(target, rows, sched, np.linspace(-1.0, 2.0, 32)[:1], 1, 1, n_iterations=n_iter)
    return n_iter * steps


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on all host cores.  A step = one
    bounded job (one chain per core, the full 20-iteration schedule) through its Scheduler."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    jobtype, kw, n_points, reps = workload_profile_kw(args.workload, args.gpus, args.reps_per_gpu)
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp(prefix='pb200_ref_')
    for w in range(args.warmup):
        reference_rate(args.workload, tmp, f'w{w}', kw, cores)
    total_steps, total_s, kind, sample = 0.0, 0.0, 'reference', ''
    for s in range(args.steps):
        rate, secs, kind, sample, _ = reference_rate(args.workload, tmp, f's{s}', kw, cores)
        total_steps += rate * secs
        total_s += secs
    value = total_steps / total_s
    ttp = None
    if kind == 'reference':           # time-to-profile of the reference on a stated subsample of K3
        _, _, _, ttp_sample, extra = reference_rate('profile30_strong', tmp, 'ttp', workload_profile_kw(
            'profile30_strong', 1)[1], cores, analyse=True)
        ttp = {'seconds': extra['seconds_to_profile'], 'job': ttp_sample + ' (1/256 of the K3 repetitions), yaml read to '
               'profile/x2.txt written by the reference\'s own AnalyseProfileTask'}
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total_s / max(args.steps, 1),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'{args.workload}: 30-D Gaussian toy (reference recipe, seed 0), {jobtype}; the CPU '
                                   f'arm runs a bounded sample per step: {sample}'},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0, 'time_to_profile': ttp}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
class Harness:
    """Device-resident annealing jobs of one workload on this rank + the timing loop around them."""

    def __init__(self, args, workload, world, rank, dev, n_jobs):
        from prospect_public_b200.config import Configuration
        from prospect_public_b200.initialise import initialise_optimiser
        from prospect_public_b200.kernels import initialise_kernel
        from prospect_public_b200.optimiser import (SimulatedAnnealing, get_initial_step_size_change,
                                                    get_temperature_change)
        from prospect_public_b200.toy import annealing_yaml
        self.workload, self.world, self.dev = workload, world, dev
        self.jobtype, kw, self.n_points, self.reps = workload_profile_kw(workload, world, args.reps_per_gpu)
        self.tmp = tempfile.mkdtemp(prefix=f'pb200_bench_r{rank}_')
        if workload == 'profile256':
            paths, self.kgold, self.rows, extra = spd_inputs(self.tmp, self.n_points)
            kw.update(extra)
        else:
            paths, self.kgold, self.rows = toy_inputs(self.tmp)
        self.kw = kw
        cfg = Configuration(annealing_yaml(paths, os.path.join(self.tmp, 'out'), jobtype=self.jobtype, write=False, **kw))
        self.cfg = cfg
        self.kernel = initialise_kernel(cfg.kernel, None, rank, dev)
        starts, values = initialise_optimiser(cfg, self.kernel)
        self.n_iter, self.steps = int(cfg.profile.max_iterations), int(cfg.profile.steps_per_iteration)
        self.chain_steps_per_job = self.n_points * self.reps * self.n_iter * self.steps          # global

        def settings(seed):
            return {'current_best_loglkl': starts.initial_loglkl, 'initial_position': starts.positions,
                    'covmat': starts.covmats, 'temperature': cfg.profile.temperature_range[0],
                    'temperature_change': get_temperature_change(cfg), 'step_size': cfg.profile.step_size_range[0],
                    'step_size_change': get_initial_step_size_change(cfg), 'iteration_number': 0,
                    'repetitions': self.reps, 'fixed_param_val': values if self.jobtype == 'profile' else None,
                    'max_rows': self.n_iter, 'seed': seed}
        self.settings = settings
        self.new_job = lambda seed: SimulatedAnnealing(cfg, self.kernel, settings(seed), keep_positions=True)
        self.jobs = [self.new_job(1234 + j) for j in range(n_jobs)]
        self.b_local = self.jobs[0].n_chains
        keys = ('x', 'temperature', 'step_size', 'current_best', 'steps_done', 'iteration', 'status', 'secant_stage')
        self.pristine = {k: getattr(self.jobs[0].batch, k).clone() for k in keys}

    def reset(self, step):
        for j, sa in enumerate(self.jobs):
            for k, v in self.pristine.items():
                getattr(sa.batch, k).copy_(v)
            sa.iterations_done = 0
            sa.step_counter = 0
            sa.seed = 1234 + 1000 * step + j
            sa.records.accepted.fill_(-1)

    def run_jobs(self):
        for sa in self.jobs:
            sa.run_schedule(self.n_iter)
            sa.reduce_over_iterations()


def timed_steps(h, steps, warmup, flush, barrier, dist, torch, engine, time_kernels=False):
    """W untimed + K timed steps (one step = every job of the harness).  -> dict(dev_ms max over ranks, kernel_ms, ...)"""
    for w in range(warmup):
        h.reset(w)
        flush.add_(1.0)
        h.run_jobs()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    kev = []
    launches0 = engine.LAUNCHES['count']
    barrier()
    t_wall0 = time.perf_counter()
    for s in range(steps):
        h.reset(warmup + s)
        flush.add_(1.0)
        ev[s][0].record()
        if time_kernels:                       # single GPU: bracket the annealing launches separately
            for sa in h.jobs:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                sa.run_schedule(h.n_iter)
                b.record()
                kev.append((a, b))
                sa.reduce_over_iterations()
        else:
            h.run_jobs()
        ev[s][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=h.dev)
    if h.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    k_ms = sum(a.elapsed_time(b) for a, b in kev) / len(kev) if kev else None
    return {'dev_ms': float(t.item()), 'kernel_ms_per_job': k_ms, 'wall_s': t_wall,
            'launches': engine.LAUNCHES['count'] - launches0}


def e2e_rate(h, n_steps, n_warm, barrier, dist, torch):
    """Host arrays in -> host arrays out, every job: (chain-steps/s, h2d bytes, d2h bytes per job).  Jobs are
    independent, so two are kept in flight through the package's own API (SimulatedAnnealing.results_async): job k + 1
    is built from its host arrays, uploaded and launched while job k runs and its records travel back to pinned host
    memory on the copy stream; every job's H2D and D2H are inside the timed region."""
    h2d = d2h = 0
    pending, out = None, None
    for s in range(n_warm + n_steps):
        if s == n_warm:                         # the first calls pay one-off pinned-buffer / allocator set-up
            if pending is not None:
                out = pending.get()
                pending = None
            barrier()
            t0 = time.perf_counter()
        sa2 = h.new_job(9000 + s)
        sa2.run_schedule(h.n_iter)
        red = sa2.reduce_over_iterations()
        nxt = sa2.results_async(h.n_iter - 1, red)
        if pending is not None:
            out = pending.get()                 # bestfit dict + reduced records of the previous job, on the host
        pending = nxt
        if s == 0:
            h2d = sum(t_.numel() * t_.element_size() for t_ in sa2.problem.t.values())
            h2d += sum(getattr(sa2.batch, k).numel() * getattr(sa2.batch, k).element_size()
                       for k in ('x', 'temperature', 'step_size', 'current_best', 'point', 'chain_id', 'steps_done'))
            d2h = nxt.d2h_bytes
    out = pending.get()
    assert out[0]['loglkl'].shape[0] == sa2.n_chains
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=h.dev)
    if h.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return h.chain_steps_per_job * n_steps / float(t.item()), int(h2d), int(d2h)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


def stream_roofline(npad, chain_steps, kernel_ms, kernel, traffic_from=None):
    """Fraction of the streaming-formulation HBM bound: 2 * 4 * npad + 16 bytes per chain-step."""
    peak, src = hbm_peak()
    bytes_per_step = 2 * 4 * npad + 16
    achieved = bytes_per_step * chain_steps / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    if traffic_from:
        try:
            with open(os.path.join(ROOT, 'profiles', traffic_from)) as f:
                met = json.load(f)['metrics']
            unit = {'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Gbyte': 1e9}
            traffic = sum(float(met[k].split()[0]) * unit[met[k].split()[1]]
                          for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
            traffic_src = f'profiles/{traffic_from} (one ncu --set full capture of this kernel; not measured in this run)'
        except Exception:
            pass
    return {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
            'traffic': traffic, 'traffic_source': traffic_src, 'kernel': kernel, 'kernel_ms': kernel_ms,
            'peak_source': src, 'algorithmic_bytes_per_chain_step': bytes_per_step,
            'note': 'streaming-formulation bound; the kernels keep chain state in registers, so DRAM traffic is ~0 and '
                    'the binding resource is instruction issue / the FMA pipe (DESIGN.md section 3)'}


def k5_record(world, rank, dev, barrier, dist, torch, engine):
    """BASELINE configs[4]: jobtype mcmc on the 30-D toy, 65 536 chains in total (sharded over the ranks), T = 1,
    proposal = stationary covariance at the wide step 2.4 / sqrt(d); 2000 steps per round."""
    from prospect_public_b200.config import Configuration
    from prospect_public_b200.kernels import initialise_kernel
    from prospect_public_b200.mcmc import initialise_mcmc
    from prospect_public_b200.toy import mcmc_yaml
    tmp = tempfile.mkdtemp(prefix=f'pb200_k5_r{rank}_')
    paths, kgold, _ = toy_inputs(tmp)
    prec = np.linalg.inv(np.asarray(kgold['covmat'], dtype=np.float64))
    cov = np.linalg.inv(0.5 * (prec + prec.T))             # stationary covariance of the free chain = the proposal
    total, rounds_steps = 65536, 2000
    lo, hi = total * rank // world, total * (rank + 1) // world
    out = {'workload': f'jobtype mcmc, 30-D toy, {total} chains in total ({hi - lo} on this GPU), T = 1, proposal = '
                       f'stationary covariance, step 2.4/sqrt(30), {rounds_steps} steps per round', 'scaling': 'strong'}
    for key, record, thin in (('no_recording', None, 1), ('thinned_100', 'thinned', 100)):
        cfg = Configuration(mcmc_yaml(paths, os.path.join(tmp, 'o'), n_chains=total, steps=rounds_steps, write=False,
                                      start_from_covmat=cov.tolist(), step_size=float(2.4 / np.sqrt(D))))
        kernel = initialise_kernel(cfg.kernel, None, rank, dev)
        s = initialise_mcmc(cfg.mcmc, kernel, seed=1, record=record, thin=thin, n_chains=hi - lo,
                            chain_ids=np.arange(lo, hi, dtype=np.uint32))
        s.run_steps(rounds_steps)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_rounds = 3
        a.record()
        for _ in range(n_rounds):
            s.run_steps(rounds_steps)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / n_rounds
        rate = total * rounds_steps / (ms * 1e-3)
        out[key] = {'value': rate, 'unit': UNIT, 'ms_per_round': ms,
                    'acceptance_rate': float(s.accepted.sum()) / ((hi - lo) * rounds_steps * (n_rounds + 1))}
        if key == 'no_recording' and world == 1:
            out['roofline'] = stream_roofline(32, total * rounds_steps, ms, 'mh_kernel (8 lanes per chain; prior box: in-block test, TF32 mma.sync filter + exact FP32 mat-vec)')
        del s
    out['value'] = out['no_recording']['value']
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='global30', choices=['global30', 'profile30', 'profile30_strong', 'profile256'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-subrecords', action='store_true', help='only the headline workload (experiments)')
    ap.add_argument('--jobs-per-step', type=int, default=None)
    ap.add_argument('--reps-per-gpu', type=int, default=4096, help='global30 only: chains per GPU (experiments)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from prospect_public_b200 import capi, engine
    from prospect_public_b200.distributed import init_from_env

    capi.lib()                                   # fail loudly if the CUDA extension is missing
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world != args.gpus:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run')
    rank, world, active = init_from_env('nccl') if world > 1 else (0, 1, False)
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = f'cuda:{local}'

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)    # 192 MiB > 126 MB L2
    headline = args.workload
    jobs = args.jobs_per_step or (JOBS_PER_STEP if headline == 'global30' else 4)

    # ---------------- headline: device-resident arm ----------------
    h = Harness(args, headline, world, rank, dev, jobs)
    sampler = ClockSampler(local)
    sampler.start()
    res = timed_steps(h, args.steps, args.warmup, flush, barrier, dist, torch, engine, time_kernels=world == 1)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    chain_steps_per_step = h.chain_steps_per_job * jobs
    value = chain_steps_per_step * args.steps / (res['dev_ms'] * 1e-3)
    sa0 = h.jobs[0]
    lanes = engine.choose_lanes(sa0.problem.npad, h.b_local)
    wide = sa0.problem.npad >= 128
    kernel_name = ('wide path: wide_steps_kernel + wide_materialise_mma_kernel (tcgen05) + wide_anchor_kernel per iteration'
                   if wide else ('anneal_ws_kernel (chain warps + producer warps)' if engine.uses_producer_warps(
                       lanes, h.b_local) else 'anneal_kernel') + f' (lanes per chain {lanes}, {h.n_iter} iterations fused '
                   'in one launch)')
    roofline = None
    if world == 1:
        roofline = stream_roofline(sa0.problem.npad, h.b_local * h.n_iter * h.steps, res['kernel_ms_per_job'], kernel_name,
                                   'r02_global30_lanes8.json' if headline == 'global30' else None)
        if sa0.problem.npad == 32:
            # measured instruction-pipe bound of this kernel (ncu opcode mix x the per-scheduler
            # throughputs of scripts/pipe_probe.cu -> ~50 FMA-pipe cycles per chain-step) at 1965 MHz
            pipe_peak = 148 * 4 * 1.965e9 / 50.0
            rate = h.b_local * h.n_iter * h.steps / (res['kernel_ms_per_job'] * 1e-3)
            roofline['pipe_bound'] = {'fma_pipe_cycles_per_chain_step': 50, 'peak_chain_steps_per_s': pipe_peak,
                                      'achieved_chain_steps_per_s': rate, 'frac': rate / pipe_peak}
    e2e_steps = max(8, min(4 * args.steps, 64))
    e2e_value, h2d, d2h = e2e_rate(h, e2e_steps, 2, barrier, dist, torch)
    fused = getattr(sa0, 'fused_exchange', False)
    del h
    torch.cuda.empty_cache()

    # ---------------- sub-records: the other configs BASELINE.json names ----------------
    sub = {}
    if not args.no_subrecords and headline == 'global30':
        for key, wl, n_jobs, n_steps in (('k3_strong', 'profile30_strong', 4, 5), ('k4', 'profile256', 2, 5)):
            if wl == 'profile30_strong' and 32 % world != 0:
                continue
            hh = Harness(args, wl, world, rank, dev, n_jobs)
            r = timed_steps(hh, n_steps, 3, flush, barrier, dist, torch, engine, time_kernels=world == 1)
            v = hh.chain_steps_per_job * n_jobs * n_steps / (r['dev_ms'] * 1e-3)
            e2e_v, e_h2d, e_d2h = e2e_rate(hh, 8, 2, barrier, dist, torch)
            rec = {'workload': f'{wl}: {hh.n_points} point(s) x {hh.reps} repetitions = {hh.n_points * hh.reps} chains '
                               f'({hh.b_local} on this GPU) x {hh.n_iter} iterations x {hh.steps} steps',
                   'value': v, 'unit': UNIT, 'scaling': 'strong' if wl == 'profile30_strong' else 'weak',
                   'ms_per_job': r['dev_ms'] / (n_jobs * n_steps),
                   'e2e': {'value': e2e_v, 'unit': UNIT, 'h2d_bytes_per_step': e_h2d, 'd2h_bytes_per_step': e_d2h}}
            if world == 1:
                npad = hh.jobs[0].problem.npad
                rec['roofline'] = stream_roofline(
                    npad, hh.b_local * hh.n_iter * hh.steps, r['kernel_ms_per_job'],
                    'wide path (steps + tcgen05 position GEMM + FP64 anchor GEMM per iteration)' if npad >= 128
                    else f'anneal_kernel (lanes per chain {engine.choose_lanes(npad, hh.b_local)})')
            if wl == 'profile30_strong':
                rec['limiter'] = ('per-GPU occupancy: 8192 / N chains per GPU are fewer warps than the schedulers can '
                                  'interleave (latency-bound below ~8192 chains per GPU, DESIGN.md section 3)')
            sub[key] = rec
            del hh
            torch.cuda.empty_cache()
        sub['k5'] = k5_record(world, rank, dev, barrier, dist, torch, engine)

    # ---------------- time-to-profile: yaml -> run() -> profile/<p>.txt (+ snapshot) on disk ----------------
    ttp = None
    if headline != 'profile256':
        from prospect_public_b200.run import run as run_job
        from prospect_public_b200.toy import annealing_yaml
        tmp = tempfile.mkdtemp(prefix=f'pb200_ttp_r{rank}_')
        pkw = dict(workload_profile_kw('profile30', world)[1])
        ttp_paths = toy_inputs(tmp)[0]
        first = None
        for attempt in range(2):       # the first run of a new job shape pays one-off set-up (kernel loading, pinned
            # staging and, for N > 1, the peer-memory rendezvous of the exchange buffers); a fresh output folder per
            # attempt (deleting the previous run's ~100 MB snapshot is not part of the job)
            ycfg = annealing_yaml(ttp_paths, os.path.join(tmp, f'ttp{attempt}'), jobtype='profile', write=True, **pkw)
            barrier()
            t0 = time.perf_counter()
            res_job = run_job(ycfg)
            barrier()
            took = res_job.get('time_to_profile_s', time.perf_counter() - t0)
            first = took if first is None else first
        ttp = {'seconds': took, 'seconds_first_run': first,
               'seconds_incl_snapshot': time.perf_counter() - t0, 'loop_seconds': res_job['loop_seconds'],
               'chain_steps': int(res_job['chain_steps']),
               'job': f'30-D toy profile, {32 * world} values x 256 repetitions x 20 iterations x 250 steps, '
                      'seconds = yaml read to profile/x2.txt written (second run of the job; seconds_first_run includes '
                      'one-off set-up); the PROSPECT-compatible config.pkl/state.pkl follow'}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and headline != 'profile256':
        cores = os.cpu_count() or 1
        tmp = tempfile.mkdtemp(prefix='pb200_cpu_')
        kw = workload_profile_kw(headline, 1)[1]
        rate, secs, kind, sample, extra = reference_rate(headline, tmp, 'cb', kw, cores, chains_per_core=8)   # ~10 s
        cpu_baseline = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': kind,
                        'sample': f'{sample}; {secs:.1f} s'}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': res['dev_ms'] / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': f'{headline}: ' + ('256-D synthetic SPD Gaussian' if wide else
                                                          '30-D Gaussian toy (reference recipe, seed 0)') +
                                       f', {sa0.config.run.jobtype}, {sa0.n_points} point(s) x {sa0.reps} repetitions = '
                                       f'{sa0.n_points * sa0.reps} chains ({sa0.n_chains} per GPU) x {int(sa0.config.profile.max_iterations)} iterations x '
                                       f'{int(sa0.config.profile.steps_per_iteration)} steps per job; one step = {jobs} jobs '
                                       'back to back on separate device-resident chain batches',
                           'chain_steps_per_step': chain_steps_per_step, 'jobs_per_step': jobs,
                           'l2': 'flushed (192 MiB write) between steps; the chain batches + records of one step exceed L2',
                           'precision': 'FP32 proposal / acceptance arithmetic, FP64 chain state and reported chi^2',
                           'wall_ms_per_step_incl_reset_and_flush': 1e3 * res['wall_s'] / args.steps,
                           **({'exchange': 'per-iteration all-gather of the per-chain scalars ' + (
                               'fused into the annealing kernel (NVLink peer-memory stores, no collective call)'
                               if fused else 'over NCCL on a side stream')} if world > 1 else {})},
                'clocks': sampler.summary(),
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'steps': e2e_steps, 'warmup': 2,
                        'step': 'one job (the call a user makes); two jobs in flight: job k + 1 is uploaded and '
                                'launched while the records of job k travel back (results_async)'},
                'gpu_launches': res['launches'], 'roofline': roofline, 'cpu_baseline': cpu_baseline,
                'time_to_profile': ttp}
        line.update(sub)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
