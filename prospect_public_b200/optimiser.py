"""Batched simulated annealing: every (profile point x repetition) chain of a job advances together.

Mirrors prospect/optimiser.py (`initialise_optimiser` :6-11, `SimulatedAnnealing` :40-162) and the
OptimiseTask loop around it (prospect/tasks/optimise_task.py:16-54).  The method names and the meaning of
`settings` / `bestfit` are kept; values that were scalars per task are arrays over the chains of this rank
(chain c = repetition * n_points + point, tasks/initialise_profile_task.py:16-26).  One call of
`optimise()` is one annealing iteration of ALL chains (one kernel launch); `run_schedule()` runs many
iterations per launch with set_bestfit / get_next_iteration_settings evaluated on the device.
"""
import math
import os

import numpy as np

from . import capi
from .distributed import dist_info, global_chain_index, shard_grid


def initialise_optimiser(config, kernel, optimise_settings, **kwargs):
    """prospect/optimiser.py:6-11.  A kernel whose likelihood is evaluated on the host (kernels.HostKernelAdapter)
    gets the optimiser that batches everything BUT the likelihood call (hosted.HostedSimulatedAnnealing)."""
    if config.profile.optimiser == 'simulated annealing':
        if getattr(kernel, 'host_evaluated', False):
            from .hosted import HostedSimulatedAnnealing
            return HostedSimulatedAnnealing(config, kernel, optimise_settings, **kwargs)
        return SimulatedAnnealing(config, kernel, optimise_settings, **kwargs)
    raise ValueError('You have specified a non-existing optimizer type.')


def get_temperature_change(config):
    """(T_end / T_start) ** (1 / max_iterations)  (tasks/initialise_optimiser_task.py:133-141)."""
    if config.profile.temperature_schedule != 'exponential':
        raise ValueError('Invalid temperature schedule.')
    tr = config.profile.temperature_range
    return (tr[1] / tr[0]) ** (1 / config.profile.max_iterations)


def get_initial_step_size_change(config):
    """tasks/initialise_optimiser_task.py:143-156."""
    prof = config.profile
    if prof.step_size_schedule == 'exponential':
        return (prof.step_size_range[1] / prof.step_size_range[0]) ** (1 / prof.max_iterations)
    if prof.step_size_schedule == 'adaptive':
        return {} if prof.step_size_adaptive_multiplier == 'adaptive' else None
    raise ValueError('Invalid step size schedule.')


def step_mode_of(profile_cfg):
    if profile_cfg.step_size_schedule == 'exponential':
        return capi.STEP_EXPONENTIAL
    if profile_cfg.step_size_adaptive_multiplier == 'adaptive':
        return capi.STEP_ADAPTIVE_SECANT
    return capi.STEP_ADAPTIVE_FIXED


def blank_rows(rows, world, scalar_len, n_alloc, device):
    """[rows, world, scalar_len] FP64 blocks in the "chain did not run this iteration" state: best / last loglkl = inf,
    accepted = -1 (the layout of engine.Records' scalar prefix)."""
    import torch
    blank = torch.zeros((rows, world, scalar_len), dtype=torch.float64, device=device)
    blank[:, :, :2 * n_alloc] = math.inf
    blank[:, :, 4 * n_alloc:].view(torch.int32).fill_(-1)
    return blank


class ExchangeUnavailable(RuntimeError):
    """Peer-mapped exchange buffers could not be set up (no symmetric memory / no peer access)."""


_EXCHANGE = {}


def _exchange_buffers(rows, world, scalar_len, n_alloc, device):
    """Two alternating sets of {exchange buffer [rows, world, scalar_len] FP64, arrival counters}, allocated in torch
    symmetric memory and mapped on every rank of the default group; cached per shape (they are scratch: results are
    copied out).  Invariant between uses: buffers hold the "did not run" pattern (loglkl = inf, accepted = -1),
    counters are zero.  Sets alternate per call, which is what makes a start-of-call barrier unnecessary: a rank
    can only begin call c + 2 (same set as call c) after every peer's kernel of call c + 1 has finished, i.e. after
    every peer restored the set behind call c."""
    import torch
    import torch.distributed as dist
    key = (int(rows), int(world), int(scalar_len), str(device))
    if key not in _EXCHANGE:
        try:
            import torch.distributed._symmetric_memory as symm_mem
            sets = []
            for _ in range(2):
                buf = symm_mem.empty((rows, world, scalar_len), dtype=torch.float64, device=device)
                prog = symm_mem.empty(max(64, rows), dtype=torch.int32, device=device)
                sets.append({'buf': buf, 'prog': prog, 'hb': symm_mem.rendezvous(buf, dist.group.WORLD),
                             'hp': symm_mem.rendezvous(prog, dist.group.WORLD)})
        except Exception as exc:            # noqa: BLE001 -- any failure here means "use the NCCL path"
            raise ExchangeUnavailable(f'{type(exc).__name__}: {exc}') from exc
        blank = blank_rows(rows, world, scalar_len, n_alloc, device)
        for st in sets:
            st['buf'].copy_(blank)
            st['prog'].zero_()
        torch.cuda.synchronize(device)
        dist.barrier()                       # every rank's buffers are blank before anyone stores remotely
        _EXCHANGE[key] = {'sets': sets, 'blank': blank, 'calls': 0}
    return _EXCHANGE[key]


_COPY_STREAMS = {}


def _copy_stream(device):
    import torch
    key = str(device)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[key]


class PendingResults:
    """A device-to-host read in flight (SimulatedAnnealing.results_async).  Keeps the device tensors alive until the
    copy has completed."""

    def __init__(self, optimiser, iteration, src, host, done):
        self.optimiser, self.iteration, self._src, self._host, self._done = optimiser, iteration, src, host, done
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in host.values())

    def get(self):
        """(bestfit dict as `set_bestfit` builds it, {name: host tensor} of the reduced records)."""
        self._done.synchronize()
        sa, h = self.optimiser, self._host
        acc = h['accepted'].numpy().copy()
        ran = acc >= 0
        names = sa.kernel.varying_param_names
        steps = sa.schedule.steps_per_iteration
        bx = h['best_x'].numpy()[:, :sa.problem.n].copy() if 'best_x' in h else None
        sa.bestfit = {
            'loglkl': np.where(ran, h['best_loglkl'].numpy(), np.inf),
            'acceptance_rate': np.where(ran, (1 + acc) / (1.0 + steps), np.nan),
            'accepted_steps': np.where(ran, 1 + acc, 0),
            'iteration_number': self.iteration,
            'position': {nm: bx[:, i] for i, nm in enumerate(names)} if bx is not None else {},
        }
        reduced = {name[len('reduced/'):]: t.clone() for name, t in h.items() if name.startswith('reduced/')}
        self._src = self._host = None
        return sa.bestfit, reduced


class SimulatedAnnealing:
    """optimise_settings (arrays over the P profile points; P = 1 for a global optimisation):
        initial_position [P, n] (varying order), covmat [P, n, n], current_best_loglkl [P],
        fixed_param_val [P] or None, repetitions R, temperature, temperature_change, step_size,
        step_size_change, iteration_number (0), max_rows (record rows to allocate), seed."""

    def __init__(self, config, kernel, optimise_settings, keep_positions=True):
        from . import engine
        self.config, self.kernel, self.settings = config, kernel, optimise_settings
        prof = config.profile
        if prof.steps_per_iteration is None:
            raise ValueError("The GPU path needs 'steps_per_iteration' (time-boxed 'minutes_per_iteration' "
                             'iterations would make results depend on the hardware).')
        s = optimise_settings
        if 'initial_position' not in s or len(s['initial_position']) == 0:
            raise ValueError('No point to start the optimisation from.')
        if 'covmat' not in s:
            raise ValueError('No covmat to start the optimisation from.')
        starts = np.atleast_2d(np.asarray(s['initial_position'], dtype=np.float64))
        covs = np.asarray(s['covmat'], dtype=np.float64)
        covs = covs[None] if covs.ndim == 2 else covs
        self.covmat = covs
        fixed = s.get('fixed_param_val')
        self.n_points = starts.shape[0]
        self.reps = int(s.get('repetitions', 1))
        self.rank, self.world, _ = dist_info()
        pt, rep = shard_grid(self.n_points, self.reps, self.rank, self.world)
        self.point_global, self.rep = pt, rep
        # a rank without chains (fewer chains than GPUs) still builds a (dummy, one-point) problem and joins every
        # collective; its launches are no-ops
        self.local_points = np.unique(pt) if len(pt) else np.array([0])
        point_local = np.searchsorted(self.local_points, pt).astype(np.int32)
        fixed_local = None if fixed is None else np.asarray(fixed, dtype=np.float64)[self.local_points]
        self.fixed_local = fixed_local
        # every point on this rank, in order: hand the caller's array through (a read-only one is cached by identity)
        all_local = len(self.local_points) == self.n_points and np.array_equal(self.local_points, np.arange(self.n_points))
        self.problem = self._device_problem(kernel, fixed_local, covs if all_local else covs[self.local_points])
        cb = np.broadcast_to(np.asarray(s.get('current_best_loglkl', math.inf), dtype=np.float64), (self.n_points,))
        gid = global_chain_index(pt, rep, self.n_points)
        resumed = s.get('chain_state')            # global-order per-chain state of an interrupted run (run.resume)
        self.batch = engine.ChainBatch(self.problem, starts[pt], point_local, chain_id=gid,
                                       temperature=s['temperature'], step_size=s['step_size'], current_best=cb[pt],
                                       state=None if resumed is None else {k: np.asarray(v)[gid]
                                                                           for k, v in resumed.items()})
        change = s.get('step_size_change')
        interval = prof.step_size_adaptive_interval or (0.0, 1.0)
        mult = prof.step_size_adaptive_multiplier
        self.schedule = engine.make_schedule(
            prof.steps_per_iteration, s['temperature_change'], step_mode_of(prof),
            change if isinstance(change, float) else 1.0, interval, mult if not isinstance(mult, str) else 0.0,
            prof.chi2_tolerance)
        self.max_rows = int(s.get('max_rows', prof.max_iterations))
        # every rank allocates record blocks of the same size so that one all-gather per iteration moves them
        self.b_alloc = max(len(shard_grid(self.n_points, self.reps, r, self.world)[0]) for r in range(self.world))
        self.keep_positions = keep_positions
        self.records = engine.Records(self.problem, self.batch.n_chains, self.max_rows, keep_positions,
                                      n_alloc=self.b_alloc)
        # lanes per chain: chosen from the size of the WHOLE job, not of this rank's shard, so that a chain's arithmetic
        # (the order of its FP32 reductions) -- and with it every record -- does not depend on the number of GPUs
        # (PB200_LANES_PER_CHAIN: A/B runs of the geometries, scripts/r02_lanes4.sh)
        self.lanes = int(s.get('lanes_per_chain', 0)) or int(os.environ.get('PB200_LANES_PER_CHAIN', '0') or 0) or \
            engine.choose_lanes(self.problem.npad, self.n_points * self.reps)
        self.seed = int(s.get('seed', 0))
        self.step_counter = 0                 # MH steps taken so far by every live chain (they advance in lockstep)
        self.iterations_done = int(s.get('iteration_number', 0))
        self.bestfit = None
        self._gatherer = None
        self.signalled = True
        self.fused_exchange = os.environ.get('PB200_EXCHANGE', 'fused') == 'fused'

    @property
    def n_chains(self):
        return self.batch.n_chains

    def _device_problem(self, kernel, fixed_local, covmats_local):
        """Device image of what the chains of this rank sample with (hosted.HostedSimulatedAnnealing overrides it)."""
        return kernel.device_problem(fixed_values=fixed_local, proposal_covmats=covmats_local)

    # --- the reference's three verbs -------------------------------------------------------------
    def optimise(self, n_iterations=1, stream=None):
        """Run `n_iterations` annealing iterations of every chain in one launch."""
        from . import engine
        engine.anneal_run(self.problem, self.batch, self.schedule, self.records, n_iterations, self.seed, stream,
                          self.lanes)
        self.iterations_done += n_iterations
        self.step_counter += n_iterations * self.schedule.steps_per_iteration

    def set_bestfit(self, iteration=None):
        """Host view of one iteration's records: loglkl, acceptance_rate, accepted_steps, position per chain
        (optimiser.py:74-85).  Chains that did not run that iteration have accepted_steps = 0 / loglkl = inf."""
        k = self.iterations_done - 1 if iteration is None else iteration
        r = self.records
        acc = r.accepted[k].cpu().numpy()
        ran = acc >= 0
        names = self.kernel.varying_param_names
        bx = r.best_x[k, :, :self.problem.n].cpu().numpy() if r.best_x is not None else None
        steps = self.schedule.steps_per_iteration
        self.bestfit = {
            'loglkl': np.where(ran, r.best_loglkl[k].cpu().numpy(), np.inf),
            'acceptance_rate': np.where(ran, (1 + acc) / (1.0 + steps), np.nan),
            'accepted_steps': np.where(ran, 1 + acc, 0),
            'iteration_number': k,
            'position': {nm: bx[:, i] for i, nm in enumerate(names)} if bx is not None else {},
        }
        return self.bestfit

    def results_async(self, iteration=None, reduced=None):
        """Start the device-to-host read of one iteration's records (what `set_bestfit` returns) and, optionally, of the
        tensors of `reduce_over_iterations()`, on a copy stream into pinned host memory, and return at once.
        `PendingResults.get()` waits for the copy.  A caller with several independent jobs (the repetitions of a
        reanneal, bench.py's end-to-end arm) builds and launches job k + 1 while job k is still on the device and its
        results are on their way back; nothing in the kernels changes."""
        import torch
        k = self.iterations_done - 1 if iteration is None else iteration
        r, dev = self.records, self.problem.device
        src = {'accepted': r.accepted[k], 'best_loglkl': r.best_loglkl[k]}
        if r.best_x is not None:
            src['best_x'] = r.best_x[k]
        for name, t in (reduced or {}).items():
            src['reduced/' + name] = t
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        cs = _copy_stream(dev)
        host = {}
        with torch.cuda.stream(cs):
            cs.wait_event(ready)
            for name, t in src.items():
                host[name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                host[name].copy_(t, non_blocking=True)
            done = torch.cuda.Event()
            done.record(cs)
        if self.world > 1 and os.environ.get('PB200_ASYNC_MULTI_GPU', '0') != '1':
            # Jobs sharded over several GPUs are kept one at a time: with two of them in flight a rank was seen to hang
            # in an 8-GPU run (bench.py e2e, 2026-10; not understood: the kernel-side exchange protocol is ordered
            # on the device and does not depend on the host), so the read is completed here, before the caller can
            # launch the next job.
            done.synchronize()
        return PendingResults(self, k, src, host, done)

    def get_next_iteration_settings(self):
        """The settings the next iteration will run with (already computed on the device, optimiser.py:87-162)."""
        b = self.batch
        out = {
            'current_best_loglkl': b.current_best.cpu().numpy(),
            'initial_position': b.positions(),
            'covmat': self.covmat,
            'temperature': b.temperature.cpu().numpy(),
            'temperature_change': self.schedule.temperature_change,
            'step_size': b.step_size.cpu().numpy(),
            'iteration_number': b.iteration.cpu().numpy(),
            'status': b.status.cpu().numpy(),
        }
        if self.settings.get('fixed_param_val') is not None:
            out['fixed_param_val'] = np.asarray(self.settings['fixed_param_val'])[self.point_global]
        return out

    # --- whole schedules ---------------------------------------------------------------------------
    def run_schedule(self, n_iterations, iterations_per_launch=None, gather=True):
        """All iterations.  Single GPU: `iterations_per_launch` (default: all) per launch.  Several GPUs: still ONE
        call, and the iteration-boundary exchange of the per-chain scalars (bestfit chi^2, last chi^2,
        temperature, step size, accepted count) is done by the kernel itself over peer memory (run_fused_exchange).
        Fallbacks: the kernel counts finished iterations and a side stream all-gathers each one over NCCL behind a
        device-side wait (run_signalled; PB200_EXCHANGE=nccl or no symmetric memory); one launch + one gather per
        iteration (`self.signalled = False`).  The bestfit positions of all iterations follow in one all-gather at
        the end (global_records)."""
        import torch
        if self.n_chains == 0 and (self.world == 1 or not gather):
            return None
        if self.world == 1 or not gather:
            per = n_iterations if iterations_per_launch is None else iterations_per_launch
            done = 0
            while done < n_iterations:
                k = min(per, n_iterations - done)
                self.optimise(k)
                done += k
            return None
        import torch.distributed as dist
        if self._gatherer is None:
            self._gatherer = True
            self._comm_stream = torch.cuda.Stream(device=self.problem.device)
            # rows that were never shipped must read "did not run" (loglkl = inf, accepted = -1), exactly like the
            # exchange buffers and a single-GPU Records: a partial run (stop_after, an interrupted snapshot) decodes
            # every row of this tensor
            self.gathered = blank_rows(self.max_rows, self.world, self.records.scalar_len, self.b_alloc,
                                       self.problem.device)

        def ship(k):
            dist.all_gather_into_tensor(self.gathered[k].view(-1), self.records.buf[k, :self.records.scalar_len])

        if self.fused_exchange:
            try:
                self.run_fused_exchange(n_iterations, self._comm_stream)
                return self.gathered
            except ExchangeUnavailable as exc:        # no peer-mapped memory on this system: NCCL gathers instead
                self.fused_exchange = False
                if self.rank == 0:
                    print(f'prospect_public_b200: fused exchange unavailable ({exc}); using NCCL all-gathers')
        if self.signalled:
            self.run_signalled(n_iterations, ship, self._comm_stream)
            return self.gathered
        compute = torch.cuda.current_stream()
        for _ in range(n_iterations):
            k = self.iterations_done
            self.optimise(1)
            ev = torch.cuda.Event()
            ev.record(compute)
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(ev)
                ship(k)
        compute.wait_stream(self._comm_stream)
        return self.gathered

    def run_signalled(self, n_iterations, ship, side_stream):
        """One fused launch of `n_iterations` iterations whose kernel signals each finished iteration; for every
        iteration, in order, `ship(row)` is issued on `side_stream` behind a device-side wait on that iteration's
        counter (pb200_anneal_run_signalled / pb200_stream_wait_geq).  All chains of a batch share the iteration
        counter here, so launch-iteration i is record row iterations_done + i."""
        import torch
        from . import engine
        compute = torch.cuda.current_stream()
        progress = torch.zeros(n_iterations, dtype=torch.int32, device=self.problem.device)
        zeroed = torch.cuda.Event()
        zeroed.record(compute)
        k0 = self.iterations_done
        target = engine.anneal_run_signalled(self.problem, self.batch, self.schedule, self.records, n_iterations,
                                             self.seed, progress, None, self.lanes)
        self.iterations_done += n_iterations
        self.step_counter += n_iterations * self.schedule.steps_per_iteration
        with torch.cuda.stream(side_stream):
            side_stream.wait_event(zeroed)
            for i in range(n_iterations):
                engine.stream_wait_geq(side_stream, progress, i, target)
                ship(k0 + i)
        compute.wait_stream(side_stream)
        self._progress = progress            # keep alive until the streams are done with it

    def run_fused_exchange(self, n_iterations, side_stream):
        """One fused launch whose kernel performs the iteration-boundary all-gather itself: every warp stores its
        chains' per-iteration scalars straight into ALL ranks' exchange buffers (peer memory over NVLink / NVSwitch,
        mapped through torch symmetric memory) and bumps an arrival counter on every rank (pb200_anneal_run_shared).
        No collective call and no extra kernel sit on the path; the side stream only waits (device-side) for the
        arrival counter to reach the number of chain warps of the whole job and then copies the finished rows into
        this optimiser's `gathered` tensor."""
        import torch
        from . import engine
        pool = _exchange_buffers(self.max_rows, self.world, self.records.scalar_len, self.b_alloc,
                                 self.problem.device)
        ex = pool['sets'][pool['calls'] % 2]
        pool['calls'] += 1
        compute = torch.cuda.current_stream()
        k0, b, n = self.iterations_done, self.b_alloc, n_iterations
        peers = engine.make_peers(self.rank, b, self.max_rows, self.records.scalar_len, ex['hb'].buffer_ptrs,
                                  ex['hp'].buffer_ptrs)
        engine.anneal_run_shared(self.problem, self.batch, self.schedule, self.records, n, self.seed, peers, None,
                                 None, self.lanes)
        self.iterations_done += n
        self.step_counter += n * self.schedule.steps_per_iteration
        # arrivals of the whole job: every rank's count for ITS number of chains, resolved by the library itself
        # (same geometry rules and environment overrides as the launch)
        total = sum(engine.exchange_arrivals(self.problem.npad, len(shard_grid(self.n_points, self.reps, r,
                                                                              self.world)[0]), self.lanes)
                    for r in range(self.world))
        rows = ex['buf'][k0:k0 + n]
        with torch.cuda.stream(side_stream):
            # every warp bumps its counters in ascending order after its completion fence: the last one suffices
            engine.stream_wait_geq(side_stream, ex['prog'], n - 1, total)
            self.gathered[k0:k0 + n].copy_(rows, non_blocking=True)
            rows.copy_(pool['blank'][k0:k0 + n], non_blocking=True)      # leave the set blank for its next use
            ex['prog'][:n].zero_()
        compute.wait_stream(side_stream)

    def global_records(self):
        """Host arrays over ALL chains of the job in global order c = rep * P + point:
        best_loglkl [K, B], accepted [K, B], step_size [K, B], temperature [K, B], best_x [K, B, n] or None."""
        from . import engine
        n, k = self.problem.n, self.max_rows
        if self.world == 1:
            r = self.records
            out = {name: getattr(r, name).cpu().numpy() for name in ('best_loglkl', 'accepted', 'step_size',
                                                                      'temperature')}
            out['best_x'] = r.best_x[:, :, :n].cpu().numpy() if r.best_x is not None else None
            return out
        import torch
        import torch.distributed as dist
        g = self.gathered.cpu().numpy()                       # [K, world, scalar_len]
        gx = None
        if self.keep_positions:
            r = self.records
            o, width = r.offsets['best_x'], r.n_alloc * r.npad
            local = r.buf[:, o:o + width].contiguous()
            allx = torch.empty((self.world,) + tuple(local.shape), dtype=torch.float64, device=local.device)
            dist.all_gather_into_tensor(allx.view(-1), local.view(-1))
            gx = allx.cpu().numpy().reshape(self.world, k, r.n_alloc, r.npad)
        b_glob = self.n_points * self.reps
        out = {'best_loglkl': np.full((k, b_glob), np.inf), 'accepted': np.full((k, b_glob), -1, dtype=np.int32),
               'step_size': np.zeros((k, b_glob)), 'temperature': np.zeros((k, b_glob)),
               'best_x': np.zeros((k, b_glob, n)) if self.keep_positions else None}
        for rank in range(self.world):
            pt, rep = shard_grid(self.n_points, self.reps, rank, self.world)
            ids = global_chain_index(pt, rep, self.n_points)
            dec = engine.Records.decode(g[:, rank], len(ids), self.b_alloc, self.problem.npad, False)
            for name in ('best_loglkl', 'accepted', 'step_size', 'temperature'):
                out[name][:, ids] = dec[name]
            if self.keep_positions:
                out['best_x'][:, ids] = gx[rank, :, :len(ids), :n]
        return out

    def global_scalars(self):
        """global_records() without the positions: best_loglkl / accepted / step_size / temperature [K, B] in global chain
        order.  Several GPUs: decoded from the per-iteration exchange (`gathered`), no further communication."""
        from . import engine
        names = ('best_loglkl', 'accepted', 'step_size', 'temperature')
        if self.world == 1:
            return {name: getattr(self.records, name).cpu().numpy() for name in names}
        g = self.gathered.cpu().numpy()
        b_glob, k = self.n_points * self.reps, self.max_rows
        out = {'best_loglkl': np.full((k, b_glob), np.inf), 'accepted': np.full((k, b_glob), -1, dtype=np.int32),
               'step_size': np.zeros((k, b_glob)), 'temperature': np.zeros((k, b_glob))}
        for rank in range(self.world):
            pt, rep = shard_grid(self.n_points, self.reps, rank, self.world)
            ids = global_chain_index(pt, rep, self.n_points)
            dec = engine.Records.decode(g[:, rank], len(ids), self.b_alloc, self.problem.npad, False)
            for name in names:
                out[name][:, ids] = dec[name]
        return out

    def global_best_positions(self):
        """Per chain of the job (global order): the position of its best iteration [B, n] -- what the result files
        need, picked on the device (first-minimum iteration from the running-bestfit reduction, one gather) and
        exchanged as B x n doubles, instead of moving every iteration's positions ([K, B, n]) to rank 0 first."""
        import torch
        from . import engine
        n, b_loc = self.problem.n, self.n_chains
        if b_loc:
            red = engine.chain_minimum(self.records, self.max_rows)
            it = red['chain_iter'].clamp(min=0).long()
            local = self.records.best_x[it, torch.arange(b_loc, device=it.device)][:, :n].contiguous()
        else:
            local = torch.zeros((0, n), dtype=torch.float64, device=self.problem.device)
        if self.world == 1:
            return local.cpu().numpy()
        import torch.distributed as dist
        pad = torch.zeros((self.b_alloc, n), dtype=torch.float64, device=self.problem.device)
        pad[:b_loc] = local
        allx = torch.empty((self.world, self.b_alloc, n), dtype=torch.float64, device=self.problem.device)
        dist.all_gather_into_tensor(allx.view(-1), pad.view(-1))
        allx = allx.cpu().numpy()
        out = np.zeros((self.n_points * self.reps, n))
        for rank in range(self.world):
            pt, rep = shard_grid(self.n_points, self.reps, rank, self.world)
            out[global_chain_index(pt, rep, self.n_points)] = allx[rank, :len(pt)]
        return out

    def global_positions(self):
        """best_x [K, B, n] of every iteration in global chain order (snapshots; one all-gather for N > 1)."""
        n, k = self.problem.n, self.max_rows
        r = self.records
        if self.world == 1:
            return r.best_x[:, :, :n].cpu().numpy()
        import torch
        import torch.distributed as dist
        o, width = r.offsets['best_x'], r.n_alloc * r.npad
        local = r.buf[:, o:o + width].contiguous()
        allx = torch.empty((self.world,) + tuple(local.shape), dtype=torch.float64, device=local.device)
        dist.all_gather_into_tensor(allx.view(-1), local.view(-1))
        gx = allx.cpu().numpy().reshape(self.world, k, r.n_alloc, r.npad)
        out = np.zeros((k, self.n_points * self.reps, n))
        for rank in range(self.world):
            pt, rep = shard_grid(self.n_points, self.reps, rank, self.world)
            ids = global_chain_index(pt, rep, self.n_points)
            out[:, ids] = gx[rank, :, :len(ids), :n]
        return out

    def global_status(self):
        """status of every chain of the job in global order (one small all-gather)."""
        import torch
        if self.world == 1:
            return self.batch.status.cpu().numpy()
        import torch.distributed as dist
        pad = torch.full((self.b_alloc,), -1, dtype=torch.int32, device=self.problem.device)
        pad[:self.n_chains] = self.batch.status
        allst = torch.empty((self.world, self.b_alloc), dtype=torch.int32, device=self.problem.device)
        dist.all_gather_into_tensor(allst.view(-1), pad)
        allst = allst.cpu().numpy()
        out = np.zeros(self.n_points * self.reps, dtype=np.int32)
        for rank in range(self.world):
            pt, rep = shard_grid(self.n_points, self.reps, rank, self.world)
            out[global_chain_index(pt, rep, self.n_points)] = allst[rank, :len(pt)]
        return out

    def global_chain_state(self):
        """Per-chain state of every chain of the job in global order (what run.resume needs to continue the run):
        position, loglkl, temperature, step size, current best, step-size history, step / iteration counters,
        status.  One small all-gather for N > 1."""
        import torch
        from .engine import ChainBatch
        local = self.batch.host_state()
        if self.world == 1:
            return local
        import torch.distributed as dist
        n = self.problem.n
        keys = [k for k in ChainBatch.STATE_KEYS if k != 'x']
        pack = np.zeros((self.b_alloc, n + len(keys)))
        pack[:self.n_chains, :n] = local['x']
        for i, k in enumerate(keys):
            pack[:self.n_chains, n + i] = local[k]          # counters are small integers: exact in FP64
        mine = torch.as_tensor(pack, device=self.problem.device)
        allp = torch.empty((self.world,) + tuple(mine.shape), dtype=torch.float64, device=self.problem.device)
        dist.all_gather_into_tensor(allp.view(-1), mine.view(-1))
        allp = allp.cpu().numpy()
        b_glob = self.n_points * self.reps
        out = {'x': np.zeros((b_glob, n))}
        out.update({k: np.zeros(b_glob, dtype=local[k].dtype) for k in keys})
        for rank in range(self.world):
            pt, rep = shard_grid(self.n_points, self.reps, rank, self.world)
            ids = global_chain_index(pt, rep, self.n_points)
            out['x'][ids] = allp[rank, :len(ids), :n]
            for i, k in enumerate(keys):
                out[k][ids] = allp[rank, :len(ids), n + i].astype(local[k].dtype)
        return out

    def reduce_over_iterations(self):
        """Running-bestfit reduction on the device (analyse_profile_task.py:91-121) for this rank's chains when
        it owns whole points (world == 1 or point sharding)."""
        from . import engine
        lp = len(self.local_points)
        reps_local = self.n_chains // lp
        return engine.profile_reduce(self.records, self.iterations_done, lp, reps_local)

    def finalize(self):
        pass
