"""A GPU batch executor with PROSPECT's executor protocol.

`prospect.run.run` hands its Scheduler a context manager whose `__enter__` returns an object with
`submit(fn, *args) -> future`; the future has `add_done_callback(cb)` and `result()` returning the finished task
(prospect/run.py:27-58, prospect/communication.py:50-68,119-121,221-242).  The reference ships three of them: an MPI
pool, a process pool and a synchronous `SerialContext`.  `GpuBatchExecutor` is a fourth: the UNMODIFIED scheduler keeps
emitting one `OptimiseTask` per (profile value, repetition, iteration), but instead of running each task's
`steps_per_iteration` Metropolis steps in a worker process, the executor collects every OptimiseTask that is ready,
advances ALL of them by one annealing iteration in ONE launch of the fused kernel (`pb200_anneal_run`) and hands each
task back with an `optimiser` object that has what the reference reads afterwards (`bestfit`, `settings`, `covmat`,
`get_next_iteration_settings()`, `mcmc.chain.last_position`; optimise_task.py:29-54, analyse_profile_task.py:62-121).
Every other task type (InitialiseProfileTask, InitialiseOptimiserTask, Analyse*Task) runs as submitted.

What a maintainer adds to prospect/run.py:35-44 is one branch:

    elif config.run.mode == 'gpu':
        from prospect_public_b200.executor import GpuBatchExecutor
        pool, pool_kwargs = GpuBatchExecutor, {}

This keeps the reference's control plane (and its per-task Python overhead: ~50 us per task in the scheduler); the
fast path for large jobs is `prospect_public_b200.run.run`, which also keeps the iterations on the device.  Nothing
here imports the reference: tasks are duck-typed (`type`, `optimise_settings`, `config`).
tests/test_gpu_dropin.py drives the unmodified reference's Scheduler with this executor.
"""
import threading
import time
from concurrent.futures import Future

import numpy as np

from . import capi


class _Chain:
    """The two things the reference reads from `optimiser.mcmc.chain` after a task has run."""

    def __init__(self, last_position, n_steps):
        self.last_position = last_position
        self.mults = [n_steps + 1]          # sum(mults) = steps + 1 (optimise_task.py:27 prints it)


class _Mcmc:
    def __init__(self, chain):
        self.chain = chain
        self.kernel = None                  # OptimiseTask.finalize deletes it (optimise_task.py:29-31)


class BatchedAnnealingResult:
    """Stands in for the reference's SimulatedAnnealing object on a finished OptimiseTask."""

    def __init__(self, config, settings, covmat, bestfit, last_position, next_temperature, next_step_size, n_steps):
        self.config, self.settings, self.covmat, self.bestfit = config, settings, covmat, bestfit
        self.mcmc = _Mcmc(_Chain(last_position, n_steps))
        self.kernel = None
        self._next = (next_temperature, next_step_size)

    def get_next_iteration_settings(self):
        """prospect/optimiser.py:87-162: temperature and step size of the next iteration were computed on the device by
        the same rule (bit-identical arithmetic, tests/test_gpu_parity.py), the rest is carried over."""
        out = {'current_best_loglkl': self.bestfit['loglkl'],
               'initial_position': self.mcmc.chain.last_position,
               'covmat': self.covmat,
               'temperature': self._next[0],
               'temperature_change': self.settings['temperature_change'],
               'step_size': self._next[1],
               'step_size_change': self.settings['step_size_change']}
        if self.config.run.jobtype == 'profile':
            out['fixed_param_val'] = self.settings['fixed_param_val']
        return out


class GpuBatchExecutor:
    """`with GpuBatchExecutor() as executor: scheduler.delegate(executor)`.

    linger: seconds the dispatcher waits for further OptimiseTasks after the last one arrived before it launches the
    batch (the scheduler submits ready tasks one by one from its own thread)."""

    def __init__(self, device='cuda:0', linger=0.003, seed=None):
        self.device, self.linger, self.seed = device, linger, seed
        self._cv = threading.Condition()
        self._queue, self._stop, self._thread = [], False, None
        self._kernels = {}
        self.launches, self.tasks_batched = 0, 0

    # --- executor protocol ------------------------------------------------------------------------------------
    def __enter__(self):
        capi.lib()                                    # fail loudly without the CUDA extension: no CPU fallback
        self._thread = threading.Thread(target=self._dispatch, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        self._thread.join()
        return False

    def submit(self, fn, *args):
        fut = Future()
        with self._cv:
            self._queue.append((fut, fn, args, time.perf_counter()))
            self._cv.notify_all()
        return fut

    # --- dispatcher ---------------------------------------------------------------------------------------------
    def _dispatch(self):
        pending = []                                   # OptimiseTasks waiting for their batch
        while True:
            with self._cv:
                while not self._queue and not self._stop and not pending:
                    self._cv.wait()
                if pending and not self._queue:        # give the scheduler a moment to submit the rest of the wave
                    self._cv.wait(timeout=self.linger)
                items, self._queue = self._queue, []
                if self._stop and not items and not pending:
                    return
            for fut, fn, args, t_in in items:
                task = getattr(fn, '__self__', None)
                if getattr(task, 'type', None) == 'OptimiseTask':
                    pending.append((fut, task))
                else:
                    fut.set_result(fn(*args))          # any other task: run it as submitted (callbacks fire here)
            if pending and not items:                  # nothing new arrived during the linger: launch
                batch, pending = pending, []
                try:
                    self._run_batch(batch)
                except Exception:                      # a failed launch fails its tasks the reference's way
                    import traceback
                    err = traceback.format_exc()
                    for fut, task in batch:
                        task.success, task.error = False, err
                        fut.set_result(task)

    # --- one annealing iteration of every collected task in one launch ---------------------------------------------
    def _kernel_for(self, config):
        from .kernels import CudaGaussianKernel
        key = (config.kernel.param, config.io.dir)
        if key not in self._kernels:
            self._kernels[key] = CudaGaussianKernel(config.kernel, 0, output_folder=config.io.dir, device=self.device)
            if config.run.jobtype == 'profile':
                self._kernels[key].set_fixed_parameters({config.profile.parameter: 0.0})
        return self._kernels[key]

    def _run_batch(self, batch):
        import torch
        from . import engine
        from .optimiser import step_mode_of
        tasks = [t for _, t in batch]
        config = tasks[0].config
        prof = config.profile
        if prof.step_size_schedule == 'adaptive' and isinstance(prof.step_size_adaptive_multiplier, str):
            raise NotImplementedError("the 'adaptive' step-size multiplier keeps its history in the task settings; "
                                      'use a numeric multiplier with the batch executor')
        kernel = self._kernel_for(config)
        names = kernel.varying_param_names
        is_profile = config.run.jobtype == 'profile'
        steps = int(prof.steps_per_iteration)
        # profile points of this batch: distinct (fixed value, proposal covariance) pairs
        point_of, fixed_values, covmats, index = [], [], [], {}
        for t in tasks:
            s = t.optimise_settings
            cov = np.asarray(s['covmat'], dtype=np.float64)
            key = (float(s['fixed_param_val']) if is_profile else None, cov.tobytes())
            if key not in index:
                index[key] = len(covmats)
                covmats.append(cov)
                fixed_values.append(key[0])
            point_of.append(index[key])
        problem = kernel.device_problem(fixed_values=np.array(fixed_values) if is_profile else None,
                                        proposal_covmats=np.stack(covmats))
        b = len(tasks)
        x0 = np.array([[t.optimise_settings['initial_position'][nm][0] for nm in names] for t in tasks])
        sets = [t.optimise_settings for t in tasks]
        # Philox stream of a chain: (repetition, profile value) -> chain id, iteration -> step counter
        vals = sorted({float(s['fixed_param_val']) for s in sets}) if is_profile else [None]
        chain_id = np.array([s['repetition_number'] * 65536 + (vals.index(float(s['fixed_param_val'])) if is_profile else 0)
                             for s in sets], dtype=np.uint32)
        chains = engine.ChainBatch(problem, x0, np.asarray(point_of, dtype=np.int32), chain_id=chain_id,
                                   temperature=np.array([s['temperature'] for s in sets], dtype=np.float64),
                                   step_size=np.array([s['step_size'] for s in sets], dtype=np.float64),
                                   current_best=np.array([s['current_best_loglkl'] for s in sets], dtype=np.float64),
                                   steps_done=np.array([s['iteration_number'] * steps for s in sets], dtype=np.int32))
        change = sets[0]['step_size_change']
        mult = prof.step_size_adaptive_multiplier
        sched = engine.make_schedule(steps, sets[0]['temperature_change'], step_mode_of(prof),
                                     change if isinstance(change, float) else 1.0,
                                     prof.step_size_adaptive_interval or (0.0, 1.0),
                                     mult if isinstance(mult, (int, float)) else 0.0, None)
        recs = engine.Records(problem, b, 1, keep_positions=True)
        seed = self.seed if self.seed is not None else int(config.run.numpy_random_seed or 0)
        engine.anneal_run(problem, chains, sched, recs, 1, seed)
        engine.launch_status()
        self.launches += 1
        self.tasks_batched += b
        n = problem.n
        best = recs.best_loglkl[0].cpu().numpy()
        acc = recs.accepted[0].cpu().numpy()
        best_x = recs.best_x[0, :, :n].cpu().numpy()
        last_x = recs.last_x[0, :, :n].cpu().numpy()
        t_next = chains.temperature.cpu().numpy()
        s_next = chains.step_size.cpu().numpy()
        for i, (fut, task) in enumerate(batch):
            s = task.optimise_settings
            fixed = {config.profile.parameter: [s['fixed_param_val']]} if is_profile else {}
            # key order of the reference's chain positions: varying parameters, then the fixed one (mcmc.py:30-46)
            position = dict({nm: [float(best_x[i, j])] for j, nm in enumerate(names)}, **fixed)
            last = dict({nm: [float(last_x[i, j])] for j, nm in enumerate(names)}, **fixed)
            bestfit = {'loglkl': float(best[i]), 'acceptance_rate': (1 + int(acc[i])) / (1.0 + steps),
                       'accepted_steps': 1 + int(acc[i]), 'iteration_number': s['iteration_number'],
                       'position': position}
            task.optimiser = BatchedAnnealingResult(config, s, s['covmat'], bestfit, last, float(t_next[i]),
                                                    float(s_next[i]), steps)
            task.success, task.error = True, None
            task.finalize()                            # what run_return_self does after run() (base_task.py:41-55)
            fut.set_result(task)
