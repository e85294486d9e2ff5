"""Harness around the UNMODIFIED reference (oracle/_ref, made by oracle/make_ref.py) -- TEST INFRASTRUCTURE.

Used by `bench.py --impl reference` / the `cpu_baseline` leg (the reference's own Scheduler in `mode: threaded` on the
host cores) and by tests/test_gpu_dropin.py (the reference's own tasks driving the CUDA kernel
class).  Nothing under prospect_public_b200/ imports this module.

What the harness has to do around the reference:
  * import shims for getdist / matplotlib (tests/golden/_shims; neither is installed and there is no network);
  * the reference never stops annealing -> OptimiseTask.emit_tasks is wrapped to stop after `max_iterations`;
  * Scheduler.finalize is broken -> the analysis task is run by hand when the profile file is wanted;
  * the kernel pickle is pre-created before the worker pool starts (start-up race of the parallel modes), and the
    run is checked for completeness (no failed task, expected number of OptimiseTasks).
"""
import contextlib
import copy
import io
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SHIMS = os.path.join(ROOT, 'tests', 'golden', '_shims')


def reference_root():
    """Directory that contains the reference's `prospect` package: oracle/_ref (travels to the GPU box) or, in the
    build container, /root/reference.  None if neither exists."""
    for cand in (os.path.join(HERE, '_ref'), '/root/reference'):
        if os.path.isdir(os.path.join(cand, 'prospect')):
            return cand
    return None


def available():
    return reference_root() is not None


def activate():
    """Put the shims and the reference on sys.path (idempotent).  Raises if the reference is not there."""
    root = reference_root()
    if root is None:
        raise RuntimeError('the reference is not available: run `python oracle/make_ref.py` where /root/reference exists')
    for path in (root, SHIMS):
        if path not in sys.path:
            sys.path.insert(0, path)
    return root


@contextlib.contextmanager
def quiet(enabled=True):
    if enabled:
        with contextlib.redirect_stdout(io.StringIO()):
            yield
    else:
        yield


@contextlib.contextmanager
def truncated_annealing(max_iter):
    """Stop every chain after `max_iter` OptimiseTasks.  emit_tasks runs in the scheduler's process (the done
    callback), so patching it here is enough for the process-pool mode too."""
    from prospect.tasks.optimise_task import OptimiseTask
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


def run_reference(config_yaml, max_iter, mode='serial', num_processes=1, analyse=False, verbose=False):
    """One job through the reference's own Configuration -> prepare_run -> Scheduler.delegate(executor).

    mode 'serial'   : SerialContext (communication.py:221-242);
    mode 'threaded' : concurrent.futures.ProcessPoolExecutor(max_workers=num_processes), what prospect/run.py:35-38
                      builds for `run.mode: threaded`.
    Returns dict(seconds, seconds_to_profile (when analyse), chain_steps, n_optimise_tasks, scheduler, config)."""
    activate()
    from prospect.communication import Scheduler, SerialContext
    from prospect.input import Configuration
    from prospect.io import prepare_run
    from prospect.kernels.initialisation import initialise_kernel
    cfg = copy.deepcopy(config_yaml)
    cfg['run']['mode'] = mode
    if mode == 'threaded':
        cfg['run']['num_processes'] = int(num_processes)
    t_start = time.perf_counter()
    with quiet(not verbose):
        config = Configuration(cfg)
        prepare_run(config)
        if config.io.write:
            initialise_kernel(config.kernel, config.io.dir, 0)      # pre-create analytical/random_gauss.pkl (F12)
        sched = Scheduler(config, False)
        if mode == 'threaded':
            from concurrent.futures import ProcessPoolExecutor
            pool = ProcessPoolExecutor(max_workers=int(num_processes))
        else:
            pool = SerialContext()
        t0 = time.perf_counter()
        with pool as executor, truncated_annealing(max_iter):
            sched.delegate(executor)
        seconds = time.perf_counter() - t0
    failed = len(sched.tasks.failed)
    if failed:
        raise RuntimeError(f'{failed} reference task(s) failed: ' +
                           str([getattr(t, "error", "")[-300:] for t in list(sched.tasks.failed.values())[:2]]))
    done = [t for t in sched.tasks.done.values() if t.type == 'OptimiseTask']
    steps = sum(int(sum(t.optimiser.mcmc.chain.mults)) - 1 for t in done)
    out = {'seconds': seconds, 'chain_steps': steps, 'n_optimise_tasks': len(done), 'scheduler': sched,
           'config': config}
    if analyse:
        with quiet(not verbose):
            task = sched.get_analysis_task()
            task.run([t for t in sched.tasks.done.values() if t.id in task.required_task_ids])
        out['seconds_to_profile'] = time.perf_counter() - t_start
    return out
