"""Drop-in proof for the kernel plugin boundary: the UNMODIFIED reference -- its
own Scheduler, InitialiseOptimiserTask, OptimiseTask, SimulatedAnnealing, MetropolisHastings and
AnalyseProfileTask (oracle/_ref, made by oracle/make_ref.py) -- runs a profile job with `CudaGaussianKernel` behind
`initialise_kernel`, which is the two-line registration INTEGRATION.md section 3 describes
(prospect/kernels/initialisation.py:1-13).  With the same numpy seed the reference draws the same proposals, so the
run on the CUDA kernel must reproduce the run on the reference's own AnalyticalKernel: every likelihood value the
sampler sees agrees to ~1e-13 relative, hence the same accept decisions and the same bestfits.
"""
import copy
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import ref_harness as rh  # noqa: E402
from prospect_public_b200.toy import annealing_yaml, write_toy_inputs  # noqa: E402
from tests.helpers import golden  # noqa: E402


def _bestfits(sched):
    out = {}
    for t in sched.tasks.done.values():
        if t.type == 'OptimiseTask':
            s, bf = t.optimise_settings, t.optimiser.bestfit
            out[(float(s['fixed_param_val']), s['iteration_number'])] = (
                bf['loglkl'], bf['accepted_steps'], [v[0] for v in bf['position'].values()], s['step_size'])
    return out


@pytest.mark.skipif(not rh.available(), reason='oracle/_ref is absent (python oracle/make_ref.py builds it where '
                                               '/root/reference exists)')
def test_reference_tasks_run_on_the_cuda_kernel(tmp_path, monkeypatch):
    rh.activate()
    import prospect.kernels.initialisation as ref_init
    import prospect.tasks.analyse_profile_task as ref_analyse
    import prospect.tasks.initialise_optimiser_task as ref_iot
    import prospect.tasks.optimise_task as ref_ot
    from prospect_public_b200.kernels import CudaGaussianKernel

    k, m = golden('toy20_kernel.npz'), golden('toy20_mcmc.npz')
    paths = write_toy_inputs(str(tmp_path / 'inp'), k['means'], k['covmat'], [m[f'chain_{i}'] for i in range(4)])

    def job(tag):
        y = annealing_yaml(paths, str(tmp_path / tag), values=[-0.4, 0.5, 1.3], repetitions=1, max_iterations=4,
                           steps_per_iteration=120, plot_profile=False, plot_schedule=False)
        y['run']['mode'] = 'serial'
        y['io']['snapshot_interval'] = 1.0e9
        return y

    plain = rh.run_reference(job('plain'), 4, mode='serial', analyse=True)        # the reference on its own kernel

    calls = {'n': 0}

    def cuda_factory(config_kernel, output_folder, task_id):
        # what the patched prospect/kernels/initialisation.py does for the new kernel type
        calls['n'] += 1
        return CudaGaussianKernel(config_kernel, task_id, output_folder=output_folder)

    for mod in (ref_init, ref_iot, ref_ot, ref_analyse):
        monkeypatch.setattr(mod, 'initialise_kernel', cuda_factory)
    cuda = rh.run_reference(job('cuda'), 4, mode='serial', analyse=True)
    assert calls['n'] >= 3 + 12                                   # every task built its kernel through the factory

    a, b = _bestfits(plain['scheduler']), _bestfits(cuda['scheduler'])
    assert a.keys() == b.keys() and len(a) == 12
    for key in a:
        ll_a, acc_a, x_a, step_a = a[key]
        ll_b, acc_b, x_b, step_b = b[key]
        assert acc_a == acc_b and step_a == step_b, key            # same accept decisions -> same schedule
        assert abs(ll_a - ll_b) <= 1e-11 * max(1.0, abs(ll_a)), key
        assert np.array_equal(np.array(x_a), np.array(x_b)), key  # positions come from the same numpy draws
    # and the reference's analysis task wrote the same profile file from the CUDA-kernel run
    txt = [open(os.path.join(str(tmp_path / tag), 'profile', 'x2.txt')).read() for tag in ('plain', 'cuda')]
    rows = [[[float(t) for t in ln.split('\t')[:5]] for ln in s.strip().split('\n')[1:]] for s in txt]
    assert np.allclose(rows[0], rows[1], rtol=1e-9, atol=1e-12)


@pytest.mark.skipif(not rh.available(), reason='oracle/_ref is absent')
def test_reference_scheduler_drives_the_gpu_batch_executor(tmp_path):
    """Executor protocol (prospect/run.py:27-58): the UNMODIFIED reference's Scheduler --
    its task objects, priorities, emit_tasks / chi2 bookkeeping, AnalyseProfileTask -- runs a 20-D profile job of 10
    values x 8 repetitions x 12 iterations with `GpuBatchExecutor` in the place of its process pool.  Every
    OptimiseTask the scheduler emits comes back from a batched launch (one launch per wave of ready tasks, not one
    per task), the reference's own analysis writes profile/x2.txt from them, and the profile agrees with the analytic
    one and with the package's own driver on the same schedule."""
    rh.activate()
    from oracle import closed_form
    from prospect.communication import Scheduler
    from prospect.input import Configuration
    from prospect.io import prepare_run
    from prospect_public_b200.executor import GpuBatchExecutor
    from prospect_public_b200.run import run as gpu_run

    k, m = golden('toy20_kernel.npz'), golden('toy20_mcmc.npz')
    paths = write_toy_inputs(str(tmp_path / 'inp'), k['means'], k['covmat'], [m[f'chain_{i}'] for i in range(4)])
    kw = dict(repetitions=8, max_iterations=12, temperature_range=[0.1, 1.0e-5], steps_per_iteration=250,
              plot_profile=False, plot_schedule=False)
    y = annealing_yaml(paths, str(tmp_path / 'ref'), **kw)
    y['run']['mode'] = 'serial'                     # (a value the reference's schema accepts; the executor is ours)
    y['io']['snapshot_interval'] = 1.0e9
    with rh.quiet():
        config = Configuration(copy.deepcopy(y))
        prepare_run(config)
        sched = Scheduler(config, False)
        with GpuBatchExecutor() as executor, rh.truncated_annealing(12):
            sched.delegate(executor)
        task = sched.get_analysis_task()
        task.run([t for t in sched.tasks.done.values() if t.id in task.required_task_ids])
    assert len(sched.tasks.failed) == 0
    opt = [t for t in sched.tasks.done.values() if t.type == 'OptimiseTask']
    assert len(opt) == 10 * 8 * 12
    assert executor.tasks_batched == len(opt) and executor.launches <= 3 * 12       # waves, not tasks
    rows = [[float(v) for v in ln.split('\t')[:3]] for ln in
            open(os.path.join(str(tmp_path / 'ref'), 'profile', 'x2.txt')).read().strip().split('\n')[1:]]
    values = np.array([r[0] for r in rows])
    cf = closed_form.profile_curve(k['means'], k['covmat'], 1, values)
    assert np.max(2 * np.abs(np.array([r[2] for r in rows]) - cf)) <= 1e-3
    ours = gpu_run(annealing_yaml(paths, str(tmp_path / 'ours'), **kw))
    assert np.max(2 * np.abs(ours['profile']['loglkls'] - np.array([r[2] for r in rows]))) <= 1e-3
