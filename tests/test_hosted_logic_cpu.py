"""CPU run of the HOST logic of prospect_public_b200/hosted.py (record layout, per-iteration bookkeeping, early stops,
evaluation masks, chain recording and continuation) -- the logic tests/test_gpu_hosted.py checks on a B200.

There is no CPU implementation of the product's kernels, so this module supplies TEST DOUBLES for the five device
entry points the hosted path calls (pb200_hosted_propose, pb200_hosted_accept, pb200_schedule_update,
pb200_rng_variates and the device allocations), written in numpy on top of oracle/device_model (Philox variates,
schedule_next).  They exist only inside the `fake_device` fixture below; nothing under prospect_public_b200/ can reach
them, and the GPU tests remain the parity tests proper.  What is verified here is everything between those calls.
"""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import device_model as dm  # noqa: E402
from prospect_public_b200 import capi, engine, hosted, problem  # noqa: E402


@pytest.fixture
def fake_device(monkeypatch):
    def on_cpu(fn):
        def wrapped(*args, **kwargs):
            kwargs.pop('pin_memory', None)
            if 'device' in kwargs and str(kwargs['device']).startswith('cuda'):
                kwargs['device'] = 'cpu'
            return fn(*args, **kwargs)
        return wrapped

    for name in ('empty', 'zeros', 'full', 'as_tensor'):
        monkeypatch.setattr(torch, name, on_cpu(getattr(torch, name)))

    class Stream:
        cuda_stream = 0

        def synchronize(self):
            pass

    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: Stream())
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, 'current_device', lambda: 0)

    def hosted_problem_init(self, n, covmats, lo, hi, device):
        self.device, self.n, self.npad = 'cuda:0', int(n), problem.pad_of(n)
        covmats = np.asarray(covmats, dtype=np.float64).reshape(-1, n, n)
        self.n_points = covmats.shape[0]
        self.factors = [problem.proposal_factor(c) for c in covmats]
        self.lo, self.hi, self.c = np.asarray(lo), np.asarray(hi), None

    def upload_packed(arrays, device):
        out = {k: torch.as_tensor(np.ascontiguousarray(v).copy()) for k, v in arrays.items()}
        out['_storage'] = None
        return out

    monkeypatch.setattr(hosted.HostedProblem, '__init__', hosted_problem_init)
    monkeypatch.setattr(problem, 'upload_packed', upload_packed)
    monkeypatch.setattr(capi, 'ptr', lambda t: None)
    monkeypatch.setattr(capi, 'Chains', lambda *a: None)
    monkeypatch.setattr(capi, 'Records', lambda *a: None)
    monkeypatch.setattr(engine, 'choose_lanes', lambda npad, b: 32)

    def propose(hp, ch, seed, out_prop, out_outside, stream=None):
        engine._count()
        for c in range(ch.n_chains):
            live = int(ch.status[c]) == capi.ACTIVE
            z, _ = dm.variates(seed, int(ch.chain_id[c]) & 0xFFFFFFFF, int(ch.steps_done[c]), 1, hp.n)
            zz = z[0].astype(np.float64) if live else np.zeros(hp.n)
            prop = ch.x[c, :hp.n].numpy() + float(ch.step_size[c]) * (hp.factors[int(ch.point[c])] @ zz)
            out_prop[c, :hp.n] = torch.as_tensor(prop)
            out_outside[c] = int(live and bool(np.any(prop < hp.lo) or np.any(prop > hp.hi)))

    def accept(hp, ch, seed, prop, prop_ll, outside, flag=None, nacc=None, best_ll=None, best_x=None, stream=None):
        engine._count()
        for c in range(ch.n_chains):
            if int(ch.status[c]) != capi.ACTIVE:
                if flag is not None:
                    flag[c] = 0
                continue
            t = int(ch.steps_done[c])
            _, e = dm.variates(seed, int(ch.chain_id[c]) & 0xFFFFFFFF, t, 1, hp.n)
            ll, llp = float(ch.loglkl[c]), float(prop_ll[c])
            acc = int(outside[c]) == 0 and (llp - ll < float(ch.temperature[c]) * float(e[0]))
            if flag is not None:
                flag[c] = int(acc)
            if acc:
                ch.x[c] = prop[c]
                ch.loglkl[c] = llp
                if nacc is not None:
                    nacc[c] += 1
                if best_ll is not None and llp < float(best_ll[c]):
                    best_ll[c] = llp
                    best_x[c] = prop[c]
            ch.steps_done[c] = t + 1

    def schedule_update(ch, sched, accepted, best, stream=None):
        engine._count()
        ms = dm.DmSchedule(sched.steps_per_iteration, sched.step_mode, sched.temperature_change, sched.step_size_change,
                           sched.adaptive_lo, sched.adaptive_hi, sched.adaptive_multiplier, sched.chi2_tolerance)
        for c in range(ch.n_chains):
            if int(ch.status[c]) != capi.ACTIVE:
                continue
            mc = dm.DmChain(None, 0.0, float(ch.temperature[c]), float(ch.step_size[c]), float(ch.current_best[c]),
                            float(ch.secant_prev_mult[c]), float(ch.secant_cur_mult[c]), float(ch.secant_prev_ar[c]), 0,
                            int(ch.iteration[c]), 0, int(ch.secant_stage[c]), 0, 0)
            dm.schedule_next(ms, mc, int(accepted[c]), float(best[c]))
            ch.temperature[c], ch.step_size[c], ch.current_best[c] = mc.temperature, mc.step_size, mc.current_best
            ch.secant_prev_mult[c], ch.secant_cur_mult[c] = mc.prev_mult, mc.cur_mult
            ch.secant_prev_ar[c], ch.secant_stage[c] = mc.prev_ar, mc.stage
            ch.status[c], ch.iteration[c] = mc.status, mc.iteration

    def rng_variates(device, seed, chain_id, step0, n_steps, n, stream=None):
        z, e = dm.variates(seed, chain_id, step0, n_steps, n)
        return torch.as_tensor(z), torch.as_tensor(e)

    monkeypatch.setattr(engine, 'hosted_propose', propose)
    monkeypatch.setattr(engine, 'hosted_accept', accept)
    monkeypatch.setattr(engine, 'schedule_update', schedule_update)
    monkeypatch.setattr(engine, 'rng_variates', rng_variates)


def _gpu_module():
    import tests.test_gpu_hosted as g
    return g


@pytest.mark.parametrize('mode', ['adaptive', 'exponential', 'tolerance', 'secant'])
def test_annealing_bookkeeping(fake_device, tmp_path, mode):
    _gpu_module().test_hosted_annealing_matches_scalar_replay(tmp_path, mode)


def test_three_ways_to_serve_a_batch(fake_device, tmp_path):
    _gpu_module().test_hosted_vectorised_likelihood_and_worker_threads_agree(tmp_path)


def test_no_box_and_failing_likelihoods(fake_device, tmp_path):
    _gpu_module().test_hosted_global_optimisation_without_box(tmp_path)


def test_mh_sampler_recording_and_continuation(fake_device):
    _gpu_module().test_hosted_mh_sampler_matches_scalar_replay()


def test_reference_kernel_behind_the_optimiser(fake_device, tmp_path):
    g = _gpu_module()
    if not g.rh.available():
        pytest.skip('the reference is not available here')
    g.test_reference_analytical_kernel_behind_the_hosted_optimiser(tmp_path)
