"""CPU checks of the device-algorithm restatement (oracle/device_model.c) against the pinned
reference port, the closed form and the Philox known answers; plus the C-ABI surface of libpb200."""
import os
import re

import numpy as np
import pytest
from scipy import stats

from oracle import closed_form, device_model as dm, reference_port as rp
from prospect_public_b200 import capi
from tests.helpers import toy_profile_problem, toy_free_problem, golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PHILOX_KAT = [   # Random123 kat_vectors, philox4x32-10
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize('ctr,key,expect', PHILOX_KAT)
def test_philox_known_answers(ctr, key, expect):
    assert dm.philox(ctr, key) == expect
    assert capi.philox4x32_10(ctr, key) == expect       # libpb200's host copy of the device function


def test_library_exports_every_declared_symbol():
    """Every function include/pb200.h declares is exported by libpb200.so and bound in capi."""
    with open(os.path.join(ROOT, 'include', 'pb200.h')) as f:
        declared = set(re.findall(r'\b(pb200_[a-z0-9_]+)\s*\(', f.read()))
    assert declared == set(capi.SIGNATURES)
    handle = capi.lib()
    for name in declared:
        assert hasattr(handle, name)
    assert handle.pb200_abi_version() == capi.ABI_VERSION


def test_struct_layouts_match_header():
    """ctypes mirrors of the ABI structs have the field order of the header."""
    with open(os.path.join(ROOT, 'include', 'pb200.h')) as f:
        text = f.read()
    for cname, cls in (('pb200_problem', capi.Problem), ('pb200_chains', capi.Chains),
                       ('pb200_schedule', capi.Schedule), ('pb200_records', capi.Records),
                       ('pb200_samples', capi.Samples), ('pb200_peers', capi.Peers), ('pb200_hosted', capi.Hosted)):
        body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (cname, cname), text, re.S).group(1)
        body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
        fields = [re.findall(r'(\w+)\s*(?:\[\w+\])?\s*;', line)[0] for line in body.split('\n') if ';' in line]
        assert fields == [f[0] for f in cls._fields_], cname
    assert f'#define PB200_MAX_PEERS {capi.MAX_PEERS}' in text


@pytest.mark.parametrize('d', [20, 30])
def test_exact_eval_matches_reference_port(d):
    hp, starts, init_ll, values, k = toy_profile_problem(d)
    m = dm.Model(hp)
    t = rp.GaussianTarget(k['means'], k['covmat'])
    rng = np.random.default_rng(1)
    for p in range(0, len(values), 3):
        assert abs(m.exact_eval(starts[p], p) - init_ll[p]) <= 1e-12 * init_ll[p]
        x = rng.uniform(-1, 1.5, hp.n)
        ref = t.fix(1, values[p]).loglkl(x)
        assert abs(m.exact_eval(x, p) - ref) <= 1e-12 * abs(ref)
        assert abs(hp.loglkl(x, p) - ref) <= 1e-12 * abs(ref)


def test_whitened_increment_is_the_true_loglkl_change():
    """Delta = s z.gt + 1/2 s^2 sum lam z^2 equals loglkl(x + s Lt z) - loglkl(x)."""
    hp, starts, _, values, k = toy_profile_problem(20, points=[3])
    m = dm.Model(hp)
    rng = np.random.default_rng(5)
    n = hp.n
    ll, gt = m.exact_eval(starts[0], 0, with_grad=True)
    lt = hp.lt_t[0, :n, :n].T.astype(np.float64)
    for s in (0.5, 0.01):
        z = rng.standard_normal(n)
        delta = s * z @ gt[:n] + 0.5 * s * s * np.sum(hp.lam[0, :n] * z * z)
        ll2 = m.exact_eval(starts[0] + s * lt @ z, 0)
        assert abs((ll2 - ll) - delta) < 1e-5 * max(abs(delta), 1e-6)


def _schedule_pair(mode, lo, hi, mult, tol=0.0, steps=250):
    py = rp.Schedule(step_size_schedule='exponential' if mode == 0 else 'adaptive',
                     step_size_range=(0.5, 0.05), step_size_adaptive_interval=(lo, hi),
                     step_size_adaptive_multiplier='adaptive' if mode == 2 else mult, steps_per_iteration=steps,
                     chi2_tolerance=tol or None)
    c = dm.DmSchedule(steps, mode, py.temperature_change(), py.initial_step_size_change() if mode == 0 else 1.0,
                      lo, hi, mult, tol)
    return py, c


@pytest.mark.parametrize('mode', [0, 1, 2])
def test_schedule_update_bit_exact_vs_reference_port(mode):
    """A5: the C restatement and the pinned Python port give identical doubles over random AR sequences,
    including the secant edge paths (failure, equal ARs)."""
    rng = np.random.default_rng(mode)
    for trial in range(200):
        py, c = _schedule_pair(mode, 0.19, 0.21, 0.3)
        settings = {'temperature': 0.1, 'temperature_change': py.temperature_change(), 'step_size': 0.37,
                    'step_size_change': py.initial_step_size_change()}
        m = dm.Model.__new__(dm.Model)
        ch = dm.DmChain(None, 0.0, 0.1, 0.37, np.inf, 0.0, 0.0, 0.0, 0, 0, 0, 0, 0, 0)
        for it in range(12):
            acc = int(rng.integers(0, 120)) if trial % 7 else int(rng.choice([49, 50, 51]))
            bf = {'acceptance_rate': np.float64(1 + acc) / np.int64(251), 'loglkl': float(rng.uniform(1, 5))}
            try:
                with np.errstate(all='ignore'):
                    settings = rp.next_settings(py, settings, bf)
                failed = False
            except UnboundLocalError:
                failed = True
            dm.schedule_next(c, ch, acc, bf['loglkl'])
            if failed:
                assert ch.status == 2
                break
            assert ch.status == 0
            assert ch.temperature == settings['temperature']
            assert ch.step_size == settings['step_size'] or (np.isnan(ch.step_size) and np.isnan(settings['step_size']))
            assert ch.current_best == settings['current_best_loglkl']


def test_chi2_tolerance_stop_matches_reference_port():
    py, c = _schedule_pair(1, 0.19, 0.21, 0.3, tol=2.5)
    for cur, best, acc in ((10.0, 9.5, 5), (10.0, 5.0, 5), (10.0, 9.5, 0), (np.inf, 9.5, 5), (9.0, 9.5, 3)):
        ch = dm.DmChain(None, 0.0, 0.1, 0.3, cur, 0.0, 0.0, 0.0, 0, 0, 0, 0, 0, 0)
        dm.schedule_next(c, ch, acc, best)
        expect = rp.converged(py, cur, {'loglkl': best, 'accepted_steps': acc + 1})
        assert (ch.status == 1) == expect


@pytest.mark.parametrize('d,iters,steps', [(20, 80, 250), (30, 400, 500)])
def test_cold_annealing_reaches_closed_form(d, iters, steps):
    """Cold schedule (T 0.1 -> 1e-12): Delta chi^2 <= 1e-3, parameters <= 1e-4 relative (|x| floored at 1e-2).
    The 30-D toy needs far more steps: its bin covariances come from 61 bootstrap rows, so the whitened
    curvature spectrum spans 1.4e-2 .. 32 and the sloppy directions equilibrate slowly (same for the reference)."""
    pts = [0, len(golden(f'toy{d}_init.npz')['values']) // 2]
    hp, starts, init_ll, values, k = toy_profile_problem(d, points=pts)
    m = dm.Model(hp)
    sched = dm.DmSchedule(steps, 1, (1e-12 / 0.1) ** (1 / iters), 1.0, 0.19, 0.21, 0.3, 0.0)
    for p in range(len(pts)):
        ch = m.new_chain(starts[p], p, chain_id=p, temperature=0.1, step_size=0.1, current_best=init_ll[p])
        out = m.anneal_chain(ch, sched, iters, seed=2024)
        cf, xcf = closed_form.profile_minimum(k['means'], k['covmat'], 1, values[p])
        best = int(np.argmin(out['best_loglkl']))
        assert 2 * abs(out['best_loglkl'][best] - cf) <= 1e-3
        assert np.max(np.abs(out['best_x'][best][:hp.n] - xcf) / np.maximum(np.abs(xcf), 1e-2)) <= 1e-4
        assert out['accepted'].min() >= 0 and ch.iteration == iters


def test_mh_stationary_distribution_ks():
    """T=1 free chain samples N(mu, S^-1): KS per coordinate on thinned samples of the model.
    Tolerance: reject at p < 1e-3 / d (Bonferroni)."""
    d = 20
    k = golden('toy20_kernel.npz')
    cov = closed_form.stationary_covariance(k['covmat'])
    from prospect_public_b200.problem import build_problem
    hp = build_problem(k['means'], k['covmat'], [-2.5] * d, [2.5] * d, None, None, cov)
    m = dm.Model(hp)
    ch = m.new_chain(k['means'], 0, chain_id=3, temperature=1.0, step_size=2.4 / np.sqrt(d))
    m.mh_chain(ch, 2000, seed=9)                                     # burn-in
    thin, n_keep = 50, 1500
    nacc, smp = m.mh_chain(ch, thin * n_keep, seed=9, mode=2, thin=thin, capacity=n_keep)
    assert smp['count'][0] == n_keep
    assert 0.15 < nacc / (thin * n_keep) < 0.40
    xs = smp['x'][:, :d]
    sig = np.sqrt(np.diag(cov))
    for i in range(d):
        p = stats.kstest((xs[:, i] - k['means'][i]) / sig[i], 'norm').pvalue
        assert p > 1e-3 / d, (i, p)


def test_mh_recording_accepted_matches_reference_chain_layout():
    """RECORD_ACCEPTED rows == Chain(mults, loglkls, positions): multiplicities sum to steps+1 and the
    recorded loglkl is the Gaussian at the recorded position."""
    hp, k = toy_free_problem(20)
    m = dm.Model(hp)
    ch = m.new_chain(np.zeros(20), 0, chain_id=0, temperature=1.0, step_size=1.0)
    n1, smp = m.mh_chain(ch, 700, seed=1, mode=1, capacity=1500)
    n2, smp = m.mh_chain(ch, 300, seed=1, mode=1, capacity=1500, samples=smp)
    cnt = smp['count'][0]
    assert cnt == 1 + n1 + n2
    assert smp['mult'][:cnt].sum() == 1001
    t = rp.GaussianTarget(k['means'], k['covmat'])
    for row in (0, cnt // 2, cnt - 1):
        ref = t.loglkl(smp['x'][row, :20])
        assert abs(smp['loglkl'][row] - ref) < 1e-5 * max(1.0, ref)
