"""GPU parity tests: the sm_100a kernels (called through the C ABI of libpb200.so) against the CPU
oracle on the same seeded inputs.

Bit-level checks replay the GPU's own normal/exponential variates (pb200_rng_variates, the same device
function the samplers call) through oracle/device_model.c, which restates the kernels' arithmetic in
scalar C.  Everything integer (accept counts, multiplicities, iteration/status bookkeeping, argmin
indices) must be identical; FP64 state must be bit-identical as well because both sides perform the
same IEEE operations in the same order.
"""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import closed_form, device_model as dm, reference_port as rp  # noqa: E402
from prospect_public_b200 import capi, engine  # noqa: E402
from prospect_public_b200.problem import DeviceProblem, build_problem  # noqa: E402
from tests.helpers import (conditional_covariance, golden, spd_target, toy_free_problem,  # noqa: E402
                           toy_profile_problem)

DEV = 'cuda:0'
SEED = 0x5EED5EED12345


def _gpu_variates(chain_id, step0, n_steps, n):
    z, e = engine.rng_variates(DEV, SEED, chain_id, step0, n_steps, n)
    torch.cuda.synchronize()
    return z.cpu().numpy(), e.cpu().numpy()


def test_device_is_blackwell_and_library_loaded():
    sm, major, minor = (capi.C.c_int32(), capi.C.c_int32(), capi.C.c_int32())
    capi.check(capi.lib().pb200_device_info(0, sm, major, minor))
    assert major.value == 10 and sm.value >= 100


@pytest.mark.parametrize('n', [19, 29, 32, 40, 255, 256])
def test_rng_variates_match_model(n):
    """Philox words are identical; the fast-math Box-Muller agrees with float64 to ~1e-6."""
    z, e = _gpu_variates(chain_id=11, step0=6, n_steps=203, n=n)
    zm, em = dm.variates(SEED, 11, 6, 203, n)
    assert np.all(np.isfinite(z)) and np.all(e >= 0)
    assert np.max(np.abs(z - zm)) < 2e-5
    assert np.max(np.abs(e - em) / (1 + em)) < 2e-6


def test_rng_variates_are_standard_normal_and_exponential():
    from scipy import stats
    z, e = _gpu_variates(chain_id=5, step0=0, n_steps=40000, n=30)
    assert stats.kstest(z[:, 3], 'norm').pvalue > 1e-4
    assert stats.kstest(z[::7].ravel(), 'norm').pvalue > 1e-4
    assert stats.kstest(e, 'expon').pvalue > 1e-4
    c = np.corrcoef(z[:, :6].T)
    assert np.max(np.abs(c - np.eye(6))) < 0.03
    assert abs(np.corrcoef(z[:-1, 0], z[1:, 0])[0, 1]) < 0.03          # consecutive steps
    z2, _ = _gpu_variates(chain_id=6, step0=0, n_steps=1000, n=30)
    assert not np.array_equal(z[:1000], z2)                              # chains differ
    z3, _ = _gpu_variates(chain_id=5, step0=100, n_steps=50, n=30)
    assert np.array_equal(z3, z[100:150])                                # keyed by absolute step


@pytest.mark.parametrize('d', [20, 30])
def test_loglkl_batch_matches_reference(d):
    """A10/A11: batch chi^2 + prior box == reference port (1e-12 rel) and == device model (bit-exact)."""
    hp, starts, init_ll, values, k = toy_profile_problem(d)
    prob = DeviceProblem(hp, DEV)
    rng = np.random.default_rng(3)
    m_rows = 257
    x = rng.uniform(-1.0, 1.5, (m_rows, hp.npad))
    x[:, hp.n:] = 0.0
    x[5, 2] = 2.6          # outside the box
    x[9, 0] = -2.5000001
    x[10, 0] = -2.5        # on the edge is inside (strict comparison)
    pts = rng.integers(0, hp.n_points, m_rows).astype(np.int32)
    ll, outside = engine.loglkl_batch(prob, torch.as_tensor(x, device=DEV), torch.as_tensor(pts, device=DEV))
    ll, outside = ll.cpu().numpy(), outside.cpu().numpy()
    model = dm.Model(hp)
    t = rp.GaussianTarget(k['means'], k['covmat'])
    for r in range(0, m_rows, 8):
        ref = t.fix(1, values[pts[r]]).loglkl(x[r, :hp.n])
        assert abs(ll[r] - ref) <= 1e-12 * abs(ref)
        assert ll[r] == model.exact_eval(x[r, :hp.n], pts[r])
    expect_out = np.zeros(m_rows, dtype=np.int32)
    expect_out[[5, 9]] = 1
    assert np.array_equal(outside, expect_out)
    # start points of the reference's InitialiseOptimiserTask reproduce its initial_loglkl
    xs = np.zeros((len(values), hp.npad))
    xs[:, :hp.n] = starts
    ll0, _ = engine.loglkl_batch(prob, torch.as_tensor(xs, device=DEV),
                                 torch.arange(len(values), dtype=torch.int32, device=DEV))
    assert np.allclose(ll0.cpu().numpy(), init_ll, rtol=1e-12, atol=0)


def _replay_anneal(hp, starts, init_ll, reps, sched_args, n_iter, launches, mode=capi.STEP_ADAPTIVE_FIXED, lanes=32):
    """Run B = reps * P chains on the GPU and through the scalar model with the GPU's variates."""
    prob = DeviceProblem(hp, DEV)
    p_count = hp.n_points
    b = reps * p_count
    point = np.tile(np.arange(p_count), reps).astype(np.int32)
    x0 = np.tile(starts, (reps, 1))
    cb = np.tile(init_ll, reps)
    chains = engine.ChainBatch(prob, x0, point, temperature=sched_args['t0'], step_size=sched_args['s0'],
                               current_best=cb)
    s = sched_args
    sched = engine.make_schedule(s['S'], s['tc'], mode, s.get('sc', 1.0), s['interval'], s.get('mult', 0.0),
                                 s.get('tol'))
    recs = engine.Records(prob, b, n_iter)
    done = 0
    for chunk in launches:
        engine.anneal_run(prob, chains, sched, recs, chunk, SEED, lanes=lanes)
        done += chunk
    assert done == n_iter
    torch.cuda.synchronize()
    model = dm.Model(hp, lanes)
    msched = dm.DmSchedule(s['S'], mode, s['tc'], s.get('sc', 1.0), s['interval'][0], s['interval'][1],
                           s.get('mult', 0.0), s.get('tol') or 0.0)
    out = {k: getattr(recs, k).cpu().numpy() for k in ('best_loglkl', 'last_loglkl', 'temperature', 'step_size',
                                                        'accepted', 'best_x', 'last_x')}
    final = {k: getattr(chains, k).cpu().numpy() for k in ('x', 'loglkl', 'temperature', 'step_size', 'iteration',
                                                            'status', 'steps_done', 'current_best')}
    for c in range(b):
        z, e = _gpu_variates(c, 0, n_iter * s['S'], hp.n)
        ch = model.new_chain(x0[c], point[c], chain_id=c, temperature=s['t0'], step_size=s['s0'],
                             current_best=cb[c])
        mo = model.anneal_chain(ch, msched, n_iter, SEED, z, e)
        rows = ch.iteration
        assert np.array_equal(out['accepted'][:rows, c], mo['accepted'][:rows]), c
        for key in ('best_loglkl', 'last_loglkl', 'temperature', 'step_size'):
            assert np.array_equal(out[key][:rows, c], mo[key][:rows]), (key, c)
        assert np.array_equal(out['best_x'][:rows, c], mo['best_x'][:rows]), c
        assert np.array_equal(out['last_x'][:rows, c], mo['last_x'][:rows]), c
        assert np.all(out['accepted'][rows:, c] == -1)
        assert final['iteration'][c] == ch.iteration and final['status'][c] == ch.status
        assert final['steps_done'][c] == ch.steps_done
        assert final['temperature'][c] == ch.temperature and final['step_size'][c] == ch.step_size
        assert np.array_equal(final['x'][c], np.ctypeslib.as_array(ch.x, (hp.npad,)))
        assert final['loglkl'][c] == ch.loglkl
    return out, final


def _toy_sched(**kw):
    base = {'S': 250, 'tc': (0.001 / 0.1) ** (1 / 20), 't0': 0.1, 's0': 0.1, 'interval': (0.19, 0.21), 'mult': 0.3}
    base.update(kw)
    return base


@pytest.mark.parametrize('producer_warps', ['0', '1'])
@pytest.mark.parametrize('lanes', [32, 16, 8])
@pytest.mark.parametrize('d', [20, 30])
def test_anneal_bit_exact_vs_model_replay(d, lanes, producer_warps, monkeypatch):
    """K1-style schedule, all profile points x 3 repetitions, 4 iterations in one launch, for every lane-group
    geometry (32 / 16 / 8 lanes per chain; 30 chains do not fill the last warp of the 16- and 8-lane grids), with the
    variates drawn in place (anneal_kernel) and by producer warps through shared memory (anneal_ws_kernel)."""
    monkeypatch.setenv('PB200_WS', producer_warps)
    hp, starts, init_ll, values, k = toy_profile_problem(d, points=None if d == 20 else list(range(0, 32, 4)))
    out, _ = _replay_anneal(hp, starts, init_ll, 3, _toy_sched(), 4, [4], lanes=lanes)
    ar = (1 + out['accepted']) / 251.0
    assert 0.02 < ar.mean() < 0.8


def test_anneal_launch_partitioning_is_invisible():
    """1 launch of 4 iterations == 1+2+1 launches (state round-trips through HBM losslessly)."""
    hp, starts, init_ll, values, k = toy_profile_problem(20, points=[0, 4, 9])
    a, fa = _replay_anneal(hp, starts, init_ll, 2, _toy_sched(S=101), 4, [4], lanes=16)
    b, fb = _replay_anneal(hp, starts, init_ll, 2, _toy_sched(S=101), 4, [1, 2, 1], lanes=16)
    for key in a:
        assert np.array_equal(a[key], b[key]), key
    for key in fa:
        assert np.array_equal(fa[key], fb[key]), key


def test_anneal_exponential_secant_and_tolerance_modes():
    hp, starts, init_ll, values, k = toy_profile_problem(20, points=[2, 6])
    _replay_anneal(hp, starts, init_ll, 2, _toy_sched(s0=0.5, sc=(0.05 / 0.5) ** (1 / 20)), 5, [5],
                   mode=capi.STEP_EXPONENTIAL)
    out, fin = _replay_anneal(hp, starts, init_ll, 4, _toy_sched(s0=0.3, interval=(0.24, 0.26), S=400), 8, [8],
                              mode=capi.STEP_ADAPTIVE_SECANT, lanes=8)
    # chains of one warp stop at different iterations (chi2_tolerance): the others must be unaffected
    _, fin = _replay_anneal(hp, starts, init_ll, 6, _toy_sched(tol=0.02, t0=0.01), 12, [12], lanes=8)
    assert (fin['status'] == capi.CONVERGED).any() and len(np.unique(fin['iteration'])) > 1
    _, fin = _replay_anneal(hp, starts, init_ll, 4, _toy_sched(tol=2.5, t0=0.01), 10, [4, 6], lanes=16)
    assert (fin['status'] == capi.CONVERGED).any()


def _spd_profile_problem(d, n_points):
    mu, cov = spd_target(d, seed=d)
    sig = float(np.sqrt(cov[1, 1]))
    vals = mu[1] + sig * np.linspace(-0.8, 0.9, n_points)
    prop = conditional_covariance(cov, 1)
    hp = build_problem(mu, cov, [-2.5] * d, [2.5] * d, 1, vals, np.stack([prop] * n_points))
    starts = np.stack([np.delete(mu, 1) + 0.01 * (1 - 2 * (p % 2)) for p in range(n_points)])
    init_ll = np.array([hp.loglkl(starts[p], p) for p in range(n_points)])
    return hp, starts, init_ll


@pytest.mark.parametrize('path', ['fused', 'wide_simt', 'wide_simt_producer_warps'])
def test_anneal_wide_path_bit_exact_npad64_and_256(path, monkeypatch):
    """CPL > 1 geometries (n = 40 -> npad 64; n = 127 -> npad 128; n = 255 -> npad 256) on synthetic SPD targets, bit for
    bit against the scalar oracle: the single fused kernel (PB200_WIDE=0) and the per-iteration pipeline of the wide
    path -- steps kernel, position GEMM on the SIMT pipe (PB200_WIDE_MMA=0), FP64 anchor GEMM -- which keeps every
    summation order of the fused kernel; its steps kernel with the variates drawn in place and by producer warps."""
    monkeypatch.setenv('PB200_WIDE', '0' if path == 'fused' else '1')
    monkeypatch.setenv('PB200_WIDE_MMA', '0')
    monkeypatch.setenv('PB200_WIDE_WS', '1' if path.endswith('producer_warps') else '0')
    for d, reps, s_steps in ((41, 3, 60), (128, 3, 40), (256, 2, 24)):
        hp, starts, init_ll = _spd_profile_problem(d, 2)
        _replay_anneal(hp, starts, init_ll, reps, _toy_sched(S=s_steps, s0=2.4 / np.sqrt(d)), 3, [2, 1])
        if d == 41:
            _replay_anneal(hp, starts, init_ll, reps, _toy_sched(S=s_steps, s0=2.4 / np.sqrt(d)), 3, [3], lanes=16)


def test_wide_path_tile_edges_and_schedule_modes_bit_exact(monkeypatch):
    """Wide pipeline with ragged GEMM tiles (19 repetitions: one full 16-row tile + 3 rows; 3 points), chains that stop
    at different iterations (chi2_tolerance), and the secant multiplier -- all bit-exact against the oracle."""
    monkeypatch.setenv('PB200_WIDE_MMA', '0')
    hp, starts, init_ll = _spd_profile_problem(130, 3)
    s0 = 2.4 / np.sqrt(130)
    _replay_anneal(hp, starts, init_ll, 19, _toy_sched(S=30, s0=s0), 3, [1, 2])
    _, fin = _replay_anneal(hp, starts, init_ll, 5, _toy_sched(S=40, s0=s0, tol=30.0, t0=0.01), 6, [6])
    assert (fin['status'] == capi.CONVERGED).any() and len(np.unique(fin['iteration'])) > 1
    _replay_anneal(hp, starts, init_ll, 4, _toy_sched(S=50, s0=s0, interval=(0.24, 0.26)), 5, [5],
                   mode=capi.STEP_ADAPTIVE_SECANT)


def _wide_mma_run(hp, starts, init_ll, reps, s, n_iter):
    prob = DeviceProblem(hp, DEV)
    b = reps * hp.n_points
    point = np.tile(np.arange(hp.n_points), reps).astype(np.int32)
    x0, cb = np.tile(starts, (reps, 1)), np.tile(init_ll, reps)
    chains = engine.ChainBatch(prob, x0, point, temperature=s['t0'], step_size=s['s0'], current_best=cb)
    sched = engine.make_schedule(s['S'], s['tc'], capi.STEP_ADAPTIVE_FIXED, 1.0, s['interval'], s['mult'])
    recs = engine.Records(prob, b, n_iter)
    engine.anneal_run(prob, chains, sched, recs, n_iter, SEED)
    engine.launch_status()
    out = {k: getattr(recs, k).cpu().numpy() for k in ('best_loglkl', 'last_loglkl', 'temperature', 'step_size',
                                                        'accepted', 'best_x', 'last_x')}
    return out, x0, point, cb


def _check_against_per_iteration_replay(hp, out, x0, point, cb, s, n_iter, chains_to_check):
    """Replay iteration k with the scalar oracle FROM THE GPU's own position after iteration k - 1 (GPU variates)."""
    model = dm.Model(hp, 32)
    msched = dm.DmSchedule(s['S'], capi.STEP_ADAPTIVE_FIXED, s['tc'], 1.0, s['interval'][0], s['interval'][1],
                           s['mult'], 0.0)
    scale = np.abs(np.asarray(hp.lt_t, dtype=np.float64)).sum(axis=1).max()      # bound of the |Lt| row sums
    for c in chains_to_check:
        ch = model.new_chain(x0[c], point[c], chain_id=c, temperature=s['t0'], step_size=s['s0'],
                             current_best=cb[c])
        for k in range(n_iter):
            z, e = _gpu_variates(c, k * s['S'], s['S'], hp.n)
            mo = model.anneal_chain(ch, msched, 1, SEED, z, e)
            assert out['accepted'][k, c] == mo['accepted'][0], (c, k)
            assert out['temperature'][k, c] == mo['temperature'][0] and out['step_size'][k, c] == mo['step_size'][0]
            for key in ('best_x', 'last_x'):
                err = np.abs(out[key][k, c] - mo[key][0]).max()
                assert err <= 3e-6 * scale * (1.0 + np.abs(mo[key][0]).max()), (key, c, k, err)
            assert out['best_loglkl'][k, c] == model.exact_eval(out['best_x'][k, c, :hp.n], point[c])
            assert out['last_loglkl'][k, c] == model.exact_eval(out['last_x'][k, c, :hp.n], point[c])
            xs = np.ctypeslib.as_array(ch.x, (hp.npad,))          # continue the oracle from where the GPU actually is
            xs[:] = out['last_x'][k, c]
            ch.current_best = out['best_loglkl'][k, c]


@pytest.mark.parametrize('d,reps', [(256, 5), (128, 130)])
def test_wide_path_tensor_core_positions(d, reps, monkeypatch):
    """The tcgen05 (3xTF32, TMA-fed) position GEMM of the wide path.  Tensor-core accumulation order is not
    specified, so positions cannot be bit-compared; everything else can: iteration k is replayed by the scalar oracle
    FROM THE GPU's own position after iteration k - 1 with the GPU's variates --
      * accept counts, temperatures and step sizes are bit-identical (the FP64 anchors at the GPU's positions are),
      * the written chi^2 values are the exact FP64 values at the written positions,
      * the GPU's positions agree with the oracle's FP32-SIMT ones to 3e-6 of the displacement scale (3xTF32 keeps
        ~21 bits per product; tolerance stated here).
    130 repetitions: a full 128-row MMA tile plus a 2-row one; 5 repetitions: a TMA box smaller than the tile."""
    monkeypatch.setenv('PB200_WIDE_MMA', '1')
    hp, starts, init_ll = _spd_profile_problem(d, 2)
    s = _toy_sched(S=32, s0=2.4 / np.sqrt(d))
    out, x0, point, cb = _wide_mma_run(hp, starts, init_ll, reps, s, 3)
    b = reps * hp.n_points
    _check_against_per_iteration_replay(hp, out, x0, point, cb, s, 3, list(range(0, b, max(1, b // 24))) + [b - 1])


@pytest.mark.parametrize('groups', ['2', '3', '8'])
def test_wide_path_sub_batches_on_parallel_streams_change_nothing(groups, monkeypatch):
    """PB200_WIDE_SUB: the per-iteration pipeline of the wide path issued as groups of profile points on parallel
    streams (the default from 4096 chains up).  Chains are independent, so every record -- positions from the tensor-core
    GEMM included -- equals the single-stream run bit for bit; 5 points in 2 / 3 / 5 (8 requested) uneven groups."""
    monkeypatch.setenv('PB200_WIDE_MMA', '1')
    hp, starts, init_ll = _spd_profile_problem(128, 5)
    s = _toy_sched(S=37, s0=2.4 / np.sqrt(128))
    monkeypatch.setenv('PB200_WIDE_SUB', '1')
    one, *_ = _wide_mma_run(hp, starts, init_ll, 20, s, 4)
    monkeypatch.setenv('PB200_WIDE_SUB', groups)
    many, *_ = _wide_mma_run(hp, starts, init_ll, 20, s, 4)
    assert (one['accepted'] >= 0).all()
    for key in one:
        assert np.array_equal(one[key], many[key]), key


def test_k4_shape_256d_8_points_x_128_repetitions():
    """BASELINE.json configs[3] at its per-GPU shape: 256-D synthetic SPD Gaussian profile, 8 points x 128
    repetitions = 1024 chains, 20 iterations x 250 steps through the default path (wide pipeline, tensor-core
    position GEMM).  32 chains spread over all points are replayed iteration by iteration against the scalar oracle;
    for ALL chains the recorded chi^2 is the exact FP64 value at the recorded position (pb200_loglkl_batch), the
    prior box is respected, and every point's best repetition closes in on the analytic profile minimum."""
    d, n_points, reps, n_iter = 256, 8, 128, 20
    hp, starts, init_ll = _spd_profile_problem(d, n_points)
    s = _toy_sched(S=250, s0=0.15)
    out, x0, point, cb = _wide_mma_run(hp, starts, init_ll, reps, s, n_iter)
    b = reps * n_points
    assert (out['accepted'] >= 0).all()
    rng = np.random.default_rng(0)
    _check_against_per_iteration_replay(hp, out, x0, point, cb, s, n_iter, sorted(rng.choice(b, 32, replace=False)))
    prob = DeviceProblem(hp, DEV)
    for k in (0, n_iter - 1):
        ll, outside = engine.loglkl_batch(prob, torch.as_tensor(out['best_x'][k], device=DEV),
                                          torch.as_tensor(point, device=DEV))
        assert np.array_equal(ll.cpu().numpy(), out['best_loglkl'][k]) and not outside.cpu().numpy().any()
    mu, cov = spd_target(d, seed=d)
    best = out['best_loglkl'].min(axis=0).reshape(reps, n_points).min(axis=0)
    t_final = s['t0'] * s['tc'] ** (n_iter - 1)
    for p in range(n_points):
        exact = closed_form.profile_minimum(mu, cov, 1, hp.fixed_values[p])[0]
        assert 0.0 <= best[p] - exact <= 1.5 * hp.n * t_final, (p, best[p] - exact)      # a sample at T sits ~ n/2 T above the minimum


@pytest.mark.parametrize('lanes', [32, 8])
def test_prior_box_rejections_bit_exact(lanes):
    """Start next to the box edge with a huge step and a flat target (T = 1e6): the deferred box test must fall
    back to exact materialisation, reject what leaves the box and re-reference what stays inside."""
    hp, k = toy_free_problem(20)
    x0 = np.full((4, 20), 2.45)
    ll0 = np.array([hp.loglkl(x0[0])] * 4)
    out, _ = _replay_anneal(hp, x0[:1], ll0[:1], 5, _toy_sched(t0=1e6, s0=3.0, S=300, mult=0.0), 2, [2], lanes=lanes)
    assert out['accepted'].min() >= 0 and out['accepted'].max() < 300
    assert np.all(out['last_x'][..., :20] <= 2.5) and np.all(out['last_x'][..., :20] >= -2.5)
    assert np.all(out['best_x'][..., :20] <= 2.5) and np.all(out['best_x'][..., :20] >= -2.5)


def test_box_is_never_left_by_any_visited_point():
    """Every recorded point of an MH chain that keeps bouncing off the box edge lies inside the box."""
    hp, k = toy_free_problem(20)
    prob = DeviceProblem(hp, DEV)
    chains = engine.ChainBatch(prob, np.full((64, 20), -2.4), 0, temperature=1e6, step_size=0.3)
    smp = engine.Samples(prob, 64, 401, capi.RECORD_THINNED, 1)
    acc = engine.mh_run(prob, chains, 400, SEED, smp, lanes=16).cpu().numpy()
    x = smp.x.cpu().numpy()[:, :400, :20]
    assert acc.min() > 0 and acc.max() < 400
    assert x.min() >= -2.5 and x.max() <= 2.5


def test_schedule_update_kernel_bit_exact():
    """The stand-alone bookkeeping kernel == scalar model for random histories, all three modes."""
    hp, starts, init_ll, values, k = toy_profile_problem(20, points=[0])
    prob = DeviceProblem(hp, DEV)
    rng = np.random.default_rng(8)
    b = 512
    for mode in (0, 1, 2):
        chains = engine.ChainBatch(prob, np.tile(starts, (b, 1)), 0, temperature=0.1, step_size=0.37,
                                   current_best=30.0)
        sched = engine.make_schedule(250, 0.8, mode, 0.9, (0.19, 0.21), 0.3, 2.5 if mode == 1 else None)
        msched = dm.DmSchedule(250, mode, 0.8, 0.9, 0.19, 0.21, 0.3, 2.5 if mode == 1 else 0.0)
        mch = [dm.DmChain(None, 0.0, 0.1, 0.37, 30.0, 0.0, 0.0, 0.0, 0, 0, 0, 0, 0, 0) for _ in range(b)]
        for it in range(6):
            acc = rng.integers(0, 110, b).astype(np.int32)
            acc[::9] = 50                                   # AR inside the interval / equal ARs
            best = rng.uniform(1.0, 30.0, b)
            engine.schedule_update(chains, sched, torch.as_tensor(acc, device=DEV), torch.as_tensor(best, device=DEV))
            for c in range(b):
                if mch[c].status == 0:
                    dm.schedule_next(msched, mch[c], int(acc[c]), float(best[c]))
        torch.cuda.synchronize()
        for key, attr in (('temperature', 'temperature'), ('step_size', 'step_size'), ('status', 'status'),
                          ('iteration', 'iteration'), ('current_best', 'current_best'),
                          ('secant_stage', 'stage')):
            got = getattr(chains, key).cpu().numpy()
            exp = np.array([getattr(mc, attr) for mc in mch])
            assert np.array_equal(got, exp, equal_nan=True) if got.dtype.kind == 'f' else np.array_equal(got, exp), \
                (mode, key)


def test_profile_reduce_matches_numpy():
    """A20 sort_tasks: first minimum over iterations per chain, first-minimum repetition, mean of finite."""
    hp, starts, init_ll, values, k = toy_profile_problem(20, points=[0])
    prob = DeviceProblem(hp, DEV)
    rng = np.random.default_rng(4)
    p_count, reps, iters = 7, 45, 9
    b = p_count * reps
    recs = engine.Records(prob, b, iters, keep_positions=False)
    best = rng.uniform(1, 2, (iters, b)).round(2)       # rounding creates exact ties
    acc = rng.integers(0, 50, (iters, b)).astype(np.int32)
    acc[5:, ::4] = -1                                   # chains that stopped early
    acc[:, 3] = -1                                      # a chain that never ran
    acc[:, [2 + p_count * r for r in range(reps)]] = -1  # a point where nothing finished
    recs.best_loglkl.copy_(torch.as_tensor(best))
    recs.accepted.copy_(torch.as_tensor(acc))
    out = {k_: v.cpu().numpy() for k_, v in engine.profile_reduce(recs, iters, p_count, reps).items()}
    masked = np.where(acc >= 0, best, np.inf)
    cb = masked.min(axis=0)
    ci = np.where(np.isfinite(cb), masked.argmin(axis=0), -1)
    assert np.array_equal(out['chain_best'], cb) and np.array_equal(out['chain_iter'], ci)
    per = cb.reshape(reps, p_count)
    assert np.array_equal(out['point_best'], per.min(axis=0))
    assert np.array_equal(out['point_rep'], per.argmin(axis=0))
    avg = np.array([per[np.isfinite(per[:, j]), j].mean() if np.isfinite(per[:, j]).any() else np.inf
                    for j in range(p_count)])
    assert np.allclose(out['point_avg'], avg, rtol=1e-14)


@pytest.mark.parametrize('n,rows', [(20, 2477), (30, 100000), (70, 5000)])
def test_weighted_moments_match_np_cov(n, rows):
    """A16: np.cov(bias=True, fweights=mults) from the FP64 moment reduction."""
    rng = np.random.default_rng(n)
    x = rng.normal(0.5, 0.2, (rows, n + 3))             # ld > n
    w = rng.integers(1, 40, rows).astype(np.int32)
    sw, swx, swxx = engine.weighted_moments(torch.as_tensor(x, device=DEV), n, torch.as_tensor(w, device=DEV))
    cov = engine.covariance_from_moments(sw, swx, swxx).cpu().numpy()
    ref = np.cov(x[:, :n], rowvar=False, bias=True, fweights=w)
    assert sw.item() == w.sum()
    assert np.allclose(cov, ref, rtol=1e-9, atol=1e-13)
    sw1, _, _ = engine.weighted_moments(torch.as_tensor(x, device=DEV), n)
    assert sw1.item() == rows


def test_weighted_moments_are_reproducible_bit_for_bit():
    """Many CTAs contribute to every sum; their partial sums are added in CTA order (no atomics), so repeated launches --
    and with them the bin covariances, proposal factors and every FP64 result built on them -- agree in every bit."""
    rng = np.random.default_rng(12)
    x = torch.as_tensor(rng.standard_normal((40000, 32)) * rng.uniform(1e-3, 1e3, 32), device=DEV)
    w = torch.as_tensor(rng.integers(1, 9, 40000).astype(np.int32), device=DEV)
    first = [t.clone() for t in engine.weighted_moments(x, 30, w)]
    for _ in range(5):
        again = engine.weighted_moments(x, 30, w)
        assert all(torch.equal(a, b) for a, b in zip(first, again))


def test_weighted_moments_on_the_shipped_bootstrap_chain():
    """The covariance the reference's InitialiseOptimiserTask derives for x2 = 0.0 (golden fixture)."""
    from tests.helpers import mcmc_rows
    rows = mcmc_rows(20)
    ini = golden('toy20_init.npz')
    i = 3
    sub = rows[ini['indices'][i]]
    sw, swx, swxx = engine.weighted_moments(torch.as_tensor(np.ascontiguousarray(sub[:, 2:]), device=DEV), 20,
                                            torch.as_tensor(sub[:, 0].astype(np.int32), device=DEV))
    cov = engine.covariance_from_moments(sw, swx, swxx).cpu().numpy()
    cov = np.delete(np.delete(cov, 1, 0), 1, 1)
    assert np.allclose(cov, ini['covmat'][i], rtol=1e-9, atol=1e-14)


@pytest.mark.parametrize('lanes', [32, 16, 8])
@pytest.mark.parametrize('mode', [capi.RECORD_ACCEPTED, capi.RECORD_THINNED])
def test_mh_recording_bit_exact_vs_model_replay(mode, lanes):
    """jobtype mcmc: reference defaults (start 0, covmat 0.005*diag(width), T = step = 1), two launches."""
    hp, k = toy_free_problem(20)
    prob = DeviceProblem(hp, DEV)
    b, n1, n2, cap, thin = 6, 700, 333, 1100, 7
    chains = engine.ChainBatch(prob, np.zeros((b, 20)), 0, temperature=1.0, step_size=1.0)
    smp = engine.Samples(prob, b, cap, mode, thin)
    a1 = engine.mh_run(prob, chains, n1, SEED, smp, lanes=lanes).cpu().numpy()
    a2 = engine.mh_run(prob, chains, n2, SEED, smp, lanes=lanes).cpu().numpy()
    model = dm.Model(hp, lanes)
    cnt = smp.count.cpu().numpy()
    gx, gl, gm = smp.x.cpu().numpy(), smp.loglkl.cpu().numpy(), smp.mult.cpu().numpy()
    for c in range(b):
        z, e = _gpu_variates(c, 0, n1 + n2, 20)
        ch = model.new_chain(np.zeros(20), 0, chain_id=c, temperature=1.0, step_size=1.0)
        m1, s = model.mh_chain(ch, n1, SEED, z, e, mode, thin, cap)
        m2, s = model.mh_chain(ch, n2, SEED, z[n1:], e[n1:], mode, thin, cap, samples=s)
        assert (a1[c], a2[c]) == (m1, m2)
        assert cnt[c] == s['count'][0]
        r = cnt[c]
        assert np.array_equal(gm[c, :r], s['mult'][:r])
        assert np.array_equal(gx[c, :r], s['x'][:r])
        assert np.array_equal(gl[c, :r], s['loglkl'][:r])
        if mode == capi.RECORD_ACCEPTED:
            assert gm[c, :r].sum() == n1 + n2 + 1           # sum(mults) = steps + 1 (optimise_task.py:27)
            assert r == 1 + a1[c] + a2[c]
    assert np.array_equal(chains.steps_done.cpu().numpy(), np.full(b, n1 + n2))


def test_misaligned_chains_in_one_warp_take_the_generic_path():
    """Chains whose step counters differ modulo 4 share a warp (8 lanes per chain): results still equal the
    per-chain model (every step redraws its Philox block)."""
    hp, k = toy_free_problem(20)
    prob = DeviceProblem(hp, DEV)
    b, n_steps = 8, 203
    offs = np.array([0, 1, 2, 3, 5, 7, 4, 8], dtype=np.int32)
    chains = engine.ChainBatch(prob, np.zeros((b, 20)), 0, temperature=1.0, step_size=1.0, steps_done=offs)
    acc = engine.mh_run(prob, chains, n_steps, SEED, lanes=8).cpu().numpy()
    model = dm.Model(hp, 8)
    xs = chains.x.cpu().numpy()
    for c in range(b):
        z, e = _gpu_variates(c, int(offs[c]), n_steps, 20)
        ch = model.new_chain(np.zeros(20), 0, chain_id=c, temperature=1.0, step_size=1.0, steps_done=int(offs[c]))
        m, _ = model.mh_chain(ch, n_steps, SEED, z, e)
        assert acc[c] == m
        assert np.array_equal(xs[c], np.ctypeslib.as_array(ch.x, (hp.npad,)))


def test_lane_choice_follows_batch_size():
    # npad = 32: more lanes per chain (more warps) for the small, latency-bound batches, 8 lanes once the GPU is filled
    assert engine.choose_lanes(32, 10) == 32 and engine.choose_lanes(32, 1024) == 32
    assert engine.choose_lanes(32, 2048) == 16
    assert engine.choose_lanes(32, 4096) == 8 and engine.choose_lanes(32, 65536) == 8
    assert engine.choose_lanes(64, 65536) == 16 and engine.choose_lanes(64, 100) == 32
    assert engine.choose_lanes(256, 65536) == 32
    assert engine.uses_producer_warps(8, 1024) and not engine.uses_producer_warps(8, 4096)
    assert not engine.uses_producer_warps(32, 1024)


def test_mh_sample_overflow_is_reported():
    hp, k = toy_free_problem(20)
    prob = DeviceProblem(hp, DEV)
    chains = engine.ChainBatch(prob, np.zeros((2, 20)), 0)
    smp = engine.Samples(prob, 2, 8, capi.RECORD_THINNED, 1)
    engine.mh_run(prob, chains, 50, SEED, smp)
    assert np.all(smp.count.cpu().numpy() == 50)            # > capacity: rows were dropped, never written OOB


def test_secant_multiplier_failure_path_stops_the_chain():
    """'adaptive' multiplier with an interval that contains every acceptance rate: the reference hits its
    unbound-local (optimiser.py:99-103) after the first iteration; here the chain gets status FAILED, keeps its
    first record and never runs again -- identically in the kernel and in the model."""
    hp, starts, init_ll, values, k = toy_profile_problem(20, points=[1, 5])
    out, fin = _replay_anneal(hp, starts, init_ll, 3, _toy_sched(interval=(0.0, 1.0)), 4, [2, 2],
                              mode=capi.STEP_ADAPTIVE_SECANT, lanes=16)
    assert np.all(fin['status'] == capi.FAILED) and np.all(fin['iteration'] == 1)
    assert np.all(out['accepted'][0] >= 0) and np.all(out['accepted'][1:] == -1)


@pytest.mark.parametrize('lanes', [32, 8])
def test_producer_warps_early_stop_box_and_wide_paths(lanes, monkeypatch):
    """anneal_ws_kernel on the paths that stress its hand-shake: chains of one warp converging at different
    iterations (the consumer leaves early and must release its producer), split launches (ring restarts at every
    block alignment), exact box tests, and a multi-coordinate-per-lane geometry."""
    monkeypatch.setenv('PB200_WS', '1')
    hp, starts, init_ll, values, k = toy_profile_problem(20, points=[2, 6])
    _, fin = _replay_anneal(hp, starts, init_ll, 6, _toy_sched(tol=0.02, t0=0.01), 12, [5, 7], lanes=lanes)
    assert (fin['status'] == capi.CONVERGED).any() and len(np.unique(fin['iteration'])) > 1
    _, fin = _replay_anneal(hp, starts, init_ll, 2, _toy_sched(tol=1e9), 6, [6], lanes=lanes)     # all stop at once
    assert np.all(fin['status'] == capi.CONVERGED) and np.all(fin['iteration'] == 1)
    hpf, _ = toy_free_problem(20)
    x0 = np.full((1, 20), 2.45)
    _replay_anneal(hpf, x0, np.array([hpf.loglkl(x0[0])]), 5, _toy_sched(t0=1e6, s0=3.0, S=300, mult=0.0), 2, [1, 1],
                   lanes=lanes)
    if lanes == 32:
        mu, cov = spd_target(41, seed=41)
        vals = np.array([mu[1] - 0.05, mu[1] + 0.08])
        prop = conditional_covariance(cov, 1)
        hpw = build_problem(mu, cov, [-2.5] * 41, [2.5] * 41, 1, vals, np.stack([prop, prop]))
        st = np.stack([np.delete(mu, 1) + 0.01, np.delete(mu, 1) - 0.01])
        ll = np.array([hpw.loglkl(st[p], p) for p in range(2)])
        _replay_anneal(hpw, st, ll, 3, _toy_sched(S=61, s0=2.4 / np.sqrt(41)), 3, [3], lanes=16)


def test_producer_warps_fall_back_for_misaligned_chains(monkeypatch):
    """Chains of one warp whose step counters differ modulo 4 cannot share a producer: that pair draws in place and
    gives the same results as the plain kernel."""
    hp, starts, init_ll, values, k = toy_profile_problem(20, points=[0, 3, 5, 8])
    prob = DeviceProblem(hp, DEV)
    b = 12
    point = np.tile(np.arange(4), 3).astype(np.int32)
    offs = np.array([0, 1, 2, 3, 4, 4, 8, 12, 5, 7, 4, 8], dtype=np.int32)       # warp 0 mixed, warp 1 aligned
    res = {}
    for ws in ('0', '1'):
        monkeypatch.setenv('PB200_WS', ws)
        chains = engine.ChainBatch(prob, np.tile(starts, (3, 1)), point, temperature=0.1, step_size=0.1,
                                   current_best=np.tile(init_ll, 3), steps_done=offs)
        recs = engine.Records(prob, b, 3)
        sched = engine.make_schedule(101, 0.8, capi.STEP_ADAPTIVE_FIXED, 1.0, (0.19, 0.21), 0.3, None)
        engine.anneal_run(prob, chains, sched, recs, 3, SEED, lanes=8)
        torch.cuda.synchronize()
        res[ws] = [getattr(recs, key).cpu().numpy() for key in ('best_loglkl', 'last_loglkl', 'accepted', 'best_x')]
        res[ws].append(chains.steps_done.cpu().numpy())
    for a, c in zip(res['0'], res['1']):
        assert np.array_equal(a, c)
    assert np.array_equal(res['1'][-1], offs + 303)


def test_chain_moments_bit_exact_and_statistics_match_numpy():
    """pb200_chain_moments: per-chain sums in row order are bit-identical to a sequential numpy loop; the flat
    statistics vector equals its numpy restatement and yields np.cov of the pooled chains."""
    from prospect_public_b200.run import convergence_from_statistics
    hp, k = toy_free_problem(20)
    prob = DeviceProblem(hp, DEV)
    b, cap, n = 37, 61, hp.n
    rng = np.random.default_rng(8)
    smp = engine.Samples(prob, b, cap, capi.RECORD_ACCEPTED, 1)
    cnt = rng.integers(0, cap + 5, b).astype(np.int32)          # some empty, some "overflowed" (count > capacity)
    cnt[3] = 0
    x = np.zeros((b, cap, hp.npad))
    mult = np.zeros((b, cap), dtype=np.int32)
    for c in range(b):
        r = min(cnt[c], cap)
        x[c, :r, :n] = rng.normal(0.3, 1.0, (r, n))
        mult[c, :r] = rng.integers(1, 9, r)
    smp.x.copy_(torch.as_tensor(x))
    smp.mult.copy_(torch.as_tensor(mult))
    smp.count.copy_(torch.as_tensor(cnt))
    sw, swx, scaled = (t.cpu().numpy() for t in engine.chain_moments(smp, n))
    for c in range(b):
        acc, w_acc = np.zeros(n), 0.0
        for r in range(min(cnt[c], cap)):
            w_acc = w_acc + float(mult[c, r])
            acc = acc + float(mult[c, r]) * x[c, r, :n]
        assert sw[c] == w_acc and np.array_equal(swx[c], acc)
        assert np.array_equal(scaled[c], acc * (1.0 / np.sqrt(w_acc)) if w_acc > 0 else np.zeros(n))
    stats = engine.chain_statistics(smp, n).cpu().numpy()
    keep = [c for c in range(b) if cnt[c] > 0]
    xs = [x[c, :min(cnt[c], cap), :n] for c in keep]
    ws = [mult[c, :min(cnt[c], cap)].astype(np.float64) for c in keep]
    allx, allw = np.concatenate(xs), np.concatenate(ws).astype(int)
    rm1, mean, cov = convergence_from_statistics(stats, n)
    assert stats[-1] == len(keep) and stats[0] == allw.sum()
    assert np.allclose(cov, np.cov(allx, rowvar=False, bias=True, fweights=allw), rtol=1e-10, atol=1e-12)
    assert np.allclose(mean, np.average(allx, axis=0, weights=allw), rtol=1e-12)


@pytest.mark.parametrize('lanes,tol', [(32, None), (8, None), (8, 0.02), (16, 1e9)])
def test_signalled_launch_ships_complete_iterations(lanes, tol):
    """pb200_anneal_run_signalled + pb200_stream_wait_geq: a side stream that waits on iteration i's counter sees
    that iteration's complete record block while the launch is still running; results equal the plain launch; the
    counters also complete when chains stop early (chi2_tolerance), so the side stream can never be left waiting."""
    hp, starts, init_ll, values, k = toy_profile_problem(20, points=[2, 6, 7])
    prob = DeviceProblem(hp, DEV)
    reps, n_iter = 5, 6
    b = reps * hp.n_points
    point = np.tile(np.arange(hp.n_points), reps).astype(np.int32)
    sched = engine.make_schedule(250, 0.7, capi.STEP_ADAPTIVE_FIXED, 1.0, (0.19, 0.21), 0.3, tol)
    out = {}
    for mode in ('plain', 'signalled'):
        chains = engine.ChainBatch(prob, np.tile(starts, (reps, 1)), point, temperature=0.01 if tol else 0.1,
                                   step_size=0.1, current_best=np.tile(init_ll, reps))
        recs = engine.Records(prob, b, n_iter)
        if mode == 'plain':
            engine.anneal_run(prob, chains, sched, recs, n_iter, SEED, lanes=lanes)
            torch.cuda.synchronize()
        else:
            side = torch.cuda.Stream()
            progress = torch.zeros(n_iter, dtype=torch.int32, device=DEV)
            snap = torch.zeros_like(recs.buf)
            torch.cuda.synchronize()
            target = engine.anneal_run_signalled(prob, chains, sched, recs, n_iter, SEED, progress, lanes=lanes)
            with torch.cuda.stream(side):
                for i in range(n_iter):
                    engine.stream_wait_geq(side, progress, i, target)
                    snap[i].copy_(recs.buf[i], non_blocking=True)
            torch.cuda.synchronize()
            assert np.all(progress.cpu().numpy() == target)
            assert target == -(-b // (32 // lanes))
            assert torch.equal(snap.view(torch.int64), recs.buf.view(torch.int64))    # bit patterns (-1 counts are NaNs)
        out[mode] = (recs.buf.clone(), chains.x.clone(), chains.status.clone(), chains.iteration.clone())
    for a, c in zip(out['plain'], out['signalled']):
        assert torch.equal(a.view(torch.int64) if a.dtype == torch.float64 else a,
                           c.view(torch.int64) if c.dtype == torch.float64 else c)
    if tol == 0.02:
        assert len(np.unique(out['plain'][3].cpu().numpy())) > 1        # chains stopped at different iterations


@pytest.mark.parametrize('lanes', [32, 8])
def test_buffers_are_never_written_out_of_bounds(lanes):
    """Canary check (compute-sanitizer is not available on this pool): sample and record buffers sit inside larger
    allocations filled with a sentinel; chains overflow the sample capacity and run more iterations than the
    records have rows; 5 chains leave the last warp of every geometry partly empty.  Nothing outside the buffers
    changes and the overflow is reported through the counters."""
    hp, k = toy_free_problem(20)
    prob = DeviceProblem(hp, DEV)
    b, cap, npad, pad, mark = 5, 7, hp.npad, 96, 777
    for mode in (capi.RECORD_ACCEPTED, capi.RECORD_THINNED):
        xs = torch.full((pad + b * cap * npad + pad,), float(mark), dtype=torch.float64, device=DEV)
        ll = torch.full((pad + b * cap + pad,), float(mark), dtype=torch.float64, device=DEV)
        mu = torch.full((pad + b * cap + pad,), mark, dtype=torch.int32, device=DEV)
        cn = torch.full((pad + b + pad,), mark, dtype=torch.int32, device=DEV)
        cn[pad:pad + b] = 0
        smp = engine.Samples(prob, b, cap, mode, 1)
        smp.c = capi.Samples(mode, 1, cap, capi.C.c_void_p(xs.data_ptr() + 8 * pad), capi.C.c_void_p(ll.data_ptr() + 8 * pad),
                             capi.C.c_void_p(mu.data_ptr() + 4 * pad), capi.C.c_void_p(cn.data_ptr() + 4 * pad))
        chains = engine.ChainBatch(prob, np.zeros((b, 20)), 0, temperature=1.0, step_size=0.5)
        engine.mh_run(prob, chains, 60, SEED, smp, lanes=lanes)
        torch.cuda.synchronize()
        for buf, n_in in ((xs, b * cap * npad), (ll, b * cap), (mu, b * cap), (cn, b)):
            assert torch.all(buf[:pad] == mark) and torch.all(buf[pad + n_in:] == mark)
        assert torch.all(cn[pad:pad + b] > cap)                  # rows were dropped and counted, not written
    # records: 2 rows allocated, 5 iterations run
    tmpl = engine.Records(prob, b, 2)
    stride = tmpl.stride
    big = torch.full(((2 + 2) * stride,), float(mark), dtype=torch.float64, device=DEV)
    big_last = torch.full(((2 + 2) * stride,), float(mark), dtype=torch.float64, device=DEV)
    base = big.data_ptr() + 8 * stride
    addr = lambda name: capi.C.c_void_p(base + 8 * tmpl.offsets[name])
    tmpl.c = capi.Records(2, stride, addr('best_loglkl'), addr('last_loglkl'), addr('temperature'), addr('step_size'),
                          addr('accepted'), addr('best_x'), capi.C.c_void_p(big_last.data_ptr() + 8 * stride))
    chains = engine.ChainBatch(prob, np.zeros((b, 20)), 0, temperature=1.0, step_size=0.5)
    sched = engine.make_schedule(37, 0.9, capi.STEP_ADAPTIVE_FIXED, 1.0, (0.19, 0.21), 0.3, None)
    engine.anneal_run(prob, chains, sched, tmpl, 5, SEED, lanes=lanes)
    torch.cuda.synchronize()
    for buf in (big, big_last):
        assert torch.all(buf[:stride] == mark) and torch.all(buf[3 * stride:] == mark)
    rows = big[stride:3 * stride].view(2, stride)
    assert torch.all(rows[:, :b] != mark)                        # both allocated rows were filled
    assert np.all(chains.iteration.cpu().numpy() == 5) and np.all(chains.steps_done.cpu().numpy() == 5 * 37)


@pytest.mark.parametrize('lanes', [8, 32])
def test_fused_exchange_fills_every_ranks_buffer(lanes):
    """pb200_anneal_run_shared with two emulated ranks on one GPU (both "peer" buffers are local tensors): every
    rank's kernel stores its per-iteration scalars into slot `rank` of BOTH buffers and bumps BOTH arrival counters;
    afterwards each buffer holds the all-gathered rows, equal to the ranks' own records, and every counter equals
    the number of chain warps of the whole job -- also when chains stop early (chi2_tolerance)."""
    hp, starts, init_ll, values, k = toy_profile_problem(20, points=[2, 6, 7])
    prob = DeviceProblem(hp, DEV)
    n_iter, world = 5, 2
    shards = [(np.array([0, 1, 2, 0, 1], dtype=np.int32), 5), (np.array([2, 0, 1, 2], dtype=np.int32), 4)]   # ragged
    n_alloc = 5
    rank_stride = 4 * n_alloc + (n_alloc + 1) // 2
    bufs = [torch.full((n_iter, world, rank_stride), 123.0, dtype=torch.float64, device=DEV) for _ in range(world)]
    progs = [torch.zeros(64, dtype=torch.int32, device=DEV) for _ in range(world)]
    sched = engine.make_schedule(101, 0.7, capi.STEP_ADAPTIVE_FIXED, 1.0, (0.19, 0.21), 0.3, 2.5)
    recs, targets = [], []
    for rank, (pts, b) in enumerate(shards):
        chains = engine.ChainBatch(prob, starts[pts], pts, chain_id=np.arange(b) + 10 * rank, temperature=0.01,
                                   step_size=0.1, current_best=init_ll[pts])
        rec = engine.Records(prob, b, n_iter, n_alloc=n_alloc)
        peers = engine.make_peers(rank, n_alloc, n_iter, rank_stride, [t.data_ptr() for t in bufs],
                                  [t.data_ptr() for t in progs])
        targets.append(engine.anneal_run_shared(prob, chains, sched, rec, n_iter, SEED, peers, lanes=lanes))
        recs.append(rec)
    torch.cuda.synchronize()
    assert targets == [-(-5 // (32 // lanes)), -(-4 // (32 // lanes))]
    for q in range(world):
        pq = progs[q].cpu().numpy()
        assert pq[n_iter - 1] == sum(targets) and pq.sum() == sum(targets)      # one arrival counter per launch
        g = bufs[q].cpu().numpy()
        for rank, (pts, b) in enumerate(shards):
            dec = engine.Records.decode(g[:, rank], b, n_alloc, hp.npad, False)
            acc = recs[rank].accepted.cpu().numpy()
            ran = acc >= 0
            assert np.array_equal(dec['accepted'][ran], acc[ran])
            for key in ('best_loglkl', 'last_loglkl', 'temperature', 'step_size'):
                assert np.array_equal(dec[key][ran], getattr(recs[rank], key).cpu().numpy()[ran]), key
            assert np.all(dec['best_loglkl'][~ran] == 123.0)        # rows of stopped chains are left to the prefill
    assert any((r.accepted.cpu().numpy() < 0).any() for r in recs)   # the tolerance did stop some chains early


@pytest.mark.parametrize('lanes', [32, 8])
@pytest.mark.parametrize('mode', [capi.RECORD_NONE, capi.RECORD_THINNED])
def test_exact_box_mode_bit_exact_and_same_decisions(mode, lanes):
    """PB200_EXACT_BOX (box test at every accepted step instead of the deferred safe-radius test): bit-exact against
    the model in the same mode, also for chains bouncing off the box edge; away from the box the accept / reject
    sequence is the one of the deferred test and the positions agree to FP32 accumulation error."""
    hp, k = toy_free_problem(20)
    prob = DeviceProblem(hp, DEV)
    b, n_steps, cap, thin = 6, 900, 200, 7
    model = dm.Model(hp, lanes)
    for start, step in ((0.0, 1.0), (2.45, 0.3)):
        res = {}
        for exact in (True, False):
            chains = engine.ChainBatch(prob, np.full((b, 20), start), 0, temperature=1.0, step_size=step)
            smp = engine.Samples(prob, b, cap, mode, thin) if mode else None
            acc = engine.mh_run(prob, chains, n_steps, SEED, smp, lanes=lanes, exact_box=exact).cpu().numpy()
            res[exact] = (acc, chains.x.cpu().numpy(), chains.loglkl.cpu().numpy())
            if not exact:
                continue
            gx = smp.x.cpu().numpy() if mode else None
            for c in range(b):
                z, e = _gpu_variates(c, 0, n_steps, 20)
                ch = model.new_chain(np.full(20, start), 0, chain_id=c, temperature=1.0, step_size=step)
                m, smod = model.mh_chain(ch, n_steps, SEED, z, e, mode | capi.EXACT_BOX, thin, cap)
                assert acc[c] == m
                assert np.array_equal(res[True][1][c], np.ctypeslib.as_array(ch.x, (hp.npad,)))
                assert res[True][2][c] == ch.loglkl
                if mode:
                    r = smod['count'][0]
                    assert smp.count.cpu().numpy()[c] == r and np.array_equal(gx[c, :r], smod['x'][:r])
        if start == 0.0:                                    # far from the box: identical decisions
            assert np.array_equal(res[True][0], res[False][0])
            assert np.max(np.abs(res[True][1] - res[False][1])) < 1e-5
        else:
            assert np.all(np.abs(res[True][1][:, :20]) <= 2.5)
