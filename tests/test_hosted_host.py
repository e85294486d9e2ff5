"""CPU tests of the host side of the batched machinery for host-evaluated likelihoods (prospect_public_b200/hosted.py)
and of its checker, oracle/hosted_model.py (no GPU: no proposal / accept call is made here)."""
import numpy as np
import pytest

from oracle import hosted_model as hm, reference_port as rp
from prospect_public_b200.hosted import HostKernelAdapter
from tests.helpers import ToyHostKernel, golden


def test_adapter_batches_the_reference_style_kernel():
    host = ToyHostKernel(6, box=(-1.0, 1.5))
    kernel = HostKernelAdapter(host, profile_parameter='x3')
    assert kernel.host_evaluated and 'x3' in host.param['fixed'] and 'x3' not in kernel.varying_param_names
    lo, hi = kernel.bounds()
    assert np.array_equal(lo, np.full(5, -1.0)) and np.array_equal(hi, np.full(5, 1.5))
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.9, 1.4, (11, 5))
    fixed = rng.uniform(-1.0, 1.0, 11)
    mask = np.ones(11, dtype=bool)
    mask[[2, 7]] = False
    out = kernel.loglkl_batch(x, fixed, mask)
    assert host.calls == 9 and kernel.evaluations == 9            # masked rows are not evaluated
    for i in range(11):
        expect = host.value(np.insert(x[i], 2, fixed[i])) if mask[i] else np.inf
        assert out[i] == expect
    x[4, 0] = 1.6                                                 # outside the kernel's own prior: rejected unevaluated
    assert kernel.loglkl_batch(x, fixed)[4] == np.inf and host.calls == 9 + 10
    with pytest.raises(ValueError):
        HostKernelAdapter(ToyHostKernel(4), profile_parameter='y1')
    open_kernel = HostKernelAdapter(ToyHostKernel(4, box=(None, 2.0)))
    lo, hi = open_kernel.bounds()
    assert np.all(lo == -np.inf) and np.all(hi == 2.0)
    lo, hi = HostKernelAdapter(ToyHostKernel(4), ignore_prior=True).bounds()
    assert np.all(lo == -np.inf) and np.all(hi == np.inf)


def test_adapter_worker_threads_and_vectorised_callable_agree():
    x = np.random.default_rng(1).uniform(-1, 1, (37, 4))
    fixed = np.linspace(-0.5, 0.5, 37)
    one = HostKernelAdapter(ToyHostKernel(5), profile_parameter='x1').loglkl_batch(x, fixed)
    pool = HostKernelAdapter([ToyHostKernel(5) for _ in range(4)], profile_parameter='x1')
    assert np.array_equal(pool.loglkl_batch(x, fixed), one)
    pool.finalize()
    vec = HostKernelAdapter(ToyHostKernel(5), profile_parameter='x1',
                            batch_loglkl=lambda xs, f: np.where(xs[:, 0] > 0.5, np.nan, 1.0 + f))
    out = vec.loglkl_batch(x, fixed)
    assert np.all(out[x[:, 0] > 0.5] == np.inf) and np.array_equal(out[x[:, 0] <= 0.5], 1.0 + fixed[x[:, 0] <= 0.5])


def test_hosted_model_step_is_the_pinned_ports_step(monkeypatch):
    """oracle/hosted_model.step == oracle/reference_port.mh_step (pinned to the reference's golden vectors) when numpy's
    two draws are the ones the variates stand for: multivariate_normal -> current + step F z, uniform -> exp(-e)."""
    k = golden('toy20_kernel.npz')
    box_hi = np.full(20, 2.5)
    box_hi[4] = k['means'][4] + 0.03                              # a wall next to the mode: some proposals leave the box
    target = rp.GaussianTarget(k['means'], k['covmat'], np.full(20, -2.5), box_hi).fix(1, 0.4)
    n = 19
    rng = np.random.default_rng(2)
    a = rng.standard_normal((n, n))
    cov = 0.003 * (a @ a.T / n + np.eye(n))
    factor = np.linalg.cholesky(cov)
    varying = [i for i in range(20) if i != 1]
    lo, hi = np.full(n, -2.5), box_hi[varying]
    chain = rp.Chain()
    x = k['means'][varying] + 0.02 * rng.standard_normal(n)
    x[3] = k['means'][4] - 0.01
    chain.push(target.loglkl(x), x)
    ll, accepted, outside_seen = chain.loglkls[-1], 0, 0
    for t in range(400):
        z, e = rng.standard_normal(n), rng.exponential()
        monkeypatch.setattr(np.random, 'multivariate_normal', lambda mean, c: mean + 0.4 * (factor @ z))
        monkeypatch.setattr(np.random, 'uniform', lambda low=0, high=1: np.exp(-e))
        acc, prop, llp, outside = hm.step(target.loglkl, x, ll, factor, lo, hi, 0.5, 0.4, z, e)
        assert rp.mh_step(target, chain, cov, 0.5, 0.4) == acc, t
        outside_seen += outside
        if acc:
            x, ll, accepted = prop, llp, accepted + 1
            assert np.array_equal(chain.positions[-1], prop) and chain.loglkls[-1] == llp
    assert 20 < accepted < 380 and outside_seen > 0


def test_external_kernel_types_validate_and_point_to_the_hosted_path(tmp_path):
    """`kernel.type: cobaya | montepython` yaml files are valid input (the reference's schema, base_kernel.py:143-186);
    the factory refuses to build those kernels and says how a host-evaluated kernel is run instead."""
    from prospect_public_b200.config import ConfigError, Configuration
    from prospect_public_b200.kernels import initialise_kernel
    from prospect_public_b200.toy import annealing_yaml, write_toy_inputs
    k = golden('toy20_kernel.npz')
    paths = write_toy_inputs(str(tmp_path / 'inp'), k['means'], k['covmat'], None)
    y = annealing_yaml(paths, str(tmp_path / 'o'), write=False)
    y['kernel']['type'] = 'cobaya'
    cfg = Configuration(y)
    assert cfg.kernel.type == 'cobaya' and cfg.kernel.path == ''
    with pytest.raises(ValueError, match='HostKernelAdapter'):
        initialise_kernel(cfg.kernel, None, 0)
    y['kernel'].update(type='montepython', path=str(tmp_path / 'no_such_dir'))
    with pytest.raises(ConfigError, match="'path'"):
        Configuration(y)
    os_path = tmp_path / 'montepython'
    os_path.mkdir()
    y['kernel'].update(path=str(os_path), conf=paths['kernel_yaml'])
    assert Configuration(y).kernel.type == 'montepython'


def test_hosted_model_step_is_the_reference_step(tmp_path, monkeypatch):
    """The same pin against the UNMODIFIED reference (oracle/_ref or /root/reference): its own AnalyticalKernel and
    MetropolisHastings.step (prospect/mcmc.py:163-190), fed the draws the variates stand for, take the decisions of
    oracle/hosted_model.step and build the same chain -- multiplicities, -log L values and positions, bit for bit."""
    from types import SimpleNamespace
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip('the reference is not available here')
    rh.activate()
    from prospect.kernels.analytical_kernel import AnalyticalKernel
    from prospect.mcmc import MetropolisHastings
    from prospect_public_b200.toy import write_toy_inputs
    k = golden('toy20_kernel.npz')
    paths = write_toy_inputs(str(tmp_path / 'inp'), k['means'], k['covmat'], None)
    (tmp_path / 'out' / 'analytical').mkdir(parents=True)
    with rh.quiet():
        kernel = AnalyticalKernel(SimpleNamespace(param=paths['kernel_yaml'], ignore_prior=False), 0,
                                  output_folder=str(tmp_path / 'out'))
    kernel.set_fixed_parameters({'x2': 0.4})
    names = kernel.varying_param_names
    kernel.param['varying']['x5']['range'] = [-2.5, float(k['means'][4]) + 0.03]    # a wall next to the mode
    n = len(names)
    rng = np.random.default_rng(4)
    a = rng.standard_normal((n, n))
    cov = 0.003 * (a @ a.T / n + np.eye(n))
    factor = np.linalg.cholesky(cov)
    temperature, step_size = 0.5, 0.4
    cfg = SimpleNamespace(start_from_covmat=cov, start_from_position=None, temperature=temperature, step_size=step_size)
    from prospect.mcmc import Chain
    start = {nm: [0.] for nm in names}
    # (the chain is seeded the way SimulatedAnnealing.__init__ does it, optimiser.py:56-62: the reference's
    # get_default_initial_position fails once a parameter is fixed)
    mh = MetropolisHastings(cfg, kernel, chain=Chain([1], [kernel.loglkl(start)], start))
    lo = np.array([kernel.param['varying'][nm]['range'][0] for nm in names], dtype=float)
    hi = np.array([kernel.param['varying'][nm]['range'][1] for nm in names], dtype=float)

    def loglkl(x):
        return kernel.loglkl({nm: [v] for nm, v in zip(names, x)})

    x = np.zeros(n)
    ll = mh.chain.loglkls[-1]
    assert ll == loglkl(x)
    # walk to the mode first (the default start is far out), then count
    accepted = outside_seen = 0
    for t in range(600):
        z, e = rng.standard_normal(n), rng.exponential()
        monkeypatch.setattr(np.random, 'multivariate_normal', lambda mean, c: mean + step_size * (factor @ z))
        monkeypatch.setattr(np.random, 'uniform', lambda low=0, high=1: np.exp(-e))
        rows_before = len(mh.chain.mults)
        acc, prop, llp, outside = hm.step(loglkl, x, ll, factor, lo, hi, temperature, step_size, z, e)
        with rh.quiet():
            mh.step()
        assert (len(mh.chain.mults) == rows_before + 1) == acc, t
        outside_seen += outside
        if acc:
            x, ll, accepted = prop, llp, accepted + 1
            assert mh.chain.loglkls[-1] == llp
            assert np.array_equal([mh.chain.positions[nm][-1] for nm in names], prop)
    assert accepted > 30 and outside_seen > 0 and sum(mh.chain.mults) == 601
