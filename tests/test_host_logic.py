"""CPU tests of the host-side mirror of the reference interfaces: yaml schema, analysis writers,
bin/index bookkeeping, chain container, sharding + gather (gloo, world size 2)."""
import json
import os

import numpy as np
import pytest

from prospect_public_b200 import analysis
from prospect_public_b200.config import ConfigError, Configuration
from prospect_public_b200.distributed import global_chain_index, shard_grid
from prospect_public_b200.initialise import get_fraction_of_points_in_bin, load_mcmc
from prospect_public_b200.mcmc import Chain, collapse_chains
from prospect_public_b200.toy import annealing_yaml, mcmc_yaml, write_toy_inputs
from tests.helpers import GOLDEN, golden


@pytest.fixture()
def toy20(tmp_path):
    k, m = golden('toy20_kernel.npz'), golden('toy20_mcmc.npz')
    return write_toy_inputs(str(tmp_path / 'inp'), k['means'], k['covmat'], [m[f'chain_{i}'] for i in range(4)])


def _plain(section):
    return {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in section.items()}


def test_schema_defaults_match_reference(toy20, tmp_path):
    """Configuration resolves the shipped example yamls to the same dicts as prospect.input.Configuration."""
    with open(os.path.join(GOLDEN, 'config_defaults.json')) as f:
        gold = json.load(f)
    shipped_profile = {'parameter': 'x2', 'values': 'np.linspace(-1.0, 2.0, 10)', 'plot_profile': True,
                       'detailed_plot': False, 'plot_Delta_chi2': False, 'plot_schedule': True,
                       'confidence_levels': [0.68, 0.95]}
    y = annealing_yaml(toy20, str(tmp_path / 'o'), **shipped_profile)
    c = Configuration(y)
    g = gold['example_toy']
    for key, val in g['profile'].items():
        if key == 'start_from_mcmc':
            continue
        assert _plain(c.config_dict['profile'])[key] == val, key
    assert set(c.config_dict['profile']) == set(g['profile'])
    for sec in ('io', 'run', 'kernel'):
        assert set(c.config_dict[sec]) == set(g[sec]), sec
    assert c.config_dict['kernel']['ignore_prior'] is False and c.config_dict['io']['snapshot_interval'] == 10.0
    # global optimisation: parameter None, values []
    y = annealing_yaml(toy20, str(tmp_path / 'o2'), jobtype='global_optimisation', repetitions=3, plot_schedule=True)
    del y['profile']['plot_profile'], y['profile']['start_bin_fraction']      # not in the shipped yaml -> defaults
    c = Configuration(y)
    for key, val in gold['example_global']['profile'].items():
        if key != 'start_from_mcmc':
            assert _plain(c.config_dict['profile'])[key] == val, key
    c = Configuration(mcmc_yaml(toy20, str(tmp_path / 'o3')))
    for key, val in gold['mcmc']['mcmc'].items():
        assert c.config_dict['mcmc'][key] == val, key
    assert c.mcmc.temperature == 1.0 and c.mcmc.step_size == 1.0


def test_schema_rejections(toy20, tmp_path):
    y = annealing_yaml(toy20, str(tmp_path / 'o'))
    del y['profile']['max_iterations']
    with pytest.raises(ConfigError):
        Configuration(y)                                     # mandatory key
    y = annealing_yaml(toy20, str(tmp_path / 'o'), bogus=1)
    with pytest.raises(ConfigError):
        Configuration(y)                                     # unknown key
    y = annealing_yaml(toy20, str(tmp_path / 'o'), repetitions='3')
    with pytest.raises(ConfigError):
        Configuration(y)                                     # wrong type
    y = annealing_yaml(toy20, str(tmp_path / 'o'), step_size_schedule='exponential')
    with pytest.raises(ConfigError):
        Configuration(y)                                     # exponential needs step_size_range
    y = annealing_yaml(toy20, str(tmp_path / 'o'), values='__import__("os").getcwd()')
    with pytest.raises(ConfigError):
        Configuration(y)                                     # only numpy in evaluated strings
    y = annealing_yaml(toy20, str(tmp_path / 'o'))
    y['run']['jobtype'] = 'nope'
    with pytest.raises(ConfigError):
        Configuration(y)
    y = annealing_yaml(toy20, str(tmp_path / 'o'))
    y['kernel']['param'] = 'does/not/exist.yaml'
    with pytest.raises(ConfigError):
        Configuration(y)


def test_output_dir_is_not_clobbered(toy20, tmp_path):
    out = tmp_path / 'run'
    out.mkdir()
    y = annealing_yaml(toy20, str(out))
    y['io']['overwrite_dir'] = False
    assert Configuration(y).io.dir == str(out) + '_0'
    (tmp_path / 'run_0').mkdir()
    assert Configuration(y).io.dir == str(out) + '_1'


def test_adaptive_initial_sets_step_size_range(toy20, tmp_path):
    c = Configuration(annealing_yaml(toy20, str(tmp_path / 'o')))
    assert c.profile.step_size_range == [0.1, None]
    assert isinstance(c.profile.values, np.ndarray) and len(c.profile.values) == 10
    assert c.profile.values[3] == np.linspace(-1.0, 2.0, 10)[3]        # exact float64 keys


def _record_from_golden(gold, n_var, reps, values):
    opt = gold['optimise']
    iters = 1 + max(o['iteration_number'] for o in opt)
    p_count = max(1, len(values))
    b = p_count * reps
    best = np.full((iters, b), np.inf)
    acc = np.full((iters, b), -1, dtype=np.int32)
    bx = np.zeros((iters, b, n_var))
    for o in opt:
        p = 0 if not len(values) else int(np.nonzero(values == o['fixed_param_val'])[0][0])
        c = o['repetition_number'] * p_count + p
        k = o['iteration_number']
        best[k, c] = o['bestfit_loglkl']
        acc[k, c] = o['accepted_steps'] - 1
        bx[k, c] = o['bestfit_position'][:n_var]
    ini = sorted(gold['initialise'], key=lambda t: t['id'])[:p_count]
    names = opt[0]['position_names'][:n_var]
    return analysis.RunRecord(values, reps, names, best, acc, bx, np.zeros(b, dtype=bool),
                              np.array([t['initial_loglkl'] for t in ini]),
                              np.array([t['initial_position'][:n_var] for t in ini]),
                              np.array([t['initial_covmat'] for t in ini]))


@pytest.mark.parametrize('d,nvals', [(20, 10), (30, 4)])
def test_profile_files_identical_to_reference(tmp_path, d, nvals):
    """A20: fed with the reference's own task results, the writers reproduce profile/x2.txt,
    x2_status.txt and x2_intervals.txt byte for byte."""
    with open(os.path.join(GOLDEN, f'toy{d}_profile_run.json')) as f:
        gold = json.load(f)
    rec = _record_from_golden(gold, d - 1, 1, np.linspace(-1.0, 2.0, nvals))
    analysis.analyse_profile(rec, str(tmp_path), 'x2', [0.68, 0.95])
    for fn in ('x2.txt', 'x2_status.txt', 'x2_intervals.txt'):
        assert (tmp_path / fn).read_text() == gold['files'][fn], fn


def test_global_results_file_identical_to_reference(tmp_path):
    with open(os.path.join(GOLDEN, 'toy20_global_run.json')) as f:
        gold = json.load(f)
    rec = _record_from_golden(gold, 20, 3, np.zeros(0))
    analysis.analyse_global_optimisation(rec, str(tmp_path))
    assert (tmp_path / 'results.txt').read_text() == gold['files']['results.txt']


def test_load_profile_round_trip(tmp_path):
    from prospect_public_b200.io import load_profile
    with open(os.path.join(GOLDEN, 'toy20_profile_run.json')) as f:
        gold = json.load(f)
    p = tmp_path / 'x2.txt'
    p.write_text(gold['files']['x2.txt'])
    prof = load_profile(str(p), direct_txt=True)
    assert list(prof)[:5] == ['x2', 'Delta_chi2', '-loglkl', 'avg_best_loglkl', 'initial_loglkl']
    assert len(prof['x2']) == 10 and prof['x2'][0] == -1.0 and len(prof) == 5 + 19


def test_parabola_vertex():
    x, f = analysis.bestfit_parabola([0.0, 1.0, 3.0], [2.0 + 1.5 * (t - 1.2) ** 2 for t in (0.0, 1.0, 3.0)])
    assert abs(x - 1.2) < 1e-14 and abs(f - 2.0) < 1e-14


def test_unfinished_points_and_ties():
    """No finished optimisation -> initial_loglkl enters the profile and the row says so; ties -> first rep."""
    vals = np.array([0.0, 1.0, 2.0])
    best = np.array([[3.0, np.inf, 5.0, 3.0, np.inf, 4.0]])       # reps=2, P=3
    acc = np.array([[4, -1, 2, 4, -1, 2]], dtype=np.int32)
    rec = analysis.RunRecord(vals, 2, ['a'], best, acc, np.zeros((1, 6, 1)), np.zeros(6, dtype=bool),
                             np.array([9.0, 2.5, 9.0]), np.zeros((3, 1)), np.zeros((3, 1, 1)))
    srt = analysis.sort_bestfits(rec)
    assert list(srt['best_rep']) == [0, 0, 1]
    assert srt['avg_loglkl'][1] == np.inf and srt['avg_loglkl'][2] == 4.5
    prof = analysis.compute_profile(rec, srt)
    assert prof['loglkls'][1] == 2.5


def test_bin_selection_bit_exact(toy20):
    """A14: the rows picked for every profile value equal the reference's (golden indices)."""
    chain = load_mcmc(toy20['mcmc_root'])
    ini = golden('toy20_init.npz')
    assert len(chain.mults) == 2477
    for i, v in enumerate(ini['values']):
        idx = get_fraction_of_points_in_bin(chain, v, 'x2', 0.1)
        assert len(idx) == ini['N_points'][i] == 248
        assert np.array_equal(idx, ini['indices'][i])
        best = idx[int(np.argmin(np.take(chain.loglkls, idx)))]
        assert np.array_equal(np.delete(np.array([chain.positions[f'x{j}'][best] for j in range(1, 21)]), 1),
                              ini['position'][i][:-1])


def test_chain_container_semantics():
    ch = Chain([1], [5.0], {'a': [0.1], 'b': [0.2]})
    ch.push_position(4.0, {'a': [0.3], 'b': [0.4]})
    ch.mults[-1] += 2
    assert ch.N == 2 and ch.last_position == {'a': [0.3], 'b': [0.4]}
    assert np.array_equal(ch.data, np.array([[1, 5.0, 0.1, 0.2], [3, 4.0, 0.3, 0.4]]))
    sub = ch.from_indices([1])
    assert sub.mults == [3] and sub[0] == {'a': [0.3], 'b': [0.4]}
    tot = collapse_chains([ch, sub])
    assert tot.N == 3 and tot.mults == [1, 3, 3]


def test_shard_grid_covers_every_chain_once():
    for p, r, w in ((32, 256, 8), (10, 3, 4), (1, 4096, 8), (3, 5, 2), (7, 2, 8)):
        seen = []
        for rank in range(w):
            pt, rep = shard_grid(p, r, rank, w)
            seen.append(global_chain_index(pt, rep, p))
            if p >= w:
                assert len(np.unique(pt)) * r == len(pt)      # whole points stay on one rank
        allc = np.concatenate(seen)
        assert len(allc) == p * r and np.array_equal(np.sort(allc), np.arange(p * r))


def _gloo_worker(rank, world, port, p, r, q):
    """What SimulatedAnnealing.run_schedule / global_records do for N > 1, on CPU tensors over gloo: every rank
    fills its per-iteration record block, all-gathers it, and decodes all blocks into global chain order."""
    import torch
    import torch.distributed as dist
    from prospect_public_b200.engine import Records
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    npad, n_alloc = 32, max(len(shard_grid(p, r, k, world)[0]) for k in range(world))
    pt, rep = shard_grid(p, r, rank, world)
    gid = global_chain_index(pt, rep, p)
    b = len(gid)
    acc_d = (n_alloc + 1) // 2
    stride = 4 * n_alloc + acc_d + n_alloc * npad
    block = torch.zeros(stride, dtype=torch.float64)
    block[0:b] = torch.as_tensor(gid, dtype=torch.float64)                       # best_loglkl := global id
    block[3 * n_alloc:3 * n_alloc + b] = float(rank)                             # step_size := owner rank
    block[4 * n_alloc:4 * n_alloc + acc_d].view(torch.int32)[:b] = torch.as_tensor(2 * gid, dtype=torch.int32)
    bx = block[4 * n_alloc + acc_d:4 * n_alloc + acc_d + n_alloc * npad].view(n_alloc, npad)
    bx[:b, 0] = torch.as_tensor(gid, dtype=torch.float64) + 0.5
    gathered = torch.zeros((world, stride), dtype=torch.float64)
    dist.all_gather_into_tensor(gathered.view(-1), block)
    g = gathered.numpy()
    out = np.zeros((p * r, 4))
    for k in range(world):
        kp, kr = shard_grid(p, r, k, world)
        ids = global_chain_index(kp, kr, p)
        dec = Records.decode(g[k], len(ids), n_alloc, npad)
        sc = Records.decode(g[k][:4 * n_alloc + acc_d], len(ids), n_alloc, npad)       # scalar prefix only
        assert 'best_x' not in sc and np.array_equal(sc['accepted'], dec['accepted'])
        out[ids, 0], out[ids, 1], out[ids, 2] = dec['best_loglkl'], dec['accepted'], dec['step_size']
        out[ids, 3] = dec['best_x'][:, 0]
    q.put((rank, out))
    dist.destroy_process_group()


@pytest.mark.parametrize('p,r', [(5, 3), (1, 7)])
def test_record_all_gather_world2_gloo(p, r):
    """N > 1 path on CPU: two gloo ranks shard the grid and all-gather their records into global order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300) + p
    procs = [ctx.Process(target=_gloo_worker, args=(rank, 2, port, p, r, q)) for rank in range(2)]
    for pr in procs:
        pr.start()
    outs = dict(q.get(timeout=120) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
    for rank in range(2):
        assert np.array_equal(outs[rank][:, 0], np.arange(p * r))
        assert np.array_equal(outs[rank][:, 1], 2.0 * np.arange(p * r))
        assert np.array_equal(outs[rank][:, 3], np.arange(p * r) + 0.5)
    owners = outs[0][:, 2]
    assert set(owners) == {0.0, 1.0}


def test_status_file_with_ragged_and_empty_chains(tmp_path):
    """<param>_status.txt (analyse_profile_task.py:62-89) when chains stopped at different iterations or never ran:
    the vectorised writer gives the text of the plain per-chain loop."""
    rng = np.random.default_rng(2)
    p_count, reps, k_rows, n = 3, 4, 6, 5
    b = p_count * reps
    best = rng.uniform(1, 9, (k_rows, b))
    acc = rng.integers(0, 50, (k_rows, b)).astype(np.int32)
    stop = rng.integers(0, k_rows + 1, b)                    # chain c ran iterations [0, stop[c])
    stop[[1, 7]] = 0                                         # two chains never ran
    for c in range(b):
        acc[stop[c]:, c] = -1
    conv = rng.random(b) < 0.3
    rec = analysis.RunRecord(np.array([0.5, -1.0, 2.0]), reps, [f'x{i}' for i in range(n)], best, acc,
                             rng.normal(size=(k_rows, b, n)), conv, np.ones(p_count), np.zeros((p_count, n)),
                             np.tile(np.eye(n), (p_count, 1, 1)))
    srt = analysis.sort_bestfits(rec)
    path = analysis.write_profile_status(rec, srt, str(tmp_path), 'x2')
    ran = acc >= 0
    expect = [f"{'x2':14.10} ", '\t\t | rep. no.  current iter \t current loglkl \t best iter \t best loglkl \t converged? ']
    for p in np.argsort(rec.values, kind='stable'):
        expect.append(f'\n{rec.values[p]:14.10} ')
        for rep in range(reps):
            c = rep * p_count + p
            if not ran[:, c].any():
                continue
            cur = int(np.nonzero(ran[:, c])[0][-1])
            line = (f"| rep {rep + 1} \t {cur:<13} \t {best[cur, c]:.10e} \t {srt['chain_iter'][rep, p]:<9} \t "
                    f"{srt['chain_best'][rep, p]:.10e} \t {bool(conv[c])} ")
            expect.append(f'\t {line}' if rep == 0 else f'\n\t\t\t\t {line}')
    assert open(path).read() == ''.join(expect)


def _chain_statistics_numpy(xs, ws):
    """numpy restatement of engine.chain_statistics for a list of chains (rows [r, n], integer weights [r])."""
    n = xs[0].shape[1]
    out = np.zeros(2 + n + 2 * n * n)
    for x, w in zip(xs, ws):
        sw, swx = w.sum(), (w[:, None] * x).sum(0)
        out[0] += sw
        out[1:1 + n] += swx
        out[1 + n:1 + n + n * n] += ((w[:, None] * x).T @ x).ravel()
        out[1 + n + n * n:1 + n + 2 * n * n] += np.outer(swx, swx).ravel() / sw
        out[-1] += 1
    return out


def _fake_chains(n_chains, n, seed):
    rng = np.random.default_rng(seed)
    xs = [rng.normal(0.1 * c, 1.0 + 0.05 * c, (40 + 3 * c, n)) for c in range(n_chains)]
    ws = [rng.integers(1, 6, len(x)).astype(np.float64) for x in xs]
    return xs, ws


def test_convergence_from_statistics_equals_per_chain_formula():
    """The summed sufficient statistics give the same R-1, mean and covariance as the per-chain formulation, and
    the pooled covariance equals np.cov(..., bias=True, fweights=mults) (prospect/mcmc.py:93-97)."""
    from prospect_public_b200.run import convergence_from_statistics, gelman_rubin, shard_chains
    xs, ws = _fake_chains(7, 5, seed=3)
    rm1, mean, cov = convergence_from_statistics(_chain_statistics_numpy(xs, ws), 5)
    means = np.array([np.average(x, axis=0, weights=w) for x, w in zip(xs, ws)])
    covs = np.array([np.cov(x, rowvar=False, bias=True, fweights=w.astype(int)) for x, w in zip(xs, ws)])
    assert rm1 == pytest.approx(gelman_rubin(means, covs, [w.sum() for w in ws]), rel=1e-10)
    allx, allw = np.concatenate(xs), np.concatenate(ws).astype(int)
    assert np.allclose(cov, np.cov(allx, rowvar=False, bias=True, fweights=allw), rtol=1e-11, atol=1e-13)
    assert np.allclose(mean, np.average(allx, axis=0, weights=allw), rtol=1e-12)
    assert convergence_from_statistics(_chain_statistics_numpy(xs[:1], ws[:1]), 5)[0] == 0.0
    for b, world in ((7, 2), (3, 8), (65536, 8)):
        blocks = [shard_chains(b, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == b
        assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
        assert max(hi - lo for lo, hi in blocks) - min(hi - lo for lo, hi in blocks) <= 1


def _gloo_stats_worker(rank, world, port, q):
    """jobtype mcmc for N > 1 on CPU tensors: every rank reduces the statistics of ITS chain block, one all-reduce
    makes the global vector, and every rank derives the same R-1 (run.run_mcmc)."""
    import torch
    import torch.distributed as dist
    from prospect_public_b200.run import convergence_from_statistics, shard_chains
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    xs, ws = _fake_chains(5, 4, seed=11)
    lo, hi = shard_chains(5, rank, world)
    stats = torch.as_tensor(_chain_statistics_numpy(xs[lo:hi], ws[lo:hi]))
    dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    q.put((rank, convergence_from_statistics(stats.numpy(), 4)))
    dist.destroy_process_group()


def test_mcmc_statistics_all_reduce_world2_gloo():
    import torch.multiprocessing as mp
    from prospect_public_b200.run import convergence_from_statistics
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29950 + (os.getpid() % 40)
    procs = [ctx.Process(target=_gloo_stats_worker, args=(rank, 2, port, q)) for rank in range(2)]
    for pr in procs:
        pr.start()
    outs = dict(q.get(timeout=120) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
    xs, ws = _fake_chains(5, 4, seed=11)
    rm1, mean, cov = convergence_from_statistics(_chain_statistics_numpy(xs, ws), 4)
    for rank in range(2):
        assert outs[rank][0] == pytest.approx(rm1, rel=1e-10)
        assert np.allclose(outs[rank][1], mean, rtol=1e-12) and np.allclose(outs[rank][2], cov, rtol=1e-10)
    assert outs[0][0] == outs[1][0]


REFERENCE = '/root/reference'


@pytest.mark.parametrize('writer', ['objects', 'bulk'])
@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='the reference checkout only exists in the build container')
def test_reference_reads_our_snapshot(toy20, tmp_path, writer):
    """N1: config.pkl / state.pkl written by compat.export_reference_snapshot (Python objects) and by
    compat.export_reference_snapshot_bulk (pickle byte stream assembled with numpy, used for large runs) are consumed
    by the UNMODIFIED reference (`prospect.run.load` -> `Scheduler.get_analysis_task().run`,
    `prospect.profile.load_profile`) and give its own profile/x2.txt back."""
    import subprocess
    import sys
    from prospect_public_b200 import compat
    from prospect_public_b200.io import prepare_run
    with open(os.path.join(GOLDEN, 'toy20_profile_run.json')) as f:
        gold = json.load(f)
    out = tmp_path / 'run'
    cfg = Configuration(annealing_yaml(toy20, str(out), plot_profile=False, plot_schedule=False))
    prepare_run(cfg)
    rec = _record_from_golden(gold, 19, 1, np.linspace(-1.0, 2.0, 10))
    rec.temperature = np.full(rec.best_loglkl.shape, 0.1)
    rec.step_size = np.full(rec.best_loglkl.shape, 0.1)
    if writer == 'objects':
        n_tasks = compat.export_reference_snapshot(cfg, rec, str(out))
        assert n_tasks == 10 + 200
    else:                                  # compact selection: each chain's best and last iteration
        n_tasks = compat.export_reference_snapshot_bulk(cfg, rec, str(out))
        assert 10 + 10 <= n_tasks <= 10 + 20
    script = (
        "import sys, json, numpy as np\n"
        "from prospect.run import load\n"
        "from prospect.profile import load_profile\n"
        f"folder = {str(out)!r}\n"
        "snap = load(folder)\n"
        # prospect.run.analyse itself raises KeyError for jobtype 'profile' (if / if / elif / else chain,
        # prospect/run.py:95-114); do what it intends
        "task = snap.get_analysis_task()\n"
        "task.run([t for t in snap.tasks.done.values() if t.id in task.required_task_ids])\n"
        "prof = load_profile(folder)\n"
        "types = sorted({t.type for t in snap.tasks.done.values()})\n"
        "print('RESULT', json.dumps({'n': len(prof['x2']), 'types': types, 'ntasks': len(snap.tasks.done)}))\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(GOLDEN, '_shims'), REFERENCE]))
    proc = subprocess.run([sys.executable, '-c', script], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert proc.returncode == 0, proc.stderr[-2000:]
    line = [ln for ln in proc.stdout.splitlines() if ln.startswith('RESULT')][0]
    res = json.loads(line[len('RESULT '):])
    assert res == {'n': 10, 'types': ['InitialiseOptimiserTask', 'OptimiseTask'], 'ntasks': n_tasks}
    assert (out / 'profile' / 'x2.txt').read_text() == gold['files']['x2.txt']
    assert (out / 'profile' / 'x2_status.txt').read_text() == gold['files']['x2_status.txt']
    # and our own reader understands PROSPECT's config.pkl without PROSPECT being importable
    from prospect_public_b200.io import load_profile
    assert len(load_profile(str(out))['x2']) == 10


def test_parabola_vertex_bit_identical_to_the_reference():
    """400 random triples through the reference's own AnalyseProfileTask.get_bestfit_parabola (frozen as exact hex
    floats by tests/golden/make_golden_r2.py): x* and f* must agree in every bit, because both are printed at full
    precision into the profile files."""
    import json
    with open(os.path.join(GOLDEN, 'ref_parabola_cases.json')) as f:
        cases = json.load(f)
    assert len(cases) == 400
    for c in cases:
        x, fv = [float.fromhex(v) for v in c['x']], [float.fromhex(v) for v in c['f']]
        xb, fb = analysis.bestfit_parabola(x, fv)
        assert xb == float.fromhex(c['x_best']) and fb == float.fromhex(c['f_best'])


def test_shard_grid_covers_every_chain_once_even_with_more_ranks_than_repetitions():
    """P >= world: contiguous point blocks; P < world: repetitions split -- with fewer repetitions than ranks some
    ranks own nothing (global_optimisation with 3 repetitions on 8 GPUs), and that must still be a partition."""
    from prospect_public_b200.distributed import global_chain_index, shard_grid
    for n_points, reps, world in ((32, 256, 8), (5, 6, 3), (1, 3, 8), (3, 2, 8), (7, 1, 2)):
        seen = np.zeros(n_points * reps, dtype=int)
        sizes = []
        for rank in range(world):
            pt, rep = shard_grid(n_points, reps, rank, world)
            sizes.append(len(pt))
            if len(pt):
                assert np.all(np.diff(global_chain_index(pt, rep, n_points)) > 0)      # rep-major, ascending
                n_local = len(np.unique(pt))
                assert np.array_equal(np.searchsorted(np.unique(pt), pt), np.arange(len(pt)) % n_local)  # c % P_local
            seen[global_chain_index(pt, rep, n_points)] += 1
        assert np.all(seen == 1)
        assert (min(sizes) == 0) == (n_points < world and reps < world)


def test_gathered_rows_start_as_did_not_run():
    """Rows of the multi-GPU `gathered` tensor that were never shipped must decode as "chain did not run"
    (loglkl = inf, accepted = -1), like a single-GPU Records: a partial run (stop_after) decodes every row."""
    from prospect_public_b200.engine import Records
    from prospect_public_b200.optimiser import blank_rows
    rows, world, b = 4, 3, 5
    scalar_len = 4 * b + (b + 1) // 2
    g = blank_rows(rows, world, scalar_len, b, 'cpu').numpy()
    dec = Records.decode(g[:, 1], b, b, 32, False)
    assert np.all(np.isinf(dec['best_loglkl'])) and np.all(np.isinf(dec['last_loglkl']))
    assert np.all(dec['accepted'] == -1)


def test_gelman_rubin_follows_the_published_getdist_definition():
    """R-1 as getdist's MCSamples.getGelmanRubin defines it (what prospect/analysis.py:27-32 calls): "the maximum
    var(mean) / mean(var) of orthogonalized parameters, c.f. Brooks and Gelman 1997" -- per-chain weighted means and
    (biased) covariances, chains weighted by their total weight, the between-chain covariance of the means scaled by
    C / (C - 1), both matrices normalised by the diagonal of the mean covariance, largest eigenvalue of
    L^-1 B L^-T with L the Cholesky factor of the mean covariance.  getdist is not installed here and the reference
    ships no value for this statistic, so parity stays UNPINNED; this test pins the implementation to
    a hand-computed case and to an independent transcription of the definition above, for both code paths
    (per-chain moments and the all-reduced sufficient statistics)."""
    from prospect_public_b200.run import convergence_from_statistics, gelman_rubin
    # two 1-D chains {0, 2} and {4, 6}: means 1 and 5, variances 1, overall mean 3;
    # between = ((1-3)^2 + (5-3)^2) / 2 * 2/(2-1) = 8, within = 1  ->  R-1 = 8
    means = np.array([[1.0], [5.0]])
    covs = np.array([[[1.0]], [[1.0]]])
    assert abs(gelman_rubin(means, covs, np.array([2.0, 2.0])) - 8.0) < 1e-12
    rng = np.random.default_rng(3)
    n, chains = 3, []
    for c in range(5):
        rows = int(rng.integers(40, 90))
        a = rng.normal(size=(n, n))
        chains.append((rng.integers(1, 6, rows).astype(float),
                       rng.normal(size=(rows, n)) @ a + rng.normal(scale=0.3, size=n)))

    def independent(chains):                       # the definition, step by step
        norms = np.array([w.sum() for w, _ in chains])
        m_c = np.array([(w[:, None] * x).sum(0) / w.sum() for w, x in chains])
        cov_c = np.array([np.cov(x, rowvar=False, bias=True, aweights=w) for w, x in chains])
        mean = (norms[:, None] * m_c).sum(0) / norms.sum()
        meancov = (norms[:, None, None] * cov_c).sum(0) / norms.sum()
        dm = m_c - mean
        meanscov = (norms[:, None, None] * dm[:, :, None] * dm[:, None, :]).sum(0) / norms.sum()
        meanscov *= len(chains) / (len(chains) - 1)
        d = np.diag(1.0 / np.sqrt(np.diag(meancov)))
        meancov, meanscov = d @ meancov @ d, d @ meanscov @ d
        li = np.linalg.inv(np.linalg.cholesky(meancov))
        return float(np.linalg.eigvalsh(li @ meanscov @ li.T).max()), m_c, cov_c, norms

    expect, m_c, cov_c, norms = independent(chains)
    assert abs(gelman_rubin(m_c, cov_c, norms) - expect) < 1e-10 * expect
    # the sufficient statistics the GPU reduces and the ranks all-reduce: [sw, swx, swxx, sum_c swx_c swx_c^T / sw_c, C]
    sw = norms.sum()
    swx = sum((w[:, None] * x).sum(0) for w, x in chains)
    swxx = sum((w[:, None, None] * x[:, :, None] * x[:, None, :]).sum(0) for w, x in chains)
    syy = sum(np.outer((w[:, None] * x).sum(0), (w[:, None] * x).sum(0)) / w.sum() for w, x in chains)
    stats = np.concatenate([[sw], swx, swxx.reshape(-1), syy.reshape(-1), [len(chains)]])
    got, mean, cov = convergence_from_statistics(stats, n)
    assert abs(got - expect) < 1e-9 * expect


@pytest.mark.parametrize('jobtype,n_points,reps', [('profile', 7, 40), ('global_optimisation', 1, 90)])
def test_bulk_snapshot_writer_builds_the_same_object_graph(toy20, tmp_path, jobtype, n_points, reps):
    """compat.export_reference_snapshot_bulk writes state.pkl as bytes (one OptimiseTask laid out as a pickle template,
    tiled and filled with numpy); pickle.load must give exactly the object graph of the object-by-object exporter in
    its compact mode -- same keys in the same order, same values, same sharing inside a task."""
    import pickle
    from prospect_public_b200 import compat
    rng = np.random.default_rng(5)
    k_rows, n = 6, 19 if jobtype == 'profile' else 20
    b = n_points * reps
    acc = rng.integers(0, 100, (k_rows, b)).astype(np.int32)
    acc[4:, ::5] = -1                              # chains that stopped early
    acc[:, 3] = -1                                 # a chain that never ran
    names = [f'x{i}' for i in range(1, 21) if not (jobtype == 'profile' and i == 2)]
    rec = analysis.RunRecord(np.linspace(-1.0, 2.0, n_points) if jobtype == 'profile' else np.zeros(0), reps, names,
                             rng.uniform(0, 50, (k_rows, b)), acc, rng.normal(size=(k_rows, b, n)), rng.random(b) < 0.2,
                             rng.uniform(0, 5, n_points), rng.normal(size=(n_points, n)),
                             np.stack([np.eye(n) * (p + 1) for p in range(n_points)]),
                             rng.uniform(0, 1, (k_rows, b)), rng.uniform(0, 1, (k_rows, b)))
    kw = {'values': f'np.linspace(-1.0, 2.0, {n_points})'} if jobtype == 'profile' else {}
    cfg = Configuration(annealing_yaml(toy20, str(tmp_path / 'o'), jobtype=jobtype, write=False, repetitions=reps, **kw))
    dirs = [tmp_path / 'a', tmp_path / 'b']
    for d_ in dirs:
        d_.mkdir()
    n_a = compat.export_reference_snapshot(cfg, rec, str(dirs[0]), compact=True)
    n_b = compat.export_reference_snapshot_bulk(cfg, rec, str(dirs[1]))
    assert n_a == n_b

    def same(a, b_, path=''):
        assert type(a) is type(b_) or (isinstance(a, (float, np.floating)) and isinstance(b_, (float, np.floating))), path
        if isinstance(a, dict):
            assert list(a.keys()) == list(b_.keys()), path
            for key in a:
                if not (isinstance(key, str) and key.startswith('time_')):
                    same(a[key], b_[key], f'{path}[{key}]')
        elif isinstance(a, (list, tuple)):
            assert len(a) == len(b_), path
            for i, (x, y) in enumerate(zip(a, b_)):
                same(x, y, f'{path}[{i}]')
        elif isinstance(a, np.ndarray):
            assert np.array_equal(a, b_), path
        elif hasattr(a, '__dict__') and not isinstance(a, type):
            same(a.__dict__, b_.__dict__, f'{path}.{type(a).__name__}')
        else:
            assert a == b_, (path, a, b_)

    with compat._Registered():
        with open(dirs[0] / 'state.pkl', 'rb') as f:
            sa = pickle.load(f)
        with open(dirs[1] / 'state.pkl', 'rb') as f:
            sb = pickle.load(f)
        same(sa.__dict__, sb.__dict__)
        task = list(sb.done.values())[-1]
        assert task.optimise_settings is task.optimiser.settings
        assert task.optimiser.bestfit['position'] is task.optimise_settings['initial_position']


def _sync_worker(rank, world, port, inputs, base, q):
    """Two ranks build their own Configuration (no seed in the yaml, an output folder that already exists for rank 1
    only ... ) and call synchronise(); the all_gather_object of the sharded initialisation is exercised with the same
    payload shape it has in initialise_optimiser."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    y = annealing_yaml(inputs, os.path.join(base, f'private_{rank}'), seed=None)
    y['run']['numpy_random_seed'] = None
    cfg = Configuration(y)
    before = (cfg.io.dir, cfg.seed)
    cfg.synchronise()
    mine = np.array_split(np.arange(5), world)[rank]
    parts = [None] * world
    dist.all_gather_object(parts, ({int(p_): np.full(3, float(p_)) for p_ in mine}, {int(p_): np.eye(2) * p_ for p_ in mine},
                                   None))
    merged = {}
    for pos, cov, idx in parts:
        merged.update(pos)
    q.put((rank, before, (cfg.io.dir, cfg.seed, cfg.seed), sorted(merged)))
    dist.destroy_process_group()


def test_configuration_synchronise_world2_gloo(toy20, tmp_path):
    """Every rank of a torch.distributed job builds its own Configuration; without a numpy_random_seed each would draw
    its own Philox key, and a late rank could resolve a different output folder.  synchronise() makes rank 0's
    values the job's; the seed is drawn ONCE per Configuration (two reads give the same key)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_sync_worker, args=(rank, 2, port, toy20, str(tmp_path), q)) for rank in range(2)]
    for pr in procs:
        pr.start()
    outs = {r: rest for r, *rest in (q.get(timeout=120) for _ in range(2))}
    for pr in procs:
        pr.join(timeout=60)
    (before0, after0, pts0), (before1, after1, pts1) = outs[0], outs[1]
    assert before0[0] != before1[0] and before0[1] != before1[1]          # they really disagreed
    assert after0[:2] == before0 and after1[:2] == before0                # rank 0's folder and key everywhere
    assert after1[1] == after1[2]                                         # cached: the same key on every read
    assert pts0 == pts1 == [0, 1, 2, 3, 4]                                # every rank ends up with every value


def test_array_key_identity_for_frozen_arrays_content_otherwise():
    """kernels.array_key (the device-problem cache key of the proposal covariances): writeable arrays by content,
    read-only ones (no writeable base) by address -- views of the same frozen data agree, copies do not alias."""
    from prospect_public_b200.kernels import array_key
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 5, 5))
    b = a.copy()
    assert array_key(a) == array_key(b) and array_key(a)[0] == 'hash'       # same content, different objects
    b[1, 2, 3] += 1e-12
    assert array_key(a) != array_key(b)
    a.setflags(write=False)
    ka = array_key(a)
    assert ka[0] == 'frozen' and array_key(a) == ka and array_key(a[:]) == ka   # a fresh view of the same data
    assert array_key(a[1:]) != ka and array_key(a.copy())[0] == 'hash'
    c = rng.standard_normal((5, 5))
    view = c[None]
    view.setflags(write=False)                                                   # read-only view of a WRITEABLE base
    assert array_key(view)[0] == 'hash'
    c.setflags(write=False)
    assert array_key(c[None]) == array_key(c[None]) and array_key(c[None])[0] == 'frozen'
    assert array_key(None) is None


def test_continue_chain_folds_the_repeated_first_row():
    """mcmc.continue_chain: a segment sampled from a chain's last point continues that chain -- the repeated point's
    multiplicity goes on counting in the open last row (Chain.mults[-1] += 1, prospect/mcmc.py:178-180)."""
    from prospect_public_b200.mcmc import Chain, continue_chain
    prev = Chain([2, 3], [5.0, 4.0], {'x1': [0.1, 0.2], 'x2': [1.0, 1.0]})
    seg = Chain([4, 1, 2], [4.0, 3.5, 3.0], {'x1': [0.2, 0.3, 0.4], 'x2': [1.0, 1.0, 1.0]})
    out = continue_chain(prev, seg)
    assert out.mults == [2, 6, 1, 2] and out.loglkls == [5.0, 4.0, 3.5, 3.0]
    assert out.positions['x1'] == [0.1, 0.2, 0.3, 0.4] and out.positions['x2'] == [1.0] * 4
    assert prev.mults == [2, 3] and seg.mults == [4, 1, 2] and out.N == 4          # inputs untouched
    assert sum(out.mults) == sum(prev.mults) + sum(seg.mults) - 1                  # one start row is not a step
    lone = continue_chain(prev, Chain([7], [4.0], {'x1': [0.2], 'x2': [1.0]}))     # nothing accepted in the segment
    assert lone.mults == [2, 9] and lone.loglkls == [5.0, 4.0]
