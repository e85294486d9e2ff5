"""torchrun check of the sharded yaml -> run() path: the profile must agree with the analytic one and rank 0
must write the PROSPECT output tree.  usage: python -m torch.distributed.run --nproc-per-node N scripts/multi_gpu_check.py"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import closed_form
from prospect_public_b200.run import run
from prospect_public_b200.toy import annealing_yaml, write_toy_inputs
from prospect_public_b200.distributed import dist_info

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
k = np.load(os.path.join(root, 'tests/golden/toy20_kernel.npz')); m = np.load(os.path.join(root, 'tests/golden/toy20_mcmc.npz'))
base = os.path.join(tempfile.gettempdir(), 'pb200_mgpu_check')
rank = int(os.environ.get('RANK', '0'))
if rank == 0:
    paths = write_toy_inputs(base, k['means'], k['covmat'], [m[f'chain_{i}'] for i in range(4)])
import torch, torch.distributed as dist
from prospect_public_b200.distributed import init_from_env
init_from_env('nccl'); dist.barrier()
paths = {'kernel_pkl': os.path.join(base, 'random_gauss.pkl'), 'kernel_yaml': os.path.join(base, 'kernel_load_random_gauss.yaml'),
         'mcmc_root': os.path.join(base, 'mcmc', 'test')}
for label, kw in (('points sharded (10 values x 6 reps)', dict(repetitions=6)),
                  ('reps sharded (1 value x 64 reps)', dict(repetitions=64, values=[0.5]))):
    res = run(annealing_yaml(paths, os.path.join(base, 'out'), temperature_range=[0.1, 1.0e-6], max_iterations=30, **kw))
    if dist_info()[0] == 0:
        vals = res['values']
        cf = closed_form.profile_curve(k['means'], k['covmat'], 1, vals)
        err = np.max(2 * np.abs(res['profile']['loglkls'] - cf))
        ran = (res['record'].accepted >= 0).all()
        files = os.path.isfile(os.path.join(base, 'out', 'profile', 'x2.txt'))
        print(f'{label}: world {dist_info()[1]} max |dchi2| vs analytic {err:.2e} all chains recorded {ran} files {files} '
              f'chain_steps {res["chain_steps"]}')
        assert err <= 1e-3 and ran and files
    dist.barrier()
# N-invariance, bit for bit: the sharded job's gathered records == the same job run by rank 0 alone as a single-rank job
# (Philox streams are keyed by the global chain id); incl. fewer repetitions than ranks (ranks without chains)
from prospect_public_b200 import optimiser as opt_mod
from prospect_public_b200 import run as run_mod
for label, kw in (('10 values x 6 reps', dict(repetitions=6, max_iterations=5)),
                  ('global optimisation, 3 repetitions (fewer chains than ranks at N = 8)',
                   dict(repetitions=3, max_iterations=5, jobtype='global_optimisation'))):
    jt = kw.pop('jobtype', 'profile')
    multi = run(annealing_yaml(paths, os.path.join(base, 'out_inv'), jobtype=jt, **kw))
    dist.barrier()
    if dist_info()[0] == 0:
        from prospect_public_b200 import distributed as dist_mod
        real, real_ws = dist_mod.dist_info, os.environ['WORLD_SIZE']
        opt_mod.dist_info = run_mod.dist_info = dist_mod.dist_info = lambda: (0, 1, False)
        os.environ['WORLD_SIZE'] = '1'
        try:
            alone = run(annealing_yaml(paths, os.path.join(base, 'out_alone'), jobtype=jt, **kw))
        finally:
            opt_mod.dist_info = run_mod.dist_info = dist_mod.dist_info = real
            os.environ['WORLD_SIZE'] = real_ws
        same = all(np.array_equal(getattr(multi['record'], key), getattr(alone['record'], key))
                   for key in ('best_loglkl', 'accepted', 'best_x', 'step_size', 'temperature'))
        sub = 'profile/x2.txt' if jt == 'profile' else 'global_optimisation/results.txt'
        a = open(os.path.join(base, 'out_inv', sub), 'rb').read()
        b = open(os.path.join(base, 'out_alone', sub), 'rb').read()
        print(f'N-invariance ({label}): world {dist_info()[1]} records identical to the single-rank run {same}, {sub} '
              f'identical {a == b}')
        assert same and a == b
    dist.barrier()
# the outputs of a run stopped early (rows that were never exchanged must read "did not run")
part = run(annealing_yaml(paths, os.path.join(base, 'out_stop'), repetitions=6, max_iterations=8), stop_after=3)
if dist_info()[0] == 0:
    acc = part['record'].accepted
    ok = bool((acc[:3] >= 0).all() and (acc[3:] == -1).all() and np.isinf(part['record'].best_loglkl[3:]).all())
    print(f'stop_after: world {dist_info()[1]} rows 0-2 ran, rows 3-7 read "did not run": {ok}; chain_steps {part["chain_steps"]}')
    assert ok and part['chain_steps'] == 60 * 3 * 250
dist.barrier()
# interrupted + resumed run (per-chain state gathered over the ranks) == uninterrupted run
kw = dict(repetitions=6, max_iterations=8)
full = run(annealing_yaml(paths, os.path.join(base, 'out_full'), **kw))
run(annealing_yaml(paths, os.path.join(base, 'out_part'), **kw), stop_after=3)
dist.barrier()
res = run(os.path.join(base, 'out_part'))
if dist_info()[0] == 0:
    same = all(np.array_equal(getattr(res['record'], key), getattr(full['record'], key))
               for key in ('best_loglkl', 'accepted', 'best_x', 'step_size'))
    a = open(os.path.join(base, 'out_part', 'profile', 'x2.txt'), 'rb').read()
    b = open(os.path.join(base, 'out_full', 'profile', 'x2.txt'), 'rb').read()
    print(f'resume: world {dist_info()[1]} records identical {same}, profile/x2.txt identical {a == b}')
    assert same and a == b
dist.barrier()
# jobtype mcmc sharded over the ranks: chain files and statistics must not depend on the number of GPUs
from prospect_public_b200.config import Configuration
from prospect_public_b200.kernels import initialise_kernel
from prospect_public_b200.mcmc import initialise_mcmc
from prospect_public_b200.toy import mcmc_yaml
from prospect_public_b200 import engine
from prospect_public_b200.run import convergence_from_statistics
mc_dir = os.path.join(base, 'out_mcmc')
y = mcmc_yaml(paths, mc_dir, n_chains=19, steps=600, rm1=1e9)
res = run(y)
dist.barrier()
if dist_info()[0] == 0:
    cfg = Configuration(mcmc_yaml(paths, os.path.join(base, 'unused'), n_chains=19, steps=600, rm1=1e9, write=False))
    kern = initialise_kernel(cfg.kernel, None, 0, 'cuda:0')
    single = initialise_mcmc(cfg.mcmc, kern, seed=cfg.seed, n_chains=19, chain_ids=np.arange(19, dtype=np.uint32))
    single.run_steps(600)
    for i, ch in enumerate(single.chains):
        disk = np.loadtxt(os.path.join(mc_dir, 'mcmc', f'test_{i}.txt'))
        assert np.array_equal(disk, ch.data), i
    rm1, mean, cov = convergence_from_statistics(engine.chain_statistics(single._samples, single.problem.n).cpu().numpy(),
                                                 single.problem.n)
    assert abs(rm1 - res['Rm1']) <= 1e-9 * abs(rm1) and np.allclose(cov, res['covmat'], rtol=1e-9, atol=1e-15)
    print(f'mcmc sharded: world {dist_info()[1]} 19 chain files identical to a single-GPU run, R-1 {res["Rm1"]:.6f} '
          f'(single {rm1:.6f}), chain range of rank 0 {res["chain_range"]}')
dist.barrier()
# jobtype mcmc interrupted after one round and resumed from its chain files == the uninterrupted three rounds
full_dir, part_dir = os.path.join(base, 'mc_full'), os.path.join(base, 'mc_part')
full = run(mcmc_yaml(paths, full_dir, n_chains=11, steps=500, rm1=1e-12), max_rounds=3)
run(mcmc_yaml(paths, part_dir, n_chains=11, steps=500, rm1=1e-12), stop_after=1)
dist.barrier()
cont = run(part_dir, max_rounds=3)
dist.barrier()
if dist_info()[0] == 0:
    same = all(open(os.path.join(full_dir, 'mcmc', f'test_{i}.txt'), 'rb').read() ==
               open(os.path.join(part_dir, 'mcmc', f'test_{i}.txt'), 'rb').read() for i in range(11))
    print(f'mcmc resume: world {dist_info()[1]} chain files identical {same}, R-1 {cont["Rm1"]:.6f} vs {full["Rm1"]:.6f}')
    assert same and abs(cont['Rm1'] - full['Rm1']) <= 1e-12 * full['Rm1'] and cont['rounds'] == 3
dist.barrier()
dist.destroy_process_group()
