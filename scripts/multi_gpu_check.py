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
dist.destroy_process_group()
