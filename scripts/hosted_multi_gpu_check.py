"""torchrun check of the hosted path (host-evaluated likelihood, prospect_public_b200/hosted.py) sharded over several GPUs:
the job's gathered records and its profile file must be identical, bit for bit, to the same job run by one rank alone.
The host likelihood is the UNMODIFIED reference AnalyticalKernel (oracle/_ref).
usage: python -m torch.distributed.run --nproc-per-node N scripts/hosted_multi_gpu_check.py"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ref_harness as rh
from prospect_public_b200.distributed import dist_info, init_from_env
from prospect_public_b200.hosted import HostedSimulatedAnnealing, HostKernelAdapter
from prospect_public_b200.run import run
from prospect_public_b200.toy import annealing_yaml, write_toy_inputs

rh.activate()
from prospect.kernels.analytical_kernel import AnalyticalKernel  # noqa: E402

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
k = np.load(os.path.join(root, 'tests/golden/toy20_kernel.npz')); m = np.load(os.path.join(root, 'tests/golden/toy20_mcmc.npz'))
base = os.path.join(tempfile.gettempdir(), 'pb200_hosted_mgpu_check')
if int(os.environ.get('RANK', '0')) == 0:
    write_toy_inputs(base, k['means'], k['covmat'], [m[f'chain_{i}'] for i in range(4)])
import torch.distributed as dist  # noqa: E402
init_from_env('nccl'); dist.barrier()
paths = {'kernel_pkl': os.path.join(base, 'random_gauss.pkl'), 'kernel_yaml': os.path.join(base, 'kernel_load_random_gauss.yaml'),
         'mcmc_root': os.path.join(base, 'mcmc', 'test')}


def factory(config_kernel, output_folder, rank, device):
    with rh.quiet():
        return HostKernelAdapter(AnalyticalKernel(config_kernel, rank, output_folder=output_folder), device=device)


from prospect_public_b200 import distributed as dist_mod, optimiser as opt_mod, run as run_mod  # noqa: E402
for label, kw in (('5 values x 3 reps (points sharded)', dict(values=[-0.6, -0.1, 0.4, 0.9, 1.4], repetitions=3)),
                  ('1 value x 5 reps (repetitions sharded)', dict(values=[0.5], repetitions=5))):
    kw.update(max_iterations=4, steps_per_iteration=60)
    multi = run(annealing_yaml(paths, os.path.join(base, 'out_multi'), **kw), kernel=factory)
    assert isinstance(multi['optimiser'], HostedSimulatedAnnealing)
    dist.barrier()
    if dist_info()[0] == 0:
        real, real_ws = dist_mod.dist_info, os.environ['WORLD_SIZE']
        opt_mod.dist_info = run_mod.dist_info = dist_mod.dist_info = lambda: (0, 1, False)
        os.environ['WORLD_SIZE'] = '1'
        try:
            alone = run(annealing_yaml(paths, os.path.join(base, 'out_alone'), **kw), kernel=factory)
        finally:
            opt_mod.dist_info = run_mod.dist_info = dist_mod.dist_info = real
            os.environ['WORLD_SIZE'] = real_ws
        same = all(np.array_equal(getattr(multi['record'], key), getattr(alone['record'], key))
                   for key in ('best_loglkl', 'accepted', 'best_x', 'step_size', 'temperature'))
        a = open(os.path.join(base, 'out_multi', 'profile', 'x2.txt'), 'rb').read()
        b = open(os.path.join(base, 'out_alone', 'profile', 'x2.txt'), 'rb').read()
        print(f'hosted N-invariance ({label}): world {dist_info()[1]} records identical to the single-rank run {same}, '
              f'profile/x2.txt identical {a == b}, chain_steps {multi["chain_steps"]}', flush=True)
        assert same and a == b and multi['chain_steps'] == alone['chain_steps']
    dist.barrier()
if dist_info()[0] == 0:
    print('hosted multi-GPU check OK', flush=True)
dist.destroy_process_group()
