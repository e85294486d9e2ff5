"""Where does time-to-profile go?  python scripts/ttp_profile.py   or   torchrun --nproc-per-node N scripts/ttp_profile.py"""
import os, sys, tempfile, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from prospect_public_b200.run import run
from prospect_public_b200.toy import annealing_yaml
world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
if world > 1:
    import torch, torch.distributed as dist
    from prospect_public_b200.distributed import init_from_env
    init_from_env('nccl')
tmp = os.path.join(tempfile.gettempdir(), 'pb200_ttp')
if rank == 0:
    bench.toy_inputs(tmp)
if world > 1:
    dist.barrier()
paths = {'kernel_pkl': os.path.join(tmp, 'random_gauss.pkl'), 'kernel_yaml': os.path.join(tmp, 'kernel_load_random_gauss.yaml'),
         'mcmc_root': os.path.join(tmp, 'mcmc', 'test')}
pkw = dict(bench.workload_profile_kw('profile30', world)[1])
run(annealing_yaml(paths, os.path.join(tmp, 'warm'), jobtype='profile', write=True, **pkw))
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
out = run(annealing_yaml(paths, os.path.join(tmp, 'ttp'), jobtype='profile', write=True, **pkw))
pr.disable()
if rank == 0:
    print('world', world, 'wall', round(time.perf_counter() - t0, 3), 'time_to_profile_s', round(out.get('time_to_profile_s', -1), 3),
          'loop_s', round(out['loop_seconds'], 4))
    pstats.Stats(pr).sort_stats('cumulative').print_stats(32)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
