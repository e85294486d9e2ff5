"""Where does the fused exchange spend its time?  torchrun --nproc-per-node N scripts/exchange_probe.py"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from prospect_public_b200 import engine, optimiser
from prospect_public_b200.config import Configuration
from prospect_public_b200.distributed import init_from_env, dist_info
from prospect_public_b200.initialise import initialise_optimiser
from prospect_public_b200.kernels import initialise_kernel
from prospect_public_b200.optimiser import SimulatedAnnealing, get_initial_step_size_change, get_temperature_change
from prospect_public_b200.toy import annealing_yaml
rank, world, _ = init_from_env('nccl')
jobtype, kw, n_points, reps = bench.workload_profile_kw('global30', world)
tmp = tempfile.mkdtemp(prefix=f'probe{rank}_')
paths, kgold, rows = bench.toy_inputs(tmp)
cfg = Configuration(annealing_yaml(paths, os.path.join(tmp, 'out'), jobtype=jobtype, write=False, **kw))
dev = f'cuda:{torch.cuda.current_device()}'
kernel = initialise_kernel(cfg.kernel, None, 0, dev)
starts, values = initialise_optimiser(cfg, kernel)
def settings(seed):
    return {'current_best_loglkl': starts.initial_loglkl, 'initial_position': starts.positions, 'covmat': starts.covmats,
            'temperature': 0.1, 'temperature_change': get_temperature_change(cfg), 'step_size': 0.1,
            'step_size_change': get_initial_step_size_change(cfg), 'iteration_number': 0, 'repetitions': reps,
            'fixed_param_val': None, 'max_rows': 20, 'seed': seed}
def timed(fn, n=10):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i in range(n + 3):
        sa = SimulatedAnnealing(cfg, kernel, settings(i)); torch.cuda.synchronize(); dist.barrier()
        if i >= 3: ev[i - 3][0].record()
        fn(sa)
        if i >= 3: ev[i - 3][1].record()
    torch.cuda.synchronize()
    return round(float(np.mean([a.elapsed_time(b) for a, b in ev])), 4)
res = {}
res['schedule fused'] = timed(lambda sa: sa.run_schedule(20))
def nccl(sa):
    sa.fused_exchange = False; sa.run_schedule(20)
res['schedule nccl'] = timed(nccl)
res['no exchange (gather=False)'] = timed(lambda sa: sa.run_schedule(20, gather=False))
if rank == 0:
    print('world', world, res, flush=True)
dist.barrier(); dist.destroy_process_group()
