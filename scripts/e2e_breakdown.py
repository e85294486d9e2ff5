"""Where does an end-to-end step (host arrays -> SimulatedAnnealing -> records on host) spend its time?"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from prospect_public_b200 import engine
from prospect_public_b200.config import Configuration
from prospect_public_b200.initialise import initialise_optimiser
from prospect_public_b200.kernels import initialise_kernel
from prospect_public_b200.optimiser import SimulatedAnnealing, get_initial_step_size_change, get_temperature_change
from prospect_public_b200.toy import annealing_yaml

jobtype, kw, n_points, reps = bench.workload_profile_kw('global30', 1)
tmp = tempfile.mkdtemp()
paths, kgold, rows = bench.toy_inputs(tmp)
cfg = Configuration(annealing_yaml(paths, os.path.join(tmp, 'out'), jobtype=jobtype, write=False, **kw))
kernel = initialise_kernel(cfg.kernel, None, 0, 'cuda:0')
starts, values = initialise_optimiser(cfg, kernel)
def settings(seed):
    return {'current_best_loglkl': starts.initial_loglkl, 'initial_position': starts.positions, 'covmat': starts.covmats,
            'temperature': 0.1, 'temperature_change': get_temperature_change(cfg), 'step_size': 0.1,
            'step_size_change': get_initial_step_size_change(cfg), 'iteration_number': 0, 'repetitions': reps,
            'fixed_param_val': None, 'max_rows': 20, 'seed': seed}
def sync(): torch.cuda.synchronize()
acc = {}
def tick(name, t0):
    sync(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
for s in range(12):
    if s == 2: acc.clear()
    t0 = time.perf_counter(); sa = SimulatedAnnealing(cfg, kernel, settings(s)); tick('construct', t0)
    t0 = time.perf_counter(); sa.run_schedule(20); tick('schedule', t0)
    t0 = time.perf_counter(); red = sa.reduce_over_iterations(); tick('reduce', t0)
    t0 = time.perf_counter(); bf = sa.set_bestfit(19); out = {k: v.cpu() for k, v in red.items()}; tick('d2h', t0)
print({k: round(1e3 * v / 10, 3) for k, v in acc.items()}, 'ms per step')
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for s in range(20):
    sa = SimulatedAnnealing(cfg, kernel, settings(s)); sync()
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(45)

# the same loop without intermediate syncs (what bench.py's e2e arm does)
sync(); t0 = time.perf_counter()
for s in range(10):
    sa = SimulatedAnnealing(cfg, kernel, settings(100 + s)); sa.run_schedule(20)
    red = sa.reduce_over_iterations(); bf = sa.set_bestfit(19); out = {k: v.cpu() for k, v in red.items()}
sync(); print('no-sync loop ms per step', round(1e3 * (time.perf_counter() - t0) / 10, 3))
import torch
for trial in range(3):
    t0 = time.perf_counter(); h = torch.empty(1 << 20, dtype=torch.uint8, pin_memory=True); t1 = time.perf_counter()
    print('pinned 1MB alloc ms', round(1e3 * (t1 - t0), 3)); del h
