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
tmp = tempfile.mkdtemp(); paths, kgold, rows = bench.toy_inputs(tmp)
cfg = Configuration(annealing_yaml(paths, os.path.join(tmp, 'out'), jobtype=jobtype, write=False, **kw))
kernel = initialise_kernel(cfg.kernel, None, 0, 'cuda:0'); starts, values = initialise_optimiser(cfg, kernel)
st = {'current_best_loglkl': starts.initial_loglkl, 'initial_position': starts.positions, 'covmat': starts.covmats,
      'temperature': 0.1, 'temperature_change': get_temperature_change(cfg), 'step_size': 0.1,
      'step_size_change': get_initial_step_size_change(cfg), 'iteration_number': 0, 'repetitions': reps,
      'fixed_param_val': None, 'max_rows': 20, 'seed': 1}
sa = SimulatedAnnealing(cfg, kernel, st)
zb = engine.variates_buffer(sa.problem, sa.n_chains, 250)
print('zbuf MB', zb.numel() / 1e6)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
print('variates_fill ms', timeit(lambda: engine.variates_fill(sa.problem, sa.batch, 0, 250, 1, zb)))
for lanes in (32, 16, 8):
    def one():
        sa.batch.iteration.zero_(); sa.batch.temperature.fill_(0.1); sa.batch.step_size.fill_(0.1)
        engine.anneal_run_streamed(sa.problem, sa.batch, sa.schedule, sa.records, 0, zb, lanes=lanes)
    def fused():
        sa.batch.iteration.zero_(); sa.batch.temperature.fill_(0.1); sa.batch.step_size.fill_(0.1); sa.batch.steps_done.zero_()
        engine.anneal_run(sa.problem, sa.batch, sa.schedule, sa.records, 1, 1, lanes=lanes)
    print('lanes', lanes, 'streamed 1 iteration ms', round(timeit(one), 4), 'fused 1 iteration ms', round(timeit(fused), 4))
