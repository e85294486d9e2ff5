"""Host-side phases of the double-buffered end-to-end loop (bench.py e2e_rate), per job, without extra syncs."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench

class A: pass
args = A(); args.reps_per_gpu = 4096; args.jobs_per_step = 2; NJ = int(os.environ.get("NJ", "2"))
h = bench.Harness(args, 'global30', 1, 0, torch.device("cuda:0"), NJ)
acc = {}
def lap(name, t0):
    t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + t1 - t0; return t1
pending = None
N = 40
for s in range(N + 4):
    if s == 4:
        pending.get(); pending = None; torch.cuda.synchronize(); acc.clear(); t_start = time.perf_counter()
    t = time.perf_counter()
    sa = h.new_job(100 + s); t = lap('new_job', t)
    sa.run_schedule(h.n_iter); t = lap('run_schedule', t)
    red = sa.reduce_over_iterations(); t = lap('reduce', t)
    nxt = sa.results_async(h.n_iter - 1, red); t = lap('results_async', t)
    if pending is not None:
        pending.get(); t = lap('get', t)
    pending = nxt
pending.get(); torch.cuda.synchronize()
total = time.perf_counter() - t_start
print({k: round(1e3 * v / N, 3) for k, v in acc.items()}, 'ms per job; total', round(1e3 * total / N, 3), 'ms per job')
# GPU-side floor: the same launches with no read-back at all
torch.cuda.synchronize(); t0 = time.perf_counter()
keep = []
for s in range(N):
    sa = h.new_job(500 + s); sa.run_schedule(h.n_iter); red = sa.reduce_over_iterations(); keep = [sa, red]
torch.cuda.synchronize(); print('no read-back:', round(1e3 * (time.perf_counter() - t0) / N, 3), 'ms per job')
