"""K4 scaling probe: the 256-D profile at P points x R repetitions, per-phase device times of the wide path
(PB200_WIDE_TIMING=1 prints them) and the total per 20-iteration call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('PB200_WIDE_TIMING', '1')
import numpy as np, torch
from prospect_public_b200 import capi, engine
from prospect_public_b200.problem import DeviceProblem, build_problem
from prospect_public_b200.toy import conditional_covariance, spd_target

d, n_iter, S = 256, 20, 250
mu, cov = spd_target(d, seed=4)
sig = float(np.sqrt(cov[1, 1]))
prop = conditional_covariance(cov)
shapes = [tuple(map(int, a.split('x'))) for a in sys.argv[1:]] or [(8, 16), (8, 64), (8, 128), (8, 256), (8, 512)]
for P, R in shapes:
    vals = mu[1] + sig * np.linspace(-2.0, 2.0, P)
    hp = build_problem(mu, cov, [-2.5] * d, [2.5] * d, 1, vals, np.stack([prop] * P))
    prob = DeviceProblem(hp, 'cuda:0')
    b = P * R
    point = np.tile(np.arange(P), R).astype(np.int32)
    x0 = np.tile(np.delete(mu, 1) + 0.01, (b, 1))
    tc = (0.001 / 0.1) ** (1 / 20)
    sched = engine.make_schedule(S, tc, capi.STEP_ADAPTIVE_FIXED, 1.0, (0.19, 0.21), 0.3)
    for rep in range(3):
        chains = engine.ChainBatch(prob, x0, point, temperature=0.1, step_size=0.15)
        recs = engine.Records(prob, b, n_iter)
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        engine.anneal_run(prob, chains, sched, recs, n_iter, 7 + rep)
        e.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(e)
    print(f'P={P} R={R} chains={b}: {ms:.3f} ms per call, {b * n_iter * S / ms / 1e-3:.3e} chain-steps/s, '
          f'mean AR {float((recs.accepted.float().mean() + 1) / (S + 1)):.3f}', flush=True)
