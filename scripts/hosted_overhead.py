"""Per-step overhead of the hosted path (propose launch + D2H + host batch + H2D + accept launch) with a likelihood that
costs (almost) nothing: what the batched machinery adds on top of the host code's own evaluation time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np
import torch
from prospect_public_b200.hosted import HostKernelAdapter
from prospect_public_b200.mcmc import initialise_mcmc
from tests.helpers import ToyHostKernel

for d, b in ((30, 64), (30, 4096), (30, 65536), (255, 1024)):
    host = ToyHostKernel(d, ignore_prior=True)
    prec, mu = host.prec, host.mu

    def batch(x, fixed, prec=prec, mu=mu):
        r = x - mu
        return 0.5 * np.einsum('ij,ij->i', r @ prec, r)

    kernel = HostKernelAdapter(host, ignore_prior=True, batch_loglkl=batch)
    cfg = SimpleNamespace(algorithm='MetropolisHastings', N_chains=b, temperature=1.0, step_size=2.4 / np.sqrt(d),
                          start_from_covmat=np.linalg.inv(prec), start_from_position=None, steps_per_iteration=50)
    mh = initialise_mcmc(cfg, kernel, seed=1)
    mh.run_steps(5)
    torch.cuda.synchronize()
    t_like = 0.0
    t0 = time.perf_counter()
    steps = 40
    mh.run_steps(steps)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    x = np.random.default_rng(0).standard_normal((b, d))
    t1 = time.perf_counter()
    for _ in range(10):
        batch(x, None)
    t_like = (time.perf_counter() - t1) / 10
    print(f'd {d} chains {b}: {dt * 1e6:.0f} us per MH step of the batch, of which the host likelihood batch '
          f'{t_like * 1e6:.0f} us -> machinery {max(dt - t_like, 0) * 1e6:.0f} us per step '
          f'({max(dt - t_like, 0) / b * 1e9:.1f} ns per chain-step); acceptance {mh.accepted.mean() / mh.steps:.3f}', flush=True)
