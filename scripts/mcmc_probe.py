"""K5 timing: 65 536 MH chains on the 30-D toy, 2000 steps per round (kernel time by CUDA events)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from oracle import closed_form
from prospect_public_b200 import engine
from prospect_public_b200.config import Configuration
from prospect_public_b200.kernels import initialise_kernel
from prospect_public_b200.mcmc import initialise_mcmc
from prospect_public_b200.toy import mcmc_yaml
tmp = tempfile.mkdtemp()
paths, kgold, rows = bench.toy_inputs(tmp)
cov = closed_form.stationary_covariance(kgold['covmat'])
for record, thin, step in ((None, 1, 2.4 / np.sqrt(30)), (None, 1, 0.2), (None, 1, 0.1), (None, 1, 0.05), ('thinned', 2000, 2.4 / np.sqrt(30)),
                           ('thinned', 100, 2.4 / np.sqrt(30)), ('accepted', 1, 2.4 / np.sqrt(30))):
    n_chains = 65536 if record != 'accepted' else 4096
    cfg = Configuration(mcmc_yaml(paths, os.path.join(tmp, 'o'), n_chains=n_chains, steps=2000, write=False,
                                  start_from_covmat=cov.tolist(), step_size=float(step)))
    kernel = initialise_kernel(cfg.kernel, None, 0)
    s = initialise_mcmc(cfg.mcmc, kernel, seed=1, record=record, thin=thin)
    s.run_steps(2000); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s.run_steps(2000); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    line = f'record={record} thin={thin} step={step:.3f} AR={s.accepted.sum() / (n_chains * 4000):.2f} chains={n_chains}: {ms:.2f} ms per 2000-step round, {n_chains * 2000 / ms / 1e-3:.3e} chain-steps/s'
    if record is not None:
        a.record(); st = engine.chain_statistics(s._samples, s.problem.n); b.record(); torch.cuda.synchronize()
        line += f'; chain_statistics {a.elapsed_time(b):.2f} ms ({int(s._samples.count.sum())} rows)'
    print(line, flush=True)
