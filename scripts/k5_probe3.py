"""K5 with thinned recording (every 100th step): the kernel instance ncu looks at for the recording overhead."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from prospect_public_b200.config import Configuration
from prospect_public_b200.kernels import initialise_kernel
from prospect_public_b200.mcmc import initialise_mcmc
from prospect_public_b200.toy import mcmc_yaml
tmp = tempfile.mkdtemp()
paths, kgold, rows = bench.toy_inputs(tmp)
prec = np.linalg.inv(kgold['covmat']); cov = np.linalg.inv(0.5 * (prec + prec.T))
for thin in (100, 1000):
    y = mcmc_yaml(paths, os.path.join(tmp, 'o'), n_chains=65536, steps=2000, write=False, start_from_covmat=cov.tolist(), step_size=float(2.4 / np.sqrt(30)))
    cfg = Configuration(y)
    kernel = initialise_kernel(cfg.kernel, None, 0)
    s = initialise_mcmc(cfg.mcmc, kernel, seed=1, record='thinned', thin=thin)
    s.run_steps(2000); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s.run_steps(2000); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f'thin={thin} AR={s.accepted.sum() / (65536 * 4000):.2f}: {ms:.2f} ms, {65536 * 2000 / ms / 1e-3:.3e} chain-steps/s', flush=True)
