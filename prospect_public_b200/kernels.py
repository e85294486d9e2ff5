"""CUDA-backed likelihood kernel with PROSPECT's kernel plugin interface.

`CudaGaussianKernel` is the drop-in for `AnalyticalKernel(BaseKernel)`
(prospect/kernels/base_kernel.py:9-138, prospect/kernels/analytical_kernel.py:9-127): same constructor,
same `param` dict, `loglkl(position)` (which inserts the fixed parameters into `position` and returns
-log L), `logprior`, `logpost`, `set_fixed_parameters`, `varying_param_names`, default / user covmat and
initial position, and the same kernel-settings yaml (`function: gaussian | random_gaussian`, `load`,
`dimension`, `std_scale`, `off_diag_factor`) including the `<out>/analytical/random_gauss.pkl` persistence.
Every likelihood value is computed on the GPU (pb200_loglkl_batch); `loglkl_batch` is the batched entry
the samplers and the analysis use.  There is no CPU evaluation path.
"""
import os
import pickle

import numpy as np
import yaml

from . import capi
from .problem import DeviceProblem, build_problem


def initialise_kernel(config_kernel, output_folder, task_id, device='cuda:0'):
    """Factory with the reference's signature (prospect/kernels/initialisation.py:1-13)."""
    if config_kernel.type in ('analytical', 'cuda_gaussian'):
        return CudaGaussianKernel(config_kernel, task_id, output_folder=output_folder, device=device)
    if config_kernel.type in ('montepython', 'cobaya'):
        raise ValueError(f"Kernel type '{config_kernel.type}' needs an external likelihood code that is evaluated on "
                         'the host: build the reference kernel object yourself, wrap it in '
                         'prospect_public_b200.hosted.HostKernelAdapter and pass it as run(yaml, kernel=...).')
    raise ValueError('You have specified a non-existing kernel type.')


def load_kernel_settings(param_file, output_folder=None):
    """Kernel-settings yaml -> dict(means {name: mu}, covmat [d, d], dimension, ranges {name: [lo, hi]}).
    `random_gaussian` follows analytical_kernel.py:18-48: reuse <out>/analytical/random_gauss.pkl if it is
    there, else `load` a pickle or draw the recipe from numpy's global legacy RNG, then persist it.  The write
    goes through a temporary file + rename so concurrent ranks never read a half-written pickle."""
    with open(os.path.join(os.getcwd(), param_file), 'r') as f:
        cfg = yaml.full_load(f)
    if cfg['function'] == 'gaussian':
        names = list(cfg['param_dict'].keys())
        return {'means': {n: float(cfg['means'][n]) for n in names}, 'covmat': np.array(cfg['covmat'], dtype=float),
                'dimension': int(cfg['dimension']), 'ranges': {n: list(cfg['param_dict'][n]) for n in names}}
    if cfg['function'] != 'random_gaussian':
        raise ValueError('No analytical kernel with the desired function type.')
    settings_file = None if output_folder is None else os.path.join(output_folder, 'analytical', 'random_gauss.pkl')
    if settings_file is not None and os.path.isfile(settings_file):
        with open(settings_file, 'rb') as f:
            kd = pickle.load(f)
    else:
        if 'load' in cfg:
            with open(cfg['load'], 'rb') as f:
                kd = pickle.load(f)
        else:
            d = int(cfg['dimension'])
            means = {f'x{i}': np.random.rand() for i in range(1, d + 1)}
            diag = np.diag(np.random.rand(d))
            offdiag = np.random.rand(d, d)
            kd = {'means': means, 'covmat': cfg['std_scale'] * (diag + offdiag * cfg['off_diag_factor']),
                  'dimension': d}
        if settings_file is not None and os.path.isdir(os.path.dirname(settings_file)):
            tmp = f'{settings_file}.{os.getpid()}.tmp'
            with open(tmp, 'wb') as f:
                pickle.dump(kd, f)
            os.replace(tmp, settings_file)
    d = int(kd['dimension'])
    names = [f'x{i}' for i in range(1, d + 1)]
    return {'means': {n: float(kd['means'][n]) for n in names}, 'covmat': np.array(kd['covmat'], dtype=float),
            'dimension': d, 'ranges': {n: [-2.5, 2.5] for n in names}}


def array_key(a):
    """Cache key of an array's CONTENT.  A read-only array (no writeable base anywhere under it: nothing can change
    what it holds) is keyed by its address -- no pass over the data; the proposal covariances initialise_optimiser returns are
    made read-only for exactly this (hashing the 4 MB of a 256-D job's covariances was 1.5 ms of every optimiser
    construction).  Anything else is keyed by a hash of its bytes."""
    if a is None:
        return None
    node, frozen = a, True
    while isinstance(node, np.ndarray):
        if node.flags.writeable:
            frozen = False
            break
        node = node.base
    if frozen and node is None:
        # (the memory address, not id(): a fresh view of the same frozen data is the same content; the cache entry holds
        # the array, so the address cannot be handed to another array while the entry lives)
        return ('frozen', a.__array_interface__['data'][0], a.shape, a.strides)
    return ('hash', hash(a.tobytes()), a.shape)


class CudaGaussianKernel:
    """Gaussian likelihood evaluated by libpb200 on a B200.

    position = {param_name: [value]} as everywhere in PROSPECT; `loglkl` returns -log L."""

    def __init__(self, config_kernel, task_id, output_folder=None, device='cuda:0'):
        self.id = task_id
        self.device = device
        self.config_kernel = config_kernel
        self.output_folder = output_folder
        self.config = load_kernel_settings(config_kernel.param, output_folder)
        self.dimension = self.config['dimension']
        self.names = list(self.config['means'].keys())
        self.ignore_prior = bool(getattr(config_kernel, 'ignore_prior', False))
        self.param = {'varying': {n: {'range': list(self.config['ranges'][n])} for n in self.names},
                      'fixed': {}, 'derived': {}}
        self._problem = None          # device image for the current fixed-parameter set (built lazily)
        self._analytic = {}

    # --- parameter bookkeeping (base_kernel.py:46-55) ---------------------------------------------
    def set_fixed_parameters(self, fixed_param_dict):
        for name, value in fixed_param_dict.items():
            if name in self.param['fixed']:
                self.param['fixed'][name]['fixed_value'] = value
            else:
                self.param['fixed'][name] = self.param['varying'].pop(name)
                self.param['fixed'][name]['fixed_value'] = value
        if len(self.param['fixed']) > 1:
            raise ValueError('The Gaussian kernel profiles one parameter at a time.')
        self._problem = None

    @property
    def varying_param_names(self):
        return list(self.param['varying'].keys())

    @property
    def means(self):
        return np.array([self.config['means'][n] for n in self.names])

    @property
    def covmat(self):
        return self.config['covmat']

    def bounds(self):
        lo = [self.config['ranges'][n][0] for n in self.names]
        hi = [self.config['ranges'][n][1] for n in self.names]
        return lo, hi

    @property
    def fixed_index(self):
        fixed = list(self.param['fixed'].keys())
        return self.names.index(fixed[0]) if fixed else None

    # --- device problem -----------------------------------------------------------------------------
    def host_problem(self, fixed_values=None, proposal_covmats=None, likelihood_only=False):
        """FP64 host preparation for a set of profile points (one per fixed value) sharing this target.
        `likelihood_only`: no proposal factors (zeros) -- enough for pb200_loglkl_batch, nothing to sample with."""
        lo, hi = self.bounds()
        fi = self.fixed_index
        if fi is not None and fixed_values is None:
            fixed_values = [self.param['fixed'][self.names[fi]]['fixed_value']]
        n = self.dimension - (fi is not None)
        points = 1 if fi is None else len(fixed_values)
        if likelihood_only:
            proposal_covmats = None
        elif proposal_covmats is None:
            proposal_covmats = np.tile(self.get_default_covmat(), (points, 1, 1))
        return build_problem(self.means, self.covmat, lo, hi, fi, fixed_values, proposal_covmats,
                             ignore_prior=self.ignore_prior)

    def device_problem(self, fixed_values=None, proposal_covmats=None, likelihood_only=False):
        """Device image of the target + per-point proposals.  The host preparation (an inverse, and per profile point
        a Cholesky factor and a symmetric eigen-decomposition: ~30 ms per point at n = 255) and the upload (9 MB for 8
        points at n = 255) only depend on (fixed parameter, fixed values, proposal covariances, prior switch), so the
        last few images are kept per kernel object: building one optimiser after another for the same job (every
        call of the reference-facing API does) re-uses the resident problem instead of paying for it again -- that
        was 8x the device time of a 256-D job."""
        fi = self.fixed_index
        fv = None if fixed_values is None else np.ascontiguousarray(fixed_values, dtype=np.float64)
        pc = None if proposal_covmats is None else np.ascontiguousarray(proposal_covmats, dtype=np.float64)
        fixed_now = None if fi is None or fv is not None else self.param['fixed'][self.names[fi]]['fixed_value']
        key = (fi, fixed_now, None if fv is None else fv.tobytes(), array_key(pc), self.ignore_prior, str(self.device),
               bool(likelihood_only))
        cache = self.__dict__.setdefault('_problem_cache', {})
        hit = cache.get(key)
        if hit is None:
            hit = (DeviceProblem(self.host_problem(fixed_values, proposal_covmats, likelihood_only), self.device), pc)
            if len(cache) >= 4:
                cache.pop(next(iter(cache)))
            cache[key] = hit          # (the array travels with the entry: an identity key must not outlive its object)
        return hit[0]

    def _scalar_problem(self):
        if self._problem is None:
            self._problem = self.device_problem()
        return self._problem

    # --- likelihood / prior (base_kernel.py:72-112) --------------------------------------------------
    def _vector(self, position):
        return np.array([position[n][0] for n in self.varying_param_names], dtype=np.float64)

    def loglkl(self, prop):
        for name, par in self.param['fixed'].items():
            prop[name] = [par['fixed_value']]
        return float(self.loglkl_batch(self._vector(prop)[None, :])[0])

    def loglkl_batch(self, x, points=None, problem=None, return_outside=False):
        """x: [M, n] positions of the varying parameters (numpy) -> -log L [M] (numpy float64), on the GPU."""
        import torch
        from . import engine
        prob = self._scalar_problem() if problem is None else problem
        x = np.asarray(x, dtype=np.float64)
        xp = np.zeros((x.shape[0], prob.npad))
        xp[:, :prob.n] = x[:, :prob.n]
        pts = None if points is None else torch.as_tensor(np.asarray(points, dtype=np.int32), device=prob.device)
        ll, outside = engine.loglkl_batch(prob, torch.as_tensor(xp, device=prob.device), pts)
        if return_outside:
            return ll.cpu().numpy(), outside.cpu().numpy()
        return ll.cpu().numpy()

    def logprior(self, position):
        if self.ignore_prior:
            return 0.0
        return np.inf if self.outside_of_prior_bound(position) else 0.0

    def outside_of_prior_bound(self, prop):
        for name, val in prop.items():
            if name in self.param['varying']:
                lo, hi = self.param['varying'][name]['range']
                if lo is not None and val[0] < lo:
                    return True
                if hi is not None and val[0] > hi:
                    return True
        return False

    def logpost(self, position):
        return self.loglkl(position) + self.logprior(position)

    # --- defaults / user input (analytical_kernel.py:88-107) ----------------------------------------
    def get_default_initial_position(self):
        return {n: [0.] for n in self.varying_param_names}

    def read_initial_position(self, config_initial_position):
        return np.array(config_initial_position)

    def get_default_covmat(self):
        widths = [self.param['varying'][n]['range'][1] - self.param['varying'][n]['range'][0]
                  for n in self.varying_param_names]
        return 0.005 * np.diag(widths)

    def read_covmat(self, config_covmat):
        return np.array(config_covmat)

    # --- closed-form profile (replaces the BFGS of get_scipy_profile, analytical_kernel.py:116-122) --
    def get_analytic_profile(self, parameter, fixed_val):
        """min over the varying parameters of -log L at parameter = fixed_val, with the reference's residual
        ordering: 1/2 (s_bb - s_ab^T S_aa^-1 s_ab) f^2."""
        fi = self.names.index(parameter)
        n = self.dimension - 1
        a = np.linalg.inv(self.covmat)
        s = 0.5 * (a + a.T)
        f = fixed_val - self.means[fi]
        return float(0.5 * (s[n, n] - s[:n, n] @ np.linalg.solve(s[:n, :n], s[:n, n])) * f * f)

    get_scipy_profile = get_analytic_profile

    def finalize(self):
        self._problem = None
        self._problem_cache = {}
