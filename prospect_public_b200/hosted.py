"""Batched proposal / accept machinery for kernels whose likelihood is evaluated on the HOST.

External likelihood codes (MontePython, cobaya, any subclass of the reference's BaseKernel:
prospect/kernels/base_kernel.py:9-138) stay on the CPU.  Everything else of the optimiser's inner loop --
the Gaussian proposal of MetropolisHastings.get_proposal (prospect/mcmc.py:182-190), the prior-box test
(base_kernel.py:89-102), the accept rule of MetropolisHastings.step (mcmc.py:163-180), the running bestfit of
SimulatedAnnealing.set_bestfit (optimiser.py:74-85) and the temperature / step-size bookkeeping
(optimiser.py:87-162) -- is the same for every kernel and runs on the GPU for ALL (profile point x repetition)
chains of a job at once.  One MH step of the whole job is

    pb200_hosted_propose  ->  host: -log L of the proposals inside the box  ->  pb200_hosted_accept

so the host sees ONE batch of positions per step (which a vectorised or multi-threaded likelihood can take as a
whole, `HostKernelAdapter.loglkl_batch`) instead of one Python task per chain and iteration.

`HostKernelAdapter` wraps an unmodified reference kernel object; `HostedSimulatedAnnealing` keeps the interface
and the record layout of `optimiser.SimulatedAnnealing`, so run.py's analysis / writers / snapshots take its results
unchanged.  The device never evaluates a likelihood here and the host never proposes or accepts: there is no CPU
fallback for the batched part (capi.lib() raises without the CUDA library).
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import capi
from .optimiser import SimulatedAnnealing
from .problem import pad_of, proposal_factor, upload_packed


class HostKernelAdapter:
    """A reference-style kernel (`param`, `loglkl(position_dict)`, `logprior`, `set_fixed_parameters`, defaults) seen
    as a batched likelihood.

    kernel            one kernel object, or a list of identically configured ones (one per worker thread: external
                      codes keep per-object state, base_kernel.py:72-75 mutates `param['fixed']`); likelihood codes
                      that release the GIL then run `len(kernels)` evaluations at a time
    profile_parameter the parameter the job profiles (None: global optimisation / mcmc); it is moved to
                      `param['fixed']` once (base_kernel.py:46-51) and re-valued per chain
    batch_loglkl      optional callable(x [M, n], fixed [M] or None) -> [M]: a likelihood that is already vectorised
                      replaces the per-position loop
    """
    host_evaluated = True

    def __init__(self, kernel, profile_parameter=None, ignore_prior=None, batch_loglkl=None, device=None):
        self.kernels = list(kernel) if isinstance(kernel, (list, tuple)) else [kernel]
        self.kernel = self.kernels[0]
        self.profile_parameter = None
        self.batch_loglkl = batch_loglkl
        self.ignore_prior = bool(getattr(self.kernel, 'ignore_prior', False)) if ignore_prior is None \
            else bool(ignore_prior)
        self.device = device                      # None: the current CUDA device when the optimiser is built
        if profile_parameter is not None:
            self.set_fixed_parameters({profile_parameter: 0.0})
        self._pool = ThreadPoolExecutor(len(self.kernels)) if len(self.kernels) > 1 else None
        self.evaluations = 0

    def set_fixed_parameters(self, fixed_param_dict):
        """base_kernel.py:46-51 for every wrapped kernel object.  The (one) profiled parameter moves to `param['fixed']`
        the first time; afterwards its value is whatever the chain being evaluated needs (loglkl_batch)."""
        if len(fixed_param_dict) != 1:
            raise ValueError('One parameter is profiled at a time.')
        (name, value), = fixed_param_dict.items()
        if self.profile_parameter not in (None, name):
            raise ValueError(f"'{self.profile_parameter}' is already the profiled parameter of this kernel.")
        for k in self.kernels:
            if name in k.param['fixed']:
                k.param['fixed'][name]['fixed_value'] = value
            elif name in k.param['varying']:
                k.set_fixed_parameters({name: value})
            else:
                raise ValueError(f'The requested profile parameter {name} is not recognized by the kernel.')
        self.profile_parameter = name

    # the reference's plugin surface is the wrapped kernel's
    @property
    def param(self):
        return self.kernel.param

    @property
    def varying_param_names(self):
        return list(self.kernel.param['varying'].keys())

    def __getattr__(self, name):          # get_default_covmat, read_covmat, logprior, ... of the wrapped kernel
        if name in ('kernels', 'kernel', 'device'):
            raise AttributeError(name)
        return getattr(self.kernel, name)

    def bounds(self):
        """(lo [n], hi [n]) of the varying parameters, open sides / ignore_prior as -inf / +inf."""
        names = self.varying_param_names
        lo, hi = np.full(len(names), -np.inf), np.full(len(names), np.inf)
        if not self.ignore_prior:
            for i, nm in enumerate(names):
                rng = self.kernel.param['varying'][nm].get('range', [None, None])
                if rng[0] is not None:
                    lo[i] = rng[0]
                if rng[1] is not None:
                    hi[i] = rng[1]
        return lo, hi

    def _one(self, kernel, names, row, fixed_value):
        if self.profile_parameter is not None:
            kernel.param['fixed'][self.profile_parameter]['fixed_value'] = fixed_value
        prop = {nm: [float(v)] for nm, v in zip(names, row)}
        if not self.ignore_prior and kernel.logprior(prop) == np.inf:
            return np.inf                 # a prior the box does not describe: rejected like mcmc.py:166-167
        value = kernel.loglkl(prop)       # inserts the fixed parameter, maps computation exceptions to inf
        return np.inf if value is None or np.isnan(value) else float(value)

    def loglkl_batch(self, x, fixed_values=None, mask=None):
        """-log L of the rows of x [M, n] (varying order); fixed_values [M] values of the profiled parameter;
        rows with mask == False are not evaluated (inf)."""
        x = np.asarray(x, dtype=np.float64)
        m = x.shape[0]
        out = np.full(m, np.inf)
        idx = np.arange(m) if mask is None else np.flatnonzero(mask)
        if len(idx) == 0:
            return out
        fixed = None if fixed_values is None else np.asarray(fixed_values, dtype=np.float64)
        self.evaluations += len(idx)
        if self.batch_loglkl is not None:
            out[idx] = np.asarray(self.batch_loglkl(x[idx], None if fixed is None else fixed[idx]), dtype=np.float64)
            out[np.isnan(out)] = np.inf
            return out
        names = self.varying_param_names

        def work(kernel, part):
            return [self._one(kernel, names, x[i], None if fixed is None else float(fixed[i])) for i in part]

        if self._pool is None:
            out[idx] = work(self.kernel, idx)
        else:
            parts = np.array_split(idx, len(self.kernels))
            for part, vals in zip(parts, self._pool.map(work, self.kernels, parts)):
                out[part] = vals
        return out

    def finalize(self):
        if self._pool is not None:
            self._pool.shutdown()
        for k in self.kernels:
            k.finalize()


class HostedProblem:
    """What the hosted kernels need on the device (struct pb200_hosted): the proposal factors of this rank's profile
    points, F F^T = covmat (tasks/initialise_optimiser_task.py:126-130), and the prior box."""

    def __init__(self, n, covmats, lo, hi, device):
        import torch
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise capi.Pb200Error('prospect_public_b200 runs on CUDA devices only (no CPU fallback)')
        self.n, self.npad = int(n), pad_of(n)
        covmats = np.asarray(covmats, dtype=np.float64).reshape(-1, n, n)
        self.n_points = covmats.shape[0]
        factor_t = np.zeros((self.n_points, self.npad, self.npad))
        for p in range(self.n_points):
            factor_t[p, :n, :n] = proposal_factor(covmats[p]).T
        lo_p, hi_p = np.full(self.npad, -np.inf), np.full(self.npad, np.inf)
        lo_p[:n], hi_p[:n] = lo, hi
        self.host_factor_t = factor_t
        self.t = upload_packed({'factor_t': factor_t, 'lo': lo_p, 'hi': hi_p}, self.device)
        self._storage = self.t.pop('_storage')
        self.c = capi.Hosted(self.n, self.npad, self.n_points, capi.ptr(self.t['factor_t']), capi.ptr(self.t['lo']),
                             capi.ptr(self.t['hi']))


class HostedSimulatedAnnealing(SimulatedAnnealing):
    """optimiser.SimulatedAnnealing for a HostKernelAdapter: same settings, verbs, records and multi-GPU sharding;
    one iteration = steps_per_iteration x (propose launch, host likelihood batch, accept launch) + the device-side
    schedule update (pb200_schedule_update)."""

    def __init__(self, config, kernel, optimise_settings, keep_positions=True, device=None):
        import torch
        self._device = device or getattr(kernel, 'device', None) or f'cuda:{torch.cuda.current_device()}'
        super().__init__(config, kernel, optimise_settings, keep_positions=keep_positions)
        self.fused_exchange = False           # several GPUs: one launch set + one NCCL all-gather per iteration
        self.signalled = False
        dev, b, npad = self.problem.device, self.batch.n_chains, self.problem.npad
        self._prop = torch.zeros((b, npad), dtype=torch.float64, device=dev)
        self._outside = torch.zeros(b, dtype=torch.int32, device=dev)
        self._prop_ll = torch.zeros(b, dtype=torch.float64, device=dev)
        self._nacc = torch.zeros(b, dtype=torch.int32, device=dev)
        self._best_ll = torch.zeros(b, dtype=torch.float64, device=dev)
        self._best_x = torch.zeros((b, npad), dtype=torch.float64, device=dev)
        self._prop_host = torch.empty((b, npad), dtype=torch.float64, pin_memory=True)
        self._outside_host = torch.empty(b, dtype=torch.int32, pin_memory=True)
        self._ll_host = torch.empty(b, dtype=torch.float64, pin_memory=True)
        fixed = self.settings.get('fixed_param_val')
        self._fixed_chain = None if fixed is None else np.asarray(fixed, dtype=np.float64)[self.point_global]
        if self.settings.get('chain_state') is None and b:
            # loglkl of the start points (optimiser.py:55), one host batch
            ll0 = kernel.loglkl_batch(self.batch.positions(), self._fixed_chain)
            self.batch.loglkl.copy_(torch.as_tensor(ll0, device=dev))

    def _device_problem(self, kernel, fixed_local, covmats_local):
        lo, hi = kernel.bounds()
        return HostedProblem(len(kernel.varying_param_names), covmats_local, lo, hi, self._device)

    def step(self, status_host=None):
        """One Metropolis-Hastings step of every active chain (mcmc.py:163-180), on the current stream."""
        import torch
        from . import engine
        b, n = self.batch.n_chains, self.problem.n
        if b == 0:
            return
        if status_host is None:
            status_host = self.batch.status.cpu().numpy()
        engine.hosted_propose(self.problem, self.batch, self.seed, self._prop, self._outside)
        self._prop_host.copy_(self._prop, non_blocking=True)
        self._outside_host.copy_(self._outside, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        evaluate = (status_host == capi.ACTIVE) & (self._outside_host.numpy() == 0)
        ll = self.kernel.loglkl_batch(self._prop_host.numpy()[:, :n], self._fixed_chain, evaluate)
        self._ll_host.numpy()[:] = ll
        self._prop_ll.copy_(self._ll_host, non_blocking=True)
        engine.hosted_accept(self.problem, self.batch, self.seed, self._prop, self._prop_ll, self._outside, None,
                             self._nacc, self._best_ll, self._best_x)

    def optimise(self, n_iterations=1, stream=None):
        """`n_iterations` annealing iterations of every chain (optimise_task.py:16-27 for the whole job).  Runs on the
        current stream (`stream` is accepted for interface compatibility and must be None)."""
        if stream is not None:
            raise ValueError('the hosted optimiser synchronises with the host every step: use the current stream')
        import torch
        from . import engine
        b = self.batch.n_chains
        for _ in range(n_iterations):
            row = self.iterations_done
            status_host = self.batch.status.cpu().numpy() if b else None
            if b and (status_host == capi.ACTIVE).any():          # (every chain stopped: nothing left to run)
                live = self.batch.status == capi.ACTIVE
                # the chain of an iteration starts at the current point: it is the first bestfit candidate
                self._best_ll.copy_(self.batch.loglkl)
                self._best_x.copy_(self.batch.x)
                self._nacc.zero_()
                for _ in range(self.schedule.steps_per_iteration):
                    self.step(status_host)
                if row < self.max_rows:
                    r = self.records
                    for dst, src in ((r.best_loglkl, self._best_ll), (r.last_loglkl, self.batch.loglkl),
                                     (r.temperature, self.batch.temperature), (r.step_size, self.batch.step_size),
                                     (r.accepted, self._nacc)):
                        dst[row] = torch.where(live, src, dst[row])
                    if r.best_x is not None:
                        r.best_x[row] = torch.where(live[:, None], self._best_x, r.best_x[row])
                        r.last_x[row] = torch.where(live[:, None], self.batch.x, r.last_x[row])
                engine.schedule_update(self.batch, self.schedule, self._nacc, self._best_ll)
            self.iterations_done += 1
            self.step_counter += self.schedule.steps_per_iteration


class HostedMetropolisHastings:
    """mcmc.MetropolisHastings for a HostKernelAdapter (prospect/mcmc.py:104-190 for `n_chains` chains at once): the
    proposal, box test and accept rule run on the device, the likelihood on the host, and the accepted points are
    recorded on the host (they are already there: the host evaluated them) as rows (multiplicity, -log L, position),
    the reference's Chain layout (mcmc.py:25-82).  Same constructor arguments as mcmc.MetropolisHastings
    (n_chains, seed, chain_ids, chains, steps_done)."""

    def __init__(self, config_mcmc, kernel, **mcmc_args):
        import torch
        from . import engine
        self.config_mcmc, self.kernel = config_mcmc, kernel
        self.seed = int(mcmc_args.get('seed', 0))
        names = kernel.varying_param_names
        n = len(names)
        if config_mcmc.start_from_covmat is not None:
            cm = config_mcmc.start_from_covmat
            self.covmat = np.array(cm) if not isinstance(cm, str) else kernel.read_covmat(cm)
        else:
            self.covmat = kernel.get_default_covmat()
        prior = mcmc_args.get('chains')
        self.n_chains = b = len(prior) if prior is not None else int(mcmc_args.get('n_chains', config_mcmc.N_chains))
        if prior is not None:
            x0 = np.array([[ch.positions[nm][-1] for nm in names] for ch in prior], dtype=np.float64)
        elif config_mcmc.start_from_position is not None:
            x0 = np.tile(np.asarray(kernel.read_initial_position(config_mcmc.start_from_position), dtype=float), (b, 1))
        else:
            start = kernel.get_default_initial_position()
            x0 = np.tile(np.array([start[nm][0] for nm in names], dtype=np.float64), (b, 1))
        device = mcmc_args.get('device') or getattr(kernel, 'device', None) or f'cuda:{torch.cuda.current_device()}'
        lo, hi = kernel.bounds()
        self.problem = HostedProblem(n, np.asarray(self.covmat, dtype=np.float64)[None], lo, hi, device)
        self.batch = engine.ChainBatch(self.problem, x0, 0, chain_id=mcmc_args.get('chain_ids'),
                                       temperature=config_mcmc.temperature, step_size=config_mcmc.step_size,
                                       steps_done=mcmc_args.get('steps_done', 0))
        fixed = kernel.param['fixed']
        self._fixed = None if not fixed else np.full(b, float(next(iter(fixed.values()))['fixed_value']))
        # the chain starts with its first point (mcmc.py:118-129); a continued chain knows that point's likelihood
        ll0 = (np.array([ch.loglkls[-1] for ch in prior], dtype=np.float64) if prior is not None
               else kernel.loglkl_batch(x0, self._fixed))
        dev = self.problem.device
        self.batch.loglkl.copy_(torch.as_tensor(ll0, device=dev))
        npad = self.problem.npad
        self._prop = torch.zeros((b, npad), dtype=torch.float64, device=dev)
        self._outside = torch.zeros(b, dtype=torch.int32, device=dev)
        self._prop_ll = torch.zeros(b, dtype=torch.float64, device=dev)
        self._flag = torch.zeros(b, dtype=torch.int32, device=dev)
        self._nacc = torch.zeros(b, dtype=torch.int32, device=dev)
        self._prop_host = torch.empty((b, npad), dtype=torch.float64, pin_memory=True)
        self._outside_host = torch.empty(b, dtype=torch.int32, pin_memory=True)
        self._flag_host = torch.empty(b, dtype=torch.int32, pin_memory=True)
        self._ll_host = torch.empty(b, dtype=torch.float64, pin_memory=True)
        self._names, self._prior = names, prior
        if prior is not None and mcmc_args.get('preload', False):
            # continue the chains in place (`prospect <folder>` for jobtype mcmc, tasks/mcmc_task.py:71-81): their
            # earlier rows go back into the record, the open last row keeps counting
            rows = np.array([len(ch.mults) for ch in prior], dtype=np.int64)
            cap = int(rows.max()) + 64
            self._x, self._ll = np.zeros((b, cap, n)), np.zeros((b, cap))
            self._mult, self._count = np.zeros((b, cap), dtype=np.int64), rows
            for c, ch in enumerate(prior):
                self._x[c, :rows[c]] = np.stack([np.asarray(ch.positions[nm], dtype=np.float64) for nm in names], axis=1)
                self._ll[c, :rows[c]] = ch.loglkls
                self._mult[c, :rows[c]] = np.asarray(ch.mults, dtype=np.int64)
            self._prior = None
        else:
            self._x = np.zeros((b, 64, n))
            self._ll = np.zeros((b, 64))
            self._mult = np.zeros((b, 64), dtype=np.int64)
            self._count = np.ones(b, dtype=np.int64)
            self._x[:, 0], self._ll[:, 0], self._mult[:, 0] = x0, ll0, 1
        self.accepted = np.zeros(b, dtype=np.int64)
        self.steps = 0

    def _grow(self, need):
        cap = self._ll.shape[1]
        if need <= cap:
            return
        new = max(2 * cap, need)
        for name in ('_x', '_ll', '_mult'):
            old = getattr(self, name)
            big = np.zeros((old.shape[0], new) + old.shape[2:], dtype=old.dtype)
            big[:, :cap] = old
            setattr(self, name, big)

    def step(self):
        """MetropolisHastings.step (mcmc.py:163-180) of every chain."""
        import torch
        from . import engine
        b, n = self.n_chains, self.problem.n
        if b == 0:
            return
        engine.hosted_propose(self.problem, self.batch, self.seed, self._prop, self._outside)
        self._prop_host.copy_(self._prop, non_blocking=True)
        self._outside_host.copy_(self._outside, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        inside = self._outside_host.numpy() == 0
        prop = self._prop_host.numpy()[:, :n]
        ll = self.kernel.loglkl_batch(prop, self._fixed, inside)
        self._ll_host.numpy()[:] = ll
        self._prop_ll.copy_(self._ll_host, non_blocking=True)
        engine.hosted_accept(self.problem, self.batch, self.seed, self._prop, self._prop_ll, self._outside, self._flag,
                             self._nacc)
        self._flag_host.copy_(self._flag, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        acc = self._flag_host.numpy() != 0
        # Chain.push_position for the accepted proposals, mults[-1] += 1 for the others (mcmc.py:176-180)
        self._grow(int(self._count.max()) + 1)
        idx = np.flatnonzero(acc)
        rows = self._count[idx]
        self._x[idx, rows], self._ll[idx, rows], self._mult[idx, rows] = prop[idx], ll[idx], 1
        self._count[idx] += 1
        rej = np.flatnonzero(~acc)
        self._mult[rej, self._count[rej] - 1] += 1
        self.accepted += acc
        self.steps += 1

    def run_steps(self, steps_per_iteration):
        for _ in range(int(steps_per_iteration)):
            self.step()

    def run_minutes(self, minutes_per_iteration):
        """BaseMCMC.run_minutes (mcmc.py:140-151): step until the wall-clock budget is used up."""
        import time
        t_end = time.perf_counter() + 60.0 * float(minutes_per_iteration)
        done = 0
        while True:
            self.step()
            done += 1
            if time.perf_counter() >= t_end:
                return done

    def samples_numpy(self):
        """(x [B, cap, n], loglkl [B, cap], mult [B, cap], count [B]) of everything recorded so far."""
        return self._x, self._ll, self._mult.astype(np.int32), self._count.astype(np.int32)

    @property
    def _samples(self):
        """The record as device buffers in engine.Samples' layout (what engine.chain_statistics reduces for R-1 and
        the pooled covariance): uploaded when asked for, once per round."""
        import torch
        from types import SimpleNamespace
        dev, n = self.problem.device, self.problem.n
        cap = int(self._count.max())
        x = torch.zeros((self.n_chains, cap, self.problem.npad), dtype=torch.float64, device=dev)
        x[:, :, :n] = torch.as_tensor(self._x[:, :cap], device=dev)
        return SimpleNamespace(x=x, loglkl=torch.as_tensor(self._ll[:, :cap], device=dev),
                               mult=torch.as_tensor(self._mult[:, :cap].astype(np.int32), device=dev),
                               count=torch.as_tensor(self._count.astype(np.int32), device=dev), capacity=cap)

    @property
    def chains(self):
        """list[mcmc.Chain]; a continued chain is its earlier rows followed by the new ones, the open last row of the
        earlier segment keeps its multiplicity count (Chain.mults[-1] += 1)."""
        from .mcmc import Chain, continue_chain
        fixed = {nm: par['fixed_value'] for nm, par in self.kernel.param['fixed'].items()}
        out = []
        for c in range(self.n_chains):
            r = int(self._count[c])
            ch = Chain.from_arrays(self._names, self._mult[c, :r], self._ll[c, :r], self._x[c, :r])
            for nm, val in fixed.items():
                ch.positions[nm] = [val] * r
            if self._prior is not None:
                ch = continue_chain(self._prior[c], ch)
            out.append(ch)
        return out

    @property
    def chain(self):
        return self.chains[0]

    def finalize(self):
        pass
