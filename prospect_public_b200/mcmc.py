"""Chain container, covariance estimator and the batched Metropolis-Hastings sampler.

Mirrors prospect/mcmc.py: `Chain` (:25-82), `collapse_chains` (:84-91), `compute_covariance_matrix`
(:93-97), `initialise_mcmc` (:13-18) and `MetropolisHastings` (:104-190) keep their names and meaning.
What changes is the unit of work: one sampler object advances MANY independent chains per call -- one
GPU warp each (pb200_mh_run) -- instead of one Python-loop chain per process.
"""
from collections import defaultdict

import numpy as np

from . import capi


class Chain:
    """Accepted points with multiplicities, -log L values and per-parameter positions (MontePython layout)."""

    def __init__(self, initial_mults=None, initial_loglkls=None, initial_positions=None):
        self.mults, self.loglkls, self.positions = [], [], defaultdict(list)
        if initial_mults is not None and initial_loglkls is not None and initial_positions is not None:
            self.mults += list(initial_mults)
            self.loglkls += list(initial_loglkls)
            for name, vec in initial_positions.items():
                self.positions[name] += list(vec)

    @classmethod
    def from_arrays(cls, names, mults, loglkls, x):
        ch = cls()
        ch.mults = list(np.asarray(mults))
        ch.loglkls = list(np.asarray(loglkls))
        for i, name in enumerate(names):
            ch.positions[name] = list(x[:, i])
        return ch

    def last_varying_position(self, names):
        return {n: [self.positions[n][-1]] for n in names}

    @property
    def last_position(self):
        return {n: [vec[-1]] for n, vec in self.positions.items()}

    @property
    def N(self):
        lengths = {len(self.mults), len(self.loglkls)} | {len(v) for v in self.positions.values()}
        if len(lengths) != 1:
            raise ValueError('Invalid dimensions in Chain: Non-compatible length of mults, loglkls and positions.')
        return len(self.mults)

    @property
    def data(self):
        """[N_points, 2 + N_params] array: multiplicity, -log L, parameters."""
        pos = np.array(list(self.positions.values()))
        return np.concatenate([np.stack([self.mults, self.loglkls], axis=0), pos], axis=0).T

    def __getitem__(self, idx):
        return {n: [vec[idx]] for n, vec in self.positions.items()}

    def push_position(self, new_loglkl, new_position):
        self.mults.append(1)
        self.loglkls.append(new_loglkl)
        for n, vec in self.positions.items():
            vec.append(new_position[n][0])

    def from_indices(self, index_list):
        pos = {n: np.take(vec, index_list) for n, vec in self.positions.items()}
        return Chain(list(np.take(self.mults, index_list)), list(np.take(self.loglkls, index_list)), pos)


def collapse_chains(chain_list):
    total = Chain()
    for ch in chain_list:
        total.mults += list(ch.mults)
        total.loglkls += list(ch.loglkls)
        for n, vec in ch.positions.items():
            total.positions[n] += list(vec)
    return total


def continue_chain(previous, segment):
    """The chain `previous` continued by `segment`, whose first row is previous' last point again (a sampler that was
    started from it): that row's multiplicity count carries on in previous' open last row (Chain.mults[-1] += 1,
    prospect/mcmc.py:178-180), the other rows follow.  Neither input is modified."""
    merged = Chain(list(previous.mults), list(previous.loglkls), {k: list(v) for k, v in previous.positions.items()})
    merged.mults[-1] += segment.mults[0] - 1
    merged.mults += segment.mults[1:]
    merged.loglkls += segment.loglkls[1:]
    for k in merged.positions:
        merged.positions[k] += segment.positions[k][1:]
    return merged


def compute_covariance_matrix(chain_list, device='cuda:0'):
    """np.cov(x, bias=True, fweights=mults) of the collapsed chains, reduced on the GPU in FP64
    (pb200_weighted_moments)."""
    import torch
    from . import engine
    total = collapse_chains(chain_list)
    data = total.data
    x = torch.as_tensor(np.ascontiguousarray(data[:, 2:]), device=device)
    w = torch.as_tensor(np.asarray(total.mults).astype(np.int32), device=device)
    cov = engine.covariance_from_moments(*engine.weighted_moments(x, x.shape[1], w))
    return cov.cpu().numpy()


def initialise_mcmc(config_mcmc, kernel, **mcmc_args):
    """prospect/mcmc.py:13-18.  A kernel whose likelihood is evaluated on the host (hosted.HostKernelAdapter) gets the
    sampler that batches everything BUT the likelihood call."""
    if config_mcmc.algorithm in ('MetropolisHastings', 'Metropolis Hastings'):
        if getattr(kernel, 'host_evaluated', False):
            from .hosted import HostedMetropolisHastings
            return HostedMetropolisHastings(config_mcmc, kernel, **mcmc_args)
        return MetropolisHastings(config_mcmc, kernel, **mcmc_args)
    raise ValueError('You have specified a non-existing MCMC type.')


class MetropolisHastings:
    """`n_chains` independent MH chains on one GPU.

    config_mcmc: the `mcmc:` section (temperature, step_size, start_from_covmat, start_from_position,
    steps_per_iteration).  mcmc_args: n_chains (default config_mcmc.N_chains), seed, chain_ids (global ids for
    the Philox streams), chains (list[Chain] to continue), record ('accepted' | 'thinned' | None), thin."""

    def __init__(self, config_mcmc, kernel, **mcmc_args):
        from . import engine
        self.config_mcmc = config_mcmc
        self.kernel = kernel
        self.seed = int(mcmc_args.get('seed', 0))
        names = kernel.varying_param_names
        n = len(names)
        if config_mcmc.start_from_covmat is not None:
            cm = config_mcmc.start_from_covmat
            self.covmat = np.array(cm) if not isinstance(cm, str) else kernel.read_covmat(cm)
        else:
            self.covmat = kernel.get_default_covmat()
        prior = mcmc_args.get('chains')
        self.n_chains = len(prior) if prior is not None else int(mcmc_args.get('n_chains', config_mcmc.N_chains))
        b = self.n_chains
        if prior is not None:
            x0 = np.array([[ch.positions[nm][-1] for nm in names] for ch in prior], dtype=np.float64)
        elif config_mcmc.start_from_position is not None:
            x0 = np.tile(np.asarray(kernel.read_initial_position(config_mcmc.start_from_position), dtype=float), (b, 1))
        else:
            start = kernel.get_default_initial_position()
            x0 = np.tile(np.array([start[nm][0] for nm in names], dtype=np.float64), (b, 1))
        self.problem = kernel.device_problem(proposal_covmats=np.asarray(self.covmat)[None])
        ids = mcmc_args.get('chain_ids')
        self.batch = engine.ChainBatch(self.problem, x0, 0, chain_id=ids, temperature=config_mcmc.temperature,
                                       step_size=config_mcmc.step_size,
                                       steps_done=mcmc_args.get('steps_done', 0))
        # Prior-box test: deferred by default -- a Philox block whose four steps provably stay inside the whitened safe
        # radius needs no test at all; otherwise the positions the block would accept are tested in place (a TF32
        # tensor-core filter in front of the exact FP32 mat-vec, see box_filter_mma in csrc/pb200_kernels.cu).
        # `exact_box=True` materialises and tests every likelihood-accepted step instead -- same decisions; measured at
        # 65 536 chains, step 2.4 / sqrt(d): 7.0e9 exact vs 1.12e10 deferred chain-steps/s.  Recording accepted points
        # always materialises every accepted step.
        exact = mcmc_args.get('exact_box', 'auto')
        if exact == 'auto':
            exact = False
        self.exact_box = bool(exact)
        # lanes per chain: chosen from the size of the WHOLE job (config N_chains), not of this rank's share, so that a
        # chain's arithmetic -- and with it every chain file -- does not depend on the number of GPUs
        self.lanes = int(mcmc_args.get('lanes_per_chain', 0)) or engine.choose_lanes(
            self.problem.npad, max(b, int(mcmc_args.get('n_chains_global', config_mcmc.N_chains or b))))
        self.record = mcmc_args.get('record', 'accepted')
        self.thin = int(mcmc_args.get('thin', 1))
        self._prior = prior
        self._samples = None
        self._names = names
        self.accepted = np.zeros(b, dtype=np.int64)
        self.steps = 0
        if prior is not None and mcmc_args.get('preload', False):
            self._preload(prior)

    def _preload(self, prior):
        """Continue chains in place (`prospect <folder>` for jobtype mcmc, tasks/mcmc_task.py:71-81): the rows of the
        earlier segments go back into the device sample buffers, so the recording cursor picks up at each chain's open
        last row and everything downstream (chain files, pooled covariance, R-1) sees ONE chain per index, exactly as
        if the run had never been interrupted."""
        import torch
        from . import engine
        if self.record != 'accepted':
            raise ValueError("chains can only be continued in place with record='accepted'")
        n, names = self.problem.n, self._names
        rows = [len(ch.mults) for ch in prior]
        cap = max(rows)
        x = np.zeros((self.n_chains, cap, self.problem.npad))
        ll = np.zeros((self.n_chains, cap))
        mult = np.zeros((self.n_chains, cap), dtype=np.int32)
        for c, ch in enumerate(prior):
            x[c, :rows[c], :n] = np.stack([np.asarray(ch.positions[nm], dtype=np.float64) for nm in names], axis=1)
            ll[c, :rows[c]] = ch.loglkls
            mult[c, :rows[c]] = np.asarray(ch.mults, dtype=np.int64)
        smp = engine.Samples(self.problem, self.n_chains, cap, capi.RECORD_ACCEPTED, self.thin)
        dev = self.problem.device
        smp.x.copy_(torch.as_tensor(x, device=dev))
        smp.loglkl.copy_(torch.as_tensor(ll, device=dev))
        smp.mult.copy_(torch.as_tensor(mult, device=dev))
        smp.count.copy_(torch.as_tensor(np.asarray(rows, dtype=np.int32), device=dev))
        self._samples, self._used, self._prior = smp, 0, None

    def _ensure_samples(self, n_steps):
        from . import engine
        if self.record is None:
            return None
        mode = capi.RECORD_ACCEPTED if self.record == 'accepted' else capi.RECORD_THINNED
        need = n_steps + 1 if mode == capi.RECORD_ACCEPTED else n_steps // self.thin + 1
        if self._samples is None:
            # thinned rows are few: room for four rounds at once (accepted points: worst case of one round, n_steps + 1)
            first = 4 * need if mode == capi.RECORD_THINNED else need
            self._samples = engine.Samples(self.problem, self.n_chains, first, mode, self.thin)
            self._used = 0
        else:
            # grow geometrically (every round of a long run would otherwise re-allocate and copy everything recorded so
            # far): copy the rows into a buffer of at least twice the size
            old = self._samples
            used = int(old.count.max().item())
            if used + need <= old.capacity:
                return old
            cap = max(2 * old.capacity, used + need)
            new = engine.Samples(self.problem, self.n_chains, cap, mode, self.thin)
            new.x[:, :old.capacity] = old.x
            new.loglkl[:, :old.capacity] = old.loglkl
            new.mult[:, :old.capacity] = old.mult
            new.count.copy_(old.count)
            self._samples = new
        return self._samples

    def step(self):
        self.run_steps(1)

    def run_steps(self, steps_per_iteration):
        from . import engine
        smp = self._ensure_samples(steps_per_iteration)
        acc = engine.mh_run(self.problem, self.batch, steps_per_iteration, self.seed, smp, lanes=self.lanes,
                            exact_box=self.exact_box)
        self.accepted += acc.cpu().numpy()
        self.steps += steps_per_iteration

    def run_minutes(self, minutes_per_iteration, chunk=1024):
        """BaseMCMC.run_minutes (prospect/mcmc.py:140-151): keep stepping until the wall-clock budget is used up.  The
        reference checks the clock after every step of its one chain; here every chain of the batch advances `chunk`
        steps per launch and the clock is read between launches, so all chains take the same number of steps (a
        multiple of `chunk`) -- how many depends on the hardware, as it does in the reference."""
        import time
        import torch
        t_end = time.perf_counter() + 60.0 * float(minutes_per_iteration)
        done = 0
        while True:
            self.run_steps(int(chunk))
            done += int(chunk)
            torch.cuda.synchronize()
            if time.perf_counter() >= t_end:
                return done

    def samples_numpy(self):
        """(x [B, cap, n], loglkl [B, cap], mult [B, cap], count [B]) of everything recorded so far."""
        s = self._samples
        n = self.problem.n
        cnt = s.count.cpu().numpy()
        if np.any(cnt > s.capacity):
            raise capi.Pb200Error('sample buffer overflow')
        return s.x[:, :, :n].cpu().numpy(), s.loglkl.cpu().numpy(), s.mult.cpu().numpy(), cnt

    @property
    def chains(self):
        """list[Chain]; continued chains are the earlier rows followed by the new ones (the open last row of
        the earlier segment keeps accumulating multiplicity, exactly like Chain.mults[-1] += 1)."""
        x, ll, mult, cnt = self.samples_numpy()
        fixed = {nm: par['fixed_value'] for nm, par in self.kernel.param['fixed'].items()}
        out = []
        for c in range(self.n_chains):
            r = cnt[c]
            ch = Chain.from_arrays(self._names, mult[c, :r], ll[c, :r], x[c, :r])
            for nm, val in fixed.items():
                ch.positions[nm] = [val] * r
            if self._prior is not None:
                ch = continue_chain(self._prior[c], ch)     # row 0 of the new segment is the previous chain's last point
            out.append(ch)
        return out

    @property
    def chain(self):
        return self.chains[0]

    def finalize(self):
        pass
