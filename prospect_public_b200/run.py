"""Entry points: `run(input.yaml | folder)`, `analyse(folder)`, CLI `python -m prospect_public_b200`.

Replaces prospect/run.py:19-117 and the task farm of prospect/communication.py:28-219 with GPU batch
dispatch: all (profile value x repetition) chains of a job are one device batch per rank; iterations are
kernel launches; ranks (one per GPU, torchrun) own fixed chain blocks and all-gather records at iteration
boundaries.  Python stays in charge of input, initialisation from the starting MCMC, analysis and files.

Deliberate differences from the reference:
  * a run stops after `max_iterations` iterations (k = 0 .. max_iterations-1) or when every chain met
    `chi2_tolerance`; the reference never terminates and its finaliser crashes;
  * repetitions are statistically independent (Philox streams keyed by the global chain id), whatever
    the number of GPUs; forked reference workers replay one RNG stream;
  * a chain whose 'adaptive' multiplier bookkeeping hits the reference's unbound-local path stops with
    status FAILED instead of aborting the whole serial run.
"""
import os
import sys
import time

import numpy as np

from . import analysis, capi
from . import io as pio
from .config import Configuration
from .distributed import dist_info, init_from_env
from .initialise import initialise_optimiser
from .kernels import initialise_kernel
from .optimiser import SimulatedAnnealing, get_initial_step_size_change, get_temperature_change

STATE_FORMAT = 'pb200-state-v1'


def _device():
    import torch
    if not torch.cuda.is_available():
        raise capi.Pb200Error('No CUDA device: prospect_public_b200 has no CPU fallback.')
    return f'cuda:{torch.cuda.current_device()}'


def _barrier():
    rank, world, active = dist_info()
    if active and world > 1:
        import torch.distributed as dist
        dist.barrier()


def run(user_inp, quiet=True):
    """Run a job.  Returns a dict with the records, the analysis products and timings."""
    t_start = time.perf_counter()
    config_yaml, resume = pio.read_user_input(user_inp) if isinstance(user_inp, str) else (user_inp, False)
    rank, world, _ = init_from_env('nccl') if int(os.environ.get('WORLD_SIZE', '1')) > 1 else dist_info()
    if resume:
        raise NotImplementedError('Resuming a run folder is not part of this round (SURVEY.md section 8f, N2).')
    config = Configuration(config_yaml)
    if rank == 0:
        pio.prepare_run(config)
    _barrier()
    if config.run.jobtype == 'mcmc':
        out = run_mcmc(config)
    else:
        out = run_annealing(config)
    out['time_to_result_s'] = time.perf_counter() - t_start
    out['config'] = config
    return out


def run_annealing(config, keep_positions=True):
    """jobtype profile / global_optimisation."""
    import torch
    rank, world, _ = dist_info()
    device = _device()
    prof = config.profile
    if config.run.numpy_random_seed is not None:
        np.random.seed(int(config.run.numpy_random_seed))     # only used if a random kernel must be generated
    out_dir = config.io.dir if config.io.write else None
    if rank == 0:
        kernel = initialise_kernel(config.kernel, out_dir, 0, device)   # may create analytical/random_gauss.pkl
    _barrier()
    if rank != 0:
        kernel = initialise_kernel(config.kernel, out_dir, rank, device)
    starts, values = initialise_optimiser(config, kernel)
    n_iter = int(prof.max_iterations)
    reps = int(prof.repetitions)
    settings = {
        'current_best_loglkl': starts.initial_loglkl,
        'initial_position': starts.positions,
        'covmat': starts.covmats,
        'temperature': prof.temperature_range[0],
        'temperature_change': get_temperature_change(config),
        'step_size': prof.step_size_range[0],
        'step_size_change': get_initial_step_size_change(config),
        'iteration_number': 0,
        'repetitions': reps,
        'fixed_param_val': values if config.run.jobtype == 'profile' else None,
        'max_rows': n_iter,
        'seed': config.seed,
    }
    sa = SimulatedAnnealing(config, kernel, settings, keep_positions=keep_positions)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gathered = sa.run_schedule(n_iter)
    torch.cuda.synchronize()
    loop_s = time.perf_counter() - t0
    g = sa.global_records()
    best_ll, acc, step, temperature, best_x = g['best_loglkl'], g['accepted'], g['step_size'], g['temperature'], g['best_x']
    status = sa.global_status()
    ran = acc >= 0
    chain_steps = int(ran.sum()) * int(prof.steps_per_iteration) * 1       # global (records are gathered)
    rec = analysis.RunRecord(values, reps, starts.names, best_ll, acc, best_x, status == capi.CONVERGED,
                             starts.initial_loglkl, starts.positions, starts.covmats, temperature, step)
    result = {'record': rec, 'chain_steps': chain_steps, 'loop_seconds': loop_s, 'optimiser': sa,
              'status': status, 'values': values}
    if rank == 0:
        if config.run.jobtype == 'profile':
            sub = os.path.join(config.io.dir, 'profile') if config.io.write else None
            if sub is not None:
                reduced = None
                if world == 1:
                    reduced = {k: v.cpu().numpy() for k, v in sa.reduce_over_iterations().items()}
                srt, profile, intervals = analysis.analyse_profile(rec, sub, prof.parameter, prof.confidence_levels,
                                                                   reduced)
            else:
                srt = analysis.sort_bestfits(rec)
                profile = analysis.compute_profile(rec, srt)
                intervals = analysis.compute_intervals_neyman(profile, prof.confidence_levels)
            result.update({'bestfits': srt, 'profile': profile, 'intervals': intervals})
        else:
            if config.io.write:
                result['bestfits'] = analysis.analyse_global_optimisation(
                    rec, os.path.join(config.io.dir, 'global_optimisation'))
            else:
                cb, ci = rec.per_chain_best()
                result['bestfits'] = {'chain_best': cb, 'chain_iter': ci, 'best_rep': int(np.argmin(cb[:, 0]))}
        if config.io.write:
            dump_snapshot(config, rec)
    _barrier()
    return result


def dump_snapshot(config, rec):
    """config.pkl + state.pkl (dense records; a few MB instead of one pickled chain per task)."""
    pio.dump_pickle(config, config.io.dir, 'config')
    state = {'format': STATE_FORMAT, 'jobtype': config.run.jobtype,
             'record': {k: getattr(rec, k) for k in ('values', 'reps', 'names', 'best_loglkl', 'accepted', 'best_x',
                                                     'converged', 'initial_loglkl', 'initial_position',
                                                     'initial_covmat', 'temperature', 'step_size')}}
    pio.dump_pickle(state, config.io.dir, 'state')
    with open(os.path.join(config.io.dir, 'status.txt'), 'w') as f:
        ran = rec.accepted >= 0
        f.write(f"Status of job '{config.io.jobname}'\nJobtype: {config.run.jobtype}\n")
        f.write(f'Chains: {ran.shape[1]}  iterations run per chain: min {ran.sum(0).min()} max {ran.sum(0).max()}  '
                f'converged: {int(np.sum(rec.converged))}\n')


def analyse(prospect_folder):
    """Re-run the analysis of a finished folder (prospect/run.py:93-117)."""
    config = pio.load_config(prospect_folder)
    config.io.dir = prospect_folder
    state = pio.load_state(prospect_folder)
    if not isinstance(state, dict) or state.get('format') != STATE_FORMAT:
        raise ValueError('state.pkl was not written by prospect_public_b200')
    s = state['record']
    rec = analysis.RunRecord(s['values'], s['reps'], s['names'], s['best_loglkl'], s['accepted'], s['best_x'],
                             s['converged'], s['initial_loglkl'], s['initial_position'], s['initial_covmat'],
                             s['temperature'], s['step_size'])
    if config.run.jobtype == 'profile':
        return analysis.analyse_profile(rec, os.path.join(prospect_folder, 'profile'), config.profile.parameter,
                                        config.profile.confidence_levels)
    if config.run.jobtype == 'global_optimisation':
        return analysis.analyse_global_optimisation(rec, os.path.join(prospect_folder, 'global_optimisation'))
    raise KeyError(f'Jobtype {config.run.jobtype} not understood.')


def gelman_rubin(means, covs, weights):
    """R-1 from per-chain weighted means [C, n], covariances [C, n, n] and total weights [C]: largest
    eigenvalue of W^-1/2 B W^-1/2 (between-chain covariance of the means in units of the mean within-chain
    covariance).  PARITY UNPINNED: the reference delegates this to getdist, which is not installed here
    (SURVEY.md O4)."""
    w = np.asarray(weights, dtype=np.float64)
    w = w / w.sum()
    mean_all = (w[:, None] * means).sum(0)
    within = (w[:, None, None] * covs).sum(0)
    dm = means - mean_all
    between = (w[:, None, None] * (dm[:, :, None] * dm[:, None, :])).sum(0) * len(w) / max(len(w) - 1, 1)
    lw = np.linalg.cholesky(within)
    inv = np.linalg.inv(lw)
    return float(np.linalg.eigvalsh(inv @ between @ inv.T).max())


def run_mcmc(config, max_rounds=64):
    """jobtype mcmc (tasks/mcmc_task.py:10-81): N_chains chains advance `steps_per_iteration` per round on the
    GPU; after each round the chains are unpacked to <dir>/mcmc/<jobname>_<i>.txt (+ .paramnames / .ranges) and
    R-1 decides whether to continue."""
    import torch
    from . import engine
    from .mcmc import initialise_mcmc
    rank, world, _ = dist_info()
    if world > 1:
        raise NotImplementedError('jobtype mcmc runs on one GPU per job in this round.')
    device = _device()
    mc = config.mcmc
    if mc.steps_per_iteration is None:
        raise ValueError("The GPU path needs 'steps_per_iteration'.")
    if config.run.numpy_random_seed is not None:
        np.random.seed(int(config.run.numpy_random_seed))
    out_dir = config.io.dir if config.io.write else None
    kernel = initialise_kernel(config.kernel, out_dir, 0, device)
    sampler = initialise_mcmc(mc, kernel, seed=config.seed)
    n = sampler.problem.n
    rounds, conv = 0, np.inf
    t0 = time.perf_counter()
    while rounds < max_rounds:
        sampler.run_steps(mc.steps_per_iteration)
        rounds += 1
        x, ll, mult, cnt = None, None, None, sampler._samples.count.cpu().numpy()
        means, covs, wsum = [], [], []
        s = sampler._samples
        for c in range(sampler.n_chains):
            sw, swx, swxx = engine.weighted_moments(s.x[c, :cnt[c]], n, s.mult[c, :cnt[c]].contiguous())
            means.append((swx / sw).cpu().numpy())
            covs.append(engine.covariance_from_moments(sw, swx, swxx).cpu().numpy())
            wsum.append(float(sw))
        conv = gelman_rubin(np.array(means), np.array(covs), wsum) if sampler.n_chains > 1 else 0.0
        if config.io.write and mc.unpack_at_dump:
            pio.unpack_mcmc(kernel.param['varying'], os.path.join(config.io.dir, 'mcmc'), config.io.jobname,
                            *sampler.chains)
        if conv < mc.convergence_Rm1:
            break
    torch.cuda.synchronize()
    return {'sampler': sampler, 'rounds': rounds, 'Rm1': conv, 'loop_seconds': time.perf_counter() - t0,
            'chain_steps': rounds * mc.steps_per_iteration * sampler.n_chains}


def run_from_shell(argv=None):
    import argparse
    parser = argparse.ArgumentParser(prog='prospect_public_b200')
    parser.add_argument('input_file', help='input .yaml')
    parser.add_argument('--analyse', action='store_true', help='re-analyse an existing run folder')
    args = parser.parse_args(argv)
    if args.analyse:
        analyse(args.input_file)
    else:
        out = run(args.input_file)
        if dist_info()[0] == 0:
            print(f"{out['chain_steps']} chain-steps in {out['loop_seconds']:.4f} s "
                  f"({out['chain_steps'] / out['loop_seconds']:.4g} chain-steps/s); "
                  f"time to result {out['time_to_result_s']:.3f} s")


if __name__ == '__main__':
    run_from_shell(sys.argv[1:])
