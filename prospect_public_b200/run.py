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
from .optimiser import initialise_optimiser as build_optimiser

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


def run(user_inp, quiet=True, stop_after=None, max_rounds=64, kernel=None):
    """Run a job (a .yaml file or dict) or resume one (the folder of an earlier run, prospect/io.py:15-24).  Returns
    a dict with the records, the analysis products and timings.  `stop_after` ends an annealing job after that
    many iterations (an mcmc job after that many rounds) with a complete snapshot, as an interrupted run would leave
    it; `max_rounds` bounds the rounds of an mcmc job that does not reach convergence_Rm1.  `kernel` (annealing jobs):
    a kernel object, or a callable (config_kernel, output_folder, rank, device) -> kernel, used instead of the yaml's
    `kernel.type` factory -- e.g. a hosted.HostKernelAdapter around a likelihood that stays on the host."""
    t_start = time.perf_counter()
    config_yaml, is_folder = pio.read_user_input(user_inp) if isinstance(user_inp, str) else (user_inp, False)
    rank, world, _ = init_from_env('nccl') if int(os.environ.get('WORLD_SIZE', '1')) > 1 else dist_info()
    config = Configuration(config_yaml).synchronise()      # rank 0's output folder and Philox key for every rank
    if is_folder:
        out = resume(config, user_inp, max_rounds=max_rounds, kernel=kernel)
        out['time_to_result_s'] = time.perf_counter() - t_start
        out['config'] = config
        return out
    if rank == 0:
        pio.prepare_run(config)
    _barrier()
    if config.run.jobtype == 'mcmc':
        out = run_mcmc(config, max_rounds=max_rounds, stop_after_rounds=stop_after, kernel=kernel)
    else:
        out = run_annealing(config, stop_after=stop_after, kernel=kernel)
    out['time_to_result_s'] = time.perf_counter() - t_start          # incl. the snapshot pickles
    if 't_results_written' in out:
        out['time_to_profile_s'] = out['t_results_written'] - t_start  # yaml read -> result files written
    out['config'] = config
    return out


def run_annealing(config, keep_positions=True, starts=None, values=None, seed=None, write=True, stop_after=None,
                  resume_from=None, kernel=None):
    """jobtype profile / global_optimisation.  `starts` / `values` / `seed` let reanneal() supply the start points;
    `resume_from` = (state dict, RunRecord) of an interrupted run continues its chains where they stopped."""
    import torch
    rank, world, _ = dist_info()
    device = _device()
    prof = config.profile
    if config.run.numpy_random_seed is not None:
        np.random.seed(int(config.run.numpy_random_seed))     # only used if a random kernel must be generated
    out_dir = config.io.dir if config.io.write else None
    make_kernel = kernel if callable(kernel) else (initialise_kernel if kernel is None else (lambda *a: kernel))
    if rank == 0:
        kernel = make_kernel(config.kernel, out_dir, 0, device)   # may create analytical/random_gauss.pkl
    _barrier()
    if rank != 0:
        kernel = make_kernel(config.kernel, out_dir, rank, device)
    if starts is None:
        starts, values = initialise_optimiser(config, kernel, shard=(rank, world))
    elif config.run.jobtype == 'profile':
        kernel.set_fixed_parameters({prof.parameter: float(values[0])})
    n_iter = int(prof.max_iterations)
    reps = int(prof.repetitions)
    done, old_rec = 0, None
    if resume_from is not None:
        done, old_rec = int(resume_from[0]['iterations_done']), resume_from[1]
    to_run = max(0, n_iter - done) if stop_after is None else max(0, min(n_iter - done, int(stop_after)))
    settings = {
        'current_best_loglkl': starts.initial_loglkl,
        'initial_position': starts.positions,
        'covmat': starts.covmats,
        'temperature': prof.temperature_range[0],
        'temperature_change': get_temperature_change(config),
        'step_size': prof.step_size_range[0],
        'step_size_change': get_initial_step_size_change(config),
        'iteration_number': done,
        'repetitions': reps,
        'fixed_param_val': values if config.run.jobtype == 'profile' else None,
        'max_rows': n_iter,
        'seed': config.seed if seed is None else seed,
    }
    if resume_from is not None:
        settings['chain_state'] = resume_from[0]['chain_state']
    sa = build_optimiser(config, kernel, settings, keep_positions=keep_positions)    # fused, or hosted for host kernels
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if to_run > 0:
        sa.run_schedule(to_run)
    torch.cuda.synchronize()
    loop_s = time.perf_counter() - t0
    result = {'loop_seconds': loop_s, 'optimiser': sa, 'values': values, 'seed': sa.seed,
              'iterations_done': sa.iterations_done}

    # --- the result files first: they need the per-iteration scalars (already exchanged at the iteration boundaries)
    # and ONE position per chain (its best iteration's), not every iteration's positions
    scal = sa.global_scalars()
    best_ll, acc, step, temperature = scal['best_loglkl'], scal['accepted'], scal['step_size'], scal['temperature']
    chain_steps = int((acc >= 0).sum()) * int(prof.steps_per_iteration)    # global; only what THIS call ran
    if old_rec is not None:                     # rows of the iterations that ran before the interruption
        k_old = min(done, old_rec.best_loglkl.shape[0], n_iter)
        best_ll[:k_old], acc[:k_old] = old_rec.best_loglkl[:k_old], old_rec.accepted[:k_old]
        step[:k_old], temperature[:k_old] = old_rec.step_size[:k_old], old_rec.temperature[:k_old]
    status = sa.global_status()
    rec = analysis.RunRecord(values, reps, starts.names, best_ll, acc, None, status == capi.CONVERGED,
                             starts.initial_loglkl, starts.positions, starts.covmats, temperature, step)
    fast = old_rec is None and keep_positions          # (a resumed run needs the old rows' positions: dense path)
    if fast:
        rec.best_positions = sa.global_best_positions()
        if rank == 0 and write:
            result.update(write_result_files(config, rec, sa if world == 1 else None))
    # --- then everything a snapshot / the caller may want: every iteration's positions, the per-chain state
    best_x = sa.global_positions() if keep_positions else None
    if old_rec is not None and best_x is not None and old_rec.best_x is not None:
        best_x[:k_old] = old_rec.best_x[:k_old]
    rec.best_x = best_x
    result.update({'record': rec, 'chain_steps': chain_steps, 'status': status})
    chain_state = sa.global_chain_state()
    if rank == 0 and write:
        if not fast:
            result.update(write_result_files(config, rec, None))
        if config.io.write:
            dump_snapshot(config, rec, resume_state={'chain_state': chain_state, 'iterations_done': sa.iterations_done,
                                                     'seed': sa.seed})
    return result


def resume(config, prospect_folder, max_rounds=64, kernel=None):
    """`prospect <folder>` (prospect/io.py:15-24, prospect/communication.py:30-36): continue the chains of an
    interrupted run: an mcmc job from its chain files (run_mcmc), an annealing run from the state in its snapshot -- position, temperature, step size and step-size
    history, Philox step counter -- until every chain has run `profile.max_iterations` iterations (read from the
    folder's log.yaml, so raising it there extends a finished run).  The continued run is step for step the run
    that was never interrupted: same random streams, same records, same output files."""
    if config.run.jobtype == 'mcmc':
        config.io.dir = config.config_dict['io']['dir'] = prospect_folder
        return run_mcmc(config, max_rounds=max_rounds, resume_folder=prospect_folder, kernel=kernel)
    config.io.dir = config.config_dict['io']['dir'] = prospect_folder
    _, rec_old = load_record(prospect_folder)
    state = pio.load_pickle(prospect_folder, 'pb200_state')
    if 'chain_state' not in state:
        raise ValueError(f'{prospect_folder}/pb200_state.pkl holds no chain state to resume from.')
    from .initialise import StartingPoints
    starts = StartingPoints(rec_old.names, rec_old.initial_position, rec_old.initial_covmat, rec_old.initial_loglkl)
    return run_annealing(config, starts=starts, values=rec_old.values, seed=int(state['seed']),
                         resume_from=(state, rec_old), kernel=kernel)


def write_result_files(config, rec, sa=None):
    """Analysis + result files of a finished (or re-annealed) run: profile/<p>.txt, <p>_intervals.txt, <p>_status.txt
    or global_optimisation/results.txt.  Returns the analysis products and `t_results_written`."""
    prof = config.profile
    result = {}
    if config.run.jobtype == 'profile':
        sub = os.path.join(config.io.dir, 'profile') if config.io.write else None
        if sub is not None:
            reduced = None
            if sa is not None:
                reduced = {k: v.cpu().numpy() for k, v in sa.reduce_over_iterations().items()}
            srt, profile, intervals = analysis.analyse_profile(rec, sub, prof.parameter, prof.confidence_levels,
                                                               reduced, stamp=result)
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
    result.setdefault('t_results_written', time.perf_counter())      # profile/<p>.txt (or results.txt) is on disk here
    return result


def write_outputs(config, rec, sa=None, resume_state=None):
    """Result files + snapshot (re-annealing, tests)."""
    result = write_result_files(config, rec, sa)
    if config.io.write:
        dump_snapshot(config, rec, resume_state=resume_state)
    return result


def dump_snapshot(config, rec, reference_compatible=True, resume_state=None):
    """Snapshot of a finished run:
      pb200_config.pkl / pb200_state.pkl   this package's own dense records (a few MB);
      config.pkl / state.pkl               PROSPECT's own pickle layout (compat.export_reference_snapshot), so that
                                           `prospect-analyse`, `prospect.run.load`, `load_profile` and `reanneal` of the
                                           reference work on the folder unchanged;
      status.txt                           one-paragraph summary."""
    from . import compat
    pio.dump_pickle(config, config.io.dir, 'pb200_config')
    state = {'format': STATE_FORMAT, 'jobtype': config.run.jobtype, 'seed': config.seed, 'epoch': 0,    # (cached draw)
             'record': {k: getattr(rec, k) for k in ('values', 'reps', 'names', 'best_loglkl', 'accepted', 'best_x',
                                                     'converged', 'initial_loglkl', 'initial_position',
                                                     'initial_covmat', 'temperature', 'step_size')}}
    if resume_state is not None:            # per-chain state: `prospect_public_b200 <folder>` continues from here
        state.update(chain_state=resume_state['chain_state'], iterations_done=int(resume_state['iterations_done']),
                     seed=int(resume_state['seed']))
    pio.dump_pickle(state, config.io.dir, 'pb200_state')
    if reference_compatible:
        n_tasks = int((rec.accepted >= 0).sum())
        if n_tasks > 20000 and rec.temperature is not None and rec.step_size is not None and rec.best_x is not None:
            compat.export_reference_snapshot_bulk(config, rec, config.io.dir)      # bytes assembled with numpy
        else:
            compat.export_reference_snapshot(config, rec, config.io.dir, compact=n_tasks > 20000)
    with open(os.path.join(config.io.dir, 'status.txt'), 'w') as f:
        ran = rec.accepted >= 0
        f.write(f"Status of job '{config.io.jobname}'\nJobtype: {config.run.jobtype}\n")
        f.write(f'Chains: {ran.shape[1]}  iterations run per chain: min {ran.sum(0).min()} max {ran.sum(0).max()}  '
                f'converged: {int(np.sum(rec.converged))}\n')


def load_record(prospect_folder):
    """(config, RunRecord) of a folder written by this package."""
    config = pio.load_pickle(prospect_folder, 'pb200_config')
    config.io.dir = prospect_folder
    state = pio.load_pickle(prospect_folder, 'pb200_state')
    if not isinstance(state, dict) or state.get('format') != STATE_FORMAT:
        raise ValueError('pb200_state.pkl was not written by prospect_public_b200')
    s = state['record']
    rec = analysis.RunRecord(s['values'], s['reps'], s['names'], s['best_loglkl'], s['accepted'], s['best_x'],
                             s['converged'], s['initial_loglkl'], s['initial_position'], s['initial_covmat'],
                             s['temperature'], s['step_size'])
    return config, rec


def analyse(prospect_folder):
    """Re-run the analysis of a finished folder (prospect/run.py:93-117)."""
    config, rec = load_record(prospect_folder)
    if config.run.jobtype == 'profile':
        return analysis.analyse_profile(rec, os.path.join(prospect_folder, 'profile'), config.profile.parameter,
                                        config.profile.confidence_levels)
    if config.run.jobtype == 'global_optimisation':
        return analysis.analyse_global_optimisation(rec, os.path.join(prospect_folder, 'global_optimisation'))
    raise KeyError(f'Jobtype {config.run.jobtype} not understood.')


def merge_records(old, new, k_offset):
    """Stack a re-annealing's records under the old ones.  Chains are matched by (value, repetition); values or
    repetitions present in only one of the two runs get rows marked `accepted = -1` in the other part."""
    values = np.array(sorted(set(old.values.tolist()) | set(new.values.tolist())))
    if len(old.values) == 0:
        values = np.zeros(0)
    n_pts, reps = max(1, len(values)), max(old.reps, new.reps)
    k_tot, n = k_offset + new.best_loglkl.shape[0], old.best_x.shape[2]
    best = np.full((k_tot, n_pts * reps), np.inf)
    acc = np.full((k_tot, n_pts * reps), -1, dtype=np.int32)
    bx = np.zeros((k_tot, n_pts * reps, n))
    temp, step = np.zeros((k_tot, n_pts * reps)), np.zeros((k_tot, n_pts * reps))
    conv = np.zeros(n_pts * reps, dtype=bool)
    init_ll, init_pos = np.zeros(n_pts), np.zeros((n_pts, n))
    init_cov = np.zeros((n_pts, n, n))
    for part, k0 in ((old, 0), (new, k_offset)):
        kk = part.best_loglkl.shape[0]
        for p_src in range(part.n_points):
            p_dst = int(np.nonzero(values == part.values[p_src])[0][0]) if len(values) else 0
            if part is old or not np.any(old.values == part.values[p_src] if len(values) else True):
                init_ll[p_dst], init_pos[p_dst] = part.initial_loglkl[p_src], part.initial_position[p_src]
                init_cov[p_dst] = part.initial_covmat[p_src]
            for rep in range(part.reps):
                src, dst = rep * part.n_points + p_src, rep * n_pts + p_dst
                best[k0:k0 + kk, dst] = part.best_loglkl[:, src]
                acc[k0:k0 + kk, dst] = part.accepted[:, src]
                bx[k0:k0 + kk, dst] = part.best_x[:, src]
                if part.temperature is not None:
                    temp[k0:k0 + kk, dst] = part.temperature[:, src]
                if part.step_size is not None:
                    step[k0:k0 + kk, dst] = part.step_size[:, src]
                conv[dst] = part.converged[src] if part is new else conv[dst]
    return analysis.RunRecord(values, reps, old.names, best, acc, bx, conv, init_ll, init_pos, init_cov, temp, step)


def reanneal(yaml_schedule, prospect_folder):
    """Restart every profile value from its best point so far with a new schedule (prospect/profile.py:17-97):
    `yaml_schedule` holds {'profile': {...overrides...}} (temperature_range, max_iterations, step-size settings,
    steps_per_iteration, repetitions, values).  Values without an old bestfit are initialised from the starting
    MCMC as in a fresh run.  Unlike the reference, which only queues the tasks for a later `prospect <folder>`,
    the new iterations run at once; their records are appended under the old ones and every output file is
    rewritten.  The new chains draw from fresh Philox streams (the seed is advanced per re-annealing)."""
    import yaml as _yaml
    new_schedule = _yaml.full_load(open(yaml_schedule, 'r')) if isinstance(yaml_schedule, str) else yaml_schedule
    if 'profile' not in new_schedule:
        raise KeyError("Reanneal yaml file must contain 'profile' as the first and only key.")
    config_old, rec_old = load_record(prospect_folder)
    cfg = {sec: dict(entries) for sec, entries in config_old.config_dict.items()}
    cfg['profile'].update(new_schedule['profile'])
    cfg['io'] = dict(cfg['io'], dir=prospect_folder, overwrite_dir=True)
    config = Configuration(cfg)
    config.io.dir = config.config_dict['io']['dir'] = prospect_folder
    is_profile = config.run.jobtype == 'profile'
    device = _device()
    kernel = initialise_kernel(config.kernel, prospect_folder, 0, device)
    srt = analysis.sort_bestfits(rec_old)
    p_idx = np.arange(rec_old.n_points)
    old_best = srt['chain_best'][srt['best_rep'], p_idx]
    old_pos = analysis._best_positions(rec_old, srt)
    values = np.asarray(config.profile.values, dtype=np.float64) if is_profile else np.zeros(0)
    n_pts = max(1, len(values))
    n = rec_old.best_x.shape[2]
    positions, covmats, init_ll = np.zeros((n_pts, n)), np.zeros((n_pts, n, n)), np.zeros(n_pts)
    fresh = []
    for p in range(n_pts):
        hit = np.nonzero(rec_old.values == values[p])[0] if is_profile else np.array([0])
        if len(hit) and np.isfinite(old_best[hit[0]]):
            positions[p], covmats[p], init_ll[p] = old_pos[hit[0]], rec_old.initial_covmat[hit[0]], old_best[hit[0]]
        else:
            fresh.append(p)
    if fresh:
        sub = Configuration(dict(cfg, profile=dict(cfg['profile'], values=values[fresh])))
        sp, _ = initialise_optimiser(sub, initialise_kernel(config.kernel, prospect_folder, 0, device))
        positions[fresh], covmats[fresh], init_ll[fresh] = sp.positions, sp.covmats, sp.initial_loglkl
    from .initialise import StartingPoints
    starts = StartingPoints(rec_old.names, positions, covmats, init_ll)
    state = pio.load_pickle(prospect_folder, 'pb200_state')
    epoch = int(state.get('epoch', 0)) + 1
    seed = (int(state.get('seed', config.seed)) + 0x9E3779B97F4A7C15 * epoch) & 0xFFFFFFFFFFFFFFFF
    out = run_annealing(config, starts=starts, values=values, seed=seed, write=False)
    rec = merge_records(rec_old, out['record'], rec_old.best_loglkl.shape[0])
    out['record'] = rec
    if dist_info()[0] == 0:
        out.update(write_outputs(config, rec))
        st = pio.load_pickle(prospect_folder, 'pb200_state')
        st.update(epoch=epoch, seed=int(state.get('seed', config.seed)))
        pio.dump_pickle(st, prospect_folder, 'pb200_state')
    return out


def gelman_rubin(means, covs, weights):
    """R-1 from per-chain weighted means [C, n], covariances [C, n, n] and total weights [C], following the
    definition of getdist's MCSamples.getGelmanRubin (what prospect/analysis.py:27-32 calls): chains weighted by
    their total weight, between-chain covariance of the means x C / (C - 1), largest eigenvalue of
    L^-1 B L^-T with L the Cholesky factor of the weight-averaged within-chain covariance (getdist rescales both
    matrices by the diagonal of the latter first, which leaves the eigenvalues unchanged).  PARITY UNPINNED: getdist
    is not installed here and the reference ships no value of this statistic;
    tests/test_host_logic.py pins the implementation to a hand-computed case and an independent transcription of
    the definition."""
    w = np.asarray(weights, dtype=np.float64)
    w = w / w.sum()
    mean_all = (w[:, None] * means).sum(0)
    within = (w[:, None, None] * covs).sum(0)
    dm = means - mean_all
    between = (w[:, None, None] * (dm[:, :, None] * dm[:, None, :])).sum(0) * len(w) / max(len(w) - 1, 1)
    lw = np.linalg.cholesky(within)
    inv = np.linalg.inv(lw)
    return float(np.linalg.eigvalsh(inv @ between @ inv.T).max())


def convergence_from_statistics(stats, n):
    """(R-1, pooled mean [n], pooled covariance [n, n]) from the flat statistics vector of engine.chain_statistics
    (summed over ranks).  Same quantity as gelman_rubin(): with sw_c, m_c the weight and mean of chain c,
    within = sum_c (swxx_c - sw_c m_c m_c^T) / sum_c sw_c and between = (sum_c sw_c m_c m_c^T / sum_c sw_c - m m^T)
    * C / (C - 1).  PARITY UNPINNED like gelman_rubin (getdist absent)."""
    st = np.asarray(stats, dtype=np.float64)
    sw, swx = st[0], st[1:1 + n]
    swxx = st[1 + n:1 + n + n * n].reshape(n, n)
    syy = st[1 + n + n * n:1 + n + 2 * n * n].reshape(n, n)
    n_chains = int(round(st[-1]))
    mean = swx / sw
    cov = swxx / sw - np.outer(mean, mean)
    if n_chains < 2:
        return 0.0, mean, cov
    within = (swxx - syy) / sw
    between = (syy / sw - np.outer(mean, mean)) * n_chains / (n_chains - 1)
    inv = np.linalg.inv(np.linalg.cholesky(within))
    return float(np.linalg.eigvalsh(inv @ between @ inv.T).max()), mean, cov


def shard_chains(n_chains, rank, world):
    """Contiguous block [lo, hi) of the global chain indices owned by `rank` (blocks differ by at most one chain)."""
    return n_chains * rank // world, n_chains * (rank + 1) // world


def run_mcmc(config, max_rounds=64, resume_folder=None, stop_after_rounds=None, kernel=None):
    """jobtype mcmc (tasks/mcmc_task.py:10-81): N_chains chains advance `steps_per_iteration` per round on the
    GPU(s); after each round the chains are unpacked to <dir>/mcmc/<jobname>_<i>.txt (+ .paramnames / .ranges) and
    R-1 decides whether to continue.

    Several GPUs (one process each): rank r owns the contiguous chain block shard_chains(); Philox streams are keyed
    by the GLOBAL chain index, so every chain is the same whatever the number of GPUs.  The only exchange is one
    all-reduce (sum, FP64) per round of the 2 + n + 2 n^2 sufficient statistics of the pooled covariance and R-1
    (SURVEY.md section 8e); every rank writes its own chain files.

    `resume_folder` (`prospect <folder>`, prospect/io.py:15-24 + tasks/mcmc_task.py:71-81): continue the chains of an
    earlier run from its chain files -- last position (written with 18 digits: exact), step counter (= sum of the
    multiplicities - 1) and the Philox key kept in pb200_mcmc_state.pkl -- until R-1 < convergence_Rm1.  The
    continued run is step for step the run that was never interrupted.  `stop_after_rounds` ends a run early, as an
    interruption would (tests)."""
    import torch
    from . import engine
    from .mcmc import initialise_mcmc
    rank, world, _ = dist_info()
    device = _device()
    mc = config.mcmc
    if mc.steps_per_iteration is None:
        raise ValueError("The GPU path needs 'steps_per_iteration'.")
    if config.run.numpy_random_seed is not None:
        np.random.seed(int(config.run.numpy_random_seed))
    out_dir = config.io.dir if config.io.write else None
    make_kernel = kernel if callable(kernel) else (initialise_kernel if kernel is None else (lambda *a: kernel))
    if rank == 0:
        kernel = make_kernel(config.kernel, out_dir, 0, device)   # may create analytical/random_gauss.pkl
    _barrier()
    if rank != 0:
        kernel = make_kernel(config.kernel, out_dir, rank, device)
    lo, hi = shard_chains(int(mc.N_chains), rank, world)
    n = len(kernel.varying_param_names)
    seed, rounds, extra = config.seed, 0, {}
    if resume_folder is not None:
        from .initialise import load_mcmc
        state = pio.load_pickle(resume_folder, 'pb200_mcmc_state')
        seed, rounds = int(state['seed']), int(state['rounds'])
        if int(state['n_chains']) != int(mc.N_chains) or int(state['steps_per_iteration']) != int(mc.steps_per_iteration):
            raise ValueError('N_chains / steps_per_iteration of the resumed job differ from the interrupted one.')
        if hi > lo:
            prior = load_mcmc(os.path.join(resume_folder, 'mcmc', config.io.jobname), collapse=False,
                              indices=range(lo, hi))
            done = {int(round(np.sum(ch.mults))) - 1 for ch in prior}
            if done != {rounds * int(mc.steps_per_iteration)}:
                raise ValueError(f'chain files hold {sorted(done)} steps, the snapshot says '
                                 f'{rounds * int(mc.steps_per_iteration)}')
            extra = dict(chains=prior, preload=True, steps_done=rounds * int(mc.steps_per_iteration))
    sampler = initialise_mcmc(mc, kernel, seed=seed, n_chains=hi - lo,
                              chain_ids=np.arange(lo, hi, dtype=np.uint32), **extra) if hi > lo else None
    conv, mean, cov = np.inf, None, None
    if world > 1:
        import torch.distributed as dist
        _barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rounds_here = 0
    while rounds < max_rounds:
        if sampler is not None:
            sampler.run_steps(mc.steps_per_iteration)
            stats = engine.chain_statistics(sampler._samples, n)
        else:
            stats = torch.zeros(2 + n + 2 * n * n, dtype=torch.float64, device=device)
        rounds += 1
        rounds_here += 1
        if world > 1:
            if dist.get_backend() != 'nccl':
                stats = stats.cpu()
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        conv, mean, cov = convergence_from_statistics(stats.cpu().numpy(), n)
        if config.io.write and mc.unpack_at_dump and sampler is not None:
            pio.unpack_mcmc(kernel.param['varying'], os.path.join(config.io.dir, 'mcmc'), config.io.jobname,
                            *sampler.chains, first_index=lo, write_names=rank == 0)
        if config.io.write and rank == 0:
            pio.dump_pickle({'format': STATE_FORMAT, 'jobtype': 'mcmc', 'seed': int(seed), 'rounds': rounds,
                             'n_chains': int(mc.N_chains), 'steps_per_iteration': int(mc.steps_per_iteration),
                             'Rm1': float(conv)}, config.io.dir, 'pb200_mcmc_state')
            pio.dump_pickle(config, config.io.dir, 'pb200_config')
        if conv < mc.convergence_Rm1 or (stop_after_rounds is not None and rounds_here >= stop_after_rounds):
            break
    torch.cuda.synchronize()
    _barrier()
    return {'sampler': sampler, 'rounds': rounds, 'Rm1': conv, 'mean': mean, 'covmat': cov,
            'loop_seconds': time.perf_counter() - t0, 'chain_range': (lo, hi),
            'chain_steps': rounds_here * mc.steps_per_iteration * int(mc.N_chains)}


def run_from_shell(argv=None):
    import argparse
    parser = argparse.ArgumentParser(prog='prospect_public_b200')
    parser.add_argument('input_file', help='input .yaml, or the folder of an earlier run to resume it')
    parser.add_argument('--analyse', action='store_true', help='re-analyse an existing run folder')
    parser.add_argument('--reanneal', metavar='YAML', help='re-anneal the run folder with the schedule in YAML')
    args = parser.parse_args(argv)
    if args.analyse:
        analyse(args.input_file)
    elif args.reanneal:
        reanneal(args.reanneal, args.input_file)
    else:
        out = run(args.input_file)
        if dist_info()[0] == 0:
            print(f"{out['chain_steps']} chain-steps in {out['loop_seconds']:.4f} s "
                  f"({out['chain_steps'] / out['loop_seconds']:.4g} chain-steps/s); "
                  f"time to result {out['time_to_result_s']:.3f} s")


if __name__ == '__main__':
    run_from_shell(sys.argv[1:])
