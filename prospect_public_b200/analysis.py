"""Profile / global-optimisation analysis and the result files, in PROSPECT's output layout.

Mirrors AnalyseProfileTask (prospect/tasks/analyse_profile_task.py: sort_tasks :91-121, compute_profile
:123-177, write_results :179-196, write_profile_status :62-89, compute_intervals_neyman :198-220,
write_intervals :222-235, get_bestfit_parabola :300-307) and AnalyseGlobalOptimisationTask
(prospect/tasks/analyse_global_optimisation_task.py:39-75).  The files `profile/<p>.txt`,
`profile/<p>_status.txt`, `profile/<p>_intervals.txt` and `global_optimisation/results.txt` are written
character for character as the reference writes them, so `load_profile` and downstream scripts keep working.

Input is the dense record of a run -- arrays over (iteration, chain) -- instead of a list of task objects;
the O(P * R * tasks) Python scans of the reference become array reductions (on the device for the
min-over-iterations / best-repetition part, see pb200_profile_reduce).  Plotting is out of scope.
"""
import os

import numpy as np
from scipy import stats
from scipy.interpolate import interp1d


class RunRecord:
    """Everything the analysis needs, in global chain order c = rep * P + point.

    best_loglkl [K, B], accepted [K, B] (-1 = chain did not run that iteration), best_x [K, B, n] or None,
    converged [B] bool, values [P] (empty for global optimisation), reps R,
    initial_loglkl [P], initial_position [P, n], initial_covmat [P, n, n], names (varying, profile param removed)."""

    def __init__(self, values, reps, names, best_loglkl, accepted, best_x, converged, initial_loglkl,
                 initial_position, initial_covmat, temperature=None, step_size=None):
        self.values = np.asarray(values, dtype=np.float64)
        self.reps, self.names = int(reps), list(names)
        self.best_loglkl, self.accepted, self.best_x = best_loglkl, accepted, best_x
        self.converged = converged
        self.initial_loglkl, self.initial_position, self.initial_covmat = initial_loglkl, initial_position, initial_covmat
        self.temperature, self.step_size = temperature, step_size
        self.n_points = max(1, len(self.values))

    def per_chain_best(self):
        """(best [R, P], iteration [R, P]): first minimum over the iterations that ran; inf / -1 if none."""
        masked = np.where(self.accepted >= 0, self.best_loglkl, np.inf)
        cb = masked.min(axis=0)
        ci = np.where(np.isfinite(cb), masked.argmin(axis=0), -1)
        return cb.reshape(self.reps, self.n_points), ci.reshape(self.reps, self.n_points)


def bestfit_parabola(x, f):
    """Vertex (x*, f*) of the parabola through three points.  The floating-point operations are those of
    analyse_profile_task.py:300-307, term for term and in the same order (Python float arithmetic, `pow` for the
    squares), because x* and f* are printed at full precision into <param>.txt / <param>_intervals.txt: a
    mathematically equal form (e.g. Lagrange) differs in the last bits and breaks byte-identity of those files
    (tests/test_host_logic.py compares randomised triples bit for bit with the pinned port)."""
    xm, xj, xp = (float(v) for v in x)
    fm, fj, fp = (float(v) for v in f)
    x_num = fp*(xj - xm)*(xj + xm) + fj*(xm - xp)*(xm + xp) + fm*(-pow(xj, 2) + pow(xp, 2))
    x_den = 2.*(fp*(xj - xm) + fj*(xm - xp) + fm*(-xj + xp))
    f_num = pow(-(fp*pow(xj - xm, 2)) + fj*pow(xm - xp, 2) + fm*(xj - xp)*(xj - 2*xm + xp), 2)
    f_den = 4.*(xj - xm)*(xj - xp)*(xm - xp)*(fp*(-xj + xm) + fm*(xj - xp) + fj*(-xm + xp))
    return x_num/x_den, fm + f_num/f_den


def sort_bestfits(rec: RunRecord, reduced=None):
    """Per point: best repetition (first minimum), its loglkl / iteration, mean over finished repetitions.
    `reduced` may carry the device reduction (chain_best, chain_iter, point_rep, point_avg as numpy)."""
    if reduced is None:
        cb, ci = rec.per_chain_best()
        best_rep = cb.argmin(axis=0)
        fin = np.isfinite(cb)
        with np.errstate(invalid='ignore'):
            avg = np.array([cb[fin[:, p], p].mean() if fin[:, p].any() else np.inf for p in range(rec.n_points)])
    else:
        cb = reduced['chain_best'].reshape(rec.reps, rec.n_points)
        ci = reduced['chain_iter'].reshape(rec.reps, rec.n_points)
        best_rep, avg = reduced['point_rep'], reduced['point_avg']
    return {'chain_best': cb, 'chain_iter': ci, 'best_rep': np.asarray(best_rep), 'avg_loglkl': np.asarray(avg)}


def compute_profile(rec: RunRecord, srt):
    """Delta chi^2 against the parabola minimum through the best point and its neighbours."""
    p_idx = np.arange(rec.n_points)
    best = srt['chain_best'][srt['best_rep'], p_idx]
    loglkls = np.where(np.isfinite(best), best, rec.initial_loglkl)
    order = np.argsort(rec.values)
    pv, ll = rec.values[order], loglkls[order]
    bi = int(np.argmin(ll))
    if bi == 0 or bi == len(ll) - 1:
        gx, gl = pv[bi], ll[bi]
    else:
        gx, gl = bestfit_parabola(pv[bi - 1:bi + 2], ll[bi - 1:bi + 2])
    return {'global_best_loglkl': gl, 'global_best_param_val': gx, 'param_vals': pv, 'loglkls': ll,
            'Delta_chi2': 2 * (ll - gl), 'Delta_chi2_by_point': 2 * (loglkls - gl),
            'Delta_chi2_chains': 2 * (srt['chain_best'] - gl)}


def _best_positions(rec, srt):
    """[P, n] position of the best repetition's best iteration (from the per-chain best positions when the record
    carries them -- rec.best_positions [B, n] -- else from the dense per-iteration positions)."""
    p_idx = np.arange(rec.n_points)
    rep = srt['best_rep']
    chain = rep * rec.n_points + p_idx
    per_chain = getattr(rec, 'best_positions', None)
    if per_chain is not None:
        return per_chain[chain]
    it = srt['chain_iter'][rep, p_idx]
    return rec.best_x[np.maximum(it, 0), chain]


def write_profile_results(rec: RunRecord, srt, profile, out_dir, parameter):
    path = os.path.join(out_dir, f'{parameter}.txt')
    p_idx = np.arange(rec.n_points)
    best = srt['chain_best'][srt['best_rep'], p_idx]
    pos = _best_positions(rec, srt)
    with open(path, 'w') as f:
        f.write(f'{parameter} \t Delta_chi2 \t -loglkl \t avg_best_loglkl \t initial_loglkl \t ')
        f.write(' \t '.join(str(nm) for nm in rec.names) + '\n')
        for p in np.argsort(rec.values, kind='stable'):
            if best[p] == np.inf:
                f.write(f'{rec.values[p]:.10f} \t --- no optimisations finished ---\n')
                continue
            f.write(f"{rec.values[p]:.10f} \t {profile['Delta_chi2_by_point'][p]:.10e} \t {best[p]:.10e} \t ")
            f.write(f"{srt['avg_loglkl'][p]:.10e} \t {rec.initial_loglkl[p]:.10e} \t ")
            f.write(' \t '.join(str(np.round(v, 10)) for v in pos[p]) + '\n')
    return path


def write_profile_status(rec: RunRecord, srt, out_dir, parameter):
    path = os.path.join(out_dir, f'{parameter}_status.txt')
    ran = rec.accepted >= 0
    k_rows = ran.shape[0]
    any_ran = ran.any(axis=0)                                             # per chain
    last = k_rows - 1 - np.argmax(ran[::-1], axis=0)                      # last iteration that ran
    cur_ll = rec.best_loglkl[last, np.arange(ran.shape[1])]
    chain_iter, chain_best, conv = srt['chain_iter'], srt['chain_best'], np.asarray(rec.converged, dtype=bool)
    out = [f'{parameter:14.10} ',
           '\t\t | rep. no.  current iter \t current loglkl \t best iter \t best loglkl \t converged? ']
    for p in np.argsort(rec.values, kind='stable'):
        out.append(f'\n{rec.values[p]:14.10} ')
        for rep in range(rec.reps):
            c = rep * rec.n_points + p
            if not any_ran[c]:
                continue
            line = (f"| rep {rep + 1} \t {int(last[c]):<13} \t {cur_ll[c]:.10e} \t "
                    f"{chain_iter[rep, p]:<9} \t {chain_best[rep, p]:.10e} \t {bool(conv[c])} ")
            out.append(f'\t {line}' if rep == 0 else f'\n\t\t\t\t {line}')
    with open(path, 'w') as f:
        f.write(''.join(out))
    return path


def compute_intervals_neyman(profile, confidence_levels):
    pv, dchi = profile['param_vals'], profile['Delta_chi2']
    cont = np.linspace(np.min(pv), np.max(pv), 10000)
    interp = interp1d(pv, dchi, fill_value='extrapolate', kind='linear')(cont)
    idx_min = int(np.argmin(np.abs(cont - profile['global_best_param_val'])))
    intervals = {}
    for cl in confidence_levels:
        chi2 = stats.chi2.ppf(cl, 1)
        intervals[cl] = [None, None]
        if idx_min > 0:
            intervals[cl][0] = cont[int(np.argmin(np.abs(interp[0:idx_min] - chi2)))]
        if idx_min < len(cont) - 1:
            intervals[cl][1] = cont[idx_min + int(np.argmin(np.abs(interp[idx_min:-1] - chi2)))]
    return intervals


def write_intervals(profile, intervals, out_dir, parameter):
    path = os.path.join(out_dir, f'{parameter}_intervals.txt')
    best = profile['global_best_param_val']
    with open(path, 'w') as f:
        f.write('C.L. \t C.I. \t bestfit (+dist. to upper bound)(-dist. to lower bound)\n')
        for cl, interval in intervals.items():
            f.write(f'{cl*100} %: \t {interval} \t {best}')
            upper = interval[1] - best if interval[1] is not None else None
            lower = best - interval[0] if interval[0] is not None else None
            f.write(f'(+{upper})(-{lower}) \n')
    return path


def analyse_profile(rec: RunRecord, out_dir, parameter, confidence_levels, reduced=None, stamp=None):
    """The whole AnalyseProfileTask.run minus the plots.  Returns (sorted bestfits, profile, intervals).  The profile
    and its intervals are written first (`stamp['t_results_written']` is taken when profile/<p>.txt is on disk), the
    per-repetition status table -- one formatted line per chain -- after them."""
    import time
    os.makedirs(out_dir, exist_ok=True)
    srt = sort_bestfits(rec, reduced)
    profile = compute_profile(rec, srt)
    write_profile_results(rec, srt, profile, out_dir, parameter)
    if stamp is not None:
        stamp['t_results_written'] = time.perf_counter()
    intervals = compute_intervals_neyman(profile, confidence_levels)
    write_intervals(profile, intervals, out_dir, parameter)
    write_profile_status(rec, srt, out_dir, parameter)
    return srt, profile, intervals


def analyse_global_optimisation(rec: RunRecord, out_dir):
    """global_optimisation/results.txt (analyse_global_optimisation_task.py:59-75): one row per repetition."""
    os.makedirs(out_dir, exist_ok=True)
    cb, ci = rec.per_chain_best()
    path = os.path.join(out_dir, 'results.txt')
    with open(path, 'w') as f:
        f.write('Repetition number \t -loglkl \t initial_loglkl \t ')
        f.write(' \t '.join(str(nm) for nm in rec.names) + '\n')
        for rep in range(rec.reps):
            if cb[rep, 0] == np.inf:
                f.write(f'{rep} \t --- no optimisations finished ---\n')
                continue
            f.write(f'{rep} \t {cb[rep, 0]:.10e} \t {rec.initial_loglkl[0]:.10e} \t ')
            per_chain = getattr(rec, 'best_positions', None)
            pos = per_chain[rep] if per_chain is not None else rec.best_x[ci[rep, 0], rep]
            f.write(' \t '.join(str(np.round(v, 10)) for v in pos) + '\n')
    return {'chain_best': cb, 'chain_iter': ci, 'best_rep': int(np.argmin(cb[:, 0]))}
