"""Scalar restatement of the annealing loop for HOST-evaluated likelihoods, with explicit variates.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): tests/test_gpu_hosted.py replays the GPU's own Philox variates
(pb200_rng_variates) through this model and compares every proposal, decision, bestfit and schedule value with what
prospect_public_b200.hosted produced through pb200_hosted_propose / pb200_hosted_accept.

Parity status: PINNED.  The step is oracle/reference_port.mh_step (itself pinned against the reference's golden vectors)
with its two numpy draws replaced by the variates handed in; tests/test_hosted_host.py checks it, decision for decision
and bit for bit, against that port and against the unmodified reference's MetropolisHastings.step (numpy's draws
patched to the values the variates stand for) --
    prop = current + step_size * F z,  F F^T = covmat      for  multivariate_normal(current, covmat * step_size**2)
                                                            (prospect/mcmc.py:182-190; equal in distribution),
    accept iff  loglkl_prop - loglkl_cur < T * e, e = -ln u  for  exp((cur - prop) / T) > u  (mcmc.py:170-180)
-- and the iteration bookkeeping IS reference_port.next_settings / reference_port.converged (optimiser.py:87-162,
tasks/optimise_task.py:33-50), called unchanged.
"""
import numpy as np

from . import reference_port as rp


def step(loglkl, x, ll, factor, lo, hi, temperature, step_size, z, e):
    """One MH step (mcmc.py:163-190).  Returns (accepted, prop, loglkl_prop or None, outside)."""
    prop = x + step_size * (factor @ np.asarray(z, dtype=np.float64))
    outside = bool(np.any(prop < lo) or np.any(prop > hi))          # base_kernel.py:89-102
    if outside:                                                     # mcmc.py:166-167: likelihood not evaluated
        return False, prop, None, True
    llp = loglkl(prop)
    return bool(llp - ll < temperature * float(e)), prop, llp, False


def anneal_chain(loglkl, start, covmat, lo, hi, schedule: rp.Schedule, temperature, step_size, current_best,
                 z, e, n_iterations):
    """All iterations of one chain.  z [n_iterations * S, n], e [n_iterations * S]: the chain's variates by absolute
    step.  Returns a list of per-iteration dicts (bestfit + the settings the iteration ran with) and the final state."""
    factor = np.linalg.cholesky(0.5 * (np.asarray(covmat) + np.asarray(covmat).T))      # lower, F F^T = covmat
    s_per = schedule.steps_per_iteration
    x, ll = np.array(start, dtype=np.float64), loglkl(np.array(start, dtype=np.float64))
    settings = {'temperature': temperature, 'temperature_change': schedule.temperature_change(),
                'step_size': step_size, 'step_size_change': schedule.initial_step_size_change(),
                'current_best_loglkl': current_best}
    rows, status = [], 'active'
    for it in range(n_iterations):
        best_ll, best_x, nacc, n_out = ll, x.copy(), 0, 0                   # optimiser.py:55: the chain starts at the start point
        for k in range(s_per):
            t = it * s_per + k
            acc, prop, llp, out = step(loglkl, x, ll, factor, lo, hi, settings['temperature'], settings['step_size'],
                                     z[t], e[t])
            n_out += out
            if acc:
                x, ll, nacc = prop, llp, nacc + 1
                if llp < best_ll:                                   # first minimum (np.argmin, optimiser.py:75)
                    best_ll, best_x = llp, prop.copy()
        bestfit = {'loglkl': best_ll, 'acceptance_rate': np.float64(1 + nacc) / (1 + s_per), 'accepted_steps': 1 + nacc,
                   'position': best_x}
        rows.append({'bestfit': bestfit, 'temperature': settings['temperature'], 'step_size': settings['step_size'],
                     'last_loglkl': ll, 'last_x': x.copy(), 'accepted': nacc, 'outside': n_out})
        if rp.converged(schedule, settings['current_best_loglkl'], bestfit):
            status = 'converged'
            break
        settings = rp.next_settings(schedule, settings, bestfit)
    return rows, {'x': x, 'loglkl': ll, 'status': status, 'temperature': settings['temperature'],
                  'step_size': settings['step_size']}
