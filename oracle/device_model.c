/*
 * device_model.c -- scalar C restatement of the batched sampler that libpb200 runs on the GPU.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/reference_port.py).  It is written independently of the
 * CUDA sources, one chain at a time, lane by lane, with the floating-point operations in the
 * order the device performs them (explicit fmaf/fma, no contraction: build with
 * -ffp-contract=off), so that -- given the same normal / exponential variates -- chain
 * trajectories, accept counts and bestfits can be compared bit for bit.  In stand-alone mode it
 * draws its own variates from Philox-4x32-10 with a float64 Box-Muller.
 *
 * What it restates (reference semantics, paths relative to the reference root):
 *   prospect/mcmc.py:163-190       MH step: n normals, one uniform, accept iff exp(-D/T) > u
 *   prospect/optimiser.py:40-85    start point counts, best = first minimum, AR = (1+acc)/(1+S)
 *   prospect/optimiser.py:87-162   next temperature / step size
 *   prospect/tasks/optimise_task.py:33-54  chi2_tolerance stop
 *   prospect/kernels/analytical_kernel.py:59-66 + base_kernel.py:89-102  Gaussian -log L, prior box
 * Parity of those semantics with the reference itself is pinned through oracle/reference_port.py
 * (tests/test_oracle_golden.py) and compared with this model in tests/test_device_model.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ACCEPT_SLOT 0xffffffffu

typedef struct {
    int32_t d, n, npad, n_points;
    const double *saa, *sab, *mu, *lo, *hi, *f, *c0;
    const float *lt_t, *lam;
    const double *g_t, *h;
} dm_problem;

typedef struct {
    int32_t steps_per_iteration, step_mode;
    double temperature_change, step_size_change, adaptive_lo, adaptive_hi, adaptive_multiplier, chi2_tolerance;
} dm_schedule;

typedef struct {          /* one chain */
    double *x;            /* [npad] */
    double loglkl, temperature, step_size, current_best;
    double prev_mult, cur_mult, prev_ar;
    int32_t point, iteration, status, stage;
    uint32_t chain_id, steps_done;
} dm_chain;

void dm_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* the device forms u in float: u = fmaf((float)w, 2^-32, 2^-33) */
static double u01(uint32_t w) { return (double)fmaf((float)w, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

static void box_muller(uint32_t w0, uint32_t w1, float *z0, float *z1) {
    const double u1 = u01(w0), u2 = u01(w1);
    const double r = sqrt(-2.0 * log(u1));
    const double th = 6.283185307179586 * u2 - 3.141592653589793;
    *z0 = (float)(r * cos(th));
    *z1 = (float)(r * sin(th));
}

/* variates of steps [step0, step0+n_steps): z [n_steps][n] float, e [n_steps] float */
void dm_variates(uint64_t seed, uint32_t chain_id, uint32_t step0, int32_t n_steps, int32_t n, float *z, float *e) {
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const uint32_t t_end = step0 + (uint32_t)n_steps;
    for (uint32_t blk = step0 >> 2; (blk << 2) < t_end; ++blk) {
        for (int c = 0; c <= n; ++c) {
            uint32_t ctr[4] = {blk, c == n ? ACCEPT_SLOT : (uint32_t)c, chain_id, 0u}, w[4];
            float zz[4];
            dm_philox4x32_10(ctr, key, w);
            if (c < n) {
                box_muller(w[0], w[1], &zz[0], &zz[1]);
                box_muller(w[2], w[3], &zz[2], &zz[3]);
            }
            for (int k = 0; k < 4; ++k) {
                const uint32_t tk = (blk << 2) + k;
                if (tk < step0 || tk >= t_end) continue;
                if (c < n) z[(size_t)(tk - step0) * n + c] = zz[k];
                else e[tk - step0] = (float)(-log(u01(w[k])));
            }
        }
    }
}

static float warp_sum_f(float v[32]) {
    float t[32];
    for (int m = 16; m >= 1; m >>= 1) {
        for (int i = 0; i < 32; ++i) t[i] = v[i] + v[i ^ m];
        memcpy(v, t, sizeof(t));
    }
    return v[0];
}
static double warp_sum_d(double v[32]) {
    double t[32];
    for (int m = 16; m >= 1; m >>= 1) {
        for (int i = 0; i < 32; ++i) t[i] = v[i] + v[i ^ m];
        memcpy(v, t, sizeof(t));
    }
    return v[0];
}

/* exact FP64 -log L at x (and gt = G q + h when gt != NULL), device operation order */
double dm_exact_eval(const dm_problem *p, int pt, const double *x, double *gt) {
    const int np = p->npad, cpl = np / 32;
    double q[256], part[32];
    for (int c = 0; c < np; ++c) q[c] = x[c] - p->mu[c];
    const double f = p->f[pt];
    for (int lane = 0; lane < 32; ++lane) {
        double acc = 0.0;
        for (int r = 0; r < cpl; ++r) {
            const int c = lane + 32 * r;
            double v = 0.0, g = 0.0;
            for (int j = 0; j < p->n; ++j) {
                v = fma(p->saa[(size_t)j * np + c], q[j], v);
                if (gt) g = fma(p->g_t[((size_t)pt * np + j) * np + c], q[j], g);
            }
            acc = fma(q[c], fma(0.5, v, f * p->sab[c]), acc);
            if (gt) gt[c] = g + p->h[(size_t)pt * np + c];
        }
        part[lane] = acc;
    }
    return warp_sum_d(part) + p->c0[pt];
}

/*
 * MH steps [t_begin, t_end).  z_replay/e_replay (may be NULL) are indexed by (step - replay_step0).
 * on_row (may be NULL) is called for chain recording: kind 0 = accepted point, 1 = rejected, 2 = any step.
 */
typedef void (*dm_row_cb)(void *user, int kind, uint32_t step, const double *x, double ll);

static void run_steps(const dm_problem *p, int pt, uint64_t seed, uint32_t chain_id, uint32_t t_begin, uint32_t t_end,
                      float Tf, float sf, double *x, double *gt, double *ll, double *best_ll, double *bx,
                      int32_t *nacc, const float *z_replay, const float *e_replay, uint32_t replay_step0,
                      dm_row_cb cb, void *user) {
    const int np = p->npad, n = p->n, cpl = np / 32;
    const float *lt = p->lt_t + (size_t)pt * np * np;
    const float *lam = p->lam + (size_t)pt * np;
    float gf[256], hl[256], z[256];
    double xn[256];
    const float half_s = 0.5f * sf;
    for (int c = 0; c < np; ++c) { gf[c] = (float)gt[c]; hl[c] = half_s * lam[c]; }
    float *zbuf = NULL, *ebuf = NULL;
    if (!z_replay) {
        zbuf = (float *)malloc(sizeof(float) * (size_t)(t_end - t_begin) * n);
        ebuf = (float *)malloc(sizeof(float) * (size_t)(t_end - t_begin));
        dm_variates(seed, chain_id, t_begin, (int32_t)(t_end - t_begin), n, zbuf, ebuf);
    }
    for (uint32_t t = t_begin; t < t_end; ++t) {
        const float *zt = z_replay ? z_replay + (size_t)(t - replay_step0) * n : zbuf + (size_t)(t - t_begin) * n;
        const float ek = e_replay ? e_replay[t - replay_step0] : ebuf[t - t_begin];
        for (int c = 0; c < np; ++c) z[c] = c < n ? zt[c] : 0.0f;
        float lanes[32];
        for (int lane = 0; lane < 32; ++lane) {
            float tt = 0.0f;
            for (int r = 0; r < cpl; ++r) {
                const int c = lane + 32 * r;
                tt = fmaf(z[c], fmaf(hl[c], z[c], gf[c]), tt);
            }
            lanes[lane] = tt;
        }
        const float delta = sf * warp_sum_f(lanes);
        int accepted = 0;
        if (delta < Tf * ek) {
            int outside = 0;
            for (int c = 0; c < np; ++c) {
                float a[4] = {0.f, 0.f, 0.f, 0.f};
                for (int j = 0; j < np; ++j) a[j & 3] = fmaf(lt[(size_t)j * np + c], z[j], a[j & 3]);
                const float acc = (a[0] + a[1]) + (a[2] + a[3]);
                xn[c] = x[c] + (double)(sf * acc);
                if (xn[c] < p->lo[c] || xn[c] > p->hi[c]) outside = 1;
            }
            if (!outside) {
                accepted = 1;
                for (int c = 0; c < np; ++c) {
                    x[c] = xn[c];
                    gt[c] = gt[c] + (double)(2.0f * (hl[c] * z[c]));
                    gf[c] = (float)gt[c];
                }
                *ll = *ll + (double)delta;
                *nacc += 1;
                if (*ll < *best_ll) { *best_ll = *ll; memcpy(bx, x, sizeof(double) * np); }
            }
        }
        if (cb) { cb(user, accepted ? 0 : 1, t, x, *ll); cb(user, 2, t, x, *ll); }
    }
    free(zbuf); free(ebuf);
}

/* optimiser.py:87-162 + optimise_task.py:33-54, IEEE doubles in Python's evaluation order */
void dm_schedule_next(const dm_schedule *s, dm_chain *st, int32_t accepted, double best_loglkl) {
    st->iteration += 1;
    if (s->chi2_tolerance > 0.0) {
        const double improvement = st->current_best - best_loglkl;
        const int keep_going = (improvement < 0.0) || (st->current_best == INFINITY) ||
                               (2.0 * improvement > s->chi2_tolerance);
        if (!keep_going && accepted >= 1) { st->status = 1; return; }
    }
    const double ar = (double)(1 + accepted) / (double)(1 + s->steps_per_iteration);
    const double new_temp = st->temperature * s->temperature_change;
    double new_step = st->step_size;
    if (s->step_mode == 0) {
        new_step = st->step_size * s->step_size_change;
    } else if (s->step_mode == 1) {
        if (ar > s->adaptive_hi) new_step = st->step_size * (1.0 + s->adaptive_multiplier);
        else if (ar < s->adaptive_lo) new_step = st->step_size * (1.0 - s->adaptive_multiplier);
    } else if (st->stage < 2) {
        const double mag = st->stage == 0 ? 0.05 : 0.1;
        double m;
        if (ar > s->adaptive_hi) m = mag;
        else if (ar < s->adaptive_lo) m = -mag;
        else { st->status = 2; return; }
        new_step = st->step_size * (1.0 + m);
        if (st->stage == 0) { st->prev_mult = m; st->prev_ar = ar; } else { st->cur_mult = m; }
        st->stage += 1;
    } else {
        const double ar_desired = (s->adaptive_lo + s->adaptive_hi) / 2.0;
        double m = (ar_desired - ar) / (ar - st->prev_ar) * (st->cur_mult - st->prev_mult) + st->cur_mult;
        if (1.0 < m) m = 1.0;
        if (!(m > -0.9)) m = -0.9;
        new_step = st->step_size * (1.0 + m);
        st->prev_ar = ar; st->prev_mult = st->cur_mult; st->cur_mult = m;
    }
    st->current_best = best_loglkl;
    st->temperature = new_temp;
    st->step_size = new_step;
}

/*
 * n_iter annealing iterations of one chain.  Per-iteration outputs are rows [row0 + k] of the rec_* arrays
 * (row0 = the chain's iteration number at entry): best/last loglkl, T, step, accepted, best_x/last_x [npad].
 */
void dm_anneal_chain(const dm_problem *p, const dm_schedule *s, dm_chain *c, int32_t n_iter, uint64_t seed,
                     const float *z_replay, const float *e_replay, double *rec_best, double *rec_last,
                     double *rec_T, double *rec_step, int32_t *rec_acc, double *rec_best_x, double *rec_last_x) {
    const int np = p->npad;
    double gt[256], bx[256], tmp[256];
    if (c->status != 0) return;
    const uint32_t replay0 = c->steps_done;
    double ll = dm_exact_eval(p, c->point, c->x, gt);
    int row = 0;
    for (int it = 0; it < n_iter && c->status == 0; ++it, ++row) {
        double best_ll = ll;
        int32_t nacc = 0;
        memcpy(bx, c->x, sizeof(double) * np);
        run_steps(p, c->point, seed, c->chain_id, c->steps_done, c->steps_done + s->steps_per_iteration,
                  (float)c->temperature, (float)c->step_size, c->x, gt, &ll, &best_ll, bx, &nacc, z_replay, e_replay,
                  replay0, NULL, NULL);
        c->steps_done += s->steps_per_iteration;
        const double best_exact = dm_exact_eval(p, c->point, bx, NULL);
        ll = dm_exact_eval(p, c->point, c->x, gt);
        (void)tmp;
        rec_best[row] = best_exact; rec_last[row] = ll; rec_T[row] = c->temperature; rec_step[row] = c->step_size;
        rec_acc[row] = nacc;
        if (rec_best_x) memcpy(rec_best_x + (size_t)row * np, bx, sizeof(double) * np);
        if (rec_last_x) memcpy(rec_last_x + (size_t)row * np, c->x, sizeof(double) * np);
        dm_schedule_next(s, c, nacc, best_exact);
    }
    c->loglkl = ll;
}

/* chain recording for dm_mh_chain */
typedef struct {
    int mode, thin, capacity, np, count, cur_mult;
    double *x, *ll;
    int32_t *mult;
} dm_recorder;

static void rec_push(dm_recorder *r, const double *x, double ll, int mult) {
    if (r->count < r->capacity) {
        memcpy(r->x + (size_t)r->count * r->np, x, sizeof(double) * r->np);
        r->ll[r->count] = ll;
        r->mult[r->count] = mult;
    }
    r->count += 1;
}
static void rec_cb(void *user, int kind, uint32_t step, const double *x, double ll) {
    dm_recorder *r = (dm_recorder *)user;
    if (r->mode == 1) {
        if (kind == 0) {
            if (r->count - 1 < r->capacity) r->mult[r->count - 1] = r->cur_mult;
            r->cur_mult = 1;
            rec_push(r, x, ll, 1);
        } else if (kind == 1) {
            r->cur_mult += 1;
        }
    } else if (r->mode == 2 && kind == 2) {
        if ((step + 1) % (uint32_t)r->thin == 0) rec_push(r, x, ll, 1);
    }
}

/* n_steps MH steps with the device's 256-step re-anchoring; returns accepted count; *count in/out */
int32_t dm_mh_chain(const dm_problem *p, dm_chain *c, int32_t n_steps, uint64_t seed, const float *z_replay,
                    const float *e_replay, int mode, int thin, int capacity, double *sx, double *sll, int32_t *smult,
                    int32_t *count) {
    const int np = p->npad;
    double gt[256], bx[256];
    const uint32_t replay0 = c->steps_done;
    double ll = dm_exact_eval(p, c->point, c->x, gt);
    dm_recorder r = {mode, thin > 0 ? thin : 1, capacity, np, count ? *count : 0, 1, sx, sll, smult};
    if (mode == 1) {
        if (r.count == 0) rec_push(&r, c->x, ll, 1);
        else r.cur_mult = r.count - 1 < capacity ? smult[r.count - 1] : 1;
    }
    double best_ll = ll;
    int32_t nacc = 0;
    uint32_t t = c->steps_done;
    const uint32_t t_end = t + (uint32_t)n_steps;
    while (t < t_end) {
        const uint32_t t_next = t_end < t + 256u ? t_end : t + 256u;
        run_steps(p, c->point, seed, c->chain_id, t, t_next, (float)c->temperature, (float)c->step_size, c->x, gt, &ll,
                  &best_ll, bx, &nacc, z_replay, e_replay, replay0, mode ? rec_cb : NULL, &r);
        t = t_next;
        ll = dm_exact_eval(p, c->point, c->x, gt);
    }
    if (mode == 1 && r.count - 1 < capacity) smult[r.count - 1] = r.cur_mult;
    if (count) *count = r.count;
    c->steps_done = t;
    c->loglkl = ll;
    return nacc;
}
