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
 *
 * Device specifics that are mirrored: a chain is spread over `lpc` lanes (coordinate c lives on lane
 * c % lpc), sums are butterfly reductions over those lanes, and the prior-box test is deferred: the
 * chain keeps position = xref + Lt*w and only materialises it when |w|^2 leaves a provably safe radius.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ACCEPT_SLOT 0xffffffffu
#define MAXN 256

typedef struct {
    int32_t d, n, npad, n_points;
    const double *saa, *sab, *mu, *lo, *hi, *f, *c0;
    const float *lt_t, *lam, *lt_norm;
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

typedef struct {          /* in-flight state of a chain (device registers) */
    double xref[MAXN], gt[MAXN], bref[MAXN];
    float w[MAXN], bw[MAXN];
    double ll, best_ll;
    float m2;
    int32_t nacc;
} dm_state;

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
            float zz[4] = {0.f, 0.f, 0.f, 0.f};
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

/* butterfly reductions over the lpc lanes of a group; every lane ends with the same value */
static float group_sum_f(float *v, int lpc) {
    float t[32];
    for (int m = lpc / 2; m >= 1; m >>= 1) {
        for (int i = 0; i < lpc; ++i) t[i] = v[i] + v[i ^ m];
        memcpy(v, t, sizeof(float) * lpc);
    }
    return v[0];
}
static double group_sum_d(double *v, int lpc) {
    double t[32];
    for (int m = lpc / 2; m >= 1; m >>= 1) {
        for (int i = 0; i < lpc; ++i) t[i] = v[i] + v[i ^ m];
        memcpy(v, t, sizeof(double) * lpc);
    }
    return v[0];
}

/* exact FP64 -log L at x (and gt = G q + h when gt != NULL), device operation order */
double dm_exact_eval(const dm_problem *p, int pt, const double *x, double *gt, int lpc) {
    const int np = p->npad, cpl = np / lpc;
    double q[MAXN], part[32];
    for (int c = 0; c < np; ++c) q[c] = x[c] - p->mu[c];
    const double f = p->f[pt];
    for (int gl = 0; gl < lpc; ++gl) {
        double acc = 0.0;
        for (int r = 0; r < cpl; ++r) {
            const int c = gl + lpc * r;
            double v = 0.0, g = 0.0;
            for (int j = 0; j < p->n; ++j) {
                v = fma(p->saa[(size_t)j * np + c], q[j], v);
                if (gt) g = fma(p->g_t[((size_t)pt * np + j) * np + c], q[j], g);
            }
            acc = fma(q[c], fma(0.5, v, f * p->sab[c]), acc);
            if (gt) gt[c] = g + p->h[(size_t)pt * np + c];
        }
        part[gl] = acc;
    }
    return group_sum_d(part, lpc) + p->c0[pt];
}

/* out = base + Lt w: FP32 matvec, four interleaved partial sums over j, one FP64 add */
static void materialise(const dm_problem *p, int pt, const double *base, const float *w, double *out) {
    const int np = p->npad;
    const float *lt = p->lt_t + (size_t)pt * np * np;
    for (int c = 0; c < np; ++c) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < np; ++j) a[j & 3] = fmaf(lt[(size_t)j * np + c], w[j], a[j & 3]);
        out[c] = base[c] + (double)((a[0] + a[1]) + (a[2] + a[3]));
    }
}

/* squared whitened radius around base inside which base + Lt w provably stays in the prior box */
static float safe_radius2(const dm_problem *p, int pt, const double *base) {
    const int np = p->npad;
    double m = INFINITY;
    for (int c = 0; c < np; ++c) {
        const double nrm = (double)p->lt_norm[(size_t)pt * np + c];
        const double dist = fmin(p->hi[c] - base[c], base[c] - p->lo[c]);
        if (nrm > 0.0) m = fmin(m, dist / nrm);
    }
    m = fmax(m, 0.0) * 0.999;
    return (float)(m * m);
}

static void anchor(const dm_problem *p, int pt, const double *x, dm_state *s, int lpc) {
    s->ll = dm_exact_eval(p, pt, x, s->gt, lpc);
    memcpy(s->xref, x, sizeof(double) * p->npad);
    memset(s->w, 0, sizeof(float) * p->npad);
    s->m2 = safe_radius2(p, pt, s->xref);
}

/*
 * MH steps [t_begin, t_begin + n_steps).  z_replay/e_replay (may be NULL) are indexed by (step - replay_step0).
 * rec_mode: 0 none, 1 accepted points (every accepted move is materialised), 2 thinned.
 */
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

/* per-run constants + scratch shared by the step routines */
typedef struct {
    const dm_problem *p;
    int pt, lpc;
    int exact;              /* PB200_EXACT_BOX: every likelihood-accepted step is materialised and box-tested */
    int boxtest;            /* block_steps tests its accepted positions in place instead of asking for a replay */
    float Tf, sf;
    float gf[MAXN], hl[MAXN];
} dm_run;

/* one sequential MH step (device: one_step) */
static void seq_step(dm_run *R, dm_state *s, const float *zt, float ek, uint32_t t, dm_recorder *rec) {
    const dm_problem *p = R->p;
    const int np = p->npad, n = p->n, lpc = R->lpc, cpl = np / lpc, pt = R->pt;
    const float sf = R->sf, Tf = R->Tf;
    float z[MAXN], wt[MAXN], lanes[32];
    double xt[MAXN];
    const int always_exact = R->exact || (rec && rec->mode == 1);
    for (int c = 0; c < np; ++c) z[c] = c < n ? zt[c] : 0.0f;
    for (int gl = 0; gl < lpc; ++gl) {
        float tt = 0.0f;
        for (int r = 0; r < cpl; ++r) {
            const int c = gl + lpc * r;
            tt = fmaf(z[c], fmaf(R->hl[c], z[c], R->gf[c]), tt);
        }
        lanes[gl] = tt;
    }
    const float delta = sf * group_sum_f(lanes, lpc);
    int acc = 0;
    if (delta < Tf * ek) {
        for (int gl = 0; gl < lpc; ++gl) {
            float rr = 0.0f;
            for (int r = 0; r < cpl; ++r) {
                const int c = gl + lpc * r;
                wt[c] = fmaf(sf, z[c], s->w[c]);
                rr = fmaf(wt[c], wt[c], rr);
            }
            lanes[gl] = rr;
        }
        const float rr = group_sum_f(lanes, lpc);
        acc = 1;
        if (always_exact || !(rr <= s->m2)) {
            materialise(p, pt, s->xref, wt, xt);
            int outside = 0;
            for (int c = 0; c < np; ++c)
                if (xt[c] < p->lo[c] || xt[c] > p->hi[c]) outside = 1;
            if (outside) {
                acc = 0;
            } else {
                memcpy(s->xref, xt, sizeof(double) * np);
                memset(wt, 0, sizeof(float) * np);
                s->m2 = safe_radius2(p, pt, s->xref);
            }
        }
        if (acc) {
            for (int c = 0; c < np; ++c) {
                s->w[c] = wt[c];
                R->gf[c] = fmaf(R->hl[c] + R->hl[c], z[c], R->gf[c]);   /* FP32 gradient between two FP64 anchors */
            }
            s->ll = s->ll + (double)delta;
            s->nacc += 1;
            if (s->ll < s->best_ll) {
                s->best_ll = s->ll;
                memcpy(s->bw, s->w, sizeof(float) * np);
                memcpy(s->bref, s->xref, sizeof(double) * np);
            }
        }
    }
    if (rec && rec->mode == 1) {
        if (acc) {
            if (rec->count - 1 < rec->capacity) rec->mult[rec->count - 1] = rec->cur_mult;
            rec->cur_mult = 1;
            rec_push(rec, s->xref, s->ll, 1);
        } else {
            rec->cur_mult += 1;
        }
    } else if (rec && rec->mode == 2) {
        if ((t + 1) % (uint32_t)rec->thin == 0) {
            materialise(p, pt, s->xref, s->w, xt);
            rec_push(rec, xt, s->ll, 1);
        }
    }
}

/*
 * Four steps of one Philox block at once (device: block_step): 16 dot products reduced together, the accept
 * decisions resolved with scalars.  zb[k] points at the n normals of step k.  Returns 1 (nothing committed) when the
 * block must be replayed step by step because the box bound is not provably safe.
 */
static int block_steps(dm_run *R, dm_state *s, const float *zb[4], const float e[4]) {
    const dm_problem *p = R->p;
    const int np = p->npad, n = p->n, lpc = R->lpc, cpl = np / lpc;
    const float sf = R->sf, Tf = R->Tf;
    float v[16][32];
    for (int gl = 0; gl < lpc; ++gl) {
        float a[16];
        for (int i = 0; i < 16; ++i) a[i] = 0.0f;
        for (int r = 0; r < cpl; ++r) {
            const int c = gl + lpc * r;
            const float g = R->gf[c], h = R->hl[c];
            float z[4], hz[4];
            for (int k = 0; k < 4; ++k) z[k] = c < n ? zb[k][c] : 0.0f;
            for (int k = 0; k < 4; ++k) {
                hz[k] = h * z[k];
                a[k] = fmaf(z[k], g, a[k]);
                a[4 + k] = fmaf(hz[k], z[k], a[4 + k]);
                a[14] = fmaf(z[k], z[k], a[14]);
            }
            a[8] = fmaf(hz[0], z[1], a[8]);
            a[9] = fmaf(hz[0], z[2], a[9]);
            a[10] = fmaf(hz[0], z[3], a[10]);
            a[11] = fmaf(hz[1], z[2], a[11]);
            a[12] = fmaf(hz[1], z[3], a[12]);
            a[13] = fmaf(hz[2], z[3], a[13]);
            a[15] = fmaf(s->w[c], s->w[c], a[15]);
        }
        for (int i = 0; i < 16; ++i) v[i][gl] = a[i];
    }
    float S[16];
    for (int i = 0; i < 16; ++i) S[i] = group_sum_f(v[i], lpc);
    float delta[4];
    int acc[4];
    delta[0] = sf * (S[0] + S[4]);
    acc[0] = delta[0] < Tf * e[0];
    const float x1 = acc[0] ? S[8] + S[8] : 0.0f;
    delta[1] = sf * ((S[1] + S[5]) + x1);
    acc[1] = delta[1] < Tf * e[1];
    const float x2 = (acc[0] ? S[9] + S[9] : 0.0f) + (acc[1] ? S[11] + S[11] : 0.0f);
    delta[2] = sf * ((S[2] + S[6]) + x2);
    acc[2] = delta[2] < Tf * e[2];
    const float x3 = ((acc[0] ? S[10] + S[10] : 0.0f) + (acc[1] ? S[12] + S[12] : 0.0f)) + (acc[2] ? S[13] + S[13] : 0.0f);
    delta[3] = sf * ((S[3] + S[7]) + x3);
    acc[3] = delta[3] < Tf * e[3];
    const int any = acc[0] || acc[1] || acc[2] || acc[3];
    const float s2 = sf + sf;
    const float half_bnd2 = fmaf(s2 * s2, S[14], S[15]);
    if (any && !((half_bnd2 + half_bnd2) <= s->m2)) {
        if (!R->boxtest) return 1;
        /* MH kernels: test the positions the block would accept (the commit loop's own FMAs, one FP32 mat-vec each);
           all inside -> commit without re-referencing, one outside -> replay step by step */
        float wt[MAXN];
        double xt[MAXN];
        memcpy(wt, s->w, sizeof(float) * np);
        for (int k = 0; k < 4; ++k) {
            if (!acc[k]) continue;
            for (int c = 0; c < np; ++c) wt[c] = fmaf(sf, c < n ? zb[k][c] : 0.0f, wt[c]);
            materialise(p, R->pt, s->xref, wt, xt);
            for (int c = 0; c < np; ++c)
                if (xt[c] < p->lo[c] || xt[c] > p->hi[c]) return 1;
        }
    }
    for (int k = 0; k < 4; ++k) {
        s->ll = s->ll + (acc[k] ? (double)delta[k] : 0.0);      /* the kernel adds +0.0 for a rejected step */
        if (!acc[k]) continue;
        for (int c = 0; c < np; ++c) {
            const float zc = c < n ? zb[k][c] : 0.0f;
            s->w[c] = fmaf(sf, zc, s->w[c]);
            R->gf[c] = fmaf(R->hl[c] + R->hl[c], zc, R->gf[c]);
        }
        s->nacc += 1;
        if (s->ll < s->best_ll) {
            s->best_ll = s->ll;
            memcpy(s->bw, s->w, sizeof(float) * np);
            memcpy(s->bref, s->xref, sizeof(double) * np);
        }
    }
    return 0;
}

static void run_steps(const dm_problem *p, int pt, uint64_t seed, uint32_t chain_id, uint32_t t_begin,
                      uint32_t n_steps, float Tf, float sf, int lpc, dm_state *s, const float *z_replay,
                      const float *e_replay, uint32_t replay_step0, dm_recorder *rec, int exact, int boxtest) {
    const int np = p->npad, n = p->n;
    const float *lam = p->lam + (size_t)pt * np;
    dm_run R;
    R.p = p; R.pt = pt; R.lpc = lpc; R.Tf = Tf; R.sf = sf; R.exact = exact; R.boxtest = boxtest;
    const float half_s = 0.5f * sf;
    for (int c = 0; c < np; ++c) { R.gf[c] = (float)s->gt[c]; R.hl[c] = half_s * lam[c]; }
    float *zbuf = NULL, *ebuf = NULL;
    if (!z_replay) {
        zbuf = (float *)malloc(sizeof(float) * (size_t)n_steps * n);
        ebuf = (float *)malloc(sizeof(float) * (size_t)n_steps);
        dm_variates(seed, chain_id, t_begin, (int32_t)n_steps, n, zbuf, ebuf);
    }
    const float *zs = z_replay ? z_replay + (size_t)(t_begin - replay_step0) * n : zbuf;
    const float *es = e_replay ? e_replay + (t_begin - replay_step0) : ebuf;
    const int recording = rec && rec->mode != 0;
    uint32_t i = 0;
    while (i < n_steps) {
        const uint32_t t = t_begin + i;
        /* a whole Philox block in the geometries that take four steps per reduction: every geometry of npad = 32, and
           32 lanes with >= 4 coordinates per lane (npad >= 128) */
        /* (thinned recording: blocks with no record step in them as well) */
        const int no_record = !recording || (rec->mode == 2 && (t % (uint32_t)rec->thin) + 4u < (uint32_t)rec->thin);
        if ((np == 32 || (lpc == 32 && np >= 128)) && no_record && !exact && (t & 3u) == 0u && n_steps - i >= 4u) {
            const float *zb[4] = {zs + (size_t)i * n, zs + (size_t)(i + 1) * n, zs + (size_t)(i + 2) * n,
                                  zs + (size_t)(i + 3) * n};
            if (block_steps(&R, s, zb, es + i))
                for (int k = 0; k < 4; ++k) seq_step(&R, s, zb[k], es[i + k], t + k, rec);
            i += 4;
        } else {
            seq_step(&R, s, zs + (size_t)i * n, es[i], t, rec);
            i += 1;
        }
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
 * n_iter annealing iterations of one chain.  Per-iteration outputs are rows [k] of the rec_* arrays
 * (k counted from the chain's iteration number at entry): best/last loglkl, T, step, accepted, best_x/last_x.
 */
void dm_anneal_chain(const dm_problem *p, const dm_schedule *s, dm_chain *c, int32_t n_iter, uint64_t seed, int lpc,
                     const float *z_replay, const float *e_replay, double *rec_best, double *rec_last,
                     double *rec_T, double *rec_step, int32_t *rec_acc, double *rec_best_x, double *rec_last_x) {
    const int np = p->npad;
    double bx[MAXN];
    dm_state st;
    if (c->status != 0) return;
    const uint32_t replay0 = c->steps_done;
    anchor(p, c->point, c->x, &st, lpc);
    int row = 0;
    for (int it = 0; it < n_iter && c->status == 0; ++it, ++row) {
        st.best_ll = st.ll;
        st.nacc = 0;
        memcpy(st.bref, st.xref, sizeof(double) * np);
        memset(st.bw, 0, sizeof(float) * np);
        run_steps(p, c->point, seed, c->chain_id, c->steps_done, (uint32_t)s->steps_per_iteration,
                  (float)c->temperature, (float)c->step_size, lpc, &st, z_replay, e_replay, replay0, NULL, 0, 0);
        c->steps_done += s->steps_per_iteration;
        materialise(p, c->point, st.bref, st.bw, bx);
        materialise(p, c->point, st.xref, st.w, c->x);
        const double best_exact = dm_exact_eval(p, c->point, bx, NULL, lpc);
        const int32_t nacc = st.nacc;
        anchor(p, c->point, c->x, &st, lpc);
        rec_best[row] = best_exact; rec_last[row] = st.ll; rec_T[row] = c->temperature; rec_step[row] = c->step_size;
        rec_acc[row] = nacc;
        if (rec_best_x) memcpy(rec_best_x + (size_t)row * np, bx, sizeof(double) * np);
        if (rec_last_x) memcpy(rec_last_x + (size_t)row * np, c->x, sizeof(double) * np);
        dm_schedule_next(s, c, nacc, best_exact);
    }
    c->loglkl = st.ll;
}

/* n_steps MH steps with the device's 256-step re-anchoring; returns accepted count; *count in/out */
int32_t dm_mh_chain(const dm_problem *p, dm_chain *c, int32_t n_steps, uint64_t seed, int lpc, const float *z_replay,
                    const float *e_replay, int mode, int thin, int capacity, double *sx, double *sll, int32_t *smult,
                    int32_t *count) {
    const int np = p->npad;
    const int exact = (mode & 0x100) != 0 || (mode & 0xff) == 1;      /* PB200_EXACT_BOX / RECORD_ACCEPTED */
    mode &= 0xff;
    dm_state st;
    const uint32_t replay0 = c->steps_done;
    anchor(p, c->point, c->x, &st, lpc);
    st.best_ll = st.ll;
    st.nacc = 0;
    memcpy(st.bref, st.xref, sizeof(double) * np);
    memset(st.bw, 0, sizeof(float) * np);
    dm_recorder r = {mode, thin > 0 ? thin : 1, capacity, np, count ? *count : 0, 1, sx, sll, smult};
    if (mode == 1) {
        if (r.count == 0) rec_push(&r, c->x, st.ll, 1);
        else r.cur_mult = r.count - 1 < capacity ? smult[r.count - 1] : 1;
    }
    uint32_t done = 0;
    while (done < (uint32_t)n_steps) {
        const uint32_t chunk = (uint32_t)n_steps - done < 256u ? (uint32_t)n_steps - done : 256u;
        run_steps(p, c->point, seed, c->chain_id, c->steps_done + done, chunk, (float)c->temperature,
                  (float)c->step_size, lpc, &st, z_replay, e_replay, replay0, mode ? &r : NULL, exact,
                  1 /* the MH kernels test a block's accepted positions in place */);
        done += chunk;
        materialise(p, c->point, st.xref, st.w, c->x);
        anchor(p, c->point, c->x, &st, lpc);
    }
    if (mode == 1 && r.count - 1 < capacity) smult[r.count - 1] = r.cur_mult;
    if (count) *count = r.count;
    c->steps_done += (uint32_t)n_steps;
    c->loglkl = st.ll;
    return st.nacc;
}
