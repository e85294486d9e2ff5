/*
 * pb200.h -- C ABI of libpb200.so: the B200 (sm_100a) implementation of PROSPECT's
 * simulated-annealing / Metropolis-Hastings hot path over the analytic Gaussian kernels.
 *
 * The reference (AarhusCosmology/prospect_public, pure Python) has no FFI; each entry point
 * below names the reference call site it replaces (paths relative to the reference root).
 * INTEGRATION.md shows the ctypes binding a PROSPECT maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; pb200_last_error() gives the text
 *     (thread-local); no exception ever crosses the ABI;
 *   - "device pointer" arguments are raw CUDA device addresses owned by the caller (the Python host
 *     allocates them through torch); `stream` is a cudaStream_t passed as void* (NULL = default);
 *     launches are asynchronous on that stream;
 *   - all matrices are zero-padded to npad = 32*ceil(n/32) and stored "input-major":
 *     M_t[j*npad + i] multiplies input j into output i, so a warp reads one input row coalesced;
 *   - chains are enumerated c = repetition * n_points + point
 *     (prospect/tasks/initialise_profile_task.py:16-26);
 *   - loglkl always means -log L (prospect/kernels/base_kernel.py:104-107).
 */
#ifndef PB200_H
#define PB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 4: struct pb200_problem gained lt_hi / lt_lo (tensor-core operands of the wide path); the stand-alone variate
 *    generator entry points of version 3 (pb200_variates_*, pb200_anneal_run_streamed: measured slower than the fused
 *    kernel at every batch size) were removed; pb200_exchange_arrivals and pb200_launch_status are new */
/* 5: additions only -- pb200_hosted, pb200_hosted_propose, pb200_hosted_accept (host-evaluated likelihoods) */
#define PB200_ABI_VERSION 5

/* chain status (pb200_chains.status) */
#define PB200_ACTIVE     0
#define PB200_CONVERGED  1   /* chi2_tolerance stop, tasks/optimise_task.py:33-50 */
#define PB200_FAILED     2   /* the reference task would have raised (optimiser.py:99-112 unbound local) */

/* step-size schedule modes (profile.step_size_schedule / step_size_adaptive_multiplier) */
#define PB200_STEP_EXPONENTIAL      0   /* optimiser.py:92-93 */
#define PB200_STEP_ADAPTIVE_FIXED   1   /* optimiser.py:138-147 */
#define PB200_STEP_ADAPTIVE_SECANT  2   /* optimiser.py:95-137 ('adaptive' multiplier) */

/* sample recording modes of pb200_mh_run */
#define PB200_RECORD_NONE      0
#define PB200_RECORD_ACCEPTED  1   /* accepted points + multiplicities == prospect/mcmc.py Chain */
#define PB200_RECORD_THINNED   2   /* current position every `thin` steps, weight 1 */
/* OR-ed into pb200_samples.mode: test the prior box exactly at EVERY likelihood-accepted step (one n x n proposal
 * mat-vec per accepted step) instead of deferring it behind the whitened safe radius.  Same decisions; faster when
 * steps are a sizeable fraction of that radius (wide-step sampling at T = 1), where the deferred test would have to
 * re-reference every few accepted steps anyway.  PB200_RECORD_ACCEPTED always works this way. */
#define PB200_EXACT_BOX        0x100

/*
 * The Gaussian target and the per-point proposal, precomputed on the host in FP64.
 * Replaces the data held by AnalyticalKernel (prospect/kernels/analytical_kernel.py:10-66)
 * plus the per-task `covmat` of optimise_settings (tasks/initialise_optimiser_task.py:78-92).
 *
 * With S = sym(inv(C)) and the reference's residual order [varying..., fixed]:
 *   loglkl(x) = 1/2 q^T S_aa q + f s_ab^T q + c0,   q = x - mu,  f = fixed_value - mu_fixed,
 *   c0 = 1/2 s_bb f^2.   (free runs: n = d, f = c0 = 0)
 * Proposal covariance of point p: Sigma_p = F F^T.  M = F^T S_aa F = Q diag(lam) Q^T and
 * Lt = F Q, so a proposal x' = x + step * Lt z (z ~ N(0, I)) changes loglkl by exactly
 *   Delta = step * z.gt + 1/2 step^2 sum_i lam_i z_i^2,   gt = Lt^T (S_aa q + f s_ab) = G q + h.
 */
typedef struct pb200_problem {
    int32_t d;            /* kernel dimension */
    int32_t n;            /* varying parameters (d-1 for jobtype profile, d otherwise) */
    int32_t npad;         /* 32, 64, 128 or 256 */
    int32_t n_points;     /* P: profile values (1 for global_optimisation / mcmc) */
    const double* saa;    /* device [npad*npad]  S_aa (symmetric) */
    const double* sab;    /* device [npad]       s_ab (zeros for free runs) */
    const double* mu;     /* device [npad]       means of the varying parameters, varying order */
    const double* lo;     /* device [npad]       prior box (pads / ignore_prior: -inf) */
    const double* hi;     /* device [npad]       prior box (pads / ignore_prior: +inf) */
    const double* f;      /* device [P] */
    const double* c0;     /* device [P] */
    const float*  lt_t;   /* device [P][npad*npad]  lt_t[j*npad+i] = Lt[i][j]  (FP32) */
    const float*  lam;    /* device [P][npad]       eigenvalues (FP32) */
    const float*  lt_norm;/* device [P][npad]       Euclidean norm of row i of Lt, rounded up (0 for pads): bounds
                             |x_i - xref_i| <= lt_norm_i * |w| for a whitened displacement w (deferred box test) */
    const double* g_t;    /* device [P][npad*npad]  g_t[j*npad+i] = G[i][j] */
    const double* h;      /* device [P][npad] */
    const float*  lt_hi;  /* device [P][npad*npad]  lt_hi[i*npad+j] = the TF32 part of Lt[i][j] (low 13 mantissa bits
                             cleared), K-major: operand of the tcgen05 materialisation GEMM of the wide path
                             (npad >= 128).  May be NULL (with lt_lo): that GEMM then runs on the FP32 SIMT pipe */
    const float*  lt_lo;  /* device [P][npad*npad]  lt_lo[i*npad+j] = Lt[i][j] - lt_hi[i*npad+j] (exact) */
} pb200_problem;

/*
 * Per-chain state, resident in HBM between launches.  Replaces the `optimise_settings` dict that
 * OptimiseTask hands from iteration to iteration (prospect/optimiser.py:151-162) and the tail of the
 * Chain object (prospect/mcmc.py:25-82).
 */
typedef struct pb200_chains {
    int32_t   n_chains;        /* B (local to this GPU) */
    double*   x;               /* device [B][npad]  current position (varying order, pads 0) */
    double*   loglkl;          /* device [B]        loglkl of x */
    double*   temperature;     /* device [B] */
    double*   step_size;       /* device [B] */
    double*   current_best;    /* device [B]  'current_best_loglkl' (optimiser.py:152) */
    const int32_t*  point;     /* device [B]  profile-point index of the chain */
    const uint32_t* chain_id;  /* device [B]  global chain id -> Philox counter word (GPU-count independent) */
    uint32_t* steps_done;      /* device [B]  MH steps taken so far -> Philox counter word */
    int32_t*  iteration;       /* device [B]  next iteration_number */
    int32_t*  status;          /* device [B]  PB200_ACTIVE / CONVERGED / FAILED */
    int32_t*  secant_stage;    /* device [B]  0,1: hard-coded multipliers; 2: secant (optimiser.py:96-137) */
    double*   secant_prev_mult;/* device [B] */
    double*   secant_cur_mult; /* device [B] */
    double*   secant_prev_ar;  /* device [B] */
} pb200_chains;

/* The `profile:` schedule knobs (prospect/profile.py:129-362) that the device needs. */
typedef struct pb200_schedule {
    int32_t steps_per_iteration;   /* S */
    int32_t step_mode;             /* PB200_STEP_* */
    double  temperature_change;    /* (T_end/T_start)^(1/max_iterations), initialise_optimiser_task.py:138 */
    double  step_size_change;      /* exponential rate, initialise_optimiser_task.py:146 */
    double  adaptive_lo;           /* step_size_adaptive_interval[0] */
    double  adaptive_hi;           /* step_size_adaptive_interval[1] */
    double  adaptive_multiplier;   /* numeric step_size_adaptive_multiplier */
    double  chi2_tolerance;        /* <= 0 or NaN: disabled (None / 0.0 in the yaml) */
} pb200_schedule;

/*
 * One record per (iteration, chain): what SimulatedAnnealing.set_bestfit produces
 * (prospect/optimiser.py:74-85) plus the settings the task ran with.  Row k = iteration_number.
 * accepted = -1 marks "chain did not run this iteration" (converged / failed earlier).
 * All arrays of one iteration may live in one contiguous block (that block is what ranks all-gather at an
 * iteration boundary): element (k, c) of an f64 array is at [k * iter_stride + c], of `accepted` at
 * [k * 2 * iter_stride + c], coordinate i of best_x / last_x at [k * iter_stride + c * npad + i].
 */
typedef struct pb200_records {
    int32_t  max_iterations;    /* rows allocated */
    int64_t  iter_stride;       /* distance between consecutive iterations, in doubles */
    double*  best_loglkl;       /* device [K][B]  exact FP64 loglkl of best_x */
    double*  last_loglkl;       /* device [K][B] */
    double*  temperature;       /* device [K][B] */
    double*  step_size;         /* device [K][B] */
    int32_t* accepted;          /* device [K][B]  accepted proposals (AR = (1+accepted)/(1+S)) */
    double*  best_x;            /* device [K][B][npad]  (may be NULL) */
    double*  last_x;            /* device [K][B][npad]  (may be NULL) */
} pb200_records;

/* Sample recording for jobtype mcmc (prospect/mcmc.py:69-73, tasks/mcmc_task.py:17-23). */
typedef struct pb200_samples {
    int32_t  mode;          /* PB200_RECORD_* [| PB200_EXACT_BOX] */
    int32_t  thin;          /* PB200_RECORD_THINNED: keep every thin-th step */
    int32_t  capacity;      /* rows per chain */
    double*  x;             /* device [B][capacity][npad] */
    double*  loglkl;        /* device [B][capacity] */
    int32_t* mult;          /* device [B][capacity] */
    int32_t* count;         /* device [B]  rows used so far (in/out; the open last row is included) */
} pb200_samples;

const char* pb200_last_error(void);
int pb200_abi_version(void);

/* Number of SMs etc. of `device` (for grid sizing reports). */
int pb200_device_info(int device, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/*
 * loglkl + prior box for M positions.  Replaces BaseKernel.loglkl / logprior
 * (prospect/kernels/base_kernel.py:72-102) -> AnalyticalKernel.Gaussian
 * (prospect/kernels/analytical_kernel.py:59-66).
 *   x [M][npad] device, point [M] device (NULL: all point 0), out_loglkl [M] device,
 *   out_outside [M] device int32 (1 = outside the prior box; may be NULL).
 */
int pb200_loglkl_batch(const pb200_problem* prob, const double* x, const int32_t* point, int32_t m,
                       double* out_loglkl, int32_t* out_outside, void* stream);

/*
 * n_iter annealing iterations of every active chain in ONE launch: proposal, quadratic form,
 * Philox accept/reject at the chain's temperature, running bestfit, then set_bestfit and
 * get_next_iteration_settings on the device.  Replaces, per chain and iteration,
 * OptimiseTask.run + emit_tasks (prospect/tasks/optimise_task.py:16-54),
 * SimulatedAnnealing.__init__/optimise/set_bestfit/get_next_iteration_settings
 * (prospect/optimiser.py:40-162) and MetropolisHastings.step/get_proposal (prospect/mcmc.py:163-190).
 *
 * Wide problems (npad >= 128): the dense contractions of an iteration boundary -- the proposal-factor product
 * x = xref + Lt w (prospect/mcmc.py:187) and the chi^2 quadratic forms / gradients (analytical_kernel.py:65) -- are
 * batched over the repetitions of a profile point: per iteration one launch of the Metropolis steps, one tcgen05
 * (3xTF32, TMA-fed) GEMM for the positions and one FP64 shared-memory-tiled GEMM for the anchors, all enqueued by
 * this one call.  That path addresses the repetitions of a point as chain = rep * n_points + point, i.e. it needs
 * chains.point[c] == c % n_points (a batch laid out differently ends with every chain in status PB200_FAILED and no
 * records).  PB200_WIDE=0 in the environment keeps the single fused kernel; PB200_WIDE_MMA=0 keeps the position GEMM
 * on the FP32 SIMT pipe (bit-identical to the fused kernel).
 */
int pb200_anneal_run(const pb200_problem* prob, const pb200_chains* chains, const pb200_schedule* sched,
                     const pb200_records* rec, int32_t n_iter, uint64_t seed, int32_t lanes_per_chain, void* stream);

/*
 * Exchange buffers of the ranks of a sharded job, all mapped into this process (CUDA peer memory over NVLink /
 * NVSwitch, e.g. torch symmetric memory).  With them the annealing kernel does the iteration-boundary all-gather
 * ITSELF: next to its own record row it stores the per-chain scalars of iteration k into slot `rank` of row k of
 * EVERY rank's buffer as soon as the iteration ends (posted stores that drain over NVLink while the next
 * iterations compute).  When a warp has finished all its iterations it issues ONE system-scope fence and adds 1 to
 * counter n_iter - 1 of every rank: on a rank, that counter reaches the total number of chain warps of the whole job
 * exactly when all gathered rows of the launch are complete there.  No collective call, no extra kernel, no packing.
 *   buffer[r]   [rows][n_peers][rank_stride] doubles; block of (row k, rank q) at k * row_stride + q * rank_stride:
 *               [best_loglkl | last_loglkl | temperature | step_size] (n_alloc doubles each), accepted (n_alloc int32)
 *   progress[r] [>= n_iter] uint32, zeroed (on every rank, before any rank launches) by the caller; only entry
 *               n_iter - 1 is used
 */
#define PB200_MAX_PEERS 16
typedef struct pb200_peers {
    int32_t   n_peers;                      /* ranks (0: no exchange) */
    int32_t   rank;                         /* this rank's slot */
    int32_t   n_alloc;                      /* chains per block */
    int32_t   rows;                         /* rows allocated in every buffer */
    int64_t   row_stride;                   /* doubles */
    int64_t   rank_stride;                  /* doubles, >= 4 * n_alloc + (n_alloc + 1) / 2 */
    double*   buffer[PB200_MAX_PEERS];
    uint32_t* progress[PB200_MAX_PEERS];
} pb200_peers;

/*
 * pb200_anneal_run_signalled with the iteration-boundary all-gather fused into the kernel (see pb200_peers).
 * `progress` / `out_target` as in pb200_anneal_run_signalled (local counters; may be NULL / NULL here).
 * Replaces the per-task result hand-off of the reference's scheduler (prospect/communication.py:106-157) for jobs
 * sharded over several GPUs.
 */
int pb200_anneal_run_shared(const pb200_problem* prob, const pb200_chains* chains, const pb200_schedule* sched,
                            const pb200_records* rec, int32_t n_iter, uint64_t seed, int32_t lanes_per_chain,
                            const pb200_peers* peers, uint32_t* progress, int32_t* out_target, void* stream);

/*
 * pb200_anneal_run that also SIGNALS its progress: `progress` is a device array of n_iter uint32 counters, zeroed
 * by the caller before the launch; counter k reaches *out_target (returned by this call) once every warp has made
 * its records of the launch's k-th iteration globally visible (or will never run it because all its chains
 * stopped).  A second stream can therefore ship iteration k's records -- pb200_stream_wait_geq(stream,
 * progress + k, target), then the all-gather -- while later iterations are still running in the same launch.
 * The kernel never waits on anything, so the scheme cannot dead-lock.  This is the iteration-boundary exchange of
 * the sharded task farm without giving up the fused multi-iteration launch.
 */
int pb200_anneal_run_signalled(const pb200_problem* prob, const pb200_chains* chains, const pb200_schedule* sched,
                               const pb200_records* rec, int32_t n_iter, uint64_t seed, int32_t lanes_per_chain,
                               uint32_t* progress, int32_t* out_target, void* stream);

/*
 * Number of arrivals a rank with `n_chains` chains adds to an arrival counter of pb200_anneal_run_signalled /
 * pb200_anneal_run_shared (= its number of chain warps for the geometry the library resolves `lanes_per_chain`
 * to, environment overrides included).  A rank of a sharded job sums this over all ranks' chain counts to get the
 * value the counters reach; 0 for n_chains <= 0, -1 (with pb200_last_error) for an invalid geometry.
 */
int pb200_exchange_arrivals(int32_t npad, int32_t n_chains, int32_t lanes_per_chain);

/*
 * Device-side protocol faults (a producer-warp ring or a TMA / tensor-core pipeline stage that never delivered within
 * its time-out) do not trap the context: the kernel records the fault, stops the affected chains (status
 * PB200_FAILED) and ends.  This call waits for the work enqueued on `stream` so far, then returns non-zero -- with
 * the description in pb200_last_error() -- if any launch since the previous call recorded a fault, and clears it.
 */
int pb200_launch_status(void* stream);

/* Enqueue on `stream` a wait until *counter >= value (cuStreamWaitValue32, CU_STREAM_WAIT_VALUE_GEQ). */
int pb200_stream_wait_geq(void* stream, const uint32_t* counter, uint32_t value);

/*
 * n_steps Metropolis-Hastings steps per chain at the chain's temperature / step size, optionally
 * recording the chain.  Replaces MCMCTask.run -> BaseMCMC.run_steps
 * (prospect/tasks/mcmc_task.py:17-23, prospect/mcmc.py:131-138).  accepted_out [B] device (may be NULL).
 */
int pb200_mh_run(const pb200_problem* prob, const pb200_chains* chains, int32_t n_steps, uint64_t seed,
                 const pb200_samples* samples, int32_t* accepted_out, int32_t lanes_per_chain, void* stream);

/*
 * lanes_per_chain (both samplers): how many lanes of a warp cooperate on one chain -- 32, 16 or 8 (16 / 8 only
 * for npad <= 64 / 32); 0 lets the library choose from the batch size.  A chain's random stream does not depend
 * on it; the floating-point summation order inside a step does (pb200_choose_lanes tells what 0 resolves to).
 */
int pb200_choose_lanes(int32_t npad, int32_t n_chains);

/*
 * 1 if pb200_anneal_run serves this (lanes_per_chain, n_chains) with the producer-warp kernel (anneal_ws_kernel:
 * every chain warp is paired with a warp that draws its variates into a shared-memory ring), 0 for the plain fused
 * kernel.  Results are identical either way; small batches run faster with producer warps.  PB200_WS=0/1 in the
 * environment forces the choice.
 */
int pb200_uses_producer_warps(int32_t lanes_per_chain, int32_t n_chains);

/*
 * Host-evaluated likelihoods: the kernels of external codes (cobaya, MontePython, any
 * BaseKernel subclass -- prospect/kernels/base_kernel.py:9-138) compute -log L on the CPU; everything else of
 * MetropolisHastings.step (prospect/mcmc.py:163-190) is batched here.  One step of all chains is
 *     pb200_hosted_propose  ->  (the host evaluates the proposals that are inside the box)  ->  pb200_hosted_accept.
 * Used fields of pb200_chains: x, loglkl, temperature, step_size, point, chain_id, steps_done, status (chains whose
 * status is not PB200_ACTIVE are left alone); pb200_schedule_update serves the annealing bookkeeping unchanged.
 */
typedef struct pb200_hosted {
    int32_t n;               /* varying parameters */
    int32_t npad;            /* row length of x / prop / best_x (>= n, a multiple of 32, <= 256) */
    int32_t n_points;        /* P: proposal covariances (one per profile point) */
    const double* factor_t;  /* device [P][npad*npad]  factor_t[j*npad+i] = F[i][j] with F F^T = covmat of the point
                                (tasks/initialise_optimiser_task.py:126-130); pads zero */
    const double* lo;        /* device [npad]  prior box (open sides / pads / ignore_prior: -inf) */
    const double* hi;        /* device [npad]  (+inf) */
} pb200_hosted;

/*
 * The proposal of every active chain, x' = x + step_size * F z with z ~ N(0, I) from the chain's Philox stream at its
 * step counter -- numpy's multivariate_normal(x, covmat * step_size^2) of prospect/mcmc.py:182-190 in distribution --
 * and its prior-box test (prospect/kernels/base_kernel.py:89-102; the reference skips the likelihood of an
 * out-of-box proposal, mcmc.py:166-167, and so may the host).  Nothing of the chain state changes.
 *   out_prop [B][npad] device (inactive chains: their current position), out_outside [B] device int32.
 */
int pb200_hosted_propose(const pb200_hosted* hosted, const pb200_chains* chains, uint64_t seed, double* out_prop,
                         int32_t* out_outside, void* stream);

/*
 * Accept / reject with the host's likelihood values: accept iff the proposal is inside the box and
 * exp((loglkl - prop_loglkl) / temperature) > u  (prospect/mcmc.py:170-180, strict), evaluated as
 * prop_loglkl - loglkl < temperature * e with e = -ln u from the chain's Philox stream (NaN / +inf never accept).
 * Accepted chains take the proposal (x, loglkl); every active chain's step counter advances by one.
 *   prop [B][npad], prop_loglkl [B], outside [B]: device (prop_loglkl of an out-of-box proposal is not read);
 *   accepted_flag [B] int32 out (may be NULL), n_accepted [B] int32 in/out (may be NULL),
 *   best_loglkl [B] / best_x [B][npad] in/out (may be NULL): running bestfit, first minimum (optimiser.py:75).
 */
int pb200_hosted_accept(const pb200_hosted* hosted, const pb200_chains* chains, uint64_t seed, const double* prop,
                        const double* prop_loglkl, const int32_t* outside, int32_t* accepted_flag,
                        int32_t* n_accepted, double* best_loglkl, double* best_x, void* stream);

/*
 * Stand-alone adaptive step-size / temperature bookkeeping for B chains (the same device function the
 * fused kernel calls).  Replaces SimulatedAnnealing.get_next_iteration_settings
 * (prospect/optimiser.py:87-162) + the chi2_tolerance test of OptimiseTask.emit_tasks.
 *   accepted [B], best_loglkl [B] device: this iteration's result.
 */
int pb200_schedule_update(const pb200_chains* chains, const pb200_schedule* sched,
                          const int32_t* accepted, const double* best_loglkl, void* stream);

/*
 * Running-bestfit reductions of the analysis side.  Replaces AnalyseProfileTask.sort_tasks
 * (prospect/tasks/analyse_profile_task.py:91-121): per chain the first minimum over iterations,
 * per point the first-minimum repetition and the mean over repetitions with a finite bestfit.
 *   best_loglkl / accepted: element (k, c) at best_loglkl[k * stride_best + c] / accepted[k * stride_accepted + c]
 *   (records: iter_stride and 2 * iter_stride; dense [K][B] arrays: b and b); n_iter rows; B = R * P (rep-major).
 *   out_chain_best [B], out_chain_iter [B], out_point_best [P], out_point_rep [P], out_point_avg [P].
 */
int pb200_profile_reduce(const double* best_loglkl, const int32_t* accepted, int64_t stride_best,
                         int64_t stride_accepted, int32_t n_iter, int32_t b, int32_t n_points, int32_t reps,
                         double* out_chain_best, int32_t* out_chain_iter,
                         double* out_point_best, int32_t* out_point_rep, double* out_point_avg, void* stream);

/*
 * Weighted moments of N rows: sum w, sum w x_i, sum w x_i x_j (FP64).  Replaces
 * compute_covariance_matrix (prospect/mcmc.py:84-97, np.cov(..., bias=True, fweights=mults)).
 *   x [N][ld] device (first n columns used), w [N] device int32 (NULL: all 1),
 *   out [1 + n + n*n] device, zeroed by the call.
 * Reproducible: per-CTA partial sums (stream-ordered scratch, <= 32 MB) added in a fixed order, no atomics -- the same
 * rows give the same bits on every run.
 */
int pb200_weighted_moments(const double* x, const int32_t* w, int64_t n_rows, int32_t ld, int32_t n,
                           double* out, void* stream);

/*
 * Per-chain weighted first moments of recorded chains (struct pb200_samples buffers): for every chain c over its
 * rows r < min(count[c], capacity), in row order,
 *   out_sw[c] = sum_r mult,   out_swx[c][i] = sum_r mult * x_i,   out_scaled[c][i] = out_swx[c][i] / sqrt(out_sw[c])
 * (out_scaled rows are 0 for an empty chain).  With pb200_weighted_moments over x (all rows, w = mult) and over
 * out_scaled (w = NULL) these are the sufficient statistics of the pooled covariance and of the Gelman-Rubin R-1 the
 * reference gets from getdist (prospect/tasks/mcmc_task.py:28-61, prospect/analysis.py:27-32); they are plain sums,
 * so ranks all-reduce them.
 *   x [B][capacity][ld] device, mult [B][capacity] device int32, count [B] device int32,
 *   out_sw [B], out_swx [B][n], out_scaled [B][n] device FP64 (out_scaled may be NULL).
 */
int pb200_chain_moments(const double* x, const int32_t* mult, const int32_t* count, int32_t n_chains,
                        int32_t capacity, int32_t ld, int32_t n, double* out_sw, double* out_swx,
                        double* out_scaled, void* stream);

/*
 * Test hook: the N(0,1) / Exp(1) variates the samplers draw for one chain, steps [step0, step0+n_steps).
 *   out_z [n_steps][n] device float, out_e [n_steps] device float.
 */
int pb200_rng_variates(uint64_t seed, uint32_t chain_id, uint32_t step0, int32_t n_steps, int32_t n,
                       float* out_z, float* out_e, void* stream);

/* Test hook: raw Philox-4x32-10 block (host side, no GPU needed). */
void pb200_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* PB200_H */
