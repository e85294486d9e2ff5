// pb200_kernels.cu -- sm_100a kernels + the C ABI of include/pb200.h.
//
// Kernels (DESIGN.md has the roofline of each):
//   anneal_kernel<CPL>        fused annealing iterations: Philox normals -> whitened quadratic form ->
//                             accept/reject at the chain temperature -> (on accept) proposal matvec, box
//                             test, state update, running bestfit -> set_bestfit + next settings.
//   mh_kernel<CPL, REC>       plain Metropolis-Hastings rounds with chain recording (jobtype mcmc).
//   loglkl_kernel<CPL>        FP64 chi^2 quadratic form + prior box for a batch of positions.
//   schedule_kernel           stand-alone adaptive step-size bookkeeping.
//   chain_min_kernel + point_reduce_kernel   min over iterations / first-min repetition / mean (warp shuffles).
//   moments_kernel            weighted first/second moments (covariance reduction), FP64.
//   rng_kernel                test hook exposing the variates.
// Chain state lives in HBM between launches and in registers inside one; there is no CPU fallback.
#include "pb200_device.cuh"

#ifndef PB200_LT_REGS
#define PB200_LT_REGS 0      // 1: keep this lane's column of the proposal factor in registers (n <= 32).
                             // 0 (measured faster on B200): read it through L1 on accept, 72 registers, 28 warps/SM
#endif
#ifndef PB200_MIN_CTAS
#define PB200_MIN_CTAS 7
#endif

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

namespace pb200 {

thread_local std::string g_error;

static int fail(const std::string& msg) {
    g_error = msg;
    return 1;
}
static int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    return fail(std::string(what) + ": " + cudaGetErrorString(e));
}

// Per-warp recording cursor of mh_kernel.
struct RecCursor {
    double* x;        // this chain's [capacity][npad]
    double* ll;       // [capacity]
    int32_t* mult;    // [capacity]
    int32_t capacity, thin, count, cur_mult;
};

template <int CPL>
__device__ __forceinline__ void record_row(RecCursor& rc, const double (&x)[CPL], double ll, int32_t mult, int lane) {
    constexpr int NPAD = 32 * CPL;
    if (rc.count < rc.capacity) {
        double* row = rc.x + (size_t)rc.count * NPAD;
#pragma unroll
        for (int r = 0; r < CPL; ++r) row[lane + 32 * r] = x[r];
        if (lane == 0) {
            rc.ll[rc.count] = ll;
            rc.mult[rc.count] = mult;
        }
    }
    rc.count += 1;   // count > capacity tells the host that rows were dropped
}

// ----------------------------------------------------------------------------------------------
// MH steps [t_begin, t_end) of one chain (prospect/mcmc.py:163-190, batched and re-ordered):
//   Delta   = step * sum_i z_i (gt_i + 1/2 step lam_i z_i)        O(n), FP32, every step
//   accept  = Delta < T * e   and   x + step * Lt z inside the box  (matvec only if the first holds)
// The reference draws the uniform even for out-of-box proposals and never evaluates the likelihood
// there; "inside AND likelihood-accept" is the same predicate evaluated in the cheaper order.
// ----------------------------------------------------------------------------------------------
// One MH step given this lane's normals zk[] and the accept variate ek.
template <int CPL, bool LT_REGS, int REC>
__device__ __forceinline__ void one_step(const float* __restrict__ lt_pt, int lane, uint32_t tk, float Tf, float sf,
                                         const float (&zk)[CPL], float ek, const float (&ltrow)[LT_REGS ? 32 : 1],
                                         const float (&hl)[CPL], float (&gf)[CPL], const double (&lo)[CPL],
                                         const double (&hi)[CPL], double (&x)[CPL], double (&gt)[CPL], double& ll,
                                         double& best_ll, double (&bx)[CPL], int32_t& nacc, float* zs, RecCursor& rc) {
    constexpr int NPAD = 32 * CPL;
    float tt = 0.0f;
#pragma unroll
    for (int r = 0; r < CPL; ++r) tt = __fmaf_rn(zk[r], __fmaf_rn(hl[r], zk[r], gf[r]), tt);
    tt = warp_sum(tt);
    const float delta = __fmul_rn(sf, tt);
    bool accepted = false;
    if (delta < __fmul_rn(Tf, ek)) {
#pragma unroll
        for (int r = 0; r < CPL; ++r) zs[lane + 32 * r] = zk[r];
        __syncwarp();
        double xn[CPL];
        bool outside = false;
        const float4* __restrict__ z4 = reinterpret_cast<const float4*>(zs);
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            if (LT_REGS) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 v = z4[j];
                    a0 = __fmaf_rn(ltrow[4 * j + 0], v.x, a0);
                    a1 = __fmaf_rn(ltrow[4 * j + 1], v.y, a1);
                    a2 = __fmaf_rn(ltrow[4 * j + 2], v.z, a2);
                    a3 = __fmaf_rn(ltrow[4 * j + 3], v.w, a3);
                }
            } else {
                const float* __restrict__ col = lt_pt + lane + 32 * r;
#pragma unroll 8
                for (int j = 0; j < NPAD / 4; ++j) {
                    const float4 v = z4[j];
                    a0 = __fmaf_rn(__ldg(col + (4 * j + 0) * NPAD), v.x, a0);
                    a1 = __fmaf_rn(__ldg(col + (4 * j + 1) * NPAD), v.y, a1);
                    a2 = __fmaf_rn(__ldg(col + (4 * j + 2) * NPAD), v.z, a2);
                    a3 = __fmaf_rn(__ldg(col + (4 * j + 3) * NPAD), v.w, a3);
                }
            }
            const float acc = __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
            xn[r] = __dadd_rn(x[r], (double)__fmul_rn(sf, acc));
            outside = outside || (xn[r] < lo[r]) || (xn[r] > hi[r]);
        }
        __syncwarp();
        if (!__any_sync(kFull, outside)) {
            accepted = true;
#pragma unroll
            for (int r = 0; r < CPL; ++r) {
                x[r] = xn[r];
                gt[r] = __dadd_rn(gt[r], (double)__fmul_rn(2.0f, __fmul_rn(hl[r], zk[r])));
                gf[r] = (float)gt[r];
            }
            ll = __dadd_rn(ll, (double)delta);
            nacc += 1;
            if (ll < best_ll) {
                best_ll = ll;
#pragma unroll
                for (int r = 0; r < CPL; ++r) bx[r] = x[r];
            }
        }
    }
    if (REC == PB200_RECORD_ACCEPTED) {
        if (accepted) {
            if (lane == 0 && rc.count - 1 < rc.capacity) rc.mult[rc.count - 1] = rc.cur_mult;
            rc.cur_mult = 1;
            record_row<CPL>(rc, x, ll, 1, lane);
        } else {
            rc.cur_mult += 1;
        }
    } else if (REC == PB200_RECORD_THINNED) {
        if ((tk + 1) % (uint32_t)rc.thin == 0) record_row<CPL>(rc, x, ll, 1, lane);
    }
}

// ----------------------------------------------------------------------------------------------
// MH steps [t_begin, t_end) of one chain (prospect/mcmc.py:163-190, batched and re-ordered):
//   Delta   = step * sum_i z_i (gt_i + 1/2 step lam_i z_i)        O(n), FP32, every step
//   accept  = Delta < T * e   and   x + step * Lt z inside the box  (matvec only if the first holds)
// The reference draws the uniform even for out-of-box proposals and never evaluates the likelihood
// there; "inside AND likelihood-accept" is the same predicate evaluated in the cheaper order.
// Philox blocks cover 4 consecutive steps: whole blocks run a branch-free unrolled body, the ragged
// head / tail of the range goes through one generic copy.
// ----------------------------------------------------------------------------------------------
template <int CPL, bool LT_REGS, int REC>
__device__ __forceinline__ void run_steps(const pb200_problem& p, const float* __restrict__ lt_pt, int lane,
                                          uint32_t k0, uint32_t k1, uint32_t chain_id, uint32_t t_begin,
                                          uint32_t t_end, float Tf, float sf, const float (&ltrow)[LT_REGS ? 32 : 1],
                                          const float (&lam)[CPL], const double (&lo)[CPL], const double (&hi)[CPL],
                                          double (&x)[CPL], double (&gt)[CPL], double& ll, double& best_ll,
                                          double (&bx)[CPL], int32_t& nacc, float* zs, RecCursor& rc) {
    float gf[CPL], hl[CPL];
    const float half_s = __fmul_rn(0.5f, sf);
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        gf[r] = (float)gt[r];
        hl[r] = __fmul_rn(half_s, lam[r]);
    }
    uint32_t t = t_begin;
    while (t < t_end) {
        const uint32_t blk = t >> 2, base = blk << 2;
        float z[CPL][4], e[4];
        draw_block<CPL>(k0, k1, chain_id, blk, p.n, lane, z, e);
        if (t == base && t_end - base >= 4u) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float zk[CPL];
#pragma unroll
                for (int r = 0; r < CPL; ++r) zk[r] = z[r][k];
                one_step<CPL, LT_REGS, REC>(lt_pt, lane, base + k, Tf, sf, zk, e[k], ltrow, hl, gf, lo, hi, x, gt, ll,
                                            best_ll, bx, nacc, zs, rc);
            }
            t = base + 4;
        } else {
            const uint32_t stop = min(base + 4u, t_end);
#pragma unroll 1
            for (uint32_t tk = t; tk < stop; ++tk) {
                const int k = (int)(tk - base);
                float zk[CPL];
#pragma unroll
                for (int r = 0; r < CPL; ++r) zk[r] = k == 0 ? z[r][0] : (k == 1 ? z[r][1] : (k == 2 ? z[r][2] : z[r][3]));
                const float ek = k == 0 ? e[0] : (k == 1 ? e[1] : (k == 2 ? e[2] : e[3]));
                one_step<CPL, LT_REGS, REC>(lt_pt, lane, tk, Tf, sf, zk, ek, ltrow, hl, gf, lo, hi, x, gt, ll, best_ll,
                                            bx, nacc, zs, rc);
            }
            t = stop;
        }
    }
}

template <int CPL>
struct LaneConst {
    float lam[CPL];
    double lo[CPL], hi[CPL];
};

template <int CPL, bool LT_REGS>
__device__ __forceinline__ void load_lane_const(const pb200_problem& p, int pt, int lane, LaneConst<CPL>& lc,
                                                float (&ltrow)[LT_REGS ? 32 : 1]) {
    constexpr int NPAD = 32 * CPL;
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        const int c = lane + 32 * r;
        lc.lam[r] = __ldg(p.lam + (size_t)pt * NPAD + c);
        lc.lo[r] = __ldg(p.lo + c);
        lc.hi[r] = __ldg(p.hi + c);
    }
    if (LT_REGS) {
        const float* lt = p.lt_t + (size_t)pt * NPAD * NPAD + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) ltrow[j] = __ldg(lt + j * NPAD);
    }
}

// ----------------------------------------------------------------------------------------------
// Fused annealing kernel: n_iter iterations x S steps per launch, one warp per chain.
// ----------------------------------------------------------------------------------------------
template <int CPL>
__global__ void __launch_bounds__(32 * kWarpsPerCta, (CPL == 1) ? PB200_MIN_CTAS : 1)
anneal_kernel(pb200_problem p, pb200_chains c, pb200_schedule s, pb200_records rec, int n_iter, uint32_t k0,
              uint32_t k1) {
    constexpr int NPAD = 32 * CPL;
    constexpr bool LT_REGS = (CPL == 1) && PB200_LT_REGS;
    __shared__ __align__(16) double stage[kWarpsPerCta][NPAD];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int chain = blockIdx.x * kWarpsPerCta + wib;
    if (chain >= c.n_chains) return;
    SchedState st;
    st.status = c.status[chain];
    if (st.status != PB200_ACTIVE) return;
    st.temperature = c.temperature[chain];
    st.step_size = c.step_size[chain];
    st.current_best = c.current_best[chain];
    st.prev_mult = c.secant_prev_mult[chain];
    st.cur_mult = c.secant_cur_mult[chain];
    st.prev_ar = c.secant_prev_ar[chain];
    st.stage = c.secant_stage[chain];
    st.iteration = c.iteration[chain];
    const int pt = c.point[chain];
    const uint32_t chain_id = c.chain_id[chain];
    uint32_t t = c.steps_done[chain];

    LaneConst<CPL> lc;
    float ltrow[LT_REGS ? 32 : 1];
    load_lane_const<CPL, LT_REGS>(p, pt, lane, lc, ltrow);
    const float* lt_pt = p.lt_t + (size_t)pt * NPAD * NPAD;

    double x[CPL], gt[CPL], bx[CPL];
#pragma unroll
    for (int r = 0; r < CPL; ++r) x[r] = c.x[(size_t)chain * NPAD + lane + 32 * r];
    double* qs = stage[wib];
    float* zs = reinterpret_cast<float*>(stage[wib]);
    double ll = exact_eval<CPL, true>(p, pt, x, lane, qs, gt);   // loglkl of the start point (optimiser.py:55)
    RecCursor rc{};

    for (int it = 0; it < n_iter && st.status == PB200_ACTIVE; ++it) {
        double best_ll = ll;
        int32_t nacc = 0;
#pragma unroll
        for (int r = 0; r < CPL; ++r) bx[r] = x[r];
        run_steps<CPL, LT_REGS, PB200_RECORD_NONE>(p, lt_pt, lane, k0, k1, chain_id, t, t + s.steps_per_iteration,
                                                   (float)st.temperature, (float)st.step_size, ltrow, lc.lam, lc.lo,
                                                   lc.hi, x, gt, ll, best_ll, bx, nacc, zs, rc);
        t += s.steps_per_iteration;
        // set_bestfit: the reported chi^2 is re-evaluated in FP64 at the reported position
        double unused[CPL];
        const double best_exact = exact_eval<CPL, false>(p, pt, bx, lane, qs, unused);
        ll = exact_eval<CPL, true>(p, pt, x, lane, qs, gt);       // re-anchor the incremental state
        const int row = st.iteration;
        if (row < rec.max_iterations) {
            const size_t o = (size_t)row * c.n_chains + chain;
            if (lane == 0) {
                rec.best_loglkl[o] = best_exact;
                rec.last_loglkl[o] = ll;
                rec.temperature[o] = st.temperature;
                rec.step_size[o] = st.step_size;
                rec.accepted[o] = nacc;
            }
#pragma unroll
            for (int r = 0; r < CPL; ++r) {
                if (rec.best_x) rec.best_x[o * NPAD + lane + 32 * r] = bx[r];
                if (rec.last_x) rec.last_x[o * NPAD + lane + 32 * r] = x[r];
            }
        }
        schedule_next(s, st, nacc, best_exact);
    }
#pragma unroll
    for (int r = 0; r < CPL; ++r) c.x[(size_t)chain * NPAD + lane + 32 * r] = x[r];
    if (lane == 0) {
        c.loglkl[chain] = ll;
        c.temperature[chain] = st.temperature;
        c.step_size[chain] = st.step_size;
        c.current_best[chain] = st.current_best;
        c.secant_prev_mult[chain] = st.prev_mult;
        c.secant_cur_mult[chain] = st.cur_mult;
        c.secant_prev_ar[chain] = st.prev_ar;
        c.secant_stage[chain] = st.stage;
        c.iteration[chain] = st.iteration;
        c.status[chain] = st.status;
        c.steps_done[chain] = t;
    }
}

// ----------------------------------------------------------------------------------------------
// Metropolis-Hastings rounds (jobtype mcmc): fixed temperature / step size per chain, chain recording.
// The incremental loglkl / gradient are re-anchored in FP64 every kResync steps.
// ----------------------------------------------------------------------------------------------
constexpr uint32_t kResync = 256;

template <int CPL, int REC>
__global__ void __launch_bounds__(32 * kWarpsPerCta, (CPL == 1) ? PB200_MIN_CTAS : 1)
mh_kernel(pb200_problem p, pb200_chains c, int n_steps, uint32_t k0, uint32_t k1, pb200_samples smp,
          int32_t* accepted_out) {
    constexpr int NPAD = 32 * CPL;
    constexpr bool LT_REGS = (CPL == 1) && PB200_LT_REGS;
    __shared__ __align__(16) double stage[kWarpsPerCta][NPAD];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int chain = blockIdx.x * kWarpsPerCta + wib;
    if (chain >= c.n_chains) return;
    const int pt = c.point[chain];
    const uint32_t chain_id = c.chain_id[chain];
    uint32_t t = c.steps_done[chain];
    const float Tf = (float)c.temperature[chain], sf = (float)c.step_size[chain];

    LaneConst<CPL> lc;
    float ltrow[LT_REGS ? 32 : 1];
    load_lane_const<CPL, LT_REGS>(p, pt, lane, lc, ltrow);
    const float* lt_pt = p.lt_t + (size_t)pt * NPAD * NPAD;

    double x[CPL], gt[CPL], bx[CPL];
#pragma unroll
    for (int r = 0; r < CPL; ++r) x[r] = c.x[(size_t)chain * NPAD + lane + 32 * r];
    double* qs = stage[wib];
    float* zs = reinterpret_cast<float*>(stage[wib]);
    double ll = exact_eval<CPL, true>(p, pt, x, lane, qs, gt);

    RecCursor rc{};
    if (REC != PB200_RECORD_NONE) {
        rc.capacity = smp.capacity;
        rc.thin = smp.thin > 0 ? smp.thin : 1;
        rc.x = smp.x + (size_t)chain * smp.capacity * NPAD;
        rc.ll = smp.loglkl + (size_t)chain * smp.capacity;
        rc.mult = smp.mult + (size_t)chain * smp.capacity;
        rc.count = smp.count[chain];
        if (REC == PB200_RECORD_ACCEPTED) {
            if (rc.count == 0) {
                record_row<CPL>(rc, x, ll, 1, lane);   // Chain([1], [loglkl(x0)], x0), mcmc.py:125
                rc.cur_mult = 1;
            } else {
                rc.cur_mult = rc.count - 1 < rc.capacity ? rc.mult[rc.count - 1] : 1;
            }
        }
    }
    double best_ll = ll;
    int32_t nacc = 0;
    const uint32_t t_end = t + (uint32_t)n_steps;
    while (t < t_end) {
        const uint32_t t_next = min(t_end, t + kResync);
        run_steps<CPL, LT_REGS, REC>(p, lt_pt, lane, k0, k1, chain_id, t, t_next, Tf, sf, ltrow, lc.lam, lc.lo, lc.hi,
                                     x, gt, ll, best_ll, bx, nacc, zs, rc);
        t = t_next;
        ll = exact_eval<CPL, true>(p, pt, x, lane, qs, gt);
    }
    if (REC != PB200_RECORD_NONE && lane == 0) {
        if (REC == PB200_RECORD_ACCEPTED && rc.count - 1 < rc.capacity) rc.mult[rc.count - 1] = rc.cur_mult;
        smp.count[chain] = rc.count;
    }
#pragma unroll
    for (int r = 0; r < CPL; ++r) c.x[(size_t)chain * NPAD + lane + 32 * r] = x[r];
    if (lane == 0) {
        c.loglkl[chain] = ll;
        c.steps_done[chain] = t;
        if (accepted_out) accepted_out[chain] = nacc;
    }
}

// ----------------------------------------------------------------------------------------------
template <int CPL>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
loglkl_kernel(pb200_problem p, const double* __restrict__ xs, const int32_t* __restrict__ point, int m,
              double* __restrict__ out, int32_t* __restrict__ outside) {
    constexpr int NPAD = 32 * CPL;
    __shared__ __align__(16) double stage[kWarpsPerCta][NPAD];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warps = gridDim.x * kWarpsPerCta;
    for (int row = blockIdx.x * kWarpsPerCta + wib; row < m; row += warps) {
        const int pt = point ? point[row] : 0;
        double x[CPL], unused[CPL];
        bool out_of_box = false;
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            const int c = lane + 32 * r;
            x[r] = xs[(size_t)row * NPAD + c];
            out_of_box = out_of_box || (x[r] < __ldg(p.lo + c)) || (x[r] > __ldg(p.hi + c));
        }
        const double ll = exact_eval<CPL, false>(p, pt, x, lane, stage[wib], unused);
        const bool any_out = __any_sync(kFull, out_of_box);
        if (lane == 0) {
            out[row] = ll;
            if (outside) outside[row] = any_out ? 1 : 0;
        }
    }
}

__global__ void schedule_kernel(pb200_chains c, pb200_schedule s, const int32_t* __restrict__ accepted,
                                const double* __restrict__ best) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n_chains || c.status[i] != PB200_ACTIVE) return;
    SchedState st{c.temperature[i], c.step_size[i], c.current_best[i], c.secant_prev_mult[i], c.secant_cur_mult[i],
                  c.secant_prev_ar[i], c.secant_stage[i], PB200_ACTIVE, c.iteration[i]};
    schedule_next(s, st, accepted[i], best[i]);
    c.temperature[i] = st.temperature;
    c.step_size[i] = st.step_size;
    c.current_best[i] = st.current_best;
    c.secant_prev_mult[i] = st.prev_mult;
    c.secant_cur_mult[i] = st.cur_mult;
    c.secant_prev_ar[i] = st.prev_ar;
    c.secant_stage[i] = st.stage;
    c.status[i] = st.status;
    c.iteration[i] = st.iteration;
}

// Running-bestfit reductions (analyse_profile_task.py:91-121) in two coalesced passes:
//   chain_min_kernel   one thread per chain: first minimum over the iterations that ran;
//   point_reduce_kernel one CTA per profile point: first-minimum repetition and the mean over the
//                       repetitions that finished, block-wide via warp shuffles + one smem hop.
__global__ void chain_min_kernel(const double* __restrict__ best, const int32_t* __restrict__ accepted, int n_iter,
                                 int b, double* __restrict__ chain_best, int32_t* __restrict__ chain_iter) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= b) return;
    double cb = INFINITY;
    int ci = -1;
    for (int k = 0; k < n_iter; ++k) {
        const size_t o = (size_t)k * b + ch;
        const double v = best[o];
        if (accepted[o] >= 0 && v < cb) { cb = v; ci = k; }
    }
    chain_best[ch] = cb;
    chain_iter[ch] = ci;
}

constexpr int kPointThreads = 256;

__global__ void __launch_bounds__(kPointThreads)
point_reduce_kernel(const double* __restrict__ chain_best, int n_points, int reps, double* __restrict__ point_best,
                    int32_t* __restrict__ point_rep, double* __restrict__ point_avg) {
    __shared__ double s_best[kPointThreads / 32], s_sum[kPointThreads / 32];
    __shared__ int s_rep[kPointThreads / 32], s_cnt[kPointThreads / 32];
    const int pt = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double my_best = INFINITY, my_sum = 0.0;
    int my_rep = 0x7fffffff, my_cnt = 0;
    for (int rep = threadIdx.x; rep < reps; rep += kPointThreads) {
        const double cb = chain_best[(size_t)rep * n_points + pt];
        if (cb < my_best) { my_best = cb; my_rep = rep; }      // ascending reps: first minimum kept
        if (cb < INFINITY) { my_sum += cb; my_cnt += 1; }
    }
    auto combine = [&](double ob, int orp, double os, int oc) {
        if (ob < my_best || (ob == my_best && orp < my_rep)) { my_best = ob; my_rep = orp; }
        my_sum += os;
        my_cnt += oc;
    };
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1)
        combine(__shfl_xor_sync(kFull, my_best, m), __shfl_xor_sync(kFull, my_rep, m),
                __shfl_xor_sync(kFull, my_sum, m), __shfl_xor_sync(kFull, my_cnt, m));
    if (lane == 0) { s_best[warp] = my_best; s_rep[warp] = my_rep; s_sum[warp] = my_sum; s_cnt[warp] = my_cnt; }
    __syncthreads();
    if (warp == 0) {
        const bool have = lane < kPointThreads / 32;
        my_best = have ? s_best[lane] : INFINITY;
        my_rep = have ? s_rep[lane] : 0x7fffffff;
        my_sum = have ? s_sum[lane] : 0.0;
        my_cnt = have ? s_cnt[lane] : 0;
#pragma unroll
        for (int m = 4; m >= 1; m >>= 1)
            combine(__shfl_xor_sync(kFull, my_best, m), __shfl_xor_sync(kFull, my_rep, m),
                    __shfl_xor_sync(kFull, my_sum, m), __shfl_xor_sync(kFull, my_cnt, m));
        if (lane == 0) {
            point_best[pt] = my_best;
            point_rep[pt] = (my_best < INFINITY) ? my_rep : 0;     // np.argmin of all-inf is 0
            point_avg[pt] = my_cnt > 0 ? my_sum / my_cnt : INFINITY;
        }
    }
}

// Weighted moments: out[0] = sum w, out[1..n] = sum w x_i, out[1+n+i*n+j] = sum w x_i x_j.
// Each CTA streams a strided set of 32-row tiles through shared memory (coalesced), every thread
// keeps kPairs (i,j) accumulators in registers, one FP64 atomicAdd per accumulator at the end.
constexpr int kMomThreads = 256, kMomTile = 32, kPairs = 4;

__global__ void __launch_bounds__(kMomThreads)
moments_kernel(const double* __restrict__ x, const int32_t* __restrict__ w, long long n_rows, int ld, int n,
               double* __restrict__ out) {
    extern __shared__ double sm[];                 // [kMomTile][n+1]: row values then the weight
    const int stride = n + 1;
    const int chunk = blockIdx.y;                  // pairs [chunk*1024, chunk*1024 + 1024)
    const int nn = n * n;
    int pi[kPairs], pj[kPairs];
    double acc[kPairs], acc1 = 0.0, acc0 = 0.0;
#pragma unroll
    for (int q = 0; q < kPairs; ++q) {
        const int pr = chunk * (kMomThreads * kPairs) + q * kMomThreads + threadIdx.x;
        pi[q] = pr < nn ? pr / n : -1;
        pj[q] = pr < nn ? pr % n : 0;
        acc[q] = 0.0;
    }
    for (long long base = (long long)blockIdx.x * kMomTile; base < n_rows; base += (long long)gridDim.x * kMomTile) {
        const int rows = (int)min((long long)kMomTile, n_rows - base);
        for (int idx = threadIdx.x; idx < rows * n; idx += kMomThreads) {
            const int r = idx / n, col = idx % n;
            sm[r * stride + col] = x[(size_t)(base + r) * ld + col];
        }
        for (int r = threadIdx.x; r < rows; r += kMomThreads) sm[r * stride + n] = w ? (double)w[base + r] : 1.0;
        __syncthreads();
        for (int r = 0; r < rows; ++r) {
            const double wr = sm[r * stride + n];
#pragma unroll
            for (int q = 0; q < kPairs; ++q)
                if (pi[q] >= 0) acc[q] = fma(wr * sm[r * stride + pi[q]], sm[r * stride + pj[q]], acc[q]);
            if (chunk == 0) {
                if (threadIdx.x < n) acc1 = fma(wr, sm[r * stride + threadIdx.x], acc1);
                acc0 += wr;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < kPairs; ++q)
        if (pi[q] >= 0) atomicAdd(out + 1 + n + pi[q] * n + pj[q], acc[q]);
    if (chunk == 0) {
        if (threadIdx.x < n) atomicAdd(out + 1 + threadIdx.x, acc1);
        if (threadIdx.x == 0) atomicAdd(out, acc0);
    }
}

__global__ void rng_kernel(uint32_t k0, uint32_t k1, uint32_t chain_id, uint32_t step0, int n_steps, int n,
                           float* __restrict__ out_z, float* __restrict__ out_e) {
    // one warp; CPL = 8 covers every n <= 256 with the production draw_block
    const int lane = threadIdx.x & 31;
    const uint32_t t_end = step0 + (uint32_t)n_steps;
    for (uint32_t blk = step0 >> 2; (blk << 2) < t_end; ++blk) {
        float z[8][4], e[4];
        draw_block<8>(k0, k1, chain_id, blk, n, lane, z, e);
        for (int k = 0; k < 4; ++k) {
            const uint32_t tk = (blk << 2) + k;
            if (tk < step0 || tk >= t_end) continue;
            for (int r = 0; r < 8; ++r) {
                const int c = lane + 32 * r;
                if (c < n) out_z[(size_t)(tk - step0) * n + c] = z[r][k];
            }
            if (lane == 0) out_e[tk - step0] = e[k];
        }
    }
}

static int validate(const pb200_problem* p) {
    if (!p) return fail("problem is NULL");
    if (p->npad != 32 && p->npad != 64 && p->npad != 128 && p->npad != 256)
        return fail("npad must be 32, 64, 128 or 256 (n <= 256)");
    if (p->n < 1 || p->n > p->npad) return fail("need 1 <= n <= npad");
    if (p->n_points < 1) return fail("n_points must be >= 1");
    if (!p->saa || !p->sab || !p->mu || !p->lo || !p->hi || !p->f || !p->c0 || !p->lt_t || !p->lam || !p->g_t || !p->h)
        return fail("problem has NULL device pointers");
    return 0;
}

}  // namespace pb200

using namespace pb200;

extern "C" {

const char* pb200_last_error(void) { return g_error.c_str(); }
int pb200_abi_version(void) { return PB200_ABI_VERSION; }

void pb200_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

int pb200_device_info(int device, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
    cudaDeviceProp prop;
    if (check_cuda(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return 1;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}

#define PB200_DISPATCH_CPL(npad, CALL)                \
    switch (npad) {                                   \
        case 32:  { constexpr int CPL = 1; CALL; } break; \
        case 64:  { constexpr int CPL = 2; CALL; } break; \
        case 128: { constexpr int CPL = 4; CALL; } break; \
        default:  { constexpr int CPL = 8; CALL; } break; \
    }

int pb200_loglkl_batch(const pb200_problem* prob, const double* x, const int32_t* point, int32_t m,
                       double* out_loglkl, int32_t* out_outside, void* stream) {
    if (validate(prob)) return 1;
    if (m <= 0) return 0;
    if (!x || !out_loglkl) return fail("loglkl_batch: NULL buffers");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (int)std::min<long long>(((long long)m + kWarpsPerCta - 1) / kWarpsPerCta, 148LL * 16);
    PB200_DISPATCH_CPL(prob->npad, (loglkl_kernel<CPL><<<blocks, 32 * kWarpsPerCta, 0, st>>>(
                                        *prob, x, point, m, out_loglkl, out_outside)));
    return check_cuda(cudaGetLastError(), "loglkl_kernel launch");
}

int pb200_anneal_run(const pb200_problem* prob, const pb200_chains* chains, const pb200_schedule* sched,
                     const pb200_records* rec, int32_t n_iter, uint64_t seed, void* stream) {
    if (validate(prob)) return 1;
    if (!chains || !sched || !rec) return fail("anneal_run: NULL argument");
    if (chains->n_chains <= 0 || n_iter <= 0) return 0;
    if (sched->steps_per_iteration <= 0) return fail("anneal_run: steps_per_iteration must be > 0");
    if (sched->step_mode < 0 || sched->step_mode > 2) return fail("anneal_run: unknown step_mode");
    if (!rec->best_loglkl || !rec->last_loglkl || !rec->temperature || !rec->step_size || !rec->accepted)
        return fail("anneal_run: records have NULL device pointers");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (chains->n_chains + kWarpsPerCta - 1) / kWarpsPerCta;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    PB200_DISPATCH_CPL(prob->npad, (anneal_kernel<CPL><<<blocks, 32 * kWarpsPerCta, 0, st>>>(
                                        *prob, *chains, *sched, *rec, n_iter, k0, k1)));
    return check_cuda(cudaGetLastError(), "anneal_kernel launch");
}

int pb200_mh_run(const pb200_problem* prob, const pb200_chains* chains, int32_t n_steps, uint64_t seed,
                 const pb200_samples* samples, int32_t* accepted_out, void* stream) {
    if (validate(prob)) return 1;
    if (!chains) return fail("mh_run: NULL chains");
    if (chains->n_chains <= 0 || n_steps <= 0) return 0;
    pb200_samples none{};
    const pb200_samples smp = samples ? *samples : none;
    if (smp.mode != PB200_RECORD_NONE) {
        if (!smp.x || !smp.loglkl || !smp.mult || !smp.count || smp.capacity <= 0)
            return fail("mh_run: sample buffers missing");
        if (smp.mode == PB200_RECORD_THINNED && smp.thin <= 0) return fail("mh_run: thin must be > 0");
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (chains->n_chains + kWarpsPerCta - 1) / kWarpsPerCta;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#define PB200_MH(REC)                                                                                          \
    PB200_DISPATCH_CPL(prob->npad, (mh_kernel<CPL, REC><<<blocks, 32 * kWarpsPerCta, 0, st>>>(                 \
                                        *prob, *chains, n_steps, k0, k1, smp, accepted_out)))
    if (smp.mode == PB200_RECORD_NONE) { PB200_MH(PB200_RECORD_NONE); }
    else if (smp.mode == PB200_RECORD_ACCEPTED) { PB200_MH(PB200_RECORD_ACCEPTED); }
    else if (smp.mode == PB200_RECORD_THINNED) { PB200_MH(PB200_RECORD_THINNED); }
    else return fail("mh_run: unknown record mode");
#undef PB200_MH
    return check_cuda(cudaGetLastError(), "mh_kernel launch");
}

int pb200_schedule_update(const pb200_chains* chains, const pb200_schedule* sched, const int32_t* accepted,
                          const double* best_loglkl, void* stream) {
    if (!chains || !sched || !accepted || !best_loglkl) return fail("schedule_update: NULL argument");
    if (chains->n_chains <= 0) return 0;
    schedule_kernel<<<(chains->n_chains + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*chains, *sched, accepted,
                                                                                       best_loglkl);
    return check_cuda(cudaGetLastError(), "schedule_kernel launch");
}

int pb200_profile_reduce(const double* best_loglkl, const int32_t* accepted, int32_t n_iter, int32_t b,
                         int32_t n_points, int32_t reps, double* out_chain_best, int32_t* out_chain_iter,
                         double* out_point_best, int32_t* out_point_rep, double* out_point_avg, void* stream) {
    if (!best_loglkl || !accepted || !out_chain_best || !out_chain_iter || !out_point_best || !out_point_rep ||
        !out_point_avg)
        return fail("profile_reduce: NULL argument");
    if (b != n_points * reps) return fail("profile_reduce: b must equal n_points * reps");
    if (n_points <= 0) return 0;
    chain_min_kernel<<<(b + 255) / 256, 256, 0, (cudaStream_t)stream>>>(best_loglkl, accepted, n_iter, b,
                                                                       out_chain_best, out_chain_iter);
    point_reduce_kernel<<<n_points, kPointThreads, 0, (cudaStream_t)stream>>>(out_chain_best, n_points, reps,
                                                                              out_point_best, out_point_rep,
                                                                              out_point_avg);
    return check_cuda(cudaGetLastError(), "profile_reduce launch");
}

int pb200_weighted_moments(const double* x, const int32_t* w, int64_t n_rows, int32_t ld, int32_t n, double* out,
                           void* stream) {
    if (!x || !out) return fail("weighted_moments: NULL argument");
    if (n < 1 || n > 256 || ld < n) return fail("weighted_moments: need 1 <= n <= 256 and ld >= n");
    cudaStream_t st = (cudaStream_t)stream;
    if (check_cuda(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)(1 + n + n * n), st), "memset")) return 1;
    if (n_rows <= 0) return 0;
    const int chunks = (n * n + kMomThreads * kPairs - 1) / (kMomThreads * kPairs);
    const long long tiles = (n_rows + kMomTile - 1) / kMomTile;
    const int gx = (int)std::min<long long>(tiles, 148LL * 4);
    const size_t smem = sizeof(double) * kMomTile * (n + 1);
    moments_kernel<<<dim3(gx, chunks), kMomThreads, smem, st>>>(x, w, (long long)n_rows, ld, n, out);
    return check_cuda(cudaGetLastError(), "moments_kernel launch");
}

int pb200_rng_variates(uint64_t seed, uint32_t chain_id, uint32_t step0, int32_t n_steps, int32_t n, float* out_z,
                       float* out_e, void* stream) {
    if (!out_z || !out_e) return fail("rng_variates: NULL argument");
    if (n < 1 || n > 256) return fail("rng_variates: need 1 <= n <= 256");
    if (n_steps <= 0) return 0;
    rng_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((uint32_t)seed, (uint32_t)(seed >> 32), chain_id, step0, n_steps,
                                                   n, out_z, out_e);
    return check_cuda(cudaGetLastError(), "rng_kernel launch");
}

}  // extern "C"
