// pb200_kernels.cu -- sm_100a kernels + the C ABI of include/pb200.h.
//
// Kernels (DESIGN.md has the roofline of each):
//   anneal_kernel<LPC,CPL>    fused annealing iterations: Philox normals -> whitened quadratic form ->
//                             accept/reject at the chain temperature -> (on accept) proposal matvec, box
//                             test, state update, running bestfit -> set_bestfit + next settings.
//   anneal_ws_kernel<LPC,CPL> the same with a producer warp per chain warp (variates through a shared-memory ring)
//                             for batches too small to fill the schedulers.  Both can signal finished iterations
//                             and perform the iteration-boundary all-gather themselves (peer-memory stores).
//   mh_kernel<LPC,CPL,REC>    plain Metropolis-Hastings rounds with chain recording (jobtype mcmc).
//   loglkl_kernel<CPL>        FP64 chi^2 quadratic form + prior box for a batch of positions.
//   schedule_kernel           stand-alone adaptive step-size bookkeeping.
//   chain_min_kernel + point_reduce_kernel   min over iterations / first-min repetition / mean (warp shuffles).
//   moments_kernel            weighted first/second moments (covariance reduction), FP64.
//   chain_moments_kernel      per-chain weighted first moments of recorded chains (Gelman-Rubin statistics).
//   wide_* (pb200_wide.cuh, pb200_umma.cuh)  the per-iteration pipeline of wide problems (npad >= 128): steps ->
//                             tcgen05 position GEMM -> FP64 anchor GEMM.
//   hosted_* (pb200_hosted.cuh)  batched proposal / accept for likelihoods evaluated on the host.
//   rng_kernel                test hook exposing the variates.
// Chain state lives in HBM between launches and in registers inside one; there is no CPU fallback.
#include "pb200_device.cuh"
#include "pb200_hosted.cuh"

// Minimum resident CTAs per SM (-> register budget 65536 / (CTAs * 128)) of the npad = 32 geometries; the values are
// the measured optimum on B200 (profiles/README.md).  Tried and rejected: drawing the next Philox block piecewise
// in front of each step (software pipelining) -- the extra live registers cost more than the overlap gained.
#ifndef PB200_MIN_CTAS
#define PB200_MIN_CTAS 7     // 32 lanes per chain: <= 72 registers, 28 warps per SM
#endif
#ifndef PB200_MIN_CTAS_16
#define PB200_MIN_CTAS_16 5  // 16 lanes per chain: <= 96 registers
#endif
#ifndef PB200_MIN_CTAS_8
#define PB200_MIN_CTAS_8 4   // 8 lanes per chain: <= 128 registers
#endif
#ifndef PB200_PIPE_RNG
#define PB200_PIPE_RNG 0     // experiment: draw the next Philox block in the shadow of the current one (8-lane geometry)
#endif

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

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

template <int LPC, int CPL>
struct Geometry {
    static constexpr int NPAD = LPC * CPL, G = 32 / LPC;
    // four-steps-per-reduction (block_step): one reduce-scatter of 16 sums (30-32 shuffles) instead of four dependent
    // reduce -> compare -> branch chains.  Measured on B200 (profiles/README.md): with 8 lanes per chain it wins at every
    // batch size (+14 % at 4096 chains); with 16 / 32 lanes it wins where the batch is latency-bound -- which is exactly
    // where choose_lanes picks those geometries for npad = 32 (<= 2560 chains: +6 % .. +13 % over the 8-lane kernel with
    // producer warps) -- and loses 10-25 % on large batches (more shuffle instructions per chain-step, issue-bound),
    // where choose_lanes does not pick them.  The wide geometries (32 lanes, >= 4 coordinates per lane: npad >= 128)
    // are latency-bound at every batch size that fits the register file (one chain per warp, <= 8 warps per SM).
    // npad = 64 keeps the step-by-step form (not measured).
    static constexpr bool kBlockMode = (NPAD == 32) || (LPC == 32 && CPL >= 4);
    // register budget per geometry (65536 / (kMinCtas * 128 threads)): 72 / 96 / 128 for 1 / 2 / 4 coordinates per lane
    static constexpr int kMinCtas =
        (NPAD != 32 ? 1 : (LPC == 32 ? PB200_MIN_CTAS : (LPC == 16 ? PB200_MIN_CTAS_16 : PB200_MIN_CTAS_8))) * 4 / kWarpsPerCta;
};

// Per-group recording cursor of mh_kernel.
struct RecCursor {
    double* x;        // this chain's [capacity][npad]
    double* ll;       // [capacity]
    int32_t* mult;    // [capacity]
    int32_t capacity, thin, count, cur_mult;
};

template <int LPC, int CPL>
__device__ __forceinline__ void record_row(RecCursor& rc, const double (&x)[CPL], double ll, int32_t mult, int gl) {
    constexpr int NPAD = LPC * CPL;
    if (rc.count < rc.capacity) {
        double* row = rc.x + (size_t)rc.count * NPAD;
#pragma unroll
        for (int r = 0; r < CPL; ++r) row[gl + LPC * r] = x[r];
        if (gl == 0) {
            rc.ll[rc.count] = ll;
            rc.mult[rc.count] = mult;
        }
    }
    rc.count += 1;   // count > capacity tells the host that rows were dropped
}

// Registers of one chain, spread over the LPC lanes of its group.
//   position = xref + Lt * w   (w: whitened displacement accumulated since xref was last materialised)
//   best     = (best_rel ? xref : bref) + Lt * bw      (bref is only written when xref moves away from under a
//                                                        best point, so a new best costs CPL moves, not 3*CPL)
template <int CPL>
struct ChainState {
    double xref[CPL], bref[CPL];
    float w[CPL], bw[CPL], gf[CPL], hl[CPL];   // gf: whitened gradient, FP32 between two FP64 anchors
    double ll, best_ll;
    float m2;             // squared whitened radius around xref that provably stays inside the prior box
    int32_t nacc;
    bool best_rel;        // the best point is expressed relative to the CURRENT xref
};

// ----------------------------------------------------------------------------------------------
// One MH step of every chain of the warp (prospect/mcmc.py:163-190, batched and re-ordered):
//   Delta  = step * sum_i z_i (gt_i + 1/2 step lam_i z_i)            O(n), FP32, every step
//   accept = Delta < T * e  and  proposal inside the prior box
// The reference draws the uniform even for out-of-box proposals and never evaluates the likelihood there;
// "likelihood-accept AND inside" is the same predicate evaluated in the cheaper order.  The box test is
// DEFERRED: a likelihood-accepted move only updates the whitened displacement w (O(n)); as long as
// |w|^2 <= m2 the position provably lies inside the box (safe_radius2), otherwise the position is
// materialised (n x n matvec), tested exactly and becomes the new reference point.
// Predicates are per group; everything containing a shuffle or __syncwarp runs under warp-uniform control.
// ----------------------------------------------------------------------------------------------
// The exact prior-box test of a proposal: materialise it, compare with the box, and (deferred mode) the safe radius
// around the point that will be the reference afterwards.  In the deferred mode this is the RARE path of a step (only
// when |w|^2 leaves the safe radius), so it is kept out of line: inlined into each of the four unrolled steps of a
// Philox block it made the hot loop several times larger than the instruction cache (ncu: `no_instruction` stalls of
// 0.5 per issue in the 256-D kernel).  Arguments and results travel by value, so the caller's arrays stay in registers.
template <int CPL>
struct BoxIn { double xref[CPL]; float wt[CPL]; };
template <int CPL>
struct BoxProbe { double xt[CPL]; float m2n; bool outside; };

template <int LPC, int CPL, bool WITH_RADIUS>
__device__ __forceinline__ BoxProbe<CPL> probe_box_body(const pb200_problem& p, int pt, const float* __restrict__ lt_pt,
                                                        int gl, int gi, bool exact, const BoxIn<CPL>& in, float* ws) {
    BoxProbe<CPL> o;
    if constexpr (LPC < 32 && LPC * CPL == 32)      // several chains per warp: only the chains that ask pay
        materialise_selected<LPC, CPL>(lt_pt, gl, gi, exact, in.xref, in.wt, ws, o.xt);
    else
        materialise<LPC, CPL>(lt_pt, gl, in.xref, in.wt, ws, o.xt);
    bool outside = false;
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        const int col = gl + LPC * r;
        outside = outside || (o.xt[r] < __ldg(p.lo + col)) || (o.xt[r] > __ldg(p.hi + col));
    }
    o.outside = group_any<LPC>(outside, gi);
    o.m2n = 0.0f;
    if (WITH_RADIUS) {                         // the safe radius around the reference point after this step
        const bool reref = exact && !o.outside;
        double ref[CPL];
#pragma unroll
        for (int r = 0; r < CPL; ++r) ref[r] = reref ? o.xt[r] : in.xref[r];
        o.m2n = safe_radius2<LPC, CPL>(p, pt, gl, ref);
    }
    return o;
}
template <int LPC, int CPL>
__device__ __noinline__ BoxProbe<CPL> probe_box_rare(const pb200_problem& p, int pt, const float* __restrict__ lt_pt,
                                                     int gl, int gi, bool exact, BoxIn<CPL> in, float* ws) {
    return probe_box_body<LPC, CPL, true>(p, pt, lt_pt, gl, gi, exact, in, ws);
}

// ----------------------------------------------------------------------------------------------
// Tensor-core box filter of the MH kernels (8 lanes per chain, npad = 32: four chains per warp).  A Philox block has up
// to 4 x 4 = 16 candidate positions per warp (chain x step); their offsets from the reference points are ONE
// [16 x 32] . [32 x 32] product, 16 `mma.sync.m16n8k8` TF32 instructions, against ~100 instructions per candidate for the
// FP32 mat-vec.  TF32 is not exact, so the product only FILTERS: a candidate whose every coordinate stays inside the box
// by more than a rigorous error margin is provably inside for the exact FP32 test as well; the rest (one in ~10^3 on
// the toy) go through the exact test.  Decisions therefore never depend on the tensor-core arithmetic -- the oracle
// knows nothing of this filter and stays bit-identical.
//   margin: operands truncated to TF32 (2^-10 relative each) => |d~ - d| <= 2^-9 sum_k |w_k Lt_kc| <= 2^-9 |w| |Lt_c|
//   (Cauchy-Schwarz; |Lt_c| = lt_norm, |w| <= sqrt(2 (R0 + 4 s^2 E)), the block's own bound); 2^-8 is used.  The FP32
//   roundings of the reference point, the box centre and the accumulation (< 2^-20 max(|lo|, |hi|) in all) are taken
//   off the half-width once, at set-up (2^-19).
// ----------------------------------------------------------------------------------------------
struct __align__(16) BoxFilter {
    float wc[32][24];       // candidate displacements [coordinate][4 * chain + step] (24: conflict-free fragment loads)
    float xm[4][32];        // FP32 (reference point - box centre) of the warp's four chains: the accumulators' start
    float2 bfrag[16][32];   // Lt of the warp's point in B-fragment order: [4 kt + nt][lane] = (b0, b1)
    float mid[32], half[32], ltn[32];     // box centre, shrunk half-width, 2^-8 |Lt_c| per coordinate
    int ok;                 // all chains of the warp share one point and every box side pair is finite or absent
};

__device__ __forceinline__ void box_filter_setup(BoxFilter& f, const pb200_problem& p, int pt, int lane) {
    const int pt0 = __shfl_sync(kFull, pt, 0);
    bool good = pt == pt0;
    float mid = 0.0f, half = INFINITY;
    if (lane < p.n) {
        const double lo = __ldg(p.lo + lane), hi = __ldg(p.hi + lane);
        if (isfinite(lo) && isfinite(hi)) {
            mid = (float)(0.5 * (lo + hi));
            half = __double2float_rd(0.5 * (hi - lo) - fmax(fabs(lo), fabs(hi)) * 0x1p-19);
        } else if (!(isinf(lo) && isinf(hi))) {
            good = false;                       // one-sided box: no bound on |x| for the rounding analysis
        }
    }
    f.mid[lane] = mid;
    f.half[lane] = half;
    f.ltn[lane] = __fmul_ru(__ldg(p.lt_norm + (size_t)pt0 * 32 + lane), 0x1p-8f);
    const float* __restrict__ lt = p.lt_t + (size_t)pt0 * 32 * 32;
    const int q = lane & 3, row = lane >> 2;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
            f.bfrag[kt * 4 + nt][lane] = make_float2(__ldg(lt + (kt * 8 + q) * 32 + nt * 8 + row),
                                                     __ldg(lt + (kt * 8 + q + 4) * 32 + nt * 8 + row));
    }
    good = __all_sync(kFull, good);
    if (lane == 0) f.ok = good ? 1 : 0;
    __syncwarp();
}

// `wanted`: this group's candidate k is to be tested (bit k).  Returns true when some wanted candidate of the group
// is not provably inside.  Called by the whole warp after every lane has staged its part of f.wc / f.xm; out of line
// so that the 16 accumulators do not join the block's register budget (inlined: 148 bytes of spills in the hot loop).
// The accumulators start at (reference point - box centre), so the product comes out as the signed distance from the
// centre and a coordinate costs one FFMA + one compare.
__device__ __noinline__ bool box_filter_mma(BoxFilter& f, int lane, float wmax, unsigned wanted) {
    const int gi = lane >> 3;
    // one word for the warp: bits [4 g, 4 g + 4) = wanted candidates of chain g = rows of the A operand
    const unsigned w0 = __shfl_sync(kFull, wanted, 0), w1 = __shfl_sync(kFull, wanted, 8),
                   w2 = __shfl_sync(kFull, wanted, 16), w3 = __shfl_sync(kFull, wanted, 24);
    const unsigned cand = w0 | (w1 << 4) | (w2 << 8) | (w3 << 12);
    __syncwarp();
    // rows `row` / `row + 8` of the product: candidate (chain, step) = (row >> 2, row & 3) / (2 + (row >> 2), row & 3)
    const int q = lane & 3, row = lane >> 2, c_lo = lane >> 4, c_hi = 2 + (lane >> 4);
    float d[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const float2 lo2 = *reinterpret_cast<const float2*>(&f.xm[c_lo][nt * 8 + 2 * q]);
        const float2 hi2 = *reinterpret_cast<const float2*>(&f.xm[c_hi][nt * 8 + 2 * q]);
        d[nt][0] = lo2.x; d[nt][1] = lo2.y; d[nt][2] = hi2.x; d[nt][3] = hi2.y;
    }
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
        const uint32_t a0 = __float_as_uint(f.wc[kt * 8 + q][row]), a1 = __float_as_uint(f.wc[kt * 8 + q][row + 8]),
                       a2 = __float_as_uint(f.wc[kt * 8 + q + 4][row]), a3 = __float_as_uint(f.wc[kt * 8 + q + 4][row + 8]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const float2 b = f.bfrag[kt * 4 + nt][lane];
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                         "{%0,%1,%2,%3};"
                         : "+f"(d[nt][0]), "+f"(d[nt][1]), "+f"(d[nt][2]), "+f"(d[nt][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)));
        }
    }
    const float wm_lo = __shfl_sync(kFull, wmax, 8 * c_lo), wm_hi = __shfl_sync(kFull, wmax, 8 * c_hi);
    bool ok_lo = true, ok_hi = true;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const float2 half = *reinterpret_cast<const float2*>(&f.half[nt * 8 + 2 * q]);
        const float2 ltn = *reinterpret_cast<const float2*>(&f.ltn[nt * 8 + 2 * q]);
        ok_lo = ok_lo && (__fmaf_rn(ltn.x, wm_lo, fabsf(d[nt][0])) < half.x) &&
                (__fmaf_rn(ltn.y, wm_lo, fabsf(d[nt][1])) < half.y);
        ok_hi = ok_hi && (__fmaf_rn(ltn.x, wm_hi, fabsf(d[nt][2])) < half.x) &&
                (__fmaf_rn(ltn.y, wm_hi, fabsf(d[nt][3])) < half.y);
    }
    const bool bad_lo = ((cand >> row) & 1u) != 0u && !ok_lo, bad_hi = ((cand >> (row + 8)) & 1u) != 0u && !ok_hi;
    const unsigned blo = __ballot_sync(kFull, bad_lo), bhi = __ballot_sync(kFull, bad_hi);
    __syncwarp();                                   // the staging rows are rewritten by the next block
    const unsigned mine = gi < 2 ? (blo >> (16 * gi)) : (bhi >> (16 * (gi - 2)));
    return (mine & 0xFFFFu) != 0u;
}

template <int LPC, int CPL>
__device__ __noinline__ bool probe_box_outside(const pb200_problem& p, int pt, const float* __restrict__ lt_pt, int gl,
                                               int gi, bool want, BoxIn<CPL> in, float* ws) {
    return probe_box_body<LPC, CPL, false>(p, pt, lt_pt, gl, gi, want, in, ws).outside;
}

// The thinned record of a step (one step in `thin`): materialise the position and append the row.  Out of line: four
// inlined copies per Philox block made the hot loop of mh_kernel<.., RECORD_THINNED> miss the instruction cache (ncu:
// `no_instruction` 1.56 stalls per issue, issue slots 55 % busy against 64 % without recording).
template <int LPC, int CPL>
__device__ __noinline__ int32_t record_thinned(const float* __restrict__ lt_pt, int gl, BoxIn<CPL> in, float* ws,
                                               RecCursor rc, double ll, bool now) {
    double xn[CPL];
    materialise<LPC, CPL>(lt_pt, gl, in.xref, in.wt, ws, xn);
    if (now) record_row<LPC, CPL>(rc, xn, ll, 1, gl);
    return rc.count;
}

template <int LPC, int CPL, int REC, bool EXACT = (REC == PB200_RECORD_ACCEPTED)>
__device__ __forceinline__ void one_step(const pb200_problem& p, int pt, const float* __restrict__ lt_pt, int gl, int gi,
                                         bool live, uint32_t tk, float Tf, float sf, const float (&zk)[CPL], float ek,
                                         ChainState<CPL>& c, float* ws, RecCursor& rc) {
    float tt = 0.0f;
#pragma unroll
    for (int r = 0; r < CPL; ++r) tt = __fmaf_rn(zk[r], __fmaf_rn(c.hl[r], zk[r], c.gf[r]), tt);
    tt = group_sum<LPC>(tt);
    const float delta = __fmul_rn(sf, tt);
    const bool lik = live && (delta < __fmul_rn(Tf, ek));
    bool acc = false;
    if (__any_sync(kFull, lik)) {
        float wt[CPL], rr = 0.0f;
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            wt[r] = __fmaf_rn(sf, zk[r], c.w[r]);
            rr = __fmaf_rn(wt[r], wt[r], rr);
        }
        if (!EXACT) rr = group_sum<LPC>(rr);      // EXACT: every accepted step is materialised, no radius test
        const bool exact = lik && (EXACT || !(rr <= c.m2));
        acc = lik;
        if (__any_sync(kFull, exact)) {
            BoxIn<CPL> in;
#pragma unroll
            for (int r = 0; r < CPL; ++r) { in.xref[r] = c.xref[r]; in.wt[r] = wt[r]; }
            const BoxProbe<CPL> probe = EXACT ? probe_box_body<LPC, CPL, false>(p, pt, lt_pt, gl, gi, exact, in, ws)
                                              : probe_box_rare<LPC, CPL>(p, pt, lt_pt, gl, gi, exact, in, ws);
            const bool reref = exact && !probe.outside;
            acc = lik && !(exact && probe.outside);
            if (reref) {
#pragma unroll
                for (int r = 0; r < CPL; ++r) {
                    if (c.best_rel) c.bref[r] = c.xref[r];      // the best point stays where it was
                    c.xref[r] = probe.xt[r];
                    wt[r] = 0.0f;
                }
                c.best_rel = false;
                if (!EXACT) c.m2 = probe.m2n;
            }
        }
        if (acc) {
#pragma unroll
            for (int r = 0; r < CPL; ++r) {
                c.w[r] = wt[r];
                c.gf[r] = __fmaf_rn(__fadd_rn(c.hl[r], c.hl[r]), zk[r], c.gf[r]);     // gt += step * lam * z
            }
            c.ll = __dadd_rn(c.ll, (double)delta);
            c.nacc += 1;
            if (c.ll < c.best_ll) {
                c.best_ll = c.ll;
                c.best_rel = true;
#pragma unroll
                for (int r = 0; r < CPL; ++r) c.bw[r] = c.w[r];
            }
        }
    }
    if (REC == PB200_RECORD_ACCEPTED) {
        if (live) {
            if (acc) {       // exact path always re-references: the new point is xref itself
                if (gl == 0 && rc.count - 1 < rc.capacity) rc.mult[rc.count - 1] = rc.cur_mult;
                rc.cur_mult = 1;
                record_row<LPC, CPL>(rc, c.xref, c.ll, 1, gl);
            } else {
                rc.cur_mult += 1;
            }
        }
    } else if (REC == PB200_RECORD_THINNED) {
        const bool now = live && ((tk + 1) % (uint32_t)rc.thin == 0);
        if (__any_sync(kFull, now)) {
            BoxIn<CPL> in;
#pragma unroll
            for (int r = 0; r < CPL; ++r) { in.xref[r] = c.xref[r]; in.wt[r] = c.w[r]; }
            rc.count = record_thinned<LPC, CPL>(lt_pt, gl, in, ws, rc, c.ll, now);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// Four steps at once.  A chain is sequential in its steps, but inside one Philox block everything a step needs
// from the state is linear in what the earlier steps of the block did: with the block-start gradient g,
//   A_k = z_k . g,   B_k = sum_i hl_i z_ki^2,   C_jk = sum_i hl_i z_ji z_ki  (j < k)
//   tt_k = A_k + B_k + 2 * sum_{accepted j < k} C_jk,       Delta_k = step * tt_k,
// so the 14 dot products (+ |w|^2 and sum_k |z_k|^2 for the box bound) are reduced TOGETHER -- one butterfly with
// 16 independent values in flight instead of four dependent reduce -> compare -> branch chains -- and the four
// accept decisions are then resolved with scalar arithmetic.  This is what keeps a small batch (few warps per
// SM) from being latency-bound.  Box safety for every intermediate point: |w_k| <= |w_0| + step * sum_j |z_j|
// <= sqrt(R0) + 2 step sqrt(E), hence |w_k|^2 <= 2 (R0 + 4 step^2 E).  Returns true, with nothing committed, when
// the group must replay the block step by step (bound not provably inside the safe radius).
//
// BOXTEST (the MH kernels): at the reference's wide step 2.4 / sqrt(d) the Cauchy-Schwarz radius is no use -- in 30-D
// it is ~5x tighter than any single face, one accepted step (|s z| ~ 2.4) already leaves it -- so every block was
// replayed step by step, each accepted step paying a mat-vec, a fresh FP64 safe radius and a re-reference (ncu / probe:
// 55 % of the K5 kernel).  With BOXTEST a block whose bound fails tests the positions it would accept instead: one
// FP32 mat-vec per likelihood-accepted step (the state the chain would hold after it, formed by the commit loop's own
// FMAs), compared with the box; all inside -> the block commits as it stands (no re-reference: w simply keeps
// growing until the next FP64 anchor), one outside -> replay.  Same predicate as the reference's (inside the box AND
// likelihood-accept), evaluated on speculative accepts.
// ----------------------------------------------------------------------------------------------
template <int LPC, int CPL, bool BOXTEST = false>
__device__ __forceinline__ bool block_step(int gl, bool live, float Tf, float sf, const float (&z)[CPL][4],
                                           const float (&e)[4], ChainState<CPL>& c, const pb200_problem* p = nullptr,
                                           int pt = 0, const float* __restrict__ lt_pt = nullptr, int gi = 0,
                                           float* ws = nullptr, BoxFilter* bf = nullptr) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.0f;
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        const float g = c.gf[r], h = c.hl[r];
        float hz[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hz[k] = __fmul_rn(h, z[r][k]);
            v[k] = __fmaf_rn(z[r][k], g, v[k]);                  // A_k
            v[4 + k] = __fmaf_rn(hz[k], z[r][k], v[4 + k]);      // B_k
            v[14] = __fmaf_rn(z[r][k], z[r][k], v[14]);          // E
        }
        v[8] = __fmaf_rn(hz[0], z[r][1], v[8]);                  // C01
        v[9] = __fmaf_rn(hz[0], z[r][2], v[9]);                  // C02
        v[10] = __fmaf_rn(hz[0], z[r][3], v[10]);                // C03
        v[11] = __fmaf_rn(hz[1], z[r][2], v[11]);                // C12
        v[12] = __fmaf_rn(hz[1], z[r][3], v[12]);                // C13
        v[13] = __fmaf_rn(hz[2], z[r][3], v[13]);                // C23
        v[15] = __fmaf_rn(c.w[r], c.w[r], v[15]);                // R0 = |w|^2
    }
    if constexpr (LPC == 8) {
        // reduce-scatter + broadcast instead of 16 full butterflies: at every level a lane keeps the half of the
        // values selected by its lane bit and ships the other half (8 + 4 + 2 exchanges), then each of the 16 sums
        // is broadcast from the lane that owns it: 30 SHFL + 14 FADD instead of 48 + 48.  Every sum is formed by the
        // same additions in the same order as the plain xor butterfly (own + partner, distances 4, 2, 1).
        const bool b4 = (gl & 4) != 0, b2 = (gl & 2) != 0, b1 = (gl & 1) != 0;
        float a8[8], a4[4], a2[2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            a8[i] = __fadd_rn(b4 ? v[i + 8] : v[i], __shfl_xor_sync(kFull, b4 ? v[i] : v[i + 8], 4));
#pragma unroll
        for (int i = 0; i < 4; ++i)
            a4[i] = __fadd_rn(b2 ? a8[i + 4] : a8[i], __shfl_xor_sync(kFull, b2 ? a8[i] : a8[i + 4], 2));
#pragma unroll
        for (int i = 0; i < 2; ++i)
            a2[i] = __fadd_rn(b1 ? a4[i + 2] : a4[i], __shfl_xor_sync(kFull, b1 ? a4[i] : a4[i + 2], 1));
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __shfl_sync(kFull, a2[j & 1], j >> 1, 8);
    } else if constexpr (LPC == 16) {
        // four levels (8 + 4 + 2 + 1 exchanges): the lane bits select the index bits, so sum j ends up in lane j of the
        // group; one broadcast per sum -- 31 SHFL + 15 FADD
        const bool b8 = (gl & 8) != 0, b4 = (gl & 4) != 0, b2 = (gl & 2) != 0, b1 = (gl & 1) != 0;
        float a8[8], a4[4], a2[2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            a8[i] = __fadd_rn(b8 ? v[i + 8] : v[i], __shfl_xor_sync(kFull, b8 ? v[i] : v[i + 8], 8));
#pragma unroll
        for (int i = 0; i < 4; ++i)
            a4[i] = __fadd_rn(b4 ? a8[i + 4] : a8[i], __shfl_xor_sync(kFull, b4 ? a8[i] : a8[i + 4], 4));
#pragma unroll
        for (int i = 0; i < 2; ++i)
            a2[i] = __fadd_rn(b2 ? a4[i + 2] : a4[i], __shfl_xor_sync(kFull, b2 ? a4[i] : a4[i + 2], 2));
        const float a1 = __fadd_rn(b1 ? a2[1] : a2[0], __shfl_xor_sync(kFull, b1 ? a2[0] : a2[1], 1));
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __shfl_sync(kFull, a1, j, 16);
    } else if constexpr (LPC == 32) {
        // the same reduce-scatter over five levels (8 + 4 + 2 + 1 exchanges, then one full exchange): sum j ends up in
        // lanes 2 j and 2 j + 1, one broadcast per sum -- 32 SHFL + 16 FADD, each sum formed by the butterfly's own
        // additions (own + partner at distances 16, 8, 4, 2, 1)
        const bool b16 = (gl & 16) != 0, b8 = (gl & 8) != 0, b4 = (gl & 4) != 0, b2 = (gl & 2) != 0;
        float a8[8], a4[4], a2[2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            a8[i] = __fadd_rn(b16 ? v[i + 8] : v[i], __shfl_xor_sync(kFull, b16 ? v[i] : v[i + 8], 16));
#pragma unroll
        for (int i = 0; i < 4; ++i)
            a4[i] = __fadd_rn(b8 ? a8[i + 4] : a8[i], __shfl_xor_sync(kFull, b8 ? a8[i] : a8[i + 4], 8));
#pragma unroll
        for (int i = 0; i < 2; ++i)
            a2[i] = __fadd_rn(b4 ? a4[i + 2] : a4[i], __shfl_xor_sync(kFull, b4 ? a4[i] : a4[i + 2], 4));
        float a1 = __fadd_rn(b2 ? a2[1] : a2[0], __shfl_xor_sync(kFull, b2 ? a2[0] : a2[1], 2));
        a1 = __fadd_rn(a1, __shfl_xor_sync(kFull, a1, 1));
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __shfl_sync(kFull, a1, 2 * j);
    } else {
#pragma unroll
        for (int m = LPC / 2; m >= 1; m >>= 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __fadd_rn(v[i], __shfl_xor_sync(kFull, v[i], m));
        }
    }
    float delta[4];
    bool acc[4];
    delta[0] = __fmul_rn(sf, __fadd_rn(v[0], v[4]));
    acc[0] = live && (delta[0] < __fmul_rn(Tf, e[0]));
    const float x1 = acc[0] ? __fadd_rn(v[8], v[8]) : 0.0f;
    delta[1] = __fmul_rn(sf, __fadd_rn(__fadd_rn(v[1], v[5]), x1));
    acc[1] = live && (delta[1] < __fmul_rn(Tf, e[1]));
    const float x2 = __fadd_rn(acc[0] ? __fadd_rn(v[9], v[9]) : 0.0f, acc[1] ? __fadd_rn(v[11], v[11]) : 0.0f);
    delta[2] = __fmul_rn(sf, __fadd_rn(__fadd_rn(v[2], v[6]), x2));
    acc[2] = live && (delta[2] < __fmul_rn(Tf, e[2]));
    const float x3 = __fadd_rn(__fadd_rn(acc[0] ? __fadd_rn(v[10], v[10]) : 0.0f,
                                         acc[1] ? __fadd_rn(v[12], v[12]) : 0.0f),
                               acc[2] ? __fadd_rn(v[13], v[13]) : 0.0f);
    delta[3] = __fmul_rn(sf, __fadd_rn(__fadd_rn(v[3], v[7]), x3));
    acc[3] = live && (delta[3] < __fmul_rn(Tf, e[3]));
    const bool any = acc[0] || acc[1] || acc[2] || acc[3];
    // (sqrt(R0) + 2 s sqrt(E))^2 <= 2 (R0 + 4 s^2 E): no square roots (each would be a MUFU + a slow-path branch)
    const float s2 = __fadd_rn(sf, sf);
    const float half_bnd2 = __fmaf_rn(__fmul_rn(s2, s2), v[14], v[15]);
    const bool unsafe = any && !(__fadd_rn(half_bnd2, half_bnd2) <= c.m2);
    if constexpr (!BOXTEST) {
        if (unsafe) return true;
    } else {
        if (__any_sync(kFull, unsafe)) {
            bool open = unsafe;                     // some accepted position of the group is not yet known to be inside
            if constexpr (LPC == 8 && CPL == 4) {
                if (bf != nullptr && bf->ok) {      // (warp-uniform)
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        float w = c.w[r], wk[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            w = acc[k] ? __fmaf_rn(sf, z[r][k], w) : w;
                            wk[k] = w;
                        }
                        *reinterpret_cast<float4*>(&bf->wc[gl + 8 * r][4 * gi]) = make_float4(wk[0], wk[1], wk[2], wk[3]);
                        bf->xm[gi][gl + 8 * r] = __fsub_rn((float)c.xref[r], bf->mid[gl + 8 * r]);
                    }
                    const float wmax = __fmul_ru(__fsqrt_ru(__fadd_ru(half_bnd2, half_bnd2)), 1.0001f);
                    const unsigned wanted = unsafe ? ((acc[0] ? 1u : 0u) | (acc[1] ? 2u : 0u) | (acc[2] ? 4u : 0u) |
                                                      (acc[3] ? 8u : 0u)) : 0u;
                    open = box_filter_mma(*bf, gi * 8 + gl, wmax, wanted);
                }
            }
            if (__any_sync(kFull, open)) {          // the exact FP32 test of what the filter could not settle
                BoxIn<CPL> in;
#pragma unroll
                for (int r = 0; r < CPL; ++r) { in.xref[r] = c.xref[r]; in.wt[r] = c.w[r]; }
                bool bad = false;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int r = 0; r < CPL; ++r) in.wt[r] = acc[k] ? __fmaf_rn(sf, z[r][k], in.wt[r]) : in.wt[r];
                    const bool want = open && acc[k];
                    if (__any_sync(kFull, want))
                        bad = (probe_box_outside<LPC, CPL>(*p, pt, lt_pt, gl, gi, want, in, ws) && want) || bad;
                }
                if (bad) return true;
            }
        }
    }
    // scalar pass, branch-free: running loglkl (+0.0 for a rejected step), accept count, and which step (if any)
    // sets a new best
    int kb = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c.ll = __dadd_rn(c.ll, acc[k] ? (double)delta[k] : 0.0);
        c.nacc += acc[k] ? 1 : 0;
        const bool better = acc[k] && (c.ll < c.best_ll);
        c.best_ll = better ? c.ll : c.best_ll;
        kb = better ? k : kb;
    }
    c.best_rel = c.best_rel || (kb >= 0);
    // vector pass: predicated, branch-free
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        const float h2 = __fadd_rn(c.hl[r], c.hl[r]);
        float w = c.w[r], g = c.gf[r], b = c.bw[r];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            w = acc[k] ? __fmaf_rn(sf, z[r][k], w) : w;
            g = acc[k] ? __fmaf_rn(h2, z[r][k], g) : g;
            b = (kb == k) ? w : b;
        }
        c.w[r] = w; c.gf[r] = g; c.bw[r] = b;
    }
    return false;
}

// MH steps [t_begin, t_begin + n_steps) of the chains of this warp whose step counter is congruent to `lead`
// modulo 4 (`live` is false for the others).  Whole Philox blocks go through block_step, the ragged head / tail of
// the range and blocks that need the exact box test through one_step.  Which path a step takes depends only on
// the chain's own step index and state, never on its warp neighbours.
// Where the variates of a Philox block come from.
constexpr int kSrcPhilox = 0;   // drawn in place by the warp that uses them
constexpr int kSrcRing = 2;     // produced by the partner warp into a shared-memory ring (anneal_ws_kernel)

template <int LPC, int CPL, int REC, int SRC, bool EXACT, bool BOXTEST>
__device__ __forceinline__ void run_steps_class(const pb200_problem& p, int pt, const float* __restrict__ lt_pt, int gl,
                                                int gi, bool live, uint32_t k0, uint32_t k1, uint32_t chain_id,
                                                uint32_t t_begin, uint32_t lead, uint32_t n_steps, float Tf, float sf,
                                                ChainState<CPL>& c, float* ws, RecCursor& rc, Ring& ring,
                                                BoxFilter* bf = nullptr) {
    float z[CPL][4], e[4], zk[CPL];
    uint32_t i = 0;                             // steps done (warp-uniform); this group's step index is t_begin + i
    uint32_t tmod = 0;                          // (t_begin + i) % thin of this group (thinned recording)
    if constexpr (REC == PB200_RECORD_THINNED) tmod = t_begin % (uint32_t)rc.thin;
#if PB200_PIPE_RNG
    // software pipeline (small batches of the 8-lane geometry, where registers do not limit occupancy): the variates of
    // block b + 1 -- independent of the chain state -- are drawn in the shadow of block b's dependent reduce ->
    // compare -> commit chain
    constexpr bool kPipe = SRC == kSrcPhilox && LPC == 8 && REC == PB200_RECORD_NONE && !EXACT;
    float zn[kPipe ? CPL : 1][4], en[4];
    if constexpr (kPipe) draw_block<LPC, CPL>(k0, k1, chain_id, t_begin >> 2, p.n, gl, zn, en);
#else
    constexpr bool kPipe = false;
#endif
    while (i < n_steps) {
        const uint32_t t = t_begin + i;
        if constexpr (SRC == kSrcRing) {
            ring_fetch<LPC, CPL>(ring, gi * LPC + gl, gi, z, e);
        } else if constexpr (kPipe) {
#if PB200_PIPE_RNG
#pragma unroll
            for (int r = 0; r < CPL; ++r) {
#pragma unroll
                for (int k = 0; k < 4; ++k) z[r][k] = zn[r][k];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) e[k] = en[k];
            draw_block<LPC, CPL>(k0, k1, chain_id, (t >> 2) + 1u, p.n, gl, zn, en);     // next block (unused after the last)
#endif
        } else {
            draw_block<LPC, CPL>(k0, k1, chain_id, t >> 2, p.n, gl, z, e);
        }
        const uint32_t off = (lead + i) & 3u, cnt = min(4u - off, n_steps - i);
        bool todo = live, cnt_done = false;
        // thinned recording: a block with no record step in it ((tk + 1) % thin == 0 for none of its four steps) takes
        // the block path as well -- per chain, the chains of a warp need not share their step counters
        bool mine = live;
        if constexpr (REC == PB200_RECORD_THINNED) mine = live && (tmod + 4u < (uint32_t)rc.thin);
        if (Geometry<LPC, CPL>::kBlockMode && REC != PB200_RECORD_ACCEPTED && !EXACT && off == 0u && cnt == 4u &&
            (REC == PB200_RECORD_NONE || __any_sync(kFull, mine))) {
            todo = block_step<LPC, CPL, BOXTEST>(gl, mine, Tf, sf, z, e, c, &p, pt, lt_pt, gi, ws, bf);
            todo = todo || (live && !mine);
            if (!__any_sync(kFull, todo)) cnt_done = true;
        }
        if constexpr (REC == PB200_RECORD_THINNED) {
            tmod += cnt;
            if (tmod >= (uint32_t)rc.thin) tmod %= (uint32_t)rc.thin;
        }
        // the (up to) four steps of this block, step index k static so that z[.][k] are plain registers (a dynamic
        // k costs a select ladder per coordinate and step); off / cnt are warp-uniform
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k) {
            if (!cnt_done && k >= off && k < off + cnt) {
#pragma unroll
                for (int r = 0; r < CPL; ++r) zk[r] = z[r][k];
                one_step<LPC, CPL, REC, EXACT>(p, pt, lt_pt, gl, gi, todo, t + (k - off), Tf, sf, zk, e[k], c, ws, rc);
            }
        }
        i += cnt;
    }
}

template <int LPC, int CPL, int REC, int SRC = kSrcPhilox, bool EXACT = (REC == PB200_RECORD_ACCEPTED),
          bool BOXTEST = false>
__device__ __forceinline__ void run_steps(const pb200_problem& p, int pt, const float* __restrict__ lt_pt, int gl, int gi,
                                          bool live, uint32_t k0, uint32_t k1, uint32_t chain_id, uint32_t t_begin,
                                          uint32_t n_steps, float Tf, float sf, const float (&lam)[CPL],
                                          ChainState<CPL>& c, float* ws, RecCursor& rc, Ring& ring,
                                          BoxFilter* bf = nullptr) {
    asm volatile("" : "+f"(Tf), "+f"(sf));      // keep the FP32 copies live instead of re-converting the doubles
    const float half_s = __fmul_rn(0.5f, sf);
#pragma unroll
    for (int r = 0; r < CPL; ++r) c.hl[r] = __fmul_rn(half_s, lam[r]);
    // chains of one warp normally advance in lockstep; if the step counters of its LIVE chains differ modulo 4 every
    // alignment class gets its own pass so that each chain still sees its own block structure
    const unsigned lm = __ballot_sync(kFull, live);
    if (SRC != kSrcRing && lm == 0u) return;
    const uint32_t lead = __shfl_sync(kFull, t_begin, lm ? __ffs(lm) - 1 : 0) & 3u;
    if (SRC == kSrcRing || __all_sync(kFull, !live || (t_begin & 3u) == lead)) {
        // (ring: the kernel only pairs a producer with warps whose live chains share one alignment class)
        run_steps_class<LPC, CPL, REC, SRC, EXACT, BOXTEST>(p, pt, lt_pt, gl, gi, live, k0, k1, chain_id, t_begin, lead,
                                                            n_steps, Tf, sf, c, ws, rc, ring, bf);
    } else {
#pragma unroll 1
        for (uint32_t a = 0; a < 4u; ++a) {
            const bool mine = live && (t_begin & 3u) == a;
            if (__any_sync(kFull, mine))
                run_steps_class<LPC, CPL, REC, SRC, EXACT, BOXTEST>(p, pt, lt_pt, gl, gi, mine, k0, k1, chain_id, t_begin, a,
                                                                    n_steps, Tf, sf, c, ws, rc, ring, bf);
        }
    }
}

// Make x the new reference point of the chain: exact loglkl + whitened gradient, w = 0, fresh safe radius.
template <int LPC, int CPL>
__device__ __forceinline__ void anchor(const pb200_problem& p, int pt, int gl, const double (&x)[CPL],
                                       ChainState<CPL>& c, double* qs) {
    double gt[CPL];
    c.ll = exact_eval<LPC, CPL, true>(p, pt, x, gl, qs, gt);
#pragma unroll
    for (int r = 0; r < CPL; ++r) { c.xref[r] = x[r]; c.w[r] = 0.0f; c.gf[r] = (float)gt[r]; }
    c.m2 = safe_radius2<LPC, CPL>(p, pt, gl, c.xref);
}


// ----------------------------------------------------------------------------------------------
// Fused annealing: n_iter iterations x S steps per launch, one lane group per chain.  `warp_slot` is the index of
// this warp among the warps that own chains, `stage_warp` its [G][NPAD] double staging rows in shared memory.
// SRC = kSrcRing: the variates come from the partner warp through `ring` (anneal_ws_kernel).
// ----------------------------------------------------------------------------------------------
template <int LPC, int CPL, int SRC>
__device__ __forceinline__ void anneal_body(const pb200_problem& p, const pb200_chains& ch, const pb200_schedule& s,
                                            const pb200_records& rec, int n_iter, uint32_t k0, uint32_t k1,
                                            int warp_slot,
                                            double* stage_warp, Ring& ring, uint32_t* __restrict__ progress,
                                            const pb200_peers& peers) {
    constexpr int NPAD = LPC * CPL, G = 32 / LPC;
    const int lane = threadIdx.x & 31, gi = lane / LPC, gl = lane % LPC;
    const int slot = warp_slot * G + gi;
    if (warp_slot * G >= ch.n_chains) return;         // whole warp idle
    const bool alive = slot < ch.n_chains;
    const int chain = alive ? slot : ch.n_chains - 1;
    SchedState st;
    st.status = ch.status[chain];
    st.temperature = ch.temperature[chain];
    st.step_size = ch.step_size[chain];
    st.current_best = ch.current_best[chain];
    st.prev_mult = ch.secant_prev_mult[chain];
    st.cur_mult = ch.secant_cur_mult[chain];
    st.prev_ar = ch.secant_prev_ar[chain];
    st.stage = ch.secant_stage[chain];
    st.iteration = ch.iteration[chain];
    bool live = alive && st.status == PB200_ACTIVE;
    const bool was_live = live;
    const int pt = ch.point[chain];
    const uint32_t chain_id = ch.chain_id[chain];
    uint32_t t = ch.steps_done[chain];
    const float* lt_pt = p.lt_t + (size_t)pt * NPAD * NPAD;
    double* qs = stage_warp + gi * NPAD;
    float* ws = reinterpret_cast<float*>(qs);

    float lam[CPL];
    double x[CPL], bx[CPL];
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        lam[r] = __ldg(p.lam + (size_t)pt * NPAD + gl + LPC * r);
        x[r] = ch.x[(size_t)chain * NPAD + gl + LPC * r];
    }
    ChainState<CPL> c;
    anchor<LPC, CPL>(p, pt, gl, x, c, qs);                 // loglkl of the start point (optimiser.py:55)
    RecCursor rc{};

    int it = 0;
    for (; it < n_iter && __any_sync(kFull, live); ++it) {
        c.best_ll = c.ll;
        c.nacc = 0;
#pragma unroll
        for (int r = 0; r < CPL; ++r) c.bw[r] = 0.0f;
        c.best_rel = true;
        run_steps<LPC, CPL, PB200_RECORD_NONE, SRC>(p, pt, lt_pt, gl, gi, live, k0, k1, chain_id, t,
                                                    (uint32_t)s.steps_per_iteration, (float)st.temperature,
                                                    (float)st.step_size, lam, c, ws, rc, ring);
        // set_bestfit: positions are materialised, the reported chi^2 is re-evaluated in FP64 at them
        if (c.best_rel) {
#pragma unroll
            for (int r = 0; r < CPL; ++r) c.bref[r] = c.xref[r];
        }
        materialise<LPC, CPL>(lt_pt, gl, c.bref, c.bw, ws, bx);
        materialise<LPC, CPL>(lt_pt, gl, c.xref, c.w, ws, x);
        double unused[CPL];
        const double best_exact = exact_eval<LPC, CPL, false>(p, pt, bx, gl, qs, unused);
        const int32_t nacc = c.nacc;
        anchor<LPC, CPL>(p, pt, gl, x, c, qs);             // re-anchor the incremental state at x
        if (live) {
            t += s.steps_per_iteration;
            const int row = st.iteration;
            if (row < rec.max_iterations) {
                const size_t o = (size_t)row * rec.iter_stride;
                if (gl == 0) {
                    rec.best_loglkl[o + chain] = best_exact;
                    rec.last_loglkl[o + chain] = c.ll;
                    rec.temperature[o + chain] = st.temperature;
                    rec.step_size[o + chain] = st.step_size;
                    rec.accepted[2 * o + chain] = nacc;
                }
#pragma unroll
                for (int r = 0; r < CPL; ++r) {
                    if (rec.best_x) rec.best_x[o + (size_t)chain * NPAD + gl + LPC * r] = bx[r];
                    if (rec.last_x) rec.last_x[o + (size_t)chain * NPAD + gl + LPC * r] = x[r];
                }
            }
            // fused all-gather: the same scalars go into slot `rank` of row `row` of every rank's exchange buffer
            if (peers.n_peers > 0 && gl == 0 && row < peers.rows) {
                const size_t blk = (size_t)row * peers.row_stride + (size_t)peers.rank * peers.rank_stride;
                const size_t na = (size_t)peers.n_alloc;
                for (int q = 0; q < peers.n_peers; ++q) {
                    double* dst = peers.buffer[q] + blk;
                    dst[chain] = best_exact;
                    dst[na + chain] = c.ll;
                    dst[2 * na + chain] = st.temperature;
                    dst[3 * na + chain] = st.step_size;
                    reinterpret_cast<int32_t*>(dst + 4 * na)[chain] = nacc;
                }
            }
            schedule_next(s, st, nacc, best_exact);
            live = st.status == PB200_ACTIVE;
        }
        if (progress) {          // this warp's records of launch-iteration `it` are globally visible: count it
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(progress + it, 1u);
        }
    }
    if (SRC == kSrcRing && ring.dead && live) {     // the producer stopped delivering: stop these chains, loudly
        st.status = PB200_FAILED;
        live = false;
    }
    if (progress && lane == 0) { // iterations this warp will never run (all its chains stopped) count as done
        for (int k = it; k < n_iter; ++k) atomicAdd(progress + k, 1u);
    }
    if (peers.n_peers > 0) {
        // The peer stores above are posted as each iteration ends and drain over NVLink while the next iterations
        // compute; their COMPLETION is established once, here: one system-scope fence per warp (a fence per
        // iteration costs ~9 us each, 10 % of the kernel), then the launch's arrival counter (index n_iter - 1) of
        // every rank is bumped -- one remote atomic per warp and rank.
        __threadfence_system();
        __syncwarp();
        if (lane == 0)
            for (int q = 0; q < peers.n_peers; ++q) atomicAdd_system(peers.progress[q] + (n_iter - 1), 1u);
    }
    if (was_live) {
#pragma unroll
        for (int r = 0; r < CPL; ++r) ch.x[(size_t)chain * NPAD + gl + LPC * r] = x[r];
        if (gl == 0) {
            ch.loglkl[chain] = c.ll;
            ch.temperature[chain] = st.temperature;
            ch.step_size[chain] = st.step_size;
            ch.current_best[chain] = st.current_best;
            ch.secant_prev_mult[chain] = st.prev_mult;
            ch.secant_cur_mult[chain] = st.cur_mult;
            ch.secant_prev_ar[chain] = st.prev_ar;
            ch.secant_stage[chain] = st.stage;
            ch.iteration[chain] = st.iteration;
            ch.status[chain] = st.status;
            ch.steps_done[chain] = t;
        }
    }
}

template <int LPC, int CPL>
__global__ void __launch_bounds__(32 * kWarpsPerCta, Geometry<LPC, CPL>::kMinCtas)
anneal_kernel(pb200_problem p, pb200_chains ch, pb200_schedule s, pb200_records rec, int n_iter, uint32_t k0,
              uint32_t k1, uint32_t* __restrict__ progress,
              const __grid_constant__ pb200_peers peers) {
    constexpr int NPAD = LPC * CPL, G = 32 / LPC;
    __shared__ __align__(16) double stage[kWarpsPerCta][G][NPAD];
    const int wib = threadIdx.x >> 5;
    Ring none{};
    anneal_body<LPC, CPL, kSrcPhilox>(p, ch, s, rec, n_iter, k0, k1, blockIdx.x * kWarpsPerCta + wib,
                                      &stage[wib][0][0], none, progress, peers);
}

// ----------------------------------------------------------------------------------------------
// Warp-specialised variant for batches too small to fill the SMs with chain warps (a few warps per scheduler): the
// chain update is one long dependent instruction chain per block of steps, the Philox + Box-Muller work is wide and
// independent.  Every chain warp (consumer) gets a partner warp (producer) that draws its variates into a
// shared-memory ring kRingDepth blocks ahead, so the scheduler always has an independent instruction stream to
// issue from.  Same counters, same arithmetic, same results as anneal_kernel.  A pair whose live chains do not share
// one step-counter alignment class falls back to drawing in place.
// ----------------------------------------------------------------------------------------------
#ifndef PB200_WS_PAIRS
#define PB200_WS_PAIRS 2          // consumer/producer pairs per CTA
#endif
#ifndef PB200_WS_INTERLEAVE
#define PB200_WS_INTERLEAVE 1     // 1: warps alternate consumer / producer; 0: consumers first
#endif
constexpr int kWsPairs = PB200_WS_PAIRS;
// register budget: 128 per thread for npad = 32 (65536 / (128 * 64 * pairs) CTAs per SM), unconstrained above
constexpr int ws_min_ctas(int npad) { return npad == 32 ? (8 / kWsPairs > 0 ? 8 / kWsPairs : 1) : 1; }

template <int LPC, int CPL>
__device__ __forceinline__ void produce_variates(int n, int n_iter, uint32_t steps, uint32_t k0, uint32_t k1,
                                                 uint32_t chain_id, uint32_t t, uint32_t lead, int gi, int gl,
                                                 Ring& ring) {
    float z[CPL][4], e[4];
#pragma unroll 1
    for (int it = 0; it < n_iter; ++it) {
        uint32_t i = 0;
        while (i < steps) {                     // mirrors run_steps_class: one block per pass, ragged head / tail
            const uint32_t off = (lead + i) & 3u, cnt = min(4u - off, steps - i);
            draw_block<LPC, CPL>(k0, k1, chain_id, (t + i) >> 2, n, gl, z, e);
            if (!ring_store<LPC, CPL>(ring, gi * LPC + gl, gi, gl, z, e)) return;
            i += cnt;
        }
        t += steps;
        lead = (lead + steps) & 3u;
    }
}

template <int LPC, int CPL>
__global__ void __launch_bounds__(64 * kWsPairs, ws_min_ctas(LPC * CPL))
anneal_ws_kernel(pb200_problem p, pb200_chains ch, pb200_schedule s, pb200_records rec, int n_iter, uint32_t k0,
                 uint32_t k1, uint32_t* __restrict__ progress, const __grid_constant__ pb200_peers peers) {
    constexpr int NPAD = LPC * CPL, G = 32 / LPC;
    __shared__ __align__(16) unsigned char slots[kWsPairs][kRingDepth][ring_slot_bytes<CPL>()];
    __shared__ __align__(16) double stage[kWsPairs][G][NPAD];
    __shared__ __align__(8) unsigned long long bars[kWsPairs][2 * kRingDepth];
    __shared__ uint32_t stop[kWsPairs];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gi = lane / LPC, gl = lane % LPC;
    const int pair = PB200_WS_INTERLEAVE ? wib >> 1 : wib % kWsPairs;
    const bool producer = PB200_WS_INTERLEAVE ? (wib & 1) != 0 : wib >= kWsPairs;
    if (threadIdx.x < kWsPairs) {
        for (int b = 0; b < 2 * kRingDepth; ++b) mbar_init(smem_addr(&bars[threadIdx.x][b]), 32u);
        stop[threadIdx.x] = 0u;
    }
    const int warp_slot = blockIdx.x * kWsPairs + pair;
    const bool idle = warp_slot * G >= ch.n_chains;
    const int slot = warp_slot * G + gi;
    const bool alive = !idle && slot < ch.n_chains;
    const int chain = alive ? slot : ch.n_chains - 1;
    const bool live = alive && ch.status[chain] == PB200_ACTIVE;
    const uint32_t t = ch.steps_done[chain];
    const uint32_t chain_id = ch.chain_id[chain];
    __syncthreads();      // barriers initialised; both warps of a pair have read the same chain state
    if (idle) return;
    const unsigned lm = __ballot_sync(kFull, live);
    const uint32_t lead_t = __shfl_sync(kFull, t, lm ? __ffs(lm) - 1 : 0);
    const bool paired = lm != 0u && __all_sync(kFull, !live || (t & 3u) == (lead_t & 3u));
    Ring ring;
    ring.data = smem_addr(&slots[pair][0][0]);
    ring.full = smem_addr(&bars[pair][0]);
    ring.empty = smem_addr(&bars[pair][kRingDepth]);
    ring.stop = smem_addr(&stop[pair]);
    ring.count = 0u;
    ring.dead = 0u;
    if (producer) {
        if (paired)
            produce_variates<LPC, CPL>(p.n, n_iter, (uint32_t)s.steps_per_iteration, k0, k1, chain_id,
                                       live ? t : lead_t, lead_t & 3u, gi, gl, ring);
        return;
    }
    if (paired)
        anneal_body<LPC, CPL, kSrcRing>(p, ch, s, rec, n_iter, k0, k1, warp_slot, &stage[pair][0][0], ring, progress,
                                        peers);
    else
        anneal_body<LPC, CPL, kSrcPhilox>(p, ch, s, rec, n_iter, k0, k1, warp_slot, &stage[pair][0][0], ring, progress,
                                          peers);
    __syncwarp();
    if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(ring.stop), "r"(1u) : "memory");
}

}  // namespace pb200
#include "pb200_wide.cuh"
#include "pb200_umma.cuh"
namespace pb200 {

// ----------------------------------------------------------------------------------------------
// Metropolis-Hastings rounds (jobtype mcmc): fixed temperature / step size per chain, chain recording.
// The incremental loglkl / gradient are re-anchored in FP64 every kResync steps.
// ----------------------------------------------------------------------------------------------
constexpr uint32_t kResync = 256;

template <int LPC, int CPL, int REC, bool EXACT>
__global__ void __launch_bounds__(32 * kWarpsPerCta, Geometry<LPC, CPL>::kMinCtas)
mh_kernel(pb200_problem p, pb200_chains ch, int n_steps, uint32_t k0, uint32_t k1, pb200_samples smp,
          int32_t* accepted_out) {
    constexpr int NPAD = LPC * CPL, G = 32 / LPC;
    __shared__ __align__(16) double stage[kWarpsPerCta][G][NPAD];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gi = lane / LPC, gl = lane % LPC;
    const int slot = (blockIdx.x * kWarpsPerCta + wib) * G + gi;
    if ((blockIdx.x * kWarpsPerCta + wib) * G >= ch.n_chains) return;
    const bool live = slot < ch.n_chains;
    const int chain = live ? slot : ch.n_chains - 1;
    const int pt = ch.point[chain];
    const uint32_t chain_id = ch.chain_id[chain];
    uint32_t t = ch.steps_done[chain];
    const float Tf = (float)ch.temperature[chain], sf = (float)ch.step_size[chain];
    const float* lt_pt = p.lt_t + (size_t)pt * NPAD * NPAD;
    double* qs = stage[wib][gi];
    float* ws = reinterpret_cast<float*>(qs);
    // tensor-core box filter of the block path (see box_filter_mma): the geometry of the large batches only
    constexpr bool kFilter = LPC == 8 && CPL == 4 && REC != PB200_RECORD_ACCEPTED && !EXACT;
    __shared__ __align__(16) BoxFilter filt[kFilter ? kWarpsPerCta : 1];
    BoxFilter* bf = nullptr;
    if constexpr (kFilter) {
        bf = &filt[wib];
        box_filter_setup(*bf, p, pt, lane);
    }

    float lam[CPL];
    double x[CPL];
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        lam[r] = __ldg(p.lam + (size_t)pt * NPAD + gl + LPC * r);
        x[r] = ch.x[(size_t)chain * NPAD + gl + LPC * r];
    }
    ChainState<CPL> c;
    anchor<LPC, CPL>(p, pt, gl, x, c, qs);
    c.best_ll = c.ll;
    c.nacc = 0;
#pragma unroll
    for (int r = 0; r < CPL; ++r) { c.bref[r] = c.xref[r]; c.bw[r] = 0.0f; }
    c.best_rel = true;

    RecCursor rc{};
    if (REC != PB200_RECORD_NONE) {
        rc.capacity = smp.capacity;
        rc.thin = smp.thin > 0 ? smp.thin : 1;
        rc.x = smp.x + (size_t)chain * smp.capacity * NPAD;
        rc.ll = smp.loglkl + (size_t)chain * smp.capacity;
        rc.mult = smp.mult + (size_t)chain * smp.capacity;
        rc.count = smp.count[chain];
        if (REC == PB200_RECORD_ACCEPTED && live) {
            if (rc.count == 0) {
                record_row<LPC, CPL>(rc, x, c.ll, 1, gl);   // Chain([1], [loglkl(x0)], x0), mcmc.py:125
                rc.cur_mult = 1;
            } else {
                rc.cur_mult = rc.count - 1 < rc.capacity ? rc.mult[rc.count - 1] : 1;
            }
        }
    }
    uint32_t done = 0;
    Ring no_ring{};
    while (done < (uint32_t)n_steps) {
        const uint32_t chunk = min((uint32_t)n_steps - done, kResync);
        run_steps<LPC, CPL, REC, kSrcPhilox, EXACT, true>(p, pt, lt_pt, gl, gi, live, k0, k1, chain_id, t + done, chunk, Tf,
                                                          sf, lam, c, ws, rc, no_ring, bf);
        done += chunk;
        materialise<LPC, CPL>(lt_pt, gl, c.xref, c.w, ws, x);
        anchor<LPC, CPL>(p, pt, gl, x, c, qs);
    }
    if (!live) return;
    if (REC != PB200_RECORD_NONE && gl == 0) {
        if (REC == PB200_RECORD_ACCEPTED && rc.count - 1 < rc.capacity) rc.mult[rc.count - 1] = rc.cur_mult;
        smp.count[chain] = rc.count;
    }
#pragma unroll
    for (int r = 0; r < CPL; ++r) ch.x[(size_t)chain * NPAD + gl + LPC * r] = x[r];
    if (gl == 0) {
        ch.loglkl[chain] = c.ll;
        ch.steps_done[chain] = t + (uint32_t)n_steps;
        if (accepted_out) accepted_out[chain] = c.nacc;
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
        const double ll = exact_eval<32, CPL, false>(p, pt, x, lane, stage[wib], unused);
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
__global__ void chain_min_kernel(const double* __restrict__ best, const int32_t* __restrict__ accepted,
                                 long long stride_best, long long stride_acc, int n_iter, int b,
                                 double* __restrict__ chain_best, int32_t* __restrict__ chain_iter) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= b) return;
    double cb = INFINITY;
    int ci = -1;
    for (int k = 0; k < n_iter; ++k) {
        const double v = best[(size_t)k * stride_best + ch];
        if (accepted[(size_t)k * stride_acc + ch] >= 0 && v < cb) { cb = v; ci = k; }
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
// keeps kPairs (i,j) accumulators in registers and stores them into the CTA's own row of `partial`
// ([gridDim.x][1 + n + n*n]); moments_reduce_kernel adds the rows in CTA order.  No atomics: the sums -- and with them
// the bin covariances every proposal factor is built from -- are reproducible run to run (FP64 atomicAdd across CTAs
// rounded differently from one run to the next, which the FP64 hosted path showed in the last bits).
constexpr int kMomThreads = 256, kMomTile = 32, kPairs = 4;
constexpr long long kMomScratchDoubles = 4LL << 20;      // <= 32 MB of partial sums per call

__global__ void __launch_bounds__(kMomThreads)
moments_kernel(const double* __restrict__ x, const int32_t* __restrict__ w, long long n_rows, int ld, int n,
               double* __restrict__ partial) {
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
    double* __restrict__ out = partial + (size_t)blockIdx.x * (size_t)(1 + n + nn);
#pragma unroll
    for (int q = 0; q < kPairs; ++q)
        if (pi[q] >= 0) out[1 + n + pi[q] * n + pj[q]] = acc[q];
    if (chunk == 0) {
        if (threadIdx.x < n) out[1 + threadIdx.x] = acc1;
        if (threadIdx.x == 0) out[0] = acc0;
    }
}

__global__ void moments_reduce_kernel(const double* __restrict__ partial, int parts, int len, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    double s = 0.0;
    for (int p = 0; p < parts; ++p) s = __dadd_rn(s, partial[(size_t)p * len + i]);      // CTA order: reproducible
    out[i] = s;
}

// Per-chain weighted first moments of the recorded chains (sufficient statistics of R-1 together with the pooled
// second moments): one warp per chain, lanes over coordinates, rows in order (plain multiply + add so that a
// sequential host loop reproduces every bit).  A row is one coalesced 8 * npad byte read.
constexpr int kChainMomWarps = 8;

__global__ void __launch_bounds__(32 * kChainMomWarps)
chain_moments_kernel(const double* __restrict__ x, const int32_t* __restrict__ mult, const int32_t* __restrict__ count,
                     int n_chains, int capacity, int ld, int n, double* __restrict__ out_sw,
                     double* __restrict__ out_swx, double* __restrict__ out_scaled) {
    const int lane = threadIdx.x & 31, chain = blockIdx.x * kChainMomWarps + (threadIdx.x >> 5);
    if (chain >= n_chains) return;
    const int rows = min(count[chain], capacity);
    const double* __restrict__ xc = x + (size_t)chain * capacity * ld;
    const int32_t* __restrict__ mc = mult + (size_t)chain * capacity;
    double acc[8], sw = 0.0;                       // n <= 256: at most 8 coordinates per lane
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0;
    for (int r = 0; r < rows; ++r) {
        const double w = (double)__ldg(mc + r);
        sw = __dadd_rn(sw, w);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            if (c < n) acc[j] = __dadd_rn(acc[j], __dmul_rn(w, __ldg(xc + (size_t)r * ld + c)));
        }
    }
    const double scale = sw > 0.0 ? 1.0 / sqrt(sw) : 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        if (c < n) {
            out_swx[(size_t)chain * n + c] = acc[j];
            if (out_scaled) out_scaled[(size_t)chain * n + c] = __dmul_rn(acc[j], scale);
        }
    }
    if (lane == 0) out_sw[chain] = sw;
}

__global__ void rng_kernel(uint32_t k0, uint32_t k1, uint32_t chain_id, uint32_t step0, int n_steps, int n,
                           float* __restrict__ out_z, float* __restrict__ out_e) {
    // one warp; CPL = 8 covers every n <= 256 with the production draw_block
    const int lane = threadIdx.x & 31;
    const uint32_t t_end = step0 + (uint32_t)n_steps;
    for (uint32_t blk = step0 >> 2; (blk << 2) < t_end; ++blk) {
        float z[8][4], e[4];
        draw_block<32, 8>(k0, k1, chain_id, blk, n, lane, z, e);
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
    if (!p->saa || !p->sab || !p->mu || !p->lo || !p->hi || !p->f || !p->c0 || !p->lt_t || !p->lam || !p->lt_norm || !p->g_t || !p->h)
        return fail("problem has NULL device pointers");
    return 0;
}

// Layout guard of the wide path: its GEMM tiles address the repetitions of a point as chain = rep * P + point (the
// enumeration every entry point documents).  A batch laid out differently is stopped, loudly at the data level:
// every chain gets status FAILED and no record is written.
__global__ void wide_check_layout_kernel(pb200_chains ch, int n_points) {
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int c = threadIdx.x; c < ch.n_chains; c += blockDim.x)
        if (ch.point[c] != c % n_points) bad = 1;
    __syncthreads();
    if (bad)
        for (int c = threadIdx.x; c < ch.n_chains; c += blockDim.x) ch.status[c] = PB200_FAILED;
}

static bool use_wide_path(const pb200_problem* prob, const pb200_chains* chains) {
    if (prob->npad < 128 || chains->n_chains % prob->n_points != 0) return false;
    if (const char* env = std::getenv("PB200_WIDE")) return env[0] != '0';
    return true;
}

constexpr int kWideMaxSub = 8;
#ifndef PB200_WIDE_SUB_DEFAULT
#define PB200_WIDE_SUB_DEFAULT 4
#endif
// Side streams of the wide pipeline's sub-batches (created once per device, never destroyed).
static int wide_sub_streams(int n, cudaStream_t* out) {
    static std::mutex mu;
    static cudaStream_t pool[16][kWideMaxSub - 1] = {};
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return 1;
    if (dev < 0 || dev >= 16) return fail("wide path: device index out of range");
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n; ++i) {
        if (!pool[dev][i] && check_cuda(cudaStreamCreateWithFlags(&pool[dev][i], cudaStreamNonBlocking), "cudaStreamCreate"))
            return 1;
        out[i] = pool[dev][i];
    }
    return 0;
}

// Annealing iterations of a wide problem as three launches each (pb200_wide.cuh).  Scratch memory is stream-ordered
// (cudaMallocAsync on `st`), so the call stays asynchronous and owns nothing afterwards.
template <int NPAD>
static int anneal_launch_wide(const pb200_problem& prob, const pb200_chains& chains, const pb200_schedule& sched,
                              const pb200_records& rec, int n_iter, uint32_t k0, uint32_t k1, uint32_t* progress,
                              cudaStream_t st, const pb200_peers& peers) {
    constexpr int CPL = NPAD / 32;
    const size_t b = (size_t)chains.n_chains;
    const int P = prob.n_points, reps = chains.n_chains / P;
    static bool pool_ready = false;
    if (!pool_ready) {                     // keep freed scratch cached in the default pool between calls
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool_ready = true;
    }
    size_t off = 0;
    auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_base = carve(2 * b * NPAD * 8), o_hi = carve(2 * b * NPAD * 4), o_lo = carve(2 * b * NPAD * 4),
                 o_bx = carve(b * NPAD * 8), o_q = carve(2 * b * NPAD * 8), o_v = carve(3 * b * NPAD * 8),
                 o_nacc = carve(b * 4), o_ran = carve(b * 4);
    char* mem = nullptr;
    if (check_cuda(cudaMallocAsync((void**)&mem, off, st), "cudaMallocAsync (wide workspace)")) return 1;
    WideWs ws{(double*)(mem + o_base), (float*)(mem + o_hi), (float*)(mem + o_lo), (double*)(mem + o_bx),
              (double*)(mem + o_q), (double*)(mem + o_v), (int32_t*)(mem + o_nacc), (int32_t*)(mem + o_ran)};
    const int n_rt = (reps + kWideTileRows - 1) / kWideTileRows, n_at = (reps + kAnchorTileRows - 1) / kAnchorTileRows;
    const size_t anchor_smem = sizeof(double) * kAnchorStages * (kAnchorTileRows * kWideKChunk + kWideKChunk * kAnchorTileCols);
    static bool attr_set[2] = {false, false};
    if (!attr_set[NPAD == 256]) {
        if (check_cuda(cudaFuncSetAttribute(wide_anchor_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)anchor_smem), "cudaFuncSetAttribute (wide_anchor_kernel)"))
            return 1;
        attr_set[NPAD == 256] = true;
    }
    WideMma mma;
    if (wide_mma_prepare<NPAD>(prob, chains, ws, reps, &mma)) return 1;
    // PB200_WIDE_TIMING=1 (diagnostics): CUDA events between the launches, one synchronising report per call on stderr
    static const bool timing = std::getenv("PB200_WIDE_TIMING") && std::getenv("PB200_WIDE_TIMING")[0] == '1';
    std::vector<cudaEvent_t> marks;
    auto mark = [&]() {
        if (!timing) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        marks.push_back(e);
    };
    mark();
    const WideSub all{0, P};
    wide_check_layout_kernel<<<1, 256, 0, st>>>(chains, P);
    wide_residual_kernel<NPAD><<<148 * 2, 256, 0, st>>>(prob, chains, ws);
    wide_anchor_kernel<NPAD><<<2 * P * n_at * (NPAD / kAnchorTileCols), 256, anchor_smem, st>>>(prob, chains, ws, reps, 0x3, all);
    mark();
    // producer warps pay while the chain warps alone leave the schedulers short of instruction streams (measured:
    // profiles/README.md); PB200_WIDE_WS=0 / 1 forces the choice
    bool ws_steps = chains.n_chains <= 148 * 4;          // up to one pair per scheduler
    if (const char* env = std::getenv("PB200_WIDE_WS")) ws_steps = env[0] != '0';
    // Sub-batches of profile points on parallel streams: the steps kernel is latency-bound (its time hardly depends
    // on the number of chains), the two GEMMs are short and throughput-bound, so with the batch cut into groups of
    // points the GEMMs of one group can run in the shadow of another group's steps.  Same results (chains are
    // independent).  Measured (scripts/r02_wide_sub.sh, ms per 20-iteration call, 1 / 2 / 4 / 8 groups): 1024 chains
    // 3.85 / 3.66 / 3.96 / 3.93 (8 points), 3.77 / 3.88 / 3.92 / 3.97 (16 points): nothing to gain, the steps kernels
    // of the groups share the same schedulers; 4096 chains 12.96 / 12.36 / 11.92 / 12.06: +9 %.  So: four groups from
    // 4096 chains up, one below.  PB200_WIDE_SUB=n forces the number of groups.
    int n_sub = (P >= 2 && chains.n_chains >= 4096 && !timing) ? std::min(P, PB200_WIDE_SUB_DEFAULT) : 1;
    if (const char* env = std::getenv("PB200_WIDE_SUB"))
        if (std::atoi(env) >= 1) n_sub = std::min(std::min(P, kWideMaxSub), std::atoi(env));
    if (!mma.enabled) n_sub = 1;            // the SIMT position GEMM works on the whole batch
    cudaStream_t streams[kWideMaxSub];
    streams[0] = st;
    if (n_sub > 1 && wide_sub_streams(n_sub - 1, streams + 1)) return 1;
    cudaEvent_t fork = nullptr, join[kWideMaxSub] = {};
    if (n_sub > 1) {
        if (check_cuda(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming), "cudaEventCreate")) return 1;
        cudaEventRecord(fork, st);
        for (int g = 1; g < n_sub; ++g) cudaStreamWaitEvent(streams[g], fork, 0);
    }
    auto sub_of = [&](int g) { return WideSub{(int)((long long)P * g / n_sub), (int)((long long)P * (g + 1) / n_sub - (long long)P * g / n_sub)}; };
    auto launch_steps = [&](int g, int flags, int it, bool ring) {
        const WideSub sub = sub_of(g);
        const int nb = sub.np * reps;
        if (ring)
            wide_steps_ws_kernel<CPL><<<(nb + kWideWsPairs - 1) / kWideWsPairs, 64 * kWideWsPairs, 0, streams[g]>>>(
                prob, chains, sched, rec, ws, k0, k1, flags, it, n_iter, progress, peers, sub);
        else
            wide_steps_kernel<CPL><<<(nb + kWarpsPerCta - 1) / kWarpsPerCta, 32 * kWarpsPerCta, 0, streams[g]>>>(
                prob, chains, sched, rec, ws, k0, k1, flags, it, n_iter, progress, peers, sub);
    };
    for (int it = 0; it < n_iter; ++it) {
        const int flags = (it > 0 ? kWideFinish : 0) | kWideRun;
        for (int g = 0; g < n_sub; ++g) {
            const WideSub sub = sub_of(g);
            launch_steps(g, flags, it, ws_steps);
            mark();
            if (mma.enabled) {
                if (wide_mma_launch<NPAD>(prob, chains, ws, reps, mma, streams[g], sub)) return 1;
            } else {                // (SIMT position GEMM: whole batch, n_sub is 1 -- see below)
                wide_materialise_simt_kernel<NPAD><<<2 * P * n_rt, NPAD, 0, st>>>(prob, chains, ws, reps);
            }
            mark();
            wide_anchor_kernel<NPAD><<<3 * sub.np * n_at * (NPAD / kAnchorTileCols), 256, anchor_smem, streams[g]>>>(
                prob, chains, ws, reps, 0x7, sub);
            mark();
        }
    }
    for (int g = 0; g < n_sub; ++g) launch_steps(g, kWideFinish | kWideLast, n_iter, false);
    for (int g = 1; g < n_sub; ++g) {
        if (check_cuda(cudaEventCreateWithFlags(&join[g], cudaEventDisableTiming), "cudaEventCreate")) return 1;
        cudaEventRecord(join[g], streams[g]);
        cudaStreamWaitEvent(st, join[g], 0);
        cudaEventDestroy(join[g]);          // (released once it has completed)
    }
    if (fork) cudaEventDestroy(fork);
    if (check_cuda(cudaGetLastError(), "wide annealing launches")) return 1;
    if (timing && marks.size() >= 2) {
        cudaStreamSynchronize(st);
        double t_steps = 0, t_pos = 0, t_anchor = 0;
        float ms = 0;
        for (size_t i = 2; i + 0 < marks.size(); ++i) {      // marks: [start, after first anchor, (steps, pos, anchor)*]
            cudaEventElapsedTime(&ms, marks[i - 1], marks[i]);
            const size_t ph = (i - 2) % 3;
            (ph == 0 ? t_steps : ph == 1 ? t_pos : t_anchor) += ms;
        }
        std::fprintf(stderr, "[pb200 wide timing] %d iterations: steps %.1f us, positions %.1f us, anchors %.1f us per iteration\n",
                     n_iter, 1e3 * t_steps / n_iter, 1e3 * t_pos / n_iter, 1e3 * t_anchor / n_iter);
        for (cudaEvent_t e : marks) cudaEventDestroy(e);
    }
    return check_cuda(cudaFreeAsync(mem, st), "cudaFreeAsync (wide workspace)");
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

// lanes per chain: small groups serve several chains per warp instruction, large groups give more warps.
static int choose_lanes(int npad, int n_chains) {
    // Measured on B200 (30-D toy, profiles/README.md): fewer lanes per chain = fewer warp instructions per chain-step.
    // npad = 32 (all geometries take four steps per reduction): more lanes per chain = more warps, which is what a small,
    // latency-bound batch needs -- 32 lanes up to 1536 chains (5.03e9 vs 4.46e9 chain-steps/s for the 8-lane kernel with
    // producer warps at 1024 chains), 16 lanes up to 2560 (8.16e9 vs 7.73e9 at 2048), 8 lanes above (1.17e10 vs 1.06e10
    // at 4096); wider problems keep more lanes per chain.
    if (npad > 64) return 32;
    if (npad == 32) return n_chains <= 1536 ? 32 : (n_chains <= 2560 ? 16 : 8);
    return n_chains >= 4096 ? 16 : 32;
}

static int resolve_lanes(int requested, int npad, int n_chains, int* out) {
    int l = requested == 0 ? choose_lanes(npad, n_chains) : requested;
    if (const char* env = std::getenv("PB200_LANES_PER_CHAIN")) {
        if (requested == 0 && std::atoi(env) > 0) l = std::atoi(env);
    }
    const bool ok = (l == 32) || (l == 16 && npad <= 64) || (l == 8 && npad == 32);
    if (!ok) return fail("lanes_per_chain must be 32, 16 (npad <= 64) or 8 (npad == 32)");
    *out = l;
    return 0;
}

// Producer warps (anneal_ws_kernel) pay while the chain warps alone leave the schedulers idle: measured on B200 for the
// 8-lane geometry, +10 % at 1024 chains, +19 % at 3072, break-even at 4096, slower above (profiles/README.md).
// PB200_WS=0 / 1 forces the choice (A/B runs, tests).
#ifndef PB200_WS_MAX_CHAINS
#define PB200_WS_MAX_CHAINS 4095
#endif
static bool use_producer_warps(int lanes, int n_chains) {
    if (const char* env = std::getenv("PB200_WS")) {
        if (env[0] == '0') return false;
        if (env[0] == '1') return true;
    }
    return lanes == 8 && n_chains <= PB200_WS_MAX_CHAINS;
}

#define PB200_DISPATCH_GEOMETRY(npad, lanes, CALL)                                   \
    switch ((npad) * 64 + (lanes)) {                                                 \
        case 32 * 64 + 32:  { constexpr int LPC = 32, CPL = 1; CALL; } break;        \
        case 32 * 64 + 16:  { constexpr int LPC = 16, CPL = 2; CALL; } break;        \
        case 32 * 64 + 8:   { constexpr int LPC = 8,  CPL = 4; CALL; } break;        \
        case 64 * 64 + 32:  { constexpr int LPC = 32, CPL = 2; CALL; } break;        \
        case 64 * 64 + 16:  { constexpr int LPC = 16, CPL = 4; CALL; } break;        \
        case 128 * 64 + 32: { constexpr int LPC = 32, CPL = 4; CALL; } break;        \
        default:            { constexpr int LPC = 32, CPL = 8; CALL; } break;        \
    }

int pb200_choose_lanes(int32_t npad, int32_t n_chains) { return choose_lanes(npad, n_chains); }
int pb200_uses_producer_warps(int32_t lanes_per_chain, int32_t n_chains) {
    return use_producer_warps(lanes_per_chain, n_chains) ? 1 : 0;
}

static int anneal_launch(const pb200_problem* prob, const pb200_chains* chains, const pb200_schedule* sched,
                         const pb200_records* rec, int32_t n_iter, uint64_t seed, int32_t lanes_per_chain,
                         uint32_t* progress, void* stream, const pb200_peers* peers_in = nullptr) {
    pb200_peers peers{};
    if (peers_in && peers_in->n_peers > 0) {
        peers = *peers_in;
        if (peers.n_peers > PB200_MAX_PEERS || peers.rank < 0 || peers.rank >= peers.n_peers)
            return fail("anneal_run_shared: need 1 <= n_peers <= PB200_MAX_PEERS and 0 <= rank < n_peers");
        if (chains && peers.n_alloc < chains->n_chains) return fail("anneal_run_shared: n_alloc < n_chains");
        if (peers.rank_stride < 4LL * peers.n_alloc + (peers.n_alloc + 1) / 2 ||
            peers.row_stride < (int64_t)peers.n_peers * peers.rank_stride)
            return fail("anneal_run_shared: exchange buffer strides too small");
        for (int q = 0; q < peers.n_peers; ++q)
            if (!peers.buffer[q] || !peers.progress[q]) return fail("anneal_run_shared: NULL peer buffer");
    }
    if (!prob || !chains || !sched || !rec) return fail("anneal_run: NULL argument");
    if (chains->n_chains <= 0 || n_iter <= 0) return 0;      // a rank without chains (fewer chains than GPUs) is fine
    if (validate(prob)) return 1;
    if (sched->steps_per_iteration <= 0) return fail("anneal_run: steps_per_iteration must be > 0");
    if (sched->step_mode < 0 || sched->step_mode > 2) return fail("anneal_run: unknown step_mode");
    if (!rec->best_loglkl || !rec->last_loglkl || !rec->temperature || !rec->step_size || !rec->accepted)
        return fail("anneal_run: records have NULL device pointers");
    if (rec->iter_stride < chains->n_chains) return fail("anneal_run: records iter_stride < n_chains");
    int lanes;
    if (resolve_lanes(lanes_per_chain, prob->npad, chains->n_chains, &lanes)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    const int per_cta = kWarpsPerCta * (32 / lanes);
    const int blocks = (chains->n_chains + per_cta - 1) / per_cta;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    if (use_wide_path(prob, chains)) {
        if (prob->npad == 128) return anneal_launch_wide<128>(*prob, *chains, *sched, *rec, n_iter, k0, k1, progress, st, peers);
        return anneal_launch_wide<256>(*prob, *chains, *sched, *rec, n_iter, k0, k1, progress, st, peers);
    }
    if (use_producer_warps(lanes, chains->n_chains)) {
        const int pairs = (chains->n_chains + (32 / lanes) - 1) / (32 / lanes);
        const int ws_blocks = (pairs + kWsPairs - 1) / kWsPairs;
        PB200_DISPATCH_GEOMETRY(prob->npad, lanes, (anneal_ws_kernel<LPC, CPL><<<ws_blocks, 64 * kWsPairs, 0, st>>>(
                                                        *prob, *chains, *sched, *rec, n_iter, k0, k1, progress,
                                                        peers)));
        return check_cuda(cudaGetLastError(), "anneal_ws_kernel launch");
    }
    PB200_DISPATCH_GEOMETRY(prob->npad, lanes, (anneal_kernel<LPC, CPL><<<blocks, 32 * kWarpsPerCta, 0, st>>>(
                                                    *prob, *chains, *sched, *rec, n_iter, k0, k1, progress, peers)));
    return check_cuda(cudaGetLastError(), "anneal_kernel launch");
}

int pb200_anneal_run(const pb200_problem* prob, const pb200_chains* chains, const pb200_schedule* sched,
                     const pb200_records* rec, int32_t n_iter, uint64_t seed, int32_t lanes_per_chain, void* stream) {
    return anneal_launch(prob, chains, sched, rec, n_iter, seed, lanes_per_chain, nullptr, stream);
}

int pb200_anneal_run_signalled(const pb200_problem* prob, const pb200_chains* chains, const pb200_schedule* sched,
                               const pb200_records* rec, int32_t n_iter, uint64_t seed, int32_t lanes_per_chain,
                               uint32_t* progress, int32_t* out_target, void* stream) {
    if (!progress || !out_target) return fail("anneal_run_signalled: NULL progress / out_target");
    int lanes;
    if (!prob || !chains) return fail("anneal_run_signalled: NULL argument");
    if (resolve_lanes(lanes_per_chain, prob->npad, chains->n_chains, &lanes)) return 1;
    *out_target = (chains->n_chains + (32 / lanes) - 1) / (32 / lanes);      // warps that own chains
    return anneal_launch(prob, chains, sched, rec, n_iter, seed, lanes_per_chain, progress, stream);
}

int pb200_anneal_run_shared(const pb200_problem* prob, const pb200_chains* chains, const pb200_schedule* sched,
                            const pb200_records* rec, int32_t n_iter, uint64_t seed, int32_t lanes_per_chain,
                            const pb200_peers* peers, uint32_t* progress, int32_t* out_target, void* stream) {
    if (!peers) return fail("anneal_run_shared: NULL peers");
    if (!prob || !chains) return fail("anneal_run_shared: NULL argument");
    int lanes;
    if (resolve_lanes(lanes_per_chain, prob->npad, chains->n_chains, &lanes)) return 1;
    if (out_target) *out_target = (chains->n_chains + (32 / lanes) - 1) / (32 / lanes);
    return anneal_launch(prob, chains, sched, rec, n_iter, seed, lanes_per_chain, progress, stream, peers);
}

int pb200_exchange_arrivals(int32_t npad, int32_t n_chains, int32_t lanes_per_chain) {
    if (n_chains <= 0) return 0;
    int lanes;
    if (resolve_lanes(lanes_per_chain, npad, n_chains, &lanes)) return -1;
    return (n_chains + (32 / lanes) - 1) / (32 / lanes);      // warps that own chains (wide path: one per chain)
}

int pb200_launch_status(void* stream) {
    uint32_t word = 0;
    if (check_cuda(cudaMemcpyFromSymbolAsync(&word, g_fault_word, sizeof(word), 0, cudaMemcpyDeviceToHost,
                                             (cudaStream_t)stream), "launch_status: read"))
        return 1;
    if (check_cuda(cudaStreamSynchronize((cudaStream_t)stream), "launch_status: synchronize")) return 1;
    if (word == 0) return 0;
    const uint32_t zero = 0;
    cudaMemcpyToSymbol(g_fault_word, &zero, sizeof(zero));
    std::string msg = "device-side protocol fault:";
    if (word & kFaultRing) msg += " a producer-warp ring timed out (anneal_ws_kernel);";
    if (word & kFaultPipe) msg += " a TMA / tcgen05 pipeline barrier timed out (wide_materialise_mma_kernel);";
    return fail(msg + " the affected chains were stopped, results of this launch are incomplete");
}

// cuStreamWaitValue32 through the runtime's driver entry point lookup: no link-time dependency on libcuda, so
// the library still loads (and exports every symbol) on a machine without a driver.
int pb200_stream_wait_geq(void* stream, const uint32_t* counter, uint32_t value) {
    typedef int (*wait_fn)(void*, unsigned long long, unsigned int, unsigned int);
    static const wait_fn fn = []() -> wait_fn {          // initialised once, thread-safe (C++11 magic static)
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult status;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &ptr, cudaEnableDefault, &status) != cudaSuccess ||
            status != cudaDriverEntryPointSuccess)
            return nullptr;
        return (wait_fn)ptr;
    }();
    if (!fn) return fail("cuStreamWaitValue32 is not available");
    if (!counter) return fail("stream_wait_geq: NULL counter");
    const int rc = fn(stream, (unsigned long long)(uintptr_t)counter, value, 0x0 /* CU_STREAM_WAIT_VALUE_GEQ */);
    if (rc != 0) return fail("cuStreamWaitValue32 failed with CUresult " + std::to_string(rc));
    return 0;
}

int pb200_mh_run(const pb200_problem* prob, const pb200_chains* chains, int32_t n_steps, uint64_t seed,
                 const pb200_samples* samples, int32_t* accepted_out, int32_t lanes_per_chain, void* stream) {
    if (validate(prob)) return 1;
    if (!chains) return fail("mh_run: NULL chains");
    if (chains->n_chains <= 0 || n_steps <= 0) return 0;
    pb200_samples none{};
    pb200_samples smp = samples ? *samples : none;
    const bool exact_box = (smp.mode & PB200_EXACT_BOX) != 0;
    smp.mode &= ~PB200_EXACT_BOX;
    if (smp.mode != PB200_RECORD_NONE) {
        if (!smp.x || !smp.loglkl || !smp.mult || !smp.count || smp.capacity <= 0)
            return fail("mh_run: sample buffers missing");
        if (smp.mode == PB200_RECORD_THINNED && smp.thin <= 0) return fail("mh_run: thin must be > 0");
    }
    int lanes;
    if (resolve_lanes(lanes_per_chain, prob->npad, chains->n_chains, &lanes)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    const int per_cta = kWarpsPerCta * (32 / lanes);
    const int blocks = (chains->n_chains + per_cta - 1) / per_cta;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#define PB200_MH(REC, EXACT)                                                                                           \
    PB200_DISPATCH_GEOMETRY(prob->npad, lanes, (mh_kernel<LPC, CPL, REC, EXACT><<<blocks, 32 * kWarpsPerCta, 0, st>>>( \
                                                    *prob, *chains, n_steps, k0, k1, smp, accepted_out)))
    if (smp.mode == PB200_RECORD_NONE) { if (exact_box) { PB200_MH(PB200_RECORD_NONE, true); } else { PB200_MH(PB200_RECORD_NONE, false); } }
    else if (smp.mode == PB200_RECORD_ACCEPTED) { PB200_MH(PB200_RECORD_ACCEPTED, true); }
    else if (smp.mode == PB200_RECORD_THINNED) { if (exact_box) { PB200_MH(PB200_RECORD_THINNED, true); } else { PB200_MH(PB200_RECORD_THINNED, false); } }
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

static int validate_hosted(const pb200_hosted* h, const pb200_chains* c, const char* who) {
    if (!h || !c) return fail(std::string(who) + ": NULL argument");
    if (h->npad < 32 || h->npad > 256 || h->npad % 32 != 0) return fail(std::string(who) + ": npad must be 32, 64, ..., 256");
    if (h->n < 1 || h->n > h->npad) return fail(std::string(who) + ": need 1 <= n <= npad");
    if (h->n_points < 1 || !h->factor_t || !h->lo || !h->hi) return fail(std::string(who) + ": incomplete pb200_hosted");
    if (!c->x || !c->loglkl || !c->temperature || !c->step_size || !c->point || !c->chain_id || !c->steps_done ||
        !c->status)
        return fail(std::string(who) + ": chains have NULL device pointers");
    return 0;
}

int pb200_hosted_propose(const pb200_hosted* hosted, const pb200_chains* chains, uint64_t seed, double* out_prop,
                         int32_t* out_outside, void* stream) {
    if (validate_hosted(hosted, chains, "hosted_propose")) return 1;
    if (!out_prop || !out_outside) return fail("hosted_propose: NULL output");
    if (chains->n_chains <= 0) return 0;
    hosted_propose_kernel<<<(chains->n_chains + kHostedWarps - 1) / kHostedWarps, 32 * kHostedWarps,
                            sizeof(double) * kHostedWarps * hosted->npad, (cudaStream_t)stream>>>(
        *hosted, *chains, (uint32_t)seed, (uint32_t)(seed >> 32), out_prop, out_outside);
    return check_cuda(cudaGetLastError(), "hosted_propose_kernel launch");
}

int pb200_hosted_accept(const pb200_hosted* hosted, const pb200_chains* chains, uint64_t seed, const double* prop,
                        const double* prop_loglkl, const int32_t* outside, int32_t* accepted_flag,
                        int32_t* n_accepted, double* best_loglkl, double* best_x, void* stream) {
    if (validate_hosted(hosted, chains, "hosted_accept")) return 1;
    if (!prop || !prop_loglkl || !outside) return fail("hosted_accept: NULL input");
    if ((best_loglkl == nullptr) != (best_x == nullptr)) return fail("hosted_accept: best_loglkl and best_x go together");
    if (chains->n_chains <= 0) return 0;
    hosted_accept_kernel<<<(chains->n_chains + kHostedWarps - 1) / kHostedWarps, 32 * kHostedWarps, 0,
                           (cudaStream_t)stream>>>(*hosted, *chains, (uint32_t)seed, (uint32_t)(seed >> 32), prop,
                                                   prop_loglkl, outside, accepted_flag, n_accepted, best_loglkl,
                                                   best_x);
    return check_cuda(cudaGetLastError(), "hosted_accept_kernel launch");
}

int pb200_profile_reduce(const double* best_loglkl, const int32_t* accepted, int64_t stride_best,
                         int64_t stride_accepted, int32_t n_iter, int32_t b, int32_t n_points, int32_t reps,
                         double* out_chain_best, int32_t* out_chain_iter, double* out_point_best,
                         int32_t* out_point_rep, double* out_point_avg, void* stream) {
    if (!best_loglkl || !accepted || !out_chain_best || !out_chain_iter || !out_point_best || !out_point_rep ||
        !out_point_avg)
        return fail("profile_reduce: NULL argument");
    if (b != n_points * reps) return fail("profile_reduce: b must equal n_points * reps");
    if (n_points <= 0) return 0;
    chain_min_kernel<<<(b + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        best_loglkl, accepted, (long long)stride_best, (long long)stride_accepted, n_iter, b, out_chain_best,
        out_chain_iter);
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
    const int len = 1 + n + n * n;
    const int gx = (int)std::max<long long>(1, std::min<long long>({tiles, 148LL * 4, kMomScratchDoubles / len}));
    const size_t smem = sizeof(double) * kMomTile * (n + 1);
    if (gx == 1) {                                   // one CTA per pair chunk: its sums are the result
        moments_kernel<<<dim3(1, chunks), kMomThreads, smem, st>>>(x, w, (long long)n_rows, ld, n, out);
        return check_cuda(cudaGetLastError(), "moments_kernel launch");
    }
    double* partial = nullptr;
    if (check_cuda(cudaMallocAsync(&partial, sizeof(double) * (size_t)gx * len, st), "weighted_moments: scratch")) return 1;
    moments_kernel<<<dim3(gx, chunks), kMomThreads, smem, st>>>(x, w, (long long)n_rows, ld, n, partial);
    moments_reduce_kernel<<<(len + 255) / 256, 256, 0, st>>>(partial, gx, len, out);
    const int rc = check_cuda(cudaGetLastError(), "moments_kernel launch");
    cudaFreeAsync(partial, st);
    return rc;
}

int pb200_chain_moments(const double* x, const int32_t* mult, const int32_t* count, int32_t n_chains,
                        int32_t capacity, int32_t ld, int32_t n, double* out_sw, double* out_swx,
                        double* out_scaled, void* stream) {
    if (!x || !mult || !count || !out_sw || !out_swx) return fail("chain_moments: NULL argument");
    if (n < 1 || n > 256 || ld < n || capacity < 1) return fail("chain_moments: need 1 <= n <= 256, ld >= n, capacity >= 1");
    if (n_chains <= 0) return 0;
    chain_moments_kernel<<<(n_chains + kChainMomWarps - 1) / kChainMomWarps, 32 * kChainMomWarps, 0,
                           (cudaStream_t)stream>>>(x, mult, count, n_chains, capacity, ld, n, out_sw, out_swx,
                                                   out_scaled);
    return check_cuda(cudaGetLastError(), "chain_moments_kernel launch");
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
