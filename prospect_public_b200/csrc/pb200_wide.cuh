// pb200_wide.cuh -- the annealing path for wide problems (npad >= 128, e.g. the 256-D profile of BASELINE.json
// configs[3]).  Included by pb200_kernels.cu (one translation unit) after the step machinery it reuses.
//
// Why a second path.  In the fused kernel a warp owns a chain for the whole launch, so everything that happens at an
// iteration boundary -- materialising x = xref + Lt w (n x n, FP32), re-anchoring chi^2 and the whitened gradient at
// x and re-evaluating the bestfit (three n x n FP64 mat-vecs) -- is done chain by chain: at n = 255 every chain
// streams ~2 MB of matrices through L1/L2 per iteration although the R repetitions of a profile point share them
// (ncu, 1024 chains: 31 GB of L2 traffic per 20-iteration launch, 43 % of the stall samples; profiles/README.md).
// Those contractions are DENSE and BATCHED over the repetitions of a point:
//
//     [X ; BX] = [Xref ; Bref] + [W ; BW] Lt^T        (2R x n) (n x n) per point, FP32 -> tensor cores (tcgen05, 3xTF32)
//     V = Q S_aa,   Gt = Q G^T + h,   Vb = Qb S_aa    (R x n) (n x n) per point, FP64 -> shared-memory tiled DFMA GEMM
//
// so here an iteration is three launches: wide_steps_kernel (the S Metropolis steps, one warp per chain, the same
// run_steps as the fused kernel) -> materialise (GEMM) -> wide_anchor_kernel (GEMM + chi^2 / gradient epilogue).
// set_bestfit / get_next_iteration_settings / records / the iteration-boundary exchange of iteration k run in the
// prologue of the steps kernel of iteration k + 1 (or of a last, steps-less launch).  A launch boundary costs ~3 us
// against ~100 us of steps; the matrices are read once per 16-row tile instead of once per chain.
//
// Results are bit-identical to the fused kernel when the FP32 materialisation runs on the SIMT pipe
// (wide_materialise_simt_kernel keeps materialise()'s summation order; the FP64 GEMM keeps exact_eval()'s: one
// accumulator per output, k ascending).  The tensor-core materialisation (pb200_umma.cuh) rounds differently (3xTF32
// products, hardware accumulation order): positions differ by ~1e-7 of the proposal scale, every chi^2 is still the
// exact FP64 value at the written position.
#pragma once

namespace pb200 {

constexpr int kWideTileRows = 16;     // rows (chains of one point) per tile of the SIMT position GEMM
constexpr int kAnchorTileRows = 32;   // rows per tile of the FP64 anchor GEMM: every warp re-reads the matrix chunk from
                                      // shared memory, so taller thread tiles (8 rows x 4 columns) halve the LSU traffic
                                      // per DFMA -- with 16-row tiles the kernel was shared-memory bound (45 us -> see
                                      // profiles/README.md)
constexpr int kWideKChunk = 16;       // k-rows of the matrix per shared-memory stage of the FP64 GEMM

struct WideWs {
    double*  base;      // [2][B][npad]  reference point of the last (0) / best (1) position after the steps
    float*   w_hi;      // [2][B][npad]  whitened displacement from it, split w = hi + lo (hi: 10 mantissa bits = TF32)
    float*   w_lo;      // [2][B][npad]
    double*  bx;        // [B][npad]     materialised best position
    double*  q;         // [2][B][npad]  residuals X - mu (0) and BX - mu (1): the A operand of the anchor GEMM, written by
                        //               the position GEMM's epilogue
    double*  v;         // [3][B][npad]  anchor GEMM outputs: 0: (X - mu) S_aa, 1: (X - mu) G_pt^T, 2: (BX - mu) S_aa
    int32_t* nacc;      // [B]           accepted proposals of the iteration just run
    int32_t* ran;       // [B]           1: the chain ran that iteration
};

// A launch of the wide pipeline covers the profile points [p0, p0 + np) of the batch (all their repetitions): the
// per-iteration pipeline of a batch is issued as several such sub-batches on parallel streams, so that the position and
// anchor GEMMs of one run in the shadow of another's (latency-bound) steps kernel.  Chains are independent, hence the
// same results bit for bit.
struct WideSub {
    int p0, np;
    // chain of local index j in this sub-batch (chain = rep * P + point), -1 past its end
    __device__ __forceinline__ int chain_of(int j, int n_points, int n_chains) const {
        const int reps = n_chains / n_points;
        if (j >= np * reps) return -1;
        return (j / np) * n_points + p0 + j % np;
    }
};

constexpr int kWideFinish = 1;   // prologue: set_bestfit / records / next settings of the iteration run by the previous launch
constexpr int kWideRun = 2;      // run the S steps of the next iteration
constexpr int kWideLast = 4;     // last launch of the call: completion of the peer exchange

__device__ __forceinline__ void split_tf32(float w, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);     // sign, exponent, 10 mantissa bits: exact in TF32
    lo = __fsub_rn(w, hi);                                       // exact (<= 13 significant bits)
}

// Producer of the wide geometry (32 lanes per chain): one coordinate at a time, straight into the ring slot.  Nothing
// is kept in registers across coordinates beyond the Philox chains in flight (PB200_WIDE_PRODUCER_UNROLL of them).  Same
// counters and the same device functions as draw_block -> the same numbers.
#ifndef PB200_WIDE_PRODUCER_UNROLL
#define PB200_WIDE_PRODUCER_UNROLL 8    // independent Philox chains in flight per producer warp: a Philox round is a
                                        // dependent IMAD.WIDE -> LOP3 pair (~12 cycles), so 2 chains left the warp at one
                                        // instruction per ~6 cycles (measured); 8 make it pipe-bound
#endif
constexpr int kProducerUnroll = PB200_WIDE_PRODUCER_UNROLL;
template <int CPL>
__device__ __forceinline__ void produce_variates_wide(int n, uint32_t steps, uint32_t k0, uint32_t k1, uint32_t chain_id,
                                                      uint32_t t, int lane, Ring& rg) {
    constexpr int NPAD = 32 * CPL;
    const int acc_r = n / 32, acc_lane = n % 32;
    uint32_t i = 0, lead = t & 3u;
    while (i < steps) {                         // mirrors run_steps_class: one Philox block per pass
        const uint32_t off = (lead + i) & 3u, cnt = min(4u - off, steps - i), blk = (t + i) >> 2;
        const uint32_t slot = rg.count % kRingDepth, parity = ((rg.count / kRingDepth) & 1u) ^ 1u;
        const uint32_t bar = rg.empty + 8u * slot;
        while (!__all_sync(kFull, mbar_try_wait(bar, parity))) {
            uint32_t stop;
            asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(stop) : "r"(rg.stop));
            if (__any_sync(kFull, stop != 0u)) return;
        }
        const uint32_t base = rg.data + slot * (uint32_t)ring_slot_bytes<CPL>();
#pragma unroll kProducerUnroll
        for (int r = 0; r < CPL; ++r) {
            const int coord = lane + 32 * r;
            uint32_t w[4];
            float z0, z1, z2, z3;
#ifdef PB200_DBG_NO_RNG          // timing experiments only: how long does the consumer take on its own?
            w[0] = w[1] = w[2] = w[3] = blk + coord; z0 = z1 = z2 = z3 = 1e-3f * (float)coord;
#else
            philox4x32_10(blk, coord == n ? kAcceptSlot : (uint32_t)coord, chain_id, 0u, k0, k1, w);
            box_muller(w[0], w[1], z0, z1);
            box_muller(w[2], w[3], z2, z3);
#endif
            if (coord >= n) { z0 = z1 = z2 = z3 = 0.0f; }
            const uint32_t dst = base + (uint32_t)(r * 128 + lane) * 4u;      // [r][k][lane]: conflict-free both ways
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst), "f"(z0) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + 128u), "f"(z1) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + 256u), "f"(z2) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + 384u), "f"(z3) : "memory");
            if (r == acc_r && lane == acc_lane)        // the spare coordinate n carries the accept variates
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)(CPL * 32) * 16u),
                             "f"(exp_variate(w[0])), "f"(exp_variate(w[1])), "f"(exp_variate(w[2])),
                             "f"(exp_variate(w[3])) : "memory");
        }
        if (n >= NPAD && lane == 0) {
            uint32_t w[4];
            philox4x32_10(blk, kAcceptSlot, chain_id, 0u, k0, k1, w);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)(CPL * 32) * 16u),
                         "f"(exp_variate(w[0])), "f"(exp_variate(w[1])), "f"(exp_variate(w[2])),
                         "f"(exp_variate(w[3])) : "memory");
        }
        mbar_arrive(rg.full + 8u * slot);
        rg.count += 1;
        i += cnt;
    }
}

// Consumer side of that ring: the S steps of one chain.  The variates of a step are read from the slot as they are
// needed (8 conflict-free 4-byte reads per step) instead of being copied to a 8 x 4 register tile first, so the step
// index is a run-time address offset and the step loop stays ROLLED: one copy of one_step, a hot loop of a few
// hundred instructions, ~30 fewer live registers.  Same arithmetic as run_steps (one_step is the same function).
template <int CPL>
__device__ __forceinline__ void wide_run_steps_ring(const pb200_problem& p, int pt, const float* __restrict__ lt_pt,
                                                    int gl, bool live, uint32_t t_begin, uint32_t n_steps, float Tf,
                                                    float sf, const float (&lam)[CPL], ChainState<CPL>& c, float* ws,
                                                    Ring& rg) {
    asm volatile("" : "+f"(Tf), "+f"(sf));
    const float half_s = __fmul_rn(0.5f, sf);
#pragma unroll
    for (int r = 0; r < CPL; ++r) c.hl[r] = __fmul_rn(half_s, lam[r]);
    RecCursor rc{};
    const uint32_t lead = t_begin & 3u;
    uint32_t i = 0;
    while (i < n_steps) {
        const uint32_t off = (lead + i) & 3u, cnt = min(4u - off, n_steps - i);
        const uint32_t slot = rg.count % kRingDepth, parity = (rg.count / kRingDepth) & 1u;
        const uint32_t bar = rg.full + 8u * slot;
        if (!rg.dead && !__all_sync(kFull, mbar_try_wait(bar, parity))) {
            const long long t0 = clock64();
            while (!__all_sync(kFull, mbar_try_wait(bar, parity))) {
                if (__any_sync(kFull, clock64() - t0 > PB200_WAIT_TIMEOUT_CYCLES)) {
                    if (gl == 0) atomicOr(&g_fault_word, kFaultRing);
                    rg.dead = 1u;
                    break;
                }
            }
        }
        const uint32_t base = rg.data + slot * (uint32_t)ring_slot_bytes<CPL>();
        bool todo = live, block_done = false;
#ifdef PB200_DBG_NO_STEPS        // timing experiments only: how long does the producer take on its own?
        block_done = true;
#endif
        if (off == 0u && cnt == 4u && !rg.dead && !block_done) {       // a whole Philox block: four steps per reduction (block_step)
            float z[CPL][4], e[4];
#pragma unroll
            for (int r = 0; r < CPL; ++r) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(z[r][k]) : "r"(base + (uint32_t)((r * 4 + k) * 32 + gl) * 4u));
            }
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(e[0]), "=f"(e[1]), "=f"(e[2]), "=f"(e[3]) : "r"(base + (uint32_t)(CPL * 32) * 16u));
            todo = block_step<32, CPL>(gl, live, Tf, sf, z, e, c);
            block_done = !__any_sync(kFull, todo);      // else: replay the block step by step (exact box test needed)
        }
#pragma unroll 1
        for (uint32_t k = off; k < off + cnt && !block_done; ++k) {
            float zk[CPL], ek = 0.0f;
#pragma unroll
            for (int r = 0; r < CPL; ++r) zk[r] = 0.0f;
            if (!rg.dead) {
#pragma unroll
                for (int r = 0; r < CPL; ++r)
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(zk[r]) : "r"(base + (uint32_t)((r * 4 + k) * 32 + gl) * 4u));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ek) : "r"(base + (uint32_t)(CPL * 32) * 16u + 4u * k));
            }
            one_step<32, CPL, PB200_RECORD_NONE, false>(p, pt, lt_pt, gl, 0, todo, t_begin + i + (k - off), Tf, sf, zk, ek, c,
                                                        ws, rc);
        }
        if (!rg.dead) {
            mbar_arrive(rg.empty + 8u * slot);
            rg.count += 1;
        }
        i += cnt;
    }
}

// ----------------------------------------------------------------------------------------------
// Steps: one warp per chain (32 lanes, CPL coordinates per lane), state in registers for the S steps.
// `chain` is this warp's chain, `stage` its [NPAD] double staging row, SRC where the variates come from.
// ----------------------------------------------------------------------------------------------
template <int CPL, int SRC>
__device__ __forceinline__ void wide_steps_body(const pb200_problem& p, const pb200_chains& ch, const pb200_schedule& s,
                                                const pb200_records& rec, const WideWs& ws, uint32_t k0, uint32_t k1,
                                                int flags, int it, int n_iter, uint32_t* __restrict__ progress,
                                                const pb200_peers& peers, int chain, double* stage, Ring& ring) {
    constexpr int LPC = 32, NPAD = LPC * CPL;
    const int gl = threadIdx.x & 31;
    const size_t b = (size_t)ch.n_chains;
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
    uint32_t t = ch.steps_done[chain];
    const int pt = ch.point[chain];
    __syncwarp();        // every lane has read the chain state before lane 0 rewrites it below

    // chi^2 from the anchor GEMM's V = Q S_aa: sum_i q_i (1/2 v_i + f s_ab,i) + c0 -- the tail of exact_eval(), same
    // lanes, same order, hence the same bits
    auto chi2_of = [&](const double* __restrict__ x_row, const double* __restrict__ v_row) {
        const double f = __ldg(p.f + pt);
        double part = 0.0;
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            const int c = gl + LPC * r;
            const double q = x_row[c] - __ldg(p.mu + c);
            part = fma(q, fma(0.5, v_row[c], f * __ldg(p.sab + c)), part);
        }
        return group_sum<LPC>(part) + __ldg(p.c0 + pt);
    };
    const double ll_x = chi2_of(ch.x + (size_t)chain * NPAD, ws.v + (size_t)chain * NPAD);

    if (flags & kWideFinish) {
        if (ws.ran[chain]) {                 // uniform over the warp
            const int32_t nacc = ws.nacc[chain];
            const double best_exact = chi2_of(ws.bx + (size_t)chain * NPAD, ws.v + (2 * b + chain) * NPAD);
            t += (uint32_t)s.steps_per_iteration;
            const int row = st.iteration;
            if (row < rec.max_iterations) {
                const size_t o = (size_t)row * rec.iter_stride;
                if (gl == 0) {
                    rec.best_loglkl[o + chain] = best_exact;
                    rec.last_loglkl[o + chain] = ll_x;
                    rec.temperature[o + chain] = st.temperature;
                    rec.step_size[o + chain] = st.step_size;
                    rec.accepted[2 * o + chain] = nacc;
                }
#pragma unroll
                for (int r = 0; r < CPL; ++r) {
                    const size_t c = (size_t)chain * NPAD + gl + LPC * r;
                    if (rec.best_x) rec.best_x[o + c] = ws.bx[c];
                    if (rec.last_x) rec.last_x[o + c] = ch.x[c];
                }
            }
            if (peers.n_peers > 0 && gl == 0 && row < peers.rows) {       // fused all-gather (see anneal_body)
                const size_t blk = (size_t)row * peers.row_stride + (size_t)peers.rank * peers.rank_stride;
                const size_t na = (size_t)peers.n_alloc;
                for (int q = 0; q < peers.n_peers; ++q) {
                    double* dst = peers.buffer[q] + blk;
                    dst[chain] = best_exact;
                    dst[na + chain] = ll_x;
                    dst[2 * na + chain] = st.temperature;
                    dst[3 * na + chain] = st.step_size;
                    reinterpret_cast<int32_t*>(dst + 4 * na)[chain] = nacc;
                }
            }
            schedule_next(s, st, nacc, best_exact);
            if (gl == 0) {
                ch.loglkl[chain] = ll_x;
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
        if (progress) {          // the records of launch-iteration it - 1 of this warp are globally visible: count it
            __threadfence();
            __syncwarp();
            if (gl == 0) atomicAdd(progress + (it - 1), 1u);
        }
    }
    if ((flags & kWideLast) && peers.n_peers > 0) {
        __threadfence_system();
        __syncwarp();
        if (gl == 0)
            for (int q = 0; q < peers.n_peers; ++q) atomicAdd_system(peers.progress[q] + (n_iter - 1), 1u);
    }
    if (!(flags & kWideRun)) return;

    const bool live = st.status == PB200_ACTIVE;
    const float* lt_pt = p.lt_t + (size_t)pt * NPAD * NPAD;
    float* wst = reinterpret_cast<float*>(stage);
    ChainState<CPL> c;
    float lam[CPL];
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        const size_t col = (size_t)chain * NPAD + gl + LPC * r;
        lam[r] = __ldg(p.lam + (size_t)pt * NPAD + gl + LPC * r);
        c.xref[r] = ch.x[col];
        c.w[r] = 0.0f;
        c.bw[r] = 0.0f;
        c.gf[r] = (float)(ws.v[b * NPAD + col] + __ldg(p.h + (size_t)pt * NPAD + gl + LPC * r));
    }
    c.ll = ll_x;
    c.m2 = safe_radius2<LPC, CPL>(p, pt, gl, c.xref);
    c.best_ll = c.ll;
    c.nacc = 0;
    c.best_rel = true;
    if constexpr (SRC == kSrcRing) {
        wide_run_steps_ring<CPL>(p, pt, lt_pt, gl, live, t, (uint32_t)s.steps_per_iteration, (float)st.temperature,
                                 (float)st.step_size, lam, c, wst, ring);
    } else {
        RecCursor rc{};
        run_steps<LPC, CPL, PB200_RECORD_NONE, SRC>(p, pt, lt_pt, gl, 0, live, k0, k1, ch.chain_id[chain], t,
                                                    (uint32_t)s.steps_per_iteration, (float)st.temperature,
                                                    (float)st.step_size, lam, c, wst, rc, ring);
    }
    const bool ok = !(SRC == kSrcRing && ring.dead);      // a dead ring: the chain is stopped, nothing it did counts
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        const size_t col = (size_t)chain * NPAD + gl + LPC * r;
        float hi, lo;
        ws.base[col] = c.xref[r];
        ws.base[b * NPAD + col] = c.best_rel ? c.xref[r] : c.bref[r];
        split_tf32(c.w[r], hi, lo);
        ws.w_hi[col] = hi;
        ws.w_lo[col] = lo;
        split_tf32(c.bw[r], hi, lo);
        ws.w_hi[b * NPAD + col] = hi;
        ws.w_lo[b * NPAD + col] = lo;
    }
    if (gl == 0) {
        ws.nacc[chain] = c.nacc;
        ws.ran[chain] = (live && ok) ? 1 : 0;
        if (!ok) ch.status[chain] = PB200_FAILED;
    }
}

template <int CPL>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
wide_steps_kernel(pb200_problem p, pb200_chains ch, pb200_schedule s, pb200_records rec, WideWs ws, uint32_t k0,
                  uint32_t k1, int flags, int it, int n_iter, uint32_t* __restrict__ progress,
                  const __grid_constant__ pb200_peers peers, WideSub sub) {
    constexpr int NPAD = 32 * CPL;
    __shared__ __align__(16) double stage[kWarpsPerCta][NPAD];
    const int wib = threadIdx.x >> 5;
    const int chain = sub.chain_of(blockIdx.x * kWarpsPerCta + wib, p.n_points, ch.n_chains);
    if (chain < 0) return;
    Ring none{};
    wide_steps_body<CPL, kSrcPhilox>(p, ch, s, rec, ws, k0, k1, flags, it, n_iter, progress, peers, chain, stage[wib],
                                     none);
}

// The same with a PRODUCER warp per chain warp (the shared-memory ring of anneal_ws_kernel): at 256-D ~70 % of the
// warp instructions of a step are Philox + Box-Muller, independent of the chain state, while the chain update is a
// short dependent chain -- and a 1024-chain batch is only 1.7 warps per scheduler.  Splitting the two doubles the
// number of instruction streams a scheduler can pick from and keeps each under 128 registers (16 warps per SM: all
// 2 x 1024 warps resident in one wave).  Same counters, same arithmetic, same results.
constexpr int kWideWsPairs = 2;

template <int CPL>
__global__ void __launch_bounds__(64 * kWideWsPairs, 8 / kWideWsPairs)
wide_steps_ws_kernel(pb200_problem p, pb200_chains ch, pb200_schedule s, pb200_records rec, WideWs ws, uint32_t k0,
                     uint32_t k1, int flags, int it, int n_iter, uint32_t* __restrict__ progress,
                     const __grid_constant__ pb200_peers peers, WideSub sub) {
    constexpr int NPAD = 32 * CPL;
    __shared__ __align__(16) unsigned char slots[kWideWsPairs][kRingDepth][ring_slot_bytes<CPL>()];
    __shared__ __align__(16) double stage[kWideWsPairs][NPAD];
    __shared__ __align__(8) unsigned long long bars[kWideWsPairs][2 * kRingDepth];
    __shared__ uint32_t stop[kWideWsPairs];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, pair = wib >> 1;
    const bool producer = (wib & 1) != 0;
    if (threadIdx.x < kWideWsPairs) {
        for (int i = 0; i < 2 * kRingDepth; ++i) mbar_init(smem_addr(&bars[threadIdx.x][i]), 32u);
        stop[threadIdx.x] = 0u;
    }
    const int chain = sub.chain_of(blockIdx.x * kWideWsPairs + pair, p.n_points, ch.n_chains);
    const bool idle = chain < 0;
    const int cc = idle ? ch.n_chains - 1 : chain;
    // the step counter the chain will have AFTER the finish phase of this launch (both warps of the pair compute it
    // from the same loads, before the consumer rewrites the chain state)
    uint32_t t = ch.steps_done[cc];
    if ((flags & kWideFinish) && ws.ran[cc]) t += (uint32_t)s.steps_per_iteration;
    const uint32_t chain_id = ch.chain_id[cc];
    __syncthreads();
    if (idle) return;
    Ring ring;
    ring.data = smem_addr(&slots[pair][0][0]);
    ring.full = smem_addr(&bars[pair][0]);
    ring.empty = smem_addr(&bars[pair][kRingDepth]);
    ring.stop = smem_addr(&stop[pair]);
    ring.count = 0u;
    ring.dead = 0u;
    if (producer) {
        if (flags & kWideRun)
            produce_variates_wide<CPL>(p.n, (uint32_t)s.steps_per_iteration, k0, k1, chain_id, t, lane, ring);
        return;
    }
    wide_steps_body<CPL, kSrcRing>(p, ch, s, rec, ws, k0, k1, flags, it, n_iter, progress, peers, chain, stage[pair],
                                   ring);
    __syncwarp();
    if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(ring.stop), "r"(1u) : "memory");
}

// ----------------------------------------------------------------------------------------------
// Materialisation on the SIMT pipe: out = base + (double)(Lt (hi + lo)) for the 2B rows [W ; BW], one tile of
// kWideTileRows repetitions of one point per CTA, one thread per output coordinate.  Lt^T rows are read coalesced,
// once per tile; four interleaved partial sums over j (j = 4 jj + m), exactly materialise()'s order.
// ----------------------------------------------------------------------------------------------
template <int NPAD>
__global__ void __launch_bounds__(NPAD)
wide_materialise_simt_kernel(pb200_problem p, pb200_chains ch, WideWs ws, int reps) {
    constexpr int TM = kWideTileRows;
    __shared__ __align__(16) float wt[TM][NPAD];
    const int n_rt = (reps + TM - 1) / TM, P = p.n_points;
    const int rt = blockIdx.x % n_rt, pt = (blockIdx.x / n_rt) % P, which = blockIdx.x / (n_rt * P);
    const int i = threadIdx.x;
    const size_t b = (size_t)ch.n_chains, half = (size_t)which * b * NPAD;
#pragma unroll
    for (int r = 0; r < TM; ++r) {
        const int rep = rt * TM + r;
        const size_t src = half + ((size_t)rep * P + pt) * NPAD + i;
        wt[r][i] = rep < reps ? __fadd_rn(ws.w_hi[src], ws.w_lo[src]) : 0.0f;
    }
    __syncthreads();
    const float* __restrict__ col = p.lt_t + (size_t)pt * NPAD * NPAD + i;
    float a[TM][4];
#pragma unroll
    for (int r = 0; r < TM; ++r) a[r][0] = a[r][1] = a[r][2] = a[r][3] = 0.0f;
#pragma unroll 2
    for (int j = 0; j < NPAD / 4; ++j) {
        const float l0 = __ldg(col + (size_t)(4 * j + 0) * NPAD), l1 = __ldg(col + (size_t)(4 * j + 1) * NPAD);
        const float l2 = __ldg(col + (size_t)(4 * j + 2) * NPAD), l3 = __ldg(col + (size_t)(4 * j + 3) * NPAD);
#pragma unroll
        for (int r = 0; r < TM; ++r) {
            const float4 v = *reinterpret_cast<const float4*>(&wt[r][4 * j]);
            a[r][0] = __fmaf_rn(l0, v.x, a[r][0]);
            a[r][1] = __fmaf_rn(l1, v.y, a[r][1]);
            a[r][2] = __fmaf_rn(l2, v.z, a[r][2]);
            a[r][3] = __fmaf_rn(l3, v.w, a[r][3]);
        }
    }
#pragma unroll
    for (int r = 0; r < TM; ++r) {
        const int rep = rt * TM + r;
        if (rep >= reps) continue;
        const size_t c = ((size_t)rep * P + pt) * NPAD + i;
        const double x = __dadd_rn(ws.base[half + c],
                                   (double)__fadd_rn(__fadd_rn(a[r][0], a[r][1]), __fadd_rn(a[r][2], a[r][3])));
        if (which == 0) ch.x[c] = x;
        else ws.bx[c] = x;
        ws.q[half + c] = x - __ldg(p.mu + i);
    }
}

// Residuals of the start positions (the anchor GEMM before the first iteration of a call).
template <int NPAD>
__global__ void wide_residual_kernel(pb200_problem p, pb200_chains ch, WideWs ws) {
    const size_t total = (size_t)ch.n_chains * NPAD;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
        ws.q[idx] = ch.x[idx] - __ldg(p.mu + (idx % NPAD));
}

// ----------------------------------------------------------------------------------------------
// FP64 anchors as a batched GEMM.  Job classes (bit `cls` of cls_mask selects them), outputs in ws.v[cls]:
//   0: V  = (X  - mu) S_aa        -> chi^2 at x            (the epilogue runs in the steps kernel: chi2_of)
//   1: G  = (X  - mu) G_pt^T      -> whitened gradient gt = g + h_pt
//   2: Vb = (BX - mu) S_aa        -> chi^2 of the bestfit
// One CTA = kAnchorTileRows (32) repetitions of one point x kAnchorTileCols (64) outputs, 256 threads with 4 x 2 register
// tiles; both operands -- the residual rows Q and the matrix -- stream through shared memory in chunks of kWideKChunk
// k-values (cp.async, kAnchorStages-deep ring, one barrier per chunk).  Every output has ONE accumulator advanced in
// ascending k with one FMA per term -- the order of exact_eval(), hence the same bits.  tcgen05 has no FP64 type and
// DMMA's accumulation order is not specified, so this stays on the DFMA pipe, which is what bounds it: measured
// ~1 warp-DFMA per clock and SM in steady state (= 32 lanes: 18.6 TFLOP/s per B200); the small tiles (384 for the
// 8-point x 128-repetition batch, up to 4 resident per SM) are for load balance over the 148 SMs.
// ----------------------------------------------------------------------------------------------
constexpr int kAnchorTileCols = 64;
constexpr int kAnchorStages = 4;

template <int NPAD>
__global__ void __launch_bounds__(256, 2)
wide_anchor_kernel(pb200_problem p, pb200_chains ch, WideWs ws, int reps, int cls_mask, WideSub sub) {
    constexpr int TM = kAnchorTileRows, TN = kAnchorTileCols, KC = kWideKChunk, NCT = NPAD / TN, ST = kAnchorStages;
    constexpr int kStageDoubles = TM * KC + KC * TN;       // residual chunk [TM][KC] + matrix chunk [KC][TN]
    extern __shared__ __align__(16) unsigned char wide_smem[];
    double* stages = reinterpret_cast<double*>(wide_smem);
    const int n_rt = (reps + TM - 1) / TM, P = p.n_points;
    int bid = blockIdx.x;
    const int ct = bid % NCT; bid /= NCT;
    const int rt = bid % n_rt; bid /= n_rt;
    const int pt = sub.p0 + bid % sub.np;
    const int cls_idx = bid / sub.np;
    int cls = 0;
    for (int seen = 0; cls < 3; ++cls)
        if ((cls_mask >> cls) & 1) { if (seen == cls_idx) break; ++seen; }
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5, n = p.n;      // warp ty: rows 4 ty .. 4 ty + 3
    const size_t b = (size_t)ch.n_chains;
    const double* __restrict__ qsrc = ws.q + (cls == 2 ? b * NPAD : 0);
    const double* __restrict__ mat = (cls == 1 ? p.g_t + (size_t)pt * NPAD * NPAD : p.saa) + (size_t)ct * TN;
    // this thread's piece of a residual chunk: row tid / 8 of the tile, doubles 2 (tid % 8) .. + 1 of the chunk
    const int q_row = tid >> 3, q_rep = min(rt * TM + q_row, reps - 1);      // rows past the batch re-read the last one
    const double* q_piece = qsrc + ((size_t)q_rep * P + pt) * NPAD + 2 * (tid & 7);
    auto load_chunk = [&](int ck) {
        if (ck * KC < n) {
            double* st = stages + (size_t)(ck % ST) * kStageDoubles;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(st) + 16u * tid), "l"(q_piece + ck * KC)
                         : "memory");
            const uint32_t dst = smem_addr(st + TM * KC);
            for (int piece = tid; piece < KC * TN / 2; piece += 256) {
                const int kr = piece / (TN / 2), cc = piece % (TN / 2);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * piece),
                             "l"(mat + (size_t)(ck * KC + kr) * NPAD + 2 * cc) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");       // (an empty group keeps the group count uniform)
    };
    double acc[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = 0.0;
    const int n_chunks = (n + KC - 1) / KC;        // rows >= n of the padded matrices are never used
#pragma unroll
    for (int ck = 0; ck < ST - 1; ++ck) load_chunk(ck);
    for (int ck = 0; ck < n_chunks; ++ck) {
        asm volatile("cp.async.wait_group %0;" ::"n"(ST - 2) : "memory");     // chunk ck has landed (this thread's pieces)
        __syncthreads();                                                        // ... everyone's; chunk ck - 1 is consumed
        load_chunk(ck + ST - 1);                                                // refills the buffer of chunk ck - 1
        const double* st = stages + (size_t)(ck % ST) * kStageDoubles;
        const double* qq = st + (size_t)(4 * ty) * KC;
        const double* bb = st + TM * KC + 2 * tx;
        const int kk_end = min(KC, n - ck * KC);
#pragma unroll 4
        for (int kk = 0; kk < kk_end; ++kk) {
            const double2 b01 = *reinterpret_cast<const double2*>(bb + kk * TN);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double q = qq[r * KC + kk];
                acc[r][0] = fma(b01.x, q, acc[r][0]);
                acc[r][1] = fma(b01.y, q, acc[r][1]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int rep = rt * TM + 4 * ty + r;
        if (rep >= reps) continue;
        double* dst = ws.v + ((size_t)cls * b + (size_t)rep * P + pt) * NPAD + (size_t)ct * TN + 2 * tx;
        *reinterpret_cast<double2*>(dst) = make_double2(acc[r][0], acc[r][1]);
    }
}

}  // namespace pb200
