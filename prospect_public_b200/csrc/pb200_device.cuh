// pb200_device.cuh -- device building blocks shared by the sampler kernels (sm_100a).
//
// Mapping: a GROUP of LPC lanes (32, 16 or 8) per chain, G = 32/LPC chains per warp.  Lane l of a group
// owns coordinates l, l+LPC, ... (CPL per lane, npad = LPC*CPL).  A chain is strictly sequential in steps,
// so the parallelism inside a step is what fills the SM: the n normals and the O(n) whitened quadratic
// form are spread over the group's lanes and reduced with xor-shuffles that never leave the group.
// Small groups make every warp instruction serve several chains (fewer instructions per chain-step);
// large groups give more warps for a small batch.  pb200_choose_lanes picks LPC from the batch size.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pb200.h"

namespace pb200 {

#ifndef PB200_WARPS_PER_CTA
#define PB200_WARPS_PER_CTA 4
#endif
constexpr int kWarpsPerCta = PB200_WARPS_PER_CTA;   // warps never synchronise with each other: small CTAs only balance the grid
constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kAcceptSlot = 0xffffffffu;   // Philox counter word of the accept/reject variates

// ----------------------------------------------------------------------------------------------
// Philox-4x32-10 (Salmon et al., SC'11).  Counter = (step_block, slot, chain_id, 0), key = seed.
// slot = coordinate index for the proposal normals, kAcceptSlot for the accept variates, so the
// stream of a chain depends only on (seed, chain_id, step) -- not on the thread mapping, the launch
// partitioning or the number of GPUs.
// ----------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        // one 32x32->64 multiply per word (IMAD.WIDE.U32) yields both halves
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 32 random bits -> uniform in (0, 1]:  u = w * 2^-32 + 2^-33 (rounds to 1.0f for the top words).
__device__ __forceinline__ float u01(uint32_t w) {
    return __fmaf_rn(__uint2float_rn(w), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

// log2 of a uniform in [2^-33, 1]: never subnormal, so the .ftz form (one MUFU.LG2, no denormal fix-up) is exact enough.
__device__ __forceinline__ float lg2_fast(float u) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(u));
    return r;
}

// Box-Muller on the MUFU pipe: r = sqrt(-2 ln u1), theta = 2 pi u2 - pi.
__device__ __forceinline__ void box_muller(uint32_t w0, uint32_t w1, float& z0, float& z1) {
    const float u1 = u01(w0), u2 = u01(w1);
    float r;                                                          // sqrt(-2 ln2 * log2(u1)), one MUFU.SQRT
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * lg2_fast(u1)));
    float s, c;
    __sincosf(__fmaf_rn(u2, 6.283185307179586f, -3.141592653589793f), &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// Exp(1) variate e = -ln u; the MH test exp(-Delta/T) > u becomes Delta < T * e.
__device__ __forceinline__ float exp_variate(uint32_t w) {
    return -0.6931471805599453f * lg2_fast(u01(w));
}

// Reductions over the LPC lanes of a group (xor distances < LPC stay inside the group).
template <int LPC>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int m = LPC / 2; m >= 1; m >>= 1) v = __fadd_rn(v, __shfl_xor_sync(kFull, v, m));
    return v;
}
template <int LPC>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int m = LPC / 2; m >= 1; m >>= 1) v = __dadd_rn(v, __shfl_xor_sync(kFull, v, m));
    return v;
}
template <int LPC>
__device__ __forceinline__ double group_min(double v) {
#pragma unroll
    for (int m = LPC / 2; m >= 1; m >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, m));
    return v;
}
template <int LPC>
__device__ __forceinline__ bool group_any(bool pred, int gi) {
    const unsigned b = __ballot_sync(kFull, pred);
    if (LPC == 32) return b != 0u;
    return ((b >> (gi * LPC)) & ((1u << (LPC & 31)) - 1u)) != 0u;
}

// Variates of the 4 steps of Philox block `blk` for this lane's CPL coordinates.
//   z[r][k]: N(0,1) for coordinate gl+LPC*r at step 4*blk+k (0 for pads);  e[k]: Exp(1) accept variate.
// When n < npad the spare coordinate n carries the accept slot, otherwise every lane makes one
// extra (identical) call -- both give the same numbers.
template <int LPC, int CPL>
__device__ __forceinline__ void draw_block(uint32_t k0, uint32_t k1, uint32_t chain_id, uint32_t blk, int n,
                                           int gl, float (&z)[CPL][4], float (&e)[4]) {
    constexpr int NPAD = LPC * CPL;
    const int acc_r = n / LPC, acc_lane = n % LPC;
    uint32_t w[4];
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        const int coord = gl + LPC * r;
        philox4x32_10(blk, coord == n ? kAcceptSlot : (uint32_t)coord, chain_id, 0u, k0, k1, w);
        box_muller(w[0], w[1], z[r][0], z[r][1]);
        box_muller(w[2], w[3], z[r][2], z[r][3]);
        if (coord >= n) { z[r][0] = z[r][1] = z[r][2] = z[r][3] = 0.0f; }
        if (r == acc_r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) e[k] = __shfl_sync(kFull, exp_variate(w[k]), acc_lane, LPC);
        }
    }
    if (n >= NPAD) {
        philox4x32_10(blk, kAcceptSlot, chain_id, 0u, k0, k1, w);
#pragma unroll
        for (int k = 0; k < 4; ++k) e[k] = exp_variate(w[k]);
    }
}

// ----------------------------------------------------------------------------------------------
// Shared-memory ring between a PRODUCER warp (Philox + Box-Muller) and the CONSUMER warp that advances the chains
// (anneal_ws_kernel).  One slot = the variates of one Philox block for the 32 lanes of the pair: float4 [CPL][32]
// normals (lane-major, so both sides touch their own 16 bytes: conflict-free LDS.128 / STS.128) + float4 [G] accept
// variates.  full[slot] / empty[slot] are mbarriers with 32 arrivals (every lane arrives for its own accesses).
// ----------------------------------------------------------------------------------------------
constexpr int kRingDepth = 4;
#ifndef PB200_MBAR_SUSPEND_NS
#define PB200_MBAR_SUSPEND_NS 20000u      // suspend-time hint of a pending mbarrier wait
#endif

__device__ __forceinline__ uint32_t smem_addr(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
#if PB200_MBAR_SUSPEND_NS > 0
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"     // suspends up to the hint, no busy spin
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(PB200_MBAR_SUSPEND_NS) : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"         // system-default suspend time
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#endif
    return ok != 0u;
}

struct Ring {
    uint32_t data, full, empty, stop;   // shared-space addresses: slots, mbarriers [kRingDepth] each, stop word
    uint32_t count;                     // blocks produced / consumed so far by this side
    uint32_t dead;                      // consumer: the producer stopped delivering (fault recorded)
};

// Device-side protocol faults (pb200_launch_status reads and clears the word).
constexpr uint32_t kFaultRing = 1u;     // a producer warp never filled its ring slot
constexpr uint32_t kFaultPipe = 2u;     // a TMA / tcgen05 pipeline barrier of the wide path never completed
__device__ uint32_t g_fault_word = 0u;
#ifndef PB200_WAIT_TIMEOUT_CYCLES
#define PB200_WAIT_TIMEOUT_CYCLES 4000000000LL      // ~2 s at 1.9 GHz
#endif

template <int CPL>
__device__ __forceinline__ constexpr int ring_slot_bytes() { return (CPL + 1) * 32 * 16; }

// Consumer side: wait for the next block, copy it to registers, hand the slot back.  A producer that never delivers
// would hang the GPU, so the wait is bounded: on a time-out the fault is recorded (pb200_launch_status turns it into
// an error on the host), the ring is marked dead and every later fetch returns variates that reject every proposal
// (z = 0, e = 0 -> Delta = 0 is not < 0); the caller stops the chains with status PB200_FAILED.
template <int LPC, int CPL>
__device__ __forceinline__ void ring_fetch(Ring& rg, int lane, int gi, float (&z)[CPL][4], float (&e)[4]) {
    const uint32_t slot = rg.count % kRingDepth, parity = (rg.count / kRingDepth) & 1u;
    const uint32_t bar = rg.full + 8u * slot;
    if (!rg.dead && !__all_sync(kFull, mbar_try_wait(bar, parity))) {
        const long long t0 = clock64();
        while (!__all_sync(kFull, mbar_try_wait(bar, parity))) {
            if (__any_sync(kFull, clock64() - t0 > PB200_WAIT_TIMEOUT_CYCLES)) {
                if (lane == 0) atomicOr(&g_fault_word, kFaultRing);
                rg.dead = 1u;
                break;
            }
        }
    }
    if (rg.dead) {
#pragma unroll
        for (int r = 0; r < CPL; ++r) z[r][0] = z[r][1] = z[r][2] = z[r][3] = 0.0f;
        e[0] = e[1] = e[2] = e[3] = 0.0f;
        return;
    }
    const uint32_t base = rg.data + slot * (uint32_t)ring_slot_bytes<CPL>();
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(z[r][0]), "=f"(z[r][1]), "=f"(z[r][2]), "=f"(z[r][3])
                     : "r"(base + (uint32_t)(r * 32 + lane) * 16u));
    }
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(e[0]), "=f"(e[1]), "=f"(e[2]), "=f"(e[3])
                 : "r"(base + (uint32_t)(CPL * 32 + gi) * 16u));
    mbar_arrive(rg.empty + 8u * slot);
    rg.count += 1;
}

// Producer side.  Returns false when the consumer has raised the stop word (it left its iteration loop early).
template <int LPC, int CPL>
__device__ __forceinline__ bool ring_store(Ring& rg, int lane, int gi, int gl, const float (&z)[CPL][4],
                                           const float (&e)[4]) {
    const uint32_t slot = rg.count % kRingDepth, parity = ((rg.count / kRingDepth) & 1u) ^ 1u;
    const uint32_t bar = rg.empty + 8u * slot;
    while (!__all_sync(kFull, mbar_try_wait(bar, parity))) {
        uint32_t stop;
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(stop) : "r"(rg.stop));
        if (__any_sync(kFull, stop != 0u)) return false;
    }
    const uint32_t base = rg.data + slot * (uint32_t)ring_slot_bytes<CPL>();
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)(r * 32 + lane) * 16u),
                     "f"(z[r][0]), "f"(z[r][1]), "f"(z[r][2]), "f"(z[r][3]) : "memory");
    }
    if (gl == 0) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)(CPL * 32 + gi) * 16u),
                     "f"(e[0]), "f"(e[1]), "f"(e[2]), "f"(e[3]) : "memory");
    }
    mbar_arrive(rg.full + 8u * slot);
    rg.count += 1;
    return true;
}

// acc[r] += sum_j m[j][gl + LPC r] * qs[j], j ascending, one FMA per term (the order the oracle uses).  The loads of
// four rows are issued before their FMAs: the loop is bound by the latency of the L2-resident matrix rows (one
// 8 * npad byte row per j), not by FP64 throughput, so memory-level parallelism is what counts.
template <int LPC, int CPL>
__device__ __forceinline__ void matvec_rows(const double* __restrict__ m, const double* qs, int n, double (&acc)[CPL]) {
    // (measured: pays for the wide geometries, 256-D profile 4.55 -> 4.40 ms; costs 4 % at npad = 32, where the 30
    // short rows are L1 hits and the extra live registers hurt the step loop -- so exact_eval only uses it for
    // rows of >= 1 KB)
    constexpr int NPAD = LPC * CPL, UN = 4;
    int j = 0;
    for (; j + UN <= n; j += UN) {
        double a[UN][CPL];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
#pragma unroll
            for (int r = 0; r < CPL; ++r) a[u][r] = __ldg(m + (size_t)(j + u) * NPAD + LPC * r);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const double qj = qs[j + u];
#pragma unroll
            for (int r = 0; r < CPL; ++r) acc[r] = fma(a[u][r], qj, acc[r]);
        }
    }
    for (; j < n; ++j) {
        const double qj = qs[j];
#pragma unroll
        for (int r = 0; r < CPL; ++r) acc[r] = fma(__ldg(m + (size_t)j * NPAD + LPC * r), qj, acc[r]);
    }
}

// ----------------------------------------------------------------------------------------------
// Exact FP64 evaluation at a position: loglkl = sum_i q_i (1/2 (S_aa q)_i + f s_ab_i) + c0 and,
// optionally, the whitened gradient gt = G q + h.  `qs` is this group's [npad] double staging row.
// ----------------------------------------------------------------------------------------------
template <int LPC, int CPL, bool WITH_GRAD>
__device__ __forceinline__ double exact_eval(const pb200_problem& p, int pt, const double (&x)[CPL], int gl,
                                             double* qs, double (&gt)[CPL]) {
    constexpr int NPAD = LPC * CPL;
    double q[CPL], v[CPL], g[CPL];
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        const int c = gl + LPC * r;
        q[r] = x[r] - __ldg(p.mu + c);
        qs[c] = q[r];
        v[r] = 0.0;
        g[r] = 0.0;
    }
    __syncwarp();
    if constexpr (NPAD >= 128) {      // long rows: batched loads, one matrix at a time (matvec_rows)
        matvec_rows<LPC, CPL>(p.saa + gl, qs, p.n, v);
        if (WITH_GRAD) matvec_rows<LPC, CPL>(p.g_t + (size_t)pt * NPAD * NPAD + gl, qs, p.n, g);
    } else {                          // short rows (L1 hits): both matrices in one pass
        const double* __restrict__ saa = p.saa + gl;
        const double* __restrict__ gtm = p.g_t + (size_t)pt * NPAD * NPAD + gl;
        for (int j = 0; j < p.n; ++j) {
            const double qj = qs[j];
#pragma unroll
            for (int r = 0; r < CPL; ++r) {
                v[r] = fma(__ldg(saa + j * NPAD + LPC * r), qj, v[r]);
                if (WITH_GRAD) g[r] = fma(__ldg(gtm + j * NPAD + LPC * r), qj, g[r]);
            }
        }
    }
    __syncwarp();
    const double f = __ldg(p.f + pt);
    double part = 0.0;
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        const int c = gl + LPC * r;
        part = fma(q[r], fma(0.5, v[r], f * __ldg(p.sab + c)), part);
        if (WITH_GRAD) gt[r] = g[r] + __ldg(p.h + (size_t)pt * NPAD + c);
    }
    return group_sum<LPC>(part) + __ldg(p.c0 + pt);
}

// x_out = base + Lt * w for this group: w staged through `ws` (float [npad]), FP32 matvec with four
// interleaved partial sums, one FP64 add.  Used whenever a position must be materialised (see run_steps).
template <int LPC, int CPL>
__device__ __forceinline__ void materialise(const float* __restrict__ lt_pt, int gl, const double (&base)[CPL],
                                            const float (&w)[CPL], float* ws, double (&x_out)[CPL]) {
    constexpr int NPAD = LPC * CPL;
#pragma unroll
    for (int r = 0; r < CPL; ++r) ws[gl + LPC * r] = w[r];
    __syncwarp();
    const float4* __restrict__ w4 = reinterpret_cast<const float4*>(ws);
    float a[CPL][4];
#pragma unroll
    for (int r = 0; r < CPL; ++r) a[r][0] = a[r][1] = a[r][2] = a[r][3] = 0.0f;
    const float* __restrict__ col = lt_pt + gl;
#pragma unroll 4
    for (int j = 0; j < NPAD / 4; ++j) {
        const float4 v = w4[j];
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            a[r][0] = __fmaf_rn(__ldg(col + (4 * j + 0) * NPAD + LPC * r), v.x, a[r][0]);
            a[r][1] = __fmaf_rn(__ldg(col + (4 * j + 1) * NPAD + LPC * r), v.y, a[r][1]);
            a[r][2] = __fmaf_rn(__ldg(col + (4 * j + 2) * NPAD + LPC * r), v.z, a[r][2]);
            a[r][3] = __fmaf_rn(__ldg(col + (4 * j + 3) * NPAD + LPC * r), v.w, a[r][3]);
        }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < CPL; ++r)
        x_out[r] = __dadd_rn(base[r], (double)__fadd_rn(__fadd_rn(a[r][0], a[r][1]), __fadd_rn(a[r][2], a[r][3])));
}

// The same for warps that hold SEVERAL chains (LPC < 32) of a 32-coordinate problem, when only some of them need their
// position: in `materialise` every lane group runs the full n x n mat-vec whenever ANY group of the warp asks for it
// (the call sits under warp-uniform control), so one accepting chain makes its three neighbours do -- and wait for --
// 128 L1 loads + FMAs per lane.  Here the warp works through the asking chains one at a time, all 32 lanes on ONE
// chain: lane i forms output coordinate i (32 coalesced loads + FMAs, the same four interleaved partial sums over j,
// hence the same bits) and the sums travel back to the owning lanes by shuffle.  ncu (65 536 chains, wide step): the
// per-group form was 62 % of the MH kernel's stall samples and held the LSU pipe at 65 %.
//   want: this group's position is needed (others get x_out = base);  stage rows of the warp's groups are contiguous.
template <int LPC, int CPL>
__device__ __forceinline__ void materialise_selected(const float* __restrict__ lt_pt, int gl, int gi, bool want,
                                                     const double (&base)[CPL], const float (&w)[CPL], float* ws,
                                                     double (&x_out)[CPL]) {
    constexpr int NPAD = LPC * CPL, G = 32 / LPC;
    static_assert(NPAD == 32 && LPC < 32, "one output coordinate per lane");
    const int lane = gi * LPC + gl;
#pragma unroll
    for (int r = 0; r < CPL; ++r) ws[gl + LPC * r] = w[r];
    __syncwarp();
    const unsigned asking = __ballot_sync(kFull, want);
    float sum[CPL];
#pragma unroll
    for (int r = 0; r < CPL; ++r) sum[r] = 0.0f;
    const unsigned long long my_lt = (unsigned long long)(uintptr_t)lt_pt;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (((asking >> (g * LPC)) & 1u) == 0u) continue;                      // warp-uniform
        const float* __restrict__ lt_g =
            (const float*)(uintptr_t)__shfl_sync(kFull, my_lt, g * LPC);         // the chain's point may differ per group
        // stage row of group g: rows are [NPAD] doubles apart
        const float4* __restrict__ w4 = reinterpret_cast<const float4*>(ws + (g - gi) * NPAD * 2);
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
        for (int j = 0; j < NPAD / 4; ++j) {
            const float4 v = w4[j];
            a0 = __fmaf_rn(__ldg(lt_g + (4 * j + 0) * NPAD + lane), v.x, a0);
            a1 = __fmaf_rn(__ldg(lt_g + (4 * j + 1) * NPAD + lane), v.y, a1);
            a2 = __fmaf_rn(__ldg(lt_g + (4 * j + 2) * NPAD + lane), v.z, a2);
            a3 = __fmaf_rn(__ldg(lt_g + (4 * j + 3) * NPAD + lane), v.w, a3);
        }
        const float s = __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            const float mine = __shfl_sync(kFull, s, gl + LPC * r);              // coordinate gl + LPC r lives on that lane
            if (gi == g) sum[r] = mine;
        }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < CPL; ++r) x_out[r] = __dadd_rn(base[r], (double)sum[r]);
}

// Squared whitened radius inside which base + Lt*w provably stays in the prior box:
//   |x_i - base_i| <= lt_norm_i * |w|  (Cauchy-Schwarz)  =>  safe while |w| < min_i dist_i / lt_norm_i.
// 0.999 absorbs FP32 rounding of |w|^2 and of the later materialisation.  +inf without a box.
template <int LPC, int CPL>
__device__ __forceinline__ float safe_radius2(const pb200_problem& p, int pt, int gl, const double (&base)[CPL]) {
    constexpr int NPAD = LPC * CPL;
    double m = INFINITY;
#pragma unroll
    for (int r = 0; r < CPL; ++r) {
        const int c = gl + LPC * r;
        const double nrm = (double)__ldg(p.lt_norm + (size_t)pt * NPAD + c);
        const double dist = fmin(__ldg(p.hi + c) - base[r], base[r] - __ldg(p.lo + c));
        if (nrm > 0.0) m = fmin(m, dist / nrm);
    }
    m = group_min<LPC>(m);
    m = fmax(m, 0.0) * 0.999;
    return (float)(m * m);
}

// ----------------------------------------------------------------------------------------------
// Adaptive step-size / temperature bookkeeping of one chain after an iteration
// (prospect/optimiser.py:87-162 + tasks/optimise_task.py:33-54).  Plain IEEE double ops in the
// reference's order, no FMA contraction, so the result is bit-identical to the Python arithmetic.
// ----------------------------------------------------------------------------------------------
struct SchedState {
    double temperature, step_size, current_best;
    double prev_mult, cur_mult, prev_ar;
    int32_t stage, status, iteration;
};

__host__ __device__ inline double pb_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
__host__ __device__ inline double pb_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
__host__ __device__ inline double pb_div(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b; return r;
#endif
}

__host__ __device__ inline void schedule_next(const pb200_schedule& s, SchedState& st, int32_t accepted,
                                              double best_loglkl) {
    st.iteration += 1;
    // chi2_tolerance stop: emit_tasks returns [] when the bestfit moved by <= tol and something was accepted
    if (s.chi2_tolerance > 0.0) {
        const double improvement = pb_add(st.current_best, -best_loglkl);
        const bool keep_going = (improvement < 0.0) || (st.current_best == (double)INFINITY) ||
                                (pb_mul(2.0, improvement) > s.chi2_tolerance);
        if (!keep_going && accepted >= 1) {
            st.status = PB200_CONVERGED;
            return;
        }
    }
    const double ar = pb_div((double)(1 + accepted), (double)(1 + s.steps_per_iteration));
    const double new_temp = pb_mul(st.temperature, s.temperature_change);
    double new_step = st.step_size;
    if (s.step_mode == PB200_STEP_EXPONENTIAL) {
        new_step = pb_mul(st.step_size, s.step_size_change);
    } else if (s.step_mode == PB200_STEP_ADAPTIVE_FIXED) {
        if (ar > s.adaptive_hi) new_step = pb_mul(st.step_size, pb_add(1.0, s.adaptive_multiplier));
        else if (ar < s.adaptive_lo) new_step = pb_mul(st.step_size, pb_add(1.0, -s.adaptive_multiplier));
    } else {
        if (st.stage < 2) {
            const double mag = st.stage == 0 ? 0.05 : 0.1;
            double m;
            if (ar > s.adaptive_hi) m = mag;
            else if (ar < s.adaptive_lo) m = -mag;
            else { st.status = PB200_FAILED; return; }        // unbound local in the reference
            new_step = pb_mul(st.step_size, pb_add(1.0, m));
            if (st.stage == 0) { st.prev_mult = m; st.prev_ar = ar; }
            else               { st.cur_mult = m; }
            st.stage += 1;
        } else {
            const double ar_desired = pb_div(pb_add(s.adaptive_lo, s.adaptive_hi), 2.0);
            double m = pb_add(pb_mul(pb_div(pb_add(ar_desired, -ar), pb_add(ar, -st.prev_ar)),
                                     pb_add(st.cur_mult, -st.prev_mult)), st.cur_mult);
            if (1.0 < m) m = 1.0;            // min(m_new, 1.0): NaN stays NaN
            if (!(m > -0.9)) m = -0.9;       // max(-0.9, m_new): NaN -> -0.9
            new_step = pb_mul(st.step_size, pb_add(1.0, m));
            st.prev_ar = ar;
            st.prev_mult = st.cur_mult;
            st.cur_mult = m;
        }
    }
    st.current_best = best_loglkl;
    st.temperature = new_temp;
    st.step_size = new_step;
}

}  // namespace pb200
