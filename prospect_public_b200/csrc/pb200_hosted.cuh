// pb200_hosted.cuh -- the batched proposal / accept machinery for HOST-evaluated likelihoods.
//
// External kernels (cobaya, MontePython, any BaseKernel subclass: prospect/kernels/base_kernel.py:9-138) compute their
// likelihood on the CPU.  Everything else of MetropolisHastings.step (prospect/mcmc.py:163-190) is the same for every
// kernel and runs here, for all chains of a batch at once:
//   hosted_propose_kernel   x' = x + step * F z,  F F^T = covmat of the chain's profile point  (mcmc.py:182-190),
//                           z from the chain's Philox stream (the numbers draw_block hands the fused kernels), and
//                           the prior-box test of x' (base_kernel.py:89-102);
//   (the host evaluates -log L of the proposals that are inside the box)
//   hosted_accept_kernel    accept iff exp((L_cur - L_prop) / T) > u  <=>  L_prop - L_cur < T e, e = -ln u ~ Exp(1)
//                           (mcmc.py:170-180; an out-of-box proposal is rejected without looking at its likelihood,
//                           :166-167), position / loglkl update, accepted count, running bestfit (first minimum,
//                           optimiser.py:75), step counter.
// One warp per chain; FP64 throughout (the likelihood call on the host dominates, nothing here is worth FP32).
#pragma once
#include "pb200_device.cuh"

namespace pb200 {

constexpr int kHostedWarps = 4;

// N(0,1) of coordinate `coord` at step t of a chain: element t % 4 of the Philox block t / 4 (draw_block's layout).
__device__ __forceinline__ float hosted_normal(uint32_t k0, uint32_t k1, uint32_t chain_id, uint32_t t, uint32_t coord) {
    uint32_t w[4];
    philox4x32_10(t >> 2, coord, chain_id, 0u, k0, k1, w);
    float a, b;
    const bool second = (t & 2u) != 0u;
    box_muller(second ? w[2] : w[0], second ? w[3] : w[1], a, b);
    return (t & 1u) ? b : a;
}

// Exp(1) accept variate of step t.
__device__ __forceinline__ float hosted_exp_variate(uint32_t k0, uint32_t k1, uint32_t chain_id, uint32_t t) {
    uint32_t w[4];
    philox4x32_10(t >> 2, kAcceptSlot, chain_id, 0u, k0, k1, w);
    const uint32_t k = t & 3u;
    return exp_variate(k == 0u ? w[0] : (k == 1u ? w[1] : (k == 2u ? w[2] : w[3])));
}

__global__ void __launch_bounds__(32 * kHostedWarps)
hosted_propose_kernel(pb200_hosted h, pb200_chains ch, uint32_t k0, uint32_t k1, double* __restrict__ prop,
                      int32_t* __restrict__ outside) {
    extern __shared__ double hosted_stage[];          // [kHostedWarps][npad]: the whitened draw of each warp's chain
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int chain = blockIdx.x * kHostedWarps + wib;
    if (chain >= ch.n_chains) return;                 // (whole warp)
    double* zs = hosted_stage + (size_t)wib * h.npad;
    const bool live = ch.status[chain] == PB200_ACTIVE;
    const uint32_t t = ch.steps_done[chain], id = ch.chain_id[chain];
    const double step = ch.step_size[chain];
    for (int j = lane; j < h.npad; j += 32)
        zs[j] = (live && j < h.n) ? (double)hosted_normal(k0, k1, id, t, (uint32_t)j) : 0.0;
    __syncwarp();
    const double* __restrict__ f = h.factor_t + (size_t)ch.point[chain] * h.npad * h.npad;
    bool out = false;
    for (int i = lane; i < h.npad; i += 32) {
        double acc = 0.0;                             // (F z)_i, j ascending, one FMA per term
        for (int j = 0; j < h.n; ++j) acc = fma(__ldg(f + (size_t)j * h.npad + i), zs[j], acc);
        const double xi = ch.x[(size_t)chain * h.npad + i];
        const double xp = i < h.n ? fma(step, acc, xi) : 0.0;
        prop[(size_t)chain * h.npad + i] = xp;
        if (i < h.n) out = out || (xp < __ldg(h.lo + i)) || (xp > __ldg(h.hi + i));
    }
    const bool any_out = __any_sync(kFull, out);
    if (lane == 0) outside[chain] = (live && any_out) ? 1 : 0;
}

__global__ void __launch_bounds__(32 * kHostedWarps)
hosted_accept_kernel(pb200_hosted h, pb200_chains ch, uint32_t k0, uint32_t k1, const double* __restrict__ prop,
                     const double* __restrict__ prop_loglkl, const int32_t* __restrict__ outside,
                     int32_t* __restrict__ accepted_flag, int32_t* __restrict__ n_accepted,
                     double* __restrict__ best_loglkl, double* __restrict__ best_x) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int chain = blockIdx.x * kHostedWarps + wib;
    if (chain >= ch.n_chains) return;
    if (ch.status[chain] != PB200_ACTIVE) {
        if (lane == 0 && accepted_flag) accepted_flag[chain] = 0;
        return;
    }
    const uint32_t t = ch.steps_done[chain];
    const double e = (double)hosted_exp_variate(k0, k1, ch.chain_id[chain], t);
    const double ll = ch.loglkl[chain], llp = prop_loglkl[chain];
    // NaN / +inf likelihoods (a computation exception of the host kernel, base_kernel.py:78-81) fail the comparison
    const bool acc = outside[chain] == 0 && (__dadd_rn(llp, -ll) < __dmul_rn(ch.temperature[chain], e));
    const bool better = acc && best_loglkl != nullptr && llp < best_loglkl[chain];
    __syncwarp();                                     // every lane has read the scalars lane 0 is about to overwrite
    if (acc) {
        for (int i = lane; i < h.npad; i += 32) {
            const double v = prop[(size_t)chain * h.npad + i];
            ch.x[(size_t)chain * h.npad + i] = v;
            if (better && best_x) best_x[(size_t)chain * h.npad + i] = v;
        }
    }
    if (lane == 0) {
        if (acc) {
            ch.loglkl[chain] = llp;
            if (n_accepted) n_accepted[chain] += 1;
            if (better) best_loglkl[chain] = llp;
        }
        if (accepted_flag) accepted_flag[chain] = acc ? 1 : 0;
        ch.steps_done[chain] = t + 1u;
    }
}

}  // namespace pb200
