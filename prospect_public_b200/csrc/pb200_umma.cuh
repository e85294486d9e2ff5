// pb200_umma.cuh -- the materialisation GEMM of the wide path on the 5th-generation tensor cores (sm_100a):
//
//     [X ; BX] = [Xref ; Bref] + [W ; BW] Lt^T      per profile point: (2R x n) (n x n), FP32
//
// as tcgen05.mma kind::tf32 with the classic three-product split for FP32-level accuracy (3xTF32):
//     w = w_hi + w_lo,  Lt = Lt_hi + Lt_lo  (hi = the 10 explicit mantissa bits TF32 keeps, lo = the exact remainder)
//     W Lt^T ~= W_lo Lt_hi^T + W_hi Lt_lo^T + W_hi Lt_hi^T          (the dropped lo x lo term is ~2^-22 relative)
// accumulated in FP32 in tensor memory.  One CTA = one (point, best/last half, 128-repetition tile, 64-column tile):
// A = 128 rows of W (K-major, read with ONE 4-D TMA box per K-chunk straight out of the strided chain layout
// chain = rep * P + point), B = 64 rows of Lt (K-major), both landed by TMA in the 128-byte-swizzled layout the MMA
// descriptors name; thread 0 issues the TMA loads, thread 32 the MMAs (12 per 32-wide K-chunk), a ring of mbarriers
// between them; the 128-lane x 64-column FP32 accumulator is read back with tcgen05.ld and added to the FP64 base.
//
// Every wait is bounded; a time-out records a fault word (pb200_launch_status) and the CTA ends without writing.
#pragma once
#include <cuda.h>

namespace pb200 {

constexpr int kMmaTileM = 128, kMmaTileN = 64, kMmaChunkK = 32, kMmaStages = 4;
constexpr int kMmaStageBytes = 2 * kMmaTileM * 128 + 2 * kMmaTileN * 128;      // A_hi, A_lo, B_hi, B_lo
constexpr int kMmaSmemBytes = kMmaStages * kMmaStageBytes + 1024;               // + alignment slack

struct WideMma {
    bool enabled;
    int box_rows;                 // rows of one A box = min(128, repetitions)
    CUtensorMap w_hi, w_lo, lt_hi, lt_lo;
};

// Bounded wait: false (with the fault recorded for pb200_launch_status) if the barrier never completed.
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > PB200_WAIT_TIMEOUT_CYCLES) {
            atomicOr(&g_fault_word, kFaultPipe);
            return false;
        }
    }
    return true;
}

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t addr) {
    // K-major, SWIZZLE_128B: 8-row x 128-byte atoms, 1024 bytes between atoms along M/N (SBO), LBO unused,
    // descriptor version 1 (sm_100)
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <int NPAD>
__global__ void __launch_bounds__(128, 1)
wide_materialise_mma_kernel(pb200_chains ch, WideWs ws, const double* __restrict__ mu, int n_points, int reps,
                            int box_rows, WideSub sub,
                            const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                            const __grid_constant__ CUtensorMap map_lt_hi,
                            const __grid_constant__ CUtensorMap map_lt_lo) {
    constexpr int S = kMmaStages, KCH = NPAD / kMmaChunkK, NT = NPAD / kMmaTileN;
    extern __shared__ unsigned char umma_smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * S + 1];
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_mt = (reps + kMmaTileM - 1) / kMmaTileM;
    int bid = blockIdx.x;
    const int nt = bid % NT; bid /= NT;
    const int mt = bid % n_mt; bid /= n_mt;
    const int pt = sub.p0 + bid % sub.np;
    const int which = bid / sub.np;
    const uint32_t smem0 = (smem_addr(umma_smem_raw) + 1023u) & ~1023u;          // swizzle-128B atoms: 1024-byte aligned
    const uint32_t full0 = smem_addr(&bars[0]), empty0 = smem_addr(&bars[S]), done = smem_addr(&bars[2 * S]);
    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(full0 + 8u * i, 1u); mbar_init(empty0 + 8u * i, 1u); }
        mbar_init(done, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_slot)),
                     "r"((uint32_t)kMmaTileN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_slot;

    if (tid == 0) {                                   // ---- TMA producer ----
        const uint32_t bytes = (uint32_t)(2 * box_rows * 128 + 2 * kMmaTileN * 128);
        for (int c = 0; c < KCH; ++c) {
            const int s = c % S;
            if (c >= S && !mbar_wait_bounded(empty0 + 8u * s, (uint32_t)((c / S) - 1) & 1u)) break;
            const uint32_t st = smem0 + (uint32_t)s * kMmaStageBytes, bar = full0 + 8u * s;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            tma_load_4d(st, &map_w_hi, bar, c * kMmaChunkK, pt, mt * kMmaTileM, which);
            tma_load_4d(st + kMmaTileM * 128, &map_w_lo, bar, c * kMmaChunkK, pt, mt * kMmaTileM, which);
            tma_load_3d(st + 2 * kMmaTileM * 128, &map_lt_hi, bar, c * kMmaChunkK, nt * kMmaTileN, pt);
            tma_load_3d(st + 2 * kMmaTileM * 128 + kMmaTileN * 128, &map_lt_lo, bar, c * kMmaChunkK, nt * kMmaTileN, pt);
        }
    } else if (tid == 32) {                           // ---- MMA issuer ----
        // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 64, M = 128
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kMmaTileN >> 3) << 17) |
                                   ((uint32_t)(kMmaTileM >> 4) << 24);
        for (int c = 0; c < KCH; ++c) {
            const int s = c % S;
            if (!mbar_wait_bounded(full0 + 8u * s, (uint32_t)(c / S) & 1u)) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t st = smem0 + (uint32_t)s * kMmaStageBytes;
            const uint64_t a_hi = umma_smem_desc(st), a_lo = umma_smem_desc(st + kMmaTileM * 128);
            const uint64_t b_hi = umma_smem_desc(st + 2 * kMmaTileM * 128);
            const uint64_t b_lo = umma_smem_desc(st + 2 * kMmaTileM * 128 + kMmaTileN * 128);
#pragma unroll
            for (int k = 0; k < kMmaChunkK / 8; ++k) {            // UMMA_K = 8 for TF32: 32 bytes along K per MMA
                const uint64_t adv = (uint64_t)(2 * k);           // (32 bytes >> 4) inside the swizzle atom
                umma_tf32(tmem, a_lo + adv, b_hi + adv, idesc, (c | k) != 0);
                umma_tf32(tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                umma_tf32(tmem, a_hi + adv, b_hi + adv, idesc, 1u);
            }
            umma_commit(empty0 + 8u * s);                         // frees the stage when these MMAs have read it
        }
        umma_commit(done);            // (after a time-out above this still arrives: nothing is in flight then)
    }
    __syncwarp();
    const bool ok = mbar_wait_bounded(done, 0u) && (*(volatile uint32_t*)&g_fault_word & kFaultPipe) == 0u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: TMEM lane = row of the tile; warp w reads lanes 32 w .. 32 w + 31 ----
    uint32_t acc[kMmaTileN];
#pragma unroll
    for (int part = 0; part < kMmaTileN / 16; ++part) {
        const uint32_t addr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(16 * part);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(acc[16 * part + 0]), "=r"(acc[16 * part + 1]), "=r"(acc[16 * part + 2]), "=r"(acc[16 * part + 3]),
              "=r"(acc[16 * part + 4]), "=r"(acc[16 * part + 5]), "=r"(acc[16 * part + 6]), "=r"(acc[16 * part + 7]),
              "=r"(acc[16 * part + 8]), "=r"(acc[16 * part + 9]), "=r"(acc[16 * part + 10]), "=r"(acc[16 * part + 11]),
              "=r"(acc[16 * part + 12]), "=r"(acc[16 * part + 13]), "=r"(acc[16 * part + 14]), "=r"(acc[16 * part + 15])
            : "r"(addr));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    // accumulator rows -> shared memory (the pipeline stages are free now), then row-contiguous, coalesced global
    // access: a warp handles one row per pass, lane l the columns 2 l, 2 l + 1 of the 64-column tile
    float* tile = reinterpret_cast<float*>(umma_smem_raw + (smem0 - smem_addr(umma_smem_raw)));      // [128][65]
    {
        float* mine = tile + (32 * warp + lane) * (kMmaTileN + 1);
#pragma unroll
        for (int c = 0; c < kMmaTileN; ++c) mine[c] = __uint_as_float(acc[c]);
    }
    __syncthreads();
    if (ok) {
        const size_t b = (size_t)ch.n_chains;
        const double2 m = __ldg(reinterpret_cast<const double2*>(mu + (size_t)nt * kMmaTileN) + lane);
        for (int r = warp; r < kMmaTileM; r += 4) {
            const int rep = mt * kMmaTileM + r;
            if (rep >= reps) break;
            const size_t row = ((size_t)rep * n_points + pt) * NPAD + (size_t)nt * kMmaTileN;
            const double2 bv = reinterpret_cast<const double2*>(ws.base + (size_t)which * b * NPAD + row)[lane];
            const float* a = tile + r * (kMmaTileN + 1) + 2 * lane;
            const double2 x = make_double2(__dadd_rn(bv.x, (double)a[0]), __dadd_rn(bv.y, (double)a[1]));
            reinterpret_cast<double2*>((which == 0 ? ch.x : ws.bx) + row)[lane] = x;
            reinterpret_cast<double2*>(ws.q + (size_t)which * b * NPAD + row)[lane] =
                make_double2(x.x - m.x, x.y - m.y);            // the anchor GEMM's A operand
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kMmaTileN) : "memory");
}

// ---- host side: tensor maps through the runtime's driver entry point (no link-time dependency on libcuda) ----
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn tensor_map_encoder() {
    static encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult status;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &status) == cudaSuccess &&
            status == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)ptr;
    }
    return fn;
}

static int encode_f32_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                          const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    encode_tiled_fn enc = tensor_map_encoder();
    if (!enc) return fail("cuTensorMapEncodeTiled is not available");
    const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    const CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims,
                            strides_bytes, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)rc));
    return 0;
}

// PB200_WIDE_MMA=0 keeps the materialisation on the SIMT pipe (bit-identical to the fused kernel).
template <int NPAD>
static int wide_mma_prepare(const pb200_problem& prob, const pb200_chains& chains, const WideWs& ws, int reps,
                            WideMma* out) {
    out->enabled = false;
    if (!prob.lt_hi || !prob.lt_lo) return 0;
    if (const char* env = std::getenv("PB200_WIDE_MMA"))
        if (env[0] == '0') return 0;
    const cuuint64_t b = (cuuint64_t)chains.n_chains, P = (cuuint64_t)prob.n_points;
    out->box_rows = reps < kMmaTileM ? reps : kMmaTileM;
    {   // W: [2][reps][P][NPAD] floats (chain = rep * P + point), box = 32 k x 1 point x box_rows reps x 1 half
        const cuuint64_t dims[4] = {(cuuint64_t)NPAD, P, (cuuint64_t)reps, 2};
        const cuuint64_t strides[3] = {(cuuint64_t)NPAD * 4, P * NPAD * 4, b * NPAD * 4};
        const cuuint32_t box[4] = {(cuuint32_t)kMmaChunkK, 1, (cuuint32_t)out->box_rows, 1};
        if (encode_f32_map(&out->w_hi, ws.w_hi, 4, dims, strides, box)) return 1;
        if (encode_f32_map(&out->w_lo, ws.w_lo, 4, dims, strides, box)) return 1;
    }
    {   // Lt: [P][NPAD (i)][NPAD (j)] floats, K-major (j contiguous), box = 32 k x 64 rows x 1 point
        const cuuint64_t dims[3] = {(cuuint64_t)NPAD, (cuuint64_t)NPAD, P};
        const cuuint64_t strides[2] = {(cuuint64_t)NPAD * 4, (cuuint64_t)NPAD * NPAD * 4};
        const cuuint32_t box[3] = {(cuuint32_t)kMmaChunkK, (cuuint32_t)kMmaTileN, 1};
        if (encode_f32_map(&out->lt_hi, prob.lt_hi, 3, dims, strides, box)) return 1;
        if (encode_f32_map(&out->lt_lo, prob.lt_lo, 3, dims, strides, box)) return 1;
    }
    static bool attr_set[2] = {false, false};
    if (!attr_set[NPAD == 256]) {
        if (check_cuda(cudaFuncSetAttribute(wide_materialise_mma_kernel<NPAD>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, kMmaSmemBytes),
                       "cudaFuncSetAttribute (wide_materialise_mma_kernel)"))
            return 1;
        attr_set[NPAD == 256] = true;
    }
    out->enabled = true;
    return 0;
}

template <int NPAD>
static int wide_mma_launch(const pb200_problem& prob, const pb200_chains& chains, const WideWs& ws, int reps,
                           const WideMma& mma, cudaStream_t st, WideSub sub) {
    const int n_mt = (reps + kMmaTileM - 1) / kMmaTileM;
    const int grid = 2 * sub.np * n_mt * (NPAD / kMmaTileN);
    wide_materialise_mma_kernel<NPAD><<<grid, 128, kMmaSmemBytes, st>>>(chains, ws, prob.mu, prob.n_points, reps,
                                                                        mma.box_rows, sub,
                                                                        mma.w_hi, mma.w_lo, mma.lt_hi, mma.lt_lo);
    return check_cuda(cudaGetLastError(), "wide_materialise_mma_kernel launch");
}

}  // namespace pb200
