// Micro-benchmark: per-scheduler (SMSP) throughput of the instruction classes the annealing kernel is made of.
// Each test runs ILP independent dependency chains per thread, 8 warps per scheduler, and reports warp-instructions
// per cycle per scheduler.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe scripts/pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ILP = 8, ITERS = 4096;

template <int OP>
__global__ void probe(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t a[ILP];
    float f[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = seed + threadIdx.x * 977u + i * 131u; f[i] = 1.0f + (float)(a[i] & 1023u) * 1e-3f; }
    const float c1 = 1.0001f + seed * 1e-9f, c2 = 0.5f + seed * 1e-9f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) f[i] = __fmaf_rn(f[i], c1, c2);
            if (OP == 1) { uint64_t p = (uint64_t)a[i] * 0xD2511F53u; a[i] = (uint32_t)(p >> 32) ^ (uint32_t)p; }   // IMAD.WIDE + LOP3
            if (OP == 2) a[i] = a[i] * 0xD2511F53u + seed;                                                   // IMAD (lo)
            if (OP == 3) a[i] = (a[i] ^ seed) | (a[i] >> 3);                                                 // LOP3/SHF
            if (OP == 4) { asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i])); }
            if (OP == 5) f[i] = __uint2float_rn(__float_as_uint(f[i]));                                      // I2F
            if (OP == 6) f[i] = __shfl_xor_sync(0xffffffffu, f[i], 4);
            if (OP == 7) f[i] = __fadd_rn(f[i], c2);
            if (OP == 8) { uint64_t p = (uint64_t)a[i] * 0xD2511F53u; a[i] = (uint32_t)(p >> 32) + (uint32_t)p; }   // IMAD.WIDE + IADD
            if (OP == 9) a[i] = __umulhi(a[i], 0xD2511F53u) + seed;                                          // IMAD.HI
            if (OP == 10) { f[i] = __fmaf_rn(f[i], c1, c2); a[i] = (a[i] ^ seed) | (a[i] >> 3); }            // FFMA + 2 alu
            if (OP == 11) { asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(f[i])); }
            if (OP == 12) f[i] = __fmul_rn(f[i], c1);
            if (OP == 13) f[i] = a[i] & 1 ? f[i] : c1;                                                         // FSEL-like
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= a[i] ^ __float_as_uint(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter, uint32_t* out, long long* cyc) {
    const int blocks = 148, threads = 1024;      // 32 warps per SM = 8 per scheduler
    probe<OP><<<blocks, threads>>>(out, cyc, 12345u);
    probe<OP><<<blocks, threads>>>(out, cyc, 12345u);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < 148; ++i) mean += (double)h[i] / 148;
    const double winstr = (double)ITERS * ILP * per_iter * 8;   // warp instructions per scheduler
    printf("%-28s cycles %10.0f   warp-instr/cycle/SMSP %.3f   (cycles per warp-instr %.2f)\n", name, mean, winstr / mean,
           mean / winstr);
}

int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    run<0>("FFMA", 1, out, cyc);
    run<7>("FADD", 1, out, cyc);
    run<12>("FMUL", 1, out, cyc);
    run<1>("IMAD.WIDE + LOP3", 2, out, cyc);
    run<8>("IMAD.WIDE + IADD", 2, out, cyc);
    run<9>("IMAD.HI (+IADD?)", 1, out, cyc);
    run<2>("IMAD lo", 1, out, cyc);
    run<3>("LOP3 + SHF", 2, out, cyc);
    run<4>("MUFU.LG2", 1, out, cyc);
    run<11>("MUFU.SIN (+FMUL)", 2, out, cyc);
    run<5>("I2F", 1, out, cyc);
    run<6>("SHFL", 1, out, cyc);
    run<10>("FFMA + LOP3 + SHF", 3, out, cyc);
    run<13>("LOP3 + ISETP + FSEL", 3, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
