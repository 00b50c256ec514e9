// pipe_peak.cu -- microbenchmark of the integer pipes the match kernel lives on.
// Measures warp-instructions per clock per SM for POPC, LOP3, IADD3, SHF, IMAD
// and for the match kernel's per-cell instruction mix, with enough independent
// chains per thread to hide latency.  Output: one JSON line per test.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_peak pipe_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 4096

template <int MODE> __global__ void __launch_bounds__(256) k(unsigned *out, unsigned seed, unsigned long long *cycles) {
    unsigned x[CHAINS], w[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
        x[c] = seed * (threadIdx.x + 1) + c * 0x9e3779b9u;
        w[c] = c;
    }
    const unsigned s = seed | 1u;
    const int bias = -11;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (MODE == 0) {            // POPC chain
                x[c] = __popc(x[c]) + s;    // popc + iadd (iadd needed to keep a dependency with entropy)
            } else if (MODE == 1) {     // LOP3
                x[c] = (x[c] ^ s) & (x[c] | w[c]);
                asm volatile("" : "+r"(x[c]));
            } else if (MODE == 2) {     // IADD3
                x[c] = x[c] + s + w[c];
                asm volatile("" : "+r"(x[c]));
            } else if (MODE == 3) {     // SHF
                x[c] = __funnelshift_l(x[c], w[c], 1);
                asm volatile("" : "+r"(x[c]));
            } else if (MODE == 4) {     // IMAD
                x[c] = x[c] * s + w[c];
                asm volatile("" : "+r"(x[c]));
            } else if (MODE == 5) {     // match mix: xor, popc, iadd, shf (1 POPC per cell)
                const int y = __popc(x[c] ^ (s + it)) + bias;
                w[c] = __funnelshift_l((unsigned)y, w[c], 1);
            } else if (MODE == 6) {     // POPC only, independent inputs (no extra iadd): popc(x ^ it) accumulated via lop
                w[c] ^= __popc(x[c] ^ it);
            } else if (MODE == 8) {     // FADD (scalar)
                x[c] = __float_as_uint(__uint_as_float(x[c]) + 1.0009765625f);
                w[c] = __float_as_uint(__uint_as_float(w[c]) + 1.0009765625f);
            } else if (MODE == 9) {     // FADD2 (packed f32x2): one instruction for the pair
                float2 r = __fadd2_rn(make_float2(__uint_as_float(x[c]), __uint_as_float(w[c])),
                                      make_float2(1.0009765625f, 1.0009765625f));
                x[c] = __float_as_uint(r.x);
                w[c] = __float_as_uint(r.y);
            } else if (MODE == 10) {    // FFMA (scalar)
                x[c] = __float_as_uint(fmaf(__uint_as_float(x[c]), 0.999f, 1.0009765625f));
                w[c] = __float_as_uint(fmaf(__uint_as_float(w[c]), 0.999f, 1.0009765625f));
            } else if (MODE == 11) {    // FFMA2
                float2 r = __ffma2_rn(make_float2(__uint_as_float(x[c]), __uint_as_float(w[c])),
                                      make_float2(0.999f, 0.999f), make_float2(1.0009765625f, 1.0009765625f));
                x[c] = __float_as_uint(r.x);
                w[c] = __float_as_uint(r.y);
            } else if (MODE == 7) {     // mix with IMAD doing the add: popc, xor, imad (fma pipe), shf
                const int y = __popc(x[c] ^ (s + it)) * 1 + bias;
                w[c] = w[c] * 2u + ((unsigned)y >> 31);
            }
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc ^= x[c] ^ w[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE> void run(const char *name, double inst_per_iter_chain, int sms) {
    const int threads = 256, blocks_per_sm = 4;
    const int grid = sms * blocks_per_sm;
    unsigned *out;
    unsigned long long *cyc;
    cudaMalloc(&out, grid * threads * sizeof(unsigned));
    cudaMalloc(&cyc, grid * sizeof(unsigned long long));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 200; i++) k<MODE><<<grid, threads>>>(out, 12345u, cyc);   // ~0.1-0.2 s: clocks up
    cudaDeviceSynchronize();
    const int REP = 20;
    cudaEventRecord(e0);
    for (int i = 0; i < REP; i++) k<MODE><<<grid, threads>>>(out, 12345u, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= REP;
    unsigned long long *h = (unsigned long long *)malloc(grid * sizeof(unsigned long long));
    cudaMemcpy(h, cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < grid; i++) mean += (double)h[i];
    mean /= grid;
    // per SM: blocks_per_sm * threads/32 warps, each issuing CHAINS*ITERS "cells"
    const double cells_per_sm = (double)blocks_per_sm * threads * CHAINS * ITERS;   // thread-level ops
    const double cells_per_clk_sm = cells_per_sm / mean;
    const double total = cells_per_sm * sms;
    printf("{\"test\": \"%s\", \"thread_ops_per_clock64_per_sm\": %.2f, \"inst_per_op\": %.1f, \"ms\": %.4f, "
           "\"Gops_per_s\": %.1f, \"ops_per_clk_per_sm_at_1965MHz\": %.2f, \"mean_clock64\": %.0f, \"clock64_mhz\": %.0f}\n",
           name, cells_per_clk_sm, inst_per_iter_chain, ms, total / ms / 1e6, total / (ms * 1e-3) / sms / 1.965e9,
           mean, mean / ms / 1e3);
    cudaFree(out);
    cudaFree(cyc);
    free(h);
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) {
        printf("{\"error\": \"no device\"}\n");
        return 1;
    }
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
    run<0>("popc+iadd", 2, p.multiProcessorCount);
    run<6>("xor+popc+xor", 3, p.multiProcessorCount);
    run<1>("lop3x2", 2, p.multiProcessorCount);
    run<2>("iadd3", 1, p.multiProcessorCount);
    run<3>("shf", 1, p.multiProcessorCount);
    run<4>("imad", 1, p.multiProcessorCount);
    run<5>("match_mix xor+popc+iadd+shf", 4, p.multiProcessorCount);
    run<7>("match_mix xor+popc+imad+shf.imad", 4, p.multiProcessorCount);
    run<8>("fadd x2 (two scalar FADD per op)", 2, p.multiProcessorCount);
    run<9>("fadd2 (one packed FADD2 per op)", 1, p.multiProcessorCount);
    run<10>("ffma x2 (two scalar FFMA per op)", 2, p.multiProcessorCount);
    run<11>("ffma2 (one packed FFMA2 per op)", 1, p.multiProcessorCount);
    return 0;
}
