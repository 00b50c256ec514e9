// k1_mix_peak.cu -- the floating-point instruction mix of K1 (fp_fft_chroma) with
// nothing else: no shared memory, no integer work, no shuffles.  One "frame" here is
// the FP32 instruction count one frame costs in the real kernel (dynamic SASS counts of
// an `ncu --set full` capture, per warp: FADD2 / FFMA2 / FMUL2 packed, FADD / FMUL / FFMA
// scalar), issued from 16 independent register chains per thread, 16 warps per SM.
// The frames per second this loop reaches is the ceiling the FMA pipe sets for that
// mix: the measured denominator of K1's roofline (bench.py: roofline.peak).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o k1_mix_peak k1_mix_peak.cu
//   ./k1_mix_peak [fadd2 ffma2 fmul2 fadd fmul ffma]      (counts per frame-warp)
// Output: JSON lines.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int CHAINS = 16;
constexpr int UNIT_A2 = 13, UNIT_F2 = 11, UNIT_M2 = 9;   // packed ops per unit (649 : 535 : 462): round 1's h32 kernel
constexpr int UNIT_A1 = 4, UNIT_M1 = 1, UNIT_F1 = 1;     // scalar ops per unit (189 : 53 : 41 + part of the rest)
// the mix of the current default kernel (revision 4: 258 FADD2, 639 FFMA2, 100 FMUL2, 128 FADD, 104 FMUL, 393 FFMA)
constexpr int V4_A2 = 5, V4_F2 = 13, V4_M2 = 2, V4_A1 = 3, V4_M1 = 2, V4_F1 = 8;

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { return ((u64)__float_as_uint(b) << 32) | __float_as_uint(a); }
// one SASS instruction each, whatever the compiler thinks of the operands
__device__ __forceinline__ void a2(u64 &x, u64 c) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(c)); }
__device__ __forceinline__ void m2(u64 &x, u64 c) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(c)); }
__device__ __forceinline__ void f2(u64 &x, u64 c, u64 d) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(c), "l"(d)); }
__device__ __forceinline__ void a1(float &x, float c) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(c)); }
__device__ __forceinline__ void m1(float &x, float c) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(c)); }
__device__ __forceinline__ void f1(float &x, float c, float d) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(c), "f"(d)); }

// MODE 0: the mix; 1: FADD2 only; 2: FFMA2 only; 3: scalar FFMA only; 4: packed + scalar FFMA alternating
template <int MODE>
__global__ void __launch_bounds__(512, 1) mix_kernel(float2 *out, int iters, float seed) {
    u64 x[CHAINS];
    float y[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
        x[c] = pk(seed + c + threadIdx.x * 1e-3f, seed - c);
        y[c] = seed + 0.5f * c;
    }
    const u64 ca = pk(1.0e-3f * seed, -1.0e-3f * seed), cm = pk(0.99951171875f * seed, 1.00048828125f * seed);
    const u64 cf = pk(0.999755859375f * seed, 0.999755859375f * seed);
    const float sa = 1.0e-3f * seed, sm = 0.99951171875f * seed;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
            // 33 packed + 6 scalar ops.  Each chain sees one kind of operation only (chains 0-5 add,
            // 6-10 fma, 11-15 mul): ptxas contracts a mul.rn.f32x2 followed by an add.rn.f32x2 of the
            // same register into one FFMA2, which would change the mix.
            int ca_ = 0, cf_ = 0, cm_ = 0, d = 0;
#pragma unroll
            for (int r = 0; r < UNIT_A2; r++) {
                a2(x[ca_], ca); ca_ = (ca_ + 1) % 6;
                if (r < UNIT_F2) { f2(x[6 + cf_], cf, ca); cf_ = (cf_ + 1) % 5; }
                if (r < UNIT_M2) { m2(x[11 + cm_], cm); cm_ = (cm_ + 1) % 5; }
                if (r < UNIT_A1) { a1(y[d], sa); d++; }
                if (r == 5) { m1(y[d], sm); d++; }
                if (r == 7) { f1(y[d], sm, sa); d++; }
            }
        } else if (MODE == 5) {
            // 20 packed + 13 scalar ops in the proportions of the revision-4 kernel; chains 0-3 add, 4-12 fma, 13-15 mul
            int ca_ = 0, cf_ = 0, cm_ = 0, d = 0;
#pragma unroll
            for (int r = 0; r < V4_F2; r++) {
                f2(x[4 + cf_], cf, ca); cf_ = (cf_ + 1) % 9;
                if (r < V4_A2) { a2(x[ca_], ca); ca_ = (ca_ + 1) % 4; }
                if (r < V4_M2) { m2(x[13 + cm_], cm); cm_ = (cm_ + 1) % 3; }
                if (r < V4_F1) { f1(y[d], sm, sa); d++; }
                if (r < V4_A1) { a1(y[d], sa); d++; }
                if (r < V4_M1) { m1(y[d], sm); d++; }
            }
        } else {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) {
                if (MODE == 1) a2(x[c], ca);
                if (MODE == 2) f2(x[c], cf, ca);
                if (MODE == 3) { f1(y[c], sm, sa); }
                if (MODE == 4) { if (c & 1) f2(x[c], cf, ca); else f1(y[c], sm, sa); }
            }
        }
    }
    u64 acc = 0;
    float accy = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { acc ^= x[c]; accy += y[c]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(__uint_as_float((unsigned)acc) + accy, __uint_as_float((unsigned)(acc >> 32)));
}

template <int MODE> static double time_ms(int sms, int iters, float2 *out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) mix_kernel<MODE><<<sms, 512>>>(out, iters, 1.0f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        mix_kernel<MODE><<<sms, 512>>>(out, iters, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv) {
    // dynamic FP32 instruction counts of one frame-warp of K1 (profiles/: ncu source page of the kernel)
    double n[6] = {649, 535, 462, 189, 53, 41};
    for (int i = 0; i < 6 && i + 1 < argc; i++) n[i] = atof(argv[i + 1]);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    float2 *out;
    cudaMalloc(&out, sizeof(float2) * sms * 512);
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, clock_khz);
    const int iters = 20000;
    const double warps = (double)sms * 16;
    {
        const double ms = time_ms<0>(sms, iters, out);
        const double packed_unit = UNIT_A2 + UNIT_F2 + UNIT_M2, scalar_unit = UNIT_A1 + UNIT_M1 + UNIT_F1;
        const double packed_frame = n[0] + n[1] + n[2], scalar_frame = n[3] + n[4] + n[5];
        // pipe cycles: a packed op holds a scheduler's FMA pipe for two cycles, a scalar op for one
        const double unit_cycles = 2 * packed_unit + scalar_unit, frame_cycles = 2 * packed_frame + scalar_frame;
        const double units_per_s = warps * iters / (ms * 1e-3);
        const double frames_per_s = units_per_s * unit_cycles / frame_cycles;
        const double pipe_cycles_per_s = units_per_s * unit_cycles;   // warp-level pipe cycles per second, whole GPU
        printf("{\"test\": \"k1 fp mix\", \"ms\": %.4f, \"packed_per_frame\": %.0f, \"scalar_per_frame\": %.0f, "
               "\"fma_pipe_cycles_per_frame_warp\": %.0f, \"frames_per_s_ceiling\": %.4e, "
               "\"pipe_cycles_per_clk_per_scheduler_at_%d_MHz\": %.4f, \"fp32_lane_ops_per_clk_per_sm\": %.2f}\n",
               ms, packed_frame, scalar_frame, frame_cycles, frames_per_s, clock_khz / 1000,
               pipe_cycles_per_s / (sms * 4.0) / (clock_khz * 1e3),
               pipe_cycles_per_s * 32.0 / sms / (clock_khz * 1e3));
    }
    {
        const double ms = time_ms<5>(sms, iters, out);
        const double unit_cycles = 2.0 * (V4_A2 + V4_F2 + V4_M2) + (V4_A1 + V4_M1 + V4_F1);
        const double frame_cycles = 2.0 * (258 + 639 + 100) + (128 + 104 + 393);
        const double units_per_s = warps * iters / (ms * 1e-3);
        const double pipe_cycles_per_s = units_per_s * unit_cycles;
        printf("{\"test\": \"k1 fp mix, revision-4 kernel\", \"ms\": %.4f, \"packed_per_frame\": 997, \"scalar_per_frame\": 625, "
               "\"fma_pipe_cycles_per_frame_warp\": %.0f, \"frames_per_s_ceiling\": %.4e, "
               "\"pipe_cycles_per_clk_per_scheduler_at_%d_MHz\": %.4f, \"fp32_lane_ops_per_clk_per_sm\": %.2f}\n",
               ms, frame_cycles, units_per_s * unit_cycles / frame_cycles, clock_khz / 1000,
               pipe_cycles_per_s / (sms * 4.0) / (clock_khz * 1e3), pipe_cycles_per_s * 32.0 / sms / (clock_khz * 1e3));
    }
    const char *names[] = {"", "fadd2 only", "ffma2 only", "ffma scalar only", "ffma2 + ffma alternating"};
    double msv[5];
    msv[1] = time_ms<1>(sms, iters, out);
    msv[2] = time_ms<2>(sms, iters, out);
    msv[3] = time_ms<3>(sms, iters, out);
    msv[4] = time_ms<4>(sms, iters, out);
    for (int m = 1; m <= 4; m++) {
        const double insts = 1.0 * CHAINS;   // warp instructions per iteration
        const double per_s = warps * iters * insts / (msv[m] * 1e-3);
        printf("{\"test\": \"%s\", \"ms\": %.4f, \"warp_inst_per_clk_per_scheduler_at_%d_MHz\": %.4f}\n", names[m], msv[m],
               clock_khz / 1000, per_s / (sms * 4.0) / (clock_khz * 1e3));
    }
    cudaFree(out);
    return 0;
}
