// tmem_bw.cu -- throughput of tcgen05.ld / tcgen05.st (32x32b shape) as K1 uses them: 16 warps per
// SM, every warp reading (writing) its own 32 TMEM lanes, 16 or 64 columns per instruction.
// Answers whether TMEM can also hold K1's constant tables (window, twiddles) next to the parked
// half-frames: that would add ~256 registers of TMEM reads per frame and lane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: ld x16, 1: ld x64 (4 x16 then one wait), 2: st x16, 3: ld x16 + st x16 alternating
__global__ void __launch_bounds__(512, 1) tmem_kernel(float *out, int iters) {
    __shared__ unsigned s_tmem;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned base = s_tmem + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)(warp >> 2) * 128u;
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = threadIdx.x + i;
    // initialise the columns
#pragma unroll
    for (int c = 0; c < 8; c++)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     ::"r"(base + 16 * c), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]),
                       "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    float acc = 0.f;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            if (MODE == 0 || MODE == 1 || (MODE == 3 && (c & 1) == 0)) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
                               "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
                             : "r"(base + 16 * c) : "memory");
                if (MODE != 1 || (c & 3) == 3) {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(r[0]), "+f"(r[5]), "+f"(r[15]) :: "memory");
                    acc += r[0] + r[5] + r[15];
                }
            } else {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                             ::"r"(base + 16 * c), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]),
                               "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]) : "memory");
                if (MODE == 3) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
        }
        if (MODE == 2) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u) : "memory");
}

template <int MODE> static void run(const char *name, int sms, float *out, int clock_khz) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) tmem_kernel<MODE><<<sms, 512>>>(out, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        tmem_kernel<MODE><<<sms, 512>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double bytes_per_sm = 16.0 * iters * 8 * 16 * 4 * 32;   // warps x iters x instrs x regs x 4 B x lanes
    printf("{\"test\": \"%s\", \"ms\": %.3f, \"bytes_per_clk_per_sm_at_%d_MHz\": %.1f, \"err\": \"%s\"}\n", name, best,
           clock_khz / 1000, bytes_per_sm / (best * 1e-3) / (clock_khz * 1e3), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    float *out;
    cudaMalloc(&out, sizeof(float) * prop.multiProcessorCount * 512);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, prop.multiProcessorCount, clock_khz);
    run<0>("tcgen05.ld 32x32b.x16, wait after each", prop.multiProcessorCount, out, clock_khz);
    run<1>("tcgen05.ld 32x32b.x16, wait after four", prop.multiProcessorCount, out, clock_khz);
    run<2>("tcgen05.st 32x32b.x16", prop.multiProcessorCount, out, clock_khz);
    run<3>("ld x16 / st x16 alternating", prop.multiProcessorCount, out, clock_khz);
    return 0;
}
