// fingerprint.cu -- K1/K2: Chromaprint-style frame hashing on sm_100a.
//
// Replaces the chromaprint::Context that Analyzer::process_frames drives
// (needle/src/audio/analyzer.rs:176,218,275,286-301: start / feed / finish /
// get_fingerprint_raw) plus the subsample-and-stamp tail (:288-318).  The
// algorithm is Chromaprint 1.5.x TEST2 (crate chromaprint-sys-next 1.5.3,
// needle/Cargo.lock:158-159); constants in fp_tables.h.
//
// K1 fp_fft_chroma: one WARP per 4096-sample frame (hop 1365).  The real FFT is
// a 2048-point complex FFT of z[n] = x[2n] + i x[2n+1] plus a split step:
//   pass 1  lane L holds z[32 m + L], m = 0..63 (coalesced loads, window
//           fused); 64-point FFT over m entirely in registers (8 x 8)
//   xchg    one transpose through shared memory (padded, conflict-free)
//   pass 2  lane L holds k1 in {L, L+32}: twiddle, two 32-point FFTs in
//           registers (8 x 4)  ->  Z[L + 32 t], t = 0..63
//   split   Z[2048-k] lives in lane (32-L)&31: one warp shuffle per value;
//           power |X[k]|^2 for k in [10, 1308), folded into 12 pitch classes
// Output: 12 floats per frame.  FP32 throughout.
//
// K2 fp_classify: FIR {.25,.75,1,.75,.25} over time, L2 normalisation, the
// 16 Haar-like classifiers on each 16-row window via a tile-local integral
// image, 3-threshold quantisation, Gray code, packing, stride-s emission and
// the timestamp.  All in FP64: it is <1 % of the work, and FP32 would flip
// hashes near quantiser thresholds.  log((1+a)/(1+b)) < t is evaluated as
// (1+a) < e^t (1+b).
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "common.h"
#include "fp_tables.h"
#include "tma.cuh"

namespace nb200 {

constexpr int FP_WARPS = 8;                 // frames in flight per CTA
constexpr int FP_HALF = FP_FRAME / 2;       // 2048-point complex FFT
constexpr int XCHG_STRIDE = 33;             // float2 row stride of the transpose buffer
constexpr int XCHG_ELEMS = 64 * XCHG_STRIDE;
constexpr int FP_KMIN = 10, FP_KMAX = 1308; // chroma bins [10, 1308)
constexpr int FP_T = 41;                    // k = L + 32 t covers [0, 1312)
constexpr int K2_TILE = 128;                // raw hashes per CTA
constexpr int K2_ROWS = K2_TILE + FP_WINDOW_ROWS - 1;   // 143 normalised rows
constexpr int K2_COLS = FP_BANDS + 1;

__constant__ float2 c_w64[64];    // exp(-2 pi i e / 64)
__constant__ float2 c_w32[32];    // exp(-2 pi i e / 32)
__constant__ float2 c_w128[FP_T]; // exp(-2 pi i t / 128) = W_4096^(32 t)
// the same twiddles as (-w.y, w.x): second operand of the packed complex multiply
__constant__ float2 c_w64r[64], c_w32r[32], c_w128r[FP_T];

struct K2Classifier {
    int type, y, h, w;
    double e0, e1, e2;   // exp(t0), exp(t1), exp(t2)
};
__constant__ K2Classifier c_cls[FP_NUM_CLASSIFIERS];
__constant__ double c_fir[FP_FIR_LEN];

uint64_t num_frames(uint64_t n_mono) {
    return n_mono >= (uint64_t)FP_FRAME ? (n_mono - FP_FRAME) / FP_HOP + 1 : 0;
}
static uint64_t num_raw(uint64_t n_mono) {
    const uint64_t f = num_frames(n_mono);
    return f > (uint64_t)FP_WARMUP ? f - FP_WARMUP : 0;
}

// ------------------------------------------------- packed complex arithmetic
// A complex number is one float2 = one aligned register pair; sm_100a has packed
// FP32x2 instructions (FADD2 / FMUL2 / FFMA2) with broadcast and swap operand
// modifiers, so a complex add is ONE instruction and a complex multiply two or
// three.  Half the issue slots of scalar code for the same flops.
using f2 = float2;
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 bx(f2 a) { return make_float2(a.x, a.x); }
__device__ __forceinline__ f2 by(f2 a) { return make_float2(a.y, a.y); }
__device__ __forceinline__ f2 swp(f2 a) { return make_float2(a.y, a.x); }
// a * w for a run-time w
__device__ __forceinline__ f2 cmul(f2 a, f2 w) { return fma2(by(a), make_float2(-w.y, w.x), mul2(bx(a), w)); }
// a * w in three packed instructions and nothing else: P = a.x * w, Q = a.y * (w.y, w.x),
// result = P + (-Q.x, Q.y) -- the (-1, 1) factor becomes an operand modifier.  cmul() above
// needs a scalar negation and a register move to build (-w.y, w.x) from a run-time w.
__device__ __forceinline__ f2 cmul3(f2 a, f2 w) {
    const f2 q = mul2(by(a), swp(w));
    return fma2(q, make_float2(-1.f, 1.f), mul2(bx(a), w));
}
// a * w with wr = (-w.y, w.x) supplied (tables of constants carry both forms)
__device__ __forceinline__ f2 cmulc(f2 a, f2 w, f2 wr) { return fma2(by(a), wr, mul2(bx(a), w)); }
// d + (-i) t   and   d - (-i) t
__device__ __forceinline__ f2 add_mi(f2 d, f2 t) { return fma2(swp(t), make_float2(1.f, -1.f), d); }
__device__ __forceinline__ f2 sub_mi(f2 d, f2 t) { return fma2(swp(t), make_float2(-1.f, 1.f), d); }

// Forward DFTs on register arrays; S = element stride.  Natural order in and out.
template <int S> __device__ __forceinline__ void fft4(f2 *v) {
    const f2 c0 = v[0], c1 = v[S], c2 = v[2 * S], c3 = v[3 * S];
    const f2 d0 = add2(c0, c2), d1 = sub2(c0, c2), d2 = add2(c1, c3), e = sub2(c1, c3);
    v[0] = add2(d0, d2);
    v[S] = add_mi(d1, e);
    v[2 * S] = sub2(d0, d2);
    v[3 * S] = sub_mi(d1, e);
}

template <int S> __device__ __forceinline__ void fft8(f2 *v) {
    constexpr float R = 0.70710678118654752440f;
    const f2 a0 = v[0], a1 = v[S], a2 = v[2 * S], a3 = v[3 * S];
    const f2 a4 = v[4 * S], a5 = v[5 * S], a6 = v[6 * S], a7 = v[7 * S];
    // X[2r] = FFT4(a_j + a_{j+4}),  X[2r+1] = FFT4((a_j - a_{j+4}) W8^j)
    const f2 b0 = add2(a0, a4), b1 = add2(a1, a5), b2 = add2(a2, a6), b3 = add2(a3, a7);
    const f2 b4 = sub2(a0, a4), t5 = sub2(a1, a5), e6 = sub2(a2, a6), t7 = sub2(a3, a7);
    // t5 (1-i)/sqrt2 = t5.x (R,-R) + t5.y (R,R);   t7 (-1-i)/sqrt2 = t7.x (-R,-R) + t7.y (R,-R)
    const f2 b5 = fma2(by(t5), make_float2(R, R), mul2(bx(t5), make_float2(R, -R)));
    const f2 b7 = fma2(by(t7), make_float2(R, -R), mul2(bx(t7), make_float2(-R, -R)));
    {
        const f2 d0 = add2(b0, b2), d1 = sub2(b0, b2), d2 = add2(b1, b3), e = sub2(b1, b3);
        v[0] = add2(d0, d2);
        v[2 * S] = add_mi(d1, e);
        v[4 * S] = sub2(d0, d2);
        v[6 * S] = sub_mi(d1, e);
    }
    {
        // b6 = (-i) e6 is folded into the butterflies
        const f2 d0 = add_mi(b4, e6), d1 = sub_mi(b4, e6), d2 = add2(b5, b7), e = sub2(b5, b7);
        v[S] = add2(d0, d2);
        v[3 * S] = add_mi(d1, e);
        v[5 * S] = sub2(d0, d2);
        v[7 * S] = sub_mi(d1, e);
    }
}

// 64-point FFT, n = n1 + 8 n2, k = 8 k1 + k2: X[8 k1 + k2] ends up at v[k1 + 8 k2].
__device__ __forceinline__ void fft64(f2 *v) {
#pragma unroll
    for (int n1 = 0; n1 < 8; n1++) fft8<8>(v + n1);
#pragma unroll
    for (int n1 = 1; n1 < 8; n1++)
#pragma unroll
        for (int k2 = 1; k2 < 8; k2++)
            v[n1 + 8 * k2] = cmulc(v[n1 + 8 * k2], c_w64[n1 * k2], c_w64r[n1 * k2]);
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) fft8<1>(v + 8 * k2);
}
__device__ __forceinline__ constexpr int pos64(int k) { return (k >> 3) + 8 * (k & 7); }

// 32-point FFT, n = n1 + 8 n2 (n2 < 4), k = 4 k1 + k2: X[4 k1 + k2] at v[k1 + 8 k2].
__device__ __forceinline__ void fft32(f2 *v) {
#pragma unroll
    for (int n1 = 0; n1 < 8; n1++) fft4<8>(v + n1);
#pragma unroll
    for (int n1 = 1; n1 < 8; n1++)
#pragma unroll
        for (int k2 = 1; k2 < 4; k2++)
            v[n1 + 8 * k2] = cmulc(v[n1 + 8 * k2], c_w32[n1 * k2], c_w32r[n1 * k2]);
#pragma unroll
    for (int k2 = 0; k2 < 4; k2++) fft8<1>(v + 8 * k2);
}
__device__ __forceinline__ constexpr int pos32(int k) { return (k >> 2) + 8 * (k & 3); }
// register slot of Z[L + 32 t]: k1 = L + 32 (t & 1), k2 = t >> 1
__device__ __forceinline__ constexpr int zslot(int t) { return (t & 1) * 32 + pos32(t >> 1); }

// two packed i16 -> two floats without I2F: 0x4B00hhhh is 2^23 + hhhh
__device__ __forceinline__ f2 s16x2_to_float(uint32_t w) {
    w ^= 0x80008000u;   // bias to unsigned
    const f2 biased = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)),
                                  __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)));
    return add2(biased, make_float2(-8421376.0f, -8421376.0f));
}

// the same through the conversion unit: two I2F.S16 (one per half of the word) on the XU pipe,
// which nothing else in K1 uses -- no bias trick, no work for the FMA pipe
__device__ __forceinline__ f2 s16x2_to_float_xu(uint32_t w) {
    return make_float2((float)(short)(w & 0xffffu), (float)(short)(w >> 16));
}

// (1-D bulk async copy + mbarrier helpers: tma.cuh)
// float2 store to shared memory as ONE st.shared.v2.f32 whose operands are the value's own
// registers.  The C++ form (`buf[i] = packed_result`) makes ptxas copy every result into a
// staging register pair first: two MOVs per store, ~17 % of this kernel's instructions.
__device__ __forceinline__ void sts2(uint32_t addr, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

struct FpSegment {
    uint64_t pcm_off;      // element offset of the segment's mono PCM
    uint32_t frame_begin;  // first global frame index
    uint32_t n_frames;
    uint32_t n_raw;        // raw sub-fingerprints = n_frames - 19
    uint32_t tile_begin;   // first K2 tile
    uint32_t out_off;      // element offset into the hashset arrays
    uint32_t pad;
    uint64_t seek_to_ns;
};

struct K1Args {
    const int16_t *pcm;
    const FpSegment *segs;
    uint32_t n_segs;
    uint32_t frame_begin, frame_end;   // this launch covers global frames [begin, end)
    const float2 *window;    // [2048]  (w[2n], w[2n+1]) / 32767
    const float2 *twiddle;   // [32][64] W_2048^(n2 k1)
    const float2 *lane_w;    // [32] W_4096^L
    const float2 *wsplit;    // [41][32] W_4096^(L + 32 t)
    const float2 *window_odd;   // [1056] (w[2n-1], w[2n]) / 32767, entry 0 = (w[4095], w[0]): frames rotated by one sample
    float *chroma;           // [total_frames][12]
};

constexpr int PCM_COPY_BYTES = FP_FRAME * 2 + 16;   // the frame plus alignment slack, multiple of 16
constexpr int K1_WARP_BYTES = XCHG_ELEMS * 8;       // 16,896 B: transpose buffer, doubles as PCM landing zone
static_assert(PCM_COPY_BYTES <= K1_WARP_BYTES, "PCM must fit the exchange buffer");
constexpr int K1_SMEM_BYTES = 2 * FP_HALF * 8 + FP_WARPS * K1_WARP_BYTES + FP_WARPS * 8;

// address of frame `frame`'s first sample (binary search over the segment table)
__device__ __forceinline__ const int16_t *frame_src(const K1Args &a, uint32_t frame) {
    uint32_t lo = 0, hi = a.n_segs;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.segs[mid].frame_begin <= frame) lo = mid; else hi = mid;
    }
    return a.pcm + a.segs[lo].pcm_off + (uint64_t)(frame - a.segs[lo].frame_begin) * FP_HOP;
}

__global__ void __launch_bounds__(FP_WARPS * 32, 1) fp_fft_chroma_kernel(const K1Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    f2 *s_window = reinterpret_cast<f2 *>(smem_raw);                     // 2048
    f2 *s_twiddle = s_window + FP_HALF;                                  // 2048
    unsigned char *s_warp = reinterpret_cast<unsigned char *>(s_twiddle + FP_HALF);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_warp + FP_WARPS * K1_WARP_BYTES);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < FP_HALF; i += blockDim.x) {
        s_window[i] = a.window[i];
        s_twiddle[i] = a.twiddle[i];
    }
    if (threadIdx.x < FP_WARPS) mbar_init(smem_u32(s_bar + threadIdx.x), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    f2 *xchg = reinterpret_cast<f2 *>(s_warp + warp * K1_WARP_BYTES);
    const uint32_t *pcm_words = reinterpret_cast<const uint32_t *>(xchg);
    const uint32_t bar = smem_u32(s_bar + warp);
    const uint32_t xchg_addr = smem_u32(xchg);
    const int partner = (32 - lane) & 31;
    const f2 lane_w = a.lane_w[lane];

    const uint32_t n_blocks = (a.frame_end - a.frame_begin + FP_WARPS - 1) / FP_WARPS;
    uint32_t blk = blockIdx.x;
    uint32_t frame = a.frame_begin + blk * FP_WARPS + warp;
    uint32_t phase = 0;
    // misalignment (bytes, even, < 16) of the frame that is in flight / resident
    uint32_t delta = 0;
    if (blk < n_blocks && frame < a.frame_end) {
        const int16_t *src = frame_src(a, frame);
        delta = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15);
        if (lane == 0) {
            mbar_expect_tx(bar, PCM_COPY_BYTES);
            bulk_g2s(xchg_addr, reinterpret_cast<const unsigned char *>(src) - delta, PCM_COPY_BYTES, bar);
        }
    }

    for (; blk < n_blocks; blk += gridDim.x) {
        frame = a.frame_begin + blk * FP_WARPS + warp;
        if (frame >= a.frame_end) break;   // warp-uniform; only warp-level sync below

        f2 v[64];
        // ---- PCM (landed by the bulk copy) -> registers, window fused:
        //      z[32 m + L] = (x[2n] w[2n], x[2n+1] w[2n+1])
        mbar_wait(bar, phase);
        phase ^= 1;
        {
            const uint32_t *p = pcm_words + (delta >> 2) + lane;
            const uint32_t shift = (delta & 2) * 8;
#pragma unroll
            for (int m = 0; m < 64; m++) {
                const uint32_t w0 = p[32 * m], w1 = p[32 * m + 1];
                const f2 s = s16x2_to_float(__funnelshift_r(w0, w1, shift));
                v[m] = mul2(s, s_window[32 * m + lane]);
            }
        }
        __syncwarp();

        // ---- pass 1: 64-point FFT over m; Y[k1] (for n2 = L) at v[pos64(k1)]
        fft64(v);

        // ---- transpose: S[k1][n2]
#pragma unroll
        for (int k1 = 0; k1 < 64; k1++) xchg[k1 * XCHG_STRIDE + lane] = v[pos64(k1)];
        __syncwarp();
        // ---- pass 2: lane holds k1 = L + 32 q; twiddle W_2048^(n2 k1); 32-point FFT over n2
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int n2 = 0; n2 < 32; n2++)
                v[q * 32 + n2] = cmul(xchg[(lane + 32 * q) * XCHG_STRIDE + n2],
                                      s_twiddle[n2 * 64 + lane + 32 * q]);
        __syncwarp();
        // ---- the exchange buffer is free: start the next frame's PCM on its way
        {
            const uint32_t nblk = blk + gridDim.x;
            const uint32_t nframe = a.frame_begin + nblk * FP_WARPS + warp;
            if (nblk < n_blocks && nframe < a.frame_end) {
                const int16_t *src = frame_src(a, nframe);
                delta = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15);
                fence_proxy_async();   // our generic reads of xchg are ordered before the async write
                if (lane == 0) {
                    mbar_expect_tx(bar, PCM_COPY_BYTES);
                    bulk_g2s(xchg_addr, reinterpret_cast<const unsigned char *>(src) - delta, PCM_COPY_BYTES, bar);
                }
            }
        }
        fft32(v);
        fft32(v + 32);
        // now Z[L + 32 t] = v[zslot(t)]

        // ---- split + power + chroma fold (pitch classes are compile-time lane ranges)
        float acc[FP_BANDS];
#pragma unroll
        for (int b = 0; b < FP_BANDS; b++) acc[b] = 0.f;
#pragma unroll
        for (int t = 0; t < FP_T; t++) {
            const f2 z = v[zslot(t)];
            const f2 ps = v[zslot(63 - t)];
            f2 zp;
            zp.x = __shfl_sync(0xffffffffu, ps.x, partner);
            zp.y = __shfl_sync(0xffffffffu, ps.y, partner);
            if (t >= 1) {   // lane 0 pairs k = 32 t with 2048 - 32 t, which it holds itself
                const f2 own = v[zslot(t >= 1 ? 64 - t : 0)];
                if (lane == 0) zp = own;
            }
            // 2E = (a+c, b-d), 2O = (b+d, c-a) with Z[k] = a+ib, Z[N-k] = c+id
            const f2 e = fma2(zp, make_float2(1.f, -1.f), z);
            const f2 o = fma2(swp(z), make_float2(1.f, -1.f), swp(zp));
            const f2 w = cmulc(lane_w, c_w128[t], c_w128r[t]);            // W_4096^(L + 32 t)
            const f2 x = add2(e, cmul(o, w));                              // 2 X[k]
            const float pw = x.x * x.x + x.y * x.y;                        // 4 |X[k]|^2
#define FOLD(T, NOTE, LO, HI) \
            if (t == T && lane >= LO && lane < HI) acc[NOTE] += pw;
#include "fp_chroma_fold.inc"
#undef FOLD
        }
#pragma unroll
        for (int b = 0; b < FP_BANDS; b++) {
            float s = acc[b];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            acc[b] = 0.25f * s;
        }
        if (lane == 0) {
            float4 *out = reinterpret_cast<float4 *>(a.chroma + (uint64_t)frame * FP_BANDS);
            out[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            out[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            out[2] = make_float4(acc[8], acc[9], acc[10], acc[11]);
        }
    }
}

// ---------------------------------------------------------------------------
// K1, half-frame-in-registers variant ("h32").  The kernel above keeps all 64
// complex values of a lane in registers (255 registers, 8 warps per SM, two per
// scheduler) and is latency bound at 0.5 IPC.  Here a lane never holds more than
// 32: each 64-point pass-1 FFT is split by radix 2 into its even and odd rows,
// and pass 2 is done one 32-point FFT at a time, with the half that is not being
// worked on parked in the warp's exchange buffer:
//
//   rows m = 2j   -> 32-point FFT -> E[k] parked in rows 32..63 of the buffer
//   rows m = 2j+1 -> 32-point FFT -> O[k];  Y[k] = E[k] + W64^k O[k] -> row k,
//                                          Y[k+32] = E[k] - W64^k O[k] -> row k+32 (over E[k])
//   k1 = L+32: twiddle, 32-point FFT over n2 -> Z[L + 32 t], t odd, parked in rows 32..63
//   k1 = L:    twiddle, 32-point FFT over n2 -> Z[L + 32 t], t even, in registers
//   split: bin k = L + 32 t pairs with 2048 - k = (32 - L) + 32 (63 - t): one of the two
//   is in registers (own, or fetched by shuffle), the other in the parked half (own
//   slot, or read directly from the partner lane's slot -- no shuffle).
//
// ~100 registers -> NW = 12 warps per SM (three per scheduler); the exchange
// buffer (16.5 KB per warp) is what bounds NW: 12 x 16.5 + 16 (twiddles) + 8
// (window, half of it: the Hamming window is symmetric) = 222 KB of 227.
template <int NW> constexpr int h32_smem_bytes() { return (FP_HALF / 2 + FP_HALF) * 8 + NW * K1_WARP_BYTES + NW * 8; }

template <int NW>
__global__ void __launch_bounds__(NW * 32, 1) fp_fft_chroma_h32_kernel(const K1Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    f2 *s_window = reinterpret_cast<f2 *>(smem_raw);                     // 1024: first half of the window
    f2 *s_twiddle = s_window + FP_HALF / 2;                              // 2048
    unsigned char *s_warp = reinterpret_cast<unsigned char *>(s_twiddle + FP_HALF);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_warp + NW * K1_WARP_BYTES);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < FP_HALF; i += blockDim.x) {
        if (i < FP_HALF / 2) s_window[i] = a.window[i];
        s_twiddle[i] = a.twiddle[i];
    }
    if (threadIdx.x < NW) mbar_init(smem_u32(s_bar + threadIdx.x), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    f2 *xchg = reinterpret_cast<f2 *>(s_warp + warp * K1_WARP_BYTES);
    f2 *park = xchg + 32 * XCHG_STRIDE;                                  // rows 32..63
    const uint32_t *pcm_words = reinterpret_cast<const uint32_t *>(xchg);
    const uint32_t bar = smem_u32(s_bar + warp);
    const uint32_t xchg_addr = smem_u32(xchg);
    const uint32_t lane_addr = xchg_addr + 8u * (uint32_t)lane;                 // &xchg[0 * STRIDE + lane]
    const uint32_t park_addr = lane_addr + 32u * XCHG_STRIDE * 8u;              // &park[0 * STRIDE + lane]
    constexpr uint32_t ROW = XCHG_STRIDE * 8u;                                  // bytes per row
    const int partner = (32 - lane) & 31;
    const f2 lane_w = a.lane_w[lane];

    const uint32_t n_blocks = (a.frame_end - a.frame_begin + NW - 1) / NW;
    uint32_t blk = blockIdx.x;
    uint32_t frame = a.frame_begin + blk * NW + warp;
    uint32_t phase = 0;
    uint32_t delta = 0;
    if (blk < n_blocks && frame < a.frame_end) {
        const int16_t *src = frame_src(a, frame);
        delta = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15);
        if (lane == 0) {
            mbar_expect_tx(bar, PCM_COPY_BYTES);
            bulk_g2s(xchg_addr, reinterpret_cast<const unsigned char *>(src) - delta, PCM_COPY_BYTES, bar);
        }
    }

    for (; blk < n_blocks; blk += gridDim.x) {
        frame = a.frame_begin + blk * NW + warp;
        if (frame >= a.frame_end) break;   // warp-uniform; only warp-level sync below

        f2 v[32];
        mbar_wait(bar, phase);
        phase ^= 1;
        const uint32_t *p = pcm_words + (delta >> 2) + lane;
        const uint32_t shift = (delta & 2) * 8;
        // ---- pass 1: even rows (half = 0), then odd rows (half = 1).  One copy of the load and
        //      of the 32-point FFT serves both: the kernel's code must stay inside the
        //      instruction cache now that three warps per scheduler run at different places in it.
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            // row m = 2 j + half at lane L: PCM word 32 m + L, window w[32 m + L] for m < 32 (j < 16),
            // else the mirrored entry w[32 (63 - m) + 31 - L], swapped
            const uint32_t *ph = p + 32 * half;
            const f2 *wlo = s_window + 32 * half + lane;
            const f2 *whi = s_window + 32 * (63 - half) + 31 - lane;
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const uint32_t w0 = ph[64 * j], w1 = ph[64 * j + 1];
                const f2 s = s16x2_to_float(__funnelshift_r(w0, w1, shift));
                const f2 wv = j < 16 ? wlo[64 * j] : swp(whi[-64 * j]);
                v[j] = mul2(s, wv);
            }
            if (half == 1) __syncwarp();   // every lane has taken its PCM: rows 0..31 may be overwritten
            fft32(v);
            if (half == 0) {
#pragma unroll
                for (int k = 0; k < 32; k++) sts2(park_addr + k * ROW, v[pos32(k)]);   // bytes >= 8448: clear of the PCM
            } else {
#pragma unroll
                for (int k = 0; k < 32; k++) {
                    const f2 o = k == 0 ? v[pos32(0)] : cmulc(v[pos32(k)], c_w64[k], c_w64r[k]);
                    const f2 e = park[k * XCHG_STRIDE + lane];
                    sts2(lane_addr + k * ROW, add2(e, o));
                    sts2(park_addr + k * ROW, sub2(e, o));
                }
            }
        }
        __syncwarp();
        // ---- pass 2: k1 = L + 32 (q = 1: odd t, parked), then k1 = L (q = 0: even t, stays in registers)
#pragma unroll 1
        for (int q = 1; q >= 0; q--) {
            const f2 *row = xchg + (lane + 32 * q) * XCHG_STRIDE;
            const f2 *tw = s_twiddle + lane + 32 * q;
#pragma unroll
            for (int n2 = 0; n2 < 32; n2++) v[n2] = cmul3(row[n2], tw[n2 * 64]);
            __syncwarp();   // this half of the rows is consumed (q = 0: and the parked half is visible)
            if (q == 0) {
                // rows 0..31 are free: start the next frame's PCM on its way
                const uint32_t nblk = blk + gridDim.x;
                const uint32_t nframe = a.frame_begin + nblk * NW + warp;
                if (nblk < n_blocks && nframe < a.frame_end) {
                    const int16_t *src = frame_src(a, nframe);
                    delta = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15);
                    fence_proxy_async();
                    if (lane == 0) {
                        mbar_expect_tx(bar, PCM_COPY_BYTES);
                        bulk_g2s(xchg_addr, reinterpret_cast<const unsigned char *>(src) - delta, PCM_COPY_BYTES, bar);
                    }
                }
            }
            fft32(v);
            if (q == 1) {
#pragma unroll
                for (int k2 = 0; k2 < 32; k2++) sts2(park_addr + k2 * ROW, v[pos32(k2)]);   // Z[L + 32 (2 k2 + 1)]
            }
        }
        __syncwarp();
        // Z[L + 32 t]: t even -> v[pos32(t / 2)], t odd -> park[(t - 1) / 2][L]

        float acc[FP_BANDS];
#pragma unroll
        for (int b = 0; b < FP_BANDS; b++) acc[b] = 0.f;
        // the 123 lane-range predicates of the fold depend on the lane only; read the lane id
        // through a volatile asm so that they are recomputed here instead of being hoisted out of
        // the frame loop and kept (spilled) for its whole body
        int lane_s;
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_s));
        // likewise W_4096^(L + 32 t) = lane_w * W_128^t: two packed ops per bin when computed here;
        // hoisted out of the frame loop it becomes 41 spilled values per thread (140 KB of local
        // memory per CTA behind an L1 that shared memory has shrunk to ~30 KB)
        f2 lw = lane_w;
        asm volatile("" : "+f"(lw.x), "+f"(lw.y));
#pragma unroll
        for (int t = 0; t < FP_T; t++) {
            f2 z, zp;
            if ((t & 1) == 0) {
                z = v[pos32(t >> 1)];
                zp = park[(31 - (t >> 1)) * XCHG_STRIDE + partner];      // t' = 63 - t, odd, from the partner's slot
                // lane 0 pairs k = 32 t with 2048 - 32 t = 32 (64 - t): even, in its own registers (t = 0: itself)
                const f2 own = v[pos32(t == 0 ? 0 : (64 - t) >> 1)];
                if (lane == 0) zp = own;
            } else {
                z = park[(t >> 1) * XCHG_STRIDE + lane];
                const f2 ps = v[pos32((63 - t) >> 1)];                    // t' = 63 - t, even, in the partner's registers
                zp.x = __shfl_sync(0xffffffffu, ps.x, partner);
                zp.y = __shfl_sync(0xffffffffu, ps.y, partner);
                const f2 own = park[((64 - t) >> 1) * XCHG_STRIDE];       // lane 0: t' = 64 - t, odd, parked at lane 0
                if (lane == 0) zp = own;
            }
            const f2 e = fma2(zp, make_float2(1.f, -1.f), z);
            const f2 o = fma2(swp(z), make_float2(1.f, -1.f), swp(zp));
            const f2 w = cmulc(lw, c_w128[t], c_w128r[t]);
            const f2 x = add2(e, cmul3(o, w));
            const float pw = x.x * x.x + x.y * x.y;
#define FOLD(T, NOTE, LO, HI) \
            if (t == T && lane_s >= LO && lane_s < HI) acc[NOTE] += pw;
#include "fp_chroma_fold.inc"
#undef FOLD
        }
        __syncwarp();   // the parked half has been read by every lane before the next frame overwrites it
        // 12 sums over 32 lanes with halving butterflies: after the exchange over lane bit 4 a lane
        // keeps 6 of the classes, after bit 3 it keeps 3, and only those 3 go through the last
        // three steps -- 18 shuffles instead of 60.  Lanes 0, 8, 16, 24 end up with 3 classes each.
        float r6[6], r3[3];
        {
            const bool up = (lane & 16) != 0;
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const float keep = up ? acc[i + 6] : acc[i], send = up ? acc[i] : acc[i + 6];
                r6[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
        }
        {
            const bool up = (lane & 8) != 0;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const float keep = up ? r6[i + 3] : r6[i], send = up ? r6[i] : r6[i + 3];
                r3[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int o = 4; o >= 1; o >>= 1) r3[i] += __shfl_xor_sync(0xffffffffu, r3[i], o);
        }
        if ((lane & 7) == 0) {
            float *out = a.chroma + (uint64_t)frame * FP_BANDS + 6 * ((lane >> 4) & 1) + 3 * ((lane >> 3) & 1);
            out[0] = 0.25f * r3[0];
            out[1] = 0.25f * r3[1];
            out[2] = 0.25f * r3[2];
        }
    }
}

// ---------------------------------------------------------------------------
// K1, tensor-memory variant ("tm"): the h32 kernel with the parked half of a frame in
// TMEM instead of shared memory.  What capped h32 at 12 warps per SM was its 16.5 KB
// exchange buffer per warp, and half of that buffer was only ever a lane-private parking
// place (E[k] during pass 1, the odd rows of Z during the split).  Blackwell has 256 KB
// of tensor memory per SM that this kernel's tensor-core-free arithmetic leaves idle:
// tcgen05.st / tcgen05.ld (32x32b shape) move registers to and from a warp's own 32
// TMEM lanes, i.e. exactly a lane-private spill space with its own data path:
//
//   pass 1, even rows -> 32-point FFT -> E[k]                      -> TMEM slot A
//   pass 1, odd rows  -> 32-point FFT -> O[k];  E[k] <- slot A
//                        Y[k+32] = E - W64^k O -> row k of the exchange buffer
//                        Y[k]    = E + W64^k O -> slot A (over E)
//   pass 2, q = 1: rows (transposed) -> twiddle -> 32-point FFT -> Z odd -> TMEM slot B
//   Y[k] <- slot A -> rows;  pass 2, q = 0 -> Z even, stays in registers
//   split: Z odd <- slot B four values at a time; the partner lane's values by shuffle
//
// The exchange buffer shrinks to 32 rows (8.25 KB per warp, still the PCM landing zone of
// the bulk copy), so 16 warps fit (4 per scheduler, 128 registers each); shared-memory
// traffic drops by the park/unpark stores and loads (~250 of ~930 wavefronts per frame).
// TMEM: 2 slots x 64 columns per warp, 4 warps per lane quadrant -> all 512 columns.
constexpr int TM_ROWS = 32;
constexpr int TM_WARP_BYTES = TM_ROWS * XCHG_STRIDE * 8;     // 8,448 B
static_assert(PCM_COPY_BYTES <= TM_WARP_BYTES, "PCM must fit the exchange buffer");
template <int NW> constexpr int tm_smem_bytes() { return (FP_HALF / 2 + FP_HALF) * 8 + NW * TM_WARP_BYTES + NW * 8 + 16 + FP_T * 32 * 8 + 1056 * 8; }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 8 complex values = 16 columns of this warp's lanes
__device__ __forceinline__ void tmem_st8(uint32_t taddr, f2 a0, f2 a1, f2 a2, f2 a3, f2 a4, f2 a5, f2 a6, f2 a7) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "f"(a0.x), "f"(a0.y), "f"(a1.x), "f"(a1.y), "f"(a2.x), "f"(a2.y), "f"(a3.x), "f"(a3.y),
                   "f"(a4.x), "f"(a4.y), "f"(a5.x), "f"(a5.y), "f"(a6.x), "f"(a6.y), "f"(a7.x), "f"(a7.y) : "memory");
}
// The load is asynchronous: its registers are valid after tcgen05.wait::ld.  The wait takes the
// registers as read-write operands so that no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, f2 (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=f"(r[0].x), "=f"(r[0].y), "=f"(r[1].x), "=f"(r[1].y), "=f"(r[2].x), "=f"(r[2].y), "=f"(r[3].x), "=f"(r[3].y),
                   "=f"(r[4].x), "=f"(r[4].y), "=f"(r[5].x), "=f"(r[5].y), "=f"(r[6].x), "=f"(r[6].y), "=f"(r[7].x), "=f"(r[7].y)
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld8(f2 (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0].x), "+f"(r[0].y), "+f"(r[1].x), "+f"(r[1].y), "+f"(r[2].x), "+f"(r[2].y), "+f"(r[3].x), "+f"(r[3].y),
                   "+f"(r[4].x), "+f"(r[4].y), "+f"(r[5].x), "+f"(r[5].y), "+f"(r[6].x), "+f"(r[6].y), "+f"(r[7].x), "+f"(r[7].y)
                 :: "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, f2 (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r[0].x), "=f"(r[0].y), "=f"(r[1].x), "=f"(r[1].y), "=f"(r[2].x), "=f"(r[2].y), "=f"(r[3].x), "=f"(r[3].y)
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld4(f2 (&r)[4]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0].x), "+f"(r[0].y), "+f"(r[1].x), "+f"(r[1].y), "+f"(r[2].x), "+f"(r[2].y), "+f"(r[3].x), "+f"(r[3].y)
                 :: "memory");
}
__device__ __forceinline__ void tmem_wait_ld4x2(f2 (&r)[4], f2 (&q)[4]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0].x), "+f"(r[0].y), "+f"(r[1].x), "+f"(r[1].y), "+f"(r[2].x), "+f"(r[2].y), "+f"(r[3].x), "+f"(r[3].y),
                   "+f"(q[0].x), "+f"(q[0].y), "+f"(q[1].x), "+f"(q[1].y), "+f"(q[2].x), "+f"(q[2].y), "+f"(q[3].x), "+f"(q[3].y)
                 :: "memory");
}
__device__ __forceinline__ f2 tmem_ld1_wait(uint32_t taddr) {
    f2 r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=f"(r.x), "=f"(r.y) : "r"(taddr) : "memory");
    return r;
}

// ---- revision 2 of the arithmetic: fewer FMA-pipe cycles for the same transform.
// On sm_100a a packed FFMA2/FADD2/FMUL2 holds a scheduler's FMA pipe for two cycles and a
// scalar FFMA for one, so what counts is the number of scalar-equivalent results.  Two things
// waste them in the (re, im)-packed form: a complex product by a run-time twiddle (three packed
// ops = 6 cycles for 4 useful multiply-adds), and a product followed by a radix-2 butterfly
// (a + w b, a - w b) computed as product, sum, difference (8 cycles packed).  Here
//   * the sum is accumulated INTO the product (two FFMA2 on top of `a`) and the difference is
//     2 a - (a + w b): 6 cycles, three instructions;
//   * products by shared-memory twiddles are four scalar FFMA (4 cycles), also accumulated.
__device__ __forceinline__ f2 neg2(f2 a) { return make_float2(-a.x, -a.y); }
// (plus, minus) = (a + w b, a - w b); wr = (-w.y, w.x)
__device__ __forceinline__ void bfly_w(const f2 a, const f2 b, const f2 w, const f2 wr, f2 &plus, f2 &minus) {
    plus = fma2(by(b), wr, fma2(bx(b), w, a));
    minus = fma2(a, make_float2(2.f, 2.f), neg2(plus));
}
// a * w and c + a * w with scalar FFMA (run-time w: no rotated copy needed)
__device__ __forceinline__ f2 cmul_s(const f2 a, const f2 w) {
    f2 r;
    r.x = __fmaf_rn(-a.y, w.y, __fmul_rn(a.x, w.x));
    r.y = __fmaf_rn(a.y, w.x, __fmul_rn(a.x, w.y));
    return r;
}
__device__ __forceinline__ f2 cmad_s(const f2 a, const f2 w, const f2 c) {
    f2 r;
    r.x = __fmaf_rn(-a.y, w.y, __fmaf_rn(a.x, w.x, c.x));
    r.y = __fmaf_rn(a.y, w.x, __fmaf_rn(a.x, w.y, c.y));
    return r;
}
// second half of fft8: v[0..3S] = b0..b3 (sums), v[4S] = b4, v[5S] = t5, v[6S] = e6, v[7S] = t7 (differences)
template <int S> __device__ __forceinline__ void fft8_tail(f2 *v) {
    constexpr float R = 0.70710678118654752440f;
    const f2 b0 = v[0], b1 = v[S], b2 = v[2 * S], b3 = v[3 * S];
    const f2 b4 = v[4 * S], t5 = v[5 * S], e6 = v[6 * S], t7 = v[7 * S];
    const f2 b5 = fma2(by(t5), make_float2(R, R), mul2(bx(t5), make_float2(R, -R)));
    const f2 b7 = fma2(by(t7), make_float2(R, -R), mul2(bx(t7), make_float2(-R, -R)));
    {
        const f2 d0 = add2(b0, b2), d1 = sub2(b0, b2), d2 = add2(b1, b3), e = sub2(b1, b3);
        v[0] = add2(d0, d2);
        v[2 * S] = add_mi(d1, e);
        v[4 * S] = sub2(d0, d2);
        v[6 * S] = sub_mi(d1, e);
    }
    {
        const f2 d0 = add_mi(b4, e6), d1 = sub_mi(b4, e6), d2 = add2(b5, b7), e = sub2(b5, b7);
        v[S] = add2(d0, d2);
        v[3 * S] = add_mi(d1, e);
        v[5 * S] = sub2(d0, d2);
        v[7 * S] = sub_mi(d1, e);
    }
}
// fft32 whose first radix-2 layer the caller has already done (fused with the window or the
// twiddles): for every n1 < 8, with c_i = x[n1 + 8 i]:
//   v[n1] = c0 + c2, v[n1 + 16] = c0 - c2, v[n1 + 8] = c1 + c3, v[n1 + 24] = c1 - c3.
// Output order as fft32(): X[4 k1 + k2] at v[k1 + 8 k2].
__device__ __forceinline__ void fft32_tail(f2 *v) {
#pragma unroll
    for (int n1 = 0; n1 < 8; n1++) {
        const f2 d0 = v[n1], d2 = v[n1 + 8], d1 = v[n1 + 16], e = v[n1 + 24];
        v[n1] = add2(d0, d2);
        v[n1 + 8] = add_mi(d1, e);
        v[n1 + 16] = sub2(d0, d2);
        v[n1 + 24] = sub_mi(d1, e);
    }
    fft8<1>(v);
#pragma unroll
    for (int k2 = 1; k2 < 4; k2++) {
        f2 *u = v + 8 * k2;
        // first layer of the 8-point FFT over n1 with the twiddles W32^(n1 k2) folded in
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const f2 aj = j == 0 ? u[0] : cmulc(u[j], c_w32[j * k2], c_w32r[j * k2]);
            const int e4 = (j + 4) * k2;
            if (e4 == 8) {   // W32^8 = -i
                const f2 b = u[j + 4];
                u[j] = add_mi(aj, b);
                u[j + 4] = sub_mi(aj, b);
            } else {
                bfly_w(aj, u[j + 4], c_w32[e4], c_w32r[e4], u[j], u[j + 4]);
            }
        }
        fft8_tail<1>(u);
    }
}
// one predicated add: acc += pw where the lane's membership bit is set (LOP3 to a predicate + FADD)
__device__ __forceinline__ void fold_add(float &acc, const float pw, const uint32_t m, const uint32_t bit) {
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 p, t, 0;\n\t@p add.rn.f32 %0, %0, %1;\n\t}"
        : "+f"(acc) : "f"(pw), "r"(m), "r"(bit));
}
// split + power + chroma fold of bin k = L + 32 t, revision 2: w = W_4096^(L + 32 t) from shared memory
__device__ __forceinline__ void tm_bin2(const int t, const f2 z, const f2 zp, const f2 w, const uint32_t (&fold_m)[4],
                                        float (&acc)[FP_BANDS]) {
    const f2 e = fma2(zp, make_float2(1.f, -1.f), z);
    const f2 o = fma2(swp(z), make_float2(1.f, -1.f), swp(zp));
    const f2 x = cmad_s(o, w, e);                    // 2 X[k] = 2E + W 2O
    const float pw = __fmaf_rn(x.x, x.x, __fmul_rn(x.y, x.y));
#define FOLDC(T, NOTE, IDX, LANES)                                                   \
    if (t == T) {                                                                    \
        if ((LANES) == 0xffffffffu) acc[NOTE] += pw;                                 \
        else fold_add(acc[NOTE], pw, fold_m[(IDX) >> 5], 1u << ((IDX) & 31));        \
    }
#include "fp_chroma_fold.inc"
#undef FOLDC
}

// split + power + chroma fold of bin k = L + 32 t (t is a compile-time constant after unrolling):
// z = Z[k], zp = Z[2048 - k]
__device__ __forceinline__ void tm_bin(const int t, const f2 z, const f2 zp, const f2 lw, const uint32_t (&fold_m)[4],
                                       float (&acc)[FP_BANDS]) {
    // 2E = (a+c, b-d), 2O = (b+d, c-a) with Z[k] = a+ib, Z[N-k] = c+id
    const f2 e = fma2(zp, make_float2(1.f, -1.f), z);
    const f2 o = fma2(swp(z), make_float2(1.f, -1.f), swp(zp));
    const f2 w = cmulc(lw, c_w128[t], c_w128r[t]);   // W_4096^(L + 32 t)
    const f2 x = add2(e, cmul3(o, w));               // 2 X[k]
    const float pw = x.x * x.x + x.y * x.y;          // 4 |X[k]|^2
#define FOLDC(T, NOTE, IDX, LANES)                                                   \
    if (t == T) {                                                                    \
        if ((LANES) == 0xffffffffu) acc[NOTE] += pw;                                 \
        else if (fold_m[(IDX) >> 5] & (1u << ((IDX) & 31))) acc[NOTE] += pw;         \
    }
#include "fp_chroma_fold.inc"
#undef FOLDC
}

template <int NW, int REV>
__global__ void __launch_bounds__(NW * 32, 1) fp_fft_chroma_tm_kernel(const K1Args a) {
    static_assert(NW % 4 == 0 && NW <= 16, "4 warps per TMEM lane quadrant, 128 columns each");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    f2 *s_window = reinterpret_cast<f2 *>(smem_raw);                     // 1024: first half of the window
    f2 *s_twiddle = s_window + FP_HALF / 2;                              // 2048
    unsigned char *s_warp = reinterpret_cast<unsigned char *>(s_twiddle + FP_HALF);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_warp + NW * TM_WARP_BYTES);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + NW);
    f2 *s_wsplit = reinterpret_cast<f2 *>(s_tmem + 4);                   // REV >= 2: [41][32] W_4096^(L + 32 t)
    f2 *s_window_odd = s_wsplit + FP_T * 32;                             // REV >= 3: 1056 pairs (w[2n-1], w[2n]), see below

    const int lane = threadIdx.x & 31;
    // the warp index through a shuffle: the compiler then knows it is warp-uniform and keeps what
    // derives from it (TMEM addresses, barrier address) in uniform registers
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    for (int i = threadIdx.x; i < FP_HALF; i += blockDim.x) {
        if (i < FP_HALF / 2) s_window[i] = a.window[i];
        s_twiddle[i] = a.twiddle[i];
        if (REV >= 2 && i < FP_T * 32) s_wsplit[i] = a.wsplit[i];
        if (REV >= 3 && i < 1056) s_window_odd[i] = a.window_odd[i];
    }
    if (threadIdx.x < NW) mbar_init(smem_u32(s_bar + threadIdx.x), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    constexpr uint32_t TM_COLS = 512;                                    // power of two >= NW / 4 * 128
    if (warp == 0) tmem_alloc(smem_u32(s_tmem), TM_COLS);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;
    // this warp's lanes (its quadrant of the 128) and its two 64-column slots
    // REV >= 4: ONE 64-column slot per warp -- Y[k] leaves it for the exchange rows as soon as pass 2
    // (q = 1) has pulled its rows into registers, before that pass's FFT, so the odd rows of Z can take
    // the same columns afterwards.  The freed half of TMEM holds two lane-private constant tables in
    // the order a lane reads them: the pass-2 twiddles (columns 256..383) and the split's
    // W_4096^(L + 32 t) (384..465).  TMEM reads are ~20x what this kernel needs (tools/tmem_bw.cu);
    // shared memory was its busiest pipe (73 % of peak wavefronts).
    const uint32_t tq = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    const uint32_t tA = tq + (uint32_t)(warp >> 2) * (REV >= 4 ? 64u : 128u);
    const uint32_t tB = REV >= 4 ? tA : tA + 64u;
    const uint32_t tT = tq + 256u, tS = tq + 384u;
    if (REV >= 4) {
        if (warp < 4) {   // one warp per lane quadrant fills that quadrant's copy
#pragma unroll 1
            for (int c = 0; c < 8; c++) {
                f2 tv[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int t = 8 * c + e;                                   // table position
                    const int hq = t >> 5, n1 = (t >> 2) & 7, i = t & 3;
                    tv[e] = a.twiddle[(n1 + 8 * i) * 64 + lane + 32 * (1 - hq)];   // first q = 1 (hq = 0), then q = 0
                }
                tmem_st8(tT + 16 * c, tv[0], tv[1], tv[2], tv[3], tv[4], tv[5], tv[6], tv[7]);
            }
#pragma unroll 1
            for (int c = 0; c < 6; c++) {
                f2 sv8[8];
#pragma unroll
                for (int e = 0; e < 8; e++) sv8[e] = a.wsplit[min(8 * c + e, FP_T - 1) * 32 + lane];   // rows t = 8 c + e
                tmem_st8(tS + 16 * c, sv8[0], sv8[1], sv8[2], sv8[3], sv8[4], sv8[5], sv8[6], sv8[7]);
            }
            tmem_wait_st();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    f2 *xchg = reinterpret_cast<f2 *>(s_warp + warp * TM_WARP_BYTES);
    const uint32_t *pcm_words = reinterpret_cast<const uint32_t *>(xchg);
    const uint32_t bar = smem_u32(s_bar + warp);
    const uint32_t xchg_addr = smem_u32(xchg);
    const uint32_t lane_addr = xchg_addr + 8u * (uint32_t)lane;                 // &xchg[0 * STRIDE + lane]
    constexpr uint32_t ROW = XCHG_STRIDE * 8u;                                  // bytes per row
    const int partner = (32 - lane) & 31;
    const f2 lane_w = a.lane_w[lane];

    // chroma fold: bit idx of fold_m[] says whether this lane's bin of that (row, pitch class) entry
    // belongs to the class (fp_chroma_fold.inc, FOLDC).  Four registers for the whole kernel; the
    // range compares they replace were ~300 instructions per frame.
    uint32_t fold_m[4] = {0u, 0u, 0u, 0u};
#define FOLDC(T, NOTE, IDX, LANES) \
    if (((LANES) >> lane) & 1u) fold_m[(IDX) >> 5] |= 1u << ((IDX) & 31);
#include "fp_chroma_fold.inc"
#undef FOLDC
    asm volatile("" : "+r"(fold_m[0]), "+r"(fold_m[1]), "+r"(fold_m[2]), "+r"(fold_m[3]));

    const uint32_t n_blocks = (a.frame_end - a.frame_begin + NW - 1) / NW;
    uint32_t blk = blockIdx.x;
    uint32_t frame = a.frame_begin + blk * NW + warp;
    uint32_t phase = 0;
    uint32_t delta = 0;
    // REV >= 3: a frame whose first sample sits at an odd 2-byte position (every other frame: the hop is
    // 1365 samples) is fetched from ONE SAMPLE EARLIER, which makes its sample pairs 4-byte aligned words,
    // and transformed as the frame rotated by one sample: v = (xw[4095], xw[0], .., xw[4094]).  A circular
    // shift multiplies the spectrum by a unit-modulus phase: the power spectrum -- all that Chromaprint
    // takes from the FFT -- is the same.  The window for those frames is the window shifted by one
    // (s_window_odd), and the one sample that wraps around (xw[4095] in place of the sample before the
    // frame) is patched into lane 0's first value.  Every frame then costs one LDS per complex value
    // and no funnel shift.  `odd` travels in bit 4 of delta.
    auto frame_bytes = [&](uint32_t f) -> const unsigned char * {
        const unsigned char *src = reinterpret_cast<const unsigned char *>(frame_src(a, f));
        uint32_t odd = 0;
        if (REV >= 3) {
            odd = (uint32_t)(reinterpret_cast<uintptr_t>(src) >> 1) & 1u;
            src -= 2 * odd;
        }
        delta = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15) | (odd << 4);
        return src - (delta & 15);
    };
    if (blk < n_blocks && frame < a.frame_end) {
        const unsigned char *base = frame_bytes(frame);
        if (lane == 0) {
            mbar_expect_tx(bar, PCM_COPY_BYTES);
            bulk_g2s(xchg_addr, base, PCM_COPY_BYTES, bar);
        }
    }

    for (; blk < n_blocks; blk += gridDim.x) {
        frame = a.frame_begin + blk * NW + warp;
        if (frame >= a.frame_end) break;   // warp-uniform; only warp-level sync below

        f2 v[32];
        mbar_wait(bar, phase);
        phase ^= 1;
        const uint32_t *p = pcm_words + ((delta & 15) >> 2) + lane;
        const uint32_t shift = (delta & 2) * 8;
        const bool odd = REV >= 3 && (delta & 16) != 0;
        // rotated frame: the sample that wraps around, x[4095], is the low half of word 2048
        float wrap = 0.f;
        if (REV >= 3) wrap = (float)(short)(pcm_words[((delta & 15) >> 2) + 2048] & 0xffffu);
        // ---- pass 1: even rows (half = 0), then odd rows (half = 1); one copy of the code
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            const uint32_t *ph = p + 32 * half;
            const f2 *wt = odd ? s_window_odd : s_window;
            const f2 *wlo = wt + 32 * half + lane;
            const f2 *whi = wt + 32 * (63 - half) + (odd ? 32 : 31) - lane;
            if (REV >= 2) {
                // window and first radix-2 layer of the 32-point FFT in one: c = s w, (c0 + s2 w2, 2 c0 - that);
                // the samples go through the conversion unit (I2F.S16 on each half of the word)
#pragma unroll
                for (int n1 = 0; n1 < 8; n1++) {
                    f2 sv[4], wv[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int j = n1 + 8 * i;
                        if (REV >= 3) {
                            sv[i] = s16x2_to_float_xu(ph[64 * j]);
                        } else {
                            const uint32_t w0 = ph[64 * j], w1 = ph[64 * j + 1];
                            sv[i] = s16x2_to_float_xu(__funnelshift_r(w0, w1, shift));
                        }
                        wv[i] = j < 16 ? wlo[64 * j] : swp(whi[-64 * j]);
                    }
                    if (REV >= 3 && n1 == 0) {
                        // value 0 of the rotated frame (lane 0, row 0, first half): (x[4095], x[0])
                        if (odd && half == 0 && lane == 0) sv[0].x = wrap;
                    }
                    const f2 c0 = mul2(sv[0], wv[0]), c1 = mul2(sv[1], wv[1]);
                    const f2 d0 = fma2(sv[2], wv[2], c0), d2 = fma2(sv[3], wv[3], c1);
                    v[n1] = d0;
                    v[n1 + 16] = fma2(c0, make_float2(2.f, 2.f), neg2(d0));
                    v[n1 + 8] = d2;
                    v[n1 + 24] = fma2(c1, make_float2(2.f, 2.f), neg2(d2));
                }
                if (half == 1) __syncwarp();   // every lane has taken its PCM: the rows may be overwritten
                fft32_tail(v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const uint32_t w0 = ph[64 * j], w1 = ph[64 * j + 1];
                    const f2 s = s16x2_to_float(__funnelshift_r(w0, w1, shift));
                    const f2 wv = j < 16 ? wlo[64 * j] : swp(whi[-64 * j]);
                    v[j] = mul2(s, wv);
                }
                if (half == 1) __syncwarp();   // every lane has taken its PCM: the rows may be overwritten
                fft32(v);
            }
            if (half == 0) {
                // E[k] sits in v[pos32(k)]; parked in register order (slot j = v[j])
#pragma unroll
                for (int c = 0; c < 4; c++)
                    tmem_st8(tA + 16 * c, v[8 * c], v[8 * c + 1], v[8 * c + 2], v[8 * c + 3], v[8 * c + 4],
                             v[8 * c + 5], v[8 * c + 6], v[8 * c + 7]);
            } else {
                tmem_wait_st();
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    f2 e[8];
                    tmem_ld8(tA + 16 * c, e);
                    tmem_wait_ld8(e);
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int k = c + 4 * i;   // pos32(k) = 8 c + i
                        if (REV >= 2 && k != 0) {
                            f2 yp, ym;
                            bfly_w(e[i], v[8 * c + i], c_w64[k], c_w64r[k], yp, ym);
                            sts2(lane_addr + k * ROW, ym);              // Y[k + 32] -> row k
                            e[i] = yp;                                  // Y[k]      -> back to slot A
                        } else {
                            const f2 o = k == 0 ? v[8 * c + i] : cmulc(v[8 * c + i], c_w64[k], c_w64r[k]);
                            sts2(lane_addr + k * ROW, sub2(e[i], o));   // Y[k + 32] -> row k
                            e[i] = add2(e[i], o);                       // Y[k]      -> back to slot A
                        }
                    }
                    tmem_st8(tA + 16 * c, e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7]);
                }
            }
        }
        __syncwarp();
        // ---- pass 2: k1 = L + 32 (q = 1: odd t, parked in slot B), then k1 = L (q = 0: even t, in registers)
#pragma unroll 1
        for (int q = 1; q >= 0; q--) {
            if (q == 0 && REV < 4) {
                // Y[k], k < 32, from slot A into the rows that pass q = 1 has consumed
                tmem_wait_st();
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    f2 y[8];
                    tmem_ld8(tA + 16 * c, y);
                    tmem_wait_ld8(y);
#pragma unroll
                    for (int i = 0; i < 8; i++) sts2(lane_addr + (c + 4 * i) * ROW, y[i]);
                }
                __syncwarp();
            }
            if (q == 0 && REV >= 4) __syncwarp();   // the rows written after pass q = 1's reads are visible
            const f2 *row = xchg + lane * XCHG_STRIDE;
            const f2 *tw = s_twiddle + lane + 32 * q;
            if (REV >= 2) {
                // twiddles and first radix-2 layer in one, scalar FFMA: c0 = r0 t0, c0 + r2 t2, 2 c0 - that
#pragma unroll
                for (int n1 = 0; n1 < 8; n1++) {
                    f2 t4[4];
                    if (REV >= 4) {
                        tmem_ld4(tT + 64 * (1 - q) + 8 * n1, t4);
                        tmem_wait_ld4(t4);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; i++) t4[i] = tw[(n1 + 8 * i) * 64];
                    }
                    const f2 c0 = n1 == 0 ? row[0] : cmul_s(row[n1], t4[0]);
                    const f2 c1 = cmul_s(row[n1 + 8], t4[1]);
                    const f2 d0 = cmad_s(row[n1 + 16], t4[2], c0);
                    const f2 d2 = cmad_s(row[n1 + 24], t4[3], c1);
                    v[n1] = d0;
                    v[n1 + 16] = fma2(c0, make_float2(2.f, 2.f), neg2(d0));
                    v[n1 + 8] = d2;
                    v[n1 + 24] = fma2(c1, make_float2(2.f, 2.f), neg2(d2));
                }
            } else {
#pragma unroll
                for (int n2 = 0; n2 < 32; n2++) v[n2] = cmul3(row[n2], tw[n2 * 64]);
            }
            __syncwarp();   // the rows are consumed
            if (q == 1 && REV >= 4) {
                // the rows are free and the FFT has not started: Y[k], k < 32, leaves the slot for them now
                tmem_wait_st();
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    f2 y[8];
                    tmem_ld8(tA + 16 * c, y);
                    tmem_wait_ld8(y);
#pragma unroll
                    for (int i = 0; i < 8; i++) sts2(lane_addr + (c + 4 * i) * ROW, y[i]);
                }
            }
            if (q == 0) {
                // the buffer is free: start the next frame's PCM on its way
                const uint32_t nblk = blk + gridDim.x;
                const uint32_t nframe = a.frame_begin + nblk * NW + warp;
                if (nblk < n_blocks && nframe < a.frame_end) {
                    const unsigned char *base = frame_bytes(nframe);
                    fence_proxy_async();
                    if (lane == 0) {
                        mbar_expect_tx(bar, PCM_COPY_BYTES);
                        bulk_g2s(xchg_addr, base, PCM_COPY_BYTES, bar);
                    }
                }
            }
            if (REV >= 2) fft32_tail(v); else fft32(v);
            if (q == 1) {
                // Z[L + 32 (2 k2 + 1)] parked in k2 order: slot B column 2 k2
#pragma unroll
                for (int c = 0; c < 4; c++)
                    tmem_st8(tB + 16 * c, v[pos32(8 * c)], v[pos32(8 * c + 1)], v[pos32(8 * c + 2)], v[pos32(8 * c + 3)],
                             v[pos32(8 * c + 4)], v[pos32(8 * c + 5)], v[pos32(8 * c + 6)], v[pos32(8 * c + 7)]);
            }
        }
        tmem_wait_st();
        // Z[L + 32 t]: t even -> v[pos32(t / 2)], t odd -> slot B entry (t - 1) / 2

        float acc[FP_BANDS];
#pragma unroll
        for (int b = 0; b < FP_BANDS; b++) acc[b] = 0.f;
        f2 lw = lane_w;
        asm volatile("" : "+f"(lw.x), "+f"(lw.y));   // keep W_4096^(L + 32 t) from being hoisted (41 spilled values)
        // rows t = 2 a and t = 2 a + 1 together: both need the parked entry 31 - a (the even row's
        // partner value is the partner lane's entry; lane 0's partner of the odd row is its own)
#pragma unroll
        for (int g = 0; g < 5; g++) {
            f2 pa[4], pb[4], ws[8];
            tmem_ld4(tB + 2 * (4 * g), pa);          // entries 4 g .. 4 g + 3
            tmem_ld4(tB + 2 * (28 - 4 * g), pb);     // entries 28 - 4 g .. 31 - 4 g: pb[3 - i] = entry 31 - (4 g + i)
            if (REV >= 4) tmem_ld8(tS + 16 * g, ws);  // W_4096^(L + 32 t), t = 8 g .. 8 g + 7
            tmem_wait_ld4x2(pa, pb);
            if (REV >= 4) tmem_wait_ld8(ws);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int a2 = 4 * g + i;
                {   // t = 2 a: own register; partner's parked entry 31 - a
                    const f2 z = v[pos32(a2)];
                    const f2 own = v[pos32(a2 == 0 ? 0 : 32 - a2)];   // lane 0: k = 32 t pairs with 32 (64 - t), even
                    f2 zp;
                    zp.x = __shfl_sync(0xffffffffu, pb[3 - i].x, partner);
                    zp.y = __shfl_sync(0xffffffffu, pb[3 - i].y, partner);
                    if (lane == 0) zp = own;
                    if (REV >= 2) tm_bin2(2 * a2, z, zp, REV >= 4 ? ws[2 * i] : s_wsplit[(2 * a2) * 32 + lane], fold_m, acc); else tm_bin(2 * a2, z, zp, lw, fold_m, acc);
                }
                {   // t = 2 a + 1: own parked entry a; partner's register; lane 0: its own parked entry 31 - a
                    const f2 z = pa[i];
                    const f2 ps = v[pos32(31 - a2)];
                    f2 zp;
                    zp.x = __shfl_sync(0xffffffffu, ps.x, partner);
                    zp.y = __shfl_sync(0xffffffffu, ps.y, partner);
                    if (lane == 0) zp = pb[3 - i];
                    if (REV >= 2) tm_bin2(2 * a2 + 1, z, zp, REV >= 4 ? ws[2 * i + 1] : s_wsplit[(2 * a2 + 1) * 32 + lane], fold_m, acc); else tm_bin(2 * a2 + 1, z, zp, lw, fold_m, acc);
                }
            }
        }
        {   // t = 40: partner's parked entry 11
            const f2 p11 = tmem_ld1_wait(tB + 2 * 11);
            const f2 z = v[pos32(20)];
            const f2 own = v[pos32(12)];
            f2 zp;
            zp.x = __shfl_sync(0xffffffffu, p11.x, partner);
            zp.y = __shfl_sync(0xffffffffu, p11.y, partner);
            if (lane == 0) zp = own;
            if (REV >= 2) tm_bin2(40, z, zp, REV >= 4 ? tmem_ld1_wait(tS + 80) : s_wsplit[40 * 32 + lane], fold_m, acc); else tm_bin(40, z, zp, lw, fold_m, acc);
        }
        // 12 sums over 32 lanes with halving butterflies (as in the h32 kernel): 18 shuffles
        float r6[6], r3[3];
        {
            const bool up = (lane & 16) != 0;
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const float keep = up ? acc[i + 6] : acc[i], send = up ? acc[i] : acc[i + 6];
                r6[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
        }
        {
            const bool up = (lane & 8) != 0;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const float keep = up ? r6[i + 3] : r6[i], send = up ? r6[i] : r6[i + 3];
                r3[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int o = 4; o >= 1; o >>= 1) r3[i] += __shfl_xor_sync(0xffffffffu, r3[i], o);
        }
        if ((lane & 7) == 0) {
            float *out = a.chroma + (uint64_t)frame * FP_BANDS + 6 * ((lane >> 4) & 1) + 3 * ((lane >> 3) & 1);
            out[0] = 0.25f * r3[0];
            out[1] = 0.25f * r3[1];
            out[2] = 0.25f * r3[2];
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TM_COLS);
}

// --------------------------------------------------------------------- K2

struct K2Args {
    const float *chroma;
    const FpSegment *segs;
    uint32_t n_segs;
    uint32_t n_tiles;
    uint32_t stride;
    float item_secs;        // Duration::as_secs_f32(item_duration)
    uint64_t delay_ns;
    uint32_t *out_hash;
    uint64_t *out_ts;
};

__device__ __forceinline__ double area(const double *I, int r1, int c1, int r2, int c2) {
    return I[r2 * K2_COLS + c2] - I[r1 * K2_COLS + c2] - I[r2 * K2_COLS + c1] + I[r1 * K2_COLS + c1];
}

__global__ void __launch_bounds__(K2_TILE) fp_classify_kernel(const K2Args a) {
    // I[r][c] = sum of rows < r, columns < c of the tile's normalised rows
    __shared__ double I[(K2_ROWS + 1) * K2_COLS];
    const uint32_t tile = blockIdx.x;
    uint32_t lo = 0, hi = a.n_segs;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.segs[mid].tile_begin <= tile) lo = mid; else hi = mid;
    }
    const FpSegment sg = a.segs[lo];
    const uint32_t x0 = (tile - sg.tile_begin) * K2_TILE;
    const int n_here = (int)min((uint32_t)K2_TILE, sg.n_raw - x0);
    const int n_rows = n_here + FP_WINDOW_ROWS - 1;
    const int tid = threadIdx.x;

    for (int c = tid; c < K2_COLS; c += K2_TILE) I[c] = 0.0;
    for (int r = tid; r < n_rows; r += K2_TILE) {
        // ChromaFilter: row r of the tile = frames x0+r .. x0+r+4
        const float *f = a.chroma + ((uint64_t)sg.frame_begin + x0 + r) * FP_BANDS;
        double y[FP_BANDS];
#pragma unroll
        for (int b = 0; b < FP_BANDS; b++) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < FP_FIR_LEN; j++) s += (double)f[j * FP_BANDS + b] * c_fir[j];
            y[b] = s;
        }
        // ChromaNormalizer: Euclidean norm, zero below the threshold
        double sq = 0.0;
#pragma unroll
        for (int b = 0; b < FP_BANDS; b++) sq += y[b] * y[b];
        const double norm = sqrt(sq);
        double *row = I + (r + 1) * K2_COLS;
        row[0] = 0.0;
        double run = 0.0;
#pragma unroll
        for (int b = 0; b < FP_BANDS; b++) {
            const double val = norm < FP_NORM_THRESHOLD ? 0.0 : y[b] / norm;
            run += val;
            row[b + 1] = run;   // prefix along the pitch classes
        }
    }
    __syncthreads();
    if (tid >= 1 && tid < K2_COLS) {   // prefix along time, one thread per column
        double run = 0.0;
        for (int r = 1; r <= n_rows; r++) {
            run += I[r * K2_COLS + tid];
            I[r * K2_COLS + tid] = run;
        }
    }
    __syncthreads();
    if (tid >= n_here) return;

    const int x = tid;   // window = rows [x, x+16) of the tile
    uint32_t bits = 0;
#pragma unroll 1
    for (int k = 0; k < FP_NUM_CLASSIFIERS; k++) {
        const K2Classifier cl = c_cls[k];
        const int y = cl.y, h = cl.h, w = cl.w;
        double pa, pb;
        switch (cl.type) {
        case 0:
            pa = area(I, x, y, x + w, y + h);
            pb = 0.0;
            break;
        case 1: {
            const int h2 = h / 2;
            pa = area(I, x, y + h2, x + w, y + h);
            pb = area(I, x, y, x + w, y + h2);
            break;
        }
        case 2: {
            const int w2 = w / 2;
            pa = area(I, x + w2, y, x + w, y + h);
            pb = area(I, x, y, x + w2, y + h);
            break;
        }
        case 3: {
            const int w2 = w / 2, h2 = h / 2;
            pa = area(I, x, y + h2, x + w2, y + h) + area(I, x + w2, y, x + w, y + h2);
            pb = area(I, x, y, x + w2, y + h2) + area(I, x + w2, y + h2, x + w, y + h);
            break;
        }
        case 4: {
            const int h3 = h / 3;
            pa = area(I, x, y + h3, x + w, y + 2 * h3);
            pb = area(I, x, y, x + w, y + h3) + area(I, x, y + 2 * h3, x + w, y + h);
            break;
        }
        default: {
            const int w3 = w / 3;
            pa = area(I, x + w3, y, x + 2 * w3, y + h);
            pb = area(I, x, y, x + w3, y + h) + area(I, x + 2 * w3, y, x + w, y + h);
            break;
        }
        }
        // Quantizer on log((1+a)/(1+b)), evaluated without the log
        const double num = 1.0 + pa, den = 1.0 + pb;
        int q;
        if (num < cl.e1 * den) q = num < cl.e0 * den ? 0 : 1;
        else q = num < cl.e2 * den ? 2 : 3;
        const uint32_t gray = (uint32_t)(q ^ (q >> 1));   // {0,1,3,2}
        bits = (bits << 2) | gray;
    }
    const uint32_t raw_idx = x0 + (uint32_t)x;
    if (raw_idx % a.stride == 0) {
        const uint32_t o = sg.out_off + raw_idx / a.stride;
        a.out_hash[o] = bits;
        // ts = delay + item.mul_f32(i as f32) + seek_to   (analyzer.rs:309,314-318)
        const float p = __fmul_rn(__uint2float_rn(raw_idx), a.item_secs);
        a.out_ts[o] = a.delay_ns + __double2ull_rn((double)p * 1e9) + sg.seek_to_ns;
    }
}

// Timestamps of stored hashes are a function of the index alone (analyzer.rs:309,314-318):
// ts[k] = delay + item.mul_f32((k * stride) as f32) + seek_to.  Every rank of a multi-GPU job
// can therefore fill them in for ALL segments itself instead of exchanging 8 bytes per hash.
struct TsSegment {
    uint64_t dev_offset, len, seek_to_ns;
};
__global__ void ts_fill_kernel(uint64_t *__restrict__ ts, const TsSegment *__restrict__ segs, uint32_t stride,
                               float item_secs, uint64_t delay_ns) {
    const TsSegment sg = segs[blockIdx.y];
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= sg.len) return;
    const float p = __fmul_rn(__ull2float_rn(k * stride), item_secs);
    ts[sg.dev_offset + k] = delay_ns + __double2ull_rn((double)p * 1e9) + sg.seek_to_ns;
}

// (L + R) / 2 with C integer division, as Chromaprint's AudioProcessor::LoadStereo
__global__ void downmix_kernel(const int16_t *__restrict__ stereo, int16_t *__restrict__ mono, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = stereo[2 * i], r = stereo[2 * i + 1];
    mono[i] = (int16_t)((l + r) / 2);
}

// ------------------------------------------------------------------- host

static float as_secs_f32(uint64_t ns) {
    volatile float s = (float)(ns / 1000000000ull);
    volatile float f = (float)(uint32_t)(ns % 1000000000ull) / 1000000000.0f;
    return s + f;
}

static int ensure_fp_tables(nb200_ctx *ctx) {
    if (ctx->fp_ready) return NB200_OK;
    const double PI = 3.14159265358979323846;
    std::vector<float2> window(FP_HALF), tw(FP_HALF), lane_w(32), w64(64), w32(32), w128(FP_T);
    for (int n = 0; n < FP_HALF; n++) {
        // PrepareHammingWindow(first, last, 1.0 / INT16_MAX)
        const double w0 = (1.0 / 32767.0) * (0.54 - 0.46 * cos((2 * n) * 2.0 * PI / (FP_FRAME - 1)));
        const double w1 = (1.0 / 32767.0) * (0.54 - 0.46 * cos((2 * n + 1) * 2.0 * PI / (FP_FRAME - 1)));
        window[n] = make_float2((float)w0, (float)w1);
    }
    for (int n2 = 0; n2 < 32; n2++)
        for (int k1 = 0; k1 < 64; k1++) {
            const double ang = -2.0 * PI * (double)(n2 * k1) / FP_HALF;
            tw[n2 * 64 + k1] = make_float2((float)cos(ang), (float)sin(ang));
        }
    for (int l = 0; l < 32; l++) {
        const double ang = -2.0 * PI * l / FP_FRAME;
        lane_w[l] = make_float2((float)cos(ang), (float)sin(ang));
    }
    // the window shifted by one sample, for frames that are transformed rotated by one (K1, REV >= 3):
    // pair n = (w[2n-1], w[2n]); pair 0 = (w[4095], w[0]); pairs above 1024 are the swapped mirror images
    std::vector<float2> window_odd(1056, make_float2(0.f, 0.f));
    {
        auto wnd = [&](int i) { return (1.0 / 32767.0) * (0.54 - 0.46 * cos(i * 2.0 * PI / (FP_FRAME - 1))); };
        for (int n = 0; n <= 1024; n++)
            window_odd[n] = make_float2((float)wnd(n == 0 ? FP_FRAME - 1 : 2 * n - 1), (float)wnd(2 * n));
    }
    std::vector<float2> wsplit(FP_T * 32);
    for (int k = 0; k < FP_T * 32; k++) {
        const double ang = -2.0 * PI * k / FP_FRAME;
        wsplit[k] = make_float2((float)cos(ang), (float)sin(ang));
    }
    for (int e = 0; e < 64; e++) w64[e] = make_float2((float)cos(-2.0 * PI * e / 64), (float)sin(-2.0 * PI * e / 64));
    for (int e = 0; e < 32; e++) w32[e] = make_float2((float)cos(-2.0 * PI * e / 32), (float)sin(-2.0 * PI * e / 32));
    for (int t = 0; t < FP_T; t++)
        w128[t] = make_float2((float)cos(-2.0 * PI * t / 128), (float)sin(-2.0 * PI * t / 128));

    // Chroma::PrepareNotes: bins [min_index, max_index) -> pitch class
    const int min_index = std::max(1, (int)lround((double)FP_FRAME * FP_MIN_FREQ / FP_SAMPLE_RATE));
    const int max_index = std::min(FP_FRAME / 2, (int)lround((double)FP_FRAME * FP_MAX_FREQ / FP_SAMPLE_RATE));
    if (min_index != FP_KMIN || max_index != FP_KMAX) {
        set_last_error_msg("chroma bin range does not match the kernel's compiled range");
        return NB200_ERR_STATE;
    }
    K2Classifier cls[FP_NUM_CLASSIFIERS];
    for (int k = 0; k < FP_NUM_CLASSIFIERS; k++) {
        const FpClassifierDef &d = FP_CLASSIFIERS_TEST2[k];
        cls[k] = {d.type, d.y, d.height, d.width, exp(d.t0), exp(d.t1), exp(d.t2)};
    }
    NB_CUDA(cudaMemcpyToSymbol(c_w64, w64.data(), sizeof(float2) * 64));
    NB_CUDA(cudaMemcpyToSymbol(c_w32, w32.data(), sizeof(float2) * 32));
    NB_CUDA(cudaMemcpyToSymbol(c_w128, w128.data(), sizeof(float2) * FP_T));
    auto rot = [](std::vector<float2> v) {
        for (auto &w : v) w = make_float2(-w.y, w.x);
        return v;
    };
    NB_CUDA(cudaMemcpyToSymbol(c_w64r, rot(w64).data(), sizeof(float2) * 64));
    NB_CUDA(cudaMemcpyToSymbol(c_w32r, rot(w32).data(), sizeof(float2) * 32));
    NB_CUDA(cudaMemcpyToSymbol(c_w128r, rot(w128).data(), sizeof(float2) * FP_T));
    NB_CUDA(cudaMemcpyToSymbol(c_cls, cls, sizeof(cls)));
    NB_CUDA(cudaMemcpyToSymbol(c_fir, FP_FIR_COEFFS, sizeof(double) * FP_FIR_LEN));
    NB_TRY(ctx->d_fp_window.alloc(sizeof(float2) * FP_HALF));
    NB_TRY(ctx->d_fp_twiddle.alloc(sizeof(float2) * FP_HALF));
    NB_TRY(ctx->d_fp_lane.alloc(sizeof(float2) * 32));
    NB_CUDA(cudaMemcpy(ctx->d_fp_window.p, window.data(), sizeof(float2) * FP_HALF, cudaMemcpyHostToDevice));
    NB_CUDA(cudaMemcpy(ctx->d_fp_twiddle.p, tw.data(), sizeof(float2) * FP_HALF, cudaMemcpyHostToDevice));
    NB_CUDA(cudaMemcpy(ctx->d_fp_lane.p, lane_w.data(), sizeof(float2) * 32, cudaMemcpyHostToDevice));
    NB_TRY(ctx->d_fp_window_odd.alloc(sizeof(float2) * window_odd.size()));
    NB_CUDA(cudaMemcpy(ctx->d_fp_window_odd.p, window_odd.data(), sizeof(float2) * window_odd.size(), cudaMemcpyHostToDevice));
    NB_TRY(ctx->d_fp_wsplit.alloc(sizeof(float2) * wsplit.size()));
    NB_CUDA(cudaMemcpy(ctx->d_fp_wsplit.p, wsplit.data(), sizeof(float2) * wsplit.size(), cudaMemcpyHostToDevice));
    NB_CUDA(cudaFuncSetAttribute(fp_fft_chroma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 K1_SMEM_BYTES));
    NB_CUDA(cudaFuncSetAttribute(fp_fft_chroma_h32_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 h32_smem_bytes<12>()));
    NB_CUDA(cudaFuncSetAttribute(fp_fft_chroma_h32_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 h32_smem_bytes<10>()));
    NB_CUDA(cudaFuncSetAttribute(fp_fft_chroma_h32_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 h32_smem_bytes<8>()));
    NB_CUDA(cudaFuncSetAttribute(fp_fft_chroma_tm_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 tm_smem_bytes<16>()));
    NB_CUDA(cudaFuncSetAttribute(fp_fft_chroma_tm_kernel<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 tm_smem_bytes<16>()));
    NB_CUDA(cudaFuncSetAttribute(fp_fft_chroma_tm_kernel<16, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 tm_smem_bytes<16>()));
    NB_CUDA(cudaFuncSetAttribute(fp_fft_chroma_tm_kernel<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 tm_smem_bytes<16>()));
    NB_CUDA(cudaFuncSetAttribute(fp_fft_chroma_tm_kernel<12, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 tm_smem_bytes<12>()));
    ctx->fp_ready = true;
    return NB200_OK;
}

// host_pcm != NULL: the PCM is not in ps->d_pcm yet.  It is copied there in chunks of
// segments on the copy stream while K1 already runs on the chunks that have landed
// (H2D over PCIe is ~10 ms for a season, K1 ~1.5 ms: everything but the last chunk hides).
// ext_hashes / ext_ts != NULL: write into these caller-owned device arrays
// (ext_capacity elements) instead of allocating; the returned set borrows them.
static int fingerprint_run_pipelined(nb200_ctx *ctx, const nb200_pcmset *ps, const int16_t *const *host_pcm,
                                     uint32_t stride, uint64_t delay_ns, uint64_t item_ns,
                                     const uint64_t *seek_to_ns, nb200_hashset **out,
                                     void *ext_hashes = nullptr, void *ext_ts = nullptr,
                                     uint64_t ext_capacity = 0, bool wait = true) {
    if (stride == 0) return NB200_ERR_INVALID_ARGUMENT;   // Rust's step_by(0) panics
    if (ps->n_segments % 2 != 0) return NB200_ERR_INVALID_ARGUMENT;
    NB_CUDA(cudaSetDevice(ctx->device));
    NB_TRY(ensure_fp_tables(ctx));
    PhaseTimer pt(ctx, NB200_PHASE_FP);

    const uint32_t nseg = ps->n_segments;
    std::vector<FpSegment> segs(nseg);
    std::vector<uint64_t> seg_offset(nseg + 1, 0);
    uint64_t frames = 0, tiles = 0;
    for (uint32_t s = 0; s < nseg; s++) {
        FpSegment &g = segs[s];
        const uint64_t nf = num_frames(ps->n_mono[s]);
        const uint64_t nr = num_raw(ps->n_mono[s]);
        g.pcm_off = ps->dev_offset[s];
        g.frame_begin = (uint32_t)frames;
        g.n_frames = (uint32_t)nf;
        g.n_raw = (uint32_t)nr;
        g.tile_begin = (uint32_t)tiles;
        g.pad = 0;
        g.seek_to_ns = seek_to_ns ? seek_to_ns[s] : 0;
        frames += nf;
        tiles += (nr + K2_TILE - 1) / K2_TILE;
        if (frames > 0x7fffffffull || tiles > 0x7fffffffull) return NB200_ERR_TOO_LARGE;
        seg_offset[s + 1] = seg_offset[s] + (nr + stride - 1) / stride;
    }

    nb200_hashset *hs = new (std::nothrow) nb200_hashset();
    if (!hs) return NB200_ERR_TOO_LARGE;
    auto fail = [&](int st) {
        delete hs;
        return st;
    };
    {
        // same aligned layout as nb200_hashset_upload
        hs->n_videos = nseg / 2;
        hs->seg_offset = seg_offset;
        hs->dev_offset.resize(nseg);
        hs->seg_len.resize(nseg);
        uint64_t dev = 0;
        for (uint32_t s = 0; s < nseg; s++) {
            const uint64_t len = seg_offset[s + 1] - seg_offset[s];
            hs->dev_offset[s] = (uint32_t)dev;
            hs->seg_len[s] = (uint32_t)len;
            segs[s].out_off = (uint32_t)dev;
            dev += (len + 3) & ~3ull;
            if (dev > 0xfffffff0ull) return fail(NB200_ERR_TOO_LARGE);
        }
        hs->total = seg_offset[nseg];
        hs->dev_total = dev;
    }
    if (ext_hashes) {
        if (hs->dev_total > ext_capacity) return fail(NB200_ERR_TOO_LARGE);
        hs->d_hashes.borrow(ext_hashes, (size_t)ext_capacity * sizeof(uint32_t));
        hs->d_ts.borrow(ext_ts, (size_t)ext_capacity * sizeof(uint64_t));
    } else if (hs->d_hashes.alloc((size_t)hs->dev_total * sizeof(uint32_t), ctx->pool) != NB200_OK ||
               hs->d_ts.alloc((size_t)hs->dev_total * sizeof(uint64_t), ctx->pool) != NB200_OK)
        return fail(NB200_ERR_CUDA);
    cudaStream_t st = ctx->stream;
#define NB_CUDA_F(expr)                      \
    do {                                     \
        cudaError_t _e = (expr);             \
        if (_e != cudaSuccess) {             \
            set_last_error(#expr, _e);       \
            return fail(NB200_ERR_CUDA);     \
        }                                    \
    } while (0)
    ctx->ev_valid[0] = ctx->ev_valid[1] = false;
    if (hs->dev_total > 0) {
        NB_CUDA_F(cudaMemsetAsync(hs->d_hashes.p, 0, hs->dev_total * sizeof(uint32_t), st));
        NB_CUDA_F(cudaMemsetAsync(hs->d_ts.p, 0, hs->dev_total * sizeof(uint64_t), st));
    }
    if (frames > 0) {
        const size_t b_seg = (size_t)nseg * sizeof(FpSegment);
        if (ctx->d_fp_desc.reserve(b_seg) != NB200_OK || ctx->h_fp_stage.reserve(b_seg) != NB200_OK ||
            ctx->d_chroma.reserve((size_t)(frames + 8) * FP_BANDS * sizeof(float)) != NB200_OK)
            return fail(NB200_ERR_CUDA);
        // a deferred-wait call may still be reading the staging area and the chroma scratch
        NB_CUDA_F(cudaEventSynchronize(ctx->ev_chunk[11]));
        memcpy(ctx->h_fp_stage.p, segs.data(), b_seg);
        NB_CUDA_F(cudaMemcpyAsync(ctx->d_fp_desc.p, ctx->h_fp_stage.p, b_seg, cudaMemcpyHostToDevice, st));

        K1Args k1;
        k1.pcm = ps->d_pcm.as<int16_t>();
        k1.segs = ctx->d_fp_desc.as<FpSegment>();
        k1.n_segs = nseg;
        k1.window = ctx->d_fp_window.as<float2>();
        k1.twiddle = ctx->d_fp_twiddle.as<float2>();
        k1.lane_w = ctx->d_fp_lane.as<float2>();
        k1.wsplit = ctx->d_fp_wsplit.as<float2>();
        k1.window_odd = ctx->d_fp_window_odd.as<float2>();
        k1.chroma = ctx->d_chroma.as<float>();
        // K1 over global frames [f0, f1)
        auto launch_k1 = [&](uint32_t f0, uint32_t f1) -> cudaError_t {
            if (f1 <= f0) return cudaSuccess;
            const uint32_t nf = f1 - f0;
            const int variant = ctx->k1_variant == 0 ? 19 : ctx->k1_variant;   // default: tensor-memory kernel, 16 warps per SM, revision 4
            if (variant >= 16) {
                // tensor-memory kernel: 16 (or, variant 112, 12) warps per SM
                k1.frame_begin = f0;
                k1.frame_end = f1;
                const uint32_t nw = variant == 112 ? 12u : 16u;
                const uint32_t grid1 = std::min<uint32_t>((uint32_t)ctx->sm_count, (nf + nw - 1) / nw);
                if (variant == 19)
                    fp_fft_chroma_tm_kernel<16, 4><<<grid1, 16 * 32, tm_smem_bytes<16>(), st>>>(k1);
                else if (variant == 18)
                    fp_fft_chroma_tm_kernel<16, 3><<<grid1, 16 * 32, tm_smem_bytes<16>(), st>>>(k1);
                else if (variant == 17)
                    fp_fft_chroma_tm_kernel<16, 2><<<grid1, 16 * 32, tm_smem_bytes<16>(), st>>>(k1);
                else if (nw == 16)
                    fp_fft_chroma_tm_kernel<16, 1><<<grid1, 16 * 32, tm_smem_bytes<16>(), st>>>(k1);
                else
                    fp_fft_chroma_tm_kernel<12, 1><<<grid1, 12 * 32, tm_smem_bytes<12>(), st>>>(k1);
            } else if (variant >= 8) {
                k1.frame_begin = f0;
                k1.frame_end = f1;
                const uint32_t nw = (uint32_t)variant;
                const uint32_t grid1 = std::min<uint32_t>((uint32_t)ctx->sm_count, (nf + nw - 1) / nw);
                if (nw == 12)
                    fp_fft_chroma_h32_kernel<12><<<grid1, 12 * 32, h32_smem_bytes<12>(), st>>>(k1);
                else if (nw == 10)
                    fp_fft_chroma_h32_kernel<10><<<grid1, 10 * 32, h32_smem_bytes<10>(), st>>>(k1);
                else
                    fp_fft_chroma_h32_kernel<8><<<grid1, 8 * 32, h32_smem_bytes<8>(), st>>>(k1);
            } else {
                k1.frame_begin = f0;
                k1.frame_end = f1;
                const uint32_t grid1 = std::min<uint32_t>((uint32_t)ctx->sm_count, (nf + FP_WARPS - 1) / FP_WARPS);
                fp_fft_chroma_kernel<<<grid1, FP_WARPS * 32, K1_SMEM_BYTES, st>>>(k1);
            }
            ctx->n_launches++;
            return cudaGetLastError();
        };
        NB_CUDA_F(cudaEventRecord(ctx->ev[0], st));
        if (!host_pcm) {
            NB_CUDA_F(launch_k1(0, (uint32_t)frames));
        } else {
            // the copy stream must not overtake earlier work on the compute stream that
            // may still read the (pooled) PCM buffer
            NB_CUDA_F(cudaEventRecord(ctx->ev_chunk[7], st));
            NB_CUDA_F(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chunk[7], 0));
            // chunks of consecutive segments, shrinking towards the end: the copy is the long pole
            // (PCIe), so what matters is how little K1 work is left when the last byte lands
            constexpr int N_CHUNKS = 6;
            static const uint32_t CUM_PERMILLE[N_CHUNKS] = {300, 550, 750, 880, 960, 1000};
            uint64_t total_mono = 0;
            for (uint32_t s2 = 0; s2 < nseg; s2++) total_mono += ps->n_mono[s2];
            uint32_t seg = 0;
            uint64_t done = 0;
            for (int c = 0; c < N_CHUNKS && seg < nseg; c++) {
                const uint64_t target = total_mono / 1000 * CUM_PERMILLE[c];
                const uint32_t first = seg;
                while (seg < nseg && (done < target || c == N_CHUNKS - 1)) {
                    if (ps->n_mono[seg] > 0)
                        NB_CUDA_F(cudaMemcpyAsync(ps->d_pcm.as<int16_t>() + ps->dev_offset[seg], host_pcm[seg],
                                                  ps->n_mono[seg] * sizeof(int16_t), cudaMemcpyHostToDevice,
                                                  ctx->copy_stream));
                    done += ps->n_mono[seg];
                    seg++;
                }
                if (seg == first) continue;
                NB_CUDA_F(cudaEventRecord(ctx->ev_chunk[c], ctx->copy_stream));
                NB_CUDA_F(cudaStreamWaitEvent(st, ctx->ev_chunk[c], 0));
                const uint32_t f0 = segs[first].frame_begin;
                const uint32_t f1 = segs[seg - 1].frame_begin + segs[seg - 1].n_frames;
                NB_CUDA_F(launch_k1(f0, f1));
            }
        }
        NB_CUDA_F(cudaEventRecord(ctx->ev[1], st));
        ctx->ev_valid[0] = true;

        if (tiles > 0) {
            K2Args k2;
            k2.chroma = ctx->d_chroma.as<float>();
            k2.segs = ctx->d_fp_desc.as<FpSegment>();
            k2.n_segs = nseg;
            k2.n_tiles = (uint32_t)tiles;
            k2.stride = stride;
            k2.item_secs = as_secs_f32(item_ns);
            k2.delay_ns = delay_ns;
            k2.out_hash = hs->d_hashes.as<uint32_t>();
            k2.out_ts = hs->d_ts.as<uint64_t>();
            NB_CUDA_F(cudaEventRecord(ctx->ev[2], st));
            fp_classify_kernel<<<(uint32_t)tiles, K2_TILE, 0, st>>>(k2);
            NB_CUDA_F(cudaGetLastError());
            NB_CUDA_F(cudaEventRecord(ctx->ev[3], st));
            ctx->ev_valid[1] = true;
            ctx->n_launches++;
        }
        // the caller's PCM (pipelined copy) and h_fp_stage are free again once the stream drains; a
        // caller that goes on to search the season on the same stream waits once, at the very end
        NB_CUDA_F(cudaEventRecord(ctx->ev_chunk[11], st));
        if (wait) NB_CUDA_F(cudaStreamSynchronize(st));
    }
#undef NB_CUDA_F
    *out = hs;
    return NB200_OK;
}

int fingerprint_run_impl(nb200_ctx *ctx, const nb200_pcmset *ps, uint32_t stride, uint64_t delay_ns,
                         uint64_t item_ns, const uint64_t *seek_to_ns, nb200_hashset **out) {
    return fingerprint_run_pipelined(ctx, ps, nullptr, stride, delay_ns, item_ns, seek_to_ns, out);
}

}  // namespace nb200

using namespace nb200;

struct nb200_fp {
    nb200_ctx *ctx = nullptr;
    bool started = false, finished = false;
    int channels = 0;
    PinnedBuf buf;            // interleaved i16 as fed
    size_t n_samples = 0;     // i16 values accumulated
    std::vector<uint32_t> raw;
};

extern "C" {

uint64_t nb200_num_raw_hashes(uint64_t n_mono_samples) { return num_raw(n_mono_samples); }

// copy == false: only lay the set out and allocate it (the caller streams the PCM in itself)
static int pcmset_make(nb200_ctx *ctx, const int16_t *const *pcm, const uint64_t *n_samples_total,
                       int channels, uint32_t n_segments, bool copy, nb200_pcmset **out) {
    if (!ctx || !out || (n_segments && (!pcm || !n_samples_total))) return NB200_ERR_NULL_ARGUMENT;
    *out = nullptr;
    if (channels != 1 && channels != 2) return NB200_ERR_INVALID_ARGUMENT;
    NB_CUDA(cudaSetDevice(ctx->device));
    PhaseTimer pt(ctx, NB200_PHASE_PCM_UPLOAD);
    nb200_pcmset *ps = new (std::nothrow) nb200_pcmset();
    if (!ps) return NB200_ERR_TOO_LARGE;
    ps->n_segments = n_segments;
    ps->n_mono.resize(n_segments);
    ps->dev_offset.resize(n_segments);
    uint64_t dev = 0, max_seg = 0;
    for (uint32_t s = 0; s < n_segments; s++) {
        if (n_samples_total[s] % (uint64_t)channels != 0 || (n_samples_total[s] && !pcm[s])) {
            delete ps;
            return NB200_ERR_INVALID_ARGUMENT;
        }
        ps->n_mono[s] = n_samples_total[s] / (uint64_t)channels;
        ps->dev_offset[s] = dev;
        dev += (ps->n_mono[s] + 7) & ~7ull;
        max_seg = std::max(max_seg, n_samples_total[s]);
    }
    ps->dev_total = dev;
    auto fail = [&](int st, const char *what, cudaError_t e) {
        if (what) set_last_error(what, e);
        delete ps;
        return st;
    };
    if (ps->d_pcm.alloc((size_t)(dev + 8) * sizeof(int16_t), ctx->pool) != NB200_OK) return fail(NB200_ERR_CUDA, nullptr, cudaSuccess);
    cudaStream_t st = ctx->stream;
    cudaError_t e;
    if (channels == 2 && ctx->d_stereo.reserve((size_t)max_seg * sizeof(int16_t)) != NB200_OK)
        return fail(NB200_ERR_CUDA, nullptr, cudaSuccess);
    for (uint32_t s = 0; copy && s < n_segments; s++) {
        if (ps->n_mono[s] == 0) continue;
        int16_t *dst = ps->d_pcm.as<int16_t>() + ps->dev_offset[s];
        if (channels == 1) {
            e = cudaMemcpyAsync(dst, pcm[s], ps->n_mono[s] * sizeof(int16_t), cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return fail(NB200_ERR_CUDA, "cudaMemcpyAsync(pcm)", e);
        } else {
            e = cudaMemcpyAsync(ctx->d_stereo.p, pcm[s], n_samples_total[s] * sizeof(int16_t),
                                cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return fail(NB200_ERR_CUDA, "cudaMemcpyAsync(stereo pcm)", e);
            const uint64_t n = ps->n_mono[s];
            downmix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ctx->d_stereo.as<int16_t>(), dst, n);
            e = cudaGetLastError();
            if (e != cudaSuccess) return fail(NB200_ERR_CUDA, "downmix_kernel", e);
            ctx->n_launches++;
        }
    }
    e = cudaStreamSynchronize(st);   // the caller may reuse its buffers
    if (e != cudaSuccess) return fail(NB200_ERR_CUDA, "cudaStreamSynchronize", e);
    *out = ps;
    return NB200_OK;
}

int nb200_pcmset_upload(nb200_ctx *ctx, const int16_t *const *pcm, const uint64_t *n_samples_total,
                        int channels, uint32_t n_segments, nb200_pcmset **out) {
    return pcmset_make(ctx, pcm, n_samples_total, channels, n_segments, true, out);
}

void nb200_pcmset_free(nb200_pcmset *ps) { delete ps; }

int nb200_pcmset_view(nb200_ctx *ctx, const void *d_pcm, const uint64_t *dev_offset, const uint64_t *n_mono_samples,
                      uint32_t n_segments, uint64_t capacity, nb200_pcmset **out) {
    if (!ctx || !out || (n_segments && (!d_pcm || !dev_offset || !n_mono_samples))) return NB200_ERR_NULL_ARGUMENT;
    *out = nullptr;
    if (reinterpret_cast<uintptr_t>(d_pcm) % 16 != 0) return NB200_ERR_INVALID_ARGUMENT;
    nb200_pcmset *ps = new (std::nothrow) nb200_pcmset();
    if (!ps) return NB200_ERR_TOO_LARGE;
    ps->n_segments = n_segments;
    ps->n_mono.assign(n_mono_samples, n_mono_samples + n_segments);
    ps->dev_offset.assign(dev_offset, dev_offset + n_segments);
    uint64_t end = 0;
    for (uint32_t s = 0; s < n_segments; s++) {
        // the bulk copy of a frame reads 16-byte units: up to 16 bytes beyond the frame's last sample
        if (dev_offset[s] + n_mono_samples[s] + 8 > capacity) {
            delete ps;
            return NB200_ERR_INVALID_ARGUMENT;
        }
        end = std::max(end, dev_offset[s] + n_mono_samples[s]);
    }
    ps->dev_total = end;
    ps->d_pcm.borrow(const_cast<void *>(d_pcm), (size_t)capacity * sizeof(int16_t));
    *out = ps;
    return NB200_OK;
}

int nb200_fingerprint_run(nb200_ctx *ctx, const nb200_pcmset *ps, uint32_t stride, uint64_t delay_ns,
                          uint64_t item_ns, const uint64_t *seek_to_ns, nb200_hashset **out) {
    if (!ctx || !ps || !out) return NB200_ERR_NULL_ARGUMENT;
    *out = nullptr;
    return fingerprint_run_impl(ctx, ps, stride, delay_ns, item_ns, seek_to_ns, out);
}

int nb200_fingerprint_layout(const uint64_t *n_mono_samples, uint32_t n_segments, uint32_t stride,
                             uint64_t *dev_offset_out, uint64_t *seg_len_out, uint64_t *total_out) {
    if (!n_mono_samples && n_segments) return NB200_ERR_NULL_ARGUMENT;
    if (stride == 0) return NB200_ERR_INVALID_ARGUMENT;
    uint64_t dev = 0;
    for (uint32_t s = 0; s < n_segments; s++) {
        const uint64_t len = (num_raw(n_mono_samples[s]) + stride - 1) / stride;
        if (dev_offset_out) dev_offset_out[s] = dev;
        if (seg_len_out) seg_len_out[s] = len;
        dev += (len + 3) & ~3ull;
    }
    if (total_out) *total_out = dev;
    return NB200_OK;
}

int nb200_fingerprint_run_into(nb200_ctx *ctx, const nb200_pcmset *ps, uint32_t stride,
                               uint64_t delay_ns, uint64_t item_ns, const uint64_t *seek_to_ns,
                               void *d_hashes_out, void *d_ts_ns_out, uint64_t capacity) {
    if (!ctx || !ps || !d_hashes_out || !d_ts_ns_out) return NB200_ERR_NULL_ARGUMENT;
    nb200_hashset *hs = nullptr;
    NB_TRY(fingerprint_run_pipelined(ctx, ps, nullptr, stride, delay_ns, item_ns, seek_to_ns, &hs,
                                     d_hashes_out, d_ts_ns_out, capacity, !ctx->defer_wait));
    nb200_hashset_free(hs);   // borrowed arrays: nothing is released
    return NB200_OK;
}

int nb200_timestamps_fill(nb200_ctx *ctx, void *d_ts_ns, const uint64_t *dev_offset, const uint64_t *seg_len,
                          const uint64_t *seek_to_ns, uint32_t n_segments, uint32_t stride, uint64_t delay_ns,
                          uint64_t item_ns) {
    if (!ctx || !d_ts_ns || !dev_offset || !seg_len) return NB200_ERR_NULL_ARGUMENT;
    if (stride == 0) return NB200_ERR_INVALID_ARGUMENT;
    if (n_segments == 0) return NB200_OK;
    NB_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)n_segments * sizeof(TsSegment);
    NB_TRY(ctx->h_tsfill.reserve(bytes));
    NB_TRY(ctx->d_tsfill.reserve(bytes));
    // the pinned staging may still be in flight from the previous call on this stream
    NB_CUDA(cudaEventSynchronize(ctx->ev_chunk[6]));
    TsSegment *tab = ctx->h_tsfill.as<TsSegment>();
    uint64_t max_len = 1;
    for (uint32_t s = 0; s < n_segments; s++) {
        tab[s] = {dev_offset[s], seg_len[s], seek_to_ns ? seek_to_ns[s] : 0};
        max_len = std::max(max_len, seg_len[s]);
    }
    NB_CUDA(cudaMemcpyAsync(ctx->d_tsfill.p, tab, bytes, cudaMemcpyHostToDevice, ctx->stream));
    NB_CUDA(cudaEventRecord(ctx->ev_chunk[6], ctx->stream));
    ts_fill_kernel<<<dim3((unsigned)((max_len + 255) / 256), n_segments), 256, 0, ctx->stream>>>(
        (uint64_t *)d_ts_ns, ctx->d_tsfill.as<TsSegment>(), stride, as_secs_f32(item_ns), delay_ns);
    NB_CUDA(cudaGetLastError());
    ctx->n_launches++;
    return NB200_OK;
}

int nb200_fingerprint_host_into(nb200_ctx *ctx, const int16_t *const *pcm, const uint64_t *n_samples,
                                uint32_t n_segments, uint32_t stride, uint64_t delay_ns, uint64_t item_ns,
                                const uint64_t *seek_to_ns, void *d_hashes_out, void *d_ts_ns_out,
                                uint64_t capacity) {
    if (!ctx || !pcm || !n_samples || !d_hashes_out || !d_ts_ns_out) return NB200_ERR_NULL_ARGUMENT;
    nb200_pcmset *ps = nullptr;
    NB_TRY(pcmset_make(ctx, pcm, n_samples, 1, n_segments, false, &ps));
    int st = ensure_fp_tables(ctx);
    nb200_hashset *hs = nullptr;
    if (st == NB200_OK)
        st = fingerprint_run_pipelined(ctx, ps, pcm, stride, delay_ns, item_ns, seek_to_ns, &hs, d_hashes_out,
                                       d_ts_ns_out, capacity, !ctx->defer_wait);
    nb200_hashset_free(hs);
    // the PCM block goes back to the context's pool; whoever takes it next is ordered behind
    // these kernels (same stream, and the copy stream waits for the compute stream first)
    nb200_pcmset_free(ps);
    return st;
}

int nb200_fingerprint_batch(nb200_ctx *ctx, const int16_t *const *pcm, const uint64_t *n_samples_total,
                            int channels, uint32_t n_segments, uint32_t stride, uint32_t **out_hashes,
                            uint64_t *out_counts) {
    if (!ctx || !out_hashes || !out_counts) return NB200_ERR_NULL_ARGUMENT;
    for (uint32_t s = 0; s < n_segments; s++) {
        out_hashes[s] = nullptr;
        out_counts[s] = 0;
    }
    if (n_segments == 0) return NB200_OK;
    // the staged pipeline works on (opening, ending) segment pairs: pad to even
    std::vector<const int16_t *> p(pcm, pcm + n_segments);
    std::vector<uint64_t> c(n_samples_total, n_samples_total + n_segments);
    if (n_segments & 1) {
        p.push_back(nullptr);
        c.push_back(0);
    }
    nb200_pcmset *ps = nullptr;
    NB_TRY(nb200_pcmset_upload(ctx, p.data(), c.data(), channels, (uint32_t)p.size(), &ps));
    nb200_hashset *hs = nullptr;
    int st = fingerprint_run_impl(ctx, ps, stride, 0, 0, nullptr, &hs);
    nb200_pcmset_free(ps);
    if (st != NB200_OK) return st;
    std::vector<uint32_t> h(std::max<uint64_t>(hs->total, 1));
    std::vector<uint64_t> t(std::max<uint64_t>(hs->total, 1));
    st = nb200_hashset_download(ctx, hs, h.data(), t.data());
    if (st == NB200_OK) {
        for (uint32_t s = 0; s < n_segments; s++) {
            const uint64_t n = hs->seg_len[s];
            uint32_t *o = (uint32_t *)malloc(std::max<uint64_t>(n, 1) * sizeof(uint32_t));
            if (!o) {
                st = NB200_ERR_TOO_LARGE;
                break;
            }
            memcpy(o, h.data() + hs->seg_offset[s], n * sizeof(uint32_t));
            out_hashes[s] = o;
            out_counts[s] = n;
        }
        if (st != NB200_OK)
            for (uint32_t s = 0; s < n_segments; s++) {
                free(out_hashes[s]);
                out_hashes[s] = nullptr;
                out_counts[s] = 0;
            }
    }
    nb200_hashset_free(hs);
    return st;
}

/* ---------------------------------------------- chromaprint-shaped shim */

int nb200_fp_new(nb200_ctx *ctx, nb200_fp **out) {
    if (!ctx || !out) return NB200_ERR_NULL_ARGUMENT;
    nb200_fp *fp = new (std::nothrow) nb200_fp();
    if (!fp) return NB200_ERR_TOO_LARGE;
    fp->ctx = ctx;
    *out = fp;
    return NB200_OK;
}

void nb200_fp_free(nb200_fp *fp) { delete fp; }

int nb200_fp_sample_rate(const nb200_fp *) { return FP_SAMPLE_RATE; }

int nb200_fp_start(nb200_fp *fp, int sample_rate, int channels) {
    if (!fp) return NB200_ERR_NULL_ARGUMENT;
    // needle resamples to the context's own rate before feeding (analyzer.rs:179-187);
    // there is no resampler on this side of the boundary
    if (sample_rate != FP_SAMPLE_RATE || (channels != 1 && channels != 2)) return NB200_ERR_INVALID_ARGUMENT;
    fp->started = true;
    fp->finished = false;
    fp->channels = channels;
    fp->n_samples = 0;
    fp->raw.clear();
    return NB200_OK;
}

int nb200_fp_feed(nb200_fp *fp, const int16_t *data, size_t n_samples_total) {
    if (!fp || (n_samples_total && !data)) return NB200_ERR_NULL_ARGUMENT;
    if (!fp->started || fp->finished) return NB200_ERR_STATE;
    const size_t need = (fp->n_samples + n_samples_total) * sizeof(int16_t);
    if (need > fp->buf.bytes) {
        // grow the pinned buffer, keeping what was fed so far
        PinnedBuf bigger;
        NB_TRY(bigger.reserve(std::max<size_t>(need * 2, 1u << 20)));
        if (fp->n_samples) memcpy(bigger.p, fp->buf.p, fp->n_samples * sizeof(int16_t));
        std::swap(fp->buf.p, bigger.p);
        std::swap(fp->buf.bytes, bigger.bytes);
    }
    memcpy(fp->buf.as<int16_t>() + fp->n_samples, data, n_samples_total * sizeof(int16_t));
    fp->n_samples += n_samples_total;
    return NB200_OK;
}

int nb200_fp_finish(nb200_fp *fp) {
    if (!fp) return NB200_ERR_NULL_ARGUMENT;
    if (!fp->started || fp->finished) return NB200_ERR_STATE;
    // a trailing odd sample of a stereo stream cannot form a frame: dropped, as
    // Chromaprint's AudioProcessor consumes whole frames only
    const uint64_t usable = fp->n_samples - fp->n_samples % (size_t)fp->channels;
    const int16_t *p = fp->buf.as<int16_t>();
    uint32_t *out = nullptr;
    uint64_t count = 0;
    NB_TRY(nb200_fingerprint_batch(fp->ctx, &p, &usable, fp->channels, 1, 1, &out, &count));
    fp->raw.assign(out, out + count);
    free(out);
    fp->finished = true;
    return NB200_OK;
}

int nb200_fp_get_delay_ms(const nb200_fp *fp, int *out) {
    if (!fp || !out) return NB200_ERR_NULL_ARGUMENT;
    *out = NB200_DELAY_MS;
    return NB200_OK;
}

int nb200_fp_get_item_duration_ms(const nb200_fp *fp, int *out) {
    if (!fp || !out) return NB200_ERR_NULL_ARGUMENT;
    *out = NB200_ITEM_DURATION_MS;
    return NB200_OK;
}

int nb200_fp_get_raw(const nb200_fp *fp, const uint32_t **hashes, size_t *n) {
    if (!fp || !hashes || !n) return NB200_ERR_NULL_ARGUMENT;
    if (!fp->finished) return NB200_ERR_STATE;
    *hashes = fp->raw.data();
    *n = fp->raw.size();
    return NB200_OK;
}

/* ------------------------------------------------ fused analyze + search */

int nb200_search_hashset(nb200_ctx *ctx, const nb200_hashset *hs, const uint64_t *hash_duration_ns,
                         const nb200_match_params *params, nb200_search_result *results);

int nb200_analyze_search(nb200_ctx *ctx, const int16_t *const *pcm, const uint64_t *n_samples_total,
                         int channels, uint32_t n_videos, const uint64_t *seek_to_ns,
                         uint64_t hash_duration_ns, const nb200_match_params *params,
                         nb200_search_result *results) {
    if (!ctx || !pcm || !n_samples_total || !params || !results) return NB200_ERR_NULL_ARGUMENT;
    if (n_videos < 2) return NB200_ERR_COMPARATOR_MINIMUM_PATHS;
    // step_by = hash_duration.as_millis() / item_duration.as_millis()   (analyzer.rs:293-297)
    const uint64_t stride = (hash_duration_ns / 1000000ull) / NB200_ITEM_DURATION_MS;
    if (stride == 0 || stride > 0xffffffffull) return NB200_ERR_INVALID_ARGUMENT;
    nb200_pcmset *ps = nullptr;
    nb200_hashset *hs = nullptr;
    int st;
    if (channels == 1) {
        // mono: stream the PCM in chunk by chunk underneath K1
        NB_TRY(pcmset_make(ctx, pcm, n_samples_total, channels, 2 * n_videos, false, &ps));
        st = ensure_fp_tables(ctx);
        if (st != NB200_OK) {
            nb200_pcmset_free(ps);
            return st;
        }
        st = fingerprint_run_pipelined(ctx, ps, pcm, (uint32_t)stride, NB200_DELAY_MS * 1000000ull,
                                       NB200_ITEM_DURATION_MS * 1000000ull, seek_to_ns, &hs, nullptr, nullptr, 0,
                                       /*wait=*/false);
    } else {
        NB_TRY(nb200_pcmset_upload(ctx, pcm, n_samples_total, channels, 2 * n_videos, &ps));
        st = fingerprint_run_pipelined(ctx, ps, nullptr, (uint32_t)stride, NB200_DELAY_MS * 1000000ull,
                                       NB200_ITEM_DURATION_MS * 1000000ull, seek_to_ns, &hs, nullptr, nullptr, 0,
                                       /*wait=*/false);
    }
    if (st != NB200_OK) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->copy_stream);
        nb200_pcmset_free(ps);
        return st;
    }
    // the match and the vote queue up behind K1/K2 on the same stream; nb200_search_hashset
    // synchronises once, after which the PCM, the season and the caller's buffers are free
    std::vector<uint64_t> hd(n_videos, hash_duration_ns);
    st = nb200_search_hashset(ctx, hs, hd.data(), params, results);
    if (st != NB200_OK) cudaStreamSynchronize(ctx->stream);
    nb200_pcmset_free(ps);
    nb200_hashset_free(hs);
    return st;
}

/* Analyzer + Comparator on PCM that is already resident in HBM: K1, K2, the match, simhash
 * and the vote are enqueued back to back; one synchronisation at the end. */
int nb200_pcmset_search(nb200_ctx *ctx, const nb200_pcmset *ps, const uint64_t *seek_to_ns,
                        uint64_t hash_duration_ns, const nb200_match_params *params,
                        nb200_search_result *results) {
    if (!ctx || !ps || !params || !results) return NB200_ERR_NULL_ARGUMENT;
    if (ps->n_segments < 4 || ps->n_segments % 2 != 0) return NB200_ERR_COMPARATOR_MINIMUM_PATHS;
    const uint64_t stride = (hash_duration_ns / 1000000ull) / NB200_ITEM_DURATION_MS;
    if (stride == 0 || stride > 0xffffffffull) return NB200_ERR_INVALID_ARGUMENT;
    nb200_hashset *hs = nullptr;
    NB_TRY(fingerprint_run_pipelined(ctx, ps, nullptr, (uint32_t)stride, NB200_DELAY_MS * 1000000ull,
                                     NB200_ITEM_DURATION_MS * 1000000ull, seek_to_ns, &hs, nullptr, nullptr, 0,
                                     /*wait=*/false));
    std::vector<uint64_t> hd(ps->n_segments / 2, hash_duration_ns);
    int st = nb200_search_hashset(ctx, hs, hd.data(), params, results);
    if (st != NB200_OK) cudaStreamSynchronize(ctx->stream);
    nb200_hashset_free(hs);
    return st;
}

}  // extern "C"
