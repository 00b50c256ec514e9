// match.cu -- K3/K4: the Comparator's all-pairs sliding match on sm_100a.
//
// Replaces Comparator::longest_common_hash_match
// (needle/src/audio/comparator.rs:157-250) for every pair of the pair loop
// (comparator.rs:532-578).  The reference fills an (n+1)x(m+1) usize table
// (table[i][j] = table[i-1][j-1]+1 iff popcount(src[i]^dst[j]) <= T, rows and
// columns 0 forced to 0, :175-187) and scans it backwards for run ends
// (:191-200).  A run lives on one diagonal d = i - j, so no table is needed:
// this kernel walks diagonals, packs the match predicate of 32 consecutive
// cells of a diagonal into one word, and extracts maximal runs from the words.
//
// Work decomposition (all integer, CUDA cores; 1 LOP3 + 1 POPC per cell):
//   problem   = one table (pair x {opening, ending})
//   row chunk = MATCH_R consecutive src rows
//   warp tile = one row chunk x 32*C adjacent diagonals; lane l owns the C
//               diagonals dband + l*C .. dband + l*C + C-1
// Persistent warps pull warp tiles from a global counter.  Per 32-row step a
// lane holds the 32 src hashes of the step (warp-uniform) and a sliding window
// of 32+C-1 dst hashes in registers (both staged through shared memory), so a
// cell costs no memory instruction.
//
// Runs that cross a row-chunk boundary are owned by the tile in which they
// START: a tile ignores a run that is already active in the row above its
// first row, and keeps walking past its last row while a run it owns is still
// active.  That makes tiles independent and the result exact.
//
// The duration filter of :212-223 is evaluated exactly, on the u64 ns
// timestamps, for every run that reaches `lcand` cells -- a per-problem lower
// bound on the length of any qualifying run computed from the same
// timestamps on the device (seg_lmin_kernel).  Runs shorter than that cannot
// pass the filter, so skipping them changes nothing.
#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

#include "common.h"
#include "tma.cuh"

namespace nb200 {

constexpr int MATCH_R = 256;          // rows per chunk
constexpr int MATCH_C = 8;            // diagonals per lane
constexpr int MATCH_DW = 32 * MATCH_C;  // diagonals per warp tile
constexpr int MATCH_WARPS = 4;        // warps per CTA (independent workers)
constexpr int SRC_TILE = MATCH_R + 8;             // rows r0-4 .. r0+R+3
constexpr int DST_TILE = MATCH_R + MATCH_DW + 8;  // cols tile_j0 .. tile_j0+R+263
constexpr int LUT_SHIFT = 7;
constexpr int FAST_DST_TILE = 2 * MATCH_R;        // fast kernel: s = 255 + kk - x in [0, 510]
constexpr int FOREIGN = -(1 << 30);   // carry marker: run started in an earlier chunk
constexpr int COOP_MAX_LANES = 8;     // fast kernel: at most this many lanes alive -> verify their words cooperatively

struct MatchProblem {
    uint32_t src_off, n;    // element offset into d_hashes / d_ts, length
    uint32_t dst_off, m;
    uint64_t min_ns;        // min_opening_duration or min_ending_duration
    uint32_t lcand;         // >= 1
    uint32_t n_bands;       // warp tiles per row chunk
    uint32_t tile_begin;    // first global tile id
    uint32_t group;         // 2 * (global pair index) + is_ending
    uint32_t src_seg, dst_seg;   // segment indices (2 * video + is_ending): where lcand comes from
    // fast kernel: the dst list's live columns 1 .. m-1, padded to `period` (a multiple of 4) and
    // repeated, so that every wrapped 512-column window is ONE 16-byte-aligned span (dst2_build_kernel)
    uint32_t dst2_off, period;
    uint32_t n_groups;      // fast kernel: units (band groups) per row chunk; tile_begin counts units
};

struct MatchArgs {
    const uint32_t *hashes;
    const uint64_t *ts;
    const MatchProblem *problems;
    const uint32_t *tile_end;   // [n_problems] exclusive end tile id per problem
    const uint32_t *lut;        // [n_tiles >> LUT_SHIFT + 1] first problem whose tile_end > (k << LUT_SHIFT)
    const uint32_t *dst2;       // fast kernel: repeated live columns of every list (MatchProblem::dst2_off)
    uint32_t n_tiles;           // work items under the cursor: tiles (general kernel) or units (fast kernel)
    uint32_t band_group;        // fast kernel: bands per unit
    uint32_t threshold;
    uint32_t run_cap;
    // 0: run; 1: run iff every problem's lcand >= 63 (the fast kernel's precondition);
    // 2: run iff not.  The decision is made on the device (problem_finalize_kernel), so the
    // host never waits for the run-length bounds: it launches both kernels, one returns.
    uint32_t mode;
    // counters[0] = tile cursor, [1] = run count, [2] = error flags, [3] = ~(min lcand)
    unsigned int *counters;
    DeviceRun *runs;
};

constexpr uint32_t FAST_MIN_LCAND = 63;
__device__ __forceinline__ bool not_my_turn(const MatchArgs &a) {
    if (a.mode == 0) return false;
    const uint32_t min_lcand = ~a.counters[3];
    return (a.mode == 1) != (min_lcand >= FAST_MIN_LCAND);
}

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}

__device__ __forceinline__ bool hamming_match(uint32_t a, uint32_t b, uint32_t T) {
    return (uint32_t)__popc(a ^ b) <= T;
}

// What the (rare) slow paths need; passed by value so that the kernel
// parameter block never has its address taken.
struct EmitCtx {
    const uint32_t *hashes;
    const uint64_t *ts;
    const MatchProblem *problems;
    unsigned int *counters;
    DeviceRun *runs;
    uint32_t run_cap;
    uint32_t threshold;
};

// Exact duration filter (comparator.rs:206-223) + emission.
__device__ __noinline__ void check_emit(const EmitCtx a, uint32_t problem, int i_end, int d, int len) {
    const MatchProblem pr = a.problems[problem];
    if (len < (int)pr.lcand) return;
    const int j_end = i_end - d;
    const uint64_t *ts_s = a.ts + pr.src_off;
    const uint64_t *ts_d = a.ts + pr.dst_off;
    const uint64_t s_end = ts_s[i_end], s_start = ts_s[i_end - len];
    const uint64_t d_end = ts_d[j_end], d_start = ts_d[j_end - len];
    if (s_end < s_start || d_end < d_start) {   // Duration - Duration panics in the reference
        atomicOr(&a.counters[2], 1u);
        return;
    }
    if (s_end - s_start >= pr.min_ns && d_end - d_start >= pr.min_ns) {
        unsigned int slot = atomicAdd(&a.counters[1], 1u);
        if (slot < a.run_cap) {
            DeviceRun r;
            r.problem = problem;
            r.i_end = (uint32_t)i_end;
            r.j_end = (uint32_t)j_end;
            r.len = (uint32_t)len;
            r.src_simhash = 0;
            r.dst_simhash = 0;
            r.group = pr.group;
            r.reserved = 0;
            r.src_start_ns = s_start;
            r.src_end_ns = s_end;
            r.dst_start_ns = d_start;
            r.dst_end_ns = d_end;
            a.runs[slot] = r;
        }
    }
}

// All maximal runs that END inside this 32-row word (i.e. are followed by a
// non-matching cell of the word).  v: bit k = match at row i0+k.  carry = length
// of the run active at row i0-1 (FOREIGN if it is not ours).  Returns the new
// carry.  Slow path: only reached when a run may qualify.
__device__ __noinline__ int runs_in_word(const EmitCtx a, uint32_t problem, uint32_t v, int carry,
                                         int i0, int d) {
    int pos = 0;
    int cur = carry;
    while (pos < 32) {
        uint32_t rest = v >> pos;
        uint32_t inv = ~rest;
        int ones = inv ? (__ffs(inv) - 1) : 32;
        if (ones > 32 - pos) ones = 32 - pos;
        if (ones == 32 - pos) {      // run reaches the end of the word
            cur += ones;
            return cur;
        }
        int total = cur + ones;
        if (total > 0) check_emit(a, problem, i0 + pos + ones - 1, d, total);
        cur = 0;
        pos += ones;
        rest = v >> pos;             // bit 0 is now a zero
        int zeros = rest ? (__ffs(rest) - 1) : 32;
        pos += zeros;
    }
    return 0;
}

// A run we own is still active at the last row of the chunk and the cell after
// it matches too: keep walking the diagonal in global memory until it ends.
__device__ __noinline__ void walk_and_emit(const EmitCtx a, uint32_t problem, int i, int d, int len) {
    const MatchProblem pr = a.problems[problem];
    const uint32_t *hs = a.hashes + pr.src_off;
    const uint32_t *hd = a.hashes + pr.dst_off;
    int j = i - d;
    // (i, j) is the first cell not yet examined
    while (i < (int)pr.n && j < (int)pr.m && hamming_match(hs[i], hd[j], a.threshold)) {
        len++;
        i++;
        j++;
    }
    check_emit(a, problem, i - 1, d, len);
}

__global__ void __launch_bounds__(MATCH_WARPS * 32, 4) match_kernel(const MatchArgs a) {
    __shared__ __align__(128) uint32_t s_src[MATCH_WARPS][SRC_TILE];
    __shared__ __align__(128) uint32_t s_dst[MATCH_WARPS][DST_TILE];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t *src_t = s_src[warp];
    uint32_t *dst_t = s_dst[warp];
    const uint32_t T = a.threshold;
    const EmitCtx ectx = {a.hashes, a.ts, a.problems, a.counters, a.runs, a.run_cap, a.threshold};
    const int ybias = -(int)(T + 1);    // popc + ybias < 0  <=>  popc <= T
    if (not_my_turn(a)) return;
    __shared__ __align__(8) uint64_t s_bar[MATCH_WARPS];
    const uint32_t bar = smem_u32(&s_bar[warp]);
    uint32_t phase = 0;
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(&a.counters[0], 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= a.n_tiles) break;

        uint32_t p = a.lut[tile >> LUT_SHIFT];
        while (tile >= a.tile_end[p]) p++;
        const MatchProblem pr = a.problems[p];
        const int n = (int)pr.n, m = (int)pr.m;
        const uint32_t local = tile - pr.tile_begin;
        const int chunk = (int)(local / pr.n_bands);
        const int band = (int)(local - (uint32_t)chunk * pr.n_bands);
        const int r0 = chunk * MATCH_R;
        // diagonals with a valid cell in rows [r0, r0+R): d in [r0-(m-1), r0+R-2];
        // start is lowered to == 1 (mod 4) so that tile_j0 is a multiple of 4
        int dstart = r0 - (m - 1);
        dstart -= (dstart - 1) & 3;
        const int dband = dstart + band * MATCH_DW;
        const int tile_j0 = r0 - dband - (MATCH_DW + 3);   // dst_t[s] = dst[tile_j0 + s]
        const int tile_i0 = r0 - 4;                        // src_t[s] = src[tile_i0 + s]

        // ---- stage the hash tiles with 1-D bulk copies (TMA): both tiles are contiguous, 16-byte
        //      aligned spans of the lists (tile_i0 and tile_j0 are multiples of 4).  Only the part that
        //      lies inside the list is fetched -- out-of-range entries are never used unmasked -- and
        //      its last 1..3 elements (a list need not end on a 16-byte boundary) by plain loads.
        __syncwarp();
        {
            const uint32_t *hs = a.hashes + pr.src_off;
            const uint32_t *hd = a.hashes + pr.dst_off;
            const int si0 = max(tile_i0, 0), si1 = min(tile_i0 + SRC_TILE, n);
            const int sj0 = max(tile_j0, 0), sj1 = min(tile_j0 + DST_TILE, m);
            const int sn = max(si1 - si0, 0) & ~3, dn = max(sj1 - sj0, 0) & ~3;
            fence_proxy_async();
            if (lane == 0) {
                mbar_expect_tx(bar, (uint32_t)(sn + dn) * 4u);
                if (sn) bulk_g2s(smem_u32(src_t + (si0 - tile_i0)), hs + si0, (uint32_t)sn * 4u, bar);
                if (dn) bulk_g2s(smem_u32(dst_t + (sj0 - tile_j0)), hd + sj0, (uint32_t)dn * 4u, bar);
            }
            if (lane < 3 && si0 + sn + lane < si1) src_t[si0 + sn + lane - tile_i0] = hs[si0 + sn + lane];
            if (lane >= 4 && lane < 7 && sj0 + dn + lane - 4 < sj1) dst_t[sj0 + dn + lane - 4 - tile_j0] = hd[sj0 + dn + lane - 4];
            mbar_wait(bar, phase);
            phase ^= 1;
        }
        __syncwarp();

        const int dlane = dband + lane * MATCH_C;   // lane's lowest diagonal; diagonal c -> dlane + c
        int carry[MATCH_C];

        // ---- is a run already active in the row above the chunk?  (then it is not ours)
#pragma unroll
        for (int c = 0; c < MATCH_C; c++) {
            const int d = dlane + c;
            const int i = r0 - 1, j = i - d;
            bool prev = false;
            if (i >= 1 && j >= 1 && j <= m - 1 && i <= n - 1)
                prev = hamming_match(src_t[i - tile_i0], dst_t[j - tile_j0], T);
            carry[c] = prev ? FOREIGN : 0;
        }

        // ---- main loop: 32 rows x C diagonals per lane per step
        // rows past n-1 are masked to "no match"; one (possibly all-zero) word
        // after the last valid row closes every carry
        const int n_steps = min(MATCH_R / 32, (n - r0) / 32 + 1);
        for (int step = 0; step < n_steps; step++) {
            const int i0 = r0 + step * 32;
            // warp-uniform bounds of j over the whole warp tile at this step
            const int jmin_w = i0 - (dband + MATCH_DW - 1);
            const int jmax_w = i0 + 31 - dband;
            uint32_t w[MATCH_C];
            const bool any_valid = (jmax_w >= 1) && (jmin_w <= m - 1) && (i0 <= n - 1);
            if (any_valid) {
                // window of dst hashes: element e <-> dst[wbase + e], e in [0, 32+C-1)
                const int wbase = i0 - dlane - (MATCH_C - 1);
                const uint4 *dptr = reinterpret_cast<const uint4 *>(dst_t + (wbase - tile_j0));
                const uint4 *sptr = reinterpret_cast<const uint4 *>(src_t + (i0 - tile_i0));
                uint32_t dw[40];
#pragma unroll
                for (int q = 0; q < 10; q++) {
                    uint4 v = dptr[q];
                    dw[4 * q + 0] = v.x;
                    dw[4 * q + 1] = v.y;
                    dw[4 * q + 2] = v.z;
                    dw[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int c = 0; c < MATCH_C; c++) w[c] = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const uint4 sv = sptr[q];
                    const uint32_t s4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        const int k = 4 * q + kk;
#pragma unroll
                        for (int c = 0; c < MATCH_C; c++) {
                            // cell (row i0+k, diagonal dlane+c): j = wbase + (C-1) + k - c
                            const int y = __popc(s4[kk] ^ dw[k + (MATCH_C - 1) - c]) + ybias;
                            // shift the sign bit of y in: row k ends up at bit 31-k
                            w[c] = __funnelshift_l((uint32_t)y, w[c], 1);
                        }
                    }
                }
                const bool interior = (i0 >= 1) && (i0 + 31 <= n - 1) && (jmin_w >= 1) && (jmax_w <= m - 1);
                if (!interior) {
#pragma unroll
                    for (int c = 0; c < MATCH_C; c++) {
                        const int d = dlane + c;
                        // valid rows: i in [max(1, d+1), min(n-1, d+m-1)]
                        int klo = max(1, d + 1) - i0;
                        int khi = min(n - 1, d + m - 1) - i0;
                        klo = max(klo, 0);
                        khi = min(khi, 31);
                        uint32_t mask = 0;
                        if (klo <= khi) {
                            // rows klo..khi  <->  bits 31-khi .. 31-klo
                            const uint32_t hi_part = 0xffffffffu >> klo;          // bits 31-klo .. 0
                            const uint32_t lo_part = 0xffffffffu << (31 - khi);   // bits 31 .. 31-khi
                            mask = hi_part & lo_part;
                        }
                        w[c] &= mask;
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < MATCH_C; c++) w[c] = 0;
            }

            // ---- run extraction on the words
#pragma unroll
            for (int c = 0; c < MATCH_C; c++) {
                const uint32_t nw = ~w[c];
                if (nw == 0) {
                    carry[c] += 32;
                    continue;
                }
                const int lead = __clz((int)nw);       // matches in rows 0 .. lead-1
                const bool maybe = (carry[c] + lead >= (int)pr.lcand) ||
                                   (pr.lcand <= 30u && w[c] != 0u);
                if (maybe) {
                    carry[c] = runs_in_word(ectx, p, __brev(w[c]), carry[c], i0, dlane + c);
                } else {
                    carry[c] = __ffs((int)nw) - 1;     // matches at the end of the word
                }
            }
        }

        // ---- runs we own that are still active at the last row of the chunk
        {
            const int i = r0 + MATCH_R;    // first row after the chunk
#pragma unroll
            for (int c = 0; c < MATCH_C; c++) {
                if (carry[c] > 0) {
                    const int d = dlane + c;
                    const int j = i - d;
                    bool next = false;
                    if (i <= n - 1 && j >= 1 && j <= m - 1)
                        next = hamming_match(src_t[i - tile_i0], dst_t[j - tile_j0], T);
                    if (!next) {
                        if (carry[c] >= (int)pr.lcand) check_emit(ectx, p, i - 1, d, carry[c]);
                    } else {
                        walk_and_emit(ectx, p, i + 1, d, carry[c] + 1);
                    }
                }
            }
        }
    }
}

// The adaptive fast kernel's first look at a 32-row word: per lane, the smallest sum of Hamming distances
// over rows 0, 12, 24 among its 8 diagonals.  dl = the lane's dst window of the word, sp = the word's src rows.
__device__ __forceinline__ int first_look(const uint32_t *dl, const uint32_t *sp) {
    const uint4 *dptr = reinterpret_cast<const uint4 *>(dl);
    uint32_t pw[24];
#pragma unroll
    for (int g = 0; g < 3; g++) {
        const uint4 v0 = dptr[3 * g], v1 = dptr[3 * g + 1];
        pw[8 * g + 0] = v0.x;
        pw[8 * g + 1] = v0.y;
        pw[8 * g + 2] = v0.z;
        pw[8 * g + 3] = v0.w;
        pw[8 * g + 4] = v1.x;
        pw[8 * g + 5] = v1.y;
        pw[8 * g + 6] = v1.z;
        pw[8 * g + 7] = v1.w;
    }
    const uint32_t s0 = sp[0], s1 = sp[12], s2 = sp[24];
    const uint32_t s01 = s0 ^ s1;
    int lo = 1 << 20;
#pragma unroll
    for (int c = 0; c < MATCH_C; c++) {
        // x_k = s_k ^ d_k;  t = x0 ^ x1;  ones = t ^ x2;  twos = maj(x0, x1, x2) = t ? ~ones : x0
        const uint32_t d0 = pw[(MATCH_C - 1) - c], d1 = pw[8 + (MATCH_C - 1) - c], d2 = pw[16 + (MATCH_C - 1) - c];
        const uint32_t t = lop3<0x96>(d0, d1, s01);
        const uint32_t ones = lop3<0x96>(t, d2, s2);
        const uint32_t twos = lop3<0x74>(d0 ^ s0, t, ones);
        lo = min(lo, __popc(ones) + 2 * __popc(twos));
    }
    return lo;
}

// ---------------------------------------------------------------------------
// Fast variant for lcand >= 63 (every default configuration: 20 s / 0.246 s = 82
// cells).  A run that long contains at least one ALIGNED 32-row word of its
// diagonal in which all 32 cells match, so the hot loop only has to find such
// words: per (32 rows x diagonal) the MAXIMUM Hamming distance, one 3-input
// max per two cells -- no match words, no run bookkeeping, no FLO/BREV on the
// XU pipe that POPC needs.  A full word is handed to the whole warp, which
// walks the diagonal backwards (is this the run's first full word? only that
// one owns the run) and forwards (where does it end?) 32 cells per ballot.
__device__ __noinline__ void coop_long_run(const EmitCtx a, uint32_t problem, int i0, int d, int lane) {
    const MatchProblem pr = a.problems[problem];
    const uint32_t *hs = a.hashes + pr.src_off;
    const uint32_t *hd = a.hashes + pr.dst_off;
    const int n = (int)pr.n, m = (int)pr.m;
    // cells (i0-1-lane, .) : contiguous matches going up from the word
    bool mt = false;
    {
        const int i = i0 - 1 - lane, j = i - d;
        if (i >= 1 && j >= 1 && j <= m - 1) mt = hamming_match(hs[i], hd[j], a.threshold);
    }
    const uint32_t up = __ballot_sync(0xffffffffu, mt);
    if (up == 0xffffffffu) return;                 // the aligned word above is full too: it owns the run
    const int above = __ffs((int)~up) - 1;
    const int i_first = i0 - above;
    int len = 32 + above;
    int base = i0 + 32;
    for (;;) {
        const int i = base + lane, j = i - d;
        bool f = false;
        if (i <= n - 1 && j >= 1 && j <= m - 1) f = hamming_match(hs[i], hd[j], a.threshold);
        const uint32_t fw = __ballot_sync(0xffffffffu, f);
        if (fw == 0xffffffffu) {
            len += 32;
            base += 32;
            continue;
        }
        len += __ffs((int)~fw) - 1;
        break;
    }
    // first matching row = i0 - (cells above the word); last = first + len - 1
    if (lane == 0) check_emit(a, problem, i_first + len - 1, d, len);
}

// A full word (32 matching cells of wrapped diagonal u = ub + x starting at row i0) was found:
// drop the look-alikes the wrapped tiling can produce, then let the warp walk the run.
__device__ __forceinline__ void full_word(const EmitCtx &ectx, uint32_t p, int i0, int u, int Mp, int M, int n, int lane) {
    if (u >= Mp || i0 == 0 || i0 + 31 > n - 1) return;    // duplicate diagonal / row 0 / past the end
    int jm = (i0 - 1 - u) % Mp;                             // j0 - 1
    if (jm < 0) jm += Mp;
    if (jm + 31 > M - 1) return;                            // the wrap point or a padding column is inside the word
    coop_long_run(ectx, p, i0, i0 - (jm + 1), lane);
}

// Few diagonals of this 32-row word are still alive after a sampling stage: instead of running
// the remaining stages over all 256 diagonals of the tile, the warp tests each survivor's whole
// word at once -- lane r takes row r -- which is exact ("all 32 cells match") and costs a dozen
// instructions per survivor against ~90 per stage.  On real audio unrelated frames of stationary
// background match with p ~ 0.15-0.2 and neighbouring cells of a diagonal correlate (0.4 at 8
// rows), so a word rarely dies in the first stage.  Measured on the bench season
// (profiles/r01_ncu_k3_match_real_v4.txt): stage 1 runs for 66 % of the words, stage 2 for 31 %,
// later stages for < 4 %; this path takes the others.  It is a serial chain per survivor
// (load, XOR, POPC, vote), so it only pays for a handful: with 16 lanes alive it loses to
// another stage.
__device__ __noinline__ void coop_verify(const EmitCtx ectx, uint32_t p, const uint32_t *src_t, const uint32_t *dst_t,
                                         int step, int i0, int ub, int Mp, int M, int n, uint32_t lanes_alive,
                                         uint32_t mine, int lane) {
    const uint32_t sv = src_t[32 * step + lane];
    while (lanes_alive) {
        const int sl = __ffs((int)lanes_alive) - 1;
        lanes_alive &= lanes_alive - 1;
        uint32_t f = __shfl_sync(0xffffffffu, mine, sl);
        while (f) {
            const int c = __ffs((int)f) - 1;
            f &= f - 1;
            const int x = sl * MATCH_C + c;
            // cell (row 32 step + lane, diagonal x) <-> dst_t[255 + 32 step + lane - x]
            const bool ok = __popc(sv ^ dst_t[255 + 32 * step + lane - x]) <= (int)ectx.threshold;
            if (__all_sync(0xffffffffu, ok)) full_word(ectx, p, i0, ub + x, Mp, M, n, lane);
        }
    }
}

//
// Geometry: diagonals are WRAPPED modulo (m - 1).  Column 0 never matches, so a
// row has m - 1 live columns j = 1 .. m-1; wrapped diagonal u holds the cells
// j - 1 = (i - 1 - u) mod (m - 1).  A 256-row chunk then needs exactly
// ceil((m-1)/256) warp tiles with no empty triangles at either end (the plain
// diagonal tiling of the general kernel computes ~19 % dead cells on the
// BASELINE shapes).  A word that contains the wrap point, row 0, rows past n-1
// or a duplicate diagonal (u >= m-1 in the last band) can look full without
// being a run; those are filtered where a full word is reported (rare), so the
// hot loop has no boundary handling at all.
//
// SAMPLED = true goes one step further: "all 32 cells match" implies "the cells of any
// 4 rows match", so the word is tested 4 rows at a time and abandoned (by the whole
// warp) as soon as no diagonal survives -- usually after the first 4 rows, 1/8 of the
// POPCs.  Survivors of all 8 stages are exactly the full words.  Same runs, bit for bit.
template <bool SAMPLED>
__global__ void __launch_bounds__(MATCH_WARPS * 32, SAMPLED ? 5 : 4) match_fast_kernel(const MatchArgs a) {
    // two tile buffers per warp: the next tile's hashes arrive by bulk copy (TMA) while this one is matched
    __shared__ __align__(128) uint32_t s_src[MATCH_WARPS][2][MATCH_R];
    __shared__ __align__(128) uint32_t s_dst[MATCH_WARPS][2][FAST_DST_TILE];
    // A second copy of the dst window, 16 bytes further into its 128-byte line.  A lane's 16-byte loads
    // are 32 bytes from its neighbour's, so the 8 lanes of one wavefront would cover only 4 of the 8
    // 16-byte bank groups twice (r02_ncu_k3_trail.txt, step 0: 11.6 M conflict wavefronts, the LSU data pipe at
    // 73 %); lanes with bit 2 set read this copy and land on the other 4 groups.
    __shared__ __align__(128) uint32_t s_dst_b[MATCH_WARPS][2][FAST_DST_TILE + 32];
    __shared__ __align__(8) uint64_t s_bar[MATCH_WARPS][2];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int T = (int)a.threshold;
    const EmitCtx ectx = {a.hashes, a.ts, a.problems, a.counters, a.runs, a.run_cap, a.threshold};
    if (not_my_turn(a)) return;
    if (lane == 0) {
        mbar_init(smem_u32(&s_bar[warp][0]), 1);
        mbar_init(smem_u32(&s_bar[warp][1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // Geometry of a tile: diagonals wrapped modulo the PERIOD Mp = (m - 1) rounded up to a multiple of 4;
    // the dst window of a tile then starts at a multiple of 4 elements (r0, ub and 256 are), i.e. 16-byte
    // aligned in the repeated list a.dst2 -- one bulk copy.  Columns M .. Mp-1 are padding: a word that
    // touches them is rejected where full words are reported, like the wrap point.
    //
    // What a warp claims from the global cursor is a UNIT: one row chunk x up to a.band_group consecutive
    // bands (tiles) of one problem.  The atomic, the LUT walk, the descriptor loads and the divisions --
    // a chain of dependent L2 round trips, a third of the kernel's stall samples when paid per tile
    // (profiles/r02_ncu_k3_trail.txt, step 1) -- are paid once per unit; the src rows are fetched once per unit;
    // from band to band only the dst window moves, by 256 columns.
    struct Unit {
        uint32_t p;
        int n, M, Mp, r0, ub0, nb, e0, estep;
    };
    auto claim = [&](Unit &t) -> bool {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(&a.counters[0], 1u);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= a.n_tiles) return false;
        uint32_t p = a.lut[unit >> LUT_SHIFT];
        while (unit >= a.tile_end[p]) p++;
        const MatchProblem &pr = a.problems[p];
        const uint32_t local = unit - pr.tile_begin;
        const uint32_t chunk = local / pr.n_groups;
        const uint32_t band0 = (local - chunk * pr.n_groups) * a.band_group;
        t.p = p;
        t.n = (int)pr.n;
        t.M = (int)pr.m - 1;
        t.Mp = (int)pr.period;
        t.r0 = (int)chunk * MATCH_R;
        t.ub0 = (int)band0 * MATCH_DW;
        t.nb = (int)min(a.band_group, pr.n_bands - band0);
        // dst_t[s] = column 1 + ((e + s) mod Mp) of dst,  cell (row r0+kk, diagonal ub+x) <-> s = 255 + kk - x
        int e0 = (t.r0 - t.ub0 - MATCH_DW) % t.Mp;
        if (e0 < 0) e0 += t.Mp;
        t.e0 = e0;
        t.estep = MATCH_DW % t.Mp;     // e of the next band = e - 256 (mod Mp)
        return true;
    };
    // bulk copies of one tile into dst buffer b: its dst window (twice, see s_dst_b) and, at the first
    // band of a unit, the unit's src rows into src buffer sb
    auto fetch = [&](const Unit &t, int e, int b, int sb, bool with_src) {
        const MatchProblem &pr = a.problems[t.p];
        const uint32_t *hs = a.hashes + pr.src_off + t.r0;
        const int rows = min(MATCH_R, t.n - t.r0), sn = with_src ? (rows & ~3) : 0;   // rows past n are never used unmasked
        fence_proxy_async();
        if (lane == 0) {
            const uint32_t bar = smem_u32(&s_bar[warp][b]);
            mbar_expect_tx(bar, (uint32_t)sn * 4u + 2u * FAST_DST_TILE * 4u);
            if (sn) bulk_g2s(smem_u32(&s_src[warp][sb][0]), hs, (uint32_t)sn * 4u, bar);
            bulk_g2s(smem_u32(&s_dst[warp][b][0]), a.dst2 + pr.dst2_off + e, FAST_DST_TILE * 4u, bar);
            bulk_g2s(smem_u32(&s_dst_b[warp][b][4]), a.dst2 + pr.dst2_off + e, FAST_DST_TILE * 4u, bar);
        }
        if (with_src && lane < 3 && sn + lane < rows) s_src[warp][sb][sn + lane] = hs[sn + lane];   // a list need not end on 16 bytes
    };

    Unit cur, nxt;
    bool have = claim(cur);
    int band = 0, sb = 0, e_cur = 0;   // band of the unit being matched; its src buffer; e of the last window fetched
    if (have) {
        e_cur = cur.e0;
        fetch(cur, e_cur, 0, 0, true);
    }
    uint32_t phase[2] = {0u, 0u};
    for (int buf = 0; have; buf ^= 1) {
        // the tile after this one: the unit's next band, or the first band of the next unit
        const bool same_unit = band + 1 < cur.nb;
        bool have_next = true;
        if (same_unit) {
            e_cur -= cur.estep;
            if (e_cur < 0) e_cur += cur.Mp;
            fetch(cur, e_cur, buf ^ 1, sb, false);   // that buffer's last readers finished before the __syncwarp below
        } else {
            have_next = claim(nxt);
            if (have_next) {
                e_cur = nxt.e0;
                fetch(nxt, e_cur, buf ^ 1, sb ^ 1, true);
            }
        }
        mbar_wait(smem_u32(&s_bar[warp][buf]), phase[buf]);
        phase[buf] ^= 1;
        __syncwarp();
        const uint32_t *src_t = s_src[warp][sb];
        const uint32_t *dst_t = s_dst[warp][buf];
        const uint32_t *dst_l = (lane & 4) ? &s_dst_b[warp][buf][4] : dst_t;   // this lane's copy (bank groups)
        const uint32_t p = cur.p;
        const int n = cur.n, M = cur.M, Mp = cur.Mp, r0 = cur.r0, ub = cur.ub0 + band * MATCH_DW;

        const int n_steps = min(MATCH_R / 32, (n - r0 + 31) / 32);
        int pf_score = 3;   // two words in a row must survive the first look before it is switched off
        const uint32_t *dl0 = dst_l + (248 - MATCH_C * lane);
        for (int step2 = 0; step2 < n_steps; step2 += 2) {
            // words are taken two at a time: two independent chains through the first look
            uint32_t todo = step2 + 1 < n_steps ? 3u : 1u;
            if (SAMPLED && pf_score > 0) {
                // First look, 2 POPCs for 3 cells of a diagonal (rows 0, 12, 24 of the word): if all three
                // match, their distances add up to at most 3 T.  The three XOR words go through one
                // carry-save adder -- ones = x0^x1^x2, twos = maj(x0,x1,x2), 4 LOP3 on the ALU pipe per diagonal --
                // and sum = popc(ones) + 2 popc(twos).  Unrelated hashes (sum ~ 48 +- 4.9 against 3 T = 30)
                // leave here at half the POPCs and 6 of the 10 window loads of the first 4-row stage.
                // Exact: a word that fails cannot be full.  pf_score turns the look off while it is not
                // paying (stationary audio, silence) and probes again at every tile.  (The second word of
                // an odd tail is computed on whatever the buffers hold and ignored.)
                const int lo0 = first_look(dl0 + 32 * step2, src_t + 32 * step2);
                const int lo1 = first_look(dl0 + 32 * step2 + 32, src_t + 32 * step2 + 32);
                // (handing the few surviving diagonals straight to coop_verify instead of the staged test was
                // measured slower: 4.82 -> 5.51 ms on configs[3] -- a serial chain per survivor, 127 registers)
                uint32_t surv = (__ballot_sync(0xffffffffu, lo0 <= 3 * T) ? 1u : 0u) |
                                (__ballot_sync(0xffffffffu, lo1 <= 3 * T) ? 2u : 0u);
                surv &= todo;
                pf_score = min(pf_score + __popc(todo ^ surv) - 2 * __popc(surv), 8);
                todo = surv;
            }
#pragma unroll 1
            for (int w = 0; w < 2; w++) {
                if (((todo >> w) & 1u) == 0) continue;
                const int step = step2 + w;
                const uint4 *dptr = reinterpret_cast<const uint4 *>(dl0 + 32 * step);
                int mx[MATCH_C];
#pragma unroll
                for (int c = 0; c < MATCH_C; c++) mx[c] = 0;
                if (SAMPLED) {
                    // Adaptive depth.  Stage s tests rows o, o+8, o+16, o+24 with o = bitrev3(s)
                    // (0,4,2,6,1,5,3,7): 4 more rows of every word.  The warp leaves as soon as no
                    // lane has a diagonal that still matches everywhere it was tested: one stage
                    // for unrelated hashes (p^4 ~ 4e-7), a few for correlated audio; a diagonal
                    // that survives all 8 stages has 32 matching cells.
                    uint32_t dw[40];
#pragma unroll
                    for (int q = 0; q < 10; q++) {
                        const uint4 v = dptr[q];
                        dw[4 * q + 0] = v.x;
                        dw[4 * q + 1] = v.y;
                        dw[4 * q + 2] = v.z;
                        dw[4 * q + 3] = v.w;
                    }
                    bool alive = true;
#pragma unroll
                    for (int stage = 0; stage < 8; stage++) {
                        if (alive) {
                            const int o = ((stage & 1) << 2) | (stage & 2) | ((stage & 4) >> 2);
                            uint32_t sr[4];
#pragma unroll
                            for (int r = 0; r < 4; r++) sr[r] = src_t[32 * step + o + 8 * r];
#pragma unroll
                            for (int r = 0; r < 4; r += 2) {
#pragma unroll
                                for (int c = 0; c < MATCH_C; c++) {
                                    const int p0 = __popc(sr[r] ^ dw[o + 8 * r + (MATCH_C - 1) - c]);
                                    const int p1 = __popc(sr[r + 1] ^ dw[o + 8 * (r + 1) + (MATCH_C - 1) - c]);
                                    mx[c] = max(mx[c], max(p0, p1));
                                }
                            }
                            if (stage < 7) {
                                int b2 = mx[0];
#pragma unroll
                                for (int c = 1; c < MATCH_C; c++) b2 = min(b2, mx[c]);
                                const uint32_t lanes_alive = __ballot_sync(0xffffffffu, b2 <= T);
                                alive = lanes_alive != 0;
                                if (stage < 3 && alive && __popc(lanes_alive) <= COOP_MAX_LANES) {
                                    uint32_t mine = 0;
#pragma unroll
                                    for (int c = 0; c < MATCH_C; c++) mine |= (mx[c] <= T ? 1u : 0u) << c;
                                    coop_verify(ectx, p, src_t, dst_t, step, r0 + 32 * step, ub, Mp, M, n, lanes_alive, mine, lane);
                                    alive = false;   // this word is done
                                }
                            }
                        }
                    }
                    if (!alive) continue;   // warp-uniform: nothing survived
                } else {
                    const uint4 *sptr = reinterpret_cast<const uint4 *>(src_t + 32 * step);
                    uint32_t dw[40];
#pragma unroll
                    for (int q = 0; q < 10; q++) {
                        const uint4 v = dptr[q];
                        dw[4 * q + 0] = v.x;
                        dw[4 * q + 1] = v.y;
                        dw[4 * q + 2] = v.z;
                        dw[4 * q + 3] = v.w;
                    }
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const uint4 sv = sptr[q];
                        const uint32_t s4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                        for (int kk = 0; kk < 4; kk += 2) {
                            const int k = 4 * q + kk;
#pragma unroll
                            for (int c = 0; c < MATCH_C; c++) {
                                // cell (row k, diagonal c of this lane): window element k + 7 - c
                                const int p0 = __popc(s4[kk] ^ dw[k + (MATCH_C - 1) - c]);
                                const int p1 = __popc(s4[kk + 1] ^ dw[k + 1 + (MATCH_C - 1) - c]);
                                mx[c] = max(mx[c], max(p0, p1));     // one 3-input max per two cells
                            }
                        }
                    }
                }
                int best = mx[0];
#pragma unroll
                for (int c = 1; c < MATCH_C; c++) best = min(best, mx[c]);
                // ---- rare: some diagonal of some lane has 32 matching cells in a row
                uint32_t pending = __ballot_sync(0xffffffffu, best <= T);
                if (pending == 0) continue;
                uint32_t full = 0;
#pragma unroll
                for (int c = 0; c < MATCH_C; c++) full |= (mx[c] <= T ? 1u : 0u) << c;
                const int i0 = r0 + 32 * step;
                while (pending) {
                    const int src_lane = __ffs((int)pending) - 1;
                    pending &= pending - 1;
                    uint32_t f = __shfl_sync(0xffffffffu, full, src_lane);
                    while (f) {
                        const int c = __ffs((int)f) - 1;
                        f &= f - 1;
                        full_word(ectx, p, i0, ub + src_lane * MATCH_C + c, Mp, M, n, lane);
                    }
                }
            }
        }
        __syncwarp();   // every lane is done with this buffer: the tile after next may land in it
        if (same_unit) {
            band++;
        } else {
            cur = nxt;
            band = 0;
            sb ^= 1;
            have = have_next;
        }
    }
}

// K4: simhash32 of both sides of every emitted run (comparator.rs:149-153,
// 226-229): per-bit majority over hashes[i_end-len ..= i_end], ties -> 0.
// One warp per run; lane b counts bit b.
// 32 x 32 bit-matrix transpose across the warp (one row per lane, five exchange steps, the
// recursive block swap of Hacker's Delight 7-3): afterwards lane l holds, one bit per source
// lane, bit 31 - l of the words the lanes started with.
__device__ __forceinline__ uint32_t warp_bit_transpose(uint32_t x, int lane) {
    uint32_t m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t p = __shfl_xor_sync(0xffffffffu, x, j);
        if ((lane & j) == 0) {
            x ^= (x ^ (p >> j)) & m;
        } else {
            x ^= ((p ^ (x >> j)) & m) << j;
        }
        m ^= m << (j >> 1);
    }
    return x;
}

__global__ void __launch_bounds__(128) simhash_kernel(const uint32_t *__restrict__ hashes,
                                                      const MatchProblem *__restrict__ problems,
                                                      DeviceRun *runs, const unsigned int *__restrict__ counters,
                                                      uint32_t run_cap) {
    // the run count lives on the device: the host has not waited for the match kernel
    const uint32_t n_runs = min(counters[1], run_cap);
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_runs; w += n_warps) {
        const DeviceRun r = runs[w];
        const MatchProblem pr = problems[r.problem];
        const uint32_t count = r.len + 1;
        const uint32_t *hs = hashes + pr.src_off + (r.i_end - r.len);
        const uint32_t *hd = hashes + pr.dst_off + (r.j_end - r.len);
        // Per 32 hashes of each side: one load per lane, a bit transpose (lane b then holds bit b of
        // all 32), one POPC.  The next chunk is loaded before this one is reduced.
        uint32_t ones_s = 0, ones_d = 0;
        uint32_t ms = lane < count ? hs[lane] : 0u, md = lane < count ? hd[lane] : 0u;
        for (uint32_t base = 0; base < count; base += 32) {
            const uint32_t nxt = base + 32 + lane;
            const uint32_t ns = nxt < count ? hs[nxt] : 0u, nd = nxt < count ? hd[nxt] : 0u;
            ones_s += __popc(warp_bit_transpose(ms, lane));
            ones_d += __popc(warp_bit_transpose(md, lane));
            ms = ns;
            md = nd;
        }
        // v[b] = ones - (count - ones) > 0; lane l counted bit 31 - l
        const uint32_t out_s = __brev(__ballot_sync(0xffffffffu, 2u * ones_s > count));
        const uint32_t out_d = __brev(__ballot_sync(0xffffffffu, 2u * ones_d > count));
        if (lane == 0) {
            runs[w].src_simhash = out_s;
            runs[w].dst_simhash = out_d;
        }
    }
}

// lcand of every problem from the per-segment bounds (both tilings of the problem list),
// and the smallest lcand overall, which decides between the fast and the general kernel.
__global__ void problem_finalize_kernel(MatchProblem *__restrict__ fast, MatchProblem *__restrict__ general,
                                        const uint32_t *__restrict__ lmin, uint32_t n_problems,
                                        unsigned int *__restrict__ counters) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_problems) return;
    const MatchProblem *any = fast ? fast : general;
    const uint32_t lc = max(1u, max(lmin[any[p].src_seg], lmin[any[p].dst_seg]));
    if (fast) fast[p].lcand = lc;
    if (general) general[p].lcand = lc;
    atomicMax(&counters[3], ~lc);
}

// Per segment: the smallest run length that can satisfy ts[i] - ts[i-len] >= min
// anywhere in the list (comparator.rs:212-218 only ever compares timestamps of
// one list).  1 when the list is not monotone (then every run is examined and
// the underflow the reference would panic on is reported), "longer than the
// list" when no window qualifies.  One thread per element: binary search for
// the latest start that is still min apart, atomicMin of the distance.
struct SegTab {
    uint32_t dev_offset, len;
};
__global__ void seg_lmin_kernel(const uint64_t *__restrict__ ts, const SegTab *__restrict__ segs,
                                uint64_t min_open_ns, uint64_t min_end_ns, uint32_t *__restrict__ lmin) {
    // blockIdx.y = segment (any placement in memory), blockIdx.x * blockDim.x + threadIdx.x = element
    const uint32_t seg = blockIdx.y;
    const SegTab sg = segs[seg];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sg.len || i == 0) return;
    const uint64_t *t = ts + sg.dev_offset;
    const uint64_t mn = (seg & 1) ? min_end_ns : min_open_ns;
    const uint64_t ti = t[i];
    if (ti < t[i - 1]) {
        atomicMin(&lmin[seg], 1u);
        return;
    }
    if (ti < mn || t[0] > ti - mn) return;          // no start is far enough from i
    // largest s in [0, i-1] with t[s] <= ti - mn  (t[0] qualifies)
    const uint64_t limit = ti - mn;
    uint32_t a = 0, b = i;                          // invariant: t[a] <= limit, b = first index known > limit (or i)
    while (b - a > 1) {
        const uint32_t mid = (a + b) >> 1;
        if (t[mid] <= limit) a = mid; else b = mid;
    }
    const uint32_t len = i - a;
    atomicMin(&lmin[seg], len < 1u ? 1u : len);
}

// The fast kernel's view of a list as a dst: its live columns 1 .. len-1, padded to a period that is a
// multiple of 4 and repeated often enough that any 512-column window of the wrapped list is one span.
struct Dst2Seg {
    uint32_t dev_offset, len, out_off, period, reps;
};
__global__ void dst2_build_kernel(const uint32_t *__restrict__ hashes, const Dst2Seg *__restrict__ segs,
                                  uint32_t *__restrict__ out) {
    const Dst2Seg sg = segs[blockIdx.y];
    const uint32_t total = sg.period * sg.reps;
    for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < total; x += gridDim.x * blockDim.x) {
        const uint32_t c = x % sg.period;
        out[sg.out_off + x] = c + 1 < sg.len ? hashes[sg.dev_offset + 1 + c] : 0u;
    }
}

// ------------------------------------------------------------------- host

// Enqueue the whole match stage on ctx->stream without waiting for anything:
// run-length bounds, problem tables (both tilings), the fast and the general
// kernel (the device decides which of the two does the work), simhash.  Runs
// are left in ctx->d_runs, their count in ctx->d_counters[1].
int match_enqueue(nb200_ctx *ctx, const nb200_hashset *hs_c, const uint32_t (*pairs)[2], uint64_t n_pairs,
                  const nb200_match_params *params, uint32_t pair_base, uint32_t run_cap_min, MatchPending *out,
                  DeviceRun *ext_runs, uint32_t ext_cap) {
    nb200_hashset *hs = const_cast<nb200_hashset *>(hs_c);
    // the reference takes any u16 threshold (comparator.rs:76); above 32 every cell with i, j >= 1
    // matches, exactly as with 32, so the kernels see min(T, 32) (the vote's bias T + T/2 keeps T)
    nb200_match_params clamped = *params;
    clamped.hash_match_threshold = std::min<uint32_t>(params->hash_match_threshold, 32u);
    params = &clamped;
    NB_CUDA(cudaSetDevice(ctx->device));
    // NULL = every (i < j) in the reference's order; the list itself is only built when the plan is
    const bool all_pairs = pairs == nullptr;
    std::vector<uint32_t> default_flat;
    if (all_pairs) n_pairs = (uint64_t)hs->n_videos * (hs->n_videos ? hs->n_videos - 1 : 0) / 2;
    if (2 * (n_pairs + (uint64_t)pair_base) >= 0xffffffffull) return NB200_ERR_TOO_LARGE;
    const bool endings = params->include_endings != 0;
    const bool want_fast = !ctx->force_general_match;
    cudaStream_t s = ctx->stream;
    out->problem_pair.clear();
    out->problem_is_ending.clear();

    PhaseTimer phase_setup(ctx, NB200_PHASE_MATCH_SETUP);
    // A search over the same lists, pairs and parameters as the previous one on this context (a
    // resident job run again and again) finds its descriptor tables still in the device arena: the
    // host neither rebuilds nor uploads them (1.5 ms for 19,900 pairs, a fifth of the whole search).
    uint64_t key = 1469598103934665603ull;
    auto mix = [&](const void *p, size_t bytes) {
        const unsigned char *c = static_cast<const unsigned char *>(p);
        uint64_t h = key;
        size_t k = 0;
        for (; k + 8 <= bytes; k += 8) {
            uint64_t w;
            memcpy(&w, c + k, 8);
            h = (h ^ w) * 1099511628211ull;
        }
        for (; k < bytes; k++) h = (h ^ c[k]) * 1099511628211ull;
        key = h;
    };
    {
        const uint64_t scalars[8] = {hs->n_videos, n_pairs, pair_base, params->hash_match_threshold,
                                     params->min_opening_ns, params->min_ending_ns,
                                     (uint64_t)endings | ((uint64_t)want_fast << 1) | ((uint64_t)ctx->match_band_group << 8),
                                     all_pairs ? 2u : 1u};
        mix(scalars, sizeof(scalars));
        mix(hs->dev_offset.data(), hs->dev_offset.size() * sizeof(uint32_t));
        mix(hs->seg_len.data(), hs->seg_len.size() * sizeof(uint32_t));
        if (!all_pairs) mix(pairs, (size_t)n_pairs * 8);
    }
    MatchPlanCache &pc = ctx->match_plan;
    const bool hit = pc.valid && pc.key == key;
    const uint32_t nseg = 2 * hs->n_videos;
    if (!hit) {
        pc.valid = false;
        if (all_pairs) {
            default_pairs(hs->n_videos, default_flat);
            pairs = reinterpret_cast<const uint32_t(*)[2]>(default_flat.data());
        }
        // every list as a dst of the fast kernel: period and place of its repeated live columns
        std::vector<Dst2Seg> d2(nseg);
        uint64_t d2_total = 0;
        uint32_t d2_max = 1;
        for (uint32_t k = 0; k < nseg; k++) {
            const uint32_t len = hs->seg_len[k];
            const uint32_t period = len >= 2 ? ((len - 1 + 3) & ~3u) : 4u;
            const uint32_t reps = std::max<uint32_t>(2u, (FAST_DST_TILE + period - 1) / period + 1);
            d2[k] = {hs->dev_offset[k], len, (uint32_t)d2_total, period, reps};
            d2_total += (uint64_t)period * reps;
            d2_max = std::max(d2_max, period * reps);
            if (d2_total > 0xfffffff0ull) return NB200_ERR_TOO_LARGE;
        }
        // problems, tiled for the fast kernel (wrapped diagonals) and for the general one
        std::vector<MatchProblem> prob_f, prob_g;
        prob_f.reserve(n_pairs * (endings ? 2 : 1));
        prob_g.reserve(n_pairs * (endings ? 2 : 1));
        pc.problem_pair.clear();
        pc.problem_is_ending.clear();
        uint64_t tiles_f = 0, tiles_g = 0, cells = 0;
        for (uint64_t k = 0; k < n_pairs; k++) {
            const uint32_t a = pairs[k][0], b = pairs[k][1];
            if (a >= hs->n_videos || b >= hs->n_videos) return NB200_ERR_INVALID_ARGUMENT;
            for (int e = 0; e < (endings ? 2 : 1); e++) {
                const uint32_t sa = 2 * a + e, sb = 2 * b + e;
                const uint32_t n = hs->seg_len[sa], m = hs->seg_len[sb];
                if (e == 1 && (n == 0 || m == 0)) return NB200_ERR_NO_ENDING;   // comparator.rs:271-273
                cells += (uint64_t)n * m;
                if (n < 2 || m < 2) continue;   // rows/cols 0 never match: no run possible (:165-167,179)
                MatchProblem pr;
                pr.n_groups = 0;
                pr.src_off = hs->dev_offset[sa];
                pr.n = n;
                pr.dst_off = hs->dev_offset[sb];
                pr.m = m;
                pr.min_ns = e ? params->min_ending_ns : params->min_opening_ns;
                pr.lcand = 1;   // problem_finalize_kernel
                pr.group = (uint32_t)(2 * (k + pair_base) + (uint64_t)e);
                pr.src_seg = sa;
                pr.dst_seg = sb;
                pr.dst2_off = d2[sb].out_off;
                pr.period = d2[sb].period;
                const uint64_t chunks = (n + MATCH_R - 1) / MATCH_R;
                pr.n_bands = (uint32_t)(((uint64_t)pr.period + MATCH_DW - 1) / MATCH_DW);
                pr.tile_begin = (uint32_t)tiles_f;
                tiles_f += chunks * pr.n_bands;
                prob_f.push_back(pr);
                // diagonals per chunk: (m-1) + (R-1) + up to 3 of alignment slack
                pr.n_bands = (uint32_t)(((uint64_t)m + MATCH_R + 1 + MATCH_DW - 1) / MATCH_DW);
                pr.tile_begin = (uint32_t)tiles_g;
                tiles_g += chunks * pr.n_bands;
                prob_g.push_back(pr);
                if (tiles_g >= 0xfffffff0ull) return NB200_ERR_TOO_LARGE;
                pc.problem_pair.push_back((uint32_t)k);
                pc.problem_is_ending.push_back((uint8_t)e);
            }
        }
        // fast kernel: bands are claimed in groups (units) so that the per-claim work is amortised, but
        // never so large that a persistent warp gets fewer than ~32 units (the last unit of the slowest
        // warp is the kernel's tail: 8 % of a rank's 0.6 ms at N = 8 with 15 units per warp)
        {
            const uint64_t warps = (uint64_t)ctx->sm_count * 5 * MATCH_WARPS;
            const uint32_t G = ctx->match_band_group ? ctx->match_band_group
                                                     : (uint32_t)std::min<uint64_t>(16, std::max<uint64_t>(1, tiles_f / (warps * 32)));
            uint64_t units = 0;
            for (MatchProblem &pr : prob_f) {
                const uint64_t chunks = (pr.n + MATCH_R - 1) / MATCH_R;
                pr.n_groups = (pr.n_bands + G - 1) / G;
                pr.tile_begin = (uint32_t)units;
                units += chunks * pr.n_groups;
            }
            tiles_f = units;
            pc.band_group = G;
        }
        pc.cells = cells;
        pc.n_problems = (uint32_t)prob_f.size();
        pc.tiles_f = tiles_f;
        pc.tiles_g = tiles_g;
        pc.d2_total = d2_total;
        pc.d2_max = d2_max;
        pc.max_len = 1;
        for (uint32_t k = 0; k < nseg; k++) pc.max_len = std::max(pc.max_len, hs->seg_len[k]);

        if (pc.n_problems > 0) {
            auto tile_tables = [](const std::vector<MatchProblem> &pb, uint32_t n_tiles, std::vector<uint32_t> &tile_end,
                                  std::vector<uint32_t> &lut) {
                const uint32_t np = (uint32_t)pb.size();
                tile_end.resize(np);
                for (uint32_t p = 0; p < np; p++) tile_end[p] = (p + 1 < np) ? pb[p + 1].tile_begin : n_tiles;
                lut.resize((n_tiles >> LUT_SHIFT) + 1);
                uint32_t p = 0;
                for (uint32_t k = 0; k < lut.size(); k++) {
                    const uint32_t t = k << LUT_SHIFT;
                    while (p + 1 < np && tile_end[p] <= t) p++;
                    lut[k] = p;
                }
            };
            std::vector<uint32_t> tend_f, lut_f, tend_g, lut_g;
            if (want_fast) tile_tables(prob_f, (uint32_t)tiles_f, tend_f, lut_f);
            tile_tables(prob_g, (uint32_t)tiles_g, tend_g, lut_g);

            // Everything this call uploads lives in ONE device arena mirrored by one pinned staging
            // area: one H2D copy, issued on the copy stream so that it runs underneath whatever is
            // still executing on the compute stream (K1/K2 of a fused analyze + search); the compute
            // stream waits for it just before the first kernel that reads it.
            size_t total = 0;
            auto carve = [&](size_t bytes) {
                const size_t off = total;
                total += (bytes + 255) & ~size_t(255);
                return off;
            };
            const size_t b_prob = (size_t)pc.n_problems * sizeof(MatchProblem), b_tend = (size_t)pc.n_problems * sizeof(uint32_t);
            pc.o_tab = carve((size_t)nseg * sizeof(SegTab));
            pc.o_lmin = carve((size_t)nseg * sizeof(uint32_t));
            pc.o_d2 = carve((size_t)nseg * sizeof(Dst2Seg));
            pc.o_prob_f = carve(b_prob);
            pc.o_tend_f = carve(b_tend);
            pc.o_lut_f = carve(lut_f.size() * sizeof(uint32_t));
            pc.o_prob_g = carve(b_prob);
            pc.o_tend_g = carve(b_tend);
            pc.o_lut_g = carve(lut_g.size() * sizeof(uint32_t));
            // The arena and the staging area are rewritten below, from another stream: the previous match
            // stage enqueued on this context must be through with them.  Normally it is (every search
            // ends with a synchronisation); two nb200_match_export calls back to back are the exception.
            NB_CUDA(cudaEventSynchronize(ctx->ev_chunk[10]));
            if (ctx->h_stage.reserve(total) != NB200_OK || ctx->d_problems.reserve(total) != NB200_OK) return NB200_ERR_CUDA;
            char *st = ctx->h_stage.as<char>();
            // run-length lower bounds per segment, computed where the timestamps live
            {
                SegTab *tab = reinterpret_cast<SegTab *>(st + pc.o_tab);
                uint32_t *init = reinterpret_cast<uint32_t *>(st + pc.o_lmin);
                for (uint32_t k = 0; k < nseg; k++) {
                    tab[k] = {hs->dev_offset[k], hs->seg_len[k]};
                    const uint64_t mn = (k & 1) ? params->min_ending_ns : params->min_opening_ns;
                    // n < 2 or min == 0: every run is a candidate; else start from "no window qualifies"
                    init[k] = (hs->seg_len[k] < 2 || mn == 0) ? 1u : std::min<uint32_t>(hs->seg_len[k], 0x7fffffffu);
                }
            }
            memcpy(st + pc.o_d2, d2.data(), (size_t)nseg * sizeof(Dst2Seg));
            if (want_fast) {
                memcpy(st + pc.o_prob_f, prob_f.data(), b_prob);
                memcpy(st + pc.o_tend_f, tend_f.data(), b_tend);
                memcpy(st + pc.o_lut_f, lut_f.data(), lut_f.size() * sizeof(uint32_t));
            }
            memcpy(st + pc.o_prob_g, prob_g.data(), b_prob);
            memcpy(st + pc.o_tend_g, tend_g.data(), b_tend);
            memcpy(st + pc.o_lut_g, lut_g.data(), lut_g.size() * sizeof(uint32_t));
            NB_CUDA(cudaMemcpyAsync(ctx->d_problems.p, st, total, cudaMemcpyHostToDevice, ctx->copy_stream));
            NB_CUDA(cudaEventRecord(ctx->ev_chunk[8], ctx->copy_stream));
            NB_CUDA(cudaStreamWaitEvent(s, ctx->ev_chunk[8], 0));
        }
        pc.key = key;
        pc.valid = true;
    }
    if (!ext_runs) {   // only the run-set path reads these
        out->problem_pair = pc.problem_pair;
        out->problem_is_ending = pc.problem_is_ending;
    }
    out->n_cells = pc.cells;
    const uint32_t n_problems = pc.n_problems;
    const uint64_t tiles_f = pc.tiles_f, tiles_g = pc.tiles_g, d2_total = pc.d2_total;
    const uint32_t d2_max = pc.d2_max, max_len = pc.max_len;
    out->n_problems = n_problems;
    uint32_t run_cap = std::max<uint32_t>(4096u, (uint32_t)std::min<uint64_t>(n_problems * 8ull, 1u << 22));
    run_cap = std::max(run_cap, run_cap_min);
    if (ext_runs) {
        run_cap = ext_cap;   // the caller's block (an exchange send buffer): runs are emitted straight into it
    } else {
        if (ctx->d_runs.bytes / sizeof(DeviceRun) > run_cap)
            run_cap = (uint32_t)std::min<uint64_t>(ctx->d_runs.bytes / sizeof(DeviceRun), 0xffffffffull);
        if (ctx->d_runs.reserve((size_t)run_cap * sizeof(DeviceRun)) != NB200_OK) return NB200_ERR_CUDA;
    }
    DeviceRun *const runs_out = ext_runs ? ext_runs : ctx->d_runs.as<DeviceRun>();
    out->run_cap = run_cap;
    if (ctx->d_counters.reserve(64) != NB200_OK) return NB200_ERR_CUDA;
    // the same immutable season as last time: its run-length bounds, the problems' lcand, the smallest
    // lcand (counters[3]) and the repeated dst lists are still on the device
    const bool same_data = hit && hs->immutable && pc.data_uid == hs->uid;
    NB_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, same_data ? 12 : 64, s));
    ctx->ev_valid[2] = ctx->ev_valid[3] = false;
    if (n_problems == 0) return NB200_OK;
    if (want_fast && ctx->d_dst2.reserve((size_t)(d2_total + 16) * sizeof(uint32_t)) != NB200_OK) return NB200_ERR_CUDA;
    char *arena = ctx->d_problems.as<char>();
    SegTab *const d_segtab = reinterpret_cast<SegTab *>(arena + pc.o_tab);
    uint32_t *const d_lmin = reinterpret_cast<uint32_t *>(arena + pc.o_lmin);
    MatchProblem *const d_prob_f = reinterpret_cast<MatchProblem *>(arena + pc.o_prob_f);
    MatchProblem *const d_prob_g = reinterpret_cast<MatchProblem *>(arena + pc.o_prob_g);
    const size_t o_d2 = pc.o_d2, o_tend_f = pc.o_tend_f, o_lut_f = pc.o_lut_f, o_tend_g = pc.o_tend_g, o_lut_g = pc.o_lut_g;

    if (!same_data) {
        seg_lmin_kernel<<<dim3((max_len + 255) / 256, nseg), 256, 0, s>>>(
            hs->d_ts.as<uint64_t>(), d_segtab, params->min_opening_ns, params->min_ending_ns, d_lmin);
        NB_CUDA(cudaGetLastError());
        ctx->n_launches++;
        problem_finalize_kernel<<<(n_problems + 255) / 256, 256, 0, s>>>(want_fast ? d_prob_f : nullptr, d_prob_g, d_lmin,
                                                                          n_problems, ctx->d_counters.as<unsigned int>());
        NB_CUDA(cudaGetLastError());
        ctx->n_launches++;
    }

    MatchArgs args;
    args.band_group = 1;
    args.hashes = hs->d_hashes.as<uint32_t>();
    args.ts = hs->d_ts.as<uint64_t>();
    args.threshold = params->hash_match_threshold;
    args.run_cap = run_cap;
    args.counters = ctx->d_counters.as<unsigned int>();
    args.runs = runs_out;
    auto grid_for = [&](uint64_t n_tiles, uint32_t ctas_per_sm = 4) {
        return std::min<uint32_t>((uint32_t)ctx->sm_count * ctas_per_sm, (uint32_t)((n_tiles + MATCH_WARPS - 1) / MATCH_WARPS));
    };
    if (want_fast && !same_data) {
        // the repeated dst lists (a few MB at most): rebuilt with every match unless the season is the
        // same immutable one -- a view may be over buffers that K2 or an all-gather has just rewritten
        dst2_build_kernel<<<dim3(std::min<uint32_t>((d2_max + 255) / 256, 64u), nseg), 256, 0, s>>>(
            hs->d_hashes.as<uint32_t>(), reinterpret_cast<const Dst2Seg *>(arena + o_d2), ctx->d_dst2.as<uint32_t>());
        NB_CUDA(cudaGetLastError());
        ctx->n_launches++;
    }
    pc.data_uid = hs->immutable ? hs->uid : 0;
    args.dst2 = ctx->d_dst2.as<uint32_t>();
    NB_CUDA(cudaEventRecord(ctx->ev[4], s));
    if (want_fast) {
        args.problems = d_prob_f;
        args.tile_end = reinterpret_cast<const uint32_t *>(arena + o_tend_f);
        args.lut = reinterpret_cast<const uint32_t *>(arena + o_lut_f);
        args.n_tiles = (uint32_t)tiles_f;
        args.band_group = pc.band_group;
        args.mode = 1;
        if (ctx->match_dense)
            match_fast_kernel<false><<<grid_for(tiles_f), MATCH_WARPS * 32, 0, s>>>(args);
        else
            match_fast_kernel<true><<<grid_for(tiles_f, 5), MATCH_WARPS * 32, 0, s>>>(args);   // 96 registers, 41 KB: 5 CTAs per SM
        NB_CUDA(cudaGetLastError());
        ctx->n_launches++;
    }
    args.problems = d_prob_g;
    args.tile_end = reinterpret_cast<const uint32_t *>(arena + o_tend_g);
    args.lut = reinterpret_cast<const uint32_t *>(arena + o_lut_g);
    args.n_tiles = (uint32_t)tiles_g;
    args.mode = want_fast ? 2 : 0;
    match_kernel<<<grid_for(tiles_g), MATCH_WARPS * 32, 0, s>>>(args);
    NB_CUDA(cudaGetLastError());
    ctx->n_launches++;
    NB_CUDA(cudaEventRecord(ctx->ev[5], s));
    ctx->ev_valid[2] = true;

    NB_CUDA(cudaEventRecord(ctx->ev[6], s));
    simhash_kernel<<<(uint32_t)ctx->sm_count * 2, 128, 0, s>>>(hs->d_hashes.as<uint32_t>(),
                                                             d_prob_g, runs_out,
                                                             ctx->d_counters.as<unsigned int>(), run_cap);
    NB_CUDA(cudaGetLastError());
    NB_CUDA(cudaEventRecord(ctx->ev[7], s));
    NB_CUDA(cudaEventRecord(ctx->ev_chunk[10], s));
    ctx->ev_valid[3] = true;
    ctx->n_launches++;
    return NB200_OK;
}

// Reads the match counters back (one synchronisation).  NB200_OK with *n_runs <= run_cap
// when every run was stored; *n_runs > run_cap tells the caller to enqueue again with room.
int match_collect(nb200_ctx *ctx, const MatchPending &mp, uint64_t *n_runs) {
    cudaStream_t s = ctx->stream;
    NB_CUDA(cudaMemcpyAsync(ctx->h_small.p, ctx->d_counters.p, 16, cudaMemcpyDeviceToHost, s));
    NB_CUDA(cudaStreamSynchronize(s));
    unsigned int h_counters[4];
    memcpy(h_counters, ctx->h_small.p, 16);
    if (h_counters[2] & 1u) return NB200_ERR_DURATION_UNDERFLOW;
    *n_runs = h_counters[1];
    (void)mp;
    return NB200_OK;
}

int match_run_impl(nb200_ctx *ctx, const nb200_hashset *hs, const uint32_t (*pairs)[2],
                   uint64_t n_pairs, const nb200_match_params *params, nb200_runset **out) {
    MatchPending mp;
    uint64_t n_runs = 0;
    uint32_t cap_min = 0;
    for (int attempt = 0;; attempt++) {
        NB_TRY(match_enqueue(ctx, hs, pairs, n_pairs, params, 0, cap_min, &mp));
        PhaseTimer pt(ctx, NB200_PHASE_MATCH);
        NB_TRY(match_collect(ctx, mp, &n_runs));
        if (n_runs <= mp.run_cap) break;
        if (attempt == 1 || n_runs > 0xffffffffull) return NB200_ERR_TOO_LARGE;
        cap_min = (uint32_t)n_runs;   // exact count is known now: rerun with room for all of them
    }
    auto rs = new nb200_runset();
    rs->n_cells = mp.n_cells;
    rs->n_runs = n_runs;
    rs->problem_pair = std::move(mp.problem_pair);
    rs->problem_is_ending = std::move(mp.problem_is_ending);
    if (n_runs > 0) {
        PhaseTimer pt(ctx, NB200_PHASE_SIMHASH);
        if (rs->d_runs.alloc((size_t)n_runs * sizeof(DeviceRun), ctx->pool) != NB200_OK) {
            delete rs;
            return NB200_ERR_CUDA;
        }
        cudaError_t e = cudaMemcpyAsync(rs->d_runs.p, ctx->d_runs.p, (size_t)n_runs * sizeof(DeviceRun),
                                        cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) {
            set_last_error("cudaMemcpyAsync(runs)", e);
            delete rs;
            return NB200_ERR_CUDA;
        }
    }
    *out = rs;
    return NB200_OK;
}

}  // namespace nb200
