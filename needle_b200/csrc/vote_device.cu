// vote_device.cu -- the tail of Comparator::run_with_frame_hashes on the device
// (SURVEY 8f3): what vote.cpp does on the host, for runs that are already in HBM,
// so that a search downloads only the per-video intervals.
//
//   * every table's runs pushed into a BinaryHeap in (i desc, j desc) order and
//     read back as the heap's internal array  (needle/src/audio/comparator.rs:191-192,
//     :231-249; Ord derived over the fields of ComparatorHeapEntry, :20-35)
//   * info_map: per video the non-empty pair results in pair order (:580-588)
//   * find_best_match: candidates (openings then endings of each pair result),
//     all-pairs signature clustering with popcount < T + T/2, score
//     -(0.3 |cluster| + 0.7 secs_f32), smallest (score, candidate index) (:405-515)
//
// Layout: runs are grouped by table g = 2 * pair + is_ending with a counting sort
// (count -> exclusive scan -> scatter), each group is ordered and heap-replayed by
// one thread (tables hold a handful of runs), then one CTA per video enumerates
// its candidates through a CSR of the pair list and reduces to the winner.  The
// candidate index of the reference is monotone in (g, heap position) = the flat
// position in the grouped array, which is what breaks score ties here.
//
// Bounds: a table with more than VOTE_MAX_GROUP runs or a video with more than
// VOTE_MAX_CAND candidates raises VOTE_FLAG_FALLBACK and the caller votes on the
// host instead (same results; vote.cpp).  Float arithmetic uses explicit
// round-to-nearest intrinsics so that no FMA contraction can differ from Rust's f32.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstring>
#include <vector>

#include "common.h"

namespace nb200 {

constexpr int VOTE_MAX_GROUP = 64;
constexpr int VOTE_MAX_CAND = 4096;
constexpr int VOTE_THREADS = 256;

struct VoteArgs {
    const DeviceRun *runs;
    const unsigned int *n_runs_ptr;
    uint32_t run_cap;
    uint32_t n_groups;          // 2 * n_pairs
    uint32_t *gcount;           // [n_groups + 1]; the last element stays 0
    uint32_t *gfill;            // [n_groups]
    uint32_t *gbegin;           // [n_groups + 1] exclusive scan of gcount
    uint32_t *order;            // [run_cap] run index at each flat heap position
    const uint32_t *vid_begin;  // [n_videos + 1] CSR over the pair list
    const uint32_t *vid_pairs;  // 2 * pair + is_source, pair order within a video
    const uint64_t *hash_duration;
    nb200_search_result *results;
    uint32_t *flags;
    uint32_t n_videos;
    uint32_t bias;              // T + T / 2
    uint32_t include_endings;
    uint64_t time_padding_ns;
};

__device__ __forceinline__ uint32_t vote_n_runs(const VoteArgs &a) { return min(*a.n_runs_ptr, a.run_cap); }

__global__ void vote_count_kernel(const VoteArgs a) {
    const uint32_t n = vote_n_runs(a);
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint32_t g = a.runs[r].group;
        if (g >= a.n_groups) {
            atomicOr(a.flags, VOTE_FLAG_FALLBACK);   // the host vote reports the bad pair index
            continue;
        }
        atomicAdd(&a.gcount[g], 1u);
    }
}

__global__ void vote_scatter_kernel(const VoteArgs a) {
    const uint32_t n = vote_n_runs(a);
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint32_t g = a.runs[r].group;
        if (g >= a.n_groups) continue;
        a.order[a.gbegin[g] + atomicAdd(&a.gfill[g], 1u)] = r;
    }
}

// derived Ord of ComparatorHeapEntry, restricted to the fields that can differ inside one table
__device__ __forceinline__ bool entry_greater_dev(const DeviceRun &x, const DeviceRun &y) {
    if (x.len != y.len) return x.len > y.len;
    if (x.src_start_ns != y.src_start_ns) return x.src_start_ns > y.src_start_ns;
    if (x.src_end_ns != y.src_end_ns) return x.src_end_ns > y.src_end_ns;
    if (x.dst_start_ns != y.dst_start_ns) return x.dst_start_ns > y.dst_start_ns;
    if (x.dst_end_ns != y.dst_end_ns) return x.dst_end_ns > y.dst_end_ns;
    if (x.src_simhash != y.src_simhash) return x.src_simhash > y.src_simhash;
    return x.dst_simhash > y.dst_simhash;
}

// One thread per table: push order (i desc, j desc), then BinaryHeap::push (append + sift up
// while greater than the parent) for each run; the heap array replaces the group's slice.
__global__ void vote_heap_kernel(const VoteArgs a) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n_groups) return;
    const uint32_t b = a.gbegin[g], k = a.gbegin[g + 1] - b;
    if (k < 2) return;
    if (k > (uint32_t)VOTE_MAX_GROUP) {
        atomicOr(a.flags, VOTE_FLAG_FALLBACK);
        return;
    }
    uint32_t idx[VOTE_MAX_GROUP];
    for (uint32_t q = 0; q < k; q++) {   // insertion sort by (i_end desc, j_end desc)
        const uint32_t r = a.order[b + q];
        const uint32_t ri = a.runs[r].i_end, rj = a.runs[r].j_end;
        uint32_t pos = q;
        while (pos > 0) {
            const DeviceRun &o = a.runs[idx[pos - 1]];
            if (o.i_end > ri || (o.i_end == ri && o.j_end > rj)) break;
            idx[pos] = idx[pos - 1];
            pos--;
        }
        idx[pos] = r;
    }
    uint32_t heap[VOTE_MAX_GROUP];
    for (uint32_t q = 0; q < k; q++) {
        const uint32_t r = idx[q];
        uint32_t pos = q;
        while (pos > 0) {
            const uint32_t parent = (pos - 1) / 2;
            if (!entry_greater_dev(a.runs[r], a.runs[heap[parent]])) break;
            heap[pos] = heap[parent];
            pos = parent;
        }
        heap[pos] = r;
    }
    for (uint32_t q = 0; q < k; q++) a.order[b + q] = heap[q];
}

// Rust Duration::as_secs_f32: (secs as f32) + (nanos as f32) / 1e9
__device__ __forceinline__ float as_secs_f32_dev(uint64_t ns) {
    const float s = __ull2float_rn(ns / 1000000000ull);
    const float f = __fdiv_rn(__uint2float_rn((uint32_t)(ns % 1000000000ull)), 1000000000.0f);
    return __fadd_rn(s, f);
}

struct Best {
    float score;
    uint32_t key;   // 2 * flat position + (0 source, 1 destination): the candidate order
    uint32_t have;
};
__device__ __forceinline__ void best_merge(Best &x, const Best &y) {
    if (!y.have) return;
    if (!x.have || y.score < x.score || (y.score == x.score && y.key < x.key)) x = y;
}

__global__ void __launch_bounds__(VOTE_THREADS) vote_video_kernel(const VoteArgs a) {
    __shared__ uint32_t s_hash[VOTE_MAX_CAND];
    __shared__ uint32_t s_key[VOTE_MAX_CAND];
    __shared__ uint32_t s_n;
    __shared__ Best s_best[2][VOTE_THREADS];

    const uint32_t v = blockIdx.x;
    const int tid = threadIdx.x;
    if (tid == 0) s_n = 0;
    __syncthreads();
    // candidates: every heap entry of every pair result this video takes part in
    for (uint32_t m = a.vid_begin[v] + tid; m < a.vid_begin[v + 1]; m += VOTE_THREADS) {
        const uint32_t kp = a.vid_pairs[m];
        const uint32_t g0 = 2 * (kp >> 1), is_source = kp & 1u;
        const uint32_t x0 = a.gbegin[g0], x1 = a.gbegin[g0 + 2];
        for (uint32_t x = x0; x < x1; x++) {
            const uint32_t slot = atomicAdd(&s_n, 1u);
            if (slot < (uint32_t)VOTE_MAX_CAND) {
                const DeviceRun &r = a.runs[a.order[x]];
                s_hash[slot] = is_source ? r.src_simhash : r.dst_simhash;
                s_key[slot] = 2 * x + (is_source ? 0u : 1u);
            }
        }
    }
    __syncthreads();
    const uint32_t c = s_n;
    if (c > (uint32_t)VOTE_MAX_CAND) {
        if (tid == 0) atomicOr(a.flags, VOTE_FLAG_FALLBACK);
        return;
    }
    nb200_search_result res;
    memset(&res, 0, sizeof(res));
    if (c == 0) {   // find_best_match -> None (:608-617)
        if (tid == 0) a.results[v] = res;
        return;
    }
    Best best[2];
    best[0].have = best[1].have = 0;
    best[0].score = best[1].score = 0.f;
    best[0].key = best[1].key = 0;
    for (uint32_t i = tid; i < c; i += VOTE_THREADS) {
        const uint32_t h = s_hash[i];
        uint32_t cluster = 0;   // |distinct_matches[i]|, the candidate itself included (:436-454)
        for (uint32_t j = 0; j < c; j++) cluster += ((uint32_t)__popc(h ^ s_hash[j]) < a.bias) ? 1u : 0u;
        if (cluster == 0) continue;
        const uint32_t key = s_key[i];
        const DeviceRun &r = a.runs[a.order[key >> 1]];
        const bool is_source = (key & 1u) == 0;
        const uint64_t start = is_source ? r.src_start_ns : r.dst_start_ns;
        const uint64_t end = is_source ? r.src_end_ns : r.dst_end_ns;
        if (end < start) {
            atomicOr(a.flags, VOTE_FLAG_UNDERFLOW);
            continue;
        }
        const float qa = __fmul_rn(__ll2float_rn((long long)cluster), 0.3f);
        const float qb = __fmul_rn(as_secs_f32_dev(end - start), 0.7f);
        Best cand;
        cand.score = -__fadd_rn(qa, qb);
        cand.key = key;
        cand.have = 1;
        best_merge(best[r.group & 1u], cand);   // [0] openings, [1] endings
    }
    s_best[0][tid] = best[0];
    s_best[1][tid] = best[1];
    __syncthreads();
    for (int step = VOTE_THREADS / 2; step >= 1; step >>= 1) {
        if (tid < step) {
            best_merge(s_best[0][tid], s_best[0][tid + step]);
            best_merge(s_best[1][tid], s_best[1][tid + step]);
        }
        __syncthreads();
    }
    if (tid != 0) return;
    res.present = 1;
    for (int e = 0; e < 2; e++) {
        if (e == 1 && !a.include_endings) break;
        const Best w = s_best[e][0];
        if (!w.have) continue;
        const DeviceRun &r = a.runs[a.order[w.key >> 1]];
        const bool is_source = (w.key & 1u) == 0;
        const uint64_t start = is_source ? r.src_start_ns : r.dst_start_ns;
        const uint64_t end = is_source ? r.src_end_ns : r.dst_end_ns;
        const uint64_t sub = a.time_padding_ns + a.hash_duration[v];
        if (end < sub) {   // end - padding - hash_duration panics in the reference
            atomicOr(a.flags, VOTE_FLAG_UNDERFLOW);
            continue;
        }
        if (e == 0) {
            res.has_opening = 1;
            res.opening_start_ns = start + a.time_padding_ns;
            res.opening_end_ns = end - sub;
        } else {
            res.has_ending = 1;
            res.ending_start_ns = start + a.time_padding_ns;
            res.ending_end_ns = end - sub;
        }
    }
    a.results[v] = res;
}

// CSR of a pair list: for every video the pairs it takes part in, in pair order, tagged
// with whether it is the source.  [n_videos + 1] begins, then 2 * n_pairs entries.
static void build_csr(const uint32_t (*pairs)[2], uint64_t n_pairs, uint32_t n_videos, std::vector<uint32_t> &out) {
    out.assign((size_t)n_videos + 1 + 2 * n_pairs, 0);
    uint32_t *begin = out.data(), *list = out.data() + n_videos + 1;
    for (uint64_t k = 0; k < n_pairs; k++) {
        begin[pairs[k][0] + 1]++;
        begin[pairs[k][1] + 1]++;
    }
    for (uint32_t v = 0; v < n_videos; v++) begin[v + 1] += begin[v];
    std::vector<uint32_t> cursor(begin, begin + n_videos);
    for (uint64_t k = 0; k < n_pairs; k++) {
        list[cursor[pairs[k][0]]++] = (uint32_t)(2 * k + 1);   // source first (a pair (v, v) keeps that order)
        list[cursor[pairs[k][1]]++] = (uint32_t)(2 * k);
    }
}

int vote_enqueue(nb200_ctx *ctx, const DeviceRun *d_runs, const unsigned int *d_n_runs, uint32_t run_cap,
                 const uint64_t *hash_duration_ns, uint32_t n_videos, const uint32_t (*pairs)[2],
                 uint64_t n_pairs, const nb200_match_params *params) {
    cudaStream_t s = ctx->stream;
    const bool default_list = pairs == nullptr;
    std::vector<uint32_t> default_flat, csr_local;
    const std::vector<uint32_t> *csr = nullptr;
    bool upload_csr = true;
    if (default_list) {
        if (ctx->vote_csr_videos != n_videos || ctx->vote_csr_host.empty()) {
            default_pairs(n_videos, default_flat);
            build_csr(reinterpret_cast<const uint32_t(*)[2]>(default_flat.data()), default_flat.size() / 2, n_videos,
                      ctx->vote_csr_host);
            ctx->vote_csr_videos = n_videos;
            ctx->vote_csr_on_device = false;
        }
        n_pairs = ((uint64_t)n_videos * (n_videos > 0 ? n_videos - 1 : 0)) / 2;
        csr = &ctx->vote_csr_host;
        upload_csr = !ctx->vote_csr_on_device;
    } else {
        for (uint64_t k = 0; k < n_pairs; k++)
            if (pairs[k][0] >= n_videos || pairs[k][1] >= n_videos) return NB200_ERR_INVALID_ARGUMENT;
        build_csr(pairs, n_pairs, n_videos, csr_local);
        csr = &csr_local;
        ctx->vote_csr_on_device = false;   // the device copy is about to be overwritten
        ctx->vote_csr_videos = 0;
    }
    if (2 * n_pairs >= 0x7fffffffull || run_cap >= 0x7fffffffu) return NB200_ERR_TOO_LARGE;
    const uint32_t n_groups = (uint32_t)(2 * n_pairs);

    const size_t b_csr = csr->size() * sizeof(uint32_t), b_hd = (size_t)n_videos * sizeof(uint64_t);
    const size_t b_groups = ((size_t)n_groups + 1) * sizeof(uint32_t);
    const size_t b_res = (size_t)n_videos * sizeof(nb200_search_result);
    size_t scan_tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)(n_groups + 1), s);
    if (ctx->d_vote_groups.reserve(3 * b_groups) != NB200_OK || ctx->d_vote_order.reserve((size_t)run_cap * 4) != NB200_OK ||
        ctx->d_vote_csr.reserve(b_csr) != NB200_OK || ctx->d_vote_hd.reserve(b_hd) != NB200_OK ||
        ctx->d_vote_results.reserve(b_res + 16) != NB200_OK || ctx->d_vote_scan_tmp.reserve(scan_tmp) != NB200_OK ||
        ctx->h_vote_stage.reserve((upload_csr ? b_csr : 0) + b_hd) != NB200_OK ||
        ctx->h_results.reserve(b_res + 16) != NB200_OK)
        return NB200_ERR_CUDA;
    // CSR (unless the device already holds it) and hash durations: one staging area, copied on
    // the copy stream underneath the match kernels; the compute stream waits before the vote
    char *st = ctx->h_vote_stage.as<char>();
    if (upload_csr) {
        memcpy(st, csr->data(), b_csr);
        NB_CUDA(cudaMemcpyAsync(ctx->d_vote_csr.p, st, b_csr, cudaMemcpyHostToDevice, ctx->copy_stream));
        if (default_list) ctx->vote_csr_on_device = true;
    }
    memcpy(st + (upload_csr ? b_csr : 0), hash_duration_ns, b_hd);
    NB_CUDA(cudaMemcpyAsync(ctx->d_vote_hd.p, st + (upload_csr ? b_csr : 0), b_hd, cudaMemcpyHostToDevice,
                            ctx->copy_stream));
    NB_CUDA(cudaEventRecord(ctx->ev_chunk[9], ctx->copy_stream));
    NB_CUDA(cudaStreamWaitEvent(s, ctx->ev_chunk[9], 0));

    VoteArgs a;
    a.runs = d_runs;
    a.n_runs_ptr = d_n_runs;
    a.run_cap = run_cap;
    a.n_groups = n_groups;
    a.gcount = ctx->d_vote_groups.as<uint32_t>();
    a.gfill = a.gcount + (n_groups + 1);
    a.gbegin = a.gfill + (n_groups + 1);
    a.order = ctx->d_vote_order.as<uint32_t>();
    a.vid_begin = ctx->d_vote_csr.as<uint32_t>();
    a.vid_pairs = a.vid_begin + (n_videos + 1);
    a.hash_duration = ctx->d_vote_hd.as<uint64_t>();
    a.results = ctx->d_vote_results.as<nb200_search_result>();
    a.flags = reinterpret_cast<uint32_t *>(ctx->d_vote_results.as<char>() + b_res);
    a.n_videos = n_videos;
    a.bias = params->hash_match_threshold + params->hash_match_threshold / 2;
    a.include_endings = params->include_endings;
    a.time_padding_ns = params->time_padding_ns;

    NB_CUDA(cudaEventRecord(ctx->ev[8], s));
    NB_CUDA(cudaMemsetAsync(a.gcount, 0, 2 * b_groups, s));
    NB_CUDA(cudaMemsetAsync(a.flags, 0, 16, s));
    const uint32_t run_grid = std::max<uint32_t>(1u, std::min<uint32_t>((run_cap + 255) / 256, (uint32_t)ctx->sm_count * 4));
    vote_count_kernel<<<run_grid, 256, 0, s>>>(a);
    NB_CUDA(cudaGetLastError());
    NB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->d_vote_scan_tmp.p, scan_tmp, a.gcount, a.gbegin, (int)(n_groups + 1), s));
    vote_scatter_kernel<<<run_grid, 256, 0, s>>>(a);
    NB_CUDA(cudaGetLastError());
    if (n_groups > 0) {
        vote_heap_kernel<<<(n_groups + 127) / 128, 128, 0, s>>>(a);
        NB_CUDA(cudaGetLastError());
    }
    if (n_videos > 0) {
        vote_video_kernel<<<n_videos, VOTE_THREADS, 0, s>>>(a);
        NB_CUDA(cudaGetLastError());
    }
    NB_CUDA(cudaEventRecord(ctx->ev[9], s));
    ctx->ev_valid[4] = true;
    ctx->n_launches += 5;   // count, scan, scatter, heap, video
    NB_CUDA(cudaMemcpyAsync(ctx->h_results.p, ctx->d_vote_results.p, b_res + 16, cudaMemcpyDeviceToHost, s));
    return NB200_OK;
}

// ------------------------------------------------------------------ run blocks
// A rank's match output as one fixed-size message: 64-byte header, then DeviceRun records.
// The match kernel emits straight into the records; this kernel fills in the header from
// the match counters once the kernels before it on the stream are done.
// (RunBlockHeader: common.h)

__global__ void block_header_kernel(RunBlockHeader *h, const unsigned int *__restrict__ counters, uint32_t cap) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t n = counters[1];
    h->n_found = n;
    h->n_stored = min(n, cap);
    h->flags = counters[2] & 1u;
}

// blocks of all ranks (an all-gather receive buffer) -> one contiguous run array + its length
__global__ void block_compact_kernel(const char *__restrict__ blocks, uint32_t n_blocks, uint64_t block_bytes,
                                     DeviceRun *__restrict__ out, unsigned int *__restrict__ total) {
    const uint32_t b = blockIdx.y;
    uint32_t prefix = 0, n = 0, all = 0;
    for (uint32_t q = 0; q < n_blocks; q++) {
        const uint32_t c = reinterpret_cast<const RunBlockHeader *>(blocks + (size_t)q * block_bytes)->n_stored;
        if (q < b) prefix += c;
        if (q == b) n = c;
        all += c;
    }
    if (b == 0 && blockIdx.x == 0 && threadIdx.x == 0) *total = all;
    const uint4 *src = reinterpret_cast<const uint4 *>(blocks + (size_t)b * block_bytes + sizeof(RunBlockHeader));
    uint4 *dst = reinterpret_cast<uint4 *>(out + prefix);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < 4 * n; k += gridDim.x * blockDim.x) dst[k] = src[k];
}

}  // namespace nb200

using namespace nb200;

extern "C" {

int nb200_match_export(nb200_ctx *ctx, const nb200_hashset *hs, const uint32_t (*pairs)[2], uint64_t n_pairs,
                       const nb200_match_params *params, uint32_t pair_base, void *d_block, uint64_t block_bytes) {
    if (!ctx || !hs || !params || !d_block) return NB200_ERR_NULL_ARGUMENT;
    if (block_bytes < 2 * sizeof(RunBlockHeader) || block_bytes % 64 != 0) return NB200_ERR_INVALID_ARGUMENT;
    const uint64_t cap = (block_bytes - sizeof(RunBlockHeader)) / sizeof(DeviceRun);
    if (cap > 0x7fffffffull) return NB200_ERR_TOO_LARGE;
    MatchPending mp;
    // an empty slice (pairs != NULL, n_pairs == 0) still produces a valid, empty block
    NB_TRY(match_enqueue(ctx, hs, pairs, n_pairs, params, pair_base, 0, &mp,
                         reinterpret_cast<DeviceRun *>((char *)d_block + sizeof(RunBlockHeader)), (uint32_t)cap));
    block_header_kernel<<<1, 32, 0, ctx->stream>>>((RunBlockHeader *)d_block, ctx->d_counters.as<unsigned int>(),
                                                   (uint32_t)cap);
    NB_CUDA(cudaGetLastError());
    ctx->n_launches++;
    return NB200_OK;
}

int nb200_vote_blocks(nb200_ctx *ctx, const void *d_blocks, uint32_t n_blocks, uint64_t block_bytes,
                      const uint64_t *hash_duration_ns, uint32_t n_videos, const uint32_t (*pairs)[2],
                      uint64_t n_pairs, const nb200_match_params *params, int do_vote,
                      nb200_search_result *results, uint64_t *max_found) {
    if (block_bytes < 2 * sizeof(RunBlockHeader) || block_bytes % 64 != 0) return NB200_ERR_INVALID_ARGUMENT;
    return vote_blocks_impl(ctx, d_blocks, n_blocks, block_bytes, (block_bytes - sizeof(RunBlockHeader)) / sizeof(DeviceRun),
                            hash_duration_ns, n_videos, pairs, n_pairs, params, do_vote, results, max_found, nullptr, nullptr);
}

}  // extern "C"

// The blocks may sit further apart than they are long (a receive area with fixed slots):
// block_bytes = distance between block starts, cap = records a block can hold.
int nb200::vote_blocks_impl(nb200_ctx *ctx, const void *d_blocks, uint32_t n_blocks, uint64_t block_bytes, uint64_t cap,
                            const uint64_t *hash_duration_ns, uint32_t n_videos, const uint32_t (*pairs)[2],
                            uint64_t n_pairs, const nb200_match_params *params, int do_vote,
                            nb200_search_result *results, uint64_t *max_found, int (*blocks_consumed)(void *),
                            void *consumed_arg) {
    if (!ctx || !d_blocks || !hash_duration_ns || !params || !max_found) return NB200_ERR_NULL_ARGUMENT;
    if (do_vote && !results) return NB200_ERR_NULL_ARGUMENT;
    if (n_blocks == 0 || block_bytes < 2 * sizeof(RunBlockHeader) || block_bytes % 64 != 0 ||
        sizeof(RunBlockHeader) + cap * sizeof(DeviceRun) > block_bytes)
        return NB200_ERR_INVALID_ARGUMENT;
    NB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t total_cap = cap * n_blocks;
    if (total_cap >= 0x7fffffffull) return NB200_ERR_TOO_LARGE;
    *max_found = 0;
    ctx->ev_valid[4] = false;
    // headers of every block for the host (overflow / underflow are decided there, by every rank alike)
    const size_t b_hdr = (size_t)n_blocks * sizeof(RunBlockHeader);
    NB_TRY(ctx->h_runs.reserve(b_hdr));
    NB_CUDA(cudaMemcpy2DAsync(ctx->h_runs.p, sizeof(RunBlockHeader), d_blocks, block_bytes, sizeof(RunBlockHeader),
                              n_blocks, cudaMemcpyDeviceToHost, s));
    if (do_vote) {
        NB_TRY(ctx->d_vote_gather.reserve(total_cap * sizeof(DeviceRun) + 64));
        unsigned int *d_total = reinterpret_cast<unsigned int *>(ctx->d_vote_gather.as<char>() + total_cap * sizeof(DeviceRun));
        block_compact_kernel<<<dim3(16, n_blocks), 256, 0, s>>>((const char *)d_blocks, n_blocks, block_bytes,
                                                               ctx->d_vote_gather.as<DeviceRun>(), d_total);
        NB_CUDA(cudaGetLastError());
        ctx->n_launches++;
        // nothing after this point reads d_blocks: whoever fills them may be told so
        if (blocks_consumed) NB_TRY(blocks_consumed(consumed_arg));
        PhaseTimer pt(ctx, NB200_PHASE_VOTE);
        NB_TRY(vote_enqueue(ctx, ctx->d_vote_gather.as<DeviceRun>(), d_total, (uint32_t)total_cap, hash_duration_ns,
                            n_videos, pairs, n_pairs, params));
    }
    {
        PhaseTimer pt(ctx, NB200_PHASE_MATCH);
        NB_CUDA(cudaStreamSynchronize(s));
    }
    const RunBlockHeader *hdr = ctx->h_runs.as<RunBlockHeader>();
    bool underflow = false, truncated = false;
    uint64_t n_total = 0;
    for (uint32_t b = 0; b < n_blocks; b++) {
        *max_found = std::max<uint64_t>(*max_found, hdr[b].n_found);
        underflow |= (hdr[b].flags & 1u) != 0;
        truncated |= hdr[b].n_found > hdr[b].n_stored;
        n_total += hdr[b].n_stored;
    }
    if (underflow) return NB200_ERR_DURATION_UNDERFLOW;
    if (truncated) return NB200_ERR_TOO_LARGE;   // *max_found tells the caller how big a block must be
    if (!do_vote) return NB200_OK;
    const size_t b_res = (size_t)n_videos * sizeof(nb200_search_result);
    uint32_t flags[4];
    memcpy(flags, ctx->h_results.as<char>() + b_res, sizeof(flags));
    if (!(flags[0] & VOTE_FLAG_FALLBACK)) {
        memcpy(results, ctx->h_results.p, b_res);
        return (flags[0] & VOTE_FLAG_UNDERFLOW) ? NB200_ERR_DURATION_UNDERFLOW : NB200_OK;
    }
    // beyond the device vote's bounds: download the gathered runs, order them, vote on the host
    std::vector<DeviceRun> raw(n_total);
    if (n_total) {
        NB_CUDA(cudaMemcpyAsync(raw.data(), ctx->d_vote_gather.p, n_total * sizeof(DeviceRun), cudaMemcpyDeviceToHost, s));
        NB_CUDA(cudaStreamSynchronize(s));
    }
    std::sort(raw.begin(), raw.end(), [](const DeviceRun &a, const DeviceRun &b) {
        if (a.group != b.group) return a.group < b.group;
        if (a.i_end != b.i_end) return a.i_end > b.i_end;
        return a.j_end > b.j_end;
    });
    std::vector<nb200_run> runs(std::max<size_t>(raw.size(), 1));
    for (size_t k = 0; k < raw.size(); k++) {
        const DeviceRun &r = raw[k];
        runs[k] = {r.group >> 1, r.group & 1u, r.i_end, r.j_end, r.len, r.src_simhash, r.dst_simhash, 0,
                   r.src_start_ns, r.src_end_ns, r.dst_start_ns, r.dst_end_ns};
    }
    PhaseTimer pt(ctx, NB200_PHASE_VOTE);
    return vote_impl(hash_duration_ns, n_videos, pairs, n_pairs, params, runs.data(), raw.size(), nullptr, results);
}
