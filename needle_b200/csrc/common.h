// common.h -- internal declarations shared by the translation units of
// libneedle_b200.so.  Not part of the C ABI (that is include/needle_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <chrono>
#include <memory>
#include <string>
#include <vector>

#include "../../include/needle_b200.h"

namespace nb200 {

void set_last_error(const char *what, cudaError_t e);
void set_last_error_msg(const std::string &msg);

#define NB_CUDA(expr)                                      \
    do {                                                   \
        cudaError_t _e = (expr);                           \
        if (_e != cudaSuccess) {                           \
            ::nb200::set_last_error(#expr, _e);            \
            return NB200_ERR_CUDA;                         \
        }                                                  \
    } while (0)

#define NB_TRY(expr)                   \
    do {                               \
        int _s = (expr);               \
        if (_s != NB200_OK) return _s; \
    } while (0)

// Per-context cache of device allocations: cudaMalloc/cudaFree cost tens of
// microseconds each and cudaFree synchronises the device, which matters when a
// whole step is a few milliseconds.  Blocks are handed back on release and
// reused for requests of similar size.  Shared ownership: objects made by a
// context may outlive it.
struct DevPool {
    struct Block {
        void *p;
        size_t bytes;
    };
    std::vector<Block> free_blocks;
    size_t cached_bytes = 0;
    void *get(size_t want, size_t *got) {
        size_t best = free_blocks.size();
        for (size_t k = 0; k < free_blocks.size(); k++)
            if (free_blocks[k].bytes >= want && free_blocks[k].bytes <= 2 * want + (1u << 16) &&
                (best == free_blocks.size() || free_blocks[k].bytes < free_blocks[best].bytes))
                best = k;
        if (best == free_blocks.size()) return nullptr;
        Block b = free_blocks[best];
        free_blocks.erase(free_blocks.begin() + (long)best);
        cached_bytes -= b.bytes;
        *got = b.bytes;
        return b.p;
    }
    void put(void *p, size_t bytes) {
        free_blocks.push_back({p, bytes});
        cached_bytes += bytes;
        // keep the cache bounded: drop the largest blocks beyond 8 GiB / 64 entries
        while (free_blocks.size() > 64 || cached_bytes > (8ull << 30)) {
            size_t big = 0;
            for (size_t k = 1; k < free_blocks.size(); k++)
                if (free_blocks[k].bytes > free_blocks[big].bytes) big = k;
            cudaFree(free_blocks[big].p);
            cached_bytes -= free_blocks[big].bytes;
            free_blocks.erase(free_blocks.begin() + (long)big);
        }
    }
    ~DevPool() {
        for (auto &b : free_blocks) cudaFree(b.p);
    }
};

// Simple RAII device buffer (optionally backed by a DevPool).
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    std::shared_ptr<DevPool> pool;
    bool borrowed = false;   // points into memory owned by the caller
    void borrow(void *q, size_t n) {
        release();
        p = q;
        bytes = n;
        borrowed = true;
    }
    int alloc(size_t n, const std::shared_ptr<DevPool> &from = nullptr) {
        release();
        if (n == 0) n = 16;
        if (from) {
            size_t got = 0;
            if (void *q = from->get(n, &got)) {
                p = q;
                bytes = got;
                pool = from;
                return NB200_OK;
            }
        }
        NB_CUDA(cudaMalloc(&p, n));
        bytes = n;
        pool = from;
        return NB200_OK;
    }
    // Grow-only (amortised scratch).
    int reserve(size_t n) {
        if (n <= bytes) return NB200_OK;
        return alloc(n + n / 4);
    }
    void release() {
        if (p && !borrowed) {
            if (pool) pool->put(p, bytes); else cudaFree(p);
        }
        p = nullptr;
        bytes = 0;
        borrowed = false;
        pool.reset();
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(DevBuf &&o) noexcept : p(o.p), bytes(o.bytes), pool(std::move(o.pool)), borrowed(o.borrowed) {
        o.p = nullptr;
        o.bytes = 0;
        o.borrowed = false;
    }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinnedBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int reserve(size_t n) {
        if (n <= bytes) return NB200_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        n += n / 4;
        NB_CUDA(cudaHostAlloc(&p, n, cudaHostAllocDefault));
        bytes = n;
        return NB200_OK;
    }
    ~PinnedBuf() {
        if (p) cudaFreeHost(p);
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

}  // namespace nb200

// ---------------------------------------------------------------- objects

namespace nb200 {
uint64_t next_object_uid();
// What the host derived for the last match stage on a context and left in its device arena
// (match.cu: match_enqueue); reused when the next stage has the same lists, pairs and parameters.
struct MatchPlanCache {
    bool valid = false;
    uint64_t key = 0;
    uint64_t data_uid = 0;   // immutable hashset whose bounds and repeated lists the arena / d_dst2 hold (0: none)
    uint32_t n_problems = 0, d2_max = 1, max_len = 1;
    uint64_t tiles_f = 0, tiles_g = 0, cells = 0, d2_total = 0;   // tiles_f counts the fast kernel's units
    uint32_t band_group = 1;
    size_t o_tab = 0, o_lmin = 0, o_d2 = 0, o_prob_f = 0, o_tend_f = 0, o_lut_f = 0, o_prob_g = 0, o_tend_g = 0, o_lut_g = 0;
    std::vector<uint32_t> problem_pair;
    std::vector<uint8_t> problem_is_ending;
};
}  // namespace nb200

struct nb200_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;     // the one in use (own or caller's)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev[10] = {};           // pairs: K1, K2, K3, K4, device vote
    cudaEvent_t ev_chunk[13] = {};     // [0..5] H2D chunk landed (pipelined analyze), [6] tsfill, [7] PCM guard,
                                       // [8] match descriptors landed, [9] vote descriptors landed,
                                       // [10] the last enqueued match stage (arena + staging readers) is done,
                                       // [11] the last enqueued fingerprint stage is done, [12] hashset upload staging drained
    bool ev_valid[5] = {false, false, false, false, false};
    uint64_t n_launches = 0;
    double host_ms[NB200_HOST_PHASES] = {};   // wall time spent in each host phase since the last reset
    uint32_t match_band_group = 0;      // NB200_OPT_MATCH_BAND_GROUP (0: chosen from the tile count)
    bool match_dense = false;           // fast match kernel evaluates every cell instead of 4 sampled rows per word
    bool force_general_match = false;   // tests: run the general kernel where the fast one would be chosen
    std::shared_ptr<nb200::DevPool> pool = std::make_shared<nb200::DevPool>();
    // scratch reused across calls
    bool defer_wait = false;            // *_into fingerprint calls return without waiting for the stream
    bool host_vote = false;             // searches download the runs and vote on the host (vote.cpp) instead of on the device
    nb200::DevBuf d_problems;   // the match stage's descriptor arena: segment table, run-length bounds, both tilings
    nb200::MatchPlanCache match_plan;   // what that arena currently holds
    nb200::DevBuf d_counters, d_runs, d_dst2;   // d_dst2: the fast match kernel's repeated dst lists
    nb200::PinnedBuf h_stage, h_small, h_runs, h_tsfill;
    nb200::PinnedBuf h_upload;   // nb200_hashset_upload's staging area (guarded by ev_chunk[12])
    nb200::DevBuf d_upload;      // ... and its device scratch for page-locked input (packed arrays + offset tables)
    nb200::PinnedBuf h_fp_stage, h_vote_stage, h_results;   // one staging area per phase: nothing waits for a copy to drain
    // device vote scratch (vote_device.cu)
    nb200::DevBuf d_vote_groups, d_vote_order, d_vote_csr, d_vote_hd, d_vote_results, d_vote_flags, d_vote_scan_tmp,
        d_vote_gather;
    std::vector<uint32_t> vote_csr_host;   // cached CSR of the default pair list
    uint32_t vote_csr_videos = 0;          // n_videos the cache was built for (0: none)
    bool vote_csr_on_device = false;
    nb200::DevBuf d_tsfill;
    // fingerprint stage: constant tables (built once per context) and scratch
    bool fp_ready = false;
    nb200::DevBuf d_fp_window, d_fp_twiddle, d_fp_lane, d_fp_notes, d_fp_wsplit, d_fp_window_odd;
    int k1_variant = 0;   // NB200_OPT_K1_VARIANT (include/needle_b200.h)
    nb200::DevBuf d_chroma, d_fp_desc, d_stereo;
};

// Device layout of a season: every segment starts at a multiple of 4 elements
// (16 B, the granularity of cp.async.bulk) in d_hashes / d_ts.
struct nb200_hashset {
    uint32_t n_videos = 0;
    std::vector<uint64_t> seg_offset;   // caller's packed offsets [2N+1]
    std::vector<uint32_t> dev_offset;   // aligned device offsets [2N]
    std::vector<uint32_t> seg_len;      // [2N]
    uint64_t total = 0;                 // packed element count
    uint64_t dev_total = 0;             // aligned element count
    nb200::DevBuf d_hashes, d_ts;
    bool sorted_layout = true;          // dev_offset ascending with the segment index (false for views)
    // Identity of the DATA, for what the match stage derives from it on the device (run-length bounds,
    // repeated dst lists): a set the library filled itself never changes after creation; a view is
    // over buffers the caller (K2, an all-gather) may rewrite between two searches.
    uint64_t uid = nb200::next_object_uid();
    bool immutable = true;
};

struct nb200_pcmset {
    uint32_t n_segments = 0;
    std::vector<uint64_t> n_mono;         // mono sample count per segment
    std::vector<uint64_t> dev_offset;     // element (i16) offset of each segment, 8-aligned
    uint64_t dev_total = 0;
    nb200::DevBuf d_pcm;                  // mono i16 (stereo input is down-mixed on upload)
};

struct nb200_runset {
    uint64_t n_runs = 0;
    uint64_t n_cells = 0;
    std::vector<uint32_t> problem_pair;     // per problem: pair index
    std::vector<uint8_t> problem_is_ending;
    nb200::DevBuf d_runs;                   // DeviceRun[n_runs], unsorted, simhashes filled
};

namespace nb200 {

// Accumulates the wall time of a scope into ctx->host_ms[phase].
struct PhaseTimer {
    nb200_ctx *ctx;
    int phase;
    std::chrono::steady_clock::time_point t0;
    PhaseTimer(nb200_ctx *c, int p) : ctx(c), phase(p), t0(std::chrono::steady_clock::now()) {}
    ~PhaseTimer() {
        ctx->host_ms[phase] +=
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};

// match.cu ---------------------------------------------------------------
struct DeviceRun {   // 64 bytes
    uint32_t problem;
    uint32_t i_end, j_end, len;
    uint32_t src_simhash, dst_simhash;
    uint32_t group;      // 2 * (global pair index) + is_ending; stamped by the simhash kernel
    uint32_t reserved;
    uint64_t src_start_ns, src_end_ns, dst_start_ns, dst_end_ns;
};
static_assert(sizeof(DeviceRun) == 64, "DeviceRun is exchanged between ranks as raw bytes");

// A rank's match output as one message: this 64-byte header, then DeviceRun records (vote_device.cu, multi.cu)
struct RunBlockHeader {
    uint32_t n_found;    // runs the match found
    uint32_t n_stored;   // min(n_found, capacity): the records that follow
    uint32_t flags;      // bit 0: a Duration subtraction would have underflowed
    uint32_t reserved[13];
};
static_assert(sizeof(RunBlockHeader) == 64, "header and records share one 64-byte grid");

// What match_enqueue left on ctx's stream and in ctx's scratch (d_runs, d_counters):
// nothing has been waited for.  counters[1] = number of runs found (may exceed
// run_cap: then only run_cap were stored and the caller re-runs with more room),
// counters[2] bit 0 = a Duration subtraction would have underflowed.
struct MatchPending {
    uint32_t n_problems = 0;
    uint32_t run_cap = 0;
    uint64_t n_cells = 0;
    std::vector<uint32_t> problem_pair;
    std::vector<uint8_t> problem_is_ending;
};
int match_enqueue(nb200_ctx *ctx, const nb200_hashset *hs, const uint32_t (*pairs)[2], uint64_t n_pairs,
                  const nb200_match_params *params, uint32_t pair_base, uint32_t run_cap_min, MatchPending *out,
                  DeviceRun *ext_runs = nullptr, uint32_t ext_cap = 0);
// Waits for the stream and reads the counters: DURATION_UNDERFLOW, or OK with the number of
// runs found (> run_cap: enqueue again with run_cap_min = that number).
int match_collect(nb200_ctx *ctx, const MatchPending &mp, uint64_t *n_runs);
int match_run_impl(nb200_ctx *ctx, const nb200_hashset *hs, const uint32_t (*pairs)[2],
                   uint64_t n_pairs, const nb200_match_params *params, nb200_runset **out);

// vote_device.cu ---------------------------------------------------------
// Device-side heap replay + find_best_match over runs that are already in HBM.
// Enqueues on ctx->stream; results land in ctx->h_results (pinned) after the
// caller synchronises: [n_videos] nb200_search_result, then 4 u32 flags
// (VOTE_FLAG_*).  d_n_runs points at the device-resident run count.
constexpr uint32_t VOTE_FLAG_UNDERFLOW = 1u;   // a Duration subtraction would have underflowed
constexpr uint32_t VOTE_FLAG_FALLBACK = 2u;    // a table or a video exceeds the device vote's bounds: vote on the host
int vote_enqueue(nb200_ctx *ctx, const DeviceRun *d_runs, const unsigned int *d_n_runs, uint32_t run_cap,
                 const uint64_t *hash_duration_ns, uint32_t n_videos, const uint32_t (*pairs)[2],
                 uint64_t n_pairs, const nb200_match_params *params);

int vote_blocks_impl(nb200_ctx *ctx, const void *d_blocks, uint32_t n_blocks, uint64_t block_bytes, uint64_t cap,
                     const uint64_t *hash_duration_ns, uint32_t n_videos, const uint32_t (*pairs)[2],
                     uint64_t n_pairs, const nb200_match_params *params, int do_vote,
                     nb200_search_result *results, uint64_t *max_found, int (*blocks_consumed)(void *) = nullptr,
                     void *consumed_arg = nullptr);

// vote.cpp ---------------------------------------------------------------
void default_pairs(uint32_t n_videos, std::vector<uint32_t> &flat);   // (i<j) reference order
int vote_impl(const uint64_t *hash_duration_ns, uint32_t n_videos, const uint32_t (*pairs)[2],
              uint64_t n_pairs, const nb200_match_params *params, const nb200_run *runs,
              uint64_t n_runs, const uint8_t *video_mask, nb200_search_result *results);

// api.cu ------------------------------------------------------------------
// A season uploaded 1/world per rank (multi.cu's search job): every rank copies ONE slice of the packed
// byte image [ts | hashes] of the caller's arrays into its scratch, the ranks all-gather the scratch in
// place over NVLink (the caller: d_pack + rank * slice_bytes -> d_pack, slice_bytes each), then every rank
// moves the image into its own aligned season.  Host staging and PCIe traffic per rank fall by `world`.
struct SliceUpload {
    char *d_pack = nullptr;
    size_t slice_bytes = 0;
};
int hashset_upload_slice(nb200_ctx *ctx, const uint32_t *hashes, const uint64_t *ts_ns, const uint64_t *seg_offset,
                         uint32_t n_videos, int rank, int world, SliceUpload *su, nb200_hashset **out);
int hashset_upload_finish(nb200_ctx *ctx, nb200_hashset *hs, const SliceUpload &su, int world);

// fingerprint.cu -----------------------------------------------------------
int fingerprint_run_impl(nb200_ctx *ctx, const nb200_pcmset *ps, uint32_t stride, uint64_t delay_ns,
                         uint64_t item_ns, const uint64_t *seek_to_ns, nb200_hashset **out);
uint64_t num_frames(uint64_t n_mono);

}  // namespace nb200
