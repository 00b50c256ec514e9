// api.cu -- the C ABI of include/needle_b200.h: status/error plumbing, contexts,
// device-resident seasons (hashsets), and the match entry points that replace
// the pair loop of Comparator::run_with_frame_hashes
// (needle/src/audio/comparator.rs:532-578).  The kernels live in match.cu and
// fingerprint.cu, the host vote in vote.cpp, .needle.dat I/O in persist.cpp.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#include "common.h"

namespace nb200 {

static thread_local std::string g_last_error;
uint64_t next_object_uid() {
    static std::atomic<uint64_t> counter{1};
    return counter.fetch_add(1);
}

void set_last_error(const char *what, cudaError_t e) {
    g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
}
void set_last_error_msg(const std::string &msg) { g_last_error = msg; }

// Packed (caller) layout -> aligned device layout, one thread per element.
__global__ void repack_kernel(const uint32_t *__restrict__ src_h, const uint64_t *__restrict__ src_t,
                              uint32_t *__restrict__ dst_h, uint64_t *__restrict__ dst_t,
                              const uint64_t *__restrict__ seg_offset,
                              const uint32_t *__restrict__ dev_offset, uint32_t n_seg, uint64_t total) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    // last segment s with seg_offset[s] <= idx
    uint32_t lo = 0, hi = n_seg;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (seg_offset[mid] <= idx) lo = mid; else hi = mid;
    }
    const uint64_t o = dev_offset[lo] + (idx - seg_offset[lo]);
    dst_h[o] = src_h[idx];
    dst_t[o] = src_t[idx];
}

__global__ void unpack_kernel(const uint32_t *__restrict__ src_h, const uint64_t *__restrict__ src_t,
                              uint32_t *__restrict__ dst_h, uint64_t *__restrict__ dst_t,
                              const uint64_t *__restrict__ seg_offset,
                              const uint32_t *__restrict__ dev_offset, uint32_t n_seg, uint64_t total) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    uint32_t lo = 0, hi = n_seg;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (seg_offset[mid] <= idx) lo = mid; else hi = mid;
    }
    const uint64_t o = dev_offset[lo] + (idx - seg_offset[lo]);
    dst_h[idx] = src_h[o];
    dst_t[idx] = src_t[o];
}

// Gathers segments that sit at arbitrary element offsets of two source arrays
// (e.g. the receive buffer of an all-gather) into the aligned device layout.
__global__ void scatter_gather_kernel(const uint32_t *__restrict__ src_h, const uint64_t *__restrict__ src_t,
                                      uint32_t *__restrict__ dst_h, uint64_t *__restrict__ dst_t,
                                      const uint64_t *__restrict__ seg_offset,   // packed prefix [n_seg+1]
                                      const uint64_t *__restrict__ src_hash_off,
                                      const uint64_t *__restrict__ src_ts_off,
                                      const uint32_t *__restrict__ dev_offset, uint32_t n_seg, uint64_t total) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    uint32_t lo = 0, hi = n_seg;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (seg_offset[mid] <= idx) lo = mid; else hi = mid;
    }
    const uint64_t local = idx - seg_offset[lo];
    const uint64_t o = dev_offset[lo] + local;
    dst_h[o] = src_h[src_hash_off[lo] + local];
    dst_t[o] = src_t[src_ts_off[lo] + local];
}

// Fills the layout vectors of a hashset from packed offsets.
static int layout_hashset(nb200_hashset *hs, const uint64_t *seg_offset, uint32_t n_videos) {
    hs->n_videos = n_videos;
    const uint32_t nseg = 2 * n_videos;
    hs->seg_offset.assign(seg_offset, seg_offset + nseg + 1);
    hs->dev_offset.resize(nseg);
    hs->seg_len.resize(nseg);
    uint64_t dev = 0;
    for (uint32_t s = 0; s < nseg; s++) {
        if (seg_offset[s + 1] < seg_offset[s]) return NB200_ERR_INVALID_ARGUMENT;
        const uint64_t len = seg_offset[s + 1] - seg_offset[s];
        if (len > 0x7fffff00ull) return NB200_ERR_TOO_LARGE;
        hs->dev_offset[s] = (uint32_t)dev;
        hs->seg_len[s] = (uint32_t)len;
        dev += (len + 3) & ~3ull;
        if (dev > 0xfffffff0ull) return NB200_ERR_TOO_LARGE;
    }
    hs->total = seg_offset[nseg] - seg_offset[0];
    hs->dev_total = dev;
    return NB200_OK;
}

}  // namespace nb200

using namespace nb200;

extern "C" {

const char *nb200_status_str(int status) {
    switch (status) {
    case NB200_OK: return "ok";
    case NB200_ERR_NULL_ARGUMENT: return "null argument";
    case NB200_ERR_INVALID_ARGUMENT: return "invalid argument";
    case NB200_ERR_CUDA: return "CUDA error (no sm_100a device, or a runtime failure; see nb200_last_error)";
    case NB200_ERR_NO_ENDING: return "frame hash data has no ending";
    case NB200_ERR_DURATION_UNDERFLOW: return "timestamp subtraction underflow (the reference panics)";
    case NB200_ERR_TOO_LARGE: return "input too large for the device layout";
    case NB200_ERR_IO: return "I/O error";
    case NB200_ERR_FORMAT: return "invalid frame hash data";
    case NB200_ERR_STATE: return "call order violated";
    case NB200_ERR_COMPARATOR_MINIMUM_PATHS: return "comparator needs at least two videos";
    case NB200_ERR_NCCL: return "NCCL error (library not found, or a collective failed; see nb200_last_error)";
    default: return "unknown status";
    }
}

const char *nb200_last_error(void) { return g_last_error.c_str(); }

/* ----------------------------------------------------------------- context */

int nb200_ctx_create(int device, nb200_ctx **out) {
    if (!out) return NB200_ERR_NULL_ARGUMENT;
    *out = nullptr;
    int count = 0;
    NB_CUDA(cudaGetDeviceCount(&count));
    if (count <= 0) {
        set_last_error_msg("no CUDA device");
        return NB200_ERR_CUDA;
    }
    if (device < 0) NB_CUDA(cudaGetDevice(&device));
    if (device >= count) return NB200_ERR_INVALID_ARGUMENT;
    NB_CUDA(cudaSetDevice(device));
    nb200_ctx *ctx = new (std::nothrow) nb200_ctx();
    if (!ctx) return NB200_ERR_TOO_LARGE;
    ctx->device = device;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        set_last_error("cudaGetDeviceProperties", e);
        delete ctx;
        return NB200_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int k = 0; ok && k < 10; k++) ok = cudaEventCreate(&ctx->ev[k]) == cudaSuccess;
    for (int k = 0; ok && k < 13; k++)
        ok = cudaEventCreateWithFlags(&ctx->ev_chunk[k], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && ctx->h_small.reserve(256) == NB200_OK;
    if (!ok) {
        set_last_error("nb200_ctx_create", cudaGetLastError());
        nb200_ctx_destroy(ctx);
        return NB200_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    // tuning knob for test and profiling runs: same values as NB200_OPT_K1_VARIANT
    if (const char *v = getenv("NB200_K1_VARIANT")) nb200_ctx_set_option(ctx, NB200_OPT_K1_VARIANT, atoll(v));
    *out = ctx;
    return NB200_OK;
}

void nb200_ctx_destroy(nb200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    for (int k = 0; k < 10; k++)
        if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
    for (int k = 0; k < 13; k++)
        if (ctx->ev_chunk[k]) cudaEventDestroy(ctx->ev_chunk[k]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

int nb200_ctx_set_stream(nb200_ctx *ctx, void *cuda_stream) {
    if (!ctx) return NB200_ERR_NULL_ARGUMENT;
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return NB200_OK;
}

int nb200_ctx_set_option(nb200_ctx *ctx, int option, int64_t value) {
    if (!ctx) return NB200_ERR_NULL_ARGUMENT;
    switch (option) {
    case NB200_OPT_FORCE_GENERAL_MATCH:
        ctx->force_general_match = value != 0;
        return NB200_OK;
    case NB200_OPT_MATCH_DENSE:
        ctx->match_dense = value != 0;
        return NB200_OK;
    case NB200_OPT_DEFER_WAIT:
        ctx->defer_wait = value != 0;
        return NB200_OK;
    case NB200_OPT_MATCH_BAND_GROUP:
        if (value < 0 || value > 16) return NB200_ERR_INVALID_ARGUMENT;
        ctx->match_band_group = (uint32_t)value;
        return NB200_OK;
    case NB200_OPT_HOST_VOTE:
        ctx->host_vote = value != 0;
        return NB200_OK;
    case NB200_OPT_K1_VARIANT:
        if (value != 0 && value != 1 && value != 8 && value != 10 && value != 12 && value != 16 && value != 17 &&
            value != 18 && value != 19 && value != 112)
            return NB200_ERR_INVALID_ARGUMENT;
        ctx->k1_variant = (int)value;
        return NB200_OK;
    default:
        return NB200_ERR_INVALID_ARGUMENT;
    }
}

int nb200_ctx_host_profile(nb200_ctx *ctx, double *ms_out, int reset) {
    if (!ctx) return NB200_ERR_NULL_ARGUMENT;
    for (int k = 0; k < NB200_HOST_PHASES; k++) {
        if (ms_out) ms_out[k] = ctx->host_ms[k];
        if (reset) ctx->host_ms[k] = 0.0;
    }
    return NB200_OK;
}

int nb200_ctx_synchronize(nb200_ctx *ctx) {
    if (!ctx) return NB200_ERR_NULL_ARGUMENT;
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

int nb200_ctx_last_kernel_ms(nb200_ctx *ctx, float ms[4], uint64_t *n_launches) {
    if (!ctx || !ms) return NB200_ERR_NULL_ARGUMENT;
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 4; k++) {
        ms[k] = 0.f;
        if (ctx->ev_valid[k]) NB_CUDA(cudaEventElapsedTime(&ms[k], ctx->ev[2 * k], ctx->ev[2 * k + 1]));
    }
    if (n_launches) *n_launches = ctx->n_launches;
    return NB200_OK;
}

int nb200_ctx_last_vote_ms(nb200_ctx *ctx, float *ms) {
    if (!ctx || !ms) return NB200_ERR_NULL_ARGUMENT;
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    *ms = 0.f;
    if (ctx->ev_valid[4]) NB_CUDA(cudaEventElapsedTime(ms, ctx->ev[8], ctx->ev[9]));
    return NB200_OK;
}

int nb200_host_alloc(void **out, size_t bytes) {
    if (!out) return NB200_ERR_NULL_ARGUMENT;
    *out = nullptr;
    NB_CUDA(cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault));
    return NB200_OK;
}

int nb200_host_free(void *p) {
    if (p) NB_CUDA(cudaFreeHost(p));
    return NB200_OK;
}

void nb200_free(void *p) { free(p); }

void nb200_match_params_default(nb200_match_params *p) {
    if (!p) return;
    p->hash_match_threshold = 10;          // DEFAULT_HASH_MATCH_THRESHOLD (audio/mod.rs:14)
    p->include_endings = 0;
    p->min_opening_ns = 20ull * 1000000000ull;   // DEFAULT_MIN_OPENING_DURATION
    p->min_ending_ns = 20ull * 1000000000ull;    // DEFAULT_MIN_ENDING_DURATION
    p->time_padding_ns = 0;
}

/* ----------------------------------------------------------------- hashset */

namespace {
struct CopySpan {
    void *dst;
    const void *src;
    size_t bytes;
};
// memcpy of a list of spans; above 2 MB shared between up to 4 threads (contiguous groups of equal bytes)
void staged_copy(const std::vector<CopySpan> &spans) {
    size_t total = 0;
    for (const CopySpan &c : spans) total += c.bytes;
    unsigned n_threads = 1;
    if (total >= (size_t(2) << 20)) n_threads = std::min(4u, std::max(1u, std::thread::hardware_concurrency()));
    auto run = [&spans](size_t from, size_t to) {
        for (size_t k = from; k < to; k++)
            if (spans[k].bytes) memcpy(spans[k].dst, spans[k].src, spans[k].bytes);
    };
    if (n_threads == 1) {
        run(0, spans.size());
        return;
    }
    std::vector<size_t> cut(n_threads + 1, spans.size());
    cut[0] = 0;
    size_t acc = 0;
    unsigned g = 1;
    for (size_t k = 0; k < spans.size() && g < n_threads; k++) {
        acc += spans[k].bytes;
        while (g < n_threads && acc >= total * g / n_threads) cut[g++] = k + 1;
    }
    std::vector<std::thread> workers;
    size_t done = cut[1];   // the caller takes group 0 ...
    try {
        for (unsigned t = 1; t < n_threads; t++) {
            workers.emplace_back(run, cut[t], cut[t + 1]);
            done = cut[t + 1];
        }
    } catch (const std::system_error &) {
        // ... and whatever no thread could be started for
    }
    run(0, cut[1]);
    if (workers.size() + 1 < n_threads) run(done, spans.size());
    for (std::thread &w : workers) w.join();
}
}   // namespace

namespace {
// True when [p, p + bytes) is page-locked host memory the device can copy from directly
// (nb200_host_alloc, cudaHostAlloc, cudaHostRegister).
bool is_pinned_host(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}
}   // namespace

// The staging area of the uploads may still feed the previous upload's copies: wait for them before the
// area is written again -- or replaced by a larger one (PinnedBuf::reserve frees the old block).
static int upload_staging_idle(nb200_ctx *ctx) {
    if (cudaEventSynchronize(ctx->ev_chunk[12]) != cudaSuccess) {
        set_last_error("cudaEventSynchronize(upload staging)", cudaGetLastError());
        return NB200_ERR_CUDA;
    }
    return NB200_OK;
}

// `wait`: the caller's arrays may be reused as soon as this returns (the public upload).  The fused calls
// (nb200_search, nb200_match_pairs) pass false: they wait for the whole pipeline before they return.
static int hashset_upload_impl(nb200_ctx *ctx, const uint32_t *hashes, const uint64_t *ts_ns,
                               const uint64_t *seg_offset, uint32_t n_videos, bool wait, nb200_hashset **out) {
    if (!ctx || !seg_offset || !out) return NB200_ERR_NULL_ARGUMENT;
    *out = nullptr;
    if (seg_offset[0] != 0) return NB200_ERR_INVALID_ARGUMENT;
    const uint64_t total = seg_offset[2 * (uint64_t)n_videos];
    if (total > 0 && (!hashes || !ts_ns)) return NB200_ERR_NULL_ARGUMENT;
    NB_CUDA(cudaSetDevice(ctx->device));
    nb200_hashset *hs = new (std::nothrow) nb200_hashset();
    if (!hs) return NB200_ERR_TOO_LARGE;
    int st = layout_hashset(hs, seg_offset, n_videos);
    if (st != NB200_OK) {
        delete hs;
        return st;
    }
    if (hs->total > 0 && is_pinned_host(hashes) && is_pinned_host(ts_ns)) {
        // Page-locked input: no host staging at all.  Two contiguous copies of the caller's PACKED arrays into
        // a device scratch, the offset tables behind them, and one kernel that moves every element to its
        // aligned place (the device -> device route of nb200_hashset_from_device).
        const uint32_t nseg = 2 * n_videos;
        const size_t p_t = (size_t)hs->total * sizeof(uint64_t), p_h = ((size_t)hs->total * sizeof(uint32_t) + 15) & ~size_t(15);
        const size_t b_off = (size_t)(nseg + 1) * sizeof(uint64_t), b_dev = (size_t)nseg * sizeof(uint32_t);
        cudaStream_t s = ctx->stream;
        if (upload_staging_idle(ctx) != NB200_OK ||
            ctx->d_upload.reserve(p_t + p_h + b_off + b_dev) != NB200_OK || ctx->h_upload.reserve(b_off + b_dev) != NB200_OK ||
            hs->d_hashes.alloc((size_t)hs->dev_total * sizeof(uint32_t), ctx->pool) != NB200_OK ||
            hs->d_ts.alloc((size_t)hs->dev_total * sizeof(uint64_t), ctx->pool) != NB200_OK) {
            delete hs;
            return NB200_ERR_CUDA;
        }
        char *scratch = ctx->d_upload.as<char>();
        memcpy(ctx->h_upload.p, hs->seg_offset.data(), b_off);
        memcpy(ctx->h_upload.as<char>() + b_off, hs->dev_offset.data(), b_dev);
        cudaError_t e = cudaMemcpyAsync(scratch + p_t, hashes, (size_t)hs->total * sizeof(uint32_t), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(scratch, ts_ns, p_t, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(scratch + p_t + p_h, ctx->h_upload.p, b_off + b_dev, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_chunk[12], s);
        if (e == cudaSuccess) e = cudaMemsetAsync(hs->d_hashes.p, 0, (size_t)hs->dev_total * sizeof(uint32_t), s);
        if (e == cudaSuccess) e = cudaMemsetAsync(hs->d_ts.p, 0, (size_t)hs->dev_total * sizeof(uint64_t), s);
        if (e == cudaSuccess) {
            repack_kernel<<<(unsigned)((hs->total + 255) / 256), 256, 0, s>>>(
                reinterpret_cast<const uint32_t *>(scratch + p_t), reinterpret_cast<const uint64_t *>(scratch),
                hs->d_hashes.as<uint32_t>(), hs->d_ts.as<uint64_t>(), reinterpret_cast<const uint64_t *>(scratch + p_t + p_h),
                reinterpret_cast<const uint32_t *>(scratch + p_t + p_h + b_off), nseg, hs->total);
            e = cudaGetLastError();
            ctx->n_launches++;
        }
        if (e == cudaSuccess && wait) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            cudaStreamSynchronize(s);
            set_last_error("nb200_hashset_upload (page-locked input)", e);
            delete hs;
            return NB200_ERR_CUDA;
        }
        *out = hs;
        return NB200_OK;
    }
    // Stage in the aligned layout (pad entries are zero and never read unmasked) in this call's own pinned
    // area and leave the two copies to the stream: nothing here waits for the device -- the next upload
    // waits for ev_chunk[12] before it overwrites the area.  Hashes go first, so that their copy runs
    // underneath the staging of the timestamps (2/3 of the bytes); large seasons are staged by a few threads
    // (one core moves ~10 GB/s: 1 ms for the 10.4 MB of a 200-episode season, as long as a fifth of its search).
    const size_t b_h = (size_t)hs->dev_total * sizeof(uint32_t);
    const size_t b_t = (size_t)hs->dev_total * sizeof(uint64_t);
    if (upload_staging_idle(ctx) != NB200_OK || ctx->h_upload.reserve(b_h + b_t + 16) != NB200_OK ||
        hs->d_hashes.alloc(b_h, ctx->pool) != NB200_OK || hs->d_ts.alloc(b_t, ctx->pool) != NB200_OK) {
        delete hs;
        return NB200_ERR_CUDA;
    }
    uint64_t *st_t = ctx->h_upload.as<uint64_t>();
    uint32_t *st_h = reinterpret_cast<uint32_t *>(st_t + hs->dev_total);
    const uint32_t nseg = 2 * n_videos;
    std::vector<CopySpan> spans_h(nseg), spans_t(nseg);
    for (uint32_t s = 0; s < nseg; s++) {
        const uint64_t at = hs->dev_offset[s], len = hs->seg_len[s];
        const uint64_t next = s + 1 < nseg ? hs->dev_offset[s + 1] : hs->dev_total;
        for (uint64_t k = at + len; k < next; k++) {   // at most 3 pad entries
            st_h[k] = 0;
            st_t[k] = 0;
        }
        spans_h[s] = {st_h + at, hashes + seg_offset[s], (size_t)len * sizeof(uint32_t)};
        spans_t[s] = {st_t + at, ts_ns + seg_offset[s], (size_t)len * sizeof(uint64_t)};
    }
    cudaError_t e = cudaSuccess;
    staged_copy(spans_h);
    if (hs->dev_total > 0) e = cudaMemcpyAsync(hs->d_hashes.p, st_h, b_h, cudaMemcpyHostToDevice, ctx->stream);
    staged_copy(spans_t);
    if (hs->dev_total > 0 && e == cudaSuccess)
        e = cudaMemcpyAsync(hs->d_ts.p, st_t, b_t, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_chunk[12], ctx->stream);
    if (e != cudaSuccess) {
        cudaStreamSynchronize(ctx->stream);   // a copy may be reading the staging area
        set_last_error("nb200_hashset_upload", e);
        delete hs;
        return NB200_ERR_CUDA;
    }
    *out = hs;
    return NB200_OK;
}

}   // extern "C"

namespace nb200 {

int hashset_upload_slice(nb200_ctx *ctx, const uint32_t *hashes, const uint64_t *ts_ns, const uint64_t *seg_offset,
                         uint32_t n_videos, int rank, int world, SliceUpload *su, nb200_hashset **out) {
    if (!ctx || !seg_offset || !su || !out || !hashes || !ts_ns) return NB200_ERR_NULL_ARGUMENT;
    *out = nullptr;
    if (seg_offset[0] != 0 || world < 1 || rank < 0 || rank >= world) return NB200_ERR_INVALID_ARGUMENT;
    NB_CUDA(cudaSetDevice(ctx->device));
    nb200_hashset *hs = new (std::nothrow) nb200_hashset();
    if (!hs) return NB200_ERR_TOO_LARGE;
    int st = layout_hashset(hs, seg_offset, n_videos);
    if (st == NB200_OK && hs->total == 0) st = NB200_ERR_INVALID_ARGUMENT;   // callers take the plain upload for an empty season
    if (st != NB200_OK) {
        delete hs;
        return st;
    }
    const uint32_t nseg = 2 * n_videos;
    const size_t p_t = (size_t)hs->total * sizeof(uint64_t), image = p_t + (size_t)hs->total * sizeof(uint32_t);
    const size_t slice = ((image + (size_t)world - 1) / (size_t)world + 15) & ~size_t(15);
    const size_t b_off = (size_t)(nseg + 1) * sizeof(uint64_t), b_dev = (size_t)nseg * sizeof(uint32_t);
    if (upload_staging_idle(ctx) != NB200_OK || ctx->d_upload.reserve(slice * (size_t)world + b_off + b_dev) != NB200_OK ||
        ctx->h_upload.reserve(slice + b_off + b_dev) != NB200_OK ||
        hs->d_hashes.alloc((size_t)hs->dev_total * sizeof(uint32_t), ctx->pool) != NB200_OK ||
        hs->d_ts.alloc((size_t)hs->dev_total * sizeof(uint64_t), ctx->pool) != NB200_OK) {
        delete hs;
        return NB200_ERR_CUDA;
    }
    cudaStream_t s = ctx->stream;
    char *scratch = ctx->d_upload.as<char>();
    char *stage = ctx->h_upload.as<char>();
    cudaError_t e = cudaSuccess;
    // this rank's bytes [lo, hi) of the image: a piece of the timestamps and/or a piece of the hashes
    const size_t lo = std::min(image, slice * (size_t)rank), hi = std::min(image, lo + slice);
    const size_t t_lo = std::min(lo, p_t), t_hi = std::min(hi, p_t), h_lo = std::max(lo, p_t) - p_t, h_hi = std::max(hi, p_t) - p_t;
    const bool direct = is_pinned_host(hashes) && is_pinned_host(ts_ns);
    const char *from_t = reinterpret_cast<const char *>(ts_ns) + t_lo, *from_h = reinterpret_cast<const char *>(hashes) + h_lo;
    if (e == cudaSuccess && !direct) {   // ordinary memory: through the pinned staging area
        if (t_hi > t_lo) memcpy(stage, from_t, t_hi - t_lo);
        if (h_hi > h_lo) memcpy(stage + (t_hi - t_lo), from_h, h_hi - h_lo);
        from_t = stage;
        from_h = stage + (t_hi - t_lo);
    }
    if (e == cudaSuccess && t_hi > t_lo) e = cudaMemcpyAsync(scratch + t_lo, from_t, t_hi - t_lo, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && h_hi > h_lo) e = cudaMemcpyAsync(scratch + p_t + h_lo, from_h, h_hi - h_lo, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        memcpy(stage + slice, hs->seg_offset.data(), b_off);
        memcpy(stage + slice + b_off, hs->dev_offset.data(), b_dev);
        e = cudaMemcpyAsync(scratch + slice * (size_t)world, stage + slice, b_off + b_dev, cudaMemcpyHostToDevice, s);
    }
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_chunk[12], s);
    if (e != cudaSuccess) {
        cudaStreamSynchronize(s);
        set_last_error("hashset_upload_slice", e);
        delete hs;
        return NB200_ERR_CUDA;
    }
    su->d_pack = scratch;
    su->slice_bytes = slice;
    *out = hs;
    return NB200_OK;
}

int hashset_upload_finish(nb200_ctx *ctx, nb200_hashset *hs, const SliceUpload &su, int world) {
    NB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t nseg = 2 * hs->n_videos;
    const size_t p_t = (size_t)hs->total * sizeof(uint64_t);
    const char *tables = su.d_pack + su.slice_bytes * (size_t)world;
    NB_CUDA(cudaMemsetAsync(hs->d_hashes.p, 0, (size_t)hs->dev_total * sizeof(uint32_t), s));
    NB_CUDA(cudaMemsetAsync(hs->d_ts.p, 0, (size_t)hs->dev_total * sizeof(uint64_t), s));
    repack_kernel<<<(unsigned)((hs->total + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<const uint32_t *>(su.d_pack + p_t), reinterpret_cast<const uint64_t *>(su.d_pack),
        hs->d_hashes.as<uint32_t>(), hs->d_ts.as<uint64_t>(), reinterpret_cast<const uint64_t *>(tables),
        reinterpret_cast<const uint32_t *>(tables + (size_t)(nseg + 1) * sizeof(uint64_t)), nseg, hs->total);
    NB_CUDA(cudaGetLastError());
    ctx->n_launches++;
    return NB200_OK;
}

}   // namespace nb200

extern "C" {

int nb200_hashset_upload(nb200_ctx *ctx, const uint32_t *hashes, const uint64_t *ts_ns,
                         const uint64_t *seg_offset, uint32_t n_videos, nb200_hashset **out) {
    return hashset_upload_impl(ctx, hashes, ts_ns, seg_offset, n_videos, true, out);
}

int nb200_hashset_info(const nb200_hashset *hs, uint32_t *n_videos, uint64_t *total,
                       uint64_t *seg_offset_out) {
    if (!hs) return NB200_ERR_NULL_ARGUMENT;
    if (n_videos) *n_videos = hs->n_videos;
    if (total) *total = hs->total;
    if (seg_offset_out) memcpy(seg_offset_out, hs->seg_offset.data(), hs->seg_offset.size() * sizeof(uint64_t));
    return NB200_OK;
}

static int pack_or_unpack(nb200_ctx *ctx, const nb200_hashset *hs, bool to_packed, const void *src_h,
                          const void *src_t, void *dst_h, void *dst_t) {
    // device copies of the two offset tables
    const uint32_t nseg = 2 * hs->n_videos;
    DevBuf d_off, d_dev;
    NB_TRY(d_off.alloc((nseg + 1) * sizeof(uint64_t), ctx->pool));
    NB_TRY(d_dev.alloc(std::max<uint32_t>(nseg, 1) * sizeof(uint32_t), ctx->pool));
    NB_CUDA(cudaMemcpyAsync(d_off.p, hs->seg_offset.data(), (nseg + 1) * sizeof(uint64_t),
                            cudaMemcpyHostToDevice, ctx->stream));
    NB_CUDA(cudaMemcpyAsync(d_dev.p, hs->dev_offset.data(), nseg * sizeof(uint32_t),
                            cudaMemcpyHostToDevice, ctx->stream));
    const uint64_t total = hs->total;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (to_packed)
        unpack_kernel<<<grid, 256, 0, ctx->stream>>>((const uint32_t *)src_h, (const uint64_t *)src_t,
                                                     (uint32_t *)dst_h, (uint64_t *)dst_t,
                                                     d_off.as<uint64_t>(), d_dev.as<uint32_t>(), nseg, total);
    else
        repack_kernel<<<grid, 256, 0, ctx->stream>>>((const uint32_t *)src_h, (const uint64_t *)src_t,
                                                     (uint32_t *)dst_h, (uint64_t *)dst_t,
                                                     d_off.as<uint64_t>(), d_dev.as<uint32_t>(), nseg, total);
    NB_CUDA(cudaGetLastError());
    ctx->n_launches++;
    NB_CUDA(cudaStreamSynchronize(ctx->stream));   // d_off / d_dev die here
    return NB200_OK;
}

int nb200_hashset_download(nb200_ctx *ctx, const nb200_hashset *hs, uint32_t *hashes_out,
                           uint64_t *ts_ns_out) {
    if (!ctx || !hs) return NB200_ERR_NULL_ARGUMENT;
    if (hs->total == 0) return NB200_OK;
    if (!hashes_out || !ts_ns_out) return NB200_ERR_NULL_ARGUMENT;
    NB_CUDA(cudaSetDevice(ctx->device));
    std::vector<uint32_t> th(hs->dev_total);
    std::vector<uint64_t> tt(hs->dev_total);
    NB_CUDA(cudaMemcpyAsync(th.data(), hs->d_hashes.p, hs->dev_total * sizeof(uint32_t),
                            cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaMemcpyAsync(tt.data(), hs->d_ts.p, hs->dev_total * sizeof(uint64_t),
                            cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (uint32_t s = 0; s < 2 * hs->n_videos; s++) {
        memcpy(hashes_out + hs->seg_offset[s], th.data() + hs->dev_offset[s],
               (size_t)hs->seg_len[s] * sizeof(uint32_t));
        memcpy(ts_ns_out + hs->seg_offset[s], tt.data() + hs->dev_offset[s],
               (size_t)hs->seg_len[s] * sizeof(uint64_t));
    }
    return NB200_OK;
}

int nb200_hashset_export_packed(nb200_ctx *ctx, const nb200_hashset *hs, void *d_hashes_out,
                                void *d_ts_ns_out) {
    if (!ctx || !hs) return NB200_ERR_NULL_ARGUMENT;
    if (hs->total == 0) return NB200_OK;
    if (!d_hashes_out || !d_ts_ns_out) return NB200_ERR_NULL_ARGUMENT;
    NB_CUDA(cudaSetDevice(ctx->device));
    return pack_or_unpack(ctx, hs, true, hs->d_hashes.p, hs->d_ts.p, d_hashes_out, d_ts_ns_out);
}

int nb200_hashset_device_ptrs(const nb200_hashset *hs, void **d_hashes, void **d_ts_ns) {
    if (!hs) return NB200_ERR_NULL_ARGUMENT;
    if (d_hashes) *d_hashes = hs->d_hashes.p;
    if (d_ts_ns) *d_ts_ns = hs->d_ts.p;
    return NB200_OK;
}

int nb200_hashset_from_device(nb200_ctx *ctx, const void *d_hashes, const void *d_ts_ns,
                              const uint64_t *seg_offset, uint32_t n_videos, nb200_hashset **out) {
    if (!ctx || !seg_offset || !out) return NB200_ERR_NULL_ARGUMENT;
    *out = nullptr;
    if (seg_offset[0] != 0) return NB200_ERR_INVALID_ARGUMENT;
    NB_CUDA(cudaSetDevice(ctx->device));
    nb200_hashset *hs = new (std::nothrow) nb200_hashset();
    if (!hs) return NB200_ERR_TOO_LARGE;
    int st = layout_hashset(hs, seg_offset, n_videos);
    if (st == NB200_OK && hs->total > 0 && (!d_hashes || !d_ts_ns)) st = NB200_ERR_NULL_ARGUMENT;
    if (st == NB200_OK) st = hs->d_hashes.alloc((size_t)hs->dev_total * sizeof(uint32_t), ctx->pool);
    if (st == NB200_OK) st = hs->d_ts.alloc((size_t)hs->dev_total * sizeof(uint64_t), ctx->pool);
    if (st == NB200_OK && hs->dev_total > 0) {
        cudaError_t e = cudaMemsetAsync(hs->d_hashes.p, 0, hs->dev_total * sizeof(uint32_t), ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(hs->d_ts.p, 0, hs->dev_total * sizeof(uint64_t), ctx->stream);
        if (e != cudaSuccess) {
            set_last_error("cudaMemsetAsync", e);
            st = NB200_ERR_CUDA;
        }
    }
    if (st == NB200_OK && hs->total > 0)
        st = pack_or_unpack(ctx, hs, false, d_hashes, d_ts_ns, hs->d_hashes.p, hs->d_ts.p);
    if (st != NB200_OK) {
        delete hs;
        return st;
    }
    *out = hs;
    return NB200_OK;
}

int nb200_hashset_from_device_scattered(nb200_ctx *ctx, const void *d_hashes, const void *d_ts_ns,
                                        const uint64_t *src_hash_off, const uint64_t *src_ts_off,
                                        const uint64_t *seg_len, uint32_t n_videos, nb200_hashset **out) {
    if (!ctx || !src_hash_off || !src_ts_off || !seg_len || !out) return NB200_ERR_NULL_ARGUMENT;
    *out = nullptr;
    NB_CUDA(cudaSetDevice(ctx->device));
    const uint32_t nseg = 2 * n_videos;
    std::vector<uint64_t> seg_offset(nseg + 1, 0);
    for (uint32_t s = 0; s < nseg; s++) seg_offset[s + 1] = seg_offset[s] + seg_len[s];
    nb200_hashset *hs = new (std::nothrow) nb200_hashset();
    if (!hs) return NB200_ERR_TOO_LARGE;
    int st = layout_hashset(hs, seg_offset.data(), n_videos);
    if (st == NB200_OK && hs->total > 0 && (!d_hashes || !d_ts_ns)) st = NB200_ERR_NULL_ARGUMENT;
    if (st == NB200_OK) st = hs->d_hashes.alloc((size_t)hs->dev_total * sizeof(uint32_t), ctx->pool);
    if (st == NB200_OK) st = hs->d_ts.alloc((size_t)hs->dev_total * sizeof(uint64_t), ctx->pool);
    DevBuf d_tab;   // seg_offset | src_hash_off | src_ts_off | dev_offset
    const size_t b64 = (size_t)(nseg + 1) * sizeof(uint64_t);
    if (st == NB200_OK) st = d_tab.alloc(3 * b64 + (size_t)std::max<uint32_t>(nseg, 1) * sizeof(uint32_t), ctx->pool);
    if (st == NB200_OK && hs->total > 0) {
        cudaStream_t s = ctx->stream;
        char *tab = d_tab.as<char>();
        cudaError_t e = cudaMemsetAsync(hs->d_hashes.p, 0, hs->dev_total * sizeof(uint32_t), s);
        if (e == cudaSuccess) e = cudaMemsetAsync(hs->d_ts.p, 0, hs->dev_total * sizeof(uint64_t), s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(tab, seg_offset.data(), b64, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(tab + b64, src_hash_off, nseg * sizeof(uint64_t), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(tab + 2 * b64, src_ts_off, nseg * sizeof(uint64_t), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(tab + 3 * b64, hs->dev_offset.data(), nseg * sizeof(uint32_t), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) {
            const unsigned grid = (unsigned)((hs->total + 255) / 256);
            scatter_gather_kernel<<<grid, 256, 0, s>>>(
                (const uint32_t *)d_hashes, (const uint64_t *)d_ts_ns, hs->d_hashes.as<uint32_t>(),
                hs->d_ts.as<uint64_t>(), (const uint64_t *)tab, (const uint64_t *)(tab + b64),
                (const uint64_t *)(tab + 2 * b64), (const uint32_t *)(tab + 3 * b64), nseg, hs->total);
            e = cudaGetLastError();
            ctx->n_launches++;
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);   // d_tab and the host tables die here
        if (e != cudaSuccess) {
            set_last_error("nb200_hashset_from_device_scattered", e);
            st = NB200_ERR_CUDA;
        }
    }
    if (st != NB200_OK) {
        delete hs;
        return st;
    }
    *out = hs;
    return NB200_OK;
}

int nb200_hashset_view(nb200_ctx *ctx, const void *d_hashes, const void *d_ts_ns,
                       const uint64_t *dev_offset, const uint64_t *seg_len, uint32_t n_videos,
                       nb200_hashset **out) {
    if (!ctx || !dev_offset || !seg_len || !out) return NB200_ERR_NULL_ARGUMENT;
    *out = nullptr;
    const uint32_t nseg = 2 * n_videos;
    nb200_hashset *hs = new (std::nothrow) nb200_hashset();
    if (!hs) return NB200_ERR_TOO_LARGE;
    hs->n_videos = n_videos;
    hs->seg_offset.assign(nseg + 1, 0);
    hs->dev_offset.resize(nseg);
    hs->seg_len.resize(nseg);
    uint64_t extent = 0;
    for (uint32_t s = 0; s < nseg; s++) {
        if ((dev_offset[s] & 3) || dev_offset[s] + seg_len[s] > 0xfffffff0ull || seg_len[s] > 0x7fffff00ull) {
            delete hs;
            return NB200_ERR_INVALID_ARGUMENT;
        }
        hs->dev_offset[s] = (uint32_t)dev_offset[s];
        hs->seg_len[s] = (uint32_t)seg_len[s];
        hs->seg_offset[s + 1] = hs->seg_offset[s] + seg_len[s];
        extent = std::max<uint64_t>(extent, dev_offset[s] + seg_len[s]);
    }
    hs->total = hs->seg_offset[nseg];
    hs->dev_total = extent;
    if (hs->total > 0 && (!d_hashes || !d_ts_ns)) {
        delete hs;
        return NB200_ERR_NULL_ARGUMENT;
    }
    hs->d_hashes.borrow(const_cast<void *>(d_hashes), extent * sizeof(uint32_t));
    hs->d_ts.borrow(const_cast<void *>(d_ts_ns), extent * sizeof(uint64_t));
    hs->sorted_layout = false;
    hs->immutable = false;
    *out = hs;
    return NB200_OK;
}

void nb200_hashset_free(nb200_hashset *hs) { delete hs; }

/* ------------------------------------------------------------------- match */

int nb200_match_run(nb200_ctx *ctx, const nb200_hashset *hs, const uint32_t (*pairs)[2],
                    uint64_t n_pairs, const nb200_match_params *params, nb200_runset **out) {
    if (!ctx || !hs || !params || !out) return NB200_ERR_NULL_ARGUMENT;
    *out = nullptr;
    return match_run_impl(ctx, hs, pairs, n_pairs, params, out);
}

int nb200_runset_count(const nb200_runset *rs, uint64_t *n_runs, uint64_t *n_cells) {
    if (!rs) return NB200_ERR_NULL_ARGUMENT;
    if (n_runs) *n_runs = rs->n_runs;
    if (n_cells) *n_cells = rs->n_cells;
    return NB200_OK;
}

int nb200_runset_download(nb200_ctx *ctx, const nb200_runset *rs, nb200_run *runs_out) {
    if (!ctx || !rs) return NB200_ERR_NULL_ARGUMENT;
    if (rs->n_runs == 0) return NB200_OK;
    if (!runs_out) return NB200_ERR_NULL_ARGUMENT;
    NB_CUDA(cudaSetDevice(ctx->device));
    PhaseTimer pt(ctx, NB200_PHASE_RUNS_DOWNLOAD);
    // through pinned staging: a pageable destination makes the copy synchronous and slow
    NB_TRY(ctx->h_runs.reserve(rs->n_runs * sizeof(DeviceRun)));
    NB_CUDA(cudaMemcpyAsync(ctx->h_runs.p, rs->d_runs.p, rs->n_runs * sizeof(DeviceRun),
                            cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    // the reference's push order: pair, opening before ending, i desc, j desc (problems are
    // numbered in (pair, is_ending) order).  Bucket by problem, then order each small bucket.
    const DeviceRun *src = ctx->h_runs.as<DeviceRun>();
    const size_t n_problems = rs->problem_pair.size();
    std::vector<uint64_t> bucket(n_problems + 1, 0);
    for (uint64_t k = 0; k < rs->n_runs; k++) {
        if (src[k].problem >= n_problems) return NB200_ERR_STATE;
        bucket[src[k].problem + 1]++;
    }
    for (size_t q = 0; q < n_problems; q++) bucket[q + 1] += bucket[q];
    std::vector<DeviceRun> tmp(rs->n_runs);
    {
        std::vector<uint64_t> cursor(bucket.begin(), bucket.end() - 1);
        for (uint64_t k = 0; k < rs->n_runs; k++) tmp[cursor[src[k].problem]++] = src[k];
    }
    for (size_t q = 0; q < n_problems; q++)
        if (bucket[q + 1] - bucket[q] > 1)
            std::sort(tmp.begin() + (long)bucket[q], tmp.begin() + (long)bucket[q + 1],
                      [](const DeviceRun &a, const DeviceRun &b) {
                          if (a.i_end != b.i_end) return a.i_end > b.i_end;
                          return a.j_end > b.j_end;
                      });
    for (uint64_t k = 0; k < rs->n_runs; k++) {
        const DeviceRun &r = tmp[k];
        nb200_run &o = runs_out[k];
        o.pair = rs->problem_pair[r.problem];
        o.is_ending = rs->problem_is_ending[r.problem];
        o.i_end = r.i_end;
        o.j_end = r.j_end;
        o.len = r.len;
        o.src_simhash = r.src_simhash;
        o.dst_simhash = r.dst_simhash;
        o.reserved = 0;
        o.src_start_ns = r.src_start_ns;
        o.src_end_ns = r.src_end_ns;
        o.dst_start_ns = r.dst_start_ns;
        o.dst_end_ns = r.dst_end_ns;
    }
    return NB200_OK;
}

void nb200_runset_free(nb200_runset *rs) { delete rs; }

int nb200_match_pairs(nb200_ctx *ctx, const uint32_t *hashes, const uint64_t *ts_ns,
                      const uint64_t *seg_offset, uint32_t n_videos, const uint32_t (*pairs)[2],
                      uint64_t n_pairs, const nb200_match_params *params, nb200_run **out_runs,
                      uint64_t *out_n) {
    if (!ctx || !seg_offset || !params || !out_runs || !out_n) return NB200_ERR_NULL_ARGUMENT;
    *out_runs = nullptr;
    *out_n = 0;
    nb200_hashset *hs = nullptr;
    NB_TRY(hashset_upload_impl(ctx, hashes, ts_ns, seg_offset, n_videos, false, &hs));
    nb200_runset *rs = nullptr;
    int st = nb200_match_run(ctx, hs, pairs, n_pairs, params, &rs);
    if (st == NB200_OK) {
        nb200_run *runs = (nb200_run *)malloc(std::max<uint64_t>(rs->n_runs, 1) * sizeof(nb200_run));
        if (!runs) st = NB200_ERR_TOO_LARGE;
        if (st == NB200_OK) st = nb200_runset_download(ctx, rs, runs);
        if (st == NB200_OK) {
            *out_runs = runs;
            *out_n = rs->n_runs;
        } else {
            free(runs);
        }
    }
    nb200_runset_free(rs);
    nb200_hashset_free(hs);
    return st;
}

int nb200_vote(const uint64_t *hash_duration_ns, uint32_t n_videos, const uint32_t (*pairs)[2],
               uint64_t n_pairs, const nb200_match_params *params, const nb200_run *runs,
               uint64_t n_runs, nb200_search_result *results) {
    if (!hash_duration_ns || !params || !results) return NB200_ERR_NULL_ARGUMENT;
    if (n_runs > 0 && !runs) return NB200_ERR_NULL_ARGUMENT;
    return vote_impl(hash_duration_ns, n_videos, pairs, n_pairs, params, runs, n_runs, nullptr, results);
}

int nb200_vote_subset(const uint64_t *hash_duration_ns, uint32_t n_videos, const uint32_t (*pairs)[2],
                      uint64_t n_pairs, const nb200_match_params *params, const nb200_run *runs,
                      uint64_t n_runs, const uint8_t *video_mask, nb200_search_result *results) {
    if (!hash_duration_ns || !params || !results) return NB200_ERR_NULL_ARGUMENT;
    if (n_runs > 0 && !runs) return NB200_ERR_NULL_ARGUMENT;
    return vote_impl(hash_duration_ns, n_videos, pairs, n_pairs, params, runs, n_runs, video_mask, results);
}

int nb200_search(nb200_ctx *ctx, const uint32_t *hashes, const uint64_t *ts_ns,
                 const uint64_t *seg_offset, const uint64_t *hash_duration_ns, uint32_t n_videos,
                 const nb200_match_params *params, nb200_search_result *results) {
    if (!ctx || !seg_offset || !hash_duration_ns || !params || !results) return NB200_ERR_NULL_ARGUMENT;
    if (n_videos < 2) return NB200_ERR_COMPARATOR_MINIMUM_PATHS;   // main.rs:305-315, capi lib.rs:569
    nb200_hashset *hs = nullptr;
    NB_TRY(hashset_upload_impl(ctx, hashes, ts_ns, seg_offset, n_videos, false, &hs));
    const int st = nb200_search_hashset(ctx, hs, hash_duration_ns, params, results);
    nb200_hashset_free(hs);
    return st;
}

/* --------------------------------------------------------------- searching
 * a device-resident season (the tail of nb200_analyze_search and of the
 * multi-GPU flow): match on the GPU, download runs + timestamps, vote. */
int nb200_search_hashset(nb200_ctx *ctx, const nb200_hashset *hs, const uint64_t *hash_duration_ns,
                         const nb200_match_params *params, nb200_search_result *results) {
    if (!ctx || !hs || !hash_duration_ns || !params || !results) return NB200_ERR_NULL_ARGUMENT;
    if (hs->n_videos < 2) return NB200_ERR_COMPARATOR_MINIMUM_PATHS;
    ctx->ev_valid[4] = false;
    if (!ctx->host_vote) {
        // everything stays on the device; ONE synchronisation, then 48 bytes per video come back
        MatchPending mp;
        uint32_t cap_min = 0;
        for (int attempt = 0;; attempt++) {
            NB_TRY(match_enqueue(ctx, hs, nullptr, 0, params, 0, cap_min, &mp));
            {
                PhaseTimer pt(ctx, NB200_PHASE_VOTE);
                NB_TRY(vote_enqueue(ctx, ctx->d_runs.as<DeviceRun>(), ctx->d_counters.as<unsigned int>() + 1, mp.run_cap,
                                    hash_duration_ns, hs->n_videos, nullptr, 0, params));
            }
            PhaseTimer pt(ctx, NB200_PHASE_MATCH);
            uint64_t n_runs = 0;
            NB_TRY(match_collect(ctx, mp, &n_runs));
            if (n_runs <= mp.run_cap) break;
            if (attempt == 1 || n_runs > 0x7ffffff0ull) return NB200_ERR_TOO_LARGE;
            cap_min = (uint32_t)n_runs;
        }
        const size_t b_res = (size_t)hs->n_videos * sizeof(nb200_search_result);
        uint32_t flags[4];
        memcpy(flags, ctx->h_results.as<char>() + b_res, sizeof(flags));
        if (!(flags[0] & VOTE_FLAG_FALLBACK)) {
            memcpy(results, ctx->h_results.p, b_res);
            return (flags[0] & VOTE_FLAG_UNDERFLOW) ? NB200_ERR_DURATION_UNDERFLOW : NB200_OK;
        }
        // a table or a video beyond the device vote's bounds: fall through to the host vote
    }
    nb200_runset *rs = nullptr;
    NB_TRY(nb200_match_run(ctx, hs, nullptr, 0, params, &rs));
    std::vector<nb200_run> runs(std::max<uint64_t>(rs->n_runs, 1));
    int st = nb200_runset_download(ctx, rs, runs.data());
    if (st == NB200_OK) {
        PhaseTimer pt(ctx, NB200_PHASE_VOTE);
        st = vote_impl(hash_duration_ns, hs->n_videos, nullptr, 0, params, runs.data(), rs->n_runs, nullptr, results);
    }
    nb200_runset_free(rs);
    return st;
}

}  // extern "C"
