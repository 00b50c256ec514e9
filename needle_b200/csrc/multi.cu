// multi.cu -- the multi-GPU job behind the C ABI.
//
// needle is ONE process that fans out over videos (Analyzer::run,
// needle/src/audio/analyzer.rs:437-445) and over pairs (Comparator::
// run_with_frame_hashes, needle/src/audio/comparator.rs:549-564) with rayon.  The
// same two fan-outs across the GPUs of a box live HERE, under include/needle_b200.h,
// so that a host which links the library gets all of them, whether it is one process
// driving N devices (needle itself: nb200_comm_init_all = ncclCommInitAll) or one
// process per device (torchrun, MPI: nb200_comm_init_rank from a unique id).
//
//   videos --shard--> K1/K2, hashes written straight into the all-gather send buffer
//   ONE ncclAllGather of the u32 hashes (NCCL over NVLink; loaded with dlopen, so the
//       library has no link-time dependency and shares libnccl.so.2 with the host)
//   pairs  --shard--> K3/K4 read the receive buffer through a hashset view and emit
//       their runs into this rank's run block
//   run blocks --> rank 0 by a push kernel over NVLink peer memory: the USED prefix
//       of the block is stored directly into rank 0's HBM (peer pointer; CUDA IPC
//       between processes), then a system-scope release flag; rank 0's vote waits on
//       the flags of all ranks.  No fixed-size collective, no host round trip.
//   rank 0: as soon as all blocks have landed an acknowledgement word (step, largest run
//       count) is stored into every peer's memory -- what the peers' call returns on, and
//       how every rank learns that a block overflowed and the step must be repeated with
//       more room -- then heap replay + find_best_match on the device, 48 bytes per video
//       to the host.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <numeric>

#include "common.h"

namespace nb200 {

// ------------------------------------------------------------------ NCCL, loaded at run time
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi &nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        // RTLD_NOLOAD first: a host that already has NCCL in the process (PyTorch) must share it
        a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!a.handle) a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle) return a;
#define NB_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name))
        NB_SYM(GetUniqueId, "ncclGetUniqueId");
        NB_SYM(CommInitRank, "ncclCommInitRank");
        NB_SYM(CommInitAll, "ncclCommInitAll");
        NB_SYM(CommDestroy, "ncclCommDestroy");
        NB_SYM(AllGather, "ncclAllGather");
        NB_SYM(GroupStart, "ncclGroupStart");
        NB_SYM(GroupEnd, "ncclGroupEnd");
        NB_SYM(GetErrorString, "ncclGetErrorString");
#undef NB_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllGather && a.GroupStart &&
               a.GroupEnd && a.GetErrorString;
        return a;
    }();
    return api;
}

#define NB_NCCL(expr)                                                                           \
    do {                                                                                        \
        ncclResult_t _r = (expr);                                                               \
        if (_r != ncclSuccess) {                                                                \
            set_last_error_msg(std::string(#expr) + ": " + nccl_api().GetErrorString(_r));      \
            return NB200_ERR_NCCL;                                                              \
        }                                                                                       \
    } while (0)

static int need_nccl() {
    if (nccl_api().ok) return NB200_OK;
    set_last_error_msg("libnccl.so.2 could not be loaded (dlopen): multi-GPU jobs need NCCL");
    return NB200_ERR_NCCL;
}

// ------------------------------------------------------------------ peer-memory tail
// Control words (u32), one block per rank, in that rank's own HBM:
//   [0 .. world)   rank 0 only: flag[r] = the last step whose run block rank r has pushed
//   [64], [65]     every rank: ack info (largest n_found of the step), ack step -- written by rank 0
constexpr uint32_t CTRL_WORDS = 128;
constexpr uint32_t CTRL_ACK_INFO = 64, CTRL_ACK_STEP = 65, CTRL_TIMEOUT = 66;   // [96], [97]: wait_ack's staging

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// The used prefix of this rank's run block (64-byte header + n_stored 64-byte records) into its
// slot of rank 0's receive area -- plain stores to a peer pointer, i.e. NVLink writes -- then the
// flag.  One CTA: a block is a few hundred records.
__global__ void __launch_bounds__(1024) push_block_kernel(const uint4 *__restrict__ block, uint4 *__restrict__ dst,
                                                          uint32_t *flag, uint32_t step, uint32_t cap) {
    const RunBlockHeader *h = reinterpret_cast<const RunBlockHeader *>(block);
    const uint32_t n = min(h->n_stored, cap);
    const uint32_t units = 4u * (1u + n);   // 16-byte units
    for (uint32_t k = threadIdx.x; k < units; k += blockDim.x) dst[k] = block[k];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(flag, step);
}

// rank 0: until every rank's block of this step has landed.  Steps only grow; compared as a
// signed difference so that the counter may wrap.
// Every wait is bounded (~10 s): a rank that never arrives must not hang the GPU; the host
// finds CTRL_TIMEOUT set and fails the call instead.
constexpr uint32_t WAIT_SPINS = 1u << 26;
__global__ void wait_flags_kernel(uint32_t *ctrl, uint32_t world, uint32_t step) {
    if (threadIdx.x < world) {
        uint32_t spins = 0;
        while ((int32_t)(ld_acquire_sys(ctrl + threadIdx.x) - step) < 0) {
            __nanosleep(128);
            if (++spins > WAIT_SPINS) {
                ctrl[CTRL_TIMEOUT] = 1u + threadIdx.x;
                break;
            }
        }
    }
}

// rank 0, as soon as every block has landed: the largest run count of the step and the step number
// into every rank's control block (its own included) -- the peers return on it while rank 0 votes
struct AckTargets {
    uint32_t *ctrl[64];
};
__global__ void __launch_bounds__(64) ack_kernel(const char *__restrict__ recv, uint64_t block_stride, const AckTargets t,
                                                 uint32_t world, uint32_t step) {
    __shared__ uint32_t s_max;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    if (threadIdx.x < world)
        atomicMax(&s_max, reinterpret_cast<const RunBlockHeader *>(recv + (size_t)threadIdx.x * block_stride)->n_found);
    __syncthreads();
    if (threadIdx.x < world) {
        uint32_t *c = t.ctrl[threadIdx.x];
        c[CTRL_ACK_INFO] = min(s_max, 0xfffffff0u);
        __threadfence_system();
        st_release_sys(c + CTRL_ACK_STEP, step);
    }
}

__global__ void wait_ack_kernel(const uint32_t *ctrl, uint32_t step, uint32_t *out) {
    if (threadIdx.x == 0) {
        uint32_t spins = 0;
        bool timeout = false;
        while ((int32_t)(ld_acquire_sys(ctrl + CTRL_ACK_STEP) - step) < 0) {
            __nanosleep(128);
            if (++spins > WAIT_SPINS) {
                timeout = true;
                break;
            }
        }
        out[0] = timeout ? 0xfffffffeu : ctrl[CTRL_ACK_INFO];
        out[1] = ld_acquire_sys(ctrl + CTRL_ACK_STEP);
    }
}

}  // namespace nb200

using namespace nb200;

// One rank's endpoint.
struct nb200_comm {
    nb200_ctx *ctx = nullptr;
    int rank = 0, world = 1;
    ncclComm_t nccl = nullptr;
    bool single_process = false;
    // peer-memory tail
    uint32_t *d_ctrl = nullptr;               // CTRL_WORDS u32, cudaMalloc'd (IPC-exportable)
    char *d_recv = nullptr;                   // rank 0: world * recv_block_bytes
    uint64_t recv_block_bytes = 0;
    uint32_t *peer_flags0 = nullptr;          // rank 0's control block as seen from this rank
    char *peer_recv0 = nullptr;               // rank 0's receive area as seen from this rank
    AckTargets ack_targets = {};              // rank 0: every rank's control block
    std::vector<void *> ipc_opened;           // to close
    std::vector<void *> retired;              // rank 0: receive areas replaced by larger ones
    uint32_t step = 0;
    nb200::PinnedBuf h_ack;
    bool tail_ready = false;
};

namespace nb200 {

constexpr uint64_t RECV_BLOCK_BYTES = 16ull << 20;   // 262,143 runs per rank and step before the job fails over

static int comm_alloc_tail(nb200_comm *c) {
    NB_CUDA(cudaSetDevice(c->ctx->device));
    NB_CUDA(cudaMalloc(&c->d_ctrl, CTRL_WORDS * sizeof(uint32_t)));
    NB_CUDA(cudaMemset(c->d_ctrl, 0, CTRL_WORDS * sizeof(uint32_t)));
    c->recv_block_bytes = RECV_BLOCK_BYTES;
    if (c->rank == 0) NB_CUDA(cudaMalloc(&c->d_recv, (size_t)c->world * c->recv_block_bytes));
    NB_TRY(c->h_ack.reserve(64));
    NB_CUDA(cudaDeviceSynchronize());
    return NB200_OK;
}

}  // namespace nb200

extern "C" {

int nb200_comm_unique_id(uint8_t id[NB200_UNIQUE_ID_BYTES]) {
    if (!id) return NB200_ERR_NULL_ARGUMENT;
    NB_TRY(need_nccl());
    static_assert(sizeof(ncclUniqueId) == NB200_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    NB_NCCL(nccl_api().GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return NB200_OK;
}

void nb200_comm_destroy(nb200_comm *c) {
    if (!c) return;
    if (c->ctx) {
        cudaSetDevice(c->ctx->device);
        cudaStreamSynchronize(c->ctx->stream);
    }
    for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    if (c->nccl && nccl_api().ok) nccl_api().CommDestroy(c->nccl);
    if (c->d_ctrl) cudaFree(c->d_ctrl);
    if (c->d_recv) cudaFree(c->d_recv);
    for (void *p : c->retired) cudaFree(p);
    delete c;
}

int nb200_comm_info(const nb200_comm *c, int *rank, int *world) {
    if (!c) return NB200_ERR_NULL_ARGUMENT;
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return NB200_OK;
}

int nb200_comm_init_rank(nb200_ctx *ctx, const uint8_t id[NB200_UNIQUE_ID_BYTES], int rank, int world,
                         nb200_comm **out) {
    if (!ctx || !out || (world > 1 && !id)) return NB200_ERR_NULL_ARGUMENT;
    *out = nullptr;
    if (world < 1 || world > 64 || rank < 0 || rank >= world) return NB200_ERR_INVALID_ARGUMENT;
    nb200_comm *c = new (std::nothrow) nb200_comm();
    if (!c) return NB200_ERR_TOO_LARGE;
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    ctx->defer_wait = true;   // the job waits once, at its end
    auto fail = [&](int st) {
        nb200_comm_destroy(c);
        return st;
    };
    if (world == 1) {
        *out = c;
        return NB200_OK;
    }
    int st = need_nccl();
    if (st != NB200_OK) return fail(st);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(NB200_ERR_CUDA);
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    {
        ncclResult_t r = nccl_api().CommInitRank(&c->nccl, world, u, rank);
        if (r != ncclSuccess) {
            set_last_error_msg(std::string("ncclCommInitRank: ") + nccl_api().GetErrorString(r));
            return fail(NB200_ERR_NCCL);
        }
    }
    if ((st = comm_alloc_tail(c)) != NB200_OK) return fail(st);
    // exchange the IPC handles of every rank's control block and of rank 0's receive area
    struct Handles {
        cudaIpcMemHandle_t ctrl, recv;
    };
    static_assert(sizeof(Handles) == 128, "two 64-byte IPC handles");
    Handles mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaIpcGetMemHandle(&mine.ctrl, c->d_ctrl) != cudaSuccess ||
        (rank == 0 && cudaIpcGetMemHandle(&mine.recv, c->d_recv) != cudaSuccess)) {
        set_last_error("cudaIpcGetMemHandle", cudaGetLastError());
        return fail(NB200_ERR_CUDA);
    }
    DevBuf d_send, d_all;
    std::vector<Handles> all((size_t)world);
    cudaStream_t s = ctx->stream;
    if (d_send.alloc(sizeof(Handles)) != NB200_OK || d_all.alloc(sizeof(Handles) * world) != NB200_OK)
        return fail(NB200_ERR_CUDA);
    if (cudaMemcpyAsync(d_send.p, &mine, sizeof(mine), cudaMemcpyHostToDevice, s) != cudaSuccess) return fail(NB200_ERR_CUDA);
    {
        ncclResult_t r = nccl_api().AllGather(d_send.p, d_all.p, sizeof(Handles), ncclUint8, c->nccl, s);
        if (r != ncclSuccess) {
            set_last_error_msg(std::string("ncclAllGather(ipc handles): ") + nccl_api().GetErrorString(r));
            return fail(NB200_ERR_NCCL);
        }
    }
    if (cudaMemcpyAsync(all.data(), d_all.p, sizeof(Handles) * world, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
        set_last_error("ipc handle exchange", cudaGetLastError());
        return fail(NB200_ERR_CUDA);
    }
    auto open = [&](const cudaIpcMemHandle_t &h, void **p) {
        cudaError_t e = cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_last_error("cudaIpcOpenMemHandle", e);
            return false;
        }
        c->ipc_opened.push_back(*p);
        return true;
    };
    if (rank == 0) {
        c->peer_flags0 = c->d_ctrl;
        c->peer_recv0 = c->d_recv;
        c->ack_targets.ctrl[0] = c->d_ctrl;
        for (int r = 1; r < world; r++) {
            void *p = nullptr;
            if (!open(all[r].ctrl, &p)) return fail(NB200_ERR_CUDA);
            c->ack_targets.ctrl[r] = (uint32_t *)p;
        }
    } else {
        void *p = nullptr, *q = nullptr;
        if (!open(all[0].ctrl, &p) || !open(all[0].recv, &q)) return fail(NB200_ERR_CUDA);
        c->peer_flags0 = (uint32_t *)p;
        c->peer_recv0 = (char *)q;
    }
    c->tail_ready = true;
    *out = c;
    return NB200_OK;
}

int nb200_comm_init_all(nb200_ctx *const *ctxs, int n, nb200_comm **out) {
    if (!ctxs || !out) return NB200_ERR_NULL_ARGUMENT;
    if (n < 1 || n > 64) return NB200_ERR_INVALID_ARGUMENT;
    for (int r = 0; r < n; r++) {
        out[r] = nullptr;
        if (!ctxs[r]) return NB200_ERR_NULL_ARGUMENT;
    }
    if (n == 1) return nb200_comm_init_rank(ctxs[0], nullptr, 0, 1, out);
    NB_TRY(need_nccl());
    std::vector<int> devs(n);
    for (int r = 0; r < n; r++) devs[r] = ctxs[r]->device;
    std::vector<ncclComm_t> comms(n, nullptr);
    NB_NCCL(nccl_api().CommInitAll(comms.data(), n, devs.data()));
    auto fail = [&](int st) {
        for (int r = 0; r < n; r++) {
            if (out[r]) nb200_comm_destroy(out[r]);
            else if (comms[r]) nccl_api().CommDestroy(comms[r]);
            out[r] = nullptr;
        }
        return st;
    };
    for (int r = 0; r < n; r++) {
        nb200_comm *c = new (std::nothrow) nb200_comm();
        if (!c) return fail(NB200_ERR_TOO_LARGE);
        c->ctx = ctxs[r];
        c->rank = r;
        c->world = n;
        c->nccl = comms[r];
        c->single_process = true;
        ctxs[r]->defer_wait = true;
        out[r] = c;
        int st = comm_alloc_tail(c);
        if (st != NB200_OK) return fail(st);
    }
    // one address space: peer pointers are the pointers themselves once peer access is on
    for (int r = 0; r < n; r++) {
        if (cudaSetDevice(devs[r]) != cudaSuccess) return fail(NB200_ERR_CUDA);
        for (int q = 0; q < n; q++) {
            if (q == r || devs[q] == devs[r]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devs[r], devs[q]);
            if (!can) {
                set_last_error_msg("GPUs of this job cannot access each other's memory (no peer access)");
                return fail(NB200_ERR_CUDA);
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(devs[q], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                set_last_error("cudaDeviceEnablePeerAccess", e);
                return fail(NB200_ERR_CUDA);
            }
            cudaGetLastError();
        }
    }
    for (int r = 0; r < n; r++) {
        out[r]->peer_flags0 = out[0]->d_ctrl;
        out[r]->peer_recv0 = out[0]->d_recv;
        out[0]->ack_targets.ctrl[r] = out[r]->d_ctrl;
        out[r]->tail_ready = true;
    }
    return NB200_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------ jobs
namespace nb200 {

// all (i < j), i-major: the reference's pair order (comparator.rs:534-545)
static void all_pairs(uint32_t n, std::vector<uint32_t> &flat) {
    flat.clear();
    for (uint32_t i = 0; i < n; i++)
        for (uint32_t j = i + 1; j < n; j++) {
            flat.push_back(i);
            flat.push_back(j);
        }
}

// contiguous slices of the pair list, balanced by table cells n*m (+ endings)
static void shard_pairs(const std::vector<uint64_t> &seg_len, const uint32_t *pairs, uint64_t n_pairs, int world,
                        bool endings, std::vector<uint64_t> &cuts) {
    std::vector<double> cum(n_pairs + 1, 0.0);
    for (uint64_t k = 0; k < n_pairs; k++) {
        const uint32_t a = pairs[2 * k], b = pairs[2 * k + 1];
        double c = (double)seg_len[2 * a] * (double)seg_len[2 * b];
        if (endings) c += (double)seg_len[2 * a + 1] * (double)seg_len[2 * b + 1];
        cum[k + 1] = cum[k] + c;
    }
    cuts.assign(world + 1, 0);
    for (int r = 1; r < world; r++) {
        const double target = cum[n_pairs] * r / world;
        cuts[r] = (uint64_t)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        cuts[r] = std::min<uint64_t>(std::max(cuts[r], cuts[r - 1]), n_pairs);
    }
    cuts[world] = n_pairs;
}

// videos -> ranks: longest first onto the least loaded rank, ties to the lower video / rank
// (every rank computes the same plan from the same sample counts)
static void plan_videos(const uint64_t *n_mono, uint32_t n_videos, int world, uint32_t *rank_of_video) {
    std::vector<uint32_t> order(n_videos);
    std::iota(order.begin(), order.end(), 0u);
    auto cost = [&](uint32_t v) { return n_mono[2 * v] + n_mono[2 * v + 1]; };
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cost(a) > cost(b); });
    std::vector<uint64_t> load(world, 0);
    for (uint32_t v : order) {
        int best = 0;
        for (int r = 1; r < world; r++)
            if (load[r] < load[best]) best = r;
        rank_of_video[v] = (uint32_t)best;
        load[best] += cost(v);
    }
}

static uint64_t stored_hashes(uint64_t n_mono, uint32_t stride) {
    const uint64_t raw = nb200_num_raw_hashes(n_mono);
    return (raw + stride - 1) / stride;
}

}  // namespace nb200

enum { MJOB_SEARCH = 0, MJOB_SEASON = 1 };

// per local rank
struct MjobRank {
    nb200_comm *comm = nullptr;
    nb200_hashset *season = nullptr;      // search job: uploaded season; season job: view over the receive buffers
    nb200_pcmset *pcm = nullptr;          // season job: resident PCM of this rank's segments
    nb200::DevBuf d_hash_send, d_ts_send, d_hash_recv, d_ts_recv, d_block;
    std::vector<uint32_t> my_segments;    // season job: global segment ids, in local order
    std::vector<uint64_t> my_seek;
    std::vector<uint64_t> my_n_mono;
    cudaEvent_t ev[6] = {};
    bool ts_filled = false;
};

struct nb200_mjob {
    int kind = MJOB_SEARCH;
    int world = 1;
    uint32_t n_videos = 0;
    uint32_t stride = 2;
    uint64_t hash_duration_ns = 0;
    nb200_match_params params = {};
    std::vector<uint64_t> hash_durations;
    std::vector<uint32_t> pairs;          // flat
    bool default_pairs = true;
    std::vector<uint64_t> cuts;           // pair slices per rank
    std::vector<MjobRank> ranks;          // local ranks, ascending
    // season job: the plan
    std::vector<uint64_t> n_mono, seek_to_ns, seg_len, recv_offset;
    std::vector<uint32_t> video_rank;
    uint64_t cap = 0;                     // elements per rank in the exchange buffers
    uint64_t run_cap = 0;                 // records per run block
    float phase_ms[NB200_MJOB_PHASES] = {};
};

namespace nb200 {

static uint64_t pick_run_cap(const nb200_mjob *j) {
    uint64_t widest = 0;
    for (int r = 0; r < j->world; r++) widest = std::max(widest, j->cuts[r + 1] - j->cuts[r]);
    uint64_t cap = 4096;
    while (cap < 16 * widest) cap *= 2;
    return cap;
}

static int mjob_common(nb200_mjob *j, nb200_comm *const *comms, int n_local, const uint32_t (*pairs)[2],
                       uint64_t n_pairs, const nb200_match_params *params) {
    j->world = comms[0]->world;
    j->params = *params;
    for (int k = 0; k < n_local; k++) {
        if (!comms[k] || comms[k]->world != j->world) return NB200_ERR_INVALID_ARGUMENT;
        if (k && comms[k]->rank <= comms[k - 1]->rank) return NB200_ERR_INVALID_ARGUMENT;   // ascending, distinct
        if (j->world > 1 && !comms[k]->tail_ready) return NB200_ERR_STATE;
    }
    if (pairs) {
        j->pairs.assign(&pairs[0][0], &pairs[0][0] + 2 * n_pairs);
        std::vector<uint32_t> def;
        all_pairs(j->n_videos, def);
        j->default_pairs = def == j->pairs;
    } else {
        all_pairs(j->n_videos, j->pairs);
        j->default_pairs = true;
    }
    for (uint32_t v : j->pairs)
        if (v >= j->n_videos) return NB200_ERR_INVALID_ARGUMENT;
    j->ranks.resize(n_local);
    for (int k = 0; k < n_local; k++) {
        j->ranks[k].comm = comms[k];
        NB_CUDA(cudaSetDevice(comms[k]->ctx->device));
        for (auto &e : j->ranks[k].ev) NB_CUDA(cudaEventCreate(&e));
    }
    return NB200_OK;
}

static int mjob_alloc_blocks(nb200_mjob *j) {
    for (auto &R : j->ranks) {
        NB_CUDA(cudaSetDevice(R.comm->ctx->device));
        NB_TRY(R.d_block.alloc(64 * (1 + j->run_cap), R.comm->ctx->pool));   // a job per search must not pay cudaMalloc + cudaFree
    }
    return NB200_OK;
}

// Rank 0's receive area has one fixed slot per rank; a step whose run blocks outgrow the slots
// (every rank learns the same run count from the acknowledgement) replaces the area.  Collective:
// every rank calls it at the same point with the same size; all streams are idle then.
static int grow_recv(nb200_mjob *j, uint64_t need_bytes) {
    uint64_t bytes = j->ranks[0].comm->recv_block_bytes;
    while (bytes < need_bytes) bytes *= 2;
    if (bytes > (1ull << 30)) {
        set_last_error_msg("a rank's run block exceeds 1 GiB");
        return NB200_ERR_TOO_LARGE;
    }
    const int world = j->world;
    nb200_comm *c0 = nullptr;
    for (auto &R : j->ranks)
        if (R.comm->rank == 0) c0 = R.comm;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (c0) {
        NB_CUDA(cudaSetDevice(c0->ctx->device));
        NB_CUDA(cudaStreamSynchronize(c0->ctx->stream));
        // the old area stays allocated until the comm goes: other processes may still have it mapped
        c0->retired.push_back(c0->d_recv);
        c0->d_recv = nullptr;
        NB_CUDA(cudaMalloc(&c0->d_recv, (size_t)world * bytes));
        c0->peer_recv0 = c0->d_recv;
    }
    if (j->ranks[0].comm->single_process) {
        for (auto &R : j->ranks) {
            R.comm->peer_recv0 = c0->d_recv;
            R.comm->recv_block_bytes = bytes;
        }
        return NB200_OK;
    }
    // one process per rank: the new area's IPC handle travels the way the first one did
    nb200_comm *c = j->ranks[0].comm;
    nb200_ctx *ctx = c->ctx;
    NB_CUDA(cudaSetDevice(ctx->device));
    if (c->rank == 0) NB_CUDA(cudaIpcGetMemHandle(&mine, c->d_recv));
    DevBuf d_send, d_all;
    std::vector<cudaIpcMemHandle_t> all((size_t)world);
    NB_TRY(d_send.alloc(sizeof(mine)));
    NB_TRY(d_all.alloc(sizeof(mine) * world));
    NB_CUDA(cudaMemcpyAsync(d_send.p, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
    NB_NCCL(nccl_api().AllGather(d_send.p, d_all.p, sizeof(mine), ncclUint8, c->nccl, ctx->stream));
    NB_CUDA(cudaMemcpyAsync(all.data(), d_all.p, sizeof(mine) * world, cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (c->rank != 0) {
        // the old mapping is the last one this rank opened for rank 0's receive area
        for (size_t k = 0; k < c->ipc_opened.size(); k++)
            if (c->ipc_opened[k] == c->peer_recv0) {
                cudaIpcCloseMemHandle(c->peer_recv0);
                c->ipc_opened.erase(c->ipc_opened.begin() + (long)k);
                break;
            }
        void *q = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&q, all[0], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_last_error("cudaIpcOpenMemHandle(grown receive area)", e);
            return NB200_ERR_CUDA;
        }
        c->ipc_opened.push_back(q);
        c->peer_recv0 = (char *)q;
    }
    c->recv_block_bytes = bytes;
    return NB200_OK;
}

// The part after "every rank holds the whole season": match this rank's slice, push, vote on rank 0.
// Phases over the local ranks so that ONE host thread can drive all of them (single-process mode):
// nothing that rank 0 waits for may sit behind a host wait of another local rank.
static int mjob_tail(nb200_mjob *j, nb200_search_result *results) {
    const int world = j->world;
    const uint32_t(*pairs)[2] = reinterpret_cast<const uint32_t(*)[2]>(j->pairs.data());
    for (int attempt = 0; attempt < 8; attempt++) {
        const uint64_t block_bytes = 64 * (1 + j->run_cap);
        if (world > 1 && block_bytes > j->ranks[0].comm->recv_block_bytes) NB_TRY(grow_recv(j, block_bytes));
        // ---- A: every local rank matches its slice into its block and pushes it to rank 0
        for (auto &R : j->ranks) {
            nb200_comm *c = R.comm;
            nb200_ctx *ctx = c->ctx;
            NB_CUDA(cudaSetDevice(ctx->device));
            const uint64_t a = j->cuts[c->rank], b = j->cuts[c->rank + 1];
            NB_TRY(nb200_match_export(ctx, R.season, pairs + a, b - a, &j->params, (uint32_t)a, R.d_block.p, block_bytes));
            NB_CUDA(cudaEventRecord(R.ev[3], ctx->stream));
            if (world > 1) {
                c->step++;
                push_block_kernel<<<1, 1024, 0, ctx->stream>>>(
                    R.d_block.as<uint4>(), reinterpret_cast<uint4 *>(c->peer_recv0 + (size_t)c->rank * c->recv_block_bytes),
                    c->peer_flags0 + c->rank, c->step, (uint32_t)j->run_cap);
                NB_CUDA(cudaGetLastError());
                ctx->n_launches++;
            }
        }
        // ---- B: rank 0 (if it is ours) waits for all blocks, votes, acknowledges
        uint64_t max_found = 0;
        int st0 = NB200_OK;
        bool have0 = false;
        for (auto &R : j->ranks) {
            nb200_comm *c = R.comm;
            if (c->rank != 0) continue;
            have0 = true;
            nb200_ctx *ctx = c->ctx;
            NB_CUDA(cudaSetDevice(ctx->device));
            const void *blocks = R.d_block.p;
            uint64_t stride_bytes = block_bytes;
            if (world > 1) {
                wait_flags_kernel<<<1, 64, 0, ctx->stream>>>(c->d_ctrl, (uint32_t)world, c->step);
                NB_CUDA(cudaGetLastError());
                ctx->n_launches++;
                NB_CUDA(cudaMemcpyAsync(c->h_ack.as<uint32_t>() + 4, c->d_ctrl + CTRL_TIMEOUT, 4, cudaMemcpyDeviceToHost, ctx->stream));
                blocks = c->d_recv;
                stride_bytes = c->recv_block_bytes;
            }
            NB_CUDA(cudaEventRecord(R.ev[4], ctx->stream));
            // the acknowledgement goes out once the blocks have been compacted out of the receive area
            // (the peers may push their next step into it from then on), while the vote is still running
            auto release_peers = [](void *arg) -> int {
                nb200_comm *c0 = static_cast<nb200_comm *>(arg);
                ack_kernel<<<1, 64, 0, c0->ctx->stream>>>(c0->d_recv, c0->recv_block_bytes, c0->ack_targets,
                                                          (uint32_t)c0->world, c0->step);
                NB_CUDA(cudaGetLastError());
                c0->ctx->n_launches++;
                return NB200_OK;
            };
            st0 = vote_blocks_impl(ctx, blocks, (uint32_t)world, stride_bytes, j->run_cap, j->hash_durations.data(),
                                   j->n_videos, j->default_pairs ? nullptr : pairs,
                                   j->default_pairs ? 0 : j->pairs.size() / 2, &j->params, 1, results, &max_found,
                                   world > 1 ? +release_peers : nullptr, c);
            NB_CUDA(cudaEventRecord(R.ev[5], ctx->stream));
            NB_CUDA(cudaStreamSynchronize(ctx->stream));
            if (world > 1) {
                const uint32_t timed_out = c->h_ack.as<uint32_t>()[4];
                if (timed_out) {
                    NB_CUDA(cudaMemset(c->d_ctrl + CTRL_TIMEOUT, 0, 4));
                    set_last_error_msg("rank " + std::to_string(timed_out - 1) + " did not deliver its run block in time");
                    return NB200_ERR_STATE;
                }
            }
        }
        // ---- C: the other local ranks return when rank 0 has acknowledged the step
        int st_peer = NB200_OK;
        for (auto &R : j->ranks) {
            nb200_comm *c = R.comm;
            if (c->rank == 0 || world == 1) continue;
            nb200_ctx *ctx = c->ctx;
            NB_CUDA(cudaSetDevice(ctx->device));
            uint32_t *h = c->h_ack.as<uint32_t>();
            wait_ack_kernel<<<1, 32, 0, ctx->stream>>>(c->d_ctrl, c->step, reinterpret_cast<uint32_t *>(c->d_ctrl + 96));
            NB_CUDA(cudaGetLastError());
            ctx->n_launches++;
            NB_CUDA(cudaMemcpyAsync(h, c->d_ctrl + 96, 8, cudaMemcpyDeviceToHost, ctx->stream));
            NB_CUDA(cudaStreamSynchronize(ctx->stream));
            if (h[0] == 0xfffffffeu) {
                set_last_error_msg("rank 0 did not acknowledge the step in time");
                st_peer = NB200_ERR_STATE;
            }
            else if (!have0) max_found = std::max<uint64_t>(max_found, h[0]);
        }
        if (have0 && st0 != NB200_OK && st0 != NB200_ERR_TOO_LARGE) return st0;
        if (st_peer != NB200_OK) return st_peer;
        if (max_found <= j->run_cap) return NB200_OK;
        // a block overflowed: every rank saw the same count and repeats the step with room for it
        while (j->run_cap < max_found) j->run_cap *= 2;
        NB_TRY(mjob_alloc_blocks(j));
    }
    return NB200_ERR_TOO_LARGE;
}

static void mjob_phase_times(nb200_mjob *j) {
    // CUDA-event times of the lowest local rank's stream
    MjobRank &R = j->ranks[0];
    float *ms = j->phase_ms;
    for (int k = 0; k < NB200_MJOB_PHASES; k++) ms[k] = 0.f;
    cudaSetDevice(R.comm->ctx->device);
    auto span = [&](int a, int b) {
        float v = 0.f;
        if (cudaEventElapsedTime(&v, R.ev[a], R.ev[b]) != cudaSuccess) {
            cudaGetLastError();
            return 0.f;
        }
        return v;
    };
    if (j->kind == MJOB_SEASON) {
        ms[NB200_MJOB_FINGERPRINT] = span(0, 1);
        ms[NB200_MJOB_HASH_ALLGATHER] = span(1, 2);
        ms[NB200_MJOB_MATCH] = span(2, 3);
    } else {
        ms[NB200_MJOB_MATCH] = span(2, 3);
    }
    if (R.comm->rank == 0) {
        ms[NB200_MJOB_RUN_PUSH_WAIT] = span(3, 4);
        ms[NB200_MJOB_VOTE] = span(4, 5);
    }
}

}  // namespace nb200

extern "C" {

void nb200_mjob_free(nb200_mjob *j) {
    if (!j) return;
    for (auto &R : j->ranks) {
        if (R.comm && R.comm->ctx) {
            cudaSetDevice(R.comm->ctx->device);
            cudaStreamSynchronize(R.comm->ctx->stream);
        }
        nb200_hashset_free(R.season);
        nb200_pcmset_free(R.pcm);
        for (auto &e : R.ev)
            if (e) cudaEventDestroy(e);
    }
    delete j;
}

int nb200_mjob_search_create(nb200_comm *const *comms, int n_local, const uint32_t *hashes, const uint64_t *ts_ns,
                             const uint64_t *seg_offset, uint32_t n_videos, const uint64_t *hash_duration_ns,
                             const uint32_t (*pairs)[2], uint64_t n_pairs, const nb200_match_params *params,
                             nb200_mjob **out) {
    if (!comms || !hashes || !ts_ns || !seg_offset || !hash_duration_ns || !params || !out) return NB200_ERR_NULL_ARGUMENT;
    *out = nullptr;
    if (n_local < 1) return NB200_ERR_INVALID_ARGUMENT;
    if (n_videos < 2) return NB200_ERR_COMPARATOR_MINIMUM_PATHS;
    nb200_mjob *j = new (std::nothrow) nb200_mjob();
    if (!j) return NB200_ERR_TOO_LARGE;
    j->kind = MJOB_SEARCH;
    j->n_videos = n_videos;
    j->hash_durations.assign(hash_duration_ns, hash_duration_ns + n_videos);
    int st = mjob_common(j, comms, n_local, pairs, n_pairs, params);
    if (st == NB200_OK) {
        j->seg_len.resize(2 * (size_t)n_videos);
        for (uint32_t s = 0; s < 2 * n_videos; s++) j->seg_len[s] = seg_offset[s + 1] - seg_offset[s];
        shard_pairs(j->seg_len, j->pairs.data(), j->pairs.size() / 2, j->world, params->include_endings != 0, j->cuts);
        j->run_cap = pick_run_cap(j);
        st = mjob_alloc_blocks(j);
    }
    // Every rank holds the whole season (KB..MB; BASELINE configs[3]: 10 MB), but no rank copies all of it
    // from the host: each takes 1/world of the packed image over PCIe and the ranks all-gather it in place
    // over NVLink -- host staging and H2D traffic per rank fall by `world`.
    if (st == NB200_OK && j->world > 1 && seg_offset[2 * (size_t)n_videos] > 0) {
        st = need_nccl();
        std::vector<SliceUpload> su(j->ranks.size());
        for (size_t k = 0; st == NB200_OK && k < j->ranks.size(); k++) {
            nb200_comm *c = j->ranks[k].comm;
            st = hashset_upload_slice(c->ctx, hashes, ts_ns, seg_offset, n_videos, c->rank, j->world, &su[k], &j->ranks[k].season);
        }
        if (st == NB200_OK) {
            ncclResult_t r = nccl_api().GroupStart();
            for (size_t k = 0; r == ncclSuccess && k < j->ranks.size(); k++) {
                nb200_comm *c = j->ranks[k].comm;
                r = nccl_api().AllGather(su[k].d_pack + su[k].slice_bytes * (size_t)c->rank, su[k].d_pack, su[k].slice_bytes,
                                         ncclUint8, c->nccl, c->ctx->stream);
            }
            const ncclResult_t r2 = nccl_api().GroupEnd();
            if (r != ncclSuccess || r2 != ncclSuccess) {
                set_last_error_msg(std::string("ncclAllGather(season image): ") + nccl_api().GetErrorString(r != ncclSuccess ? r : r2));
                st = NB200_ERR_NCCL;
            }
        }
        for (size_t k = 0; st == NB200_OK && k < j->ranks.size(); k++)
            st = hashset_upload_finish(j->ranks[k].comm->ctx, j->ranks[k].season, su[k], j->world);
        // the caller's arrays and the staging areas are free again when this returns
        for (size_t k = 0; st == NB200_OK && k < j->ranks.size(); k++) {
            if (cudaSetDevice(j->ranks[k].comm->ctx->device) != cudaSuccess ||
                cudaStreamSynchronize(j->ranks[k].comm->ctx->stream) != cudaSuccess) {
                set_last_error("nb200_mjob_search_create", cudaGetLastError());
                st = NB200_ERR_CUDA;
            }
        }
    } else {
        for (size_t k = 0; st == NB200_OK && k < j->ranks.size(); k++)
            st = nb200_hashset_upload(j->ranks[k].comm->ctx, hashes, ts_ns, seg_offset, n_videos, &j->ranks[k].season);
    }
    if (st != NB200_OK) {
        nb200_mjob_free(j);
        return st;
    }
    *out = j;
    return NB200_OK;
}

int nb200_mjob_season_create(nb200_comm *const *comms, int n_local, const uint64_t *n_mono_samples,
                             const uint64_t *seek_to_ns, uint32_t n_videos, uint64_t hash_duration_ns,
                             const uint32_t (*pairs)[2], uint64_t n_pairs, const nb200_match_params *params,
                             nb200_mjob **out) {
    if (!comms || !n_mono_samples || !params || !out) return NB200_ERR_NULL_ARGUMENT;
    *out = nullptr;
    if (n_local < 1) return NB200_ERR_INVALID_ARGUMENT;
    if (n_videos < 2) return NB200_ERR_COMPARATOR_MINIMUM_PATHS;
    const uint64_t stride = (hash_duration_ns / 1000000ull) / NB200_ITEM_DURATION_MS;   // analyzer.rs:293-297
    if (stride == 0 || stride > 0xffffffffull) return NB200_ERR_INVALID_ARGUMENT;
    nb200_mjob *j = new (std::nothrow) nb200_mjob();
    if (!j) return NB200_ERR_TOO_LARGE;
    j->kind = MJOB_SEASON;
    j->n_videos = n_videos;
    j->stride = (uint32_t)stride;
    j->hash_duration_ns = hash_duration_ns;
    j->hash_durations.assign(n_videos, hash_duration_ns);
    int st = mjob_common(j, comms, n_local, pairs, n_pairs, params);
    auto fail = [&](int s) {
        nb200_mjob_free(j);
        return s;
    };
    if (st != NB200_OK) return fail(st);
    const int world = j->world;
    const size_t nseg = 2 * (size_t)n_videos;
    j->n_mono.assign(n_mono_samples, n_mono_samples + nseg);
    j->seek_to_ns.assign(nseg, 0);
    if (seek_to_ns) j->seek_to_ns.assign(seek_to_ns, seek_to_ns + nseg);
    j->seg_len.resize(nseg);
    for (size_t s = 0; s < nseg; s++) j->seg_len[s] = stored_hashes(j->n_mono[s], j->stride);
    j->video_rank.assign(n_videos, 0);
    plan_videos(j->n_mono.data(), n_videos, world, j->video_rank.data());
    // each rank's block of the exchange buffers is laid out like a hashset of its own segments
    std::vector<std::vector<uint32_t>> segs_of(world);
    for (uint32_t v = 0; v < n_videos; v++) {
        segs_of[j->video_rank[v]].push_back(2 * v);
        segs_of[j->video_rank[v]].push_back(2 * v + 1);
    }
    j->recv_offset.assign(nseg, 0);
    std::vector<std::vector<uint64_t>> local_off(world);
    j->cap = 4;
    for (int r = 0; r < world; r++) {
        uint64_t pos = 0;
        for (uint32_t s : segs_of[r]) {
            local_off[r].push_back(pos);
            pos += (j->seg_len[s] + 3) & ~3ull;
        }
        j->cap = std::max(j->cap, pos);
    }
    for (int r = 0; r < world; r++)
        for (size_t k = 0; k < segs_of[r].size(); k++) j->recv_offset[segs_of[r][k]] = (uint64_t)r * j->cap + local_off[r][k];
    if (j->cap * (uint64_t)world > 0xfffffff0ull) return fail(NB200_ERR_TOO_LARGE);
    shard_pairs(j->seg_len, j->pairs.data(), j->pairs.size() / 2, world, params->include_endings != 0, j->cuts);
    j->run_cap = pick_run_cap(j);
    if ((st = mjob_alloc_blocks(j)) != NB200_OK) return fail(st);
    for (auto &R : j->ranks) {
        nb200_ctx *ctx = R.comm->ctx;
        const int r = R.comm->rank;
        if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(NB200_ERR_CUDA);
        R.my_segments = segs_of[r];
        for (uint32_t s : R.my_segments) {
            R.my_seek.push_back(j->seek_to_ns[s]);
            R.my_n_mono.push_back(j->n_mono[s]);
        }
        const size_t total = (size_t)j->cap * world;
        const auto &pool = R.comm->ctx->pool;
        if (R.d_hash_send.alloc(j->cap * 4, pool) != NB200_OK || R.d_ts_send.alloc(j->cap * 8, pool) != NB200_OK ||
            R.d_hash_recv.alloc(total * 4, pool) != NB200_OK || R.d_ts_recv.alloc(total * 8, pool) != NB200_OK)
            return fail(NB200_ERR_CUDA);
        if (cudaMemsetAsync(R.d_hash_send.p, 0, j->cap * 4, ctx->stream) != cudaSuccess ||
            cudaMemsetAsync(R.d_hash_recv.p, 0, total * 4, ctx->stream) != cudaSuccess)
            return fail(NB200_ERR_CUDA);
        // world == 1: the send buffers ARE the season
        void *h = world > 1 ? R.d_hash_recv.p : R.d_hash_send.p, *t = world > 1 ? R.d_ts_recv.p : R.d_ts_send.p;
        st = nb200_hashset_view(ctx, h, t, j->recv_offset.data(), j->seg_len.data(), n_videos, &R.season);
        if (st != NB200_OK) return fail(st);
    }
    *out = j;
    return NB200_OK;
}

int nb200_mjob_video_rank(const nb200_mjob *j, uint32_t *rank_of_video) {
    if (!j || !rank_of_video) return NB200_ERR_NULL_ARGUMENT;
    if (j->kind != MJOB_SEASON) return NB200_ERR_STATE;
    memcpy(rank_of_video, j->video_rank.data(), sizeof(uint32_t) * j->n_videos);
    return NB200_OK;
}

int nb200_mjob_upload_pcm(nb200_mjob *j, const int16_t *const *pcm) {
    if (!j || !pcm) return NB200_ERR_NULL_ARGUMENT;
    if (j->kind != MJOB_SEASON) return NB200_ERR_STATE;
    for (auto &R : j->ranks) {
        nb200_pcmset_free(R.pcm);
        R.pcm = nullptr;
        std::vector<const int16_t *> p;
        for (uint32_t s : R.my_segments) {
            if (!pcm[s] && j->n_mono[s]) return NB200_ERR_NULL_ARGUMENT;
            p.push_back(pcm[s]);
        }
        NB_TRY(nb200_pcmset_upload(R.comm->ctx, p.data(), R.my_n_mono.data(), 1, (uint32_t)p.size(), &R.pcm));
    }
    return NB200_OK;
}

int nb200_mjob_run(nb200_mjob *j, const int16_t *const *host_pcm, nb200_search_result *results) {
    if (!j) return NB200_ERR_NULL_ARGUMENT;
    bool have0 = false;
    for (auto &R : j->ranks) have0 |= R.comm->rank == 0;
    if (have0 && !results) return NB200_ERR_NULL_ARGUMENT;
    const int world = j->world;
    if (j->kind == MJOB_SEASON) {
        // ---- fingerprint this rank's segments straight into its send buffer
        for (auto &R : j->ranks) {
            nb200_ctx *ctx = R.comm->ctx;
            NB_CUDA(cudaSetDevice(ctx->device));
            NB_CUDA(cudaEventRecord(R.ev[0], ctx->stream));
            if (R.my_segments.empty()) {
                // more ranks than videos: this one only matches
            } else if (host_pcm) {
                std::vector<const int16_t *> p;
                for (uint32_t s : R.my_segments) {
                    if (!host_pcm[s] && j->n_mono[s]) return NB200_ERR_NULL_ARGUMENT;
                    p.push_back(host_pcm[s]);
                }
                NB_TRY(nb200_fingerprint_host_into(ctx, p.data(), R.my_n_mono.data(), (uint32_t)p.size(), j->stride,
                                                   NB200_DELAY_MS * 1000000ull, NB200_ITEM_DURATION_MS * 1000000ull,
                                                   R.my_seek.data(), R.d_hash_send.p, R.d_ts_send.p, j->cap));
            } else {
                if (!R.pcm) return NB200_ERR_STATE;
                NB_TRY(nb200_fingerprint_run_into(ctx, R.pcm, j->stride, NB200_DELAY_MS * 1000000ull,
                                                  NB200_ITEM_DURATION_MS * 1000000ull, R.my_seek.data(),
                                                  R.d_hash_send.p, R.d_ts_send.p, j->cap));
            }
            NB_CUDA(cudaEventRecord(R.ev[1], ctx->stream));
            // timestamps are a function of the index (analyzer.rs:309): every rank fills in all of them
            // itself, once -- the layout of a job never changes
            if (world > 1 && !R.ts_filled) {
                NB_TRY(nb200_timestamps_fill(ctx, R.d_ts_recv.p, j->recv_offset.data(), j->seg_len.data(),
                                             j->seek_to_ns.data(), 2 * j->n_videos, j->stride,
                                             NB200_DELAY_MS * 1000000ull, NB200_ITEM_DURATION_MS * 1000000ull));
                R.ts_filled = true;
            }
        }
        // ---- ONE all-gather: the hashes
        if (world > 1) {
            NB_NCCL(nccl_api().GroupStart());
            for (auto &R : j->ranks) {
                ncclResult_t r = nccl_api().AllGather(R.d_hash_send.p, R.d_hash_recv.p, j->cap, ncclUint32, R.comm->nccl,
                                                      R.comm->ctx->stream);
                if (r != ncclSuccess) {
                    nccl_api().GroupEnd();
                    set_last_error_msg(std::string("ncclAllGather(hashes): ") + nccl_api().GetErrorString(r));
                    return NB200_ERR_NCCL;
                }
            }
            NB_NCCL(nccl_api().GroupEnd());
        }
    }
    for (auto &R : j->ranks) {
        NB_CUDA(cudaSetDevice(R.comm->ctx->device));
        NB_CUDA(cudaEventRecord(R.ev[2], R.comm->ctx->stream));
    }
    int st = mjob_tail(j, results);
    if (st == NB200_OK) mjob_phase_times(j);
    return st;
}

int nb200_mjob_phase_ms(const nb200_mjob *j, float *ms) {
    if (!j || !ms) return NB200_ERR_NULL_ARGUMENT;
    memcpy(ms, j->phase_ms, sizeof(j->phase_ms));
    return NB200_OK;
}

/* Host-only: the sharding rules of the jobs above, for hosts that want to know them in advance
 * (which PCM to decode where) and for tests. */
int nb200_plan_videos(const uint64_t *n_mono_samples, uint32_t n_videos, int world, uint32_t *rank_of_video) {
    if ((!n_mono_samples || !rank_of_video) && n_videos) return NB200_ERR_NULL_ARGUMENT;
    if (world < 1) return NB200_ERR_INVALID_ARGUMENT;
    plan_videos(n_mono_samples, n_videos, world, rank_of_video);
    return NB200_OK;
}

int nb200_plan_pairs(const uint64_t *seg_len, uint32_t n_videos, const uint32_t (*pairs)[2], uint64_t n_pairs, int world,
                     int include_endings, uint64_t *cuts) {
    if (!seg_len || !cuts) return NB200_ERR_NULL_ARGUMENT;
    if (world < 1) return NB200_ERR_INVALID_ARGUMENT;
    std::vector<uint32_t> flat;
    if (pairs) flat.assign(&pairs[0][0], &pairs[0][0] + 2 * n_pairs);
    else all_pairs(n_videos, flat);
    for (uint32_t v : flat)
        if (v >= n_videos) return NB200_ERR_INVALID_ARGUMENT;
    std::vector<uint64_t> len(seg_len, seg_len + 2 * (size_t)n_videos), c;
    shard_pairs(len, flat.data(), flat.size() / 2, world, include_endings != 0, c);
    memcpy(cuts, c.data(), sizeof(uint64_t) * (world + 1));
    return NB200_OK;
}

}  // extern "C"
