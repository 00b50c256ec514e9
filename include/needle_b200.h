/*
 * needle_b200.h -- C ABI of the B200-native fingerprint-and-match path for
 * aksiksi/needle.  Plain pointers and sizes only; every entry point returns an
 * int status (NB200_OK == 0), never throws or aborts across the boundary.
 *
 * What each group replaces in the reference (paths relative to the needle repo):
 *
 *   B1 fingerprint  nb200_fp_*            the chromaprint_rust::Context calls made by
 *                                         Analyzer::process_frames
 *                                         (needle/src/audio/analyzer.rs:176,179,218,275,
 *                                          286,288,289,299-301)
 *                   nb200_fingerprint_*   the per-video loop of Analyzer::run
 *                                         (analyzer.rs:437-445) + the subsample/timestamp
 *                                         tail of process_frames (analyzer.rs:288-318)
 *   B2 match        nb200_match_*         Comparator::longest_common_hash_match for every
 *                                         pair (needle/src/audio/comparator.rs:157-250,
 *                                          pair loop :532-578)
 *                   nb200_search          Comparator::run_with_frame_hashes
 *                                         (comparator.rs:524-629) without display / skip files
 *   persistence     nb200_framehashes_*   FrameHashes bincode (.needle.dat,
 *                                         needle/src/audio/data.rs:15-26,104-115;
 *                                         analyzer.rs:413-417)
 *
 * There is no CPU fallback: every compute entry point needs a CUDA device
 * (sm_100a) and fails with NB200_ERR_CUDA otherwise.
 *
 * Threading: a context owns one CUDA stream and scratch buffers; calls on the
 * same context must not overlap, distinct contexts are independent (needle's
 * rayon workers would each own one, or share one behind a mutex).
 */
#ifndef NEEDLE_B200_H
#define NEEDLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status */
enum {
    NB200_OK = 0,
    NB200_ERR_NULL_ARGUMENT = 1,     /* NeedleError_NullArgument */
    NB200_ERR_INVALID_ARGUMENT = 2,  /* NeedleError_InvalidArgument */
    NB200_ERR_CUDA = 3,              /* any CUDA runtime failure / no device */
    NB200_ERR_NO_ENDING = 4,         /* Error::FrameHashDataNoEnding (comparator.rs:271-273) */
    NB200_ERR_DURATION_UNDERFLOW = 5,/* the reference would panic on Duration - Duration */
    NB200_ERR_TOO_LARGE = 6,         /* an index does not fit the device layout */
    NB200_ERR_IO = 7,                /* NeedleError_IOError */
    NB200_ERR_FORMAT = 8,            /* NeedleError_InvalidFrameHashData / InvalidVersion */
    NB200_ERR_STATE = 9,             /* call order violated (e.g. feed before start) */
    NB200_ERR_COMPARATOR_MINIMUM_PATHS = 10, /* NeedleError_ComparatorMinimumPaths */
    NB200_ERR_NCCL = 11,             /* NCCL missing (dlopen) or a collective failed; see nb200_last_error */
};
const char *nb200_status_str(int status);
/* Last CUDA error text seen by this thread ("" if none). */
const char *nb200_last_error(void);

/* ----------------------------------------------------------------- context */
typedef struct nb200_ctx nb200_ctx;
typedef struct nb200_hashset nb200_hashset; /* a season of hashes + timestamps in HBM */
typedef struct nb200_pcmset nb200_pcmset;   /* decoded PCM segments in HBM */
typedef struct nb200_runset nb200_runset;   /* match output in HBM */
/* device < 0: use the calling thread's current CUDA device. */
int nb200_ctx_create(int device, nb200_ctx **out);
void nb200_ctx_destroy(nb200_ctx *ctx);
/* Run on a caller-owned cudaStream_t instead of the context's own stream
 * (NULL restores it).  Lets a host time the kernels with its own events. */
int nb200_ctx_set_stream(nb200_ctx *ctx, void *cuda_stream);
int nb200_ctx_synchronize(nb200_ctx *ctx);
/* Milliseconds of the kernels of the last *_run call, from CUDA events on the
 * context's stream: [0] fingerprint K1 (fft+chroma), [1] K2 (classify),
 * [2] match K3, [3] simhash K4.  Counters of launches since context creation
 * in n_launches (may be NULL). */
int nb200_ctx_last_kernel_ms(nb200_ctx *ctx, float ms[4], uint64_t *n_launches);
/* Milliseconds of the device vote kernels (group, heap replay, find_best_match) of the
 * last search; 0 when the host voted. */
int nb200_ctx_last_vote_ms(nb200_ctx *ctx, float *ms);

/* Tuning / test switches.  NB200_OPT_FORCE_GENERAL_MATCH (value 0|1): always run
 * the general match kernel, also where the fast variant (all candidate run
 * lengths >= 63 cells) would be chosen.  Results are identical either way. */
enum {
    NB200_OPT_FORCE_GENERAL_MATCH = 1,
    /* NB200_OPT_K1_VARIANT: which fingerprint FFT kernel runs (all: one warp per 4096-sample frame, at
     * most 32 complex values per lane; same hashes up to FP32 rounding order, >= 99.99 % identical
     * frames; a tuning switch).  0 = the default (19).  16..19 = tensor-memory kernel, 16 warps per SM:
     * the parked half of a frame lives in TMEM (tcgen05.st / tcgen05.ld); 17 adds FMA-fused butterflies
     * and the conversion unit, 18 fetches odd-aligned frames rotated by one sample (aligned words), 19
     * keeps the pass-2 and split twiddles in TMEM as well (one slot per warp); 112 = 16 with 12 warps;
     * 8, 10, 12 = the parked half in shared memory, that many warps per SM; 1 = all 64 values of a
     * lane in registers (8 warps per SM). */
    NB200_OPT_K1_VARIANT = 2,
    /* NB200_OPT_MATCH_DENSE (0|1): the fast match kernel normally tests 4 sampled rows of
     * every 32-row word and verifies the words that pass; 1 makes it evaluate all 32 cells
     * of every word (the POPC-roofline kernel).  Results are identical. */
    NB200_OPT_MATCH_DENSE = 3,
    /* NB200_OPT_HOST_VOTE (0|1): nb200_search* / nb200_analyze_search normally replay the
     * heaps and run find_best_match on the device (vote_device.cu) and download only the
     * per-video results; 1 downloads the runs and votes on the host (vote.cpp) instead.
     * Results are identical; the host vote is also the automatic fallback when a table
     * has more runs, or a video more candidates, than the device kernels are sized for. */
    NB200_OPT_HOST_VOTE = 4,
    /* NB200_OPT_DEFER_WAIT (0|1): nb200_fingerprint_run_into / nb200_fingerprint_host_into return
     * as soon as their work is enqueued on the context's stream instead of waiting for it.  For
     * pipelines that go on to exchange and match on the same stream (the multi-GPU job) and
     * synchronise once at the end; host PCM passed to nb200_fingerprint_host_into must then stay
     * untouched until that synchronisation. */
    NB200_OPT_DEFER_WAIT = 5,
    /* NB200_OPT_MATCH_BAND_GROUP (0 = automatic, 1..16): how many consecutive 256-diagonal bands of a
     * row chunk the fast match kernel's warps claim as one unit of work.  Results never depend on it;
     * tests force it to cover the multi-band path on small seasons. */
    NB200_OPT_MATCH_BAND_GROUP = 6
};
int nb200_ctx_set_option(nb200_ctx *ctx, int option, int64_t value);

/* Wall-clock milliseconds the library spent in each host-side phase (including
 * the waits on the GPU inside it) since the last reset; for finding where a
 * step's time goes beyond the kernels.  ms_out[NB200_HOST_PHASES]. */
enum {
    NB200_PHASE_PCM_UPLOAD = 0,   /* nb200_pcmset_upload: H2D (+ down-mix) */
    NB200_PHASE_FP = 1,           /* nb200_fingerprint_run: K1 + K2 */
    NB200_PHASE_LMIN = 2,         /* run-length bounds kernel + read-back */
    NB200_PHASE_MATCH_SETUP = 3,  /* problem/tile tables built and uploaded */
    NB200_PHASE_MATCH = 4,        /* K3 (+ retry) until the run count is known */
    NB200_PHASE_SIMHASH = 5,      /* K4 + copy into the run set */
    NB200_PHASE_RUNS_DOWNLOAD = 6,/* D2H of runs + sort */
    NB200_PHASE_VOTE = 7,         /* heap replay + find_best_match */
    NB200_HOST_PHASES = 8
};
int nb200_ctx_host_profile(nb200_ctx *ctx, double *ms_out, int reset);

/* Pinned host memory (cudaHostAlloc) for PCM / hash staging. */
int nb200_host_alloc(void **out, size_t bytes);
int nb200_host_free(void *p);
/* Frees any buffer this library returned through an out-pointer. */
void nb200_free(void *p);

/* ---------------------------------------------------------- shared structs */

/* Comparator configuration (comparator.rs:74-82, defaults audio/mod.rs:14-45) */
typedef struct nb200_match_params {
    uint32_t hash_match_threshold;   /* default 10, <= 32 */
    uint32_t include_endings;        /* 0/1 */
    uint64_t min_opening_ns;         /* default 20 s */
    uint64_t min_ending_ns;          /* default 20 s */
    uint64_t time_padding_ns;        /* default 0 */
} nb200_match_params;
void nb200_match_params_default(nb200_match_params *p);

/* One maximal diagonal run that passed the duration filter == one
 * ComparatorHeapEntry (comparator.rs:20-35, built at :231-243):
 * score = len, src run = ts_src[i_end-len]..ts_src[i_end], likewise dst. */
typedef struct nb200_run {
    uint32_t pair;        /* index into the pair list */
    uint32_t is_ending;   /* 0: opening x opening table, 1: ending x ending */
    uint32_t i_end;       /* src index of the last matching cell */
    uint32_t j_end;       /* dst index of the last matching cell */
    uint32_t len;         /* table[i_end][j_end] */
    uint32_t src_simhash; /* simhash32(src[i_end-len ..= i_end]) (comparator.rs:149-153) */
    uint32_t dst_simhash;
    uint32_t reserved;
    /* src_longest_run / dst_longest_run (comparator.rs:208-209,233-234), ns */
    uint64_t src_start_ns; /* ts_src[i_end - len] */
    uint64_t src_end_ns;   /* ts_src[i_end] */
    uint64_t dst_start_ns; /* ts_dst[j_end - len] */
    uint64_t dst_end_ns;   /* ts_dst[j_end] */
} nb200_run;

/* SearchResult (comparator.rs:65-69) with the video index kept: element v
 * of the results array belongs to video v; present == 0 means the reference's
 * results Vec has no element for it (find_best_match returned None). */
typedef struct nb200_search_result {
    uint32_t present, has_opening, has_ending, reserved;
    uint64_t opening_start_ns, opening_end_ns;
    uint64_t ending_start_ns, ending_end_ns;
} nb200_search_result;

/* ------------------------------------------------- B2: match, host buffers */

/* A season in SoA form: video k's opening list is
 * hashes[seg_offset[2k] .. seg_offset[2k+1]), its ending list
 * hashes[seg_offset[2k+1] .. seg_offset[2k+2]); ts_ns likewise.
 * (FrameHashes::opening_data / ending_data, data.rs:143-155.)
 *
 * pairs == NULL: all (i < j) in the reference's order (comparator.rs:534-545).
 * Runs are returned sorted by (pair, is_ending, i_end desc, j_end desc) = the
 * reference's heap-push order (comparator.rs:191-192); *out_runs is owned by
 * the caller (nb200_free). */
int nb200_match_pairs(nb200_ctx *ctx, const uint32_t *hashes, const uint64_t *ts_ns,
                      const uint64_t *seg_offset, uint32_t n_videos,
                      const uint32_t (*pairs)[2], uint64_t n_pairs,
                      const nb200_match_params *params, nb200_run **out_runs, uint64_t *out_n);

/* Comparator::run_with_frame_hashes: match every pair on the GPU, then replay
 * the reference's BinaryHeap pushes and find_best_match (on the device too, unless
 * NB200_OPT_HOST_VOTE; identical results).
 * hash_duration_ns[n_videos] = FrameHashes::hash_duration per video.
 * results[n_videos]. */
int nb200_search(nb200_ctx *ctx, const uint32_t *hashes, const uint64_t *ts_ns,
                 const uint64_t *seg_offset, const uint64_t *hash_duration_ns, uint32_t n_videos,
                 const nb200_match_params *params, nb200_search_result *results);

/* nb200_search on a season that already lives in HBM (made by
 * nb200_fingerprint_run or nb200_hashset_from_device). */
int nb200_search_hashset(nb200_ctx *ctx, const nb200_hashset *hs, const uint64_t *hash_duration_ns,
                         const nb200_match_params *params, nb200_search_result *results);

/* Host-only tail of nb200_search for callers that gathered runs themselves
 * (multi-GPU: every rank matches its share of the pair list, rank 0 votes).
 * runs must be sorted as nb200_match_pairs returns them; pairs as above.  A run
 * record carries its own timestamps, so no hash or timestamp array is needed. */
int nb200_vote(const uint64_t *hash_duration_ns, uint32_t n_videos, const uint32_t (*pairs)[2],
               uint64_t n_pairs, const nb200_match_params *params, const nb200_run *runs,
               uint64_t n_runs, nb200_search_result *results);

/* nb200_vote for the videos v with video_mask[v] != 0 only (NULL = all); the others get
 * present = 0.  All runs are still needed (a video's candidates come from every pair it
 * is in).  Lets the ranks of a multi-GPU job split the per-video work. */
int nb200_vote_subset(const uint64_t *hash_duration_ns, uint32_t n_videos, const uint32_t (*pairs)[2],
                      uint64_t n_pairs, const nb200_match_params *params, const nb200_run *runs,
                      uint64_t n_runs, const uint8_t *video_mask, nb200_search_result *results);

/* ------------------------------------------ device-resident staged pipeline */

/* A season from host arrays (layout as for nb200_match_pairs).  The caller's arrays are free again when
 * the call returns; the copies themselves are left to the context's stream.  Page-locked arrays
 * (nb200_host_alloc, cudaHostAlloc, cudaHostRegister) are copied from directly -- two contiguous H2D
 * copies and a device-side move into the aligned layout; ordinary memory is staged through the
 * library's own pinned area first (about 1 ms per 10 MB).  nb200_search / nb200_match_pairs take the
 * same route for their inputs. */
int nb200_hashset_upload(nb200_ctx *ctx, const uint32_t *hashes, const uint64_t *ts_ns,
                         const uint64_t *seg_offset, uint32_t n_videos, nb200_hashset **out);
/* Sizes: total = number of hashes; seg_offset_out[2*n_videos+1] (may be NULL). */
int nb200_hashset_info(const nb200_hashset *hs, uint32_t *n_videos, uint64_t *total,
                       uint64_t *seg_offset_out);
int nb200_hashset_download(nb200_ctx *ctx, const nb200_hashset *hs, uint32_t *hashes_out,
                           uint64_t *ts_ns_out);
/* Raw device pointers of the packed (contiguous, seg_offset-indexed) arrays,
 * for collectives run by the host framework (e.g. NCCL all-gather). */
int nb200_hashset_device_ptrs(const nb200_hashset *hs, void **d_hashes, void **d_ts_ns);
/* Device-to-device copy of the set into caller-owned device arrays in the
 * PACKED layout of nb200_hashset_upload's inputs (total u32 / total u64) --
 * the send buffers of an all-gather. */
int nb200_hashset_export_packed(nb200_ctx *ctx, const nb200_hashset *hs, void *d_hashes_out,
                                void *d_ts_ns_out);
/* Wrap device arrays already laid out like nb200_hashset_upload's inputs
 * (e.g. the result of an all-gather).  Copies device-to-device. */
int nb200_hashset_from_device(nb200_ctx *ctx, const void *d_hashes, const void *d_ts_ns,
                              const uint64_t *seg_offset, uint32_t n_videos, nb200_hashset **out);
/* Same, for segments scattered over two device arrays: segment s (2k = opening
 * of video k, 2k+1 = its ending) has seg_len[s] elements starting at element
 * src_hash_off[s] of d_hashes and src_ts_off[s] of d_ts_ns.  This is how the
 * receive buffer of the multi-GPU all-gather (rank-major, padded) becomes a
 * season in the reference's video order. */
int nb200_hashset_from_device_scattered(nb200_ctx *ctx, const void *d_hashes, const void *d_ts_ns,
                                        const uint64_t *src_hash_off, const uint64_t *src_ts_off,
                                        const uint64_t *seg_len, uint32_t n_videos, nb200_hashset **out);
/* A season that lives in caller-owned device memory, not copied: segment s has
 * seg_len[s] elements starting at element dev_offset[s] of BOTH arrays (offsets
 * multiples of 4).  The memory must stay valid while the set is in use.  This is
 * how the multi-GPU flow matches straight out of the all-gather receive
 * buffers (one for hashes, one for timestamps, rank-major, equal capacity). */
int nb200_hashset_view(nb200_ctx *ctx, const void *d_hashes, const void *d_ts_ns,
                       const uint64_t *dev_offset, const uint64_t *seg_len, uint32_t n_videos,
                       nb200_hashset **out);
void nb200_hashset_free(nb200_hashset *hs);

/* Launch the match kernels for the given pairs (NULL = all i<j); returns once
 * the kernels are enqueued and the run count is known. */
int nb200_match_run(nb200_ctx *ctx, const nb200_hashset *hs, const uint32_t (*pairs)[2],
                    uint64_t n_pairs, const nb200_match_params *params, nb200_runset **out);
int nb200_runset_count(const nb200_runset *rs, uint64_t *n_runs, uint64_t *n_cells);
/* Sorted like nb200_match_pairs' output; runs_out sized by nb200_runset_count. */
int nb200_runset_download(nb200_ctx *ctx, const nb200_runset *rs, nb200_run *runs_out);
void nb200_runset_free(nb200_runset *rs);

/* --------------------------------------------------------- B1: fingerprint */

/* Chromaprint constants needle reads back (analyzer.rs:288-289) */
#define NB200_SAMPLE_RATE 11025
#define NB200_DELAY_MS 2600          /* chromaprint_get_delay_ms */
#define NB200_ITEM_DURATION_MS 123   /* chromaprint_get_item_duration_ms */

/* Number of raw sub-fingerprints for n mono samples: frames - 19 */
uint64_t nb200_num_raw_hashes(uint64_t n_mono_samples);

/* H2D of n_segments PCM buffers (interleaved i16 @ 11025 Hz, channels 1|2;
 * n_samples_total[k] counts i16 values, i.e. frames*channels, as needle passes
 * to feed, analyzer.rs:263-275).  Pinned source buffers copy at full PCIe rate. */
int nb200_pcmset_upload(nb200_ctx *ctx, const int16_t *const *pcm, const uint64_t *n_samples_total,
                        int channels, uint32_t n_segments, nb200_pcmset **out);
void nb200_pcmset_free(nb200_pcmset *ps);
/* A PCM set over mono S16 @ 11025 Hz that is ALREADY in device memory (decoded or generated on the
 * GPU, or placed there by the caller's own copies): nothing is copied, the arrays stay the
 * caller's.  d_pcm 16-byte aligned; dev_offset[s] = element offset of segment s (any value);
 * capacity = elements allocated at d_pcm, at least 8 beyond the end of every segment (the frame
 * loads are 16-byte bulk copies).  Segments come in (opening, ending) pairs like everywhere else. */
int nb200_pcmset_view(nb200_ctx *ctx, const void *d_pcm, const uint64_t *dev_offset, const uint64_t *n_mono_samples,
                      uint32_t n_segments, uint64_t capacity, nb200_pcmset **out);

/* Fingerprint every segment on the GPU and keep raw index 0, stride, 2*stride..
 * (step_by, analyzer.rs:293-305), stamped
 *   ts = delay + item.mul_f32(i as f32) + seek_to   (analyzer.rs:309,314-318).
 * Segments 2k and 2k+1 become video k's opening and ending lists; pass
 * n_segments = 2*n_videos with empty ending segments when endings are off.
 * seek_to_ns may be NULL (all zero). */
int nb200_fingerprint_run(nb200_ctx *ctx, const nb200_pcmset *ps, uint32_t stride,
                          uint64_t delay_ns, uint64_t item_ns, const uint64_t *seek_to_ns,
                          nb200_hashset **out);

/* Same, writing into caller-owned device arrays (capacity elements each: u32
 * hashes, u64 timestamps) in the library's device layout -- every segment
 * starts at the next multiple of 4 elements, in segment order.  dev_offset_out
 * [n_segments] and seg_len_out[n_segments] (may be NULL) receive that layout;
 * nb200_fingerprint_layout computes it without running anything.  With the
 * arrays being all-gather send buffers, fingerprints go to the wire without a copy. */
int nb200_fingerprint_layout(const uint64_t *n_mono_samples, uint32_t n_segments, uint32_t stride,
                             uint64_t *dev_offset_out, uint64_t *seg_len_out, uint64_t *total_out);
int nb200_fingerprint_run_into(nb200_ctx *ctx, const nb200_pcmset *ps, uint32_t stride,
                               uint64_t delay_ns, uint64_t item_ns, const uint64_t *seek_to_ns,
                               void *d_hashes_out, void *d_ts_ns_out, uint64_t capacity);

/* Fills in the timestamps of stored hashes from their index alone,
 *   ts[k] = delay + item.mul_f32((k * stride) as f32) + seek_to   (analyzer.rs:309,314-318),
 * for n_segments segments of seg_len[s] elements at element dev_offset[s] of d_ts_ns.
 * Asynchronous on the context's stream.  A rank of a multi-GPU job regenerates the
 * timestamps of every segment this way instead of receiving them. */
int nb200_timestamps_fill(nb200_ctx *ctx, void *d_ts_ns, const uint64_t *dev_offset, const uint64_t *seg_len,
                          const uint64_t *seek_to_ns, uint32_t n_segments, uint32_t stride, uint64_t delay_ns,
                          uint64_t item_ns);

/* nb200_pcmset_upload + nb200_fingerprint_run_into in one call for MONO host PCM
 * (pinned for full rate): the H2D copy runs in chunks on a second stream while K1
 * already works on the chunks that have landed.  n_segments must be even. */
int nb200_fingerprint_host_into(nb200_ctx *ctx, const int16_t *const *pcm, const uint64_t *n_samples,
                                uint32_t n_segments, uint32_t stride, uint64_t delay_ns, uint64_t item_ns,
                                const uint64_t *seek_to_ns, void *d_hashes_out, void *d_ts_ns_out,
                                uint64_t capacity);

/* One-shot with host buffers: raw sub-fingerprints (stride 1, no timestamps)
 * of each segment == chromaprint_get_raw_fingerprint.  out_hashes[k] is
 * nb200_free'd by the caller. */
int nb200_fingerprint_batch(nb200_ctx *ctx, const int16_t *const *pcm,
                            const uint64_t *n_samples_total, int channels, uint32_t n_segments,
                            uint32_t stride, uint32_t **out_hashes, uint64_t *out_counts);

/* Chromaprint-shaped streaming shim: a type swap for chromaprint::Context in
 * Analyzer::process_frames.  feed() accumulates (down-mixing stereo) into a
 * pinned buffer; finish() runs the kernels. */
typedef struct nb200_fp nb200_fp;
int nb200_fp_new(nb200_ctx *ctx, nb200_fp **out);          /* Context::default()        :176 */
void nb200_fp_free(nb200_fp *fp);
int nb200_fp_sample_rate(const nb200_fp *fp);              /* Context::sample_rate()    :179 */
int nb200_fp_start(nb200_fp *fp, int sample_rate, int channels); /* start(rate, 2)      :218 */
int nb200_fp_feed(nb200_fp *fp, const int16_t *data, size_t n_samples_total); /* feed   :275 */
int nb200_fp_finish(nb200_fp *fp);                         /* finish()                  :286 */
int nb200_fp_get_delay_ms(const nb200_fp *fp, int *out);   /* get_delay()               :288 */
int nb200_fp_get_item_duration_ms(const nb200_fp *fp, int *out); /* get_item_duration() :289 */
/* Borrowed pointer, valid until the next start()/free() (get_fingerprint_raw :299-301). */
int nb200_fp_get_raw(const nb200_fp *fp, const uint32_t **hashes, size_t *n);

/* ------------------------------------------------ fused analyze + search */

/* `needle search --analyze`: PCM segments (2 per video: opening, ending) ->
 * fingerprints -> all-pairs match -> vote, hashes never leaving HBM.
 * hash_duration_ns: what Analyzer was asked for (Duration::from_secs_f32(0.3)
 * = 300000012); stride = hash_duration_ms / 123 (analyzer.rs:293-297). */
int nb200_analyze_search(nb200_ctx *ctx, const int16_t *const *pcm, const uint64_t *n_samples_total,
                         int channels, uint32_t n_videos, const uint64_t *seek_to_ns,
                         uint64_t hash_duration_ns, const nb200_match_params *params,
                         nb200_search_result *results);
/* The same on PCM that is already in HBM (nb200_pcmset_upload; 2 segments per video):
 * K1, K2, match, simhash and the vote are enqueued back to back on the context's
 * stream, the host waits once and downloads 48 bytes per video. */
int nb200_pcmset_search(nb200_ctx *ctx, const nb200_pcmset *ps, const uint64_t *seek_to_ns,
                        uint64_t hash_duration_ns, const nb200_match_params *params,
                        nb200_search_result *results);

/* ------------------------------------------- multi-GPU: run blocks in HBM
 *
 * A rank's share of the pair loop as ONE fixed-size device message that can be
 * handed to a collective as it is: 64-byte header {u32 n_found, u32 n_stored,
 * u32 flags, ...} followed by 64-byte run records (the device form of
 * nb200_run, carrying the GLOBAL pair index pair_base + k and all four
 * timestamps).  Nothing is waited for: the match kernels emit straight into
 * the block on the context's stream.  block_bytes = 64 * (1 + capacity). */
int nb200_match_export(nb200_ctx *ctx, const nb200_hashset *hs, const uint32_t (*pairs)[2], uint64_t n_pairs,
                       const nb200_match_params *params, uint32_t pair_base, void *d_block, uint64_t block_bytes);
/* The blocks of all ranks, back to back in device memory (an all-gather receive
 * buffer): with do_vote != 0 the heaps are replayed and find_best_match runs on
 * the device over all of them (pairs / n_pairs = the GLOBAL pair list, NULL = all
 * i < j) and results[n_videos] is filled; with do_vote == 0 only the headers are
 * checked.  Synchronises the stream once.  *max_found = the largest n_found of any
 * block; NB200_ERR_TOO_LARGE when a block could not hold its runs (every rank sees
 * the same headers, so every rank can repeat the step with larger blocks). */
int nb200_vote_blocks(nb200_ctx *ctx, const void *d_blocks, uint32_t n_blocks, uint64_t block_bytes,
                      const uint64_t *hash_duration_ns, uint32_t n_videos, const uint32_t (*pairs)[2],
                      uint64_t n_pairs, const nb200_match_params *params, int do_vote,
                      nb200_search_result *results, uint64_t *max_found);

/* ------------------------------------------------ multi-GPU jobs (one box)
 *
 * The reference fans out over videos and over pairs inside ONE process with rayon
 * (Analyzer::run, analyzer.rs:437-445; Comparator::run_with_frame_hashes,
 * comparator.rs:549-564).  These calls are that fan-out across the GPUs of a box:
 * videos are sharded for fingerprinting, the hash arrays are all-gathered once
 * (NCCL, loaded with dlopen: no link-time dependency), the pair list is sharded by
 * table cells for matching, every rank's runs are pushed into rank 0's memory over
 * NVLink (peer stores from a kernel, not a collective) and rank 0 votes.
 *
 * nb200_comm = one rank: a context plus its NCCL communicator.  Two ways to get them:
 *   - one process, N devices (needle's own shape): nb200_comm_init_all over N contexts
 *     (ncclCommInitAll); pass all N comms to the job calls;
 *   - one process per device (torchrun, MPI): rank 0 calls nb200_comm_unique_id, the host
 *     framework hands the 128 bytes to every process, each calls nb200_comm_init_rank and
 *     passes its one comm to the job calls.
 * Job calls are collective: every rank makes the same calls in the same order with the same
 * global metadata; per-video results land in `results` of the process that drives rank 0.
 * A comm switches its context to NB200_OPT_DEFER_WAIT. */
typedef struct nb200_comm nb200_comm;
typedef struct nb200_mjob nb200_mjob;
#define NB200_UNIQUE_ID_BYTES 128
int nb200_comm_unique_id(uint8_t id[NB200_UNIQUE_ID_BYTES]);
int nb200_comm_init_rank(nb200_ctx *ctx, const uint8_t id[NB200_UNIQUE_ID_BYTES], int rank, int world,
                         nb200_comm **out);
int nb200_comm_init_all(nb200_ctx *const *ctxs, int n, nb200_comm **out /* [n] */);
int nb200_comm_info(const nb200_comm *comm, int *rank, int *world);
void nb200_comm_destroy(nb200_comm *comm);

/* `needle search` from precomputed frame hashes (.needle.dat) across the ranks: every rank is
 * given the same season (KB..MB, uploaded once, resident for the life of the job) and matches its
 * slice of the pair list.  comms[n_local] = the ranks this process drives, ascending.
 * Replaces the pair loop of Comparator::run_with_frame_hashes (comparator.rs:532-578). */
int nb200_mjob_search_create(nb200_comm *const *comms, int n_local, const uint32_t *hashes, const uint64_t *ts_ns,
                             const uint64_t *seg_offset, uint32_t n_videos, const uint64_t *hash_duration_ns,
                             const uint32_t (*pairs)[2], uint64_t n_pairs, const nb200_match_params *params,
                             nb200_mjob **out);
/* `needle search --analyze` for a library of videos: n_mono_samples / seek_to_ns [2 * n_videos]
 * describe ALL segments (opening_k, ending_k); the library decides which rank fingerprints which
 * video (nb200_mjob_video_rank) and each process supplies PCM only for its own ranks' videos.
 * Replaces Analyzer::run's per-video loop (analyzer.rs:437-445) + the pair loop. */
int nb200_mjob_season_create(nb200_comm *const *comms, int n_local, const uint64_t *n_mono_samples,
                             const uint64_t *seek_to_ns, uint32_t n_videos, uint64_t hash_duration_ns,
                             const uint32_t (*pairs)[2], uint64_t n_pairs, const nb200_match_params *params,
                             nb200_mjob **out);
int nb200_mjob_video_rank(const nb200_mjob *job, uint32_t *rank_of_video /* [n_videos] */);
/* Season job: mono PCM of this process's videos into HBM, for nb200_mjob_run(job, NULL, ..).
 * pcm[2 * n_videos], entries of videos that belong to other processes' ranks are ignored. */
int nb200_mjob_upload_pcm(nb200_mjob *job, const int16_t *const *pcm);
/* One pass of the job.  Season job: host_pcm != NULL streams the PCM from (pinned) host memory
 * underneath K1, NULL uses the uploaded PCM.  Search job: host_pcm is ignored.
 * results[n_videos]: required where rank 0 is local. */
int nb200_mjob_run(nb200_mjob *job, const int16_t *const *host_pcm, nb200_search_result *results);
/* CUDA-event milliseconds of the phases of the last run on this process's lowest rank. */
enum {
    NB200_MJOB_FINGERPRINT = 0,   /* K1 + K2 (+ pipelined H2D) */
    NB200_MJOB_HASH_ALLGATHER = 1,/* ncclAllGather of the hashes (includes waiting for the slowest rank) */
    NB200_MJOB_MATCH = 2,         /* K3 + K4 on this rank's slice of the pair list */
    NB200_MJOB_RUN_PUSH_WAIT = 3, /* rank 0: until every rank's run block has landed */
    NB200_MJOB_VOTE = 4,          /* rank 0: heap replay + find_best_match + results to the host */
    NB200_MJOB_PHASES = 5
};
int nb200_mjob_phase_ms(const nb200_mjob *job, float *ms /* [NB200_MJOB_PHASES] */);
void nb200_mjob_free(nb200_mjob *job);
/* The jobs' sharding rules on their own (host only, no device needed): which rank fingerprints
 * which video (greedy by sample count, n_mono_samples[2 * n_videos]), and where the pair list
 * is cut (cuts[world + 1], slices balanced by table cells; pairs == NULL: all i < j). */
int nb200_plan_videos(const uint64_t *n_mono_samples, uint32_t n_videos, int world, uint32_t *rank_of_video);
int nb200_plan_pairs(const uint64_t *seg_len, uint32_t n_videos, const uint32_t (*pairs)[2], uint64_t n_pairs, int world,
                     int include_endings, uint64_t *cuts);

/* ------------------------------------------------------------ persistence */

/* .needle.dat (bincode 1.3, fixint little-endian) of one video.
 * write: opening/ending as (hash, ts_ns) SoA; md5 = 32 hex chars. */
int nb200_framehashes_write(const char *path, const uint32_t *open_hash, const uint64_t *open_ts_ns,
                            uint64_t n_open, const uint32_t *end_hash, const uint64_t *end_ts_ns,
                            uint64_t n_end, uint64_t hash_duration_ns, const char *md5);
/* read: out arrays are nb200_free'd by the caller; md5_out holds >= 33 bytes. */
int nb200_framehashes_read(const char *path, uint32_t **open_hash, uint64_t **open_ts_ns,
                           uint64_t *n_open, uint32_t **end_hash, uint64_t **end_ts_ns,
                           uint64_t *n_end, uint64_t *hash_duration_ns, char *md5_out);

/* ------------------------------------- libneedle.so (include/needle.h) hooks
 *
 * The needle-capi ABI of include/needle.h is served by libneedle.so (capi.cpp),
 * which links against this library.  In the reference, FFmpeg demuxes, decodes and
 * resamples to 11025 Hz S16 before Chromaprint sees a sample (analyzer.rs:170-283);
 * that stays host code.  libneedle.so gets its PCM from this callback table; when
 * none is installed it reads RIFF/WAVE files that already carry 11025 Hz S16 mono or
 * stereo PCM.  All callbacks return 0 on success.
 *   probe:   stream duration (the reference's stream/format duration, analyzer.rs:352-371)
 *            and whether the file has video / audio streams (util.rs:41-53)
 *   decode:  interleaved S16 at 11025 Hz of [from_ns, until_ns) (UINT64_MAX = to the
 *            end); every file of one analyzer run must use the same channel count (1|2)
 *   release: frees a buffer returned by decode */
typedef struct nb200_capi_decoder {
    void *user;
    int (*probe)(void *user, const char *path, uint64_t *duration_ns, int *has_video, int *has_audio);
    int (*decode)(void *user, const char *path, uint64_t from_ns, uint64_t until_ns, int16_t **pcm,
                  uint64_t *n_samples_total, int *channels);
    void (*release)(void *user, int16_t *pcm);
} nb200_capi_decoder;
/* NULL restores the built-in WAVE reader.  Exported by libneedle.so. */
int nb200_capi_set_decoder(const nb200_capi_decoder *decoder);
/* Read access to a FrameHashes handle of needle.h (opaque there, as in the reference):
 * the opening (ending = 0) or ending (1) list as SoA, the hash duration and the header md5.
 * Pointers are borrowed from the handle.  Exported by libneedle.so. */
struct FrameHashes;
int nb200_capi_frame_hashes_view(const struct FrameHashes *f, int ending, const uint32_t **hashes,
                                 const uint64_t **ts_ns, uint64_t *n, uint64_t *hash_duration_ns,
                                 const char **md5);

/* util::compute_header_md5sum (needle/src/util.rs:99-105): lower-case hex md5 of the first 8 KiB
 * of a file, the key that ties a .needle.dat / .needle.skip.json to its video.  NB200_ERR_IO
 * when the file is missing or shorter than 8 KiB (read_exact fails in the reference).
 * Exported by libneedle.so. */
int nb200_capi_header_md5(const char *path, char out[33]);

#ifdef __cplusplus
}
#endif
#endif /* NEEDLE_B200_H */
