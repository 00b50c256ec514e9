/*
 * needle_oracle.h -- CPU restatement of aksiksi/needle's fingerprint-and-match
 * path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the checker the CUDA path is compared against.  Nothing under
 * needle_b200/ (the product) may include, link or call it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the reference ships no golden vectors or known-answer tests
 * for either stage (needle/src/audio/analyzer.rs:472-480 is #[ignore]d with a
 * stale snapshot; comparator.rs has no tests), and neither the Rust crate nor
 * Chromaprint 1.5.x (crate chromaprint-sys-next 1.5.3, needle/Cargo.lock:158)
 * can be built here (no cargo/rustc, sources un-vendored).  The match stage
 * follows needle/src/audio/comparator.rs line by line; the fingerprint stage
 * restates Chromaprint's published TEST2 algorithm from recall.  The recalled
 * upstream unit-test vectors for Chromaprint's building blocks (chroma note
 * mapping, chroma filter, quantizer, gray code, simhash) are checked in
 * tests/test_oracle_kat.py and labelled as recalled.
 *
 * All times are u64 nanoseconds (Rust Duration = u64 s + u32 ns; its derived
 * Ord equals the order of the total ns for every value that fits).
 */
#ifndef NEEDLE_ORACLE_H
#define NEEDLE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Rust Duration / f32 semantics (needle/src/audio/analyzer.rs:293-318) -- */

/* Duration::from_secs_f32: exact value of the f32, rounded to the nearest ns,
 * ties to even.  Returns UINT64_MAX for negative / non-finite input (Rust
 * panics). */
uint64_t orc_duration_from_secs_f32(float secs);
/* Duration::as_secs_f32: (secs as f32) + (nanos as f32) / 1e9f32 */
float orc_duration_as_secs_f32(uint64_t ns);
/* Duration::mul_f32(rhs) = from_secs_f32(rhs * self.as_secs_f32()) */
uint64_t orc_duration_mul_f32(uint64_t ns, float rhs);
/* Timestamp of raw chromaprint item i: delay + item.mul_f32(i as f32) + seek_to
 * (analyzer.rs:309,314-318). */
uint64_t orc_hash_timestamp(uint64_t delay_ns, uint64_t item_ns, uint32_t raw_index,
                            uint64_t seek_to_ns);

/* ---- match stage (needle/src/audio/comparator.rs) ------------------------- */

/* chromaprint-rust 0.1.3 simhash::simhash32 (== Chromaprint src/simhash.cpp):
 * per-bit majority, ties -> 0. */
uint32_t orc_simhash32(const uint32_t *hashes, size_t n);

/* One ComparatorHeapEntry (comparator.rs:20-35).  The four is_* flags collapse
 * to is_ending because longest_common_hash_match sets them all from is_opening
 * (comparator.rs:231-243). */
typedef struct orc_entry {
    uint64_t score;          /* table[i][j] = run length */
    uint64_t src_start_ns, src_end_ns;
    uint64_t dst_start_ns, dst_end_ns;
    uint32_t src_match_hash, dst_match_hash;
    uint32_t is_ending;      /* 0 = opening, 1 = ending */
    uint64_t src_hash_duration_ns, dst_hash_duration_ns;
    /* not part of the Rust struct: where the run ended, for kernel parity */
    uint32_t i_end, j_end;
} orc_entry;

/* Comparator::longest_common_hash_match (comparator.rs:157-250): the literal
 * (n+1)x(m+1) usize table, forward fill, reverse scan, duration filter,
 * simhash, BinaryHeap push; returns the heap's internal array (heap.into()).
 * *out is malloc'd (free with orc_free).  Returns the entry count, or -1 when
 * a Duration subtraction would underflow (Rust panics). */
int64_t orc_longest_common_hash_match(const uint32_t *src_hash, const uint64_t *src_ts, size_t n,
                                      const uint32_t *dst_hash, const uint64_t *dst_ts, size_t m,
                                      uint32_t threshold, uint64_t min_opening_ns,
                                      uint64_t min_ending_ns, uint64_t src_hash_duration_ns,
                                      uint64_t dst_hash_duration_ns, int is_opening,
                                      orc_entry **out);
void orc_free(void *p);

/* A season: N videos, each with an opening and an ending hash list, SoA. */
typedef struct orc_season {
    size_t n_videos;
    const uint32_t *hashes;       /* concatenated */
    const uint64_t *ts_ns;        /* concatenated, same indexing */
    const uint64_t *seg_offset;   /* [2*N+1]: opening_k at 2k, ending_k at 2k+1 */
    const uint64_t *hash_duration_ns; /* [N] FrameHashes::hash_duration */
} orc_season;

typedef struct orc_params {
    uint32_t hash_match_threshold;
    uint64_t min_opening_ns, min_ending_ns;
    uint64_t time_padding_ns;
    int include_endings;
} orc_params;

/* SearchResult with the video index retained (comparator.rs:65-69; Q7). */
typedef struct orc_result {
    uint32_t present;      /* find_best_match returned Some(..) */
    uint32_t has_opening, has_ending;
    uint64_t opening_start_ns, opening_end_ns;
    uint64_t ending_start_ns, ending_end_ns;
} orc_result;

/* Error codes of orc_run_with_frame_hashes */
#define ORC_OK 0
#define ORC_ERR_NO_ENDING (-2)   /* Error::FrameHashDataNoEnding (comparator.rs:271-273) */
#define ORC_ERR_UNDERFLOW (-3)   /* Duration subtraction would panic */

/* Comparator::run_with_frame_hashes (comparator.rs:524-629) minus display and
 * skip files: pair list, search per pair (n_threads worker threads pulling
 * pairs from a shared counter -- the rayon par_iter of :553-563), info_map,
 * find_best_match per video.  results[N].  If entries_out != NULL it receives
 * every heap entry of every pair, in pair order then opening-before-ending
 * then heap-array order, with pair indices in entry_pair_out (both malloc'd).
 */
int orc_run_with_frame_hashes(const orc_season *season, const orc_params *params, int n_threads,
                              orc_result *results, orc_entry **entries_out,
                              uint32_t **entry_pair_out, uint64_t *n_entries_out);

/* ---- fingerprint stage (Chromaprint 1.5.x TEST2, restated) ---------------- */

#define ORC_FRAME_SIZE 4096
#define ORC_FRAME_HOP 1365       /* frame_size - overlap, overlap = 4096 - 4096/3 */
#define ORC_NUM_BANDS 12
#define ORC_SAMPLE_RATE 11025
#define ORC_DELAY_SAMPLES 28666  /* (5-1 + 16-1)*1365 + 2731 */

/* Number of FFT frames / raw sub-fingerprints produced from n mono samples. */
size_t orc_num_frames(size_t n_mono_samples);
size_t orc_num_raw_hashes(size_t n_mono_samples);

/* chromaprint_start(11025, channels) + feed(pcm) + finish + get_raw_fingerprint.
 * pcm: interleaved i16, n_samples_total = frames * channels (what needle passes
 * to feed, analyzer.rs:263-275).  channels 1 or 2.  out: malloc'd raw hashes.
 * If chroma_out != NULL it receives the 12 raw chroma sums per FFT frame
 * (malloc'd, num_frames*12 doubles) for stage-level parity diagnostics. */
int64_t orc_fingerprint(const int16_t *pcm, size_t n_samples_total, int channels,
                        uint32_t **out, double **chroma_out);

/* Analyzer::process_frames tail (analyzer.rs:288-318): subsample raw hashes by
 * step_by and stamp them.  Returns count written (caller sizes out_* to
 * ceil(n_raw/step_by)). */
size_t orc_subsample_and_stamp(const uint32_t *raw, size_t n_raw, uint32_t step_by,
                               uint64_t delay_ns, uint64_t item_ns, uint64_t seek_to_ns,
                               uint32_t *out_hash, uint64_t *out_ts);

/* Building blocks exposed for the recalled upstream known-answer tests. */
void orc_chroma_notes(int min_freq, int max_freq, int frame_size, int sample_rate,
                      int *min_index, int *max_index, signed char *notes /*[frame_size]*/);
void orc_chroma_fold(const double *power, int frame_size, int min_freq, int max_freq,
                     int sample_rate, double *features /*[12]*/);
/* ChromaFilter over n_rows x 12 rows; returns output rows (n_rows-len+1 or 0). */
size_t orc_chroma_filter(const double *coeffs, int len, const double *rows, size_t n_rows,
                         double *out);
void orc_normalize(double *features /*[12]*/, double threshold);
int orc_quantize(double value, double t0, double t1, double t2);
int orc_gray_code(int i);
/* Filter(type,y,h,w).Apply on a plain n_rows x n_cols image at row offset x */
double orc_filter_apply(int type, int y, int height, int width, const double *image,
                        size_t n_rows, size_t n_cols, size_t x);
/* Real DFT power spectrum of one windowed frame (for FFT self-checks). */
void orc_power_spectrum(const int16_t *frame /*[4096]*/, double *power /*[2049]*/);

/* Multi-threaded fingerprinting of many segments (one segment per worker
 * thread, the rayon par_iter of analyzer.rs:440-444) -- CPU baseline. */
int orc_fingerprint_many(const int16_t *const *pcm, const uint64_t *n_samples_total, int channels,
                         size_t n_segments, int n_threads, uint32_t **out, uint64_t *out_counts);

#ifdef __cplusplus
}
#endif
#endif
