/* needle.h -- the needle-capi C ABI, served by the B200 library.
 *
 * This header declares, symbol for symbol and value for value, the interface that
 * aksiksi/needle's `needle-capi` crate generates (reference: needle-capi/needle.h:12-248,
 * implemented by needle-capi/src/lib.rs): the NeedleError enum (same order, Ok = 0), the
 * three opaque handle types and the 13 functions.  A C program written against the
 * reference's needle.h compiles and links against libneedle.so from this repository
 * without change; underneath, analysis and search run on sm_100a kernels through
 * libneedle_b200.so (include/needle_b200.h).
 *
 * Conventions kept from the reference (file:line are in needle-capi/src/lib.rs):
 *   - every fallible function returns NeedleError; out-parameters come last;
 *   - NULL pointer arguments -> NeedleError_NullArgument (:216, :383, :420, :471, :566, :620);
 *   - handles are heap objects released by the matching *_free, which accepts NULL;
 *   - needle_audio_analyzer_get_frame_hashes lends a pointer owned by the analyzer (:430-432);
 *   - a handle must not be used from two threads at once; distinct handles may be.
 *
 * What differs: demux/decode/resample is FFmpeg in the reference and stays host code
 * there.  Here the library asks a decoder callback for 11025 Hz S16 PCM
 * (include/needle_b200.h: nb200_capi_set_decoder); without one it reads RIFF/WAVE files
 * carrying exactly that format.
 */
#ifndef NEEDLE_H
#define NEEDLE_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum NeedleError {
  NeedleError_Ok = 0,                      /* success */
  NeedleError_InvalidUtf8String,           /* a path is not valid UTF-8 */
  NeedleError_NullArgument,                /* a required pointer was NULL */
  NeedleError_InvalidArgument,             /* an argument was out of range (usually zero) */
  NeedleError_FrameHashDataNotFound,       /* no <video>.needle.dat next to the video */
  NeedleError_FrameHashDataInvalidVersion, /* .needle.dat written by another format version */
  NeedleError_InvalidFrameHashData,        /* .needle.dat does not decode */
  NeedleError_ComparatorMinimumPaths,      /* fewer than two videos */
  NeedleError_AnalyzerInvalidHashPeriod,   /* hash period <= 0 */
  NeedleError_AnalyzerInvalidHashDuration, /* hash duration <= 0 */
  NeedleError_IOError,                     /* file system error */
  NeedleError_Unknown,                     /* anything else (details on stderr) */
} NeedleError;

typedef struct FrameHashes FrameHashes;                     /* one video's (hash, time) lists */
typedef struct NeedleAudioAnalyzer NeedleAudioAnalyzer;     /* needle::audio::Analyzer */
typedef struct NeedleAudioComparator NeedleAudioComparator; /* needle::audio::Comparator */

/* Static, NUL-terminated description of an error value. */
const char *needle_error_to_str(enum NeedleError error);

/* Expands files and directories (one level deep) into the list of video files.
 * full = false sniffs the first 8 KiB for a video container signature; full = true asks the
 * decoder to open the file, and with audio = true also requires an audio stream.
 * The returned array and its strings belong to the caller until needle_util_video_files_free. */
enum NeedleError needle_util_find_video_files(const char *const *paths, size_t num_paths, bool full,
                                              bool audio, const char *const **videos,
                                              size_t *num_videos);
void needle_util_video_files_free(const char *const *videos, size_t num_videos);

/* Analyzer with the default search percentages (0.50 opening, 0.25 ending), no endings,
 * no forced re-analysis. */
enum NeedleError needle_audio_analyzer_new_default(const char *const *paths, size_t num_paths,
                                                   struct NeedleAudioAnalyzer **output);
enum NeedleError needle_audio_analyzer_new(const char *const *paths, size_t num_paths,
                                           float opening_search_percentage,
                                           float ending_search_percentage, bool include_endings,
                                           bool threaded_decoding, bool force,
                                           struct NeedleAudioAnalyzer **output);
/* Frame hashes of the index-th video of the last run; the pointer stays owned by the analyzer. */
enum NeedleError needle_audio_analyzer_get_frame_hashes(const struct NeedleAudioAnalyzer *analyzer,
                                                        size_t index,
                                                        const struct FrameHashes **output);
void needle_audio_analyzer_free(const struct NeedleAudioAnalyzer *analyzer);
/* One path per line on stdout. */
void needle_audio_analyzer_print_paths(const struct NeedleAudioAnalyzer *analyzer);
/* Fingerprints every video (hash_duration in seconds, e.g. 0.3); persist = true writes
 * <video>.needle.dat.  `threading` chose rayon in the reference; here all videos that need
 * analysis go to the GPU as one batch either way. */
enum NeedleError needle_audio_analyzer_run(struct NeedleAudioAnalyzer *analyzer, float hash_duration,
                                           bool persist, bool threading);

/* Comparator with threshold 10, minimum opening and ending duration 20 s, no padding, no endings. */
enum NeedleError needle_audio_comparator_new_default(const char *const *paths, size_t num_paths,
                                                     const struct NeedleAudioComparator **output);
/* Durations in whole seconds, time_padding in seconds. */
enum NeedleError needle_audio_comparator_new(const char *const *paths, size_t num_paths,
                                             bool include_endings, uint16_t hash_match_threshold,
                                             uint16_t min_opening_duration,
                                             uint16_t min_ending_duration, float time_padding,
                                             const struct NeedleAudioComparator **output);
void needle_audio_comparator_free(const struct NeedleAudioComparator *comparator);
/* Searches all pairs of videos for shared openings (and endings).  analyze = true fingerprints
 * in place instead of reading <video>.needle.dat; display prints the intervals; the skip-file
 * flags read / write <video>.needle.skip.json. */
enum NeedleError needle_audio_comparator_run(const struct NeedleAudioComparator *comparator,
                                             bool analyze, bool display, bool use_skip_files,
                                             bool write_skip_files, bool threading);

#ifdef __cplusplus
}
#endif

#endif /* NEEDLE_H */
