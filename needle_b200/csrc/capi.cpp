// capi.cpp -- libneedle.so: the needle-capi C ABI (include/needle.h) on top of
// libneedle_b200.so.  SURVEY 8(f4).
//
// Mirrors, function by function, needle-capi/src/lib.rs (argument checks and error
// values) and the library calls underneath it:
//   needle/src/util.rs:8-105           format_time, is_valid_video_file, find_video_files,
//                                      compute_header_md5sum
//   needle/src/audio/analyzer.rs:325-456   Analyzer::run_single / run
//   needle/src/audio/comparator.rs:310-378, 524-663   skip files, display, run
//   needle/src/audio/data.rs:8-13, 116-141  SkipFile, FrameHashes::from_video
// Host code only: the fingerprinting and the search are calls into the C ABI of
// include/needle_b200.h (nb200_pcmset_upload / nb200_fingerprint_run / nb200_search);
// without a CUDA device those fail and the error is reported, nothing is computed here.
//
// The reference decodes with FFmpeg; here a decoder callback supplies 11025 Hz S16 PCM
// (nb200_capi_set_decoder), with a RIFF/WAVE reader as the built-in stand-in.
#include <dirent.h>
#include <sys/stat.h>

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/needle.h"
#include "../../include/needle_b200.h"

namespace {

constexpr const char *FRAME_HASH_DATA_FILE_NAME = "needle.dat";   // needle/src/lib.rs:154
constexpr const char *SKIP_FILE_NAME = "needle.skip.json";        // needle/src/lib.rs:155
constexpr float DEFAULT_OPENING_SEARCH_PERCENTAGE = 0.50f;        // needle/src/audio/mod.rs:14-45
constexpr float DEFAULT_ENDING_SEARCH_PERCENTAGE = 0.25f;
constexpr uint16_t DEFAULT_HASH_MATCH_THRESHOLD = 10;
constexpr uint16_t DEFAULT_MIN_OPENING_DURATION = 20;
constexpr uint16_t DEFAULT_MIN_ENDING_DURATION = 20;
constexpr float DEFAULT_HASH_DURATION = 0.3f;
constexpr uint64_t SAMPLE_RATE = 11025;
constexpr uint64_t NS = 1000000000ull;

// ------------------------------------------------------------------ Duration (Rust semantics)
uint64_t duration_from_secs_f32(float x) {   // exact value of the f32 in ns, ties to even
    return (uint64_t)std::nearbyint((double)x * 1e9);
}
float duration_as_secs_f32(uint64_t ns) {
    volatile float s = (float)(ns / NS);
    volatile float f = (float)(uint32_t)(ns % NS) / 1000000000.0f;
    return s + f;
}
uint64_t duration_mul_f32(uint64_t ns, float rhs) {
    volatile float p = rhs * duration_as_secs_f32(ns);
    return duration_from_secs_f32(p);
}

// ------------------------------------------------------------------ MD5 (RFC 1321)
struct Md5 {
    uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
    static uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
    void block(const unsigned char *p) {
        static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,
                                  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
        uint32_t m[16];
        for (int i = 0; i < 16; i++)
            m[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
                   ((uint32_t)p[4 * i + 3] << 24);
        uint32_t A = a, B = b, C = c, D = d;
        for (int i = 0; i < 64; i++) {
            uint32_t f;
            int g;
            if (i < 16) {
                f = (B & C) | (~B & D);
                g = i;
            } else if (i < 32) {
                f = (D & B) | (~D & C);
                g = (5 * i + 1) & 15;
            } else if (i < 48) {
                f = B ^ C ^ D;
                g = (3 * i + 5) & 15;
            } else {
                f = C ^ (B | ~D);
                g = (7 * i) & 15;
            }
            const uint32_t k = (uint32_t)(int64_t)std::floor(std::fabs(std::sin((double)(i + 1))) * 4294967296.0);
            const uint32_t t = D;
            D = C;
            C = B;
            B = B + rol(A + f + k + m[g], S[i]);
            A = t;
        }
        a += A;
        b += B;
        c += C;
        d += D;
    }
    static std::string hex(const unsigned char *data, size_t n) {
        Md5 h;
        size_t i = 0;
        for (; i + 64 <= n; i += 64) h.block(data + i);
        unsigned char tail[128] = {0};
        const size_t rest = n - i;
        memcpy(tail, data + i, rest);
        tail[rest] = 0x80;
        const size_t padded = rest < 56 ? 64 : 128;
        const uint64_t bits = (uint64_t)n * 8;
        for (int k = 0; k < 8; k++) tail[padded - 8 + k] = (unsigned char)(bits >> (8 * k));
        h.block(tail);
        if (padded == 128) h.block(tail + 64);
        char out[33];
        const uint32_t w[4] = {h.a, h.b, h.c, h.d};
        for (int k = 0; k < 16; k++) snprintf(out + 2 * k, 3, "%02x", (w[k / 4] >> (8 * (k % 4))) & 0xff);
        return std::string(out, 32);
    }
};

// util::compute_header_md5sum (util.rs:99-105): md5 of the first 8 KiB; read_exact semantics
NeedleError header_md5(const std::string &video, std::string *out) {
    FILE *f = fopen(video.c_str(), "rb");
    if (!f) {
        fprintf(stderr, "needle error: IO error: cannot open %s\n", video.c_str());
        return NeedleError_IOError;
    }
    unsigned char buf[8192];
    const size_t got = fread(buf, 1, sizeof(buf), f);
    fclose(f);
    if (got < sizeof(buf)) {
        fprintf(stderr, "needle error: IO error: failed to fill whole buffer (%s)\n", video.c_str());
        return NeedleError_IOError;
    }
    *out = Md5::hex(buf, sizeof(buf));
    return NeedleError_Ok;
}

// ------------------------------------------------------------------ paths
bool path_exists(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}
bool path_is_dir(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
// Path::with_extension: everything after the last '.' of the file name is replaced
std::string with_extension(const std::string &path, const char *ext) {
    const size_t slash = path.find_last_of('/');
    const size_t name = slash == std::string::npos ? 0 : slash + 1;
    const size_t dot = path.find_last_of('.');
    std::string stem = (dot == std::string::npos || dot <= name) ? path : path.substr(0, dot);
    return stem + "." + ext;
}
bool ends_with(const std::string &s, const char *suffix) {
    const size_t n = strlen(suffix);
    return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}
// util::format_time (util.rs:8-12)
std::string format_time(uint64_t ns) {
    char b[48];
    const uint64_t s = ns / NS;
    snprintf(b, sizeof(b), "%02llu:%02llus", (unsigned long long)(s / 60), (unsigned long long)(s % 60));
    return b;
}
bool valid_utf8(const char *s) {
    const unsigned char *p = (const unsigned char *)s;
    while (*p) {
        int n = *p < 0x80 ? 0 : (*p >> 5) == 6 ? 1 : (*p >> 4) == 14 ? 2 : (*p >> 3) == 30 ? 3 : -1;
        if (n < 0) return false;
        p++;
        for (; n > 0; n--, p++)
            if ((*p >> 6) != 2) return false;
    }
    return true;
}
// get_paths_from_raw (needle-capi/src/lib.rs:283-303)
NeedleError paths_from_raw(const char *const *raw, size_t n, std::vector<std::string> *out) {
    for (size_t k = 0; k < n; k++) {
        if (!raw[k]) return NeedleError_NullArgument;
        if (!valid_utf8(raw[k])) return NeedleError_InvalidUtf8String;
        out->push_back(raw[k]);
    }
    return NeedleError_Ok;
}

// ------------------------------------------------------------------ decoder
nb200_capi_decoder g_decoder = {nullptr, nullptr, nullptr, nullptr};

struct Wav {
    int channels = 0;
    uint64_t frames = 0;       // samples per channel
    long data_offset = 0;
};
bool wav_open(const char *path, Wav *w) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    unsigned char h[12];
    bool ok = fread(h, 1, 12, f) == 12 && !memcmp(h, "RIFF", 4) && !memcmp(h + 8, "WAVE", 4);
    bool have_fmt = false;
    while (ok) {
        unsigned char c[8];
        if (fread(c, 1, 8, f) != 8) {
            ok = false;
            break;
        }
        const uint32_t size = c[4] | (c[5] << 8) | (c[6] << 16) | ((uint32_t)c[7] << 24);
        if (!memcmp(c, "fmt ", 4)) {
            unsigned char q[16];
            if (size < 16 || fread(q, 1, 16, f) != 16) {
                ok = false;
                break;
            }
            const int format = q[0] | (q[1] << 8), channels = q[2] | (q[3] << 8), bits = q[14] | (q[15] << 8);
            const uint32_t rate = q[4] | (q[5] << 8) | (q[6] << 16) | ((uint32_t)q[7] << 24);
            if (format != 1 || bits != 16 || rate != SAMPLE_RATE || (channels != 1 && channels != 2)) {
                ok = false;
                break;
            }
            w->channels = channels;
            have_fmt = true;
            fseek(f, (long)(size - 16 + (size & 1)), SEEK_CUR);
        } else if (!memcmp(c, "data", 4)) {
            if (!have_fmt) ok = false;
            w->data_offset = ftell(f);
            w->frames = size / (2u * (uint32_t)w->channels);
            break;
        } else {
            fseek(f, (long)(size + (size & 1)), SEEK_CUR);
        }
    }
    fclose(f);
    return ok && have_fmt && w->data_offset > 0;
}
int wav_probe(void *, const char *path, uint64_t *duration_ns, int *has_video, int *has_audio) {
    Wav w;
    if (!wav_open(path, &w)) return 1;
    *duration_ns = w.frames * NS / SAMPLE_RATE;
    *has_video = 1;   // the stand-in for "a video whose audio has been decoded"
    *has_audio = 1;
    return 0;
}
int wav_decode(void *, const char *path, uint64_t from_ns, uint64_t until_ns, int16_t **pcm, uint64_t *n_samples_total,
               int *channels) {
    Wav w;
    if (!wav_open(path, &w)) return 1;
    const uint64_t first = std::min<uint64_t>(w.frames, from_ns * SAMPLE_RATE / NS);
    const uint64_t last = until_ns == UINT64_MAX ? w.frames : std::min<uint64_t>(w.frames, until_ns * SAMPLE_RATE / NS);
    const uint64_t n = last > first ? (last - first) * (uint64_t)w.channels : 0;
    int16_t *buf = (int16_t *)malloc(std::max<uint64_t>(n, 1) * sizeof(int16_t));
    if (!buf) return 1;
    FILE *f = fopen(path, "rb");
    bool ok = f && fseek(f, w.data_offset + (long)(first * 2 * (uint64_t)w.channels), SEEK_SET) == 0 &&
              fread(buf, sizeof(int16_t), n, f) == n;   // little-endian host
    if (f) fclose(f);
    if (!ok) {
        free(buf);
        return 1;
    }
    *pcm = buf;
    *n_samples_total = n;
    *channels = w.channels;
    return 0;
}
void wav_release(void *, int16_t *pcm) { free(pcm); }
const nb200_capi_decoder &decoder() {
    static const nb200_capi_decoder builtin = {nullptr, wav_probe, wav_decode, wav_release};
    return g_decoder.decode ? g_decoder : builtin;
}

// the video container signatures util::is_valid_video_file's header sniff accepts
// (`infer::is_video`, util.rs:34-39): mp4/m4v/mov (ftyp box), mkv/webm (EBML), avi, wmv/asf,
// mpeg program stream / video, flv
bool sniff_video(const unsigned char *b, size_t n) {
    if (n >= 12 && !memcmp(b + 4, "ftyp", 4)) return true;
    if (n >= 4 && b[0] == 0x1A && b[1] == 0x45 && b[2] == 0xDF && b[3] == 0xA3) return true;
    if (n >= 12 && !memcmp(b, "RIFF", 4) && !memcmp(b + 8, "AVI ", 4)) return true;
    if (n >= 10 && b[0] == 0x30 && b[1] == 0x26 && b[2] == 0xB2 && b[3] == 0x75 && b[4] == 0x8E && b[5] == 0x66 &&
        b[6] == 0xCF && b[7] == 0x11 && b[8] == 0xA6 && b[9] == 0xD9)
        return true;
    if (n >= 4 && b[0] == 0 && b[1] == 0 && b[2] == 1 && b[3] >= 0xB0 && b[3] <= 0xBF) return true;
    if (n >= 4 && !memcmp(b, "FLV", 3) && b[3] == 1) return true;
    return false;
}
bool is_valid_video_file(const std::string &path, bool full, bool audio) {
    if (ends_with(path, FRAME_HASH_DATA_FILE_NAME)) return false;
    if (path_is_dir(path)) return false;
    if (!full) {
        unsigned char buf[8192];
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) return false;
        const size_t n = fread(buf, 1, sizeof(buf), f);
        fclose(f);
        if (sniff_video(buf, n)) return true;
        // decoded-audio stand-in: what the built-in decoder can read counts as a video here
        return n >= 12 && !memcmp(buf, "RIFF", 4) && !memcmp(buf + 8, "WAVE", 4);
    }
    uint64_t dur = 0;
    int has_video = 0, has_audio = 0;
    const nb200_capi_decoder &d = decoder();
    if (d.probe(d.user, path.c_str(), &dur, &has_video, &has_audio) != 0) return false;
    return has_video && (!audio || has_audio);
}

// ------------------------------------------------------------------ GPU context
nb200_ctx *gpu_ctx(NeedleError *err) {
    static nb200_ctx *ctx = nullptr;
    if (!ctx) {
        const char *dev = getenv("NEEDLE_B200_DEVICE");
        const int st = nb200_ctx_create(dev ? atoi(dev) : -1, &ctx);
        if (st != NB200_OK) {
            fprintf(stderr, "needle error: GPU: %s (%s)\n", nb200_status_str(st), nb200_last_error());
            ctx = nullptr;
            *err = NeedleError_Unknown;
        }
    }
    return ctx;
}
NeedleError from_status(int st, const char *where) {
    if (st == NB200_OK) return NeedleError_Ok;
    fprintf(stderr, "needle error: %s: %s (%s)\n", where, nb200_status_str(st), nb200_last_error());
    switch (st) {
    case NB200_ERR_NULL_ARGUMENT: return NeedleError_NullArgument;
    case NB200_ERR_IO: return NeedleError_IOError;
    case NB200_ERR_FORMAT: return NeedleError_InvalidFrameHashData;
    case NB200_ERR_COMPARATOR_MINIMUM_PATHS: return NeedleError_ComparatorMinimumPaths;
    default: return NeedleError_Unknown;   // needle::Error::* without a NeedleError of its own (lib.rs:121-134)
    }
}

// serde_json's f32 formatting: shortest representation that round-trips, always with a
// fraction or an exponent
std::string json_f32(float x) {
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), x);
    std::string s(buf, r.ptr);
    // ryu (what serde_json prints with) writes exponents without sign padding: 1e-5, 1e16
    const size_t e = s.find('e');
    if (e != std::string::npos) {
        size_t k = e + 1;
        std::string sign;
        if (k < s.size() && (s[k] == '+' || s[k] == '-')) {
            if (s[k] == '-') sign = "-";
            k++;
        }
        while (k + 1 < s.size() && s[k] == '0') k++;
        s = s.substr(0, e + 1) + sign + s.substr(k);
    }
    if (s.find_first_of(".eE") == std::string::npos) s += ".0";
    return s;
}

}  // namespace

// The opaque handle types of needle.h ------------------------------------------
struct FrameHashes {   // FrameHashesV1 (data.rs:20-26) in SoA form
    std::vector<uint32_t> opening_hash, ending_hash;
    std::vector<uint64_t> opening_ts, ending_ts;
    uint64_t hash_duration_ns = 0;
    std::string md5;
};
struct NeedleAudioAnalyzer {   // audio::Analyzer + the results of the last run (lib.rs:332-335)
    std::vector<std::string> videos;
    float opening_search_percentage = DEFAULT_OPENING_SEARCH_PERCENTAGE;
    float ending_search_percentage = DEFAULT_ENDING_SEARCH_PERCENTAGE;
    bool include_endings = false, threaded_decoding = false, force = false;
    std::vector<FrameHashes> frame_hashes;
};
struct NeedleAudioComparator {   // audio::Comparator (comparator.rs:74-82)
    std::vector<std::string> videos;
    bool include_endings = false;
    uint32_t hash_match_threshold = DEFAULT_HASH_MATCH_THRESHOLD;
    uint64_t min_opening_ns = DEFAULT_MIN_OPENING_DURATION * NS, min_ending_ns = DEFAULT_MIN_ENDING_DURATION * NS;
    uint64_t time_padding_ns = 0;
};

namespace {

NeedleError read_frame_hashes(const std::string &path, FrameHashes *out) {
    if (!path_exists(path)) {
        fprintf(stderr, "needle error: frame hash data not found at: %s\n", path.c_str());
        return NeedleError_FrameHashDataNotFound;
    }
    uint32_t *oh = nullptr, *eh = nullptr;
    uint64_t *ot = nullptr, *et = nullptr, no = 0, ne = 0, hd = 0;
    char md5[64] = {0};
    const int st = nb200_framehashes_read(path.c_str(), &oh, &ot, &no, &eh, &et, &ne, &hd, md5);
    if (st == NB200_ERR_FORMAT) {
        fprintf(stderr, "needle error: invalid frame hash data in %s\n", path.c_str());
        return NeedleError_InvalidFrameHashData;
    }
    if (st != NB200_OK) return from_status(st, "nb200_framehashes_read");
    out->opening_hash.assign(oh, oh + no);
    out->opening_ts.assign(ot, ot + no);
    out->ending_hash.assign(eh, eh + ne);
    out->ending_ts.assign(et, et + ne);
    out->hash_duration_ns = hd;
    out->md5 = md5;
    nb200_free(oh);
    nb200_free(ot);
    nb200_free(eh);
    nb200_free(et);
    return NeedleError_Ok;
}

struct AnalyzerConfig {
    float opening_pct, ending_pct;
    bool include_endings, force;
};

// Analyzer::run_single for a list of videos (analyzer.rs:325-421): cached results are
// reused as the reference does; everything else is decoded on the host and fingerprinted
// on the GPU in ONE batch.
NeedleError analyze_videos(const std::vector<std::string> &videos, const AnalyzerConfig &cfg, uint64_t hash_duration_ns,
                           bool persist, std::vector<FrameHashes> *out) {
    // step_by = hash_duration.as_millis() / item_duration.as_millis()   (analyzer.rs:293-297)
    const uint64_t stride = (hash_duration_ns / 1000000ull) / NB200_ITEM_DURATION_MS;
    if (stride == 0) {
        fprintf(stderr, "needle error: hash duration shorter than one fingerprint item\n");
        return NeedleError_AnalyzerInvalidHashDuration;   // step_by(0) panics in the reference
    }
    out->assign(videos.size(), FrameHashes());
    const nb200_capi_decoder &dec = decoder();
    std::vector<size_t> todo;
    std::vector<int16_t *> pcm;
    std::vector<uint64_t> counts, seeks;
    int channels = 0;
    NeedleError err = NeedleError_Ok;
    auto release_all = [&]() {
        for (int16_t *p : pcm)
            if (p) dec.release(dec.user, p);
    };
    for (size_t k = 0; k < videos.size() && err == NeedleError_Ok; k++) {
        const std::string &path = videos[k];
        std::string md5;
        if ((err = header_md5(path, &md5)) != NeedleError_Ok) break;
        const std::string dat = with_extension(path, FRAME_HASH_DATA_FILE_NAME);
        if (!cfg.force && path_exists(dat)) {
            FrameHashes cached;
            if (read_frame_hashes(dat, &cached) == NeedleError_Ok && cached.md5 == md5) {
                printf("Skipping analysis for %s...\n", path.c_str());
                (*out)[k] = std::move(cached);
                continue;
            }
        }
        uint64_t dur = 0;
        int has_video = 0, has_audio = 0;
        if (dec.probe(dec.user, path.c_str(), &dur, &has_video, &has_audio) != 0 || !has_audio) {
            fprintf(stderr, "needle error: cannot open an audio stream in %s\n", path.c_str());
            err = NeedleError_Unknown;   // needle::Error::FFmpegError -> Unknown
            break;
        }
        const uint64_t opening_ns = duration_mul_f32(dur, cfg.opening_pct);
        int16_t *o = nullptr, *e = nullptr;
        uint64_t no = 0, ne = 0, seek = 0;
        int ch = 0, ch2 = 0;
        if (dec.decode(dec.user, path.c_str(), 0, opening_ns, &o, &no, &ch) != 0) {
            err = NeedleError_Unknown;
            break;
        }
        pcm.push_back(o);
        if (cfg.include_endings) {
            seek = duration_mul_f32(dur, 1.0f - cfg.ending_pct);
            if (dec.decode(dec.user, path.c_str(), seek, UINT64_MAX, &e, &ne, &ch2) != 0 || ch2 != ch) {
                err = NeedleError_Unknown;
                pcm.push_back(e);
                break;
            }
        }
        pcm.push_back(e);
        if (channels == 0) channels = ch;
        if (ch != channels) {   // needle's resampler always yields one layout; a decoder must too
            fprintf(stderr, "needle error: decoder returned %d and %d channels in one run\n", channels, ch);
            err = NeedleError_Unknown;
            break;
        }
        (*out)[k].md5 = md5;
        todo.push_back(k);
        counts.push_back(no);
        counts.push_back(ne);
        seeks.push_back(0);
        seeks.push_back(seek);
    }
    if (err != NeedleError_Ok || todo.empty()) {
        release_all();
        return err;
    }
    nb200_ctx *ctx = gpu_ctx(&err);
    if (!ctx) {
        release_all();
        return err;
    }
    nb200_pcmset *ps = nullptr;
    nb200_hashset *hs = nullptr;
    static const int16_t none = 0;
    std::vector<const int16_t *> ptrs;
    for (int16_t *p : pcm) ptrs.push_back(p ? p : &none);
    int st = nb200_pcmset_upload(ctx, ptrs.data(), counts.data(), channels, (uint32_t)ptrs.size(), &ps);
    if (st == NB200_OK)
        st = nb200_fingerprint_run(ctx, ps, (uint32_t)stride, NB200_DELAY_MS * 1000000ull,
                                   NB200_ITEM_DURATION_MS * 1000000ull, seeks.data(), &hs);
    release_all();
    std::vector<uint32_t> hashes;
    std::vector<uint64_t> ts, off(ptrs.size() + 1, 0);
    if (st == NB200_OK) {
        uint64_t total = 0;
        st = nb200_hashset_info(hs, nullptr, &total, off.data());
        hashes.resize(std::max<uint64_t>(total, 1));
        ts.resize(std::max<uint64_t>(total, 1));
        if (st == NB200_OK) st = nb200_hashset_download(ctx, hs, hashes.data(), ts.data());
    }
    if (hs) nb200_hashset_free(hs);
    if (ps) nb200_pcmset_free(ps);
    if (st != NB200_OK) return from_status(st, "fingerprint");
    for (size_t j = 0; j < todo.size(); j++) {
        FrameHashes &f = (*out)[todo[j]];
        const uint64_t a = off[2 * j], b = off[2 * j + 1], c = off[2 * j + 2];
        f.opening_hash.assign(hashes.begin() + (long)a, hashes.begin() + (long)b);
        f.opening_ts.assign(ts.begin() + (long)a, ts.begin() + (long)b);
        f.ending_hash.assign(hashes.begin() + (long)b, hashes.begin() + (long)c);
        f.ending_ts.assign(ts.begin() + (long)b, ts.begin() + (long)c);
        f.hash_duration_ns = hash_duration_ns;   // the REQUESTED duration (analyzer.rs:321)
        if (persist) {
            const std::string dat = with_extension(videos[todo[j]], FRAME_HASH_DATA_FILE_NAME);
            st = nb200_framehashes_write(dat.c_str(), f.opening_hash.data(), f.opening_ts.data(), f.opening_hash.size(),
                                         f.ending_hash.data(), f.ending_ts.data(), f.ending_hash.size(),
                                         f.hash_duration_ns, f.md5.c_str());
            if (st != NB200_OK) return from_status(st, "nb200_framehashes_write");
        }
    }
    return NeedleError_Ok;
}

// Comparator::check_skip_file (comparator.rs:310-327)
NeedleError check_skip_file(const std::string &video, bool *skip) {
    *skip = false;
    const std::string path = with_extension(video, SKIP_FILE_NAME);
    if (!path_exists(path)) return NeedleError_Ok;
    std::string md5;
    const NeedleError e = header_md5(video, &md5);
    if (e != NeedleError_Ok) return e;
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return NeedleError_IOError;
    std::string text;
    char buf[512];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
    fclose(f);
    const size_t key = text.find("\"md5\"");
    if (key == std::string::npos) return NeedleError_Unknown;   // serde_json error -> unwrap panic in the reference
    const size_t q0 = text.find('"', text.find(':', key) + 1);
    const size_t q1 = q0 == std::string::npos ? q0 : text.find('"', q0 + 1);
    if (q1 == std::string::npos) return NeedleError_Unknown;
    *skip = text.substr(q0 + 1, q1 - q0 - 1) == md5;
    return NeedleError_Ok;
}

// Comparator::create_skip_file (comparator.rs:329-354); serde_json::to_writer's compact form
NeedleError create_skip_file(const std::string &video, const nb200_search_result &r) {
    if (!r.has_opening && !r.has_ending) return NeedleError_Ok;
    std::string md5;
    const NeedleError e = header_md5(video, &md5);
    if (e != NeedleError_Ok) return e;
    auto pair = [](bool have, uint64_t a, uint64_t b) {
        if (!have) return std::string("null");
        return "[" + json_f32(duration_as_secs_f32(a)) + "," + json_f32(duration_as_secs_f32(b)) + "]";
    };
    const std::string text = "{\"opening\":" + pair(r.has_opening, r.opening_start_ns, r.opening_end_ns) +
                             ",\"ending\":" + pair(r.has_ending, r.ending_start_ns, r.ending_end_ns) + ",\"md5\":\"" +
                             md5 + "\"}";
    FILE *f = fopen(with_extension(video, SKIP_FILE_NAME).c_str(), "wb");
    if (!f) return NeedleError_IOError;
    const bool ok = fwrite(text.data(), 1, text.size(), f) == text.size();
    fclose(f);
    return ok ? NeedleError_Ok : NeedleError_IOError;
}

}  // namespace

extern "C" {

int nb200_capi_set_decoder(const nb200_capi_decoder *decoder) {
    if (!decoder) {
        g_decoder = {nullptr, nullptr, nullptr, nullptr};
        return NB200_OK;
    }
    if (!decoder->probe || !decoder->decode || !decoder->release) return NB200_ERR_NULL_ARGUMENT;
    g_decoder = *decoder;
    return NB200_OK;
}

int nb200_capi_frame_hashes_view(const FrameHashes *f, int ending, const uint32_t **hashes, const uint64_t **ts_ns,
                                 uint64_t *n, uint64_t *hash_duration_ns, const char **md5) {
    if (!f) return NB200_ERR_NULL_ARGUMENT;
    const std::vector<uint32_t> &h = ending ? f->ending_hash : f->opening_hash;
    const std::vector<uint64_t> &t = ending ? f->ending_ts : f->opening_ts;
    if (hashes) *hashes = h.data();
    if (ts_ns) *ts_ns = t.data();
    if (n) *n = h.size();
    if (hash_duration_ns) *hash_duration_ns = f->hash_duration_ns;
    if (md5) *md5 = f->md5.c_str();
    return NB200_OK;
}

int nb200_capi_header_md5(const char *path, char out[33]) {
    if (!path || !out) return NB200_ERR_NULL_ARGUMENT;
    std::string md5;
    if (header_md5(path, &md5) != NeedleError_Ok) return NB200_ERR_IO;
    memcpy(out, md5.c_str(), 33);
    return NB200_OK;
}

const char *needle_error_to_str(enum NeedleError error) {
    switch (error) {   // needle-capi/src/lib.rs:139-203
    case NeedleError_Ok: return "No error";
    case NeedleError_InvalidUtf8String: return "Invalid UTF-8 string";
    case NeedleError_NullArgument: return "Input argument is NULL";
    case NeedleError_InvalidArgument: return "One or more input arguments were invalid (usually zero)";
    case NeedleError_FrameHashDataNotFound: return "Frame hash data not found on disk";
    case NeedleError_FrameHashDataInvalidVersion: return "Frame hash data has an invalid version.";
    case NeedleError_InvalidFrameHashData: return "Invalid frame hash data read from disk";
    case NeedleError_ComparatorMinimumPaths: return "Comparator requires at least 2 video paths";
    case NeedleError_AnalyzerInvalidHashPeriod: return "Analyzer hash period must be greater than 0";
    case NeedleError_AnalyzerInvalidHashDuration: return "Analyzer hash duration must be greater than 3 seconds";
    case NeedleError_IOError: return "I/O error";
    case NeedleError_Unknown: break;
    }
    return "Unknown error occurred; please re-run with logging enabled";
}

enum NeedleError needle_util_find_video_files(const char *const *paths, size_t num_paths, bool full, bool audio,
                                              const char *const **videos, size_t *num_videos) {
    if (!paths || !videos || !num_videos) return NeedleError_NullArgument;
    if (num_paths == 0) return NeedleError_InvalidArgument;
    std::vector<std::string> in;
    const NeedleError e = paths_from_raw(paths, num_paths, &in);
    if (e != NeedleError_Ok) return e;
    for (const std::string &p : in)   // util.rs:66-72
        if (!path_exists(p)) {
            fprintf(stderr, "needle error: path does not exist: %s\n", p.c_str());
            return NeedleError_Unknown;   // Error::PathNotFound has no NeedleError of its own
        }
    std::vector<std::string> found;
    for (const std::string &p : in) {
        if (path_is_dir(p)) {
            if (DIR *d = opendir(p.c_str())) {
                while (struct dirent *ent = readdir(d)) {
                    if (!strcmp(ent->d_name, ".") || !strcmp(ent->d_name, "..")) continue;
                    const std::string child = p + (ends_with(p, "/") ? "" : "/") + ent->d_name;
                    if (is_valid_video_file(child, full, audio)) found.push_back(child);
                }
                closedir(d);
            }
        } else if (is_valid_video_file(p, full, audio)) {
            found.push_back(p);
        }
    }
    char **arr = (char **)malloc(std::max<size_t>(found.size(), 1) * sizeof(char *));
    if (!arr) return NeedleError_Unknown;
    for (size_t k = 0; k < found.size(); k++) arr[k] = strdup(found[k].c_str());
    *videos = (const char *const *)arr;
    *num_videos = found.size();
    return NeedleError_Ok;
}

void needle_util_video_files_free(const char *const *videos, size_t num_videos) {
    if (!videos) return;
    for (size_t k = 0; k < num_videos; k++) free((void *)videos[k]);
    free((void *)videos);
}

enum NeedleError needle_audio_analyzer_new(const char *const *paths, size_t num_paths, float opening_search_percentage,
                                           float ending_search_percentage, bool include_endings,
                                           bool threaded_decoding, bool force, struct NeedleAudioAnalyzer **output) {
    if (!paths || !output) return NeedleError_NullArgument;
    auto *a = new NeedleAudioAnalyzer();
    const NeedleError e = paths_from_raw(paths, num_paths, &a->videos);
    if (e != NeedleError_Ok) {
        delete a;
        return e;
    }
    a->opening_search_percentage = opening_search_percentage;
    a->ending_search_percentage = ending_search_percentage;
    a->include_endings = include_endings;
    a->threaded_decoding = threaded_decoding;
    a->force = force;
    *output = a;
    return NeedleError_Ok;
}

enum NeedleError needle_audio_analyzer_new_default(const char *const *paths, size_t num_paths,
                                                   struct NeedleAudioAnalyzer **output) {
    return needle_audio_analyzer_new(paths, num_paths, DEFAULT_OPENING_SEARCH_PERCENTAGE,
                                     DEFAULT_ENDING_SEARCH_PERCENTAGE, false, false, false, output);
}

enum NeedleError needle_audio_analyzer_get_frame_hashes(const struct NeedleAudioAnalyzer *analyzer, size_t index,
                                                        const struct FrameHashes **output) {
    if (!analyzer || !output) return NeedleError_NullArgument;
    if (index >= analyzer->frame_hashes.size()) return NeedleError_InvalidArgument;
    *output = &analyzer->frame_hashes[index];
    return NeedleError_Ok;
}

void needle_audio_analyzer_free(const struct NeedleAudioAnalyzer *analyzer) { delete analyzer; }

void needle_audio_analyzer_print_paths(const struct NeedleAudioAnalyzer *analyzer) {
    if (!analyzer) return;
    for (const std::string &p : analyzer->videos) printf("%s\n", p.c_str());
    fflush(stdout);
}

enum NeedleError needle_audio_analyzer_run(struct NeedleAudioAnalyzer *analyzer, float hash_duration, bool persist,
                                           bool threading) {
    (void)threading;
    if (!analyzer) return NeedleError_NullArgument;
    if (hash_duration <= 0.0f) return NeedleError_AnalyzerInvalidHashDuration;
    if (analyzer->videos.empty()) {
        fprintf(stderr, "needle error: no paths provided to analyzer\n");
        return NeedleError_Unknown;   // Error::AnalyzerMissingPaths -> Unknown (lib.rs:130)
    }
    const AnalyzerConfig cfg = {analyzer->opening_search_percentage, analyzer->ending_search_percentage,
                                analyzer->include_endings, analyzer->force};
    std::vector<FrameHashes> out;
    const NeedleError e = analyze_videos(analyzer->videos, cfg, duration_from_secs_f32(hash_duration), persist, &out);
    fflush(stdout);
    if (e != NeedleError_Ok) return e;
    analyzer->frame_hashes = std::move(out);
    return NeedleError_Ok;
}

enum NeedleError needle_audio_comparator_new(const char *const *paths, size_t num_paths, bool include_endings,
                                             uint16_t hash_match_threshold, uint16_t min_opening_duration,
                                             uint16_t min_ending_duration, float time_padding,
                                             const struct NeedleAudioComparator **output) {
    if (!paths || !output) return NeedleError_NullArgument;
    if (num_paths < 2) return NeedleError_ComparatorMinimumPaths;
    auto *c = new NeedleAudioComparator();
    const NeedleError e = paths_from_raw(paths, num_paths, &c->videos);
    if (e != NeedleError_Ok) {
        delete c;
        return e;
    }
    c->include_endings = include_endings;
    c->hash_match_threshold = hash_match_threshold;
    c->min_opening_ns = (uint64_t)min_opening_duration * NS;
    c->min_ending_ns = (uint64_t)min_ending_duration * NS;
    c->time_padding_ns = duration_from_secs_f32(time_padding);
    *output = c;
    return NeedleError_Ok;
}

enum NeedleError needle_audio_comparator_new_default(const char *const *paths, size_t num_paths,
                                                     const struct NeedleAudioComparator **output) {
    return needle_audio_comparator_new(paths, num_paths, false, DEFAULT_HASH_MATCH_THRESHOLD,
                                       DEFAULT_MIN_OPENING_DURATION, DEFAULT_MIN_ENDING_DURATION, 0.0f, output);
}

void needle_audio_comparator_free(const struct NeedleAudioComparator *comparator) { delete comparator; }

enum NeedleError needle_audio_comparator_run(const struct NeedleAudioComparator *comparator, bool analyze, bool display,
                                             bool use_skip_files, bool write_skip_files, bool threading) {
    (void)threading;
    if (!comparator) return NeedleError_NullArgument;
    const NeedleAudioComparator &c = *comparator;
    // FrameHashes::from_video for every video (comparator.rs:650-654, data.rs:124-141)
    std::vector<FrameHashes> fh;
    if (analyze) {
        // Analyzer::default().with_force(true): default percentages, NO endings
        const AnalyzerConfig cfg = {DEFAULT_OPENING_SEARCH_PERCENTAGE, DEFAULT_ENDING_SEARCH_PERCENTAGE, false, true};
        const NeedleError e = analyze_videos(c.videos, cfg, duration_from_secs_f32(DEFAULT_HASH_DURATION), false, &fh);
        if (e != NeedleError_Ok) return e;
    } else {
        fh.resize(c.videos.size());
        for (size_t k = 0; k < c.videos.size(); k++) {
            const NeedleError e = read_frame_hashes(with_extension(c.videos[k], FRAME_HASH_DATA_FILE_NAME), &fh[k]);
            if (e != NeedleError_Ok) return e;
        }
    }
    // run_with_frame_hashes (comparator.rs:524-629): one search on the GPU
    std::vector<uint32_t> hashes;
    std::vector<uint64_t> ts, off(1, 0), hd;
    for (const FrameHashes &f : fh) {
        hashes.insert(hashes.end(), f.opening_hash.begin(), f.opening_hash.end());
        ts.insert(ts.end(), f.opening_ts.begin(), f.opening_ts.end());
        off.push_back(hashes.size());
        hashes.insert(hashes.end(), f.ending_hash.begin(), f.ending_hash.end());
        ts.insert(ts.end(), f.ending_ts.begin(), f.ending_ts.end());
        off.push_back(hashes.size());
        hd.push_back(f.hash_duration_ns);
    }
    NeedleError err = NeedleError_Ok;
    nb200_ctx *ctx = gpu_ctx(&err);
    if (!ctx) return err;
    nb200_match_params p;
    p.hash_match_threshold = c.hash_match_threshold;
    p.include_endings = c.include_endings ? 1u : 0u;
    p.min_opening_ns = c.min_opening_ns;
    p.min_ending_ns = c.min_ending_ns;
    p.time_padding_ns = c.time_padding_ns;
    std::vector<nb200_search_result> res(fh.size());
    if (hashes.empty()) {
        hashes.push_back(0);
        ts.push_back(0);
    }
    const int st = nb200_search(ctx, hashes.data(), ts.data(), off.data(), hd.data(), (uint32_t)fh.size(), &p, res.data());
    if (st != NB200_OK) return from_status(st, "nb200_search");   // e.g. FrameHashDataNoEnding -> Unknown
    for (size_t idx = 0; idx < res.size(); idx++) {   // comparator.rs:592-626
        const std::string &path = c.videos[idx];
        if (display) printf("\n%s\n\n", path.c_str());
        if (use_skip_files) {
            bool skip = false;
            const NeedleError e = check_skip_file(path, &skip);
            if (e != NeedleError_Ok) return e;
            if (skip) {
                if (display) printf("Skipping due to existing skip file...\n");
                continue;
            }
        }
        const nb200_search_result &r = res[idx];
        if (!r.present) {
            if (display) printf(c.include_endings ? "No opening or ending found.\n" : "No opening found.\n");
            continue;
        }
        if (display) {   // display_opening_ending_info (comparator.rs:356-378)
            if (r.has_opening)
                printf("* Opening - \"%s\"-\"%s\"\n", format_time(r.opening_start_ns).c_str(),
                       format_time(r.opening_end_ns).c_str());
            else
                printf("* Opening - N/A\n");
            if (c.include_endings) {
                if (r.has_ending)
                    printf("* Ending - \"%s\"-\"%s\"\n", format_time(r.ending_start_ns).c_str(),
                           format_time(r.ending_end_ns).c_str());
                else
                    printf("* Ending - N/A\n");
            }
        }
        if (write_skip_files) {
            const NeedleError e = create_skip_file(path, r);
            if (e != NeedleError_Ok) return e;
        }
    }
    fflush(stdout);
    return NeedleError_Ok;
}

}  // extern "C"
