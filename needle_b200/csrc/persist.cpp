// persist.cpp -- FrameHashes on disk (.needle.dat), the file Analyzer::run_single
// writes with bincode::serialize_into (needle/src/audio/analyzer.rs:413-417) and
// FrameHashes::from_path reads back (needle/src/audio/data.rs:104-115).
//
// bincode 1.3 default options = fixed-width little-endian integers, u64 lengths,
// enum variants as u32 indices.  Layout of FrameHashes (data.rs:15-26,60-80):
//
//   u32  version       FrameHashesVersion::V1 is variant INDEX 0 (the `= 12345`
//                      discriminant is not serialised by serde)
//   u32  data variant  FrameHashesData::V1 = 0
//   u64  n_open ; n_open x { u32 hash, u64 secs, u32 nanos }     Vec<(u32, Duration)>
//   u64  n_end  ; n_end  x { u32 hash, u64 secs, u32 nanos }
//   u64  secs, u32 nanos                                          hash_duration
//   u64  md5 length ; bytes                                       String
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.h"

namespace {

struct Writer {
    std::vector<unsigned char> buf;
    void u32(uint32_t v) {
        for (int k = 0; k < 4; k++) buf.push_back((unsigned char)(v >> (8 * k)));
    }
    void u64(uint64_t v) {
        for (int k = 0; k < 8; k++) buf.push_back((unsigned char)(v >> (8 * k)));
    }
    void duration(uint64_t ns) {
        u64(ns / 1000000000ull);
        u32((uint32_t)(ns % 1000000000ull));
    }
};

struct Reader {
    const unsigned char *p, *end;
    bool ok = true;
    uint32_t u32() {
        if (end - p < 4) {
            ok = false;
            return 0;
        }
        uint32_t v = 0;
        for (int k = 0; k < 4; k++) v |= (uint32_t)p[k] << (8 * k);
        p += 4;
        return v;
    }
    uint64_t u64() {
        if (end - p < 8) {
            ok = false;
            return 0;
        }
        uint64_t v = 0;
        for (int k = 0; k < 8; k++) v |= (uint64_t)p[k] << (8 * k);
        p += 8;
        return v;
    }
    // Duration deserialisation rejects nanos that carry past u64 seconds; we
    // additionally need secs*1e9+nanos to fit our u64 ns representation.
    uint64_t duration() {
        const uint64_t secs = u64();
        const uint32_t nanos = u32();
        if (!ok) return 0;
        if (secs > (0xffffffffffffffffull - nanos) / 1000000000ull) {
            ok = false;
            return 0;
        }
        return secs * 1000000000ull + nanos;
    }
};

bool read_list(Reader &r, uint32_t **hash, uint64_t **ts, uint64_t *n) {
    const uint64_t count = r.u64();
    if (!r.ok || count > (uint64_t)(r.end - r.p) / 16) return false;
    uint32_t *h = (uint32_t *)malloc((count ? count : 1) * sizeof(uint32_t));
    uint64_t *t = (uint64_t *)malloc((count ? count : 1) * sizeof(uint64_t));
    if (!h || !t) {
        free(h);
        free(t);
        return false;
    }
    for (uint64_t k = 0; k < count; k++) {
        h[k] = r.u32();
        t[k] = r.duration();
    }
    if (!r.ok) {
        free(h);
        free(t);
        return false;
    }
    *hash = h;
    *ts = t;
    *n = count;
    return true;
}

}  // namespace

extern "C" {

int nb200_framehashes_write(const char *path, const uint32_t *open_hash, const uint64_t *open_ts_ns,
                            uint64_t n_open, const uint32_t *end_hash, const uint64_t *end_ts_ns,
                            uint64_t n_end, uint64_t hash_duration_ns, const char *md5) {
    if (!path || !md5) return NB200_ERR_NULL_ARGUMENT;
    if ((n_open && (!open_hash || !open_ts_ns)) || (n_end && (!end_hash || !end_ts_ns)))
        return NB200_ERR_NULL_ARGUMENT;
    Writer w;
    w.buf.reserve(76 + 16 * (n_open + n_end));
    w.u32(0);
    w.u32(0);
    w.u64(n_open);
    for (uint64_t k = 0; k < n_open; k++) {
        w.u32(open_hash[k]);
        w.duration(open_ts_ns[k]);
    }
    w.u64(n_end);
    for (uint64_t k = 0; k < n_end; k++) {
        w.u32(end_hash[k]);
        w.duration(end_ts_ns[k]);
    }
    w.duration(hash_duration_ns);
    const size_t md5_len = strlen(md5);
    w.u64(md5_len);
    w.buf.insert(w.buf.end(), md5, md5 + md5_len);
    FILE *f = fopen(path, "wb");
    if (!f) return NB200_ERR_IO;
    const bool ok = fwrite(w.buf.data(), 1, w.buf.size(), f) == w.buf.size();
    if (fclose(f) != 0 || !ok) return NB200_ERR_IO;
    return NB200_OK;
}

int nb200_framehashes_read(const char *path, uint32_t **open_hash, uint64_t **open_ts_ns,
                           uint64_t *n_open, uint32_t **end_hash, uint64_t **end_ts_ns,
                           uint64_t *n_end, uint64_t *hash_duration_ns, char *md5_out) {
    if (!path || !open_hash || !open_ts_ns || !n_open || !end_hash || !end_ts_ns || !n_end ||
        !hash_duration_ns || !md5_out)
        return NB200_ERR_NULL_ARGUMENT;
    *open_hash = *end_hash = nullptr;
    *open_ts_ns = *end_ts_ns = nullptr;
    *n_open = *n_end = 0;
    md5_out[0] = 0;
    FILE *f = fopen(path, "rb");
    if (!f) return NB200_ERR_IO;   // Error::FrameHashDataNotFound / io
    std::vector<unsigned char> buf;
    unsigned char chunk[65536];
    size_t got;
    while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) buf.insert(buf.end(), chunk, chunk + got);
    const bool read_err = ferror(f) != 0;
    fclose(f);
    if (read_err) return NB200_ERR_IO;

    Reader r{buf.data(), buf.data() + buf.size()};
    const uint32_t version = r.u32();
    const uint32_t variant = r.u32();
    // unknown variant indices fail bincode deserialisation; is_version_valid
    // (data.rs:96-101) then requires V1 data under a V1 tag
    if (!r.ok || version != 0 || variant != 0) return NB200_ERR_FORMAT;
    bool ok = read_list(r, open_hash, open_ts_ns, n_open) && read_list(r, end_hash, end_ts_ns, n_end);
    if (ok) {
        *hash_duration_ns = r.duration();
        const uint64_t md5_len = r.u64();
        ok = r.ok && md5_len <= 32 && md5_len <= (uint64_t)(r.end - r.p);
        if (ok) {
            memcpy(md5_out, r.p, md5_len);
            md5_out[md5_len] = 0;
        }
    }
    if (!ok) {
        free(*open_hash);
        free(*open_ts_ns);
        free(*end_hash);
        free(*end_ts_ns);
        *open_hash = *end_hash = nullptr;
        *open_ts_ns = *end_ts_ns = nullptr;
        *n_open = *n_end = 0;
        return NB200_ERR_FORMAT;
    }
    return NB200_OK;
}

}  // extern "C"
