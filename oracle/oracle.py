"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/needle_oracle.h.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never by needle_b200/.  Parity unpinned (the reference has no golden vectors
for this path; see the header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (make -C oracle)."""
    srcs = [os.path.join(_HERE, f) for f in
            ("match_ref.c", "chromaprint_ref.c", "needle_oracle.h", "chromaprint_tables.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


class Entry(C.Structure):
    _fields_ = [
        ("score", C.c_uint64),
        ("src_start_ns", C.c_uint64), ("src_end_ns", C.c_uint64),
        ("dst_start_ns", C.c_uint64), ("dst_end_ns", C.c_uint64),
        ("src_match_hash", C.c_uint32), ("dst_match_hash", C.c_uint32),
        ("is_ending", C.c_uint32),
        ("src_hash_duration_ns", C.c_uint64), ("dst_hash_duration_ns", C.c_uint64),
        ("i_end", C.c_uint32), ("j_end", C.c_uint32),
    ]

    def astuple(self):
        return (self.score, self.src_start_ns, self.src_end_ns, self.dst_start_ns, self.dst_end_ns,
                self.src_match_hash, self.dst_match_hash, self.is_ending,
                self.src_hash_duration_ns, self.dst_hash_duration_ns, self.i_end, self.j_end)


class _Season(C.Structure):
    _fields_ = [("n_videos", C.c_size_t), ("hashes", C.c_void_p), ("ts_ns", C.c_void_p),
                ("seg_offset", C.c_void_p), ("hash_duration_ns", C.c_void_p)]


class _Params(C.Structure):
    _fields_ = [("hash_match_threshold", C.c_uint32), ("min_opening_ns", C.c_uint64),
                ("min_ending_ns", C.c_uint64), ("time_padding_ns", C.c_uint64),
                ("include_endings", C.c_int)]


class Result(C.Structure):
    _fields_ = [("present", C.c_uint32), ("has_opening", C.c_uint32), ("has_ending", C.c_uint32),
                ("opening_start_ns", C.c_uint64), ("opening_end_ns", C.c_uint64),
                ("ending_start_ns", C.c_uint64), ("ending_end_ns", C.c_uint64)]

    def astuple(self):
        return (self.present, self.has_opening, self.has_ending, self.opening_start_ns,
                self.opening_end_ns, self.ending_start_ns, self.ending_end_ns)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_duration_from_secs_f32.restype = C.c_uint64
        L.orc_duration_from_secs_f32.argtypes = [C.c_float]
        L.orc_duration_as_secs_f32.restype = C.c_float
        L.orc_duration_as_secs_f32.argtypes = [C.c_uint64]
        L.orc_duration_mul_f32.restype = C.c_uint64
        L.orc_duration_mul_f32.argtypes = [C.c_uint64, C.c_float]
        L.orc_hash_timestamp.restype = C.c_uint64
        L.orc_hash_timestamp.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64]
        L.orc_simhash32.restype = C.c_uint32
        L.orc_simhash32.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_longest_common_hash_match.restype = C.c_int64
        L.orc_longest_common_hash_match.argtypes = [
            C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32,
            C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.POINTER(Entry))]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_run_with_frame_hashes.restype = C.c_int
        L.orc_run_with_frame_hashes.argtypes = [
            C.POINTER(_Season), C.POINTER(_Params), C.c_int, C.POINTER(Result),
            C.POINTER(C.POINTER(Entry)), C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_uint64)]
        L.orc_num_frames.restype = C.c_size_t
        L.orc_num_frames.argtypes = [C.c_size_t]
        L.orc_num_raw_hashes.restype = C.c_size_t
        L.orc_num_raw_hashes.argtypes = [C.c_size_t]
        L.orc_fingerprint.restype = C.c_int64
        L.orc_fingerprint.argtypes = [C.c_void_p, C.c_size_t, C.c_int,
                                      C.POINTER(C.POINTER(C.c_uint32)),
                                      C.POINTER(C.POINTER(C.c_double))]
        L.orc_subsample_and_stamp.restype = C.c_size_t
        L.orc_subsample_and_stamp.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64,
                                              C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_chroma_notes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                       C.POINTER(C.c_int), C.c_void_p]
        L.orc_chroma_fold.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_chroma_filter.restype = C.c_size_t
        L.orc_chroma_filter.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_normalize.argtypes = [C.c_void_p, C.c_double]
        L.orc_quantize.restype = C.c_int
        L.orc_quantize.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_gray_code.restype = C.c_int
        L.orc_gray_code.argtypes = [C.c_int]
        L.orc_filter_apply.restype = C.c_double
        L.orc_filter_apply.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                       C.c_size_t, C.c_size_t]
        L.orc_power_spectrum.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_fingerprint_many.restype = C.c_int
        L.orc_fingerprint_many.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int,
                                           C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- durations

def duration_from_secs_f32(x: float) -> int:
    return lib().orc_duration_from_secs_f32(float(np.float32(x)))


def duration_as_secs_f32(ns: int) -> float:
    return lib().orc_duration_as_secs_f32(int(ns))


def duration_mul_f32(ns: int, rhs: float) -> int:
    return lib().orc_duration_mul_f32(int(ns), float(np.float32(rhs)))


def hash_timestamp(delay_ns: int, item_ns: int, raw_index: int, seek_to_ns: int = 0) -> int:
    return lib().orc_hash_timestamp(delay_ns, item_ns, raw_index, seek_to_ns)


# -------------------------------------------------------------------- match

def simhash32(hashes: np.ndarray) -> int:
    h = np.ascontiguousarray(hashes, dtype=np.uint32)
    return lib().orc_simhash32(_p(h), h.size)


def longest_common_hash_match(src_hash, src_ts, dst_hash, dst_ts, threshold=10,
                              min_opening_ns=20_000_000_000, min_ending_ns=20_000_000_000,
                              src_hash_duration_ns=300_000_012, dst_hash_duration_ns=300_000_012,
                              is_opening=True):
    """comparator.rs:157-250.  Returns the heap array as a list of tuples
    (see Entry.astuple), or raises OverflowError where Rust would panic."""
    sh = np.ascontiguousarray(src_hash, dtype=np.uint32)
    st = np.ascontiguousarray(src_ts, dtype=np.uint64)
    dh = np.ascontiguousarray(dst_hash, dtype=np.uint32)
    dt = np.ascontiguousarray(dst_ts, dtype=np.uint64)
    out = C.POINTER(Entry)()
    n = lib().orc_longest_common_hash_match(_p(sh), _p(st), sh.size, _p(dh), _p(dt), dh.size,
                                            threshold, min_opening_ns, min_ending_ns,
                                            src_hash_duration_ns, dst_hash_duration_ns,
                                            1 if is_opening else 0, C.byref(out))
    if n < 0:
        raise OverflowError("Duration subtraction underflow (the reference panics)")
    res = [out[k].astuple() for k in range(n)]
    lib().orc_free(out)
    return res


@dataclass
class Season:
    """N videos; per video an opening and an ending list of (hash, ts_ns)."""
    hashes: np.ndarray          # u32, concatenated
    ts_ns: np.ndarray           # u64
    seg_offset: np.ndarray      # u64 [2N+1]
    hash_duration_ns: np.ndarray  # u64 [N]

    @property
    def n_videos(self) -> int:
        return (self.seg_offset.size - 1) // 2

    def segment(self, video: int, ending: bool):
        a = int(self.seg_offset[2 * video + (1 if ending else 0)])
        b = int(self.seg_offset[2 * video + (1 if ending else 0) + 1])
        return self.hashes[a:b], self.ts_ns[a:b]

    @staticmethod
    def from_lists(openings, endings, hash_duration_ns=300_000_012):
        """openings/endings: per video (hash array, ts array)."""
        hs, ts, off = [], [], [0]
        for (oh, ot), (eh, et) in zip(openings, endings):
            hs += [np.asarray(oh, dtype=np.uint32), np.asarray(eh, dtype=np.uint32)]
            ts += [np.asarray(ot, dtype=np.uint64), np.asarray(et, dtype=np.uint64)]
            off.append(off[-1] + len(oh))
            off.append(off[-1] + len(eh))
        n = len(openings)
        hd = np.full(n, hash_duration_ns, dtype=np.uint64) if np.isscalar(hash_duration_ns) \
            else np.asarray(hash_duration_ns, dtype=np.uint64)
        return Season(np.concatenate(hs) if hs else np.zeros(0, np.uint32),
                      np.concatenate(ts) if ts else np.zeros(0, np.uint64),
                      np.asarray(off, dtype=np.uint64), hd)


def run_with_frame_hashes(season: Season, threshold=10, min_opening_ns=20_000_000_000,
                          min_ending_ns=20_000_000_000, time_padding_ns=0, include_endings=False,
                          n_threads=1, want_entries=False):
    """comparator.rs:524-629.  Returns (status, [Result tuples per video], entries)
    where entries = [(pair_index, Entry tuple)] in reference order."""
    N = season.n_videos
    hashes = np.ascontiguousarray(season.hashes, dtype=np.uint32)
    ts = np.ascontiguousarray(season.ts_ns, dtype=np.uint64)
    off = np.ascontiguousarray(season.seg_offset, dtype=np.uint64)
    hd = np.ascontiguousarray(season.hash_duration_ns, dtype=np.uint64)
    s = _Season(N, _p(hashes), _p(ts), _p(off), _p(hd))
    p = _Params(threshold, min_opening_ns, min_ending_ns, time_padding_ns, 1 if include_endings else 0)
    results = (Result * max(N, 1))()
    ent = C.POINTER(Entry)()
    pidx = C.POINTER(C.c_uint32)()
    n_ent = C.c_uint64(0)
    st = lib().orc_run_with_frame_hashes(
        C.byref(s), C.byref(p), n_threads, results,
        C.byref(ent) if want_entries else None, C.byref(pidx) if want_entries else None,
        C.byref(n_ent) if want_entries else None)
    entries = None
    if want_entries:
        entries = [(pidx[k], ent[k].astuple()) for k in range(n_ent.value)]
        lib().orc_free(ent)
        lib().orc_free(pidx)
    return st, [results[v].astuple() for v in range(N)], entries


# -------------------------------------------------------------- fingerprint

def num_frames(n_mono: int) -> int:
    return lib().orc_num_frames(n_mono)


def num_raw_hashes(n_mono: int) -> int:
    return lib().orc_num_raw_hashes(n_mono)


def fingerprint(pcm: np.ndarray, channels: int = 1, want_chroma: bool = False):
    """Raw Chromaprint TEST2 sub-fingerprints of interleaved i16 PCM @ 11025 Hz."""
    x = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
    out = C.POINTER(C.c_uint32)()
    chroma = C.POINTER(C.c_double)()
    n = lib().orc_fingerprint(_p(x), x.size, channels, C.byref(out),
                              C.byref(chroma) if want_chroma else None)
    if n < 0:
        raise ValueError("bad channels / sample count")
    h = np.ctypeslib.as_array(out, shape=(max(n, 1),))[:n].copy()
    lib().orc_free(out)
    if want_chroma:
        nf = num_frames(x.size // channels)
        c = np.ctypeslib.as_array(chroma, shape=(max(nf, 1) * 12,))[:nf * 12].copy().reshape(nf, 12)
        lib().orc_free(chroma)
        return h, c
    return h


def subsample_and_stamp(raw: np.ndarray, step_by: int = 2, delay_ns: int = 2_600_000_000,
                        item_ns: int = 123_000_000, seek_to_ns: int = 0):
    raw = np.ascontiguousarray(raw, dtype=np.uint32)
    n = (raw.size + step_by - 1) // step_by
    oh = np.zeros(n, np.uint32)
    ot = np.zeros(n, np.uint64)
    c = lib().orc_subsample_and_stamp(_p(raw), raw.size, step_by, delay_ns, item_ns, seek_to_ns,
                                      _p(oh), _p(ot))
    assert c == n
    return oh, ot


def fingerprint_many(segments, channels: int = 1, n_threads: int = 1):
    """One segment per worker thread (analyzer.rs:440-444).  Returns list of u32 arrays."""
    segs = [np.ascontiguousarray(s, dtype=np.int16).reshape(-1) for s in segments]
    n = len(segs)
    ptrs = (C.c_void_p * n)(*[s.ctypes.data for s in segs])
    cnts = np.asarray([s.size for s in segs], dtype=np.uint64)
    outs = (C.POINTER(C.c_uint32) * n)()
    out_counts = np.zeros(n, np.uint64)
    st = lib().orc_fingerprint_many(ptrs, _p(cnts), channels, n, n_threads, outs, _p(out_counts))
    if st != 0:
        raise ValueError("orc_fingerprint_many failed")
    res = []
    for k in range(n):
        c = int(out_counts[k])
        res.append(np.ctypeslib.as_array(outs[k], shape=(max(c, 1),))[:c].copy())
        lib().orc_free(outs[k])
    return res


def power_spectrum(frame: np.ndarray) -> np.ndarray:
    f = np.ascontiguousarray(frame, dtype=np.int16)
    assert f.size == 4096
    out = np.zeros(2049, np.float64)
    lib().orc_power_spectrum(_p(f), _p(out))
    return out


def chroma_notes(min_freq=28, max_freq=3520, frame_size=4096, sample_rate=11025):
    notes = np.zeros(frame_size, np.int8)
    lo, hi = C.c_int(), C.c_int()
    lib().orc_chroma_notes(min_freq, max_freq, frame_size, sample_rate, C.byref(lo), C.byref(hi), _p(notes))
    return lo.value, hi.value, notes


def chroma_fold(power: np.ndarray, frame_size, min_freq, max_freq, sample_rate) -> np.ndarray:
    p = np.ascontiguousarray(power, dtype=np.float64)
    out = np.zeros(12, np.float64)
    lib().orc_chroma_fold(_p(p), frame_size, min_freq, max_freq, sample_rate, _p(out))
    return out


def chroma_filter(coeffs, rows: np.ndarray) -> np.ndarray:
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    r = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, 12)
    out = np.zeros((max(r.shape[0] - c.size + 1, 0), 12), np.float64)
    n = lib().orc_chroma_filter(_p(c), c.size, _p(r), r.shape[0], _p(out))
    return out[:n]


def normalize(features: np.ndarray, threshold: float = 0.01) -> np.ndarray:
    f = np.array(features, dtype=np.float64).copy()
    lib().orc_normalize(_p(f), threshold)
    return f


def quantize(v, t0, t1, t2) -> int:
    return lib().orc_quantize(v, t0, t1, t2)


def gray_code(i: int) -> int:
    return lib().orc_gray_code(i)


def filter_apply(ftype, y, height, width, image: np.ndarray, x: int) -> float:
    im = np.ascontiguousarray(image, dtype=np.float64)
    return lib().orc_filter_apply(ftype, y, height, width, _p(im), im.shape[0], im.shape[1], x)
