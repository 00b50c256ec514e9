"""Host-side mirror of needle's `needle::audio` API over the C ABI:
`Analyzer`, `Comparator`, `FrameHashes`, `SearchResult` with the reference's
names, builder methods, defaults, argument meaning and error behaviour
(needle/src/audio/{analyzer,comparator,data,mod}.rs), so that the parity tests
read like the reference's own usage (needle/src/lib.rs:20-100).

What is NOT here: FFmpeg.  In needle, demux/decode/swresample stay on the host
in Rust and hand 11025 Hz S16 PCM to the fingerprinter (analyzer.rs:179-187,
275).  This mirror reads that PCM from `.wav` files (S16, 11025 Hz, mono or
stereo) -- the stand-in for "a video whose audio has been decoded".  All
compute goes through libneedle_b200.so; there is no CPU fallback.
"""
from __future__ import annotations

import hashlib
import json
import os
import wave
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib, engine
from ._lib import Nb200Error, check, lib, ptr

# needle/src/audio/mod.rs:14-45
DEFAULT_HASH_MATCH_THRESHOLD = 10
DEFAULT_OPENING_SEARCH_PERCENTAGE = 0.50
DEFAULT_ENDING_SEARCH_PERCENTAGE = 0.25
DEFAULT_MIN_OPENING_DURATION = 20
DEFAULT_MIN_ENDING_DURATION = 20
DEFAULT_HASH_DURATION = 0.3
DEFAULT_OPENING_AND_ENDING_TIME_PADDING = 0.0
FRAME_HASH_DATA_FILE_NAME = "needle.dat"       # needle/src/lib.rs:154
SKIP_FILE_NAME = "needle.skip.json"            # needle/src/lib.rs:155
SAMPLE_RATE = 11025


class NeedleError(Exception):
    """needle::Error (needle/src/lib.rs:116-149)."""


class FrameHashDataNotFound(NeedleError):
    pass


class FrameHashDataInvalidVersion(NeedleError):
    pass


class FrameHashDataNoEnding(NeedleError):
    pass


class AnalyzerMissingPaths(NeedleError):
    pass


class PathNotFound(NeedleError):
    pass


def json_f32(x) -> str:
    """serde_json's rendering of an f32 (ryu): the shortest decimal that round-trips, always with a
    fraction or an exponent; exponents without padding (1e-5, 1e16)."""
    v = np.float32(x)
    s = np.format_float_positional(v, unique=True, trim="0") if (v == 0 or 1e-5 <= abs(float(v)) < 1e16) else \
        np.format_float_scientific(v, unique=True, trim="-", exp_digits=1).replace("e+", "e")
    return s


class ComparatorMinimumPaths(NeedleError):
    pass


# ----------------------------------------------------------- Duration helpers

def duration_from_secs_f32(x: float) -> int:
    """Duration::from_secs_f32 -> ns (exact value of the f32, ties to even)."""
    return int(np.rint(np.float64(np.float32(x)) * 1e9))


def duration_as_secs_f32(ns: int) -> float:
    secs, nanos = divmod(int(ns), 1_000_000_000)
    return float(np.float32(np.float32(secs) + np.float32(nanos) / np.float32(1e9)))


def duration_mul_f32(ns: int, rhs: float) -> int:
    return duration_from_secs_f32(np.float32(rhs) * np.float32(duration_as_secs_f32(ns)))


# -------------------------------------------------------------------- util.rs

def format_time(ns: int) -> str:
    """util::format_time: "MM:SSs" (needle/src/util.rs:8-12)."""
    s = int(ns) // 1_000_000_000
    return "%02d:%02ds" % (s // 60, s % 60)


def compute_header_md5sum(video: str) -> str:
    """md5 of the first 8 KiB; read_exact fails on shorter files (util.rs:99-105)."""
    with open(video, "rb") as f:
        buf = f.read(8192)
    if len(buf) < 8192:
        raise NeedleError("IO error: failed to fill whole buffer")
    return hashlib.md5(buf).hexdigest()


def is_valid_video_file(path: str) -> bool:
    """The stand-in for util::is_valid_video_file: a readable S16 11025 Hz WAV that
    is not a frame-hash file (util.rs:22-32)."""
    if path.endswith(FRAME_HASH_DATA_FILE_NAME):
        return False
    try:
        with wave.open(path, "rb") as w:
            return w.getsampwidth() == 2 and w.getframerate() == SAMPLE_RATE and w.getnchannels() in (1, 2)
    except Exception:
        return False


def find_video_files(paths) -> list[str]:
    """util::find_video_files: files as given, directories one level deep (util.rs:60-96)."""
    for p in paths:
        if not os.path.exists(p):
            raise PathNotFound("path does not exist: %r" % p)
    out = []
    for p in paths:
        if os.path.isdir(p):
            out += [os.path.join(p, e) for e in os.listdir(p) if is_valid_video_file(os.path.join(p, e))]
        elif is_valid_video_file(p):
            out.append(p)
    return out


def _with_extension(path: str, ext: str) -> str:
    """Path::with_extension: replaces everything after the last '.' of the file name."""
    d, name = os.path.split(path)
    stem = name.rsplit(".", 1)[0] if "." in name.lstrip(".") else name
    return os.path.join(d, stem + "." + ext)


def read_decoded_audio(path: str):
    """-> (interleaved i16 array, channels).  The host decoder's output."""
    with wave.open(path, "rb") as w:
        if w.getsampwidth() != 2 or w.getframerate() != SAMPLE_RATE or w.getnchannels() not in (1, 2):
            raise NeedleError("FFmpeg error: unsupported stream (want S16 %d Hz mono/stereo)" % SAMPLE_RATE)
        ch = w.getnchannels()
        data = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.int16)
    return data, ch


# ------------------------------------------------------------------- data.rs

@dataclass
class FrameHashes:
    """FrameHashes V1 (needle/src/audio/data.rs:20-26,74-80): (hash, timestamp)
    lists for the opening and ending segments, SoA here."""
    opening_hashes: np.ndarray
    opening_ts_ns: np.ndarray
    ending_hashes: np.ndarray
    ending_ts_ns: np.ndarray
    hash_duration_ns: int
    md5: str = ""

    def opening_data(self):
        return self.opening_hashes, self.opening_ts_ns

    def ending_data(self):
        return self.ending_hashes, self.ending_ts_ns

    def hash_duration(self) -> int:
        return self.hash_duration_ns

    def save(self, path: str):
        oh = np.ascontiguousarray(self.opening_hashes, np.uint32)
        ot = np.ascontiguousarray(self.opening_ts_ns, np.uint64)
        eh = np.ascontiguousarray(self.ending_hashes, np.uint32)
        et = np.ascontiguousarray(self.ending_ts_ns, np.uint64)
        check(lib().nb200_framehashes_write(path.encode(), ptr(oh), ptr(ot), oh.size, ptr(eh), ptr(et), eh.size,
                                            int(self.hash_duration_ns), self.md5.encode()),
              "nb200_framehashes_write")

    @staticmethod
    def from_path(path: str) -> "FrameHashes":
        import ctypes as C
        if not os.path.exists(path):
            raise FrameHashDataNotFound("frame hash data not found at: %r" % path)
        L = lib()
        oh, ot, eh, et = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        no, ne, hd = C.c_uint64(), C.c_uint64(), C.c_uint64()
        md5 = C.create_string_buffer(64)
        st = L.nb200_framehashes_read(path.encode(), C.byref(oh), C.byref(ot), C.byref(no), C.byref(eh),
                                      C.byref(et), C.byref(ne), C.byref(hd), md5)
        if st == _lib.ERR_FORMAT:
            raise FrameHashDataInvalidVersion("invalid frame hash data version")
        check(st, "nb200_framehashes_read")

        def arr(p, n, t, dt):
            a = np.ctypeslib.as_array(C.cast(p, C.POINTER(t)), shape=(max(n, 1),))[:n].astype(dt)
            L.nb200_free(p)
            return a
        return FrameHashes(arr(oh, no.value, C.c_uint32, np.uint32), arr(ot, no.value, C.c_uint64, np.uint64),
                           arr(eh, ne.value, C.c_uint32, np.uint32), arr(et, ne.value, C.c_uint64, np.uint64),
                           hd.value, md5.value.decode())

    @staticmethod
    def from_video(video: str, analyze: bool, ctx: Optional[engine.Context] = None) -> "FrameHashes":
        """data.rs:124-141: load `<video>.needle.dat`, or analyze in place with
        Analyzer::default().with_force(true) and the default hash duration."""
        if not analyze:
            return FrameHashes.from_path(_with_extension(video, FRAME_HASH_DATA_FILE_NAME))
        a = Analyzer.from_files([video], False, True, ctx=ctx)
        return a.run_single(video, duration_from_secs_f32(DEFAULT_HASH_DURATION), False)


@dataclass
class SearchResult:
    """comparator.rs:65-69: optional (start, end) in ns."""
    opening: Optional[tuple] = None
    ending: Optional[tuple] = None


# --------------------------------------------------------------- analyzer.rs

class Analyzer:
    """needle::audio::Analyzer (analyzer.rs:86-456)."""

    def __init__(self, ctx: Optional[engine.Context] = None):
        self.videos: list[str] = []
        self.opening_search_percentage = DEFAULT_OPENING_SEARCH_PERCENTAGE
        self.ending_search_percentage = DEFAULT_ENDING_SEARCH_PERCENTAGE
        self.include_endings = False
        self.threaded_decoding = False
        self.force = False
        self._ctx = ctx

    @staticmethod
    def from_files(videos, threaded_decoding: bool = False, force: bool = False,
                   ctx: Optional[engine.Context] = None) -> "Analyzer":
        a = Analyzer(ctx)
        a.videos = list(videos)
        a.threaded_decoding = threaded_decoding
        a.force = force
        return a

    def with_opening_search_percentage(self, v: float):
        self.opening_search_percentage = v
        return self

    def with_ending_search_percentage(self, v: float):
        self.ending_search_percentage = v
        return self

    def with_include_endings(self, v: bool):
        self.include_endings = v
        return self

    def with_threaded_decoding(self, v: bool):
        self.threaded_decoding = v
        return self

    def with_force(self, v: bool):
        self.force = v
        return self

    def _context(self) -> engine.Context:
        if self._ctx is None:
            self._ctx = engine.Context(-1)
        return self._ctx

    def _segments(self, path: str):
        """What process_frames is asked to hash (analyzer.rs:378-402): the first
        opening_search_percentage of the stream, and from 1 - ending_search_percentage on."""
        data, ch = read_decoded_audio(path)
        n = data.size // ch
        dur_ns = (n * 1_000_000_000) // SAMPLE_RATE
        open_ns = duration_mul_f32(dur_ns, self.opening_search_percentage)
        n_open = min(n, open_ns * SAMPLE_RATE // 1_000_000_000)
        opening = data[:n_open * ch]
        if not self.include_endings:
            return opening, np.zeros(0, np.int16), 0, ch
        seek_ns = duration_mul_f32(dur_ns, np.float32(1.0) - np.float32(self.ending_search_percentage))
        s_end = min(n, seek_ns * SAMPLE_RATE // 1_000_000_000)
        return opening, data[s_end * ch:], seek_ns, ch

    def _cached(self, path: str, md5: str):
        """analyzer.rs:338-348: reuse `<video>.needle.dat` if its md5 matches."""
        dat = _with_extension(path, FRAME_HASH_DATA_FILE_NAME)
        if self.force or not os.path.exists(dat):
            return None
        data = FrameHashes.from_path(dat)
        if data.md5 == md5:
            print("Skipping analysis for %s..." % path)
            return data
        return None

    def run_single(self, path: str, hash_duration_ns: int, persist: bool) -> FrameHashes:
        return self._run_many([path], hash_duration_ns, persist)[0]

    def run(self, hash_duration_ns: int, persist: bool, threading: bool = True) -> list:
        """Analyzer::run (analyzer.rs:425-456).  `threading` selected rayon in the
        reference; here every video that needs analysis goes to the GPU in one batch."""
        if len(self.videos) == 0:
            raise AnalyzerMissingPaths("no paths provided to analyzer")
        if hash_duration_ns <= 0:
            raise NeedleError("invalid hash duration")      # NeedleError_AnalyzerInvalidHashDuration (capi :474)
        return self._run_many(self.videos, hash_duration_ns, persist)

    def _run_many(self, paths, hash_duration_ns: int, persist: bool) -> list:
        # step_by = hash_duration.as_millis() / item_duration.as_millis()   (analyzer.rs:293-297)
        stride = (hash_duration_ns // 1_000_000) // 123
        if stride == 0:
            raise NeedleError("hash duration shorter than one fingerprint item (step_by(0) panics in the reference)")
        out: list = [None] * len(paths)
        todo, segs, seeks, md5s, channels = [], [], [], {}, None
        for k, path in enumerate(paths):
            md5s[k] = compute_header_md5sum(path)
            cached = self._cached(path, md5s[k])
            if cached is not None:
                out[k] = cached
                continue
            o, e, seek_ns, ch = self._segments(path)
            if channels is None:
                channels = ch
            if ch != channels:       # mixed layouts: bring everything to stereo like needle's resampler does
                raise NeedleError("all videos of one run must have the same channel count here")
            todo.append(k)
            segs += [o, e]
            seeks += [0, seek_ns]
        if todo:
            ctx = self._context()
            ps = engine.PcmSet.upload(ctx, segs, channels=channels)
            hs = ps.fingerprint(stride=stride, seek_to_ns=seeks)
            h, t, off = hs.download()
            hs.free()
            ps.free()
            for j, k in enumerate(todo):
                a, b, c = int(off[2 * j]), int(off[2 * j + 1]), int(off[2 * j + 2])
                fh = FrameHashes(h[a:b].copy(), t[a:b].copy(), h[b:c].copy(), t[b:c].copy(),
                                 hash_duration_ns, md5s[k])
                if persist:
                    fh.save(_with_extension(paths[k], FRAME_HASH_DATA_FILE_NAME))
                out[k] = fh
        return out


# ------------------------------------------------------------- comparator.rs

class Comparator:
    """needle::audio::Comparator (comparator.rs:74-664)."""

    def __init__(self, ctx: Optional[engine.Context] = None):
        self.videos: list[str] = []
        self.include_endings = False
        self.hash_match_threshold = DEFAULT_HASH_MATCH_THRESHOLD
        self.min_opening_duration_ns = DEFAULT_MIN_OPENING_DURATION * 1_000_000_000
        self.min_ending_duration_ns = DEFAULT_MIN_ENDING_DURATION * 1_000_000_000
        self.time_padding_ns = 0
        self._ctx = ctx

    @staticmethod
    def from_files(videos, ctx: Optional[engine.Context] = None) -> "Comparator":
        c = Comparator(ctx)
        c.videos = list(videos)
        return c

    @staticmethod
    def from_analyzer(analyzer: Analyzer) -> "Comparator":
        c = Comparator(analyzer._ctx)
        c.videos = list(analyzer.videos)
        return c

    def with_include_endings(self, v: bool):
        self.include_endings = v
        return self

    def with_hash_match_threshold(self, v: int):
        self.hash_match_threshold = v
        return self

    def with_min_opening_duration(self, ns: int):
        self.min_opening_duration_ns = ns
        return self

    def with_min_ending_duration(self, ns: int):
        self.min_ending_duration_ns = ns
        return self

    def with_time_padding(self, ns: int):
        self.time_padding_ns = ns
        return self

    def _context(self) -> engine.Context:
        if self._ctx is None:
            self._ctx = engine.Context(-1)
        return self._ctx

    def _params(self):
        return engine.match_params(self.hash_match_threshold, self.include_endings, self.min_opening_duration_ns,
                                   self.min_ending_duration_ns, self.time_padding_ns)

    # ---- skip files (comparator.rs:310-354, data.rs:8-13)
    @staticmethod
    def check_skip_file(video: str) -> bool:
        skip = _with_extension(video, SKIP_FILE_NAME)
        if not os.path.exists(skip):
            return False
        md5 = compute_header_md5sum(video)
        with open(skip) as f:
            return json.load(f)["md5"] == md5

    @staticmethod
    def create_skip_file(video: str, result: SearchResult):
        to_f32 = lambda r: None if r is None else [duration_as_secs_f32(r[0]), duration_as_secs_f32(r[1])]
        opening, ending = to_f32(result.opening), to_f32(result.ending)
        if opening is None and ending is None:
            return
        md5 = compute_header_md5sum(video)
        # the bytes serde_json writes (comparator.rs:329-354): f32 values in their shortest form
        pair = lambda r: "null" if r is None else "[%s,%s]" % (json_f32(r[0]), json_f32(r[1]))
        with open(_with_extension(video, SKIP_FILE_NAME), "w") as f:
            f.write('{"opening":%s,"ending":%s,"md5":"%s"}' % (pair(opening), pair(ending), md5))

    def display_opening_ending_info(self, r: SearchResult):
        if r.opening is not None:
            print('* Opening - "%s"-"%s"' % (format_time(r.opening[0]), format_time(r.opening[1])))
        else:
            print("* Opening - N/A")
        if self.include_endings:
            if r.ending is not None:
                print('* Ending - "%s"-"%s"' % (format_time(r.ending[0]), format_time(r.ending[1])))
            else:
                print("* Ending - N/A")

    def run_with_frame_hashes(self, frame_hashes, display: bool = False, use_skip_files: bool = False,
                              write_skip_files: bool = False, threading: bool = True) -> list:
        """comparator.rs:524-629.  Returns the reference's Vec<SearchResult>: only
        videos with a result and not skipped, in video order."""
        n = len(frame_hashes)
        if n < 2:
            raise ComparatorMinimumPaths("comparator needs at least two videos")
        hs, ts, off = [], [], [0]
        for f in frame_hashes:
            for h, t in (f.opening_data(), f.ending_data()):
                hs.append(np.asarray(h, np.uint32))
                ts.append(np.asarray(t, np.uint64))
                off.append(off[-1] + len(h))
        hd = np.array([f.hash_duration() for f in frame_hashes], dtype=np.uint64)
        try:
            res = self._context().search(np.concatenate(hs), np.concatenate(ts), np.asarray(off, np.uint64), hd,
                                         self._params())
        except Nb200Error as e:
            if e.status == _lib.ERR_NO_ENDING:
                raise FrameHashDataNoEnding("no ending hash data present") from e
            raise
        results = []
        for idx, r in enumerate(res):
            path = self.videos[idx] if idx < len(self.videos) else str(idx)
            if display:
                print("\n%s\n" % path)
            if use_skip_files and self.check_skip_file(path):
                if display:
                    print("Skipping due to existing skip file...")
                continue
            if not r[0]:
                if display:
                    print("No opening or ending found." if self.include_endings else "No opening found.")
                continue
            sr = SearchResult((r[3], r[4]) if r[1] else None, (r[5], r[6]) if r[2] else None)
            if display:
                self.display_opening_ending_info(sr)
            if write_skip_files:
                self.create_skip_file(path, sr)
            results.append(sr)
        return results

    def run(self, analyze: bool, display: bool = False, use_skip_files: bool = False,
            write_skip_files: bool = False, threading: bool = True) -> list:
        """comparator.rs:637-663.  `analyze` runs Analyzer::default() in place, i.e.
        WITHOUT endings (data.rs:134): combined with include_endings this is the
        reference's FrameHashDataNoEnding failure."""
        if analyze:
            a = Analyzer.from_files(self.videos, False, True, ctx=self._context())
            frame_hashes = a._run_many(self.videos, duration_from_secs_f32(DEFAULT_HASH_DURATION), False)
        else:
            frame_hashes = [FrameHashes.from_video(v, False) for v in self.videos]
        return self.run_with_frame_hashes(frame_hashes, display, use_skip_files, write_skip_files, threading)
