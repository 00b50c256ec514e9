"""Thin object layer over the C ABI: contexts, device-resident seasons, PCM
sets and run sets.  Everything here forwards to libneedle_b200.so; numpy is
only used to hold host buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import MatchParams, RUN_DTYPE, SearchResultC, check, lib, ptr

DELAY_NS = 2_600_000_000        # chromaprint get_delay(): 2600 ms
ITEM_NS = 123_000_000           # chromaprint get_item_duration(): 123 ms
HASH_DURATION_NS = 300_000_012  # Duration::from_secs_f32(0.3)  (DEFAULT_HASH_DURATION)


def match_params(threshold: int = 10, include_endings: bool = False,
                 min_opening_ns: int = 20_000_000_000, min_ending_ns: int = 20_000_000_000,
                 time_padding_ns: int = 0) -> MatchParams:
    return MatchParams(threshold, 1 if include_endings else 0, min_opening_ns, min_ending_ns,
                       time_padding_ns)


def _pairs_arg(pairs):
    if pairs is None:
        return None, None, 0
    a = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
    return a, ptr(a), a.shape[0]


class ResultList:
    """Per-video result tuples (present, has_opening, has_ending, opening_start_ns, opening_end_ns,
    ending_start_ns, ending_end_ns) of a search: a read-only sequence over a copy of the C result array
    that builds the Python tuples only when they are looked at.  (A Rust or C host reads the array in
    place; building 200 tuples eagerly costs this Python host 0.05-0.12 ms per search -- up to a seventh
    of an 8-GPU search step.)  Compares equal to a list of the same tuples."""
    __slots__ = ("_raw", "_n", "_list")

    def __init__(self, res, n):
        self._n = n
        self._raw = bytes(memoryview(res).cast("B")[:48 * n]) if n else b""
        self._list = None

    def tolist(self):
        if self._list is None:
            n = self._n
            if n == 0:
                self._list = []
            else:
                w = np.frombuffer(self._raw, dtype=np.uint32, count=12 * n).reshape(n, 12)
                t = np.frombuffer(self._raw, dtype=np.uint64, count=6 * n).reshape(n, 6)
                self._list = list(zip(w[:, 0].tolist(), w[:, 1].tolist(), w[:, 2].tolist(), t[:, 2].tolist(),
                                      t[:, 3].tolist(), t[:, 4].tolist(), t[:, 5].tolist()))
        return self._list

    def as_array(self) -> np.ndarray:
        """The records as a structured numpy array (_lib.RESULT_DTYPE)."""
        return np.frombuffer(self._raw, dtype=_lib.RESULT_DTYPE, count=self._n)

    def __len__(self):
        return self._n

    def __getitem__(self, k):
        return self.tolist()[k]

    def __iter__(self):
        return iter(self.tolist())

    def __eq__(self, other):
        if isinstance(other, ResultList):
            return self._raw == other._raw or self.tolist() == other.tolist()
        if isinstance(other, (list, tuple)):
            return self.tolist() == list(other)
        return NotImplemented

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    __hash__ = None

    def __repr__(self):
        return "ResultList(%r)" % (self.tolist(),)


def _results_list(res, n):
    return ResultList(res, n)


class Context:
    """nb200_ctx: one CUDA stream + scratch on one device."""

    def __init__(self, device: int = -1):
        self._h = C.c_void_p()
        check(lib().nb200_ctx_create(device, C.byref(self._h)), "nb200_ctx_create")

    def close(self):
        if self._h:
            lib().nb200_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream: int | None):
        check(lib().nb200_ctx_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None),
              "nb200_ctx_set_stream")

    OPT_FORCE_GENERAL_MATCH, OPT_K1_VARIANT, OPT_MATCH_DENSE, OPT_HOST_VOTE, OPT_DEFER_WAIT, OPT_MATCH_BAND_GROUP = 1, 2, 3, 4, 5, 6

    def set_option(self, option: int, value: int):
        check(lib().nb200_ctx_set_option(self._h, option, value), "nb200_ctx_set_option")

    HOST_PHASES = ("pcm_upload", "fingerprint", "lmin", "match_setup", "match", "simhash", "runs_download", "vote")

    def host_profile(self, reset: bool = True):
        """Wall ms per host phase since the last reset (nb200_ctx_host_profile)."""
        ms = (C.c_double * len(self.HOST_PHASES))()
        check(lib().nb200_ctx_host_profile(self._h, ms, 1 if reset else 0), "nb200_ctx_host_profile")
        return dict(zip(self.HOST_PHASES, list(ms)))

    def synchronize(self):
        check(lib().nb200_ctx_synchronize(self._h), "nb200_ctx_synchronize")

    def last_kernel_ms(self):
        ms = (C.c_float * 4)()
        n = C.c_uint64(0)
        check(lib().nb200_ctx_last_kernel_ms(self._h, ms, C.byref(n)), "nb200_ctx_last_kernel_ms")
        vote = C.c_float(0)
        check(lib().nb200_ctx_last_vote_ms(self._h, C.byref(vote)), "nb200_ctx_last_vote_ms")
        return {"fp_fft_chroma": ms[0], "fp_classify": ms[1], "match": ms[2], "simhash": ms[3],
                "vote": vote.value, "n_launches": n.value}

    # ---------------------------------------------------------- host-buffer calls

    def match_pairs(self, hashes, ts_ns, seg_offset, params: MatchParams, pairs=None) -> np.ndarray:
        """nb200_match_pairs: structured array of runs in the reference's push order."""
        h = np.ascontiguousarray(hashes, dtype=np.uint32)
        t = np.ascontiguousarray(ts_ns, dtype=np.uint64)
        off = np.ascontiguousarray(seg_offset, dtype=np.uint64)
        n_videos = (off.size - 1) // 2
        _keep, pp, npairs = _pairs_arg(pairs)
        out = C.POINTER(_lib.Run)()
        n = C.c_uint64(0)
        check(lib().nb200_match_pairs(self._h, ptr(h), ptr(t), ptr(off), n_videos, pp, npairs,
                                      C.byref(params), C.byref(out), C.byref(n)), "nb200_match_pairs")
        try:
            if n.value == 0:
                return np.zeros(0, dtype=RUN_DTYPE)
            buf = C.cast(out, C.POINTER(C.c_uint8 * (n.value * C.sizeof(_lib.Run)))).contents
            return np.frombuffer(buf, dtype=RUN_DTYPE).copy()
        finally:
            lib().nb200_free(out)

    def search(self, hashes, ts_ns, seg_offset, hash_duration_ns, params: MatchParams):
        """nb200_search: per-video result tuples (present, has_opening, has_ending,
        opening_start_ns, opening_end_ns, ending_start_ns, ending_end_ns)."""
        h = np.ascontiguousarray(hashes, dtype=np.uint32)
        t = np.ascontiguousarray(ts_ns, dtype=np.uint64)
        off = np.ascontiguousarray(seg_offset, dtype=np.uint64)
        hd = np.ascontiguousarray(hash_duration_ns, dtype=np.uint64)
        n_videos = (off.size - 1) // 2
        res = (SearchResultC * max(n_videos, 1))()
        check(lib().nb200_search(self._h, ptr(h), ptr(t), ptr(off), ptr(hd), n_videos, C.byref(params),
                                 res), "nb200_search")
        return _results_list(res, n_videos)

    def fingerprint_batch(self, segments, channels: int = 1, stride: int = 1):
        """nb200_fingerprint_batch: list of u32 arrays (raw sub-fingerprints, every stride-th)."""
        segs = [np.ascontiguousarray(s, dtype=np.int16).reshape(-1) for s in segments]
        n = len(segs)
        ptrs = (C.c_void_p * max(n, 1))(*[s.ctypes.data for s in segs])
        cnt = np.asarray([s.size for s in segs], dtype=np.uint64)
        outs = (C.c_void_p * max(n, 1))()
        out_counts = np.zeros(max(n, 1), np.uint64)
        check(lib().nb200_fingerprint_batch(self._h, ptrs, ptr(cnt), channels, n, stride, outs,
                                            ptr(out_counts)), "nb200_fingerprint_batch")
        res = []
        for k in range(n):
            c = int(out_counts[k])
            if c:
                a = np.ctypeslib.as_array(C.cast(outs[k], C.POINTER(C.c_uint32)), shape=(c,)).copy()
            else:
                a = np.zeros(0, np.uint32)
            res.append(a)
            lib().nb200_free(outs[k])
        return res

    def fingerprint_host_into(self, segments, d_hashes: int, d_ts_ns: int, capacity: int, stride: int = 2,
                              delay_ns: int = DELAY_NS, item_ns: int = ITEM_NS, seek_to_ns=None):
        """nb200_fingerprint_host_into: mono host PCM streamed in under K1, hashes and
        timestamps written into caller-owned device arrays."""
        segs = [np.ascontiguousarray(s, dtype=np.int16).reshape(-1) for s in segments]
        n = len(segs)
        ptrs = (C.c_void_p * max(n, 1))(*[s.ctypes.data for s in segs])
        cnt = np.asarray([s.size for s in segs], dtype=np.uint64)
        seek = None if seek_to_ns is None else np.ascontiguousarray(seek_to_ns, dtype=np.uint64)
        check(lib().nb200_fingerprint_host_into(self._h, ptrs, ptr(cnt), n, stride, delay_ns, item_ns, ptr(seek),
                                                C.c_void_p(d_hashes), C.c_void_p(d_ts_ns), capacity),
              "nb200_fingerprint_host_into")

    def timestamps_fill(self, d_ts_ns: int, dev_offset, seg_len, seek_to_ns, stride: int = 2,
                        delay_ns: int = DELAY_NS, item_ns: int = ITEM_NS):
        """nb200_timestamps_fill: timestamps of stored hashes from their index (asynchronous)."""
        do = np.ascontiguousarray(dev_offset, dtype=np.uint64)
        ln = np.ascontiguousarray(seg_len, dtype=np.uint64)
        sk = None if seek_to_ns is None else np.ascontiguousarray(seek_to_ns, dtype=np.uint64)
        check(lib().nb200_timestamps_fill(self._h, C.c_void_p(d_ts_ns), ptr(do), ptr(ln), ptr(sk), ln.size, stride,
                                          delay_ns, item_ns), "nb200_timestamps_fill")

    def vote_blocks(self, d_blocks: int, n_blocks: int, block_bytes: int, hash_duration_ns, params: MatchParams,
                    pairs=None, do_vote: bool = True):
        """nb200_vote_blocks -> (results or None, max_found, truncated).  One stream synchronisation."""
        hd = np.ascontiguousarray(hash_duration_ns, dtype=np.uint64)
        _keep, pp, npairs = _pairs_arg(pairs)
        res = (SearchResultC * max(hd.size, 1))()
        found = C.c_uint64(0)
        st = lib().nb200_vote_blocks(self._h, C.c_void_p(d_blocks), n_blocks, block_bytes, ptr(hd), hd.size, pp, npairs,
                                     C.byref(params), 1 if do_vote else 0, res, C.byref(found))
        if st == _lib.ERR_TOO_LARGE and found.value > (block_bytes - 64) // 64:
            return None, found.value, True
        check(st, "nb200_vote_blocks")
        return (_results_list(res, hd.size) if do_vote else None), found.value, False

    def analyze_search(self, segments, channels, seek_to_ns, hash_duration_ns, params: MatchParams):
        """nb200_analyze_search: segments = [opening_0, ending_0, opening_1, ...] PCM arrays."""
        segs = [np.ascontiguousarray(s, dtype=np.int16).reshape(-1) for s in segments]
        n = len(segs)
        assert n % 2 == 0
        ptrs = (C.c_void_p * max(n, 1))(*[s.ctypes.data for s in segs])
        cnt = np.asarray([s.size for s in segs], dtype=np.uint64)
        seek = None if seek_to_ns is None else np.ascontiguousarray(seek_to_ns, dtype=np.uint64)
        res = (SearchResultC * max(n // 2, 1))()
        check(lib().nb200_analyze_search(self._h, ptrs, ptr(cnt), channels, n // 2, ptr(seek),
                                         hash_duration_ns, C.byref(params), res), "nb200_analyze_search")
        return _results_list(res, n // 2)


def vote_array(hash_duration_ns, params: MatchParams, runs: np.ndarray, pairs=None, video_mask=None) -> np.ndarray:
    """nb200_vote / nb200_vote_subset (host only) -> structured array, one row per video."""
    hd = np.ascontiguousarray(hash_duration_ns, dtype=np.uint64)
    r = np.ascontiguousarray(runs, dtype=RUN_DTYPE)
    n_videos = hd.size
    _keep, pp, npairs = _pairs_arg(pairs)
    out = np.zeros(max(n_videos, 1), dtype=_lib.RESULT_DTYPE)
    res = C.cast(out.ctypes.data_as(C.c_void_p), C.POINTER(SearchResultC))
    mask = None if video_mask is None else np.ascontiguousarray(video_mask, dtype=np.uint8)
    check(lib().nb200_vote_subset(ptr(hd), n_videos, pp, npairs, C.byref(params), ptr(r), r.shape[0], ptr(mask),
                                  res), "nb200_vote_subset")
    return out[:n_videos]


def results_as_tuples(arr: np.ndarray):
    f = ("present", "has_opening", "has_ending", "opening_start_ns", "opening_end_ns", "ending_start_ns",
         "ending_end_ns")
    return list(zip(*[arr[k].tolist() for k in f])) if arr.shape[0] else []


def vote(hash_duration_ns, params: MatchParams, runs: np.ndarray, pairs=None):
    """nb200_vote (host only): runs as returned by match_pairs / RunSet.download;
    one hash duration per video."""
    hd = np.ascontiguousarray(hash_duration_ns, dtype=np.uint64)
    r = np.ascontiguousarray(runs, dtype=RUN_DTYPE)
    n_videos = hd.size
    _keep, pp, npairs = _pairs_arg(pairs)
    res = (SearchResultC * max(n_videos, 1))()
    check(lib().nb200_vote(ptr(hd), n_videos, pp, npairs, C.byref(params), ptr(r), r.shape[0], res),
          "nb200_vote")
    return _results_list(res, n_videos)


class HashSet:
    """nb200_hashset: a season of hashes + timestamps resident in HBM."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = handle

    @staticmethod
    def upload(ctx: Context, hashes, ts_ns, seg_offset) -> "HashSet":
        h = np.ascontiguousarray(hashes, dtype=np.uint32)
        t = np.ascontiguousarray(ts_ns, dtype=np.uint64)
        off = np.ascontiguousarray(seg_offset, dtype=np.uint64)
        out = C.c_void_p()
        check(lib().nb200_hashset_upload(ctx.handle, ptr(h), ptr(t), ptr(off), (off.size - 1) // 2,
                                         C.byref(out)), "nb200_hashset_upload")
        return HashSet(ctx, out)

    @staticmethod
    def from_device(ctx: Context, d_hashes: int, d_ts_ns: int, seg_offset) -> "HashSet":
        off = np.ascontiguousarray(seg_offset, dtype=np.uint64)
        out = C.c_void_p()
        check(lib().nb200_hashset_from_device(ctx.handle, C.c_void_p(d_hashes), C.c_void_p(d_ts_ns),
                                              ptr(off), (off.size - 1) // 2, C.byref(out)),
              "nb200_hashset_from_device")
        return HashSet(ctx, out)

    @staticmethod
    def from_device_scattered(ctx: Context, d_hashes: int, d_ts_ns: int, src_hash_off, src_ts_off,
                              seg_len) -> "HashSet":
        ho = np.ascontiguousarray(src_hash_off, dtype=np.uint64)
        to = np.ascontiguousarray(src_ts_off, dtype=np.uint64)
        ln = np.ascontiguousarray(seg_len, dtype=np.uint64)
        out = C.c_void_p()
        check(lib().nb200_hashset_from_device_scattered(ctx.handle, C.c_void_p(d_hashes), C.c_void_p(d_ts_ns),
                                                        ptr(ho), ptr(to), ptr(ln), ln.size // 2, C.byref(out)),
              "nb200_hashset_from_device_scattered")
        return HashSet(ctx, out)

    @staticmethod
    def view(ctx: Context, d_hashes: int, d_ts_ns: int, dev_offset, seg_len, keepalive=None) -> "HashSet":
        """nb200_hashset_view: a season over caller-owned device arrays (not copied)."""
        do = np.ascontiguousarray(dev_offset, dtype=np.uint64)
        ln = np.ascontiguousarray(seg_len, dtype=np.uint64)
        out = C.c_void_p()
        check(lib().nb200_hashset_view(ctx.handle, C.c_void_p(d_hashes), C.c_void_p(d_ts_ns), ptr(do), ptr(ln),
                                       ln.size // 2, C.byref(out)), "nb200_hashset_view")
        hs = HashSet(ctx, out)
        hs._keepalive = keepalive
        return hs

    def info(self):
        n = C.c_uint32(0)
        total = C.c_uint64(0)
        check(lib().nb200_hashset_info(self._h, C.byref(n), C.byref(total), None), "nb200_hashset_info")
        off = np.zeros(2 * n.value + 1, np.uint64)
        check(lib().nb200_hashset_info(self._h, None, None, ptr(off)), "nb200_hashset_info")
        return n.value, total.value, off

    def download(self):
        n, total, off = self.info()
        h = np.zeros(total, np.uint32)
        t = np.zeros(total, np.uint64)
        check(lib().nb200_hashset_download(self.ctx.handle, self._h, ptr(h), ptr(t)),
              "nb200_hashset_download")
        return h, t, off

    def export_packed(self, d_hashes: int, d_ts_ns: int):
        check(lib().nb200_hashset_export_packed(self.ctx.handle, self._h, C.c_void_p(d_hashes),
                                                C.c_void_p(d_ts_ns)), "nb200_hashset_export_packed")

    def match(self, params: MatchParams, pairs=None) -> "RunSet":
        _keep, pp, npairs = _pairs_arg(pairs)
        out = C.c_void_p()
        check(lib().nb200_match_run(self.ctx.handle, self._h, pp, npairs, C.byref(params), C.byref(out)),
              "nb200_match_run")
        return RunSet(self.ctx, out)

    def match_export(self, params: MatchParams, pairs, pair_base: int, d_block: int, block_bytes: int):
        """nb200_match_export: this slice of the pair list matched straight into a run block in
        device memory (asynchronous; pairs must be an explicit, possibly empty, list)."""
        a = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        keep = a if a.shape[0] else np.zeros((1, 2), np.uint32)     # an empty slice still needs a non-NULL list
        check(lib().nb200_match_export(self.ctx.handle, self._h, keep.ctypes.data_as(C.c_void_p), a.shape[0],
                                       C.byref(params), pair_base, C.c_void_p(d_block), block_bytes),
              "nb200_match_export")

    def search(self, hash_duration_ns, params: MatchParams):
        n, _, _ = self.info()
        hd = np.ascontiguousarray(hash_duration_ns, dtype=np.uint64)
        res = (SearchResultC * max(n, 1))()
        check(lib().nb200_search_hashset(self.ctx.handle, self._h, ptr(hd), C.byref(params), res),
              "nb200_search_hashset")
        return _results_list(res, n)

    def free(self):
        if self._h:
            lib().nb200_hashset_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def fingerprint_layout(n_mono_per_segment, stride: int = 2):
    """nb200_fingerprint_layout -> (dev_offset[], seg_len[], total) of the device layout."""
    n = np.ascontiguousarray(n_mono_per_segment, dtype=np.uint64)
    off = np.zeros(n.size, np.uint64)
    ln = np.zeros(n.size, np.uint64)
    total = C.c_uint64(0)
    check(lib().nb200_fingerprint_layout(ptr(n), n.size, stride, ptr(off), ptr(ln), C.byref(total)),
          "nb200_fingerprint_layout")
    return off, ln, total.value


class RunSet:
    """nb200_runset: match output in HBM."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = handle

    def count(self):
        n = C.c_uint64(0)
        cells = C.c_uint64(0)
        check(lib().nb200_runset_count(self._h, C.byref(n), C.byref(cells)), "nb200_runset_count")
        return n.value, cells.value

    def download(self) -> np.ndarray:
        n, _ = self.count()
        out = np.zeros(n, dtype=RUN_DTYPE)
        check(lib().nb200_runset_download(self.ctx.handle, self._h, ptr(out)), "nb200_runset_download")
        return out

    def free(self):
        if self._h:
            lib().nb200_runset_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PcmSet:
    """nb200_pcmset: decoded PCM segments in HBM (2 per video: opening, ending)."""

    def __init__(self, ctx: Context, handle, n_segments: int):
        self.ctx = ctx
        self._h = handle
        self.n_segments = n_segments

    @staticmethod
    def upload(ctx: Context, segments, channels: int = 1) -> "PcmSet":
        segs = [np.ascontiguousarray(s, dtype=np.int16).reshape(-1) for s in segments]
        n = len(segs)
        ptrs = (C.c_void_p * max(n, 1))(*[s.ctypes.data for s in segs])
        cnt = np.asarray([s.size for s in segs], dtype=np.uint64)
        out = C.c_void_p()
        check(lib().nb200_pcmset_upload(ctx.handle, ptrs, ptr(cnt), channels, n, C.byref(out)),
              "nb200_pcmset_upload")
        return PcmSet(ctx, out, n)

    @staticmethod
    def view(ctx: "Context", d_pcm: int, dev_offset, n_mono, capacity: int, keepalive=None) -> "PcmSet":
        """nb200_pcmset_view: mono S16 PCM that is already in device memory (caller-owned)."""
        off = np.ascontiguousarray(dev_offset, dtype=np.uint64)
        cnt = np.ascontiguousarray(n_mono, dtype=np.uint64)
        out = C.c_void_p()
        check(lib().nb200_pcmset_view(ctx.handle, C.c_void_p(d_pcm), ptr(off), ptr(cnt), cnt.size, capacity,
                                      C.byref(out)), "nb200_pcmset_view")
        ps = PcmSet(ctx, out, cnt.size)
        ps._keepalive = keepalive
        return ps

    def fingerprint(self, stride: int = 2, delay_ns: int = DELAY_NS, item_ns: int = ITEM_NS,
                    seek_to_ns=None) -> HashSet:
        seek = None if seek_to_ns is None else np.ascontiguousarray(seek_to_ns, dtype=np.uint64)
        out = C.c_void_p()
        check(lib().nb200_fingerprint_run(self.ctx.handle, self._h, stride, delay_ns, item_ns, ptr(seek),
                                          C.byref(out)), "nb200_fingerprint_run")
        return HashSet(self.ctx, out)

    def search(self, seek_to_ns, hash_duration_ns: int, params: MatchParams):
        """nb200_pcmset_search: fingerprint + match + vote on the resident PCM, one host wait."""
        seek = None if seek_to_ns is None else np.ascontiguousarray(seek_to_ns, dtype=np.uint64)
        n_videos = self.n_segments // 2
        res = (SearchResultC * max(n_videos, 1))()
        check(lib().nb200_pcmset_search(self.ctx.handle, self._h, ptr(seek), hash_duration_ns, C.byref(params), res),
              "nb200_pcmset_search")
        return _results_list(res, n_videos)

    def fingerprint_into(self, d_hashes: int, d_ts_ns: int, capacity: int, stride: int = 2,
                         delay_ns: int = DELAY_NS, item_ns: int = ITEM_NS, seek_to_ns=None):
        """nb200_fingerprint_run_into: hashes + timestamps straight into caller-owned device arrays."""
        seek = None if seek_to_ns is None else np.ascontiguousarray(seek_to_ns, dtype=np.uint64)
        check(lib().nb200_fingerprint_run_into(self.ctx.handle, self._h, stride, delay_ns, item_ns, ptr(seek),
                                               C.c_void_p(d_hashes), C.c_void_p(d_ts_ns), capacity),
              "nb200_fingerprint_run_into")

    def free(self):
        if self._h:
            lib().nb200_pcmset_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


@dataclass
class PinnedArray:
    """cudaHostAlloc'd numpy array (nb200_host_alloc)."""
    array: np.ndarray
    _ptr: int

    @staticmethod
    def empty(n: int, dtype) -> "PinnedArray":
        dt = np.dtype(dtype)
        p = C.c_void_p()
        check(lib().nb200_host_alloc(C.byref(p), max(n, 1) * dt.itemsize), "nb200_host_alloc")
        buf = (C.c_uint8 * (max(n, 1) * dt.itemsize)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dt)[:n]
        return PinnedArray(arr, p.value)

    def free(self):
        if self._ptr:
            self.array = None
            check(lib().nb200_host_free(C.c_void_p(self._ptr)), "nb200_host_free")
            self._ptr = 0


# ------------------------------------------------------------------ multi-GPU jobs (nb200_comm / nb200_mjob)

MJOB_PHASES = ("fingerprint", "hash_allgather", "match", "run_push_wait", "vote")


def comm_unique_id() -> bytes:
    """nb200_comm_unique_id: 128 bytes that rank 0 hands to every process of the job."""
    buf = (C.c_uint8 * 128)()
    check(lib().nb200_comm_unique_id(buf), "nb200_comm_unique_id")
    return bytes(buf)


class Comm:
    """nb200_comm: one rank of a multi-GPU job (a context + its NCCL communicator)."""

    def __init__(self, ctx: Context, handle, rank: int, world: int):
        self.ctx, self._h, self.rank, self.world = ctx, handle, rank, world

    @staticmethod
    def init_rank(ctx: Context, unique_id: bytes | None, rank: int, world: int) -> "Comm":
        """One process per GPU: every process calls this with the id rank 0 made."""
        out = C.c_void_p()
        idbuf = (C.c_uint8 * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        check(lib().nb200_comm_init_rank(ctx.handle, idbuf, rank, world, C.byref(out)), "nb200_comm_init_rank")
        return Comm(ctx, out, rank, world)

    @staticmethod
    def init_all(ctxs) -> list:
        """One process, N devices (ncclCommInitAll): the shape of needle itself."""
        n = len(ctxs)
        hs = (C.c_void_p * n)(*[c.handle for c in ctxs])
        out = (C.c_void_p * n)()
        check(lib().nb200_comm_init_all(hs, n, out), "nb200_comm_init_all")
        return [Comm(ctxs[r], C.c_void_p(out[r]), r, n) for r in range(n)]

    def destroy(self):
        if self._h:
            lib().nb200_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class MultiJob:
    """nb200_mjob: a search-only or analyze+search job over all ranks; `comms` = the ranks this
    process drives (one under torchrun, all of them in a single-process host)."""

    def __init__(self, comms, handle, n_videos: int, n_segments_pcm: int = 0):
        self.comms, self._h, self.n_videos = list(comms), handle, n_videos
        self.has_rank0 = any(c.rank == 0 for c in self.comms)
        self._n_seg = n_segments_pcm
        self._keep = None

    @staticmethod
    def _comm_array(comms):
        return (C.c_void_p * len(comms))(*[c._h for c in comms])

    @staticmethod
    def search(comms, hashes, ts_ns, seg_offset, hash_duration_ns, params: MatchParams, pairs=None) -> "MultiJob":
        h = np.ascontiguousarray(hashes, dtype=np.uint32)
        t = np.ascontiguousarray(ts_ns, dtype=np.uint64)
        off = np.ascontiguousarray(seg_offset, dtype=np.uint64)
        n_videos = (off.size - 1) // 2
        hd = np.ascontiguousarray(np.broadcast_to(np.asarray(hash_duration_ns, dtype=np.uint64), (n_videos,)))
        _keep, pp, npairs = _pairs_arg(pairs)
        out = C.c_void_p()
        check(lib().nb200_mjob_search_create(MultiJob._comm_array(comms), len(comms), ptr(h), ptr(t), ptr(off), n_videos,
                                             ptr(hd), pp, npairs, C.byref(params), C.byref(out)),
              "nb200_mjob_search_create")
        return MultiJob(comms, out, n_videos)

    @staticmethod
    def season(comms, n_mono_per_segment, seek_to_ns, hash_duration_ns: int, params: MatchParams,
               pairs=None) -> "MultiJob":
        nm = np.ascontiguousarray(n_mono_per_segment, dtype=np.uint64)
        n_videos = nm.size // 2
        sk = None if seek_to_ns is None else np.ascontiguousarray(seek_to_ns, dtype=np.uint64)
        _keep, pp, npairs = _pairs_arg(pairs)
        out = C.c_void_p()
        check(lib().nb200_mjob_season_create(MultiJob._comm_array(comms), len(comms), ptr(nm), ptr(sk), n_videos,
                                             int(hash_duration_ns), pp, npairs, C.byref(params), C.byref(out)),
              "nb200_mjob_season_create")
        return MultiJob(comms, out, n_videos, nm.size)

    def video_rank(self) -> np.ndarray:
        out = np.zeros(self.n_videos, np.uint32)
        check(lib().nb200_mjob_video_rank(self._h, ptr(out)), "nb200_mjob_video_rank")
        return out

    def _pcm_ptrs(self, segments):
        """segments: {global segment id: int16 array} (or a full list); others stay NULL."""
        arr = (C.c_void_p * max(self._n_seg, 1))()
        keep = []
        items = segments.items() if isinstance(segments, dict) else enumerate(segments)
        for s, x in items:
            if x is None:
                continue
            a = x if (isinstance(x, np.ndarray) and x.dtype == np.int16 and x.flags.c_contiguous) else \
                np.ascontiguousarray(x, dtype=np.int16)
            keep.append(a)
            arr[s] = a.ctypes.data
        return arr, keep

    def upload_pcm(self, segments):
        arr, keep = self._pcm_ptrs(segments)
        check(lib().nb200_mjob_upload_pcm(self._h, arr), "nb200_mjob_upload_pcm")

    def run(self, host_segments=None):
        """One pass.  Returns the per-video result tuples where rank 0 is local, else None."""
        arr, keep = (None, None) if host_segments is None else self._pcm_ptrs(host_segments)
        res = (SearchResultC * max(self.n_videos, 1))() if self.has_rank0 else None
        check(lib().nb200_mjob_run(self._h, arr, res), "nb200_mjob_run")
        return _results_list(res, self.n_videos) if self.has_rank0 else None

    def phase_ms(self) -> dict:
        ms = (C.c_float * len(MJOB_PHASES))()
        check(lib().nb200_mjob_phase_ms(self._h, ms), "nb200_mjob_phase_ms")
        return dict(zip(MJOB_PHASES, [float(x) for x in ms]))

    def free(self):
        if self._h:
            lib().nb200_mjob_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def plan_videos(n_mono_per_segment, world: int) -> np.ndarray:
    """nb200_plan_videos: rank that fingerprints each video in a multi-GPU season job (host only)."""
    nm = np.ascontiguousarray(n_mono_per_segment, dtype=np.uint64)
    out = np.zeros(nm.size // 2, np.uint32)
    check(lib().nb200_plan_videos(ptr(nm), nm.size // 2, world, ptr(out)), "nb200_plan_videos")
    return out


def plan_pairs(seg_len, world: int, include_endings: bool, pairs=None) -> np.ndarray:
    """nb200_plan_pairs: cuts[world + 1] of the pair list, slices balanced by table cells (host only)."""
    sl = np.ascontiguousarray(seg_len, dtype=np.uint64)
    _keep, pp, npairs = _pairs_arg(pairs)
    cuts = np.zeros(world + 1, np.uint64)
    check(lib().nb200_plan_pairs(ptr(sl), sl.size // 2, pp, npairs, world, 1 if include_endings else 0, ptr(cuts)),
          "nb200_plan_pairs")
    return cuts
