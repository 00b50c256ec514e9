"""Multi-GPU flow: one process per GPU (torch.distributed; NCCL over NVLink on
the box, gloo in CPU tests).

  episodes --shard--> fingerprint (K1/K2)        no collective
  hashes   --ONE all-gather--> every rank holds the whole season (KB..MB)
  pairs    --shard--> match (K3/K4)              no collective
  runs     --gather--> rank 0, host vote (find_best_match)

which is how the reference's two rayon par_iter loops (per video,
needle/src/audio/analyzer.rs:437-445; per pair, comparator.rs:549-564) map to
ranks.  Everything that decides *what* each rank does (plans, offsets, run
merging) is plain host code here and is exercised with world_size-2 gloo tests;
the compute itself is behind the `Backend` protocol -- the product backend is
`GpuBackend` (libneedle_b200.so), tests plug a checker in.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ._lib import RESULT_DTYPE, RUN_DTYPE

FRAME, HOP, WARMUP = 4096, 1365, 19


def pair_list(n_videos: int) -> np.ndarray:
    """All (i < j) in the reference's order (comparator.rs:534-545)."""
    i, j = np.triu_indices(n_videos, 1)
    return np.stack([i, j], axis=1).astype(np.uint32)


def stored_hashes(n_mono: int, stride: int) -> int:
    frames = (n_mono - FRAME) // HOP + 1 if n_mono >= FRAME else 0
    raw = max(frames - WARMUP, 0)
    return (raw + stride - 1) // stride


def shard_videos(costs, world: int) -> list[list[int]]:
    """Greedy longest-first assignment of videos to ranks by cost (sample
    count); deterministic, every rank computes the same plan.  Each rank's
    list is in ascending video order."""
    order = sorted(range(len(costs)), key=lambda v: (-int(costs[v]), v))
    load = [0] * world
    out = [[] for _ in range(world)]
    for v in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(v)
        load[r] += int(costs[v])
    return [sorted(x) for x in out]


def shard_pairs(seg_len: np.ndarray, pairs: np.ndarray, world: int, include_endings: bool):
    """Contiguous slices [start, end) of the pair list, balanced by the number
    of table cells n*m (+ endings)."""
    seg_len = np.asarray(seg_len, dtype=np.int64)
    a, b = pairs[:, 0].astype(np.int64), pairs[:, 1].astype(np.int64)
    cost = seg_len[2 * a] * seg_len[2 * b]
    if include_endings:
        cost = cost + seg_len[2 * a + 1] * seg_len[2 * b + 1]
    cum = np.concatenate([[0], np.cumsum(cost)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * r / world, side="left")))
    cuts.append(len(pairs))
    cuts = np.maximum.accumulate(np.minimum(cuts, len(pairs)))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def device_layout(n_mono, stride: int):
    """The library's device layout of a list of segments (nb200_fingerprint_layout):
    every segment starts at the next multiple of 4 elements.  -> (offsets, lengths, total)"""
    off, ln, pos = [], [], 0
    for x in n_mono:
        n = stored_hashes(int(x), stride)
        off.append(pos)
        ln.append(n)
        pos += (n + 3) & ~3
    return off, ln, pos


@dataclass
class SeasonPlan:
    """Who fingerprints what, and where each segment sits in the all-gather
    receive buffers.  Two buffers, rank-major, `cap` elements per rank: u32
    hashes and u64 timestamps, each rank's block in the library's own device
    layout -- so K2 writes straight into the send buffers and the match kernel
    reads straight out of the receive buffers (nb200_hashset_view): the
    exchange adds no copy kernel on either side."""
    world: int
    n_videos: int
    stride: int
    videos_of_rank: list          # list[list[int]]
    seg_len: np.ndarray           # u64 [2N] stored hashes per segment (opening_k, ending_k)
    cap: int                      # elements per rank block (multiple of 4)
    dev_offset: np.ndarray        # u64 [2N] element offset of each segment in the receive buffers

    @staticmethod
    def build(n_mono_per_segment, world: int, stride: int = 2) -> "SeasonPlan":
        n_mono = np.asarray(n_mono_per_segment, dtype=np.int64)
        assert n_mono.size % 2 == 0
        n_videos = n_mono.size // 2
        seg_len = np.array([stored_hashes(int(x), stride) for x in n_mono], dtype=np.uint64)
        costs = n_mono[0::2] + n_mono[1::2]
        vor = shard_videos(costs, world)
        layouts = []
        for vs in vor:
            segs = [2 * v + e for v in vs for e in (0, 1)]
            layouts.append((segs,) + device_layout([n_mono[s] for s in segs], stride))
        cap = max([total for _s, _o, _l, total in layouts] + [4])
        dev_offset = np.zeros(2 * n_videos, np.uint64)
        for r, (segs, off, _ln, _total) in enumerate(layouts):
            for s, o in zip(segs, off):
                dev_offset[s] = r * cap + o
        return SeasonPlan(world, n_videos, stride, vor, seg_len, cap, dev_offset)

    def local_segments(self, rank: int) -> list[int]:
        """Global segment ids this rank fingerprints, in its local order."""
        return [2 * v + e for v in self.videos_of_rank[rank] for e in (0, 1)]


def merge_runs(per_rank_runs, slices) -> np.ndarray:
    """Concatenate per-rank run lists (each sorted, pair indices relative to
    the rank's slice) into the global list in the reference's push order."""
    out = []
    for runs, (start, _end) in zip(per_rank_runs, slices):
        r = np.array(runs, dtype=RUN_DTYPE, copy=True)
        r["pair"] += np.uint32(start)
        out.append(r)
    return np.concatenate(out) if out else np.zeros(0, dtype=RUN_DTYPE)


def globalize_runs(runs: np.ndarray, slice_start: int) -> np.ndarray:
    """Pair indices relative to this rank's slice -> indices into the global pair list
    (done on the sending side, so the receiver only concatenates)."""
    if slice_start:
        runs = runs.copy()
        runs["pair"] += np.uint32(slice_start)
    return runs


def bind_host_to_gpu_numa(device_index: int) -> bool:
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity), so that the pinned
    PCM buffers it allocates afterwards are first-touched on that NUMA node and the H2D copies
    of several ranks do not all cross the same socket link.  Best effort: False if NVML or the
    affinity call is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device_index]) if vis and vis.replace(",", "").isdigit() else device_index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
            cpus &= set(os.sched_getaffinity(0))
            if not cpus:
                return False
            os.sched_setaffinity(0, cpus)
            return True
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return False


class GpuBackend:
    """The product compute path: libneedle_b200.so on this rank's GPU, running
    on torch's current stream so that NCCL collectives order with the kernels."""

    def __init__(self, device_index: int):
        import torch
        from . import engine
        self.torch = torch
        self.engine = engine
        self.device = torch.device("cuda", device_index)
        torch.cuda.set_device(self.device)
        self.ctx = engine.Context(device_index)
        # One real stream for the library's kernels AND torch's collectives: torch's default
        # stream is the NULL stream, which the library's non-blocking stream would not be ordered
        # with.  Every method that touches torch tensors runs under `with self._on_stream()`.
        self.stream = torch.cuda.Stream(device=self.device)
        self.ctx.set_stream(self.stream.cuda_stream)
        # the *_into fingerprint calls only enqueue: the job synchronises once, after the vote
        self.ctx.set_option(engine.Context.OPT_DEFER_WAIT, 1)
        self._run_bufs = None

    def _on_stream(self):
        return self.torch.cuda.stream(self.stream)

    def upload_hashes(self, hashes, ts_ns, seg_offset):
        """A whole library of precomputed hashes (KB..MB) on this rank's GPU."""
        return self.engine.HashSet.upload(self.ctx, hashes, ts_ns, seg_offset)

    def upload_pcm(self, segments):
        return self.engine.PcmSet.upload(self.ctx, segments)

    def fingerprint_search(self, pcmset, plan: SeasonPlan, seek_to_ns, hash_duration_ns, params):
        """Single-GPU path: nb200_pcmset_search (K1, K2, match, simhash and the vote enqueued back
        to back, one host wait); two calls when the videos' hash durations differ."""
        hd = np.asarray(hash_duration_ns, dtype=np.uint64)
        if hd.size and np.all(hd == hd[0]) and int(hd[0]) // 1_000_000 // 123 == plan.stride:
            return pcmset.search(seek_to_ns, int(hd[0]), params)
        hs = pcmset.fingerprint(stride=plan.stride, seek_to_ns=seek_to_ns)
        try:
            return hs.search(hash_duration_ns, params)
        finally:
            hs.free()

    def alloc_exchange(self, plan: SeasonPlan):
        t = self.torch
        with self._on_stream():
            return {"hash_send": t.zeros(plan.cap, dtype=t.int32, device=self.device),
                    "ts_send": t.zeros(plan.cap, dtype=t.int64, device=self.device),
                    "hash_recv": t.zeros(plan.cap * plan.world, dtype=t.int32, device=self.device),
                    "ts_recv": t.zeros(plan.cap * plan.world, dtype=t.int64, device=self.device)}

    def fingerprint_into(self, pcmset, plan: SeasonPlan, seek_to_ns, ex):
        """K1/K2 on this rank's segments, written directly into the send buffers."""
        pcmset.fingerprint_into(ex["hash_send"].data_ptr(), ex["ts_send"].data_ptr(), plan.cap,
                                stride=plan.stride, seek_to_ns=seek_to_ns)

    def fingerprint_host_into(self, segments, plan: SeasonPlan, seek_to_ns, ex):
        """Host PCM -> (H2D in chunks underneath K1) -> K2 into the send buffers."""
        self.ctx.fingerprint_host_into(segments, ex["hash_send"].data_ptr(), ex["ts_send"].data_ptr(), plan.cap,
                                       stride=plan.stride, seek_to_ns=seek_to_ns)

    def exchange(self, dist, ex, plan: SeasonPlan, seek_to_ns, group=None):
        """ONE collective: the hashes.  Timestamps are a function of the index
        (analyzer.rs:309), so every rank fills in all of them itself."""
        # the timestamps do not depend on anyone's hashes: filled first, while slower ranks are still
        # fingerprinting, so that nothing but the match stands behind the collective
        self.ctx.timestamps_fill(ex["ts_recv"].data_ptr(), plan.dev_offset, plan.seg_len, seek_to_ns,
                                 stride=plan.stride)
        with self._on_stream():
            dist.all_gather_into_tensor(ex["hash_recv"], ex["hash_send"], group=group)

    def season_from_gathered(self, ex, plan: SeasonPlan, world: int):
        h, t = (ex["hash_recv"], ex["ts_recv"]) if world > 1 else (ex["hash_send"], ex["ts_send"])
        return self.engine.HashSet.view(self.ctx, h.data_ptr(), t.data_ptr(), plan.dev_offset, plan.seg_len,
                                        keepalive=(h, t))

    # ---- device-resident tail: runs never visit the host ------------------------------------
    def run_block_capacity(self, slices) -> int:
        """Run records per rank block: the same on every rank (it sizes the collective)."""
        cap = getattr(self, "_block_cap", 0)
        need = max(4096, 16 * max((e - s) for s, e in slices))
        while cap < need:
            cap = max(4096, cap * 2)
        self._block_cap = cap
        return cap

    def search_gathered(self, dist, season, params, my_pairs, pair_base, all_pairs, hash_duration_ns, slices,
                        group=None):
        """match (this rank's slice, emitted straight into its run block) -> ONE all-gather of
        the blocks -> heap replay + find_best_match on rank 0's GPU -> 48 bytes per video come
        back.  One host wait per job; if a block overflowed every rank sees it in the gathered
        headers and the step is repeated with larger blocks."""
        t = self.torch
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        while True:
            cap = self.run_block_capacity(slices)
            block = 64 * (1 + cap)
            with self._on_stream():
                if self._run_bufs is None or self._run_bufs[0].numel() != block:
                    self._run_bufs = (t.zeros(block, dtype=t.uint8, device=self.device),
                                      t.zeros(block * world, dtype=t.uint8, device=self.device))
                send, recv = self._run_bufs
                season.match_export(params, my_pairs, pair_base, send.data_ptr(), block)
                dist.all_gather_into_tensor(recv, send, group=group)
            results, found, truncated = self.ctx.vote_blocks(recv.data_ptr(), world, block, hash_duration_ns, params,
                                                             pairs=all_pairs, do_vote=(rank == 0))
            if not truncated:
                return results
            self._block_cap = max(self._block_cap * 2, int(found))

    def match(self, season, params, pairs) -> np.ndarray:
        if len(pairs) == 0:
            return np.zeros(0, dtype=RUN_DTYPE)
        rs = season.match(params, pairs=pairs)
        runs = rs.download()
        rs.free()
        return runs

    def release(self, obj):
        obj.free()


def results_rows_as_tuples(arr: np.ndarray):
    f = ("present", "has_opening", "has_ending", "opening_start_ns", "opening_end_ns", "ending_start_ns",
         "ending_end_ns")
    return list(zip(*[arr[k].tolist() for k in f])) if arr.shape[0] else []


def unpack_run_blocks(buf: np.ndarray, counts, block: int) -> np.ndarray:
    """Padded per-rank blocks -> one run array (rank order = pair order)."""
    parts = [buf[r * block:r * block + int(n) * RUN_DTYPE.itemsize] for r, n in enumerate(counts) if n]
    if not parts:
        return np.zeros(0, dtype=RUN_DTYPE)
    return np.frombuffer(np.concatenate(parts).tobytes(), dtype=RUN_DTYPE)


def search_tail(job, season, n_videos: int, t):
    """What every rank does once it holds the whole season: match its slice of the pair list,
    bring the runs together, vote.  Device-resident when the backend offers it (run blocks, one
    all-gather, vote on rank 0's GPU); otherwise the generic form: runs to the host, gathered,
    every rank votes for the videos v with v % world == rank, rank 0 collects the rows."""
    b = job.backend
    start, end = job.slices[job.rank]
    if job.world > 1 and hasattr(b, "search_gathered"):
        results = b.search_gathered(job.dist, season, job.params, job.pairs[start:end], start,
                                    None if job._all_pairs else job.pairs, job.hash_duration_ns, job.slices,
                                    job.group)
        b.release(season)
        job._tick("match+gather+vote", t)
        return results
    runs = globalize_runs(b.match(season, job.params, job.pairs[start:end]), start)
    b.release(season)
    t = job._tick("match", t)
    all_runs = b.gather_runs(job.dist, runs, job.group) if job.world > 1 else runs
    t = job._tick("gather_runs", t)
    if job.world == 1:
        table = b.vote(job.hash_duration_ns, job.params, all_runs, job.pairs)
        job._tick("vote", t)
        return results_rows_as_tuples(table)
    mask = np.zeros(n_videos, np.uint8)
    mask[job.rank::job.world] = 1
    table = b.vote(job.hash_duration_ns, job.params, all_runs, job.pairs, video_mask=mask)
    t = job._tick("vote", t)
    cap = (n_videos + job.world - 1) // job.world
    rows = b.gather_results(job.dist, table[job.rank::job.world], cap, job.group)
    if job.rank != 0:
        return None
    full = np.zeros(n_videos, dtype=RESULT_DTYPE)
    for r in range(job.world):
        k = len(range(r, n_videos, job.world))
        full[r::job.world] = rows[r, :k]
    job._tick("gather_results", t)
    return results_rows_as_tuples(full)


class SearchJob:
    """`needle search` from precomputed frame hashes (.needle.dat) across the ranks of a process
    group: every rank is given the same season (it is KB..MB), matches its n*m-balanced slice of
    the pair list and the tail of SeasonJob does the rest.  Rank 0 receives the per-video results."""

    def __init__(self, backend, dist, hashes, ts_ns, seg_offset, hash_duration_ns, params, pairs=None, group=None):
        self.backend, self.dist, self.group, self.params = backend, dist, group, params
        self.world = dist.get_world_size(group) if dist is not None else 1
        self.rank = dist.get_rank(group) if dist is not None else 0
        self.hashes = np.ascontiguousarray(hashes, dtype=np.uint32)
        self.ts_ns = np.ascontiguousarray(ts_ns, dtype=np.uint64)
        self.seg_offset = np.ascontiguousarray(seg_offset, dtype=np.uint64)
        self.n_videos = (self.seg_offset.size - 1) // 2
        self.hash_duration_ns = np.asarray(hash_duration_ns, dtype=np.uint64)
        self.pairs = pair_list(self.n_videos) if pairs is None else \
            np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        self._all_pairs = np.array_equal(self.pairs, pair_list(self.n_videos))
        seg_len = np.diff(self.seg_offset.astype(np.int64))
        self.slices = shard_pairs(seg_len, self.pairs, self.world, bool(params.include_endings))
        self.phase_s = {}

    _tick = None   # set below: shared with SeasonJob

    def run(self):
        import time
        t = time.perf_counter()
        season = self.backend.upload_hashes(self.hashes, self.ts_ns, self.seg_offset)
        t = self._tick("upload", t)
        return search_tail(self, season, self.n_videos, t)


class SeasonJob:
    """`needle search --analyze` for a library of videos across the ranks of a
    process group (or on one GPU when dist is None): the public multi-GPU call.

    Every rank constructs it with the same global metadata (per-segment sample
    counts, seeks, pair list); each rank passes only ITS segments' PCM to
    run_host / upload.  Rank 0 receives the per-video results."""

    def __init__(self, backend, dist, n_mono_per_segment, seek_to_ns, hash_duration_ns, params,
                 pairs=None, stride: int = 2, group=None):
        self.backend = backend
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist is not None else 1
        self.rank = dist.get_rank(group) if dist is not None else 0
        self.params = params
        self.plan = SeasonPlan.build(n_mono_per_segment, self.world, stride)
        n_videos = self.plan.n_videos
        self.seek_to_ns = np.zeros(2 * n_videos, np.uint64) if seek_to_ns is None else \
            np.asarray(seek_to_ns, dtype=np.uint64)
        self.hash_duration_ns = np.asarray(hash_duration_ns, dtype=np.uint64)
        self.pairs = pair_list(n_videos) if pairs is None else \
            np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        self._all_pairs = np.array_equal(self.pairs, pair_list(n_videos))
        self.slices = shard_pairs(self.plan.seg_len, self.pairs, self.world, bool(params.include_endings))
        self.local_segs = self.plan.local_segments(self.rank)
        self.local_seek = self.seek_to_ns[self.local_segs] if self.local_segs else np.zeros(0, np.uint64)
        self._ex = None
        self.phase_s = {}      # wall seconds per phase of run_resident / run_host on this rank (accumulated)

    def _tick(self, name, t0):
        import time
        t1 = time.perf_counter()
        self.phase_s[name] = self.phase_s.get(name, 0.0) + (t1 - t0)
        return t1

    # which of the global segments this rank must be given
    def local_segment_ids(self):
        return list(self.local_segs)

    def upload(self, local_segments):
        assert len(local_segments) == len(self.local_segs)
        return self.backend.upload_pcm(local_segments)

    def _buffers(self):
        if self._ex is None:
            self._ex = self.backend.alloc_exchange(self.plan)
        return self._ex

    def run_resident(self, pcmset):
        """fingerprint -> all-gather -> match -> gather -> vote.  Returns the list of
        per-video result tuples on rank 0, None elsewhere."""
        b = self.backend
        if self.world == 1 and self._all_pairs and hasattr(b, "fingerprint_search"):
            # one GPU, default pair list: no exchange buffers, two C-ABI calls
            return b.fingerprint_search(pcmset, self.plan, self.local_seek, self.hash_duration_ns, self.params)
        import time
        t = time.perf_counter()
        ex = self._buffers()
        b.fingerprint_into(pcmset, self.plan, self.local_seek, ex)
        t = self._tick("fingerprint", t)
        return self._after_fingerprint(ex, t)

    def _after_fingerprint(self, ex, t):
        b = self.backend
        if self.world > 1:
            b.exchange(self.dist, ex, self.plan, self.seek_to_ns, self.group)
        season = b.season_from_gathered(ex, self.plan, self.world)
        t = self._tick("exchange_enqueue", t)
        return search_tail(self, season, self.plan.n_videos, t)

    def run_host(self, local_segments):
        """From host PCM (pinned for full PCIe rate).  The copy is pipelined under K1
        when the backend can stream (GpuBackend.fingerprint_host_into)."""
        b = self.backend
        one_gpu = self.world == 1 and self._all_pairs and hasattr(b, "fingerprint_search")
        if one_gpu or not hasattr(b, "fingerprint_host_into"):
            ps = self.upload(local_segments)
            try:
                return self.run_resident(ps)
            finally:
                b.release(ps)
        import time
        t = time.perf_counter()
        ex = self._buffers()
        b.fingerprint_host_into(local_segments, self.plan, self.local_seek, ex)
        t = self._tick("upload+fingerprint", t)
        return self._after_fingerprint(ex, t)


SearchJob._tick = SeasonJob._tick
