"""Host-side mirror of the multi-GPU job, for CPU tests.

The product's multi-GPU path is behind the C ABI (nb200_comm_* / nb200_mjob_*,
csrc/multi.cu; engine.Comm / engine.MultiJob): one process driving N devices or one
process per device, the library issuing the hash all-gather (NCCL) and pushing the run
blocks to rank 0 over peer memory itself.  `comm_from_torch` below is all a torchrun
host needs.

  episodes --shard--> fingerprint (K1/K2)        no collective
  hashes   --ONE all-gather--> every rank holds the whole season (KB..MB)
  pairs    --shard--> match (K3/K4)              no collective
  runs     --gather--> rank 0, vote (find_best_match)

which is how the reference's two rayon par_iter loops (per video,
needle/src/audio/analyzer.rs:437-445; per pair, comparator.rs:549-564) map to ranks.
None of that can run without GPUs, so the same flow is written out here in plain Python
over a `Backend` protocol: world_size-2 gloo tests (tests/test_dist_cpu.py) run it with
the oracle playing the compute, and check that the sharding rules here (videos by sample
count, pair slices by table cells, buffer layout) are the library's own
(nb200_plan_videos, nb200_plan_pairs, nb200_fingerprint_layout).
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


def comm_from_torch(ctx, dist, device=None, group=None):
    """One nb200_comm per process from a torch.distributed process group: rank 0 makes NCCL's unique
    id (nb200_comm_unique_id), torch carries the 128 bytes, every rank joins
    (nb200_comm_init_rank).  From here on the collectives are the library's own."""
    from . import engine
    if dist is None:
        return engine.Comm.init_rank(ctx, None, 0, 1)
    import torch
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if dist.get_backend(group) == "nccl":
        buf = torch.zeros(128, dtype=torch.uint8, device=device)
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(engine.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = buf.cpu().numpy().tobytes()
    else:
        box = [engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = box[0]
    return engine.Comm.init_rank(ctx, uid, rank, world)


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
        when the backend can stream (a fingerprint_host_into method)."""
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
