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

from ._lib import RUN_DTYPE

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


@dataclass
class SeasonPlan:
    """Who fingerprints what, and where each segment lands in the all-gather
    receive buffer.  Per-rank block = [u32 hashes x cap | u64 ts x cap] bytes."""
    world: int
    n_videos: int
    stride: int
    videos_of_rank: list          # list[list[int]]
    seg_len: np.ndarray           # u64 [2N] stored hashes per segment (opening_k, ending_k)
    cap: int                      # hashes per rank block (even)
    block_bytes: int
    src_hash_off: np.ndarray      # u64 [2N] element offset in the receive buffer viewed as u32
    src_ts_off: np.ndarray        # u64 [2N] element offset in the receive buffer viewed as u64

    @staticmethod
    def build(n_mono_per_segment, world: int, stride: int = 2) -> "SeasonPlan":
        n_mono = np.asarray(n_mono_per_segment, dtype=np.int64)
        assert n_mono.size % 2 == 0
        n_videos = n_mono.size // 2
        seg_len = np.array([stored_hashes(int(x), stride) for x in n_mono], dtype=np.uint64)
        costs = n_mono[0::2] + n_mono[1::2]
        vor = shard_videos(costs, world)
        per_rank = [int(sum(int(seg_len[2 * v]) + int(seg_len[2 * v + 1]) for v in vs)) for vs in vor]
        cap = max(per_rank + [2])
        cap += cap & 1
        block_bytes = 12 * cap
        src_h = np.zeros(2 * n_videos, np.uint64)
        src_t = np.zeros(2 * n_videos, np.uint64)
        for r, vs in enumerate(vor):
            local = 0
            for v in vs:
                for e in (0, 1):
                    s = 2 * v + e
                    src_h[s] = (r * block_bytes) // 4 + local
                    src_t[s] = (r * block_bytes + 4 * cap) // 8 + local
                    local += int(seg_len[s])
        return SeasonPlan(world, n_videos, stride, vor, seg_len, cap, block_bytes, src_h, src_t)

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
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def upload_pcm(self, segments):
        return self.engine.PcmSet.upload(self.ctx, segments)

    def fingerprint_into(self, pcmset, plan: SeasonPlan, seek_to_ns, send):
        """K1/K2 on this rank's segments, packed into the send block (uint8 tensor)."""
        hs = pcmset.fingerprint(stride=plan.stride, seek_to_ns=seek_to_ns)
        hs.export_packed(send.data_ptr(), send.data_ptr() + 4 * plan.cap)
        hs.free()

    def fingerprint_search(self, pcmset, plan: SeasonPlan, seek_to_ns, hash_duration_ns, params):
        """Single-GPU path: nb200_fingerprint_run + nb200_search_hashset."""
        hs = pcmset.fingerprint(stride=plan.stride, seek_to_ns=seek_to_ns)
        try:
            return hs.search(hash_duration_ns, params)
        finally:
            hs.free()

    def season_from_gathered(self, recv, plan: SeasonPlan):
        return self.engine.HashSet.from_device_scattered(self.ctx, recv.data_ptr(), recv.data_ptr(),
                                                         plan.src_hash_off, plan.src_ts_off, plan.seg_len)

    def match(self, season, params, pairs) -> np.ndarray:
        if len(pairs) == 0:
            return np.zeros(0, dtype=RUN_DTYPE)
        rs = season.match(params, pairs=pairs)
        runs = rs.download()
        rs.free()
        return runs

    def alloc_exchange(self, plan: SeasonPlan):
        t = self.torch
        send = t.zeros(plan.block_bytes, dtype=t.uint8, device=self.device)
        recv = t.zeros(plan.block_bytes * plan.world, dtype=t.uint8, device=self.device)
        return send, recv

    def gather_runs(self, dist, runs, group=None):
        return gather_runs(dist, runs, self.device, self.torch, group)

    def vote(self, hash_duration_ns, params, runs, pairs):
        return self.engine.vote(hash_duration_ns, params, runs, pairs=pairs)

    def release(self, obj):
        obj.free()


def gather_runs(dist, runs: np.ndarray, device, torch, group=None) -> list | None:
    """All ranks contribute a run list; rank 0 gets the list of arrays.  Two
    fixed-size collectives (counts, then padded payload): no pickling."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    cnt = torch.tensor([runs.shape[0]], dtype=torch.int64, device=device)
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, cnt, group=group)
    counts = counts.cpu().numpy()
    cap = int(counts.max())
    if cap == 0:
        return [np.zeros(0, dtype=RUN_DTYPE) for _ in range(world)] if rank == 0 else None
    send = torch.zeros(cap * RUN_DTYPE.itemsize, dtype=torch.uint8, device=device)
    if runs.shape[0]:
        raw = torch.from_numpy(np.frombuffer(runs.tobytes(), dtype=np.uint8).copy())
        send[:raw.numel()] = raw.to(device)
    recv = torch.zeros(world * cap * RUN_DTYPE.itemsize, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    if rank != 0:
        return None
    host = recv.cpu().numpy()
    out = []
    for r in range(world):
        blk = host[r * cap * RUN_DTYPE.itemsize:(r * cap + int(counts[r])) * RUN_DTYPE.itemsize]
        out.append(np.frombuffer(blk.tobytes(), dtype=RUN_DTYPE).copy())
    return out


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
        self._send = self._recv = None

    # which of the global segments this rank must be given
    def local_segment_ids(self):
        return list(self.local_segs)

    def upload(self, local_segments):
        assert len(local_segments) == len(self.local_segs)
        return self.backend.upload_pcm(local_segments)

    def _buffers(self):
        if self._send is None:
            self._send, self._recv = self.backend.alloc_exchange(self.plan)
        return self._send, self._recv

    def run_resident(self, pcmset):
        """fingerprint -> all-gather -> match -> gather -> vote.  Returns the list of
        per-video result tuples on rank 0, None elsewhere."""
        b = self.backend
        if self.world == 1 and self._all_pairs and hasattr(b, "fingerprint_search"):
            # one GPU, default pair list: no exchange buffers, two C-ABI calls
            return b.fingerprint_search(pcmset, self.plan, self.local_seek, self.hash_duration_ns, self.params)
        send, recv = self._buffers()
        b.fingerprint_into(pcmset, self.plan, self.local_seek, send)
        if self.world > 1:
            self.dist.all_gather_into_tensor(recv, send, group=self.group)
        else:
            recv = send
        season = b.season_from_gathered(recv, self.plan)
        start, end = self.slices[self.rank]
        runs = b.match(season, self.params, self.pairs[start:end])
        if self.world > 1:
            gathered = b.gather_runs(self.dist, runs, self.group)
        else:
            gathered = [runs]
        results = None
        if self.rank == 0:
            all_runs = merge_runs(gathered, self.slices)
            results = b.vote(self.hash_duration_ns, self.params, all_runs, self.pairs)
        b.release(season)
        return results

    def run_host(self, local_segments):
        ps = self.upload(local_segments)
        try:
            return self.run_resident(ps)
        finally:
            self.backend.release(ps)
