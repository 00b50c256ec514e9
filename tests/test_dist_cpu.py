"""The N > 1 plumbing on CPU: world_size-2 gloo process group, the real
SeasonJob / SeasonPlan / pair sharding / run merging / nb200_vote, with the
compute steps played by the oracle (a test-only checker backend).  Rank 0's
results must equal the oracle's single-process run of the same library."""
import os
import socket

import numpy as np
import pytest

from needle_b200 import dist as nd
from needle_b200 import engine, synth
from needle_b200._lib import RUN_DTYPE
from tests import helpers as H


def test_pair_list_is_reference_order():
    from needle_b200._lib import lib
    assert nd.pair_list(4).tolist() == [[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]]
    assert nd.pair_list(1).shape == (0, 2)


def test_stored_hashes_formula(oracle):
    for n in (0, 4096, 30_000, 6_615_000, 3_307_500, 7_938_000):
        for stride in (1, 2, 3):
            raw = oracle.num_raw_hashes(n)
            assert nd.stored_hashes(n, stride) == (raw + stride - 1) // stride


def test_plans_cover_everything_once():
    rng = np.random.default_rng(0)
    n_mono = rng.integers(0, 2_000_000, 2 * 13)
    for world in (1, 2, 3, 8):
        plan = nd.SeasonPlan.build(n_mono, world)
        seen = sorted(v for vs in plan.videos_of_rank for v in vs)
        assert seen == list(range(13))
        # segments do not overlap inside the receive buffers, start 4-aligned, stay in their rank's block
        spans = sorted((int(plan.dev_offset[s]), int(plan.dev_offset[s] + plan.seg_len[s])) for s in range(26))
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
        assert all(int(o) % 4 == 0 for o in plan.dev_offset) and plan.cap % 4 == 0
        for r, vs in enumerate(plan.videos_of_rank):
            for v in vs:
                for e in (0, 1):
                    s = 2 * v + e
                    assert r * plan.cap <= plan.dev_offset[s] and plan.dev_offset[s] + plan.seg_len[s] <= (r + 1) * plan.cap
        # the python layout equals the library's (nb200_fingerprint_layout)
        off, ln, total = engine.fingerprint_layout(n_mono, 2)
        o2, l2, t2 = nd.device_layout(n_mono, 2)
        assert off.tolist() == o2 and ln.tolist() == l2 and total == t2
        pairs = nd.pair_list(13)
        sl = nd.shard_pairs(plan.seg_len, pairs, world, True)
        assert sl[0][0] == 0 and sl[-1][1] == len(pairs)
        assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
        # ... and so are the sharding rules (multi.cu: nb200_plan_videos / nb200_plan_pairs)
        want = np.zeros(13, np.uint32)
        for r, vs in enumerate(plan.videos_of_rank):
            want[vs] = r
        assert engine.plan_videos(n_mono, world).tolist() == want.tolist()
        for endings in (False, True):
            sl = nd.shard_pairs(plan.seg_len, pairs, world, endings)
            assert engine.plan_pairs(plan.seg_len, world, endings).tolist() == [a for a, _ in sl] + [len(pairs)]
        sub = pairs[::3]
        sl = nd.shard_pairs(plan.seg_len, sub, world, True)
        assert engine.plan_pairs(plan.seg_len, world, True, pairs=sub).tolist() == [a for a, _ in sl] + [len(sub)]
    loads = [sum(int(n_mono[2 * v] + n_mono[2 * v + 1]) for v in vs) for vs in nd.SeasonPlan.build(n_mono, 4).videos_of_rank]
    assert max(loads) - min(loads) <= int(n_mono.reshape(-1, 2).sum(1).max())


class OracleBackend:
    """TEST-ONLY backend of the dist.py mirror: compute by the oracle."""

    def __init__(self):
        import torch
        from oracle import oracle as orc
        self.torch, self.orc = torch, orc

    def upload_pcm(self, segments):
        return list(segments)

    def upload_hashes(self, hashes, ts_ns, seg_offset):
        return synth.HashSeason(np.asarray(hashes, np.uint32), np.asarray(ts_ns, np.uint64),
                                np.asarray(seg_offset, np.uint64), None)

    def alloc_exchange(self, plan):
        t = self.torch
        return {"hash_send": t.zeros(plan.cap, dtype=t.int32), "ts_send": t.zeros(plan.cap, dtype=t.int64),
                "hash_recv": t.zeros(plan.cap * plan.world, dtype=t.int32),
                "ts_recv": t.zeros(plan.cap * plan.world, dtype=t.int64)}

    def fingerprint_into(self, segs, plan, seek, ex):
        hv = ex["hash_send"].numpy().view(np.uint32)
        tv = ex["ts_send"].numpy().view(np.uint64)
        off, ln, _total = nd.device_layout([s.size for s in segs], plan.stride)
        for pcm, sk, o, n in zip(segs, seek, off, ln):
            h, t = self.orc.subsample_and_stamp(self.orc.fingerprint(pcm), plan.stride, seek_to_ns=int(sk))
            assert h.size == n
            hv[o:o + n] = h
            tv[o:o + n] = t

    def exchange(self, dist, ex, plan, seek_to_ns, group=None):
        dist.all_gather_into_tensor(ex["hash_recv"], ex["hash_send"], group=group)
        tv = ex["ts_recv"].numpy().view(np.uint64)      # timestamps from the index, like nb200_timestamps_fill
        for s in range(2 * plan.n_videos):
            n, a = int(plan.seg_len[s]), int(plan.dev_offset[s])
            tv[a:a + n] = synth.hash_timestamps(n * plan.stride, plan.stride, seek_to_ns=int(seek_to_ns[s]))[:n]

    def season_from_gathered(self, ex, plan, world):
        h, t = (ex["hash_recv"], ex["ts_recv"]) if world > 1 else (ex["hash_send"], ex["ts_send"])
        hv, tv = h.numpy().view(np.uint32), t.numpy().view(np.uint64)
        segs = []
        for s in range(2 * plan.n_videos):
            n, a = int(plan.seg_len[s]), int(plan.dev_offset[s])
            segs.append((hv[a:a + n].copy(), tv[a:a + n].copy()))
        return H.season_from_lists(segs[0::2], segs[1::2])

    def match(self, season, params, pairs):
        rows = []
        for k, (a, b) in enumerate(pairs):
            for e in ((0, 1) if params.include_endings else (0,)):
                sa, sb = 2 * int(a) + e, 2 * int(b) + e
                o = season.seg_offset.astype(np.int64)
                ent = self.orc.longest_common_hash_match(
                    season.hashes[o[sa]:o[sa + 1]], season.ts_ns[o[sa]:o[sa + 1]],
                    season.hashes[o[sb]:o[sb + 1]], season.ts_ns[o[sb]:o[sb + 1]],
                    threshold=params.hash_match_threshold, min_opening_ns=params.min_opening_ns,
                    min_ending_ns=params.min_ending_ns, is_opening=(e == 0))
                ent.sort(key=lambda x: (-x[10], -x[11]))
                rows += [(k, e, x[10], x[11], x[0], x[5], x[6], x[1], x[2], x[3], x[4]) for x in ent]
        return H.rows_to_runs(rows)

    def gather_runs(self, dist, runs, group=None):
        """Pad-and-all-gather of the run lists, on CPU tensors."""
        t = self.torch
        world = dist.get_world_size(group)
        cnts = t.zeros(world, dtype=t.int64)
        dist.all_gather_into_tensor(cnts, t.tensor([runs.shape[0]], dtype=t.int64), group=group)
        counts = cnts.numpy()
        cap = 64
        while cap < int(counts.max()):
            cap *= 2
        block = cap * RUN_DTYPE.itemsize
        send = t.zeros(block, dtype=t.uint8)
        if runs.shape[0]:
            send.numpy()[:runs.shape[0] * RUN_DTYPE.itemsize] = np.frombuffer(runs.tobytes(), dtype=np.uint8)
        recv = t.zeros(block * world, dtype=t.uint8)
        dist.all_gather_into_tensor(recv, send, group=group)
        return nd.unpack_run_blocks(recv.numpy(), counts, block)

    def gather_results(self, dist, mine, cap, group=None):
        t = self.torch
        from needle_b200._lib import RESULT_DTYPE
        world = dist.get_world_size(group)
        nbytes = cap * RESULT_DTYPE.itemsize
        send = t.zeros(nbytes, dtype=t.uint8)
        send.numpy()[:mine.shape[0] * RESULT_DTYPE.itemsize] = np.frombuffer(mine.tobytes(), dtype=np.uint8)
        recv = t.zeros(nbytes * world, dtype=t.uint8)
        dist.all_gather_into_tensor(recv, send, group=group)
        if dist.get_rank(group) != 0:
            return None
        return np.frombuffer(recv.numpy().tobytes(), dtype=RESULT_DTYPE).reshape(world, cap)

    def vote(self, hd, params, runs, pairs, video_mask=None):
        return engine.vote_array(hd, params, runs, pairs=pairs, video_mask=video_mask)

    def release(self, obj):
        pass


def library(n_seasons, per_season):
    """Small PCM library: n_seasons shows, each with its own shared intro/credits."""
    eps, segs, seek = [], [], []
    for s in range(n_seasons):
        eps += synth.make_pcm_season(per_season, 2.6, season_seed=50 + s, intro_s=36.0, credits_s=26.0)
    for ep in eps:
        a, b, sk = synth.split_segments(ep.pcm)
        segs += [a, b]
        seek += [0, sk]
    pairs = [(s * per_season + i, s * per_season + j) for s in range(n_seasons)
             for i in range(per_season) for j in range(i + 1, per_season)]
    return segs, seek, np.array(pairs, dtype=np.uint32)


def _worker(rank, world, port, n_seasons, per_season, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        segs, seek, pairs = library(n_seasons, per_season)
        n_mono = [s.size for s in segs]
        params = engine.match_params(include_endings=True, min_opening_ns=15_000_000_000,
                                     min_ending_ns=10_000_000_000)
        hd = np.full(len(segs) // 2, synth.HASH_DURATION_NS, np.uint64)
        job = nd.SeasonJob(OracleBackend(), dist, n_mono, seek, hd, params, pairs=pairs)
        mine = [segs[s] for s in job.local_segment_ids()]
        res = job.run_host(mine)
        res2 = job.run_host(mine)          # buffers are reused across steps
        if rank == 0:
            assert res == res2
            np.save(out_path, np.array(res, dtype=np.uint64))
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_seasons,per_season", [(1, 4), (2, 3)])
def test_world2_gloo_equals_single_process_oracle(oracle, tmp_path, n_seasons, per_season):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_seasons, per_season, out), nprocs=2, join=True)
    got = [tuple(int(x) for x in row) for row in np.load(out)]
    # the oracle, one season at a time, single process
    segs, seek, _ = library(n_seasons, per_season)
    want = []
    for s in range(n_seasons):
        op, en = [], []
        for v in range(s * per_season, (s + 1) * per_season):
            op.append(oracle.subsample_and_stamp(oracle.fingerprint(segs[2 * v]), 2))
            en.append(oracle.subsample_and_stamp(oracle.fingerprint(segs[2 * v + 1]), 2, seek_to_ns=seek[2 * v + 1]))
        st, res, _ = H.oracle_run(oracle, H.season_from_lists(op, en),
                                  **H.params_kw(include_endings=True, min_opening_ns=15_000_000_000,
                                                min_ending_ns=10_000_000_000))
        assert st == 0
        want += res
    assert got == want
    assert sum(r[1] for r in got) >= len(got) - 1      # openings found


def _search_worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        season = synth.make_hash_season(5, 260, 140, seed=7, run_len=120)
        params = engine.match_params(include_endings=True)
        job = nd.SearchJob(OracleBackend(), dist, season.hashes, season.ts_ns, season.seg_offset,
                           season.hash_duration_ns, params)
        assert job.slices[0][0] == 0 and job.slices[-1][1] == 10 and job.slices[0][1] == job.slices[1][0]
        res = job.run()
        if rank == 0:
            np.save(out_path, np.array(res, dtype=np.uint64))
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


def test_world2_search_job_from_precomputed_hashes(oracle, tmp_path):
    """BASELINE configs[3] in miniature: search-only from precomputed hashes, pairs sharded over 2 ranks."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.npy")
    mp.spawn(_search_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = [tuple(int(x) for x in row) for row in np.load(out)]
    season = synth.make_hash_season(5, 260, 140, seed=7, run_len=120)
    st, want, _ = H.oracle_run(oracle, season, **H.params_kw(include_endings=True))
    assert st == 0 and got == want
    assert sum(r[1] for r in got) == 5
