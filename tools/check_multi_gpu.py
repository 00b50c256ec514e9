"""Parity check of the multi-GPU jobs behind the C ABI (nb200_comm_* / nb200_mjob_*).

    python tools/check_multi_gpu.py [N]                  one process, N devices (ncclCommInitAll)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29518 tools/check_multi_gpu.py     one process per device (unique id)

Both shapes run the same jobs and compare rank 0's per-video results with the
single-GPU calls on the whole input (nb200_analyze_search / nb200_search) and
with the CPU oracle:
  * season job (analyze + search) from resident PCM and from host PCM, default and
    explicit pair lists;
  * search-only job from precomputed hashes;
  * a run whose run blocks are too small (min durations 0 -> thousands of runs):
    every rank must see the overflow and repeat the step.
Prints "MULTI_OK <mode> <world>"."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from needle_b200 import engine, synth  # noqa: E402


def workload():
    n_videos = 7
    eps = synth.make_pcm_season(n_videos, 3.0, season_seed=21, intro_s=40.0, credits_s=30.0)
    segs, seeks = [], []
    for ep in eps:
        a, b, sk = synth.split_segments(ep.pcm)
        segs += [a, b]
        seeks += [0, sk]
    return n_videos, segs, seeks


def tiny_season():
    h = np.arange(1, 13, dtype=np.uint32) * np.uint32(2654435761)
    e = np.arange(1, 6, dtype=np.uint32) * np.uint32(40503)
    ts = lambda n, seek=0: synth.hash_timestamps(2 * n, 2, seek_to_ns=seek)[:n]
    hashes = np.concatenate([h, e, h, e])
    ts_ns = np.concatenate([ts(12), ts(5, 10 ** 12), ts(12), ts(5, 10 ** 12)])
    off = np.array([0, 12, 17, 29, 34], np.uint64)
    season = synth.HashSeason(hashes, ts_ns, off, np.full(2, synth.HASH_DURATION_NS, np.uint64))
    return season, engine.match_params(threshold=2, include_endings=True, min_opening_ns=10 ** 9, min_ending_ns=5 * 10 ** 8)


def single_gpu_truth(device, n_videos, segs, seeks, params, pairs, season, params0):
    from oracle import oracle as orc
    with engine.Context(device) as ctx:
        want = ctx.analyze_search(segs, 1, seeks, synth.HASH_DURATION_NS, params)
        ps = engine.PcmSet.upload(ctx, segs)
        hs = ps.fingerprint(stride=2, seek_to_ns=seeks)
        runs = hs.match(params, pairs=pairs).download()
        sub = engine.vote(np.full(n_videos, synth.HASH_DURATION_NS, np.uint64), params, runs, pairs=pairs)
        want_search = ctx.search(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, params)
        want_dense = ctx.search(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, params0)
        want_few = ctx.analyze_search(segs[:6], 1, seeks[:6], synth.HASH_DURATION_NS, params)
        tiny, tiny_params = tiny_season()
        want_tiny = ctx.search(tiny.hashes, tiny.ts_ns, tiny.seg_offset, tiny.hash_duration_ns, tiny_params)
        assert sum(r[1] for r in want_tiny) == 2 and sum(r[2] for r in want_tiny) == 2, want_tiny
    # the oracle on the precomputed-hash season (bit-exact bar)
    s = orc.Season(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns)
    st, ref, _ = orc.run_with_frame_hashes(s, include_endings=True)
    assert st == 0 and [tuple(int(x) for x in r) for r in ref] == [tuple(int(x) for x in r) for r in want_search], \
        "single-GPU search differs from the oracle"
    assert sum(r[1] for r in want) == n_videos, want
    return want, sub, want_search, want_dense, want_few, want_tiny


def run_jobs(comms, n_videos, segs, seeks, params, pairs, season, params0):
    """-> dict of results (rank 0's process) or Nones."""
    n_mono = [s.size for s in segs]
    out = {}
    for name, pl in (("default", None), ("explicit", pairs)):
        job = engine.MultiJob.season(comms, n_mono, seeks, synth.HASH_DURATION_NS, params, pairs=pl)
        vr = job.video_rank()
        mine = {s: segs[s] for s in range(len(segs)) if any(c.rank == vr[s // 2] for c in comms)}
        job.upload_pcm(mine)
        a = job.run()
        a2 = job.run()
        b = job.run(mine)
        assert a == a2
        out[name + "_resident"], out[name + "_host"] = a, b
        out[name + "_phases"] = job.phase_ms()
        job.free()
    # fewer videos than ranks (world >= 4): some ranks fingerprint nothing and only match
    few = 3
    job = engine.MultiJob.season(comms, n_mono[:2 * few], seeks[:2 * few], synth.HASH_DURATION_NS, params)
    vr = job.video_rank()
    mine = {s: segs[s] for s in range(2 * few) if any(c.rank == vr[s // 2] for c in comms)}
    out["few_videos"] = job.run(mine)
    job.free()
    job = engine.MultiJob.search(comms, season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, params)
    out["search"] = job.run()
    job.free()
    # the same season from page-locked arrays (each rank copies its slice of the image directly)
    pin_h = engine.PinnedArray.empty(season.hashes.size, np.uint32)
    pin_t = engine.PinnedArray.empty(season.ts_ns.size, np.uint64)
    pin_h.array[:] = season.hashes
    pin_t.array[:] = season.ts_ns
    job = engine.MultiJob.search(comms, pin_h.array, pin_t.array, season.seg_offset, season.hash_duration_ns, params)
    pin_h.array[:] = 0    # released by the create call
    pin_t.array[:] = 0
    out["search_pinned"] = job.run()
    job.free()
    pin_h.free()
    pin_t.free()
    # a season smaller than the world: 2 videos (1 pair), 12 + 5 hashes each -- most ranks get an empty slice
    # of the image and no pair at all
    tiny, tiny_params = tiny_season()
    job = engine.MultiJob.search(comms, tiny.hashes, tiny.ts_ns, tiny.seg_offset, tiny.hash_duration_ns, tiny_params)
    out["search_tiny"] = job.run()
    job.free()
    # thousands of runs per rank: the first attempt's blocks (4096 records) overflow
    job = engine.MultiJob.search(comms, season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, params0)
    out["search_overflow"] = job.run()
    out["search_overflow_again"] = job.run()
    job.free()
    return out


def compare(out, want, sub, want_search, want_dense, want_few, want_tiny):
    assert out["default_resident"] == want, (out["default_resident"], want)
    assert out["default_host"] == want
    assert out["explicit_resident"] == sub and out["explicit_host"] == sub
    assert out["search"] == want_search
    assert out["search_pinned"] == want_search
    assert out["search_tiny"] == want_tiny
    assert out["few_videos"] == want_few
    assert out["search_overflow"] == want_dense and out["search_overflow_again"] == want_dense
    assert sum(r[1] for r in want_search) >= 8


def main():
    n_videos, segs, seeks = workload()
    params = engine.match_params(include_endings=True)
    params0 = engine.match_params(include_endings=True, min_opening_ns=0, min_ending_ns=0)
    pairs = np.array([(i, j) for i in range(n_videos) for j in range(i + 1, n_videos) if (i < 4) == (j < 4)],
                     dtype=np.uint32)
    season = synth.make_hash_season(12, 900, 500, seed=3, run_len=200)
    if "RANK" in os.environ:
        import torch
        import torch.distributed as dist
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        dist.init_process_group("gloo")
        from needle_b200 import dist as nd
        ctx = engine.Context(local)
        comm = nd.comm_from_torch(ctx, dist)
        out = run_jobs([comm], n_videos, segs, seeks, params, pairs, season, params0)
        if rank == 0:
            compare(out, *single_gpu_truth(local, n_videos, segs, seeks, params, pairs, season, params0))
            print("phases", out["default_phases"], flush=True)
            print("MULTI_OK processes %d" % world, flush=True)
        else:
            assert all(v is None for k, v in out.items() if not k.endswith("_phases"))
        dist.barrier()
        comm.destroy()
        dist.destroy_process_group()
    else:
        import torch
        n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
        ctxs = [engine.Context(d) for d in range(n)]
        comms = engine.Comm.init_all(ctxs)
        out = run_jobs(comms, n_videos, segs, seeks, params, pairs, season, params0)
        compare(out, *single_gpu_truth(0, n_videos, segs, seeks, params, pairs, season, params0))
        print("phases", out["default_phases"], flush=True)
        for c in comms:
            c.destroy()
        print("MULTI_OK single-process %d" % n, flush=True)


if __name__ == "__main__":
    main()
