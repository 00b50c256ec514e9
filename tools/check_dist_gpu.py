"""Multi-GPU parity check, launched under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tools/check_dist_gpu.py

A small PCM season goes through SeasonJob on all ranks (episodes sharded for
fingerprinting, one all-gather of hashes, pair slices matched into run blocks, one
all-gather of the blocks, device vote on rank 0), from resident PCM and from host
PCM, with the default pair list and with an explicit one; rank 0 compares every
result with the single-GPU call nb200_analyze_search on the whole season, and
with a run whose blocks are deliberately too small (overflow -> repeat).
Prints "DIST_OK <world>" on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from needle_b200 import dist as nd  # noqa: E402
from needle_b200 import engine, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_videos = 7
    eps = synth.make_pcm_season(n_videos, 3.0, season_seed=21, intro_s=40.0, credits_s=30.0)
    segs, seeks = [], []
    for ep in eps:
        a, b, sk = synth.split_segments(ep.pcm)
        segs += [a, b]
        seeks += [0, sk]
    n_mono = [s.size for s in segs]
    hd = np.full(n_videos, synth.HASH_DURATION_NS, np.uint64)
    params = engine.match_params(include_endings=True)
    backend = nd.GpuBackend(local)

    want = None
    if rank == 0:
        with engine.Context(local) as ctx:
            want = ctx.analyze_search(segs, 1, seeks, synth.HASH_DURATION_NS, params)
        assert sum(r[1] for r in want) == n_videos, want

    def run(pairs, tiny_blocks=False):
        job = nd.SeasonJob(backend, dist, n_mono, seeks, hd, params, pairs=pairs)
        mine = [segs[s] for s in job.local_segment_ids()]
        if tiny_blocks:
            backend._block_cap = 0
            backend.run_block_capacity = lambda slices: max(backend._block_cap, 2)   # 2 records: must overflow
        ps = job.upload(mine)
        a = job.run_resident(ps)
        backend.release(ps)
        b = job.run_host(mine)
        if tiny_blocks:
            del backend.run_block_capacity
        return a, b

    a, b = run(None)
    if rank == 0:
        assert a == want and b == want, (a, want)
    else:
        assert a is None and b is None
    a, b = run(None, tiny_blocks=True)
    if rank == 0:
        assert a == want and b == want, "overflow retry"
        assert backend._block_cap >= 4
    # an explicit pair list: only pairs inside {0..3} and inside {4..6}
    pairs = np.array([(i, j) for i in range(n_videos) for j in range(i + 1, n_videos) if (i < 4) == (j < 4)],
                     dtype=np.uint32)
    a, b = run(pairs)
    if rank == 0:
        with engine.Context(local) as ctx:
            ps = engine.PcmSet.upload(ctx, segs)
            hs = ps.fingerprint(stride=2, seek_to_ns=seeks)
            runs = hs.match(params, pairs=pairs).download()
            sub = engine.vote(hd, params, runs, pairs=pairs)
        assert a == sub and b == sub, (a, sub)
    # search-only from precomputed hashes (BASELINE configs[3] in miniature): SearchJob
    season = synth.make_hash_season(12, 900, 500, seed=3, run_len=200)
    got = nd.SearchJob(backend, dist, season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns,
                       params).run()
    if rank == 0:
        with engine.Context(local) as ctx:
            assert got == ctx.search(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, params)
        assert sum(r[1] for r in got) >= 8
        print("DIST_OK %d" % world, flush=True)
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
