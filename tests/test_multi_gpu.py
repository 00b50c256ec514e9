"""The multi-GPU jobs behind the C ABI (nb200_comm_* / nb200_mjob_*, multi.cu).

One GPU is enough for the world-size-1 forms (what a host gets when it links the library on
a single-GPU machine: same calls, no NCCL, no peer memory); with two or more GPUs
tools/check_multi_gpu.py runs the real thing in both shapes -- one process driving N devices
(ncclCommInitAll, how needle itself would use it) and one process per device -- and compares
with the single-GPU calls and the oracle.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from needle_b200 import engine, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pcm_season(n_videos=5):
    eps = synth.make_pcm_season(n_videos, 3.0, season_seed=33, intro_s=40.0, credits_s=30.0)
    segs, seeks = [], []
    for ep in eps:
        a, b, sk = synth.split_segments(ep.pcm)
        segs += [a, b]
        seeks += [0, sk]
    return segs, seeks


def test_world_of_one_search_job_equals_nb200_search(oracle):
    season = synth.make_hash_season(10, 900, 500, seed=5, run_len=200)
    p = engine.match_params(include_endings=True)
    with engine.Context(0) as ctx:
        comm = engine.Comm.init_rank(ctx, None, 0, 1)
        job = engine.MultiJob.search([comm], season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, p)
        got = job.run()
        assert job.run() == got
        ms = job.phase_ms()
        assert ms["match"] > 0.0 and ms["vote"] > 0.0 and ms["hash_allgather"] == 0.0
        job.free()
        comm.destroy()
    with engine.Context(0) as c1:
        assert got == c1.search(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, p)
    s = oracle.Season(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns)
    st, ref, _ = oracle.run_with_frame_hashes(s, include_endings=True)
    assert st == 0 and [tuple(int(x) for x in r) for r in ref] == [tuple(r) for r in got]


def test_world_of_one_season_job_equals_analyze_search():
    segs, seeks = pcm_season()
    p = engine.match_params(include_endings=True)
    pairs = np.array([(0, 1), (0, 2), (1, 2), (3, 4)], dtype=np.uint32)
    with engine.Context(0) as ctx:
        comm = engine.Comm.init_all([ctx])[0]
        for pl in (None, pairs):
            job = engine.MultiJob.season([comm], [s.size for s in segs], seeks, synth.HASH_DURATION_NS, p, pairs=pl)
            assert job.video_rank().tolist() == [0] * 5
            host = job.run(segs)
            job.upload_pcm(segs)
            assert job.run() == host
            job.free()
            if pl is None:
                want = host
        comm.destroy()
    with engine.Context(0) as c1:
        assert want == c1.analyze_search(segs, 1, seeks, synth.HASH_DURATION_NS, p)
    assert sum(r[1] for r in want) == 5


def test_run_block_overflow_repeats_the_step():
    """min durations 0: far more runs than the first block holds; the step is repeated with room."""
    season = synth.make_hash_season(6, 500, 300, seed=9, run_len=120)
    p = engine.match_params(include_endings=True, min_opening_ns=0, min_ending_ns=0)
    with engine.Context(0) as ctx:
        comm = engine.Comm.init_rank(ctx, None, 0, 1)
        job = engine.MultiJob.search([comm], season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, p)
        got = job.run()
        job.free()
        comm.destroy()
    with engine.Context(0) as c1:
        assert len(c1.match_pairs(season.hashes, season.ts_ns, season.seg_offset, p)) > 4096
        assert got == c1.search(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, p)


def test_job_argument_errors():
    from needle_b200._lib import Nb200Error, ERR_COMPARATOR_MINIMUM_PATHS, ERR_STATE
    season = synth.make_hash_season(1, 100, 50, seed=1)
    p = engine.match_params()
    with engine.Context(0) as ctx:
        comm = engine.Comm.init_rank(ctx, None, 0, 1)
        with pytest.raises(Nb200Error) as e:
            engine.MultiJob.search([comm], season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, p)
        assert e.value.status == ERR_COMPARATOR_MINIMUM_PATHS      # comparator.rs: at least two videos
        s2 = synth.make_hash_season(3, 100, 50, seed=1)
        job = engine.MultiJob.search([comm], s2.hashes, s2.ts_ns, s2.seg_offset, s2.hash_duration_ns, p)
        with pytest.raises(Nb200Error) as e:
            job.video_rank()
        assert e.value.status == ERR_STATE                         # a search job has no video plan
        job.free()
        comm.destroy()


def _gpus():
    import torch
    return torch.cuda.device_count()


def test_single_process_n_devices():
    n = min(_gpus(), 4)
    if n < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_multi_gpu.py"), str(n)],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MULTI_OK single-process %d" % n in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_one_process_per_device():
    n = min(_gpus(), 4)
    if n < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29518", os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MULTI_OK processes %d" % n in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
