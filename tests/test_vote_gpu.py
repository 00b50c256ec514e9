"""The device vote (needle_b200/csrc/vote_device.cu: heap replay + find_best_match
on runs that never leave HBM) against the oracle's transcription of
needle/src/audio/comparator.rs:231-249 (heap order) and :405-515 (vote), and
against the library's own host vote (vote.cpp).  Bit-exact intervals; the cases
here aim at what the device version could get wrong: score ties (candidate
order), several runs per table (heap order), tables and videos beyond the
device kernels' bounds (host fallback), threshold 0 (empty clusters).
"""
import numpy as np
import pytest

from needle_b200 import engine, synth
from needle_b200._lib import OPT_HOST_VOTE
from tests import helpers as H

pytestmark = pytest.mark.gpu


def searches(ctx, season, p):
    """(device vote, host vote) results of the same resident season."""
    hs = engine.HashSet.upload(ctx, season.hashes, season.ts_ns, season.seg_offset)
    dev = hs.search(season.hash_duration_ns, p)
    dev_ms = ctx.last_kernel_ms()["vote"]
    ctx.set_option(OPT_HOST_VOTE, 1)
    try:
        host = hs.search(season.hash_duration_ns, p)
        host_ms = ctx.last_kernel_ms()["vote"]
    finally:
        ctx.set_option(OPT_HOST_VOTE, 0)
    hs.free()
    return dev, host, dev_ms, host_ms


def check(ctx, orc, season, **kw):
    kw = H.params_kw(**kw)
    p = engine.match_params(**kw)
    st, want, _ = H.oracle_run(orc, season, **kw)
    assert st == 0
    dev, host, dev_ms, host_ms = searches(ctx, season, p)
    assert dev == want
    assert host == want
    assert host_ms == 0.0
    return want, dev_ms


def test_identical_episodes_tie_on_score(ctx, oracle):
    """Every video is the same episode: all candidates of a video have the same cluster size
    and duration, so the winner is decided by the candidate order alone."""
    rng = np.random.default_rng(5)
    n_open, n_end = 500, 300
    oh = rng.integers(0, 2 ** 32, n_open, dtype=np.uint64).astype(np.uint32)
    eh = rng.integers(0, 2 ** 32, n_end, dtype=np.uint64).astype(np.uint32)
    ot = synth.hash_timestamps(2 * n_open, 2)[:n_open]
    et = synth.hash_timestamps(2 * n_end, 2, seek_to_ns=900_000_000_000)[:n_end]
    season = H.season_from_lists([(oh, ot)] * 6, [(eh, et)] * 6)
    want, dev_ms = check(ctx, oracle, season, include_endings=True)
    assert dev_ms > 0.0
    assert all(r[0] == 1 and r[1] == 1 and r[2] == 1 for r in want)


@pytest.mark.parametrize("seed", range(4))
def test_shifted_copies_equal_durations(ctx, oracle, seed):
    """The same theme at different offsets in every video, and twice in some: equal
    durations from different pairs, two runs in one table (heap order)."""
    rng = np.random.default_rng(40 + seed)
    season = H.random_season(rng, 7, 900, 500, jitter=bool(seed & 1))
    H.plant(rng, season, 120, [(v, int(rng.integers(2, 300))) for v in range(7)], flips=seed % 3)
    H.plant(rng, season, 120, [(v, int(rng.integers(450, 700))) for v in (0, 2, 3, 6)], flips=0)
    H.plant(rng, season, 100, [(v, int(rng.integers(2, 250))) for v in range(7)], ending=True, flips=1)
    check(ctx, oracle, season, include_endings=True, threshold=[10, 6, 14, 0][seed])


def test_many_runs_per_table_heap_order(ctx, oracle):
    """Short minimum duration on correlated hashes: tables with several runs each, so the
    BinaryHeap's internal order (not the push order) decides the candidate order."""
    rng = np.random.default_rng(77)
    season = H.random_season(rng, 5, 300, 160)
    for k in range(6):
        H.plant(rng, season, 12 + 2 * k, [(v, 10 + 45 * k + 3 * v) for v in range(5)], flips=1)
    ns = int(1.0e9)
    check(ctx, oracle, season, include_endings=True, threshold=12, min_opening_ns=ns, min_ending_ns=ns)


def test_table_beyond_device_bounds_falls_back(ctx, oracle):
    """> 64 runs in one table: the device raises its fallback flag and the library votes on
    the host; the caller sees the same, correct, result."""
    rng = np.random.default_rng(9)
    season = H.random_season(rng, 3, 400, 100, jitter=False)
    # 80 separate 6-hash copies on the main diagonal of pair (0, 1)
    a0, a1 = int(season.seg_offset[0]), int(season.seg_offset[2])
    for k in range(80):
        season.hashes[a1 + 2 + 4 * k: a1 + 2 + 4 * k + 3] = season.hashes[a0 + 2 + 4 * k: a0 + 2 + 4 * k + 3]
    ns = int(0.4e9)
    want, dev_ms = check(ctx, oracle, season, include_endings=False, threshold=1, min_opening_ns=ns, min_ending_ns=ns)
    assert want[0][0] == 1 and want[1][0] == 1


def test_bias_zero_has_no_clusters(ctx, oracle):
    """hash_match_threshold 0 => T + T/2 = 0: no candidate is in any cluster (popcount < 0
    is never true), videos are present but carry no interval (comparator.rs:441)."""
    rng = np.random.default_rng(3)
    season = H.random_season(rng, 4, 300, 200)
    H.plant(rng, season, 150, [(v, 20 + v) for v in range(4)], flips=0)
    want, _ = check(ctx, oracle, season, include_endings=True, threshold=0)
    assert any(r[0] == 1 for r in want) and not any(r[1] for r in want)


def test_season_28x20min_shapes(ctx, oracle):
    """BASELINE configs[1] shapes (378 pairs, openings 2,413 / endings 1,201 hashes)."""
    season = synth.make_hash_season(28, 2413, 1201, seed=11)
    want, dev_ms = check(ctx, oracle, season, include_endings=True)
    assert sum(r[1] for r in want) == 28 and sum(r[2] for r in want) == 28
    assert 0.0 < dev_ms < 5.0


def test_pcmset_search_equals_two_calls(ctx):
    """nb200_pcmset_search (one host wait) = nb200_fingerprint_run + nb200_search_hashset."""
    eps = synth.make_pcm_season(4, 3.0, season_seed=8, intro_s=40.0, credits_s=30.0)
    segs, seeks = [], []
    for ep in eps:
        a, b, sk = synth.split_segments(ep.pcm)
        segs += [a, b]
        seeks += [0, sk]
    p = engine.match_params(include_endings=True)
    ps = engine.PcmSet.upload(ctx, segs)
    hs = ps.fingerprint(stride=2, seek_to_ns=seeks)
    two = hs.search(np.full(4, synth.HASH_DURATION_NS, np.uint64), p)
    one = ps.search(seeks, synth.HASH_DURATION_NS, p)
    assert one == two
    assert ctx.analyze_search(segs, 1, seeks, synth.HASH_DURATION_NS, p) == two
    assert sum(r[1] for r in one) == 4


def test_200_video_library_device_equals_host_vote(ctx):
    """BASELINE configs[3] shapes (200 x 24 min: 19,900 pairs, 39,800 tables, ~400 candidates per
    video): too slow for the oracle's table fill, so the two independent votes of the library (GPU
    kernels, host C++) are checked against each other, and against the planted structure."""
    season = synth.make_hash_season(200, 2897, 1443, seed=4)
    p = engine.match_params(include_endings=True)
    dev, host, dev_ms, _ = searches(ctx, season, p)
    assert dev == host
    assert sum(r[1] for r in dev) >= 190 and sum(r[2] for r in dev) >= 190   # occasional hard breaks split a plant
    assert 0.0 < dev_ms < 5.0
    # what is found is (a piece of) the planted run: at most 366 hashes of 0.246 s, at least the 20 s minimum
    for r in dev:
        if r[1]:
            assert 19e9 < r[4] - r[3] < 95e9
