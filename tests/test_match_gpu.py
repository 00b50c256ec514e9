"""GPU parity of the match stage (K3/K4 + host vote) against the CPU oracle,
through the C ABI.  Bit-exact: identical run lists in the reference's push
order (needle/src/audio/comparator.rs:191-192), identical simhashes
(:149-153), identical final SearchResult intervals (:405-515).
"""
import numpy as np
import pytest

from needle_b200 import engine, synth
from needle_b200._lib import OPT_FORCE_GENERAL_MATCH, OPT_MATCH_DENSE, OPT_HOST_VOTE, OPT_MATCH_BAND_GROUP, ERR_DURATION_UNDERFLOW, ERR_NO_ENDING, ERR_COMPARATOR_MINIMUM_PATHS, Nb200Error
from tests import helpers as H

pytestmark = pytest.mark.gpu


def check_season(ctx, orc, season, vote=True, **kw):
    """Run lists (and, with vote=True, final results) of every library path against the
    oracle.  vote=False skips find_best_match, whose O(candidates^2) (in the reference
    too) takes minutes when a zero minimum duration lets ~10^5 runs through."""
    kw = H.params_kw(**kw)
    p = engine.match_params(**kw)
    if vote:
        st, want_results, entries = H.oracle_run(orc, season, **kw)
        assert st == 0
        want_rows = H.entries_as_runs(entries)
    else:
        want_results, want_rows = None, H.oracle_pair_runs(orc, season, **kw)
    runs = ctx.match_pairs(season.hashes, season.ts_ns, season.seg_offset, p)
    assert H.runs_as_rows(runs) == want_rows
    # the same list from the other kernels that could have been chosen: the exhaustive fast
    # kernel (every cell instead of 4 sampled rows per word) and the general kernel
    for opt in (OPT_MATCH_DENSE, OPT_FORCE_GENERAL_MATCH):
        ctx.set_option(opt, 1)
        try:
            assert np.array_equal(ctx.match_pairs(season.hashes, season.ts_ns, season.seg_offset, p), runs)
        finally:
            ctx.set_option(opt, 0)
    # device-resident path gives the same answer
    hs = engine.HashSet.upload(ctx, season.hashes, season.ts_ns, season.seg_offset)
    rs = hs.match(p)
    n_runs, n_cells = rs.count()
    assert n_runs == len(want_rows)
    assert n_cells == season.n_cells(kw["include_endings"])
    assert np.array_equal(rs.download(), runs)
    # page-locked input takes the direct route (two packed copies + a device-side move into the aligned layout)
    pin_h = engine.PinnedArray.empty(season.hashes.size, np.uint32)
    pin_t = engine.PinnedArray.empty(season.ts_ns.size, np.uint64)
    try:
        pin_h.array[:] = season.hashes
        pin_t.array[:] = season.ts_ns
        assert np.array_equal(ctx.match_pairs(pin_h.array, pin_t.array, season.seg_offset, p), runs)
        hp = engine.HashSet.upload(ctx, pin_h.array, pin_t.array, season.seg_offset)
        pin_h.array[:] = 0     # the upload has released the caller's arrays
        pin_t.array[:] = 0
        got_h, got_t, _off = hp.download()
        assert np.array_equal(got_h, season.hashes) and np.array_equal(got_t, season.ts_ns)
        hp.free()
    finally:
        pin_h.free()
        pin_t.free()
    if vote:
        assert ctx.search(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, p) == want_results
        assert hs.search(season.hash_duration_ns, p) == want_results       # device vote (vote_device.cu)
        ctx.set_option(OPT_HOST_VOTE, 1)
        try:
            assert hs.search(season.hash_duration_ns, p) == want_results   # runs downloaded, host vote (vote.cpp)
        finally:
            ctx.set_option(OPT_HOST_VOTE, 0)
        # host-only vote on the downloaded runs
        assert engine.vote(season.hash_duration_ns, p, runs) == want_results
    return runs, want_results


@pytest.mark.parametrize("seed", range(6))
def test_random_small_no_min_duration(ctx, oracle, seed):
    """min duration 0: every maximal run qualifies (general path, lcand = 1)."""
    rng = np.random.default_rng(seed)
    season = H.random_season(rng, 4, 90, 50)
    thr = [10, 12, 8, 14, 0, 16][seed]
    runs, _ = check_season(ctx, oracle, season, threshold=thr, include_endings=True,
                           min_opening_ns=0, min_ending_ns=0)
    if thr >= 10:
        assert len(runs) > 0


@pytest.mark.parametrize("min_s", [0.2, 0.5, 1.0, 3.0])
def test_random_short_min_durations(ctx, oracle, min_s):
    """Short minima: candidate lengths between 1 and 30 cells."""
    rng = np.random.default_rng(100 + int(min_s * 10))
    season = H.random_season(rng, 5, 400, 200)
    H.plant(rng, season, 40, [(0, 5), (1, 300), (2, 100), (4, 260)])
    H.plant(rng, season, 25, [(0, 100), (3, 10)], ending=True)
    ns = int(min_s * 1e9)
    check_season(ctx, oracle, season, vote=min_s >= 0.5, threshold=13, include_endings=True, min_opening_ns=ns,
                 min_ending_ns=ns // 2)


@pytest.mark.parametrize("group", [1, 2, 3, 16])
def test_band_groups(ctx, oracle, group):
    """The fast kernel's unit of work = `group` consecutive bands of a row chunk (src rows fetched once,
    the dst window moved band by band): lists of 3-5 bands, ragged, with runs that cross band and chunk
    borders; every group size gives the oracle's runs."""
    season = synth.make_hash_season(5, 1100, 600, seed=30 + group, run_len=300, jitter_len=True)
    ctx.set_option(OPT_MATCH_BAND_GROUP, group)
    try:
        check_season(ctx, oracle, season, include_endings=True)
    finally:
        ctx.set_option(OPT_MATCH_BAND_GROUP, 0)


def test_planted_season_defaults(ctx, oracle):
    season = synth.make_hash_season(8, 700, 420, seed=3, run_len=200, jitter_len=True)
    runs, results = check_season(ctx, oracle, season, include_endings=True)
    assert len(runs) > 0
    assert sum(r[1] for r in results) >= 6 and sum(r[2] for r in results) >= 6


def test_planted_season_correlated_background(ctx, oracle):
    season = synth.make_hash_season(5, 500, 300, seed=5, run_len=150, correlated=True)
    check_season(ctx, oracle, season, include_endings=True, min_opening_ns=2_000_000_000,
                 min_ending_ns=1_000_000_000)


def test_runs_cross_row_chunks(ctx, oracle):
    """Runs longer than one 256-row chunk, starting/ending on chunk edges,
    touching index 1 and the last row / column."""
    rng = np.random.default_rng(11)
    season = H.random_season(rng, 6, 1100, 0, jitter=False)
    H.plant(rng, season, 700, [(0, 200), (1, 390)])          # spans 3 chunks of src
    H.plant(rng, season, 256, [(2, 256), (3, 1)])            # exactly one chunk; dst starts at 1
    H.plant(rng, season, 300, [(4, 800), (5, 800)])          # ends at the last element
    H.plant(rng, season, 120, [(0, 950), (5, 30)])
    check_season(ctx, oracle, season)
    check_season(ctx, oracle, season, vote=False, min_opening_ns=0)


def test_fast_kernel_word_ownership(ctx, oracle):
    """Runs of every length 60..140 at every alignment mod 32 (the fast kernel finds a
    run through its first fully matching aligned 32-row word), plus runs of many
    words and runs that touch the table edges."""
    rng = np.random.default_rng(21)
    n = 4200
    vids = [rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32) for _ in range(4)]
    ca, cb, planted = 3, 40, 0
    for ln in range(60, 141, 3):                      # disjoint runs, every alignment of start and diagonal
        base = rng.integers(0, 2 ** 32, ln, dtype=np.uint64).astype(np.uint32)
        vids[0][ca:ca + ln] = base
        vids[1][cb:cb + ln] = base
        ca += ln + 7
        cb += ln + 12
        planted += 1
    assert max(ca, cb) < n
    long_run = rng.integers(0, 2 ** 32, 700, dtype=np.uint64).astype(np.uint32)
    vids[2][1:701] = long_run            # starts at index 1 (row 0 never matches)
    vids[3][n - 700:] = long_run         # ends at the last column
    ts = synth.hash_timestamps(2 * n, 2)[:n]
    season = H.season_from_lists([(v, ts) for v in vids], [(np.zeros(0, np.uint32), np.zeros(0, np.uint64))] * 4)
    runs, _ = check_season(ctx, oracle, season, threshold=0)                                   # lcand 82
    assert len(runs) >= 15
    runs, _ = check_season(ctx, oracle, season, threshold=0, min_opening_ns=16_000_000_000)    # lcand 66
    assert len(runs) >= 20


def test_sampled_rows_false_positives(ctx, oracle):
    """Hashes that match on rows 0, 8, 16, 24 of many words but not in between (period-8
    pattern): every such word passes the sampled filter and must be rejected by the
    32-row verification; plus genuine long runs."""
    rng = np.random.default_rng(31)
    n = 1200
    a = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    key = np.uint32(0x5A5A1234)
    a[::8] = key                      # every 8th hash identical in both videos, all others random
    b[::8] = key
    c = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    c[100:400] = a[300:600]           # a genuine 300-run between videos 0 and 2
    ts = synth.hash_timestamps(2 * n, 2)[:n]
    none = (np.zeros(0, np.uint32), np.zeros(0, np.uint64))
    season = H.season_from_lists([(a, ts), (b, ts), (c, ts)], [none] * 3)
    runs, _ = check_season(ctx, oracle, season, threshold=0)
    assert len(runs) == 1 and int(runs[0]["len"]) >= 299
    check_season(ctx, oracle, season, threshold=10)


def test_threshold_extremes(ctx, oracle):
    rng = np.random.default_rng(12)
    season = H.random_season(rng, 3, 300, 150)
    check_season(ctx, oracle, season, threshold=32, include_endings=True)   # every cell matches
    check_season(ctx, oracle, season, threshold=32, include_endings=True, min_opening_ns=0, min_ending_ns=0)
    check_season(ctx, oracle, season, threshold=0, include_endings=True, min_opening_ns=0, min_ending_ns=0)
    # the reference takes any u16 threshold (comparator.rs:76): above 32 the match is what 32 gives, the
    # vote's clustering bias T + T/2 keeps the caller's T
    check_season(ctx, oracle, season, threshold=40, include_endings=True)


def test_identical_videos_threshold_zero(ctx, oracle):
    rng = np.random.default_rng(13)
    no, ne = 333, 129
    oh = rng.integers(0, 2 ** 32, no, dtype=np.uint64).astype(np.uint32)
    eh = rng.integers(0, 2 ** 32, ne, dtype=np.uint64).astype(np.uint32)
    ot = synth.hash_timestamps(2 * no, 2)[:no]
    et = synth.hash_timestamps(2 * ne, 2, seek_to_ns=900_000_000_000)[:ne]
    season = H.season_from_lists([(oh, ot)] * 4, [(eh, et)] * 4)
    runs, results = check_season(ctx, oracle, season, threshold=0, include_endings=True)
    # one main-diagonal run per table: i_end = j_end = n-1, len = n-1 (row/col 0 never match)
    assert {(r["i_end"], r["len"]) for r in runs if r["is_ending"] == 0} == {(no - 1, no - 1)}


@pytest.mark.parametrize("sizes", [(0, 0), (1, 1), (2, 2), (1, 5), (2, 300), (3, 3), (257, 2), (2, 257)])
def test_tiny_and_ragged(ctx, oracle, sizes):
    """Empty, single-element and 2-element lists next to ordinary ones."""
    rng = np.random.default_rng(sum(sizes))
    n0, n1 = sizes
    mk = lambda n, seek=0: (rng.integers(0, 4, n, dtype=np.uint64).astype(np.uint32),   # few distinct hashes
                            synth.hash_timestamps(2 * n, 2, seek_to_ns=seek)[:n])
    season = H.season_from_lists([mk(n0), mk(n1), mk(40)], [mk(5, 10 ** 12), mk(7, 10 ** 12), mk(3, 10 ** 12)])
    check_season(ctx, oracle, season, threshold=1, min_opening_ns=0, min_ending_ns=0, include_endings=True)
    check_season(ctx, oracle, season, threshold=1, min_opening_ns=0)


def test_missing_ending_is_an_error(ctx, oracle):
    rng = np.random.default_rng(14)
    season = H.random_season(rng, 3, 50, 0, jitter=False)
    st, _, _ = H.oracle_run(oracle, season, **H.params_kw(include_endings=True))
    assert st == -2   # ORC_ERR_NO_ENDING
    with pytest.raises(Nb200Error) as e:
        ctx.match_pairs(season.hashes, season.ts_ns, season.seg_offset, engine.match_params(include_endings=True))
    assert e.value.status == ERR_NO_ENDING


def test_minimum_paths(ctx):
    with pytest.raises(Nb200Error) as e:
        ctx.search(np.zeros(4, np.uint32), np.arange(4, dtype=np.uint64), np.array([0, 4, 4], np.uint64),
                   np.array([1], np.uint64), engine.match_params())
    assert e.value.status == ERR_COMPARATOR_MINIMUM_PATHS


def test_non_monotone_timestamps_underflow(ctx, oracle):
    """Duration - Duration panics in the reference; both sides must flag it."""
    rng = np.random.default_rng(15)
    season = H.random_season(rng, 2, 120, 0, jitter=False)
    H.plant(rng, season, 60, [(0, 10), (1, 30)], flips=0)
    season.ts_ns[40] = 5     # inside the run of video 0: ts[i_end] < ts[start] for some run
    season.ts_ns[69] = 1
    st, _, _ = H.oracle_run(oracle, season, **H.params_kw(min_opening_ns=0))
    assert st == -3
    with pytest.raises(Nb200Error) as e:
        ctx.match_pairs(season.hashes, season.ts_ns, season.seg_offset, engine.match_params(min_opening_ns=0))
    assert e.value.status == ERR_DURATION_UNDERFLOW


def test_arbitrary_monotone_timestamps(ctx, oracle):
    """.needle.dat files can carry any timestamps: duration filter is on the ns values."""
    rng = np.random.default_rng(16)
    season = H.random_season(rng, 4, 500, 0, jitter=False)
    H.plant(rng, season, 90, [(0, 50), (1, 200), (2, 400), (3, 7)])
    # irregular spacing: some 90-hash spans last 20 s, others do not
    off = season.seg_offset.astype(np.int64)
    for v in range(4):
        n = off[2 * v + 1] - off[2 * v]
        gaps = rng.integers(150_000_000, 300_000_000, n).astype(np.uint64)
        season.ts_ns[off[2 * v]:off[2 * v + 1]] = np.cumsum(gaps)
    check_season(ctx, oracle, season)
    check_season(ctx, oracle, season, min_opening_ns=15_000_000_000)


def test_explicit_pair_list(ctx, oracle):
    rng = np.random.default_rng(17)
    season = synth.make_hash_season(5, 400, 200, seed=17, run_len=120)
    kw = H.params_kw(include_endings=True)
    _, _, entries = H.oracle_run(oracle, season, **kw)
    all_rows = H.entries_as_runs(entries)
    # reference pair order for 5 videos
    ref_pairs = [(i, j) for i in range(5) for j in range(i + 1, 5)]
    pick = [7, 2, 9]
    pairs = np.array([ref_pairs[k] for k in pick], dtype=np.uint32)
    runs = ctx.match_pairs(season.hashes, season.ts_ns, season.seg_offset, engine.match_params(**kw), pairs=pairs)
    want = []
    for new_idx, k in enumerate(pick):
        want += [(new_idx,) + r[1:] for r in all_rows if r[0] == k]
    assert H.runs_as_rows(runs) == want


def test_time_padding_and_hash_durations(ctx, oracle):
    season = synth.make_hash_season(4, 600, 300, seed=21, run_len=180)
    season.hash_duration_ns[:] = [300_000_012, 500_000_000, 123_000_000, 1_000_000_000]
    check_season(ctx, oracle, season, include_endings=True, time_padding_ns=1_500_000_000)


@pytest.mark.parametrize("kind", ["correlated", "silence60", "jingle20", "quiet_half", "all_silence"])
def test_adversarial_seasons_equal_oracle(ctx, oracle, kind):
    """Seasons that keep diagonals alive (silence = one constant hash, repeated jingles, correlated
    background): every kernel's run list against the oracle, bit for bit.  The vote is skipped where
    the candidate count makes find_best_match quadratic in tens of thousands (in the reference too)."""
    season = synth.make_adversarial_season(kind, 4, 700, 400, seed=5)
    vote = kind in ("correlated", "jingle20")
    runs, _ = check_season(ctx, oracle, season, vote=vote, include_endings=True)
    if kind in ("silence60", "quiet_half", "all_silence"):
        assert len(runs) > 100      # blocks of matching cells: many qualifying diagonals


def test_adversarial_silence_short_minimum(ctx, oracle):
    """The same with a 2 s minimum (general kernel: candidate runs shorter than 63 cells)."""
    season = synth.make_adversarial_season("silence60", 3, 400, 300, seed=6)
    check_season(ctx, oracle, season, vote=False, include_endings=True, min_opening_ns=2_000_000_000,
                 min_ending_ns=2_000_000_000)


def test_full_size_24min_season(ctx, oracle):
    """BASELINE config sizes (24-min episodes: 2897 / 1443 hashes), 6 videos."""
    season = synth.make_hash_season(6, 2897, 1443, seed=24)
    runs, results = check_season(ctx, oracle, season, include_endings=True)
    assert len(runs) >= 15


def test_size_independent_properties_large(ctx):
    """60-min sizes (7259 / 3624), 12 videos = config 3: too slow for the oracle's
    full table in a unit test, so check properties the domain gives us --
    every run is a maximal diagonal run of matches that meets the duration
    minimum, swapping src/dst transposes the runs, and results are reproducible."""
    season = synth.make_hash_season(12, 7259, 3624, seed=60)
    p = engine.match_params(include_endings=True)
    hs = engine.HashSet.upload(ctx, season.hashes, season.ts_ns, season.seg_offset)
    rs = hs.match(p)
    runs = rs.download()
    n_runs, n_cells = rs.count()
    assert n_cells == season.n_cells(True) and n_runs == runs.shape[0] and n_runs >= 66
    pairs = [(i, j) for i in range(12) for j in range(i + 1, 12)]
    off = season.seg_offset.astype(np.int64)
    pc = lambda x: bin(int(x)).count("1")
    for r in runs:
        a, b = pairs[r["pair"]]
        e = int(r["is_ending"])
        sh = season.hashes[off[2 * a + e]:off[2 * a + e + 1]]
        dh = season.hashes[off[2 * b + e]:off[2 * b + e + 1]]
        st = season.ts_ns[off[2 * a + e]:off[2 * a + e + 1]]
        dt = season.ts_ns[off[2 * b + e]:off[2 * b + e + 1]]
        i, j, ln = int(r["i_end"]), int(r["j_end"]), int(r["len"])
        assert all(pc(sh[i - k] ^ dh[j - k]) <= 10 for k in range(ln))            # all cells match
        assert i - ln == 0 or j - ln == 0 or pc(sh[i - ln] ^ dh[j - ln]) > 10       # maximal at the start
        assert i == len(sh) - 1 or j == len(dh) - 1 or pc(sh[i + 1] ^ dh[j + 1]) > 10   # and at the end
        assert int(st[i] - st[i - ln]) >= 20_000_000_000 and int(dt[j] - dt[j - ln]) >= 20_000_000_000
    # transposition: (b, a) pairs give the same runs with i and j swapped
    swapped = np.array([(b, a) for a, b in pairs], dtype=np.uint32)
    runs_t = hs.match(p, pairs=swapped).download()
    key = lambda rr, sw: sorted((int(x["pair"]), int(x["is_ending"]),
                                 int(x["j_end"] if sw else x["i_end"]), int(x["i_end"] if sw else x["j_end"]),
                                 int(x["len"]), int(x["dst_simhash"] if sw else x["src_simhash"])) for x in rr)
    assert key(runs, False) == key(runs_t, True)
    assert np.array_equal(hs.match(p).download(), runs)
