"""The oracle's match stage checked three ways (CPU only):
  * hand-checkable known answers for the quirks of
    needle/src/audio/comparator.rs:157-250;
  * Rust Duration / f32 semantics (analyzer.rs:293-318) on known values;
  * the C restatement (oracle/match_ref.c) against an independent pure-Python
    transcription (oracle/pyref.py) on seeded random inputs.
The reference itself ships no golden vectors for this path (parity unpinned)."""
import numpy as np
import pytest

from needle_b200 import synth
from oracle import pyref
from tests import helpers as H

S = 1_000_000_000


def lists(h, t):
    return list(zip([int(x) for x in h], [int(x) for x in t]))


# ---------------------------------------------------------------- Duration

def test_duration_known_values(oracle):
    assert oracle.duration_from_secs_f32(0.3) == 300_000_012          # DEFAULT_HASH_DURATION
    assert oracle.duration_from_secs_f32(0.5) == 500_000_000
    assert oracle.duration_from_secs_f32(0.0) == 0
    assert oracle.duration_from_secs_f32(1.0) == S
    assert oracle.duration_as_secs_f32(123_000_000) == float(np.float32(0.123))
    # item.mul_f32(i): one f32 multiply, then exact conversion
    for i in (0, 1, 2, 3, 100, 2897 * 2, 16_777_217):
        want = pyref.hash_timestamp(2_600_000_000, 123_000_000, i, 7)
        assert oracle.hash_timestamp(2_600_000_000, 123_000_000, i, 7) == want
    # round-half-even at the ns: 2^-1 ns cases do not occur for f32 < 2^23 * 1e-9 s; spot check monotonic
    ts = [oracle.hash_timestamp(2_600_000_000, 123_000_000, i, 0) for i in range(0, 4000, 2)]
    d = np.diff(ts)
    assert d.min() > 245_900_000 and d.max() < 246_100_000


def test_synth_timestamps_match_oracle(oracle):
    ts = synth.hash_timestamps(3000, 2, seek_to_ns=5 * S)
    want = [oracle.hash_timestamp(synth.DELAY_NS, synth.ITEM_NS, i, 5 * S) for i in range(0, 3000, 2)]
    assert ts.tolist() == want


# ----------------------------------------------------------------- simhash

def test_simhash_known_answers(oracle):
    assert oracle.simhash32(np.array([], np.uint32)) == 0
    assert oracle.simhash32(np.array([0xDEADBEEF], np.uint32)) == 0xDEADBEEF
    assert oracle.simhash32(np.array([0xFFFFFFFF, 0], np.uint32)) == 0            # ties -> 0
    assert oracle.simhash32(np.array([0xF0F0F0F0, 0xFF00FF00, 0xFFFF0000], np.uint32)) == 0xFFF0F000
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 10, 83, 367):
        h = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
        assert oracle.simhash32(h) == pyref.simhash32(h)


# --------------------------------------------- longest_common_hash_match KATs

def lcs(oracle, sh, st, dh, dt, **kw):
    kw.setdefault("threshold", 0)
    kw.setdefault("min_opening_ns", 0)
    kw.setdefault("min_ending_ns", 0)
    return oracle.longest_common_hash_match(np.array(sh, np.uint32), np.array(st, np.uint64),
                                            np.array(dh, np.uint32), np.array(dt, np.uint64), **kw)


def test_q1_row_and_column_zero_never_match(oracle):
    # identical lists: the diagonal run cannot include index 0
    h = [7, 7, 7, 7]
    t = [0, S, 2 * S, 3 * S]
    out = lcs(oracle, h, t, h, t)
    by_end = {(e[10], e[11]): e for e in out}
    e = by_end[(3, 3)]
    assert e[0] == 3                       # cells (1,1),(2,2),(3,3)
    assert (e[1], e[2]) == (0, 3 * S)      # Q3: start index = 3 - 3 = 0, one before the first matching cell
    # every other diagonal too: (3,2) has cells (2,1),(3,2)
    assert by_end[(3, 2)][0] == 2 and by_end[(2, 3)][0] == 2
    assert all(e[10] >= 1 and e[11] >= 1 for e in out)


def test_q2_runs_ending_on_last_row_or_column_are_emitted(oracle):
    sh = [1, 2, 3, 4, 5]
    dh = [9, 9, 3, 4]          # matches at (2,2),(3,3); dst ends at j = 3
    st = [k * S for k in range(5)]
    dt = [k * S for k in range(4)]
    out = lcs(oracle, sh, st, dh, dt)
    assert [(e[0], e[10], e[11]) for e in out] == [(2, 3, 3)]
    e = out[0]
    assert (e[1], e[2], e[3], e[4]) == (1 * S, 3 * S, 1 * S, 3 * S)
    # Q3: the simhash range includes the element before the run: src[1..=3], dst[1..=3]
    assert e[5] == oracle.simhash32(np.array(sh[1:4], np.uint32))
    assert e[6] == oracle.simhash32(np.array(dh[1:4], np.uint32))


def test_q5_duration_filter_is_inclusive_and_two_sided(oracle):
    sh = [0, 5, 5, 5, 0]
    dh = [1, 5, 5, 5, 1]
    st = [0, 10, 20, 30, 40]
    dt = [0, 10, 20, 29, 40]
    # run (1,1),(2,2),(3,3): src span ts[3]-ts[0] = 30, dst span 29; other diagonals give shorter runs
    spans = lambda out: sorted((e[0], e[10], e[11]) for e in out)
    assert (3, 3, 3) in spans(lcs(oracle, sh, st, dh, dt, min_opening_ns=29))
    assert (3, 3, 3) not in spans(lcs(oracle, sh, st, dh, dt, min_opening_ns=30))     # dst fails
    assert (3, 3, 3) in spans(lcs(oracle, sh, st, dh, dt, min_opening_ns=1000, min_ending_ns=29, is_opening=False))
    with pytest.raises(OverflowError):
        lcs(oracle, sh, [0, 10, 20, 5, 40][:0] + [50, 10, 20, 30, 40], dh, dt)        # ts[3] - ts[0] underflows


def test_q4_heap_array_order(oracle):
    """heap.into() is the BinaryHeap's array after pushes in (i desc, j desc) order."""
    # three disjoint runs of lengths 2, 4, 3 on different diagonals
    sh = [0, 1, 2, 0, 3, 4, 5, 6, 0, 7, 8, 9, 0]
    dh = [100, 7, 8, 9, 100, 3, 4, 5, 6, 100, 1, 2, 100]
    st = [k * S for k in range(len(sh))]
    dt = [k * S for k in range(len(dh))]
    out = lcs(oracle, sh, st, dh, dt)
    got = [(e[0], e[10], e[11]) for e in out]
    # push order: i=11 (len 3, j=3), i=7 (len 4, j=8), i=2 (len 2, j=11)
    # heap after pushes 3, 4, 2 -> [4, 3, 2]
    assert got == [(4, 7, 8), (3, 11, 3), (2, 2, 11)]
    assert out == [tuple(x) for x in map(lambda e: e[:7] + (0 if e[7] else 1,) + e[11:13] + e[13:15],
                                         pyref.longest_common_hash_match(lists(sh, st), lists(dh, dt), 0, 0, 0,
                                                                         300_000_012, 300_000_012, True))]


def test_empty_inputs(oracle):
    assert lcs(oracle, [], [], [1, 2], [0, 1]) == []
    assert lcs(oracle, [1], [0], [1], [0]) == []


@pytest.mark.parametrize("seed", range(8))
def test_c_oracle_equals_python_transcription(oracle, seed):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(1, 40)), int(rng.integers(1, 40))
    alphabet = rng.integers(0, 2 ** 32, 3, dtype=np.uint64).astype(np.uint32)
    sh = alphabet[rng.integers(0, 3, n)] ^ (np.uint32(1) << rng.integers(0, 32, n).astype(np.uint32))
    dh = alphabet[rng.integers(0, 3, m)] ^ (np.uint32(1) << rng.integers(0, 32, m).astype(np.uint32))
    st = np.cumsum(rng.integers(1, 400_000_000, n)).astype(np.uint64)
    dt = np.cumsum(rng.integers(1, 400_000_000, m)).astype(np.uint64)
    thr = int(rng.integers(0, 5))
    mn = int(rng.integers(0, 900_000_000))
    is_opening = bool(seed % 2)
    got = oracle.longest_common_hash_match(sh, st, dh, dt, threshold=thr, min_opening_ns=mn, min_ending_ns=mn // 2,
                                           src_hash_duration_ns=11, dst_hash_duration_ns=22, is_opening=is_opening)
    want = pyref.longest_common_hash_match(lists(sh, st), lists(dh, dt), thr, mn, mn // 2, 11, 22, is_opening)
    want = [e[:7] + (0 if e[7] else 1,) + e[11:13] + e[13:15] for e in want]
    assert got == want
    assert len(got) > 0 or mn > 0


@pytest.mark.parametrize("seed", range(4))
def test_run_with_frame_hashes_equals_python_transcription(oracle, seed):
    season = synth.make_hash_season(4, 48, 30, seed=seed, run_len=20, flip_p=0.04)
    season.hash_duration_ns[:] = [300_000_012, 123_000_000, 300_000_012, 500_000_000]
    kw = H.params_kw(threshold=8 + seed, include_endings=seed % 2 == 0, min_opening_ns=2 * S, min_ending_ns=S,
                     time_padding_ns=seed * 100_000_000)
    st, got, _ = H.oracle_run(oracle, season, **kw)
    assert st == 0
    off = season.seg_offset.astype(np.int64)
    videos = []
    for v in range(4):
        videos.append(dict(opening=lists(season.hashes[off[2 * v]:off[2 * v + 1]], season.ts_ns[off[2 * v]:off[2 * v + 1]]),
                           ending=lists(season.hashes[off[2 * v + 1]:off[2 * v + 2]], season.ts_ns[off[2 * v + 1]:off[2 * v + 2]]),
                           hash_duration_ns=int(season.hash_duration_ns[v])))
    want = pyref.run_with_frame_hashes(videos, **kw)
    assert got == want
    assert any(r[0] for r in got)


def test_threaded_oracle_equals_serial(oracle):
    season = synth.make_hash_season(6, 300, 150, seed=9, run_len=120)
    s = oracle.Season(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns)
    a = oracle.run_with_frame_hashes(s, include_endings=True, n_threads=1, want_entries=True)
    b = oracle.run_with_frame_hashes(s, include_endings=True, n_threads=4, want_entries=True)
    assert a == b


def _identical_pair_season():
    """Two videos with the same 12-hash opening and 5-hash ending."""
    h = np.arange(1, 13, dtype=np.uint32) * np.uint32(2654435761)
    e = np.arange(1, 6, dtype=np.uint32) * np.uint32(40503)
    ts = lambda n, seek=0: synth.hash_timestamps(2 * n, 2, seek_to_ns=seek)[:n]
    return synth.HashSeason(np.concatenate([h, e, h, e]),
                            np.concatenate([ts(12), ts(5, 10 ** 12), ts(12), ts(5, 10 ** 12)]),
                            np.array([0, 12, 17, 29, 34], np.uint64), np.full(2, synth.HASH_DURATION_NS, np.uint64))


def test_q6_threshold_zero_finds_runs_but_votes_for_nothing(oracle):
    """comparator.rs:441 skips a candidate pair when dist >= T + T/2: with T = 0 that is every pair, the
    candidate itself included, so every cluster is empty and find_best_match has no winner although the
    pair tables hold runs (both videos are `present`).  Both the C oracle and the Python transcription."""
    season = _identical_pair_season()
    kw = H.params_kw(threshold=0, include_endings=True, min_opening_ns=S, min_ending_ns=S // 2)
    st, got, entries = H.oracle_run(oracle, season, **kw)
    assert st == 0
    assert got == [(1, 0, 0, 0, 0, 0, 0), (1, 0, 0, 0, 0, 0, 0)]
    assert len(H.entries_as_runs(entries)) == 2      # the opening run (11 cells: row/column 0 never match) and the ending run
    kw["threshold"] = 2
    st, got, _ = H.oracle_run(oracle, season, **kw)
    assert st == 0 and all(r[:3] == (1, 1, 1) for r in got)
    # opening: from the timestamp of index 0 (Q3: one before the first matching cell) to index 11 minus the hash duration
    assert got[0][3] == int(season.ts_ns[0]) and got[0][4] == int(season.ts_ns[11]) - synth.HASH_DURATION_NS
