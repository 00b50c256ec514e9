"""Host-only part of the product (vote.cpp through nb200_vote) against the
oracle, on CPU: feed the oracle's own run list (re-sorted into the library's
push order) to nb200_vote and require identical SearchResults.  This pins the
BinaryHeap replay and find_best_match (needle/src/audio/comparator.rs:231-249,
405-515) without a GPU."""
import numpy as np
import pytest

from needle_b200 import engine, synth
from tests import helpers as H


rows_to_runs = H.rows_to_runs


CASES = [
    dict(threshold=10, include_endings=True),
    dict(threshold=10, include_endings=False),
    dict(threshold=14, include_endings=True, min_opening_ns=0, min_ending_ns=0),
    dict(threshold=12, include_endings=True, min_opening_ns=3_000_000_000, min_ending_ns=1_000_000_000,
         time_padding_ns=700_000_000),
    dict(threshold=0, include_endings=True, min_opening_ns=0, min_ending_ns=0),
    dict(threshold=1, include_endings=True, min_opening_ns=0, min_ending_ns=0),
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("seed", [0, 1])
def test_vote_matches_oracle(oracle, case, seed):
    kw = H.params_kw(**CASES[case])
    if kw["min_opening_ns"] == 0:
        season = synth.make_hash_season(5, 70, 40, seed=seed, run_len=30, flip_p=0.05)
    else:
        season = synth.make_hash_season(7, 420, 260, seed=seed, run_len=150, jitter_len=True)
    season.hash_duration_ns[:] = 300_000_012 + np.arange(season.n_videos, dtype=np.uint64) * np.uint64(1000)
    st, want, entries = H.oracle_run(oracle, season, **kw)
    assert st == 0
    runs = rows_to_runs(H.entries_as_runs(entries))
    got = engine.vote(season.hash_duration_ns, engine.match_params(**kw), runs)
    assert got == want
    if kw["threshold"] >= 10:
        assert any(r[0] for r in want)


def test_vote_rejects_unsorted_runs(oracle):
    season = synth.make_hash_season(4, 300, 0, seed=4, run_len=120)
    kw = H.params_kw()
    _, _, entries = H.oracle_run(oracle, season, **kw)
    rows = H.entries_as_runs(entries)
    assert len(rows) >= 2
    runs = rows_to_runs(rows[::-1])
    from needle_b200._lib import Nb200Error, ERR_INVALID_ARGUMENT
    with pytest.raises(Nb200Error) as e:
        engine.vote(season.hash_duration_ns, engine.match_params(**kw), runs)
    assert e.value.status == ERR_INVALID_ARGUMENT


def test_heap_order_matters(oracle):
    """Many equal-length runs in one table: the result depends on the replayed
    heap layout (ties in find_best_match broken by candidate index)."""
    rng = np.random.default_rng(5)
    n = 200
    base = rng.integers(0, 2 ** 32, 40, dtype=np.uint64).astype(np.uint32)
    vids = []
    for v in range(3):
        h = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
        for at in (5 + 3 * v, 60 + v, 120 + 2 * v):     # three copies each -> 9 equal runs per pair
            h[at:at + 40] = base
        vids.append((h, synth.hash_timestamps(2 * n, 2)[:n]))
    season = H.season_from_lists(vids, [(np.zeros(0, np.uint32), np.zeros(0, np.uint64))] * 3)
    kw = H.params_kw(threshold=0, min_opening_ns=5_000_000_000)
    st, want, entries = H.oracle_run(oracle, season, **kw)
    assert st == 0 and len(entries) >= 27
    runs = rows_to_runs(H.entries_as_runs(entries))
    got = engine.vote(season.hash_duration_ns, engine.match_params(**kw), runs)
    assert got == want


def test_vote_subset_partitions_the_videos(oracle):
    """nb200_vote_subset over disjoint video masks reassembles nb200_vote (the
    multi-GPU job splits the vote this way)."""
    season = synth.make_hash_season(7, 420, 260, seed=6, run_len=150, jitter_len=True)
    kw = H.params_kw(include_endings=True)
    _, want, entries = H.oracle_run(oracle, season, **kw)
    runs = rows_to_runs(H.entries_as_runs(entries))
    p = engine.match_params(**kw)
    full = engine.vote_array(season.hash_duration_ns, p, runs)
    assert engine.results_as_tuples(full) == want
    merged = np.zeros_like(full)
    for r in range(3):
        mask = np.zeros(7, np.uint8)
        mask[r::3] = 1
        part = engine.vote_array(season.hash_duration_ns, p, runs, video_mask=mask)
        assert not part["present"][mask == 0].any()
        merged[r::3] = part[r::3]
    assert np.array_equal(merged, full)
