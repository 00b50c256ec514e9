"""Oracle parity at the FULL sizes of BASELINE configs[2] and configs[3].

The reference's table fill (comparator.rs:175-187: 8 bytes per cell, two passes) is what
makes these slow on the CPU -- the oracle restates it literally -- so they run once each,
multi-threaded like the reference's rayon pair loop (comparator.rs:549-564): about 5 s
for 12 x 60 min and about a minute for the 19,900 pairs of 200 x 24 min on a 16-core box.
Bar: bit-exact run lists (where compared) and bit-exact per-video intervals.
"""
import os

import numpy as np
import pytest

from needle_b200 import engine, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def oracle_results(orc, season, **kw):
    s = orc.Season(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns)
    st, res, entries = orc.run_with_frame_hashes(s, n_threads=os.cpu_count() or 1, want_entries=True, **kw)
    assert st == 0
    return [tuple(int(x) for x in r) for r in res], entries


def test_config2_12x60min_runs_and_intervals_equal_oracle(ctx, oracle):
    """BASELINE configs[2]: 12 episodes x 60 min (7,259 / 3,624 hashes), 66 pairs, openings + endings:
    every run (indices, length, simhashes, timestamps, order) and every interval."""
    season = synth.make_hash_season(12, 7259, 3624, seed=60)
    p = engine.match_params(include_endings=True)
    want, entries = oracle_results(oracle, season, include_endings=True)
    runs = ctx.match_pairs(season.hashes, season.ts_ns, season.seg_offset, p)
    assert H.runs_as_rows(runs) == H.entries_as_runs(entries)
    got = ctx.search(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, p)
    assert [tuple(r) for r in got] == want
    assert sum(r[1] for r in got) == 12 and sum(r[2] for r in got) >= 11


def test_config3_200x24min_intervals_equal_oracle(ctx, oracle):
    """BASELINE configs[3]: 200 episodes x 24 min, 19,900 pairs, 2.08e11 cells: the device path
    (match, simhash, device vote) against the oracle's full tables and find_best_match --
    all 200 per-video results, and the run lists of the first 400 pairs."""
    season = synth.make_hash_season(200, 2897, 1443, seed=4)
    p = engine.match_params(include_endings=True)
    want, entries = oracle_results(oracle, season, include_endings=True)
    hs = engine.HashSet.upload(ctx, season.hashes, season.ts_ns, season.seg_offset)
    got = hs.search(season.hash_duration_ns, p)
    assert [tuple(r) for r in got] == want
    assert sum(r[1] for r in got) >= 190 and sum(r[2] for r in got) >= 190
    pairs = np.array([(i, j) for i in range(200) for j in range(i + 1, 200)][:400], dtype=np.uint32)
    runs = hs.match(p, pairs=pairs).download()
    rows = [r for r in H.entries_as_runs(entries) if r[0] < 400]
    assert H.runs_as_rows(runs) == rows
    hs.free()
