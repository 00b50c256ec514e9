"""Shared helpers of the parity tests: build seasons, run the oracle, and put
oracle entries and library runs into one comparable form."""
from __future__ import annotations

import numpy as np

from needle_b200 import synth


def season_from_lists(openings, endings, hash_duration_ns=synth.HASH_DURATION_NS):
    """openings / endings: per video (u32 hashes, u64 ts_ns).  -> synth.HashSeason"""
    hs, ts, off = [], [], [0]
    for (oh, ot), (eh, et) in zip(openings, endings):
        for h, t in ((oh, ot), (eh, et)):
            hs.append(np.asarray(h, dtype=np.uint32))
            ts.append(np.asarray(t, dtype=np.uint64))
            off.append(off[-1] + len(h))
    n = len(openings)
    return synth.HashSeason(
        np.concatenate(hs) if hs else np.zeros(0, np.uint32),
        np.concatenate(ts) if ts else np.zeros(0, np.uint64),
        np.asarray(off, dtype=np.uint64),
        np.full(n, hash_duration_ns, dtype=np.uint64))


def random_season(rng, n_videos, n_open, n_end, jitter=True, seek_ns=1_080_000_000_000):
    """Uniform random hashes, analyzer-formula timestamps, ragged lengths."""
    openings, endings = [], []
    for _ in range(n_videos):
        no = int(max(0, n_open - (rng.integers(0, max(1, n_open // 3)) if jitter else 0)))
        ne = int(max(0, n_end - (rng.integers(0, max(1, n_end // 3)) if jitter else 0)))
        openings.append((rng.integers(0, 2 ** 32, no, dtype=np.uint64).astype(np.uint32),
                         synth.hash_timestamps(2 * no, 2)[:no]))
        endings.append((rng.integers(0, 2 ** 32, ne, dtype=np.uint64).astype(np.uint32),
                        synth.hash_timestamps(2 * ne, 2, seek_to_ns=seek_ns)[:ne]))
    return season_from_lists(openings, endings)


def plant(rng, season, length, video_positions, ending=False, flips=2):
    """Copies one random run of `length` hashes into the given videos at the
    given start indices (each copy with `flips` random bit flips per hash)."""
    base = rng.integers(0, 2 ** 32, length, dtype=np.uint64).astype(np.uint32)
    for v, at in video_positions:
        a = int(season.seg_offset[2 * v + (1 if ending else 0)])
        copy = base.copy()
        for _ in range(flips):
            copy ^= (np.uint32(1) << rng.integers(0, 32, length).astype(np.uint32))
        season.hashes[a + at:a + at + length] = copy


def oracle_run(orc, season, **kw):
    """-> (status, results, entries) from oracle.run_with_frame_hashes."""
    s = orc.Season(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns)
    return orc.run_with_frame_hashes(s, want_entries=True, **kw)


def entries_as_runs(entries):
    """Oracle entries -> sorted list of (pair, is_ending, i_end, j_end, len, src_simhash, dst_simhash,
    src_start_ns, src_end_ns, dst_start_ns, dst_end_ns) in the library's order: pair, opening first,
    i desc, j desc."""
    rows = [(int(p), int(e[7]), int(e[10]), int(e[11]), int(e[0]), int(e[5]), int(e[6]),
             int(e[1]), int(e[2]), int(e[3]), int(e[4])) for p, e in entries]
    rows.sort(key=lambda r: (r[0], r[1], -r[2], -r[3]))
    return rows


RUN_FIELDS = ("pair", "is_ending", "i_end", "j_end", "len", "src_simhash", "dst_simhash",
              "src_start_ns", "src_end_ns", "dst_start_ns", "dst_end_ns")


def runs_as_rows(runs: np.ndarray):
    cols = [runs[f].tolist() for f in RUN_FIELDS]
    return list(zip(*cols))


def rows_to_runs(rows):
    from needle_b200._lib import RUN_DTYPE
    runs = np.zeros(len(rows), dtype=RUN_DTYPE)
    for k, f in enumerate(RUN_FIELDS):
        runs[f] = [r[k] for r in rows]
    return runs


def oracle_pair_runs(orc, season, pairs=None, **kw):
    """Run lists only (no vote): orc.longest_common_hash_match per pair, in the
    library's order.  For cases where the vote's O(candidates^2) would dominate."""
    n = season.n_videos
    if pairs is None:
        pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
    off = season.seg_offset.astype(np.int64)
    rows = []
    for k, (a, b) in enumerate(pairs):
        for e in ((0, 1) if kw.get("include_endings") else (0,)):
            sa, sb = 2 * a + e, 2 * b + e
            ent = orc.longest_common_hash_match(
                season.hashes[off[sa]:off[sa + 1]], season.ts_ns[off[sa]:off[sa + 1]],
                season.hashes[off[sb]:off[sb + 1]], season.ts_ns[off[sb]:off[sb + 1]],
                threshold=kw.get("threshold", 10), min_opening_ns=kw.get("min_opening_ns", 20_000_000_000),
                min_ending_ns=kw.get("min_ending_ns", 20_000_000_000), is_opening=(e == 0))
            rows += [(k, e, x[10], x[11], x[0], x[5], x[6], x[1], x[2], x[3], x[4]) for x in ent]
    rows.sort(key=lambda r: (r[0], r[1], -r[2], -r[3]))
    return rows


def params_kw(threshold=10, include_endings=False, min_opening_ns=20_000_000_000,
              min_ending_ns=20_000_000_000, time_padding_ns=0):
    return dict(threshold=threshold, include_endings=include_endings, min_opening_ns=min_opening_ns,
                min_ending_ns=min_ending_ns, time_padding_ns=time_padding_ns)
