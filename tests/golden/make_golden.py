"""Generates the committed fixtures in this directory from the CPU oracle.

    python tests/golden/make_golden.py

The reference has no golden vectors for this path and
cannot be run here (Rust, no cargo), so these pin the ORACLE's outputs on seeded
inputs: a regression guard for the oracle itself (tests/test_golden.py, CPU) and
a fixed target for the CUDA path (tests/test_golden.py -m gpu)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from needle_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def match_fixture(name, season, **kw):
    s = orc.Season(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns)
    st, results, entries = orc.run_with_frame_hashes(s, want_entries=True, **kw)
    assert st == 0
    ent = np.array([(p,) + e for p, e in entries], dtype=np.uint64).reshape(-1, 13)
    np.savez_compressed(os.path.join(HERE, name), hashes=season.hashes, ts_ns=season.ts_ns,
                        seg_offset=season.seg_offset, hash_duration_ns=season.hash_duration_ns,
                        params=np.array([kw.get("threshold", 10), int(kw.get("include_endings", False)),
                                         kw.get("min_opening_ns", 20_000_000_000),
                                         kw.get("min_ending_ns", 20_000_000_000),
                                         kw.get("time_padding_ns", 0)], dtype=np.uint64),
                        entries=ent, results=np.array(results, dtype=np.uint64))
    print(name, "entries", len(entries), "videos with opening", sum(r[1] for r in results))


def fingerprint_fixture(name, seed, seconds, stereo=False):
    rng = np.random.default_rng(seed)
    n = int(seconds * synth.SAMPLE_RATE)
    x = synth._noise(rng, n, amp=0.1) + np.pad(synth._chords(rng, seconds), (0, n))[:n]
    pcm = np.clip(np.rint(x * 32767), -32768, 32767).astype(np.int16)
    if stereo:
        right = np.clip(pcm.astype(np.int32) // 3 + rng.integers(-2000, 2000, n), -32768, 32767).astype(np.int16)
        pcm = np.stack([pcm, right], axis=1).reshape(-1)
    raw, chroma = orc.fingerprint(pcm, channels=2 if stereo else 1, want_chroma=True)
    h, t = orc.subsample_and_stamp(raw, 2, seek_to_ns=77_000_000_000)
    np.savez_compressed(os.path.join(HERE, name), pcm=pcm, channels=np.array([2 if stereo else 1]),
                        raw=raw, chroma=chroma.astype(np.float64), stored_hash=h, stored_ts=t)
    print(name, "raw hashes", raw.size)


if __name__ == "__main__":
    match_fixture("match_defaults.npz", synth.make_hash_season(6, 520, 300, seed=101, run_len=170, jitter_len=True),
                  include_endings=True)
    match_fixture("match_short_runs.npz", synth.make_hash_season(5, 120, 80, seed=102, run_len=40, flip_p=0.05,
                                                                 correlated=True),
                  threshold=12, include_endings=True, min_opening_ns=1_000_000_000, min_ending_ns=0,
                  time_padding_ns=250_000_000)
    fingerprint_fixture("fingerprint_mono.npz", 201, 16.0)
    fingerprint_fixture("fingerprint_stereo.npz", 202, 9.0, stereo=True)
