"""Committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py
from the oracle): the oracle must keep reproducing them (CPU), the host vote
must reproduce the results from the stored runs (CPU), and the CUDA path must
reproduce them through the C ABI (GPU)."""
import os

import numpy as np
import pytest

from needle_b200 import engine, synth
from tests import helpers as H

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MATCH = ["match_defaults.npz", "match_short_runs.npz"]
FP = ["fingerprint_mono.npz", "fingerprint_stereo.npz"]


def load_match(name):
    z = np.load(os.path.join(G, name))
    season = synth.HashSeason(z["hashes"], z["ts_ns"], z["seg_offset"], z["hash_duration_ns"])
    p = [int(x) for x in z["params"]]
    kw = H.params_kw(threshold=p[0], include_endings=bool(p[1]), min_opening_ns=p[2], min_ending_ns=p[3],
                     time_padding_ns=p[4])
    entries = [(int(r[0]), tuple(int(x) for x in r[1:])) for r in z["entries"]]
    results = [tuple(int(x) for x in r) for r in z["results"]]
    return season, kw, entries, results


def golden_runs(entries):
    rows = H.entries_as_runs(entries)
    return rows, H.rows_to_runs(rows)


@pytest.mark.parametrize("name", MATCH)
def test_oracle_reproduces_match_golden(oracle, name):
    season, kw, entries, results = load_match(name)
    st, got_results, got_entries = H.oracle_run(oracle, season, **kw)
    assert st == 0 and got_results == results
    assert [(p, tuple(e)) for p, e in got_entries] == entries
    assert len(entries) > 5


@pytest.mark.parametrize("name", MATCH)
def test_host_vote_reproduces_match_golden(name):
    season, kw, entries, results = load_match(name)
    _, runs = golden_runs(entries)
    assert engine.vote(season.hash_duration_ns, engine.match_params(**kw), runs) == results


@pytest.mark.parametrize("name", FP)
def test_oracle_reproduces_fingerprint_golden(oracle, name):
    z = np.load(os.path.join(G, name))
    raw, chroma = oracle.fingerprint(z["pcm"], channels=int(z["channels"][0]), want_chroma=True)
    assert np.array_equal(raw, z["raw"]) and np.allclose(chroma, z["chroma"], rtol=1e-12, atol=0)
    h, t = oracle.subsample_and_stamp(raw, 2, seek_to_ns=77_000_000_000)
    assert np.array_equal(h, z["stored_hash"]) and np.array_equal(t, z["stored_ts"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", MATCH)
def test_gpu_reproduces_match_golden(ctx, name):
    season, kw, entries, results = load_match(name)
    rows, _ = golden_runs(entries)
    p = engine.match_params(**kw)
    assert H.runs_as_rows(ctx.match_pairs(season.hashes, season.ts_ns, season.seg_offset, p)) == rows
    assert ctx.search(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, p) == results


@pytest.mark.gpu
@pytest.mark.parametrize("name", FP)
def test_gpu_reproduces_fingerprint_golden(ctx, name):
    z = np.load(os.path.join(G, name))
    ch = int(z["channels"][0])
    got = ctx.fingerprint_batch([z["pcm"]], channels=ch)[0]
    assert got.shape == z["raw"].shape
    # FP32 FFT vs the FP64 oracle: >= 99.5 % of frames (north_star); on these short streams at most 1 frame
    assert int(np.sum(got != z["raw"])) <= max(1, int(0.005 * got.size))
    # stored form: stride 2 + timestamps, exact given the GPU's own raw hashes
    seg = z["pcm"]
    ps = engine.PcmSet.upload(ctx, [seg, np.zeros(0, np.int16)], channels=ch)
    h, t, off = ps.fingerprint(stride=2, seek_to_ns=[77_000_000_000, 0]).download()
    assert np.array_equal(h[:int(off[1])], got[::2])
    assert np.array_equal(t[:int(off[1])], z["stored_ts"])
