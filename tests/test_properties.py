"""Property tests (hypothesis): the C oracle against the independent Python
transcription on arbitrary small inputs (CPU), and the CUDA match path against
the oracle on arbitrary seasons (GPU).  Small alphabets of hashes make long
accidental runs, ties and boundary cases likely."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from needle_b200 import engine, synth
from oracle import pyref
from tests import helpers as H

SETTINGS = dict(deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


def hash_list(draw, n, alphabet, flip_bits):
    base = draw(st.lists(st.integers(0, len(alphabet) - 1), min_size=n, max_size=n))
    flips = draw(st.lists(st.integers(0, 31), min_size=n, max_size=n))
    do = draw(st.lists(st.booleans(), min_size=n, max_size=n))
    return np.array([alphabet[b] ^ ((1 << f) if (d and flip_bits) else 0) for b, f, d in zip(base, flips, do)],
                    dtype=np.uint32)


@st.composite
def lcs_case(draw):
    n = draw(st.integers(0, 24))
    m = draw(st.integers(0, 24))
    alphabet = draw(st.lists(st.integers(0, 2 ** 32 - 1), min_size=1, max_size=3))
    sh = hash_list(draw, n, alphabet, True)
    dh = hash_list(draw, m, alphabet, True)
    st_ = np.cumsum(draw(st.lists(st.integers(1, 5 * 10 ** 8), min_size=n, max_size=n)), dtype=np.uint64) \
        if n else np.zeros(0, np.uint64)
    dt_ = np.cumsum(draw(st.lists(st.integers(1, 5 * 10 ** 8), min_size=m, max_size=m)), dtype=np.uint64) \
        if m else np.zeros(0, np.uint64)
    return sh, st_, dh, dt_, draw(st.integers(0, 4)), draw(st.integers(0, 2 * 10 ** 9)), draw(st.booleans())


@settings(max_examples=150, **SETTINGS)
@given(lcs_case())
def test_c_oracle_equals_python_transcription_property(oracle, case):
    sh, st_, dh, dt_, thr, mn, is_opening = case
    got = oracle.longest_common_hash_match(sh, st_, dh, dt_, threshold=thr, min_opening_ns=mn, min_ending_ns=mn // 3,
                                           src_hash_duration_ns=7, dst_hash_duration_ns=9, is_opening=is_opening)
    lists = lambda h, t: list(zip([int(x) for x in h], [int(x) for x in t]))
    want = pyref.longest_common_hash_match(lists(sh, st_), lists(dh, dt_), thr, mn, mn // 3, 7, 9, is_opening)
    want = [e[:7] + (0 if e[7] else 1,) + e[11:13] + e[13:15] for e in want]
    assert got == want
    # structural invariants of any result
    for e in got:
        ln, i, j = e[0], e[10], e[11]
        assert 1 <= ln <= min(i, j) and i <= len(sh) - 1 and j <= len(dh) - 1
        assert all(bin(int(sh[i - k]) ^ int(dh[j - k])).count("1") <= thr for k in range(ln))


@st.composite
def season_case(draw):
    n_videos = draw(st.integers(2, 5))
    alphabet = draw(st.lists(st.integers(0, 2 ** 32 - 1), min_size=1, max_size=4))
    openings, endings = [], []
    for _ in range(n_videos):
        no = draw(st.integers(0, 330))
        ne = draw(st.integers(1, 90))
        openings.append((hash_list(draw, no, alphabet, True), synth.hash_timestamps(2 * no, 2)[:no]))
        endings.append((hash_list(draw, ne, alphabet, True), synth.hash_timestamps(2 * ne, 2, seek_to_ns=10 ** 12)[:ne]))
    thr = draw(st.sampled_from([0, 1, 2, 10]))
    mn = draw(st.sampled_from([0, 3 * 10 ** 8, 2 * 10 ** 9, 20 * 10 ** 9]))
    return openings, endings, thr, mn


@pytest.mark.gpu
@settings(max_examples=25, **SETTINGS)
@given(season_case())
def test_gpu_runs_equal_oracle_property(ctx, oracle, case):
    openings, endings, thr, mn = case
    season = H.season_from_lists(openings, endings)
    kw = H.params_kw(threshold=thr, include_endings=True, min_opening_ns=mn, min_ending_ns=mn)
    want = H.oracle_pair_runs(oracle, season, **kw)
    got = ctx.match_pairs(season.hashes, season.ts_ns, season.seg_offset, engine.match_params(**kw))
    assert H.runs_as_rows(got) == want


@st.composite
def vote_case(draw):
    """Small seasons over a tiny hash alphabet: many equal runs, equal signatures and equal
    durations, i.e. the ties that candidate order and heap order decide."""
    n_videos = draw(st.integers(2, 6))
    alphabet = draw(st.lists(st.integers(0, 2 ** 32 - 1), min_size=1, max_size=3))
    openings, endings = [], []
    for _ in range(n_videos):
        no = draw(st.integers(2, 160))
        ne = draw(st.integers(1, 60))
        openings.append((hash_list(draw, no, alphabet, True), synth.hash_timestamps(2 * no, 2)[:no]))
        endings.append((hash_list(draw, ne, alphabet, True), synth.hash_timestamps(2 * ne, 2, seek_to_ns=10 ** 12)[:ne]))
    thr = draw(st.sampled_from([0, 1, 3, 10]))
    mn = draw(st.sampled_from([10 ** 9, 3 * 10 ** 9, 8 * 10 ** 9]))
    pad = draw(st.sampled_from([0, 0, 250_000_000, 10 ** 9]))
    hd = [draw(st.sampled_from([300_000_012, 123_000_000, 10 ** 9])) for _ in range(n_videos)]
    return openings, endings, thr, mn, pad, hd


@pytest.mark.gpu
@settings(max_examples=40, **SETTINGS)
@given(vote_case())
def test_gpu_vote_equals_oracle_property(ctx, oracle, case):
    """Final per-video intervals of the device vote (and of the host vote) against the oracle's
    find_best_match, bit for bit, including the reference's panic cases as an error status."""
    from needle_b200._lib import ERR_DURATION_UNDERFLOW, OPT_HOST_VOTE, Nb200Error
    openings, endings, thr, mn, pad, hd = case
    season = H.season_from_lists(openings, endings)
    season.hash_duration_ns[:] = np.asarray(hd, np.uint64)
    kw = H.params_kw(threshold=thr, include_endings=True, min_opening_ns=mn, min_ending_ns=mn // 2, time_padding_ns=pad)
    st_, want, _ = H.oracle_run(oracle, season, **kw)
    p = engine.match_params(**kw)
    hs = engine.HashSet.upload(ctx, season.hashes, season.ts_ns, season.seg_offset)
    try:
        for host_vote in (0, 1):
            ctx.set_option(OPT_HOST_VOTE, host_vote)
            try:
                got = hs.search(season.hash_duration_ns, p)
                assert st_ == 0 and got == want
            except Nb200Error as e:
                # end - padding - hash_duration underflows: a panic in the reference, a status here
                assert e.status == ERR_DURATION_UNDERFLOW and st_ != 0
    finally:
        ctx.set_option(OPT_HOST_VOTE, 0)
        hs.free()
