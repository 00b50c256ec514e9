"""GPU parity of the fingerprint stage (K1/K2) against the double-precision CPU
oracle (Chromaprint TEST2 restated), through the C ABI.  The FFT is FP32 on
the GPU and FP64 in the oracle, so hashes must agree bitwise on >= 99.5 % of
frames (BASELINE.json north_star); everything downstream of the hashes
(subsampling, timestamps) is exact."""
import numpy as np
import pytest

from needle_b200 import engine, synth
from needle_b200._lib import Nb200Error, ERR_INVALID_ARGUMENT, ERR_STATE
from tests import helpers as H

pytestmark = pytest.mark.gpu

FRAME_AGREEMENT = 0.995


def agreement(a, b):
    assert a.shape == b.shape
    return float(np.mean(a == b)) if a.size else 1.0


def make_pcm(seed, seconds, kind="mix"):
    rng = np.random.default_rng(seed)
    n = int(seconds * synth.SAMPLE_RATE)
    if kind == "noise":
        x = synth._noise(rng, n)
    elif kind == "chords":
        x = synth._chords(rng, seconds)[:n]
    else:
        x = synth._noise(rng, n, amp=0.1) + np.pad(synth._chords(rng, seconds), (0, n))[:n]
    return np.clip(np.rint(x * 32767.0), -32768, 32767).astype(np.int16)


@pytest.mark.parametrize("variant", [0, 1, 8, 12, 16, 17, 18, 112])
@pytest.mark.parametrize("kind", ["noise", "chords", "mix"])
def test_raw_hashes_match_oracle(ctx, oracle, kind, variant):
    """Every K1 kernel variant (NB200_OPT_K1_VARIANT: 0 = the default tensor-memory kernel; 16/112 its
    first arithmetic revision with 16/12 warps; 8/12: parked half in shared memory; 1: 64 values per lane)."""
    from needle_b200._lib import OPT_K1_VARIANT
    pcm = make_pcm(1, 120.0, kind)
    want = oracle.fingerprint(pcm)
    ctx.set_option(OPT_K1_VARIANT, variant)
    try:
        got = ctx.fingerprint_batch([pcm[3:], pcm])[1]     # the first segment makes this one start mid-buffer
    finally:
        ctx.set_option(OPT_K1_VARIANT, 0)
    assert got.shape == want.shape == (oracle.num_raw_hashes(pcm.size),)
    assert agreement(got, want) >= FRAME_AGREEMENT
    # the disagreeing frames differ in a few classifier bits only
    bad = got != want
    if bad.any():
        bits = np.array([bin(int(x)).count("1") for x in (got[bad] ^ want[bad])])
        assert bits.max() <= 4


def test_batch_of_ragged_segments(ctx, oracle):
    lens = [0, 100, 4095, 4096, 4096 + 1365 * 19 - 1, 4096 + 1365 * 19, 4096 + 1365 * 19 + 1,
            4096 + 1365 * 146, 4096 + 1365 * 147, 200_001, 333_333]
    segs = [make_pcm(10 + k, n / synth.SAMPLE_RATE + 0.001)[:n] for k, n in enumerate(lens)]
    got = ctx.fingerprint_batch(segs)
    agree = total = 0
    for pcm, g in zip(segs, got):
        w = oracle.fingerprint(pcm)
        assert g.shape == w.shape
        agree += int(np.sum(g == w))
        total += w.size
    assert total > 0 and agree / total >= FRAME_AGREEMENT


def test_stereo_downmix(ctx, oracle):
    rng = np.random.default_rng(3)
    left = make_pcm(20, 30.0, "mix")
    right = np.clip(left.astype(np.int32) // 2 + rng.integers(-3000, 3000, left.size), -32768, 32767).astype(np.int16)
    right[:100] = -32768   # (L+R)/2 truncates toward zero for negative odd sums
    left[:100:2] = 32767
    inter = np.empty(2 * left.size, np.int16)
    inter[0::2] = left
    inter[1::2] = right
    want = oracle.fingerprint(inter, channels=2)
    got = ctx.fingerprint_batch([inter], channels=2)[0]
    assert agreement(got, want) >= FRAME_AGREEMENT
    # equals fingerprinting the oracle's own mono mix
    mono = ((left.astype(np.int32) + right.astype(np.int32)) / 2).astype(np.int64)   # trunc toward zero
    mono = np.trunc((left.astype(np.int32) + right.astype(np.int32)) / 2).astype(np.int16)
    assert np.array_equal(ctx.fingerprint_batch([mono])[0], got)


def test_silence_and_full_scale(ctx, oracle):
    n = 4096 + 1365 * 60
    silence = np.zeros(n, np.int16)
    got = ctx.fingerprint_batch([silence])[0]
    want = oracle.fingerprint(silence)
    assert np.array_equal(got, want)           # all-zero features: one fixed hash, exact
    assert len(set(got.tolist())) == 1
    sq = np.where((np.arange(n) // 25) % 2 == 0, 32767, -32768).astype(np.int16)   # 220.5 Hz square wave
    assert agreement(ctx.fingerprint_batch([sq])[0], oracle.fingerprint(sq)) >= FRAME_AGREEMENT


def test_upstream_chromaprint_silence_known_answer(ctx):
    """Chromaprint's own API test (tests/test_api.cpp, Test2SilenceRawFp): 130 x 1024 zero samples at
    44100 Hz -> raw fingerprint [627964279] * 3.  At 11025 Hz that input is 33280 zeros.  The literal
    is upstream's, not the oracle's (tests/test_oracle_kat.py)."""
    for channels in (1, 2):
        pcm = np.zeros(33280 * channels, np.int16)
        assert ctx.fingerprint_batch([pcm], channels=channels)[0].tolist() == [627964279] * 3


def test_stride_and_timestamps_exact(ctx, oracle):
    """The subsample/timestamp tail of process_frames (analyzer.rs:288-318) is integer/f32-exact."""
    pcm_o = make_pcm(30, 95.0)
    pcm_e = make_pcm(31, 41.0)
    seek = [0, 1_234_567_890_123]
    ps = engine.PcmSet.upload(ctx, [pcm_o, pcm_e])
    for stride in (1, 2, 3, 7):
        hs = ps.fingerprint(stride=stride, seek_to_ns=seek)
        h, t, off = hs.download()
        raw = ctx.fingerprint_batch([pcm_o, pcm_e])
        for k, (r, sk) in enumerate(zip(raw, seek)):
            wh, wt = oracle.subsample_and_stamp(r, stride, seek_to_ns=sk)
            a, b = int(off[k]), int(off[k + 1])
            assert np.array_equal(h[a:b], wh)
            assert np.array_equal(t[a:b], wt)


def test_streaming_shim_matches_batch(ctx, oracle):
    """nb200_fp_* = the chromaprint::Context call sequence of process_frames."""
    import ctypes as C
    from needle_b200._lib import lib, check
    pcm = make_pcm(40, 20.0)
    stereo = np.repeat(pcm, 2)
    L = lib()
    fp = C.c_void_p()
    check(L.nb200_fp_new(ctx.handle, C.byref(fp)), "fp_new")
    try:
        assert L.nb200_fp_sample_rate(fp) == 11025
        assert L.nb200_fp_feed(fp, stereo.ctypes.data_as(C.c_void_p), 10) == ERR_STATE     # before start
        assert L.nb200_fp_start(fp, 44100, 2) == ERR_INVALID_ARGUMENT
        check(L.nb200_fp_start(fp, 11025, 2), "fp_start")
        pos = 0
        rng = np.random.default_rng(0)
        while pos < stereo.size:   # irregular chunks, like resampler output
            n = min(stereo.size - pos, 2 * int(rng.integers(1, 5000)))
            chunk = np.ascontiguousarray(stereo[pos:pos + n])
            check(L.nb200_fp_feed(fp, chunk.ctypes.data_as(C.c_void_p), n), "fp_feed")
            pos += n
        check(L.nb200_fp_finish(fp), "fp_finish")
        d, it = C.c_int(), C.c_int()
        check(L.nb200_fp_get_delay_ms(fp, C.byref(d)), "delay")
        check(L.nb200_fp_get_item_duration_ms(fp, C.byref(it)), "item")
        assert (d.value, it.value) == (2600, 123)
        hp, n = C.c_void_p(), C.c_size_t()
        check(L.nb200_fp_get_raw(fp, C.byref(hp), C.byref(n)), "get_raw")
        got = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint32)), shape=(n.value,)).copy()
    finally:
        L.nb200_fp_free(fp)
    assert np.array_equal(got, ctx.fingerprint_batch([pcm])[0])     # L == R: same mono signal
    assert agreement(got, oracle.fingerprint(pcm)) >= FRAME_AGREEMENT


def test_analyze_search_intervals_vs_oracle(ctx, oracle):
    """PCM season -> fingerprints -> match -> vote, against the oracle end to end:
    intervals within one hash period (north_star), here 2 x 123 ms."""
    eps = synth.make_pcm_season(4, 6.0, season_seed=7, intro_s=45.0, credits_s=40.0)
    segs, seek, o_open, o_end = [], [], [], []
    for ep in eps:
        a, b, sk = synth.split_segments(ep.pcm)
        segs += [a, b]
        seek += [0, sk]
        o_open.append(oracle.subsample_and_stamp(oracle.fingerprint(a), 2))
        o_end.append(oracle.subsample_and_stamp(oracle.fingerprint(b), 2, seek_to_ns=sk))
    p = engine.match_params(include_endings=True)
    got = ctx.analyze_search(segs, 1, seek, synth.HASH_DURATION_NS, p)
    season = H.season_from_lists(o_open, o_end)
    st, want, _ = H.oracle_run(oracle, season, **H.params_kw(include_endings=True))
    assert st == 0
    tol = 2 * 123_000_000
    for g, w, ep in zip(got, want, eps):
        assert g[:3] == w[:3] == (1, 1, 1)
        for a, b in zip(g[3:], w[3:]):
            assert abs(int(a) - int(b)) <= tol
        # and the detected opening really is where the intro was spliced
        assert abs(g[3] / 1e9 - ep.intro_at) < 4.0


def test_fingerprint_into_and_hashset_view(ctx):
    """The zero-copy multi-GPU plumbing: K2 writing into caller-owned device arrays
    (nb200_fingerprint_run_into) and matching out of caller-owned arrays in any segment
    placement (nb200_hashset_view) give what the ordinary owned path gives."""
    import torch
    eps = synth.make_pcm_season(4, 4.0, season_seed=9, intro_s=40.0, credits_s=30.0)
    segs, seek = [], []
    for ep in eps:
        a, b, sk = synth.split_segments(ep.pcm)
        segs += [a, b]
        seek += [0, sk]
    ps = engine.PcmSet.upload(ctx, segs)
    hs = ps.fingerprint(stride=2, seek_to_ns=seek)
    h, t, off = hs.download()
    doff, dlen, total = engine.fingerprint_layout([s.size for s in segs], 2)
    assert [int(x) for x in dlen] == [int(off[k + 1] - off[k]) for k in range(len(segs))]
    hb = torch.zeros(total + 8, dtype=torch.int32, device="cuda")
    tb = torch.zeros(total + 8, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ps.fingerprint_into(hb.data_ptr(), tb.data_ptr(), total + 8, stride=2, seek_to_ns=seek)
    ctx.synchronize()
    hv, tv = hb.cpu().numpy().view(np.uint32), tb.cpu().numpy().view(np.uint64)
    for k in range(len(segs)):
        a, n = int(doff[k]), int(dlen[k])
        assert np.array_equal(hv[a:a + n], h[int(off[k]):int(off[k + 1])])
        assert np.array_equal(tv[a:a + n], t[int(off[k]):int(off[k + 1])])
    # the timestamps are a function of the index: nb200_timestamps_fill reproduces K2's
    tf = torch.zeros(total + 8, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.timestamps_fill(tf.data_ptr(), doff, dlen, seek, stride=2)
    ctx.synchronize()
    assert np.array_equal(tf.cpu().numpy().view(np.uint64), tv)
    with pytest.raises(Nb200Error):
        ps.fingerprint_into(hb.data_ptr(), tb.data_ptr(), total - 4, stride=2, seek_to_ns=seek)   # too small
    p = engine.match_params(include_endings=True)
    want = hs.match(p).download()
    assert want.shape[0] >= 6
    view = engine.HashSet.view(ctx, hb.data_ptr(), tb.data_ptr(), doff, dlen, keepalive=(hb, tb))
    assert np.array_equal(view.match(p).download(), want)
    hd = np.full(4, synth.HASH_DURATION_NS, np.uint64)
    assert view.search(hd, p) == hs.search(hd, p)
    # videos in reverse order through the offsets alone == uploading the reversed season
    order = [3, 2, 1, 0]
    roff = np.array([doff[2 * v + e] for v in order for e in (0, 1)], np.uint64)
    rlen = np.array([dlen[2 * v + e] for v in order for e in (0, 1)], np.uint64)
    rview = engine.HashSet.view(ctx, hb.data_ptr(), tb.data_ptr(), roff, rlen, keepalive=(hb, tb))
    hs_l, ts_l, off_l = [], [], [0]
    for v in order:
        for e in (0, 1):
            k = 2 * v + e
            hs_l.append(h[int(off[k]):int(off[k + 1])])
            ts_l.append(t[int(off[k]):int(off[k + 1])])
            off_l.append(off_l[-1] + hs_l[-1].size)
    rev = engine.HashSet.upload(ctx, np.concatenate(hs_l), np.concatenate(ts_l), np.asarray(off_l, np.uint64))
    assert np.array_equal(rview.match(p).download(), rev.match(p).download())
    assert rview.search(hd, p) == rev.search(hd, p)
    with pytest.raises(Nb200Error) as e:
        engine.HashSet.view(ctx, hb.data_ptr(), tb.data_ptr(), doff + np.uint64(1), dlen)     # not 4-aligned
    assert e.value.status == ERR_INVALID_ARGUMENT
