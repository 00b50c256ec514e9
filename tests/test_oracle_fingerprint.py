"""The oracle's fingerprint stage (oracle/chromaprint_ref.c, Chromaprint 1.5.x
TEST2 restated from its published algorithm -- the crate chromaprint-sys-next
1.5.3 is not vendored in the reference) checked on CPU:
  * known answers of Chromaprint's own unit tests for the building blocks
    (chroma note mapping, chroma filter, quantizer, Gray code) -- vectors
    RECALLED from upstream's tests/, each also verified by arithmetic below;
  * the FFT against numpy's;
  * the C restatement against the independent numpy transcription
    (oracle/pyref.py) on seeded PCM.
Parity with the real Chromaprint build stays unpinned (no golden vectors in
the reference; see oracle/needle_oracle.h)."""
import numpy as np
import pytest

from needle_b200 import synth
from oracle import pyref


def test_chroma_note_mapping_upstream_kat(oracle):
    # chroma tests (upstream test_chroma.cpp): frame_size 256, 1000 Hz, band 10..510 Hz
    def fold(idx):
        p = np.zeros(129)
        p[idx] = 1.0
        return oracle.chroma_fold(p, 256, 10, 510, 1000)
    a = np.zeros(12); a[0] = 1.0          # NormalA: bin 113 = 441.4 Hz -> A
    gs = np.zeros(12); gs[11] = 1.0       # NormalGSharp: bin 112 = 437.5 Hz -> G#
    b = np.zeros(12); b[2] = 1.0          # NormalB: bin 64 = 250 Hz -> B
    assert np.array_equal(fold(113), a)
    assert np.array_equal(fold(112), gs)
    assert np.array_equal(fold(64), b)
    # arithmetic check of the recalled vectors
    for idx, note in ((113, 0), (112, 11), (64, 2)):
        octv = np.log2(idx * 1000 / 256 / 27.5)
        assert int(12 * (octv - np.floor(octv))) == note


def test_chroma_bin_range_of_needle_config(oracle):
    lo, hi, notes = oracle.chroma_notes()
    assert (lo, hi) == (10, 1308)
    plo, phi, pnotes = pyref.chroma_notes()
    assert (plo, phi) == (lo, hi) and np.array_equal(notes[lo:hi], pnotes[lo:hi])
    assert notes[lo:hi].min() == 0 and notes[lo:hi].max() == 11
    # 440 Hz = bin 163.5: bins 164.. are A (0), bin 163 is G# (11)
    assert notes[164] == 0 and notes[163] == 11


def test_chroma_filter_upstream_kat(oracle):
    # upstream test_chroma_filter.cpp "Blur2": coefficients {0.5, 0.5}
    rows = np.zeros((3, 12))
    rows[:, 0] = [0.0, 1.0, 2.0]
    rows[:, 1] = [5.0, 6.0, 7.0]
    out = oracle.chroma_filter([0.5, 0.5], rows)
    assert out.shape == (2, 12)
    assert out[:, 0].tolist() == [0.5, 1.5] and out[:, 1].tolist() == [5.5, 6.5]
    # "Blur3": {0.5, 0.7, 0.5}
    rows = np.zeros((4, 12))
    rows[:, 0] = [0.0, 1.0, 2.0, 3.0]
    rows[:, 1] = [5.0, 6.0, 7.0, 8.0]
    out = oracle.chroma_filter([0.5, 0.7, 0.5], rows)
    assert np.allclose(out[:, 0], [1.7, 3.4]) and np.allclose(out[:, 1], [10.2, 11.9])
    assert oracle.chroma_filter([0.25, 0.75, 1.0, 0.75, 0.25], np.ones((4, 12))).shape[0] == 0


def test_quantizer_and_gray_code_upstream_kat(oracle):
    # upstream test_quantizer.cpp: Quantizer(0.0, 0.1, 0.3)
    q = lambda v: oracle.quantize(v, 0.0, 0.1, 0.3)
    assert [q(-0.1), q(0.0), q(0.03), q(0.1), q(0.13), q(0.3), q(0.33), q(1000.0)] == [0, 1, 1, 2, 2, 3, 3, 3]
    assert [oracle.gray_code(i) for i in range(4)] == [0, 1, 3, 2]


def test_normalizer(oracle):
    v = np.arange(12, dtype=np.float64)
    out = oracle.normalize(v)
    assert np.allclose(out, v / np.sqrt((v ** 2).sum())) and abs((out ** 2).sum() - 1) < 1e-12
    assert np.array_equal(oracle.normalize(np.full(12, 0.001)), np.zeros(12))     # norm 0.0035 < 0.01
    assert np.array_equal(oracle.normalize(np.zeros(12)), np.zeros(12))


def test_filters_against_direct_sums(oracle):
    rng = np.random.default_rng(0)
    img = rng.random((20, 12))
    x = 3

    def A(r1, c1, r2, c2):
        return img[x + r1:x + r2, c1:c2].sum()
    L = lambda a, b: np.log((1 + a) / (1 + b))
    y, h, w = 2, 6, 12
    want = {
        0: L(A(0, y, w, y + h), 0.0),
        1: L(A(0, y + 3, w, y + h), A(0, y, w, y + 3)),
        2: L(A(6, y, w, y + h), A(0, y, 6, y + h)),
        3: L(A(0, y + 3, 6, y + h) + A(6, y, w, y + 3), A(0, y, 6, y + 3) + A(6, y + 3, w, y + h)),
        4: L(A(0, y + 2, w, y + 4), A(0, y, w, y + 2) + A(0, y + 4, w, y + h)),
        5: L(A(4, y, 8, y + h), A(0, y, 4, y + h) + A(8, y, w, y + h)),
    }
    for t, v in want.items():
        assert abs(oracle.filter_apply(t, y, h, w, img, x) - v) < 1e-12
    # upstream test_filter.cpp flavour: 2x2 image {1,2;3,4}... Filter0 over the whole image
    small = np.array([[1.0, 2.0], [3.0, 4.0]])
    assert abs(oracle.filter_apply(0, 0, 2, 2, small, 0) - np.log(11.0)) < 1e-12
    # odd sizes use integer halves / thirds (h/2, w/3 ...)
    assert abs(oracle.filter_apply(1, 0, 3, 5, img, 0)
               - np.log((1 + img[0:5, 1:3].sum()) / (1 + img[0:5, 0:1].sum()))) < 1e-12


def test_power_spectrum_against_numpy(oracle):
    rng = np.random.default_rng(1)
    frame = rng.integers(-32768, 32768, 4096).astype(np.int16)
    win = (1.0 / 32767.0) * (0.54 - 0.46 * np.cos(np.arange(4096) * 2.0 * np.pi / 4095))
    spec = np.fft.rfft(frame.astype(np.float64) * win)
    want = spec.real ** 2 + spec.imag ** 2
    got = oracle.power_spectrum(frame)
    assert got.shape == (2049,)
    assert np.max(np.abs(got - want) / (want + 1e-9)) < 1e-9
    # a pure tone lands in its bin: 1000 Hz -> bin 371.5
    t = np.arange(4096) / 11025
    tone = np.rint(20000 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.int16)
    p = oracle.power_spectrum(tone)
    assert int(np.argmax(p)) in (371, 372)


def test_frame_counts(oracle):
    for n, frames in ((0, 0), (4095, 0), (4096, 1), (4096 + 1364, 1), (4096 + 1365, 2), (11025 * 720, 5813)):
        assert oracle.num_frames(n) == frames == synth.num_frames(n)
        assert oracle.num_raw_hashes(n) == max(frames - 19, 0) == synth.num_raw_hashes(n)


@pytest.mark.parametrize("seed,seconds", [(0, 14.0), (1, 9.5)])
def test_c_oracle_equals_numpy_transcription(oracle, seed, seconds):
    rng = np.random.default_rng(seed)
    n = int(seconds * 11025)
    x = synth._noise(rng, n, amp=0.08) + np.pad(synth._chords(rng, seconds), (0, n))[:n]
    pcm = np.clip(np.rint(x * 32767), -32768, 32767).astype(np.int16)
    got, chroma = oracle.fingerprint(pcm, want_chroma=True)
    want, pchroma = pyref.fingerprint(pcm)
    assert got.shape == want.shape and got.size == oracle.num_raw_hashes(n) > 50
    assert np.max(np.abs(chroma - pchroma) / (np.abs(pchroma) + 1e-12)) < 1e-9
    assert np.array_equal(got, want)


def test_stereo_downmix_truncates_toward_zero(oracle):
    rng = np.random.default_rng(2)
    n = 4096 + 1365 * 30
    l = rng.integers(-32768, 32768, n).astype(np.int16)
    r = rng.integers(-32768, 32768, n).astype(np.int16)
    inter = np.empty(2 * n, np.int16)
    inter[0::2], inter[1::2] = l, r
    s = l.astype(np.int32) + r.astype(np.int32)
    mono = (np.sign(s) * (np.abs(s) // 2)).astype(np.int16)
    assert np.array_equal(oracle.fingerprint(inter, channels=2), oracle.fingerprint(mono))
    assert np.array_equal(oracle.fingerprint(inter, channels=2), pyref.fingerprint(inter, channels=2)[0])


def test_silence_gives_one_fixed_hash(oracle):
    h = oracle.fingerprint(np.zeros(4096 + 1365 * 40, np.int16))
    assert h.size == 22 and len(set(h.tolist())) == 1
    # all areas are 0 -> every classifier sees log(1/1) = 0
    bits = 0
    for (t, y, hh, w, t0, t1, t2) in pyref.CLASSIFIERS:
        q = (0 if 0.0 < t0 else 1) if 0.0 < t1 else (2 if 0.0 < t2 else 3)
        bits = (bits << 2) | pyref.GRAY[q]
    assert int(h[0]) == bits


def test_subsample_and_stamp(oracle):
    raw = np.arange(11, dtype=np.uint32) + 100
    h, t = oracle.subsample_and_stamp(raw, 2, seek_to_ns=9)
    assert h.tolist() == [100, 102, 104, 106, 108, 110]
    assert t.tolist() == [pyref.hash_timestamp(2_600_000_000, 123_000_000, i, 9) for i in range(0, 11, 2)]
    h, _ = oracle.subsample_and_stamp(raw, 3)
    assert h.tolist() == [100, 103, 106, 109]


def test_fingerprint_many_threads(oracle):
    rng = np.random.default_rng(3)
    segs = [rng.integers(-9000, 9000, n).astype(np.int16) for n in (50_000, 0, 4096, 80_001)]
    a = oracle.fingerprint_many(segs, n_threads=1)
    b = oracle.fingerprint_many(segs, n_threads=3)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert np.array_equal(a[0], oracle.fingerprint(segs[0]))
