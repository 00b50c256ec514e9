"""Known answers that come from OUTSIDE this repository, asserted as literals.

The reference holds no golden vectors for the fingerprint stage
(needle/src/audio/analyzer.rs:472-480 is #[ignore]d with a stale snapshot), but
upstream Chromaprint -- the library needle calls through chromaprint-sys-next
1.5.3 (needle/Cargo.lock:158-159) -- pins its TEST2 algorithm in its own API
test (tests/test_api.cpp, `Test2SilenceFp` / `Test2SilenceRawFp`): 130 blocks
of 1024 zero samples at 44100 Hz mono give a raw fingerprint of length 3 whose
every item is 627964279, and the fingerprint hash (simhash32 of the raw
fingerprint, which needle's compute_hash_for_match uses, comparator.rs:149-153)
is 627964279 too.  Resampling zeros to 11025 Hz gives 33280 zeros, which is
what the oracle and the kernels are fed here.

These literals pin: the classifier table (types, positions, sizes), all 48
quantiser thresholds' signs around zero, the Gray code, the packing order, the
19-frame warm-up and the frame count formula.  They do not pin the FFT, the
chroma fold or the filter coefficients (silence is zero everywhere): for those
the recalled building-block vectors of tests/test_oracle_fingerprint.py apply.
"""
import numpy as np
import pytest

from oracle import pyref

UPSTREAM_TEST2_SILENCE_ITEM = 627964279          # chromaprint tests/test_api.cpp
UPSTREAM_TEST2_SILENCE_LENGTH = 3
UPSTREAM_TEST2_SILENCE_SAMPLES_11025 = 130 * 1024 * 11025 // 44100   # 33280


def test_literal_is_what_the_docstring_says():
    assert UPSTREAM_TEST2_SILENCE_ITEM == 0x256DF977
    assert UPSTREAM_TEST2_SILENCE_SAMPLES_11025 == 33280


def test_oracle_reproduces_upstream_silence_raw_fingerprint(oracle):
    pcm = np.zeros(UPSTREAM_TEST2_SILENCE_SAMPLES_11025, np.int16)
    raw = oracle.fingerprint(pcm)
    assert raw.tolist() == [UPSTREAM_TEST2_SILENCE_ITEM] * UPSTREAM_TEST2_SILENCE_LENGTH


def test_numpy_transcription_reproduces_upstream_silence_raw_fingerprint():
    pcm = np.zeros(UPSTREAM_TEST2_SILENCE_SAMPLES_11025, np.int16)
    raw, _ = pyref.fingerprint(pcm)
    assert raw.tolist() == [UPSTREAM_TEST2_SILENCE_ITEM] * UPSTREAM_TEST2_SILENCE_LENGTH


def test_stereo_silence_same_answer(oracle):
    # needle feeds interleaved stereo (analyzer.rs:218: start(11025, 2))
    pcm = np.zeros(2 * UPSTREAM_TEST2_SILENCE_SAMPLES_11025, np.int16)
    assert oracle.fingerprint(pcm, channels=2).tolist() == [UPSTREAM_TEST2_SILENCE_ITEM] * 3


def test_simhash_of_upstream_silence_fingerprint(oracle):
    # chromaprint_get_fingerprint_hash = SimHash(raw fingerprint); upstream asserts 627964279
    raw = np.full(3, UPSTREAM_TEST2_SILENCE_ITEM, np.uint32)
    assert oracle.simhash32(raw) == UPSTREAM_TEST2_SILENCE_ITEM
    assert pyref.simhash32(raw) == UPSTREAM_TEST2_SILENCE_ITEM


def test_literal_from_the_classifier_table_by_hand():
    """Silence: every area is 0, every filter value is log(1/1) = 0; the 2-bit code of a
    classifier is Gray(number of thresholds <= 0).  Written out from the table so that a
    wrong sign or a swapped row shows up as a different literal."""
    below_or_equal_zero = [0, 3, 1, 1, 1, 3, 2, 1, 2, 2, 3, 1, 1, 2, 1, 2]   # t0<=0, t1<=0, t2<=0 counted per classifier
    gray = [0, 1, 3, 2]
    bits = 0
    for q in below_or_equal_zero:
        bits = (bits << 2) | gray[q]
    assert bits == UPSTREAM_TEST2_SILENCE_ITEM
    # and the table in the oracle says the same about its thresholds
    for (t, y, h, w, t0, t1, t2), q in zip(pyref.CLASSIFIERS, below_or_equal_zero):
        assert sum(1 for th in (t0, t1, t2) if th <= 0.0) == q


# Rust std's own documentation examples of Duration::from_secs_f32 (library/core/src/time.rs,
# "conversion uses rounding"), which needle relies on through Duration::mul_f32 (analyzer.rs:309)
# (the 3e10 s example exceeds the u64 nanoseconds the oracle counts in)
RUST_STD_FROM_SECS_F32 = [(0.0, 0), (1e-20, 0), (4.2e-7, 420), (2.7, 2_700_000_048),
                          (0.999e-9, 1)]


@pytest.mark.parametrize("secs,nanos", RUST_STD_FROM_SECS_F32)
def test_rust_duration_from_secs_f32_std_doc_examples(oracle, secs, nanos):
    assert oracle.duration_from_secs_f32(np.float32(secs)) == nanos
    assert pyref.duration_from_secs_f32(np.float32(secs)) == nanos


def test_hash_duration_constant():
    # needle's hash_duration = Duration::from_secs_f32(0.3) (main.rs:287, data.rs:135):
    # 0.3f32 is exactly 0.300000011920928955078125
    from fractions import Fraction
    exact = Fraction(float(np.float32(0.3)))
    assert exact == Fraction(10066330, 2 ** 25)
    assert round(exact * 10 ** 9) == 300_000_012 == pyref.duration_from_secs_f32(np.float32(0.3))
