"""The product's copy of the Chromaprint TEST2 constants
(needle_b200/csrc/fp_tables.h) equals the oracle's (oracle/chromaprint_tables.h)
and pyref's: one edit must change all three or this fails."""
import os
import re

from oracle import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROW = re.compile(r"\{\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(-?[\d.]+),\s*(-?[\d.]+),\s*(-?[\d.]+)\s*\}")


def rows(path):
    return [tuple(float(x) for x in m) for m in ROW.findall(open(os.path.join(ROOT, path)).read())]


def test_classifier_tables_agree():
    a = rows("needle_b200/csrc/fp_tables.h")
    b = rows("oracle/chromaprint_tables.h")
    assert len(a) == 16 and a == b
    assert a == [tuple(float(x) for x in c) for c in pyref.CLASSIFIERS]
    for (t, y, h, w, t0, t1, t2) in a:
        assert 0 <= t <= 5 and y + h <= 12 and 1 <= w <= 16 and t0 < t1 < t2


def test_scalar_constants_agree():
    prod = open(os.path.join(ROOT, "needle_b200/csrc/fp_tables.h")).read()
    orc = open(os.path.join(ROOT, "oracle/chromaprint_tables.h")).read() + \
        open(os.path.join(ROOT, "oracle/needle_oracle.h")).read()
    val = lambda src, name: re.search(name + r"\s*=?\s*(\d+)", src).group(1)
    assert val(prod, "FP_FRAME") == val(orc, "ORC_FRAME_SIZE") == "4096"
    assert val(orc, "ORC_FRAME_HOP") == "1365"
    assert val(prod, "FP_MIN_FREQ") == val(orc, "ORC_MIN_FREQ") == "28"
    assert val(prod, "FP_MAX_FREQ") == val(orc, "ORC_MAX_FREQ") == "3520"
    assert val(prod, "FP_FIR_LEN") == val(orc, "ORC_CHROMA_FILTER_LEN") == "5"
    assert "{0.25, 0.75, 1.0, 0.75, 0.25}" in prod and "{0.25, 0.75, 1.0, 0.75, 0.25}" in orc


def test_generated_chroma_fold_matches_oracle_notes():
    """needle_b200/csrc/fp_chroma_fold.inc (compile-time pitch classes of K1) against
    the oracle's Chroma::PrepareNotes table: every bin in [10, 1308) exactly once."""
    from oracle import oracle as orc
    lo, hi, notes = orc.chroma_notes()
    src = open(os.path.join(ROOT, "needle_b200/csrc/fp_chroma_fold.inc")).read()
    seen = {}
    for t, n, a, b in re.findall(r"FOLD\((\d+), (\d+), (\d+), (\d+)\)", src):
        for lane in range(int(a), int(b)):
            k = lane + 32 * int(t)
            assert k not in seen
            seen[k] = int(n)
    assert sorted(seen) == list(range(lo, hi))
    assert all(seen[k] == int(notes[k]) for k in seen)
