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
