"""Host-side pieces of the Python binding that need no GPU."""
import ctypes as C

import numpy as np

from needle_b200 import _lib
from needle_b200.engine import ResultList, SearchResultC


def _filled(n):
    res = (SearchResultC * max(n, 1))()
    for v in range(n):
        res[v].present = 1 if v % 5 else 0
        res[v].has_opening = v & 1
        res[v].has_ending = (v >> 1) & 1
        res[v].opening_start_ns = v * 10 ** 9 + 7
        res[v].opening_end_ns = v * 10 ** 9 + 90 * 10 ** 9
        res[v].ending_start_ns = 2 ** 40 + v
        res[v].ending_end_ns = 2 ** 63 + v        # beyond i64: must stay unsigned
    return res


def test_result_record_layout():
    assert C.sizeof(SearchResultC) == 48 == _lib.RESULT_DTYPE.itemsize


def test_result_list_is_a_lazy_sequence_of_tuples():
    n = 37
    res = _filled(n)
    want = [res[v].astuple() for v in range(n)]
    got = ResultList(res, n)
    res[3].present = 99                       # a copy: later writes to the C array do not show
    assert len(got) == n
    assert got == want and want == got and not (got != want)
    assert list(got) == want and got[5] == want[5] and got[-1] == want[-1] and got[2:4] == want[2:4]
    assert all(isinstance(x, int) for x in got[7])
    assert got == ResultList(_filled(n), n)
    assert got != want[:-1] and got != ResultList(res, n)
    arr = got.as_array()
    assert arr.dtype == _lib.RESULT_DTYPE and arr["ending_end_ns"][4] == 2 ** 63 + 4
    assert sum(r[1] for r in got) == sum(w[1] for w in want)


def test_result_list_empty():
    empty = ResultList(_filled(0), 0)
    assert len(empty) == 0 and empty == [] and list(empty) == []
    assert np.asarray(empty.as_array()).size == 0
