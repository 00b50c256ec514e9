"""The drop-in boundary on CPU: libneedle_b200.so loads, exports exactly what
include/needle_b200.h declares, refuses to compute without a CUDA device (no
CPU fallback), and its host-only entry points (persistence) work."""
import ctypes as C
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from needle_b200 import _lib, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "needle_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nb200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    # nb200_capi_* are the hooks of libneedle.so (tests/test_capi.py); the rest is libneedle_b200.so
    names = [n for n in header_functions() if not n.startswith("nb200_capi_")]
    assert len(names) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True)
    exported = set(re.findall(r" T (nb200_[a-z0-9_]+)", out.stdout))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert exported <= set(names), "undeclared exports: %s" % sorted(exported - set(names))
    assert sorted(_lib.PROTOTYPES) == names
    L = _lib.lib()
    for n in names:
        assert getattr(L, n) is not None


def test_header_compiles_as_c():
    src = '#include "needle_b200.h"\nint main(void) { nb200_match_params p; nb200_run r; (void)p; (void)r; return NB200_OK; }\n'
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.dirname(HEADER),
                        "-x", "c", "-"], input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Run) == 64 and _lib.RUN_DTYPE.itemsize == 64
    assert C.sizeof(_lib.MatchParams) == 32
    assert C.sizeof(_lib.SearchResultC) == 48
    p = _lib.MatchParams()
    _lib.lib().nb200_match_params_default(C.byref(p))
    # audio/mod.rs:14-45 defaults
    assert (p.hash_match_threshold, p.include_endings, p.min_opening_ns, p.min_ending_ns, p.time_padding_ns) == \
        (10, 0, 20_000_000_000, 20_000_000_000, 0)


def test_status_strings():
    L = _lib.lib()
    assert L.nb200_status_str(0) == b"ok"
    for s in range(1, 11):
        assert len(L.nb200_status_str(s)) > 3
    assert L.nb200_num_raw_hashes(4096 + 1365 * 19) == 1
    assert L.nb200_num_raw_hashes(4096 + 1365 * 19 - 1) == 0
    assert L.nb200_num_raw_hashes(11025 * 720) == 5794


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_compute_fails_loudly_without_a_device():
    with pytest.raises(_lib.Nb200Error) as e:
        engine.Context(0)
    assert e.value.status == _lib.ERR_CUDA
    assert "cuda" in str(e.value).lower()


def test_null_arguments():
    L = _lib.lib()
    assert L.nb200_ctx_create(0, None) == _lib.ERR_NULL_ARGUMENT
    assert L.nb200_match_run(None, None, None, 0, None, None) == _lib.ERR_NULL_ARGUMENT
    assert L.nb200_fp_feed(None, None, 0) == _lib.ERR_NULL_ARGUMENT
    assert L.nb200_framehashes_write(None, None, None, 0, None, None, 0, 0, None) == _lib.ERR_NULL_ARGUMENT
    L.nb200_ctx_destroy(None)
    L.nb200_hashset_free(None)
    L.nb200_runset_free(None)
    L.nb200_pcmset_free(None)
    L.nb200_fp_free(None)
    L.nb200_free(None)


# ------------------------------------------------------------- .needle.dat

def write_dat(path, oh, ot, eh, et, hd, md5):
    oh, ot = np.asarray(oh, np.uint32), np.asarray(ot, np.uint64)
    eh, et = np.asarray(eh, np.uint32), np.asarray(et, np.uint64)
    return _lib.lib().nb200_framehashes_write(str(path).encode(), _lib.ptr(oh), _lib.ptr(ot), oh.size,
                                              _lib.ptr(eh), _lib.ptr(et), eh.size, hd, md5.encode())


def read_dat(path):
    L = _lib.lib()
    oh, ot, eh, et = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    no, ne, hd = C.c_uint64(), C.c_uint64(), C.c_uint64()
    md5 = C.create_string_buffer(64)
    st = L.nb200_framehashes_read(str(path).encode(), C.byref(oh), C.byref(ot), C.byref(no), C.byref(eh),
                                  C.byref(et), C.byref(ne), C.byref(hd), md5)
    if st != 0:
        return st, None
    arr = lambda p, n, t: np.ctypeslib.as_array(C.cast(p, C.POINTER(t)), shape=(max(n, 1),))[:n].copy()
    out = (arr(oh, no.value, C.c_uint32), arr(ot, no.value, C.c_uint64), arr(eh, ne.value, C.c_uint32),
           arr(et, ne.value, C.c_uint64), hd.value, md5.value.decode())
    for p in (oh, ot, eh, et):
        L.nb200_free(p)
    return 0, out


def test_needle_dat_byte_layout(tmp_path):
    """bincode 1.3 fixint LE of FrameHashes (needle/src/audio/data.rs:15-26,60-80)."""
    path = tmp_path / "ep.needle.dat"
    md5 = "759c6a520c5ce70359fdff38c4be6b98"
    assert write_dat(path, [0xAABBCCDD, 7], [2_600_000_000, 3_123_456_789], [9], [1_081_000_000_001],
                     300_000_012, md5) == 0
    want = struct.pack("<II", 0, 0)                       # FrameHashesVersion::V1 -> variant 0; FrameHashesData::V1 -> 0
    want += struct.pack("<Q", 2)
    want += struct.pack("<IQI", 0xAABBCCDD, 2, 600_000_000) + struct.pack("<IQI", 7, 3, 123_456_789)
    want += struct.pack("<Q", 1) + struct.pack("<IQI", 9, 1081, 1)
    want += struct.pack("<QI", 0, 300_000_012)
    want += struct.pack("<Q", 32) + md5.encode()
    data = path.read_bytes()
    assert data == want
    assert len(data) == 76 + 16 * 3


def test_needle_dat_roundtrip_and_errors(tmp_path):
    rng = np.random.default_rng(0)
    oh = rng.integers(0, 2 ** 32, 2897, dtype=np.uint64).astype(np.uint32)
    ot = np.cumsum(rng.integers(1, 10 ** 9, 2897)).astype(np.uint64)
    path = tmp_path / "a.needle.dat"
    assert write_dat(path, oh, ot, [], [], 123, "x" * 32) == 0
    st, (roh, rot, reh, ret, hd, md5) = read_dat(path)
    assert st == 0 and np.array_equal(roh, oh) and np.array_equal(rot, ot)
    assert reh.size == 0 and ret.size == 0 and hd == 123 and md5 == "x" * 32
    assert read_dat(tmp_path / "missing.dat")[0] == _lib.ERR_IO
    raw = path.read_bytes()
    (tmp_path / "trunc.dat").write_bytes(raw[:100])
    assert read_dat(tmp_path / "trunc.dat")[0] == _lib.ERR_FORMAT
    (tmp_path / "ver.dat").write_bytes(struct.pack("<I", 1) + raw[4:])
    assert read_dat(tmp_path / "ver.dat")[0] == _lib.ERR_FORMAT
    (tmp_path / "huge.dat").write_bytes(raw[:8] + struct.pack("<Q", 2 ** 60) + raw[16:])
    assert read_dat(tmp_path / "huge.dat")[0] == _lib.ERR_FORMAT
    assert write_dat(tmp_path / "no_such_dir" / "x.dat", [], [], [], [], 0, "") == _lib.ERR_IO
