"""ctypes binding of libneedle.so, the needle-capi C ABI (include/needle.h) served by
the B200 library.  Used by the tests the way a C program uses needle.h; nothing here
computes anything."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libneedle.so")

# enum NeedleError (include/needle.h; reference needle-capi/needle.h:12-61)
(OK, INVALID_UTF8_STRING, NULL_ARGUMENT, INVALID_ARGUMENT, FRAME_HASH_DATA_NOT_FOUND,
 FRAME_HASH_DATA_INVALID_VERSION, INVALID_FRAME_HASH_DATA, COMPARATOR_MINIMUM_PATHS,
 ANALYZER_INVALID_HASH_PERIOD, ANALYZER_INVALID_HASH_DURATION, IO_ERROR, UNKNOWN) = range(12)

_P, _PP = C.c_void_p, C.POINTER(C.c_void_p)
_PATHS = C.POINTER(C.c_char_p)
PROTOTYPES = {
    "needle_error_to_str": (C.c_char_p, [C.c_int]),
    "needle_util_find_video_files": (C.c_int, [_PATHS, C.c_size_t, C.c_bool, C.c_bool,
                                               C.POINTER(C.POINTER(C.c_char_p)), C.POINTER(C.c_size_t)]),
    "needle_util_video_files_free": (None, [C.POINTER(C.c_char_p), C.c_size_t]),
    "needle_audio_analyzer_new_default": (C.c_int, [_PATHS, C.c_size_t, _PP]),
    "needle_audio_analyzer_new": (C.c_int, [_PATHS, C.c_size_t, C.c_float, C.c_float, C.c_bool, C.c_bool, C.c_bool,
                                            _PP]),
    "needle_audio_analyzer_get_frame_hashes": (C.c_int, [_P, C.c_size_t, _PP]),
    "needle_audio_analyzer_free": (None, [_P]),
    "needle_audio_analyzer_print_paths": (None, [_P]),
    "needle_audio_analyzer_run": (C.c_int, [_P, C.c_float, C.c_bool, C.c_bool]),
    "needle_audio_comparator_new_default": (C.c_int, [_PATHS, C.c_size_t, _PP]),
    "needle_audio_comparator_new": (C.c_int, [_PATHS, C.c_size_t, C.c_bool, C.c_uint16, C.c_uint16, C.c_uint16,
                                              C.c_float, _PP]),
    "needle_audio_comparator_free": (None, [_P]),
    "needle_audio_comparator_run": (C.c_int, [_P, C.c_bool, C.c_bool, C.c_bool, C.c_bool, C.c_bool]),
    "nb200_capi_set_decoder": (C.c_int, [_P]),
    "nb200_capi_header_md5": (C.c_int, [C.c_char_p, C.c_char_p]),
    "nb200_capi_frame_hashes_view": (C.c_int, [_P, C.c_int, _PP, _PP, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                               C.POINTER(C.c_char_p)]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libneedle.so is not built: python -m needle_b200.build")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def c_paths(paths):
    arr = (C.c_char_p * max(len(paths), 1))(*[os.fsencode(p) for p in paths])
    return arr


def frame_hashes(handle, ending: bool):
    """-> (hashes list, ts_ns list, hash_duration_ns, md5) of a FrameHashes handle."""
    h, t, md5 = C.c_void_p(), C.c_void_p(), C.c_char_p()
    n, hd = C.c_uint64(), C.c_uint64()
    st = lib().nb200_capi_frame_hashes_view(handle, 1 if ending else 0, C.byref(h), C.byref(t), C.byref(n),
                                            C.byref(hd), C.byref(md5))
    assert st == 0
    hh = C.cast(h, C.POINTER(C.c_uint32))
    tt = C.cast(t, C.POINTER(C.c_uint64))
    return [hh[k] for k in range(n.value)], [tt[k] for k in range(n.value)], hd.value, md5.value.decode()
