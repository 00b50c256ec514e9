"""ctypes binding of libneedle_b200.so (include/needle_b200.h).

The library is the product: there is no CPU fallback here.  If the .so is
missing, or no CUDA device is present, the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libneedle_b200.so")

OK = 0
ERR_NULL_ARGUMENT = 1
ERR_INVALID_ARGUMENT = 2
ERR_CUDA = 3
ERR_NO_ENDING = 4
ERR_DURATION_UNDERFLOW = 5
ERR_TOO_LARGE = 6
ERR_IO = 7
ERR_FORMAT = 8
ERR_STATE = 9
ERR_COMPARATOR_MINIMUM_PATHS = 10
ERR_NCCL = 11
OPT_FORCE_GENERAL_MATCH = 1
OPT_K1_VARIANT = 2
OPT_MATCH_DENSE = 3
OPT_HOST_VOTE = 4
OPT_DEFER_WAIT = 5
OPT_MATCH_BAND_GROUP = 6


class Nb200Error(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        msg = "%s: %s" % (where, lib().nb200_status_str(status).decode())
        if detail:
            msg += " [" + detail + "]"
        super().__init__(msg)


class MatchParams(C.Structure):
    _fields_ = [("hash_match_threshold", C.c_uint32), ("include_endings", C.c_uint32),
                ("min_opening_ns", C.c_uint64), ("min_ending_ns", C.c_uint64),
                ("time_padding_ns", C.c_uint64)]


class Run(C.Structure):
    _fields_ = [("pair", C.c_uint32), ("is_ending", C.c_uint32), ("i_end", C.c_uint32),
                ("j_end", C.c_uint32), ("len", C.c_uint32), ("src_simhash", C.c_uint32),
                ("dst_simhash", C.c_uint32), ("reserved", C.c_uint32),
                ("src_start_ns", C.c_uint64), ("src_end_ns", C.c_uint64),
                ("dst_start_ns", C.c_uint64), ("dst_end_ns", C.c_uint64)]


RUN_DTYPE = np.dtype([("pair", "<u4"), ("is_ending", "<u4"), ("i_end", "<u4"), ("j_end", "<u4"),
                      ("len", "<u4"), ("src_simhash", "<u4"), ("dst_simhash", "<u4"),
                      ("reserved", "<u4"), ("src_start_ns", "<u8"), ("src_end_ns", "<u8"),
                      ("dst_start_ns", "<u8"), ("dst_end_ns", "<u8")])


class SearchResultC(C.Structure):
    _fields_ = [("present", C.c_uint32), ("has_opening", C.c_uint32), ("has_ending", C.c_uint32),
                ("reserved", C.c_uint32),
                ("opening_start_ns", C.c_uint64), ("opening_end_ns", C.c_uint64),
                ("ending_start_ns", C.c_uint64), ("ending_end_ns", C.c_uint64)]

    def astuple(self):
        return (self.present, self.has_opening, self.has_ending, self.opening_start_ns,
                self.opening_end_ns, self.ending_start_ns, self.ending_end_ns)


RESULT_DTYPE = np.dtype([("present", "<u4"), ("has_opening", "<u4"), ("has_ending", "<u4"), ("reserved", "<u4"),
                         ("opening_start_ns", "<u8"), ("opening_end_ns", "<u8"),
                         ("ending_start_ns", "<u8"), ("ending_end_ns", "<u8")])

# name -> (restype, argtypes); every symbol include/needle_b200.h declares
_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_U64P = C.POINTER(C.c_uint64)
PROTOTYPES = {
    "nb200_status_str": (C.c_char_p, [C.c_int]),
    "nb200_last_error": (C.c_char_p, []),
    "nb200_ctx_create": (C.c_int, [C.c_int, _PP]),
    "nb200_ctx_destroy": (None, [_P]),
    "nb200_ctx_set_stream": (C.c_int, [_P, _P]),
    "nb200_ctx_set_option": (C.c_int, [_P, C.c_int, C.c_int64]),
    "nb200_ctx_host_profile": (C.c_int, [_P, C.POINTER(C.c_double), C.c_int]),
    "nb200_ctx_synchronize": (C.c_int, [_P]),
    "nb200_ctx_last_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float), _U64P]),
    "nb200_ctx_last_vote_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "nb200_host_alloc": (C.c_int, [_PP, C.c_size_t]),
    "nb200_host_free": (C.c_int, [_P]),
    "nb200_free": (None, [_P]),
    "nb200_match_params_default": (None, [C.POINTER(MatchParams)]),
    "nb200_match_pairs": (C.c_int, [_P, _P, _P, _P, C.c_uint32, _P, C.c_uint64,
                                    C.POINTER(MatchParams), C.POINTER(C.POINTER(Run)), _U64P]),
    "nb200_search": (C.c_int, [_P, _P, _P, _P, _P, C.c_uint32, C.POINTER(MatchParams),
                               C.POINTER(SearchResultC)]),
    "nb200_search_hashset": (C.c_int, [_P, _P, _P, C.POINTER(MatchParams), C.POINTER(SearchResultC)]),
    "nb200_vote": (C.c_int, [_P, C.c_uint32, _P, C.c_uint64, C.POINTER(MatchParams),
                             _P, C.c_uint64, C.POINTER(SearchResultC)]),
    "nb200_vote_subset": (C.c_int, [_P, C.c_uint32, _P, C.c_uint64, C.POINTER(MatchParams),
                                    _P, C.c_uint64, _P, C.POINTER(SearchResultC)]),
    "nb200_hashset_upload": (C.c_int, [_P, _P, _P, _P, C.c_uint32, _PP]),
    "nb200_hashset_info": (C.c_int, [_P, C.POINTER(C.c_uint32), _U64P, _P]),
    "nb200_hashset_download": (C.c_int, [_P, _P, _P, _P]),
    "nb200_hashset_export_packed": (C.c_int, [_P, _P, _P, _P]),
    "nb200_hashset_device_ptrs": (C.c_int, [_P, _PP, _PP]),
    "nb200_hashset_from_device": (C.c_int, [_P, _P, _P, _P, C.c_uint32, _PP]),
    "nb200_hashset_from_device_scattered": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_uint32, _PP]),
    "nb200_hashset_view": (C.c_int, [_P, _P, _P, _P, _P, C.c_uint32, _PP]),
    "nb200_hashset_free": (None, [_P]),
    "nb200_match_run": (C.c_int, [_P, _P, _P, C.c_uint64, C.POINTER(MatchParams), _PP]),
    "nb200_runset_count": (C.c_int, [_P, _U64P, _U64P]),
    "nb200_runset_download": (C.c_int, [_P, _P, _P]),
    "nb200_runset_free": (None, [_P]),
    "nb200_num_raw_hashes": (C.c_uint64, [C.c_uint64]),
    "nb200_pcmset_upload": (C.c_int, [_P, _P, _P, C.c_int, C.c_uint32, _PP]),
    "nb200_pcmset_free": (None, [_P]),
    "nb200_pcmset_view": (C.c_int, [_P, _P, _P, _P, C.c_uint32, C.c_uint64, _PP]),
    "nb200_fingerprint_run": (C.c_int, [_P, _P, C.c_uint32, C.c_uint64, C.c_uint64, _P, _PP]),
    "nb200_fingerprint_layout": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, _P, _U64P]),
    "nb200_fingerprint_run_into": (C.c_int, [_P, _P, C.c_uint32, C.c_uint64, C.c_uint64, _P, _P, _P, C.c_uint64]),
    "nb200_timestamps_fill": (C.c_int, [_P, _P, _P, _P, _P, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64]),
    "nb200_fingerprint_host_into": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, _P,
                                              _P, _P, C.c_uint64]),
    "nb200_fingerprint_batch": (C.c_int, [_P, _P, _P, C.c_int, C.c_uint32, C.c_uint32, _P, _P]),
    "nb200_fp_new": (C.c_int, [_P, _PP]),
    "nb200_fp_free": (None, [_P]),
    "nb200_fp_sample_rate": (C.c_int, [_P]),
    "nb200_fp_start": (C.c_int, [_P, C.c_int, C.c_int]),
    "nb200_fp_feed": (C.c_int, [_P, _P, C.c_size_t]),
    "nb200_fp_finish": (C.c_int, [_P]),
    "nb200_fp_get_delay_ms": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "nb200_fp_get_item_duration_ms": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "nb200_fp_get_raw": (C.c_int, [_P, _PP, C.POINTER(C.c_size_t)]),
    "nb200_analyze_search": (C.c_int, [_P, _P, _P, C.c_int, C.c_uint32, _P, C.c_uint64,
                                       C.POINTER(MatchParams), C.POINTER(SearchResultC)]),
    "nb200_pcmset_search": (C.c_int, [_P, _P, _P, C.c_uint64, C.POINTER(MatchParams), C.POINTER(SearchResultC)]),
    "nb200_match_export": (C.c_int, [_P, _P, _P, C.c_uint64, C.POINTER(MatchParams), C.c_uint32, _P, C.c_uint64]),
    "nb200_vote_blocks": (C.c_int, [_P, _P, C.c_uint32, C.c_uint64, _P, C.c_uint32, _P, C.c_uint64,
                                    C.POINTER(MatchParams), C.c_int, C.POINTER(SearchResultC), _U64P]),
    "nb200_comm_unique_id": (C.c_int, [_P]),
    "nb200_comm_init_rank": (C.c_int, [_P, _P, C.c_int, C.c_int, _PP]),
    "nb200_comm_init_all": (C.c_int, [_P, C.c_int, _P]),
    "nb200_comm_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nb200_comm_destroy": (None, [_P]),
    "nb200_mjob_search_create": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_uint32, _P, _P, C.c_uint64,
                                           C.POINTER(MatchParams), _PP]),
    "nb200_mjob_season_create": (C.c_int, [_P, C.c_int, _P, _P, C.c_uint32, C.c_uint64, _P, C.c_uint64,
                                           C.POINTER(MatchParams), _PP]),
    "nb200_mjob_video_rank": (C.c_int, [_P, _P]),
    "nb200_mjob_upload_pcm": (C.c_int, [_P, _P]),
    "nb200_mjob_run": (C.c_int, [_P, _P, C.POINTER(SearchResultC)]),
    "nb200_mjob_phase_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "nb200_mjob_free": (None, [_P]),
    "nb200_plan_videos": (C.c_int, [_P, C.c_uint32, C.c_int, _P]),
    "nb200_plan_pairs": (C.c_int, [_P, C.c_uint32, _P, C.c_uint64, C.c_int, C.c_int, _P]),
    "nb200_framehashes_write": (C.c_int, [C.c_char_p, _P, _P, C.c_uint64, _P, _P, C.c_uint64,
                                          C.c_uint64, C.c_char_p]),
    "nb200_framehashes_read": (C.c_int, [C.c_char_p, _PP, _PP, _U64P, _PP, _PP, _U64P, _U64P,
                                         C.c_char_p]),
}

_lib = None


def lib():
    """Loads the C-ABI library.  Raises if it has not been built: the product
    has no other code path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libneedle_b200.so is missing (%s): build it with `python -m needle_b200.build`; "
                "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)   # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int, where: str):
    if status != OK:
        detail = lib().nb200_last_error().decode() if status in (ERR_CUDA, ERR_NCCL, ERR_STATE, ERR_TOO_LARGE) else ""
        raise Nb200Error(status, where, detail)


def ptr(a):
    """void* of a numpy array (None -> NULL)."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)
