"""libneedle.so = the needle-capi C ABI (include/needle.h) over the B200 library.

CPU part: what the reference's own capi tests check (needle-capi/src/lib.rs:650-760:
find_video_files / analyzer / comparator construct and free), plus its argument and
error conventions (NULL -> NullArgument :216,:383,:566; num_paths < 2 ->
ComparatorMinimumPaths :569; hash_duration <= 0 -> AnalyzerInvalidHashDuration :474;
index out of range -> InvalidArgument :424), the error strings (:139-203), the header
itself compiled as C against a small program written like examples/full.c, and the
host helpers (header md5, Path::with_extension, skip files).
GPU part: analyze -> .needle.dat -> search with display and skip files through the C
ABI equals the Python mirror of the same API (needle_b200/audio.py), which the other
tests tie to the oracle."""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import wave

import numpy as np
import pytest

from needle_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")


def test_exports_match_header():
    src = open(os.path.join(INCLUDE, "needle.h")).read()
    import re
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = sorted(set(re.findall(r"\b(needle_[a-z0-9_]+)\s*\(", src)))
    assert len(declared) == 13            # needle-capi/needle.h declares 13 functions
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True, check=True)
    exported = set(re.findall(r" T ((?:needle|nb200_capi)_[a-z0-9_]+)", out.stdout))
    assert set(declared) | {"nb200_capi_set_decoder", "nb200_capi_frame_hashes_view",
                            "nb200_capi_header_md5"} == exported
    assert sorted(capi.PROTOTYPES) == sorted(exported)


def test_error_enum_and_strings():
    L = capi.lib()
    want = ["No error", "Invalid UTF-8 string", "Input argument is NULL",
            "One or more input arguments were invalid (usually zero)", "Frame hash data not found on disk",
            "Frame hash data has an invalid version.", "Invalid frame hash data read from disk",
            "Comparator requires at least 2 video paths", "Analyzer hash period must be greater than 0",
            "Analyzer hash duration must be greater than 3 seconds", "I/O error",
            "Unknown error occurred; please re-run with logging enabled"]
    assert [L.needle_error_to_str(k).decode() for k in range(12)] == want
    assert capi.OK == 0 and capi.UNKNOWN == 11


def test_analyzer_new_and_free_like_the_reference_tests():
    L = capi.lib()
    paths = capi.c_paths(["/tmp/abcd.mkv"])
    for make in (lambda out: L.needle_audio_analyzer_new_default(paths, 1, C.byref(out)),
                 lambda out: L.needle_audio_analyzer_new(paths, 1, 0.33, 0.2, True, False, True, C.byref(out))):
        a = C.c_void_p()
        assert make(a) == capi.OK and a.value
        fh = C.c_void_p()
        assert L.needle_audio_analyzer_get_frame_hashes(a, 0, C.byref(fh)) == capi.INVALID_ARGUMENT   # nothing run yet
        L.needle_audio_analyzer_free(a)
    L.needle_audio_analyzer_free(None)


def test_comparator_new_and_free_like_the_reference_tests():
    L = capi.lib()
    paths = capi.c_paths(["/tmp/abcd.mkv", "/tmp/efgh.mp4"])
    c = C.c_void_p()
    assert L.needle_audio_comparator_new(paths, 2, False, 10, 10, 10, 0.0, C.byref(c)) == capi.OK and c.value
    L.needle_audio_comparator_free(c)
    c = C.c_void_p()
    assert L.needle_audio_comparator_new_default(paths, 2, C.byref(c)) == capi.OK and c.value
    L.needle_audio_comparator_free(c)
    L.needle_audio_comparator_free(None)


def test_argument_conventions():
    L = capi.lib()
    paths = capi.c_paths(["/tmp/abcd.mkv", "/tmp/efgh.mp4"])
    out = C.c_void_p()
    assert L.needle_audio_analyzer_new_default(None, 1, C.byref(out)) == capi.NULL_ARGUMENT
    assert L.needle_audio_analyzer_new_default(paths, 1, None) == capi.NULL_ARGUMENT
    assert L.needle_audio_comparator_new_default(None, 2, C.byref(out)) == capi.NULL_ARGUMENT
    assert L.needle_audio_comparator_new_default(paths, 1, C.byref(out)) == capi.COMPARATOR_MINIMUM_PATHS
    assert L.needle_audio_comparator_run(None, False, False, False, False, True) == capi.NULL_ARGUMENT
    assert L.needle_audio_analyzer_run(None, 0.3, False, True) == capi.NULL_ARGUMENT
    a = C.c_void_p()
    assert L.needle_audio_analyzer_new_default(paths, 2, C.byref(a)) == capi.OK
    assert L.needle_audio_analyzer_run(a, 0.0, False, True) == capi.ANALYZER_INVALID_HASH_DURATION
    assert L.needle_audio_analyzer_run(a, -1.0, False, True) == capi.ANALYZER_INVALID_HASH_DURATION
    assert L.needle_audio_analyzer_get_frame_hashes(None, 0, C.byref(out)) == capi.NULL_ARGUMENT
    L.needle_audio_analyzer_free(a)
    bad = (C.c_char_p * 2)(b"/tmp/ok", b"/tmp/\xff\xfe")
    assert L.needle_audio_analyzer_new_default(bad, 2, C.byref(out)) == capi.INVALID_UTF8_STRING
    holes = (C.c_char_p * 2)(b"/tmp/ok", None)
    assert L.needle_audio_analyzer_new_default(holes, 2, C.byref(out)) == capi.NULL_ARGUMENT
    vids, n = C.POINTER(C.c_char_p)(), C.c_size_t()
    assert L.needle_util_find_video_files(None, 1, True, True, C.byref(vids), C.byref(n)) == capi.NULL_ARGUMENT
    assert L.needle_util_find_video_files(paths, 0, True, True, C.byref(vids), C.byref(n)) == capi.INVALID_ARGUMENT
    assert L.needle_util_find_video_files(paths, 2, True, True, C.byref(vids), C.byref(n)) == capi.UNKNOWN  # PathNotFound
    L.needle_util_video_files_free(None, 0)


def test_header_md5_is_md5_of_the_first_8k(tmp_path):
    """util::compute_header_md5sum (util.rs:99-105) against hashlib, incl. RFC 1321's padding edge
    (8192 bytes = 128 full blocks + one padding block) and read_exact's failure on short files."""
    L = capi.lib()
    rng = np.random.default_rng(3)
    out = C.create_string_buffer(33)
    for size in (8192, 8193, 100_000):
        p = tmp_path / ("f%d.bin" % size)
        data = rng.integers(0, 256, size, dtype=np.uint8).tobytes()
        p.write_bytes(data)
        assert L.nb200_capi_header_md5(os.fsencode(str(p)), out) == 0
        assert out.value.decode() == hashlib.md5(data[:8192]).hexdigest()
    zeros = tmp_path / "zeros.bin"
    zeros.write_bytes(bytes(8192))
    assert L.nb200_capi_header_md5(os.fsencode(str(zeros)), out) == 0
    assert out.value.decode() == hashlib.md5(bytes(8192)).hexdigest()
    short = tmp_path / "short.bin"
    short.write_bytes(b"x" * 8191)
    assert L.nb200_capi_header_md5(os.fsencode(str(short)), out) != 0
    assert L.nb200_capi_header_md5(os.fsencode(str(tmp_path / "missing")), out) != 0
    assert L.nb200_capi_header_md5(None, out) != 0


def write_wav(path, pcm, channels=1, rate=11025):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(2)
        w.setframerate(rate)
        w.writeframes(np.asarray(pcm, dtype="<i2").tobytes())


def test_find_video_files(tmp_path):
    L = capi.lib()
    rng = np.random.default_rng(0)
    d = tmp_path / "season"
    d.mkdir()
    write_wav(d / "e1.wav", rng.integers(-100, 100, 20000))
    write_wav(d / "e2.wav", rng.integers(-100, 100, 20000), channels=2)
    write_wav(d / "wrong_rate.wav", rng.integers(-100, 100, 20000), rate=44100)
    (d / "notes.txt").write_bytes(b"hello" * 2000)
    (d / "e1.needle.dat").write_bytes(b"\0" * 100)
    mp4 = bytes([0, 0, 0, 0x20]) + b"ftypisom" + bytes(9000)          # an ISO-BMFF header, no streams
    (d / "clip.mp4").write_bytes(mp4)
    (d / "sub").mkdir()
    write_wav(d / "sub" / "deep.wav", rng.integers(-100, 100, 20000))  # one level only (util.rs:58)
    single = tmp_path / "lone.wav"
    write_wav(single, rng.integers(-100, 100, 20000))

    def find(paths, full, audio):
        vids, n = C.POINTER(C.c_char_p)(), C.c_size_t()
        assert L.needle_util_find_video_files(capi.c_paths(paths), len(paths), full, audio, C.byref(vids),
                                              C.byref(n)) == capi.OK
        got = sorted(os.path.basename(vids[k].decode()) for k in range(n.value))
        L.needle_util_video_files_free(vids, n.value)
        return got
    # header sniff: container signatures (+ the WAVE stand-in); full: the decoder must open it
    assert find([str(d), str(single)], False, False) == ["clip.mp4", "e1.wav", "e2.wav", "lone.wav", "wrong_rate.wav"]
    assert find([str(d), str(single)], True, True) == ["e1.wav", "e2.wav", "lone.wav"]
    assert find([str(d / "notes.txt")], False, False) == []


def test_c_program_against_the_header(tmp_path):
    """A C caller written the way needle-capi/examples/full.c is, compiled against
    include/needle.h and linked with libneedle.so; runs the non-GPU part."""
    src = tmp_path / "full.c"
    src.write_text(r'''
#include <stdio.h>
#include <needle.h>
int main(int argc, char **argv) {
    NeedleError err;
    NeedleAudioAnalyzer *analyzer = NULL;
    const NeedleAudioComparator *comparator = NULL;
    const char *const *video_paths = NULL;
    size_t num_video_paths = 0;
    const char *paths[] = { argv[1] };
    err = needle_util_find_video_files(paths, 1, false, true, &video_paths, &num_video_paths);
    if (err != 0) { printf("find: %s\n", needle_error_to_str(err)); return 1; }
    err = needle_audio_analyzer_new_default(video_paths, num_video_paths, &analyzer);
    if (err != 0) { printf("analyzer: %s\n", needle_error_to_str(err)); return 2; }
    err = needle_audio_comparator_new_default(video_paths, num_video_paths, &comparator);
    if (err != 0) { printf("comparator: %s\n", needle_error_to_str(err)); return 3; }
    needle_audio_analyzer_print_paths(analyzer);
    if (argc > 2) {
        err = needle_audio_analyzer_run(analyzer, 0.3f, true, true);
        if (err != 0) { printf("run: %s\n", needle_error_to_str(err)); return 4; }
        err = needle_audio_comparator_run(comparator, false, true, false, true, true);
        if (err != 0) { printf("search: %s\n", needle_error_to_str(err)); return 5; }
    }
    needle_audio_analyzer_free(analyzer);
    needle_audio_comparator_free(comparator);
    needle_util_video_files_free(video_paths, num_video_paths);
    return 0;
}
''')
    exe = tmp_path / "full"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", INCLUDE, str(src), "-o", str(exe),
                        "-L", os.path.dirname(capi.LIB_PATH), "-lneedle",
                        "-Wl,-rpath," + os.path.dirname(capi.LIB_PATH)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d = tmp_path / "s"
    d.mkdir()
    rng = np.random.default_rng(1)
    for k in range(3):
        write_wav(d / ("e%d.wav" % k), rng.integers(-100, 100, 30000))
    r = subprocess.run([str(exe), str(d)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert sorted(r.stdout.split()) == sorted(str(d / ("e%d.wav" % k)) for k in range(3))


# ------------------------------------------------------------------ GPU: the whole flow

def make_season(tmp_path, n=4, stereo=False):
    from needle_b200 import synth
    eps = synth.make_pcm_season(n, 3.0, season_seed=31, intro_s=40.0, credits_s=30.0)
    paths = []
    for k, ep in enumerate(eps):
        p = tmp_path / ("ep%02d.wav" % k)
        pcm = np.repeat(ep.pcm, 2) if stereo else ep.pcm
        write_wav(p, pcm, channels=2 if stereo else 1)
        paths.append(str(p))
    return paths


@pytest.mark.gpu
@pytest.mark.parametrize("stereo", [False, True])
def test_analyze_persist_search_equals_the_python_mirror(tmp_path, capfd, stereo):
    from needle_b200 import audio
    L = capi.lib()
    paths = make_season(tmp_path, 4, stereo)
    cp = capi.c_paths(paths)
    a = C.c_void_p()
    assert L.needle_audio_analyzer_new(cp, len(paths), 0.5, 0.25, True, False, False, C.byref(a)) == capi.OK
    assert L.needle_audio_analyzer_run(a, 0.3, True, True) == capi.OK
    # the Python mirror of the same API on the same files (without persisting)
    mirror = audio.Analyzer.from_files(paths, False, True).with_include_endings(True)
    want = mirror.run(audio.duration_from_secs_f32(0.3), False)
    for k, p in enumerate(paths):
        fh = C.c_void_p()
        assert L.needle_audio_analyzer_get_frame_hashes(a, k, C.byref(fh)) == capi.OK
        oh, ot, hd, md5 = capi.frame_hashes(fh, False)
        eh, et, _, _ = capi.frame_hashes(fh, True)
        assert oh == want[k].opening_hashes.tolist() and ot == want[k].opening_ts_ns.tolist()
        assert eh == want[k].ending_hashes.tolist() and et == want[k].ending_ts_ns.tolist()
        assert hd == 300_000_012 and md5 == hashlib.md5(open(p, "rb").read(8192)).hexdigest()
        # .needle.dat on disk: the bytes the mirror writes (bincode layout pinned in test_abi.py)
        dat = p[:-4] + ".needle.dat"
        want[k].save(str(tmp_path / "mirror.dat"))
        assert open(dat, "rb").read() == open(tmp_path / "mirror.dat", "rb").read()
    fh = C.c_void_p()
    assert L.needle_audio_analyzer_get_frame_hashes(a, len(paths), C.byref(fh)) == capi.INVALID_ARGUMENT
    # a second run finds the files: "Skipping analysis for ..."
    capfd.readouterr()
    assert L.needle_audio_analyzer_run(a, 0.3, False, True) == capi.OK
    out = capfd.readouterr().out
    assert out.count("Skipping analysis for") == len(paths)
    L.needle_audio_analyzer_free(a)

    # search from the .needle.dat files, display + skip files
    c = C.c_void_p()
    assert L.needle_audio_comparator_new(cp, len(paths), True, 10, 20, 20, 0.0, C.byref(c)) == capi.OK
    capfd.readouterr()
    assert L.needle_audio_comparator_run(c, False, True, False, True, True) == capi.OK
    shown = capfd.readouterr().out
    comp = audio.Comparator.from_files(paths).with_include_endings(True)
    import contextlib
    import io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        results = comp.run(False, display=True)
    assert shown == buf.getvalue()
    assert shown.count("* Opening - \"") == len(paths) and shown.count("* Ending - \"") == len(paths)
    for p, r in zip(paths, results):
        skip = json.load(open(p[:-4] + ".needle.skip.json"))
        assert skip["md5"] == hashlib.md5(open(p, "rb").read(8192)).hexdigest()
        f32 = lambda ns: float(np.float32(np.float32(ns // 10 ** 9) + np.float32(ns % 10 ** 9) / np.float32(1e9)))
        assert [float(np.float32(x)) for x in skip["opening"]] == [f32(r.opening[0]), f32(r.opening[1])]
        assert [float(np.float32(x)) for x in skip["ending"]] == [f32(r.ending[0]), f32(r.ending[1])]
    # with use_skip_files every video is now skipped
    assert L.needle_audio_comparator_run(c, False, True, True, False, True) == capi.OK
    assert capfd.readouterr().out.count("Skipping due to existing skip file...") == len(paths)
    L.needle_audio_comparator_free(c)


@pytest.mark.gpu
def test_search_analyze_in_place_and_missing_data(tmp_path, capfd):
    L = capi.lib()
    paths = make_season(tmp_path, 3)
    cp = capi.c_paths(paths)
    c = C.c_void_p()
    assert L.needle_audio_comparator_new_default(cp, len(paths), C.byref(c)) == capi.OK
    # no .needle.dat yet
    assert L.needle_audio_comparator_run(c, False, False, False, False, True) == capi.FRAME_HASH_DATA_NOT_FOUND
    # analyze in place (openings only: Analyzer::default)
    capfd.readouterr()
    assert L.needle_audio_comparator_run(c, True, True, False, False, True) == capi.OK
    out = capfd.readouterr().out
    assert out.count("* Opening - \"") == 3 and "Ending" not in out
    L.needle_audio_comparator_free(c)
    # --analyze with include_endings: FrameHashDataNoEnding in the reference (an unwrap panic there)
    assert L.needle_audio_comparator_new(cp, len(paths), True, 10, 20, 20, 0.0, C.byref(c)) == capi.OK
    assert L.needle_audio_comparator_run(c, True, False, False, False, True) == capi.UNKNOWN
    L.needle_audio_comparator_free(c)


@pytest.mark.gpu
def test_custom_decoder_callback(tmp_path, capfd):
    """nb200_capi_set_decoder: the host supplies decoded audio (what FFmpeg + swresample do in
    needle, analyzer.rs:170-283) for files the library cannot read itself.  Here the "videos" are
    .mkv files holding only an EBML signature; the PCM comes from this process."""
    from needle_b200 import audio, synth
    L = capi.lib()
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    eps = synth.make_pcm_season(3, 3.0, season_seed=5, intro_s=40.0, credits_s=30.0)
    rng = np.random.default_rng(2)
    pcm_of, paths = {}, []
    for k, ep in enumerate(eps):
        p = tmp_path / ("ep%d.mkv" % k)
        p.write_bytes(bytes([0x1A, 0x45, 0xDF, 0xA3]) + rng.integers(0, 256, 9000, dtype=np.uint8).tobytes())
        pcm_of[os.fsencode(str(p))] = ep.pcm
        paths.append(str(p))
    calls = {"probe": 0, "decode": 0, "release": 0}

    PROBE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_int))
    DECODE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p),
                         C.POINTER(C.c_uint64), C.POINTER(C.c_int))
    RELEASE = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)

    def probe(_user, path, duration_ns, has_video, has_audio):
        calls["probe"] += 1
        if path not in pcm_of:
            return 1
        duration_ns[0] = pcm_of[path].size * 10 ** 9 // 11025
        has_video[0], has_audio[0] = 1, 1
        return 0

    def decode(_user, path, from_ns, until_ns, pcm, n, channels):
        calls["decode"] += 1
        x = pcm_of[path]
        a = min(x.size, from_ns * 11025 // 10 ** 9)
        b = x.size if until_ns == 2 ** 64 - 1 else min(x.size, until_ns * 11025 // 10 ** 9)
        part = np.ascontiguousarray(x[a:b])
        buf = libc.malloc(max(part.nbytes, 2))
        C.memmove(buf, part.ctypes.data, part.nbytes)
        pcm[0], n[0], channels[0] = buf, part.size, 1
        return 0

    def release(_user, p):
        calls["release"] += 1
        libc.free(p)

    class Decoder(C.Structure):
        _fields_ = [("user", C.c_void_p), ("probe", PROBE), ("decode", DECODE), ("release", RELEASE)]
    dec = Decoder(None, PROBE(probe), DECODE(decode), RELEASE(release))
    assert L.nb200_capi_set_decoder(C.byref(dec)) == 0
    try:
        cp = capi.c_paths(paths)
        vids, n = C.POINTER(C.c_char_p)(), C.c_size_t()
        assert L.needle_util_find_video_files(capi.c_paths([str(tmp_path)]), 1, True, True, C.byref(vids),
                                              C.byref(n)) == capi.OK
        assert sorted(vids[k].decode() for k in range(n.value)) == sorted(paths)
        L.needle_util_video_files_free(vids, n.value)
        a = C.c_void_p()
        assert L.needle_audio_analyzer_new(cp, 3, 0.5, 0.25, True, False, True, C.byref(a)) == capi.OK
        assert L.needle_audio_analyzer_run(a, 0.3, True, True) == capi.OK
        assert calls["decode"] == 6 and calls["release"] == 6      # opening + ending per video
        # the same PCM through the library's batched fingerprint API: identical hashes
        for k, ep in enumerate(eps):
            # the segments Analyzer::run_single asks for (analyzer.rs:378-402), in samples
            n_s = ep.pcm.size
            dur = n_s * 10 ** 9 // 11025
            n_open = min(n_s, audio.duration_mul_f32(dur, 0.5) * 11025 // 10 ** 9)
            s_end = min(n_s, audio.duration_mul_f32(dur, np.float32(1.0) - np.float32(0.25)) * 11025 // 10 ** 9)
            o, e = ep.pcm[:n_open], ep.pcm[s_end:]
            with audio.engine.Context(-1) as ctx:
                want = ctx.fingerprint_batch([o, e])
            fh = C.c_void_p()
            assert L.needle_audio_analyzer_get_frame_hashes(a, k, C.byref(fh)) == capi.OK
            assert capi.frame_hashes(fh, False)[0] == want[0][::2].tolist()
            assert capi.frame_hashes(fh, True)[0] == want[1][::2].tolist()
        L.needle_audio_analyzer_free(a)
        c = C.c_void_p()
        assert L.needle_audio_comparator_new(cp, 3, True, 10, 20, 20, 0.0, C.byref(c)) == capi.OK
        capfd.readouterr()
        assert L.needle_audio_comparator_run(c, False, True, False, False, True) == capi.OK
        out = capfd.readouterr().out
        assert out.count('* Opening - "') == 3 and out.count('* Ending - "') == 3
        L.needle_audio_comparator_free(c)
    finally:
        assert L.nb200_capi_set_decoder(None) == 0
